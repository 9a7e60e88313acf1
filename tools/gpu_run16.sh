#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --mode train --steps 5 --warmup 3 --no-cpu 2>/dev/null | cut -c1-260
