#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/pytest_gpu_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_train.log
tail -25 gpurun_out/pytest_gpu_train.log
