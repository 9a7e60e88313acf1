#!/bin/bash
# Round 2, call A: GPU suite (incl. the CTC-network parity tests), default bench line, then the cta_group::2 bring-up.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
nproc
timeout -k 10 1800 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/a_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -5 gpurun_out/a_pytest.log
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/a_bench.json; tail -3 gpurun_out/a_bench.err
timeout -k 10 1500 bash tools/gpu_pair.sh > gpurun_out/a_pair.log 2>&1; echo "pair rc=$?"; grep "==\|STOP\|ALL STEPS\|rc=" gpurun_out/a_pair.log | tail -30
