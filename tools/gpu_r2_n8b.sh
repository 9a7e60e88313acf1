#!/bin/bash
# Round 2, eight GPUs (gpurun --gpus 8): C5 (1024x1024, T=16, one sequence per GPU) full train step with the NCCL gradient
# all-reduce (overlapped / single / none).  The default bench line at N = 8 is left to the driver's scaling run.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout -k 10 400 $TR bench.py --gpus 8 --mode train --no-parity --steps 4 --warmup 3 --no-cpu --size 1024 --unroll 16 --batch 1 > gpurun_out/n8_c5_train.json 2> gpurun_out/n8_c5_train.err; echo "n8 c5_train rc=$?"
grep -v "^\*\|NCCL version\|OMP_NUM\|^$" gpurun_out/n8_c5_train.err | tail -3
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/n8_c5_train.json').read()); t=d['train']
    print('c5 n8: %.1f fps %.1f ms' % (t['value'], t['ms_per_step']), t['allreduce'], d.get('clocks'))
except Exception as e: print('unreadable', e)
PY
