#!/bin/bash
# Round 2, two GPUs (gpurun --gpus 2): the NCCL test of the multi-GPU semantics, then the default bench line at N=2
# (C2 batch-sharded inference + C4-style train block: gradient all-reduce overlapped / single / none).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout -k 10 900 python -m pytest tests/test_gpu_nccl.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -15 > gpurun_out/n2_pytest.log; echo "pytest nccl rc=${PIPESTATUS[0]}"; tail -5 gpurun_out/n2_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout -k 10 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench n2 rc=$?"; cut -c1-300 gpurun_out/n2_bench.json; grep -v "^\*\|NCCL version\|^$" gpurun_out/n2_bench.err | tail -3
timeout -k 10 900 $TR bench.py --gpus 2 --mode train --no-parity --steps 8 --warmup 3 --sync-bn > gpurun_out/n2_train_syncbn.json 2> gpurun_out/n2_train_syncbn.err; echo "train syncbn rc=$?"; cut -c1-200 gpurun_out/n2_train_syncbn.json
python - <<'PY'
import json
for f in ('gpurun_out/n2_bench.json','gpurun_out/n2_train_syncbn.json'):
    try:
        d=json.loads(open(f).read()); t=d.get('train',{})
        print(f, 'value %.1f e2e %.1f | train %.1f fps %.1f ms' % (d['value'], d['e2e']['value'], t.get('value',0), t.get('ms_per_step',0)), t.get('allreduce'))
    except Exception as e: print(f, 'unreadable', e)
PY
