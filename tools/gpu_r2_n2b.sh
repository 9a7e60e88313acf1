#!/bin/bash
# Round 2, two GPUs (gpurun --gpus 2): the stand-alone block tests + the NCCL tests of the multi-GPU semantics, the default
# bench line at N=2 (C2 batch-sharded inference + C4-style train block with the NCCL gradient all-reduce: overlapped /
# single / none), and the C5 shape (1024x1024, T=16, one sequence per GPU) as a 2-rank train step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout -k 10 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_nccl.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -15 > gpurun_out/n2_pytest.log; echo "pytest blocks+nccl rc=${PIPESTATUS[0]}"; tail -6 gpurun_out/n2_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
run() { tag=$1; shift; t0=$SECONDS; timeout -k 10 900 $TR bench.py --gpus 2 "$@" > gpurun_out/n2_$tag.json 2> gpurun_out/n2_$tag.err; echo "n2 $tag rc=$? wall $((SECONDS-t0)) s"; grep -v "^\*\|NCCL version\|OMP_NUM\|^$" gpurun_out/n2_$tag.err | tail -2; }
run bench --steps 10 --warmup 3
run c5_train --mode train --no-parity --steps 4 --warmup 3 --no-cpu --size 1024 --unroll 16 --batch 1
python - <<'PY'
import json
for tag in ('bench','c5_train'):
    try:
        d=json.loads(open('gpurun_out/n2_%s.json'%tag).read()); t=d.get('train',{})
        print(tag, 'value %.1f ms %.2f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), '| train %.1f fps %.1f ms' % (t.get('value',0), t.get('ms_per_step',0)), t.get('allreduce',{}).get('ms_per_step_by_mode'), d.get('clocks'))
    except Exception as e: print(tag, 'unreadable', e)
PY
