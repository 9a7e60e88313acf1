#!/bin/bash
# Round 2, eight GPUs (gpurun --gpus 8): NCCL semantics test (2 ranks), the default bench line at N=8 (C2 inference
# replicas + C4 train block: 4 sequences per GPU, NCCL all-reduce of the 298 MB gradients, overlapped / single / none),
# and C5 (1024x1024, T=16, one sequence per GPU): train step and inference.
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout -k 10 600 python -m pytest tests/test_gpu_nccl.py -m gpu -q -x -s 2>&1 | grep "DIAG\|two identical\|passed\|failed\|Error" | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
run() { tag=$1; shift; timeout -k 10 900 $TR bench.py --gpus $N "$@" > gpurun_out/n${N}_$tag.json 2> gpurun_out/n${N}_$tag.err; echo "n$N $tag rc=$?"; grep -v "^\*\|NCCL version\|OMP_NUM\|^$" gpurun_out/n${N}_$tag.err | tail -2; }
run bench --steps 10 --warmup 3
run c5_train --mode train --no-parity --steps 4 --warmup 3 --no-cpu --size 1024 --unroll 16 --batch 1
run c5_infer --mode infer --no-parity --no-variants --steps 5 --warmup 3 --no-cpu --size 1024 --unroll 16 --batch 1
run train_syncbn --mode train --no-parity --steps 6 --warmup 3 --no-cpu --sync-bn
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for tag in ('bench','c5_train','c5_infer','train_syncbn'):
    try:
        d=json.loads(open('gpurun_out/n%s_%s.json'%(N,tag)).read()); t=d.get('train',{})
        print(tag, 'value %.1f ms %.2f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), '| train %.1f fps %.1f ms' % (t.get('value',0), t.get('ms_per_step',0)), t.get('allreduce',{}).get('ms_per_step_by_mode'), d.get('clocks'))
    except Exception as e: print(tag, 'unreadable', e)
PY
