#!/bin/bash
# Round-2 bring-up of the cta_group::2 variants (compiled in round 1, never run: the GPU budget was spent).
#   LU_PAIR=1            conv / ConvLSTM / data-gradient kernel, cluster mode 3 (lu_conv_tc_kernel<.., .., 3>)
#   LU_WGRAD_CLUSTER=3   weight-gradient kernel, cluster mode 3 (lu_wgrad_tc_kernel<3>)
# Order = DESIGN 14.1: single small network vs the scalar mirror (one process per case, every step under its own
# `timeout` so a barrier dead-lock cannot hold the box) -> GPU suite with the switch on -> A/B benches.  The script stops
# at the first step that fails or times out: nothing after it would be meaningful.
#   gpurun --timeout 1500 -- 'bash tools/gpu_pair.sh > gpurun_out/pair.log 2>&1'
mkdir -p gpurun_out
step() {  # step <seconds> <label> <command...>
  local secs=$1 label=$2; shift 2
  timeout -k 10 "$secs" "$@"; local rc=$?
  echo "== $label rc=$rc"
  if [ $rc -ne 0 ]; then echo "STOP at: $label"; nvidia-smi --query-gpu=name,memory.used --format=csv,noheader; exit 1; fi
}
export LU_PAIR=1
# ConvLSTM launches of a small net take the pair path (any LSTM launch with a tap table does); bf16x3 gives 1e-3 headroom
step 180 "diag one  pair"  python tools/diag_tc.py one tcgen05 halo bf16x3 1 32 24
step 180 "diag two  pair"  python tools/diag_tc.py two tcgen05 halo bf16x3 2 32 24
step 180 "diag two  pair odd tail (33 M tiles)" python tools/diag_tc.py two tcgen05 halo bf16x3 2 88 48
step 180 "diag odd  pair"  python tools/diag_tc.py odd tcgen05 halo bf16x3 2 64 48
step 900 "forward suite pair" python -m pytest tests/test_gpu_forward.py -m gpu -x -q
step 900 "train suite pair (data-gradient launches)" python -m pytest tests/test_gpu_train.py -m gpu -x -q
for p in 0 1; do
  LU_PAIR=$p step 600 "bench infer LU_PAIR=$p" bash -c "python bench.py --mode infer --no-parity --no-variants --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_infer_pair$p.json"
  cut -c1-200 gpurun_out/bench_infer_pair$p.json
done
export LU_PAIR=0
export LU_WGRAD_CLUSTER=3
step 900 "wgrad pair vs scalar wgrad" python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "wgrad or train_step"
for c in 1 3; do
  LU_WGRAD_CLUSTER=$c step 900 "bench train LU_WGRAD_CLUSTER=$c" bash -c "python bench.py --mode train --no-parity --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_train_wg$c.json"
  cut -c1-200 gpurun_out/bench_train_wg$c.json
done
echo "ALL STEPS PASSED"
# ncu --set full of the same launches round 1 profiled with the default kernels (tools/gpu_final2.sh: level-1 ConvLSTM launch,
# level-1 weight-gradient launch), now in pair mode -- compare tensor-pipe activity and the tensor memory pipe with
# profiles/r1_ncu_prof_lstm_l1.txt / r1_ncu_prof_wgrad_l1.txt
NCU="ncu --set full --clock-control none --import-source on"
LU_PAIR=1 LU_WGRAD_CLUSTER=1 timeout -k 10 1500 $NCU -k regex:lu_conv_tc_kernel -s 117 -c 1 -o gpurun_out/prof_lstm_l1_pair python bench.py --mode infer --no-parity --no-variants --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu lstm pair rc=$?"
LU_PAIR=0 LU_WGRAD_CLUSTER=3 timeout -k 10 1500 $NCU -k regex:lu_wgrad_tc_kernel -s 69 -c 1 -o gpurun_out/prof_wgrad_l1_pair python bench.py --mode train --no-parity --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu wgrad pair rc=$?"
ls -la gpurun_out/*.ncu-rep
