#!/bin/bash
mkdir -p gpurun_out
for cfg in "" "LU_WGRAD_PAIR=0" "LU_PAIR=0" "LU_WGRAD_ENGINE=simt"; do echo "== $cfg"; env $cfg timeout 300 python tools/diag_determinism.py 4 2>&1 | grep -v "^$" | tail -14; done
echo "== bf16"; timeout 300 python tools/diag_determinism.py 3 bf16 2>&1 | tail -10
