#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== config3 parity diagnosis"
timeout 600 python tools/diag_parity.py --tag default 2>&1 | tail -2 | tee $O/diag_parity.log
for k in RAFT_B200_NO_FH2_FUSE RAFT_B200_NO_STASH RAFT_B200_LOOKUP_V4 RAFT_B200_NO_HOIST; do env $k=1 timeout 300 python tools/diag_parity.py --tag $k 2>&1 | tail -1; done | tee -a $O/diag_parity.log
timeout 600 python tools/diag_parity.py --simt --tag simt-fp32-convs 2>&1 | tail -1 | tee -a $O/diag_parity.log
echo "== config2 for comparison"
timeout 600 python tools/diag_parity.py --H 436 --W 1024 --seed0 1000 --tag cfg2-default 2>&1 | tail -2 | tee -a $O/diag_parity.log
timeout 600 python tools/diag_parity.py --H 436 --W 1024 --seed0 1000 --simt --tag cfg2-simt 2>&1 | tail -1 | tee -a $O/diag_parity.log
echo "== batched == per-sample"
timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -m gpu --timeout 600 --tb=line 2>&1 | tail -4
