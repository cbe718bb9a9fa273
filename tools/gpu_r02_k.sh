#!/bin/bash
# TMA-store epilogue for EPI_F32 tiles (corr build, fnet convs, mask head): parity + same-box A/B against the previous build.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
PREV=$PWD/raft-tf_b200/lib/libraft_b200_prev.so
echo "== parity (kernels)"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 --tb=short -x 2>&1 | tail -15 | tee $O/tmastore_parity.log
echo "== timings (ABAB)"
for r in 1 2; do
  for w in corr encoder update iterate forward; do
    echo -n "new      $w: "; timeout 200 python tools/micro.py $w 2>&1 | tail -1
    echo -n "no-store $w: "; RAFT_B200_NO_TMA_STORE=1 timeout 200 python tools/micro.py $w 2>&1 | tail -1
    echo -n "previous $w: "; RAFT_B200_LIB=$PREV timeout 200 python tools/micro.py $w 2>&1 | tail -1
  done
done | tee $O/tmastore_ab.log
echo "== e2e / fullsize / configs"
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py -q -m gpu --timeout 600 --tb=line 2>&1 | tail -5 | tee $O/tmastore_e2e.log
echo "== racecheck / memcheck (small)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize_small.py > $O/r02_sanitizer_memcheck.log 2>&1; tail -3 $O/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize_small.py > $O/r02_sanitizer_racecheck.log 2>&1; tail -3 $O/r02_sanitizer_racecheck.log
