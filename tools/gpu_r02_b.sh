#!/bin/bash
# Round 2, diagnostic pass: what is illegal in lookup v5 (TMA or not), everything else validated on the v4 lookup.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== v5 without TMA (plain staging)"
RAFT_B200_LOOKUP_NOTMA=1 timeout 120 python tools/micro.py lookup --n 2 --reps 1 2>&1 | tail -2
echo "== v5 volume-free tests (same kernel, no TMA)"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "volume_free" --timeout 200 2>&1 | tail -4
echo "== v5 with TMA under compute-sanitizer"
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/micro.py lookup --n 1 --reps 1 2>&1 | grep -vE "^\s*$" | head -40 | tee $O/sanitizer_v5.log
echo "== v5 NOTMA lookup tests"
RAFT_B200_LOOKUP_NOTMA=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "lookup" --timeout 200 2>&1 | tail -4
export RAFT_B200_LOOKUP_V4=1
echo "== full suite on the v4 lookup"
: > $O/pytest_gpu.log
for f in tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py tests/test_gpu_variants.py tests/test_gpu_multi.py; do
  echo "-- $f" | tee -a $O/pytest_gpu.log
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 --maxfail=4 -s 2>&1 | grep -vE "^\s*$" | tail -12 | tee -a $O/pytest_gpu.log
done
echo "== stage timings (v4 lookup)"
for w in corr encoder update iterate forward; do timeout 200 python tools/micro.py $w 2>&1 | tail -1; done | tee $O/stages.log
for knob in RAFT_B200_NO_FH2_FUSE RAFT_B200_NO_STASH RAFT_B200_ZR1_SIDE; do
  for w in update iterate; do echo -n "$knob=1 $w: "; env $knob=1 timeout 200 python tools/micro.py $w 2>&1 | tail -1; done
done | tee -a $O/stages.log
echo -n "B=8 update: "; timeout 200 python tools/micro.py update --B 8 2>&1 | tail -1 | tee -a $O/stages.log
echo -n "volume-free forward: "; RAFT_B200_VOLUME_FREE=1 timeout 200 python tools/micro.py forward 2>&1 | tail -1 | tee -a $O/stages.log
echo "== bench (ours, v4 lookup)"
timeout 900 python bench.py 2>$O/bench_err.log | tail -1 | tee $O/bench_default.json | cut -c1-1200
echo "== ncu launch list (one forward, no graph)"
RAFT_B200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/r02_launches.csv \
    python tools/micro.py forward > $O/ncu_launches.log 2>&1
echo "== ncu full: update-step convs"
RAFT_B200_NO_PDL=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_tc|flow_conv7|fh2_gather" -s 15 -c 11 -f -o $O/r02_update \
    python tools/micro.py update --reps 2 --n 1 > $O/ncu_update.log 2>&1
ls -la $O | tail -8
