#!/bin/bash
# Round 2, GPU pass C (one GPU): parity suite (one process per file, stop early on failures), A/B timings, bench, ncu evidence.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
: > $O/pytest_gpu.log
ok=1
for f in tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py; do
  echo "-- $f" | tee -a $O/pytest_gpu.log
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 --maxfail=3 --tb=short -s 2>&1 | grep -vE "^\s*$" | tail -40 | tee -a $O/pytest_gpu.log
  if ! tail -3 $O/pytest_gpu.log | grep -qE "passed" || tail -3 $O/pytest_gpu.log | grep -qE "failed|error"; then ok=0; fi
done
if [ $ok = 1 ]; then
  for f in tests/test_gpu_variants.py tests/test_gpu_multi.py; do
    echo "-- $f" | tee -a $O/pytest_gpu.log
    timeout 600 python -m pytest $f -q -m gpu --timeout 300 --maxfail=2 --tb=short 2>&1 | grep -vE "^\s*$" | tail -15 | tee -a $O/pytest_gpu.log
  done
fi
echo "== lookup A/B"
for B in 1 8; do
  for fl in "" "--flush"; do
    echo -n "default       B=$B $fl: "; timeout 200 python tools/micro.py lookup --B $B $fl 2>&1 | tail -1
    echo -n "v5 (opt-in)   B=$B $fl: "; RAFT_B200_LOOKUP_V5=1 timeout 200 python tools/micro.py lookup --B $B $fl 2>&1 | tail -1
  done
done 2>&1 | tee $O/lookup_ab.log
echo "== stage timings"
for w in corr encoder update iterate forward; do timeout 200 python tools/micro.py $w 2>&1 | tail -1; done | tee $O/stages.log
for knob in RAFT_B200_NO_FH2_FUSE RAFT_B200_NO_STASH RAFT_B200_NO_DELTA_FUSE RAFT_B200_LOOKUP_V5; do
  for w in update iterate; do echo -n "$knob=1 $w: "; env $knob=1 timeout 200 python tools/micro.py $w 2>&1 | tail -1; done
done | tee -a $O/stages.log
echo -n "B=8 update: "; timeout 200 python tools/micro.py update --B 8 2>&1 | tail -1 | tee -a $O/stages.log
echo -n "volume-free forward: "; RAFT_B200_VOLUME_FREE=1 timeout 200 python tools/micro.py forward 2>&1 | tail -1 | tee -a $O/stages.log
echo "== bench (ours)"
timeout 900 python bench.py 2>$O/bench_err.log | tail -1 | tee $O/bench_default.json | cut -c1-1500
echo "== ncu launch list (one forward, no graph)"
RAFT_B200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/r02_launches.csv \
    python tools/micro.py forward > $O/ncu_launches.log 2>&1
echo "== ncu full: lookup v5 B=1 / B=8"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_lookup -s 2 -c 1 -f -o $O/r02_lookup_b1 \
    python tools/micro.py lookup --reps 3 --n 1 > $O/ncu_lookup1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_lookup -s 2 -c 1 -f -o $O/r02_lookup_b8 \
    python tools/micro.py lookup --B 8 --reps 3 --n 1 > $O/ncu_lookup8.log 2>&1
echo "== ncu full: corr build"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_tc|corr_prep" -s 5 -c 5 -f -o $O/r02_corr \
    python tools/micro.py corr --reps 1 --n 1 > $O/ncu_corr.log 2>&1
echo "== ncu full: update-step convs"
RAFT_B200_NO_PDL=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_tc|flow_conv7|fh2_gather" -s 15 -c 11 -f -o $O/r02_update \
    python tools/micro.py update --reps 2 --n 1 > $O/ncu_update.log 2>&1
ls -la $O | tail -8
