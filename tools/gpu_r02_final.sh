#!/bin/bash
# Round 2 final single-GPU pass: parity suite (one process per file), stage timings, both bench arms, ncu evidence.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.sm --format=csv > $O/gpu.txt 2>&1
: > $O/pytest_gpu.log
for f in tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py tests/test_gpu_variants.py tests/test_gpu_multi.py; do
  echo "-- $f" | tee -a $O/pytest_gpu.log
  timeout 1200 python -m pytest $f -q -m gpu --timeout 900 --maxfail=3 --tb=short 2>&1 | grep -vE "^\s*$" | tail -25 | tee -a $O/pytest_gpu.log
done
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.log
echo "== lookup timings"
for B in 1 8; do
  for fl in "" "--flush"; do
    echo -n "default       B=$B $fl: "; timeout 200 python tools/micro.py lookup --B $B $fl 2>&1 | tail -1
  done
done 2>&1 | tee $O/lookup_ab.log
echo "== stage timings"
for w in corr encoder update iterate forward; do timeout 200 python tools/micro.py $w 2>&1 | tail -1; done | tee $O/stages.log
for knob in RAFT_B200_NO_STASH RAFT_B200_NO_SPLITK RAFT_B200_NO_PDL RAFT_B200_NO_HOIST; do
  for w in update iterate; do echo -n "$knob=1 $w: "; env $knob=1 timeout 200 python tools/micro.py $w 2>&1 | tail -1; done
done | tee -a $O/stages.log
echo -n "B=8 update: "; timeout 200 python tools/micro.py update --B 8 2>&1 | tail -1 | tee -a $O/stages.log
echo -n "B=8 encoder: "; timeout 200 python tools/micro.py encoder --B 8 2>&1 | tail -1 | tee -a $O/stages.log
echo -n "volume-free forward: "; RAFT_B200_VOLUME_FREE=1 timeout 200 python tools/micro.py forward 2>&1 | tail -1 | tee -a $O/stages.log
echo "== bench (ours)"
timeout 900 python bench.py 2>$O/bench_err.log | tail -1 | tee $O/bench_default.json | cut -c1-1200
echo "== bench (reference arm)"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>>$O/bench_err.log | tail -1 | tee $O/bench_reference.json | cut -c1-600
echo "== ncu launch list (one forward, no graph)"
RAFT_B200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/r02_launches.csv \
    python tools/micro.py forward > $O/ncu_launches.log 2>&1
echo "== ncu full: lookup B=1 / B=8"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_lookup -s 2 -c 1 -f -o $O/r02_lookup_b1 \
    python tools/micro.py lookup --reps 3 --n 1 > $O/ncu_lookup1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_lookup -s 2 -c 1 -f -o $O/r02_lookup_b8 \
    python tools/micro.py lookup --B 8 --reps 3 --n 1 > $O/ncu_lookup8.log 2>&1
echo "== ncu full: corr build"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_tc|corr_prep" -s 5 -c 5 -f -o $O/r02_corr \
    python tools/micro.py corr --reps 1 --n 1 > $O/ncu_corr.log 2>&1
echo "== ncu full: update-step convs"
RAFT_B200_NO_PDL=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_tc|flow_conv7" -s 15 -c 11 -f -o $O/r02_update \
    python tools/micro.py update --reps 2 --n 1 > $O/ncu_update.log 2>&1
echo "== ncu full: encoder convs (stem view, strided view, layer1)"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_tc|enc_stem|inorm" -s 0 -c 40 -f -o $O/r02_encoder \
    python tools/micro.py encoder --reps 1 --n 1 > $O/ncu_encoder.log 2>&1
ls -la $O/*.ncu-rep | tail -8
