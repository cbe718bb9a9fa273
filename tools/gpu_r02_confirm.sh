#!/bin/bash
# Last confirmation at HEAD: every -m gpu test file (one process each), smoke(), both bench arms, then the kernel evidence
# that changes with the update step (launch list of one forward, ncu --set full of the update-step convs, summarised here).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
: > $O/pytest_gpu.log
for f in tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py tests/test_gpu_variants.py tests/test_gpu_multi.py; do
  echo "-- $f" | tee -a $O/pytest_gpu.log
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 --maxfail=3 --tb=short 2>&1 | grep -vE "^\s*$" | tail -12 | tee -a $O/pytest_gpu.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.log
for w in corr encoder update iterate forward; do timeout 200 python tools/micro.py $w 2>&1 | tail -1; done | tee $O/stages.log
timeout 900 python bench.py 2>$O/bench_err.log | tail -1 > $O/bench_default.json; cut -c1-700 $O/bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>$O/bench_err.log | tail -1 > $O/bench_reference.json; cut -c1-300 $O/bench_reference.json
RAFT_B200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/r02_launches.csv \
    python tools/micro.py forward > $O/ncu_launches.log 2>&1
python tools/launch_summary.py $O/r02_launches.csv > $O/r02_launch_list_summary.txt 2>&1; head -28 $O/r02_launch_list_summary.txt
RAFT_B200_NO_PDL=1 timeout 600 ncu --set full --clock-control none -k "regex:conv_tc|flow_prep" -s 16 -c 12 -f -o $O/r02_update \
    python tools/micro.py update --reps 2 --n 1 > $O/ncu_update.log 2>&1
python tools/ncu_summary.py $O/r02_update.ncu-rep > $O/r02_update_convs_ncu_full.txt 2>&1; cut -c1-330 $O/r02_update_convs_ncu_full.txt; rm -f $O/r02_update.ncu-rep
