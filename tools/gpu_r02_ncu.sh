#!/bin/bash
# Round 2 ncu evidence (one GPU): launch list of one forward pass, --set full captures summarised ON the box (reports > 64 MiB
# do not travel back), plus the A/B of the PDL-chained instance-norm kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
PREV=$PWD/raft-tf_b200/lib/libraft_b200_prev.so
echo "== encoder parity + A/B (PDL norm kernels)"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "encoder" --timeout 300 --tb=short 2>&1 | tail -4
for r in 1 2; do
  echo -n "new      encoder: "; timeout 200 python tools/micro.py encoder 2>&1 | tail -1
  echo -n "previous encoder: "; RAFT_B200_LIB=$PREV timeout 200 python tools/micro.py encoder 2>&1 | tail -1
done | tee $O/enc_pdl_ab.log
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu --timeout 300 --tb=line 2>&1 | tail -3
echo "== ncu launch list (one forward, no graph)"
RAFT_B200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/r02_launches.csv \
    python tools/micro.py forward > $O/ncu_launches.log 2>&1
python tools/launch_summary.py $O/r02_launches.csv > $O/r02_launch_list_summary.txt 2>&1; head -30 $O/r02_launch_list_summary.txt
echo "== ncu full: lookup B=1 / B=8"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_lookup -s 2 -c 1 -f -o $O/r02_lookup_b1 \
    python tools/micro.py lookup --reps 3 --n 1 > $O/ncu_lookup1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_lookup -s 2 -c 1 -f -o $O/r02_lookup_b8 \
    python tools/micro.py lookup --B 8 --reps 3 --n 1 > $O/ncu_lookup8.log 2>&1
(echo "# corr_lookup_kernel<4,split>, B=1 (55x128 grid), cold"; python tools/ncu_summary.py $O/r02_lookup_b1.ncu-rep; echo "# B=8"; python tools/ncu_summary.py $O/r02_lookup_b8.ncu-rep) > $O/r02_lookup_ncu_full.txt 2>&1
cat $O/r02_lookup_ncu_full.txt
echo "== ncu full: corr build"
timeout 600 ncu --set full --clock-control none -k "regex:conv_tc|corr_prep" -s 5 -c 5 -f -o $O/r02_corr \
    python tools/micro.py corr --reps 1 --n 1 > $O/ncu_corr.log 2>&1
python tools/ncu_summary.py $O/r02_corr.ncu-rep > $O/r02_corr_ncu_full.txt 2>&1; cat $O/r02_corr_ncu_full.txt; rm -f $O/r02_corr.ncu-rep
echo "== ncu full: update-step convs"
RAFT_B200_NO_PDL=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_tc|flow_conv7" -s 15 -c 11 -f -o $O/r02_update \
    python tools/micro.py update --reps 2 --n 1 > $O/ncu_update.log 2>&1
python tools/ncu_summary.py $O/r02_update.ncu-rep > $O/r02_update_convs_ncu_full.txt 2>&1; cat $O/r02_update_convs_ncu_full.txt
echo "== ncu full: encoder (first kernels of fnet / cnet: stem view, layer1, norm passes)"
timeout 900 ncu --set full --clock-control none -k "regex:conv_tc|enc_stem|inorm" -s 0 -c 40 -f -o $O/r02_encoder \
    python tools/micro.py encoder --reps 1 --n 1 > $O/ncu_encoder.log 2>&1
python tools/ncu_summary.py $O/r02_encoder.ncu-rep > $O/r02_encoder_ncu_full.txt 2>&1; cat $O/r02_encoder_ncu_full.txt; rm -f $O/r02_encoder.ncu-rep
du -sh $O
