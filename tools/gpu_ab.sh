#!/bin/bash
# same-box A/B of the opt-in conv kernel variants after the issue-loop fixes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for knob in RAFT_B200_HALO RAFT_B200_CTA2 RAFT_B200_PAIR; do
  echo -n "$knob parity: "; env $knob=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x --timeout 600 -k "(conv2d or update_block or encoder) and tc" 2>&1 | tail -1
done
for args in "" "--B 8"; do
  echo -n "default $args : "; timeout 300 python tools/micro.py update $args 2>&1 | tail -1
  for knob in RAFT_B200_HALO RAFT_B200_CTA2 RAFT_B200_PAIR RAFT_B200_PDL RAFT_B200_FUSED; do
    echo -n "$knob $args : "; env $knob=1 timeout 300 python tools/micro.py update $args 2>&1 | tail -1
  done
  echo -n "default $args : "; timeout 300 python tools/micro.py update $args 2>&1 | tail -1
done
RAFT_B200_HALO=1 timeout 300 python tools/phase_times.py 2>&1 | tail -11 | cut -c1-220
echo -n "bench default: "; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
