#!/bin/bash
# PDL default: parity + A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -4
for args in "" "--B 8"; do
  echo -n "default (PDL) $args : "; timeout 300 python tools/micro.py update $args 2>&1 | tail -1
  echo -n "NO_PDL        $args : "; RAFT_B200_NO_PDL=1 timeout 300 python tools/micro.py update $args 2>&1 | tail -1
  echo -n "PDL_EARLY     $args : "; RAFT_B200_PDL_EARLY=1 timeout 300 python tools/micro.py update $args 2>&1 | tail -1
  echo -n "default (PDL) $args : "; timeout 300 python tools/micro.py update $args 2>&1 | tail -1
done
timeout 300 python tools/phase_times.py 2>&1 | tail -11 | cut -c1-220
echo -n "bench default: "; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
