#!/bin/bash
# tools/ab.sh LIB_A LIB_B [...] — same-box A/B of tuning builds: the headline bench line (value only) for every library,
# two rounds, alternating.  Libraries are paths relative to the repo root (build/variants/*.so travel to the GPU box).
for round in 1 2; do
  for lib in "$@"; do
    v=$(VMIS_LIB=$PWD/$lib python bench.py --steps 6 --warmup 3 --no-cpu-baseline --sections none 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print(round(d['value']/1e6,3), round(d['e2e']['value']/1e6,3))")
    echo "round $round $lib: value/e2e M qps = $v"
  done
done
