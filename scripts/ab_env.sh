#!/bin/bash
# A/B of process-wide knobs inside ONE gpurun call: runs `bench.py --no-cpu --no-train --no-sweep --no-precisions --no-kernels`
# once per environment setting given as arguments ("VAR=value ..." strings; "" = defaults), two rounds, and prints LPs/s.
for round in 1 2; do
  for setting in "$@"; do
    out=$(env $setting python bench.py --no-cpu --no-train --no-sweep --no-precisions --no-kernels --steps 400 --warmup 20 2>/dev/null | tail -1)
    python - "$setting" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2])
print(f"[{sys.argv[1] or 'defaults'}] value {d['value']:.1f} LPs/s  e2e {d['e2e']['value']:.1f}  clocks {d['clocks']['sm_mhz']}")
PY
  done
done
