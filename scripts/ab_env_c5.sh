#!/bin/bash
# A/B of process-wide knobs on the C5 sweep inside ONE gpurun call (see ab_env.sh)
for round in 1 2; do
  for setting in "$@"; do
    out=$(env $setting python bench.py --workload C5 --no-cpu 2>/dev/null | tail -1)
    python - "$setting" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2])
print(f"[{sys.argv[1] or 'defaults'}] one call per LP {d['value']:.0f} LPs/s   packed e2e {d['e2e']['value']:.0f} LPs/s  launches/LP {d['gpu_launches']/d['steps']:.1f}")
PY
  done
done
