#!/bin/bash
# GPU call 6: compute-sanitizer on the kernels of the second session; the bench line once more with the hardened e2e leg
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02b
mkdir -p $O
exec > $O/call6.log 2>&1
set -x
date
timeout 120 python scripts/sanitize_case.py
date
timeout 400 compute-sanitizer --tool memcheck python scripts/sanitize_case.py > $O/sanitizer_memcheck.log 2>&1
tail -n 12 $O/sanitizer_memcheck.log
date
timeout 500 compute-sanitizer --tool racecheck python scripts/sanitize_case.py > $O/sanitizer_racecheck.log 2>&1
tail -n 12 $O/sanitizer_racecheck.log
date
timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 1 > $O/bench6_n1.json 2> $O/bench6_n1.err
python - $O/bench6_n1.json <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d.get("e2e"), "ps", (d.get("e2e_point_sources") or {}).get("value"))
PY
tail -n 5 $O/bench6_n1.err
date
