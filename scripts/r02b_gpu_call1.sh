#!/bin/bash
# GPU call 1 of the second round-2 session: correctness of the Krylov savings, then A/B bench lines.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02b
mkdir -p $O
exec > $O/call1.log 2>&1
set -x
date
nvidia-smi -L
free -g | head -2
nproc
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" || echo SMOKE_FAILED
date
# the new regression test first, then the suite without the 257^3 CPU-port file (its iteration counts are checked by
# krylov_switch_check.py against the path that file verified)
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_headline_parity.py 2>&1 | tail -15
date
timeout 600 python scripts/krylov_switch_check.py 257 > $O/switch_check.jsonl 2> $O/switch_check.err || echo SWITCH_CHECK_FAILED
cat $O/switch_check.jsonl
date
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B --e2e-steps 2 > $O/bench_new_16.json 2> $O/bench_new_16.err
date
HH_SKIP_LAST_UPDATE=0 HH_SMALL_FUSED=0 $B --no-e2e > $O/bench_plain_16.json 2> $O/bench_plain_16.err
date
$B --nrhs 32 --e2e-steps 2 > $O/bench_new_32.json 2> $O/bench_new_32.err
date
HH_HOST_CHUNKS=2 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --nrhs 32 --e2e-steps 2 > $O/bench_new_32_halves.json 2> $O/bench_new_32_halves.err
date
for f in $O/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", (d.get("e2e") or {}).get("value"),
              "ps", (d.get("e2e_point_sources") or {}).get("value"), d["config"].get("iterations_mean"), d["config"].get("true_relres_max_last_step"), d.get("clocks"))
        pk = d["roofline"]["per_kernel"]
        for k, v in pk.items():
            print("   %-22s share %.3f avg_ms %.4f gbs %s n %d" % (k, v["share"], v["avg_ms"], v["gbs"], v["launches"]))
PY
done
tail -3 $O/*.err
date
