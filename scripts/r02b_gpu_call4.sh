#!/bin/bash
# GPU call 4: cached interpolation in k_fine3d_tma_prob, in-place x1 conversion in k_fine3d_tma_first, level buffer swap
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02b
mkdir -p $O
exec > $O/call4.log 2>&1
set -x
date
export HH_PRO_CACHE=1 HH_FIRST_CONVERT=1 HH_LEVEL_SWAP=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" || echo SMOKE_FAILED
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_headline_parity.py 2>&1 | tail -40
date
unset HH_PRO_CACHE HH_FIRST_CONVERT HH_LEVEL_SWAP
HH_CHECK_ALL=1 timeout 600 python scripts/krylov_switch_check.py 257 > $O/switch_check4.jsonl 2> $O/switch_check4.err || echo SWITCH_CHECK_FAILED
cat $O/switch_check4.jsonl
tail -n 5 $O/switch_check4.err
date
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --nrhs 16"
$B > $O/bench4_base.json 2> $O/bench4_base.err
HH_PRO_CACHE=1 $B > $O/bench4_cache.json 2> $O/bench4_cache.err
HH_FIRST_CONVERT=1 $B > $O/bench4_convert.json 2> $O/bench4_convert.err
HH_LEVEL_SWAP=1 $B > $O/bench4_swap.json 2> $O/bench4_swap.err
HH_PRO_CACHE=1 HH_FIRST_CONVERT=1 HH_LEVEL_SWAP=1 $B > $O/bench4_all.json 2> $O/bench4_all.err
date
HH_HOST_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $O/bench4_trace32.json 2> $O/bench4_trace32.err
date
for f in $O/bench4_*.json; do echo "== $f"; python - "$f" <<'PY'
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
for f in $O/bench4_*.err; do echo "== $f"; tail -n 60 $f | cut -c1-200; done
date
