#!/bin/bash
# GPU call 2: the paths still behind switches (multi-warp scalar kernels, recomputed first sweep, device-side HO stencil)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02b
mkdir -p $O
exec > $O/call2.log 2>&1
set -x
date
export HH_SCALAR_FAST=1 HH_FUSE_RECOMPUTE=1 HH_HO_BUILD=device
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" || echo SMOKE_FAILED
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_headline_parity.py 2>&1 | tail -40
date
unset HH_SCALAR_FAST HH_FUSE_RECOMPUTE HH_HO_BUILD
HH_CHECK_ALL=1 timeout 600 python scripts/krylov_switch_check.py 257 > $O/switch_check2.jsonl 2> $O/switch_check2.err || echo SWITCH_CHECK_FAILED
cat $O/switch_check2.jsonl
tail -5 $O/switch_check2.err
date
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
HH_SCALAR_FAST=1 HH_FUSE_RECOMPUTE=1 $B --e2e-steps 2 > $O/bench2_all_16.json 2> $O/bench2_all_16.err
date
HH_SCALAR_FAST=1 $B --no-e2e > $O/bench2_scalar_16.json 2> $O/bench2_scalar_16.err
HH_FUSE_RECOMPUTE=1 $B --no-e2e > $O/bench2_recompute_16.json 2> $O/bench2_recompute_16.err
$B --no-e2e > $O/bench2_base_16.json 2> $O/bench2_base_16.err
date
HH_SCALAR_FAST=1 HH_FUSE_RECOMPUTE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --prec c64 --tol 1e-5 > $O/bench2_all_16_c64.json 2> $O/bench2_all_16_c64.err
date
for f in $O/bench2_*.json; do echo "== $f"; python - "$f" <<'PY'
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
for f in $O/bench2_*.err; do echo "== $f"; tail -n 3 $f; done
date
