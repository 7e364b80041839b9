#!/bin/bash
# GPU call 5: the defaults of this session on the whole GPU suite (incl. the 257^3 CPU-port parity file), the bench line
# with its CPU leg, the mapped-state A/B of the e2e pipeline, the ncu launch list of one step and ncu --set full captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02b
mkdir -p $O
exec > $O/call5.log 2>&1
set -x
date
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" || echo SMOKE_FAILED
HH_TEST_LOG=$O/parity_257.log timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30
cat $O/parity_257.log
date
timeout 800 python bench.py --steps 5 --warmup 3 --e2e-steps 2 > $O/bench5_n1.json 2> $O/bench5_n1.err
date
HH_MAPPED_STATE=0 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 2 > $O/bench5_n1_dmaflags.json 2> $O/bench5_n1_dmaflags.err
date
HH_HOST_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $O/bench5_trace.json 2> $O/bench5_trace.err
date
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches_r02b.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_launch_bench.log 2>&1
date
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fine3d_tma_prob|k_fine3d_tma_first|k_gmres_small_step_mw|k_combine|k_multiaxpy|k_multidot" --launch-skip 60 -c 14 -o $O/full_r02b -f python scripts/explore.py --n 257 --nrhs 16 --cycle W --pre 1 --post 2 --maxit 1 > $O/ncu_full.log 2>&1
ncu -i $O/full_r02b.ncu-rep --page raw --csv > $O/ncu_full.csv 2>/dev/null
ls -la $O/*.ncu-rep; rm -f $O/*.ncu-rep
date
for f in $O/bench5_*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", (d.get("e2e") or {}).get("value"),
              "ps", (d.get("e2e_point_sources") or {}).get("value"), d["config"].get("iterations_mean"), d["config"].get("true_relres_max_last_step"), d.get("clocks"))
        print("   host_link", (d.get("e2e") or {}).get("host_link"))
        print("   cpu", d.get("cpu_baseline"))
        pk = d["roofline"]["per_kernel"]
        for k, v in pk.items():
            print("   %-22s share %.3f avg_ms %.4f gbs %s frac %s n %d" % (k, v["share"], v["avg_ms"], v["gbs"], v["frac"], v["launches"]))
PY
done
for f in $O/bench5_*.err; do echo "== $f"; tail -n 40 $f | cut -c1-200; done
tail -n 5 $O/ncu_launch_bench.log $O/ncu_full.log
date
