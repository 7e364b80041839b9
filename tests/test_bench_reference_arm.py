"""bench.py --impl reference (the CPU arm the driver runs next to the product arm): one JSON line with the contract's
keys, on a small grid so that the CPU suite stays short.  No GPU is needed for this arm."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "33", "--steps", "1", "--warmup", "1",
           "--cpu-rhs", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "RHS/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"]
    # every step is a full solve to the tolerance (a measurement, not an extrapolation), on inputs the oracle built
    assert d["config"]["converged"] is True and d["config"]["relres_max"] <= 1e-6 and d["config"]["iterations_max"] >= 5
    assert d["steps"] >= 1 and d["steps_requested"] == 1


def test_reference_arm_does_not_load_the_product_library():
    """The CPU arm builds its inputs with oracle/helm_oracle.py and loads workloads.py by path: libhelmholtz_b200.so must not
    be mapped into that process (a defect in the product's host helpers must not be able to hide on both arms)."""
    code = ("import sys, os; sys.argv=['bench.py','--impl','reference','--n','17','--steps','1','--warmup','0','--cpu-rhs','1'];"
            "import runpy; runpy.run_path('bench.py', run_name='__main__');"
            "maps=open('/proc/self/maps').read(); print('PRODUCT_LOADED' if 'libhelmholtz_b200' in maps else 'PRODUCT_ABSENT')")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "PRODUCT_ABSENT" in r.stdout


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--n", "33"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
