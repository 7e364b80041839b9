#!/usr/bin/env python
"""bench.py -- RHS solves/s of the shifted-Laplacian multigrid Helmholtz solve on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[3], SURVEY.md section 8d "config 4"): 3-D 257^3-node random-smooth velocity
model, 10 points per wavelength, absorbing layer + Sommerfeld, 256 point sources on a 16 x 16 top-plane
grid, shift 0.2, 3-level W(1,2) damped-Jacobi Galerkin multigrid with an inexact Jacobi-GMRES(10) coarsest
solve, right-preconditioned FGMRES(5) to a 1e-6 relative residual, ComplexF64.

A step = one batched solve of `--nrhs` right-hand sides (a slice of the 256 sources) on every GPU.  Right-hand
sides are independent, so ranks shard them with no data-path collective (weak scaling: per-GPU batch fixed).

  python bench.py [--gpus N] [--steps K] [--warmup W]            product arm (one JSON line on rank 0)
  python bench.py --impl reference ...                            CPU arm: the oracle's C/OpenMP port of the
                                                                  reference algorithm on the host cores
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

METRIC = "rhs_solves_per_sec_to_1e-6_3d_257cubed"
UNIT = "RHS/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=257, help="nodes per dimension (257 = the named config)")
    ap.add_argument("--nrhs", type=int, default=16, help="right-hand sides per step per GPU")
    ap.add_argument("--prec", default="c128", choices=["c128", "c64", "mixed"],
                    help="mixed = ComplexF64 solve whose multigrid cycle runs in ComplexF32 (opt-in extension)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-rhs", type=int, default=4, help="RHS block of the CPU sample")
    ap.add_argument("--cpu-iters", type=int, default=5, help="preconditioned iterations per CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tol", type=float, default=1e-6, help="relative residual tolerance (1e-6 = the named metric)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------
def workload(pkg, n, tol=1e-6):
    """config 4 at n^3 nodes (sigma and pad scale with the grid so that small smoke sizes stay sensible)."""
    cfg = pkg.workloads.config4(n=n, sigma=8.0 * (n - 1) / 256, seed=1234, pad=max(4, 16 * (n - 1) // 256))
    mesh = pkg.getRegularMesh(cfg["domain"], cfg["n_cells"])
    m = cfg["m"]
    w = pkg.getMaximalFrequency(m, mesh)  # 10 points per wavelength
    gamma = cfg["gamma0_frac"] * w * np.ones(m.shape) + pkg.getABL(mesh.n + 1, True, cfg["pad"], w)
    srcs = pkg.workloads.point_sources_top_grid(mesh.n + 1, 16, 16)
    return dict(cfg=cfg, mesh=mesh, m=m, w=w, gamma=gamma, srcs=srcs,
                settings=dict(levels=3, shift=0.2, relax="Jac", relax_param=0.8, pre=1, post=2, cycle="W", coarse="GMRES",
                              coarse_iters=10, krylov="GMRES", inner=5, tol=tol, max_cycles=30))


def workload_name(n, nrhs, prec):
    return (f"config4: 3-D {n}^3 nodes, random-smooth velocity 1.5-4.5 km/s (seed 1234), 10 ppw, ABL+Sommerfeld, 256 point "
            f"sources on a 16x16 top-plane grid ({nrhs} per step per GPU), shift 0.2, 3-level W(1,2) Jacobi(0.8) Galerkin MG, "
            f"coarsest Jacobi-GMRES(10), FGMRES(5), tol 1e-6, {'ComplexF64' if prec == 'c128' else 'ComplexF32'}")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm = []
        self.reasons = set()
        self.sm_max = None
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag.is_set():
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(tag):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(tag)
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------------
def cpu_sample(wl, nrhs, iters, threads_hint=None):
    """Time the oracle's C/OpenMP port (assembled CSR Galerkin MG + FGMRES) on a bounded sample:
    `iters` preconditioned FGMRES iterations on the first `nrhs` sources.  Returns seconds per
    (iteration x RHS), set-up seconds and the thread count."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_c  # test infrastructure: the timed CPU arm only

    s = wl["settings"]
    mesh = wl["mesh"]
    oc = oracle_c.OracleC(mesh.n + 1, mesh.h, wl["m"], wl["gamma"], wl["w"], True, True, s["shift"], s["levels"],
                          s["relax_param"], s["pre"], s["post"], s["cycle"], s["coarse_iters"])
    N = int(np.prod(mesh.n + 1))
    B = np.zeros((N, nrhs), dtype=np.complex128, order="F")
    pkg = graft.load_package()
    for c, src in enumerate(wl["srcs"][:nrhs]):
        B[pkg.loc2cs(mesh.n + 1, src) - 1, c] = 1.0 / mesh.h[0] ** 2
    X, it, rr, secs = oc.solve(B, inner=s["inner"], max_cycles=s["max_cycles"], tol=s["tol"], max_prec=iters)
    done_iters = int(it.max())
    per = secs / max(done_iters, 1) / nrhs
    out = dict(sec_per_iter_rhs=per, setup_seconds=oc.setup_seconds, threads=oc.threads, iters_done=done_iters, secs=secs)
    oc.close()
    return out


def run_reference(a):
    """CPU arm: rank 0 only.  Each step is a bounded sample (cpu_iters preconditioned FGMRES iterations on a
    block of cpu_rhs sources); RHS/s = 1 / (seconds per iteration-RHS x iterations to 1e-6)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = graft.load_package()
    wl = workload(pkg, a.n, a.tol)
    iters_needed = iterations_to_tol(a.n)
    vals = []
    setup = None
    threads = None
    for stp in range(a.warmup + a.steps):
        r = cpu_sample(wl, a.cpu_rhs, a.cpu_iters)
        setup, threads = r["setup_seconds"], r["threads"]
        if stp >= a.warmup:
            vals.append(r["sec_per_iter_rhs"])
        if stp == 0 and a.warmup + a.steps > 2 and r["secs"] + r["setup_seconds"] > 60:
            # keep the whole run within minutes on slow hosts: one warm-up + one timed sample
            r2 = cpu_sample(wl, a.cpu_rhs, a.cpu_iters)
            vals = [r2["sec_per_iter_rhs"]]
            break
    per = float(np.mean(vals))
    value = 1.0 / (per * iters_needed)
    sample = (f"{a.cpu_iters} preconditioned FGMRES(5) iterations on a block of {a.cpu_rhs} sources of the same 257^3 workload; "
              f"RHS/s = 1/(s per iteration-RHS x {iters_needed} iterations to 1e-6, the count this algorithm needs on this "
              f"workload: identical on GPU and CPU port)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * per * a.cpu_iters * a.cpu_rhs, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": workload_name(a.n, a.nrhs, a.prec), "note": "restated reference CPU path (C/OpenMP port of the "
                   "oracle: assembled CSR operator, Galerkin RAP hierarchy, OpenMP SpMV); the Julia reference cannot run here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "setup_seconds": setup},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def iterations_to_tol(n):
    """Preconditioner applications FGMRES(5) needs to reach 1e-6 on this workload (mean over sources).  Measured
    by the product arm (identical counts in the CPU port, tests/test_oracle_c.py); recorded in profiles/."""
    p = os.path.join(ROOT, "profiles", "iterations.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if str(n) in d:
                return float(d[str(n)])
        except Exception:
            pass
    return {257: 29.0, 129: 24.0, 65: 18.0}.get(n, 29.0)


# ----------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG=VERSION makes NCCL print its banner on stdout, ahead of the one JSON line this script owes
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = graft.load_package()
    lib = pkg._lib.load()
    wl = workload(pkg, a.n, a.tol)
    s = wl["settings"]
    mesh = wl["mesh"]
    prec = np.complex64 if a.prec == "c64" else np.complex128
    tdt = torch.complex64 if a.prec == "c64" else torch.complex128
    MG = pkg.getMGparam(prec, pkg.Int64, s["levels"], 1, s["max_cycles"], s["tol"], s["relax"], s["relax_param"], s["pre"],
                        s["post"], s["cycle"], s["coarse"], coarseIters=s["coarse_iters"])
    if a.prec == "mixed":
        MG.cyclePrecision = pkg.ComplexF32
    hp = pkg.HelmholtzParam(mesh, wl["gamma"], wl["m"].ravel(order="F"), wl["w"], True, True)
    Ainv = pkg.getShiftedLaplacianMultigridSolver(hp, MG, s["shift"], s["krylov"], s["inner"])
    Ainv.devices = [local]
    hd = pkg.api._ensure_hierarchy(Ainv, 0)
    N = int(np.prod(mesh.n + 1))
    nodes = mesh.n + 1
    amp = 1.0 / mesh.h[0] ** 2
    all_idx = np.array([pkg.loc2cs(nodes, src) - 1 for src in wl["srcs"]], dtype=np.int64)
    nsrc = len(all_idx)

    def step_sources(step):
        # every rank works on its own slice of the 256 sources (column sharding, no collective)
        base = (step * world + rank) * a.nrhs
        return [(base + c) % nsrc for c in range(a.nrhs)]

    B = torch.zeros((a.nrhs, N), dtype=tdt, device="cuda")
    X = torch.empty_like(B)
    rows = torch.arange(a.nrhs, device="cuda")
    prev = None

    def load_sources(step):
        nonlocal prev
        cols = torch.as_tensor(all_idx[step_sources(step)], device="cuda")
        if prev is not None:
            B[rows, prev] = 0
        B[rows, cols] = amp
        prev = cols

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    iters_log = []

    def one_step(step):
        load_sources(step)
        pkg.solveLinearSystem_(None, B, X, Ainv)
        iters_log.append(Ainv.iterations.copy())

    for w_ in range(a.warmup):
        one_step(w_)
    lib.hh_profile_enable(hd.h, 1)
    lib.hh_profile_reset(hd.h)
    l0 = C.c_int64()
    lib.hh_get_counters(hd.h, None, None, None, C.byref(l0))
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters_log.clear()
    ev0.record()
    for k in range(a.steps):
        one_step(a.warmup + k)
    ev1.record()
    barrier()
    sampler.stop_flag.set()
    ms = ev0.elapsed_time(ev1)
    l1 = C.c_int64()
    lib.hh_get_counters(hd.h, None, None, None, C.byref(l1))
    launches = int(l1.value - l0.value)
    # parity guard on the last step: true residual of the un-shifted operator
    Hop = pkg.HelmholtzOperator(hd)
    R = Hop.matvec(X) - B
    true_res = float((torch.linalg.vector_norm(R, dim=1) / torch.linalg.vector_norm(B, dim=1)).max())
    del R
    # per-kernel device time (CUDA events on the launching stream, recorded during the timed region)
    tags = []
    for t in range(lib.hh_profile_num_tags()):
        cnt, tms, by = C.c_int64(), C.c_double(), C.c_double()
        lib.hh_profile_get(hd.h, t, C.byref(cnt), C.byref(tms), C.byref(by))
        if cnt.value:
            tags.append(dict(kernel=lib.hh_profile_tag_name(t).decode(), launches=int(cnt.value), ms=tms.value, bytes=by.value))
    # finer view: launches doing identical work (same kernel class and bytes) -> the dominant kernel
    entries = []
    for e in range(lib.hh_profile_num_entries(hd.h)):
        tg, cnt, tms, by = C.c_int(), C.c_int64(), C.c_double(), C.c_double()
        lib.hh_profile_entry(hd.h, e, C.byref(tg), C.byref(cnt), C.byref(tms), C.byref(by))
        if cnt.value and by.value > 0:
            entries.append(dict(kernel=lib.hh_profile_tag_name(tg.value).decode(), launches=int(cnt.value), ms=tms.value,
                                bytes_per_launch=by.value))
    lib.hh_profile_enable(hd.h, 0)
    tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max = float(tmax.item())
    value = a.nrhs * a.steps * world / (ms_max / 1e3)

    # ---- e2e: the same steps through the plugin call with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not a.no_e2e:
        Bh = torch.zeros((a.nrhs, N), dtype=tdt).pin_memory()
        Xh_t = torch.empty((a.nrhs, N), dtype=tdt).pin_memory()
        Bh_np, Xh = Bh.numpy().T, Xh_t.numpy().T  # N x nrhs column-major views of the pinned buffers
        es = 8 if a.prec == "c64" else 16
        t_e2e = []
        for k in range(1 + a.e2e_steps):
            Bh.zero_()
            for c, sidx in enumerate(step_sources(1000 + k)):
                Bh[c, all_idx[sidx]] = amp
            barrier()
            t0 = time.perf_counter()
            pkg.solveLinearSystem_(None, Bh_np, Xh, Ainv)
            chk = float(abs(Xh[all_idx[step_sources(1000 + k)[0]], 0]))  # the result is in host memory: consume it
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if k > 0:
                t_e2e.append(dt)
            assert np.isfinite(chk)
        te = torch.tensor([float(np.mean(t_e2e))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": a.nrhs * world / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": int(N * a.nrhs * es),
               "d2h_bytes_per_step": int(N * a.nrhs * es), "steps": a.e2e_steps,
               "api": "solveLinearSystem!(A, B_host, X_host, Ainv) -> hh_solve (pinned host B and X)"}

    if rank == 0:
        peak, peak_src = measured_peak()
        tags.sort(key=lambda d: -d["ms"])
        tot = sum(d["ms"] for d in tags)
        entries.sort(key=lambda d: -d["ms"])
        de = entries[0]  # dominant kernel = the set of identical launches with the largest share of the step
        dom = dict(kernel=de["kernel"], launches=de["launches"], ms=de["ms"], bytes=de["bytes_per_launch"] * de["launches"])
        achieved = dom["bytes"] / dom["ms"] / 1e6  # GB/s
        per_kernel = {d["kernel"]: {"launches": d["launches"], "share": round(d["ms"] / tot, 4),
                                    "avg_ms": round(d["ms"] / d["launches"], 4),
                                    "gbs": round(d["bytes"] / d["ms"] / 1e6, 1) if d["bytes"] else None} for d in tags}
        its = np.concatenate(iters_log) if iters_log else np.zeros(1)
        cpu = None
        if not a.no_cpu_baseline and world == 1:
            try:
                r = cpu_sample(wl, a.cpu_rhs, a.cpu_iters)
                v = 1.0 / (r["sec_per_iter_rhs"] * float(its.mean()))
                cpu = {"value": v, "unit": UNIT, "cores": r["threads"], "kind": "port",
                       "sample": f"{r['iters_done']} preconditioned FGMRES(5) iterations on a block of {a.cpu_rhs} sources of "
                                 f"the same workload ({r['secs']:.1f} s, set-up {r['setup_seconds']:.1f} s excluded); RHS/s = "
                                 f"1/(s per iteration-RHS x {its.mean():.1f} iterations to 1e-6 as measured on the GPU run)"}
            except Exception as e:  # the CPU arm must never take the GPU number down with it
                cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {e}"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128 (ComplexF32 multigrid cycle)" if a.prec == "mixed" else a.prec, "data": "synthetic",
            "config": {"workload": workload_name(a.n, a.nrhs, a.prec), "rhs_per_step_per_gpu": a.nrhs,
                       "parallelism": f"rhs-sharding x{world} (independent columns, no data-path collective)",
                       "l2": "inputs larger than L2: every vector block is %.0f MB, no flush needed" % (N * a.nrhs * (8 if a.prec == "c64" else 16) / 1e6),
                       "rel_tol": a.tol, "iterations_mean": float(its.mean()), "iterations_max": int(its.max()),
                       "true_relres_max_last_step": true_res},
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": sampler.result(),
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "algorithmic_bytes_per_launch": de["bytes_per_launch"],
                         "traffic": ncu_traffic(dom["kernel"]), "peak_source": peak_src,
                         "share_of_step": dom["ms"] / tot, "avg_launch_ms": dom["ms"] / dom["launches"],
                         "whole_step_algorithmic_gbs": sum(d["bytes"] for d in tags) / tot / 1e6,
                         "per_kernel": per_kernel},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
