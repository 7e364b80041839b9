#!/usr/bin/env python
"""bench.py -- RHS solves/s of the shifted-Laplacian multigrid Helmholtz solve on B200 (BASELINE.json metric).

Default workload (BASELINE.json configs[3], SURVEY.md section 8d "config 4"): 3-D 257^3-node random-smooth velocity
model, 10 points per wavelength, absorbing layer + Sommerfeld, 256 point sources on a 16 x 16 top-plane
grid, shift 0.2, 3-level W(1,2) damped-Jacobi Galerkin multigrid with an inexact Jacobi-GMRES(10) coarsest
solve, right-preconditioned FGMRES(5) to a 1e-6 relative residual, ComplexF64.  `--config 3` / `--config 2` give the
same line for BASELINE configs[2] (3-D 129^3 layered model with attenuation, 16 sources) and configs[1] (2-D SEG salt
model, 64 sources).

A step = one batched solve of `--nrhs` right-hand sides (a slice of the sources) on every GPU.  Right-hand
sides are independent, so ranks shard them with no data-path collective (weak scaling: per-GPU batch fixed).
With more than one rank the line also carries "slab": ONE 257^3 problem split into slabs over all ranks (NCCL halo
exchange + all-reduced dots, BASELINE config 5's decomposition) timed and compared with the whole-grid solve.

  python bench.py [--gpus N] [--steps K] [--warmup W]            product arm (one JSON line on rank 0)
  python bench.py --impl reference ...                            CPU arm: the oracle's C/OpenMP port of the
                                                                  reference algorithm on the host cores; every step is a
                                                                  FULL solve to the tolerance of one RHS block
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib.util
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

UNIT = "RHS/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4], help="BASELINE config (4 = the named headline)")
    ap.add_argument("--n", type=int, default=0, help="nodes per dimension of the 3-D configs (default: the named size)")
    ap.add_argument("--nrhs", type=int, default=0, help="right-hand sides per step per GPU (default: 32 for config 4 = 256 sources "
                                                         "over 8 GPUs in one step, 16 for config 3, 64 for config 2)")
    ap.add_argument("--prec", default="c128", choices=["c128", "c64", "mixed"],
                    help="mixed = ComplexF64 solve whose multigrid cycle runs in ComplexF32 (opt-in extension)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-rhs", type=int, default=0, help="RHS block of one CPU step (default: 16 in the reference arm, "
                    "2 in the product arm's cpu_baseline sample)")
    ap.add_argument("--ref-budget-s", type=float, default=270.0, help="wall-clock budget of the reference arm's steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-slab", action="store_true")
    ap.add_argument("--slab-nrhs", type=int, default=8)
    ap.add_argument("--tol", type=float, default=1e-6, help="relative residual tolerance (1e-6 = the named metric)")
    a = ap.parse_args()
    if a.n == 0:
        a.n = {4: 257, 3: 129, 2: 257}[a.config]
    if a.nrhs == 0:
        a.nrhs = {2: 64, 3: 16, 4: 32}[a.config]
    return a


def metric_name(a):
    if a.config == 4:
        return f"rhs_solves_per_sec_to_{a.tol:g}_3d_{a.n}cubed"
    if a.config == 3:
        return f"rhs_solves_per_sec_to_{a.tol:g}_3d_{a.n}cubed_layered"
    return f"rhs_solves_per_sec_to_{a.tol:g}_2d_seg_salt_257x129"


# ----------------------------------------------------------------------------------------------------
def load_workloads_standalone():
    """helmholtz.jl_b200/workloads.py by path (numpy only): the reference arm must not load the product package."""
    spec = importlib.util.spec_from_file_location("hh_workloads_standalone", os.path.join(ROOT, "helmholtz.jl_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload(a, wl_mod, host):
    """The benchmark problem.  `host` supplies the reference's set-up functions (getRegularMesh, getMaximalFrequency,
    getABL, loc2cs): the product package in the product arm, oracle/helm_oracle.py in the CPU arm -- so that neither arm's
    inputs depend on the other's code."""
    n, tol = a.n, a.tol
    if a.config == 4:
        # sigma and pad scale with the grid so that small smoke sizes stay sensible
        cfg = wl_mod.config4(n=n, sigma=8.0 * (n - 1) / 256, seed=1234, pad=max(4, 16 * (n - 1) // 256))
        settings = dict(levels=3, shift=0.2, relax="Jac", relax_param=0.8, pre=1, post=2, cycle="W", coarse="GMRES",
                        coarse_iters=10, krylov="GMRES", inner=5, tol=tol, max_cycles=30)
        gamma0 = None
        grid = (16, 16)
    elif a.config == 3:
        cfg = wl_mod.config3(n=n)
        settings = dict(levels=3, shift=0.2, relax="Jac", relax_param=0.8, pre=1, post=2, cycle="W", coarse="GMRES",
                        coarse_iters=10, krylov="GMRES", inner=5, tol=tol, max_cycles=30)
        gamma0 = "att"
        grid = (4, 4)
    else:
        vp = np.load(os.path.join(ROOT, "tests", "golden", "seg_salt_vp.npz"))["vp_ms"]
        cfg = wl_mod.config2(vp)
        settings = dict(levels=3, shift=0.2, relax="Jac", relax_param=0.75, pre=2, post=2, cycle="V", coarse="GMRES",
                        coarse_iters=10, krylov="GMRES", inner=5, tol=tol, max_cycles=60)
        gamma0 = None
        grid = (64,)
    mesh = host.getRegularMesh(cfg["domain"], cfg["n_cells"])
    m = cfg["m"]
    w = host.getMaximalFrequency(m, mesh)  # 10 points per wavelength
    nodes = np.asarray(mesh.n) + 1
    base = cfg["gamma0_frac"] * w * (cfg["att_profile"] if gamma0 == "att" else np.ones(m.shape))
    gamma = base + host.getABL(nodes, True, cfg["pad"], w)
    srcs = wl_mod.point_sources_top_grid(nodes, *grid)
    idx = np.array([host.loc2cs(nodes, s) - 1 for s in srcs], dtype=np.int64)
    return dict(cfg=cfg, mesh=mesh, nodes=nodes, m=m, w=w, gamma=gamma, srcs=srcs, idx=idx, settings=settings)


def workload_name(a):
    p = {"c128": "ComplexF64", "c64": "ComplexF32", "mixed": "ComplexF64 (ComplexF32 multigrid cycle)"}[a.prec]
    if a.config == 4:
        return (f"config4: 3-D {a.n}^3 nodes, random-smooth velocity 1.5-4.5 km/s (seed 1234), 10 ppw, ABL+Sommerfeld, 256 point "
                f"sources on a 16x16 top-plane grid ({a.nrhs} per step per GPU), shift 0.2, 3-level W(1,2) Jacobi(0.8) Galerkin MG, "
                f"coarsest Jacobi-GMRES(10), FGMRES(5), tol {a.tol:g}, {p}")
    if a.config == 3:
        return (f"config3: 3-D {a.n}^3 nodes, layered velocity 1.5-5 km/s with depth-dependent attenuation, 10 ppw, ABL+Sommerfeld, "
                f"16 point sources on a 4x4 top-plane grid ({a.nrhs} per step), shift 0.2, 3-level W(1,2) Jacobi(0.8) Galerkin MG, "
                f"coarsest Jacobi-GMRES(10), FGMRES(5), tol {a.tol:g}, {p}")
    return (f"config2: 2-D SEG salt model on 257x129 nodes, 10 ppw, ABL+Sommerfeld, 64 point sources along the top row "
            f"({a.nrhs} per step per GPU), shift 0.2, 3-level V(2,2) Jacobi(0.75) Galerkin MG, coarsest Jacobi-GMRES(10), FGMRES(5), "
            f"tol {a.tol:g}, {p}")


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


def ncu_traffic(cls):
    """Average DRAM bytes per launch of a kernel class from the committed ncu pass over one bench step
    (profiles/ncu_traffic.json, written by scripts/ncu_class_traffic.py), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            v = json.load(open(p)).get(cls)
            return v.get("dram_bytes_per_launch") if isinstance(v, dict) else v
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------------
def cpu_arm_inputs(a):
    """Inputs of the CPU arm, built by the oracle's own restatement of the reference's set-up functions."""
    ho = graft.load_oracle()

    class Host:
        getRegularMesh = staticmethod(ho.getRegularMesh)
        getMaximalFrequency = staticmethod(ho.getMaximalFrequency)
        getABL = staticmethod(ho.getABL)
        loc2cs = staticmethod(ho.loc2cs)

    return workload(a, load_workloads_standalone(), Host)


def cpu_full_solve(wl, cols, oc=None):
    """One step of the CPU arm: the oracle's C/OpenMP port (assembled CSR Galerkin MG + FGMRES) solves the block of
    sources `cols` to the tolerance.  Returns (seconds, iterations, relres, oracle handle)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_c  # test infrastructure: the timed CPU arm only

    s = wl["settings"]
    mesh = wl["mesh"]
    if oc is None:
        oc = oracle_c.OracleC(wl["nodes"], mesh.h, wl["m"], wl["gamma"], wl["w"], True, True, s["shift"], s["levels"],
                              s["relax_param"], s["pre"], s["post"], s["cycle"], s["coarse_iters"])
    N = int(np.prod(wl["nodes"]))
    B = np.zeros((N, len(cols)), dtype=np.complex128, order="F")
    for c, sidx in enumerate(cols):
        B[wl["idx"][sidx], c] = 1.0 / mesh.h[0] ** 2
    X, it, rr, secs = oc.solve(B, inner=s["inner"], max_cycles=s["max_cycles"], tol=s["tol"])
    return secs, it, rr, oc


def run_reference(a):
    """CPU arm: rank 0 only.  A step = a FULL FGMRES solve to the tolerance of a block of `cpu_rhs` sources of the same
    workload by the C/OpenMP restatement of the reference's CPU algorithm on all host cores; RHS/s = block / seconds.
    The requested steps are run for as long as the wall-clock budget allows (at least one timed step); `steps` is the
    number actually timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = cpu_arm_inputs(a)
    block = a.cpu_rhs or 16
    nsrc = len(wl["idx"])
    t_start = time.perf_counter()
    oc = None
    secs_all, its_all, rr_all = [], [], []
    warm_done = 0
    last = None
    stp = 0
    while len(secs_all) < a.steps:
        cols = [(stp * block + c) % nsrc for c in range(block)]
        elapsed = time.perf_counter() - t_start
        if last is not None and elapsed + 1.1 * last > a.ref_budget_s and secs_all:
            break  # the next full solve would not fit the budget
        secs, it, rr, oc = cpu_full_solve(wl, cols, oc)
        last = secs
        stp += 1
        # warm-up steps are only taken when they are cheap next to the budget (a CPU solve has no clock ramp to wait for)
        if warm_done < a.warmup and (warm_done + 2) * secs * 1.1 + oc.setup_seconds < a.ref_budget_s / 3:
            warm_done += 1
            continue
        secs_all.append(secs)
        its_all.append(it.copy())
        rr_all.append(rr.copy())
    per_step = float(np.mean(secs_all))
    value = block / per_step
    its = np.concatenate(its_all)
    rr = np.concatenate(rr_all)
    sample = (f"every step = full FGMRES(5) solve to {a.tol:g} of a block of {block} sources of the same workload "
              f"({len(secs_all)} timed step(s) of {a.steps} requested within a {a.ref_budget_s:.0f} s budget, {warm_done} warm-up; "
              f"set-up {oc.setup_seconds:.1f} s excluded): measured iterations {int(its.min())}-{int(its.max())}, relres max {rr.max():.2e}")
    line = {
        "impl": "reference", "metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": len(secs_all),
        "steps_requested": a.steps, "warmup": warm_done, "ms_per_step": 1e3 * per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": workload_name(a), "rhs_per_step": block,
                   "iterations_mean": float(its.mean()), "iterations_max": int(its.max()), "relres_max": float(rr.max()),
                   "converged": bool((rr <= a.tol).all()),
                   "note": "restated reference CPU path (C/OpenMP port of the oracle: assembled CSR operator, Galerkin RAP hierarchy, "
                           "OpenMP SpMV over the RHS block); inputs built by oracle/helm_oracle.py; the Julia reference cannot run here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": oc.threads, "kind": "port", "sample": sample,
                         "setup_seconds": oc.setup_seconds},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    oc.close()
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
def profile_tables(lib, hd):
    tags = []
    for t in range(lib.hh_profile_num_tags()):
        cnt, tms, by = C.c_int64(), C.c_double(), C.c_double()
        lib.hh_profile_get(hd.h, t, C.byref(cnt), C.byref(tms), C.byref(by))
        if cnt.value:
            tags.append(dict(kernel=lib.hh_profile_tag_name(t).decode(), launches=int(cnt.value), ms=tms.value, bytes=by.value))
    return tags


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
    # pinned staging buffers of the e2e legs on the GPU's own NUMA node (PCIe copies of 8 ranks crossing the socket link
    # were the e2e limiter at 8 GPUs in round 1)
    all_cpus = os.sched_getaffinity(0)
    numa = pkg.sharding.bind_to_gpu_numa_node(local)
    wl = workload(a, pkg.workloads, pkg)
    s = wl["settings"]
    mesh = wl["mesh"]
    prec = np.complex64 if a.prec == "c64" else np.complex128
    tdt = torch.complex64 if a.prec == "c64" else torch.complex128
    es = 8 if a.prec == "c64" else 16

    def new_solver(slabs=None):
        MG = pkg.getMGparam(prec, pkg.Int64, s["levels"], 1, s["max_cycles"], s["tol"], s["relax"], s["relax_param"], s["pre"],
                            s["post"], s["cycle"], s["coarse"], coarseIters=s["coarse_iters"])
        if a.prec == "mixed":
            MG.cyclePrecision = pkg.ComplexF32
        hp = pkg.HelmholtzParam(mesh, wl["gamma"], wl["m"].ravel(order="F"), wl["w"], True, True)
        A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, s["shift"], s["krylov"], s["inner"])
        A.devices = [local]
        A.slabs = slabs
        return A

    Ainv = new_solver()
    hd = pkg.api._ensure_hierarchy(Ainv, 0)
    N = int(np.prod(wl["nodes"]))
    nodes = wl["nodes"]
    amp = 1.0 / mesh.h[0] ** 2
    all_idx = wl["idx"]
    nsrc = len(all_idx)

    def step_sources(step):
        # every rank works on its own slice of the sources (column sharding, no collective)
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
    torch.cuda.empty_cache()  # the library allocates with cudaMalloc: hand torch's cached blocks back before the e2e legs
    # per-kernel-class device time (CUDA events on the launching stream, recorded during the timed region)
    tags = profile_tables(lib, hd)
    lib.hh_profile_enable(hd.h, 0)
    tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max = float(tmax.item())
    value = a.nrhs * a.steps * world / (ms_max / 1e3)

    # ---- e2e: the same steps through the plugin call with HOST buffers (pinned), copies inside the timed region.
    #      (1) dense B as the reference's callers pass it; (2) point sources as (index, value) lists: no dense B upload.
    e2e = None
    e2e_ps = None
    if not a.no_e2e:
        # pinned B and X of one call; every rank of the node pins its own pair, so the block of one call is halved until the
        # node's free host memory holds all of them twice over (the device-resident `value` above is not affected)
        ne = a.nrhs
        try:
            avail = next(int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
            while ne > 1 and 1.3 * (2 * N * ne * es) * world > 0.8 * avail:
                ne //= 2
        except Exception:
            pass
        # every rank must end up with the same block: a failed pinned allocation on any rank halves it for all of them
        while True:
            try:
                Bh = torch.zeros((ne, N), dtype=tdt, pin_memory=True)
                Xh_t = torch.empty((ne, N), dtype=tdt, pin_memory=True)
                ok = 1
            except Exception:
                Bh = Xh_t = None
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) or ne == 1:
                break
            Bh = Xh_t = None
            ne //= 2
        if Bh is None:
            raise SystemExit("bench.py: cannot pin host memory for one right-hand side")
        Bh_np, Xh = Bh.numpy().T, Xh_t.numpy().T  # N x nrhs column-major views of the pinned buffers
        t_e2e, t_ps = [], []
        for k in range(1 + a.e2e_steps):
            cols = step_sources(1000 + k)[:ne]
            Bh.zero_()
            for c, sidx in enumerate(cols):
                Bh[c, all_idx[sidx]] = amp
            barrier()
            t0 = time.perf_counter()
            pkg.solveLinearSystem_(None, Bh_np, Xh, Ainv)
            chk = float(abs(Xh[all_idx[cols[0]], 0]))  # the result is in host memory: consume it
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if k > 0:
                t_e2e.append(dt)
            assert np.isfinite(chk)
            barrier()
            t0 = time.perf_counter()
            pkg.solvePointSources_(Ainv, [wl["srcs"][sidx] for sidx in cols], Xh, np.full(ne, amp))
            chk = float(abs(Xh[all_idx[cols[0]], 0]))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if k > 0:
                t_ps.append(dt)
            assert np.isfinite(chk)
        # the host link itself: plain pinned copies of one block, every rank at once (what bounds any e2e figure here)
        link = []
        for src, dst in ((X[:ne], Xh_t), (Bh, X[:ne])):
            barrier()
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            link.append(time.perf_counter() - t0)
        te = torch.tensor([float(np.mean(t_e2e)), float(np.mean(t_ps)), link[0], link[1]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        blk = N * ne * es / 1e9
        host_link = {"d2h_gbs_all_ranks": blk * world / float(te[2]), "h2d_gbs_all_ranks": blk * world / float(te[3]),
                     "note": "pinned copies of one B / X block per rank, all ranks concurrently; the result block of a step "
                             "(d2h_bytes_per_step) cannot reach the host faster than this"}
        e2e = {"value": ne * world / float(te[0]), "unit": UNIT, "h2d_bytes_per_step": int(N * ne * es),
               "d2h_bytes_per_step": int(N * ne * es), "steps": a.e2e_steps, "rhs_per_call_per_gpu": ne,
               "api": "solveLinearSystem!(A, B_host, X_host, Ainv) -> hh_solve (pinned host B and X)", "host_link": host_link}
        e2e_ps = {"value": ne * world / float(te[1]), "unit": UNIT, "h2d_bytes_per_step": int(ne * 24),
                  "d2h_bytes_per_step": int(N * ne * es), "steps": a.e2e_steps, "rhs_per_call_per_gpu": ne,
                  "api": "solvePointSources!(Ainv, srcs, X_host) -> hh_solve_point_sources (no dense B; pinned host X)"}
        del Bh, Xh_t

    # ---- one grid split into slabs over all ranks (config 5's decomposition; NCCL halos + all-reduced dots)
    slab = None
    if world > 1 and not a.no_slab and a.config != 2:
        slab = slab_section(a, pkg, lib, torch, dist, wl, new_solver, Ainv, rank, world, local, barrier, tdt)

    if rank == 0:
        peak, peak_src = measured_peak()
        tags.sort(key=lambda d: -d["ms"])
        tot = sum(d["ms"] for d in tags)
        # dominant kernel = the kernel CLASS with the largest share of the step's device time (classes with algorithmic
        # bytes only: scalar kernels and set-up have none)
        dom = next(d for d in tags if d["bytes"] > 0)
        achieved = dom["bytes"] / dom["ms"] / 1e6  # GB/s
        per_kernel = {d["kernel"]: {"launches": d["launches"], "share": round(d["ms"] / tot, 4),
                                    "avg_ms": round(d["ms"] / d["launches"], 4),
                                    "gbs": round(d["bytes"] / d["ms"] / 1e6, 1) if d["bytes"] else None,
                                    "frac": round(d["bytes"] / d["ms"] / 1e6 / peak, 3) if d["bytes"] else None,
                                    "traffic": ncu_traffic(d["kernel"])} for d in tags}
        its = np.concatenate(iters_log) if iters_log else np.zeros(1)
        cpu = None
        if not a.no_cpu_baseline and world == 1:
            try:
                os.sched_setaffinity(0, all_cpus)  # the CPU arm uses every host core
                blk = a.cpu_rhs or 2
                wl_cpu = cpu_arm_inputs(a)  # built by the oracle, independent of the product's host helpers
                secs, itc, rrc, oc = cpu_full_solve(wl_cpu, [(7 * c) % nsrc for c in range(blk)])
                cpu = {"value": blk / secs, "unit": UNIT, "cores": oc.threads, "kind": "port",
                       "sample": f"full FGMRES(5) solve to {a.tol:g} of a block of {blk} sources of the same workload by the oracle's "
                                 f"C/OpenMP port ({secs:.1f} s, set-up {oc.setup_seconds:.1f} s excluded; iterations "
                                 f"{int(itc.min())}-{int(itc.max())}, relres max {rrc.max():.2e}); the reference arm "
                                 f"(--impl reference) times blocks of 16"}
                oc.close()
            except Exception as e:  # the CPU arm must never take the GPU number down with it
                cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {e}"}
        line = {
            "metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128 (ComplexF32 multigrid cycle)" if a.prec == "mixed" else a.prec, "data": "synthetic",
            "config": {"workload": workload_name(a), "rhs_per_step_per_gpu": a.nrhs,
                       "parallelism": f"rhs-sharding x{world} (independent columns, no data-path collective)",
                       "l2": "inputs larger than L2: every vector block is %.0f MB, no flush needed" % (N * a.nrhs * es / 1e6),
                       "rel_tol": a.tol, "iterations_mean": float(its.mean()), "iterations_max": int(its.max()),
                       "true_relres_max_last_step": true_res, "host_numa_binding_rank0": numa},
            "e2e": e2e,
            "e2e_point_sources": e2e_ps,
            "gpu_launches": launches,
            "clocks": sampler.result(),
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "algorithmic_bytes_per_launch": dom["bytes"] / dom["launches"],
                         "traffic": ncu_traffic(dom["kernel"]), "peak_source": peak_src,
                         "share_of_step": dom["ms"] / tot, "avg_launch_ms": dom["ms"] / dom["launches"],
                         "launches": dom["launches"],
                         "rule": "dominant = kernel class with the largest share of the step's device time; achieved = the class's "
                                 "algorithmic bytes (SURVEY 8d formulas, DESIGN.md section 3) / its CUDA-event time; per-launch figures "
                                 "are class averages",
                         "whole_step_algorithmic_gbs": sum(d["bytes"] for d in tags) / tot / 1e6,
                         "per_kernel": per_kernel},
            "cpu_baseline": cpu,
        }
        if slab is not None:
            line["slab"] = slab
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def slab_section(a, pkg, lib, torch, dist, wl, new_solver, Ainv, rank, world, local, barrier, tdt):
    """ONE grid of the workload split into slabs along the last dimension over all ranks (hh_create_slab_nccl): the
    same sources solved (a) slab-decomposed, timed on the device, max over ranks, and (b) whole-grid on every rank's own
    GPU; the slab solution of this rank's planes is compared with the whole-grid one."""
    try:
        mesh, nodes = wl["mesh"], wl["nodes"]
        k = a.slab_nrhs
        amp = 1.0 / mesh.h[0] ** 2
        # the timed region sized the solver's work memory and staging slots for the whole step block: release them (the
        # hierarchy is rebuilt on demand for the whole-grid solve of the k slab sources below)
        pkg.clear(Ainv.MG)
        torch.cuda.empty_cache()
        cols = [(13 * c) % len(wl["idx"]) for c in range(k)]
        gidx = wl["idx"][cols]
        N = int(np.prod(nodes))
        # (b) whole grid, same tolerance
        Bw = torch.zeros((k, N), dtype=tdt, device="cuda")
        Bw[torch.arange(k, device="cuda"), torch.as_tensor(gidx, device="cuda")] = amp
        Xw = torch.empty_like(Bw)
        pkg.solveLinearSystem_(None, Bw, Xw, Ainv)
        it_whole = Ainv.iterations.copy()
        del Bw
        # (a) slabs
        A = new_solver(pkg.sharding.nccl_slabs())
        t0 = time.perf_counter()
        hd = pkg.api._ensure_hierarchy(A, 0)
        barrier()
        t_setup = time.perf_counter() - t0
        k0, k1 = hd.planes
        plane = int(nodes[0] * nodes[1])
        Nown = plane * (k1 - k0)
        B = torch.zeros((k, Nown), dtype=tdt, device="cuda")
        for c, gi in enumerate(gidx):
            if plane * k0 <= gi < plane * k1:
                B[c, gi - plane * k0] = amp
        X = torch.empty_like(B)
        pkg.solveLinearSystem_(None, B, X, A)  # warm-up
        lib.hh_profile_enable(hd.h, 1)
        lib.hh_profile_reset(hd.h)
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        nsteps = 2
        for _ in range(nsteps):
            pkg.solveLinearSystem_(None, B, X, A)
        ev1.record()
        barrier()
        sampler.stop_flag.set()
        ms = ev0.elapsed_time(ev1)
        tags = profile_tables(lib, hd)
        lib.hh_profile_enable(hd.h, 0)
        it_slab = A.iterations.copy()
        # parity: this rank's planes of the slab solution against the whole-grid solution
        diff = X - Xw[:, plane * k0:plane * k1]
        num = torch.linalg.vector_norm(diff, dim=1) ** 2
        den = torch.linalg.vector_norm(Xw[:, plane * k0:plane * k1], dim=1) ** 2
        R = pkg.HelmholtzOperator(hd).matvec(X) - B
        rn = torch.linalg.vector_norm(R, dim=1) ** 2
        bn = torch.linalg.vector_norm(B, dim=1) ** 2
        red = torch.stack([num, den, rn, bn])
        dist.all_reduce(red)
        tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        err = float(torch.sqrt(red[0] / red[1]).max())
        true_res = float(torch.sqrt(red[2] / red[3]).max())
        tot = sum(d["ms"] for d in tags) or 1.0
        share = {d["kernel"]: d["ms"] / tot for d in tags}
        out = {"workload": f"ONE {int(nodes[0])}x{int(nodes[1])}x{int(nodes[2])} grid of the step's workload in {world} slabs along the last "
                           f"dimension (one per rank), {k} sources, same solver settings",
               "rhs_per_s": k * nsteps / (float(tmax.item()) / 1e3), "ms_per_batch": float(tmax.item()) / nsteps, "rhs_per_batch": k,
               "halo_share": share.get("halo_exchange", 0.0), "allreduce_share": share.get("allreduce", 0.0),
               "rel_err_vs_whole_grid": err, "true_relres_max": true_res,
               "iterations_slab": it_slab.tolist(), "iterations_whole_grid": it_whole.tolist(),
               "setup_seconds": t_setup, "planes_rank0": [int(k0), int(k1)], "clocks": sampler.result(),
               "transport": "ncclSend/ncclRecv halo planes + ncclAllReduce of the per-RHS dot partials (hh_create_slab_nccl)"}
        pkg.clear(A.MG)
        del X, Xw, B
        return out
    except Exception as e:  # never take the headline down; the failure is reported in the line
        return {"error": str(e)[:500]}


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
