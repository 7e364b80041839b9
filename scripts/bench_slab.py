#!/usr/bin/env python
"""Config 5 (BASELINE.json configs[4]): ONE 3-D 513^3 high-frequency problem split into slabs
along the last dimension over the GPUs of one box, halos and dot products over NCCL (hh_create_slab_nccl).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/bench_slab.py [--grid 513] [--nrhs 8] [--steps 2] [--warmup 1] [--prec c128|mixed|c64]

Every rank generates the model, keeps its planes, builds its slab of the hierarchy and solves all right-hand sides in
lockstep with the others.  Timing: CUDA events around the solves, max over ranks; rank 0 prints one JSON line
(RHS/s of the whole job = strong scaling: the problem is fixed, the GPUs split it)."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
from bench import ClockSampler  # noqa: E402  (NVML clock / throttle-reason sampling during the timed region)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", dest="n", type=int, default=513, help="nodes per dimension")
    ap.add_argument("--nrhs", type=int, default=8)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--prec", default="c128", choices=["c128", "mixed", "c64"])
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--tol", type=float, default=1e-6)
    ap.add_argument("--max-cycles", type=int, default=40)
    ap.add_argument("--ppw", type=float, default=10.0, help="points per wavelength at the slowest velocity")
    ap.add_argument("--variants", default="default", help="comma-separated runs in one process; a variant may set A/B "
                    "switches of the library as ENV=value joined by '+', e.g. HH_HALO_OVERLAP=1,HH_HALO_OVERLAP=0")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    pkg = graft.load_package()
    lib = pkg._lib.load()
    n = a.n
    geo = pkg.slabPartition(n, a.levels, world, rank)[0]
    mk0, mk1 = geo["koff"], geo["koff"] + geo["nloc"]  # planes this slab needs of the model (halo planes included)
    t0 = time.perf_counter()
    scale = (n - 1) / 512.0
    cfg = pkg.workloads.config5(n=n, sigma=16.0 * scale, seed=1234, pad=max(4, int(24 * scale)), planes=(mk0, mk1))
    mesh = pkg.getRegularMesh(cfg["domain"], cfg["n_cells"])
    nodes = mesh.n + 1
    w = (10.0 / a.ppw) * 0.1 * 2 * np.pi / (float(mesh.h.max()) * np.sqrt(cfg["max_m"]))  # getMaximalFrequency at `ppw`
    gamma = cfg["gamma0_frac"] * w + pkg.workloads.abl3d_planes(nodes, True, cfg["pad"], w, mk0, mk1)
    t_model = time.perf_counter() - t0
    prec = np.complex64 if a.prec == "c64" else np.complex128
    tdt = torch.complex64 if a.prec == "c64" else torch.complex128
    MG = pkg.getMGparam(prec, pkg.Int64, a.levels, 1, a.max_cycles, a.tol, "Jac", 0.8, 1, 2, "W", "GMRES", coarseIters=10)
    if a.prec == "mixed":
        MG.cyclePrecision = pkg.ComplexF32
    hp = pkg.HelmholtzParam(mesh, np.asfortranarray(gamma).ravel(order="F"), cfg["m"].ravel(order="F"), w, True, True)
    variants = [v for v in a.variants.split(",") if v]
    for variant in variants:
        # A/B switches of the library are read when a handle is created: one handle per variant, same model
        for kv in variant.split("+"):
            if "=" in kv:
                k, v = kv.split("=")
                if k == "nrhs":
                    a.nrhs = int(v)
                elif k == "prec":  # c128 | mixed
                    a.prec = v
                    MG.cyclePrecision = pkg.ComplexF32 if v == "mixed" else None
                else:
                    os.environ[k] = v
        run_variant(a, pkg, lib, torch, dist, rank, world, local, mesh, nodes, w, hp, MG, (mk0, mk1), t_model, variant, n)
        pkg.clear(MG)
    dist.barrier()
    dist.destroy_process_group()


def run_variant(a, pkg, lib, torch, dist, rank, world, local, mesh, nodes, w, hp, MG, model_planes, t_model, variant, n):
    tdt = torch.complex64 if a.prec == "c64" else torch.complex128
    A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "GMRES", 5)
    A.devices = [local]
    A.slabs = dict(pkg.sharding.nccl_slabs(), model_planes=model_planes)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    hd = pkg.api._ensure_hierarchy(A, 0)
    torch.cuda.synchronize()
    dist.barrier()
    t_setup = time.perf_counter() - t0
    k0, k1 = hd.planes
    plane = int(nodes[0] * nodes[1])
    Nown = plane * (k1 - k0)
    amp = 1.0 / mesh.h[0] ** 2
    g = int(np.ceil(np.sqrt(a.nrhs)))
    srcs = pkg.workloads.point_sources_top_grid(nodes, g, g)[:a.nrhs]
    gidx = np.array([pkg.loc2cs(nodes, s) - 1 for s in srcs], dtype=np.int64)
    B = torch.zeros((a.nrhs, Nown), dtype=tdt, device="cuda")
    for c, gi in enumerate(gidx):
        if plane * k0 <= gi < plane * k1:
            B[c, gi - plane * k0] = amp
    X = torch.empty_like(B)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        pkg.solveLinearSystem_(None, B, X, A)
    lib.hh_profile_enable(hd.h, 1)
    lib.hh_profile_reset(hd.h)
    l0 = C.c_int64()
    lib.hh_get_counters(hd.h, None, None, None, C.byref(l0))
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):
        pkg.solveLinearSystem_(None, B, X, A)
    ev1.record()
    barrier()
    sampler.stop_flag.set()
    ms = ev0.elapsed_time(ev1)
    l1 = C.c_int64()
    lib.hh_get_counters(hd.h, None, None, None, C.byref(l1))
    tags = []
    for t in range(lib.hh_profile_num_tags()):
        cnt, tms, by = C.c_int64(), C.c_double(), C.c_double()
        lib.hh_profile_get(hd.h, t, C.byref(cnt), C.byref(tms), C.byref(by))
        if cnt.value:
            tags.append(dict(kernel=lib.hh_profile_tag_name(t).decode(), launches=int(cnt.value), ms=tms.value, bytes=by.value))
    lib.hh_profile_enable(hd.h, 0)
    # true residual of the un-shifted operator, assembled over the slabs
    R = pkg.HelmholtzOperator(hd).matvec(X) - B
    num = torch.linalg.vector_norm(R, dim=1) ** 2
    den = torch.linalg.vector_norm(B, dim=1) ** 2
    dist.all_reduce(num)
    dist.all_reduce(den)
    true_res = float(torch.sqrt(num / den).max())
    del R
    # e2e: host blocks of this rank's planes through hh_solve (copies inside the timed region)
    Bh = B.cpu().pin_memory()
    Xh = torch.empty_like(Bh).pin_memory()
    barrier()
    t0 = time.perf_counter()
    pkg.solveLinearSystem_(None, Bh.numpy().T, Xh.numpy().T, A)
    torch.cuda.synchronize()
    dist.barrier()
    t_e2e = time.perf_counter() - t0
    tmax = torch.tensor([ms, t_e2e], dtype=torch.float64, device="cuda")
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    free, total = torch.cuda.mem_get_info()
    if rank == 0:
        ms_max = float(tmax[0])
        tot = sum(d["ms"] for d in tags)
        tags.sort(key=lambda d: -d["ms"])
        es = 8 if a.prec == "c64" else 16
        line = {
            "metric": f"rhs_solves_per_sec_to_{a.tol:g}_3d_{n}cubed_slab", "value": a.nrhs * a.steps / (ms_max / 1e3), "unit": "RHS/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "c128 (ComplexF32 multigrid cycle)" if a.prec == "mixed" else a.prec, "data": "synthetic",
            "config": {"workload": f"config5: ONE 3-D {n}^3-node problem, random-smooth velocity 1.5-4.5 km/s (seed 1234), h = "
                                   f"{mesh.h[0]:.4g} km, {a.ppw:g} points per wavelength (omega = {w:.4g}), ABL+Sommerfeld, {a.nrhs} point sources, shift 0.2, "
                                   f"{a.levels}-level W(1,2) Jacobi(0.8) Galerkin MG, coarsest Jacobi-GMRES(10), FGMRES(5), tol {a.tol:g}",
                       "parallelism": f"slab decomposition x{world} along the last dimension: one NCCL halo exchange per stencil-type "
                                      f"kernel, all-reduced dot products; rank 0 owns planes [{k0},{k1})",
                       "iterations": A.iterations.tolist(), "true_relres_max": true_res, "setup_seconds": t_setup,
                       "model_seconds": t_model, "device_memory_used_gb_rank0": (total - free) / 1e9},
            "e2e": {"value": a.nrhs / float(tmax[1]), "unit": "RHS/s", "h2d_bytes_per_step": int(Nown * a.nrhs * es),
                    "d2h_bytes_per_step": int(Nown * a.nrhs * es), "note": "bytes of rank 0; every rank copies its own planes"},
            "gpu_launches": int(l1.value - l0.value),
            "clocks": sampler.result(),
            "per_kernel_rank0": {d["kernel"]: {"launches": d["launches"], "share": round(d["ms"] / tot, 4),
                                               "avg_ms": round(d["ms"] / d["launches"], 4),
                                               "gbs": round(d["bytes"] / d["ms"] / 1e6, 1) if d["bytes"] else None} for d in tags},
        }
        line["config"]["variant"] = variant
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
