"""ncu launch list of one bench.py step -> per-kernel-class time share and DRAM traffic.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline
  python scripts/ncu_class_traffic.py gpurun_out/launches_r02.csv profiles/r02_launches.md profiles/ncu_traffic.json

Classes are the ones bench.py reports (hh_profile tags), so that `roofline.traffic` (average DRAM bytes per launch of the
class, from this file) can be set against `roofline.algorithmic_bytes_per_launch` (same average, from the bench run).
ncu times are cold-cache and serialised: compare SHARES with the bench line, not absolutes."""
import csv
import json
import re
import sys
from collections import defaultdict

MODE3 = {"0": "apply", "1": "resid", "2": "jacobi"}


def classify(name):
    m = re.search(r"k_fine3d_tma_first<\w+, (\d)", name)
    if m:
        return "fine_first_resid" if m.group(1) == "0" else "fine_first_jacobi"
    if "k_fine3d_tma_pro2<" in name:
        return "fine_prolong_jacobi2"
    if "k_fine3d_tma_pro<" in name or "k_fine3d_tma_prob<" in name:
        return "fine_prolong_jacobi"
    m = re.search(r"k_fine3d_(?:tma|zmarch)<\w+, (\d)", name)
    if m:
        return "fine_" + MODE3[m.group(1)]
    m = re.search(r"k_fine_stencil<\w+, \d, (\d)", name)
    if m:
        return "fine_" + MODE3[m.group(1)]
    m = re.search(r"k_stencil2d_tma<\w+, (\d), \d+, (\w+)>", name)
    if m:
        return ("coarse_" if m.group(2) in ("true", "1") else "fine_") + MODE3[m.group(1)]
    m = re.search(r"k_coarse3d_(?:tma|zmarch)<\w+, (\d)", name)
    if m:
        return "coarse_" + MODE3[m.group(1)]
    m = re.search(r"k_coarse_stencil<\w+, \d, (\d)", name)
    if m:
        return "coarse_" + MODE3[m.group(1)]
    if "k_restrict" in name:
        return "restrict"
    if "k_prolong_add" in name:
        return "prolong"
    if "k_multidot" in name:
        return "krylov_dot"
    if "k_multiaxpy" in name or "k_bicg_p" in name or "k_combine" in name:
        return "krylov_axpy"
    if "k_diag_scale" in name:
        return "coarse_jacobi0"
    if "k_fine_jacobi0" in name:
        return "fine_jacobi0"
    if "k_dense_apply" in name:
        return "coarsest_dense"
    if re.search(r"k_gmres_|k_bicg_scalars|k_sum_partials|k_publish_state", name):
        return "scalar"
    if re.search(r"k_repitch|k_convert|k_point_sources", name):
        return "copy"
    if re.search(r"k_galerkin|k_coarse_dinv|k_fine_precompute|k_fine_dinv|k_scale_columns|k_band|k_inverse|k_gamma_abl|k_max_partial|k_fine_diag|k_ho_stencil", name):
        return "setup"
    return "other"


def main(src, dst_md, dst_json):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    launches = defaultdict(dict)  # id -> {name, metric: value}
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        lid = r[idx["ID"]]
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        met = r[idx["Metric Name"]]
        if met == "gpu__time_duration.sum":
            v = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v)
        else:
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
            v *= mult
        launches[lid]["name"] = r[idx["Kernel Name"]]
        launches[lid][met] = v
    per = defaultdict(lambda: dict(n=0, us=0.0, rd=0.0, wr=0.0, kernels=defaultdict(int)))
    tot = 0.0
    for l in launches.values():
        c = classify(l["name"])
        p = per[c]
        p["n"] += 1
        p["us"] += l.get("gpu__time_duration.sum", 0.0)
        p["rd"] += l.get("dram__bytes_read.sum", 0.0)
        p["wr"] += l.get("dram__bytes_write.sum", 0.0)
        p["kernels"][re.sub(r"\(.*", "", l["name"]).replace("void ", "").replace("hh::", "")[:70]] += 1
        tot += l.get("gpu__time_duration.sum", 0.0)
    solve_tot = sum(p["us"] for c, p in per.items() if c != "setup")
    out = {}
    with open(dst_md, "w") as f:
        f.write(f"# ncu launch list of one bench.py step, by kernel class ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` over "
                "`bench.py --steps 1 --warmup 0` (cold-cache, serialised: compare SHARES with the bench line, not absolutes; "
                "`setup` = hierarchy construction, outside the timed region of the bench).\n\n")
        f.write(f"total kernel time {tot / 1e3:.1f} ms over {len(launches)} launches; solve classes {solve_tot / 1e3:.1f} ms\n\n")
        f.write("| class | launches | total ms | share of solve | avg us | DRAM read GB | DRAM write GB | DRAM bytes / launch (MB) | DRAM GB/s under ncu | kernels |\n")
        f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
        for c, p in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
            share = p["us"] / solve_tot if c != "setup" else float("nan")
            bpl = (p["rd"] + p["wr"]) / p["n"]
            out[c] = {"launches": p["n"], "dram_bytes_per_launch": bpl, "share": None if c == "setup" else share,
                      "avg_us": p["us"] / p["n"]}
            ks = ", ".join(f"`{k}` x{n}" for k, n in sorted(p["kernels"].items(), key=lambda kv: -kv[1])[:4])
            f.write(f"| {c} | {p['n']} | {p['us'] / 1e3:.2f} | {100 * share:.1f}% | {p['us'] / p['n']:.1f} | {p['rd'] / 1e9:.2f} | "
                    f"{p['wr'] / 1e9:.2f} | {bpl / 1e6:.1f} | {(p['rd'] + p['wr']) / max(p['us'], 1e-9) / 1e3:.0f} | {ks} |\n")
    json.dump(out, open(dst_json, "w"), indent=1)
    print(open(dst_md).read())


if __name__ == "__main__":
    main(*sys.argv[1:4])
