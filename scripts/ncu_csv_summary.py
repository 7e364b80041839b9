"""Turn `ncu -i x.ncu-rep --page raw --csv` exports (gpurun_out/*.csv) into profiles/r01_ncu_full_summary.md and
profiles/ncu_traffic.json (DRAM bytes per launch of each kernel class, keyed like bench.py's roofline.kernel)."""
import csv, json, re, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_sample_count"]
# kernel function (+ mode template argument) -> bench.py kernel class
CLASS = [(r"k_fine3d_tma_first<double, 0", "fine_first_resid"), (r"k_fine3d_tma_first<double, 1", "fine_first_jacobi"),
         (r"k_fine3d_tma_pro<double", "fine_prolong_jacobi"), (r"k_fine3d_tma_prob<double", "fine_prolong_jacobi"), (r"k_fine3d_tma<double, 0", "fine_apply"),
         (r"k_fine3d_tma<double, 1", "fine_resid"), (r"k_fine3d_tma<double, 2", "fine_jacobi"),
         (r"k_coarse3d_tma<double, 0", "coarse_apply"), (r"k_coarse3d_tma<double, 1", "coarse_resid"),
         (r"k_coarse3d_tma<double, 2", "coarse_jacobi"), (r"k_restrict", "restrict"), (r"k_prolong_add", "prolong")]


def load(p):
    rows = list(csv.reader(open(p)))
    k = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[k], rows[k + 1], rows[k + 2:]


out, traffic = [], {}
for p in sys.argv[1:]:
    hdr, units, data = load(p)
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for d in data:
        name = d[idx["Kernel Name"]]
        m = re.match(r"(?:void )?(?:hh::)?([A-Za-z0-9_]+)(<[^>]*>)?", name)
        fn = m.group(1) + (m.group(2) or "")
        key = fn + " grid " + d[idx["launch__grid_size"]]
        if key in seen:
            continue
        seen.add(key)

        def val(n):
            v = float(d[idx[n]].replace(",", ""))
            u = units[idx[n]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9,
                        "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}.get(u, 1)

        tb = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        t = val("gpu__time_duration.sum")
        cls = next((c for pat, c in CLASS if fn.startswith(pat)), None)
        out.append(f"## `{key}`" + (f"  (bench kernel class `{cls}`)" if cls else "") + "\n\n| metric | value | unit |\n|---|---:|---|\n" +
                   "".join(f"| {w} | {d[idx[w]]} | {units[idx[w]]} |\n" for w in WANT if w in idx) +
                   f"| **DRAM traffic (read+write)** | {tb / 1e9:.3f} | GB |\n| **DRAM GB/s under ncu** | {tb / t / 1e9:.0f} | GB/s |\n")
        if cls and (cls not in traffic or tb > traffic[cls]):
            traffic[cls] = tb  # the largest-grid (level with most nodes) launch of the class
open("profiles/r01_ncu_full_summary.md", "w").write(
    "# ncu --set full --clock-control none: one launch per distinct kernel / grid\n\nconfig 4 (257^3, 16 RHS, W(1,2), ComplexF64); "
    "command `ncu --set full --clock-control none -k regex:<kernels> -c N python scripts/explore.py --n 257 --nrhs 16 --cycle W "
    "--pre 1 --post 2 --maxit 1`, raw page exported on the GPU box with `ncu -i ... --page raw --csv` (the .ncu-rep files exceed the "
    "64 MiB return limit).\n\n" + "\n".join(out))
json.dump(traffic, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(traffic, indent=1))
