"""Summarise ncu outputs brought back in gpurun_out/ into tracked files under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
  python scripts/summarize_ncu.py full gpurun_out/prof_r01_stencils.ncu-rep profiles/r01_stencils_ncu.md
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import defaultdict


def short(name):
    m = re.match(r"(?:void )?(?:hh::)?([A-Za-z0-9_]+)(<.*>)?", name)
    base = m.group(1) if m else name
    targs = ""
    if m and m.group(2):
        targs = m.group(2)
        targs = re.sub(r"hh::", "", targs)
        targs = targs[:60]
    return base, targs


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rows[1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        base, targs = short(r[idx["Kernel Name"]])
        per[base + targs][0] += 1
        per[base + targs][1] += us
        tot += us
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over one bench.py step "
                "(cold-cache, serialised: compare SHARES, not absolutes).\n\n")
        f.write(f"total kernel time {tot / 1e3:.1f} ms over {sum(v[0] for v in per.values())} launches\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {us / 1e3:.2f} | {100 * us / tot:.1f}% | {us / n:.1f} |\n")
    print(open(dst).read())


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_sample_count",
]


def full(src, dst, traffic_json=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = {}
    for d in data:
        base, targs = short(d[idx["Kernel Name"]])
        key = base + targs
        seen.setdefault(key, d)  # first launch of each distinct kernel
    traffic = {}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n`ncu --set full --clock-control none --import-source on`; one launch per kernel.\n\n")
        for key, d in seen.items():
            f.write(f"## `{key}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in WANT:
                if w in idx:
                    f.write(f"| {w} | {d[idx[w]]} | {units[idx[w]]} |\n")
            try:
                def val(name):
                    v = float(d[idx[name]].replace(",", ""))
                    u = units[idx[name]]
                    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
                tb = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
                t = float(d[idx["gpu__time_duration.sum"]].replace(",", ""))
                tu = units[idx["gpu__time_duration.sum"]]
                tsec = t * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "s": 1, "second": 1}.get(tu, 1e-9)
                f.write(f"| **DRAM traffic (read+write)** | {tb / 1e9:.3f} | GB |\n| **DRAM GB/s under ncu** | {tb / tsec / 1e9:.0f} | GB/s |\n")
                traffic[key] = tb
            except Exception as e:  # noqa
                pass
            f.write("\n")
    if traffic_json:
        json.dump(traffic, open(traffic_json, "w"), indent=1)
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
