"""`ncu -i x.ncu-rep --page raw --csv` export -> a tracked markdown summary: one block per distinct kernel / grid with the
metrics the roofline discussion uses and the top warp-stall reasons (pc sampling).

  python scripts/ncu_raw_summary.py gpurun_out/r02b/ncu_full.csv profiles/r02b_ncu_full_summary.md "title line"
"""
import csv
import re
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__pcsamp_sample_count"]


def main(src, dst, title):
    rows = list(csv.reader(open(src)))
    k = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[k], rows[k + 1], rows[k + 2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    seen, out = set(), []
    for d in data:
        if len(d) < len(hdr):
            continue
        name = d[idx["Kernel Name"]]
        m = re.match(r"(?:void )?(?:hh::)?([A-Za-z0-9_]+)(<[^>]*>)?", name)
        fn = m.group(1) + (m.group(2) or "")
        key = fn + " grid " + d[idx["launch__grid_size"]]
        if key in seen:
            continue
        seen.add(key)
        stalls = []
        for c in stall_cols:
            try:
                stalls.append((float(d[idx[c]].replace(",", "")), c.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
        stalls.sort(reverse=True)
        out.append(f"## `{key}`\n\n| metric | value | unit |\n|---|---:|---|\n" +
                   "".join(f"| {w} | {d[idx[w]]} | {units[idx[w]]} |\n" for w in WANT if w in idx) +
                   "| top stall reasons (pc samples) | " + ", ".join(f"{n} {int(v)}" for v, n in stalls[:6]) + " | |\n")
    open(dst, "w").write(f"# {title}\n\n" + "\n".join(out))
    print(f"{len(out)} kernels -> {dst}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "ncu --set full --clock-control none")
