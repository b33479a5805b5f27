"""Turns the two ncu outputs of the profiling recipe into the tracked summaries under profiles/:
  python scripts/ncu_summary.py <tag> <launches.csv> <raw.csv from `ncu -i X.ncu-rep --page raw --csv`> "<command>"
writes profiles/<tag>_launches.csv (copy), profiles/<tag>_kernels_full.csv (selected metrics) and profiles/<tag>_summary.md."""
import collections
import csv
import shutil
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]


def main():
    tag, launches, raw, cmd = sys.argv[1:5]
    shutil.copy(launches, f"profiles/{tag}_launches.csv")
    rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            d[r[ki]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    out = [f"# {tag}: ncu evidence for `{cmd}` on B200\n",
           "Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv ...` (after the same command exited 0 "
           f"without ncu); raw: `{tag}_launches.csv`. Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n",
           "| kernel | launches | avg us | share of captured GPU time |", "|---|---|---|---|"]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]))[:20]:
        out.append(f"| `{k[:72]}` | {len(v)} | {sum(v) / len(v) / 1000:.1f} | {100 * sum(v) / tot:.1f}% |")
    rr = list(csv.reader(open(raw)))
    h, u = rr[0], rr[1]
    cols = [h.index(k) for k in KEEP if k in h]
    with open(f"profiles/{tag}_kernels_full.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([h[i] for i in cols])
        w.writerow([u[i] for i in cols])
        for r in rr[2:]:
            w.writerow([r[i] for i in cols])
    out += ["", f"`ncu --set full --clock-control none --import-source on` of the hot kernels (selected metrics: `{tag}_kernels_full.csv`):\n",
            "| " + " | ".join(h[i].split(".")[0] for i in cols) + " |", "|" + "---|" * len(cols)]
    for r in rr[2:]:
        out.append("| " + " | ".join(f"{r[i][:48]} {u[i]}".strip() for i in cols) + " |")
    open(f"profiles/{tag}_summary.md", "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
