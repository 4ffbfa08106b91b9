"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_r01b.csv profiles/r01b_launches.md "note"
    python tools/summarize_profiles.py full gpurun_out/prof_gemm_r01b.ncu-rep profiles/r01b_gemm_full.md "note"
"""
import collections
import csv
import re
import subprocess
import sys


def us(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)


def launches(src, dst, note):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us(r)
        tot += us(r)
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({len(rows)} launches = one reverse step), `gpu__time_duration.sum`, --clock-control none\n\n{note}\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {v:.1f} | {100 * v / tot:.1f}% |\n")
        f.write(f"| **total** | {len(rows)} | {tot:.1f} | 100% |\n\n")
        f.write("Per-launch durations (us) in launch order:\n\n```\n")
        for i, r in enumerate(rows):
            kn = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")[:40]
            f.write(f"{i:3d} {us(r):8.1f} {kn:40s} grid={r['Grid Size']}\n")
        f.write("```\n")


METRICS = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
           "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "sm__cycles_elapsed.max"]


def full(src, dst, note):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full, --clock-control none ({len(data)} launches)\n\n{note}\n\n")
        cols = [m for m in METRICS if m in idx]
        f.write("| " + " | ".join(c.split(".")[-2] if c.count(".") > 1 else c for c in cols) + " |\n|" + "---|" * len(cols) + "\n")
        for d in data:
            f.write("| " + " | ".join((d[idx[c]] + " " + units[idx[c]]).strip()[:48] for c in cols) + " |\n")
        f.write("\nColumns: " + ", ".join(f"`{c}`" for c in cols) + "\n")


def metrics(src, dst, note):
    """Per-launch metric CSV of the tile kernel (one eager step) -> markdown table + profiles/<prefix>_traffic.json."""
    import json
    import os
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = per.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""), "grid": r["Grid Size"]})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            d[r["Metric Name"]] = None
    rows = list(per.values())
    rd = sum(r.get("dram__bytes_read.sum") or 0 for r in rows)
    wr = sum(r.get("dram__bytes_write.sum") or 0 for r in rows)
    tot_us = sum((r.get("gpu__time_duration.sum") or 0) / 1000 for r in rows)
    with open(dst, "w") as f:
        f.write(f"# {note}\n\n{len(rows)} launches, sum of durations {tot_us:.0f} us (cold caches per launch); DRAM read {rd / 1e9:.2f} GB + write {wr / 1e9:.2f} GB per step.\n\n")
        f.write("| # | kernel | grid | us | DRAM rd MB | DRAM wr MB | xbar->SM rd MB | L2 % | DRAM % |\n|---|---|---|---:|---:|---:|---:|---:|---:|\n")
        for i, r in enumerate(rows):
            g = lambda k, sc=1.0: ("%.1f" % ((r.get(k) or 0) / sc))
            f.write(f"| {i} | `{r['name']}` | {r['grid']} | {g('gpu__time_duration.sum', 1e3)} | {g('dram__bytes_read.sum', 1e6)} | {g('dram__bytes_write.sum', 1e6)} | "
                    f"{g('l1tex__m_xbar2l1tex_read_bytes.sum', 1e6)} | {g('lts__throughput.avg.pct_of_peak_sustained_elapsed')} | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} |\n")
    tj = os.path.join(os.path.dirname(dst), "r01_traffic.json")
    json.dump({"gemm_tile_kernel_dram_bytes_per_step": rd + wr, "launches": len(rows),
               "note": "sum of dram__bytes_read.sum + dram__bytes_write.sum over the tile-kernel launches of one eager reverse step, ncu --clock-control none "
                       "(caches flushed before every launch), B=16 16->128; per-launch table in " + dst}, open(tj, "w"), indent=1)
    print(f"{len(rows)} launches, {tot_us:.0f} us, DRAM {(rd + wr) / 1e9:.2f} GB -> {dst}, {tj}")


if __name__ == "__main__":
    {"launches": launches, "full": full, "metrics": metrics}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
