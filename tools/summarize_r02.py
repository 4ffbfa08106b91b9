"""ncu CSV of one reverse step (tools/gpu_profile_run.sh) -> profiles/r02_step_metrics.md + profiles/r02_traffic.json.

    python tools/summarize_r02.py gpurun_out/r2_<tag>_step_metrics.csv [profiles/r02]
"""
import collections, csv, json, re, sys


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("sr3::", ""), "grid": r.get("Grid Size", "")})
        v, u = num(r["Metric Value"]), r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)      # -> us
        if u in ("Kbyte",): v *= 1e3
        if u in ("Mbyte",): v *= 1e6
        if u in ("Gbyte",): v *= 1e9
        d[r["Metric Name"]] = v
    return list(launches.values())


def main():
    src = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else "profiles/r02"
    L = load(src)
    T, DR, DW, OPS, PIPE = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"
    agg = collections.OrderedDict()
    for d in L:
        a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "dram": 0.0, "ops": 0.0, "pipe_w": 0.0})
        a["n"] += 1; a["us"] += d.get(T, 0.0); a["dram"] += d.get(DR, 0.0) + d.get(DW, 0.0); a["ops"] += d.get(OPS, 0.0)
        a["pipe_w"] += d.get(PIPE, 0.0) * d.get(T, 0.0)
    tot_us = sum(a["us"] for a in agg.values()); tot_dram = sum(a["dram"] for a in agg.values()); tot_ops = sum(a["ops"] for a in agg.values())
    with open(out + "_step_metrics.md", "w") as f:
        f.write("# One reverse step (16->128, batch 16) under ncu: per-launch time, DRAM bytes and tensor-pipe counters\n\n"
                "`ncu --metrics gpu__time_duration.sum,dram__bytes_{read,write}.sum,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum,"
                "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime... --clock-control none --profile-from-start off python tools/profile_one_step.py 4 16`.\n"
                "Times under ncu are serialised and cold-cache: compare SHARES with bench.py, not absolutes.  `tensor ops` is the hardware count of\n"
                "bf16 tensor-path math operations (UTCHMMA included; 2 ops per MAC), `tensor pipe %` the time-weighted HMMA-subpipe activity.\n\n")
        f.write("| kernel | launches | us | share | DRAM MB | tensor TFLOP | TFLOP/s (ncu time) | tensor pipe % |\n|---|---:|---:|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            f.write(f"| `{k}` | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot_us:.1f}% | {a['dram'] / 1e6:.1f} | {a['ops'] / 1e12:.4f} | "
                    f"{(a['ops'] / (a['us'] * 1e-6) / 1e12) if a['us'] else 0:.0f} | {(a['pipe_w'] / a['us']) if a['us'] else 0:.1f} |\n")
        f.write(f"| **step** | {len(L)} | {tot_us:.1f} | 100% | {tot_dram / 1e6:.1f} | {tot_ops / 1e12:.4f} | {tot_ops / (tot_us * 1e-6) / 1e12:.0f} | |\n\n")
        f.write("Per launch, in launch order (us, DRAM MB, tensor GFLOP, tensor pipe %):\n\n```\n")
        for i, d in enumerate(L):
            f.write(f"{i:3d} {d['name'][:34]:34s} {d.get(T, 0):8.1f} {(d.get(DR, 0) + d.get(DW, 0)) / 1e6:9.2f} {d.get(OPS, 0) / 1e9:9.2f} {d.get(PIPE, 0):6.1f}  grid={d['grid']}\n")
        f.write("```\n")
    tile = sum(a["dram"] for k, a in agg.items() if "gemm_tile_kernel" in k or "attn_kernel" in k)
    step_k = sum(a["dram"] for k, a in agg.items() if "step_kernel" in k)
    json.dump({"source": src, "step_dram_bytes": tot_dram, "gemm_tile_kernel_dram_bytes_per_step": tile, "step_kernel_dram_bytes_per_launch": step_k or None,
               "tensor_ops_per_step": tot_ops, "launches": len(L)}, open(out + "_traffic.json", "w"), indent=1)
    print("wrote", out + "_step_metrics.md", out + "_traffic.json", "launches", len(L), "DRAM GB", tot_dram / 1e9)


if __name__ == "__main__":
    main()
