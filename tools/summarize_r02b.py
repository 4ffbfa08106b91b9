"""Round-2 profile summaries for profiles/ from the ncu CSV / reports of tools/gpu_profile_r02.sh.

    python tools/summarize_r02b.py gpurun_out/r2_p
"""
import collections, csv, json, re, subprocess, sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    L = collections.OrderedDict()
    for r in rows:
        d = L.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("sr3::", ""), "grid": r.get("Grid Size", "")})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            v = float("nan")
        u = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d[r["Metric Name"]] = v
    return list(L.values())


T, DR, DW, PIPE, INST = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum"


def table(L, with_pipe):
    agg = collections.OrderedDict()
    for d in L:
        a = agg.setdefault(d["name"], collections.Counter())
        t = d.get(T, 0.0)
        a["n"] += 1; a["us"] += t; a["dram"] += d.get(DR, 0.0) + d.get(DW, 0.0); a["pipe_w"] += d.get(PIPE, 0.0) * t; a["inst"] += d.get(INST, 0.0)
    tot = sum(a["us"] for a in agg.values())
    out = ["| kernel | launches | total us | share | DRAM GB (read + write) | GB/s |" + (" tensor pipe active (time-weighted) | tensor-pipe instructions |" if with_pipe else ""),
           "|---|---:|---:|---:|---:|---:|" + ("---:|---:|" if with_pipe else "")]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        row = "| `%s` | %d | %.1f | %.1f%% | %.3f | %.0f |" % (k, a["n"], a["us"], 100 * a["us"] / tot, a["dram"] / 1e9, a["dram"] / max(a["us"], 1e-9) / 1e3)
        if with_pipe:
            row += " %.1f%% | %d |" % (a["pipe_w"] / max(a["us"], 1e-9), a["inst"])
        out.append(row)
    out.append("| **total** | %d | %.1f | 100%% | %.3f | |" % (len(L), tot, sum(a["dram"] for a in agg.values()) / 1e9) + (" | |" if with_pipe else ""))
    return "\n".join(out), agg, tot


def full_report(rep, keys):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    out = ["| metric | value | unit |", "|---|---:|---|"]
    for k in hdr:
        if any(re.search(p, k) for p in keys) and d[k] not in ("", "n/a"):
            out.append("| `%s` | %s | %s |" % (k, d[k], u[k]))
    return "\n".join(out), d


def main():
    pre = sys.argv[1]
    L = load(pre + "_launches.csv")
    tab, agg, tot = table(L, False)
    with open("profiles/r02_launches.md", "w") as f:
        f.write("# ncu launch list of ONE reverse step (%d launches), round 2 build, `gpu__time_duration.sum`, --clock-control none\n\n" % len(L))
        f.write("`tools/profile_one_step.py 4 16` (graph replay of the benchmark step, B=16, 16->128) under `ncu --profile-from-start off`.  Per-launch times under ncu are "
                "cold-cache and serialised: compare SHARES with bench.py's event timings.\n\n" + tab + "\n\nPer-launch durations (us) in launch order:\n\n```\n")
        for i, d in enumerate(L):
            f.write("%3d %8.1f %-40s grid=%s\n" % (i, d.get(T, 0.0), d["name"], d["grid"]))
        f.write("```\n")
    M = load(pre + "_step_metrics.csv")
    tab, agg, tot = table(M, True)
    step_dram = sum(a["dram"] for a in agg.values())
    tc = sum(a["dram"] for k, a in agg.items() if k.startswith("gemm_tile_kernel") or k.startswith("attn_kernel"))
    prep = sum(a["dram"] for k, a in agg.items() if k.startswith("prep_kernel"))
    json.dump({"whole_step_dram_bytes": step_dram, "gemm_tile_kernel_dram_bytes_per_step": tc, "prep_kernel_dram_bytes_per_step": prep, "algorithmic_bytes_per_step": 2.98e9,
               "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the 146 launches of one step (tools/gpu_profile_r02.sh), round 2 build"},
              open("profiles/r02_traffic.json", "w"), indent=1)
    with open("profiles/r02_step_metrics.md", "w") as f:
        f.write("# One reverse step (B=16, 16->128): time, DRAM traffic and tensor-pipe activity of every launch, round 2 build\n\n"
                "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,"
                "sm__inst_executed_pipe_tensor.sum ...` (--clock-control none).  `sm__pipe_tensor_cycles_active` DOES count tcgen05 work on sm_100 "
                "(`ncu --query-metrics` lists `sm__inst_executed_pipe_tensor_subpipe_hmma` as HMMA/UTCHMMA/UTCQMMA/UTCOMMA); round 1 used a metric that did not.\n\n"
                + tab + "\n\nWhole-step DRAM traffic %.2f GB vs 2.98 GB algorithmic (x%.2f); the GroupNorm-apply pass (`prep_kernel`) moves %.2f GB of it.\n"
                % (step_dram / 1e9, step_dram / 2.98e9, prep / 1e9))
    R = load(pre + "_train_launches.csv")
    tab, agg, tot = table(R, False)
    with open("profiles/r02_train_launches.md", "w") as f:
        f.write("# ONE training iteration (forward + backward + Adam + weight re-pack), 16->128 config, 8 images: ncu launch list\n\n"
                "`tools/profile_one_train_step.py 8` under `ncu --profile-from-start off` (first training build of the round: before the coalesced packers / "
                "single-wave wgrad / merged GroupNorm parameter gradients -- see DESIGN.md section 8 for the event-timed numbers of the final build).\n\n" + tab + "\n")
    keys = [r"^gpu__time_duration.sum$", r"^launch__grid_size$", r"^launch__registers_per_thread$", r"^dram__bytes_(read|write).sum$", r"^lts__t_bytes.sum$",
            r"sm__pipe_tensor_cycles_active", r"pipe_tensor_subpipe_hmma_cycles_active_realtime", r"^sm__throughput.avg.pct", r"^sm__cycles_elapsed.max$",
            r"^smsp__cycles_active.avg$", r"^lts__throughput.avg.pct", r"sm__mem_tensor_cycles_active.avg.pct", r"l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_[ab].*\.sum$",
            r"^sm__inst_executed_pipe_tensor.sum$", r"^smsp__warp_issue_stalled_.*_per_warp_active.pct$"]
    for name, title in (("wgrad_full", "wgrad_kernel (tcgen05, MN-major operands), launch 40 of a training iteration at 8 images"),
                        ("hi64_full", "gemm_tile_kernel<64, 2>, conv3x3 64->64 @ 128x128, B=16 (the epilogue-bound tile shape)")):
        try:
            tab, d = full_report(pre + "_" + name + ".ncu-rep", keys)
            with open("profiles/r02_%s.md" % name, "w") as f:
                f.write("# `ncu --set full --import-source on`: %s\n\n%s\n" % (title, tab))
        except Exception as e:
            print("skip", name, e)


if __name__ == "__main__":
    main()
