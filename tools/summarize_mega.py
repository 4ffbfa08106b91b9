"""gpurun_out/mega_check.json (tools/gpu_mega_check.py) -> profiles/r02_step_kernel.md: the persistent step kernel against the CUDA graph of
per-layer launches, same plan, same bits.

    python tools/summarize_mega.py gpurun_out/mega_check.json profiles/r02_step_kernel.md
"""
import json, sys

NAMES = {0: "gemm_tile", 1: "groupnorm_apply", 2: "attention", 3: "softmax", 4: "embed_film", 5: "stats_clear"}


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/mega_check.json"
    dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/r02_step_kernel.md"
    d = json.load(open(src))
    out = ["# Persistent step kernel (SR3_MEGA=1) vs CUDA graph of per-layer launches (default) -- measured on one B200", "",
           "`python tools/gpu_mega_check.py` : both paths built from the same plan, 20 resident reverse steps timed with CUDA events after 5 warm-up",
           "steps; `eps` of one UNet forward and the sampler state after the 25 steps compared bit for bit.  Per-op times of the step kernel are",
           "device `globaltimer` stamps taken by CTA 0 (previous op done / barrier arrived + op set up / barrier passed / body done).", "",
           "| config | batch | step kernel ms/step | graph path ms/step | launches/step (graph) | bit-identical (eps, state) | repeat runs bit-identical |",
           "|---|---:|---:|---:|---:|---|---|"]
    for key, r in d.items():
        cfg, b = key.rsplit("_B", 1)
        out.append(f"| {cfg} | {b} | {r['mega']['ms_per_step']:.3f} | {r['layers']['ms_per_step']:.3f} | {r['layers']['launches']} | "
                   f"{r['eps_bit_equal']}, {r['state_bit_equal']} | {r['mega']['repeat_bit_equal'] and r['layers']['repeat_bit_equal']} |")
    out += ["", "## Where the step kernel's time goes (sum over the ops of one launch, us)", ""]
    for key, r in d.items():
        m = r["mega"]
        out.append(f"### {key}: {m['ms_per_step']:.3f} ms/step, ops sum {m['ops_total_us']} us")
        out += ["", "| op class | ops | total us | us per op | set-up | barrier wait | body | end fence |", "|---|---:|---:|---:|---:|---:|---:|---:|"]
        ph = m.get("phase_us(setup,wait,body,fence)", {})
        for name, (n, us) in m["ops"].items():
            p4 = ph.get(name, [float("nan")] * 4)
            out.append(f"| {name} | {n} | {us} | {us / n:.1f} | {p4[0]} | {p4[1]} | {p4[2]} | {p4[3]} |")
        out.append("")
        if "per_op_us" in m:
            out += ["Per op in execution order (class:us): " + " ".join(f"{NAMES[t][0]}{NAMES[t][1] if t == 1 else ''}:{us}" for t, us in m["per_op_us"]), ""]
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
