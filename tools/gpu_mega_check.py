"""First-contact check of the persistent step kernel: same engine config built twice (step kernel vs per-layer graph path),
outputs compared bit for bit, steps timed, per-op device times of the step kernel printed.

    python tools/gpu_mega_check.py [tiny|full|all] [batch ...]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sr3_b200

SCHED = {"schedule": "linear", "n_timestep": 2000, "linear_start": 1e-6, "linear_end": 1e-2}
TINY = (dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2], attn_res=[16], res_blocks=1, dropout=0.0), 32)
FULL = (dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2, dropout=0.2), 128)
BIG = (dict(in_channel=6, out_channel=3, inner_channel=64, norm_groups=16, channel_multiplier=[1, 2, 4, 8, 16], attn_res=[], res_blocks=1, dropout=0), 512)
dev = torch.device("cuda", 0)


def build(unet, size, B, mega):
    if mega:
        os.environ["SR3_MEGA"] = "1"
    else:
        os.environ.pop("SR3_MEGA", None)
    opt = {"phase": "val", "gpu_ids": [0], "distributed": False,
           "model": {"which_model_G": "sr3", "finetune_norm": False, "unet": dict(unet), "beta_schedule": {"train": SCHED, "val": SCHED},
                     "diffusion": {"image_size": size, "channels": 3, "conditional": True}}}
    torch.manual_seed(0)
    net = sr3_b200.define_G(opt).to(dev)
    net.set_new_noise_schedule(SCHED, dev)
    eng = net.denoise_fn.engine(B, conditional=True, channels=3)
    os.environ.pop("SR3_MEGA", None)
    return net, eng


def run(name, unet, size, B, K=20):
    g = torch.Generator().manual_seed(1)
    cond = (torch.rand(B, 3, size, size, generator=g) * 2 - 1).to(dev)
    xT = torch.randn(B, 3, size, size, generator=g).to(dev)
    res = {}
    outs = {}
    for mega in (True, False):
        net, eng = build(unet, size, B, mega)
        assert eng.uses_step_kernel() == mega, (eng.uses_step_kernel(), mega)
        x = torch.cat([cond, xT], 1)
        nl = torch.full((B, 1), 0.7, device=dev)
        eps = eng.unet_forward(x, nl)
        eps2 = eng.unet_forward(x, nl)
        torch.cuda.synchronize()
        eng.loop_begin(cond, xT, seed=7)
        eng.steps(1999, 5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.steps(1994, K); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        st = eng.read_state()
        outs[mega] = (eps.cpu(), st.cpu())
        res["mega" if mega else "layers"] = {"ms_per_step": round(ms, 4), "launches": eng.launches_per_step(), "repeat_bit_equal": bool(torch.equal(eps, eps2)),
                                             "finite": bool(torch.isfinite(st).all())}
        if mega:
            prof = eng.step_kernel_profile()
            by = {}
            for t, us in prof:
                d = by.setdefault(eng.STEP_OP_NAMES[t], [0, 0.0]); d[0] += 1; d[1] += us
            res["mega"]["ops"] = {k: [v[0], round(v[1], 1)] for k, v in by.items()}
            res["mega"]["ops_total_us"] = round(sum(us for _, us in prof), 1)
            res["mega"]["per_op_us"] = [[t, round(us, 1)] for t, us in prof]
            ph = eng.step_kernel_profile(phases=True)
            agg = {}
            for t, us, p4 in ph:
                d = agg.setdefault(eng.STEP_OP_NAMES[t], [0.0, 0.0, 0.0, 0.0])
                for k in range(4):
                    d[k] += p4[k]
            res["mega"]["phase_us(setup,wait,body,fence)"] = {k: [round(x, 1) for x in v] for k, v in agg.items()}
        del eng, net
        torch.cuda.empty_cache()
    res["eps_bit_equal"] = bool(torch.equal(outs[True][0], outs[False][0]))
    res["eps_rel_diff"] = float((outs[True][0] - outs[False][0]).norm() / outs[False][0].norm())
    res["state_bit_equal"] = bool(torch.equal(outs[True][1], outs[False][1]))
    res["state_rel_diff"] = float((outs[True][1] - outs[False][1]).norm() / outs[False][1].norm())
    po = res["mega"].pop("per_op_us")
    print(name, "B=%d" % B, json.dumps(res), flush=True)
    print("   per-op us:", po, flush=True)
    res["mega"]["per_op_us"] = po
    return res


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    batches = [int(a) for a in sys.argv[2:]] or None
    out = {}
    if which in ("tiny", "all"):
        for B in (batches or [2, 3]):
            out["tiny_B%d" % B] = run("tiny", *TINY, B)
    if which in ("full", "all"):
        for B in (batches or [16, 2]):
            out["full_B%d" % B] = run("full", *FULL, B)
    if which in ("big",):
        for B in (batches or [4]):
            out["big_B%d" % B] = run("big", *BIG, B, K=10)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(os.path.join("gpurun_out", "mega_check.json"), "w"), indent=1)
