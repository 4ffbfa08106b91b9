"""Experiment: the images of a batch are independent (GroupNorm and attention are per sample, diffusion.py:169-174 has no cross-image
term), so a batch can be cut into S sub-batches driven from S CUDA streams: the launch gaps, pipeline fill / drain and exposed epilogues
of one sub-batch's dependent launch chain are then filled by the other sub-batches' kernels.

    python tools/gpu_streams_check.py [total_batch ...]

Times K reverse steps for S in {1, 2, 4} (S engines of total_batch / S images each, one stream per engine) and checks that the sampler
state equals the S = 1 run bit for bit (Philox streams are keyed by global sample index).
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sr3_b200

SCHED = {"schedule": "linear", "n_timestep": 2000, "linear_start": 1e-6, "linear_end": 1e-2}
FULL = dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2, dropout=0.2)
dev = torch.device("cuda", 0)


def build_net():
    opt = {"phase": "val", "gpu_ids": [0], "distributed": False,
           "model": {"which_model_G": "sr3", "finetune_norm": False, "unet": dict(FULL), "beta_schedule": {"train": SCHED, "val": SCHED},
                     "diffusion": {"image_size": 128, "channels": 3, "conditional": True}}}
    torch.manual_seed(0)
    net = sr3_b200.define_G(opt).to(dev)
    net.set_new_noise_schedule(SCHED, dev)
    return net


def run(B, S, K=20, W=5):
    g = torch.Generator().manual_seed(1)
    cond = (torch.rand(B, 3, 128, 128, generator=g) * 2 - 1).to(dev)
    xT = torch.randn(B, 3, 128, 128, generator=g).to(dev)
    nets = [build_net() for _ in range(S)]
    b = B // S
    engs = [n.denoise_fn.engine(b, conditional=True, channels=3) for n in nets]
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    torch.cuda.synchronize()
    for i, (e, s) in enumerate(zip(engs, streams)):
        with torch.cuda.stream(s):
            e.loop_begin(cond[i * b:(i + 1) * b].contiguous(), xT[i * b:(i + 1) * b].contiguous(), seed=7, first_index=i * b)
            e.steps(1999, W)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    # interleave the launches so that no stream runs ahead on the host side
    for k in range(K):
        for e, s in zip(engs, streams):
            with torch.cuda.stream(s):
                e.steps(1999 - W - k, 1)
    for s in streams:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    state = torch.cat([e.read_state() for e in engs], 0).cpu()
    del engs, nets
    torch.cuda.empty_cache()
    return ms, state


if __name__ == "__main__":
    batches = [int(a) for a in sys.argv[1:]] or [16, 8, 4, 2]
    out = {}
    for B in batches:
        ref = None
        for S in (1, 2, 4, 8):
            if B % S or B // S < 1:
                continue
            ms, st = run(B, S)
            if ref is None:
                ref = st
            rec = {"ms_per_step": round(ms, 4), "steps_per_s": round(1000.0 / ms, 1), "bit_equal_to_S1": bool(torch.equal(st, ref)),
                   "rel_diff": float((st - ref).norm() / ref.norm())}
            out["B%d_S%d" % (B, S)] = rec
            print("B=%d S=%d" % (B, S), json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(os.path.join("gpurun_out", "streams_check_%s.json" % os.environ.get("TAG", "x")), "w"), indent=1)
