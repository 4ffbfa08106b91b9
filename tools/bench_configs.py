"""Steps/s of the other sampling configurations of BASELINE.json on one GPU (resident state, CUDA events)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sr3_b200

SCHED = {"schedule": "linear", "n_timestep": 2000, "linear_start": 1e-6, "linear_end": 1e-2}
CONFIGS = {
    "sr_sr3_16_128 B=16": (dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2, dropout=0.2), 128, True, 16),
    "sr_sr3_64_512 B=4": (dict(in_channel=6, out_channel=3, inner_channel=64, norm_groups=16, channel_multiplier=[1, 2, 4, 8, 16], attn_res=[], res_blocks=1, dropout=0), 512, True, 4),
    "sample_sr3_128 (uncond) B=32": (dict(in_channel=3, out_channel=3, inner_channel=64, channel_multiplier=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2, dropout=0.2), 128, False, 32),
    "sr_sr3_16_128 B=8 (one of 2 GPUs)": (dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2, dropout=0.2), 128, True, 8),
    "sr_sr3_16_128 B=4 (one of 4 GPUs)": (dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2, dropout=0.2), 128, True, 4),
    "sr_sr3_16_128 B=2 (one of 8 GPUs)": (dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2, dropout=0.2), 128, True, 2),
}
dev = torch.device("cuda", 0)
out = {}
for name, (unet, size, cond, B) in CONFIGS.items():
    opt = {"phase": "val", "gpu_ids": [0], "distributed": False,
           "model": {"which_model_G": "sr3", "finetune_norm": False, "unet": dict(unet), "beta_schedule": {"train": SCHED, "val": SCHED},
                     "diffusion": {"image_size": size, "channels": 3, "conditional": cond}}}
    torch.manual_seed(0)
    net = sr3_b200.define_G(opt).to(dev)
    net.set_new_noise_schedule(SCHED, dev)
    eng = net.denoise_fn.engine(B, conditional=cond, channels=3)
    c = (torch.rand(B, 3, size, size) * 2 - 1).to(dev) if cond else None
    eng.loop_begin(c, torch.randn(B, 3, size, size).to(dev), seed=1)
    eng.steps(1999, 5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 30
    e0.record(); eng.steps(1994, K); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    ok = bool(torch.isfinite(eng.read_state()).all())
    out[name] = {"ms_per_step": ms, "steps_per_s": 1000.0 / ms, "launches_per_step": eng.launches_per_step(), "finite": ok,
                 "workspace_GB": eng.workspace_bytes() / 2 ** 30}
    print(name, out[name], flush=True)
    del eng, net
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join("gpurun_out", "configs.json"), "w"), indent=1)
