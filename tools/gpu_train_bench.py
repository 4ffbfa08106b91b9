"""Training step (config 4) timing on one GPU: full 16->128 config, per-GPU batch 8 (the 8-GPU B=64 shard), forward / backward / Adam split.

    python tools/gpu_train_bench.py [batch] [steps]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sr3_b200
from sr3_b200 import parallel

SCHED = {"schedule": "linear", "n_timestep": 2000, "linear_start": 1e-6, "linear_end": 1e-2}
FULL = dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2, dropout=0.2)

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda", 0)
    opt = {"phase": "train", "gpu_ids": [0], "distributed": False,
           "model": {"which_model_G": "sr3", "finetune_norm": False, "unet": dict(FULL), "beta_schedule": {"train": SCHED, "val": SCHED},
                     "diffusion": {"image_size": 128, "channels": 3, "conditional": True}}}
    torch.manual_seed(0)
    net = sr3_b200.define_G(opt).to(dev)
    net.set_loss(dev); net.set_new_noise_schedule(SCHED, dev); net.train()
    g = torch.Generator().manual_seed(1)
    hr = (torch.rand(B, 3, 128, 128, generator=g) * 2 - 1).to(dev)
    sr = (torch.rand(B, 3, 128, 128, generator=g) * 2 - 1).to(dev)
    tr = parallel.DataParallelTrainer(net, lr=1e-4)
    t0 = time.time()
    losses = [tr.step(hr, sr) for _ in range(2)]
    torch.cuda.synchronize()
    print("warm-up 2 steps: %.2f s, losses %s, mem %.1f GB" % (time.time() - t0, losses, torch.cuda.max_memory_allocated() / 2**30), flush=True)
    eng = tr._eng
    print("engine workspace %.2f GB, backward blocks %d" % (eng.workspace_bytes() / 2**30, eng.num_backward_blocks()), flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(K):
        losses.append(tr.step(hr, sr))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    # split: forward only / backward only, device-timed
    gamma = torch.rand(B) * 0.5 + 0.3
    noise = torch.randn(B, 3, 128, 128, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record(); eng.train_forward(hr, sr, gamma, noise, "l1", 1, want_loss=False); ev[1].record()
    eng.backward_begin(1.0, tr.buckets.views)
    for i in range(eng.num_backward_blocks() - 1, -1, -1):
        eng.backward_block(i)
    eng.backward_finish(); ev[2].record()
    tr.opt.step(); ev[3].record()
    torch.cuda.synchronize()
    fwd, bwd, adam = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
    eng.train_forward(hr, sr, gamma, noise, "l1", 1, want_loss=False)
    prof = eng.train_backward_profile(1.0, tr.buckets.views)
    print("backward by op kind (ms):", {k: round(v, 2) for k, v in prof.items()}, flush=True)
    fprof = {}
    for kind, ms_, fl_, by_ in eng.profile_step(1000, reps=2):
        fprof[kind] = fprof.get(kind, 0.0) + ms_
    print("forward by op kind (ms) [0 tile 1 groupnorm 2 cast 3 softmax 4 other]:", {k: round(v, 2) for k, v in fprof.items()}, flush=True)
    fl = 92.353e9 * 3 * B
    out = {"batch": B, "ms_per_step": round(ms, 2), "steps_per_s": round(1000 / ms, 3), "fwd_ms": round(fwd, 2), "bwd_ms": round(bwd, 2), "adam_ms": round(adam, 2),
           "algorithmic_tflop_per_step": round(fl / 1e12, 3), "achieved_tflops": round(fl / (ms * 1e-3) / 1e12, 1), "losses": [round(l, 1) for l in losses]}
    print(json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/train_bench_%s.json" % os.environ.get("TAG", "x"), "w"))
