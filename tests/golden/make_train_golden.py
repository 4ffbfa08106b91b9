"""Generate tests/golden/sr3_train_golden.pt by running the UNMODIFIED reference (imported from /root/reference, CPU fp32):
the training row of SURVEY.md section 8f -- loss, gradients of all parameters and three Adam iterations of
DDPM.optimize_parameters (model/model.py:39-58) on a tiny SR3 UNet, with every random draw recorded so the oracle (and later the
CUDA path) can replay it:

  * numpy draws of p_losses (t, gamma)         -> np.random.seed(k) before each call, replayed through oracle.draw_gamma
  * Gaussian noise                              -> passed in (p_losses(x_in, noise=...))
  * Dropout masks (block2 of every ResnetBlock) -> captured with forward hooks, stored bit-packed

    python tests/golden/make_train_golden.py

Gradients are stored as signatures (norm, sum, 16 strided samples per parameter), not in full.
"""
import os
import sys

sys.dont_write_bytecode = True
REF = os.environ.get("SR3_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

import model.networks as ref_networks  # noqa: E402  (the reference)
from oracle import sr3_oracle as orc  # noqa: E402

SCHED = {"schedule": "linear", "n_timestep": 100, "linear_start": 1e-6, "linear_end": 1e-2}
TINY = dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2], attn_res=[16], res_blocks=1, dropout=0.0)
TINY_DROP = dict(TINY, dropout=0.2)
SEED = 5
B, RES = 2, 32


def make_opt(unet):
    return {"phase": "train", "gpu_ids": None, "distributed": False,
            "model": {"which_model_G": "sr3", "finetune_norm": False, "unet": dict(unet),
                      "beta_schedule": {"train": dict(SCHED), "val": dict(SCHED)},
                      "diffusion": {"image_size": RES, "channels": 3, "conditional": True}}}


def signature(t):
    f = t.detach().flatten()
    stride = max(1, f.numel() // 16)
    return {"norm": f.norm().item(), "sum": f.double().sum().item(), "samples": f[::stride][:16].clone(), "numel": f.numel()}


def build(unet):
    torch.manual_seed(SEED)
    g = ref_networks.define_G(make_opt(unet))          # train phase: orthogonal init (networks.py:110-112)
    g.set_new_noise_schedule(SCHED, "cpu")
    g.set_loss("cpu")
    g.train()
    return g


def batch(i):
    gen = torch.Generator().manual_seed(1000 + i)
    hr = torch.rand(B, 3, RES, RES, generator=gen) * 2 - 1
    sr = torch.rand(B, 3, RES, RES, generator=gen) * 2 - 1
    noise = torch.randn(B, 3, RES, RES, generator=gen)
    return hr, sr, noise


def main():
    out = {"sched": SCHED, "seed": SEED, "unet": TINY, "unet_dropout": TINY_DROP, "batch": B, "res": RES, "lr": 1e-4}

    # ---- A. dropout = 0: loss, gradient signatures, three Adam iterations -----------------------------------------------
    g = build(TINY)
    opt = torch.optim.Adam(list(g.parameters()), lr=1e-4)           # model/model.py:39-40
    steps = []
    for i in range(3):
        hr, sr, noise = batch(i)
        np.random.seed(40 + i)
        opt.zero_grad()
        l_pix = g.p_losses({"HR": hr, "SR": sr}, noise=noise)       # forward(): diffusion.py:248-249
        b, c, h, w = hr.shape
        l_pix = l_pix.sum() / int(b * c * h * w)                    # model/model.py:50-53
        l_pix.backward()
        rec = {"np_seed": 40 + i, "loss": l_pix.item()}
        if i == 0:
            rec["grads"] = {k[len("denoise_fn."):]: signature(p.grad) for k, p in g.named_parameters()}
        opt.step()
        steps.append(rec)
        print("step", i, "l_pix", rec["loss"])
    out["steps"] = steps
    out["params_after"] = {k: signature(v) for k, v in g.denoise_fn.state_dict().items()}

    # ---- B. dropout = 0.2 (training mode): loss + gradients with the captured masks ----------------------------------------
    g = build(TINY_DROP)
    masks = {}
    hooks = []
    for name, m in g.denoise_fn.named_modules():
        if isinstance(m, nn.Dropout):
            # name = e.g. "downs.1.res_block.block2.block.2" -> key "downs.1.res_block.block2"
            key = name[: -len(".block.2")]
            def hook(mod, inp, outp, key=key):
                x = inp[0]
                keep = (outp != 0) | (x == 0)          # where x == 0 the mask is unobservable (and irrelevant)
                masks[key] = keep.clone()
            hooks.append(m.register_forward_hook(hook))
    hr, sr, noise = batch(10)
    np.random.seed(77)
    torch.manual_seed(4242)                            # drives nn.Dropout
    l_pix = g.p_losses({"HR": hr, "SR": sr}, noise=noise)
    l_pix = l_pix.sum() / int(B * 3 * RES * RES)
    l_pix.backward()
    for h in hooks:
        h.remove()
    out["dropout"] = {"np_seed": 77, "batch_index": 10, "p": 0.2, "loss": l_pix.item(),
                      "masks": {k: (torch.from_numpy(np.packbits(v.numpy().reshape(-1))), tuple(v.shape)) for k, v in masks.items()},
                      "grads": {k[len("denoise_fn."):]: signature(p.grad) for k, p in g.named_parameters()}}
    print("dropout loss", l_pix.item(), "masked blocks", sorted(masks))

    # ---- cross-check with the oracle before writing ------------------------------------------------------------------------
    cfg = orc.UNetConfig(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2), attn_res=(16,), res_blocks=1,
                         dropout=0.0, image_size=RES)
    sd = orc.init_state_dict(cfg, SEED, orthogonal=True)
    sch = orc.make_schedule(SCHED)
    oopt = orc.make_adam(sd, 1e-4)
    for i in range(3):
        hr, sr, noise = batch(i)
        _, gamma = orc.draw_gamma(sch, B, np.random.RandomState(40 + i))
        l = orc.train_step(sd, oopt, cfg, sch, hr, sr, gamma, noise)
        print("oracle step", i, l, "ref", steps[i]["loss"])
        assert abs(l - steps[i]["loss"]) <= 1e-5 * abs(steps[i]["loss"])
    torch.save(out, os.path.join(HERE, "sr3_train_golden.pt"))
    print("sr3_train_golden.pt", os.path.getsize(os.path.join(HERE, "sr3_train_golden.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
