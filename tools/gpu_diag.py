"""First-contact diagnostics for a GPU box: every check runs in its own process (a device trap poisons the
CUDA context) under a timeout, and prints enough to debug descriptor / layout mistakes from one run.
Usage: python tools/gpu_diag.py [check ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHECKS = {}


def check(fn):
    CHECKS[fn.__name__] = fn
    return fn


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _gemm(M, N, K, bn):
    import torch
    from sr3_b200 import _native
    g = torch.Generator().manual_seed(1)
    a = torch.randn(M, K, generator=g).bfloat16()
    b = torch.randn(N, K, generator=g).bfloat16()
    d = _native.test_gemm(a.cuda(), b.cuda(), bn).cpu()
    ref = a.float() @ b.float().t()
    e = _rel(d, ref)
    print(f"gemm M={M} N={N} K={K} bn={bn}: rel={e:.3e} finite={bool(torch.isfinite(d).all())}")
    if e > 1e-4:
        print(" got row0[:8]", d[0, :8].tolist())
        print(" ref row0[:8]", ref[0, :8].tolist())
        print(" got col0[:8]", d[:8, 0].tolist())
        print(" ref col0[:8]", ref[:8, 0].tolist())
        # is it a K-slice problem?  compare against partial sums over the first 16/32/64 of K
        for kk in (16, 32, 48, 64):
            if kk <= K:
                part = a[:, :kk].float() @ b[:, :kk].float().t()
                print(f"  vs first {kk} of K: rel={_rel(d, part):.3e}")
        rowerr = (d - ref).abs().amax(dim=1)
        print(" rows with err>1e-2:", (rowerr > 1e-2).nonzero().flatten()[:32].tolist())
    return e < 1e-4


@check
def gemm64():
    return _gemm(128, 64, 64, 64)


@check
def gemm64_k256():
    return _gemm(128, 64, 256, 64)


@check
def gemm128():
    return _gemm(256, 256, 512, 128)


@check
def gemm256():
    return _gemm(256, 512, 1024, 256)


def _conv(B, H, W, Cin, Cout, k, s):
    import torch
    import torch.nn.functional as F
    from sr3_b200 import _native
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, Cin, H, W, generator=g).bfloat16()
    w = torch.randn(Cout, Cin, k, k, generator=g) * (1.0 / (Cin * k * k) ** 0.5)
    bias = torch.randn(Cout, generator=g)
    y, st = _native.test_conv(x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), bias.cuda(), k, s, want_stats=True)
    y = y.cpu().permute(0, 3, 1, 2)
    ref = F.conv2d(x.float(), w.bfloat16().float(), bias, stride=s, padding=k // 2)
    e = _rel(y, ref)
    es = _rel(st.cpu()[..., 0], ref.sum(dim=(2, 3)))
    eq = _rel(st.cpu()[..., 1], (ref * ref).sum(dim=(2, 3)))
    print(f"conv B={B} {H}x{W} {Cin}->{Cout} k={k} s={s}: rel={e:.3e} stats_sum={es:.3e} stats_sq={eq:.3e}")
    if e > 1e-4:
        err = (y - ref).abs()
        print(" per-image max err", err.amax(dim=(1, 2, 3)).tolist())
        print(" err map (b=0, c=0) rows 0..7 cols 0..7\n", err[0, 0, :8, :8])
        # which taps are missing?  compare with interior-only
        inner = err[:, :, 1:-1, 1:-1].amax().item() if H > 2 else -1
        print(" interior max err", inner, " border max err", err.amax().item())
    return e < 1e-4 and es < 1e-3 and eq < 1e-3


@check
def conv1x1():
    return _conv(2, 16, 16, 64, 64, 1, 1)


@check
def conv3x3():
    return _conv(2, 16, 16, 64, 64, 3, 1)


@check
def conv3x3_8x8():
    return _conv(2, 8, 8, 128, 128, 3, 1)


@check
def conv3x3_big():
    return _conv(2, 64, 64, 192, 128, 3, 1)


@check
def conv_stride2():
    return _conv(2, 32, 32, 64, 64, 3, 2)


@check
def tiny_unet():
    import torch
    import sr3_b200
    from oracle import sr3_oracle as orc
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "sr3_golden.pt"), map_location="cpu", weights_only=False)["tiny_unet"]
    sched = {"schedule": "linear", "n_timestep": 2000, "linear_start": 1e-6, "linear_end": 1e-2}
    opt = {"phase": "val", "gpu_ids": [0], "distributed": False,
           "model": {"which_model_G": "sr3", "finetune_norm": False,
                     "unet": dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2], attn_res=[16], res_blocks=1, dropout=0.0),
                     "beta_schedule": {"train": sched, "val": sched}, "diffusion": {"image_size": 32, "channels": 3, "conditional": True}}}
    torch.manual_seed(0)
    net = sr3_b200.define_G(opt).cuda()
    net.set_new_noise_schedule(sched, "cuda")
    eps = net.denoise_fn(gold["x"].cuda(), gold["noise_level"].cuda()).cpu()
    eng = net.denoise_fn.engine(2)
    print("launches/step:", eng.launches_per_step(), "workspace MB:", eng.workspace_bytes() / 2 ** 20)
    ok = True
    for name, ref in gold["taps"].items():
        e = _rel(eng.read_activation(name).cpu(), ref)
        print(f"  {name:10s} rel={e:.3e}")
        ok &= e < 1e-2
    e = _rel(eps, gold["eps"])
    print(f"  eps rel={e:.3e}")
    return ok and e < 1e-2


def main():
    names = sys.argv[1:] or list(CHECKS)
    if len(names) == 1 and names[0].startswith("--run="):
        ok = CHECKS[names[0][6:]]()
        sys.exit(0 if ok else 3)
    summary = {}
    for n in names:
        print(f"===== DIAG {n}", flush=True)
        r = subprocess.run(["timeout", "300", sys.executable, os.path.abspath(__file__), f"--run={n}"], cwd=ROOT)
        summary[n] = r.returncode
        print(f"===== DIAG {n} -> rc={r.returncode}", flush=True)
    print("DIAG SUMMARY", summary)


if __name__ == "__main__":
    main()
