"""The training row of the oracle, pinned to the unmodified reference: loss, gradients of all parameters, three
Adam iterations of DDPM.optimize_parameters (model/model.py:39-58) and a training-mode (Dropout) forward/backward with the reference's
masks.  Fixture: tests/golden/sr3_train_golden.pt (tests/golden/make_train_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import sr3_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def train_golden():
    return torch.load(os.path.join(HERE, "golden", "sr3_train_golden.pt"), weights_only=False)


def _cfg(g, dropout):
    u = g["unet_dropout"] if dropout else g["unet"]
    return orc.UNetConfig(in_channel=u["in_channel"], out_channel=u["out_channel"], inner_channel=u["inner_channel"], norm_groups=32,
                          channel_mults=tuple(u["channel_multiplier"]), attn_res=tuple(u["attn_res"]), res_blocks=u["res_blocks"],
                          dropout=u["dropout"], image_size=g["res"])


def _batch(g, i):
    gen = torch.Generator().manual_seed(1000 + i)
    B, R = g["batch"], g["res"]
    hr = torch.rand(B, 3, R, R, generator=gen) * 2 - 1
    sr = torch.rand(B, 3, R, R, generator=gen) * 2 - 1
    noise = torch.randn(B, 3, R, R, generator=gen)
    return hr, sr, noise


def _check_signature(t, sig, rtol):
    f = t.detach().flatten()
    assert f.numel() == sig["numel"]
    stride = max(1, f.numel() // 16)
    scale = max(sig["norm"] / max(f.numel(), 1) ** 0.5, 1e-12)          # rms of the tensor: absolute floor for tiny entries
    assert abs(f.norm().item() - sig["norm"]) <= rtol * max(sig["norm"], 1e-12) + 1e-12
    assert torch.allclose(f[::stride][:16], sig["samples"], rtol=rtol, atol=rtol * 10 * scale)


def test_three_adam_steps_match_reference(train_golden):
    g = train_golden
    cfg = _cfg(g, False)
    sd = orc.init_state_dict(cfg, g["seed"], orthogonal=True)
    sch = orc.make_schedule(g["sched"])
    opt = orc.make_adam(sd, g["lr"])
    for i, rec in enumerate(g["steps"]):
        hr, sr, noise = _batch(g, i)
        _, gamma = orc.draw_gamma(sch, g["batch"], np.random.RandomState(rec["np_seed"]))
        if i == 0:      # gradients of the first iteration, before the update
            opt.zero_grad()
            loss = orc.train_loss(sd, cfg, sch, hr, sr, gamma, noise)
            loss.backward()
            assert abs(loss.item() - rec["loss"]) <= 1e-6 * abs(rec["loss"])
            assert set(rec["grads"]) == set(sd)
            for k, sig in rec["grads"].items():
                _check_signature(sd[k].grad, sig, 2e-4)
        l = orc.train_step(sd, opt, cfg, sch, hr, sr, gamma, noise)
        assert abs(l - rec["loss"]) <= 2e-6 * abs(rec["loss"]), (i, l, rec["loss"])
    for k, sig in g["params_after"].items():
        _check_signature(sd[k], sig, 1e-5)


def test_training_mode_dropout_forward_backward(train_golden):
    g = train_golden
    d = g["dropout"]
    cfg = _cfg(g, True)
    sd = orc.init_state_dict(cfg, g["seed"], orthogonal=True)
    for v in sd.values():
        v.requires_grad_(True)
    sch = orc.make_schedule(g["sched"])
    masks = {}
    for k, (bits, shape) in d["masks"].items():
        keep = np.unpackbits(bits.numpy())[: int(np.prod(shape))].reshape(shape)
        masks[k] = torch.from_numpy(keep.astype(np.float32)) / (1.0 - d["p"])
    # Dropout sits in block2 of every ResnetBlock and nowhere else (unet.py:100-101)
    downs, mid, ups = orc.unet_topology(cfg)
    assert sorted(masks) == sorted(s.name + ".res_block.block2" for s in downs + mid + ups if s.kind == "res")
    hr, sr, noise = _batch(g, d["batch_index"])
    _, gamma = orc.draw_gamma(sch, g["batch"], np.random.RandomState(d["np_seed"]))
    loss = orc.train_loss(sd, cfg, sch, hr, sr, gamma, noise, dropout_masks=masks)
    assert abs(loss.item() - d["loss"]) <= 1e-6 * abs(d["loss"]), (loss.item(), d["loss"])
    loss.backward()
    for k, sig in d["grads"].items():
        _check_signature(sd[k].grad, sig, 2e-4)
    # and the eval-mode loss differs (the masks matter)
    with torch.no_grad():
        ev = orc.train_loss(sd, cfg, sch, hr, sr, gamma, noise)
    assert abs(ev.item() - d["loss"]) > 1e-5


def test_dgrad_identity_used_by_the_gpu_test():
    """conv3x3(dY, W') with W'[ci, co, r, s] = W[co, ci, 2-r, 2-s] is the data gradient of conv3x3(X, W) (stride 1, pad 1): the
    identity behind tests/test_gpu_kernels.py::test_conv_dgrad_is_the_forward_kernel_on_mirrored_weights, checked on CPU."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 6, 9, 7, generator=g, requires_grad=True)
    w = torch.randn(5, 6, 3, 3, generator=g)
    dy = torch.randn(2, 5, 9, 7, generator=g)
    (ref,) = torch.autograd.grad(F.conv2d(x, w, None, padding=1), x, dy)
    wp = w.flip(2, 3).transpose(0, 1).contiguous()
    assert torch.allclose(F.conv2d(dy, wp, None, padding=1), ref, atol=1e-5, rtol=1e-5)
