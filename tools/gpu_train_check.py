"""Training row on the GPU next to the oracle's CPU autograd: per-parameter gradient errors (backward order), for debugging.

    python tools/gpu_train_check.py [tiny|three] [l2|l1]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _train_util as tu

CFGS = {
    "tiny": (dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2], attn_res=[16], res_blocks=1, dropout=0.0), 32, 2),
    "full": (dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2, dropout=0.0), 128, 2),
    "three": (dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2, 2], attn_res=[8], res_blocks=1, dropout=0.0), 32, 3),
}

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    loss_type = sys.argv[2] if len(sys.argv) > 2 else "l2"
    unet, R, B = CFGS[which]
    net = tu.build_train_net(unet, R, 5, loss_type)
    hr, sr, noise = tu.batch(B, R, 1000)
    gamma = tu.draw_gamma(B, 7)
    lo, go = tu.ours_loss_and_grads(net, hr, sr, gamma, noise)
    lr_, gr = tu.oracle_loss_and_grads(net, unet, R, hr, sr, gamma, noise, loss_type)
    print("loss ours %.6f oracle %.6f rel %.3e" % (lo, lr_, abs(lo - lr_) / abs(lr_)), flush=True)
    rows = tu.compare(go, gr)
    bad = 0
    for name, e, c, n in rows:
        flag = "" if e < 2e-2 else "   <<<<"
        bad += e >= 2e-2
        print("%-60s rel %.3e cos %.5f |ref| %.3e%s" % (name, e, c, n, flag), flush=True)
    print("params: %d, rel >= 2e-2: %d, worst %.3e" % (len(rows), bad, max(r[1] for r in rows)))
