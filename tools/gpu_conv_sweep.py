"""Timing experiments on single conv shapes (SR3_DBG / SR3_STAGES / SR3_MAX_CTAS / SR3_NO_TMA_EPI knobs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sr3_b200 import _native
torch.zeros(1).cuda()
SHAPES = {"hi64": (16, 128, 128, 64, 64), "hi128": (16, 128, 128, 128, 64), "mid128": (16, 64, 64, 128, 128), "mid256": (16, 32, 32, 256, 256),
          "lo512": (16, 8, 8, 512, 512), "lo16": (16, 16, 16, 512, 512)}
def run(tag, shape, env=None, **kw):
    for k in ("SR3_DBG", "SR3_STAGES", "SR3_MAX_CTAS", "SR3_NO_TMA_EPI", "SR3_BLOCK_N", "SR3_NO_TALL", "SR3_TALL_BN", "SR3_NO_PDL"):
        os.environ.pop(k, None)
    for k, v in (env or {}).items():
        os.environ[k] = str(v)
    B, H, W, ci, co = SHAPES[shape]
    ms = _native.bench_conv(B, H, W, ci, co, **kw)
    gf = 2.0 * B * H * W * ci * co * 9 / 1e9
    print(f"{shape:7s} {tag:34s} {ms*1000:8.1f} us  {gf/ms:8.1f} TF/s", flush=True)
for shape in ("hi64", "hi128", "mid128", "mid256", "lo16", "lo512"):
    run("graph: default", shape)
    run("graph: no-work (dbg57)", shape, {"SR3_DBG": 57})
    run("graph: no epi body (dbg1)", shape, {"SR3_DBG": 1})
    run("graph: skip stats (dbg2)", shape, {"SR3_DBG": 2})
    run("graph: skip out store (dbg4)", shape, {"SR3_DBG": 4})
    run("graph: skip stats+store (dbg6)", shape, {"SR3_DBG": 6})
    run("graph: no loads (dbg24)", shape, {"SR3_DBG": 24})
    run("graph: no MMA, no epi body (dbg33)", shape, {"SR3_DBG": 33})
    run("graph: resid", shape, resid=True)
    run("graph: no stats arg", shape, stats=False)
