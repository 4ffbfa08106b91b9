"""Run a few eager (SR3_NO_GRAPH=1) or graph reverse steps of the benchmark configuration; meant to be wrapped by ncu."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import sr3_b200

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = sr3_b200.define_G(bench.make_opt(bench.SCHED)).to(dev)
net.set_new_noise_schedule(bench.SCHED, dev)
g = torch.Generator().manual_seed(0)
cond = torch.rand(batch, 3, 128, 128, generator=g) * 2 - 1
xT = torch.randn(batch, 3, 128, 128, generator=g)
eng = net.denoise_fn.engine(batch, conditional=True, channels=3)
eng.loop_begin(cond.to(dev), xT.to(dev), seed=1, first_index=0)
eng.steps(1999, steps)
torch.cuda.synchronize()
print("launches/step", eng.launches_per_step(), "state finite", bool(torch.isfinite(eng.read_state()).all()))
