"""Run a few reverse steps of the benchmark configuration, then ONE more between cudaProfilerStart / Stop; meant to be wrapped by
`ncu --profile-from-start off` (every launch of that one step is then listed / measured, nothing else)."""
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
eng.steps(1999, steps)                      # warm-up steps (not profiled with `ncu --profile-from-start off`)
torch.cuda.synchronize()
torch.cuda.profiler.start()                 # cudaProfilerStart: the next step is the one ncu sees
eng.steps(1999 - steps, 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches/step", eng.launches_per_step(), "state finite", bool(torch.isfinite(eng.read_state()).all()))
