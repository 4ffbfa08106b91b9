"""One training iteration (forward + backward + Adam, 16->128 config, batch 8) between cudaProfilerStart / Stop; wrap with
`ncu --profile-from-start off`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import sr3_b200
from sr3_b200 import parallel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
torch.manual_seed(0)
opt = bench.make_opt(bench.SCHED); opt["phase"] = "train"
net = sr3_b200.define_G(opt).to(dev)
net.set_loss(dev); net.set_new_noise_schedule(bench.SCHED, dev); net.train()
g = torch.Generator().manual_seed(1)
hr = (torch.rand(B, 3, 128, 128, generator=g) * 2 - 1).to(dev)
sr = (torch.rand(B, 3, 128, 128, generator=g) * 2 - 1).to(dev)
tr = parallel.DataParallelTrainer(net, lr=1e-4)
for _ in range(2):
    tr.step(hr, sr)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = tr.step(hr, sr)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", loss)
