import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, sr3_b200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = sr3_b200.define_G(bench.make_opt(bench.SCHED)).to(dev)
net.set_new_noise_schedule(bench.SCHED, dev)
eng = net.denoise_fn.engine(B, conditional=True, channels=3)
eng.loop_begin((torch.rand(B, 3, 128, 128) * 2 - 1).to(dev), torch.randn(B, 3, 128, 128).to(dev), seed=1)
eng.steps(1999, 5); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.steps(1994, 30); e1.record(); torch.cuda.synchronize()
print("graph ms/step", e0.elapsed_time(e1) / 30)
prof = eng.profile_step(1000, reps=5)
names = {0: "gemm", 1: "prep", 2: "cast", 3: "softmax", 4: "other"}
tot = {}
for k, m, fl, by in prof:
    d = tot.setdefault(names[k], [0, 0.0]); d[0] += 1; d[1] += m
print({k: (n, round(v, 3)) for k, (n, v) in tot.items()}, "eager sum", sum(m for _, m, _, _ in prof))
for i, (k, m, fl, by) in enumerate(prof):
    print(f"{i:3d} {names[k]:7s} {m*1000:7.1f}us GF={fl/1e9:7.2f} MB={by/1e6:7.1f}")
