"""A/B of environment knobs on the benchmark configuration (resident state, CUDA events):  python tools/ab_env.py B "K=V,K2=V2" "..." """
import os, sys, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch, bench, sr3_b200
    B = int(sys.argv[2])
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = sr3_b200.define_G(bench.make_opt(bench.SCHED)).to(dev)
    net.set_new_noise_schedule(bench.SCHED, dev)
    eng = net.denoise_fn.engine(B, conditional=True, channels=3)
    eng.loop_begin((torch.rand(B, 3, 128, 128) * 2 - 1).to(dev), torch.randn(B, 3, 128, 128).to(dev), seed=1)
    eng.steps(1999, 5)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.steps(1990 - 40 * rep, 40); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 40)
    print(json.dumps({"ms_per_step": best, "launches": eng.launches_per_step()}))
else:
    B = sys.argv[1]
    for spec in sys.argv[2:]:
        env = dict(os.environ)
        for kv in spec.split(","):
            if "=" in kv:
                k, v = kv.split("=", 1); env[k] = v
        r = subprocess.run([sys.executable, __file__, "--child", B], env=env, capture_output=True, text=True)
        print(f"B={B} {spec or 'default':40s} {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]}", flush=True)
