#!/usr/bin/env python
"""bench.py -- diffusion steps/sec of the SR3 16->128 sampler at batch 16 (BASELINE.json metric, configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our arm   (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference's algorithm on the host cores

A "step" is one reverse-diffusion step (p_sample) of a batch of 16 images: UNet forward + posterior update.  With N GPUs the
images are partitioned (no per-step exchange): --scaling weak (default) gives every GPU its own batch of 16 and `value` sums
the batch-16 steps of all ranks; --scaling strong shards ONE batch of 16 (16/N images per GPU, latency bound below ~4 images).
`value`   : K steps of the captured step graph with the sampler state resident in HBM, CUDA events, max over ranks.
`e2e`     : the same metric through the public API call a user makes (GaussianDiffusion.super_resolution on a HOST
            tensor, schedule length K): H2D of the condition + K steps + D2H of the images (+ all-gather for N>1)
            inside the timed region.
`roofline`: tensor-core tile kernel -- algorithmic conv+attention FLOPs of one step / the summed CUDA-event durations
            of that kernel's launches in one (eager, per-launch timed) step, against the measured bf16 peak.
`cpu_baseline`: the oracle (CPU restatement of the reference, torch-CPU fp32) timed on this host's cores on a bounded
            sample.  Only this leg and --impl reference execute oracle/ ; the GPU path never does.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCHED = {"schedule": "linear", "n_timestep": 2000, "linear_start": 1e-6, "linear_end": 1e-2}
UNET = dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2, 4, 8, 8], attn_res=[16], res_blocks=2, dropout=0.2)
GLOBAL_BATCH = 16
IMAGE = 128
METRIC = "diffusion steps/sec (batch16, 16->128 SR3)"


def make_opt(sched):
    return {"phase": "val", "gpu_ids": [0], "distributed": False,
            "model": {"which_model_G": "sr3", "finetune_norm": False, "unet": dict(UNET),
                      "beta_schedule": {"train": dict(sched), "val": dict(sched)},
                      "diffusion": {"image_size": IMAGE, "channels": 3, "conditional": True}}}


def algorithmic_flops_per_image():
    """2*MACs of every conv and of QK^T / PV in one UNet forward (92.353 GFLOP for 16->128)."""
    from sr3_b200.model.sr3_modules.unet import layer_table
    layers = layer_table(UNET["in_channel"], UNET["inner_channel"], UNET["channel_multiplier"], UNET["attn_res"], UNET["res_blocks"], IMAGE)
    res, fl = IMAGE, 0.0
    for name, kind, cin, cout, attn in layers:
        if kind == "conv":
            fl += 2.0 * res * res * cin * cout * 9
        elif kind == "down":
            res //= 2
            fl += 2.0 * res * res * cin * cout * 9
        elif kind == "up":
            res *= 2
            fl += 2.0 * res * res * cin * cout * 9
        else:
            fl += 2.0 * res * res * (cin * cout * 9 + cout * cout * 9 + (cin * cout if cin != cout else 0))
            if attn:
                hw = res * res
                fl += 2.0 * hw * cout * 3 * cout + 2.0 * hw * cout * cout + 2.0 * 2.0 * hw * hw * cout
    fl += 2.0 * IMAGE * IMAGE * UNET["inner_channel"] * UNET["out_channel"] * 9
    return fl


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "hbm_gbs": d.get("hbm_gbs"), "src": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self._stop, self._th = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(s[0])) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 3 + i and s[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU legs (the only code here that touches oracle/)
# ----------------------------------------------------------------------------------------------------------------------
_CPU_THREADS = None


def tune_cpu_threads():
    """torch-CPU convs stop scaling (and can collapse) far below the core count of a 100+ core host: probe a few thread counts
    on a tiny UNet forward and keep the fastest, as anyone running the reference on this box would."""
    global _CPU_THREADS
    if _CPU_THREADS is not None:
        return _CPU_THREADS
    import torch
    from oracle import sr3_oracle as orc
    n = os.cpu_count() or 1
    cands = sorted({c for c in (n, n // 2, 64, 32, 16, 8) if 1 <= c <= n}, reverse=True)
    cfg = orc.UNetConfig(6, 3, 64, 32, (1, 2, 4, 8, 8), (16,), 2, 0.2, 128)
    sd = orc.init_state_dict(cfg, 0)
    x = torch.randn(2, 6, 128, 128)
    nl = torch.full((2, 1), 0.5)
    best, best_t = cands[0], float("inf")
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            orc.unet_forward(sd, cfg, x, nl)
            t0 = time.perf_counter()
            orc.unet_forward(sd, cfg, x, nl)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
    _CPU_THREADS = best
    torch.set_num_threads(best)
    return best


def cpu_p_sample_time(batch, reps, warm=1):
    import torch
    from oracle import sr3_oracle as orc
    tune_cpu_threads()
    cfg = orc.UNetConfig(6, 3, 64, 32, (1, 2, 4, 8, 8), (16,), 2, 0.2, 128)
    sd = orc.init_state_dict(cfg, 0)
    sch = orc.make_schedule(SCHED)
    torch.manual_seed(0)
    cond = torch.rand(batch, 3, IMAGE, IMAGE) * 2 - 1
    x = torch.randn(batch, 3, IMAGE, IMAGE)
    ts = []
    with torch.no_grad():
        for i in range(warm + reps):
            t0 = time.perf_counter()
            x = orc.p_sample(sd, cfg, sch, x, 1999 - i, torch.randn_like(x), cond)
            if i >= warm:
                ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args, rank, world):
    """Reference arm: the reference's algorithm (oracle port, torch-CPU fp32, all host threads) on the same config.
    The Python reference itself cannot travel to the GPU box, so kind = "port" (the oracle is pinned to it by
    tests/test_oracle.py)."""
    if rank != 0:
        return
    import torch
    probe = cpu_p_sample_time(2, 1, warm=1)[0]                       # seconds per step at 2 images
    budget = 150.0
    b = int(max(1, min(GLOBAL_BATCH, (budget / max(args.steps + args.warmup, 1)) / (probe / 2.0))))
    ts = cpu_p_sample_time(b, args.steps, warm=args.warmup)
    per_step_full = (sum(ts) / len(ts)) * (GLOBAL_BATCH / b)
    v = 1.0 / per_step_full
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step_full * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "sr_sr3_16_128.json sampling, global batch 16, p_sample on host cores", "global_batch": GLOBAL_BATCH},
            "cpu_baseline": {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
                             "sample": f"p_sample on {b} of the 16 images per step, time scaled by 16/{b}; {args.steps} steps after {args.warmup} warm-up"},
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = every GPU samples its own batch of 16 images (global batch 16N, value in batch-16 steps/s); "
                         "strong = ONE batch of 16 images sharded over the GPUs")
    ap.add_argument("--profile-out", default=None, help="write the per-launch timing table of one step to this JSON file")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    import sr3_b200
    assert torch.cuda.is_available(), "bench.py (our arm) needs a B200; there is no CPU fallback"
    assert world == args.gpus or world == 1, f"WORLD_SIZE={world} but --gpus {args.gpus}"
    W = max(args.warmup, 3)
    K = args.steps
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("SR3_NCCL_DEBUG", "WARN")     # keep NCCL's version banner off stdout (one JSON line only)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    # The path partitions by image (no per-step exchange).  Weak scaling (default, tier rule 5): each rank runs the configuration the metric
    # is quoted on (16 images); `value` counts batch-16 steps of ALL ranks.  Strong scaling shards one batch of 16.
    weak = args.scaling == "weak"
    if weak:
        per, global_batch = GLOBAL_BATCH, GLOBAL_BATCH * world
    else:
        assert GLOBAL_BATCH % world == 0
        per, global_batch = GLOBAL_BATCH // world, GLOBAL_BATCH
    units = global_batch / GLOBAL_BATCH          # batch-16 steps done per reverse step of the whole job

    torch.manual_seed(0)
    net = sr3_b200.define_G(make_opt(SCHED)).to(dev)
    net.set_new_noise_schedule(SCHED, dev)
    net.eval()
    g = torch.Generator().manual_seed(0)
    cond_all = torch.rand(global_batch, 3, IMAGE, IMAGE, generator=g) * 2 - 1
    xT_all = torch.randn(global_batch, 3, IMAGE, IMAGE, generator=g)
    lo = rank * per
    cond_h = cond_all[lo:lo + per].contiguous().pin_memory()
    xT_h = xT_all[lo:lo + per].contiguous().pin_memory()
    eng = net.denoise_fn.engine(per, conditional=True, channels=3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident steps: `value`
    T = SCHED["n_timestep"]

    def resident_steps(engine, c_h, x_h, first_index, sampler=None):
        """K graph-launched reverse steps with the sampler state resident in HBM; CUDA events, max over ranks (ms)."""
        engine.loop_begin(c_h.to(dev), x_h.to(dev), seed=1234, first_index=first_index)
        engine.steps(T - 1, W)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        remaining, t = K, T - 1 - W
        while remaining > 0:                 # restart from T-1 if K is longer than the schedule
            n = min(remaining, t + 1)
            engine.steps(t, n)
            remaining -= n
            t = T - 1
        e1.record()
        barrier()
        t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        assert torch.isfinite(engine.read_state()).all(), "sampler state is not finite"
        return float(t_ms.item())

    with ClockSampler(local) as clk:
        ms = resident_steps(eng, cond_h, xT_h, lo)
    value = units * K / (ms * 1e-3)

    # the other scaling mode at N > 1, as a supplementary number (same timing rules): weak run -> also time ONE batch of 16 sharded
    # over the ranks; strong run -> also time 16 images per rank
    other = None
    if world > 1:
        o_per = GLOBAL_BATCH // world if weak else GLOBAL_BATCH
        o_lo = rank * o_per
        go = torch.Generator().manual_seed(1)
        o_cond = (torch.rand(o_per, 3, IMAGE, IMAGE, generator=go) * 2 - 1).pin_memory()
        o_x = torch.randn(o_per, 3, IMAGE, IMAGE, generator=go).pin_memory()
        o_eng = net.denoise_fn.engine(o_per, conditional=True, channels=3)
        o_ms = resident_steps(o_eng, o_cond, o_x, o_lo)
        o_units = 1.0 if weak else float(world)
        other = {"scaling": "strong" if weak else "weak", "value": o_units * K / (o_ms * 1e-3), "unit": "steps/s", "ms_per_step": o_ms / K,
                 "per_gpu_batch": o_per, "global_batch": o_per * world}
        del o_eng

    # ---------------- end to end through the public API on host tensors: `e2e`
    schedK = dict(SCHED, n_timestep=K)
    net.set_new_noise_schedule(schedK, dev)
    out_h = None
    gathered = torch.empty(global_batch, 3, IMAGE, IMAGE, device=dev) if world > 1 else None
    e2e_s = []
    for it in range(4):                      # first pass warms the allocator / graph for this schedule; median of the other three
        barrier()
        t0 = time.perf_counter()
        out_h = eng.super_resolution_host(cond_h, xT_h, seed=1234, first_index=lo)     # H2D + K steps + D2H inside one native call
        if world > 1:
            dist.all_gather_into_tensor(gathered, out_h.to(dev, non_blocking=True))
        barrier()
        e2e_s.append(time.perf_counter() - t0)
    e2e_t = torch.tensor([sorted(e2e_s[1:])[1]], device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = units * K / float(e2e_t.item())
    img_bytes = per * 3 * IMAGE * IMAGE * 4
    net.set_new_noise_schedule(SCHED, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (tensor-core tile kernel), per-launch CUDA events, rank 0
    prof = eng.profile_step(1000, reps=3)
    kind_names = {0: "gemm_tile_kernel", 1: "prep_kernel(groupnorm+silu)", 2: "cast_kernel", 3: "softmax_kernel", 4: "other", 5: "attn_kernel"}
    by_kind = {}
    for k, m, fl, by in prof:
        d = by_kind.setdefault(kind_names[k], {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["launches"] += 1; d["ms"] += m; d["flops"] += fl; d["bytes"] += by
    gemm = by_kind["gemm_tile_kernel"]
    alg_flops_step = algorithmic_flops_per_image() * per
    peaks = measured_peaks()
    attn_ms = by_kind.get("attn_kernel", {"ms": 0.0})["ms"]          # fused attention core: its FLOPs are part of the algorithmic count
    tensor_ms = gemm["ms"] + attn_ms
    achieved = alg_flops_step / (tensor_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("gemm_tile_kernel_dram_bytes_per_step")
    roof = {"bound": "tensor", "kernel": "gemm_tile_kernel (all %d launches of one step%s)" % (gemm["launches"], " + attn_kernel" if attn_ms > 0 else ""), "achieved": achieved, "peak": peaks["tflops"],
            "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": traffic, "peak_source": peaks["src"],
            "algorithmic_flops_per_step": alg_flops_step, "executed_flops_per_step": gemm["flops"] + by_kind.get("attn_kernel", {"flops": 0.0})["flops"], "kernel_ms_per_step": tensor_ms,
            "step_ms_eager_sum": sum(m for _, m, _, _ in prof),
            "step_frac_of_tensor_roofline": (alg_flops_step / peaks["tflops"] / 1e12) / (ms * 1e-3 / K),
            "by_kernel": {k: {"launches": v["launches"], "ms": round(v["ms"], 4), "GB_per_s": (v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None)}
                          for k, v in by_kind.items()}}
    if args.profile_out:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        json.dump({"per_launch": [{"kind": kind_names[k], "ms": m, "flops": fl, "bytes": by} for k, m, fl, by in prof], "summary": roof}, open(args.profile_out, "w"), indent=1)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        b = 4
        ts = cpu_p_sample_time(b, 3, warm=1)
        per_full = (sum(ts) / len(ts)) * (GLOBAL_BATCH / b)
        cpu = {"value": 1.0 / per_full, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"oracle p_sample on {b} of 16 images, 3 steps after 1 warm-up, time scaled by 16/{b}"}

    line = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "sr_sr3_16_128.json sampling (configs[1]): batch 16 per GPU, T=2000 linear schedule, random-init weights" if weak else
                                   "sr_sr3_16_128.json sampling (configs[1]): ONE batch of 16 sharded over the GPUs, T=2000 linear schedule, random-init weights",
                       "global_batch": global_batch, "per_gpu_batch": per, "parallelism": f"batch-sharded x{world}, no per-step collective",
                       "value_unit_note": "steps/s of batch-16 work: (images x reverse steps per second) / 16, summed over all ranks",
                       "l2": "per-step working set (~%.1f GB of activations+weights) exceeds the 126 MB L2; no explicit flush" % (eng.workspace_bytes() / 2 ** 30),
                       "image_steps_per_s": value * GLOBAL_BATCH},
            "e2e": {"value": e2e_val, "unit": "steps/s", "h2d_bytes_per_step": 2 * img_bytes / K, "d2h_bytes_per_step": img_bytes / K,
                    "api": "GaussianDiffusion.super_resolution on host tensors (sr3_super_resolution_host), schedule length = steps; median of 3 calls",
                    "calls_s": [round(x, 6) for x in e2e_s[1:]]},
            "gpu_launches": eng.launches_per_step() * K, "launches_per_step": eng.launches_per_step(),
            "clocks": clk.summary(), "roofline": roof, "cpu_baseline": cpu}
    if other is not None:
        line["other_scaling_mode"] = other
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
