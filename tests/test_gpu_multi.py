"""Multi-GPU sampling on real GPUs (needs >= 2 visible devices; skipped otherwise): two NCCL ranks run
sr3_b200.parallel.sharded_super_resolution and the gathered batch must equal, bit for bit, what a single process produces when it runs
the same shards itself (Philox streams are keyed by the GLOBAL sample index, so the images do not depend on the rank count)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

SCHED = {"schedule": "linear", "n_timestep": 6, "linear_start": 1e-6, "linear_end": 1e-2}
TINY_UNET = dict(in_channel=6, out_channel=3, inner_channel=64, channel_multiplier=[1, 2], attn_res=[16], res_blocks=1, dropout=0.0)


def _opt():
    return {"phase": "val", "gpu_ids": [0], "distributed": False,
            "model": {"which_model_G": "sr3", "finetune_norm": False, "unet": dict(TINY_UNET),
                      "beta_schedule": {"train": dict(SCHED), "val": dict(SCHED)},
                      "diffusion": {"image_size": 32, "channels": 3, "conditional": True}}}


def _inputs(n):
    g = torch.Generator().manual_seed(21)
    return torch.rand(n, 3, 32, 32, generator=g) * 2 - 1, torch.randn(n, 3, 32, 32, generator=g)


def _build(dev):
    import sr3_b200
    torch.manual_seed(0)
    net = sr3_b200.define_G(_opt()).to(dev)
    net.set_new_noise_schedule(SCHED, dev)
    net.eval()
    return net


def _worker(rank, world, port, n, path):
    import torch.distributed as dist
    from sr3_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        net = _build(dev)
        cond, xT = _inputs(n)
        out = parallel.sharded_super_resolution(net, cond, x_T=xT, seed=9)
        torch.save(out.cpu(), f"{path}.rank{rank}")
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(600)
@pytest.mark.parametrize("n", [4, 5])
def test_two_rank_sampling_matches_single_process(n, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from sr3_b200 import parallel
    path = str(tmp_path / "out")
    mp.spawn(_worker, args=(2, _free_port(), n, path), nprocs=2, join=True)
    outs = [torch.load(f"{path}.rank{r}") for r in range(2)]
    assert torch.equal(outs[0], outs[1]) and outs[0].shape == (n, 3, 32, 32)      # every rank holds the same gathered batch
    # single process, same shards one after the other
    net = _build(torch.device("cuda", 0))
    cond, xT = _inputs(n)
    parts = []
    for r in range(2):
        lo, hi = parallel.shard_bounds(n, 2, r)
        final, _ = net._engine(hi - lo).p_sample_loop(cond[lo:hi].cuda(), xT[lo:hi].cuda(), None, 9, lo, want_snapshots=False)
        parts.append(final.cpu())
    assert torch.equal(outs[0], torch.cat(parts, 0))
    # and the un-sharded batch agrees within the bf16 tolerance (different tile shapes / split-K factors, same Philox streams)
    whole, _ = net._engine(n).p_sample_loop(cond.cuda(), xT.cuda(), None, 9, 0, want_snapshots=False)
    rel = ((whole.cpu() - outs[0]).norm() / outs[0].norm()).item()
    assert rel < 1e-2, rel


# ---- training (config 4: data-parallel fwd + bwd + Adam, gradient all-reduce overlapped with the backward) -------------------------
def _train_inputs(n):
    g = torch.Generator().manual_seed(33)
    hr = torch.rand(n, 3, 32, 32, generator=g) * 2 - 1
    sr = torch.rand(n, 3, 32, 32, generator=g) * 2 - 1
    noise = torch.randn(n, 3, 32, 32, generator=g)
    gamma = torch.rand(n, generator=g) * 0.5 + 0.3
    return hr, sr, noise, gamma


def _build_train(dev):
    import sr3_b200
    torch.manual_seed(0)
    o = _opt(); o["phase"] = "train"
    net = sr3_b200.define_G(o).to(dev)
    net.loss_type = "l2"                       # smooth loss: the comparison below is not blurred by sign flips of the L1 gradient
    net.set_loss(dev)
    net.set_new_noise_schedule(SCHED, dev)
    net.eval()                                 # no Dropout: both runs see the same network
    return net


def _train_worker(rank, world, port, n, path):
    import torch.distributed as dist
    from sr3_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        net = _build_train(dev)
        hr, sr, noise, gamma = _train_inputs(n)
        lo, hi = parallel.shard_bounds(n, world, rank)
        tr = parallel.DataParallelTrainer(net, lr=1e-4, bucket_mb=0.25)          # small buckets: several all-reduces in flight
        loss = tr.step(hr[lo:hi].to(dev), sr[lo:hi].to(dev), gamma=gamma[lo:hi], noise=noise[lo:hi].to(dev), global_batch=n)
        torch.cuda.synchronize()
        torch.save({"loss": loss, "grad": tr.buckets.flat.cpu(), "n_buckets": len(tr.buckets.slices), "comm_ms": tr.comm_window_ms(),
                    "params": {k: v.cpu() for k, v in net.denoise_fn.state_dict().items()}}, f"{path}.rank{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_training_step_matches_single_process(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from sr3_b200 import parallel
    n = 4
    path = str(tmp_path / "tr")
    mp.spawn(_train_worker, args=(2, _free_port(), n, path), nprocs=2, join=True)
    outs = [torch.load(f"{path}.rank{r}") for r in range(2)]
    assert outs[0]["n_buckets"] >= 3
    assert torch.equal(outs[0]["grad"], outs[1]["grad"])                       # the all-reduced gradient is the same on both ranks
    for k in outs[0]["params"]:
        assert torch.equal(outs[0]["params"][k], outs[1]["params"][k]), k      # ... and so are the parameters after Adam
    # one process, whole batch
    net = _build_train(torch.device("cuda", 0))
    hr, sr, noise, gamma = _train_inputs(n)
    tr = parallel.DataParallelTrainer(net, lr=1e-4, bucket_mb=0.25)
    loss = tr.step(hr.cuda(), sr.cuda(), gamma=gamma, noise=noise.cuda(), global_batch=n)
    torch.cuda.synchronize()
    assert abs((outs[0]["loss"] + outs[1]["loss"]) - loss) < 1e-3 * abs(loss)
    g1, g2 = tr.buckets.flat.cpu(), outs[0]["grad"]
    assert ((g1 - g2).norm() / g1.norm()).item() < 1e-2
