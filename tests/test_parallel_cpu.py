"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding + gather for sampling, bucketed
overlapped gradient all-reduce with the reference's loss normalisation for training."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import wsr

par = wsr.sub("parallel")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeDiffusion:
    """Stands in for ResDiffDiffusion on CPU: a deterministic, batch-COUPLED function of the local batch (like the
    reference's 4-D FFT), so the test also pins the 'independent unit = local batch' semantics."""

    def super_resolution(self, x_in):
        sr = x_in["SR"]
        return sr * 2.0 + sr.mean() + 1.0


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        n = 5                                              # ragged: 3 + 2
        batch = {"SR": torch.randn(n, 1, 4, 8), "LR": torch.randn(n, 1, 1, 2), "name": "x"}
        out = par.sharded_super_resolution(_FakeDiffusion(), batch)
        exp = []
        for r in range(world):
            lo, hi = par.shard_bounds(n, r, world)
            loc = batch["SR"][lo:hi]
            exp.append(loc * 2.0 + loc.mean() + 1.0)
        ok_sample = torch.allclose(out, torch.cat(exp, 0))

        # data-parallel training step on a toy model: per-rank sum-loss / GLOBAL numel, SUM all-reduce
        torch.manual_seed(1)
        model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
        ref = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
        ref.load_state_dict(model.state_dict())
        x, y = torch.randn(8, 6), torch.randn(8, 3)
        lo, hi = par.shard_bounds(8, rank, world)
        bucketer = par.GradBucketer(model.parameters(), bucket_mb=1e-4)      # tiny buckets -> several of them
        loss = (model(x[lo:hi]) - y[lo:hi]).abs().sum() / y.numel()
        loss.backward()
        bucketer.finish()
        (ref(x) - y).abs().sum().div(y.numel()).backward()
        ok_grad = all(torch.allclose(a.grad, b.grad, atol=1e-6) for a, b in zip(model.parameters(), ref.parameters()))
        # flat-buffer reducer (the train plan's gradient layout): buckets are launched as ranges become final
        class _Plan:
            gflat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
            on_ready = None
        plan = _Plan()
        red = par.FlatGradReducer(plan, bucket_mb=4 * 300 / (1 << 20))       # 300-element buckets
        plan.on_ready(0, 650)
        plan.on_ready(650, 1000)
        done = red.finish()
        ok_flat = torch.allclose(plan.gflat, torch.arange(1000, dtype=torch.float32) * 3) and \
            done == [(0, 300), (300, 600), (600, 650), (650, 950), (950, 1000)]
        # per-rank Philox streams: same base seed / same torch generator state on every rank, different keys
        D = wsr.sub("models.diffusion_models.diffusion")
        gd = D.GaussianDiffusion(denoise_fn=None)
        torch.manual_seed(0)
        drawn = gd._next_seed()
        gd.sample_seed = 11
        fixed = gd._next_seed()
        seeds = [None, None]
        dist.all_gather_object(seeds, (drawn, fixed, D.rank_stream_seed(7)))
        ok_seed = (seeds[0][0] != seeds[1][0] and seeds[0][1] != seeds[1][1] and seeds[0][2] != seeds[1][2]
                   and seeds[0][1] == 11 and seeds[0][2] == 7)
        q.put((rank, bool(ok_sample), bool(ok_grad) and bool(ok_flat) and bool(ok_seed), len(bucketer.buckets)))
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    assert [par.shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [par.shard_bounds(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert par.shard_bounds(0, 0, 2) == (0, 0)


def test_world2_gloo_sharded_sampling_and_grad_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok_sample, ok_grad, nb in res:
        assert ok_sample, "rank %d: sharded sampling mismatch" % rank
        assert ok_grad, "rank %d: all-reduced gradients differ from the single-process reference" % rank
        assert nb >= 2


def test_rank_stream_seed_is_identity_on_rank0_and_spreads_other_ranks():
    D = wsr.sub("models.diffusion_models.diffusion")
    assert D.rank_stream_seed(1234, rank=0) == 1234
    keys = {D.rank_stream_seed(1234, rank=r) for r in range(64)}
    assert len(keys) == 64 and all(0 <= k < 2 ** 62 for k in keys)
    assert D.rank_stream_seed(1234, rank=3) != D.rank_stream_seed(1235, rank=3)
