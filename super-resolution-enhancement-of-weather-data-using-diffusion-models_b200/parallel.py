"""Multi-GPU plumbing: one process per GPU (torchrun), ``torch.distributed`` over NCCL/NVLink.

The reference's only parallelism is single-process ``nn.DataParallel`` (models/diffusion_models/networks.py:166-168):
per iteration it broadcasts all 98.9 M parameters, scatters the batch dict and reduces the gradients to GPU 0; sampling
always runs on one GPU (model.py:79-82).  Here:

  * sampling shards the batch across ranks -- independent units, NO collective inside the T-step loop, one all_gather of
    the (B, C, H, W) result at the end.  For the ``resdiff`` architecture the independent unit is the LOCAL batch, not
    the sample: FD_Info_Spliter runs a 4-D FFT over (B, C, H, W) (resdiff/fd_info_spliter.py:63), so results depend on
    how the batch is partitioned; parity is defined per local batch.
  * training is data-parallel with a bucketed gradient all-reduce launched from autograd hooks as soon as a bucket's
    last gradient is ready, i.e. overlapped with the rest of backward.  Loss semantics of the reference are kept:
    every rank contributes sum-loss / GLOBAL numel (model.py:64-66) and gradients are SUM-reduced.
"""
import os
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous slice [lo, hi) of n items for ``rank``; remainders go to the lowest ranks."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x_in, rank, world):
    """Slice every tensor of a batch dict (and an optional list of months) along dim 0."""
    if isinstance(x_in, dict):
        n = next(v for v in x_in.values() if torch.is_tensor(v)).shape[0]
        lo, hi = shard_bounds(n, rank, world)
        return {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n else v) for k, v in x_in.items()}
    n = x_in.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return x_in[lo:hi]


def gather_batch(local, n_total, group=None):
    """all_gather of per-rank results with (possibly) unequal local batch sizes; every rank gets the full batch."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    max_b = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_b,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)


def sharded_super_resolution(diffusion, x_in, group=None):
    """Batch-sharded ``super_resolution``: rank r super-resolves its slice; the result is gathered on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return diffusion.super_resolution(x_in)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = x_in["SR"].shape[0]
    local = diffusion.super_resolution(shard_batch(x_in, rank, world))
    return gather_batch(local, n, group)


class GradBucketer:
    """Bucketed, overlapped gradient all-reduce (SUM).

    Parameters are grouped in REVERSE registration order (the order their gradients become ready in backward: for the
    ResDiff UNet final_conv, ups, mid, downs, ...) into buckets of ~``bucket_mb`` MB.  A post-accumulate-grad hook
    counts ready gradients; when a bucket is complete its gradients are flattened and all-reduced asynchronously, and
    ``finish()`` (called before ``optimizer.step()``) waits and scatters the results back.
    """

    def __init__(self, params, bucket_mb=32.0, group=None):
        self.group = group
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, cur_bytes = [], [], 0
        limit = int(bucket_mb * (1 << 20))
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= limit:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._ready = [0] * len(self.buckets)
        self._pending = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def _on_grad(self, p):
        i = self._bucket_of[id(p)]
        self._ready[i] += 1
        if self._ready[i] == len(self.buckets[i]):
            self._launch(i)

    def _launch(self, i):
        grads = [p.grad for p in self.buckets[i]]
        flat = torch.cat([g.reshape(-1) for g in grads])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending.append((work, flat, grads))

    def finish(self):
        """Wait for all in-flight buckets, write the reduced gradients back, reset for the next iteration."""
        for i, n in enumerate(self._ready):           # parameters that received no gradient this step
            if 0 < n < len(self.buckets[i]) or (n == 0 and any(p.grad is not None for p in self.buckets[i])):
                for p in self.buckets[i]:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                self._launch(i)
        for work, flat, grads in self._pending:
            work.wait()
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        self._pending.clear()
        self._ready = [0] * len(self.buckets)

    def remove(self):
        for h in self._hooks:
            h.remove()


class FlatGradReducer:
    """Bucketed, overlapped SUM all-reduce over the flat gradient buffer of a ``UNetTrainPlan``.

    The plan lays its parameter gradients out in the order the backward pass finishes them and calls
    ``on_ready(lo, hi)`` as soon as ``gflat[lo:hi]`` is final; this class cuts that range into ~``bucket_mb`` MB buckets
    and launches one asynchronous all-reduce per bucket right away, so the exchange of the up-path gradients overlaps
    the backward of the down path (NCCL orders each collective after the kernels already enqueued on the compute
    stream).  ``finish()`` (before ``optimizer.step()``) waits for all of them.  Each rank must scale its loss by
    1 / (GLOBAL number of elements) -- the reference's ``l_pix.sum() / (b*c*h*w)`` over the whole DataParallel batch
    (models/diffusion_models/model.py:64-66) -- so that the SUM of the per-rank gradients is the reference's gradient.
    """

    def __init__(self, plan, bucket_mb=None, group=None):
        if bucket_mb is None:
            bucket_mb = float(os.environ.get("WSR_BUCKET_MB", "32"))
        self.plan, self.group = plan, group
        self.bucket_elems = max(1, int(bucket_mb * (1 << 20)) // 4)
        self._pending = []
        self.launched = []                 # (lo, hi) of every bucket of the last step, for tests / logging
        plan.on_ready = self._on_ready

    def _on_ready(self, lo, hi):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        a = lo
        while a < hi:
            b = min(hi, a + self.bucket_elems)
            work = dist.all_reduce(self.plan.gflat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append(work)
            self.launched.append((a, b))
            a = b

    def finish(self):
        for w in self._pending:
            w.wait()
        self._pending.clear()
        done, self.launched = self.launched, []
        return done

    def remove(self):
        self.plan.on_ready = None
