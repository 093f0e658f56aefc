"""Data-parallel plumbing: one process per GPU, clips shard by batch, per-replica
BatchNorm (the reference's DataParallel semantics, engine/defaults/trainer.py:58), and ONE bucketed
gradient all-reduce per step over NCCL (NVLink 5 / NVSwitch).  Forward/eval needs no collective.

torch.distributed is plumbing here: NCCL on GPUs, gloo on CPU for the world_size-2 host-logic tests.
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialises the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns
    (rank, world, local_rank).  No-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_range(global_batch, rank, world):
    """[start, stop) of this rank's clips: global batch = per-GPU batch x #GPUs
    (datasets/zoo/build.py:40 semantics); remainders go to the lowest ranks."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def rank_seed(base_seed, rank):
    """Per-rank synthetic-data seed (seed = base + rank)."""
    return int(base_seed) + int(rank)


class GradBuckets:
    """Flat gradient buckets: grads of the trainable parameters are views into a few contiguous buffers, so the whole
    exchange is `len(buckets)` all-reduce calls (reference default: only the 1.06 M alignment-head parameters are
    trainable -> one 4.2 MB bucket; un-frozen HRNet: 258.6 MB in 64 MB buckets).

    Buckets are filled in REVERSE parameter order -- the order in which backward produces gradients -- and, with
    `overlap=True`, each bucket's all-reduce is launched from an autograd post-accumulate hook the moment its last
    gradient has been accumulated, so the collective of the head's bucket runs while backward is still inside the
    backbone (the 258.6 MB case of SURVEY.md 8e).  `finish()` waits for the outstanding collectives."""

    def __init__(self, params, bucket_bytes=64 << 20, overlap=False):
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nb = p.numel() * p.element_size()
            if cur and cur_bytes + nb > bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nb
        if cur:
            self.buckets.append(cur)
        self.flat = []
        self._bucket_of = {}
        for bi, b in enumerate(self.buckets):
            n = sum(p.numel() for p in b)
            buf = torch.zeros(n, dtype=b[0].dtype, device=b[0].device)
            off = 0
            for p in b:
                p.grad = buf[off:off + p.numel()].view_as(p)
                off += p.numel()
                self._bucket_of[id(p)] = bi
            self.flat.append(buf)
        self.overlap = bool(overlap)
        self._pending = [0] * len(self.buckets)
        self._works = []
        self._hooks = []
        self.launch_order = []          # bucket indices in the order their all-reduce was issued (tests / tracing)
        if self.overlap:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def zero(self):
        for f in self.flat:
            f.zero_()
        self._pending = [len(b) for b in self.buckets]
        self._works = []
        self.launch_order = []

    def _reduce_bucket(self, bi, async_op):
        world = dist.get_world_size()
        f = self.flat[bi]
        f.div_(world)
        self.launch_order.append(bi)
        return dist.all_reduce(f, op=dist.ReduceOp.SUM, async_op=async_op)

    def _on_grad(self, p):
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and dist.is_initialized() and dist.get_world_size() > 1:
            self._works.append(self._reduce_bucket(bi, True))

    def finish(self):
        """Waits for the all-reduces issued by the hooks and reduces any bucket whose hooks did not all fire (a
        parameter that received no gradient this step)."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        for bi, n in enumerate(self._pending):
            if n > 0 or not self.overlap:
                self._works.append(self._reduce_bucket(bi, True))
                self._pending[bi] = 0
        for w in self._works:
            w.wait()
        self._works = []

    def allreduce_mean(self, async_op=False):
        """sum over ranks / world, in place, one collective per bucket (non-overlapped form)."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return []
        if self.overlap:
            self.finish()
            return []
        works = []
        for bi in range(len(self.flat)):
            w = self._reduce_bucket(bi, async_op)
            if async_op:
                works.append(w)
        return works


def max_over_ranks(value, device):
    """Timing helper: the max over ranks of a python float (device-side all-reduce)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
