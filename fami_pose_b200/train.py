"""Sync-free training step: forward + the loss of
engine/core/functions/alignment_mi_function_term6_1.py:119-148 + backward on our kernels, ONE bucketed
gradient all-reduce (parallel.GradBuckets, NCCL), and a fused Adam update per bucket (fami_adam_step;
posetimation/optimizer/optimizer.py:66-72, MultiStepLR of scheduler.py:14-26).  No .item()/.cpu() inside:
the loss comes back as a device scalar.  fp32 arm; the backbone is frozen as in the reference default."""
import torch

from . import _lib, ops
from .loss import JointMSELoss, combine_losses
from .parallel import GradBuckets


def multistep_lr(base_lr, epoch, milestones, factor):
    """torch.optim.lr_scheduler.MultiStepLR(optimizer, LR_STEP, LR_FACTOR) evaluated at `epoch`."""
    return base_lr * factor ** sum(1 for m in milestones if epoch >= m)


class TrainStep:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, w_mse=1.0, alpha=0.5, beta=0.1,
                 bucket_bytes=64 << 20):
        self.model = model
        self.lr, self.betas, self.eps = lr, betas, eps
        self.w_mse, self.alpha, self.beta = w_mse, alpha, beta
        self.criterion = JointMSELoss()
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise ValueError("no trainable parameters")
        for p in params:
            if p.dtype != torch.float32 or not p.is_cuda:
                raise ValueError("TrainStep expects fp32 CUDA parameters")
        self.buckets = GradBuckets(params, bucket_bytes)
        # parameters become views into flat buffers laid out like the gradient buckets
        self.flat_p, self.exp_avg, self.exp_avg_sq = [], [], []
        for b, g in zip(self.buckets.buckets, self.buckets.flat):
            buf = torch.empty_like(g)
            off = 0
            for p in b:
                n = p.numel()
                buf[off:off + n].copy_(p.detach().reshape(-1))
                p.data = buf[off:off + n].view_as(p)
                off += n
            self.flat_p.append(buf)
            self.exp_avg.append(torch.zeros_like(g))
            self.exp_avg_sq.append(torch.zeros_like(g))
        self.step = 0

    def loss(self, kf_x, sup_x, target, target_weight):
        out = self.model(kf_x, sup_x)
        final_hm, mi = out[0], out[2]
        mse = self.criterion(final_hm, target, target_weight)
        return combine_losses(mse, mi, self.w_mse, self.alpha, self.beta), final_hm

    def __call__(self, kf_x, sup_x, target, target_weight):
        self.buckets.zero()
        loss, final_hm = self.loss(kf_x, sup_x, target, target_weight)
        loss.backward()
        self.buckets.allreduce_mean()
        self.step += 1
        for p, g, m, v in zip(self.flat_p, self.buckets.flat, self.exp_avg, self.exp_avg_sq):
            _lib.call("fami_adam_step", ops._ptr(p), ops._ptr(g), ops._ptr(m), ops._ptr(v), p.numel(), float(self.lr),
                      float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step, ops._stream())
        # the update went through raw pointers: bump the version counters so cached packed weights /
        # folded BatchNorm affines (ops.packed_weight, ops.folded_affine) of the trained layers refresh
        torch.autograd.graph.increment_version(self.buckets.params)
        return loss.detach(), final_hm.detach()
