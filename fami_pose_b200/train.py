"""Sync-free training step: forward + the loss of
engine/core/functions/alignment_mi_function_term6_1.py:119-148 + backward on our kernels, a bucketed gradient
all-reduce (parallel.GradBuckets, NCCL) issued from autograd hooks while backward is still running, and a fused Adam
update per bucket (fami_adam_step; posetimation/optimizer/optimizer.py:66-72, MultiStepLR of scheduler.py:14-26).
No .item()/.cpu() inside: the loss comes back as a device scalar.  fp32 storage ('fp32' exact / 'tf32' tensor-core
arm); the backbone is frozen as in the reference default, or trained when FREEZE_HRNET_WEIGHTS is false.

Also here: checkpoints in the reference's own format (engine/defaults/checkpoints.py:45-107) and the engine-facing
core-function wrapper (engine/core/base.py:17-40, alignment_mi_function_term6_1.py:72-220)."""
import math
import os
import os.path as osp

import torch

from . import _lib, ops
from .loss import JointMSELoss, combine_losses
from .parallel import GradBuckets


def multistep_lr(base_lr, epoch, milestones, factor):
    """torch.optim.lr_scheduler.MultiStepLR(optimizer, LR_STEP, LR_FACTOR) evaluated at `epoch`."""
    return base_lr * factor ** sum(1 for m in milestones if epoch >= m)


class TrainStep:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, w_mse=1.0, alpha=0.5, beta=0.1,
                 bucket_bytes=64 << 20, overlap=True):
        self.model = model
        self.lr, self.betas, self.eps = lr, betas, eps
        self.w_mse, self.alpha, self.beta = w_mse, alpha, beta
        self.criterion = JointMSELoss()
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("no trainable parameters")
        for _, p in named:
            if p.dtype != torch.float32 or not p.is_cuda:
                raise ValueError("TrainStep expects fp32 CUDA parameters")
        self.param_names = [n for n, _ in named]          # model.parameters() order == torch.optim's param indices
        self.params = [p for _, p in named]
        self.buckets = GradBuckets(self.params, bucket_bytes, overlap=overlap)
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
        self._graph = None
        self._hyper = None

    # ---- one step -------------------------------------------------------------------------------------------------
    def loss(self, kf_x, sup_x, target, target_weight):
        out = self.model(kf_x, sup_x)
        final_hm, mi = out[0], out[2]
        mse = self.criterion(final_hm, target, target_weight)
        return combine_losses(mse, mi, self.w_mse, self.alpha, self.beta), final_hm

    def _fwd_bwd_reduce(self, kf_x, sup_x, target, target_weight):
        self.buckets.zero()
        loss, final_hm = self.loss(kf_x, sup_x, target, target_weight)
        loss.backward()
        self.buckets.allreduce_mean()      # waits for the hook-issued collectives (or runs them, overlap=False)
        return loss, final_hm

    def __call__(self, kf_x, sup_x, target, target_weight):
        loss, final_hm = self._fwd_bwd_reduce(kf_x, sup_x, target, target_weight)
        self.step += 1
        for p, g, m, v in zip(self.flat_p, self.buckets.flat, self.exp_avg, self.exp_avg_sq):
            _lib.call("fami_adam_step", ops._ptr(p), ops._ptr(g), ops._ptr(m), ops._ptr(v), p.numel(), float(self.lr),
                      float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step, ops._stream())
        # the update went through raw pointers: bump the version counters so cached packed weights /
        # folded BatchNorm affines (ops.packed_weight, ops.folded_affine) of the trained layers refresh
        torch.autograd.graph.increment_version(self.buckets.params)
        return loss.detach(), final_hm.detach()

    # ---- CUDA-graph form --------------------------------------------------------------------------------------------
    def capture(self, kf_x, sup_x, target, target_weight, warmup=2):
        """Captures forward + loss + backward + gradient all-reduce + Adam over the given STATIC input tensors in one
        CUDA graph (no per-launch host work on replay: the eager step issues ~2000 launches).  Learning rate and Adam
        bias corrections are read from a small device buffer refreshed before every replay, so MultiStepLR and the step
        count stay live.  Returns self; call replay() afterwards (refill the static inputs in place between replays)."""
        import gc
        dev = kf_x.device
        self._static = (kf_x, sup_x, target, target_weight)
        self._hyper = torch.zeros(3, dtype=torch.float32, device=dev)
        self._hyper_host = torch.zeros(3, dtype=torch.float32).pin_memory()
        # autograd caches each leaf's AccumulateGrad node together with the stream it was created on; nodes left over from
        # eager steps on another stream would invalidate the capture.  Drop dead graphs, then warm up AND capture on one
        # dedicated stream.
        torch.cuda.synchronize()
        if hasattr(self.model, "_last"):
            self.model._last = None
        gc.collect()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):         # allocator warm-up on the capture stream; these are real optimisation steps
                self.__call__(*self._static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gc.collect()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            loss, final_hm = self._fwd_bwd_reduce(*self._static)
            for p, gr, m, v in zip(self.flat_p, self.buckets.flat, self.exp_avg, self.exp_avg_sq):
                _lib.call("fami_adam_step_graph", ops._ptr(p), ops._ptr(gr), ops._ptr(m), ops._ptr(v), p.numel(),
                          ops._ptr(self._hyper), float(self.betas[0]), float(self.betas[1]), float(self.eps), ops._stream())
            self._out = (loss.detach(), final_hm.detach())
        self._graph = g
        return self

    def replay(self):
        self.step += 1
        h = self._hyper_host
        h[0] = float(self.lr)
        h[1] = 1.0 - self.betas[0] ** self.step
        h[2] = math.sqrt(1.0 - self.betas[1] ** self.step)
        self._hyper.copy_(h, non_blocking=True)
        self._graph.replay()
        torch.autograd.graph.increment_version(self.buckets.params)
        return self._out

    # ---- checkpoints in the reference's format (engine/defaults/checkpoints.py:45-107) ----------------------------
    def optimizer_state_dict(self):
        """A torch.optim.Adam state_dict over the trainable parameters in model.parameters() order -- what
        `optimizer.state_dict()` returns for the optimizer built by posetimation/optimizer/optimizer.py:66-72, so the
        reference's resume() (checkpoints.py:70-107) can load it and vice versa."""
        state = {}
        for i, p in enumerate(self.params):
            bi = self.buckets._bucket_of[id(p)]
            off = self._offset_of(p, bi)
            n = p.numel()
            state[i] = {"step": torch.tensor(float(self.step)),
                        "exp_avg": self.exp_avg[bi][off:off + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[bi][off:off + n].view_as(p).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(self.params)))}
        return {"state": state if self.step > 0 else {}, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("optimizer state has %d parameters, the model has %d trainable tensors"
                             % (sum(len(g["params"]) for g in groups), len(self.params)))
        g0 = groups[0]
        self.lr, self.betas, self.eps = float(g0["lr"]), tuple(g0["betas"]), float(g0["eps"])
        self.step = 0
        for i, p in enumerate(self.params):
            st = sd["state"].get(i, sd["state"].get(str(i)))
            bi = self.buckets._bucket_of[id(p)]
            off = self._offset_of(p, bi)
            n = p.numel()
            if st is None:
                self.exp_avg[bi][off:off + n].zero_()
                self.exp_avg_sq[bi][off:off + n].zero_()
                continue
            self.exp_avg[bi][off:off + n].copy_(st["exp_avg"].reshape(-1).to(p.device))
            self.exp_avg_sq[bi][off:off + n].copy_(st["exp_avg_sq"].reshape(-1).to(p.device))
            self.step = max(self.step, int(float(st["step"])))

    def _offset_of(self, p, bi):
        off = 0
        for q in self.buckets.buckets[bi]:
            if q is p:
                return off
            off += q.numel()
        raise KeyError("parameter not in bucket")

    def save_checkpoint(self, epoch, save_folder):
        """checkpoints.py:45-67: {begin_epoch, state_dict, optimizer: [adam.state_dict()]} -> epoch_{n}_state.pth."""
        os.makedirs(save_folder, exist_ok=True)
        path = osp.join(save_folder, "epoch_{}_state.pth".format(epoch))
        sd = self.model.state_dict()
        if list(sd.keys())[0].startswith("module."):
            sd = {k[7:]: v for k, v in sd.items()}
        torch.save({"begin_epoch": epoch, "state_dict": {k: v.detach().cpu() for k, v in sd.items()},
                    "optimizer": [self.optimizer_state_dict()]}, path)
        return path

    def resume(self, checkpoint_file):
        """checkpoints.py:70-107: returns begin_epoch (= saved epoch + 1).  Parameters are flat-buffer views, so the
        state is copied in place and the packed-weight / folded-BN caches are invalidated."""
        ck = torch.load(checkpoint_file, map_location="cpu")
        sd = {(k.replace("module.", "") if k.find("module") == 0 else k): v for k, v in ck["state_dict"].items()}
        sd = {(k.replace("preact.", "") if k.find("preact") == 0 else k): v for k, v in sd.items()}
        own = self.model.state_dict()
        missing = set(own) - set(sd)
        if missing or set(sd) - set(own):
            raise RuntimeError("checkpoint keys do not match the model (missing %d, unexpected %d)"
                               % (len(missing), len(set(sd) - set(own))))
        with torch.no_grad():
            for k, v in own.items():
                v.copy_(sd[k].to(v.device))          # in place: parameters stay views of the flat buffers
        self.load_optimizer_state_dict(ck["optimizer"][0])
        return ck["begin_epoch"] + 1


class AlignmentMIFunction_Term6_V1:
    """Engine-facing core function (engine/core/base.py:17-40; alignment_mi_function_term6_1.py:72-220) over TrainStep:
    the same constructor keywords the engine passes (cfg, criterion, ...), a train(model, epoch, optimizer, dataloader,
    tb_writer_dict, **kwargs) method that iterates the loader and an eval-side predict() helper -- without the
    reference's per-iteration .item()/.cpu() synchronisations (loss and accuracy stay on the device; they are read once
    per `PRINT_FREQ` iterations).  Register it with fami_pose_b200.patch_reference(); cfg.CORE_FUNCTION selects it by
    this class name, exactly as the reference resolves its own (engine/core/base.py:65)."""

    def __init__(self, cfg=None, criterion=None, **kwargs):
        self.cfg = cfg
        self.criterion = criterion
        self.output_dir = kwargs.get("output_dir")
        self.PE_Name = kwargs.get("PE_Name")
        self.max_iter_num = 0
        self.dataloader_iter = None
        self.tb_writer = None
        self.global_steps = 0
        self.alpha = getattr(getattr(cfg, "LOSS", None), "ALPHA", 0.5) if cfg is not None else 0.5
        self.beta = getattr(getattr(cfg, "LOSS", None), "BETA", 0.1) if cfg is not None else 0.1
        self.w_mse = 1.0
        if cfg is not None:
            try:
                self.w_mse = float(cfg.LOSS.HEATMAP_MSE.WEIGHT)
            except Exception:
                pass
        self.print_freq = 100
        if cfg is not None:
            try:
                self.print_freq = int(cfg.PRINT_FREQ)
            except Exception:
                pass
        self._step = None
        self.history = []          # (epoch, iteration, loss, accuracy) read back every print_freq iterations

    def _train_step(self, model, lr):
        if self._step is None or self._step.model is not model:
            m = model.module if hasattr(model, "module") else model
            self._step = TrainStep(m, lr=lr, w_mse=self.w_mse, alpha=self.alpha, beta=self.beta)
        self._step.lr = lr
        return self._step

    def train(self, model, epoch, optimizer, dataloader, tb_writer_dict, **kwargs):
        """alignment_mi_function_term6_1.py:94-217.  `optimizer` supplies the learning rate of the epoch (the engine's
        MultiStepLR steps it, trainer.py); the parameter update itself is the fused Adam of TrainStep.  Batches are
        the reference loader's (input_x, input_sup, target_heatmaps, target_heatmaps_weight, meta) tuples."""
        from .decode import accuracy
        opt = optimizer[0] if isinstance(optimizer, (list, tuple)) else optimizer
        lr = opt.param_groups[0]["lr"] if opt is not None else 1e-4
        model.train()
        step = self._train_step(model, lr)
        dev = next(model.parameters()).device
        it, last = 0, None
        acc_sum = torch.zeros((), dtype=torch.float64, device=dev)
        for batch in dataloader:
            kf_x, sup_x, tgt, tw = (t.to(dev, non_blocking=True) for t in batch[:4])
            loss, final_hm = step(kf_x, sup_x, tgt, tw)
            _, avg_acc, _, _ = accuracy(final_hm, tgt)
            acc_sum += avg_acc
            it += 1
            self.global_steps += 1
            if it % self.print_freq == 0:
                last = (epoch, it, float(loss), float(acc_sum) / it)      # the only host read-back
                self.history.append(last)
        if it and (last is None or last[1] != it):
            self.history.append((epoch, it, float(loss), float(acc_sum) / it))
        return self.history[-1] if self.history else None

    @torch.no_grad()
    def predict(self, model, kf_x, sup_x, center, scale):
        """Eval-side decode of alignment_mi_function_term6_1.py:222-328: heatmaps -> final keypoints on the device."""
        from .decode import get_final_preds
        model.eval()
        out = model(kf_x, sup_x)
        return get_final_preds(out[0], center, scale)
