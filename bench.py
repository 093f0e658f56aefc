#!/usr/bin/env python
"""bench.py -- FAMI-Pose hot-path throughput on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl fami|reference] [--batch B] [--precision fp32|bf16]

A step = one pass of the hot path over one batch of synthetic clips on each rank:
Alignment_V15 forward (HRNet-W48 on 5 frames -> global warp -> 4x modulated deformable conv -> head)
+ JointsMSE loss + keypoint argmax, at BASELINE config 2 (384x288, 5-frame window, 17 joints, 32
clips per GPU).  One process per GPU; clips shard by batch (weak scaling, no data-path collective).
Rank 0 prints ONE JSON line.
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

METRIC = "clips/sec (5-frame 384x288, 17 joints)"
UNIT = "clips/s"
H_IN, W_IN, NUM_SUP, J, WIDTH = 384, 288, 4, 17, 48


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops", 1590.0)),
                    "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                    "source": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PFLOP/s)"}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [s.strip() for s in out.strip().split(",")]
                if len(parts) >= 8:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": sorted(reasons), "samples": len(self.rows)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.  The
    reference is Python and cannot travel to the GPU box (/root/reference does not exist there), so
    this is its restatement oracle.FunctionalFami -- the same torch-CPU conv / torchvision CPU
    deform_conv2d calls the reference's modules make (pinned against the reference in tests/golden).
    Each step is a bounded sample of the B=32 workload (--ref-batch clips, default 2); --steps / --warmup are honoured."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import fami_oracle as fo
    from fami_pose_b200 import synth
    import fami_pose_b200.zoo as zoo  # only for the state_dict shapes (no CUDA needed to construct)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = synth.make_cfg(WIDTH, J)
    m = zoo.Alignment_V15(cfg, "validate")
    sd = synth.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()})
    del m
    f = fo.FunctionalFami(sd)
    Bs = args.ref_batch
    kf, sup, tgt, tw = synth.synthetic_clip(Bs)

    def step():
        with torch.no_grad():
            hm, _ = f.alignment(kf, sup)
            fo.joint_mse(hm.numpy(), tgt.numpy(), tw.numpy())
            fo.get_max_preds(hm.numpy())

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = Bs * args.steps / dt
    sample = "B=%d clips/step (bounded sample of the B=32 workload), %d steps" % (Bs, args.steps)
    emit(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE config 2: Alignment_V15 HRNet-W48 384x288, 5-frame window, 17 joints; forward "
                               "(eval-mode BN) + JointsMSE + keypoint argmax on the host CPU (torch threads = %d)" % cores,
                   "batch_per_step": Bs},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


_OUT = None


def emit(obj):
    """The ONE JSON line goes to the process's original stdout; everything else any library prints on fd 1
    (NCCL's version banner, torchrun chatter) has been diverted to stderr by main()."""
    out = _OUT if _OUT is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


ARM_DTYPE = {"fp32": "f32", "tf32": "tf32 (f32 storage, kind::tf32 tensor cores, f32 accumulate)", "fp16": "f16", "bf16": "bf16",
             "fp16s": "f16s (f16 multiplicands, f32 residual stream, f32 accumulate)"}


def main():
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fami", choices=["fami", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU")
    ap.add_argument("--precision", default=os.environ.get("FAMI_PRECISION", "tf32"), choices=["fp32", "tf32", "fp16", "fp16s", "bf16"],
                    help="the arm `value` / `e2e` / `dtype` belong to.  Default tf32: the reference computes in fp32 and its "
                         "convolutions run as TF32 on a GPU, so this is the like-for-like arm (1e-3 tier)")
    ap.add_argument("--arms", default=None,
                    help="comma-separated extra arms reported under `arms` (default: fp16 when the headline is tf32)")
    ap.add_argument("--ref-batch", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the `train` sub-record")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the same-box eager-PyTorch leg")
    ap.add_argument("--profile-step", action="store_true",
                    help="run ONE eager step between cudaProfilerStart/Stop (ncu --profile-from-start off) and exit")
    ap.add_argument("--mode", default="forward", choices=["forward", "train"],
                    help="forward (default, the headline metric; the line also carries a `train` sub-record) | train: "
                         "print the training-step line alone")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import fami_pose_b200 as fp
    from fami_pose_b200 import ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = {"torch": torch, "dist": dist, "fp": fp, "ops": ops, "synth": synth, "world": world, "rank": rank, "local": local,
           "dev": dev, "args": args}
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    try:
        if args.mode == "train":
            rec = train_record(ctx, args.steps, max(args.warmup, 3))
            if rank == 0:
                rec.update({"metric": METRIC, "mode": "train", "unit": UNIT, "n_gpus": world, "higher_is_better": True,
                            "scaling": "weak", "vs_baseline": None, "data": "synthetic", "clocks": sampler.summary()})
                emit(rec)
            return
        B = args.batch
        extra = [a for a in (args.arms.split(",") if args.arms is not None else (["fp16s", "fp16"] if args.precision == "tf32" else []))
                 if a and a != args.precision]
        fb = ForwardBench(ctx, B)
        if args.profile_step:
            return fb.profile_step(args.precision)
        head = fb.run(args.precision, args.steps, max(args.warmup, 3))
        arms = {args.precision: head}
        for a in extra:
            arms[a] = fb.run(a, args.steps, max(args.warmup, 3))
        train = None if args.no_train else train_record(ctx, min(args.steps, 10), 3)
        clocks = sampler.summary()
        sampler.stop_flag = True
        if rank == 0:
            pk = peaks()
            for a in arms:
                fp.set_precision(a)
                arms[a]["roofline"] = dcn_roofline(fp, ops, dev, fb.stream, B, pk)
                arms[a]["roofline_conv"] = conv_roofline(fp, ops, dev, fb.stream, B, pk, a)
            line = {
                "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": ARM_DTYPE[args.precision], "data": "synthetic",
                "config": {"workload": "BASELINE config 2: Alignment_V15 HRNet-W48 384x288, 5-frame window, 17 joints, "
                                       "batch 32 per GPU; forward (eval-mode BN) + JointsMSE + keypoint argmax",
                           "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d (clips shard by batch, no "
                           "data-path collective in forward; the `train` record carries the gradient all-reduce)" % world,
                           "cuda_graph": head["cuda_graph"], "e2e_overlap": "H2D of step k+1 double-buffered against compute of step k",
                           "l2": "working set (inputs 106 MB + activations > 2 GB per step) exceeds the 126 MB L2; no flush",
                           "arm": "%s: %s" % (args.precision, ARM_NOTE[args.precision])},
                "e2e": head["e2e"], "gpu_launches": head["launches_per_step"] * args.steps,
                "launches_per_step": head["launches_per_step"], "clocks": clocks,
                "roofline": arms[args.precision]["roofline"], "roofline_conv": arms[args.precision]["roofline_conv"],
                "peaks": pk["source"],
                # every measured precision arm of the SAME workload (value = device-resident inputs, e2e = host buffers);
                # the headline keys above repeat arms[--precision]
                "arms": {ARM_DTYPE[a].split(" ")[0]: arms[a] for a in arms},
            }
            if train is not None:
                line["train"] = train
            if not args.no_reference_gpu and world == 1:
                line["reference_gpu"] = reference_gpu(ctx, fb.sd)
            if not args.no_cpu_baseline and world == 1:
                from oracle import fami_oracle as fo   # cpu_baseline leg only
                line["cpu_baseline"] = cpu_baseline(fo, fb.sd)
            emit(line)
    finally:
        sampler.stop_flag = True
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()


ARM_NOTE = {
    "fp32": "exact-fp32 SIMT kernels",
    "tf32": "fp32 storage (activations, residual stream, offsets, heatmaps); every convolution on tcgen05.mma.kind::tf32, "
            "multiplicands rounded to nearest TF32 (weights at pack time, activations by the TMA load), fp32 accumulate -- "
            "the arithmetic cuDNN applies to the reference's fp32 convs on a GPU; parity tier 1e-3",
    "fp16": "fp16 activations, tcgen05.mma.kind::f16, fp32 accumulate, fp32 offsets / masks / heatmaps; parity tier 1e-2",
    "bf16": "bf16 activations, tcgen05.mma.kind::f16, fp32 accumulate, fp32 offsets / masks / heatmaps; parity tier 1e-2",
    "fp16s": "fp16 multiplicands (the 11-bit significand of TF32) on tcgen05.mma.kind::f16, every residual / running sum stored "
             "and added in fp32 (fp32 twin of each block output), fp32 accumulate, fp32 offsets / masks / heatmaps; measured "
             "5.5e-4 / 8.3e-4 against the reference (the tf32 arm's error), tested at 1e-3",
}


class ForwardBench:
    """Forward step (Alignment_V15 eval forward + JointsMSE + keypoint argmax) at config 2 on one rank: device-resident
    timing (`value`) and end-to-end timing through the public API with pinned HOST buffers (`e2e`)."""

    def __init__(self, ctx, B):
        torch, fp, synth = ctx["torch"], ctx["fp"], ctx["synth"]
        self.ctx, self.B = ctx, B
        dev, rank = ctx["dev"], ctx["rank"]
        cfg = synth.make_cfg(WIDTH, J)
        model = fp.Alignment_V15(cfg, "validate")
        self.sd = synth.seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
        model.load_state_dict(self.sd)
        self.model = model.to(dev).eval()
        self.loss_fn = fp.JointMSELoss()
        # synthetic clips: per-rank seed = base + rank; pinned host copies for the e2e leg
        self.host = tuple(t.pin_memory() for t in synth.synthetic_clip(B, seed=synth.SEED + rank))
        d0 = tuple(t.to(dev) for t in self.host)
        # two device input sets (A/B): the e2e leg double-buffers the host->device copies against compute
        self.sets = [d0, tuple(t.clone() for t in d0)]
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.host)
        self.out_host = torch.empty((B, J), dtype=torch.int32).pin_memory()
        self.loss_host = torch.empty((), dtype=torch.float32).pin_memory()
        self.d2h_bytes = self.out_host.numel() * 4 + 4
        self.stream = torch.cuda.Stream(device=dev)

    def step_fn(self, kf, sup, tgt, tw):
        fp = self.ctx["fp"]
        hm, kfhm = self.model(kf, sup)
        loss = self.loss_fn(hm, tgt, tw)
        idx = fp.argmax_indices(hm)
        return loss, idx

    def profile_step(self, precision):
        torch, fp = self.ctx["torch"], self.ctx["fp"]
        fp.set_precision(precision)
        with torch.no_grad(), torch.cuda.stream(self.stream):
            for _ in range(2):
                self.step_fn(*self.sets[0])
            self.stream.synchronize()
            torch.cuda.profiler.start()
            self.step_fn(*self.sets[0])
            self.stream.synchronize()
            torch.cuda.profiler.stop()

    def run(self, precision, steps, warmup):
        ctx = self.ctx
        torch, dist, fp, world, dev, args = ctx["torch"], ctx["dist"], ctx["fp"], ctx["world"], ctx["dev"], ctx["args"]
        fp.set_precision(precision)
        stream, sets, B = self.stream, self.sets, self.B
        graphs, outs = [None, None], [None, None]
        with torch.no_grad(), torch.cuda.stream(stream):
            self.step_fn(*sets[0])           # builds packed-weight / folded-BN caches
            stream.synchronize()
            l0 = fp._lib.launch_count()
            self.step_fn(*sets[0])
            launches_per_step = fp._lib.launch_count() - l0
            if not args.no_graph:
                pool = None
                for i in range(2):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=stream, pool=pool):
                        outs[i] = self.step_fn(*sets[i])
                    pool = g.pool()
                    graphs[i] = g
        stream.synchronize()

        def run_step(i=0):
            if graphs[i] is not None:
                graphs[i].replay()
                return outs[i]
            with torch.no_grad():
                return self.step_fn(*sets[i])

        ms = timed_region(ctx, stream, lambda k: run_step(0), steps, warmup)

        # e2e: public API with HOST buffers.  Every step copies its inputs (227 MB) from pinned host memory and
        # reads its results (loss + keypoint indices) back; the copy of step k+1 runs on a second stream while
        # step k computes (double-buffered device inputs), as a real input pipeline would.
        copy_stream = torch.cuda.Stream(device=dev)
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]

        def issue_copy(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[i])
                for t_dst, t_src in zip(sets[i], self.host):
                    t_dst.copy_(t_src, non_blocking=True)
                ready[i].record(copy_stream)

        state = {"primed": False}

        def e2e_step(k):
            i = k & 1
            if not state["primed"]:
                freed[0].record(stream)
                freed[1].record(stream)
                issue_copy(i)
                state["primed"] = True
            stream.wait_event(ready[i])
            issue_copy(i ^ 1)                       # next step's inputs, overlapping this step's compute
            loss, idx = run_step(i)
            freed[i].record(stream)
            self.out_host.copy_(idx, non_blocking=True)
            self.loss_host.copy_(loss, non_blocking=True)

        ms_e2e = timed_region(ctx, stream, e2e_step, steps, warmup)
        copy_stream.synchronize()
        del graphs, outs
        clips = B * world * steps
        return {"dtype": ARM_DTYPE[precision], "value": clips / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms / steps,
                "e2e": {"value": clips / (ms_e2e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": self.h2d_bytes,
                        "d2h_bytes_per_step": self.d2h_bytes, "ms_per_step": ms_e2e / steps},
                "launches_per_step": launches_per_step, "cuda_graph": not args.no_graph}


def timed_region(ctx, stream, fn, steps, warmup):
    """W untimed steps, then exactly K steps bracketed by barrier + synchronize; device time by CUDA events on the
    launching stream, MAX over ranks."""
    torch, dist, world, dev = ctx["torch"], ctx["dist"], ctx["world"], ctx["dev"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for k in range(warmup):
            fn(k)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(steps):
            fn(warmup + k)
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


def train_record(ctx, steps, warmup):
    """Training step at config 2 / 3 shape (alignment_mi_function_term6_1.py:104-156 under trainer.py:57-58): one process
    per GPU, per-replica BatchNorm, forward (train-mode BN) + JointsMSE + six MI terms + backward + ONE bucketed NCCL
    all-reduce of the trainable gradients INSIDE the timed region + fused Adam.  HRNet frozen (reference default)."""
    torch, dist, fp, synth = ctx["torch"], ctx["dist"], ctx["fp"], ctx["synth"]
    world, rank, dev, args = ctx["world"], ctx["rank"], ctx["dev"], ctx["args"]
    from fami_pose_b200.train import TrainStep
    B = args.batch
    # the frozen backbone runs on the headline arm (tf32 by default: the configuration tests/test_gpu_train.py pins against
    # the reference's float64 autograd); FAMI_TRAIN_BACKBONE=fp16 puts it on the 16-bit arm
    backbone = os.environ.get("FAMI_TRAIN_BACKBONE", "tf32" if args.precision == "fp32" else args.precision)
    head = os.environ.get("FAMI_TRAIN_HEAD", "tf32")      # 'tf32': tensor-core forward + dgrad of the trainable head; 'fp32': SIMT
    fp.set_precision(head)
    model = fp.Alignment_V15(synth.make_cfg(WIDTH, J), "train")
    sd = synth.seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
    model.load_state_dict(sd)
    model = model.to(dev).train()
    if backbone != head:
        model.backbone_precision = backbone
    host = tuple(t.pin_memory() for t in synth.synthetic_clip(B, seed=synth.SEED + rank))
    devs = tuple(t.to(dev) for t in host)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    step = TrainStep(model)
    n_train = sum(p.numel() for p in step.buckets.params)
    stream = torch.cuda.current_stream()
    l0 = fp._lib.launch_count()
    step(*devs)
    torch.cuda.synchronize()
    launches = fp._lib.launch_count() - l0
    graphed, graph_note = False, ""
    if not args.no_graph and os.environ.get("FAMI_TRAIN_GRAPH", "1") != "0":
        try:      # the whole step (forward, loss, backward, gradient all-reduce, Adam) as ONE CUDA graph
            step.capture(*devs)
            graphed = True
        except Exception as e:   # e.g. a collective that cannot be captured on this NCCL build: time the eager step
            graph_note = "capture failed, eager step timed: " + repr(e)[:160]
            torch.cuda.synchronize()
    run = (lambda: step.replay()) if graphed else (lambda: step(*devs))

    def e2e(_k):
        for d_, h_ in zip(devs, host):
            d_.copy_(h_, non_blocking=True)
        loss, _ = run()
        loss_host.copy_(loss, non_blocking=True)

    ms = timed_region(ctx, stream, lambda k: run(), steps, warmup)
    ms_e2e = timed_region(ctx, stream, e2e, steps, 1)
    clips = B * world * steps
    rec = {"value": clips / (ms / 1000.0), "unit": UNIT, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
           "dtype": "%s head (f32 storage) / %s frozen backbone" % (head, backbone),
           "config": {"workload": "BASELINE config 2/3 TRAIN step: forward (train-mode BN) + JointsMSE + MI losses + backward of "
                                  "the head (HRNet frozen, reference default) + gradient all-reduce + Adam; batch 32 per GPU",
                      "batch_per_gpu": B, "global_batch": B * world,
                      "collective": "one bucketed NCCL all-reduce of %d fp32 gradients per step over %d rank(s), inside the "
                                    "timed region" % (n_train, world),
                      "cuda_graph": graphed, "note": graph_note},
           "e2e": {"value": clips / (ms_e2e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                   "ms_per_step": ms_e2e / steps},
           "gpu_launches": launches * steps, "launches_per_step": launches}
    del step, model
    torch.cuda.empty_cache()
    return rec


def reference_gpu(ctx, sd):
    """Same-box GPU baseline of the REFERENCE computation: the oracle port -- the torch / torchvision
    calls the reference's modules make, pinned to the reference at 0.0 max-abs -- run eagerly on cuda in fp32, with cuDNN
    TF32 on (PyTorch's default for convolutions) and off.  A reported bar, not part of the product path."""
    torch, synth, dev = ctx["torch"], ctx["synth"], ctx["dev"]
    from oracle import fami_oracle as fo
    Bs = 8
    out = {"batch": Bs, "unit": UNIT, "what": "oracle port (eager PyTorch: cuDNN convs, torchvision deform_conv2d) on the same "
           "B200, forward + JointsMSE + argmax on device-resident inputs, B=%d per step" % Bs}
    try:
        sdd = {k: v.to(dev) for k, v in sd.items()}
        f = fo.FunctionalFami(sdd)
        kf, sup, tgt, tw = (t.to(dev) for t in synth.synthetic_clip(Bs))
        prev = torch.backends.cudnn.allow_tf32
        for name, flag in (("cudnn_tf32", True), ("strict_fp32", False)):
            torch.backends.cudnn.allow_tf32 = flag
            with torch.no_grad():
                for _ in range(2):
                    hm, _ = f.alignment(kf, sup)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 3
                e0.record()
                for _ in range(n):
                    hm, _ = f.alignment(kf, sup)
                    loss = ((hm - tgt) * tw.view(Bs, J, 1, 1)).pow(2).mean() * 0.5
                    idx = hm.reshape(Bs, J, -1).argmax(2)
                e1.record()
                torch.cuda.synchronize()
            out[name] = {"value": Bs * n / (e0.elapsed_time(e1) / 1000.0), "ms_per_step": e0.elapsed_time(e1) / n}
        torch.backends.cudnn.allow_tf32 = prev
        del f, sdd
        torch.cuda.empty_cache()
    except Exception as e:   # a missing torchvision CUDA op must not take the bench line down
        out["error"] = repr(e)[:200]
    return out


def dcn_roofline(fp, ops, dev, stream, B, pk):
    """The north-star kernel timed alone at the config-2 shape (x [B,48,96,72], G=12) in the active
    precision.  Algorithmic bytes: B*H*W*(s_x*(Cin+Cout) + 4*27*G) + weights, with
    s_x = 4 (fp32 arm) or 2 (16-bit arm; offsets and masks are fp32 in both)."""
    import torch
    C, G, H, W = 48, 12, 96, 72
    dt = ops.act_dtype()
    half = ops.dcn_fused_supported(C, G, dt)      # 16-bit arms, and the tf32 arm (fp32 x cast to fp16 on the way in, fp32 out)
    g = torch.Generator(device="cpu").manual_seed(1)
    x = ops.to_nhwc(torch.randn(B, C, H, W, generator=g).to(dev), dt)
    off = (2 * torch.randn(B, 18 * G, H, W, generator=g)).to(dev)     # sigma = 2 px
    msk = torch.randn(B, 9 * G, H, W, generator=g).to(dev)
    dcn = fp.DeformConv2d(C, C, 3, padding=3, dilation=3).to(dev)
    out = ops.empty_nhwc(B, C, H, W, dt, dev)
    if half:   # fused [offset|mask] buffer in the row-blocked layout, as the alignment head's producer conv writes it
        om = ops.to_nhwc(torch.cat([off, msk], 1)[:, ops.tap_major_perm(G)].contiguous(), torch.float32)
        blk = ops.om_to_blocked(om, G, layout=ops.dcn_blocked_layout(C, C, G))
        del om
        run = lambda: dcn(x, None, None, out=out, blocked_om=blk, groups=G)
    else:
        offn, mskn = ops.to_nhwc(off, torch.float32), ops.to_nhwc(msk, torch.float32)
        run = lambda: dcn(x, offn, mskn, out=out)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times = []
    with torch.no_grad(), torch.cuda.stream(stream):
        for i in range(13):
            flush.zero_()  # evict L2 between launches (inputs alone exceed L2 at B=32, flushed anyway)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run()
            e1.record(stream)
            stream.synchronize()
            if i >= 3:
                times.append(e0.elapsed_time(e1))
    t = sorted(times)[len(times) // 2] / 1000.0
    sx = dt.itemsize        # storage of x and out as the caller holds them (2: fp16 / bf16 arms; 4: fp32 and tf32 arms)
    alg = B * H * W * (sx * (C + C) + 4 * 27 * G) + sx * 9 * C * C + 4 * C
    ach = alg / t / 1e9
    traffic = None
    try:   # DRAM bytes of the same launch from the committed ncu --set full capture (16-bit arm, B=32)
        if half and B == 32:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))["dcn_wp_kernel_fp16_B32"]
    except Exception:
        traffic = None
    return {"kernel": "fami_dcn_fwd (modulated deformable conv, C=48 G=12 96x72 B=%d, x/out %s, offsets fp32%s)"
                      % (B, str(dt).replace("torch.", ""), "; fp32 x is cast to fp16 by one elementwise launch inside the timed call"
                         if (half and sx == 4) else ""),
            "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
            "traffic": traffic, "algorithmic_bytes": alg, "us_per_launch": t * 1e6}


def conv_roofline(fp, ops, dev, stream, B, pk, precision):
    """Dominant kernel class by time: HRNet stage-4 3x3 s1 conv (48->48 @96x72, N=5B images)."""
    import torch
    N, C, H, W = 5 * B, 48, 96, 72
    dt = ops.act_dtype()
    x = ops.empty_nhwc(N, C, H, W, dt, dev).normal_()
    conv = torch.nn.Conv2d(C, C, 3, 1, 1, bias=False).to(dev)
    bn = torch.nn.BatchNorm2d(C).to(dev).eval()
    out = ops.empty_nhwc(N, C, H, W, dt, dev)
    res = ops.empty_nhwc(N, C, H, W, dt, dev).normal_()     # a separate residual tensor (no aliasing with x)
    times = []
    with torch.no_grad(), torch.cuda.stream(stream):
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ops.conv_bn_act(x, conv, bn, relu=True, residual=res, out=out)
            e1.record(stream)
            stream.synchronize()
            if i >= 3:
                times.append(e0.elapsed_time(e1))
    t = sorted(times)[len(times) // 2] / 1000.0
    flops = 2.0 * N * H * W * 9 * C * C
    ach = flops / t / 1e12
    return {"kernel": "fami_conv2d_bn_act_fwd 3x3 s1 48->48 @96x72 N=%d (%s)" % (N, precision), "bound": "tensor",
            "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
            "traffic": None, "flops": flops, "us_per_launch": t * 1e6}


def cpu_baseline(fo, sd):
    """Oracle port (torch-CPU restatement of the reference forward) on the host cores, bounded sample."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    f = fo.FunctionalFami(sd)
    Bs = 2
    kf, sup, tgt, tw = fo.synthetic_clip(Bs)
    with torch.no_grad():
        f.alignment(kf, sup)
        t0 = time.perf_counter()
        n = 2
        for _ in range(n):
            hm, _ = f.alignment(kf, sup)
            fo.joint_mse(hm.numpy(), tgt.numpy(), tw.numpy())
            fo.get_max_preds(hm.numpy())
        dt = time.perf_counter() - t0
    return {"value": Bs * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "B=2 clips x %d forward passes of the same workload (fp32, torch CPU threads=%d)" % (n, cores)}


if __name__ == "__main__":
    main()
