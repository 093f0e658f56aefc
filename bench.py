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
    deform_conv2d calls the reference's modules make (pinned against the reference in tests/golden)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import fami_oracle as fo
    from oracle import ref_harness as rh
    import fami_pose_b200.zoo as zoo  # only for the state_dict shapes (no CUDA needed to construct)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = rh.make_cfg(WIDTH, J)
    m = zoo.Alignment_V15(cfg, "validate")
    sd = fo.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()})
    del m
    f = fo.FunctionalFami(sd)
    Bs = args.ref_batch
    kf, sup, tgt, tw = fo.synthetic_clip(Bs)

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
        "config": {"workload": "Alignment_V15 HRNet-W48 384x288 5-frame 17-joint forward+JointsMSE+argmax, CPU",
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
    ap.add_argument("--precision", default=os.environ.get("FAMI_PRECISION", "fp16"), choices=["fp32", "tf32", "fp16", "bf16"])
    ap.add_argument("--ref-batch", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-step", action="store_true",
                    help="run ONE eager step between cudaProfilerStart/Stop (ncu --profile-from-start off) and exit")
    ap.add_argument("--mode", default="forward", choices=["forward", "train"],
                    help="forward (default, the headline metric) | train: fwd + loss + bwd + gradient all-reduce + Adam "
                         "with the backbone frozen (reference default), fp32 head, --precision arm for the backbone")
    args = ap.parse_args()
    if args.mode == "train" and args.impl == "fami":
        return run_train(args)
    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5
        args.warmup = min(args.warmup, 1)
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import fami_pose_b200 as fp
    from fami_pose_b200 import ops
    from oracle import fami_oracle as fo   # synthetic inputs + seeded weights + cpu_baseline leg only
    from oracle import ref_harness as rh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fp.set_precision(args.precision)
    B = args.batch

    cfg = rh.make_cfg(WIDTH, J)
    model = fp.Alignment_V15(cfg, "validate")
    sd = fo.seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    loss_fn = fp.JointMSELoss()

    # synthetic clips: per-rank seed = base + rank; pinned host copies for the e2e leg
    kf_h, sup_h, tgt_h, tw_h = fo.synthetic_clip(B, seed=19970808 + rank)
    kf_h, sup_h, tgt_h, tw_h = (t.pin_memory() for t in (kf_h, sup_h, tgt_h, tw_h))
    kf_d, sup_d, tgt_d, tw_d = (t.to(dev) for t in (kf_h, sup_h, tgt_h, tw_h))
    h2d_bytes = sum(t.numel() * t.element_size() for t in (kf_h, sup_h, tgt_h, tw_h))
    out_host = torch.empty((B, J), dtype=torch.int32).pin_memory()
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    d2h_bytes = out_host.numel() * 4 + 4

    def step_fn(kf, sup, tgt, tw):
        hm, kfhm = model(kf, sup)
        loss = loss_fn(hm, tgt, tw)
        idx = fp.argmax_indices(hm)
        return loss, idx

    stream = torch.cuda.Stream(device=dev)
    if args.profile_step:
        with torch.no_grad(), torch.cuda.stream(stream):
            for _ in range(2):
                step_fn(kf_d, sup_d, tgt_d, tw_d)
            stream.synchronize()
            torch.cuda.profiler.start()
            step_fn(kf_d, sup_d, tgt_d, tw_d)
            stream.synchronize()
            torch.cuda.profiler.stop()
        return
    # two device input sets (A/B): the e2e leg double-buffers the host->device copies against compute
    sets = [(kf_d, sup_d, tgt_d, tw_d), tuple(torch.empty_like(t) for t in (kf_d, sup_d, tgt_d, tw_d))]
    for t_src, t_dst in zip(sets[0], sets[1]):
        t_dst.copy_(t_src)
    graphs, outs = [None, None], [None, None]
    with torch.no_grad(), torch.cuda.stream(stream):
        step_fn(*sets[0])           # builds packed-weight / folded-BN caches
        stream.synchronize()
        l0 = fp._lib.launch_count()
        step_fn(*sets[0])
        launches_per_step = fp._lib.launch_count() - l0
        if not args.no_graph:
            pool = None
            for i in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream, pool=pool):
                    outs[i] = step_fn(*sets[i])
                pool = g.pool()
                graphs[i] = g
    stream.synchronize()

    def run_step(i=0):
        if graphs[i] is not None:
            graphs[i].replay()
            return outs[i]
        with torch.no_grad():
            return step_fn(*sets[i])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
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

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda k: run_step(0), args.steps, max(args.warmup, 3))

    # e2e: public API with HOST buffers.  Every step copies its inputs (227 MB) from pinned host memory and
    # reads its results (loss + keypoint indices) back; the copy of step k+1 runs on a second stream while
    # step k computes (double-buffered device inputs), as a real input pipeline would.
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    host = (kf_h, sup_h, tgt_h, tw_h)

    def issue_copy(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[i])
            for t_dst, t_src in zip(sets[i], host):
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
        out_host.copy_(idx, non_blocking=True)
        loss_host.copy_(loss, non_blocking=True)

    ms_e2e = timed(e2e_step, args.steps, max(args.warmup, 3))
    copy_stream.synchronize()
    sampler.stop_flag = True

    clips = B * world * args.steps
    value = clips / (ms / 1000.0)
    e2e_value = clips / (ms_e2e / 1000.0)

    if rank == 0:
        pk = peaks()
        roof = dcn_roofline(fp, ops, dev, stream, B, pk)
        conv_roof = conv_roofline(fp, ops, dev, stream, B, pk, args.precision)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32": "tf32 (f32 storage)", "fp16": "f16", "bf16": "bf16"}[args.precision], "data": "synthetic",
            "config": {"workload": "BASELINE config 2: Alignment_V15 HRNet-W48 384x288, 5-frame window, 17 joints, "
                                   "batch 32 per GPU; forward (eval-mode BN) + JointsMSE + keypoint argmax",
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d (clips shard by batch)" % world,
                       "cuda_graph": graphs[0] is not None, "e2e_overlap": "H2D of step k+1 double-buffered against compute of step k",
                       "l2": "working set (inputs 106 MB + activations > 2 GB per step) exceeds the 126 MB L2; no flush"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "clocks": sampler.summary(),
            "roofline": roof, "roofline_conv": conv_roof, "peaks": pk["source"],
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(fo, sd)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train(args):
    """Training step at config 2 / 3 shape: one process per GPU, per-replica BatchNorm, ONE bucketed NCCL
    all-reduce of the 1.06 M trainable gradients per step, fused Adam.  Eager launches."""
    import torch
    import torch.distributed as dist
    import fami_pose_b200 as fp
    from fami_pose_b200.train import TrainStep
    from oracle import fami_oracle as fo   # synthetic inputs + seeded weights only
    from oracle import ref_harness as rh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fp.set_precision("fp32")
    B = args.batch
    model = fp.Alignment_V15(rh.make_cfg(WIDTH, J), "train")
    sd = fo.seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
    model.load_state_dict(sd)
    model = model.to(dev).train()
    if args.precision != "fp32":
        model.backbone_precision = args.precision
    host = tuple(t.pin_memory() for t in fo.synthetic_clip(B, seed=19970808 + rank))
    devs = tuple(t.to(dev) for t in host)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    step = TrainStep(model)
    n_train = sum(p.numel() for p in step.buckets.params)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    def e2e():
        for d_, h_ in zip(devs, host):
            d_.copy_(h_, non_blocking=True)
        loss, _ = step(*devs)
        loss_host.copy_(loss, non_blocking=True)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = fp._lib.launch_count()
    step(*devs)
    torch.cuda.synchronize()
    launches = fp._lib.launch_count() - l0
    ms = timed(lambda: step(*devs), args.steps, max(args.warmup, 3))
    ms_e2e = timed(e2e, args.steps, 1)
    sampler.stop_flag = True
    clips = B * world * args.steps
    if rank == 0:
        emit(({
            "metric": METRIC, "mode": "train", "value": clips / (ms / 1000.0), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 head / %s backbone" % args.precision, "data": "synthetic",
            "config": {"workload": "BASELINE config 2/3: Alignment_V15 HRNet-W48 384x288, 5-frame window, 17 joints, batch 32 "
                                   "per GPU; TRAIN step = forward (train-mode BN) + JointsMSE + MI losses + backward of the "
                                   "head (HRNet frozen, reference default) + gradient all-reduce + Adam",
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d, one all-reduce of %d fp32 "
                       "gradients per step" % (world, n_train), "cuda_graph": False,
                       "l2": "working set exceeds the 126 MB L2; no flush"},
            "e2e": {"value": clips / (ms_e2e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches * args.steps, "launches_per_step": launches, "clocks": sampler.summary()}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def dcn_roofline(fp, ops, dev, stream, B, pk):
    """The north-star kernel timed alone at the config-2 shape (x [B,48,96,72], G=12) in the active
    precision.  Algorithmic bytes: B*H*W*(s_x*(Cin+Cout) + 4*27*G) + weights, with
    s_x = 4 (fp32 arm) or 2 (16-bit arm; offsets and masks are fp32 in both)."""
    import torch
    C, G, H, W = 48, 12, 96, 72
    dt = ops.act_dtype()
    half = dt != torch.float32
    g = torch.Generator(device="cpu").manual_seed(1)
    x = ops.to_nhwc(torch.randn(B, C, H, W, generator=g).to(dev), dt)
    off = (2 * torch.randn(B, 18 * G, H, W, generator=g)).to(dev)     # sigma = 2 px
    msk = torch.randn(B, 9 * G, H, W, generator=g).to(dev)
    dcn = fp.DeformConv2d(C, C, 3, padding=3, dilation=3).to(dev)
    out = ops.empty_nhwc(B, C, H, W, dt, dev)
    if half:   # fused [offset|mask] buffer in the warp-blocked layout, as the alignment head's producer conv writes it
        om = ops.to_nhwc(torch.cat([off, msk], 1)[:, ops.tap_major_perm(G)].contiguous(), torch.float32)
        blk = ops.om_to_blocked(om, G)
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
    sx = 2 if half else 4
    alg = B * H * W * (sx * (C + C) + 4 * 27 * G) + sx * 9 * C * C + 4 * C
    ach = alg / t / 1e9
    traffic = None
    try:   # DRAM bytes of the same launch from the committed ncu --set full capture (16-bit arm, B=32)
        if half and B == 32:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["dcn_tc_kernel_fp16_B32"]
    except Exception:
        traffic = None
    return {"kernel": "fami_dcn_fwd (modulated deformable conv, C=48 G=12 96x72 B=%d, x/out %s, offsets fp32)" % (B, str(dt).replace("torch.", "")),
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
    times = []
    with torch.no_grad(), torch.cuda.stream(stream):
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ops.conv_bn_act(x, conv, bn, relu=True, residual=x, out=out)
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
