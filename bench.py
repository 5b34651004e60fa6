#!/usr/bin/env python
"""LRCN train clips/s on B200 (BASELINE.json metric) -- see DESIGN.md section "Measurement".

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (hand-written sm_100a kernels)
  python bench.py --impl reference ...                           the reference path on the host CPU cores

Workload (BASELINE.json configs[1]): medsos LRCN classifier -- frozen ResNet-50 frame encoder in
train-mode BN, 3x(Linear+GELU+LayerNorm) adapts, 3-layer LSTM H=32 over T frames, LN/GELU head,
4 classes -- 16 frames x 112x112 RGB, 64 clips per GPU, full train step
(zero_grad, forward, CrossEntropy, backward, Adam).  Synthetic clips, random-init weights.

One JSON line on stdout (rank 0).  `value` = clips/s with the float32 clips resident in HBM;
`e2e` = the same step through the public host API: pinned uint8 host clips -> H2D -> ingest kernel ->
train step -> loss read back, every step.  N > 1: one process per GPU (torchrun), clips sharded by
rank, gradient buckets all-reduced over NCCL, time = max over ranks."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="medsos-lrcn-resnet50", frames=16, size=112, clips_per_gpu=64, num_classes=4, hidden=32,
                rnn_input=8, rnn_layers=3)
METRIC = "lrcn_train_clips_per_sec"
RESNET50_GFLOP_PER_FRAME_112 = 2.152     # SURVEY.md section 8(d), torch.utils.flop_counter, fwd


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-clips", type=int, default=0, help="clips per step of the CPU arms (0 = the workload's 64)")
    ap.add_argument("--cpu-budget", type=float, default=280.0, help="wall-clock bound (s) of the --impl reference run")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the stock-PyTorch-on-this-GPU yardstick")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs only)")
    ap.add_argument("--no-graph", action="store_true", help="launch the encoder kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="run the frozen encoder pass and the trainable tail of each step back to back on one stream")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops_sustained"], d["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernels, from the committed ncu summary of this build
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from an ncu capture of `bench.py`); {} when absent."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (one streaming
    `nvidia-smi -lms 50` process; only samples taken between start() and summary() count)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows, self.recording = [], False
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def run(self):
        if self.proc is None:
            return
        for line in self.proc.stdout:
            if self.recording:
                self.rows.append([c.strip() for c in line.strip().split(",")])

    def begin(self):
        self.recording = True

    def summary(self):
        self.recording = False
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arms: the reference's OWN LRCN class (medsos_lrcn/src/models.py:121-234, staged unmodified into
# oracle/_ref by oracle/build_ref.py) driven by the inner body of its train loop (train_eval.py:20-43)
# ------------------------------------------------------------------------------------------------

def reference_model_and_kind():
    """(model, kind): the reference class when its sources are reachable (kind "reference"), else None."""
    import torch
    from oracle import refload
    W = WORKLOAD
    if not refload.available():
        return None, "port"
    mm = refload.medsos_models(CONF_RNN_LAYER=W["rnn_layers"], CONF_RNN_OUT="all", CONF_CLASSIF_MODE="multiclass",
                               CONF_DROPOUT=0.25, CONF_BIDIR=False, CONF_RNN_TYPE="lstm")
    torch.manual_seed(0)
    m = mm.LRCN(W["num_classes"], W["frames"], W["hidden"], W["rnn_input"], cnn_backbone="resnet50", rnn_type="lstm",
                rnn_out="all", bidirectional=False)       # random init: no network for the ImageNet weights
    return m, "reference"


def reference_train_step(model, opt, x, y):
    """train_eval.py:20-43 without the per-step `.item()` host syncs (they only feed the epoch print)."""
    import torch
    opt.zero_grad()
    out = model(x)
    loss = torch.nn.functional.cross_entropy(out, y)        # nn.CrossEntropyLoss(), main.py:147
    _, pred = torch.max(out, 1)
    loss.backward()
    opt.step()
    return loss, pred


def cpu_reference_step_rate(clips, steps, warmup, budget_s=None):
    """The reference train step on the host cores, fp32, every host thread, `clips` clips per step.
    Returns (clips/s, median s/step, cores, kind, steps actually timed)."""
    import torch
    W = WORKLOAD
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, kind = reference_model_and_kind()
    g = torch.Generator().manual_seed(1234)
    x = torch.randint(0, 256, (clips, W["frames"], 3, W["size"], W["size"]), generator=g).float() / 255.0
    y = torch.randint(0, W["num_classes"], (clips,), generator=g)
    if model is not None:
        model.train()                                            # train_eval.py:12
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)      # main.py:151
        one = lambda: reference_train_step(model, opt, x, y)
    else:                                                        # no staged reference: the oracle port
        import video_classif_b200 as vc
        from oracle import lrcn_oracle as O
        torch.manual_seed(0)
        shell = vc.LRCN(W["num_classes"], W["frames"], W["hidden"], W["rnn_input"], cnn_backbone="resnet50",
                        rnn_layers=W["rnn_layers"], dropout=0.0)
        sd = {k: v.detach().clone() for k, v in shell.state_dict().items()}
        keys = [k for k, p in shell.named_parameters() if p.requires_grad]
        for k in keys:
            sd[k].requires_grad_(True)
        opt = torch.optim.Adam([sd[k] for k in keys], lr=1e-4)

        def one():
            opt.zero_grad()
            logits, ns = O.medsos_lrcn_forward(sd, x, "resnet50", W["hidden"], W["rnn_layers"], False)
            O.cross_entropy_mean(logits, y).backward()
            opt.step()
            sd.update(ns)
    times = []
    t_start = time.perf_counter()
    done = 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            done += 1
        # bounded: never let the CPU arm run past its budget (the line reports the steps actually timed)
        if budget_s is not None and done >= 1 and (time.perf_counter() - t_start) + dt > budget_s:
            break
    times.sort()
    med = times[len(times) // 2]
    return clips / med, med, cores, kind, done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = WORKLOAD
    clips = args.cpu_clips or W["clips_per_gpu"]
    rate, sec, cores, kind, done = cpu_reference_step_rate(clips, max(1, args.steps), max(0, args.warmup),
                                                           budget_s=args.cpu_budget)
    what = ("the reference's own LRCN class (medsos_lrcn/src/models.py:121-234, staged unmodified in oracle/_ref) + "
            "train_eval.py:20-43 step" if kind == "reference" else "oracle port of the reference train step")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{W['name']}: {W['frames']} frames x {W['size']}x{W['size']}, {clips} clips/step, frozen "
                               f"ResNet-50 (train-mode BN) + GELU/LN adapts + {W['rnn_layers']}-layer LSTM H={W['hidden']} + head, "
                               "full train step (fwd, CE, bwd, Adam); BASELINE.json configs[1]",
                   "global_batch": clips, "engine": "torch CPU fp32 kernels, all host threads"},
        "cpu_baseline": {"value": rate, "unit": "clips/s", "cores": cores, "kind": kind,
                         "sample": f"{clips} clips/step x {done} timed steps (median), {what}"},
        "e2e": {"value": rate, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def torch_gpu_baseline(dev, steps=8, warmup=3):
    """Second yardstick (SURVEY.md:100): the reference's own LRCN class executed by STOCK PyTorch
    (cuDNN / cuBLASLt) on the same B200, same shape, same train step, inputs resident in HBM -- fp32 as the reference
    runs it, and bf16 autocast + channels_last as a tuned user would.  Returns a dict for the bench line."""
    import torch
    W = WORKLOAD
    B, T, S = W["clips_per_gpu"], W["frames"], W["size"]
    model, kind = reference_model_and_kind()
    if model is None:
        return {"unavailable": "reference sources not staged (oracle/_ref missing)"}
    model = model.to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(1234)
    xs = [(torch.randint(0, 256, (B, T, 3, S, S), generator=g).float() / 255.0).to(dev) for _ in range(2)]
    ys = [torch.randint(0, W["num_classes"], (B,), generator=g).to(dev) for _ in range(2)]
    out = {"what": "reference LRCN class (oracle/_ref) on stock PyTorch " + torch.__version__ + f", cuDNN {torch.backends.cudnn.version()}, "
                   f"same GPU, {B} clips x {T} x {S}x{S}, train step, inputs resident", "unit": "clips/s"}

    def timed(fn):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize(dev)
        return B / (e0.elapsed_time(e1) / steps * 1e-3)

    out["fp32"] = timed(lambda i: reference_train_step(model, opt, xs[i % 2], ys[i % 2]))
    out["fp32_note"] = "torch defaults (cuDNN convs may use TF32, matmuls fp32)"
    try:
        model.cnn_backbone.to(memory_format=torch.channels_last)

        def step_bf16(i):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                # clips are [B,T,C,H,W]; the class folds them to frames itself, channels_last weights make cuDNN pick NHWC
                return reference_train_step(model, opt, xs[i % 2], ys[i % 2])
        out["bf16_autocast_channels_last"] = timed(step_bf16)
    except Exception as e:      # the yardstick must never take the bench line down
        out["bf16_autocast_channels_last"] = None
        out["bf16_error"] = repr(e)[:200]
    del model, opt, xs, ys
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    import video_classif_b200 as vc
    from video_classif_b200 import _lib, ops
    from video_classif_b200.ingest import ingest_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = WORKLOAD
    B, T, S = W["clips_per_gpu"], W["frames"], W["size"]
    torch.manual_seed(0)
    model = vc.LRCN(W["num_classes"], T, W["hidden"], W["rnn_input"], cnn_backbone="resnet50",
                    rnn_layers=W["rnn_layers"], dropout=0.25, precision="bf16").to(dev).train()
    if not args.no_graph:
        model.enable_encoder_graph()          # frozen encoder pass replayed from a CUDA graph (captured in the warm-up)
    dp = None
    if world > 1:
        from video_classif_b200.dp import GradBucketAllReduce, broadcast_parameters
        broadcast_parameters(model)
        dp = GradBucketAllReduce(model)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4, fused=True)
    g = torch.Generator().manual_seed(1234 + rank)
    NBUF = 4                                                   # distinct synthetic batches, cycled
    host_u8 = [torch.randint(0, 256, (B, T, S, S, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(NBUF)]
    host_y = [torch.randint(0, W["num_classes"], (B,), generator=g).pin_memory() for _ in range(NBUF)]
    dev_x = [ingest_batch(h.to(dev), S, S) for h in host_u8]   # float32 [B,T,3,S,S] resident in HBM
    dev_y = [h.to(dev) for h in host_y]
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(x, y, h=None):
        """One train step; h = handle of model.encode_async(x) when the frozen encoder pass was prefetched."""
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(x, features=h), y)
        loss.backward()
        if dp is not None:
            dp.finish()
        opt.step()
        return loss

    pipelined = not args.no_pipeline

    def run_steps(n):
        """n steps over the resident batches.  Pipelined: the frozen encoder pass of batch i+1 is launched on the
        model's side stream before the trainable tail of batch i (same results, same work per step)."""
        if not pipelined:
            for i in range(n):
                step(dev_x[i % NBUF], dev_y[i % NBUF])
            return
        h = model.encode_async(dev_x[0])
        for i in range(n):
            h_next = model.encode_async(dev_x[(i + 1) % NBUF]) if i + 1 < n else None
            step(dev_x[i % NBUF], dev_y[i % NBUF], h)
            h = h_next

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident-input throughput (`value`) ----------------
    run_steps(args.warmup)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)                      # let the nvidia-smi stream start before the timed region
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.begin()
    e0.record()
    run_steps(args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    clocks = sampler.summary() if sampler else None
    t = torch.tensor([ms, float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, launches = tmax[0].item(), tsum[1].item()
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---------------- end to end from pinned host uint8 clips (`e2e`) ----------------
    copy_stream = torch.cuda.Stream(device=dev)
    NSTG = 3
    stage_u8 = [torch.empty((B, T, S, S, 3), dtype=torch.uint8, device=dev) for _ in range(NSTG)]
    stage_y = [torch.empty((B,), dtype=torch.int64, device=dev) for _ in range(NSTG)]
    ready = [torch.cuda.Event() for _ in range(NSTG)]
    freed_u8 = [torch.cuda.Event() for _ in range(NSTG)]     # the ingest kernel has read the staged clips
    freed_y = [torch.cuda.Event() for _ in range(NSTG)]      # the step has read the staged labels
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def prefetch(i):
        s = i % NSTG
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed_u8[s])
            copy_stream.wait_event(freed_y[s])
            stage_u8[s].copy_(host_u8[i % NBUF], non_blocking=True)
            stage_y[s].copy_(host_y[i % NBUF], non_blocking=True)
            ready[s].record(copy_stream)

    def ingest_slot(i):
        """K1 (uint8 -> float32 CHW /255) on the encoder's stream when pipelined, so the float clips never change pools."""
        s = i % NSTG
        st = model.side_stream(dev) if pipelined else torch.cuda.current_stream()
        with torch.cuda.stream(st):
            st.wait_event(ready[s])
            # straight into the encoder graph's static input when there is one (stream order keeps the previous pass ahead)
            buf = model.encoder_input_buffer((B, T, 3, S, S)) if pipelined else None
            x = ingest_batch(stage_u8[s], S, S, out=buf)
            freed_u8[s].record(st)
        return x

    def e2e_loop(n):
        """Every step: H2D of its uint8 clips (copy stream, two batches ahead), ingest kernel, encoder pass (one batch
        ahead on the side stream when pipelined), trainable tail, loss read back."""
        cur = torch.cuda.current_stream()
        for s in range(NSTG):
            freed_u8[s].record(cur)
            freed_y[s].record(cur)
        prefetch(0)
        if n > 1:
            prefetch(1)
        x = ingest_slot(0)
        h = model.encode_async(x) if pipelined else None
        for i in range(n):
            if i + 2 < n:
                prefetch(i + 2)
            x_next = h_next = None
            if i + 1 < n:
                x_next = ingest_slot(i + 1)
                h_next = model.encode_async(x_next) if pipelined else None
            cur.wait_event(ready[i % NSTG])                      # labels of this step
            loss = step(x, stage_y[i % NSTG], h)
            freed_y[i % NSTG].record(cur)
            loss_host.copy_(loss.detach(), non_blocking=True)    # D2H read of the step's loss
            x, h = x_next, h_next
        torch.cuda.synchronize()

    if args.no_e2e:
        ms_e2e = float("nan")
    else:
        e2e_loop(2)
        barrier()
        e0.record()
        e2e_loop(args.steps)
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t[0].item()
    e2e_value = world * B / (ms_e2e / args.steps * 1e-3)
    h2d = host_u8[0].numel() + host_y[0].numel() * 8
    d2h = 4

    # ---------------- roofline of the dominant kernel (tcgen05 GEMM / implicit-GEMM conv) -----------
    roof = None
    tflops_peak, hbm_peak, peak_src = peaks()
    if not args.no_roofline:       # every rank runs these steps (they contain the gradient all-reduce); rank 0 reports
        events = []
        orig_gemm, orig_conv, orig_convbn, orig_gram = ops.gemm_tn, ops.conv2d_nhwc, ops.conv2d_bn_nhwc, ops.conv1x1_gram_bnstats
        orig_halo = ops.conv3x3_halo_bn
        import video_classif_b200.backbone as bb

        post_events = []      # the HBM-bound instantiation: conv3 + BN3 + shortcut + ReLU (EPI_POST)

        def timed(fn, flops_of):
            def wrap(*a, **k):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                out = fn(*a, **k)
                e.record()
                events.append((s, e, flops_of(a, k, out)))
                if k.get("res") is not None and out is not None:    # bytes: A tile stream + shortcut in + output out
                    post_events.append((s, e, 2.0 * (a[0].numel() + 2 * out.numel())))
                return out
            return wrap

        def gemm_flops(a, k, out):
            return 2.0 * out.shape[0] * out.shape[1] * a[0].shape[1]

        def conv_flops(a, k, out):
            if k.get("store", True) is False:      # statistics-only pass: its time counts, its flops are recompute
                return 0.0
            x, w = a[0], a[1]
            stride, pad = a[2], a[3]
            P = (x.shape[1] + 2 * pad - w.shape[1]) // stride + 1
            Q = (x.shape[2] + 2 * pad - w.shape[2]) // stride + 1
            return 2.0 * x.shape[0] * P * Q * w.shape[0] * w.shape[1] * w.shape[2] * w.shape[3]

        ops.gemm_tn = timed(orig_gemm, gemm_flops)
        ops.conv2d_nhwc = timed(orig_conv, conv_flops)
        ops.conv2d_bn_nhwc = timed(orig_convbn, conv_flops)      # a statistics-only pass counts its flops too
        ops.conv1x1_gram_bnstats = timed(orig_gram, lambda a, k, out: 0.0)   # tensor-core statistics: time, no algorithmic flops
        ops.conv3x3_halo_bn = timed(orig_halo, lambda a, k, out: 2.0 * out.numel() * 9 * a[0].shape[-1])
        bb.conv3x3_halo_bn = ops.conv3x3_halo_bn
        bb.gemm_tn, bb.conv2d_nhwc, bb.conv2d_bn_nhwc, bb.conv1x1_gram_bnstats = (ops.gemm_tn, ops.conv2d_nhwc, ops.conv2d_bn_nhwc,
                                                                                 ops.conv1x1_gram_bnstats)
        model.enable_encoder_graph(False)                       # per-kernel CUDA events need the eager encoder
        try:
            for i in range(2):
                events.clear()
                post_events.clear()
                l2_flush.zero_()
                step(dev_x[i % NBUF], dev_y[i % NBUF])
                torch.cuda.synchronize()
        finally:
            ops.gemm_tn, ops.conv2d_nhwc, ops.conv2d_bn_nhwc, ops.conv1x1_gram_bnstats = orig_gemm, orig_conv, orig_convbn, orig_gram
            bb.gemm_tn, bb.conv2d_nhwc, bb.conv2d_bn_nhwc, bb.conv1x1_gram_bnstats = orig_gemm, orig_conv, orig_convbn, orig_gram
            ops.conv3x3_halo_bn = bb.conv3x3_halo_bn = orig_halo
            model.enable_encoder_graph(not args.no_graph)
        tot_ms = sum(s.elapsed_time(e) for s, e, _ in events)
        tot_fl = sum(f for _, _, f in events)
        achieved = tot_fl / (tot_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": tflops_peak, "unit": "TFLOP/s",
                "frac": achieved / tflops_peak, "traffic": (ncu_traffic().get("family") or {}).get("dram_bytes_per_launch"),
                "traffic_note": (ncu_traffic().get("family") or {}).get("note"),
                "kernel": "gemm_tc_kernel (tcgen05 GEMM + implicit-GEMM conv)",
                "launches_per_step": len(events), "kernel_ms_per_step": tot_ms, "algorithmic_gflop_per_step": tot_fl / 1e9,
                "share_of_step": tot_ms / ms_per_step, "peak_source": peak_src} if rank == 0 else None
        if roof is not None and post_events:
            p_ms = sum(s.elapsed_time(e) for s, e, _ in post_events)
            p_b = sum(b for _, _, b in post_events)
            gbs = p_b / (p_ms * 1e-3) / 1e9
            # the family's largest member is HBM bound: reported against the measured copy bandwidth as well
            roof["hbm_member"] = {"kernel": "gemm_tc_kernel<256,EPI_POST,TF> (conv3 + BN3 + shortcut + ReLU)", "bound": "hbm",
                                  "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                  "launches_per_step": len(post_events), "kernel_ms_per_step": p_ms,
                                  "algorithmic_mb_per_step": p_b / 1e6,
                                  "traffic": (ncu_traffic().get("hbm_member") or {}).get("dram_bytes_per_launch"),
                                  "traffic_note": (ncu_traffic().get("hbm_member") or {}).get("note")}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        clips = args.cpu_clips or B
        rate, sec, cores, kind, done = cpu_reference_step_rate(clips, 2, 1, budget_s=60.0)
        cpu = {"value": rate, "unit": "clips/s", "cores": cores, "kind": kind,
               "sample": f"{clips} clips/step x {done} timed steps (median) of the same train step, "
                         + ("the reference's own LRCN class (oracle/_ref) on torch CPU fp32 kernels" if kind == "reference"
                            else "oracle port (torch CPU fp32 kernels)")}
    torch_gpu = None
    if rank == 0 and world == 1 and not args.no_torch_baseline:
        try:
            torch_gpu = torch_gpu_baseline(dev)
        except Exception as e:
            torch_gpu = {"unavailable": repr(e)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{W['name']}: {T} frames x {S}x{S}, {B} clips/GPU, frozen ResNet-50 (train-mode BN) + "
                                   f"GELU/LN adapts + {W['rnn_layers']}-layer LSTM H={W['hidden']} + head, full train step "
                                   "(fwd, CE, bwd, Adam); BASELINE.json configs[1]",
                       "global_batch": world * B, "parallelism": f"dp{world}",
                       "timing": f"{NBUF} distinct input batches cycled; per-step working set (~6 GB activations) exceeds L2",
                       "pipeline": ("frozen encoder pass of batch i+1 on a side stream under the trainable tail of batch i "
                                    "(model.encode_async); same work per step" if pipelined else "none"),
                       "encoder_launch": "kernel by kernel" if args.no_graph else "CUDA graph replay (captured once per shape)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "path": "pinned uint8 host clips -> H2D (copy stream, two batches ahead) -> b2_ingest_u8 writing the encoder graph's "
                            "static input (the `value` leg instead copies its resident float32 clips, 154 MB, into that buffer every "
                            "step) -> encoder pass -> trainable tail -> loss D2H"},
            "gpu_launches": int(launches),
            "gflop_per_clip_fwd_backbone": RESNET50_GFLOP_PER_FRAME_112 * T,
        }
        if roof:
            line["roofline"] = roof
        if cpu:
            line["cpu_baseline"] = cpu
        if torch_gpu:
            line["torch_gpu_baseline"] = torch_gpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
