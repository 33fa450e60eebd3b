#!/usr/bin/env python
"""Headline benchmark: MobileNetLarge3D bf16 training clips/s (BASELINE.json config 3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one optimiser step over a global batch of 512 synthetic clips (3x16x224x224): every rank runs
512/(64*N) micro-batches of 64 clips (forward under bf16 autocast, cross-entropy, backward -- by default one
replay of a CUDA graph captured by picklebot_b200.graph.GraphedTrainStep per micro-batch; --no-graphs issues the
launches eagerly), the in-place NCCL all-reduce of the gradient buckets when N>1, then one multi-tensor AdamW
step (picklebot_b200.optim.AdamW; --torch-optim for torch's fused one).  Work per step is fixed, so
scaling is "strong"; the per-GPU micro-batch stays 64 so BatchNorm statistics do not depend on N
(SURVEY.md section 8d, config 3).

value   clips/s with the uint8 clips already resident in HBM (distinct buffers per micro-batch, 154 MB each,
        i.e. larger than the 126 MB L2, so no cache flush is needed between iterations).
e2e     the same step fed from pinned HOST memory through the public module API: per micro-batch H2D copy
        (issued one micro-batch ahead on a side stream, across step boundaries too, like a prefetching loader)
        of the uint8 clip batch + labels (on a side stream, double buffered) and a D2H read of the loss.
roofline  per-kernel-family GB/s from CUDA events recorded around every C-ABI launch on the launching
        stream (separate instrumented steps after the timed region), against MEASURED_PEAKS.json.
cpu_baseline  the oracle (a restatement of the reference's torch ops; the reference itself cannot travel
        to the GPU box) running the same train step on the host cores on a bounded sample.

--impl reference times that CPU path alone (rank 0; other ranks exit 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

MODEL = "MobileNetLarge3D"
NUM_CLASSES = 2
GLOBAL_BATCH = 512
MICRO = 64
CLIP = (16, 224, 224)
METRIC = "MobileNetLarge3D train clips/s"
FALLBACK_HBM_GBS = 6650.0

# BASELINE.json configs (1-based like the JSON list; configs[0] is the CPU golden case of the tests).  The driver's
# plain `python bench.py` is config 3, the configuration the headline metric is quoted on; the others are run with
# --config N and their lines are committed under profiles/.
CONFIGS = {
    2: dict(model="MobileNetLarge3D", mode="infer", clip=(16, 224, 224), micro=64, global_batch=512, nc=2,
            metric="MobileNetLarge3D inference clips/s", scaling="strong",
            workload="MobileNetLarge3D inference (eval mode), bf16 autocast, synthetic uint8 clips 3x16x224x224, batch 64 "
                     "(BASELINE.json configs[1])"),
    3: dict(model="MobileNetLarge3D", mode="train", clip=(16, 224, 224), micro=64, global_batch=512, nc=2,
            metric=METRIC, scaling="strong",
            workload="MobileNetLarge3D training step, bf16 autocast, synthetic uint8 clips 3x16x224x224 "
                     "(BASELINE.json configs[2])"),
    4: dict(model="MoViNetA2", mode="stream", clip=(64, 224, 224), chunk=8, micro=32, global_batch=32, nc=13,
            metric="MoViNetA2 causal streaming inference clips/s (64-frame clips, 8-frame chunks)", scaling="weak",
            workload="MoViNetA2 causal streaming inference, bf16, 64-frame 224x224 uint8 clips fed as 8 chunks of 8 frames, "
                     "stream buffers and cumulative pooling state resident in HBM (BASELINE.json configs[3])"),
    5: dict(model="MobileNetSmall3D", mode="train", clip=(32, 224, 224), micro=64, global_batch=512, nc=2,
            metric="MobileNetSmall3D train clips/s (32-frame clips)", scaling="strong",
            workload="MobileNetSmall3D training step, bf16 autocast, long synthetic uint8 clips 3x32x224x224 "
                     "(BASELINE.json configs[4])"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-b200"])
    ap.add_argument("--torch-mode", default="eager", choices=["eager", "compiled"],
                    help="--impl torch-b200 only: the reference's torch ops eagerly or under torch.compile")
    ap.add_argument("--no-torch-b200", action="store_true", help="skip the torch/cuDNN-on-this-B200 legs")
    ap.add_argument("--torch-compile-budget", type=float, default=240.0,
                    help="seconds the torch.compile leg may take before it is abandoned (0 disables it)")
    ap.add_argument("--torch-compile-mode", default="default",
                    help="torch.compile mode of the compiled leg.  The reference uses max-autotune-no-cudagraphs "
                         "(train.py:181), whose autotuning takes ~630 s on the bench box: measured once per round "
                         "(profiles/), while the default run uses mode='default' so that bench.py ends within minutes")
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS),
                    help="BASELINE.json configuration (1-based); 3 = the headline MobileNetLarge3D training step")
    ap.add_argument("--micro", type=int, default=None)
    ap.add_argument("--global-batch", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=1)
    ap.add_argument("--dp", default="buckets", choices=["buckets", "ddp"],
                    help="gradient exchange for N>1: picklebot_b200.dp.GradientBuckets or torch DDP")
    ap.add_argument("--torch-optim", action="store_true",
                    help="torch.optim.AdamW(fused=True) instead of picklebot_b200.optim.AdamW")
    ap.add_argument("--no-graphs", action="store_true",
                    help="issue every micro-batch eagerly instead of replaying picklebot_b200.graph.GraphedTrainStep")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.micro is None:
        args.micro = cfg["micro"]
    if args.global_batch is None:
        args.global_batch = cfg["global_batch"] * (args.gpus if cfg["scaling"] == "weak" else 1)
    return args


def measure_write_only_gbs(dev):
    """HBM write-only bandwidth (a 1 GiB fill, best of 5): on B200 it is ~3.9 TB/s against ~6.5 TB/s for a copy, which
    is the ceiling of the write-dominated kernels (the 1x1x1 expand GEMMs write 3-6x what they read)."""
    buf = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    best = None
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); buf.zero_(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return buf.numel() / best / 1e6


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference's torch ops) on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_clips_per_s(cfg, batch, reps, warm=1):
    """One 'step' of the configuration on the CPU, fp32: a training step (fwd+CE+bwd), an eval forward, or a
    64-frame clip streamed in 8-frame chunks.  Returns (best, mean) clips/s and the individual times."""
    from oracle import picklebot_oracle as O          # bench's cpu_baseline / reference leg only
    import picklebot_b200 as pb
    from picklebot_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    model, nc = cfg["model"], cfg["nc"]
    sd0 = synth.synthetic_state_dict(pb.valid_models[model](num_classes=nc).state_dict())
    clips = synth.synthetic_clips_u8(batch, *cfg["clip"], seed=0)
    x = synth.clips_to_features(clips, torch.float32).contiguous()
    labels = synth.synthetic_labels(batch, nc)
    times = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        if cfg["mode"] == "train":
            sd = O.clone_state(sd0, requires_grad=True)
            t0 = time.perf_counter()
            torch.manual_seed(7)
            O.train_step(model, sd, x, labels)
        elif cfg["mode"] == "infer":
            with torch.no_grad():
                O.MODELS[model](sd0, x)
        else:
            ch = cfg["chunk"]
            with torch.no_grad():
                O.movinet_a2_stream(sd0, [x[:, :, t:t + ch].contiguous() for t in range(0, cfg["clip"][0], ch)])
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return batch / min(times), batch / (sum(times) / len(times)), times


CPU_SAMPLE = {"train": "train steps (fwd+CE+bwd)", "infer": "eval forwards", "stream": "64-frame clips streamed in 8-frame chunks"}


def cpu_batch(cfg):
    return 1 if cfg["mode"] == "stream" else (2 if cfg["clip"][0] > 16 else 4)


def run_reference(args, rank):
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    batch = cpu_batch(cfg)
    t0 = time.perf_counter()
    best, mean, times = cpu_clips_per_s(cfg, batch, reps=max(1, args.steps), warm=max(1, min(args.warmup, 2)))
    value = mean
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * batch / value,
        "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"] + " -- CPU fp32 arm", "global_batch": args.global_batch,
                   "micro_batch": batch, "clip": list(cfg["clip"])},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{len(times)} {CPU_SAMPLE[cfg['mode']]} of {batch} clips each (one step = a bounded "
                                   f"sample of the {args.global_batch}-clip global batch); oracle = torch-op restatement "
                                   f"of the reference, which is pure Python and absent on the GPU box"},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# torch/cuDNN on this same B200: the reference's own GPU path (SURVEY section 8d "the real bar")
# --------------------------------------------------------------------------------------------------
def run_torch_b200(args):
    """Child process (``--impl torch-b200``): the reference's ops (oracle restatement = the same ATen/cuDNN calls as
    the reference modules) doing the same training step on cuda:LOCAL_RANK: micro-batches of 64 uint8 clips ->
    ``permute/.to(bf16)/255`` (train.py:106) -> forward under autocast(bf16) -> CrossEntropyLoss -> backward
    (train.py:264-269), fused AdamW per 512 clips; eager with cudnn.benchmark (train.py:193) or under
    ``torch.compile(mode='max-autotune-no-cudagraphs')`` (train.py:181).  CUDA-event timed; prints one JSON line."""
    from oracle import picklebot_oracle as O          # a measured BASELINE leg, like cpu_baseline; never the product
    import picklebot_b200 as pb
    from picklebot_b200 import synth
    t_start = time.perf_counter()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    cfg = CONFIGS[args.config]
    train = cfg["mode"] == "train"
    micro, accum = args.micro, args.global_batch // args.micro
    sd = synth.synthetic_state_dict(pb.valid_models[cfg["model"]](num_classes=cfg["nc"]).state_dict())
    sd = O.clone_state(sd, requires_grad=train, device=dev)
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.AdamW(params, lr=3e-4, weight_decay=5e-4, fused=True) if train else None
    clips = [synth.synthetic_clips_u8_device(micro, *cfg["clip"], seed=a, device=dev) for a in range(2)]
    labels = [synth.synthetic_labels(micro, cfg["nc"], seed=77 + a).to(dev) for a in range(2)]
    fwd = O.MODELS[cfg["model"]]
    if args.torch_mode == "compiled":
        fwd = torch.compile(fwd, mode=args.torch_compile_mode)

    def step():
        for a in range(accum):
            x = clips[a & 1].permute(0, 4, 1, 2, 3).to(torch.bfloat16) / 255            # train.py:106
            if train:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    loss = F.cross_entropy(fwd(sd, x, True), labels[a & 1]) / accum
                loss.backward()
            else:
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):   # estimate_loss, train.py:123-153
                    fwd(sd, x, False)
        if train:
            opt.step()
            opt.zero_grad(set_to_none=True)

    step()                                    # warm-up: cudnn.benchmark autotuning / compilation happen here
    torch.cuda.synchronize()
    warm_s = time.perf_counter() - t_start
    step()
    steps = max(1, min(args.steps, 3))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"mode": args.torch_mode + (f" ({args.torch_compile_mode})" if args.torch_mode == "compiled" else ""),
                      "value": args.global_batch / (ms / 1000.0), "unit": "clips/s",
                      "ms_per_step": ms, "steps": steps, "micro_batch": micro, "warmup_and_setup_s": warm_s,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "torch": torch.__version__,
                      "cudnn": torch.backends.cudnn.version()}), flush=True)


def torch_b200_legs(args, local):
    """Run the two torch legs as child processes of rank 0 (own CUDA context on the same GPU, which is idle while the
    parent waits) so that a slow torch.compile cannot wedge the bench: the compiled leg gets a wall-clock budget."""
    out = {}
    for mode, budget in (("eager", 240.0), ("compiled", args.torch_compile_budget)):
        if budget <= 0:
            out[mode] = {"unavailable": "disabled (budget 0)"}
            continue
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "torch-b200", "--torch-mode", mode, "--steps",
               str(args.steps), "--micro", str(args.micro), "--global-batch", str(args.global_batch),
               "--config", str(args.config), "--torch-compile-mode", args.torch_compile_mode]
        env = dict(os.environ, LOCAL_RANK=str(local))
        for k in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        t0 = time.perf_counter()
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=budget, env=env)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode == 0 and line:
                out[mode] = json.loads(line[-1])
            else:
                out[mode] = {"unavailable": f"exit {r.returncode}: " + (r.stderr.strip().splitlines() or ["?"])[-1][:300]}
        except subprocess.TimeoutExpired:
            out[mode] = {"unavailable": f"not finished within the {budget:.0f} s budget (--torch-compile-budget)"}
        out[mode]["wall_s"] = time.perf_counter() - t0
    out["what"] = ("the reference's torch ops (ATen/cuDNN) on this same B200: uint8 clips -> .to(bf16)/255 -> autocast(bf16) "
                   "forward + CE + backward at micro-batch 64, fused AdamW per 512 clips; eager with cudnn.benchmark "
                   f"(train.py:193,264-269) and torch.compile(mode='{args.torch_compile_mode}') (train.py:181 uses "
                   "max-autotune-no-cudagraphs: ~630 s of autotuning here, measured once per round, see profiles/)")
    return out


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.impl == "torch-b200":
        run_torch_b200(args)
        return
    import torch.distributed as dist
    import picklebot_b200 as pb
    from picklebot_b200 import _lib, synth

    # stdout carries exactly ONE JSON line: library chatter (e.g. "NCCL version ...") goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert args.gpus == world, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    assert _lib.lib().pb_device_check() == 0, _lib.lib().pb_last_error_string().decode()
    micro = args.micro
    assert args.global_batch % (micro * world) == 0
    accum = args.global_batch // (micro * world)

    cfg = CONFIGS[args.config]
    mode, clip_shape, nc = cfg["mode"], cfg["clip"], cfg["nc"]
    train = mode == "train"
    torch.manual_seed(1234)
    model = pb.valid_models[cfg["model"]](num_classes=nc)
    model.load_state_dict(synth.synthetic_state_dict(model.state_dict()))
    model = model.to(dev)
    model = model.train() if train else model.eval()
    net = model
    buckets = None
    if world > 1 and train:
        if args.dp == "ddp":
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
        else:
            from picklebot_b200 import dp as pbdp
            pbdp.broadcast_module(model)
            # 1/world is folded into the loss scale below: the SUM all-reduce is the average, no division kernels
            buckets = pbdp.GradientBuckets(model.parameters(), grad_as_bucket_view=True, average=False)
    opt = None
    if not train:
        pass
    elif args.torch_optim:
        opt = torch.optim.AdamW(model.parameters(), lr=3e-4, weight_decay=5e-4, fused=True)
    else:
        from picklebot_b200.optim import AdamW          # one pb_adamw_step launch over all ~170 tensors
        opt = AdamW(model.parameters(), lr=3e-4, weight_decay=5e-4)
    use_graph = not args.no_graphs and args.dp == "buckets" and mode != "stream"      # (stream: GraphedStream below)
    from picklebot_b200 import loss as pbloss               # cross-entropy + accuracy count in one kernel (pb_ce_loss)

    # synthetic uint8 clips: this rank's shard of each global batch, distinct per micro-batch
    clips = [synth.synthetic_clips_u8_device(micro, *clip_shape, seed=1000 * rank + a, device=dev) for a in range(accum)]
    labels = [synth.synthetic_labels(micro, nc, seed=77 + 1000 * rank + a).to(dev) for a in range(accum)]

    # One micro-batch = forward + loss + backward.  Default: captured once in a CUDA graph and replayed (issuing the
    # ~370 launches from Python costs the host ~11 ms per 14 ms of GPU work, which starves the GPUs once eight
    # processes share the box's cores); gradients accumulate in place, the exchange runs after the last replay.
    gstep = gfwd = None
    gstream = None
    # mean over the micro-batch / accumulation steps (/ ranks when GradientBuckets sums instead of averaging)
    loss_scale = 1.0 / accum / (world if buckets is not None else 1)
    if mode == "stream" and not args.no_graphs:
        from picklebot_b200.graph import GraphedStream
        gstream = GraphedStream(model, clips[0][:, :cfg["chunk"]].permute(0, 4, 1, 2, 3))
    if use_graph and not train:
        from picklebot_b200.graph import GraphedForward
        gfwd = GraphedForward(model, clips[0].permute(0, 4, 1, 2, 3))
    if use_graph and train:
        from contextlib import nullcontext
        from picklebot_b200.graph import GraphedTrainStep
        with (buckets.no_sync() if buckets is not None else nullcontext()):
            gstep = GraphedTrainStep(model, clips[0].permute(0, 4, 1, 2, 3), labels[0],
                                     loss_fn=lambda logits, y: pbloss.cross_entropy(logits, y, scale=loss_scale))
        grad_list = [p.grad for p in model.parameters() if p.grad is not None]

    def zero_grads():
        if buckets is not None:
            buckets.zero_grad()                 # one fill per bucket; .grad tensors are views of the buckets
        elif gstep is None:
            opt.zero_grad(set_to_none=True)
        else:
            torch._foreach_zero_(grad_list)     # the graph accumulates into these very tensors

    def stream_clip(x_u8, eager=False):
        """One batch of long clips through the causal streaming path, chunk by chunk; the stream state (tail frames
        of every temporal conv, cumulative squeeze-excite / head sums) lives on the device between the calls."""
        if gstream is not None and not eager:
            gstream.reset()
            logits = None
            for t0 in range(0, clip_shape[0], cfg["chunk"]):
                logits = gstream(x_u8[:, t0:t0 + cfg["chunk"]].permute(0, 4, 1, 2, 3))
            return logits
        state = model.init_stream_state()
        logits = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            for t0 in range(0, clip_shape[0], cfg["chunk"]):
                logits, state = model.forward_stream(x_u8[:, t0:t0 + cfg["chunk"]].permute(0, 4, 1, 2, 3), state)
        return logits

    def micro_step(x_u8, y, sync_grads, eager=False, first=True):
        if _lib.PROFILER is not None:
            # instrumented pass: park the GPU for ~15 ms first, so that every launch of this micro-batch is already
            # queued when the GPU gets to it -- otherwise the event pairs also time the host's ~10 us between two
            # ctypes calls whenever the GPU has caught up (it inflated the short kernels by 5-15 %, box-dependent)
            torch.cuda._sleep(30_000_000)
        if mode == "stream":
            return stream_clip(x_u8, eager)
        if mode == "infer":
            if gfwd is not None and not eager:
                return gfwd(x_u8.permute(0, 4, 1, 2, 3))
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return model(x_u8.permute(0, 4, 1, 2, 3))
        if gstep is not None and not eager:
            # the optimizer moved the weights before the first micro-batch of a step only: the others replay the
            # graph that skips the weight casts
            loss = gstep(x_u8.permute(0, 4, 1, 2, 3), y, weights_changed=first)
            if sync_grads and buckets is not None:
                buckets.reduce_all()
                buckets.finish()
            return loss
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = net(x_u8.permute(0, 4, 1, 2, 3))        # (B,3,T,H,W) view of the uint8 NTHWC batch
            loss = pbloss.cross_entropy(logits, y, scale=loss_scale)
        if world > 1 and not sync_grads:
            with (buckets.no_sync() if buckets is not None else net.no_sync()):
                loss.backward()
        else:
            loss.backward()
            if buckets is not None:
                buckets.finish()            # wait for the bucketed all-reduces, averaged grads back in .grad
        return loss.detach()

    def step_resident(eager=False):
        tot = None
        for a in range(accum):
            l = micro_step(clips[a], labels[a], a == accum - 1, eager, first=a == 0)
            tot = l if (tot is None or not train) else tot + l
        if train:
            opt.step()
            zero_grads()
        return tot

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # the clock sampler starts with the warm-up (nvidia-smi needs ~0.5 s to deliver its first line) and runs
    # through the timed region: every sample is taken under the same load
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step_resident()
    n0 = _lib.launch_count()
    ms = timed(step_resident, args.steps)
    launches = _lib.launch_count() - n0
    if gstep is not None:                       # replays launch the captured kernels without passing the C ABI
        launches += args.steps * (gstep.launches + (accum - 1) * gstep.launches_warm)
    if gfwd is not None:
        launches += args.steps * accum * gfwd.launches
    if gstream is not None:
        launches += args.steps * accum * (clip_shape[0] // cfg["chunk"]) * gstream.launches
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = args.global_batch / (ms_per_step / 1000.0)

    # ---- end to end: pinned host clips -> H2D (side stream, double buffered) -> step -> D2H loss ----
    e2e = None
    if not args.no_e2e:
        host_clips = [c.cpu().pin_memory() for c in clips]
        host_labels = [l.cpu().pin_memory() for l in labels]
        copy_stream = torch.cuda.Stream(device=dev)
        dbuf = [torch.empty_like(clips[0]) for _ in range(2)]
        lbuf = [torch.empty_like(labels[0]) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        # what the user reads back every step: the loss (training) or the logits of the last micro-batch (inference)
        host_loss = (torch.zeros((), dtype=torch.float32) if train else
                     torch.zeros((micro, nc), dtype=torch.float32)).pin_memory()

        def upload(a, slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])
                dbuf[slot].copy_(host_clips[a], non_blocking=True)
                lbuf[slot].copy_(host_labels[a], non_blocking=True)
                ready[slot].record(copy_stream)

        seq = [0]      # running micro-batch number: its buffer slot is seq & 1, uploaded one micro-batch ahead

        def step_e2e():
            cur = torch.cuda.current_stream()
            tot = None
            for a in range(accum):
                slot = seq[0] & 1
                upload((a + 1) % accum, slot ^ 1)      # prefetch the next micro-batch (possibly the next step's first)
                cur.wait_event(ready[slot])
                l = micro_step(dbuf[slot], lbuf[slot], a == accum - 1, first=a == 0)
                consumed[slot].record(cur)
                tot = l if (tot is None or not train) else tot + l
                seq[0] += 1
            if train:
                opt.step()
                zero_grads()
            host_loss.copy_(tot.float(), non_blocking=True)
            torch.cuda.current_stream().synchronize()          # the user reads the result every step
            return host_loss

        for s in range(2):
            consumed[s].record(torch.cuda.current_stream())
        upload(0, 0)                                           # the pipeline starts one upload ahead
        step_e2e()
        ms_e = timed(step_e2e, args.steps) / args.steps
        e2e = {"value": args.global_batch / (ms_e / 1000.0), "unit": "clips/s",
               "h2d_bytes_per_step": accum * (clips[0].numel() + labels[0].numel() * 8),
               "d2h_bytes_per_step": host_loss.numel() * 4, "ms_per_step": ms_e}
        del host_clips, dbuf

    # ---- per-kernel roofline: instrumented steps (CUDA events around every launch) -----------------
    peak, peak_kind = peaks()
    kernels, roofline = {}, None
    if rank == 0 or world > 1:
        prof = _lib.KernelProfiler()
        _lib.PROFILER = prof
        for _ in range(max(1, args.profile_steps)):
            step_resident(eager=True)           # events around every C-ABI call need the eager path
        _lib.PROFILER = None
        summ = prof.summary()
        if os.environ.get("PB_BENCH_DETAIL"):
            det = prof.summary(by_tag=True)
            rows = sorted(det.items(), key=lambda kv: -kv[1]["ms"])
            with open(os.environ["PB_BENCH_DETAIL"], "w") as f:
                for k, v in rows:
                    gb = v["bytes"] / (v["ms"] * 1e6) if v["ms"] > 0 else 0
                    f.write(f"{v['ms']:9.3f} ms  {v['launches']:4d}x  {gb:8.1f} GB/s  {k}\n")
        tot_ms = sum(v["ms"] for v in summ.values())
        write_peak = measure_write_only_gbs(dev)
        for name, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
            gbs = v["bytes"] / (v["ms"] * 1e6) if v["ms"] > 0 else 0.0
            # read/write-aware bound per launch: all bytes at the copy rate, or the written bytes at the write-only rate
            bound_ms = sum(max(nb / (peak * 1e6), wb / (write_peak * 1e6)) for nb, wb in v["per_launch"])
            kernels[name] = {"launches_per_step": v["launches"] // max(1, args.profile_steps),
                             "ms_per_step": v["ms"] / max(1, args.profile_steps), "share": v["ms"] / tot_ms,
                             "GBps": gbs, "frac_of_hbm_peak": gbs / peak,
                             "frac_of_rw_bound": bound_ms / v["ms"] if v["ms"] > 0 else 0.0,
                             "written_share_of_bytes": v["wbytes"] / v["bytes"] if v["bytes"] else 0.0,
                             "avg_us": 1000.0 * v["ms"] / v["launches"],
                             "alg_bytes_per_launch": v["bytes"] / v["launches"]}
        top = next(iter(kernels))
        kt = kernels[top]
        # DRAM bytes per launch of the dominant kernel from the committed ncu capture of this same command
        # (profiles/roofline_traffic.json; null if that kernel has no capture)
        traffic = None
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "roofline_traffic.json")) as f:
                traffic = json.load(f).get(top, {}).get("dram_bytes_per_launch")
        except OSError:
            pass
        roofline = {"kernel": top, "bound": "hbm", "achieved": kt["GBps"], "peak": peak, "unit": "GB/s",
                    "frac": kt["frac_of_hbm_peak"], "traffic": traffic, "peak_source": peak_kind + " copy bandwidth",
                    "share_of_step": kt["share"], "avg_launch_us": kt["avg_us"],
                    "alg_bytes_per_launch": kt["alg_bytes_per_launch"]}
        roofline["hbm_write_only_GBps"] = write_peak
        roofline["frac_of_rw_bound"] = kt["frac_of_rw_bound"]
        roofline["note"] = ("frac is against the copy bandwidth of MEASURED_PEAKS.json.  A write-only stream reaches "
                            "hbm_write_only_GBps on this GPU (measured here); frac_of_rw_bound uses, per launch, the "
                            "larger of (all bytes / copy rate) and (written bytes / write-only rate)")
        dw = {k: v for k, v in kernels.items() if k.startswith("pb_dwconv3d")}
        if dw:
            b = sum(v["GBps"] * v["ms_per_step"] for v in dw.values())
            t = sum(v["ms_per_step"] for v in dw.values())
            roofline["depthwise_conv3d_GBps"] = b / t
            roofline["depthwise_conv3d_frac"] = b / t / peak

    # Evidence capture: under `ncu --profile-from-start off` (tools/collect_evidence.sh) exactly one eager micro-batch
    # (forward + loss + backward, every launch of the hot path once) runs between cudaProfilerStart/Stop.
    if os.environ.get("PB_NCU_RANGE") and rank == 0:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        micro_step(clips[0], labels[0], True, eager=True)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if train:
            zero_grads()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb = cpu_batch(cfg)
        best, mean, times = cpu_clips_per_s(cfg, cb, reps=3 if train else 2, warm=1)
        cpu = {"value": mean, "unit": "clips/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{len(times)} {CPU_SAMPLE[mode]} of {cb} clips 3x{'x'.join(map(str, clip_shape))}, fp32, oracle "
                         f"(torch-op restatement of the reference) on the host cores"}

    torch_b200 = None
    if rank == 0 and world == 1 and not args.no_torch_b200 and mode != "stream":    # (no reference streaming path exists)
        torch.cuda.empty_cache()
        torch_b200 = torch_b200_legs(args, local)
        for leg in ("eager", "compiled"):
            if "value" in torch_b200.get(leg, {}):
                torch_b200[leg]["ours_over_torch"] = value / torch_b200[leg]["value"]

    if rank == 0:
        clip_mb = clips[0].numel() / 1e6
        conf = {"workload": cfg["workload"], "baseline_config": args.config,
                "global_batch": args.global_batch, "micro_batch_per_gpu": micro, "accum_steps": accum,
                "parallelism": f"dp{world}", "clip": list(clip_shape),
                "l2": f"inputs larger than L2 ({clip_mb:.0f} MB uint8 per micro-batch, distinct buffers); no flush",
                "num_classes": nc}
        if train:
            conf.update({
                "gradient_exchange": ("none (1 GPU)" if world == 1 else
                                      "picklebot_b200.dp.GradientBuckets: .grad tensors are views of the bucket "
                                      "buffers, in-place NCCL all-reduce once per optimizer step" + (
                                          " after the last graph replay" if gstep is not None else
                                          ", launched from post-accumulate-grad hooks") if args.dp == "buckets"
                                      else "torch DDP, no_sync on all but the last micro-batch"),
                "launch": ("forward+loss+backward of a micro-batch captured once in a CUDA graph "
                           "(picklebot_b200.graph.GraphedTrainStep) and replayed; clips are copied device-to-"
                           "device into the graph's static input" if gstep is not None else "eager"),
                "criterion": "picklebot_b200.loss.cross_entropy (pb_ce_loss), mean over the micro-batch / accum_steps",
                "optimizer": ("torch.optim.AdamW(fused=True)" if args.torch_optim else
                              "picklebot_b200.optim.AdamW (multi-tensor pb_adamw_step)") + " inside the timed region"})
        elif mode == "infer":
            conf["launch"] = ("eval forward of a micro-batch captured once in a CUDA graph (picklebot_b200.graph."
                              "GraphedForward) and replayed" if gfwd is not None else "eager")
        else:
            conf["launch"] = (f"{clip_shape[0] // cfg['chunk']} forward_stream chunk steps per clip batch, " +
                              ("one captured CUDA graph replayed per chunk (picklebot_b200.graph.GraphedStream)"
                               if gstream is not None else "eager") +
                              "; stream buffers and cumulative pooling state resident in HBM, updated in place")
        line = {
            "metric": cfg["metric"], "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": conf,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "torch_b200": torch_b200, "kernels": kernels,
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
