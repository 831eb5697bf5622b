#!/usr/bin/env python
"""Benchmark of the image->video sampling path (BASELINE.json metric: frames/sec, BAIR 64x64 seq16).

    python bench.py --gpus 1 --steps K --warmup W                       # this framework on 1 B200 (headline config)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                                # reference algorithm on host cores
    python bench.py --impl reference-gpu ...                            # reference algorithm, eager PyTorch on the GPU
    python bench.py --config landscape_b32_fast | dtdb_fire_seq24_b32 | iper128_transfer_b64 | bair_b6 | bair_b1

A "step" is one pass of the sampling path (embedder -> inverse cINN -> 3-D conv decoder; the transfer config adds the
3-D video encoder and the forward cINN of one query clip) over one batch of synthetic start frames with random-init
weights in the reference's checkpoint format.  The default config is BASELINE.json configs[1] (BAIR 64x64,
seq_length 16, batch 64 per GPU, fp32 parity arithmetic); the other BASELINE configs are selectable with --config.
One JSON line is printed by rank 0:

  value      frames/s with inputs resident in HBM (CUDA events, barrier+sync both sides, max over ranks;
             N>1: per-GPU batch fixed = weak scaling, the all-gather of finished frames is inside: it is issued on a
             side stream behind the step that produced the frames and the timed region ends when it has completed)
  e2e        same metric through the public API call with HOST buffers: pinned H2D of start frames + residual and
             D2H of the frames inside the timed region (the read-back of step i runs on a copy stream under the kernels
             of step i + 1; the region ends when the last copy has landed)
  ab         the two timings repeated alternately (device, e2e) x 3, value / e2e = the median leg: resolves the host-copy cost from the
             box's clock noise
  roofline   the dominant kernel family: algorithmic FLOPs of its launches / their summed CUDA-event time over K
             more steps run with events around every launch, against the measured peak in MEASURED_PEAKS.json
  cpu_baseline  the oracle port (oracle/oracle_torch.py, = the reference's PyTorch arithmetic) timed on the host
             cores on a bounded sample (B = 6, the scripts' default, and B = 1; rank 0, N=1 only) at the fastest
             of a few tuned thread counts (`cores`)
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

UNIT = "frames/s"

# BASELINE.json configs (SURVEY 8d "Configs restated").  batch = start frames per GPU per step.
CONFIGS = {
    # configs[1]: the headline
    "bair_b64": dict(dataset="bair", batch=64, seq_length=16, conv_engine=1, micro_batch=64, mode="sample",
                     metric="frames/sec (BAIR 64x64 seq16)"),
    # configs[0]: the scripts' operating points (generate_samples.py -bs 6; B = 1)
    "bair_b6": dict(dataset="bair", batch=6, seq_length=16, conv_engine=1, micro_batch=6, mode="sample",
                    metric="frames/sec (BAIR 64x64 seq16, batch 6)"),
    "bair_b1": dict(dataset="bair", batch=1, seq_length=16, conv_engine=1, micro_batch=1, mode="sample",
                    metric="frames/sec (BAIR 64x64 seq16, batch 1)"),
    # configs[2]: reduced-precision decoder (single fp16 product per MAC, flow + embedder stay fp32-grade)
    "landscape_b32_fast": dict(dataset="landscape", batch=32, seq_length=16, conv_engine=2, micro_batch=32, mode="sample",
                               metric="frames/sec (Landscape 128x128 seq16, reduced-precision decoder)"),
    "landscape_b32": dict(dataset="landscape", batch=32, seq_length=16, conv_engine=1, micro_batch=32, mode="sample",
                          metric="frames/sec (Landscape 128x128 seq16)"),
    # configs[3]: seq_length 24 -> two decoder passes, 32 frames per sample; 256 over 8 GPUs = 32 per GPU
    "dtdb_fire_seq24_b32": dict(dataset="dtdb_fire", batch=32, seq_length=24, conv_engine=1, micro_batch=32, mode="sample",
                                metric="frames/sec (DTDB fire 128x128 seq24)"),
    # configs[4]: transfer path on the declared 128x128 iPER geometry; 512 over 8 GPUs = 64 per GPU
    "iper128_transfer_b64": dict(dataset="iper128", batch=64, seq_length=16, conv_engine=1, micro_batch=16, mode="transfer",
                                 metric="frames/sec (iPER 128x128 seq16 transfer)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--config", default="bair_b64", choices=sorted(CONFIGS))
    ap.add_argument("--dataset", default=None)
    ap.add_argument("--batch", type=int, default=None, help="start frames per GPU per step")
    ap.add_argument("--seq-length", type=int, default=None)
    ap.add_argument("--micro-batch", type=int, default=None)
    ap.add_argument("--conv-engine", type=int, default=None, help="1 tcgen05 split-fp16 (parity), 0 fp32 SIMT, 2 fp16 fast")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ckpt-dir", default=None, help="reuse / create synthetic checkpoints here")
    ap.add_argument("--streams", type=int, default=1, help="decoder micro-batches alternate over this many CUDA streams")
    ap.add_argument("--dump-launches", default=None, help="CSV of per-launch device times of the timed steps")
    ap.add_argument("--opt", action="append", default=[], help="name=value tuning switch (i2v_set_option), A/B runs")
    ap.add_argument("--gather", default="overlap", choices=["overlap", "inline", "uint8", "p2p"],
                    help="N>1: NCCL all-gather of the frames on a side stream behind the next step (default), on the compute stream "
                         "(inline), of uint8 pixels, or EXPERIMENTAL peer-to-peer copies into symmetric memory (p2p: 52.8 ms "
                         "per step at N = 2, but one unexplained hang -- see dist.FrameGather)")
    ap.add_argument("--graph", type=int, default=0, help="1: replay each decoder micro-batch from a CUDA graph")
    ap.add_argument("--legs", type=int, default=3, help="alternating (device, e2e) timing legs of K steps each; value / e2e = the median "
                                                        "leg (1 under ncu, where every launch costs ~0.2 s of tool overhead)")
    ap.add_argument("--e2e-probe", action="store_true",
                    help="diagnostic: time device-resident / H2D-only / D2H-only / full end-to-end steps alternately and exit")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    for k in ("dataset", "batch", "seq_length", "micro_batch", "conv_engine"):
        if getattr(a, k) is None:
            setattr(a, k, c[k])
    a.mode = c["mode"]
    a.metric = c["metric"]
    return a


# ------------------------------------------------------------------------------------------ helpers
def synthetic_ckpt(args, rank, barrier):
    """Full-size synthetic checkpoints in the reference layout (rank 0 writes, everyone reads)."""
    from image2video_synthesis_using_cinns_b200 import synthetic
    enc = args.mode == "transfer"
    base = args.ckpt_dir or os.path.join(tempfile.gettempdir(), f"i2v_bench_ckpt_{args.dataset}{'_enc' if enc else ''}")
    done = os.path.join(base, ".done")
    if rank == 0 and not os.path.exists(done):
        synthetic.write_synthetic_checkpoints(base, args.dataset, seed=0, with_encoder=enc)
        open(done, "w").close()
    barrier()
    return os.path.join(base, "stage2") + "/"


def make_inputs(total_batch, img, z_dim):
    g = torch.Generator().manual_seed(1234)
    x0 = torch.rand(total_batch, 3, img, img, generator=g) * 2 - 1     # SURVEY 8d synthetic inputs
    torch.manual_seed(4321)
    residual = torch.randn(total_batch, z_dim)                          # CPU RNG like get_model.py:59
    return x0, residual


def make_query(img):
    g = torch.Generator().manual_seed(77)
    return torch.rand(1, 16, 3, img, img, generator=g) * 2 - 1          # SURVEY 8d transfer query clip


def workload_config(args, img, world):
    """The `config` object BOTH arms print (the reference arm runs bounded samples of the same workload)."""
    what = "transfer path (3-D encoder + forward cINN of one query clip, inverse cINN, decoder)" if args.mode == "transfer" \
        else "sampling path (embedder, inverse cINN, decoder)"
    return {"workload": f"{args.dataset.upper()} {img}x{img} seq_length={args.seq_length}, batch={args.batch} per GPU, {what}",
            "name": args.config, "l2": "per-step working set (activations+weights, GBs) exceeds the 126 MB L2; no explicit flush",
            "parallelism": f"batch-sharded x{world}, one all-gather of frames" if world > 1 else "single GPU"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Drop what was sampled so far (warm-up): the timed region starts here."""
        self.rows = []

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = sorted(float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": pw[len(pw) // 2] if pw else None, "reasons": reasons, "samples": len(sm)}


def _host_cpus():
    """CPUs this process may really use: affinity mask, capped by the cgroup CPU quota when there is one."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(-(-int(quota) // int(period)))))
    except (OSError, ValueError):
        pass
    return n


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_torch as ot     # CPU baseline / reference legs only
    return ot


def cpu_reference_rate(args, mp, n_calls=3, warm=1, sample_batch=6, budget_s=40.0):
    """The reference's arithmetic (oracle port) on the host cores, bounded sample of the benchmark workload.

    The thread count is tuned first on B=1 calls (ascending candidates, stop once it gets slower): all logical CPUs is
    not the fastest setting for this model -- 128 threads on the GPU box ran a B=1 call in 31 s against 1.5 s on 8 --
    and the baseline should be the CPU's best.  `cores` in the result is the thread count used.  Sample: `n_calls`
    calls at B = `sample_batch` (generate_samples.py's default -bs 6, BASELINE configs[0]) plus the B = 1 rate;
    the calls stop early once `budget_s` seconds of CPU work are spent."""
    ot = _oracle()
    avail = _host_cpus()
    transfer = args.mode == "transfer"
    om = ot.OracleModel(mp, args.seq_length, transfer=transfer)
    img = om.opt["Data"]["img_size"]
    x1, r1 = make_inputs(1, img, om.z_dim)
    query = make_query(img) if transfer else None

    def timed(x, r):
        t0 = time.perf_counter()
        out = om.transfer(query, x) if transfer else om.forward(x, r, batch_slice=False)
        return time.perf_counter() - t0, out

    t_begin = time.perf_counter()
    best_t, best_n, worse = None, None, 0
    for n in sorted({c for c in (4, 8, 16, 32, 64, 128, avail) if c <= avail} | {min(avail, 4)}):
        torch.set_num_threads(n)
        timed(x1, r1)                       # first call at a thread count pays primitive creation
        dt, out1 = timed(x1, r1)
        if best_t is None or dt < best_t:
            best_t, best_n, worse = dt, n, 0
        else:
            worse += 1
            if worse >= 2 or dt > 3 * best_t:
                break
        if time.perf_counter() - t_begin > budget_s / 2:
            break
    torch.set_num_threads(best_n)
    b1_rate = out1.shape[0] * out1.shape[1] / best_t
    xb, rb = make_inputs(sample_batch, img, om.z_dim)
    times, out = [], out1
    for i in range(warm + n_calls):
        if times and time.perf_counter() - t_begin > budget_s:
            break
        dt, out = timed(xb, rb)
        if i >= warm:
            times.append(dt)
    if times:
        frames = out.shape[0] * out.shape[1]
        med = sorted(times)[len(times) // 2]
        rate = frames / med
    else:
        rate, sample_batch = b1_rate, 1
    what = "Model.transfer" if transfer else "Model.forward"
    sample = (f"{len(times)} x {what}(B={sample_batch}, {args.dataset} {img}x{img}, seq {args.seq_length}) after {warm} warm-up, median; "
              f"B=1 rate {b1_rate:.1f} frames/s; {best_n} threads = fastest of the tuned counts on {avail} usable CPUs")
    return rate, best_n, sample, b1_rate


# ------------------------------------------------------------------------------------------ reference arms
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mp = synthetic_ckpt(args, 0, lambda: None)
    t0 = time.perf_counter()
    fps, cores, sample, b1 = cpu_reference_rate(args, mp, n_calls=max(1, args.steps), warm=max(1, min(args.warmup, 2)),
                                                budget_s=150.0)
    from image2video_synthesis_using_cinns_b200.config import DATASETS
    img = DATASETS[args.dataset]["img_size"]
    T = 16 * -(-args.seq_length // 16)
    line = {
        "impl": "reference", "metric": args.metric, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * args.batch * T / fps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic start frames U[-1,1], random-init weights in the reference checkpoint format",
        "config": workload_config(args, img, int(os.environ.get("WORLD_SIZE", "1"))),
        "impl_detail": "reference PyTorch arithmetic (oracle port, pinned bit-exactly to the reference modules; the Python reference tree "
                       "cannot travel to the GPU box) on the host cores; each step = one call on a bounded sample of the batch, "
                       "ms_per_step extrapolates the per-frame rate to the full batch",
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "b1_value": b1},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def run_reference_gpu(args):
    """The reference's eager PyTorch path on the B200: the oracle port with its state-dicts on
    CUDA -- cuDNN / cuBLAS kernels, fp32 and TF32 -- on the same workload.  Checker-side leg, never the product path."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ot = _oracle()
    mp = synthetic_ckpt(args, 0, lambda: None)
    dev = torch.device("cuda", 0)
    transfer = args.mode == "transfer"
    om = ot.OracleModel(mp, args.seq_length, transfer=transfer)
    om.to(dev)
    img = om.opt["Data"]["img_size"]
    x0, res = make_inputs(args.batch, img, om.z_dim)
    x0_d, res_d = x0.to(dev), res.to(dev)
    query = make_query(img).to(dev) if transfer else None
    chunk = min(args.batch, 16)             # eager mode materialises ~10 full tensors per block: bound the activations

    def step():
        outs = []
        for b0 in range(0, args.batch, chunk):
            xs, rs = x0_d[b0:b0 + chunk], res_d[b0:b0 + chunk]
            outs.append(om.transfer(query, xs) if transfer else om.forward(xs, rs, batch_slice=False))
        return torch.cat(outs)

    out = {}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        for _ in range(max(2, min(args.warmup, 3))):
            seq = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            seq = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[name] = {"ms_per_step": ms, "value": seq.shape[0] * seq.shape[1] / (ms / 1e3)}
    line = {
        "impl": "reference-gpu", "metric": args.metric, "value": out["fp32"]["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": out["fp32"]["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic start frames U[-1,1], random-init weights in the reference checkpoint format",
        "config": workload_config(args, img, 1),
        "impl_detail": f"reference PyTorch arithmetic (oracle port) in eager mode on cuda:0, cuDNN/cuBLAS kernels, chunks of {chunk} samples; "
                       "value = fp32 (allow_tf32=False), tf32 = the same with TF32 tensor cores allowed",
        "tf32": out["tf32"], "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch.distributed as dist
    from image2video_synthesis_using_cinns_b200 import cli, lib
    from image2video_synthesis_using_cinns_b200.dist import FrameGather, HostFrameSink
    from image2video_synthesis_using_cinns_b200.get_model import Model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)     # NCCL_DEBUG is left as the caller set it
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    L = lib.load()
    for kv in args.opt:
        k, v = kv.split("=", 1)
        lib.set_option(k, float(v))
    transfer = args.mode == "transfer"
    mp = synthetic_ckpt(args, rank, barrier)
    model = Model(mp, args.seq_length, transfer=transfer, device=dev, micro_batch=args.micro_batch, conv_engine=args.conv_engine,
                  streams=args.streams, graph=bool(args.graph))
    img = model.config.Data["img_size"]
    B = args.batch
    x0_all, res_all = make_inputs(B * world, img, model.z_dim)        # one global draw, sliced per rank
    x0_h = x0_all[rank * B:(rank + 1) * B].contiguous().pin_memory()
    res_h = res_all[rank * B:(rank + 1) * B].contiguous().pin_memory()
    x0_d, res_d = x0_h.to(dev), res_h.to(dev)
    q_h = make_query(img).pin_memory() if transfer else None
    q_d = q_h.to(dev) if transfer else None
    passes = -(-args.seq_length // 16)
    T = 16 * passes
    gather = FrameGather(dev, mode="p2p" if args.gather == "p2p" else "nccl") if world > 1 else None
    sink = HostFrameSink(dev)           # pinned double buffer + copy stream: read-back of step i under the compute of step i + 1
    out_numel = B * T * 3 * img * img

    def run_model(x, r, q):
        return model.transfer(q, x) if transfer else model.sample(x, residual=r)

    def finish(seq):
        if world == 1:
            return
        if args.gather == "uint8":
            mx = cli.frames_max(seq)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            gather.submit(cli.frames_to_u8(seq, mx, "video"))
        else:
            gather.submit(seq)
        if args.gather == "inline":
            gather.wait()

    def step_device():
        seq = run_model(x0_d, res_d, q_d)
        finish(seq)
        return seq

    def step_e2e():
        x = x0_h.to(dev, non_blocking=True)
        r = res_h.to(dev, non_blocking=True)
        q = q_h.to(dev, non_blocking=True) if transfer else None
        seq = run_model(x, r, q)
        finish(seq)
        sink.put(seq)
        return seq

    def timed(fn, steps, profile=False):
        barrier()
        torch.cuda.synchronize()
        n0 = L.i2v_launch_count()
        if profile:
            if args.dump_launches:
                L.i2v_prof_dump_path(args.dump_launches.encode())
            L.i2v_prof_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if gather is not None:
            gather.wait()          # the compute stream owns every gathered batch before the closing event
        torch.cuda.current_stream(dev).wait_stream(sink.stream)     # ... and every read-back has landed in host memory
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        prof = None
        if profile:
            arr = [(ctypes.c_double * 7)(), (ctypes.c_double * 7)(), (ctypes.c_double * 7)(), (ctypes.c_longlong * 7)()]
            lib.check(L.i2v_prof_collect(*arr), "prof_collect")
            L.i2v_prof_enable(0)
            prof = [list(a) for a in arr]
        launches = L.i2v_launch_count() - n0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, launches, prof

    # nvidia-smi starts BEFORE the warm-up (its NVML initialisation takes ~0.5 s and must not land in a timed leg); only the rows
    # it prints from the first timed leg on are used
    sampler = ClockSampler(local) if rank == 0 and not args.e2e_probe else None
    if sampler:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    step_e2e()
    step_e2e()                     # both pinned read-back buffers exist before the timed regions
    torch.cuda.synchronize()

    if args.e2e_probe:
        def step_h2d():
            x = x0_h.to(dev, non_blocking=True)
            r = res_h.to(dev, non_blocking=True)
            q = q_h.to(dev, non_blocking=True) if transfer else None
            return run_model(x, r, q)

        def step_d2h():
            seq = run_model(x0_d, res_d, q_d)
            sink.put(seq)
            return seq

        out = {}
        for rep in range(3):
            for name, fn in (("device", step_device), ("h2d", step_h2d), ("d2h", step_d2h), ("e2e", step_e2e)):
                out.setdefault(name, []).append(round(timed(fn, args.steps)[0] / args.steps, 3))
        if rank == 0:
            print(json.dumps({"e2e_probe_ms_per_step": out, "steps": args.steps}), flush=True)
        return

    if sampler:
        sampler.mark()
    # the timed region: K steps back to back, nothing between the launches (the value) ...
    # ... the same K steps through host buffers (e2e), then both twice more: A B A B A B.  The board sits at its power cap and the
    # clock wanders by a few % between legs, so `value` and `e2e` are the MEDIAN leg of three (every leg is listed under "ab")
    legs_dev, legs_e2e, launches = [], [], 0
    for _ in range(max(1, args.legs)):
        m, n, _ = timed(step_device, args.steps)
        launches = launches or n
        legs_dev.append(m)
        legs_e2e.append(timed(step_e2e, args.steps)[0])
    ms = sorted(legs_dev)[(len(legs_dev) - 1) // 2]
    ms_e2e = sorted(legs_e2e)[(len(legs_e2e) - 1) // 2]
    # and K steps with CUDA events around every launch: per-kernel-family device times (the roofline);
    # the events serialise the programmatic-dependent-launch overlap, so this pass is a little slower than the value
    ms_prof, _, prof = timed(step_device, args.steps, profile=True)
    clocks = sampler.stop() if sampler else None

    frames_per_step = world * B * T
    value = frames_per_step * args.steps / (ms / 1e3)
    e2e = frames_per_step * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1590.0 * 1441.7 / 1690.8)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md)"
        cms, cfl, cby, cn = prof
        names = ["conv_tc_halo", "stats", "modulate", "flow", "other", "conv_tc_pertap", "conv_simt"]
        fam = {names[i]: {"ms": cms[i] / args.steps, "launches": cn[i] / args.steps,
                          "tflops": (cfl[i] / (cms[i] * 1e-3) / 1e12) if cms[i] > 0 else 0.0,
                          "gbs": (cby[i] / (cms[i] * 1e-3) / 1e9) if cms[i] > 0 else 0.0} for i in range(7)}
        conv_fams = ["conv_tc_halo", "conv_tc_pertap", "conv_simt"]
        dom = max(conv_fams, key=lambda k: fam[k]["ms"])
        achieved = fam[dom]["tflops"]
        ncu = {}
        if args.config == "bair_b64":
            for name in ("r02_roofline_traffic.json", "r01_roofline_traffic.json"):
                try:
                    ncu = json.load(open(os.path.join(ROOT, "profiles", name)))
                    break
                except (OSError, ValueError):
                    pass
        arith = {1: "fp32-parity mode issues 3 fp16 MMAs per algorithmic MAC (hi*hi+hi*lo+lo*hi) and the phase form "
                    "skips 1/3 of conv_0's taps: tensor-pipe FLOP/s = achieved x 3 x (issued/nominal taps)",
                 2: "single fp16 product per MAC (reduced-precision decoder)", 0: "fp32 FFMA engine (no tensor cores)"}[args.conv_engine]
        kernel_names = {"conv_tc_halo": "conv_tc_pair_kernel + conv_tc_halo_kernel (the 256-voxel-tile family: CTA pairs with "
                                        "tcgen05.mma.cta_group::2 where eligible, single-CTA halo tiles otherwise)",
                        "conv_tc_pertap": "conv_tc_kernel (per-tap tiles)", "conv_simt": "conv_simt_kernel (fp32 FFMA engine)"}
        roofline = {"bound": "tensor",
                    "kernel": f"{kernel_names[dom]}: all its launches in the timed steps (decoder Conv3d/Conv2d stack), algorithmic FLOPs = "
                              "2*taps*Cin*Cout per output voxel (reference's nominal count, SURVEY 8d) / summed CUDA-event time",
                    "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "peak_source": peak_src, "note": arith,
                    "traffic": ncu.get("dram_bytes_per_launch"), "traffic_note": ncu.get("note"),
                    "share_of_step": fam[dom]["ms"] / (ms_prof / args.steps), "profiled_ms_per_step": ms_prof / args.steps,
                    "step_tflops": sum(cfl) / args.steps / (ms / args.steps * 1e-3) / 1e12,
                    "families": fam}
        h2d = world * (x0_h.numel() + res_h.numel() + (q_h.numel() if transfer else 0)) * 4
        dtype = {1: "f32", 2: "f16", 0: "f32"}[args.conv_engine]
        cfg = workload_config(args, img, world)
        cfg.update({"arithmetic": {1: "fp32-grade: fp16 hi/lo operand split, 3 tensor-core products per MAC, fp32 accumulate",
                                   2: "fp16 operands, one tensor-core product per MAC, fp32 accumulate (flow/embedder fp32-grade)",
                                   0: "fp32 FFMA"}[args.conv_engine],
                    "conv_engine": args.conv_engine, "micro_batch": args.micro_batch, "streams": args.streams,
                    "gather": (args.gather if args.gather != "p2p" else gather.mode) if world > 1 else None,
                    "gather_fallback": getattr(gather, "fallback_reason", None) if world > 1 else None,
                    "graph": args.graph, "options": args.opt})
        line = {
            "metric": args.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic start frames U[-1,1], random-init weights in the reference checkpoint format",
            "config": cfg,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": world * out_numel * 4,
                    "note": "H2D of the step's inputs on the compute stream; D2H of its frames on a copy stream behind the next step's "
                            "kernels (dist.HostFrameSink), all copies complete inside the timed region"},
            "ab": {"device_ms_per_step": [m / args.steps for m in legs_dev],
                   "e2e_ms_per_step": [m / args.steps for m in legs_e2e],
                   "note": "legs of K steps each, alternating; value / e2e = the median leg"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            fps, cores, sample, b1 = cpu_reference_rate(args, mp)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "b1_value": b1}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if world > 1:
            time.sleep(1.0)       # NCCL_DEBUG=INFO: let the other ranks' closing lines come first, the JSON line stays last
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    if os.environ.get("I2V_BENCH_WATCHDOG"):      # debugging aid: dump every thread's Python stack after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["I2V_BENCH_WATCHDOG"]), repeat=False, file=sys.stderr)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "reference-gpu":
        run_reference_gpu(a)
    else:
        run_b200(a)
