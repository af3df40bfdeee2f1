#!/usr/bin/env python
"""Benchmark of the wave-RNN hot path (BASELINE.json metric: Gcell-updates/s = B*Nx*Ny*T / s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the UNMODIFIED reference (baseline/_ref) on the host cores

Workload (config.workload = "vowel64"): BASELINE config 3 -- study/example.yml geometry, 150x100 grid, 3 intensity
probes, synthetic vowel-length waveforms, batch 64 PER GPU (weak scaling: the headline `value`), T = 1000, float32.
One "step" is one training iteration of wavetorch/train.py:59-72: forward, loss = CrossEntropy(normalize_power(sum_t I)),
backward (adjoint kernel), all-reduce of the loop gradient over ranks, Adam step on rho, constrain_to_design_region.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
    fwd            forward-only throughput
    roofline       the dominant time-loop call (wt_forward with tape / wt_backward, timed alone through the C ABI with CUDA
                   events) against the measured HBM copy bandwidth: `frac` in ALGORITHMIC bytes of an unfused implementation
                   (the fields never leave the SM, so it can exceed 1) and `frac_dram` in the bytes the kernel
                   really moves (the adjoint tape: plan.history_bytes)
    strong         BASELINE config 3 as stated: global batch 64 split over the N GPUs (64/N waveforms each)
    rank_agreement N > 1: rho is bitwise identical on all ranks after the timed loops
    config5_window BASELINE config 5 window (4096^2): one simulation split by rows over the N GPUs (csrc/wt_slab.cu)
    cpu_baseline   the reference on the host cores (N = 1 only), bounded sample
    e2e            the same step fed from pinned host memory with a device->host read of the loss every step
"""
import argparse
import hashlib
import json
import os
import statistics
import sys
import threading
import time

# stdout carries exactly one JSON line: NCCL's version banner and warnings (printed to stdout whenever NCCL_DEBUG is set)
# go to stderr instead
# (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, so VERSION is raised to WARN)
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX, NY, BATCH, T_STEPS = 150, 100, 64, 1000
H_VOWEL = 1.4283556979968262
METRIC = "Gcell-updates/s, fwd+bwd training step (batch x Nx x Ny x steps / s)"
UNIT = "Gcell-updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="waveforms per GPU")
    ap.add_argument("--T", type=int, default=T_STEPS)
    ap.add_argument("--workload", default="vowel64", choices=["vowel64", "large", "config5"],
                    help="vowel64 = BASELINE config 3 (the headline); large = config 5 window (4096^2, B=8); config5 = config 5 "
                         "at its stated size (4096^2, T=10000, B=32); both domain-decomposed for N>1")
    ap.add_argument("--grid", type=int, default=4096, help="large workload: grid edge")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling and config-5-window blocks")
    ap.add_argument("--cpu-T", type=int, default=0, help="time steps of the bounded CPU sample (0 = auto)")
    return ap.parse_args()


def config_dict(args, world, T=None):
    T = T or args.T
    return {"workload": "vowel64", "source": "study/example.yml (BASELINE config 3)", "grid": [NX, NY],
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "time_steps": T, "probes": 3,
            "step": "fwd + CE(normalize_power(sum_t I)) + adjoint + grad all-reduce + Adam + constrain",
            "parallelism": "batch-sharded x%d" % world,
            "l2": "no explicit flush: every step writes and re-reads a %.2f GB adjoint tape (>> 126 MB L2)"
                  % (args.batch * T * NX * NY * 4 / 1e9)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference itself (baseline/_ref, pip-installed copy of /root/reference) on the
# host cores; oracle/torch_port.py only if that copy is missing
# --------------------------------------------------------------------------------------------------
class CpuReference:
    """One training iteration of wavetorch/train.py:59-72 on BASELINE config 3, on the host, through the reference's own
    public API (WaveGeometryFreeForm / WaveCell / WaveSource / WaveIntensityProbe / WaveRNN, torch autograd, Adam)."""

    def __init__(self):
        import torch
        from oracle import ref_loader
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        root = ref_loader.reference_root(prefer_installed=True)
        self.kind = "reference" if root else "port"
        self.where = root
        if root:
            wt = ref_loader.load_reference(root)
            N = 20
            src = wt.WaveSource(N + 20, NY // 2)
            y0 = int((NY - 40) / 2)
            probes = [wt.WaveIntensityProbe(NX - N - 20, y0 + 20 * i) for i in range(3)]
            design = torch.zeros(NX, NY, dtype=torch.uint8)
            design[src.x.item() + 5:probes[0].x.item() - 5] = 1
            geom = wt.WaveGeometryFreeForm((NX, NY), H_VOWEL, c0=1.0, c1=0.5, eta=0.5, beta=100, abs_sig=3.0, abs_N=N,
                                           abs_p=4.0, rho="half", blur_radius=1, blur_N=1, design_region=design)
            self.model = wt.WaveRNN(wt.WaveCell(1.0, geom), [src], probes)
            self.opt = torch.optim.Adam(self.model.parameters(), lr=4e-4)
            self.normalize_power = wt.utils.normalize_power

    def step(self, B, T):
        """Seconds of one iteration on B waveforms of T samples."""
        import numpy as np
        import torch
        from wavetorch_b200 import synth
        x = torch.tensor(synth.synthetic_vowels(B, T, dtype=np.float32))
        labels = torch.arange(B) % 3
        t0 = time.perf_counter()
        if self.kind == "reference":
            self.opt.zero_grad()
            pred = self.normalize_power(self.model(x).sum(dim=1))
            loss = torch.nn.functional.cross_entropy(pred, labels)
            loss.backward()
            self.opt.step()
            self.model.cell.geom.constrain_to_design_region()
        else:
            from oracle import torch_port as tp
            from oracle import wave_oracle as wo
            tp.training_step(wo.vowel_config(np.float32, NX, NY), x, labels)
        return time.perf_counter() - t0

    def describe(self):
        if self.kind == "reference":
            return ("unmodified fancompute/wavetorch 0.2.1 from baseline/_ref (pip --target install of /root/reference), its "
                    "own WaveRNN/WaveCell/TimeStep on PyTorch CPU with autograd, skimage/librosa/matplotlib imports stubbed")
        return "oracle/torch_port.py (baseline/_ref not present): the same ATen ops per step as the reference"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference()
    B = args.batch
    # full T = 1000 does not fit: the reference's autograd keeps ~15 [B,Nx,Ny] tensors per step (57 GB at B=64, T=1000),
    # and one such iteration takes about a minute.  The sample is sized from a probe so that the whole run stays within a
    # few minutes; throughput per cell update does not depend on T.
    t_probe = ref.step(B, 20)
    per_step_s = t_probe / 20
    budget = 150.0 / max(args.steps + max(args.warmup, 1), 1)
    Tc = args.cpu_T or int(max(20, min(args.T, 200, budget / per_step_s)))
    for _ in range(max(args.warmup, 1)):
        ref.step(B, Tc)
    t0 = time.perf_counter()
    times = [ref.step(B, Tc) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    cells = B * Tc * NX * NY
    ms = 1e3 * sum(times) / len(times)
    val = cells / (ms * 1e-3) / 1e9
    sample = "B=%d, T=%d of %d steps per timed step (throughput per cell-update does not depend on T)" % (B, Tc, args.T)
    cfg = config_dict(args, 1, T=Tc)
    cfg["time_steps_of_workload"] = args.T
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall, "note": ref.describe()}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU through NVML every few milliseconds while the timed
    regions run (the profiling recipe's nvidia-smi clocks line, taken in-process so that short regions get samples)."""

    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4,
               "hw_power_brake_slowdown": 0x80}

    def __init__(self, index, period_s=0.004):
        self.index, self.period, self.samples, self.stop_flag, self.thread = index, period_s, [], False, None
        self.err = None

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                try:
                    phys = int(vis.split(",")[self.index])
                except Exception:
                    phys = self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                ut = pynvml.nvmlDeviceGetUtilizationRates(h).gpu
                self.samples.append((sm, pw, rs, ut))
                time.sleep(self.period)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples: %s" % self.err]}
        busy = [s for s in self.samples if s[3] > 0] or self.samples
        bits = 0
        for s in busy:
            bits |= s[2]
        reasons = sorted(n for n, b in self.REASONS.items() if bits & b)
        return {"sm_mhz": statistics.median(s[0] for s in busy), "sm_max_mhz": float(self.max_sm),
                "power_w_max": max(s[1] for s in busy), "samples": len(self.samples), "samples_under_load": len(busy),
                "reasons": reasons, "how": "NVML polled every %d ms across all timed regions" % int(self.period * 1e3)}


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def build_model(dev):
    import torch
    import wavetorch_b200 as wt
    N = 20
    src = wt.WaveSource(N + 20, NY // 2)
    y0 = int((NY - 40) / 2)
    probes = [wt.WaveIntensityProbe(NX - N - 20, y0 + 20 * i) for i in range(3)]
    design = torch.zeros(NX, NY, dtype=torch.uint8)
    design[src.x.item() + 5:probes[0].x.item() - 5] = 1
    geom = wt.WaveGeometryFreeForm((NX, NY), H_VOWEL, c0=1.0, c1=0.5, eta=0.5, beta=100, abs_sig=3.0,
                                   abs_N=N, abs_p=4.0, rho="half", blur_radius=1, blur_N=1, design_region=design)
    return wt.WaveRNN(wt.WaveCell(1.0, geom), [src], probes).to(dev)


def build_large_model(dev, N):
    """BASELINE config 5: N x N grid, smooth deterministic rho, point source, three intensity probes."""
    import math
    import torch
    import wavetorch_b200 as wt
    ii = torch.arange(N, dtype=torch.float32)[:, None]
    jj = torch.arange(N, dtype=torch.float32)[None, :]
    rho = 0.5 + 0.5 * torch.sin(2 * math.pi * ii / 97) * torch.cos(2 * math.pi * jj / 61)
    geom = wt.WaveGeometryFreeForm((N, N), H_VOWEL, 1.0, 0.5, abs_N=20, abs_sig=3.0, abs_p=4.0, rho=rho)
    probes = [wt.WaveIntensityProbe(N - 60, N // 2 + 20 * k) for k in (-1, 0, 1)]
    return wt.WaveRNN(wt.WaveCell(1.0, geom), [wt.WaveSource(60, N // 2)], probes).to(dev)


def peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic():
    """DRAM bytes per cell update of the time-loop kernels from the committed ncu --set full captures (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class Dist:
    """World bookkeeping + device-side timing (CUDA events on the launching stream, max over ranks)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)

    def sync_all(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, warm=0):
        """CUDA-event time of `steps` calls, max over ranks; returns ms per step."""
        import torch
        import torch.distributed as dist
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.sync_all()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    def teardown(self):
        """Captured graphs hold NCCL / peer-memory work when world > 1; with graphs alive destroy_process_group() was seen
        to hang, so multi-rank runs synchronise and leave without it."""
        import gc
        import torch
        import torch.distributed as dist
        gc.collect()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)


class TrainLoop:
    """Model + optimiser + CUDA-graph replay of one training iteration on `B` waveforms per rank."""

    def __init__(self, D, B, T, first_sample, global_batch, warmup):
        import torch
        import wavetorch_b200 as wt
        from wavetorch_b200 import synth
        from wavetorch_b200.distributed import BatchShardedWaveRNN
        self.D, self.B, self.T = D, B, T
        dev = D.dev
        self.model = build_model(dev)
        # the loss is the mean over the GLOBAL batch: each rank contributes sum(CE of its samples) / global_batch and the
        # gradient all-reduce sums the ranks
        self.runner = BatchShardedWaveRNN(self.model, average=False) if D.world > 1 else self.model
        self.opt = torch.optim.Adam(self.model.parameters(), lr=4e-4, fused=True)
        self.x_host = torch.tensor(synth.synthetic_vowels(B, T, first=first_sample)).pin_memory()
        self.labels_host = ((torch.arange(B) + first_sample) % 3).pin_memory()
        self.x_dev, self.labels = self.x_host.to(dev), self.labels_host.to(dev)
        gb = global_batch
        self.loss_head = lambda out, y: wt.power_cross_entropy(out, y, gb)[0]
        for _ in range(max(warmup, 3)):
            self.train_step(self.x_dev, self.labels)
        self.graphed, self.mode = None, "eager"
        if os.environ.get("WT_BENCH_EAGER", "0") != "1":
            try:
                from wavetorch_b200.graph import GraphedTrainStep
                opt_g = torch.optim.Adam(self.model.parameters(), lr=4e-4, capturable=True, fused=True)
                self.graphed = GraphedTrainStep(self.runner, opt_g, self.loss_head, self.x_dev, self.labels,
                                                warmup=max(warmup, 3))
                self.graphed(self.x_dev, self.labels)
                self.mode = "cuda-graph"
            except Exception as exc:  # pragma: no cover
                self.graphed = None
                sys.stderr.write("bench: CUDA-graph capture failed (%r); timing the eager loop\n" % (exc,))

    def train_step(self, x, y):
        self.opt.zero_grad(set_to_none=True)
        loss = self.loss_head(self.runner(x), y)
        loss.backward()
        self.opt.step()
        self.model.cell.geom.constrain_to_design_region()
        return loss

    def step_resident(self):
        return self.graphed(self.x_dev, self.labels) if self.graphed is not None else self.train_step(self.x_dev, self.labels)

    def step_e2e(self):      # inputs from pinned host memory, loss back to the host, every step
        if self.graphed is not None:
            return self.graphed(self.x_host, self.labels_host).item()
        xb = self.x_host.to(self.D.dev, non_blocking=True)
        yb = self.labels_host.to(self.D.dev, non_blocking=True)
        return self.train_step(xb, yb).item()

    def rank_agreement(self):
        """rho bitwise identical on every rank (the gradient all-reduce returns the same bits everywhere)."""
        import torch
        import torch.distributed as dist
        if self.D.world == 1:
            return None
        rho = self.model.cell.geom.rho.detach().contiguous()
        digest = hashlib.sha256(rho.cpu().numpy().tobytes()).digest()[:8]
        mine = torch.tensor(list(digest), dtype=torch.int64, device=self.D.dev)
        allv = [torch.empty_like(mine) for _ in range(self.D.world)]
        dist.all_gather(allv, mine)
        return all(torch.equal(allv[0], v) for v in allv)

    def drop_graph(self):
        self.graphed = None


def time_loop_calls(D, model, x_dev, reps):
    """wt_forward (with tape) and wt_backward of the bench problem, each timed ALONE through the C ABI with CUDA events on
    the launching stream (no autograd, no geometry chain, no launch gaps of other kernels)."""
    import ctypes
    import torch
    from wavetorch_b200 import _lib
    lib = _lib.load()
    dev = D.dev
    B, T = x_dev.shape
    geom = model.cell.geom
    with torch.no_grad():
        c32, b32 = geom.c.detach().float().contiguous(), geom.b.detach().float().contiguous()
    tab = model._pixel_tables(dev)
    p = _lib.make_problem(NX, NY, B, T, tab["src_ij"].shape[0], tab["prb_ij"].shape[0], 1.0, H_VOWEL,
                          flags=_lib.WT_F_ZERO_INIT, device=D.local)
    plan = _lib.query_plan(p)
    n_prb = tab["prb_ij"].shape[0]
    u1, u2 = torch.empty((B, NX, NY), device=dev), torch.empty((B, NX, NY), device=dev)
    po, pr = torch.empty((B, T, n_prb), device=dev), torch.empty((B, T, n_prb), device=dev)
    gp = torch.full((B, T, n_prb), 1e-3, device=dev)
    gc = torch.empty((NX, NY), device=dev)
    ws = torch.empty(max(int(plan.workspace_fwd_bytes), int(plan.workspace_bwd_bytes), 16), dtype=torch.uint8, device=dev)
    hist = torch.empty(max(int(plan.history_bytes), 16), dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr(dev)

    def fwd():
        _lib.check(lib.wt_forward(ctypes.byref(p), _lib.ptr(c32), _lib.ptr(b32), None, _lib.ptr(x_dev), _lib.ptr(tab["src_ij"]),
                                  _lib.ptr(tab["prb_ij"]), _lib.ptr(tab["prb_sq"]), _lib.ptr(u1), _lib.ptr(u2), _lib.ptr(po),
                                  _lib.ptr(pr), None, _lib.ptr(hist), hist.numel(), _lib.ptr(ws), ws.numel(), st), "wt_forward")

    def bwd():
        _lib.check(lib.wt_backward(ctypes.byref(p), _lib.ptr(c32), _lib.ptr(b32), None, _lib.ptr(tab["src_ij"]),
                                   _lib.ptr(tab["prb_ij"]), _lib.ptr(tab["prb_sq"]), _lib.ptr(gp), _lib.ptr(pr), None,
                                   _lib.ptr(hist), hist.numel(), None, None, _lib.ptr(gc), None, None, None, _lib.ptr(ws),
                                   ws.numel(), st), "wt_backward")

    def med(fn):
        ts = []
        for _ in range(2):
            fn()
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    ms_f = med(fwd)
    ms_b = med(bwd)
    _lib.count_launches(0)
    return ms_f, ms_b, plan


def config5_window(D, args, steps):
    """BASELINE config 5 window: 4096^2, B=8, T=128, halo 16, checkpoints every 64 steps; one simulation split by rows over
    the ranks (wavetorch_b200/domain.py: in-stream NVLink peer-store ghost exchange), N = 1: the same kernels on one slab."""
    import torch
    from wavetorch_b200.domain import DomainDecomposedWaveRNN, memory_model
    N, B, T, H, S = args.grid, 8, 128, 16, 64
    dev = D.dev
    model = build_large_model(dev, N)
    model.checkpoint_every = S
    runner = DomainDecomposedWaveRNN(model, halo=H, checkpoint_every=S) if D.world > 1 else model
    torch.manual_seed(0)
    x = (0.1 * torch.randn(B, T)).to(dev)
    w = torch.randn(B, T, 3).to(dev)
    cells = B * T * N * N

    def train():
        out = runner(x)
        (out * w).sum().backward()
        model.zero_grad(set_to_none=True)

    def fwd():
        with torch.no_grad():
            runner(x)

    torch.cuda.reset_peak_memory_stats(dev)
    base = torch.cuda.memory_allocated(dev)
    train()
    train()
    ms = D.timed(train, steps)
    peak_bytes = torch.cuda.max_memory_allocated(dev) - base
    fwd()
    ms_f = D.timed(fwd, steps)
    mm = memory_model(N, N, B, T, D.world, H, S)
    res = {"workload": "BASELINE config 5 window", "grid": [N, N], "global_batch": B, "time_steps": T, "halo": H,
           "checkpoint_every": S, "scaling": "strong",
           "parallelism": ("row slabs x%d, ghost rows by in-stream NVLink peer stores every %d steps" % (D.world, H))
           if D.world > 1 else "1 GPU",
           "value": cells / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms,
           "fwd": {"value": cells / (ms_f * 1e-3) / 1e9, "ms_per_step": ms_f},
           "what": "forward + (recompute with tape + adjoint) per checkpoint segment, gradient w.r.t. rho",
           "peak_bytes_measured": int(peak_bytes), "peak_bytes_model": int(mm["total"]),
           "nvlink_bytes_per_exchange_per_neighbour": (int(runner.exchange_bytes) if D.world > 1 else 0),
           "nvlink_bytes_model": (2 * H * N * B * 4 if D.world > 1 else 0),
           "exchanges_per_step": (3 * ((T + H - 1) // H)) if D.world > 1 else 0}
    del runner, model
    return res


def run_ours(args):
    import torch
    import wavetorch_b200 as wt  # noqa: F401
    from wavetorch_b200 import _lib

    D = Dist()
    world, rank, dev = D.world, D.rank, D.dev
    B, T = args.batch, args.T
    cells_per_step = B * T * NX * NY
    sampler = ClockSampler(D.local)
    if rank == 0:
        sampler.start()

    main = TrainLoop(D, B, T, rank * B, B * world, args.warmup)
    l0 = _lib.launch_count
    t_wall = time.perf_counter()
    ms_step = D.timed(main.step_resident, args.steps, warm=1)
    t_wall = time.perf_counter() - t_wall
    ms_e2e = D.timed(main.step_e2e, args.steps, warm=2)
    # the eager loop, for reference (and to count the launches one iteration makes)
    l0 = _lib.launch_count
    ms_eager = D.timed(lambda: main.train_step(main.x_dev, main.labels), args.steps, warm=1)
    launches = (_lib.launch_count - l0) // (args.steps + 1) * args.steps   # a replay launches what the captured iteration did

    def fwd_only():
        with torch.no_grad():
            return main.model(main.x_dev)

    for _ in range(3):
        fwd_only()
    ms_fwd = D.timed(fwd_only, args.steps)
    agree = main.rank_agreement()
    ms_fwd_tape, ms_bwd, plan = time_loop_calls(D, main.model, main.x_dev, max(5, args.steps // 2))

    # ---- strong scaling: BASELINE config 3 as stated, 64 waveforms in total -------------------------------------------
    strong = None
    if not args.no_extras:
        if world == 1:
            strong = {"global_batch": B, "batch_per_gpu": B, "value": cells_per_step / (ms_step * 1e-3) / 1e9, "unit": UNIT,
                      "ms_per_step": ms_step, "note": "N = 1: the headline run"}
        elif BATCH % world == 0:
            main.drop_graph()
            Bs = BATCH // world
            sl = TrainLoop(D, Bs, T, rank * Bs, BATCH, args.warmup)
            ms_s = D.timed(sl.step_resident, args.steps, warm=1)
            ms_se = D.timed(sl.step_e2e, args.steps, warm=2)
            ps = _lib.query_plan(_lib.make_problem(NX, NY, Bs, T, 1, 3, 1.0, H_VOWEL, device=D.local))
            strong = {"global_batch": BATCH, "batch_per_gpu": Bs, "value": BATCH * T * NX * NY / (ms_s * 1e-3) / 1e9,
                      "unit": UNIT, "ms_per_step": ms_s, "e2e_value": BATCH * T * NX * NY / (ms_se * 1e-3) / 1e9,
                      "speedup_vs_64_on_one_gpu": ms_step / ms_s,
                      "speedup_note": "ms of this run's 64-per-GPU step / ms of the 64-in-total step (the 64-per-GPU step costs "
                                      "what one GPU alone needs for the global batch, plus the all-reduce)",
                      "rank_agreement": sl.rank_agreement(), "mode": sl.mode,
                      "plan": {"cluster": ps.cluster, "rows_per_thread": ps.rows_per_thread, "threads": ps.threads,
                               "clusters": ps.n_clusters, "tape_ring": ps.reserved[0]}}
            sl.drop_graph()
            del sl
    c5 = None
    if not args.no_extras:
        try:
            main.drop_graph()
            torch.cuda.empty_cache()
            c5 = config5_window(D, args, max(2, min(args.steps, 4)))
        except Exception as exc:  # pragma: no cover
            c5 = {"error": repr(exc)}
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        D.teardown()
        return

    peak, peak_src = peak_hbm()
    traffic = committed_traffic()
    dom_is_bwd = ms_bwd >= ms_fwd_tape
    dom_ms = ms_bwd if dom_is_bwd else ms_fwd_tape
    alg_bytes = 16.0 * cells_per_step        # fwd: read 2 write 1 field + tape write; adjoint: the same in reverse
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    tape_bytes = int(plan.history_bytes)
    dram = tape_bytes / (dom_ms * 1e-3) / 1e9
    tr = traffic.get("k_res_adj" if dom_is_bwd else "k_res_fwd")
    value = world * cells_per_step / (ms_step * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": dict(config_dict(args, world), grad_allreduce=(
            getattr(main.runner, "reduce_mode", "none") if world > 1 else "none")), "clocks": clocks,
        "e2e": {"value": world * cells_per_step / (ms_e2e * 1e-3) / 1e9, "unit": UNIT,
                "h2d_bytes_per_step": int(main.x_host.numel() * 4 + main.labels_host.numel() * 8), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e, "mode": main.mode},
        "eager": {"value": world * cells_per_step / (ms_eager * 1e-3) / 1e9, "ms_per_step": ms_eager,
                  "what": "same iteration launched from Python without graph capture, inputs resident"},
        "mode": main.mode,
        "gpu_launches": launches,
        "rank_agreement": agree,
        "fwd": {"value": world * cells_per_step / (ms_fwd * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_fwd,
                "what": "forward only (torch.no_grad, no tape), same workload"},
        "kernels": {"wt_forward_with_tape_ms": ms_fwd_tape, "wt_backward_ms": ms_bwd, "fwd_no_tape_ms": ms_fwd,
                    "how": "each C-ABI call timed alone with CUDA events on its stream (median); wt_backward = k_coeff + "
                           "k_res_adj + k_finish_grad_p",
                    "plan": {"path": "resident" if plan.path == 1 else "stream", "cluster": plan.cluster,
                             "rows_per_thread": plan.rows_per_thread, "threads": plan.threads,
                             "clusters": plan.n_clusters, "smem_fwd": plan.smem_fwd, "smem_bwd": plan.smem_bwd,
                             "tape_ring": plan.reserved[0]}},
        "roofline": {"bound": "hbm", "kernel": "k_res_adj (wt_backward)" if dom_is_bwd else "k_res_fwd (wt_forward with tape)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_cell_update": 16.0, "ms_per_launch": dom_ms,
                     "achieved_dram": dram, "frac_dram": dram / peak, "dram_bytes_per_launch": tape_bytes,
                     "dram_bytes_what": "the adjoint tape this launch reads (adjoint) or writes (forward): plan.history_bytes; "
                                        "fields, coefficients and gradients stay on-chip",
                     "traffic": int(tr["dram_bytes_per_cell_update"] * cells_per_step) if tr else None,
                     "traffic_unit": "bytes per launch (dram read + write)",
                     "traffic_source": ("NOT measured in this run: committed ncu --set full capture " + tr["source"] +
                                        ", taken at " + tr["shape"] + ", scaled by cell updates") if tr else None,
                     "note": "frac uses the bytes a non-fused implementation must move (3 field passes + 1 tape pass per cell "
                             "update) and exceeds 1 because the fields never leave the SM; frac_dram is the fraction of the "
                             "HBM peak the kernel really sustains"},
        "roofline_fwd_only": {"achieved": 12.0 * cells_per_step / (ms_fwd * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": 12.0 * cells_per_step / (ms_fwd * 1e-3) / 1e9 / peak,
                              "algorithmic_bytes_per_cell_update": 12.0},
        "strong": strong,
        "config5_window": c5,
        "wall_s_timed_region": t_wall,
    }
    if world == 1 and not args.no_cpu_baseline:
        ref = CpuReference()
        Tc = args.cpu_T or 100
        ref.step(min(B, 8), 10)
        times = [ref.step(B, Tc) for _ in range(2)]
        line["cpu_baseline"] = {"value": B * Tc * NX * NY / min(times) / 1e9, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                                "sample": "B=%d, T=%d of %d steps, whole training iteration, best of 2; %s"
                                          % (B, Tc, T, ref.describe())}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    D.teardown()


def run_large(args):
    """--workload large: the BASELINE config 5 window on its own line; --workload config5: config 5 at its stated size
    (4096^2, T = 10000, B = 32): checkpoints every 128 steps, batch chunks so that everything fits one GPU's HBM."""
    import torch
    from wavetorch_b200 import _lib
    from wavetorch_b200.domain import DomainDecomposedWaveRNN, memory_model
    D = Dist()
    world, rank, dev = D.world, D.rank, D.dev
    full = args.workload == "config5"
    N = args.grid
    if full:
        B, T, H, S = 32, 10000, 16, 128
        chunk = {1: 4, 2: 8, 4: 16}.get(world, 32)
    else:
        B, T, H, S, chunk = (args.batch if args.batch != BATCH else 8), (args.T if args.T != T_STEPS else 128), 16, 64, 0
    if args.T != T_STEPS:
        T = args.T
    model = build_large_model(dev, N)
    model.checkpoint_every, model.batch_chunk = S, chunk
    runner = DomainDecomposedWaveRNN(model, halo=H, checkpoint_every=S, batch_chunk=chunk) if world > 1 else model
    torch.manual_seed(0)
    x = (0.1 * torch.randn(B, T)).to(dev)
    w = torch.randn(B, T, 3).to(dev)
    cells = B * T * N * N

    def train():
        out = runner(x)
        (out * w).sum().backward()
        model.zero_grad(set_to_none=True)

    def fwd():
        with torch.no_grad():
            runner(x)

    sampler = ClockSampler(D.local)
    if rank == 0:
        sampler.start()
    steps = 1 if full else args.steps
    torch.cuda.reset_peak_memory_stats(dev)
    for _ in range(1 if full else max(args.warmup, 3)):
        train()
    l0 = _lib.launch_count
    ms = D.timed(train, steps)
    launches = _lib.launch_count - l0
    peak_bytes = torch.cuda.max_memory_allocated(dev)
    fwd()
    ms_f = D.timed(fwd, steps)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        peak, peak_src = peak_hbm()
        mm = memory_model(N, N, B, T, world, H, S, chunk)
        # unfused algorithmic bytes: fwd 12 B, recompute with tape 16 B, adjoint 16 B per cell update
        line = {"metric": METRIC, "value": cells / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": 1 if full else max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "source": "BASELINE config 5" + ("" if full else " window"), "grid": [N, N],
                           "global_batch": B, "time_steps": T, "checkpoint_every": S, "batch_chunk": chunk or B, "halo": H,
                           "parallelism": ("row-slab domain decomposition x%d, in-stream NVLink peer-store ghost exchange every "
                                           "%d steps" % (world, H)) if world > 1 else "1 GPU",
                           "l2": "fields are %.1f GiB per time level (>> 126 MB L2)" % (B * N * N * 4 / 2 ** 30)},
                "clocks": clocks, "gpu_launches": launches,
                "fwd": {"value": cells / (ms_f * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_f},
                "memory": {"peak_bytes_measured": int(peak_bytes), "peak_bytes_model": int(mm["total"]), "model": mm},
                "roofline": {"bound": "hbm", "kernel": "k_tile_fwd (forward only)", "achieved": 12.0 * cells / (ms_f * 1e-3) / 1e9 / world,
                             "peak": peak, "unit": "GB/s", "frac": 12.0 * cells / (ms_f * 1e-3) / 1e9 / world / peak,
                             "peak_source": peak_src, "algorithmic_bytes_per_cell_update": 12.0, "traffic": None,
                             "note": "per GPU; temporal blocking moves fewer bytes than the 12 B/update of an unblocked sweep"},
                "e2e": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    D.teardown()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("large", "config5"):
        run_large(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
