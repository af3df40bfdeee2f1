#!/usr/bin/env python
"""Benchmark of the wave-RNN hot path (BASELINE.json metric: Gcell-updates/s = B*Nx*Ny*T / s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

Workload (config.workload = "vowel64"): BASELINE config 3 -- study/example.yml geometry, 150x100 grid, 3 intensity
probes, synthetic vowel-length waveforms, batch 64 PER GPU (weak scaling), T = 1000, float32.  One "step" is one
training iteration of wavetorch/train.py:59-72: forward, loss = CrossEntropy(normalize_power(sum_t I)), backward
(adjoint kernel), all-reduce of the loop gradient over ranks, Adam step on rho, constrain_to_design_region.

Prints ONE JSON line (rank 0).  Keys beyond the base contract: fwd (forward-only throughput), roofline (dominant
kernel vs the measured HBM copy bandwidth), cpu_baseline (oracle/torch_port.py on the host cores, N=1 only),
e2e (same step fed from pinned host memory with a device->host read of the loss every step).
"""
import argparse
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
    ap.add_argument("--workload", default="vowel64", choices=["vowel64", "large"],
                    help="vowel64 = BASELINE config 3 (the headline); large = config 5 window (4096^2, domain-decomposed for N>1)")
    ap.add_argument("--grid", type=int, default=4096, help="large workload: grid edge")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-T", type=int, default=0, help="time steps of the bounded CPU sample (0 = auto)")
    return ap.parse_args()


def config_dict(args, world):
    return {"workload": "vowel64", "source": "study/example.yml (BASELINE config 3)", "grid": [NX, NY],
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "time_steps": args.T, "probes": 3,
            "step": "fwd + CE(normalize_power(sum_t I)) + adjoint + grad all-reduce + Adam + constrain",
            "parallelism": "batch-sharded x%d" % world,
            "l2": "no explicit flush: every step writes and re-reads a %.2f GB adjoint tape (>> 126 MB L2)"
                  % (args.batch * args.T * NX * NY * 4 / 1e9)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference algorithm on the host cores (oracle/torch_port.py)
# --------------------------------------------------------------------------------------------------
def cpu_training_throughput(B, T, repeats, warm=True):
    import numpy as np
    import torch
    from oracle import torch_port as tp
    from oracle import wave_oracle as wo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = wo.vowel_config(np.float32, NX, NY)
    if warm:
        tp.time_cpu(cfg, min(B, 8), 10, True)
    times = []
    for _ in range(repeats):
        s, cells = tp.time_cpu(cfg, B, T, True)
        times.append(s)
    return cells, times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.batch
    Tc = args.cpu_T or 100
    cells, _, cores = cpu_training_throughput(B, Tc, max(args.warmup, 1), warm=True)
    t0 = time.perf_counter()
    _, times, _ = cpu_training_throughput(B, Tc, args.steps, warm=False)
    wall = time.perf_counter() - t0
    ms = 1e3 * sum(times) / len(times)
    val = cells / (ms * 1e-3) / 1e9
    sample = "B=%d, T=%d of %d steps per timed step (throughput per cell-update does not depend on T)" % (B, Tc, args.T)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(args, 1),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall,
            "note": "PyTorch-CPU port of the reference loop (oracle/torch_port.py): the reference is pure Python and "
                    "does not travel to the GPU box; the port issues the same ATen ops per step"}
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
    geom = wt.WaveGeometryFreeForm((NX, NY), 1.4283556979968262, c0=1.0, c1=0.5, eta=0.5, beta=100, abs_sig=3.0,
                                   abs_N=N, abs_p=4.0, rho="half", blur_radius=1, blur_N=1, design_region=design)
    return wt.WaveRNN(wt.WaveCell(1.0, geom), [src], probes).to(dev)


def peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import wavetorch_b200 as wt
    from wavetorch_b200 import _lib
    from wavetorch_b200.distributed import BatchShardedWaveRNN
    from oracle import wave_oracle as wo   # synthetic input generator only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, T = args.batch, args.T
    cells_per_step = B * T * NX * NY

    model = build_model(dev)
    runner = BatchShardedWaveRNN(model, average=True) if world > 1 else model
    opt = torch.optim.Adam(model.parameters(), lr=4e-4, fused=True)    # torch's single-kernel Adam: same update rule
    x_host = torch.tensor(wo.synthetic_vowels(B, T, first=rank * B)).pin_memory()
    labels_host = ((torch.arange(B) + rank * B) % 3).pin_memory()
    x_dev, labels = x_host.to(dev), labels_host.to(dev)

    def loss_head(out, y):
        """train.py:61-62 -- CrossEntropy(normalize_power(sum_t I), y) -- as the fused head (wavetorch_b200/loss.py)."""
        return wt.power_cross_entropy(out, y)[0]

    def train_step(x, y):
        opt.zero_grad(set_to_none=True)
        out = runner(x)
        loss = loss_head(out, y)
        loss.backward()
        opt.step()
        model.cell.geom.constrain_to_design_region()
        return loss

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm=0):
        """CUDA-event time of `steps` calls, max over ranks; returns ms per step."""
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        train_step(x_dev, labels)

    # The iteration is replayed from a CUDA graph (wavetorch_b200.graph.GraphedTrainStep, part of the public API):
    # same kernels, no host launch latency between them.  If capture is not possible the eager loop is timed.
    graphed, mode = None, "eager"
    if os.environ.get("WT_BENCH_EAGER", "0") != "1":
        try:
            from wavetorch_b200.graph import GraphedTrainStep
            opt_g = torch.optim.Adam(model.parameters(), lr=4e-4, capturable=True, fused=True)
            graphed = GraphedTrainStep(
                runner, opt_g, loss_head,
                x_dev, labels, warmup=max(args.warmup, 3))
            graphed(x_dev, labels)
            mode = "cuda-graph"
        except Exception as exc:  # pragma: no cover
            graphed = None
            sys.stderr.write("bench: CUDA-graph capture failed (%r); timing the eager loop\n" % (exc,))

    def step_resident():
        return graphed(x_dev, labels) if graphed is not None else train_step(x_dev, labels)

    def step_e2e():      # inputs from pinned host memory, loss back to the host, every step
        if graphed is not None:
            return graphed(x_host, labels_host).item()
        xb = x_host.to(dev, non_blocking=True)
        yb = labels_host.to(dev, non_blocking=True)
        return train_step(xb, yb).item()

    l0 = _lib.launch_count
    n_launch_eager = None
    t_wall = time.perf_counter()
    ms_step = timed(step_resident, args.steps, warm=1)
    t_wall = time.perf_counter() - t_wall
    launches = _lib.launch_count - l0
    ms_e2e = timed(step_e2e, args.steps, warm=2)
    # the eager loop, for reference (and to count the launches one iteration makes)
    l0 = _lib.launch_count
    ms_eager = timed(lambda: train_step(x_dev, labels), args.steps, warm=1)
    launches_eager = (_lib.launch_count - l0) // (args.steps + 1) * args.steps
    if graphed is not None:
        launches = launches_eager      # a replay launches the same kernels as the iteration it captured

    # forward only (inference, no tape)
    def fwd_only():
        with torch.no_grad():
            return runner(x_dev) if world == 1 else model(x_dev)

    for _ in range(3):
        fwd_only()
    ms_fwd = timed(fwd_only, args.steps)

    # dominant kernels, timed alone with CUDA events on the launching stream
    out = model(x_dev)
    loss = loss_head(out, labels)
    (gout,) = torch.autograd.grad(loss, out, retain_graph=True)

    def fwd_tape():
        return model(x_dev)

    ms_fwd_tape = timed(fwd_tape, max(3, args.steps // 2), warm=2)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    bw_ms = []
    for _ in range(max(3, args.steps // 2)):
        o = model(x_dev)
        torch.cuda.synchronize()
        ev[0].record()
        o.backward(gout)
        ev[1].record()
        torch.cuda.synchronize()
        bw_ms.append(ev[0].elapsed_time(ev[1]))
        model.zero_grad(set_to_none=True)
    ms_bwd = statistics.median(bw_ms)
    clocks = sampler.stop() if rank == 0 else None

    def teardown():
        # Drop the captured graph (it holds NCCL work when world > 1) before leaving; with graphs alive
        # destroy_process_group() was seen to hang, so multi-rank runs synchronise and exit without it.
        nonlocal graphed
        graphed = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        teardown()
        return

    peak, peak_src = peak_hbm()
    traffic = ncu_traffic()
    dom_is_bwd = ms_bwd >= ms_fwd_tape
    dom_ms = ms_bwd if dom_is_bwd else ms_fwd_tape
    alg_bytes = 16.0 * cells_per_step        # fwd: read 2 write 1 field + tape write; adjoint: the same in reverse
    # DRAM bytes of the same kernel from the committed ncu --set full capture, scaled by cell updates if the shape differs
    tr = traffic.get("k_res_adj" if dom_is_bwd else "k_res_fwd")
    traffic_bytes = int(tr["dram_bytes_per_cell_update"] * cells_per_step) if tr else None
    traffic_src = (tr["source"] + ", captured at " + tr["shape"]) if tr else None
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    p = _lib.make_problem(NX, NY, B, T, 1, 3, 1.0, 1.4283556979968262, device=local)
    plan = _lib.query_plan(p)
    value = world * cells_per_step / (ms_step * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": dict(config_dict(args, world), grad_allreduce=(
            getattr(runner, "reduce_mode", "none") if world > 1 else "none")), "clocks": clocks,
        "e2e": {"value": world * cells_per_step / (ms_e2e * 1e-3) / 1e9, "unit": UNIT,
                "h2d_bytes_per_step": int(x_host.numel() * 4 + labels_host.numel() * 8), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e, "mode": mode},
        "eager": {"value": world * cells_per_step / (ms_eager * 1e-3) / 1e9, "ms_per_step": ms_eager,
                  "what": "same iteration launched from Python without graph capture, inputs resident"},
        "mode": mode,
        "gpu_launches": launches,
        "fwd": {"value": world * cells_per_step / (ms_fwd * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_fwd,
                "what": "forward only (torch.no_grad, no tape), same workload"},
        "kernels": {"fwd_with_tape_ms": ms_fwd_tape, "adjoint_ms": ms_bwd, "fwd_no_tape_ms": ms_fwd,
                    "plan": {"path": "resident" if plan.path == 1 else "stream", "cluster": plan.cluster,
                             "rows_per_thread": plan.rows_per_thread, "threads": plan.threads,
                             "clusters": plan.n_clusters, "smem_fwd": plan.smem_fwd, "smem_bwd": plan.smem_bwd}},
        "roofline": {"bound": "hbm", "kernel": "k_res_adj" if dom_is_bwd else "k_res_fwd", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_cell_update": 16.0, "ms_per_launch": dom_ms,
                     "traffic": traffic_bytes, "traffic_unit": "bytes per launch (dram read + write)",
                     "traffic_source": traffic_src,
                     "note": "fields stay on-chip: algorithmic bytes (3 field passes + 1 tape pass per cell update) "
                             "are what a non-fused implementation must move; see traffic for the real DRAM bytes"},
        "roofline_fwd_only": {"achieved": 12.0 * cells_per_step / (ms_fwd * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": 12.0 * cells_per_step / (ms_fwd * 1e-3) / 1e9 / peak,
                              "algorithmic_bytes_per_cell_update": 12.0},
        "wall_s_timed_region": t_wall,
    }
    if world == 1 and not args.no_cpu_baseline:
        Tc = args.cpu_T or 100
        cells, times, cores = cpu_training_throughput(B, Tc, 2)
        line["cpu_baseline"] = {"value": cells / min(times) / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "B=%d, T=%d of %d steps, fwd+bwd, best of 2 (oracle/torch_port.py)" % (B, Tc, T)}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    teardown()


def run_large(args):
    """BASELINE config 5 window: 4096x4096 grid, batch 8, T-step window, fwd + adjoint with checkpoints every 16 steps;
    N > 1: ONE simulation split by rows over the GPUs (halo exchange every 16 steps), strong scaling."""
    import math
    import torch
    import torch.distributed as dist
    import wavetorch_b200 as wt
    from wavetorch_b200 import _lib
    from wavetorch_b200.domain import DomainDecomposedWaveRNN
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N, B, T, S = args.grid, (args.batch if args.batch != BATCH else 8), (args.T if args.T != T_STEPS else 64), 16
    ii = torch.arange(N, dtype=torch.float32)[:, None]
    jj = torch.arange(N, dtype=torch.float32)[None, :]
    rho = 0.5 + 0.5 * torch.sin(2 * math.pi * ii / 97) * torch.cos(2 * math.pi * jj / 61)
    geom = wt.WaveGeometryFreeForm((N, N), 1.4283556979968262, 1.0, 0.5, abs_N=20, abs_sig=3.0, abs_p=4.0, rho=rho)
    probes = [wt.WaveIntensityProbe(N - 60, N // 2 + 20 * k) for k in (-1, 0, 1)]
    model = wt.WaveRNN(wt.WaveCell(1.0, geom), [wt.WaveSource(60, N // 2)], probes).to(dev)
    model.checkpoint_every = S
    runner = DomainDecomposedWaveRNN(model, halo=S) if world > 1 else model
    torch.manual_seed(0)
    x = (0.1 * torch.randn(B, T)).to(dev)
    w = torch.randn(B, T, 3).to(dev)
    cells = B * T * N * N

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    def train():
        out = runner(x)
        (out * w).sum().backward()
        model.zero_grad(set_to_none=True)

    def fwd():
        with torch.no_grad():
            runner(x)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        train()
    l0 = _lib.launch_count
    ms = timed(train, args.steps)
    launches = _lib.launch_count - l0
    fwd()
    ms_f = timed(fwd, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        peak, peak_src = peak_hbm()
        # unfused algorithmic bytes: fwd 12 B, recompute with tape 16 B, adjoint 16 B per cell update
        line = {"metric": METRIC, "value": cells / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "large", "source": "BASELINE config 5 window", "grid": [N, N], "global_batch": B,
                           "time_steps": T, "checkpoint_every": S,
                           "parallelism": ("row-slab domain decomposition x%d, halo %d" % (world, S)) if world > 1 else "1 GPU",
                           "l2": "fields are %.1f GiB per time level (>> 126 MB L2)" % (B * N * N * 4 / 2 ** 30)},
                "clocks": clocks, "gpu_launches": launches,
                "fwd": {"value": cells / (ms_f * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_f},
                "roofline": {"bound": "hbm", "kernel": "k_tile_fwd (forward only)", "achieved": 12.0 * cells / (ms_f * 1e-3) / 1e9 / world,
                             "peak": peak, "unit": "GB/s", "frac": 12.0 * cells / (ms_f * 1e-3) / 1e9 / world / peak,
                             "peak_source": peak_src, "algorithmic_bytes_per_cell_update": 12.0, "traffic": None,
                             "note": "per GPU; temporal blocking moves fewer bytes than the 12 B/update of an unblocked sweep"},
                "e2e": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "large":
        run_large(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
