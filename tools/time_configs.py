"""Time BASELINE configs 1-4 on the GPU path (config 5: bench.py --workload large) and stress the cluster kernels."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import wavetorch_b200 as wt
from wavetorch_b200 import _lib
from oracle import wave_oracle as wo
from test_gpu_parity import _vowel_model, _lens_model
def tm(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def report(name, m, x, labels, cells, Nx, Ny, nl=(0, 0, 0)):
    p = _lib.make_problem(Nx, Ny, x.shape[0], x.shape[1], 1, 3, 1.0, 1.0, *nl, flags=_lib.WT_F_ZERO_INIT)
    plan = _lib.query_plan(p)
    def fwd():
        with torch.no_grad(): m(x)
    def full():
        o = m(x); torch.nn.functional.cross_entropy(wt.utils.normalize_power(o.sum(1)), labels).backward(); m.zero_grad(set_to_none=True)
    tf, tb = tm(fwd), tm(full)
    print(f"| {name} | {'on-chip' if plan.path else 'stream'} C={plan.cluster} R={plan.rows_per_thread} | {tf:.3f} | {cells/tf/1e6:.1f} | {tb:.3f} | {cells/tb/1e6:.1f} |", flush=True)
def report_graphed(name, m, x, labels, cells):
    """The whole training iteration (forward, loss head, adjoint, Adam, constrain) replayed from a CUDA graph: what is left when
    the host launch latency of the ~60 small launches is taken out (it dominates the eager numbers of the small configs)."""
    from wavetorch_b200.graph import GraphedTrainStep
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True)
    step = GraphedTrainStep(m, opt, lambda o, y: torch.nn.functional.cross_entropy(wt.utils.normalize_power(o.sum(1)), y), x, labels)
    t = tm(lambda: step(x, labels), n=20)
    print(f"| {name}, training iteration from a CUDA graph | | | | {t:.3f} | {cells/t/1e6:.1f} |", flush=True)
print("| config | plan | fwd ms | fwd Gcell/s | fwd+bwd ms | fwd+bwd Gcell/s |\n|---|---|---|---|---|---|")
m = _lens_model(0.5); x = torch.tensor(wo.propagate_waveform(500), device="cuda")
report("1/2 lens 151x151 B=1 T=500", m, x, torch.tensor([2], device="cuda"), 151 * 151 * 500, 151, 151)
report_graphed("1/2 lens 151x151 B=1 T=500", _lens_model(0.5), x, torch.tensor([2], device="cuda"), 151 * 151 * 500)
if os.environ.get("ONLY_SMALL"):
    m = _vowel_model(); x = torch.tensor(wo.synthetic_vowels(8, 1000), device="cuda")
    report_graphed("3 vowel 150x100 B=8 T=1000", m, x, torch.arange(8, device="cuda") % 3, 8 * 1000 * 15000)
    sys.exit(0)
for B, T in ((64, 1000), (8, 1000), (64, 5469)):
    m = _vowel_model(); x = torch.tensor(wo.synthetic_vowels(B, T), device="cuda")
    report(f"3 vowel 150x100 B={B} T={T}", m, x, torch.arange(B, device="cuda") % 3, B * T * 15000, 150, 100)
m = _vowel_model(); m.checkpoint_every = 128; x = torch.tensor(wo.synthetic_vowels(64, 1000), device="cuda")
report("3 vowel B=64 T=1000, on-chip checkpoints every 128", m, x, torch.arange(64, device="cuda") % 3, 64 * 1000 * 15000, 150, 100)
m = _vowel_model(); m.checkpoint_every = 256; x = torch.tensor(wo.synthetic_vowels(64, 5469), device="cuda")
report("3 vowel B=64 T=5469, on-chip checkpoints every 256", m, x, torch.arange(64, device="cuda") % 3, 64 * 5469 * 15000, 150, 100)
m = _vowel_model(0.1, 1.0, 0.0); m.checkpoint_every = 256; x = torch.tensor(wo.synthetic_vowels(64, 3000), device="cuda")
report("4(i) satdamp B=64 T=3000, on-chip checkpoints every 256", m, x, torch.arange(64, device="cuda") % 3, 64 * 3000 * 15000, 150, 100, (0.1, 1.0, 0.0))
for name, nl, T in (("4(i) satdamp b0=.1 uth=1", (0.1, 1.0, 0.0), 3000), ("4(ii) satdamp+kerr", (0.1, 1.0, -30.0), 1000), ("4(iii) satdamp uth=1.8e-4", (0.1, 0.00018, 0.0), 1000)):
    B = 64
    m = _vowel_model(*nl); x = torch.tensor(wo.synthetic_vowels(B, T), device="cuda")
    report(f"{name} B={B} T={T}", m, x, torch.arange(B, device="cuda") % 3, B * T * 15000, 150, 100, nl)
# stress: many short runs over cluster sizes (intermittent-hang detector)
t0 = time.time(); n = 0
x = torch.tensor(wo.synthetic_vowels(5, 96), device="cuda"); lab = torch.arange(5, device="cuda") % 3
ref = None
for rep in range(40):
    for C, R in ((2, 5), (4, 4), (8, 2), (8, 1), (4, 2)):
        m = _vowel_model(); m.cluster, m.rows_per_thread, m.plan_flags = C, R, _lib.WT_F_FORCE_RESIDENT
        o = m(x); (o.sum()).backward(); n += 1
        g = m.cell.geom.rho.grad
        if rep == 0 and C == 2: ref = (o.detach().clone(), g.clone())
torch.cuda.synchronize()
print(f"stress: {n} fwd+bwd runs over 5 decompositions in {time.time()-t0:.1f} s, no hang", flush=True)
