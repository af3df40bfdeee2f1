"""Planner choices for the grids of the reference's study configs (which (R, pitch, threads) instantiations matter)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wavetorch_b200 import _lib
shapes = [("example.yml / example_nonlinearity.yml", 150, 100), ("satdamp.yml / nonlinear_speed.yml", 160, 100),
          ("linear.yml", 140, 140), ("propagate.py / optimize_lens.py", 151, 151)]
print("| study config | grid | B | nl | path | C | R | threads | rows/CTA | clusters | ring | smem fwd/bwd |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for name, Nx, Ny in shapes:
    for B in (1, 8, 16, 32, 64, 128):
        for nl in ((0.0, 0.0, 0.0), (0.1, 1.0, -30.0)):
            p = _lib.make_problem(Nx, Ny, B, 1000, 1, 3, 1.0, 1.4283556979968262, *nl, flags=_lib.WT_F_ZERO_INIT)
            pl = _lib.query_plan(p)
            print(f"| {name} | {Nx}x{Ny} | {B} | {int(nl[0] > 0) + 2 * int(nl[2] != 0)} | {'on-chip' if pl.path else 'stream'} | {pl.cluster} | "
                  f"{pl.rows_per_thread} | {pl.threads} | {pl.rows_per_cta} | {pl.n_clusters} | {pl.reserved[0]} | {pl.smem_fwd}/{pl.smem_bwd} |")
