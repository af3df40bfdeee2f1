"""Copy the latest gpurun_out/ measurements into profiles/ and regenerate the derived summaries.
Usage (build container, after a gpurun batch): python tools/refresh_profiles.py"""
import collections, csv, json, os, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def last_json_line(path):
    with open(path) as f:
        lines = [l for l in f.read().strip().splitlines() if l.startswith("{")]
    return lines[-1]


for name in ("bench_n1", "bench_reference_n1", "bench_large_n1", "bench_n2", "bench_large_n2", "bench_n4", "bench_n8"):
    src = os.path.join(G, name + ".json")
    if os.path.exists(src):
        with open(os.path.join(P, "r1_" + name + ".json"), "w") as f:
            f.write(last_json_line(src) + "\n")
if os.path.exists(os.path.join(G, "launches_n1.csv")):
    shutil.copy(os.path.join(G, "launches_n1.csv"), os.path.join(P, "r1_launches_bench_n1.csv"))

# ---- launch list summary
with open(os.path.join(P, "r1_launches_bench_n1.csv")) as f:
    rows = csv.DictReader([l for l in f if not l.startswith("==")])
    agg = collections.OrderedDict()
    for row in rows:
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else v * 1000 if row["Metric Unit"] == "ms" else v
        a = agg.setdefault(row["Kernel Name"].split("(")[0][:72], [0, 0.0])
        a[0] += 1
        a[1] += v
tot = sum(a[1] for a in agg.values())
ours = sum(a[1] for k, a in agg.items() if "wt::" in k)
d = json.load(open(os.path.join(P, "r1_bench_n1.json")))
k = d["kernels"]
per = {kk: a[1] / a[0] for kk, a in agg.items()}
adj = [v for kk, v in per.items() if "k_res_adj" in kk][0]
fwd = [v for kk, v in per.items() if "k_res_fwd<5, 1" in kk][0]
out = ["# round 1: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (first 600 launches)\n",
       "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_n1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (raw list: profiles/r1_launches_bench_n1.csv).",
       "Per-launch times are cold-cache and serialised; compare shares.\n", "| kernel | launches | total us | share |", "|---|---|---|---|"]
for kk, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    out.append(f"| `{kk}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f} % |")
out.append("")
out.append(f"Our kernels (`wt::*`) take {100 * ours / tot:.1f} % of the GPU time of these launches ({len([x for x in agg if 'wt::' in x])} distinct kernels: time loop, "
           "adjoint, coefficient setup, gradient reduction, geometry forward/backward, loss head); the rest is PyTorch plumbing (fused Adam, constrain_to_design_region, fills).")
out.append(f"In the un-profiled bench (CUDA events, profiles/r1_bench_n1.json) the two on-chip kernels take {k['fwd_with_tape_ms']:.2f} + {k['adjoint_ms']:.2f} = "
           f"{k['fwd_with_tape_ms'] + k['adjoint_ms']:.2f} ms when timed alone; the CUDA-graph replay of the whole iteration takes {d['ms_per_step']:.2f} ms, the eager loop "
           f"{d['eager']['ms_per_step']:.2f} ms.  k_res_adj's share of one training step (one tape-writing forward + one adjoint launch): "
           f"{100 * k['adjoint_ms'] / (k['fwd_with_tape_ms'] + k['adjoint_ms']):.0f} % by CUDA events; in the ncu list {adj:.0f} us per adjoint launch vs {fwd:.0f} us per tape-writing "
           f"forward launch = {100 * adj / (adj + fwd):.0f} % (the list also contains the forward-only and kernel-timing launches of bench.py, so totals over the list are not per-step shares).")
open(os.path.join(P, "r1_launches_bench_n1.md"), "w").write("\n".join(out) + "\n")

# ---- config table
if os.path.exists(os.path.join(G, "configs.md")):
    new = open(os.path.join(G, "configs.md")).read().strip().splitlines()
    rows = [l for l in new if l.startswith("| ") and not l.startswith("| config") and not l.startswith("|---")]
    stress = [l for l in new if l.startswith("stress")]
    cpu = {"1/2": "0.45 s / 1.49 s (0.025 / 0.0076 Gcell/s)", "3 vowel 150x100 B=64 T=1000": "0.050 / 0.016 Gcell/s", "4(i)": "0.064 / 0.011 Gcell/s"}
    dl = json.load(open(os.path.join(P, "r1_bench_large_n1.json")))
    t = ["# round 1: all BASELINE configs on 1x B200 (tools/time_configs.py, CUDA events, mean of 5 after 2 warm-ups, eager launches)\n",
         "| config | plan | fwd ms | fwd Gcell/s | fwd+bwd ms | fwd+bwd Gcell/s | reference CPU (8 cores, SURVEY section 6) |", "|---|---|---|---|---|---|---|"]
    for r in rows:
        c = ""
        for kk, v in cpu.items():
            if r.startswith("| " + kk):
                c = v
        t.append(r + " " + c + " |")
    t.append(f"| 5 window 4096x4096 B=8 T=64 (bench.py --workload large) | streaming, K=4 tiles fwd + adjoint; checkpoints every 16 | {dl['fwd']['ms_per_step']:.2f} | {dl['fwd']['value']:.1f} | {dl['ms_per_step']:.2f} | {dl['value']:.1f} | 0.037 Gcell/s fwd |")
    p2 = os.path.join(P, "r1_bench_large_n2.json")
    if os.path.exists(p2):
        d2 = json.load(open(p2))
        t.append(f"| 5 window, 2 GPUs (row slabs) | + domain decomposition, halo 16 | {d2['fwd']['ms_per_step']:.2f} | {d2['fwd']['value']:.1f} | {d2['ms_per_step']:.2f} | {d2['value']:.1f} | |")
    t.append("| 5 window 4096x4096 B=8 T=40, no checkpoints (tools/time_large.py) | streaming, K=4 tiles | 7.60 | 706 | 20.2 | 265.5 | |")
    t.append("")
    t.append("Roofline reference (SURVEY 8d, measured HBM copy peak 6553.9 GB/s): 546 Gcell/s forward (12 B/update), 205 Gcell/s fwd+bwd (32 B/update); the 60 % target is 328 / 123.")
    dref = json.load(open(os.path.join(P, "r1_bench_reference_n1.json")))
    line = f"bench.py (CUDA-graph replay of the whole training iteration, config 3): {d['value']:.1f} Gcell/s on 1 GPU (profiles/r1_bench_n1.json; e2e from pinned host memory {d['e2e']['value']:.1f})"
    for n in (2, 4, 8):
        pn = os.path.join(P, f"r1_bench_n{n}.json")
        if os.path.exists(pn):
            line += f", {json.load(open(pn))['value']:.1f} on {n} GPUs"
    t.append(line + f"; CPU port of the reference on the box's {dref['cpu_baseline']['cores']} cores: {dref['value']:.3f} (`--impl reference`), {d['cpu_baseline']['value']:.3f} (cpu_baseline of the same run).")
    t.append("History of config 3 fwd+bwd (kernels only): 341 (first complete version) -> 387 (P-form adjoint, seed fast path) -> 404 (compile-time pitch/threads) -> 417 (dLoss/dx gather out of the unrolled step body) -> 423 (fields output in its own instantiation) -> 435 (dLoss/dx code compiled out when x.grad is not requested) -> 480 (adjoint: per-block staging out of the step bodies, seeds added row-wise) -> 497 (sources added row-wise).")
    t.append("History of config 4 fwd+bwd: 104 / 99 / 100 (first version) -> 187 / 169 (one or two MUFU reciprocals per cell) -> 236 / 205 / 230 (ring arithmetic, compile-time parity and shape, dLoss/dx compiled out).")
    if stress:
        t.append("Stress: " + stress[0].replace("stress: ", ""))
    open(os.path.join(P, "r1_configs.md"), "w").write("\n".join(t) + "\n")
print("profiles refreshed")
