"""Generate tests/golden/*.npz by running the UNMODIFIED reference (fancompute/wavetorch 0.2.1).

Run in the build container only:   python oracle/gen_golden.py
It imports /root/reference through oracle/ref_loader.py (stubbed skimage/librosa/matplotlib), runs the
reference's own CPU path (WaveRNN -> WaveCell -> TimeStep, autograd for the gradients) in float32 and
float64 on deterministic inputs, and stores inputs + outputs.  The fixtures travel to the GPU box; the
reference does not.  Provenance of every array: "ref" = produced by reference code, "in" = input we fed.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402
from oracle import wave_oracle as wo  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
wt = load_reference()
torch.set_num_threads(8)


def _np(t):
    return t.detach().cpu().numpy()


def _dtype(name):
    wt.utils.set_dtype(name)
    return torch.get_default_dtype(), (np.float32 if name == "float32" else np.float64)


def _loss(model, X, labels):
    # train.py:61-62
    out = model(X)
    u = wt.utils.normalize_power(out.sum(dim=1))
    return out, torch.nn.functional.cross_entropy(u, labels)


def lens_case(name, rho_val, with_grad):
    """study/propagate.py (rho=1, fwd only) and study/optimize_lens.py (rho=0.5, fwd+bwd)."""
    res = {}
    for dname in ("float32", "float64"):
        tdt, ndt = _dtype(dname)
        import skimage
        domain = torch.zeros(151, 151)
        rr, cc = skimage.draw.circle(75, 75, 30)
        domain[rr, cc] = rho_val
        geom = wt.WaveGeometryFreeForm((151, 151), 1.0, c0=1.0, c1=0.5, rho=domain, design_region=None)
        cell = wt.WaveCell(0.707, geom)
        src = wt.WaveLineSource(25, 50, 25, 100)
        probes = [wt.WaveIntensityProbe(125, 100), wt.WaveIntensityProbe(125, 75), wt.WaveIntensityProbe(125, 50)]
        model = wt.WaveRNN(cell, src, probes)
        X = torch.tensor(wo.propagate_waveform(500, 0.707, np.float64), dtype=tdt)
        sfx = "_f32" if dname == "float32" else "_f64"
        if with_grad:
            out, loss = _loss(model, X, torch.tensor([2]))
            loss.backward()
            res["loss" + sfx] = np.asarray(loss.item())
            res["rho_grad" + sfx] = _np(geom.rho.grad)
        else:
            with torch.no_grad():
                out = model(X)
                fields = model(X, output_fields=True)
            res["maxabs_u" + sfx] = np.asarray(fields.abs().max().item())
            res["u_final" + sfx] = _np(fields[0, -1])
        res["out" + sfx] = _np(out)
        res["c" + sfx] = _np(geom.c)
        res["b" + sfx] = _np(geom.b)
        res["rho" + sfx] = _np(geom.rho)
        res["src_x"] = _np(src.x); res["src_y"] = _np(src.y)
    res["x_f64"] = wo.propagate_waveform(500, 0.707, np.float64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    print(name, {k: (v.shape if v.ndim else float(v)) for k, v in res.items() if "out" in k or "loss" in k})


def vowel_case(name, B, T, b0, uth, cnl, x_grad=False):
    """study/example.yml geometry via study/vowel_train.py:89-122, loss of train.py:61-62."""
    sys.path.insert(0, "/root/reference/study")
    from vowel_helpers import setup_src_coords, setup_probe_coords
    res = {}
    Nx, Ny, N = 150, 100, 20
    for dname in ("float32", "float64"):
        tdt, ndt = _dtype(dname)
        probes = setup_probe_coords(3, None, None, 20, Nx, Ny, N)
        source = setup_src_coords(None, None, Nx, Ny, N)
        design_region = torch.zeros(Nx, Ny, dtype=torch.uint8)
        design_region[source[0].x.item() + 5:probes[0].x.item() - 5] = 1
        geom = wt.WaveGeometryFreeForm((Nx, Ny), 1.4283556979968262, c0=1.0, c1=0.5, eta=0.5, beta=100,
                                       abs_sig=3.0, abs_N=N, abs_p=4.0, rho="half", blur_radius=1, blur_N=1,
                                       design_region=design_region)
        cell = wt.WaveCell(1.0, geom, satdamp_b0=b0, satdamp_uth=uth, c_nl=cnl)
        model = wt.WaveRNN(cell, source, probes)
        X = torch.tensor(wo.synthetic_vowels(B, T, dtype=np.float64), dtype=tdt, requires_grad=x_grad)
        labels = torch.arange(B) % 3
        out, loss = _loss(model, X, labels)
        loss.backward()
        sfx = "_f32" if dname == "float32" else "_f64"
        res["out" + sfx] = _np(out)
        res["loss" + sfx] = np.asarray(loss.item())
        res["rho_grad" + sfx] = _np(geom.rho.grad)
        if x_grad:
            res["x_grad" + sfx] = _np(X.grad)
        res["c" + sfx] = _np(geom.c)
        res["b" + sfx] = _np(geom.b)
        res["rho" + sfx] = _np(geom.rho)
        res["src_xy"] = np.array([[s.x.item(), s.y.item()] for s in source])
        res["prb_xy"] = np.array([[p.x.item(), p.y.item()] for p in probes])
        res["design_region"] = _np(design_region)
    res["x_f64"] = wo.synthetic_vowels(B, T, dtype=np.float64)
    res["params"] = np.array([b0, uth, cnl])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    print(name, float(res["loss_f32"]), float(res["loss_f64"]), float(np.linalg.norm(res["rho_grad_f64"])))


def vowel_sums_case(name, B, T):
    """Full BASELINE config-3 batch (B=64, T=1000), forward only: the per-sample probe energies sum_t I the loss head
    consumes (train.py:61), float32 and float64.  64x3 numbers pin the full-size run of the CUDA path."""
    sys.path.insert(0, "/root/reference/study")
    from vowel_helpers import setup_src_coords, setup_probe_coords
    res = {}
    Nx, Ny, N = 150, 100, 20
    for dname in ("float32", "float64"):
        tdt, ndt = _dtype(dname)
        probes = setup_probe_coords(3, None, None, 20, Nx, Ny, N)
        source = setup_src_coords(None, None, Nx, Ny, N)
        design_region = torch.zeros(Nx, Ny, dtype=torch.uint8)
        design_region[source[0].x.item() + 5:probes[0].x.item() - 5] = 1
        geom = wt.WaveGeometryFreeForm((Nx, Ny), 1.4283556979968262, c0=1.0, c1=0.5, eta=0.5, beta=100,
                                       abs_sig=3.0, abs_N=N, abs_p=4.0, rho="half", blur_radius=1, blur_N=1,
                                       design_region=design_region)
        model = wt.WaveRNN(wt.WaveCell(1.0, geom), source, probes)
        X = torch.tensor(wo.synthetic_vowels(B, T, dtype=np.float64), dtype=tdt)
        with torch.no_grad():
            out = model(X)
        sfx = "_f32" if dname == "float32" else "_f64"
        res["sums" + sfx] = _np(out.sum(dim=1))
        res["out_last" + sfx] = _np(out[:, -1])
    res["BT"] = np.array([B, T])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    print(name, res["sums_f32"][:2], res["sums_f64"][:2])


def small_case(name, b0, uth, cnl, seed, xamp=0.3):
    """Small irregular grid: mixed plain/intensity probes, two sources sharing one pixel (SURVEY B-3),
    x.grad, final fields.  Loss = weighted sum of outputs (weights stored)."""
    res = {}
    Nx, Ny, N, B, T = 27, 22, 3, 3, 48
    rng = np.random.RandomState(seed)
    rho0 = rng.rand(Nx, Ny)
    x0 = xamp * rng.randn(B, T)
    w0 = rng.randn(B, T, 4)
    for dname in ("float32", "float64"):
        tdt, ndt = _dtype(dname)
        geom = wt.WaveGeometryFreeForm((Nx, Ny), 1.2, c0=1.0, c1=0.6, eta=0.5, beta=8.0, abs_sig=2.0, abs_N=N,
                                       abs_p=2.0, rho=torch.tensor(rho0, dtype=tdt), blur_radius=1, blur_N=2,
                                       design_region=None)
        cell = wt.WaveCell(0.8, geom, satdamp_b0=b0, satdamp_uth=uth, c_nl=cnl)
        sources = [wt.WaveSource(6, 5), wt.WaveSource(6, 5), wt.WaveSource(9, 14)]
        probes_i = [(20, 4, True), (19, 11, False), (21, 17, True), (13, 10, False)]
        # the reference stacks probe outputs as given; plain = WaveProbe, intensity = WaveIntensityProbe
        probes = [wt.WaveIntensityProbe(i, j) if sq else wt.WaveProbe(i, j) for (i, j, sq) in probes_i]
        model = wt.WaveRNN(cell, sources, probes)
        X = torch.tensor(x0, dtype=tdt, requires_grad=True)
        out = model(X)
        W = torch.tensor(w0, dtype=tdt)
        loss = (out * W).sum()
        loss.backward()
        with torch.no_grad():
            fields = model(X, output_fields=True)
        sfx = "_f32" if dname == "float32" else "_f64"
        res["out" + sfx] = _np(out)
        res["loss" + sfx] = np.asarray(loss.item())
        res["rho_grad" + sfx] = _np(geom.rho.grad)
        res["x_grad" + sfx] = _np(X.grad)
        res["u_last" + sfx] = _np(fields[:, -1])
        res["u_mid" + sfx] = _np(fields[:, T // 2])
        res["c" + sfx] = _np(geom.c)
        res["b" + sfx] = _np(geom.b)
    res["rho_f64"] = rho0
    res["x_f64"] = x0
    res["w_f64"] = w0
    res["src_xy"] = np.array([[6, 5], [6, 5], [9, 14]])
    res["prb_xy"] = np.array([[p[0], p[1]] for p in probes_i])
    res["prb_intensity"] = np.array([p[2] for p in probes_i])
    res["params"] = np.array([b0, uth, cnl, 0.8, 1.2, 1.0, 0.6, 0.5, 8.0, 2.0, N, 2.0, 1, 2])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    print(name, float(res["loss_f64"]))


def single_step_case(name):
    """TimeStep.apply forward/backward on random inputs (the reference's study/utils/test_grad.py setup,
    seeded, batched and unbatched coefficients)."""
    res = {}
    from wavetorch.cell import TimeStep
    rng = np.random.RandomState(7)
    B, Nx, Ny = 3, 9, 7
    raw = dict(b=rng.rand(Nx, Ny), c=rng.rand(Nx, Ny), bB=rng.rand(B, Nx, Ny), cB=rng.rand(B, Nx, Ny),
               y1=rng.rand(B, Nx, Ny), y2=rng.rand(B, Nx, Ny), g=rng.randn(B, Nx, Ny))
    for k, v in raw.items():
        res[k + "_f64"] = v
    for dname in ("float32", "float64"):
        tdt, ndt = _dtype(dname)
        sfx = "_f32" if dname == "float32" else "_f64"
        for tag, bk, ck in (("shared", "b", "c"), ("batched", "bB", "cB")):
            b = torch.tensor(raw[bk], dtype=tdt, requires_grad=True)
            c = torch.tensor(raw[ck], dtype=tdt, requires_grad=True)
            y1 = torch.tensor(raw["y1"], dtype=tdt, requires_grad=True)
            y2 = torch.tensor(raw["y2"], dtype=tdt, requires_grad=True)
            y = TimeStep.apply(b, c, y1, y2, torch.tensor(0.10, dtype=tdt), torch.tensor(0.25, dtype=tdt))
            y.backward(torch.tensor(raw["g"], dtype=tdt))
            res[f"{tag}_y{sfx}"] = _np(y)
            res[f"{tag}_gb{sfx}"] = _np(b.grad)
            res[f"{tag}_gc{sfx}"] = _np(c.grad)
            res[f"{tag}_gy1{sfx}"] = _np(y1.grad)
            res[f"{tag}_gy2{sfx}"] = _np(y2.grad)
    res["dt_h"] = np.array([0.10, 0.25])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    print(name, "ok")


def geometry_case(name):
    """WaveGeometryHoley / FreeForm parameterisation outputs (geom.py:89-233)."""
    res = {}
    for dname in ("float32", "float64"):
        tdt, ndt = _dtype(dname)
        sfx = "_f32" if dname == "float32" else "_f64"
        gh = wt.WaveGeometryHoley((40, 36), 1.0, 1.0, 0.5, abs_N=5, abs_sig=4.0, abs_p=3.0, eta=0.5, beta=20.0,
                                  x=[12.0, 25.5], y=[10.0, 22.25], r=[3.0, 4.5])
        c = gh.c
        w = torch.tensor(np.cos(np.arange(40 * 36).reshape(40, 36) * 0.37), dtype=tdt)
        (c * w).sum().backward()
        res["holey_c" + sfx] = _np(c)
        res["holey_rho" + sfx] = _np(gh.rho)
        res["holey_b" + sfx] = _np(gh.b)
        res["holey_gx" + sfx] = _np(gh.x.grad)
        res["holey_gy" + sfx] = _np(gh.y.grad)
        res["holey_gr" + sfx] = _np(gh.r.grad)
        rng = np.random.RandomState(3)
        rho0 = rng.rand(31, 29)
        dr = torch.zeros(31, 29, dtype=torch.uint8)
        dr[8:24, 7:22] = 1
        gf = wt.WaveGeometryFreeForm((31, 29), 1.0, 1.0, 0.5, abs_N=4, abs_sig=5.0, abs_p=2.0, eta=0.45, beta=12.0,
                                     design_region=dr, rho=torch.tensor(rho0, dtype=tdt), blur_radius=2, blur_N=2)
        c = gf.c
        w = torch.tensor(np.sin(np.arange(31 * 29).reshape(31, 29) * 0.11), dtype=tdt)
        (c * w).sum().backward()
        res["free_rho_in"] = rho0
        res["free_design"] = _np(dr)
        res["free_rho" + sfx] = _np(gf.rho)
        res["free_c" + sfx] = _np(c)
        res["free_b" + sfx] = _np(gf.b)
        res["free_blur_kernel" + sfx] = _np(gf.blur_kernel)
        res["free_grho" + sfx] = _np(gf.rho.grad)
        res["free_w" + sfx] = _np(w)
        res["holey_w" + sfx] = _np(torch.tensor(np.cos(np.arange(40 * 36).reshape(40, 36) * 0.37), dtype=tdt))
        res["state_keys"] = np.array(sorted(gf.state_dict().keys()))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    print(name, "ok")


def state_dict_case(name):
    """Buffer / parameter names of a full model: the checkpoint compat surface (io.py:67-70)."""
    _dtype("float32")
    geom = wt.WaveGeometryFreeForm((60, 50), 1.0, 1.0, 0.5, abs_N=5)
    cell = wt.WaveCell(0.5, geom, satdamp_b0=0.1, satdamp_uth=1.0, c_nl=-3.0)
    model = wt.WaveRNN(cell, [wt.WaveSource(10, 25)], [wt.WaveIntensityProbe(50, 20), wt.WaveIntensityProbe(50, 30)])
    sd = model.state_dict()
    keys = sorted(sd.keys())
    shapes = [str(tuple(sd[k].shape)) + ":" + str(sd[k].dtype) for k in keys]
    params = sorted(n for n, _ in model.named_parameters())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), keys=np.array(keys), shapes=np.array(shapes),
                        params=np.array(params))
    print(name, keys)


def train_case(name):
    """wavetorch.train (train.py:13-133) for 2 epochs on a small deterministic problem, plus the checkpoint written by
    io.save_model (io.py:13-41): loss / accuracy history and the final rho.  pandas >= 2 removed DataFrame.append, which
    train.py:116 still calls; the shim below restores exactly that method for this run, nothing else is touched."""
    import tempfile
    import pandas as pd
    from torch.utils.data import TensorDataset, DataLoader
    if not hasattr(pd.DataFrame, "append"):
        def _append(self, row, ignore_index=True):
            return pd.concat([self, pd.DataFrame([row])], ignore_index=True)
        pd.DataFrame.append = _append
    _dtype("float32")
    torch.manual_seed(0)
    Nx, Ny, N, T, n_train, n_test, bs = 44, 36, 5, 80, 9, 6, 3
    design_region = torch.zeros(Nx, Ny, dtype=torch.uint8)
    design_region[14:30] = 1
    geom = wt.WaveGeometryFreeForm((Nx, Ny), 1.0, c0=1.0, c1=0.6, eta=0.5, beta=100, abs_sig=3.0, abs_N=N, abs_p=4.0,
                                   rho="half", blur_radius=1, blur_N=1, design_region=design_region)
    cell = wt.WaveCell(0.6, geom)
    probes = [wt.WaveIntensityProbe(36, y) for y in (10, 18, 26)]
    model = wt.WaveRNN(cell, [wt.WaveSource(8, 18)], probes)
    Xall = torch.tensor(wo.synthetic_vowels(n_train + n_test, T, dtype=np.float64), dtype=torch.float32)
    lab = torch.arange(n_train + n_test) % 3
    Yall = torch.nn.functional.one_hot(lab, 3).to(torch.float32)
    train_dl = DataLoader(TensorDataset(Xall[:n_train], Yall[:n_train]), batch_size=bs, shuffle=False)
    test_dl = DataLoader(TensorDataset(Xall[n_train:], Yall[n_train:]), batch_size=bs)
    optimizer = torch.optim.Adam(model.parameters(), lr=0.02)
    rho0 = _np(geom.rho).copy()
    cfg = {"dtype": "float32"}
    with tempfile.TemporaryDirectory() as d:
        history, states = wt.train(model, optimizer, torch.nn.CrossEntropyLoss(), train_dl, test_dl, 2, bs,
                                   name="ck", savedir=d + "/", cfg=cfg, accuracy=wt.utils.accuracy_onehot)
        data = torch.load(d + "/ck.pt", weights_only=False)
    res = {"x": _np(Xall), "labels": lab.numpy(), "rho0": rho0, "rho_final": _np(geom.rho),
           "design_region": _np(design_region),
           "loss_train": history["loss_train"].to_numpy(dtype=np.float64),
           "loss_test": history["loss_test"].to_numpy(dtype=np.float64),
           "acc_train": history["acc_train"].to_numpy(dtype=np.float64),
           "acc_test": history["acc_test"].to_numpy(dtype=np.float64),
           "cm_train_last": np.asarray(history["cm_train"].iloc[-1]), "cm_test_last": np.asarray(history["cm_test"].iloc[-1]),
           "epochs": history["epoch"].to_numpy(dtype=np.int64),
           "ckpt_keys": np.array(sorted(data.keys())), "ckpt_state_keys": np.array(sorted(data["model_state"].keys())),
           "ckpt_geom_class": np.array(data["model_geom_class_str"]),
           "ckpt_geom_args": np.array(sorted(data["history_geom_state"][-1].keys())),
           "n_states": np.asarray(len(states))}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    print(name, res["loss_train"], res["loss_test"], res["acc_train"], res["acc_test"], res["ckpt_keys"])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "train":      # regenerate only the training-loop fixture
        train_case("train_small")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "round2":     # the fixtures added in round 2 only
        vowel_sums_case("vowel_linear_B64_sums", 64, 1000)
        vowel_case("vowel_satdamp_T3000", 3, 3000, 0.1, 1.0, 0.0)   # config 4 at the yml's window_size (example_nonlinearity.yml:38)
        sys.exit(0)
    single_step_case("single_step")
    geometry_case("geometry")
    state_dict_case("state_dict")
    small_case("small_linear", 0.0, 0.0, 0.0, 11)
    small_case("small_satdamp", 0.4, 0.7, 0.0, 12)
    small_case("small_kerr", 0.0, 0.0, -0.12, 13, xamp=0.1)   # larger amplitudes blow up (SURVEY B-9)
    small_case("small_both", 0.4, 0.7, -0.12, 14, xamp=0.1)
    lens_case("lens_propagate", 1.0, with_grad=False)       # BASELINE config 1
    lens_case("lens_optimize", 0.5, with_grad=True)         # BASELINE config 2
    vowel_case("vowel_linear", 6, 1000, 0.0, 1.0, 0.0, x_grad=True)       # config 3 geometry, B=6
    vowel_case("vowel_satdamp", 6, 1000, 0.1, 1.0, 0.0)      # config 4 (i): example_nonlinearity.yml as is
    vowel_case("vowel_both", 6, 1000, 0.1, 1.0, -30.0)       # config 4 (ii)
    vowel_case("vowel_satdamp_uth", 6, 1000, 0.1, 0.00018, 0.0)  # config 4 (iii)
    vowel_case("vowel_kerr", 6, 1000, 0.0, 1.0, -30.0)
    vowel_sums_case("vowel_linear_B64_sums", 64, 1000)
    vowel_case("vowel_satdamp_T3000", 3, 3000, 0.1, 1.0, 0.0)
    train_case("train_small")
