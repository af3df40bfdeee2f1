"""Host-side logic of the drop-in API (no GPU): module surface, buffer names, geometry parameterisation,
error behaviour -- checked against fixtures produced by the reference (tests/golden/)."""
import numpy as np
import pytest
import torch

import wavetorch_b200 as wt
from conftest import load_golden, rel_l2


def test_public_names_match_reference_init():
    # wavetorch/__init__.py:9-10
    assert set(wt.__all__) == {"WaveCell", "WaveGeometryHoley", "WaveGeometryFreeForm", "WaveProbe",
                               "WaveIntensityProbe", "WaveRNN", "WaveSource", "WaveLineSource"}
    for sub in ("cell", "geom", "probe", "rnn", "source", "utils"):
        assert hasattr(wt, sub)


def test_state_dict_and_parameters_match_reference():
    """Checkpoint compat surface (io.py:67-70 parses these names)."""
    g = load_golden("state_dict")
    geom = wt.WaveGeometryFreeForm((60, 50), 1.0, 1.0, 0.5, abs_N=5)
    cell = wt.WaveCell(0.5, geom, satdamp_b0=0.1, satdamp_uth=1.0, c_nl=-3.0)
    model = wt.WaveRNN(cell, [wt.WaveSource(10, 25)], [wt.WaveIntensityProbe(50, 20), wt.WaveIntensityProbe(50, 30)])
    sd = model.state_dict()
    assert sorted(sd.keys()) == list(g["keys"])
    shapes = [str(tuple(sd[k].shape)) + ":" + str(sd[k].dtype) for k in sorted(sd.keys())]
    assert shapes == list(g["shapes"])
    assert sorted(n for n, _ in model.named_parameters()) == list(g["params"])
    assert [p.shape for p in model.cell.parameters()] == [torch.Size([60, 50])]     # cell.py:75-77
    model2 = wt.WaveRNN(wt.WaveCell(0.5, wt.WaveGeometryFreeForm((60, 50), 1.0, 1.0, 0.5, abs_N=5)),
                        [wt.WaveSource(0, 0)], [wt.WaveIntensityProbe(0, 0), wt.WaveIntensityProbe(0, 0)])
    model2.load_state_dict(sd)
    assert model2.sources[0].x.item() == 10 and model2.probes[1].y.item() == 30
    assert abs(model2.cell.host_scalars()["c_nl"] + 3.0) < 1e-6


@pytest.mark.parametrize("sfx,dtype", [("f32", "float32"), ("f64", "float64")])
def test_geometry_parameterisation(sfx, dtype):
    g = load_golden("geometry")
    wt.utils.set_dtype(dtype)
    try:
        tdt = torch.get_default_dtype()
        tol = 2e-6 if sfx == "f32" else 1e-12
        gh = wt.WaveGeometryHoley((40, 36), 1.0, 1.0, 0.5, abs_N=5, abs_sig=4.0, abs_p=3.0, eta=0.5, beta=20.0,
                                  x=[12.0, 25.5], y=[10.0, 22.25], r=[3.0, 4.5])
        c = gh.c
        (c * torch.tensor(g["holey_w_" + sfx], dtype=tdt)).sum().backward()
        assert rel_l2(c.detach().numpy(), g["holey_c_" + sfx]) < tol
        assert rel_l2(gh.rho.detach().numpy(), g["holey_rho_" + sfx]) < tol
        assert rel_l2(gh.b.numpy(), g["holey_b_" + sfx]) < tol
        # the hole centred exactly on a pixel has a NaN position gradient (d sqrt(0)) in the reference too
        np.testing.assert_allclose(gh.x.grad.numpy(), g["holey_gx_" + sfx], rtol=100 * tol, equal_nan=True)
        np.testing.assert_allclose(gh.y.grad.numpy(), g["holey_gy_" + sfx], rtol=100 * tol, equal_nan=True)
        assert rel_l2(gh.r.grad.numpy(), g["holey_gr_" + sfx]) < 20 * tol
        gf = wt.WaveGeometryFreeForm((31, 29), 1.0, 1.0, 0.5, abs_N=4, abs_sig=5.0, abs_p=2.0, eta=0.45, beta=12.0,
                                     design_region=torch.tensor(g["free_design"]),
                                     rho=torch.tensor(g["free_rho_in"], dtype=tdt), blur_radius=2, blur_N=2)
        c = gf.c
        (c * torch.tensor(g["free_w_" + sfx], dtype=tdt)).sum().backward()
        assert rel_l2(gf.rho.detach().numpy(), g["free_rho_" + sfx]) < tol
        assert rel_l2(gf.blur_kernel.numpy(), g["free_blur_kernel_" + sfx]) < tol
        assert rel_l2(c.detach().numpy(), g["free_c_" + sfx]) < tol
        assert rel_l2(gf.b.numpy(), g["free_b_" + sfx]) < tol
        assert rel_l2(gf.rho.grad.numpy(), g["free_grho_" + sfx]) < 20 * tol
        assert sorted(gf.state_dict().keys()) == list(g["state_keys"])
    finally:
        wt.utils.set_dtype("float32")


def test_config_geometries_match_reference_fields():
    """c and b of BASELINE configs 1-3 as the reference builds them."""
    g = load_golden("lens_optimize")
    rho = torch.zeros(151, 151)
    rr, cc = wt.geom.disk_pixels(75, 75, 30)
    assert len(rr) == 2809
    rho[rr, cc] = 0.5
    geom = wt.WaveGeometryFreeForm((151, 151), 1.0, c0=1.0, c1=0.5, rho=rho, design_region=None)
    assert rel_l2(geom.c.detach().numpy(), g["c_f32"]) < 1e-6
    assert rel_l2(geom.b.numpy(), g["b_f32"]) < 1e-6
    src = wt.WaveLineSource(25, 50, 25, 100)
    assert np.array_equal(src.x.numpy(), g["src_x"]) and np.array_equal(src.y.numpy(), g["src_y"])
    g3 = load_golden("vowel_linear")
    design = torch.tensor(g3["design_region"])
    geom3 = wt.WaveGeometryFreeForm((150, 100), 1.4283556979968262, c0=1.0, c1=0.5, eta=0.5, beta=100, abs_sig=3.0,
                                    abs_N=20, abs_p=4.0, rho="half", design_region=design)
    assert rel_l2(geom3.c.detach().numpy(), g3["c_f32"]) < 1e-6
    assert rel_l2(geom3.rho.detach().numpy(), g3["rho_f32"]) == 0
    assert int((geom3.rho > 0).sum()) == 3600


def test_line_pixels_general():
    r, c = wt.source.line_pixels(2, 3, 9, 6)       # steep
    assert len(r) == 8 and (r[0], c[0], r[-1], c[-1]) == (2, 3, 9, 6) and np.all(np.diff(r) == 1)
    r, c = wt.source.line_pixels(5, 9, 3, 1)       # shallow, reversed
    assert len(c) == 9 and (r[0], c[0], r[-1], c[-1]) == (5, 9, 3, 1) and np.all(np.diff(c) == -1)


def test_error_behaviour():
    geom = wt.WaveGeometryFreeForm((60, 50), 1.0, 1.0, 0.5, abs_N=5)
    with pytest.raises(ValueError, match="CFL"):                       # cell.py:70-73
        wt.WaveCell(0.9, geom)
    with pytest.raises(AssertionError):                                # geom.py:68-71
        wt.WaveGeometryFreeForm((30, 50), 1.0, 1.0, 0.5, abs_N=20)
    with pytest.raises(AssertionError):                                # geom.py:16
        wt.WaveGeometryFreeForm((30, 50, 2), 1.0, 1.0, 0.5)
    with pytest.raises(ValueError):                                    # geom.py:197
        wt.WaveGeometryFreeForm((60, 50), 1.0, 1.0, 0.5, abs_N=5, rho="checkerboard")
    with pytest.raises(NotImplementedError):                           # geom.py:42-45
        geom.forward()
    with pytest.raises(ValueError):
        wt.utils.set_dtype("float16")
    model = wt.WaveRNN(wt.WaveCell(0.5, geom), wt.WaveSource(10, 10), wt.WaveProbe(40, 40))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(2, 5))
    bad = wt.WaveRNN(wt.WaveCell(0.5, geom), wt.WaveSource(10, 10), wt.WaveProbe(80, 40))
    with pytest.raises((IndexError, RuntimeError)):
        bad(torch.zeros(2, 5))


def test_source_and_probe_modules_standalone_on_cpu():
    """WaveSource/WaveProbe keep working as plain modules (source.py:15-22, probe.py:14-27)."""
    Y = torch.arange(2 * 5 * 6, dtype=torch.float32).reshape(2, 5, 6)
    X = torch.tensor([10.0, -1.0])
    s = wt.WaveSource(2, 3)
    out = s(Y, X)
    ref = Y.clone(); ref[:, 2, 3] += X
    assert torch.equal(out, ref)
    line = wt.WaveLineSource(1, 1, 1, 4)
    out = line(Y[:1], X[:1])
    ref = Y[:1].clone(); ref[:, 1, 1:5] += 10.0
    assert torch.equal(out, ref)
    assert torch.equal(wt.WaveProbe(4, 5)(Y), Y[:, 4, 5])
    assert torch.equal(wt.WaveIntensityProbe(4, 5)(Y), Y[:, 4, 5] ** 2)
    assert abs(wt.utils.normalize_power(torch.rand(4, 3)).sum(1) - 1).max() < 1e-6
    assert wt.utils.accuracy_onehot(torch.tensor([[0.1, 0.9], [0.8, 0.2]]), torch.tensor([1, 1])) == 0.5
    assert len(wt.utils.window_data(np.arange(100), 10)) == 10


def test_confusion_matrix_matches_sklearn():
    """wavetorch_b200.train.confusion_matrix replaces sklearn.metrics.confusion_matrix as called at train.py:93,106."""
    from sklearn.metrics import confusion_matrix as sk_cm
    from wavetorch_b200.train import confusion_matrix
    rng = np.random.RandomState(0)
    for labels in ([0, 1, 2], [0, 2], [1], [0, 1, 2, 5]):
        yt = rng.choice(labels, size=40)
        yp = rng.choice(labels, size=40)
        np.testing.assert_array_equal(confusion_matrix(torch.tensor(yt), torch.tensor(yp)), sk_cm(yt, yp))


def test_checkpoint_schema_round_trip(tmp_path):
    """io.save_model / load_model keep the reference's .pt schema (io.py:28-36) and rebuild the same geometry; a
    checkpoint written in the reference's layout (fixture key list) loads."""
    g = load_golden("train_small")
    geom = wt.WaveGeometryFreeForm((44, 36), 1.0, c0=1.0, c1=0.6, abs_N=5, rho="half", design_region=torch.tensor(g["design_region"]))
    model = wt.WaveRNN(wt.WaveCell(0.6, geom), [wt.WaveSource(8, 18)], [wt.WaveIntensityProbe(36, y) for y in (10, 18, 26)])
    wt.io.save_model(model, "m", str(tmp_path) + "/", cfg={"dtype": "float32"}, verbose=False)
    data = torch.load(str(tmp_path) + "/m.pt", weights_only=False)
    assert sorted(data.keys()) == list(g["ckpt_keys"])
    assert sorted(data["model_state"].keys()) == list(g["ckpt_state_keys"])
    m2, hist, states, cfg = wt.io.load_model(str(tmp_path) + "/m.pt", verbose=False)
    assert hist is None and cfg == {"dtype": "float32"} and len(states) == 1
    assert torch.equal(m2.cell.geom.rho.detach(), geom.rho.detach())
    assert torch.equal(m2.cell.geom.b, geom.b)
    assert [(int(p.x), int(p.y)) for p in m2.probes] == [(36, 10), (36, 18), (36, 26)]
    assert (int(m2.sources[0].x), int(m2.sources[0].y)) == (8, 18)
