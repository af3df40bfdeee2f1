"""CPU oracle for the wave-RNN hot path (numpy restatement of the reference algorithm).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg as the *checker*.  The product path (wavetorch_b200/) never imports this file
and has no CPU fallback.

Parity status: PINNED.  The reference ships no golden vectors; this oracle
is pinned against outputs of the unmodified reference itself, generated in the build container by
oracle/gen_golden.py (imports /root/reference through oracle/ref_loader.py) and committed under
tests/golden/.  tests/test_oracle_golden.py holds the comparison.

Each function cites the reference lines it restates (paths relative to /root/reference).
All arithmetic is done in the dtype of the inputs (float32 or float64), following the reference's
association order in the forward step so that the float32 oracle tracks the float32 reference
to rounding noise.  The adjoint is an explicit reverse-time recursion
and does not use any autograd.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------
# geometry / coefficients (evaluated once per forward; feeds the hot loop)
# --------------------------------------------------------------------------------------
def pml_damping(Nx, Ny, N=20, sig=11.0, p=4.0, dtype=np.float32):
    """b(x,y) of the absorbing layer -- wavetorch/geom.py:63-85."""
    dt = np.dtype(dtype).type
    bx = np.zeros((Nx, Ny), dtype=dtype)
    by = np.zeros((Nx, Ny), dtype=dtype)
    if N > 0:
        assert Nx > 2 * N + 1 and Ny > 2 * N + 1          # geom.py:68-71
        # torch.linspace(0,1,N+1)**p in the default dtype, times sig (geom.py:73)
        ramp = (dt(sig) * np.linspace(0.0, 1.0, N + 1, dtype=np.float64).astype(dtype) ** dt(p)).astype(dtype)
        bx[0:N + 1, :] = ramp[::-1, None]                 # geom.py:79
        bx[Nx - N - 1:Nx, :] = ramp[:, None]              # geom.py:80
        by[:, 0:N + 1] = ramp[None, ::-1]                 # geom.py:82
        by[:, Ny - N - 1:Ny] = ramp[None, :]              # geom.py:83
    return np.sqrt(bx * bx + by * by).astype(dtype)       # geom.py:85


def disk_kernel(radius, dtype=np.float32):
    """Normalised blur stencil -- geom.py:149-152 (skimage.draw.circle(r, r, r+1))."""
    n = 2 * radius + 1
    ii, jj = np.mgrid[0:n, 0:n]
    k = ((((ii - radius) / (radius + 1.0)) ** 2 + ((jj - radius) / (radius + 1.0)) ** 2) < 1.0).astype(dtype)
    return (k / k.sum()).astype(dtype)


def _corr2d_zero(a, k):
    """Zero-padded 'same' cross-correlation (what F.conv2d(padding=r) computes) -- geom.py:213."""
    r = k.shape[0] // 2
    Nx, Ny = a.shape
    pad = np.zeros((Nx + 2 * r, Ny + 2 * r), dtype=a.dtype)
    pad[r:r + Nx, r:r + Ny] = a
    out = np.zeros_like(a)
    for di in range(k.shape[0]):
        for dj in range(k.shape[1]):
            if k[di, dj] != 0:
                out += k[di, dj] * pad[di:di + Nx, dj:dj + Ny]
    return out


def blur(rho, kernel, n_pass):
    """geom.py:207-215."""
    for _ in range(int(n_pass)):
        rho = _corr2d_zero(rho, kernel)
    return rho


def project(rho, eta, beta):
    """tanh projection -- geom.py:217-222 (scalars evaluated in float64 by np.tanh there too)."""
    t = rho.dtype.type
    num0 = np.tanh(beta * eta)
    den = np.tanh(beta * eta) + np.tanh(beta * (1.0 - eta))
    return ((t(num0) + np.tanh(t(beta) * (rho - t(eta)))) / t(den)).astype(rho.dtype)


def wave_speed(rho, c0, c1, eta=0.5, beta=100.0, blur_radius=1, blur_N=1):
    """c = c0 + (c1-c0) * proj(blur(rho)) -- geom.py:224-233."""
    t = rho.dtype.type
    k = disk_kernel(blur_radius, rho.dtype)
    return (t(c0) + t(c1 - c0) * project(blur(rho, k, blur_N), eta, beta)).astype(rho.dtype)


def wave_speed_vjp(rho, grad_c, c0, c1, eta=0.5, beta=100.0, blur_radius=1, blur_N=1):
    """d(sum(grad_c * c))/d rho: reverse of geom.py:207-233 (the reference leaves this to autograd)."""
    t = rho.dtype.type
    k = disk_kernel(blur_radius, rho.dtype)
    rb = blur(rho, k, blur_N)
    den = np.tanh(beta * eta) + np.tanh(beta * (1.0 - eta))
    th = np.tanh(t(beta) * (rb - t(eta)))
    dproj = t(beta) * (t(1.0) - th * th) / t(den)
    g = grad_c * t(c1 - c0) * dproj
    kt = k[::-1, ::-1]                      # adjoint of a zero-padded correlation
    for _ in range(int(blur_N)):
        g = _corr2d_zero(g, kt)
    return g


def constrain_to_design_region(rho, design_region, b):
    """geom.py:201-205."""
    rho = rho.copy()
    if design_region is not None:
        rho[design_region == 0] = 0
    rho[b > 0] = 0
    return rho


# --------------------------------------------------------------------------------------
# the step
# --------------------------------------------------------------------------------------
def laplacian(u, h):
    """5-point Laplacian, zero outside the domain -- wavetorch/operators.py:5-11.

    u: [B, Nx, Ny].  Stencil weights are h**-2 * [[0,1,0],[1,-4,1],[0,1,0]] (operators.py:7).
    """
    t = u.dtype.type
    w = t(1.0) / (t(h) * t(h))            # ATen evaluates x**-2 as 1/(x*x); keep 2/dt^2 == 2*dt^-2 exactly
    out = (t(-4.0) * w) * u
    out[:, 1:, :] += w * u[:, :-1, :]
    out[:, :-1, :] += w * u[:, 1:, :]
    out[:, :, 1:] += w * u[:, :, :-1]
    out[:, :, :-1] += w * u[:, :, 1:]
    return out


def saturable_damping(u, uth, b0):
    """b0 / (1 + |u/uth|^2) -- wavetorch/cell.py:8-9."""
    t = u.dtype.type
    a = np.abs(u / t(uth))
    return t(b0) / (t(1.0) + a * a)


def time_step(b, c, y1, y2, dt, h):
    """One leapfrog update -- wavetorch/cell.py:12-17 (same association order)."""
    t = y1.dtype.type
    dt = t(dt)
    # ATen's pow special-cases the exponents used here: x**2 = x*x, x**-2 = 1/(x*x), x**-1 = 1/x.
    # That makes fl(2/dt^2) == 2*fl(dt^-2) bit-exactly, i.e. the y1 and y2 weights sum to one in the
    # undamped interior.  A libm powf() that is 1 ulp off breaks this and the float32 trajectory drifts
    # by ~1e-5 in 50 steps (measured while pinning this oracle), so the same evaluation is used here.
    idt2 = t(1.0) / (dt * dt)
    idt1 = t(1.0) / dt
    return (t(1.0) / (idt2 + b * idt1)) * (
        t(2.0) / (dt * dt) * y1 - (idt2 - b * idt1) * y2 + (c * c) * laplacian(y1, h))


def time_step_vjp(b, c, y1, y2, dt, h, g):
    """Single-step adjoint -- wavetorch/cell.py:27-44.  Returns per-sample (grad_b, grad_c, grad_y1, grad_y2)."""
    t = y1.dtype.type
    dt = t(dt)
    lap = laplacian(y1, h)
    q = t(1.0) / (b * dt + t(1.0))
    dt2 = dt * dt
    grad_b = -(q * q) * dt * ((c * c) * dt2 * lap + t(2.0) * y1 - t(2.0) * y2) * g          # cell.py:33-34
    grad_c = q * (t(2.0) * c * dt2 * lap) * g                                               # cell.py:36
    grad_y1 = dt2 * laplacian(q * (c * c) * g, h) + t(2.0) * g * q                          # cell.py:39-40
    grad_y2 = (b * dt - t(1.0)) * q * g                                                     # cell.py:42
    return grad_b, grad_c, grad_y1, grad_y2


def _cell_coefficients(u1, c_lin, b_pml, rho, dt, b0, uth, c_nl):
    """WaveCell.forward's choice of b and c -- wavetorch/cell.py:94-102."""
    t = u1.dtype.type
    if b0 > 0:
        b = b_pml + rho * saturable_damping(u1, uth, b0)
    else:
        b = b_pml
    if c_nl != 0:
        c = c_lin + rho * t(c_nl) * (u1 * u1)
    else:
        c = c_lin
    return b, c


# --------------------------------------------------------------------------------------
# the time loop  (wavetorch/rnn.py:21-72)
# --------------------------------------------------------------------------------------
def forward(c_lin, b_pml, rho, x, src, prb, dt, h, b0=0.0, uth=0.0, c_nl=0.0,
            keep_fields=False, u_init=None):
    """WaveRNN.forward on raw arrays.

    c_lin, b_pml, rho : [Nx, Ny]      (rho may be None in linear mode)
    x                 : [B, T]        waveform; every source pixel receives the same x[b, t]
                                      (rnn.py:56-57, source.py:15-22 with its dt=1.0 default)
    src, prb          : int arrays [n, 2] of (row, col); duplicates in src add twice (SURVEY B-3)
    Returns dict(raw=[B,T,P] field at the probes after injection, u=[T+2,B,Nx,Ny] if keep_fields
    (u[k] = field after step k-2; u[0], u[1] = initial h2, h1), u1, u2 = final state).
    """
    dtype = x.dtype
    B, T = x.shape
    Nx, Ny = c_lin.shape
    src = np.asarray(src, dtype=np.int64).reshape(-1, 2)
    prb = np.asarray(prb, dtype=np.int64).reshape(-1, 2)
    if u_init is None:
        u1 = np.zeros((B, Nx, Ny), dtype=dtype)       # rnn.py:39-41
        u2 = np.zeros((B, Nx, Ny), dtype=dtype)
    else:
        u1, u2 = (np.array(a, dtype=dtype) for a in u_init)
    raw = np.zeros((B, T, prb.shape[0]), dtype=dtype)
    hist = None
    if keep_fields:
        hist = np.zeros((T + 2, B, Nx, Ny), dtype=dtype)
        hist[0], hist[1] = u2, u1
    for t in range(T):                                 # rnn.py:50
        b, c = _cell_coefficients(u1, c_lin, b_pml, rho, dt, b0, uth, c_nl)
        y = time_step(b, c, u1, u2, dt, h)             # rnn.py:53 -> cell.py:104
        for (i, j) in src:                             # rnn.py:56-57; source.py:19-22
            y[:, i, j] = y[:, i, j] + x[:, t]
        u2, u1 = u1, y                                 # cell.py:107
        if prb.shape[0]:
            raw[:, t, :] = u1[:, prb[:, 0], prb[:, 1]]  # probe.py:15
        if keep_fields:
            hist[t + 2] = u1
    return {"raw": raw, "u": hist, "u1": u1, "u2": u2}


def probe_outputs(raw, intensity):
    """probe.py:15 (plain) / probe.py:27 (intensity = square).  intensity: bool [P]."""
    intensity = np.asarray(intensity, dtype=bool)
    return np.where(intensity[None, None, :], raw * raw, raw)


def adjoint(c_lin, b_pml, rho, x, src, prb, intensity, dt, h, grad_out, fwd,
            b0=0.0, uth=0.0, c_nl=0.0):
    """Reverse-time adjoint of `forward` (what autograd does through cell.py:27-44, cell.py:94-102,
    source.py:22, probe.py:15/27, rnn.py:50-70).

    grad_out : [B, T, P]  dLoss/d(probe output) (w.r.t. the squared value for intensity probes)
    fwd      : result of forward(..., keep_fields=True)
    Returns dict(grad_c, grad_b, grad_rho [Nx,Ny] summed over batch and time, grad_x [B,T]).
    """
    dtype = x.dtype
    t_ = dtype.type
    B, T = x.shape
    Nx, Ny = c_lin.shape
    src = np.asarray(src, dtype=np.int64).reshape(-1, 2)
    prb = np.asarray(prb, dtype=np.int64).reshape(-1, 2)
    intensity = np.asarray(intensity, dtype=bool)
    U, raw = fwd["u"], fwd["raw"]
    gc = np.zeros((Nx, Ny), dtype=dtype)
    gb = np.zeros((Nx, Ny), dtype=dtype)
    gr = np.zeros((Nx, Ny), dtype=dtype)
    gx = np.zeros((B, T), dtype=dtype)
    carry1 = np.zeros((B, Nx, Ny), dtype=dtype)    # dLoss/du_t from later steps
    carry2 = np.zeros((B, Nx, Ny), dtype=dtype)    # dLoss/du_{t-1} from step t+1's y2 path
    nonlinear = (b0 > 0) or (c_nl != 0)
    rho_ = rho if rho is not None else np.zeros((Nx, Ny), dtype=dtype)
    for t in range(T - 1, -1, -1):
        lam = carry1
        for p in range(prb.shape[0]):               # probe readout adjoint (probe.py:15,27)
            i, j = prb[p]
            seed = grad_out[:, t, p]
            if intensity[p]:
                seed = t_(2.0) * raw[:, t, p] * seed
            lam[:, i, j] = lam[:, i, j] + seed
        for (i, j) in src:                           # source.py:22 passes grad through; x gets the gather
            gx[:, t] += lam[:, i, j]
        u1, u2 = U[t + 1], U[t]                      # inputs of the step that produced u_t = U[t+2]
        b, c = _cell_coefficients(u1, c_lin, b_pml, rho_, dt, b0, uth, c_nl)
        g_b, g_c, g_u1, g_u2 = time_step_vjp(b, c, u1, u2, dt, h, lam)
        gc += g_c.sum(axis=0)                        # grad w.r.t. c_linear (cell.py:102: c = c_linear [+ ...])
        gb += g_b.sum(axis=0)                        # grad w.r.t. geom.b (cell.py:95/97)
        if nonlinear:                                # autograd through cell.py:94-100
            if b0 > 0:
                d = t_(1.0) + (u1 / t_(uth)) * (u1 / t_(uth))
                gr += (g_b * (t_(b0) / d)).sum(axis=0)
                g_u1 = g_u1 + g_b * rho_ * t_(b0) * (t_(-2.0) * u1 / (t_(uth) * t_(uth))) / (d * d)
            if c_nl != 0:
                gr += (g_c * t_(c_nl) * (u1 * u1)).sum(axis=0)
                g_u1 = g_u1 + g_c * (t_(2.0) * rho_ * t_(c_nl) * u1)
        carry1 = carry2 + g_u1
        carry2 = g_u2
    return {"grad_c": gc, "grad_b": gb, "grad_rho": gr, "grad_x": gx,
            "grad_u1_init": carry1, "grad_u2_init": carry2}


# --------------------------------------------------------------------------------------
# loss head used by the reference's training scripts
# --------------------------------------------------------------------------------------
def loss_head(out, labels):
    """CrossEntropy(normalize_power(sum_t out), labels) -- train.py:61-62, utils.py:35-36.

    Returns (loss, dLoss/d out [B,T,P]).  Mean reduction over the batch (torch default).
    """
    dtype = out.dtype
    B, T, P = out.shape
    S = out.sum(axis=1)                                   # [B,P]
    tot = S.sum(axis=1, keepdims=True)
    pwr = S / tot                                         # utils.py:36
    z = pwr - pwr.max(axis=1, keepdims=True)
    lse = np.log(np.exp(z).sum(axis=1, keepdims=True))
    logp = z - lse
    labels = np.asarray(labels, dtype=np.int64)
    loss = -logp[np.arange(B), labels].mean()
    soft = np.exp(logp)
    dp = soft.copy()
    dp[np.arange(B), labels] -= 1.0
    dp = (dp / B).astype(dtype)
    dS = dp / tot - (dp * S).sum(axis=1, keepdims=True) / (tot * tot)
    gout = np.broadcast_to(dS[:, None, :], (B, T, P)).astype(dtype)
    return dtype.type(loss), gout


# --------------------------------------------------------------------------------------
# deterministic synthetic inputs
# --------------------------------------------------------------------------------------
_FORMANTS = np.array([[730.0, 1090.0, 2440.0], [270.0, 2290.0, 3010.0], [300.0, 870.0, 2240.0]])
_AMPS = np.array([1.0, 0.5, 0.25])


def synthetic_vowels(B, T, sr=10000.0, dtype=np.float32, first=0):
    """RNG-free vowel-like waveforms, unit energy like data/vowels.py:12-17.  Sample index = first+b."""
    n = np.arange(T, dtype=np.float64)
    env = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / (T - 1))
    x = np.zeros((B, T), dtype=np.float64)
    for bb in range(B):
        b = first + bb
        k = b % 3
        jit = 1.0 + 0.02 * (np.modf(0.7548776662466927 * (b + 1))[0] - 0.5)
        for j in range(3):
            phi = 2.0 * np.pi * np.modf(0.6180339887498949 * (3 * b + j + 1))[0]
            x[bb] += _AMPS[j] * np.sin(2.0 * np.pi * _FORMANTS[k, j] * jit * n / sr + phi)
        x[bb] *= env
        x[bb] /= np.sqrt((x[bb] ** 2).sum())
    return x.astype(dtype)


def propagate_waveform(T=500, dt=0.707, dtype=np.float32):
    """study/propagate.py:48-51."""
    t = np.arange(0, T * dt, dt)[:T]
    omega1 = 2 * np.pi * 1 / dt / 15
    return (np.sin(omega1 * t) * t / (1 + t)).astype(dtype)[None, :]


def lens_config(rho_in_disk, dtype=np.float32):
    """Geometry of study/propagate.py:16-33 (rho=1 in the disk) / study/optimize_lens.py:14-33 (rho=0.5)."""
    Nx = Ny = 151
    ii, jj = np.mgrid[0:Nx, 0:Ny]
    rho = np.zeros((Nx, Ny), dtype=dtype)
    rho[((ii - 75) / 30.0) ** 2 + ((jj - 75) / 30.0) ** 2 < 1.0] = rho_in_disk
    b = pml_damping(Nx, Ny, 20, 11.0, 4.0, dtype)
    rho = constrain_to_design_region(rho, None, b)          # geom.py:158 (constructor)
    src = np.stack([np.full(51, 25), np.arange(50, 101)], axis=1)   # WaveLineSource(25,50,25,100)
    prb = np.array([[125, 100], [125, 75], [125, 50]])
    return dict(Nx=Nx, Ny=Ny, dt=0.707, h=1.0, c0=1.0, c1=0.5, rho=rho, b=b, src=src, prb=prb,
                eta=0.5, beta=100.0, blur_radius=1, blur_N=1, intensity=np.array([True] * 3))


def vowel_config(dtype=np.float32, Nx=150, Ny=100):
    """Geometry of study/example.yml through study/vowel_train.py:89-114 and vowel_helpers.py:4-35."""
    N = 20
    b = pml_damping(Nx, Ny, N, 3.0, 4.0, dtype)
    src = np.array([[N + 20, Ny // 2]])
    span = 2 * 20
    y0 = int((Ny - span) / 2)
    prb = np.array([[Nx - N - 20, y0 + 20 * i] for i in range(3)])
    design = np.zeros((Nx, Ny), dtype=np.uint8)
    design[src[0, 0] + 5:prb[0, 0] - 5] = 1
    rho = constrain_to_design_region(np.full((Nx, Ny), 0.5, dtype=dtype), design, b)
    return dict(Nx=Nx, Ny=Ny, dt=1.0, h=1.4283556979968262, c0=1.0, c1=0.5, rho=rho, b=b, src=src, prb=prb,
                eta=0.5, beta=100.0, blur_radius=1, blur_N=1, intensity=np.array([True] * 3),
                design_region=design)


def run_training_step(cfg, x, labels, b0=0.0, uth=0.0, c_nl=0.0):
    """fwd + loss + adjoint + geometry chain: what one closure() of train.py:59-64 computes.

    Returns dict(out, S, loss, grad_rho (= rho.grad of the reference), grad_c, grad_x).
    """
    rho = cfg["rho"]
    c = wave_speed(rho, cfg["c0"], cfg["c1"], cfg["eta"], cfg["beta"], cfg["blur_radius"], cfg["blur_N"])
    f = forward(c, cfg["b"], rho, x, cfg["src"], cfg["prb"], cfg["dt"], cfg["h"], b0, uth, c_nl, keep_fields=True)
    out = probe_outputs(f["raw"], cfg["intensity"])
    loss, gout = loss_head(out, labels)
    a = adjoint(c, cfg["b"], rho, x, cfg["src"], cfg["prb"], cfg["intensity"], cfg["dt"], cfg["h"], gout, f,
                b0, uth, c_nl)
    grad_rho = wave_speed_vjp(rho, a["grad_c"], cfg["c0"], cfg["c1"], cfg["eta"], cfg["beta"],
                              cfg["blur_radius"], cfg["blur_N"]) + a["grad_rho"]
    return dict(out=out, raw=f["raw"], S=out.sum(axis=1), loss=loss, grad_rho=grad_rho, grad_c=a["grad_c"],
                grad_b=a["grad_b"], grad_rho_direct=a["grad_rho"], grad_x=a["grad_x"], c=c)
