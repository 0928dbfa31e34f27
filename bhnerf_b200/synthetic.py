"""Synthetic hot-path inputs of the shapes named in BASELINE.json (no network, no kgeo on the GPU box).

Flat-space stand-in for the Kerr geodesics kgeo produces (bhnerf/kgeo.py:6-63): straight rays from an
observer at r_o = 1000 M with inclination i through an image plane alpha,beta in [-fov/2, fov/2], sampled
at G EQUAL MINO-TIME steps like kgeo (kerr_raytracing_ana.py:105-106): for a straight line of impact
parameter b, d tau = d lambda / r^2 gives lambda = b tan(phi) with phi uniform, which crowds the samples
near the black hole exactly as real geodesics do (in-domain fraction ~0.4 for the Tutorial3 geometry).  Value ranges follow what real
geodesics give: t_geos ~ -(r_o - lambda), dtau*Sigma = d lambda with dtau[...,0] = 0,
Keplerian Omega, Doppler factor g = 1/(u^t (1 - lam*Omega)) with g = 0 where no circular orbit exists.
The arrays have the reference's layout: coords (3, A, B, G), everything else (A, B, G), float32."""
import numpy as np

CONFIGS = {
    # name: spin, inclination [deg], fov_M, A=B, G, frames, rmin, rmax, z_width, S, loss, sigma
    'cfg1_tutorial3': dict(spin=0.2, inc=60.0, fov=16.0, n=64, G=64, nt=64, rmin=2.0, rmax=8.0, z_width=4.0, S=1,
                           loss='full', sigma=1.0, t_span=(0.0, 1.0)),
    'cfg2_lp_flare': dict(spin=0.0, inc=12.0, fov=40.0, n=128, G=128, nt=100, rmin=6.0, rmax=20.0, z_width=4.0, S=2,
                          loss='lc', sigma=0.01, t_span=(9.3, 11.3)),
    'cfg3_ngeht': dict(spin=0.2, inc=60.0, fov=16.0, n=128, G=128, nt=64, rmin=2.0, rmax=8.0, z_width=4.0, S=1,
                       loss='vis', sigma=0.01, t_span=(2.0, 2.667), V=190),
    'cfg4_highres': dict(spin=0.2, inc=60.0, fov=16.0, n=256, G=256, nt=128, rmin=2.0, rmax=8.0, z_width=4.0, S=1,
                         loss='full', sigma=1.0, t_span=(0.0, 1.0)),
    'cfg5_alma': dict(spin=0.0, inc=12.0, fov=40.0, n=64, G=100, nt=100, rmin=6.0, rmax=20.0, z_width=4.0, S=3,
                      loss='lc', sigma=(0.15, 0.01, 0.01), t_span=(9.3, 11.3)),
    'tiny': dict(spin=0.2, inc=60.0, fov=16.0, n=16, G=32, nt=4, rmin=2.0, rmax=8.0, z_width=4.0, S=1,
                 loss='full', sigma=1.0, t_span=(0.0, 1.0)),
}


def smooth_J(S, shape, seed=0):
    rng = np.random.default_rng(seed)
    A, B, G = shape
    a = np.linspace(0, 1, A, dtype=np.float32)[:, None, None]
    b = np.linspace(0, 1, B, dtype=np.float32)[None, :, None]
    k = np.linspace(0, 1, G, dtype=np.float32)[None, None, :]
    J = np.empty((S, A, B, G), dtype=np.float32)
    for s in range(S):
        ph = rng.uniform(0, 2 * np.pi, 3); fr = rng.uniform(1, 4, 3)
        J[s] = (np.cos(fr[0] * a * 2 * np.pi + ph[0]) * np.cos(fr[1] * b * 2 * np.pi + ph[1])
                * np.cos(fr[2] * k * 2 * np.pi + ph[2]))
    if S == 3:
        J[0] = 0.5 + 0.5 * np.abs(J[0])
    return J


def flat_geodesics(spin, inc_deg, fov, n, G, r_o=1000.0, n_beta=None):
    """Returns dict(coords, Omega, g, dtau, Sigma, t_geos, r) in float32."""
    A = n; B = n_beta or n
    inc = np.deg2rad(inc_deg)
    nhat = np.array([np.sin(inc), 0.0, np.cos(inc)])
    e_a = np.array([0.0, 1.0, 0.0])
    e_b = np.array([-np.cos(inc), 0.0, np.sin(inc)])
    alpha = np.linspace(-fov / 2, fov / 2, A)[:, None, None]
    beta = np.linspace(-fov / 2, fov / 2, B)[None, :, None]
    bimp = np.maximum(np.sqrt(alpha ** 2 + beta ** 2), 1e-3)        # impact parameter (A,B,1)
    phi0 = np.arctan(r_o / bimp)
    kk = np.arange(G)[None, None, :]
    phi = phi0 - kk * (2 * phi0 / (G - 1))                          # equal Mino-time steps, observer -> far side
    lam = bimp * np.tan(phi)
    pts = [alpha * e_a[c] + beta * e_b[c] + lam * nhat[c] for c in range(3)]
    x, y, z = [np.broadcast_to(p, (A, B, G)).copy() for p in pts]
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    costh = z / np.maximum(r, 1e-9)
    Sigma = r ** 2 + spin ** 2 * costh ** 2
    dtau = np.broadcast_to(2 * phi0 / (G - 1) / bimp, (A, B, G)).copy()   # per-ray constant Mino step
    dtau[..., 0] = 0.0                                              # kerr_raytracing_utils.py:275
    t_geos = -(r_o - lam)
    sgn = np.sign(spin + np.finfo(float).eps)
    Omega = sgn / (r ** 1.5 + spin)
    lam_c = -np.broadcast_to(alpha, (A, B, G)) * np.sin(inc)        # conserved angular momentum of the ray
    with np.errstate(invalid='ignore', divide='ignore'):
        ut = 1.0 / np.sqrt(1.0 - 3.0 / r)
        g = 1.0 / (ut * (1.0 - lam_c * Omega))
    g = np.where(np.isfinite(g) & (g > 0), g, 0.0)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(coords=f32(np.stack([x, y, z])), Omega=f32(Omega), g=f32(g), dtau=f32(dtau), Sigma=f32(Sigma),
                t_geos=f32(t_geos), r=f32(r))


def make_config(name, seed=0, frame_offset=0.0, nt=None, inc=None):
    """Inputs for one BASELINE.json config: raytracing args (reference positional order), predictor
    constants, frame times and seeded targets.  `inc` overrides the inclination [deg] (ensemble sweeps)."""
    c = dict(CONFIGS[name])
    if nt is not None:
        c['nt'] = nt
    if inc is not None:
        c['inc'] = float(inc)
    geo = flat_geodesics(c['spin'], c['inc'], c['fov'], c['n'], c['G'])
    A = B = c['n']
    rng = np.random.default_rng(seed)
    S = c['S']
    J = 1.0 if S == 1 else smooth_J(S, (A, B, c['G']), seed + 5)
    t_frames = (np.linspace(c['t_span'][0], c['t_span'][1], c['nt']) + frame_offset).astype(np.float32)
    r_o = 1000.0
    t_injection = -(r_o + c['fov'] / 4.0) if c['inc'] < 30 else -r_o
    rt = dict(coords=geo['coords'], Omega=geo['Omega'], J=J, g=geo['g'], dtau=geo['dtau'], Sigma=geo['Sigma'],
              t_start_obs=float(c['t_span'][0]), t_geos=geo['t_geos'], t_injection=float(t_injection))
    pred = dict(scale=c['rmax'], rmin=c['rmin'], rmax=c['rmax'], z_width=c['z_width'])
    out = dict(cfg=c, rt=rt, predictor=pred, t_frames=t_frames, A=A, B=B, G=c['G'], S=S, P=A * B)
    if c['loss'] == 'full':
        out['target'] = rng.uniform(0.0, 0.5, size=(c['nt'], S, A * B)).astype(np.float32)
        out['sigma'] = np.full_like(out['target'], c['sigma'])
    elif c['loss'] == 'lc':
        out['target'] = rng.normal(0.1, 0.2, size=(c['nt'], S)).astype(np.float32)
        out['sigma'] = np.broadcast_to(np.asarray(c['sigma'], dtype=np.float32), (c['nt'], S)).copy()
    else:
        V = c['V']
        psize = c['fov'] / A
        xx, yy = np.meshgrid((np.arange(A) - A / 2) * psize, (np.arange(B) - B / 2) * psize, indexing='ij')
        uv = rng.uniform(-0.25, 0.25, size=(c['nt'], V, 2)).astype(np.float32)
        ph = -2 * np.pi * (uv[..., 0:1] * xx.reshape(1, 1, -1) + uv[..., 1:2] * yy.reshape(1, 1, -1))
        out['Amat'] = (np.cos(ph) + 1j * np.sin(ph)).astype(np.complex64)
        out['uv'] = uv                      # the same matrix as baselines: network.SeparableDFT.from_fov(uv, (A, B), fov)
        out['target'] = (rng.normal(0, 1, (c['nt'], V)) + 1j * rng.normal(0, 1, (c['nt'], V))).astype(np.complex64)
        out['sigma'] = np.full((c['nt'], V), c['sigma'], dtype=np.float32)
    out['offset'] = np.zeros_like(out['sigma'])
    return out


def trained_like_flat_params(seed=7):
    """Flat parameter vector with emission O(0.1-1) in places (same construction as the oracle's
    trained_like_params, restated so the product package does not import oracle/)."""
    import math
    shapes = [(21, 128), (128, 128), (128, 128), (149, 128), (128, 1)]
    rng = np.random.default_rng(seed)
    ks, bs = [], []
    for fi, fo in shapes:
        lim = math.sqrt(6.0 / fi)
        ks.append(rng.uniform(-lim, lim, size=(fi, fo)).astype(np.float32))
        bs.append(np.zeros((fo,), dtype=np.float32))
    rng2 = np.random.default_rng(seed + 1000)
    for i in range(5):
        bs[i] = rng2.normal(0, 0.1, bs[i].shape).astype(np.float32)
    ks[4] = (ks[4] * 6.0).astype(np.float32)
    bs[4] = np.full((1,), 9.0, dtype=np.float32)
    return np.concatenate([np.concatenate([k.reshape(-1), b]) for k, b in zip(ks, bs)])
