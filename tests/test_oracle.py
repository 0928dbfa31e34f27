"""CPU tests: the oracle (oracle/bhnerf_oracle.py) against the golden vectors produced by the
REFERENCE'S OWN SOURCE (tests/golden/ref_stages.npz, made by tests/golden/make_golden.py), and
against the reference itself when /root/reference is mounted (build container only)."""
import os

import numpy as np
import pytest
import torch

from oracle import bhnerf_oracle as O
from oracle import ref_shim

G = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def geo():
    return np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))


@pytest.fixture(scope='module')
def ref():
    return np.load(os.path.join(G, 'ref_stages.npz'))


def test_posenc_known_answer():
    # SURVEY s8.0 known-answer vector (reference source under numpy)
    want = [0.1, -0.2, 0.3, 0.0998334, -0.1986693, 0.2955202, 0.1986693, -0.3894183, 0.5646425, 0.3894183,
            -0.7173561, 0.9320391, 0.9950042, 0.9800666, 0.9553365, 0.9800666, 0.921061, 0.8253356, 0.921061,
            0.6967067, 0.3623577]
    got = O.posenc(torch.tensor([0.1, -0.2, 0.3], dtype=torch.float64), 3).numpy()
    np.testing.assert_allclose(got, want, atol=5e-8)


def test_posenc_vs_reference_golden(ref):
    got = O.posenc(torch.as_tensor(ref['posenc_x']), 3).numpy()
    np.testing.assert_allclose(got, ref['posenc'], rtol=0, atol=1e-13)


def test_posenc_ref32_vs_reference_float32_golden(ref):
    """posenc_ref32 = the reference source run with float32 promotion (what JAX executes), up to libm ulps;
    and it differs from the pure-f64 formula by the 100*pi-modulo rounding (~2e-5) on negative arguments."""
    x32 = torch.as_tensor(ref['posenc_x'].astype(np.float32))
    got = O.posenc_ref32(x32.to(torch.float64), 3).numpy()
    np.testing.assert_allclose(got, ref['posenc_f32'].astype(np.float64), rtol=0, atol=2.5e-7)
    pure = O.posenc(x32.to(torch.float64), 3).numpy()
    d = np.abs(got - pure).max()
    assert 5e-6 < d < 4e-5, d


def test_rotation_matrix_vs_reference_golden(ref):
    got = O.rotation_matrix([0, 0, 1], torch.tensor([0.3, -2.0, 40.0], dtype=torch.float64)).numpy()
    np.testing.assert_allclose(got, ref['rot'], rtol=0, atol=1e-14)


def test_warp_vs_reference_golden(geo, ref):
    w = O.velocity_warp_coords(geo['coords'].astype(np.float64), geo['Omega'].astype(np.float64), ref['t_frames'],
                               float(ref['t_start_obs']), geo['t_geos'].astype(np.float64), float(ref['t_injection']),
                               float(ref['GM_c3']), torch.float64).numpy()
    assert (np.isnan(w) == np.isnan(ref['warp'])).all()
    assert np.isnan(w).any() and not np.isnan(w).all()      # the t_M<0 mask is exercised
    np.testing.assert_allclose(np.nan_to_num(w), np.nan_to_num(ref['warp']), rtol=0, atol=1e-9)


def test_fill_and_ray_integral_vs_reference_golden(geo, ref):
    co = torch.as_tensor(geo['coords'].astype(np.float64))
    fill = O.fill_unsupervised_emission(torch.as_tensor(ref['e']), co, float(ref['rmin']), float(ref['rmax']),
                                        float(ref['z_width']))
    np.testing.assert_array_equal(fill.numpy(), ref['fill'])
    f64 = lambda k: torch.as_tensor(geo[k].astype(np.float64))
    rt = O.radiative_trasfer(fill, f64('g'), f64('dtau'), f64('Sigma')).numpy()
    np.testing.assert_allclose(rt, ref['rt'], rtol=1e-13)
    J = torch.as_tensor(ref['J'])
    rtJ = O.radiative_trasfer(J.unsqueeze(0) * fill.unsqueeze(1), f64('g'), f64('dtau'), f64('Sigma')).numpy()
    np.testing.assert_allclose(rtJ, ref['rtJ'], rtol=1e-12, atol=1e-15)


def test_adam_known_answer():
    kat = np.load(os.path.join(G, 'adam_kat.npz'))['traj']
    p0 = np.linspace(-1, 1, 7); mu = np.zeros(7); nu = np.zeros(7)
    for k in range(3):
        p0, mu, nu = O.adam_step(p0, np.cos(p0 * (k + 1)) * 0.1, mu, nu, k, 1e-2, 1e-4, 10)
        np.testing.assert_allclose(p0, kat[k], rtol=1e-14)
    # first Adam step moves every coordinate by exactly lr*sign(g) (bias-corrected m/sqrt(v) = +-1)
    p1, _, _ = O.adam_step(np.zeros(3), np.array([0.3, -2.0, 1e-3]), np.zeros(3), np.zeros(3), 0, 1e-2, 1e-4, 10)
    np.testing.assert_allclose(p1, -1e-2 * np.sign([0.3, -2.0, 1e-3]), rtol=1e-4)
    assert O.polynomial_schedule(0, 1e-4, 1e-6, 1, 100) == 1e-4
    assert abs(O.polynomial_schedule(100, 1e-4, 1e-6, 1, 100) - 1e-6) < 1e-18
    assert abs(O.polynomial_schedule(1000, 1e-4, 1e-6, 1, 100) - 1e-6) < 1e-18


def test_adam_restatement_agrees_with_an_independent_implementation():
    """optax is not installed, so the Adam + linear-schedule restatement (reference network.py:171-182) is additionally
    pinned on an INDEPENDENT implementation of the same published update: torch.optim.Adam (same bias correction, eps
    outside the square root) driven by a LambdaLR with the polynomial(power=1) schedule evaluated at the pre-increment
    count, as optax does."""
    rng = np.random.default_rng(3)
    p0 = rng.standard_normal(50)
    lr0, lr1, n = 3e-3, 1e-5, 7
    tp = torch.tensor(p0, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([tp], lr=1.0, betas=(0.9, 0.999), eps=1e-8)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda k: O.polynomial_schedule(k, lr0, lr1, 1, n))
    p, mu, nu = p0.copy(), np.zeros(50), np.zeros(50)
    for k in range(10):                                   # runs past num_iters: the schedule clips at lr_final
        g = np.sin(3.0 * p + k) * (1.0 + 0.1 * k)
        p, mu, nu = O.adam_step(p, g, mu, nu, k, lr0, lr1, n)
        opt.zero_grad()
        tp.grad = torch.tensor(np.sin(3.0 * tp.detach().numpy() + k) * (1.0 + 0.1 * k))
        opt.step(); sched.step()
        np.testing.assert_allclose(p, tp.detach().numpy(), rtol=1e-12, atol=1e-15)


def test_mlp_gradient_restatement_agrees_with_finite_differences(geo):
    """The oracle's parameter gradient is torch float64 autograd through the restated MLP / predictor / ray integral / loss;
    a central finite difference of the oracle's OWN loss along random parameter directions is an independent check of the
    pull-back (value_and_grad, reference network.py:617)."""
    d = np.load(os.path.join(G, 'case_lc_QU.npz'))
    params = O.unflatten_params(d['params_flat'])
    prd = dict(scale=float(d['scale']), rmin=float(d['rmin']), rmax=float(d['rmax']), z_width=float(d['z_width']))
    rt = dict(coords=geo['coords'], Omega=geo['Omega'], J=d['J'], g=geo['g'], dtau=geo['dtau'], Sigma=geo['Sigma'],
              t_start_obs=float(d['t_start_obs']), t_geos=geo['t_geos'], t_injection=float(d['t_injection']))
    args = ('image', 'lc', d['target'], d['sigma'], np.zeros_like(d['target']), d['t_frames'], rt, prd)
    ref = O.value_and_grad(params, *args, time_dtype=torch.float64)
    rng = np.random.default_rng(0)
    for _ in range(3):
        v = rng.standard_normal(d['params_flat'].shape)
        v /= np.linalg.norm(v)
        h = 1e-5
        lp = O.value_and_grad(O.unflatten_params(d['params_flat'] + h * v), *args, time_dtype=torch.float64)['loss']
        lm = O.value_and_grad(O.unflatten_params(d['params_flat'] - h * v), *args, time_dtype=torch.float64)['loss']
        fd = (lp - lm) / (2 * h)
        an = float(np.dot(ref['grads'], v))
        assert abs(fd - an) <= 1e-5 * max(abs(an), 1e-3 * np.linalg.norm(ref['grads'])), (fd, an)


def test_param_flatten_roundtrip():
    p = O.trained_like_params(3)
    flat = O.flatten_params(p)
    assert flat.shape == (O.N_PARAMS,) == (55169,)
    q = O.unflatten_params(flat)
    for i in range(5):
        np.testing.assert_array_equal(q['MLP_0'][f'Dense_{i}']['kernel'], p['MLP_0'][f'Dense_{i}']['kernel'])
    # Dense_3 rows 128.. are the posenc skip inputs (concat order network.py:61)
    assert q['MLP_0']['Dense_3']['kernel'].shape == (149, 128)


@pytest.mark.parametrize('case,kind,dtype_str', [('case_image_full', 'image', 'full'), ('case_lc_QU', 'image', 'lc'),
                                                 ('case_lc_IQU', 'image', 'lc'), ('case_vis', 'eht', 'vis')])
def test_full_path_golden_reproducible_and_fp32_noise_floor(geo, case, kind, dtype_str):
    """The committed full-path goldens are what the GPU tests compare against; re-derive them here and
    record the reference's own float32 noise floor (oracle fp32 vs fp64) next to the 1e-4 / 1e-3 tolerances."""
    d = np.load(os.path.join(G, case + '.npz'))
    params = O.unflatten_params(d['params_flat'])
    rt = dict(coords=geo['coords'], Omega=geo['Omega'], g=geo['g'], dtau=geo['dtau'], Sigma=geo['Sigma'],
              t_geos=geo['t_geos'], t_start_obs=float(d['t_start_obs']), t_injection=float(d['t_injection']),
              J=d['J'] if 'J' in d.files else 1.0)
    pred = dict(scale=float(d['scale']), rmin=float(d['rmin']), rmax=float(d['rmax']), z_width=float(d['z_width']))
    tgt = d['target']
    sig = d['sigma'] if 'sigma' in d.files else np.ones_like(tgt)
    third = d['A'] if kind == 'eht' else np.zeros_like(tgt)
    out = O.value_and_grad(params, kind, dtype_str, tgt, sig, third, d['t_frames'], rt, pred)
    np.testing.assert_allclose(out['images'], d['images'], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(out['grads'], d['grads'], rtol=1e-9, atol=1e-9 * np.abs(d['grads']).max())
    # all-float32 restatement (what the reference's JAX-CPU path computes) vs the committed golden
    o32 = O.value_and_grad(params, kind, dtype_str, tgt, sig, third, d['t_frames'], rt, pred, dtype=torch.float32)
    img_noise = np.abs(o32['images'] - d['images']).max() / np.abs(d['images']).max()
    grad_noise = np.abs(o32['grads'] - d['grads']).max() / np.abs(d['grads']).max()
    # pure-f64 formula (no float32 modulo rounding): how far the reference's own arithmetic is from exact math
    o64p = O.value_and_grad(params, kind, dtype_str, tgt, sig, third, d['t_frames'], rt, pred, feature_mode='pure')
    gp = np.abs(o64p['grads'] - d['grads']).max() / np.abs(d['grads']).max()
    print(f'{case}: float32 restatement vs golden: images {img_noise:.2e} grads {grad_noise:.2e}; '
          f'pure-f64 math vs golden: grads {gp:.2e}')
    assert img_noise < 1e-5 and grad_noise < 1e-4


@pytest.mark.skipif(not ref_shim.available(), reason='/root/reference not mounted (GPU box)')
def test_oracle_vs_reference_source_live(geo):
    """Execute the reference's own emission/kgeo/utils source (numpy shim) on fresh random inputs."""
    ns = ref_shim.load_bhnerf()
    rng = np.random.default_rng(123)
    co = geo['coords'].astype(np.float64); Om = geo['Omega'].astype(np.float64); tg = geo['t_geos'].astype(np.float64)
    tf = np.sort(rng.uniform(0, 2, 5)); c = O.GM_C3_SGRA_HR; t0 = 0.3; tinj = -990.0
    w_ref = ns.emission.velocity_warp_coords(co, Om, (tf - t0) / c, 0.0, tg, tinj, t_units=None, use_jax=True)
    w = O.velocity_warp_coords(co, Om, tf, t0, tg, tinj, c, torch.float64).numpy()
    assert (np.isnan(w) == np.isnan(w_ref)).all()
    np.testing.assert_allclose(np.nan_to_num(w), np.nan_to_num(w_ref), atol=1e-9)
    x = rng.uniform(-3, 3, (100, 3))
    np.testing.assert_allclose(O.posenc(torch.as_tensor(x), 3).numpy(), ns.posenc(x, 3), atol=1e-13)
    e = rng.uniform(0, 1, (2,) + co.shape[1:])
    np.testing.assert_array_equal(
        O.fill_unsupervised_emission(torch.as_tensor(e), torch.as_tensor(co), 2.5, 7.0, 3.0).numpy(),
        ns.emission.fill_unsupervised_emission(e, co, 2.5, 7.0, 3.0, use_jax=True))


def test_image_plane_dynamics_oracle_vs_reference_golden():
    """oracle.image_plane_dynamics against the output of the reference's OWN emission.image_plane_dynamics
    (tests/golden/grid_dynamics.npz:ref_images, generated by make_golden_grid.py under the numpy shim)."""
    d = np.load(os.path.join(G, 'grid_dynamics.npz'))
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    rt = dict(coords=geo['coords'], Omega=geo['Omega'], dtau=geo['dtau'], Sigma=geo['Sigma'], t_geos=geo['t_geos'],
              t_start_obs=float(d['t_start_obs']), t_injection=float(d['t_injection']), g=np.ones_like(geo['g']), J=1.0)
    out = O.image_plane_dynamics(d['emission_0'].astype(np.float64), float(d['fov']), rt, d['t_frames'])
    assert np.abs(out - d['ref_images']).max() / np.abs(d['ref_images']).max() < 1e-6       # grid stored as float32
    assert np.abs(d['ref_images']).max() > 0.1


def test_jax_map_coordinates_restatement():
    """Known answers of jax.scipy.ndimage.map_coordinates(order=1, cval=0): equals scipy inside the grid, and blends
    with cval corner by corner up to one cell outside (where scipy's 'constant' mode already returns cval)."""
    import scipy.ndimage
    import torch
    rng = np.random.default_rng(0)
    grid = rng.normal(size=(5, 6, 7))
    pts = rng.uniform(0, 4, size=(3, 50)); pts[1] *= 5 / 4; pts[2] *= 6 / 4
    ours = O.jax_map_coordinates_order1(torch.as_tensor(grid), torch.as_tensor(pts)).numpy()
    np.testing.assert_allclose(ours, scipy.ndimage.map_coordinates(grid, pts, order=1, cval=0.0), rtol=1e-12, atol=1e-14)
    edge = torch.tensor([[-0.25, 4.5, -1.0, 2.0], [0.0, 0.0, 0.0, 5.75], [0.0, 0.0, 0.0, 6.5]], dtype=torch.float64)
    got = O.jax_map_coordinates_order1(torch.as_tensor(grid), edge).numpy()
    want = [0.75 * grid[0, 0, 0], 0.5 * grid[4, 0, 0], 0.0, 0.25 * 0.5 * grid[2, 5, 6]]
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)
    assert scipy.ndimage.map_coordinates(grid, edge[:, :1].numpy(), order=1, cval=0.0)[0] == 0.0


def test_grid_predictor_oracle_golden_is_reproducible():
    import torch
    d = np.load(os.path.join(G, 'grid_predictor.npz'))
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    rt = dict(coords=geo['coords'], Omega=geo['Omega'], dtau=geo['dtau'], Sigma=geo['Sigma'], t_geos=geo['t_geos'],
              g=geo['g'], J=1.0, t_start_obs=float(d['t_start_obs']), t_injection=float(d['t_injection']))
    pred = {k: float(d[k]) for k in ('scale', 'rmin', 'rmax', 'z_width')}
    img = O.grid_predictor_images(torch.as_tensor(d['grid'].astype(np.float64)), d['t_frames'], rt, pred).numpy()
    np.testing.assert_allclose(img, d['images'], rtol=1e-10, atol=1e-12)
    assert (img[0] == 0).all() and img[-1].max() > 1.0        # first frame is entirely before injection


def _pol_case(d, tag):
    a, inc, rmin, rmax, zw = d[tag + '_consts']
    shape = d[tag + '_r'].shape
    bc = lambda v: np.broadcast_to(d[tag + '_' + v][..., None], shape)
    geos = dict(r=d[tag + '_r'], theta=d[tag + '_theta'], affine=d[tag + '_affine'], lam=bc('lam'), eta=bc('eta'),
                alpha=bc('alpha'), beta=bc('beta'), spin=float(a), inc=float(inc), M=1.0)
    Om = np.sign(a + np.finfo(float).eps) / (geos['r'] ** 1.5 + a)
    return geos, Om, float(rmin), float(rmax), float(zw)


def test_polarization_factors_oracle_vs_reference_functions():
    """oracle.polarization_factors (restatement of alma.image_plane_model's J chain, bhnerf/alma.py:47-60 +
    bhnerf/kgeo.py:199-248,274-313,438-519) against the committed output of the reference's OWN functions
    (tests/golden/pol_factors.npz, made by make_golden_pol.py through the mini-xarray of oracle/ref_shim.py), on real Kerr
    geodesics of two geometries and two field configurations -- and live against the reference when it is present."""
    from oracle import bhnerf_oracle as O
    d = np.load(os.path.join(G, 'pol_factors.npz'))
    for tag in ('a', 'b'):
        geos, Om, rmin, rmax, zw = _pol_case(d, tag)
        dom = (np.abs(geos['r'] * np.cos(geos['theta'])) < zw) & (geos['r'] > rmin) & (geos['r'] < rmax)
        assert dom.sum() > 100
        for j in (0, 1):
            b = d['%s_b%d' % (tag, j)]
            J = O.polarization_factors(geos, Om, dict(arad=b[0], avert=b[1], ator=b[2]), 0.5, rmin, rmax, zw)
            ref = d['%s_J%d' % (tag, j)]
            assert J.shape == ref.shape == (3,) + geos['r'].shape
            assert np.abs(J[:, dom] - ref[:, dom]).max() / np.abs(ref[:, dom]).max() < 1e-12
            sane = np.abs(ref) < 1e6                      # near the horizon the factors blow up (and are culled by the domain)
            assert np.abs(J - ref)[sane].max() / np.abs(ref[sane]).max() < 1e-9
            assert np.abs(ref[1:, dom]).max() > 0.05 and np.abs(ref[2, dom]).max() > 1e-3     # Q and U are exercised
    from oracle import ref_shim
    if ref_shim.available():
        g = ref_shim.kerr_geodesics(0.5, np.deg2rad(40.0), 20.0, 8, 8, 16)
        Om = ref_shim.keplerian_omega(g)
        bcs = dict(arad=0.2, avert=0.7, ator=-0.4)
        Jr, _ = ref_shim.reference_polarization_factors(g, Om, bcs, 0.3, 4.5, 10.0, 3.0)
        Jo = O.polarization_factors(g, Om, bcs, 0.3, 4.5, 10.0, 3.0)
        dom = (np.abs(g['z']) < 3.0) & (g['r'] > 4.5) & (g['r'] < 10.0)
        assert np.abs(Jr[:, dom] - Jo[:, dom]).max() / np.abs(Jr[:, dom]).max() < 1e-12
