"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(libbhnerf_b200.so via ctypes); the checker is the committed float64-oracle golden output
(tests/golden/case_*.npz) or the oracle run live on small seeded inputs.

Tolerances (BASELINE.json north_star): images / lightcurves / visibilities <= 1e-4 relative,
gradients <= 1e-3 relative, measured as max|a-b| / max|b| over the whole array."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), 'golden')
IMG_TOL, GRAD_TOL = 1e-4, 1e-3
CASES = ['case_image_full', 'case_lc_QU', 'case_lc_IQU', 'case_vis', 'case_amp', 'case_cphase']


def _impls():
    return ['simt', 'tc']


@pytest.fixture(scope='module', autouse=True)
def _lib(built_lib):
    assert torch.cuda.is_available()
    import ctypes
    sm = ctypes.c_int(); mj = ctypes.c_int(); mn = ctypes.c_int()
    rc = built_lib.bhnerf_device_check(ctypes.byref(sm), ctypes.byref(mj), ctypes.byref(mn))
    assert rc == 0, built_lib.bhnerf_last_error()
    assert mj.value == 10
    return built_lib


@pytest.mark.parametrize('impl', _impls())
@pytest.mark.parametrize('case', CASES)
def test_golden_case(case, impl):
    from bhnerf_b200 import testing
    r = testing.run_golden_case(case, impl=impl)
    print(case, impl, r)
    assert r['img_err'] < IMG_TOL, r
    assert r['loss_err'] < IMG_TOL, r
    assert r['grad_err'] < GRAD_TOL, r
    if 'vis_err' in r:
        assert r['vis_err'] < IMG_TOL, r


@pytest.mark.parametrize('impl', _impls())
def test_frame_chunking_and_recompute_give_same_gradient(impl):
    """A workspace that only holds one frame (chunked fused step) and a backward that recomputes its
    residuals must reproduce the all-at-once result."""
    from bhnerf_b200 import engine, testing
    scene, d = testing.load_golden_scene('case_lc_IQU')
    dev = scene.device
    params = torch.as_tensor(d['params_flat'], device=dev)
    tf = torch.as_tensor(d['t_frames'].astype(np.float32), device=dev)
    off = np.zeros_like(d['target'])
    l0, i0, g0 = engine.train_step_image(scene, params, tf, d['target'], d['sigma'], off, 1.0, 'lc', impl)
    l1, i1, g1 = engine.train_step_image(scene, params, tf, d['target'], d['sigma'], off, 1.0, 'lc', impl,
                                         max_workspace=1)
    assert torch.equal(i0, i1)
    assert testing.rel_err(g1.cpu().numpy(), g0.cpu().numpy()) < 2e-5
    assert abs(l0.item() - l1.item()) / abs(l0.item()) < 1e-5
    # unfused: fwd -> loss -> bwd with saved residuals, and with recompute
    images, e, acts = engine.render_fwd(scene, params, tf, impl, save_acts=True)
    assert torch.equal(images, i0)
    loss, dI = engine.loss_image(images, d['target'], d['sigma'], off, 1.0, 'lc')
    g2 = engine.render_bwd(scene, params, tf, dI, e, acts, impl)
    g3 = engine.render_bwd(scene, params, tf, dI, None, None, impl)
    g4 = engine.render_bwd(scene, params, tf, dI, None, None, impl, max_workspace=1)
    for g in (g2, g3, g4):
        assert testing.rel_err(g.cpu().numpy(), g0.cpu().numpy()) < 2e-5


def test_prepack_compaction_matches_numpy():
    from bhnerf_b200 import testing
    scene, d = testing.load_golden_scene('case_lc_QU')
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    co = geo['coords'].reshape(3, -1, 32)
    r2 = (co.astype(np.float64) ** 2).sum(0)
    w = geo['g'].astype(np.float32) ** 2 * geo['dtau'] * geo['Sigma']
    act = ~((r2 < float(d['rmin']) ** 2) | (r2 > float(d['rmax']) ** 2) | (np.abs(co[2]) > float(d['z_width'])))
    act &= (w.reshape(-1, 32) != 0)
    assert abs(scene.n_active - int(act.sum())) <= 2        # fp32 vs fp64 r^2 at the domain boundary
    rp = scene.row_ptr.cpu().numpy()
    assert rp[0] == 0 and rp[-1] == scene.n_active and (np.diff(rp) >= 0).all()
    np.testing.assert_allclose(np.diff(rp), act.sum(1), atol=1)
    di = scene.dense_index.cpu().numpy()
    assert (np.diff(di) > 0).all()                          # ray-major, k ascending = dense order
    ray = scene.ray_index.cpu().numpy()
    assert (ray[:scene.n_active] == di // 32).all() and (ray[scene.n_active:] == -1).all()
    assert scene.n_pad % 128 == 0 and scene.n_pad - scene.n_active < 128


def test_standalone_stages_vs_reference_goldens():
    """emission.velocity_warp_coords / fill_unsupervised_emission / kgeo.radiative_trasfer against the
    outputs of the reference's own source (ref_stages.npz)."""
    from bhnerf_b200 import emission, kgeo
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    ref = np.load(os.path.join(G, 'ref_stages.npz'))
    w = emission.velocity_warp_coords(geo['coords'], geo['Omega'], ref['t_frames'], float(ref['t_start_obs']),
                                      geo['t_geos'], float(ref['t_injection']), t_units='hr').cpu().numpy()
    assert (np.isnan(w) == np.isnan(ref['warp'])).mean() > 0.9999      # fp32 t_M exactly at 0 may flip
    both = ~np.isnan(w) & ~np.isnan(ref['warp'])
    # fp32 theta = t_M*Omega carries ~ulp(t_M)*Omega ~ 1e-5 rad (the reference's own fp32 noise)
    assert np.abs(w[both] - ref['warp'][both]).max() < 2e-3
    fill = emission.fill_unsupervised_emission(ref['e'].astype(np.float32), geo['coords'], float(ref['rmin']),
                                               float(ref['rmax']), float(ref['z_width'])).cpu().numpy()
    assert (fill == ref['fill'].astype(np.float32)).mean() > 0.9999
    rt = kgeo.radiative_trasfer(ref['fill'].astype(np.float32), geo['g'], geo['dtau'], geo['Sigma']).cpu().numpy()
    np.testing.assert_allclose(rt, ref['rt'], rtol=2e-5, atol=1e-6 * np.abs(ref['rt']).max())
    Je = (ref['J'][None] * ref['fill'][:, None]).astype(np.float32)
    rtJ = kgeo.radiative_trasfer(Je, geo['g'], geo['dtau'], geo['Sigma']).cpu().numpy()
    np.testing.assert_allclose(rtJ, ref['rtJ'], rtol=1e-4, atol=2e-6 * np.abs(ref['rtJ']).max())


def test_adam_kernel_vs_oracle():
    from bhnerf_b200 import engine
    from oracle import bhnerf_oracle as O
    rng = np.random.default_rng(0)
    n = 55169
    p = rng.normal(0, 0.1, n); mu = np.zeros(n); nu = np.zeros(n)
    pd = torch.as_tensor(p.astype(np.float32)).cuda(); mud = torch.zeros(n, device='cuda'); nud = torch.zeros(n, device='cuda')
    for k in range(4):
        g = rng.normal(0, 1, n) * 10.0 ** rng.uniform(-6, 2, n)
        gs = 0.5 if k % 2 else 1.0
        p, mu, nu = O.adam_step(p, g.astype(np.float32).astype(np.float64) * gs, mu, nu, k, 1e-3, 1e-5, 3)
        engine.adam_step(pd, torch.as_tensor(g.astype(np.float32)).cuda(), mud, nud, k, 1e-3, 1e-5, 3, grad_scale=gs)
    np.testing.assert_allclose(pd.cpu().numpy(), p, rtol=0, atol=2e-6)
    np.testing.assert_allclose(mud.cpu().numpy(), mu, rtol=2e-5, atol=2e-6 * np.abs(mu).max())


@pytest.mark.parametrize('impl', _impls())
def test_reference_api_train_step_matches_oracle(impl):
    """TrainStep.image -> network.gradient_step_image (the reference's call stack, optimization.py:176,
    network.py:566-622): loss, images and the Adam-updated parameters against the oracle."""
    from bhnerf_b200 import network, optimization
    from oracle import bhnerf_oracle as O
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    d = np.load(os.path.join(G, 'case_image_full.npz'))
    pred = network.NeRF_Predictor(float(d['scale']), float(d['rmin']), float(d['rmax']), float(d['z_width']))
    params = network.unflatten_params(d['params_flat'])
    state = pred.init_state(params, num_iters=100, lr_init=1e-3, lr_final=1e-5)
    rt = {'coords': geo['coords'], 'Omega': geo['Omega'], 'J': 1.0, 'g': geo['g'], 'dtau': geo['dtau'],
          'Sigma': geo['Sigma'], 't_start_obs': float(d['t_start_obs']), 't_geos': geo['t_geos'],
          't_injection': float(d['t_injection'])}
    from collections import OrderedDict
    rt = OrderedDict(rt)
    ts = optimization.TrainStep.image(d['t_frames'], d['target'], sigma=1.0, dtype='full')
    os.environ['BHNERF_IMPL'] = impl
    try:
        loss, state, images = ts(state, rt, np.arange(4))
    finally:
        os.environ.pop('BHNERF_IMPL')
    assert abs(loss.item() - float(d['loss'])) / float(d['loss']) < IMG_TOL
    assert tuple(images.shape) == (4, 16, 16)
    assert np.abs(images.cpu().numpy() - d['images']).max() / np.abs(d['images']).max() < IMG_TOL
    want, _, _ = O.adam_step(d['params_flat'].astype(np.float64), d['grads'], np.zeros(55169), np.zeros(55169), 0,
                             1e-3, 1e-5, 100)
    got = state.flat.cpu().numpy()
    # first Adam step moves each weight by ~lr*sign(g): compare the UPDATE, not the weights
    upd_err = np.abs((got - d['params_flat']) - (want - d['params_flat'])).max() / 1e-3
    assert upd_err < 2e-2, upd_err
    assert state.step == 1
    # forward-only path (test_image / total_movie_loss), reference semantics loss/nt
    tot, frames = optimization.total_movie_loss(2, pred.init_state(params), ts, rt, return_frames=True)
    assert abs(tot - float(d['loss']) / 4) / (float(d['loss']) / 4) < IMG_TOL
    assert frames.shape == (4, 16, 16)


@pytest.mark.parametrize('dtype', ['vis', 'amp', 'cphase'])
def test_reference_api_eht_step_matches_oracle(dtype):
    """TrainStep.eht -> network.gradient_step_eht / test_eht (optimization.py:218-268, network.py:624-682, :741-795):
    loss and the first Adam update against the oracle's value_and_grad for every eht dtype."""
    from collections import OrderedDict
    from bhnerf_b200 import network, optimization
    from oracle import bhnerf_oracle as O
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    d = np.load(os.path.join(G, 'case_%s.npz' % dtype))
    pred = network.NeRF_Predictor(float(d['scale']), float(d['rmin']), float(d['rmax']), float(d['z_width']))
    params = network.unflatten_params(d['params_flat'])
    state = pred.init_state(params, num_iters=100, lr_init=1e-3, lr_final=1e-5)
    rt = OrderedDict(coords=geo['coords'], Omega=geo['Omega'], J=1.0, g=geo['g'], dtau=geo['dtau'], Sigma=geo['Sigma'],
                     t_start_obs=float(d['t_start_obs']), t_geos=geo['t_geos'], t_injection=float(d['t_injection']))
    ts = optimization.TrainStep.eht_arrays(d['t_frames'], d['target'], d['sigma'], d['A'], dtype=dtype)
    loss0, _, images = ts(state, rt, np.arange(4), update_state=False)
    assert abs(loss0.item() - float(d['loss'])) / abs(float(d['loss'])) < IMG_TOL
    assert np.abs(images.cpu().numpy() - d['images'].reshape(images.shape)).max() / np.abs(d['images']).max() < IMG_TOL
    loss, state, images = ts(state, rt, np.arange(4))
    assert abs(loss.item() - float(d['loss'])) / abs(float(d['loss'])) < IMG_TOL
    want, _, _ = O.adam_step(d['params_flat'].astype(np.float64), d['grads'], np.zeros(55169), np.zeros(55169), 0,
                             1e-3, 1e-5, 100)
    got = state.flat.cpu().numpy()
    upd_err = np.abs((got - d['params_flat']) - (want - d['params_flat'])).max() / 1e-3
    assert upd_err < 2e-2, upd_err
    with pytest.raises(AttributeError):
        optimization.TrainStep.eht_arrays(d['t_frames'], d['target'], d['sigma'], d['A'], dtype='bogus')(state, rt, np.arange(4))


@pytest.mark.parametrize('impl', _impls())
def test_polarized_visibilities_match_oracle(impl):
    """A with a polarization axis (optimization.py:235-251 stacks I,Q,U on axis 1; network.py:542-548 multiplies each
    Stokes image with its own DFT matrix): loss, visibilities, images and the first Adam update through
    TrainStep.eht_arrays -> gradient_step_eht against the float64 oracle golden case_vis_IQU."""
    from collections import OrderedDict
    from bhnerf_b200 import engine, network, optimization
    from oracle import bhnerf_oracle as O
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    d = np.load(os.path.join(G, 'case_vis_IQU.npz'))
    pred = network.NeRF_Predictor(float(d['scale']), float(d['rmin']), float(d['rmax']), float(d['z_width']))
    state = pred.init_state(network.unflatten_params(d['params_flat']), num_iters=100, lr_init=1e-3, lr_final=1e-5)
    rt = OrderedDict(coords=geo['coords'], Omega=geo['Omega'], J=d['J'], g=geo['g'], dtau=geo['dtau'], Sigma=geo['Sigma'],
                     t_start_obs=float(d['t_start_obs']), t_geos=geo['t_geos'], t_injection=float(d['t_injection']))
    assert d['A'].shape == (4, 3, 12, 256) and d['target'].shape == (4, 3, 12)
    # the visibilities themselves, through the C ABI with the pol axis folded into the frame axis
    scene = network._scene_for(pred, *[rt[k] for k in ('coords', 'Omega', 'J', 'g', 'dtau', 'Sigma', 't_start_obs', 't_geos',
                                                       't_injection')], 'hr', device=state.flat.device)
    images, _, _ = engine.render_fwd(scene, state.flat, d['t_frames'].astype(np.float32), impl)
    vis = engine.vis_fwd(engine._c64(d['A'], scene.device).reshape(12, 12, 256), images.reshape(12, 1, 256))
    assert np.abs(vis.cpu().numpy().reshape(4, 3, 12) - d['vis']).max() / np.abs(d['vis']).max() < IMG_TOL
    ts = optimization.TrainStep.eht_arrays(d['t_frames'], d['target'], d['sigma'], d['A'], dtype='vis')
    loss0, _, images = network.test_eht(state, 'hr', 'vis', d['target'], d['sigma'], d['A'], d['t_frames'], *rt.values(), 1.0,
                                        impl=impl)
    assert abs(loss0.item() - float(d['loss'])) / float(d['loss']) < IMG_TOL
    assert tuple(images.shape) == (4, 3, 16, 16)
    assert np.abs(images.cpu().numpy() - d['images']).max() / np.abs(d['images']).max() < IMG_TOL
    loss, state, _ = network.gradient_step_eht(state, 'hr', 'vis', d['target'], d['sigma'], d['A'], d['t_frames'],
                                               *rt.values(), 1.0, impl=impl)
    assert abs(loss.item() - float(d['loss'])) / float(d['loss']) < IMG_TOL
    want, _, _ = O.adam_step(d['params_flat'].astype(np.float64), d['grads'], np.zeros(55169), np.zeros(55169), 0,
                             1e-3, 1e-5, 100)
    got = state.flat.cpu().numpy()
    upd_err = np.abs((got - d['params_flat']) - (want - d['params_flat'])).max() / 1e-3
    assert upd_err < 2e-2, upd_err
    loss_ts, _, _ = ts(state, rt, np.arange(4), update_state=False)      # the TrainStep route accepts the same arrays
    assert np.isfinite(loss_ts.item())
    with pytest.raises(AttributeError):      # Stokes images but an A without the pol axis: the reference's matmul cannot broadcast
        network.test_eht(state, 'hr', 'vis', d['target'][:, 0], d['sigma'][:, 0], d['A'][:, 0], d['t_frames'], *rt.values(), 1.0)


def test_predictor_apply_and_sample_3d_grid_vs_oracle():
    from bhnerf_b200 import network
    from oracle import bhnerf_oracle as O
    d = np.load(os.path.join(G, 'case_image_full.npz'))
    pred = network.NeRF_Predictor(8.0, 2.0, 8.0, 4.0)
    params = network.unflatten_params(d['params_flat'])
    e = network.sample_3d_grid(pred.apply, params, fov=16.0, resolution=12)
    g1 = np.linspace(-8, 8, 12)
    coords = np.array(np.meshgrid(g1, g1, g1, indexing='ij'))
    want = O.predict_emission(O._params_t(params, torch.float64), np.array([0.0]), coords.astype(np.float32), 0.0, 0.0,
                              0.0, 0.0, 8.0, 2.0, 8.0, 4.0, GM_c3=1.0)[0].numpy()
    assert e.shape == (12, 12, 12)
    assert (e == 0).sum() == (want == 0).sum()
    assert np.abs(e - want).max() / want.max() < IMG_TOL


def test_vis_head_is_adjoint_pair():
    from bhnerf_b200 import engine
    rng = np.random.default_rng(1)
    Bt, V, P = 3, 37, 1024
    A = torch.as_tensor((rng.normal(size=(Bt, V, P)) + 1j * rng.normal(size=(Bt, V, P))).astype(np.complex64)).cuda()
    x = torch.as_tensor(rng.normal(size=(Bt, 1, P)).astype(np.float32)).cuda()
    y = torch.as_tensor((rng.normal(size=(Bt, V)) + 1j * rng.normal(size=(Bt, V))).astype(np.complex64)).cuda()
    Ax = engine.vis_fwd(A, x)
    ref = torch.einsum('bvp,bp->bv', A.to(torch.complex128), x[:, 0].to(torch.complex128))
    assert (Ax.to(torch.complex128) - ref).abs().max() / ref.abs().max() < 1e-5
    Aty = engine.vis_bwd(A, y, P)
    lhs = (Ax.conj() * y).real.sum().item()               # <Ax, y>
    rhs = (x * Aty).sum().item()                          # <x, Re(A^H y)>
    assert abs(lhs - rhs) / abs(lhs) < 1e-4
    # 'amp' loss gradient against finite differences of its own value (targets near |Ax| keep the fp32 loss small
    # enough for a difference quotient to resolve)
    amp = Ax.abs().cpu().numpy()
    t = (amp * (1 + 0.01 * rng.normal(size=(Bt, V)))).astype(np.float32); s = np.full((Bt, V), 0.3, np.float32)
    l0, dv = engine.loss_vis(Ax, t, s, 1.0, 'amp')
    h = 0.05
    pert = torch.zeros_like(Ax); pert[1, 5] = h + 0j
    l1, _ = engine.loss_vis(Ax + pert, t, s, 1.0, 'amp')
    l2, _ = engine.loss_vis(Ax - pert, t, s, 1.0, 'amp')
    fd = (l1 - l2).item() / (2 * h)
    assert abs(fd - dv[1, 5].real.item()) / abs(dv[1, 5].real.item()) < 2e-2, (fd, dv[1, 5])


@pytest.mark.parametrize('impl', _impls())
def test_size_independent_properties_at_config_shapes(impl):
    """BASELINE.json shapes (cfg1 full size; cfg2 with 6 of its 100 frames): properties that need no oracle.
    (1) frames before injection render exactly 0; (2) images are linear in the Stokes factors J;
    (3) a zero cotangent gives a zero gradient; (4) the render is bitwise deterministic;
    (5) the fused step's gradient equals the unfused fwd/loss/bwd chain."""
    from bhnerf_b200 import engine, synthetic
    from bhnerf_b200 import constants
    c = synthetic.make_config('cfg2_lp_flare', nt=6)
    rt, pr = c['rt'], c['predictor']
    params = torch.as_tensor(synthetic.trained_like_flat_params(7)).cuda()
    mk = lambda J, t0: engine.PackedScene(rt['coords'], rt['Omega'], J, rt['g'], rt['dtau'], rt['Sigma'], rt['t_geos'],
                                          t0, rt['t_injection'], pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'],
                                          constants.GM_c3(t_units='hr'))
    scene = mk(rt['J'], rt['t_start_obs'])
    assert 0.05 < scene.n_active / (c['P'] * c['G']) < 0.3
    tf = torch.as_tensor(c['t_frames']).cuda()
    img, e, acts = engine.render_fwd(scene, params, tf, impl, save_acts=True)
    img2, _, _ = engine.render_fwd(scene, params, tf, impl)
    assert torch.equal(img, img2) and torch.isfinite(img).all() and img.abs().max() > 0
    scene2 = mk(2.5 * rt['J'], rt['t_start_obs'])
    img3, _, _ = engine.render_fwd(scene2, params, tf, impl)
    assert (img3 - 2.5 * img).abs().max() / img.abs().max() < 1e-5
    early = mk(rt['J'], rt['t_start_obs'] + 100.0)         # every frame is > 1e4 M before injection
    img4, _, _ = engine.render_fwd(early, params, tf, impl)
    assert (img4 == 0).all()
    g0 = engine.render_bwd(scene, params, tf, torch.zeros_like(img), e, acts, impl)
    assert (g0 == 0).all()
    off = np.zeros_like(c['target'])
    loss, images, grads = engine.train_step_image(scene, params, tf, c['target'], c['sigma'], off, 1.0, 'lc', impl)
    l2, dI = engine.loss_image(img, c['target'], c['sigma'], off, 1.0, 'lc')
    g2 = engine.render_bwd(scene, params, tf, dI, e, acts, impl)
    assert torch.equal(images, img)
    assert abs(loss.item() - l2.item()) / abs(l2.item()) < 1e-5
    assert (grads - g2).abs().max() / g2.abs().max() < 2e-5 and torch.isfinite(grads).all()


@pytest.mark.parametrize('nt', [8, 1, 3])
def test_simt_and_tc_agree_at_config_shape(nt):
    """The fp32 SIMT family is the on-device reference for the tcgen05 family at sizes the CPU oracle cannot
    reach: cfg1 (64x64x64, 461 tiles per frame) with 8 frames, and with 1 / 3 frames -- an ODD number of tiles, so the
    last round of the persistent kernels has one tile (fused backward: one-plane plan of the per-pixel loss)."""
    from bhnerf_b200 import constants, engine, synthetic
    c = synthetic.make_config('cfg1_tutorial3', nt=nt)
    rt, pr = c['rt'], c['predictor']
    params = torch.as_tensor(synthetic.trained_like_flat_params(7)).cuda()
    scene = engine.PackedScene(rt['coords'], rt['Omega'], rt['J'], rt['g'], rt['dtau'], rt['Sigma'], rt['t_geos'],
                               rt['t_start_obs'], rt['t_injection'], pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'],
                               constants.GM_c3(t_units='hr'))
    tf = torch.as_tensor(c['t_frames']).cuda()
    off = np.zeros_like(c['target'])
    ls, is_, gs = engine.train_step_image(scene, params, tf, c['target'], c['sigma'], off, 1.0, 'full', 'simt')
    lt, it_, gt = engine.train_step_image(scene, params, tf, c['target'], c['sigma'], off, 1.0, 'full', 'tc')
    assert (it_ - is_).abs().max() / is_.abs().max() < IMG_TOL
    assert (gt - gs).abs().max() / gs.abs().max() < GRAD_TOL
    assert abs(lt.item() - ls.item()) / ls.item() < IMG_TOL


def test_fp16_operand_overflow_is_flagged_not_hidden():
    """The tcgen05 forward rounds hidden activations to fp16 operands.  Weights that push |h| past 65504 must raise
    through bhnerf_workspace_status (status word 3), healthy weights must leave every flag clear, the flag is STICKY (a
    later healthy step does not hide it; it is reported once, then cleared), and the guarded Adam update refuses the
    gradient of a flagged step."""
    from bhnerf_b200 import _lib, engine, testing
    scene, d = testing.load_golden_scene('case_image_full')
    tf = torch.as_tensor(d['t_frames'].astype(np.float32)).cuda()
    good = torch.as_tensor(d['params_flat']).cuda()
    engine.render_fwd(scene, good, tf, 'tc')
    assert engine.workspace_status(impl='tc')[:5] == [0, 0, 0, 0, 0]
    bad = good.clone()
    bad[: 21 * 128 + 128] *= 3.0e4            # W0, b0: h0 ~ 3e4 x O(1..10)
    bad[21 * 128 + 128: 21 * 128 + 128 + 128 * 128] *= 30.0
    engine.render_fwd(scene, bad, tf, 'tc')
    engine.render_fwd(scene, good, tf, 'tc')  # a healthy step in between: the per-step words are reset, the sticky ones not
    with pytest.raises(_lib.BhnerfError, match='fp16 operand range'):
        engine.workspace_status(impl='tc')
    assert engine.workspace_status(impl='tc')[:5] == [0, 0, 0, 0, 0]      # reported once
    # guarded Adam: the update after a flagged step leaves the parameters alone, after a healthy step it moves them
    tgt = d['target']; one = np.ones_like(tgt); zero = np.zeros_like(tgt)
    for params, moves in ((bad, False), (good, True)):
        p = params.clone(); mu = torch.zeros_like(p); nu = torch.zeros_like(p)
        _, _, grads = engine.train_step_image(scene, p, tf, tgt, one, zero, 1.0, 'full', 'tc')
        engine.adam_step(p, grads, mu, nu, 0, 1e-3, 1e-5, 10, guard=engine.step_guard(scene.device, 'tc'))
        torch.cuda.synchronize()
        assert bool((p != params).any().item()) == moves
    assert engine.workspace_status(impl='tc', raise_on_error=False)[3] == 1   # the flagged train step is on record
    assert engine.workspace_status(impl='tc')[:5] == [0, 0, 0, 0, 0]


def _geo_ns(geo, with_g=True):
    import types
    ns = types.SimpleNamespace(x=geo['coords'][0], y=geo['coords'][1], z=geo['coords'][2], t=geo['t_geos'],
                               dtau=geo['dtau'], Sigma=geo['Sigma'])
    if with_g:
        ns.g = geo['g']
    return ns


def test_image_plane_dynamics_vs_reference_and_oracle():
    """emission.image_plane_dynamics (csrc/grid.cu, mode 0) against the output of the reference's own function
    (doppler=False) and against the float64 oracle with Doppler factor, Stokes factors and pre-injection frames."""
    from bhnerf_b200 import constants, emission
    d = np.load(os.path.join(G, 'grid_dynamics.npz'))
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    tfM = (d['t_frames'] - float(d['t_start_obs'])) / float(d['GM_c3'])        # plain numbers = units of M (reference)
    fov = float(d['fov'])
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    out = emission.image_plane_dynamics(d['emission_0'], _geo_ns(geo), geo['Omega'], tfM, float(d['t_injection']),
                                        t_start_obs=0.0, doppler=False, fov=fov).cpu().numpy()
    assert out.shape == d['ref_images'].shape and rel(out, d['ref_images']) < IMG_TOL, rel(out, d['ref_images'])
    out = emission.image_plane_dynamics(d['emission_0'], _geo_ns(geo), geo['Omega'], tfM, float(d['t_injection']),
                                        t_start_obs=0.0, fov=fov).cpu().numpy()
    assert rel(out, d['images_g']) < IMG_TOL, rel(out, d['images_g'])
    out = emission.image_plane_dynamics(d['emission_0'], _geo_ns(geo), geo['Omega'], tfM, float(d['t_injection']),
                                        J=d['J'], t_start_obs=0.0, fov=fov).cpu().numpy()
    assert out.shape == d['images_J'].shape and rel(out, d['images_J']) < IMG_TOL, rel(out, d['images_J'])
    # hours + t_units (same conversion as the train path), frames before injection render 0
    out = emission.image_plane_dynamics(d['emission_0'], _geo_ns(geo), geo['Omega'], d['t_frames'],
                                        float(d['t_injection_early']), t_start_obs=float(d['t_start_early']), fov=fov,
                                        t_units='hr').cpu().numpy()
    assert abs(constants.GM_c3(t_units='hr') - float(d['GM_c3'])) < 1e-9
    assert rel(out, d['images_early']) < IMG_TOL and (out[0] == 0).all(), rel(out, d['images_early'])
    # the stand-alone stages compose to the same thing: velocity_warp_coords -> interpolate_coords -> radiative_trasfer
    from bhnerf_b200 import kgeo
    warped = emission.velocity_warp_coords(geo['coords'], geo['Omega'], tfM, 0.0, geo['t_geos'], float(d['t_injection']))
    e = emission.interpolate_coords(d['emission_0'], warped, fov=fov)
    import scipy.ndimage
    ic = np.moveaxis((warped.cpu().numpy().astype(np.float64) + fov / 2) / fov * (d['emission_0'].shape[0] - 1), -1, 0)
    e_ref = scipy.ndimage.map_coordinates(d['emission_0'].astype(np.float64), ic, order=1, cval=0.0)
    assert e.shape == e_ref.shape and np.abs(e.cpu().numpy() - e_ref).max() < 2e-4 * np.abs(e_ref).max()
    staged = kgeo.radiative_trasfer(e, geo['g'], geo['dtau'], geo['Sigma']).cpu().numpy()
    assert rel(staged, d['images_g']) < 5e-4
    # flat-space propagation of the grid itself (propogate_flatspace_emission): identity at t = t_start, finite and
    # mass-preserving to a few per cent later (a rigid-ish shear of a compact blob well inside the grid)
    ax = np.linspace(-fov / 2, fov / 2, d['emission_0'].shape[0])
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing='ij')
    Om3 = 1.0 / (np.maximum(np.sqrt(X ** 2 + Y ** 2 + Z ** 2), 2.0) ** 1.5)
    blob = np.exp(-((X - 3.0) ** 2 + Y ** 2 + Z ** 2) / 2.0).astype(np.float32)
    mov3 = emission.propogate_flatspace_emission(blob, Om3, np.array([0.0, 20.0]), fov=fov).cpu().numpy()
    assert mov3.shape == (2,) + blob.shape and np.abs(mov3[0] - blob).max() < 1e-5
    assert abs(mov3[1].sum() / blob.sum() - 1.0) < 0.05 and np.abs(mov3[1] - blob).max() > 0.1
    qu = emission.rotate_evpa(np.stack([np.ones(3), np.zeros(3)]), np.pi / 4)
    assert np.allclose(qu, [[0, 0, 0], [1, 1, 1]], atol=1e-12)
    # a movie of grids: (T, nt, A, B) as in the reference
    mov = np.stack([d['emission_0'], 2.0 * d['emission_0']])
    out = emission.image_plane_dynamics(mov, _geo_ns(geo), geo['Omega'], tfM, float(d['t_injection']), t_start_obs=0.0,
                                        fov=fov).cpu().numpy()
    assert out.shape == (2,) + d['images_g'].shape and rel(out[1], 2.0 * d['images_g']) < IMG_TOL


def test_grid_predictor_forward_loss_gradient_and_step_vs_oracle():
    """network.GRID_Predictor through the reference-facing API: images, 'full' loss, gradient w.r.t. the grid
    (pull-back kernel, atomics) and one optax-Adam update, against the float64 oracle golden."""
    from bhnerf_b200 import engine, network
    from oracle import bhnerf_oracle as O
    d = np.load(os.path.join(G, 'grid_predictor.npz'))
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    pred = network.GRID_Predictor(float(d['scale']), float(d['rmin']), float(d['rmax']), float(d['z_width']),
                                  grid_res=d['grid'].shape[0])
    assert (pred.init_params()['grid'] == -10).all()
    rta = network.raytracing_args(_geo_ns(geo), geo['Omega'], float(d['t_injection']), float(d['t_start_obs']))
    params = {'grid': d['grid']}
    rel = lambda a, b: float(np.abs(np.asarray(a) - b).max() / np.abs(b).max())
    img = network.image_plane_prediction(params, pred.apply, d['t_frames'], *rta.values(), 'hr').cpu().numpy()
    assert rel(img, d['images']) < IMG_TOL, rel(img, d['images'])
    rtaJ = network.raytracing_args(_geo_ns(geo), geo['Omega'], float(d['t_injection']), float(d['t_start_obs']), J=d['J'])
    imgJ = network.image_plane_prediction(params, pred.apply, d['t_frames'], *rtaJ.values(), 'hr').cpu().numpy()
    assert imgJ.shape == d['images_J'].shape and rel(imgJ, d['images_J']) < IMG_TOL
    sig = np.ones_like(d['target']); off = np.zeros_like(d['target'])
    loss, _ = network.loss_fn_image(params, pred.apply, d['target'], sig, off, d['t_frames'], *rta.values(), 1.0, 'hr', 'full')
    assert abs(loss.item() - float(d['loss'])) / float(d['loss']) < IMG_TOL
    state = pred.init_state(params, num_iters=100, lr_init=1e-2, lr_final=1e-4)
    scene = network._scene_for(pred, *rta.values(), 'hr', device=state.flat.device)
    tf = torch.as_tensor(d['t_frames'].astype(np.float32)).cuda()
    images, _, _ = pred._render_fwd(scene, state.flat, tf)
    _, dI = engine.loss_image(images, d['target'].reshape(4, -1), sig.reshape(4, -1), off.reshape(4, -1), 1.0, 'full')
    g = pred._render_bwd(scene, state.flat, tf, dI).cpu().numpy().reshape(d['grad'].shape)
    assert rel(g, d['grad']) < GRAD_TOL, rel(g, d['grad'])
    assert ((g != 0) == (d['grad'] != 0)).mean() > 0.999
    l2, state, _ = network.gradient_step_image(state, 'hr', 'full', d['target'], sig, off, d['t_frames'], *rta.values(), 1.0)
    want, _, _ = O.adam_step(d['grid'].reshape(-1).astype(np.float64), d['grad'].reshape(-1), np.zeros(d['grid'].size),
                             np.zeros(d['grid'].size), 0, 1e-2, 1e-4, 100)
    got = state.flat.cpu().numpy().astype(np.float64)
    moved = d['grad'].reshape(-1) != 0
    assert np.abs(got - want)[moved].max() / 1e-2 < 2e-3 and state.step == 1
    assert abs(l2.item() - float(d['loss'])) / float(d['loss']) < IMG_TOL
    # GRID_Predictor.apply: dense emission on the geodesic points, zero outside the domain / before injection
    e = pred.apply({'params': params}, d['t_frames'], 'hr', geo['coords'], geo['Omega'], float(d['t_start_obs']),
                   geo['t_geos'], float(d['t_injection']))
    assert e.shape == (4,) + geo['coords'].shape[1:] and (e[0] == 0).all() and 0 < e.max() <= 1


def test_geodesic_inputs_vs_reference_algebra():
    """kgeo.geodesic_inputs (bhnerf_geodesic_inputs: get_dataset algebra + Keplerian Omega + Doppler factor, float64
    on the GPU) against the arrays the reference's own tracer + algebra produced (kerr_a0.2_i60_16x16x32.npz)."""
    from bhnerf_b200 import kgeo
    raw = np.load(os.path.join(G, 'kerr_raw_a0.2_i60_16x16x32.npz'))
    ref = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    geos = {k: raw[k] for k in ('r', 'theta', 'phi', 't', 'mino', 'lam')}
    geos.update(spin=float(raw['spin']), M=float(raw['M']))
    out = kgeo.geodesic_inputs(geos)
    for k, want in (('coords', ref['coords']), ('Omega', ref['Omega']), ('g', ref['g']), ('dtau', ref['dtau']),
                    ('Sigma', ref['Sigma']), ('t_geos', ref['t_geos'])):
        got = out[k].cpu().numpy()
        assert got.shape == want.shape, k
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6 * np.abs(want).max(), err_msg=k)
    assert (out['g'].cpu().numpy() == 0).sum() == (ref['g'] == 0).sum() > 0          # NaN -> 0 where no circular orbit
    # a user-supplied velocity field is taken as given (here: the Keplerian one, in float64)
    Om = np.sign(0.2) / (raw['r'] ** 1.5 + 0.2)
    out2 = kgeo.geodesic_inputs(geos, Omega=Om)
    assert torch.equal(out2['g'], out['g']) and torch.equal(out2['Omega'], out['Omega'])
    # the result feeds the prepack directly
    from bhnerf_b200 import engine, constants
    scene = engine.PackedScene(out['coords'], out['Omega'], 1.0, out['g'], out['dtau'], out['Sigma'], out['t_geos'], 0.0,
                               -1000.0, 8.0, 2.5, 8.0, 4.0, constants.GM_c3(t_units='hr'))
    scene_ref = engine.PackedScene(ref['coords'], ref['Omega'], 1.0, ref['g'], ref['dtau'], ref['Sigma'], ref['t_geos'], 0.0,
                                   -1000.0, 8.0, 2.5, 8.0, 4.0, constants.GM_c3(t_units='hr'))
    assert scene.n_active == scene_ref.n_active > 0


def test_cuda_graph_replay_of_the_train_step_matches_direct_calls():
    """gradient_step_image replays a captured [train step -> Adam (device-side counter)] graph; the parameters after
    several iterations with changing batches must match the direct-call path, and the step counter / learning-rate
    schedule must advance (also across a counter change made behind the graph's back, as restore_checkpoint does)."""
    from collections import OrderedDict
    from bhnerf_b200 import network, optimization, synthetic
    c = synthetic.make_config('tiny')
    rt, pr = c['rt'], c['predictor']
    rta = OrderedDict((k, rt[k]) for k in ('coords', 'Omega', 'J', 'g', 'dtau', 'Sigma', 't_start_obs', 't_geos', 't_injection'))
    pred = network.NeRF_Predictor(pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'])
    nt, A, B = len(c['t_frames']), c['A'], c['B']
    ts = optimization.TrainStep.image(c['t_frames'], c['target'].reshape(nt, A, B), sigma=c['sigma'].reshape(nt, A, B))
    batches = [np.array([0, 1]), np.array([2, 3]), np.array([1, 3]), np.array([0, 2]), np.array([3, 1])]
    res = {}
    for use in (True, False):
        network._USE_GRAPHS = use
        try:
            state = pred.init_state(network.unflatten_params(synthetic.trained_like_flat_params(7)), num_iters=50,
                                    lr_init=1e-3, lr_final=1e-5)
            losses = []
            for i, b in enumerate(batches):
                if i == 3:
                    state.step = 20                     # e.g. a restored checkpoint: the schedule must follow
                loss, state, images = ts(state, rta, b)
                losses.append(float(loss.item()))
            res[use] = (state.flat.clone(), losses, int(state.step), images.clone())
        finally:
            network._USE_GRAPHS = True
    (p1, l1, s1, i1), (p0, l0, s0, i0) = res[True], res[False]
    assert s1 == s0 == 22
    np.testing.assert_allclose(l1, l0, rtol=1e-5)
    assert (i1 - i0).abs().max() / i0.abs().max() < 1e-5
    moved = (p0 - torch.as_tensor(synthetic.trained_like_flat_params(7)).cuda()).abs().max().item()
    assert moved > 1e-3 and (p1 - p0).abs().max().item() < 2e-3 * moved


def test_optimizer_run_fits_a_synthetic_movie_and_resumes_from_a_flax_checkpoint(tmp_path):
    """The reference's whole loop through its own API (optimization.Optimizer.run -> TrainStep -> gradient_step_image):
    a NeRF is fitted to the movie another (trained-like) NeRF renders; the movie loss must drop, checkpoints are written
    in flax's msgpack format with the reference's keep policy, and a new Optimizer resumes from the latest one."""
    from collections import OrderedDict
    from bhnerf_b200 import network, optimization, synthetic
    c = synthetic.make_config('tiny')
    rt, pr = c['rt'], c['predictor']
    rta = OrderedDict((k, rt[k]) for k in ('coords', 'Omega', 'J', 'g', 'dtau', 'Sigma', 't_start_obs', 't_geos', 't_injection'))
    pred = network.NeRF_Predictor(pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'])
    truth = network.unflatten_params(synthetic.trained_like_flat_params(7))
    movie = network.image_plane_prediction(truth, pred.apply, c['t_frames'], *rta.values(), 'hr').cpu().numpy()
    assert movie.shape == (4, c['A'], c['B']) and movie.max() > 0
    ts = optimization.TrainStep.image(c['t_frames'], movie, sigma=1.0, dtype='full')
    ck = str(tmp_path / 'run')
    np.random.seed(1)
    opt = optimization.Optimizer({'num_iters': 60, 'lr_init': 2e-3, 'lr_final': 1e-4, 'seed': 1}, pred, rta, save_period=20,
                                 checkpoint_dir=ck, keep=2)
    loss0 = optimization.total_movie_loss(2, opt.state, ts, rta)
    seen = []
    opt.run(2, ts, rta, log_fns=[optimization.LogFn(lambda o: seen.append(float(o.loss.item())), log_period=10)])
    loss1 = optimization.total_movie_loss(2, opt.state, ts, rta)
    assert opt.state.step == 60 and len(seen) == 7 and np.isfinite(seen).all()
    assert loss1 < 0.5 * loss0, (loss0, loss1)
    files = sorted(os.listdir(ck))
    assert 'NeRF_Predictor_params.yml' in files and [f for f in files if f.startswith('checkpoint_')] == ['checkpoint_40', 'checkpoint_60']
    import msgpack
    tree = msgpack.unpackb(open(os.path.join(ck, 'checkpoint_60'), 'rb').read(), ext_hook=optimization._flax_ext_hook, raw=False)
    assert tree['step'] == 60 and tree['params']['MLP_0']['Dense_0']['kernel'].shape == (21, 128)
    opt2 = optimization.Optimizer({'num_iters': 10, 'lr_init': 2e-3, 'lr_final': 1e-4, 'seed': 5}, pred, rta, checkpoint_dir=ck)
    assert opt2.state.step == 60 and torch.equal(opt2.state.flat, opt.state.flat) and torch.equal(opt2.state.nu, opt.state.nu)
    opt2.run(2, ts, rta)
    assert opt2.state.step == 70 and optimization.total_movie_loss(2, opt2.state, ts, rta) < loss0
    restored = network.NeRF_Predictor.from_yml(ck)
    assert restored.domain() == pred.domain()
    # forward-only model selection on the finished run (network.image_plane_checkpoint, alma.chi2_lightcurves / chi2_df)
    from bhnerf_b200 import alma
    frames = network.image_plane_checkpoint(rta, ck, c['t_frames'], batchsize=2)
    direct = network.image_plane_prediction(opt2.state.params, pred.apply, c['t_frames'], *rta.values(), 'hr').cpu().numpy()
    assert frames.shape == direct.shape and np.abs(frames - direct).max() <= 1e-6 * np.abs(direct).max()
    lc = movie.sum(axis=(-1, -2))                                # unpolarized: data (nt,), as image_plane.sum((-1,-2))
    chi2 = alma.chi2_lightcurves(rta, ck, c['t_frames'], lc, sigma=0.5)
    want = float(np.sum(((direct.sum(axis=(-1, -2)) - lc) / 0.5) ** 2) / 4)
    assert abs(chi2 - want) <= 1e-4 * max(want, 1e-12)
    df = alma.chi2_df([30.0, 60.0], 0.2, [0, 1], {}, str(tmp_path) + '/{}_{}', c['t_frames'], lc, sigma=0.5,
                      raytracing_args_fn=lambda inc, spin: rta, final_step=70)
    assert df.shape == (2, 2) and df.isna().all().all()          # no run directories of that pattern: all NaN, as the reference
    os.rename(ck, str(tmp_path) + '/60.0_1')
    df = alma.chi2_df([30.0, 60.0], 0.2, [0, 1], {}, str(tmp_path) + '/{}_{}', c['t_frames'], lc, sigma=0.5,
                      raytracing_args_fn=lambda inc, spin: rta, final_step=70)
    assert abs(df.loc[60.0, 'seed 1'] - want) <= 1e-4 * max(want, 1e-12) and df.isna().sum().sum() == 3


def test_reference_call_patterns_of_the_standalone_stages():
    """Argument shapes the reference's callers use: scalar frame time (no frame axis in the result), scalar Omega /
    t_geos (sample_3d_grid passes 0), coords as (3, N), emission with leading (frame, Stokes) axes, single-frame polarized
    prediction (jnp.squeeze semantics of network.py:418)."""
    from bhnerf_b200 import emission, kgeo, network
    from oracle import bhnerf_oracle as O
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    d = np.load(os.path.join(G, 'ref_stages.npz'))
    c = float(d['GM_c3'])
    co = geo['coords']
    t0, tinj = float(d['t_start_obs']), float(d['t_injection'])
    # scalar frame time, hours with t_units -> (16,16,32,3); compare with the batched call's slice
    w1 = emission.velocity_warp_coords(co, geo['Omega'], 0.5, t0, geo['t_geos'], tinj, t_units='hr').cpu().numpy()
    wb = emission.velocity_warp_coords(co, geo['Omega'], np.array([0.13, 0.5]), t0, geo['t_geos'], tinj, t_units='hr').cpu().numpy()
    assert w1.shape == co.shape[1:] + (3,) and wb.shape == (2,) + co.shape[1:] + (3,)
    assert np.array_equal(np.nan_to_num(w1), np.nan_to_num(wb[1]))
    ref = O.velocity_warp_coords(co, geo['Omega'], np.array([0.5]), t0, geo['t_geos'], tinj, c, torch.float64).numpy()[0]
    ok = np.isfinite(ref)
    assert np.array_equal(np.isfinite(w1), ok) and np.abs(w1[ok] - ref[ok]).max() < 2e-3      # fp32 theta at t_M ~ 1e3
    # scalar Omega and t_geos on a flat (3, N) point list
    pts = co.reshape(3, -1)[:, :500]
    w2 = emission.velocity_warp_coords(pts, 0.05, np.array([0.0, 0.2]), 0.0, 0.0, 0.0, t_units='hr').cpu().numpy()
    ref2 = O.velocity_warp_coords(pts, np.float64(0.05), np.array([0.0, 0.2]), 0.0, np.float64(0.0), 0.0, c, torch.float64).numpy()
    assert w2.shape == (2, 500, 3) and np.abs(w2 - ref2).max() < 1e-4 * np.abs(ref2).max()
    # radiative transfer with leading (frame, Stokes) axes
    rng = np.random.default_rng(2)
    e = rng.uniform(size=(3, 2) + co.shape[1:]).astype(np.float32)
    rtv = kgeo.radiative_trasfer(e, geo['g'], geo['dtau'], geo['Sigma']).cpu().numpy()
    want = (geo['g'].astype(np.float64) ** 2 * e * geo['dtau'] * geo['Sigma']).sum(-1)
    assert rtv.shape == (3, 2, 16, 16) and np.abs(rtv - want).max() < 1e-5 * np.abs(want).max()
    # fill with the default arguments (only z_width = 2 bites) and an explicit fill value
    f = emission.fill_unsupervised_emission(e, co, fill_value=-1.0).cpu().numpy()
    outside = np.abs(co[2]) > 2.0
    assert (f[..., outside] == -1).all() and np.array_equal(f[..., ~outside], e[..., ~outside])
    # single polarized frame: (S, A, B) after the squeeze
    pred = network.NeRF_Predictor(8.0, 2.5, 8.0, 4.0)
    params = pred.init_params(seed=2)
    J = np.stack([np.ones_like(geo['g']), 0.5 * np.ones_like(geo['g'])])
    rta = network.raytracing_args(_geo_ns(geo), geo['Omega'], tinj, t0, J=J)
    img = network.image_plane_prediction(params, pred.apply, np.array([0.7]), *rta.values(), 'hr')
    assert tuple(img.shape) == (2, 16, 16) and torch.allclose(img[1], 0.5 * img[0], rtol=1e-5, atol=0)
    e3 = pred.apply({'params': params}, 0.7, 'hr', pts, 0.0, t0, 0.0, 0.0)
    assert e3.shape == (500,) and np.isfinite(e3).all()


def test_multi_loss_train_step_image_plus_visibilities():
    """TrainStep.__add__ (optimization.py:181-187): an image loss and a visibility loss in one iteration = two sequential
    gradient steps on the same state, the second one at the parameters the first one produced.  Both updates against the
    oracle (value_and_grad + optax Adam restated), the second gradient evaluated live at the oracle's updated parameters."""
    from collections import OrderedDict
    from bhnerf_b200 import network, optimization
    from oracle import bhnerf_oracle as O
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    d1 = np.load(os.path.join(G, 'case_image_full.npz'))
    d2 = np.load(os.path.join(G, 'case_vis.npz'))
    assert np.array_equal(d1['params_flat'], d2['params_flat']) and np.array_equal(d1['t_frames'], d2['t_frames'])
    prd = dict(scale=float(d1['scale']), rmin=float(d1['rmin']), rmax=float(d1['rmax']), z_width=float(d1['z_width']))
    pred = network.NeRF_Predictor(prd['scale'], prd['rmin'], prd['rmax'], prd['z_width'])
    rt = OrderedDict(coords=geo['coords'], Omega=geo['Omega'], J=1.0, g=geo['g'], dtau=geo['dtau'], Sigma=geo['Sigma'],
                     t_start_obs=float(d1['t_start_obs']), t_geos=geo['t_geos'], t_injection=float(d1['t_injection']))
    ts = optimization.TrainStep.image(d1['t_frames'], d1['target'], sigma=1.0, dtype='full') + optimization.TrainStep.eht_arrays(
        d2['t_frames'], d2['target'], d2['sigma'], d2['A'], dtype='vis', scale=1.0)
    assert ts.num_losses == 2
    state = pred.init_state(network.unflatten_params(d1['params_flat']), num_iters=100, lr_init=1e-3, lr_final=1e-5)
    loss, state, images = ts(state, rt, np.arange(4))
    assert state.step == 2
    # oracle: step 1 at params0 (golden gradient), step 2 at the updated parameters
    p0 = d1['params_flat'].astype(np.float64)
    z = np.zeros(55169)
    p1, mu, nu = O.adam_step(p0, d1['grads'], z, z, 0, 1e-3, 1e-5, 100)
    out2 = O.value_and_grad(O.unflatten_params(p1.astype(np.float32)), 'eht', 'vis', d2['target'], d2['sigma'], d2['A'],
                            d2['t_frames'], dict(rt), prd)
    p2, _, _ = O.adam_step(p1.astype(np.float32).astype(np.float64), out2['grads'], mu, nu, 1, 1e-3, 1e-5, 100)
    assert abs(loss.item() - (float(d1['loss']) + out2['loss'])) / (float(d1['loss']) + out2['loss']) < 2 * IMG_TOL
    got = state.flat.cpu().numpy().astype(np.float64)
    upd_err = np.abs((got - p0) - (p2 - p0)).max() / 2e-3            # two steps of ~lr each
    assert upd_err < 3e-2, upd_err


def test_two_rank_nccl_step_matches_oracle():
    """Frame-sharded step on 2 GPUs (bhnerf_allreduce_mean below the C ABI, inside the captured CUDA graph, + Adam) and the
    ray-sharded steps ('full', 'lc', 'vis': 3 frames on 2 ranks) against the oracle; needs a 2-GPU box (skipped otherwise;
    last run recorded in profiles/r2_multirank_check.log)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', '29547', os.path.join(root, 'scripts', 'multirank_check.py')],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('ranks identical: True') == 2
    assert r.stdout.count('ray-sharded') == 3 and 'FAILED' not in r.stdout


def test_jax_scene_prepack_attributes_and_sizes():
    """bhnerf_b200.jax_scene (the repository side of the jax.ffi binding): prepack through ctypes, the typed scalar
    attributes equal the fields of the C scene struct, and a render driven ONLY by (packed buffer, attrs, size queries) --
    what an XLA-FFI handler receives -- reproduces the golden images."""
    import ctypes as C
    from bhnerf_b200 import _lib, engine, jax_scene, network
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    d = np.load(os.path.join(G, 'case_lc_IQU.npz'))
    pred = network.NeRF_Predictor(float(d['scale']), float(d['rmin']), float(d['rmax']), float(d['z_width']))
    cache = {}
    args = (geo['coords'], geo['Omega'], d['J'], geo['g'], geo['dtau'], geo['Sigma'], float(d['t_start_obs']), geo['t_geos'],
            float(d['t_injection']), 'hr')
    sc = jax_scene.prepack(pred, *args, scene_cache=cache)
    assert jax_scene.prepack(pred, *args, scene_cache=cache) is sc
    a = sc.attrs
    assert [k for k in a] == [n for n, _ in jax_scene.ATTRS] and a['P'] == 256 and a['G'] == 32 and a['S'] == 3
    assert a['n_pad'] % 128 == 0 and 0 < a['n_active'] <= a['n_pad'] and a['scale'].dtype == np.float32
    Bt = 4
    assert sc.acts_bytes(Bt) > 0 and sc.fwd_workspace_bytes > 0 and sc.bwd_workspace_bytes(Bt) > sc.bwd_workspace_bytes(1)
    # what the handler does: rebuild the struct from (buffer pointer, attrs) and call the C ABI
    st = _lib.Scene(packed=sc.packed.data_ptr(), **{k: (int(v) if np.issubdtype(type(v), np.integer) else float(v)) for k, v in a.items()})
    lib = _lib.load()
    dev = sc.packed.device
    params = torch.as_tensor(d['params_flat'], device=dev); tf = torch.as_tensor(d['t_frames'].astype(np.float32), device=dev)
    images = torch.empty((Bt, 3, 256), device=dev); e = torch.empty((Bt, int(a['n_pad'])), device=dev)
    ws = torch.empty(sc.fwd_workspace_bytes, dtype=torch.uint8, device=dev)
    rc = lib.bhnerf_render_fwd(C.byref(st), params.data_ptr(), tf.data_ptr(), Bt, images.data_ptr(), e.data_ptr(), None,
                               ws.data_ptr(), ws.numel(), _lib.IMPL_TC, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.bhnerf_last_error()
    torch.cuda.synchronize()
    assert np.abs(images.cpu().numpy().reshape(d['images'].shape) - d['images']).max() / np.abs(d['images']).max() < IMG_TOL


def test_polarization_factors_vs_reference_functions():
    """bhnerf_polarization_factors (kgeo.polarization_factors) against the output of the reference's own
    azimuthal_velocity_vector / doppler_factor / magnetic_field_fluid_frame / parallel_transport (tests/golden/pol_factors.npz):
    float64 arithmetic, float32 result; compared inside the recovery domain (outside it the factors diverge towards the
    horizon and the prepack culls the samples) and, with a looser bound, wherever the reference is finite."""
    from bhnerf_b200 import kgeo
    d = np.load(os.path.join(G, 'pol_factors.npz'))
    for tag in ('a', 'b'):
        a, inc, rmin, rmax, zw = d[tag + '_consts']
        geos = dict(r=d[tag + '_r'], theta=d[tag + '_theta'], affine=d[tag + '_affine'], lam=d[tag + '_lam'], eta=d[tag + '_eta'],
                    alpha=d[tag + '_alpha'], beta=d[tag + '_beta'], spin=float(a), inc=float(inc))
        dom = (np.abs(geos['r'] * np.cos(geos['theta'])) < zw) & (geos['r'] > rmin) & (geos['r'] < rmax)
        for j in (0, 1):
            b = d['%s_b%d' % (tag, j)]
            J = kgeo.polarization_factors(geos, None, dict(arad=b[0], avert=b[1], ator=b[2]), 0.5, rmin, rmax, zw).cpu().numpy()
            ref = d['%s_J%d' % (tag, j)]
            assert J.shape == ref.shape and J.dtype == np.float32 and np.isfinite(J).all()
            err = np.abs(J[:, dom] - ref[:, dom]).max() / np.abs(ref[:, dom]).max()
            sane = np.abs(ref) < 1e4
            err_all = np.abs(J - ref)[sane].max() / np.abs(ref[sane]).max()
            print(tag, j, 'in-domain err %.2e  everywhere %.2e' % (err, err_all))
            assert err < 1e-6 and err_all < 1e-5
    with pytest.raises(Exception, match='Q_frac'):
        kgeo.polarization_factors(geos, None, None, 1.5, rmin, rmax, zw)


@pytest.mark.parametrize('NA,NB,V,Bt', [(16, 16, 20, 3), (128, 128, 190, 4), (96, 80, 130, 2)])
def test_separable_tensor_core_dft_head_matches_the_explicit_matrix(NA, NB, V, Bt):
    """bhnerf_vis_dft_fwd / _bwd (tcgen05 GEMM with DFT factors generated on chip from (u, v)) against the explicit-matrix
    head (bhnerf_vis_fwd / _bwd) fed with A[b,k,(i,j)] = pulse * exp(-2 pi i (u x_i + v y_j)) -- the matrix ehtim builds for
    loss_fn_eht (bhnerf/network.py:542-544) -- and against float64 numpy.  Visibilities <= 1e-4, pull-back <= 1e-3."""
    from bhnerf_b200 import engine
    rng = np.random.default_rng(NA + V)
    psize = 16.0 / NA
    x = (np.arange(NA) - NA / 2) * psize; y = (np.arange(NB) - NB / 2) * psize
    uv = rng.uniform(-0.25, 0.25, size=(Bt, V, 2))
    pulse = (rng.normal(1, 0.1, (Bt, V)) + 1j * rng.normal(0, 0.1, (Bt, V)))
    ph = -2 * np.pi * (uv[..., 0][:, :, None, None] * x[None, None, :, None] + uv[..., 1][:, :, None, None] * y[None, None, None, :])
    A64 = (pulse[:, :, None, None] * np.exp(1j * ph)).reshape(Bt, V, NA * NB)
    img = rng.uniform(0, 1, size=(Bt, NA, NB)) * np.exp(rng.normal(0, 2, size=(Bt, NA, NB)))      # several decades
    dv = (rng.normal(0, 1, (Bt, V)) + 1j * rng.normal(0, 1, (Bt, V))) * 3e4                          # sigma = 0.01-sized cotangents
    vis64 = np.einsum('bkp,bp->bk', A64, img.reshape(Bt, -1))
    dI64 = np.einsum('bkp,bk->bp', np.conj(A64), dv).real.reshape(Bt, NA, NB)
    grid = (x[0], psize, y[0], psize)
    img_d = torch.as_tensor(img.astype(np.float32)).cuda()
    vis = engine.vis_dft_fwd(uv.astype(np.float32), img_d, grid, pulse=pulse.astype(np.complex64))
    dI = engine.vis_dft_bwd(uv.astype(np.float32), torch.as_tensor(dv.astype(np.complex64)).cuda(), grid, NA, NB,
                            pulse=pulse.astype(np.complex64))
    A_d = torch.as_tensor(A64.astype(np.complex64)).cuda()
    vis_x = engine.vis_fwd(A_d, img_d.reshape(Bt, 1, -1))
    dI_x = engine.vis_bwd(A_d, torch.as_tensor(dv.astype(np.complex64)).cuda(), NA * NB)
    torch.cuda.synchronize()
    ev = np.abs(vis.cpu().numpy() - vis64).max() / np.abs(vis64).max()
    ed = np.abs(dI.cpu().numpy() - dI64).max() / np.abs(dI64).max()
    ev_x = np.abs(vis_x.cpu().numpy() - vis64).max() / np.abs(vis64).max()
    ed_x = np.abs(dI_x.cpu().numpy().reshape(Bt, NA, NB) - dI64).max() / np.abs(dI64).max()
    print('separable: vis %.2e dI %.2e   explicit A: vis %.2e dI %.2e' % (ev, ed, ev_x, ed_x))
    assert ev < IMG_TOL and ev_x < IMG_TOL
    assert ed < GRAD_TOL and ed_x < GRAD_TOL


@pytest.mark.parametrize('dtype', ['vis', 'amp', 'cphase'])
def test_eht_step_with_separable_dft_equals_the_explicit_matrix_step(dtype):
    """network.SeparableDFT (baselines instead of the matrix) through TrainStep.eht_arrays -> gradient_step_eht: same loss,
    images and parameter update as the explicit-A step on the matrix built from the same baselines."""
    from collections import OrderedDict
    from bhnerf_b200 import network, optimization
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    d = np.load(os.path.join(G, 'case_vis.npz'))
    rng = np.random.default_rng(5)
    nt, NA, NB, V = 4, 16, 16, 15
    fov = float(geo['fov_M'])
    x = (np.arange(NA) - NA / 2) * fov / NA
    lead = (nt, 3, V) if dtype == 'cphase' else (nt, V)
    uv = rng.uniform(-0.25, 0.25, size=lead + (2,))
    if dtype == 'cphase':
        uv[:, 2] = -(uv[:, 0] + uv[:, 1])
    A = np.exp(-2j * np.pi * (uv[..., 0][..., None, None] * x[:, None] + uv[..., 1][..., None, None] * x[None, :])).reshape(lead + (NA * NB,))
    tshape = (nt, V)
    if dtype == 'vis':
        target = (rng.normal(0, 1, tshape) + 1j * rng.normal(0, 1, tshape)).astype(np.complex64)
    elif dtype == 'amp':
        target = np.abs(rng.normal(0.5, 0.3, tshape)).astype(np.float32)
    else:
        target = rng.uniform(-np.pi, np.pi, tshape).astype(np.float32)
    sigma = np.full(tshape, 0.3, dtype=np.float32)
    pred = network.NeRF_Predictor(float(d['scale']), float(d['rmin']), float(d['rmax']), float(d['z_width']))
    rt = OrderedDict(coords=geo['coords'], Omega=geo['Omega'], J=1.0, g=geo['g'], dtau=geo['dtau'], Sigma=geo['Sigma'],
                     t_start_obs=float(d['t_start_obs']), t_geos=geo['t_geos'], t_injection=float(d['t_injection']))
    out = []
    for Aarg in (A.astype(np.complex64), network.SeparableDFT.from_fov(uv, (NA, NB), fov)):
        state = pred.init_state(network.unflatten_params(d['params_flat']), num_iters=100, lr_init=1e-3, lr_final=1e-5)
        ts = optimization.TrainStep.eht_arrays(d['t_frames'], target, sigma, Aarg, dtype=dtype)
        loss, state, images = ts(state, rt, np.arange(nt))
        out.append((loss.item(), images.cpu().numpy(), state.flat.cpu().numpy() - d['params_flat']))
    assert abs(out[0][0] - out[1][0]) / abs(out[0][0]) < IMG_TOL
    assert np.abs(out[0][1] - out[1][1]).max() / np.abs(out[0][1]).max() < IMG_TOL
    assert np.abs(out[0][2] - out[1][2]).max() / 1e-3 < 2e-2            # Adam update ~ lr per parameter


def test_empty_and_tiny_recovery_domain():
    """Edge cases of the prepack (emission.fill_unsupervised_emission zeroes everything outside the domain, emission.py:370-373):
    a recovery domain that contains no sample gives zero images, zero loss against a zero target and a zero gradient; one that
    keeps 156 samples (two tiles, mostly padding) agrees between the two kernel families."""
    from bhnerf_b200 import constants, engine
    geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
    d = np.load(os.path.join(G, 'case_image_full.npz'))
    params = torch.as_tensor(d['params_flat']).cuda()
    tf = torch.as_tensor(d['t_frames'].astype(np.float32)).cuda()
    zero = np.zeros_like(d['target']); one = np.ones_like(zero)

    def scene(rmin, rmax, zw):
        return engine.PackedScene(geo['coords'], geo['Omega'], 1.0, geo['g'], geo['dtau'], geo['Sigma'], geo['t_geos'],
                                  float(d['t_start_obs']), float(d['t_injection']), float(d['scale']), rmin, rmax, zw,
                                  constants.GM_c3(t_units='hr'))
    sc = scene(50.0, 60.0, 1e-3)
    assert sc.n_active == 0
    for impl in ('simt', 'tc'):
        loss, img, g = engine.train_step_image(sc, params, tf, zero, one, zero, 1.0, 'full', impl)
        assert float(loss) == 0.0 and float(img.abs().max()) == 0.0 and float(g.abs().max()) == 0.0
    sc = scene(float(d['rmin']), float(d['rmin']) + 0.3, 4.0)
    assert 0 < sc.n_active < 256
    res = {impl: [t.clone() for t in engine.train_step_image(sc, params, tf, zero, one, zero, 1.0, 'full', impl)]
           for impl in ('simt', 'tc')}
    assert engine.workspace_status(impl='tc')[:5] == [0, 0, 0, 0, 0]
    for a, b, tol in zip(res['tc'], res['simt'], (1e-4, 1e-4, 1e-3)):
        assert float((a - b).abs().max()) <= tol * float(b.abs().max())
