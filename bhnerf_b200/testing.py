"""Shared helpers for tests/, smoke() and bench.py: golden-case runner and error metrics.
The comparisons are against committed oracle outputs (tests/golden/*.npz)."""
import os

import numpy as np
import torch

from . import engine

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def rel_err(a, b):
    """max-norm relative error  max|a-b| / max|b|  (the metric the parity tolerances are stated in)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def rel_err_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def load_golden_scene(case, device='cuda'):
    geo = np.load(os.path.join(GOLDEN, 'kerr_a0.2_i60_16x16x32.npz'))
    d = np.load(os.path.join(GOLDEN, case + '.npz'))
    J = d['J'] if 'J' in d.files else 1.0
    scene = engine.PackedScene(geo['coords'], geo['Omega'], J, geo['g'], geo['dtau'], geo['Sigma'], geo['t_geos'],
                               float(d['t_start_obs']), float(d['t_injection']), float(d['scale']), float(d['rmin']),
                               float(d['rmax']), float(d['z_width']), float(d['GM_c3']), device=device)
    return scene, d


def run_golden_case(case, impl=None, max_workspace=None):
    """Runs one committed golden case through the C ABI and returns error metrics vs the oracle."""
    impl_i = engine.resolve_impl(impl)
    scene, d = load_golden_scene(case)
    dev = scene.device
    params = torch.as_tensor(d['params_flat'], device=dev)
    t_frames = torch.as_tensor(d['t_frames'].astype(np.float32), device=dev)
    A_, B_ = scene.image_shape
    if case in ('case_vis', 'case_amp', 'case_cphase'):
        kind = case[5:]
        A = torch.as_tensor(d['A'], device=dev)
        if kind == 'cphase':                                   # (nt,3,V,P): triangle axis folded into the row axis
            A = A.reshape(A.shape[0], 3 * A.shape[2], A.shape[3])
        images, e, acts = engine.render_fwd(scene, params, t_frames, impl_i, save_acts=True)
        vis = engine.vis_fwd(A, images)
        loss, dvis = engine.loss_vis(vis, d['target'], d['sigma'], 1.0, kind)
        dI = engine.vis_bwd(A, dvis, scene.P)
        grads = engine.render_bwd(scene, params, t_frames, dI, e, acts, impl_i)
        v = vis.cpu().numpy().reshape(d['vis'].shape)
        extra = {'vis_err': float(np.abs(v - d['vis']).max() / np.abs(d['vis']).max())}
    else:
        kind = 'full' if case == 'case_image_full' else 'lc'
        tgt = d['target']
        sig = d['sigma'] if 'sigma' in d.files else np.ones_like(tgt)
        off = np.zeros_like(tgt)
        loss, images, grads = engine.train_step_image(scene, params, t_frames, tgt, sig, off, 1.0, kind, impl_i,
                                                      max_workspace=max_workspace)
        extra = {}
    torch.cuda.synchronize()
    img = images.cpu().numpy().reshape(d['images'].shape if d['images'].ndim == 4 else
                                       (d['images'].shape[0], 1) + d['images'].shape[1:])
    ref_img = d['images'] if d['images'].ndim == 4 else d['images'][:, None]
    # our ray index p = alpha-major flattening of the (A,B) image axes = the golden's axes 1,2
    out = {'impl': 'tc' if impl_i == engine.IMPL_TC else 'simt',
           'img_err': rel_err(img, ref_img), 'img_err_l2': rel_err_l2(img, ref_img),
           'grad_err': rel_err(grads.cpu().numpy(), d['grads']),
           'grad_err_l2': rel_err_l2(grads.cpu().numpy(), d['grads']),
           'loss_err': abs(float(loss.item()) - float(d['loss'])) / abs(float(d['loss'])),
           'n_active': scene.n_active}
    out.update(extra)
    g = grads.cpu().numpy().astype(np.float64); gr = d['grads']; o = 0; per = {}
    for i, (fi, fo) in enumerate([(21, 128), (128, 128), (128, 128), (149, 128), (128, 1)]):
        for nm, sz in (('W', fi * fo), ('b', fo)):
            per['%s%d' % (nm, i)] = float(np.abs(g[o:o + sz] - gr[o:o + sz]).max() / np.abs(gr).max()); o += sz
    out['per_layer'] = per
    return out
