"""Float64-oracle golden of POLARIZED visibilities (bhnerf/network.py:542-548 with the pol axis that
optimization.py:235-251 stacks on axis 1: A (nt, npol, nvis, npix), target / sigma (nt, npol, nvis)).  Needs only the
committed Kerr-geodesic fixture and the oracle:   python tests/golden/make_golden_eht_pol.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..')))
from oracle import bhnerf_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    geo = np.load(os.path.join(HERE, 'kerr_a0.2_i60_16x16x32.npz'))
    base = np.load(os.path.join(HERE, 'case_lc_IQU.npz'))          # same scene constants and Stokes factors J (I,Q,U)
    A_ = B_ = 16
    flat = base['params_flat']
    params = O.unflatten_params(flat)
    t_frames = base['t_frames']
    pred = dict(scale=float(base['scale']), rmin=float(base['rmin']), rmax=float(base['rmax']), z_width=float(base['z_width']))
    rt = dict(coords=geo['coords'], Omega=geo['Omega'], g=geo['g'], dtau=geo['dtau'], Sigma=geo['Sigma'],
              t_geos=geo['t_geos'], t_start_obs=float(base['t_start_obs']), t_injection=float(base['t_injection']),
              J=base['J'])
    rng = np.random.default_rng(33)
    fov = float(geo['fov_M']); psize = fov / A_
    xx, yy = np.meshgrid((np.arange(A_) - A_ / 2) * psize, (np.arange(B_) - B_ / 2) * psize, indexing='ij')
    nt, S, V = len(t_frames), 3, 12
    uv = rng.uniform(-0.3, 0.3, size=(nt, S, V, 2))
    Amat = np.exp(-2j * np.pi * (uv[..., 0:1] * xx.reshape(1, 1, 1, -1) + uv[..., 1:2] * yy.reshape(1, 1, 1, -1))
                  ).astype(np.complex64)                               # (nt, npol, V, P): one DFT matrix per frame and pol
    tgt = (rng.normal(0, 1, size=(nt, S, V)) + 1j * rng.normal(0, 1, size=(nt, S, V))).astype(np.complex64)
    sig = np.broadcast_to(np.array([0.5, 0.1, 0.1], dtype=np.float32)[None, :, None], (nt, S, V)).copy()
    out = O.value_and_grad(params, 'eht', 'vis', tgt, sig, Amat, t_frames, rt, pred, scale=1.0)
    assert out['vis'].shape == (nt, S, V) and out['images'].shape == (nt, S, A_, B_)
    np.savez_compressed(os.path.join(HERE, 'case_vis_IQU.npz'), params_flat=flat, t_frames=t_frames, GM_c3=float(base['GM_c3']),
                        **{k: np.asarray(v) for k, v in pred.items()}, loss=out['loss'], images=out['images'],
                        grads=out['grads'], vis=out['vis'], t_start_obs=rt['t_start_obs'], t_injection=rt['t_injection'],
                        target=tgt, sigma=sig, A=Amat, J=base['J'])
    print('written case_vis_IQU.npz  loss', out['loss'])


if __name__ == '__main__':
    main()
