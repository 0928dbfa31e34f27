"""Extra float64-oracle goldens for the EHT heads ('amp', 'cphase'; bhnerf/network.py:550-559).  Needs only the
committed Kerr-geodesic fixture and the oracle (no /root/reference):   python tests/golden/make_golden_eht_extra.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..')))
from oracle import bhnerf_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    geo = np.load(os.path.join(HERE, 'kerr_a0.2_i60_16x16x32.npz'))
    base_case = np.load(os.path.join(HERE, 'case_vis.npz'))
    A_ = B_ = 16
    flat = base_case['params_flat']
    params = O.unflatten_params(flat)
    t_frames = base_case['t_frames']
    pred = dict(scale=float(base_case['scale']), rmin=float(base_case['rmin']), rmax=float(base_case['rmax']),
                z_width=float(base_case['z_width']))
    rt = dict(coords=geo['coords'], Omega=geo['Omega'], g=geo['g'], dtau=geo['dtau'], Sigma=geo['Sigma'],
              t_geos=geo['t_geos'], t_start_obs=float(base_case['t_start_obs']),
              t_injection=float(base_case['t_injection']), J=1.0)
    rng = np.random.default_rng(21)
    fov = float(geo['fov_M']); psize = fov / A_
    xx, yy = np.meshgrid((np.arange(A_) - A_ / 2) * psize, (np.arange(B_) - B_ / 2) * psize, indexing='ij')

    def dft(uv):
        return np.exp(-2j * np.pi * (uv[..., 0:1] * xx.reshape(1, 1, -1) + uv[..., 1:2] * yy.reshape(1, 1, -1))
                      ).astype(np.complex64)

    def save(name, out, extra):
        np.savez_compressed(os.path.join(HERE, name), params_flat=flat, t_frames=t_frames, GM_c3=float(base_case['GM_c3']),
                            **{k: np.asarray(v) for k, v in pred.items()}, loss=out['loss'], images=out['images'],
                            grads=out['grads'], vis=out['vis'], t_start_obs=rt['t_start_obs'],
                            t_injection=rt['t_injection'], **extra)

    # 'amp': visibility amplitudes
    V = 24
    Aamp = dft(rng.uniform(-0.3, 0.3, size=(4, V, 2)))
    tgt = np.abs(rng.normal(0.5, 0.3, size=(4, V))).astype(np.float32)
    sig = np.full((4, V), 0.2, dtype=np.float32)
    out = O.value_and_grad(params, 'eht', 'amp', tgt, sig, Aamp, t_frames, rt, pred, scale=1.0)
    save('case_amp.npz', out, dict(target=tgt, sigma=sig, A=Aamp))

    # 'cphase': closure phases over baseline triangles u1+u2+u3 = 0
    V = 15
    u1 = rng.uniform(-0.25, 0.25, size=(4, V, 2)); u2 = rng.uniform(-0.25, 0.25, size=(4, V, 2)); u3 = -(u1 + u2)
    Acp = np.stack([dft(u1), dft(u2), dft(u3)], axis=1)                 # (nt, 3, V, P)
    tgt = rng.uniform(-np.pi, np.pi, size=(4, V)).astype(np.float32)
    sig = np.full((4, V), 0.3, dtype=np.float32)
    out = O.value_and_grad(params, 'eht', 'cphase', tgt, sig, Acp, t_frames, rt, pred, scale=1.0)
    save('case_cphase.npz', out, dict(target=tgt, sigma=sig, A=Acp))
    print('written case_amp.npz, case_cphase.npz')


if __name__ == '__main__':
    main()
