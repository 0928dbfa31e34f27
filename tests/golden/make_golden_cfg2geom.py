"""Golden case of the BENCHMARKED geometry and kernel plan.   Build container only:
    python tests/golden/make_golden_cfg2geom.py

* ``kerr_a0_i12_48x48x48.npz``: REAL Kerr geodesics of the cfg2 / cfg5 physics (spin 0, inclination 12 deg, fov 40 M --
  scripts/Fit_Synthetic_LP_Flares.yaml, notebooks "Synthetic lightcurves 0") traced by the reference's own kgeo through
  oracle/ref_shim.py, + get_dataset / Keplerian Omega / doppler_factor algebra of the reference.
* ``case_cfg2geom_lc_QU.npz``: float64 oracle images / loss / gradient of a polarized (Q,U) lightcurve step over enough
  frames (> 2^20 evaluated sample-frames) that the tcgen05 family takes its one-plane FUSED backward
  (tc_bwd_fused_kernel) -- the plan bench.py times.  rmin = 6 (ISCO), rmax = 20, z_width = 4, sigma = 0.01.
  The 'lc' chi^2 is separable per frame, so the oracle runs frame chunks and sums loss and gradient (exact in f64)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..')))
from oracle import bhnerf_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402
from make_golden import smooth_J  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert ref_shim.available(), 'needs /root/reference'
    A = B = 48; G = 48; NT = 88
    spin, inc, fov = 0.0, np.deg2rad(12.0), 40.0
    geos = ref_shim.kerr_geodesics(spin, inc, fov, A, B, G)
    Omega = ref_shim.keplerian_omega(geos)
    g = ref_shim.doppler_factor(geos, Omega)
    f32 = np.float32
    inp = dict(coords=np.array([geos['x'], geos['y'], geos['z']]).astype(f32), Omega=Omega.astype(f32),
               g=g.astype(f32), dtau=geos['dtau'].astype(f32), Sigma=geos['Sigma'].astype(f32),
               t_geos=geos['t'].astype(f32), spin=spin, inclination=inc, fov_M=fov, r_o=geos['r_o'])
    np.savez_compressed(os.path.join(HERE, 'kerr_a0_i12_48x48x48.npz'), **inp)
    rmin, rmax, zw = 6.0, 20.0, 4.0
    r2 = (inp['coords'].astype(np.float64) ** 2).sum(0)
    active = (r2 >= rmin ** 2) & (r2 <= rmax ** 2) & (np.abs(inp['coords'][2]) <= zw)
    n_active = int(active.sum())
    print('in-domain samples %d of %d (%.3f); x %d frames = %d sample-frames (2^20 = %d)' % (
        n_active, active.size, n_active / active.size, NT, n_active * NT, 1 << 20))
    t_start = 9.3
    t_frames = np.linspace(9.3, 11.3, NT)
    t_inj = -(float(geos['r_o']) + fov / 4.0)
    J2 = smooth_J(2, (A, B, G), 21)
    rt = dict(coords=inp['coords'], Omega=inp['Omega'], g=inp['g'], dtau=inp['dtau'], Sigma=inp['Sigma'],
              t_geos=inp['t_geos'], t_start_obs=t_start, t_injection=t_inj, J=J2)
    pred = dict(scale=rmax, rmin=rmin, rmax=rmax, z_width=zw)
    params = O.trained_like_params(seed=7)
    flat = O.flatten_params(params).astype(f32)
    rng = np.random.default_rng(3)
    tgt = rng.normal(0.1, 0.2, size=(NT, 2)).astype(f32)
    sig = np.full_like(tgt, 0.01); off = np.zeros_like(tgt)
    loss, grads, images = 0.0, 0.0, []
    CH = 4
    for b0 in range(0, NT, CH):
        sl = slice(b0, min(b0 + CH, NT))
        out = O.value_and_grad(params, 'image', 'lc', tgt[sl], sig[sl], off[sl], t_frames[sl], rt, pred, scale=1.0)
        loss += out['loss']; grads = grads + out['grads']; images.append(out['images'].astype(f32))
        print('frames', sl, 'loss so far', loss, flush=True)
    images = np.concatenate(images)
    np.savez_compressed(os.path.join(HERE, 'case_cfg2geom_lc_QU.npz'), params_flat=flat, t_frames=t_frames, GM_c3=O.GM_C3_SGRA_HR,
                        **{k: np.asarray(v) for k, v in pred.items()}, loss=loss, images=images, grads=grads,
                        lightcurves=images.astype(np.float64).sum((-1, -2)), target=tgt, sigma=sig, J=J2,
                        t_start_obs=t_start, t_injection=t_inj, n_active=n_active)
    print('written; loss', loss, 'grad max', np.abs(grads).max())


if __name__ == '__main__':
    main()
