"""Generate the committed golden vectors under tests/golden/.   Run from the repo root, in the
build container only (needs /root/reference):   python tests/golden/make_golden.py

What is pinned by what
----------------------
* ``kerr_a0.2_i60_16x16x32.npz`` : REAL Kerr geodesics traced by the reference's own kgeo
  (kgeo/kgeo/kerr_raytracing_ana.py:48-154) + its get_dataset/doppler algebra -> the float32
  hot-path inputs (coords, Omega, g, dtau, Sigma, t_geos) of network.raytracing_args
  (bhnerf/network.py:874-892).
* ``ref_stages.npz`` : outputs of the REFERENCE'S OWN SOURCE executed under numpy
  (oracle/ref_shim.py): velocity_warp_coords (emission.py:143), fill_unsupervised_emission
  (emission.py:343), radiative_trasfer (kgeo.py:595), posenc (network.py:98), expand_dims-based
  J broadcast (network.py:415-418).  These pin the oracle.
* ``case_*.npz`` : full-path outputs (images, loss, flat gradients, visibilities) of the
  float64 oracle (oracle/bhnerf_oracle.py) for fixed seeds.  These pin the CUDA path on the GPU
  box, where neither the reference nor kgeo exist.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..')))
from oracle import bhnerf_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def smooth_J(S, shape, seed):
    """Seeded smooth Stokes factors in [-1,1] (stand-in for parallel_transport output,
    bhnerf/kgeo.py:438-519: an INPUT of the hot path)."""
    rng = np.random.default_rng(seed)
    A, B, G = shape
    a = np.linspace(0, 1, A)[:, None, None]; b = np.linspace(0, 1, B)[None, :, None]
    k = np.linspace(0, 1, G)[None, None, :]
    J = []
    for s in range(S):
        ph = rng.uniform(0, 2 * np.pi, 3); fr = rng.uniform(1, 4, 3)
        J.append(np.cos(fr[0] * a * 2 * np.pi + ph[0]) * np.cos(fr[1] * b * 2 * np.pi + ph[1])
                 * np.cos(fr[2] * k * 2 * np.pi + ph[2]))
    J = np.stack(J)
    if S == 3:
        J[0] = 0.5 + 0.5 * np.abs(J[0])      # Stokes I factor positive
    return J.astype(np.float32)


def main():
    assert ref_shim.available(), 'needs /root/reference'
    A = B = 16; G = 32
    spin, inc, fov = 0.2, np.deg2rad(60.0), 16.0
    geos = ref_shim.kerr_geodesics(spin, inc, fov, A, B, G)
    Omega = ref_shim.keplerian_omega(geos)
    g = ref_shim.doppler_factor(geos, Omega)
    f32 = np.float32
    inp = dict(coords=np.array([geos['x'], geos['y'], geos['z']]).astype(f32), Omega=Omega.astype(f32),
               g=g.astype(f32), dtau=geos['dtau'].astype(f32), Sigma=geos['Sigma'].astype(f32),
               t_geos=geos['t'].astype(f32), r_min=np.float64(geos['r'].min()),
               spin=spin, inclination=inc, fov_M=fov, r_o=geos['r_o'])
    np.savez_compressed(os.path.join(HERE, 'kerr_a0.2_i60_16x16x32.npz'), **inp)

    # ---------------- reference-source stage outputs -----------------------------------
    ns = ref_shim.load_bhnerf()
    c = O.GM_C3_SGRA_HR
    t_frames = np.array([0.0, 0.13, 0.5, 1.0])
    t_start = 0.1            # two frames before / after start: exercises the t_M<0 NaN mask
    t_inj = -float(geos['r_o']) + 30.0
    co64 = inp['coords'].astype(np.float64)
    tfM = (t_frames - t_start) / c          # same op order as emission.py:200 with GM_c3=1
    warp = ns.emission.velocity_warp_coords(co64, inp['Omega'].astype(np.float64), tfM, 0.0,
                                            inp['t_geos'].astype(np.float64), t_inj, t_units=None,
                                            use_jax=True)
    rng = np.random.default_rng(0)
    e = rng.uniform(0, 1, size=(len(t_frames), A, B, G))
    rmin, rmax, zw = float(geos['r'].min()) + 0.5, 8.0, 4.0
    fill = ns.emission.fill_unsupervised_emission(e, co64, rmin, rmax, zw, use_jax=True)
    rt = ns.kgeo.radiative_trasfer(fill, inp['g'].astype(np.float64), inp['dtau'].astype(np.float64),
                                   inp['Sigma'].astype(np.float64), use_jax=True)
    J = smooth_J(3, (A, B, G), 5).astype(np.float64)
    Je = ns.utils.expand_dims(J, fill.ndim + 1, 0, use_jax=True) * ns.utils.expand_dims(fill, fill.ndim + 1, 1, use_jax=True)
    rtJ = ns.kgeo.radiative_trasfer(np.squeeze(Je), inp['g'].astype(np.float64), inp['dtau'].astype(np.float64),
                                    inp['Sigma'].astype(np.float64), use_jax=True)
    xs = rng.uniform(-1.2, 1.2, size=(64, 3))
    pe = ns.posenc(xs, 3)
    pe32 = ns.posenc_f32(xs.astype(np.float32), 3)       # reference source with JAX-like float32 promotion
    assert pe32.dtype == np.float32
    rot = ns.utils.rotation_matrix([0, 0, 1], np.array([0.3, -2.0, 40.0]), use_jax=True)
    np.savez_compressed(os.path.join(HERE, 'ref_stages.npz'), t_frames=t_frames, t_start_obs=t_start,
                        t_injection=t_inj, GM_c3=c, warp=warp, e=e, rmin=rmin, rmax=rmax, z_width=zw,
                        fill=fill, rt=rt, J=J, rtJ=rtJ, posenc_x=xs, posenc=pe, posenc_f32=pe32, rot=rot)

    # ---------------- full-path oracle cases -------------------------------------------
    params = O.trained_like_params(seed=7)
    flat = O.flatten_params(params).astype(np.float32)
    pred = dict(scale=rmax, rmin=rmin, rmax=rmax, z_width=zw)
    base = dict(coords=inp['coords'], Omega=inp['Omega'], g=inp['g'], dtau=inp['dtau'], Sigma=inp['Sigma'],
                t_geos=inp['t_geos'], t_start_obs=t_start, t_injection=t_inj)
    P = A * B

    def save(name, out, extra):
        np.savez_compressed(os.path.join(HERE, name), params_flat=flat, t_frames=t_frames, GM_c3=c,
                            **{k: np.asarray(v) for k, v in pred.items()},
                            loss=out['loss'], images=out['images'], grads=out['grads'], **extra)

    # case 1: unpolarized, 'full' image loss
    rt1 = dict(base, J=1.0)
    tgt = rng.uniform(0, 0.5, size=(4, A, B)).astype(f32)
    out = O.value_and_grad(params, 'image', 'full', tgt, np.ones_like(tgt), np.zeros_like(tgt), t_frames,
                           rt1, pred, scale=1.0)
    save('case_image_full.npz', out, dict(target=tgt, t_start_obs=t_start, t_injection=t_inj))

    # case 2: polarized (Q,U), 'lc' loss, sigma=0.01
    J2 = smooth_J(2, (A, B, G), 11)
    rt2 = dict(base, J=J2)
    tgt = rng.normal(0, 0.2, size=(4, 2)).astype(f32)
    sig = np.full_like(tgt, 0.01); off = np.zeros_like(tgt)
    out = O.value_and_grad(params, 'image', 'lc', tgt, sig, off, t_frames, rt2, pred, scale=1.0)
    save('case_lc_QU.npz', out, dict(target=tgt, sigma=sig, J=J2, t_start_obs=t_start, t_injection=t_inj))

    # case 3: I,Q,U 'lc' with per-channel sigma (ALMA shape)
    J3 = smooth_J(3, (A, B, G), 12)
    rt3 = dict(base, J=J3)
    tgt = rng.normal(0.3, 0.2, size=(4, 3)).astype(f32)
    sig = np.broadcast_to(np.array([0.15, 1e-2, 1e-2], dtype=f32), tgt.shape).copy(); off = np.zeros_like(tgt)
    out = O.value_and_grad(params, 'image', 'lc', tgt, sig, off, t_frames, rt3, pred, scale=1.0)
    save('case_lc_IQU.npz', out, dict(target=tgt, sigma=sig, J=J3, t_start_obs=t_start, t_injection=t_inj))

    # case 4: visibilities, 'vis' chi^2 with a per-frame complex DFT matrix
    V = 20
    psize = fov / A
    xx, yy = np.meshgrid((np.arange(A) - A / 2) * psize, (np.arange(B) - B / 2) * psize, indexing='ij')
    uv = rng.uniform(-0.3, 0.3, size=(4, V, 2))
    Amat = np.exp(-2j * np.pi * (uv[..., 0:1] * xx.reshape(1, 1, -1) + uv[..., 1:2] * yy.reshape(1, 1, -1))
                  ).astype(np.complex64)
    tgtv = (rng.normal(0, 1, size=(4, V)) + 1j * rng.normal(0, 1, size=(4, V))).astype(np.complex64)
    sigv = np.full((4, V), 0.5, dtype=f32)
    out = O.value_and_grad(params, 'eht', 'vis', tgtv, sigv, Amat, t_frames, rt1, pred, scale=1.0)
    save('case_vis.npz', out, dict(target=tgtv, sigma=sigv, A=Amat, vis=out['vis'], t_start_obs=t_start,
                                   t_injection=t_inj))

    # optimiser known-answer: three adam steps on a tiny vector
    p0 = np.linspace(-1, 1, 7); mu = np.zeros(7); nu = np.zeros(7)
    traj = []
    for k in range(3):
        gk = np.cos(p0 * (k + 1)) * 0.1
        p0, mu, nu = O.adam_step(p0, gk, mu, nu, k, 1e-2, 1e-4, 10)
        traj.append(p0.copy())
    np.savez_compressed(os.path.join(HERE, 'adam_kat.npz'), traj=np.array(traj))
    print('golden vectors written to', HERE)


if __name__ == '__main__':
    main()
