"""Golden vectors of the voxel-grid renderers (build container only; needs /root/reference):
    python tests/golden/make_golden_grid.py

* ``grid_dynamics.npz``: ``ref_images`` = output of the REFERENCE'S OWN ``emission.image_plane_dynamics``
  (bhnerf/emission.py:234-303, executed under oracle/ref_shim.py with doppler=False, which needs no xarray algebra)
  on the real Kerr geodesics of kerr_a0.2_i60_16x16x32.npz and a seeded emission grid -- pins the oracle's
  image_plane_dynamics and the CUDA kernel (mode 0).  ``images_g`` / ``images_J`` = float64 oracle with the Doppler
  factor and with Stokes factors; ``images_early`` has frames before the injection time.
* ``grid_predictor.npz``: float64 oracle of GRID_Predictor (network.py:254-357) + 'full' image loss and its gradient
  w.r.t. the grid (torch autograd) -- pins mode 1 and the pull-back kernel."""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..')))
from oracle import bhnerf_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


class _Val(float):
    data = property(lambda self: float(self))

    def __sub__(self, o):
        return _Val(float(self) - float(o))


class _Coord:
    def __init__(self, a):
        self.a = np.asarray(a); self.size = self.a.size

    def max(self):
        return _Val(self.a.max())

    def min(self):
        return _Val(self.a.min())


class GridDA:
    """Minimal stand-in for the xr.DataArray the reference indexes (dims, coordinate arrays, ndarray protocol)."""

    def __init__(self, values, fov):
        self.values = np.asarray(values, dtype=np.float64)
        self.dims = ('x', 'y', 'z')
        self.ndim = self.values.ndim
        self.shape = self.values.shape
        self.coords = {d: _Coord(np.linspace(-fov / 2, fov / 2, n)) for d, n in zip(self.dims, self.values.shape)}

    def __getitem__(self, k):
        return self.coords[k]

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)


def hotspot_grid(n, fov, seed):
    rng = np.random.default_rng(seed)
    ax = np.linspace(-fov / 2, fov / 2, n)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing='ij')
    e = np.exp(-((X - 5.0) ** 2 + (Y + 1.0) ** 2 + Z ** 2) / (2 * 1.2 ** 2)) + 0.3 * np.exp(
        -((X + 4.0) ** 2 + (Y - 3.0) ** 2 + (Z - 0.5) ** 2) / (2 * 0.9 ** 2))
    return e + 0.05 * rng.uniform(size=e.shape)          # non-zero up to the faces: exercises the edge rule


def main():
    assert ref_shim.available(), 'needs /root/reference'
    geo = np.load(os.path.join(HERE, 'kerr_a0.2_i60_16x16x32.npz'))
    ns = ref_shim.load_bhnerf()
    c = O.GM_C3_SGRA_HR
    fov, n = 14.0, 24
    E0 = hotspot_grid(n, fov, 3)
    t_frames = np.array([0.0, 0.13, 0.5, 1.0])
    t_start = 0.0
    t_inj = -float(geo['r_o']) - 120.0                    # every sample is after injection (the reference's callers)
    g64 = lambda k: geo[k].astype(np.float64)
    geos = types.SimpleNamespace(x=g64('coords')[0], y=g64('coords')[1], z=g64('coords')[2], t=g64('t_geos'),
                                 dtau=g64('dtau'), Sigma=g64('Sigma'))
    tfM = (t_frames - t_start) / c
    ref = ns.emission.image_plane_dynamics(GridDA(E0, fov), geos, g64('Omega'), tfM, t_inj, J=1.0, t_start_obs=0.0,
                                           slow_light=True, doppler=False)
    base = dict(coords=geo['coords'], Omega=geo['Omega'], dtau=geo['dtau'], Sigma=geo['Sigma'], t_geos=geo['t_geos'],
                t_start_obs=t_start, t_injection=t_inj)
    ours = O.image_plane_dynamics(E0, fov, dict(base, g=np.ones_like(geo['g']), J=1.0), t_frames)
    err = np.abs(ours - ref).max() / np.abs(ref).max()
    print('oracle vs reference image_plane_dynamics (doppler=False): rel err %.3e, max %.3e' % (err, np.abs(ref).max()))
    assert err < 1e-12
    img_g = O.image_plane_dynamics(E0, fov, dict(base, g=geo['g'], J=1.0), t_frames)
    from make_golden import smooth_J
    J = smooth_J(3, geo['g'].shape, 21)
    img_J = O.image_plane_dynamics(E0, fov, dict(base, g=geo['g'], J=J), t_frames)
    t_inj_early = -float(geo['r_o']) + 30.0               # part of the samples / frames are before injection
    img_early = np.nan_to_num(O.image_plane_dynamics(E0, fov, dict(base, g=geo['g'], J=1.0, t_injection=t_inj_early,
                                                                   t_start_obs=0.1), t_frames))
    np.savez_compressed(os.path.join(HERE, 'grid_dynamics.npz'), emission_0=E0.astype(np.float32), fov=fov,
                        t_frames=t_frames, t_start_obs=t_start, t_injection=t_inj, GM_c3=c, ref_images=ref,
                        images_g=img_g, J=J, images_J=img_J, t_injection_early=t_inj_early, t_start_early=0.1,
                        images_early=img_early)

    # ---- GRID_Predictor: forward, 'full' loss, gradient w.r.t. the grid ----
    rng = np.random.default_rng(5)
    res = 20
    rmin, rmax, zw = float(geo['r_min']) + 0.5, 8.0, 4.0
    pred = dict(scale=rmax, rmin=rmin, rmax=rmax, z_width=zw)
    grid0 = (10.0 + 2.5 * rng.normal(size=(res, res, res))).astype(np.float32)     # sigmoid(v-10) off its tail
    rt = dict(base, g=geo['g'], J=1.0, t_injection=t_inj_early, t_start_obs=0.1)
    grid = torch.tensor(grid0.astype(np.float64), requires_grad=True)
    images = O.grid_predictor_images(grid, t_frames, rt, pred)
    target = rng.uniform(0, 0.5, size=tuple(images.shape)).astype(np.float32)
    loss = (((images - torch.as_tensor(target, dtype=torch.float64)) / 1.0) ** 2).sum()
    loss.backward()
    rtJ = dict(rt, J=J[1:])
    imagesJ = O.grid_predictor_images(torch.tensor(grid0.astype(np.float64)), t_frames, rtJ, pred)
    np.savez_compressed(os.path.join(HERE, 'grid_predictor.npz'), grid=grid0, t_frames=t_frames, t_start_obs=0.1,
                        t_injection=t_inj_early, GM_c3=c, images=images.detach().numpy(), target=target,
                        loss=float(loss), grad=grid.grad.numpy(), J=J[1:], images_J=imagesJ.numpy(),
                        **{k: np.float64(v) for k, v in pred.items()})
    print('grid_predictor: loss %.6e, |grad| max %.3e, images max %.3e, nonzero grad voxels %d' % (
        float(loss), np.abs(grid.grad.numpy()).max(), float(images.max()), int((grid.grad != 0).sum())))


if __name__ == '__main__':
    sys.path.insert(0, HERE)
    main()
