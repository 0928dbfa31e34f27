"""Mirror of the hot-path functions of the reference's ``bhnerf/emission.py`` (device execution)."""
import ctypes as C

import numpy as np
import torch

from . import _lib, constants, engine, utils
from ._lib import check


def velocity_warp_coords(coords, Omega, t_frames, t_start_obs, t_geos, t_injection, rot_axis=[0, 0, 1],
                         M=constants.sgra_mass, t_units=None, use_jax=False):
    """bhnerf/emission.py:143-211.  Returns warped coords (Bt, *pts, 3) (no Bt axis for scalar t_frames),
    NaN before injection.  Only rot_axis = z is supported (the reference docstring says the same)."""
    if list(np.asarray(rot_axis, dtype=float)) != [0.0, 0.0, 1.0]:
        raise NotImplementedError('Currently only equitorial plane rotation is supported')
    lib = _lib.load()
    if hasattr(t_start_obs, 'unit'):
        t_units = t_start_obs.unit
    GM_c3 = constants.GM_c3(M, t_units) if t_units is not None else 1.0
    dev = torch.device('cuda')
    coords = engine._dev_f32(np.asarray(coords) if not isinstance(coords, torch.Tensor) else coords, dev)
    pts = tuple(coords.shape[1:])
    N = int(np.prod(pts))
    c = coords.reshape(3, N).contiguous()
    Om = engine._dev_f32(Omega, dev); Om = (Om.reshape(()).expand(pts) if Om.numel() == 1 else Om).reshape(N).contiguous()
    tg = engine._dev_f32(t_geos, dev); tg = (tg.reshape(()).expand(pts) if tg.numel() == 1 else tg).reshape(N).contiguous()
    scalar_t = np.ndim(t_frames) == 0 and not isinstance(t_frames, torch.Tensor)
    tf = engine._dev_f32(np.atleast_1d(utils.time_value(t_frames, t_units or 'hr')), dev)
    out = torch.empty((tf.numel(), N, 3), dtype=torch.float32, device=dev)
    check(lib.bhnerf_velocity_warp_coords(engine._ptr(c), engine._ptr(Om), engine._ptr(tg), N, engine._ptr(tf),
                                          tf.numel(), float(utils.time_value(t_start_obs, t_units or 'hr')),
                                          float(GM_c3), float(t_injection), engine._ptr(out), engine._stream()))
    out = out.reshape((tf.numel(),) + pts + (3,))
    return out[0] if scalar_t else out


def fill_unsupervised_emission(emission, coords, rmin=0, rmax=np.inf, z_width=2.0, fill_value=0.0, use_jax=False):
    """bhnerf/emission.py:343-374.  emission (..., *pts) with coords (3, *pts); returns a new tensor."""
    lib = _lib.load()
    dev = torch.device('cuda')
    coords = engine._dev_f32(np.asarray(coords) if not isinstance(coords, torch.Tensor) else coords, dev)
    N = int(np.prod(coords.shape[1:]))
    e = engine._dev_f32(emission, dev).clone()
    R = e.numel() // N
    check(lib.bhnerf_fill_unsupervised_emission(engine._ptr(e), engine._ptr(coords.reshape(3, N).contiguous()), R, N,
                                                float(rmin), float(min(rmax, 3.0e38)), float(min(z_width, 3.0e38)),
                                                float(fill_value), engine._stream()))
    return e


def _geo(geos, k):
    return np.asarray(geos[k] if hasattr(geos, '__getitem__') and not hasattr(geos, k) else getattr(geos, k))


def _grid_fov(emission_0, fov):
    """Extent of the voxel grid per axis.  The reference reads it off the xarray coordinates
    (emission[dim].max() - emission[dim].min(), bhnerf/emission.py:229); plain arrays need ``fov``."""
    if fov is not None:
        return [float(f) for f in np.broadcast_to(np.asarray(fov, dtype=np.float64), (3,))]
    if hasattr(emission_0, 'dims'):
        dims = list(emission_0.dims)[-3:]
        return [float(np.asarray(emission_0[d]).max() - np.asarray(emission_0[d]).min()) for d in dims]
    raise AttributeError('emission_0 carries no coordinates: pass fov=(fx, fy, fz) (extent of the grid in M)')


def image_plane_dynamics(emission_0, geos, Omega, t_frames, t_injection, J=1.0, t_start_obs=None, slow_light=True,
                         doppler=True, rot_axis=[0, 0, 1], M=constants.sgra_mass, fov=None, t_units=None):
    """bhnerf/emission.py:234-303 on the GPU (csrc/grid.cu, mode 0): velocity warp -> trilinear lookup of the
    voxel grid (scipy map_coordinates order=1, cval=0) -> J broadcast -> ray integral.  Returns a device tensor
    (nt, [S,] A, B); an emission movie (ndim 4) gives (T, nt, [S,] A, B) as in the reference.

    ``geos``: mapping / namespace with x, y, z, t, dtau, Sigma of shape (A, B, G) and, for doppler=True, either
    ``g`` or the fields kgeo.doppler_factor needs.  ``fov`` / ``t_units`` are extensions for inputs without
    xarray coordinates / astropy units (plain numbers are taken in units of M, as the reference does)."""
    from . import kgeo as _kgeo
    if list(np.asarray(rot_axis, dtype=float)) != [0.0, 0.0, 1.0]:
        raise NotImplementedError('Currently only equitorial plane rotation is supported')
    if hasattr(t_start_obs, 'unit'):
        t_units = t_start_obs.unit
    elif t_start_obs is None and hasattr(t_frames, 'unit'):
        t_units = t_frames.unit
    GM_c3 = constants.GM_c3(M, t_units) if t_units is not None else 1.0
    tf = np.atleast_1d(np.asarray(utils.time_value(t_frames, t_units or 'hr'), dtype=np.float64))
    t0 = float(tf[0]) if t_start_obs is None else float(utils.time_value(t_start_obs, t_units or 'hr'))
    x, y, z = [np.asarray(_geo(geos, k), dtype=np.float32) for k in ('x', 'y', 'z')]
    t_geos = np.asarray(_geo(geos, 't'), dtype=np.float32) if slow_light else np.zeros_like(x)
    if doppler:
        try:
            g = np.asarray(_geo(geos, 'g'), dtype=np.float32)
        except (KeyError, AttributeError):
            g = _kgeo.doppler_factor(geos, _kgeo.azimuthal_velocity_vector(geos, Omega))
    else:
        g = np.ones_like(x)
    e0 = np.asarray(emission_0, dtype=np.float32)
    if e0.ndim not in (3, 4):
        raise AttributeError('emission_0 must be a 3D grid or a 4D movie of grids')
    fx, fy, fz = _grid_fov(emission_0, fov)
    # samples outside the grid's bounding sphere / slab give exactly 0: let the prepack drop them
    eps = 1.0 + 1e-5
    rmax = eps * 0.5 * float(np.sqrt(fx * fx + fy * fy + fz * fz))
    scene = engine.PackedScene(np.stack([x, y, z]), Omega, J, g, _geo(geos, 'dtau'), _geo(geos, 'Sigma'), t_geos, t0,
                               float(t_injection), 1.0, 0.0, rmax, eps * 0.5 * fz, GM_c3)
    grids = e0[None] if e0.ndim == 3 else e0
    outs = []
    for gr in grids:
        images, _ = engine.grid_render_fwd(scene, gr, (fx, fy, fz), tf.astype(np.float32), engine.GRID_MODE_DYNAMICS)
        Bt = images.shape[0]
        img = images.reshape((Bt,) + scene.image_shape) if not scene.polarized else \
            images.reshape((Bt, scene.S) + scene.image_shape)
        outs.append(img)
    out = outs[0] if e0.ndim == 3 else torch.stack(outs)
    if np.ndim(utils.time_value(t_frames, t_units or 'hr')) == 0:
        out = out[0] if e0.ndim == 3 else out[:, 0]
    return out


def interpolate_coords(emission, coords, fov=None):
    """bhnerf/emission.py:213-232 on the GPU (C ABI: bhnerf_interpolate_coords): trilinear lookup of a 3D emission grid at
    world coordinates ``coords`` (..., 3) -- the layout velocity_warp_coords returns -- with scipy's order-1 / cval=0 rule.
    ``fov`` as in image_plane_dynamics for grids without xarray coordinates.  Returns a device tensor of shape coords[:-1]."""
    lib = _lib.load()
    dev = torch.device('cuda')
    fx, fy, fz = _grid_fov(emission, fov)
    grid = engine._dev_f32(np.asarray(emission, dtype=np.float32) if not isinstance(emission, torch.Tensor) else emission, dev)
    assert grid.dim() == 3, 'emission must be a 3D grid'
    c = engine._dev_f32(coords, dev)
    assert c.shape[-1] == 3, 'coords must have x,y,z on the last axis'
    shape = tuple(c.shape[:-1])
    c = c.reshape(-1, 3).contiguous()
    out = torch.empty(c.shape[0], dtype=torch.float32, device=dev)
    check(lib.bhnerf_interpolate_coords(engine._ptr(grid), grid.shape[0], grid.shape[1], grid.shape[2], fx, fy, fz, 0,
                                        engine._ptr(c), c.shape[0], engine._ptr(out), engine._stream()))
    return out.reshape(shape)


def propogate_flatspace_emission(emission_0, Omega_3D, t_frames, t_start_obs=None, rot_axis=[0, 0, 1],
                                 M=constants.sgra_mass, fov=None, t_units=None):
    """bhnerf/emission.py:305-341 (sic): the 3D movie of an initial emission grid sheared by the velocity field in flat
    space -- the two GPU stages velocity_warp_coords (t_geos = t_injection = 0) and interpolate_coords on the grid's own
    voxel centres.  Returns a device tensor (nt, nx, ny, nz)."""
    e0 = np.asarray(emission_0, dtype=np.float32)
    fx, fy, fz = _grid_fov(emission_0, fov)
    axes = [np.linspace(-f / 2, f / 2, n, dtype=np.float32) for f, n in zip((fx, fy, fz), e0.shape)]
    x, y, z = np.meshgrid(*axes, indexing='ij')
    if t_start_obs is None:
        t_start_obs = t_frames[0] if np.ndim(t_frames) else t_frames
    warped = velocity_warp_coords([x, y, z], Omega_3D, t_frames, t_start_obs, 0.0, 0.0, rot_axis=rot_axis, M=M, t_units=t_units)
    return interpolate_coords(e0, warped, fov=(fx, fy, fz))


def rotate_evpa(stokes, angle, axis=0):
    """bhnerf/emission.py:395-407: rotate the linear-polarization pair (Q, U) of a Stokes stack by ``angle`` (EVPA rotates
    by the angle, Q + iU by twice it); stacks of 2 (Q,U), 3 (I,Q,U) or 4 (I,Q,U,V) components along ``axis``."""
    stokes = np.asarray(stokes)
    n = stokes.shape[axis]
    if n not in (2, 3, 4):
        raise AttributeError('Shape of stokes vector along axis={} not supported'.format(axis))
    iq = 0 if n == 2 else 1
    comps = [np.take(stokes, k, axis) for k in range(n)]
    p = np.exp(2j * angle) * (comps[iq] + 1j * comps[iq + 1])
    comps[iq], comps[iq + 1] = p.real, p.imag
    return np.stack(comps, axis=axis)
