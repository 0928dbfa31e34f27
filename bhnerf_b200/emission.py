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
    Om = engine._dev_f32(Omega, dev); Om = (Om.expand(pts) if Om.dim() == 0 else Om).reshape(N).contiguous()
    tg = engine._dev_f32(t_geos, dev); tg = (tg.expand(pts) if tg.dim() == 0 else tg).reshape(N).contiguous()
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
