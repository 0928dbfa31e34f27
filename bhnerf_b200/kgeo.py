"""Mirror of the hot-path functions of the reference's ``bhnerf/kgeo.py``."""
import numpy as np
import torch

from . import _lib, engine
from ._lib import check


def radiative_trasfer(emission, g, dtau, Sigma, use_jax=False):
    """bhnerf/kgeo.py:595-622 (sic): sum over the last axis of g^2 * emission * dtau * Sigma.
    emission (..., *img, G) with g/dtau/Sigma (*img, G); returns a device tensor (..., *img)."""
    lib = _lib.load()
    dev = torch.device('cuda')
    e = engine._dev_f32(emission, dev)
    gg, dt, Sg = [engine._dev_f32(a, dev) for a in (g, dtau, Sigma)]
    G = e.shape[-1]
    P = int(np.prod(gg.shape[:-1])) if gg.dim() > 1 else 1
    R = e.numel() // (P * G)
    out = torch.empty((R, P), dtype=torch.float32, device=dev)
    check(lib.bhnerf_radiative_transfer(engine._ptr(e), engine._ptr(gg.expand(gg.shape).contiguous()),
                                        engine._ptr(dt.contiguous()), engine._ptr(Sg.contiguous()), R, P, G,
                                        engine._ptr(out), engine._stream()))
    return out.reshape(tuple(e.shape[:-1]))


def _get(geos, k):
    return np.asarray(geos[k] if hasattr(geos, '__getitem__') else getattr(geos, k), dtype=np.float64)


def azimuthal_velocity_vector(geos, Omega):
    """bhnerf/kgeo.py:199-223 (host, numpy): contravariant u^mu = (u^t, 0, 0, u^t*Omega) from the Kerr metric.
    Returns an array (..., 4).  Setup-time helper (runs once per (spin, inclination))."""
    r, th, a, M = _get(geos, 'r'), _get(geos, 'theta'), _get(geos, 'spin'), _get(geos, 'M')
    Sigma = r ** 2 + a ** 2 * np.cos(th) ** 2
    Delta = r ** 2 + a ** 2 - 2 * M * r
    Xi = (r ** 2 + a ** 2) ** 2 - a ** 2 * Delta * np.sin(th) ** 2
    g_tt = -(1 - 2 * M * r / Sigma)
    g_phph = Xi * np.sin(th) ** 2 / Sigma
    g_tph = -2 * M * a * r * np.sin(th) ** 2 / Sigma
    Om = np.asarray(Omega, dtype=np.float64)
    with np.errstate(invalid='ignore', divide='ignore'):
        ut = 1 / np.sqrt(-(g_tt + 2 * Om * g_tph + g_phph * Om ** 2))
    z = np.zeros_like(ut)
    return np.stack([ut, z, z, ut * Om], axis=-1)


def doppler_factor(geos, umu, fillna=0.0):
    """bhnerf/kgeo.py:225-248: g = E / -(k_mu u^mu); with u^r = u^theta = 0 only k_t = -E and
    k_phi = E*lam survive (bhnerf/kgeo.py:111-114)."""
    E, lam = _get(geos, 'E'), _get(geos, 'lam')
    umu = np.asarray(umu, dtype=np.float64)
    with np.errstate(invalid='ignore', divide='ignore'):
        g = E / -(-E * umu[..., 0] + E * lam * umu[..., 3])
    if not ((isinstance(fillna, bool) and fillna is False) or fillna is None):
        g = np.where(np.isnan(g), fillna, g)
    return g


def geodesic_inputs(geos, Omega=None, omega_sign=None, fillna=0.0, device=None):
    """Tracer output -> the geometry entries of network.raytracing_args, on the GPU (C ABI: bhnerf_geodesic_inputs).
    ``geos``: mapping / namespace with float64 r, theta, phi, t, mino of shape (*img, G), ``lam`` (per ray, shape *img
    or broadcastable (*img, G)), ``spin`` and optionally ``M``.  ``Omega=None`` selects the Keplerian field
    sign(spin + eps) sqrt(M) / (r^1.5 + spin sqrt(M)) (Tutorial3 cell 2); an array is used as given.
    Returns a dict of float32 device tensors: coords (3,*img,G), Omega, g, dtau, Sigma, t_geos (*img,G)."""
    lib = _lib.load()
    dev = torch.device(device if device is not None else 'cuda')
    f64 = lambda a: torch.as_tensor(np.array(a, dtype=np.float64, order='C'), device=dev)
    r = f64(_get(geos, 'r'))
    shape = tuple(r.shape)
    G = shape[-1]
    P = int(np.prod(shape[:-1]))
    th, ph, t, mino = [f64(_get(geos, k)).reshape(P, G).contiguous() for k in ('theta', 'phi', 't', 'mino')]
    lam = np.asarray(_get(geos, 'lam'), dtype=np.float64)
    lam = np.broadcast_to(lam, shape)[..., 0] if lam.ndim == len(shape) else np.broadcast_to(lam, shape[:-1])
    lam = f64(lam).reshape(P).contiguous()
    a = float(_get(geos, 'spin'))
    try:
        M = float(_get(geos, 'M'))
    except (KeyError, AttributeError):
        M = 1.0
    Om_in = None if Omega is None else f64(np.broadcast_to(np.asarray(Omega, dtype=np.float64), shape)).reshape(P, G).contiguous()
    if omega_sign is None:
        omega_sign = float(np.sign(a + np.finfo(float).eps))
    out = {k: torch.empty((P, G), dtype=torch.float32, device=dev) for k in ('Omega', 'g', 'dtau', 'Sigma', 't_geos')}
    coords = torch.empty((3, P, G), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.bhnerf_geodesic_inputs(engine._ptr(r.reshape(P, G).contiguous()), engine._ptr(th), engine._ptr(ph),
                                         engine._ptr(t), engine._ptr(mino), engine._ptr(lam), engine._ptr(Om_in), P, G, a, M,
                                         float(omega_sign), float(fillna), engine._ptr(coords), engine._ptr(out['Omega']),
                                         engine._ptr(out['g']), engine._ptr(out['dtau']), engine._ptr(out['Sigma']),
                                         engine._ptr(out['t_geos']), engine._stream()))
    res = {k: v.reshape(shape) for k, v in out.items()}
    res['coords'] = coords.reshape((3,) + shape)
    return res


def polarization_factors(geos, Omega=None, b_consts=None, Q_frac=0.5, rmin=0.0, rmax=np.inf, z_width=np.inf,
                         spectral_index=1, omega_sign=None, device=None):
    """The Stokes factors ``J = (I, Q, U)`` of ``network.raytracing_args`` on the GPU (C ABI: bhnerf_polarization_factors):
    alma.image_plane_model's chain (bhnerf/alma.py:47-60) -- azimuthal_velocity_vector, doppler_factor,
    magnetic_field_fluid_frame(**b_consts) normalised by its mean strength in the recovery domain, parallel_transport(Q_frac,
    V_frac=0) (bhnerf/kgeo.py:199-248, 274-313, 438-519), nan_to_num.  ``geos``: mapping / namespace with float64 r, theta,
    affine of shape (*img, G), per-ray lam, eta, alpha, beta (shape *img or broadcast (*img, G)), spin, inc.
    ``b_consts`` = dict(arad, avert, ator).  Returns a float32 device tensor (3, *img, G); apply emission.rotate_evpa for
    the reference's rot_angle."""
    lib = _lib.load()
    dev = torch.device(device if device is not None else 'cuda')
    b_consts = b_consts or dict(arad=0.0, avert=1.0, ator=0.0)
    f64 = lambda a: torch.as_tensor(np.array(a, dtype=np.float64, order='C'), device=dev)
    r = _get(geos, 'r')
    shape = tuple(r.shape)
    G = shape[-1]
    P = int(np.prod(shape[:-1]))

    def per_ray(k):
        v = np.asarray(_get(geos, k), dtype=np.float64)
        v = np.broadcast_to(v, shape)[..., 0] if v.ndim == len(shape) else np.broadcast_to(v, shape[:-1])
        return f64(v).reshape(P).contiguous()
    rr, th, aff = [f64(_get(geos, k)).reshape(P, G).contiguous() for k in ('r', 'theta', 'affine')]
    lam, eta, alpha, beta = [per_ray(k) for k in ('lam', 'eta', 'alpha', 'beta')]
    a = float(_get(geos, 'spin')); inc = float(_get(geos, 'inc'))
    Om_in = None if Omega is None else f64(np.broadcast_to(np.asarray(Omega, dtype=np.float64), shape)).reshape(P, G).contiguous()
    if omega_sign is None:
        omega_sign = float(np.sign(a + np.finfo(float).eps))
    J = torch.empty((3, P, G), dtype=torch.float32, device=dev)
    nws = lib.bhnerf_polarization_workspace_bytes(P, G)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.bhnerf_polarization_factors(engine._ptr(rr), engine._ptr(th), engine._ptr(aff), engine._ptr(lam),
                                              engine._ptr(eta), engine._ptr(alpha), engine._ptr(beta), engine._ptr(Om_in), P, G,
                                              a, inc, float(omega_sign), float(b_consts['arad']), float(b_consts['avert']),
                                              float(b_consts['ator']), float(Q_frac), float(rmin), float(min(rmax, 1e300)),
                                              float(min(z_width, 1e300)), int(spectral_index), engine._ptr(J), engine._ptr(ws),
                                              nws, engine._stream()))
    return J.reshape((3,) + shape)
