"""Small host helpers with the reference's names (bhnerf/utils.py:97-132, :215-219)."""
import numpy as np


def expand_dims(x, ndim, axis=0, use_jax=False):
    """bhnerf/utils.py:215-219."""
    for _ in range(ndim - np.array(x).ndim):
        x = np.expand_dims(x, axis=min(axis, np.array(x).ndim))
    return x


def rotation_matrix(axis, angle, use_jax=False):
    """Euler-Rodrigues rotation matrix, shape (3,3,...) (bhnerf/utils.py:97-132)."""
    axis = np.array(axis, dtype=np.float64)
    axis = axis / np.sqrt(np.dot(axis, axis))
    angle = np.asarray(angle)
    a = np.cos(angle / 2.0)
    b, c, d = np.stack([-ax * np.sin(angle / 2.0) for ax in axis])
    aa, bb, cc, dd = a * a, b * b, c * c, d * d
    bc, ad, ac, ab, bd, cd = b * c, a * d, a * c, a * b, b * d, c * d
    return np.array([[aa + bb - cc - dd, 2 * (bc + ad), 2 * (bd - ac)],
                     [2 * (bc - ad), aa + cc - bb - dd, 2 * (cd + ab)],
                     [2 * (bd + ac), 2 * (cd - ab), aa + dd - bb - cc]])


def time_value(t, t_units='hr'):
    """Strip astropy-like units if present (value in `t_units`); plain numbers are taken as `t_units`."""
    if hasattr(t, 'to') and hasattr(t, 'unit'):
        return np.asarray(t.to(t_units).value, dtype=np.float64)
    return np.asarray(t, dtype=np.float64)


def world_to_image_coords(coords, fov, npix, use_jax=False):
    """bhnerf/utils.py:160-166: (coords + fov/2) / fov * (npix - 1) per axis (last axis of coords)."""
    coords = np.asarray(coords)
    return np.stack([(coords[..., i] + fov[i] / 2.0) / fov[i] * (npix[i] - 1) for i in range(coords.shape[-1])], axis=-1)
