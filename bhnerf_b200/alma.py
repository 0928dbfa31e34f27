"""Mirror of the forward-only model-selection helpers of the reference's ``bhnerf/alma.py`` (the 39 x 4 sweep users wait
on after training): everything here is a thin loop over network.image_plane_checkpoint, i.e. the forward kernel."""
import os

import numpy as np

from . import network


def chi2_lightcurves(raytracing_args, checkpoint_dir, t, data, sigma=1.0, rmin=0.0, rmax=np.inf, batchsize=20):
    """bhnerf/alma.py:83-86: chi^2 per frame of the checkpoint's lightcurves against ``data`` (nt, S)."""
    image_plane = network.image_plane_checkpoint(raytracing_args, checkpoint_dir, t, rmin, rmax, batchsize)
    return float(np.sum(((image_plane.sum(axis=(-1, -2)) - data) / sigma) ** 2) / len(t))


def chi2_df(inclinations, spins, seeds, params, checkpoint_fmt, t, data, stokes=['I', 'Q', 'U'], sigma=1.0, rot_angle=0.0,
            num_subpixel_rays=1, raytracing_args_fn=None, final_step=50000):
    """bhnerf/alma.py:88-117: table of chi^2 over an inclination (or spin) grid x seeds.  The reference builds the
    geodesics of every grid point with ``alma.get_raytracing_args`` (kgeo ray tracing + parallel transport: setup code
    outside this package); here that step is the callable ``raytracing_args_fn(inc_rad, spin)``."""
    import pandas as pd
    if raytracing_args_fn is None:
        raise NotImplementedError('pass raytracing_args_fn(inc_rad, spin) -> raytracing args (alma.get_raytracing_args '
                                  'of the reference: kgeo tracing + parallel transport, out of this package\'s scope)')
    inclinations, spins = np.atleast_1d(inclinations), np.atleast_1d(spins)
    if len(inclinations) == 1 and len(spins) > 1:
        indices, index_name = spins, 'spin'
        inclinations = np.full_like(spins, inclinations)
    elif len(inclinations) > 1 and len(spins) == 1:
        indices, index_name = inclinations, 'inc'
        spins = np.full_like(inclinations, spins)
    elif len(inclinations) > 1 and len(spins) > 1:
        raise AttributeError('not implemented')
    else:
        indices, index_name = inclinations, 'inc'
    inc_prev = spin_prev = np.nan
    rta = None
    data_fit = np.full((len(indices), len(seeds)), fill_value=np.nan)
    for i, (inc, spin) in enumerate(zip(inclinations, spins)):
        for j, seed in enumerate(seeds):
            checkpoint_dir = checkpoint_fmt.format(indices[i], seed)
            if os.path.exists(os.path.join(checkpoint_dir, 'checkpoint_%d' % final_step)):
                if (inc_prev != inc) or (spin_prev != spin):
                    rta = raytracing_args_fn(np.deg2rad(inc), spin)
                    inc_prev, spin_prev = inc, spin
                data_fit[i, j] = chi2_lightcurves(rta, checkpoint_dir, t, data, sigma)
    df = pd.DataFrame(data_fit, index=indices, columns=['seed %d' % k for k in range(len(seeds))])
    df.index.name = index_name
    return df
