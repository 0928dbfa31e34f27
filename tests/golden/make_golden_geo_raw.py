"""Raw tracer output (float64 r, theta, phi, t, mino, lam) of the SAME Kerr geodesics whose post-processed float32
arrays are in kerr_a0.2_i60_16x16x32.npz (both come from the reference's own kgeo tracer through oracle/ref_shim.py).
Pins bhnerf_geodesic_inputs (get_dataset algebra + Keplerian Omega + Doppler factor on the GPU).
Build container only:  python tests/golden/make_golden_geo_raw.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..')))
from oracle import ref_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
assert ref_shim.available(), 'needs /root/reference'
geos = ref_shim.kerr_geodesics(0.2, np.deg2rad(60.0), 16.0, 16, 16, 32)
ref = np.load(os.path.join(HERE, 'kerr_a0.2_i60_16x16x32.npz'))
assert np.array_equal(np.array([geos['x'], geos['y'], geos['z']]).astype(np.float32), ref['coords'])
np.savez_compressed(os.path.join(HERE, 'kerr_raw_a0.2_i60_16x16x32.npz'),
                    **{k: np.asarray(geos[k], dtype=np.float64) for k in ('r', 'theta', 'phi', 't', 'mino')},
                    lam=np.asarray(geos['lam'][..., 0], dtype=np.float64), spin=geos['spin'], M=geos['M'])
print('saved', {k: np.asarray(geos[k]).shape for k in ('r', 'lam')})
