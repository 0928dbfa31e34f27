"""Golden of the polarization factors J = (I, Q, U): output of the REFERENCE'S OWN kgeo.azimuthal_velocity_vector,
doppler_factor, magnetic_field_fluid_frame and parallel_transport (bhnerf/kgeo.py:199-248, 274-313, 438-519), chained as in
alma.image_plane_model (bhnerf/alma.py:47-60) and executed under numpy with the mini-xarray of oracle/ref_shim.py, on real
Kerr geodesics of the reference's tracer.  Pins the oracle restatement (tests/test_oracle.py) and bhnerf_polarization_factors
(GPU test).  Build container only:  python tests/golden/make_golden_pol.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..')))
from oracle import ref_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
assert ref_shim.available(), 'needs /root/reference'
out = {}
for tag, (a, inc_deg, fov, rmin, rmax) in {'a': (0.2, 60.0, 16.0, 5.5, 8.0), 'b': (0.0, 12.0, 40.0, 6.0, 20.0)}.items():
    geos = ref_shim.kerr_geodesics(a, np.deg2rad(inc_deg), fov, 12, 12, 24)
    Om = ref_shim.keplerian_omega(geos)
    for k in ('r', 'theta', 'affine'):
        out['%s_%s' % (tag, k)] = geos[k]
    for k in ('lam', 'eta', 'alpha', 'beta'):
        out['%s_%s' % (tag, k)] = geos[k][..., 0]
    out['%s_consts' % tag] = np.array([a, np.deg2rad(inc_deg), rmin, rmax, 4.0])
    for j, bc in enumerate((dict(arad=0.0, avert=1.0, ator=0.0), dict(arad=0.3, avert=0.5, ator=0.8))):
        J, g = ref_shim.reference_polarization_factors(geos, Om, bc, 0.5, rmin, rmax, 4.0)
        out['%s_J%d' % (tag, j)] = J
        out['%s_b%d' % (tag, j)] = np.array([bc['arad'], bc['avert'], bc['ator']])
np.savez_compressed(os.path.join(HERE, 'pol_factors.npz'), **out)
print('saved', sorted(out))
