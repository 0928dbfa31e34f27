"""Physical constants used by the hot path (reference: bhnerf/constants.py, astropy-free).

astropy is not required: the reference only needs GM/c^3 of Sgr A* to convert frame times to M
(emission.py:183-185).  Values are astropy's (CODATA2018 c, IAU2015 nominal GM_sun)."""
import numpy as np

GM_SUN = 1.3271244e20          # m^3 s^-2
C_LIGHT = 299792458.0          # m s^-1
SECONDS = {'s': 1.0, 'min': 60.0, 'hr': 3600.0, 'h': 3600.0, 'hour': 3600.0, 'day': 86400.0}

# ISCO (bhnerf/constants.py:7-10)
z1 = lambda a: 1 + (1 - a**2)**(1 / 3) * ((1 + a)**(1 / 3) + (1 - a)**(1 / 3))
z2 = lambda a: np.sqrt(3 * a**2 + z1(a)**2)
isco_pro = lambda a: (3 + z2(a) - np.sqrt((3 - z1(a)) * (3 + z1(a) + 2 * z2(a))))
isco_retro = lambda a: (3 + z2(a) + np.sqrt((3 - z1(a)) * (3 + z1(a) + 2 * z2(a))))

sgra_mass = 4.154e6            # solar masses (bhnerf/constants.py:17)


def GM_c3(M=sgra_mass, t_units='hr'):
    """GM/c^3 for a mass of M solar masses, in `t_units` (bhnerf/constants.py:13)."""
    name = getattr(t_units, 'name', None) or str(t_units)
    name = {'h': 'hr'}.get(name, name)
    return GM_SUN * float(M) / C_LIGHT**3 / SECONDS[name]
