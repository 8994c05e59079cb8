"""quaternion.means (scri/waveform_base.py:585).  TEST INFRASTRUCTURE."""
import numpy as np


def mean_rotor_in_chordal_metric(R, t=None):
    import quaternion

    if not t:
        return np.sum(R).normalized()
    mean = np.empty((4,), dtype=float)
    definite = quaternion.calculus.spline_definite_integral(quaternion.as_float_array(R), t)
    mean[:] = definite
    return quaternion.quaternion(*mean).normalized()
