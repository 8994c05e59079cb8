"""quaternion.calculus as scri uses it: spline-based derivative / antiderivative / definite integral along an axis
(InterpolatedUnivariateSpline of degree 3 per real component, as numpy-quaternion does when scipy is present).
TEST INFRASTRUCTURE - see oracle/refshim/quaternion/__init__.py."""
import numpy as np
from scipy.interpolate import InterpolatedUnivariateSpline


def spline_evaluation(f, t, t_out=None, axis=None, spline_degree=3, derivative_order=0, definite_integral_bounds=None):
    f = np.asarray(f)
    t = np.asarray(t, dtype=float)
    if axis is None:
        axis = [i for i, n in enumerate(f.shape) if n == t.size]
        if not axis:
            raise ValueError(f"no axis of f (shape {f.shape}) matches t (size {t.size})")
        axis = axis[0]
    fm = np.moveaxis(f, axis, 0)
    is_complex = np.iscomplexobj(fm)
    cols = fm.reshape(t.size, -1)
    parts = [cols.real, cols.imag] if is_complex else [cols]
    if definite_integral_bounds is not None:
        lo, hi = definite_integral_bounds
        shape_out = fm.shape[1:]
    else:
        if t_out is None:
            t_out = t
        t_out = np.asarray(t_out, dtype=float)
        shape_out = (t_out.size,) + fm.shape[1:]
    res = []
    for p in parts:
        outp = np.empty((p.shape[1],) if definite_integral_bounds is not None else (t_out.size, p.shape[1]))
        for j in range(p.shape[1]):
            s = InterpolatedUnivariateSpline(t, p[:, j], k=spline_degree)
            if definite_integral_bounds is not None:
                outp[j] = s.integral(lo, hi)
            elif derivative_order > 0:
                outp[:, j] = s.derivative(derivative_order)(t_out)
            elif derivative_order < 0:
                outp[:, j] = s.antiderivative(-derivative_order)(t_out)
            else:
                outp[:, j] = s(t_out)
        res.append(outp)
    out = (res[0] + 1j * res[1]) if is_complex else res[0]
    out = out.reshape(shape_out)
    if definite_integral_bounds is not None:
        return out
    return np.moveaxis(out, 0, axis)


def spline_derivative(f, t, derivative_order=1, axis=None):
    return spline_evaluation(f, t, axis=axis, derivative_order=derivative_order)


def spline_indefinite_integral(f, t, integral_order=1, axis=None):
    return spline_evaluation(f, t, axis=axis, derivative_order=-integral_order)


def spline_definite_integral(f, t, t1=None, t2=None, axis=None):
    t = np.asarray(t, dtype=float)
    if t1 is None:
        t1 = t[0]
    if t2 is None:
        t2 = t[-1]
    return spline_evaluation(f, t, axis=axis, definite_integral_bounds=(t1, t2))


derivative = spline_derivative
antiderivative = indefinite_integral = spline_indefinite_integral
definite_integral = spline_definite_integral
