"""cmpotentials - host-side potentials with the reference's names and call contract.

Every function is ``name(y, params=None) -> (U, dUdphi, d2Udphi2, d3Udphi3)`` with the shape
contract of pyflation/cmpotentials.py:52-59: ``U`` a scalar, ``dUdphi`` (nfields,),
``d2Udphi2`` (nfields, nfields), ``d3Udphi3`` (nfields,)*3 (or None where the reference
returns None).  The integration kernels use their own fp64 device versions
(``csrc/pf_potentials.cuh``, selected by ``potential_func`` name); the functions here serve
what stays on the host: the initial Hubble value (``findH``) and off-path analysis.

Potentials are declared in one table (name, fields, parameter defaults, formula); parameter
names and default values are the reference's (cmpotentials.py, lines cited per entry).
"""
import numpy as np

__all__ = []
_REGISTRY = {}


def _fields(y, nf):
    y = np.asarray(y)
    if y.ndim > 1:
        y = y[:, 0]
    return [y[2 * i] for i in range(nf)]


def _declare(name, nfields, defaults, d3_none=False):
    """Register ``formula(fields, p)`` as a reference-compatible potential function."""
    def wrap(formula):
        def potential(y, params=None):
            p = dict(defaults)
            if params:
                for key in defaults:
                    if key in params:
                        p[key] = params[key]
            U, dU, d2U, d3U = formula(_fields(y, nfields), p)
            U = np.asarray(U).item()
            if nfields == 1:
                return (U, np.atleast_1d(dU), np.atleast_2d(d2U), np.atleast_3d(d3U))
            d3 = None if d3_none else np.zeros((nfields,) * 3)
            return (U, np.array(dU), np.array(d2U), d3)
        potential.__name__ = name
        potential.__doc__ = "Return (V, dV/dphi, d2V/dphi2, d3V/dphi3) for %s; parameters %s." % (
            name, ", ".join("%s=%r" % kv for kv in defaults.items()))
        _REGISTRY[name] = potential
        globals()[name] = potential
        __all__.append(name)
        return potential
    return wrap


# ---- single field -------------------------------------------------------------------------
@_declare("msqphisq", 1, {"mass": 6.3267e-6})                       # cmpotentials.py:16-72
def _(f, p):
    m2 = p["mass"] ** 2
    return 0.5 * m2 * f[0] ** 2, m2 * f[0], m2, 0


@_declare("lambdaphi4", 1, {"lambda": 1.5506e-13})                  # :74-129
def _(f, p):
    l = p["lambda"]
    return 0.25 * l * f[0] ** 4, l * f[0] ** 3, 3 * l * f[0] ** 2, 6 * l * f[0]


@_declare("linde", 1, {"mass": 5e-8, "lambda": 1.55009e-13})        # :131-199
def _(f, p):
    m, l = p["mass"], p["lambda"]
    return (-0.5 * m ** 2 * f[0] ** 2 + 0.25 * l * f[0] ** 4 + m ** 4 / (4 * l),
            -m ** 2 * f[0] + l * f[0] ** 3, -m ** 2 + 3 * l * f[0] ** 2, 6 * l * f[0])


@_declare("hybrid2and4", 1, {"mass": 5e-8, "lambda": 1.55123e-13})  # :201-269
def _(f, p):
    m2, l = p["mass"] ** 2, p["lambda"]
    return (0.5 * m2 * f[0] ** 2 + 0.25 * l * f[0] ** 4, m2 * f[0] + l * f[0] ** 3,
            m2 + 3 * l * f[0] ** 2, 6 * l * f[0])


@_declare("phi2over3", 1, {"sigma": 3.81686e-10})                   # :271-326
def _(f, p):
    s = p["sigma"]
    return (s * f[0] ** (2.0 / 3), (2.0 / 3) * s * f[0] ** (-1.0 / 3),
            -(2.0 / 9) * s * f[0] ** (-4.0 / 3), (8.0 / 27) * s * f[0] ** (-7.0 / 3))


@_declare("msqphisq_withV0", 1, {"mass": 1.7403553e-06, "V0": 5e-10})  # :328-389
def _(f, p):
    m2 = p["mass"] ** 2
    return 0.5 * m2 * f[0] ** 2 + p["V0"], m2 * f[0], m2, 0


@_declare("step_potential", 1, {"mass": 6.3267e-6, "c": 0.0018, "d": 0.022, "phi_s": 14.84})  # :391-466
def _(f, p):
    m2, c, d = p["mass"] ** 2, p["c"], p["d"]
    phi = f[0]
    arg = (phi - p["phi_s"]) / d
    s, t = 1 / np.cosh(arg), np.tanh(arg)
    w = 1 + c * (t - 1)                      # (t - 1) as coded, not the docstring's (t + 1)
    return (0.5 * m2 * phi ** 2 * w,
            m2 * phi * w + c * m2 * phi ** 2 * s ** 2 / (2 * d),
            0.5 * m2 * (4 * c * phi * s ** 2 / d - 2 * c * phi ** 2 * s ** 2 * t / d ** 2 + 2 * w),
            0.5 * m2 * (6 * c * s ** 2 / d - 12 * c * phi * s ** 2 * t / d ** 2
                        + c * phi ** 2 * (-2 * s ** 4 / d ** 3 + 4 * s ** 2 * t ** 2 / d ** 3)))


def _bump_terms(f, p):
    m2, c, d = p["mass"] ** 2, p["c"], p["d"]
    phi = f[0]
    arg = (phi - p["phi_b"]) / d
    s, t = 1 / np.cosh(arg), np.tanh(arg)
    w = 1 + c * s
    U = 0.5 * m2 * phi ** 2 * w
    dU = m2 * phi * w - c * m2 * phi ** 2 * s * t / (2 * d)
    d2U = 0.5 * m2 * (-4 * c * phi * s * t / d
                      + c * phi ** 2 * (-s ** 3 / d ** 2 + s * t ** 2 / d ** 2) + 2 * w)
    d3U = 0.5 * m2 * (-6 * c * s * t / d + 6 * c * phi * (-s ** 3 / d ** 2 + s * t ** 2 / d ** 2)
                      + c * phi ** 2 * (5 * s ** 3 * t / d ** 3 - s * t ** 3 / d ** 3))
    return U, dU, d2U, d3U


_BUMP = {"mass": 6.3267e-6, "c": 0.0005, "d": 0.01, "phi_b": 14.84}


@_declare("bump_potential", 1, _BUMP)                               # :468-543
def _(f, p):
    return _bump_terms(f, p)


@_declare("resonance", 1, {"mass": 6.3267e-6, "c": 5e-7, "d": 0.0007})  # :545-619
def _(f, p):
    m2, c, d = p["mass"] ** 2, p["c"], p["d"]
    phi = f[0]
    sp, cp = np.sin(phi / d), np.cos(phi / d)
    w = 1 + c * sp
    return (0.5 * m2 * phi ** 2 * w,
            m2 * phi * w + c * m2 * phi ** 2 * cp / (2 * d),
            m2 * (w + 2 * c / d * cp * phi),        # the reference's expression at :615
            m2 * (3 * c / d * cp - 3 * c / d ** 2 * sp * phi - 0.5 * c / d ** 3 * cp * phi ** 2))


@_declare("bump_nothirdderiv", 1, _BUMP)                            # :621-694
def _(f, p):
    U, dU, d2U, _unused = _bump_terms(f, p)
    return U, dU, d2U, 0.0


# ---- two fields ---------------------------------------------------------------------------
@_declare("hybridquadratic", 2, {"m1": 1.395464769e-6, "m2": 9.768253382e-6})  # :696-742
def _(f, p):
    a, b = p["m1"] ** 2, p["m2"] ** 2
    return (0.5 * (a * f[0] ** 2 + b * f[1] ** 2), [a * f[0], b * f[1]], [[a, 0.0], [0.0, b]], None)


@_declare("ridge_twofield", 2, {"g": 1e-5, "m": 12e-5, "V0": 1})    # :744-789
def _(f, p):
    g, m2 = p["g"], p["m"] ** 2
    return (p["V0"] - g * f[0] - 0.5 * m2 * f[1] ** 2, [-g, -m2 * f[1]], [[0, 0], [0, -m2]], None)


@_declare("quartictwofield", 2, {"m1": 5e-6, "m2": 5e-8, "l1": 5e-10, "l2": 5e-14},
          d3_none=True)                                             # :848-894
def _(f, p):
    a, b, l1, l2 = p["m1"] ** 2, p["m2"] ** 2, p["l1"], p["l2"]
    return (0.5 * (a * f[0] ** 2 + 0.5 * l1 * f[0] ** 4 + b * f[1] ** 2 + 0.5 * l2 * f[1] ** 4),
            [a * f[0] + l1 * f[0] ** 3, b * f[1] + l2 * f[1] ** 3],
            [[a + 3 * l1 * f[0] ** 2, 0.0], [0.0, b + 3 * l2 * f[1] ** 2]], None)


@_declare("hybridquartic", 2, {"lambda": 2.3644e-6, "v": 0.1, "mu": 1e3, "phi_c": 0.01})  # :896-961
def _(f, p):
    phi, chi = f
    l4, v2, mu2 = p["lambda"] ** 4, p["v"] ** 2, p["mu"] ** 2
    pcv2 = (p["phi_c"] * p["v"]) ** 2
    off = 8 * phi * chi / pcv2
    return (l4 * ((1 - chi ** 2 / v2) ** 2 + phi ** 2 / mu2 + 2 * (phi * chi) ** 2 / pcv2),
            [l4 * (2 * phi / mu2 + 4 * phi * chi ** 2 / pcv2),
             l4 * (-4 * chi / v2 * (1 - chi ** 2 / v2) + 4 * phi ** 2 * chi / pcv2)],
            [[l4 * (2 / mu2 + 4 * chi ** 2 / pcv2), l4 * off],
             [l4 * off, l4 * (-4 / v2 * (1 - 3 * chi ** 2 / v2) + 4 * phi ** 2 / pcv2)]], None)


@_declare("inflection", 2, {"lambda": 3e3, "g": 3e-2, "r": 0.14, "m": 1.0})  # :963-1024
def _(f, p):
    phi, chi = f
    l, g, r, m2 = p["lambda"], p["g"], p["r"], p["m"] ** 2
    V0 = 0.75 * g * r + l / 24.0 * r ** 3              # as coded at :1007
    return (V0 + 0.5 * m2 * phi ** 2 + g * chi + 1 / 6.0 * l * chi ** 3
            + (g / (4 * r ** 3) + l / (8 * r)) * chi ** 4,
            [m2 * phi, g + 0.5 * l * chi ** 2 + (g / (2 * r ** 3) + l / (2 * r)) * chi ** 3],
            [[m2, 0.0], [0.0, l * chi + 1.5 * (g / r ** 3 + l / r) * chi ** 2]], None)


@_declare("hilltopaxion", 2, {"Lambda": np.sqrt(6e-6 / (4 * np.pi)), "f": 1.0, "m": 6e-6})  # :1026-1086
def _(f, p):
    phi, chi = f
    L4, m2 = p["Lambda"] ** 4, p["m"] ** 2
    w = 2 * np.pi / p["f"]
    return (0.5 * m2 * phi ** 2 + L4 * (1 - np.cos(w * chi)),
            [m2 * phi, L4 * w * np.sin(w * chi)],
            [[m2, 0.0], [0.0, L4 * w ** 2 * np.cos(w * chi)]], None)


@_declare("productexponential", 2, {"lambda": 0.05, "V_0": 5.3705e-13})  # :1088-1144
def _(f, p):
    phi, chi = f
    l, V0 = p["lambda"], p["V_0"]
    e = np.exp(-l * chi ** 2)
    off = -4 * l * chi * V0 * phi * e
    return (V0 * phi ** 2 * e, [2 * V0 * phi * e, -2 * l * chi * V0 * phi ** 2 * e],
            [[2 * V0 * e, off],
             [off, -2 * l * V0 * phi ** 2 * e * (1 - 2 * l * chi)]], None)   # (1-2 l chi) as at :1140


# ---- any number of fields -----------------------------------------------------------------
def nflation(y, params=None):                                       # :791-846
    """Return (V, dV/dphi, d2V/dphi2, None) for N-flation, V = sum 1/2 m^2 phi_a^2.
    ``params["nfields"]`` is required, as in the reference (:828)."""
    m2 = (params["mass"] if params is not None and "mass" in params else 6.3267e-6) ** 2
    nf = params["nfields"]
    y = np.asarray(y)
    if y.ndim > 1:
        y = y[:, 0]
    phis = y[0:2 * nf:2]
    return (np.sum(0.5 * m2 * phis ** 2), m2 * phis, m2 * np.eye(nf, dtype=np.complex128), None)


_REGISTRY["nflation"] = nflation
__all__.append("nflation")
del _
