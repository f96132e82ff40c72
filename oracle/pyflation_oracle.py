"""CPU oracle for the first-order hot path of ihuston/pyflation -- TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the reference's algorithm for the path this repository accelerates
(fixed-step RK4 over a batch of k-modes that each start at their own time index, the
canonical background / first-order right-hand sides, the ``cmpotentials`` functions, the
horizon-crossing / Bunch-Davies initial-condition glue and the ``P_R`` reduction), plus the
second-order stage that reuses the same driver (``CanonicalSecondOrder`` /
``CanonicalRampedSecondOrder`` right-hand sides, ``SOCanonicalThreeStage`` set-up).

Who may use it: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- as the checker or the timed CPU arm, never as
the product.  ``pyflation_b200`` must not import it.

Parity status: PINNED.  ``tests/golden/*.npz`` hold outputs of the unmodified reference run
in the build container through ``oracle/ref_shim.py`` (generator: ``tests/golden/make_golden.py``);
``tests/test_oracle_golden.py`` checks every function here against them, together with the
reference's own known answers (``cosmomodels.py:2043`` ainit, ``solutions/fixtures.py:18``
etainit, ``analysis/test/test_adiabatic.py:81-260`` P_R / Pphi vectors).  The second-order functions
are pinned the same way (``tests/golden/so_*.npz``: the reference's ``SOCanonicalThreeStage`` runs and
single ``derivs`` calls; ``tests/test_second_order.py`` -- bit-identical on them).

All ``file:line`` citations are relative to the reference tree.  The arithmetic mirrors the
reference deliberately (complex128 carried for every row, each column integrating its own
copy of the background, the potential evaluated on column 0 only) because this is what the
CUDA path is judged against; the CUDA path itself uses a different, cheaper formulation.
"""
import numpy as np
from scipy import interpolate

NAN = np.nan


# --------------------------------------------------------------------------------------
# potentials  (cmpotentials.py:16-1144) -- (U, dU[nf], d2U[nf,nf], d3U or None)
# --------------------------------------------------------------------------------------
def _col0(y):
    y = np.asarray(y)
    return y[:, 0] if y.ndim > 1 else y


def _p(params, key, default):
    # the reference mixes "params is not None and key in params" with "params.get(key, d)"
    # guarded by truthiness; both reduce to: use params[key] when present, else default
    if params and key in params:
        return params[key]
    return default


def _single(U, dU, d2U, d3U):
    return (np.asarray(U).item(), np.atleast_1d(dU), np.atleast_2d(d2U), np.atleast_3d(d3U))


def msqphisq(y, params=None):                       # cmpotentials.py:16-72
    y = _col0(y)
    m2 = _p(params, "mass", 6.3267e-6) ** 2
    return _single(0.5 * m2 * y[0] ** 2, m2 * y[0], m2, 0)


def lambdaphi4(y, params=None):                     # cmpotentials.py:74-129
    y = _col0(y)
    lam = _p(params, "lambda", 1.5506e-13)
    return _single(0.25 * lam * y[0] ** 4, lam * y[0] ** 3, 3 * lam * y[0] ** 2, 6 * lam * y[0])


def linde(y, params=None):                          # cmpotentials.py:131-199
    y = _col0(y)
    m = _p(params, "mass", 5e-8)
    lam = _p(params, "lambda", 1.55009e-13)
    m2 = m ** 2
    return _single(-0.5 * m2 * y[0] ** 2 + 0.25 * lam * y[0] ** 4 + m ** 4 / (4 * lam),
                   -m2 * y[0] + lam * y[0] ** 3, -m2 + 3 * lam * y[0] ** 2, 6 * lam * y[0])


def hybrid2and4(y, params=None):                    # cmpotentials.py:201-269
    y = _col0(y)
    m2 = _p(params, "mass", 5e-8) ** 2
    lam = _p(params, "lambda", 1.55123e-13)
    return _single(0.5 * m2 * y[0] ** 2 + 0.25 * lam * y[0] ** 4,
                   m2 * y[0] + lam * y[0] ** 3, m2 + 3 * lam * y[0] ** 2, 6 * lam * y[0])


def phi2over3(y, params=None):                      # cmpotentials.py:271-326
    y = _col0(y)
    s = _p(params, "sigma", 3.81686e-10)
    return _single(s * y[0] ** (2.0 / 3), (2.0 / 3) * s * y[0] ** (-1.0 / 3),
                   -(2.0 / 9) * s * y[0] ** (-4.0 / 3), (8.0 / 27) * s * y[0] ** (-7.0 / 3))


def msqphisq_withV0(y, params=None):                # cmpotentials.py:328-389
    y = _col0(y)
    m2 = _p(params, "mass", 1.7403553e-06) ** 2
    V0 = _p(params, "V0", 5e-10)
    return _single(0.5 * m2 * y[0] ** 2 + V0, m2 * y[0], m2, 0)


def step_potential(y, params=None):                 # cmpotentials.py:391-466
    y = _col0(y)
    m2 = _p(params, "mass", 6.3267e-6) ** 2
    c = _p(params, "c", 0.0018)
    d = _p(params, "d", 0.022)
    phi_s = _p(params, "phi_s", 14.84)
    phi = y[0]
    phisq = phi ** 2
    arg = (phi - phi_s) / d
    s = 1 / np.cosh(arg)
    t = np.tanh(arg)
    U = 0.5 * m2 * phi ** 2 * (1 + c * (t - 1))
    dU = m2 * phi * (1 + c * (t - 1)) + c * m2 * phisq * s ** 2 / (2 * d)
    d2U = 0.5 * m2 * (4 * c * phi * s ** 2 / d - 2 * c * phisq * s ** 2 * t / d ** 2
                      + 2 * (1 + c * (t - 1)))
    d3U = 0.5 * m2 * (6 * c * s ** 2 / d - 12 * c * phi * s ** 2 * t / d ** 2
                      + c * phisq * (-2 * s ** 4 / d ** 3 + 4 * s ** 2 * t ** 2 / d ** 3))
    return _single(U, dU, d2U, d3U)


def _bump(y, params, third):
    y = _col0(y)
    m2 = _p(params, "mass", 6.3267e-6) ** 2
    c = _p(params, "c", 0.0005)
    d = _p(params, "d", 0.01)
    phi_b = _p(params, "phi_b", 14.84)
    phi = y[0]
    phisq = phi ** 2
    arg = (phi - phi_b) / d
    s = 1 / np.cosh(arg)
    t = np.tanh(arg)
    U = 0.5 * m2 * phi ** 2 * (1 + c * s)
    dU = m2 * phi * (1 + c * s) - c * m2 * phisq * s * t / (2 * d)
    d2U = 0.5 * m2 * (-4 * c * phi * s * t / d
                      + c * phisq * (-s ** 3 / d ** 2 + s * t ** 2 / d ** 2) + 2 * (1 + c * s))
    if third:
        d3U = 0.5 * m2 * (-6 * c * s * t / d + 6 * c * phi * (-s ** 3 / d ** 2 + s * t ** 2 / d ** 2)
                          + c * phisq * (5 * s ** 3 * t / d ** 3 - s * t ** 3 / d ** 3))
    else:
        d3U = 0.0
    return _single(U, dU, d2U, d3U)


def bump_potential(y, params=None):                 # cmpotentials.py:468-543
    return _bump(y, params, True)


def bump_nothirdderiv(y, params=None):              # cmpotentials.py:621-694
    return _bump(y, params, False)


def resonance(y, params=None):                      # cmpotentials.py:545-619
    y = _col0(y)
    m2 = _p(params, "mass", 6.3267e-6) ** 2
    c = _p(params, "c", 5e-7)
    d = _p(params, "d", 0.0007)
    phi = y[0]
    phisq = phi ** 2
    sp = np.sin(phi / d)
    cp = np.cos(phi / d)
    U = 0.5 * m2 * phisq * (1 + c * sp)
    dU = m2 * phi * (1 + c * sp) + c * m2 * phisq * cp / (2 * d)
    d2U = m2 * ((1 + c * sp) + 2 * c / d * cp * phi)          # as coded at :615
    d3U = m2 * (3 * c / d * cp - 3 * c / d ** 2 * sp * phi - 0.5 * c / d ** 3 * cp * phisq)
    return _single(U, dU, d2U, d3U)


def hybridquadratic(y, params=None):                # cmpotentials.py:696-742
    y = _col0(y)
    m1 = _p(params, "m1", 1.395464769e-6)
    m2 = _p(params, "m2", 9.768253382e-6)
    msq = np.array([m1, m2]) ** 2
    U = np.asarray(0.5 * (m1 ** 2 * y[0] ** 2 + m2 ** 2 * y[2] ** 2)).item()
    return U, msq * np.array([y[0], y[2]]), msq * np.eye(2), np.zeros((2, 2, 2))


def ridge_twofield(y, params=None):                 # cmpotentials.py:744-789
    y = _col0(y)
    g = _p(params, "g", 1e-5)
    m = _p(params, "m", 12e-5)
    V0 = _p(params, "V0", 1)
    U = np.asarray(V0 - g * y[0] - 0.5 * m ** 2 * y[2] ** 2).item()
    return (U, np.array([-g, -m ** 2 * y[2]]), np.array([[0, 0], [0, -m ** 2]]),
            np.zeros((2, 2, 2)))


def nflation(y, params=None):                       # cmpotentials.py:791-846
    m2 = _p(params, "mass", 6.3267e-6) ** 2
    nf = params["nfields"]                          # required, as in the reference (:828)
    y = _col0(y)
    phis = y[0:2 * nf:2]
    return (np.sum(0.5 * m2 * phis ** 2), m2 * phis, m2 * np.eye(nf, dtype=np.complex128), None)


def quartictwofield(y, params=None):                # cmpotentials.py:848-894
    y = _col0(y)
    m1 = _p(params, "m1", 5e-6)
    m2 = _p(params, "m2", 5e-8)
    l1 = _p(params, "l1", 5e-10)
    l2 = _p(params, "l2", 5e-14)
    U = np.asarray(0.5 * (m1 ** 2 * y[0] ** 2 + 0.5 * l1 * y[0] ** 4
                          + m2 ** 2 * y[2] ** 2 + 0.5 * l2 * y[2] ** 4)).item()
    dU = np.array([m1 ** 2 * y[0] + l1 * y[0] ** 3, m2 ** 2 * y[2] + l2 * y[2] ** 3])
    d2U = np.eye(2) * np.array([m1 ** 2 + 3 * l1 * y[0] ** 2, m2 ** 2 + 3 * l2 * y[2] ** 2])
    return U, dU, d2U, None


def hybridquartic(y, params=None):                  # cmpotentials.py:896-961
    y = _col0(y)
    lam = _p(params, "lambda", 2.3644e-6)
    v = _p(params, "v", 0.1)
    mu = _p(params, "mu", 1e3)
    phi_c = _p(params, "phi_c", 0.01)
    phi, chi = y[0], y[2]
    l4 = lam ** 4
    pcv2 = (phi_c * v) ** 2
    U = np.asarray(l4 * ((1 - chi ** 2 / v ** 2) ** 2 + phi ** 2 / mu ** 2
                         + 2 * (phi * chi) ** 2 / pcv2)).item()
    dU = l4 * np.array([2 * phi / mu ** 2 + 4 * phi * chi ** 2 / pcv2,
                        -4 * chi / v ** 2 * (1 - chi ** 2 / v ** 2) + 4 * phi ** 2 * chi / pcv2])
    off = 8 * phi * chi / pcv2
    d2U = l4 * np.array([[2 / mu ** 2 + 4 * chi ** 2 / pcv2, off],
                         [off, -4 / v ** 2 * (1 - 3 * chi ** 2 / v ** 2) + 4 * phi ** 2 / pcv2]])
    return U, dU, d2U, np.zeros((2, 2, 2))


def inflection(y, params=None):                     # cmpotentials.py:963-1024
    y = _col0(y)
    lam = _p(params, "lambda", 3e3)
    g = _p(params, "g", 3e-2)
    r = _p(params, "r", 0.14)
    m = _p(params, "m", 1.0)
    phi, chi = y[0], y[2]
    V0 = 0.75 * g * r + lam / 24.0 * r ** 3          # as coded at :1007
    U = np.asarray(V0 + 0.5 * m ** 2 * phi ** 2 + g * chi + 1 / 6.0 * lam * chi ** 3
                   + (g / (4 * r ** 3) + lam / (8 * r)) * chi ** 4).item()
    dU = np.array([m ** 2 * phi,
                   g + 0.5 * lam * chi ** 2 + (g / (2 * r ** 3) + lam / (2 * r)) * chi ** 3])
    d2U = np.array([[m ** 2, 0.0],
                    [0.0, lam * chi + 3 / 2 * (g / r ** 3 + lam / r) * chi ** 2]])
    return U, dU, d2U, np.zeros((2, 2, 2))


def hilltopaxion(y, params=None):                   # cmpotentials.py:1026-1086
    y = _col0(y)
    L = _p(params, "Lambda", np.sqrt(6e-6 / (4 * np.pi)))
    f = _p(params, "f", 1.0)
    m = _p(params, "m", 6e-6)
    phi, chi = y[0], y[2]
    w = 2 * np.pi / f
    U = np.asarray(0.5 * m ** 2 * phi ** 2 + L ** 4 * (1 - np.cos(w * chi))).item()
    dU = np.array([m ** 2 * phi, L ** 4 * w * np.sin(w * chi)])
    d2U = np.array([[m ** 2, 0.0], [0.0, L ** 4 * w ** 2 * np.cos(w * chi)]])
    return U, dU, d2U, np.zeros((2, 2, 2))


def productexponential(y, params=None):             # cmpotentials.py:1088-1144
    y = _col0(y)
    lam = _p(params, "lambda", 0.05)
    V0 = _p(params, "V_0", 5.3705e-13)
    phi, chi = y[0], y[2]
    e = np.exp(-lam * chi ** 2)
    U = np.asarray(V0 * phi ** 2 * e).item()
    dU = np.array([2 * V0 * phi * e, -2 * lam * chi * V0 * phi ** 2 * e])
    off = -4 * lam * chi * V0 * phi * e
    d2U = np.array([[2 * V0 * e, off],
                    [off, -2 * lam * V0 * phi ** 2 * e * (1 - 2 * lam * chi)]])   # as coded at :1140
    return U, dU, d2U, np.zeros((2, 2, 2))


POTENTIALS = {f.__name__: f for f in (
    msqphisq, lambdaphi4, linde, hybrid2and4, phi2over3, msqphisq_withV0, step_potential,
    bump_potential, resonance, bump_nothirdderiv, hybridquadratic, ridge_twofield, nflation,
    quartictwofield, hybridquartic, inflection, hilltopaxion, productexponential)}


# --------------------------------------------------------------------------------------
# RK4  (rk4.py:27-138)
# --------------------------------------------------------------------------------------
class SimRunError(Exception):
    pass


class ModelError(Exception):
    pass


def rk4_step(x, y, h, dydx, f):
    """One classical RK4 step; stage order of rk4.py:27-55."""
    half, sixth = h * 0.5, h / 6.0
    xm = x + half
    k2 = f(y + half * dydx, xm)
    k3 = f(y + half * k2, xm)
    k23 = k3 + k2
    k4 = f(y + h * k3, x + h)
    return y + sixth * (dydx + k4 + 2 * k23)


def number_of_steps(simtstart, tend, h):
    return int(np.around((tend - simtstart) / h) + 1)          # rk4.py:73


def rk4_drive(ystart, simtstart, tsix, tend, h, f):
    """rk4.py:58-138.  ``ystart`` (nvar,) or (nvar, nk); ``tsix`` (nk,) start rows.
    Returns (x (T,), y (T, nvar, nk)) with rows before a column's start left NaN."""
    if h is None:
        raise SimRunError("Need to specify h.")
    tsix = np.asarray(tsix)
    T = number_of_steps(simtstart, tend, h)
    if np.any(tsix > T):
        raise SimRunError("Start times outside range of steps.")
    ystart = np.asarray(ystart)
    if ystart.ndim == 1:
        ystart = ystart[:, None]
    first, last_start = int(tsix.min()), int(tsix.max())
    x = np.zeros(T)
    x[0] = simtstart
    if first > 0:
        x[1:first + 1] = simtstart + np.arange(1, first + 1) * h
    y = np.full((T,) + ystart.shape, NAN, dtype=ystart.dtype)
    y[tsix, :, np.arange(ystart.shape[1])] = ystart.T            # rk4.py:106-107
    for n in range(first + 1, T):
        x_prev = simtstart + (n - 1) * h
        prev = y[n - 1]
        new = rk4_step(x_prev, prev, h, f(prev, x_prev), f)
        if n <= last_start:
            keep = ~np.isnan(new)                                 # rk4.py:127-130
            y[n, keep] = new[keep]
        else:
            y[n] = new
        x[n] = simtstart + n * h
    return x, y


# --------------------------------------------------------------------------------------
# right-hand sides  (cosmomodels.py:551-571, 625-675)
# --------------------------------------------------------------------------------------
def nvar_of(nf):
    return 2 * nf * nf + 2 * nf + 1


def make_bg_rhs(potential, params, nf):
    pot = POTENTIALS[potential] if isinstance(potential, str) else potential

    def rhs(y, t):
        U, dU = pot(y, params)[0:2]
        pd = y[1:2 * nf:2]
        H = y[2 * nf]
        out = np.zeros_like(y)
        out[0:2 * nf:2] = pd
        out[1:2 * nf:2] = -(U * pd + dU[..., None]) / H ** 2
        out[2 * nf] = -0.5 * np.sum(pd ** 2, axis=0) * H
        return out
    return rhs


def make_fo_rhs(potential, params, nf, k, ainit):
    pot = POTENTIALS[potential] if isinstance(potential, str) else potential
    k = np.asarray(k)
    nb = 2 * nf + 1

    def rhs(y, t):
        pd = y[1:2 * nf:2]
        H = y[2 * nf]
        a = ainit * np.exp(t)
        U, dU, d2U = pot(y[0:nb, 0], params)[0:3]                # column 0 only (:641)
        out = np.zeros((nvar_of(nf), k.shape[0]), dtype=y.dtype)
        out[0:2 * nf:2] = pd
        out[1:2 * nf:2] = -(U * pd + dU[..., None]) / H ** 2
        out[2 * nf] = -0.5 * np.sum(pd ** 2, axis=0) * H
        dp = y[nb::2]
        dpdot = y[nb + 1::2]
        out[nb::2] = dpdot
        modes = dp.reshape(nf, nf, -1)
        # M_il = V_il + phidot_i V_l + V_i phidot_l + phidot_i phidot_l U   (:663-669)
        M = (d2U[:, :, None] + pd[:, None, :] * dU[None, :, None]
             + dU[:, None, None] * pd[None, :, :] + pd[:, None, :] * pd[None, :, :] * U)
        inner = np.einsum("ilk,ljk->ijk", M, modes).reshape(nf * nf, -1)
        out[nb + 1::2] = -(U * dpdot / H ** 2 + (k / (a * H)) ** 2 * dp + inner / H ** 2)
        return out
    return rhs


# --------------------------------------------------------------------------------------
# background helpers and initial conditions
# --------------------------------------------------------------------------------------
def find_H(U, y, nf):
    """cosmomodels.py:481-487."""
    pd = y[1:2 * nf:2]
    return np.sqrt(U / (3.0 - 0.5 * np.sum(pd ** 2)))


def default_bgystart(nf):
    """cosmomodels.py:1241-1243."""
    return np.array([18.0 / np.sqrt(nf), -0.1 / np.sqrt(nf)] * nf + [0.0])


def run_background(potential, params, nf, bgystart, tend=83.0, h=0.01, simtstart=0.0):
    """CanonicalBackground.__init__ + run (cosmomodels.py:525-571, 1343-1374)."""
    y0 = np.array(bgystart, dtype=float)
    if y0[2 * nf] == 0.0:
        y0[2 * nf] = find_H(POTENTIALS[potential](y0, params)[0], y0, nf)
    t, y = rk4_drive(y0, simtstart, np.array([0]), tend, h, make_bg_rhs(potential, params, nf))
    return t, y


def epsilon_of(bgy, nf):
    """cosmomodels.py:513-522."""
    return 0.5 * np.sum(bgy[:, 1:2 * nf:2, 0] ** 2, axis=1)


def find_inflation_end(t, eps):
    """cosmomodels.py:493-511 -- (efold refined by a cubic spline, index)."""
    if not np.any(eps > 1):
        raise ModelError("Inflation did not end during specified number of efoldings.")
    ix = int(np.where(eps >= 1)[0][0])
    tck = interpolate.splrep(t[:ix], eps[:ix])
    fine = np.linspace(t[ix - 1], t[ix], 100)
    val = interpolate.splev(fine, tck)
    return fine[np.where(val > 1)[0][0]], ix


def a_end_of(Hend, Hreh=None):
    """cosmomodels.py:987-1040 (instantaneous reheating when Hreh is None)."""
    if Hreh is None:
        Hreh = Hend
    return np.exp(-(72.3 + 2.0 / 3.0 * np.log(Hend) - 1.0 / 6.0 * np.log(Hreh)))


def crossing_indices(k, t, H, ainit, cq):
    """cosmomodels.py:1074-1131: first row with k - cq*a*H < 0, per mode (loop form)."""
    aH = ainit * np.exp(t) * H
    ix = np.empty(len(k), dtype=np.int64)
    for i, kk in enumerate(k):
        hit = np.where(np.sign(kk - cq * aH) < 0)[0]
        if hit.size == 0:
            raise ModelError("k mode %s crosses horizon after end of inflation!" % kk)
        ix[i] = hit[0]
    return ix, t[ix]


def bunch_davies(k, ts, tsix, bgy2d, eps, ainit, etainit, nf, phase=True, suppress_ix=None):
    """cosmomodels.py:1443-1492 (+ :2083-2132 no phase, :2144-2197 suppress one field)."""
    k = np.asarray(k)
    nb = 2 * nf + 1
    y0 = np.zeros((nvar_of(nf), len(k)), dtype=np.complex128)
    astar = ainit * np.exp(ts)
    Hstar = bgy2d[tsix, 2 * nf]
    y0[0:nb] = bgy2d[tsix, :].T
    amp = 1 / (astar * np.sqrt(2 * k))
    if phase:
        etastar = -1 / (astar * Hstar * (1 - eps[tsix]))
        osc = np.exp(-(k * (etastar - etainit)) * 1j)
        dp = amp * osc
        dpdot = -amp * osc * (1 + (k / (astar * Hstar)) * 1j)
    else:
        dp = amp + 0j
        dpdot = -amp * (1 + (k / (astar * Hstar)) * 1j)
    for i in range(nf):
        if suppress_ix is not None and i == suppress_ix:
            continue
        y0[nb + 2 * (i * nf + i)] = dp
        y0[nb + 2 * (i * nf + i) + 1] = dpdot
    return y0


# --------------------------------------------------------------------------------------
# P_R  (analysis/adiabatic.py:45-81, 153-285; analysis/utilities.py:12-37, 104-184)
# --------------------------------------------------------------------------------------
def pphi_matrix(modes, axis):
    """P_IJ = sum_K chi_IK conj(chi_JK); ``modes`` has its two nf-long axes at axis, axis+1."""
    m = np.moveaxis(modes, (axis, axis + 1), (0, 1))
    P = np.einsum("ik...,jk...->ij...", m, m.conj())
    return np.moveaxis(P, (0, 1), (axis, axis + 1)).astype(modes.dtype)


def pr_spectrum(phidot, modes, axis):
    phidot = np.atleast_1d(phidot)
    norm = np.sum(phidot ** 2, axis=axis) ** 2
    P = pphi_matrix(modes, axis)
    w = np.expand_dims(phidot, axis) * np.expand_dims(phidot, axis + 1) * P
    total = np.sum(w, axis=(axis, axis + 1))
    return np.real(total / norm).astype(float)


def pr_of_yresult(yresult, nf, tix=None):
    """adiabatic.Pr(m, tix): yresult (T, nvar, nk) -> (T or 1, nk) float."""
    if tix is not None:
        if tix < 0:
            tix += yresult.shape[0]
        yresult = yresult[tix:tix + 1]
    phidot = np.real(yresult[:, 1:2 * nf:2, :]).astype(np.float64)
    modes = yresult[:, 2 * nf + 1::2, :]
    modes = modes.reshape(modes.shape[0], nf, nf, modes.shape[2])
    return pr_spectrum(phidot, modes, 1)


def kscaling(k):
    return np.atleast_1d(k) ** 3 / (2 * np.pi ** 2)


# --------------------------------------------------------------------------------------
# the full two-stage first-order run (cosmomodels.py:1406-1441) as one function
# --------------------------------------------------------------------------------------
def first_order_run(potential, k, nf=1, params=None, bgystart=None, cq=50, tend=83.0, h=0.01,
                    simtstart=0.0, phase=True, suppress_ix=None, fixed_ainit=None,
                    integrate=True):
    """Returns a dict with the public result surface of FOCanonicalTwoStage."""
    params = {} if params is None else params
    k = np.asarray(k, dtype=float)
    if bgystart is None:
        bgystart = default_bgystart(nf)
    bgt, bgy = run_background(potential, params, nf, bgystart, tend, h, simtstart)
    eps = epsilon_of(bgy, nf)
    fotend, fotendindex = find_inflation_end(bgt, eps)
    Hend = bgy[fotendindex, 2 * nf]
    if fixed_ainit is None:
        ainit = a_end_of(Hend) * np.exp(-bgt[fotendindex])       # shape (1,), as the reference
    else:
        ainit = fixed_ainit
    etainit = -1 / (ainit * bgy[0, 2 * nf] * (1 - eps[0]))
    tsix, ts = crossing_indices(k, bgt[:fotendindex], bgy[:fotendindex, 2 * nf, 0], ainit, cq)
    if np.any(ts == 0):
        raise ModelError("Some k modes crossed horizon before simulation began.")
    y0 = bunch_davies(k, ts, tsix, bgy[..., 0], eps, ainit, etainit, nf, phase, suppress_ix)
    out = dict(k=k, bgt=bgt, bgy=bgy, bgepsilon=eps, fotend=fotend, fotendindex=fotendindex,
               ainit=ainit, etainit=etainit, fotstart=ts, fotstartindex=tsix, foystart=y0)
    if integrate:
        t, y = rk4_drive(y0, simtstart, tsix, fotend, h,
                         make_fo_rhs(potential, params, nf, k, ainit))
        out.update(tresult=t, yresult=y)
    return out


# --------------------------------------------------------------------------------------
# second-order stage: rkdriver_tsix driving CanonicalSecondOrder.derivs /
# CanonicalRampedSecondOrder.derivs (cosmomodels.py:678-766, 850-971) from a first-order result
# and a given source term, as SOCanonicalThreeStage sets it up (cosmomodels.py:1680-1790)
# --------------------------------------------------------------------------------------
DEFAULT_RAMP = {"a": 15.0, "b": 0.3, "c": 1, "d": 2, "e": 1}          # cosmomodels.py:875-880


def synthetic_source(t, k):
    """A smooth, deterministic stand-in for the second-order source term S(t, k) (the reference
    computes it with its `sourceterm` package, out of scope here): (T,) x (nk,) -> (T, nk) complex.
    Used by tests/golden/make_golden.py (fed to the unmodified reference) and by the tests."""
    t = np.asarray(t, dtype=float)[:, None]
    k = np.asarray(k, dtype=float)[None, :]
    amp = (k / 1e-60) ** -1.5 * np.exp(-0.08 * t)
    return amp * (np.cos(0.9 * t + 3.0 * (k / 1e-58)) + 1j * np.sin(1.3 * t + 0.5))


def so_start_indices(fotstartindex, fotstep):
    """cosmomodels.py:1706-1709: first-order start rows on the second-order (2x) time grid."""
    sotstep = fotstep * 2
    return np.around(np.asarray(fotstartindex) * (fotstep / sotstep) + sotstep / 2).astype(int), sotstep


def make_so_rhs(potential, params, k, ainit, fo_yresult, fo_simtstart, fo_tstep, bgepsilon, source,
                ramp=None, tstart=None, tstartindex=None, so_tstep=None):
    """Right-hand side of the real 4-vector (Re dphi2, Re dphi2', Im dphi2, Im dphi2') per mode.
    ``fo_yresult`` (T_fo, 5, nk) complex is the single-field first-order trajectory, ``source``
    (T_fo, nk) complex the source term (None: homogeneous equation), both sampled on the
    first-order time grid; ``ramp`` (dict a, b, c, d, e) switches on the tanh ramp of
    CanonicalRampedSecondOrder, which needs the per-mode first-order start times ``tstart``
    (e-folds), the second-order start rows ``tstartindex`` and step ``so_tstep``."""
    pot = globals()[potential] if isinstance(potential, str) else potential
    k = np.asarray(k, dtype=float)

    def rhs(y, t):
        fotix = int(np.around((t - fo_simtstart) / fo_tstep))    # :724 / :911
        fovars = fo_yresult[fotix]
        phidot, H = fovars[1], fovars[2]                          # :741 (phi is only used by the potential)
        eps = bgepsilon[fotix]
        d2U = np.asarray(pot(fovars, params)[2]).reshape(-1)[0]   # column 0 of the first-order row
        a = ainit * np.exp(t)
        if source is None:
            src = np.zeros(len(k), dtype=complex)
        else:
            src = np.array(source[fotix], dtype=complex)
            if ramp is not None:                                  # :925-944
                arg = t - tstart - ramp["b"]
                need = np.abs(arg) < ramp["e"]
                if np.any(need):
                    r = (np.tanh(ramp["a"] * arg) + ramp["c"]) / ramp["d"]
                    r[tstartindex == t / so_tstep] = 0
                    src[need] = r[need] * src[need]
        out = np.zeros((4, len(k)))
        with np.errstate(all="ignore"):
            damp, kah2 = 3 - eps, (k / (a * H)) ** 2
            mass = d2U / H ** 2 - 3 * (phidot ** 2)
            out[0] = y[1]
            out[1] = np.real(-damp * y[1] - kah2 * y[0] - mass * y[0] - src.real)   # :754-755
            out[2] = y[3]
            out[3] = np.real(-damp * y[3] - kah2 * y[2] - mass * y[2] - src.imag)   # :761-762
        return out
    return rhs


def second_order_run(potential, params, k, ainit, fo_tresult, fo_yresult, fotstart, fotstartindex,
                     bgepsilon, source, fo_tstep=0.01, fo_simtstart=0.0, ramp=None, ystart=None,
                     tstartindex=None):
    """SOCanonicalThreeStage(...).run() (cosmomodels.py:1680-1790): returns (tresult, yresult
    (T_so, 4, nk) float64, tstartindex)."""
    k = np.asarray(k, dtype=float)
    if tstartindex is None:
        tstartindex, sotstep = so_start_indices(fotstartindex, fo_tstep)
    else:
        sotstep = fo_tstep * 2
    if ystart is None:
        ystart = np.zeros((4, len(k)))
    simtstart, tend = fo_tresult[0], fo_tresult[-1]
    f = make_so_rhs(potential, params, k, ainit, fo_yresult, fo_simtstart, fo_tstep, bgepsilon, source,
                    ramp=ramp, tstart=np.asarray(fotstart, dtype=float), tstartindex=tstartindex,
                    so_tstep=sotstep)
    t, y = rk4_drive(np.asarray(ystart, dtype=float), simtstart, tstartindex, tend, sotstep, f)
    return t, y, tstartindex
