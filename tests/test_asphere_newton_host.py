"""CPU: the even-asphere intersection the device uses (csrc/pyr_shapes.cuh asphere_t_n: Newton on the
conic's implicit function g(t) = c r^2 + c (1 + cc) w^2 - 2 w, w = z - sum a_i r^(2i+2), seeded
with the base-conic hit, predicted-convergence exit, far-branch check) restated in NumPy and
checked against a bracketing root finder on the explicit sag of surface_shape.py:529-537.
The GPU parity tests check the CUDA code; this checks the ALGORITHM over a much wider range of
surfaces than the configurations hold."""
import numpy as np
from scipy.optimize import brentq


def explicit_sag(c, cc, co, r2):
    s = 1.0 - c * c * (1.0 + cc) * r2
    if s <= 0.0:
        return np.nan
    p = 0.0
    for a in co[::-1]:
        p = p * r2 + a
    return c * r2 / (1.0 + np.sqrt(s)) + r2 * p


def device_newton(c, cc, co, r0, d, tol=1e-14, maxit=30):
    ck1 = c * (1.0 + cc)
    cc1 = 1.0 + cc
    f = d[2] - c * (d[0] * r0[0] + d[1] * r0[1] + d[2] * r0[2] * cc1)
    g = c * (r0[0] ** 2 + r0[1] ** 2 + r0[2] ** 2 * cc1) - 2.0 * r0[2]
    h = -c - cc * c * d[2] ** 2
    sq = f * f + h * g
    # MUFU-only seed: ~2^-20 relative
    t = g / (f + np.sqrt(sq)) * (1.0 + 2.0 ** -20) if sq > 0.0 else np.nan
    if not np.isfinite(t):
        t = 0.0
    prev = 0.0
    its = 0
    for _ in range(maxit):
        its += 1
        (x, y) = (r0[0] + t * d[0], r0[1] + t * d[1])
        r2 = x * x + y * y
        (p, dp) = (0.0, 0.0)
        for a in co[::-1]:
            dp = dp * r2 + p
            p = p * r2 + a
        w = (r0[2] + t * d[2]) - r2 * p
        gv = c * r2 + w * (ck1 * w - 2.0)
        xy = x * d[0] + y * d[1]
        wp = d[2] - 2.0 * (r2 * dp + p) * xy
        hg = c * xy + (ck1 * w - 1.0) * wp
        step = 0.5 * gv / hg * (1.0 + 2.0 ** -40)             # reciprocal to ~2^-40
        bad = not np.isfinite(step)
        if bad:
            step = 0.0
        t -= step
        (a_, lim) = (abs(step), tol * (1.0 + abs(t)))
        early = prev > 0.0 and a_ ** 3 <= 0.01 * lim * prev * prev
        conv = a_ <= lim
        prev = a_
        if bad or conv or early:
            break
    (x, y) = (r0[0] + t * d[0], r0[1] + t * d[1])
    r2 = x * x + y * y
    p = 0.0
    for a in co[::-1]:
        p = p * r2 + a
    w = (r0[2] + t * d[2]) - r2 * p
    if not (1.0 - ck1 * w > 0.0):
        return (np.nan, its)                                   # far branch: the explicit sag is undefined
    return (t, its)


def _cases(n, seed, coeff_scale, r_max):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        c = rng.uniform(-0.05, 0.05)
        cc = rng.uniform(-3.0, 2.0)
        scale = 10.0 ** rng.uniform(-1.0, coeff_scale)
        co = [rng.uniform(-1, 1) * 2e-3 * scale, rng.uniform(-1, 1) * 1e-5 * scale,
              rng.uniform(-1, 1) * 5e-8 * scale, rng.uniform(-1, 1) * 2e-10 * scale][:rng.integers(1, 5)]
        r0 = np.array([rng.uniform(-r_max, r_max), rng.uniform(-r_max, r_max), rng.uniform(-8.0, -2.0)])
        (ax, ay) = rng.uniform(-0.3, 0.3, 2)
        d = np.array([np.sin(ax), np.sin(ay) * np.cos(ax), np.cos(ay) * np.cos(ax)])
        yield (c, cc, co, r0, d)


def test_implicit_newton_finds_the_explicit_root():
    """Polynomial sag up to a few millimetres at 12 mm height (every asphere of the reference's
    demos is far inside): the root of the explicit equation to 1e-13, at most 6 evaluations, never
    invalid."""
    (checked, worst, itmax) = (0, 0.0, 0)
    for (c, cc, co, r0, d) in _cases(4000, 1, 0.3, 9.0):
        def f(t):
            return (r0[2] + t * d[2]) - explicit_sag(c, cc, co, (r0[0] + t * d[0]) ** 2 + (r0[1] + t * d[1]) ** 2)
        t0 = -r0[2] / d[2]
        ts = np.linspace(t0 - 15.0, t0 + 15.0, 301)
        v = np.array([f(t) for t in ts])
        sc = [i for i in range(len(ts) - 1) if np.isfinite(v[i]) and np.isfinite(v[i + 1]) and v[i] * v[i + 1] < 0]
        if not sc:
            continue
        i = min(sc, key=lambda j: abs(ts[j] - t0))
        tref = brentq(f, ts[i], ts[i + 1], xtol=1e-15, rtol=1e-15)
        (t, its) = device_newton(c, cc, co, r0, d)
        assert np.isfinite(t), (c, cc, co, r0, d)
        err = abs(t - tref) / (1.0 + abs(tref))
        assert err < 1e-13 or abs(f(t)) < 1e-13, (err, c, cc, co)
        (checked, worst, itmax) = (checked + 1, max(worst, err if err < 1.0 else 0.0), max(itmax, its))
    assert checked > 3500 and itmax <= 6


def test_far_branch_is_reported_invalid_not_wrong():
    """Surfaces whose polynomial sag dwarfs the conic (tens of millimetres at the ray height) can
    put the conic seed on the other sheet of the implicit function; the final check then returns
    NaN (an invalid ray).  Whatever CONVERGES to a finite value is a root of the explicit
    equation; a ray that misses such a surface altogether runs into the iteration cap and returns
    its last iterate, like the reference's fsolve does (surface_shape.py:448-465 marks every
    ray valid)."""
    (checked, invalid, capped) = (0, 0, 0)
    for (c, cc, co, r0, d) in _cases(4000, 2, 1.2, 12.0):
        def f(t):
            return (r0[2] + t * d[2]) - explicit_sag(c, cc, co, (r0[0] + t * d[0]) ** 2 + (r0[1] + t * d[1]) ** 2)
        (t, its) = device_newton(c, cc, co, r0, d)
        checked += 1
        if not np.isfinite(t):
            invalid += 1
            continue
        if its >= 30:
            capped += 1
            continue
        assert abs(f(t)) < 1e-10 * (1.0 + abs(t)), (c, cc, co, r0, d, t, f(t))
    assert checked == 4000 and invalid < 80 and capped < 80
