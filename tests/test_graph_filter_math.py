"""CPU check of the arithmetic behind the graph kernel's exact sqrt-free filter
(sac_cot_b200/csrc/kernels_graph.cu, DESIGN.md "S1 exact filter").

The kernel decides  P(x,y) = |RN(RN(sqrt x) - RN(sqrt y))| < tau  from
    S = x+y;  U = S - tau2f;  Q = fma(x*y, -4, U*U);  Theta = (U*U) * 3 * 2^-20
and trusts sign(Q) only when |Q| > Theta and S > lo = max(4 tau2f, 2^-50).  Here the same fp32
sequence is replayed with numpy (individually rounded fp32 ops; the single-rounding fma is
emulated in float64, where UU - 4*XY is exact for operands this close in magnitude) on millions
of pairs concentrated around the decision boundary, and every "sure" verdict must equal the
literal predicate."""
import numpy as np
import pytest

f32 = np.float32


def literal(x, y, tau):
    return np.abs(np.sqrt(x) - np.sqrt(y)) < tau  # fp32 sqrt / sub are correctly rounded in numpy


def fast_filter(x, y, tau):
    tau2f = f32(tau) * f32(tau)
    lo = max(f32(4.0) * tau2f, f32(2.0 ** -50))
    S = x + y
    U = S - tau2f
    UU = U * U
    XY = x * y
    Q = (UU.astype(np.float64) - 4.0 * XY.astype(np.float64)).astype(f32)  # fma: one rounding
    T = UU * f32(3.0 * 2.0 ** -20)   # re-uses the product Q needs; S > 4 tau2f gives U >= 0.75 S, so T >= 26.9u S^2
    with np.errstate(invalid="ignore"):
        sure = (np.abs(Q) > T) & (S > lo)
    return sure, Q < 0


def near_boundary_pairs(rng, n, tau, scale):
    """(x, y) squared lengths whose roots differ by tau * (1 + tiny), over a range of magnitudes."""
    a = (rng.random(n) * scale + tau * 2.5).astype(np.float64)
    eps = (rng.standard_normal(n) * 4e-7) + rng.choice([0.0, 0.0, 1e-4, -1e-4, 1e-2, -1e-2], n)
    sign = rng.choice([-1.0, 1.0], n)
    b = np.abs(a + sign * tau * (1.0 + eps))
    # squares rounded to fp32 plus a few ulps of jitter, as the kernel's s2/d2 would be
    x = (a * a).astype(f32)
    y = (b * b).astype(f32)
    x = np.nextafter(x, f32(np.inf) * rng.choice([-1, 1], n).astype(f32)) if n else x
    return x, y


@pytest.mark.parametrize("tau,scale", [(0.1, 5.0), (0.6, 90.0), (0.01, 0.3), (1e-3, 1e3), (5.0, 1e4), (1e-6, 1e-3)])
def test_sure_verdicts_equal_literal_predicate(tau, scale):
    rng = np.random.default_rng(1234)
    n = 2_000_000
    x, y = near_boundary_pairs(rng, n, tau, scale)
    sure, neg = fast_filter(x, y, f32(tau))
    lit = literal(x, y, f32(tau))
    wrong = sure & (neg != lit)
    assert not wrong.any(), (int(wrong.sum()), x[wrong][:4], y[wrong][:4])
    # the filter must also be useful: almost everything away from the boundary is decided
    xa = (rng.random(n) * scale).astype(f32) ** 2
    ya = (rng.random(n) * scale).astype(f32) ** 2
    sure, neg = fast_filter(xa, ya, f32(tau))
    assert not (sure & (neg != literal(xa, ya, f32(tau)))).any()
    big = (xa + ya) > f32(8 * tau * tau)
    assert sure[big].mean() > 0.995


@pytest.mark.parametrize("tau,scale", [(0.1, 5.0), (0.6, 90.0), (0.01, 0.3), (1e-3, 1e3)])
def test_filter_on_fused_squared_lengths_still_agrees_with_the_literal_predicate(tau, scale):
    """The kernel's filter forms x and y with fused multiply-adds (within 4u of the specified values, i.e. at most two
    ulps away); the literal predicate — and the fallback — use the specified ones.  A sure verdict on the perturbed
    values must still equal the literal predicate on the exact ones."""
    rng = np.random.default_rng(4321)
    n = 2_000_000
    x, y = near_boundary_pairs(rng, n, tau, scale)
    lit = literal(x, y, f32(tau))
    for _ in range(2):   # up to two ulps, independent directions per element and per operand
        xp = np.nextafter(x, np.where(rng.random(n) < 0.5, f32(np.inf), f32(-np.inf)).astype(f32))
        yp = np.nextafter(y, np.where(rng.random(n) < 0.5, f32(np.inf), f32(-np.inf)).astype(f32))
        xp = np.nextafter(xp, np.where(rng.random(n) < 0.5, f32(np.inf), f32(-np.inf)).astype(f32))
        yp = np.nextafter(yp, np.where(rng.random(n) < 0.5, f32(np.inf), f32(-np.inf)).astype(f32))
        sure, neg = fast_filter(xp, yp, f32(tau))
        wrong = sure & (neg != lit)
        assert not wrong.any(), (int(wrong.sum()), x[wrong][:4], y[wrong][:4])
        assert sure.any()          # (the sample sits right on the boundary: most of it is NOT sure, by design)


def test_exact_ties_and_degenerate_inputs_fall_through():
    tau = f32(0.5)
    # |a-b| == tau exactly: Q* = 0 -> never "sure"
    a = np.arange(1, 2000, dtype=np.float64)
    x = (a * a).astype(f32)
    y = ((a + 0.5) ** 2).astype(f32)
    sure, _ = fast_filter(x, y, tau)
    assert not sure.any()
    assert not literal(x, y, tau).any()  # strict inequality
    # zeros, denormals, infinities and NaNs are never "sure"
    specials = np.array([0.0, 1e-45, 1e-40, 1e-30, np.inf, np.nan, 3e38], f32)
    X, Y = np.meshgrid(specials, specials)
    with np.errstate(over="ignore", invalid="ignore"):
        sure, neg = fast_filter(X.ravel(), Y.ravel(), tau)
        lit = literal(X.ravel(), Y.ravel(), tau)
    assert not (sure & (neg != lit)).any()
    assert not sure[np.isnan(X.ravel()) | np.isnan(Y.ravel())].any()


@pytest.mark.parametrize("tau", [0.1, 0.6, 1e-3, 5.0])
def test_short_lengths_just_above_the_lower_limit(tau):
    """Theta = (U*U) * 3 * 2^-20 is smallest relative to (x+y)^2 where S is just above lo = 4 tau2f (U ~ 0.75 S): lengths of
    a few tau, difference ~ tau.  Sure verdicts there, on values up to two ulps off, must still equal the literal
    predicate on the exact ones."""
    rng = np.random.default_rng(77)
    n = 2_000_000
    a = (rng.random(n) * 3.0 * tau).astype(np.float64)
    eps = (rng.standard_normal(n) * 4e-7) + rng.choice([0.0, 0.0, 1e-4, -1e-4], n)
    b = np.abs(a + rng.choice([-1.0, 1.0], n) * tau * (1.0 + eps))
    x, y = (a * a).astype(f32), (b * b).astype(f32)
    lit = literal(x, y, f32(tau))
    up = lambda v: np.nextafter(v, np.where(rng.random(n) < 0.5, f32(np.inf), f32(-np.inf)).astype(f32))
    xp, yp = up(up(x)), up(up(y))
    for xs, ys in ((x, y), (xp, yp)):
        sure, neg = fast_filter(xs, ys, f32(tau))
        wrong = sure & (neg != lit)
        assert not wrong.any(), (int(wrong.sum()), x[wrong][:4], y[wrong][:4])
    # the regime is exercised: some of the sample lies above the limit and is decided
    sure, _ = fast_filter(x, y, f32(tau))
    assert sure.any() and (~sure).any()
