"""include/mpgos_b200/fast_math.cuh (opt-in transcendentals): accuracy in ulp against 200-bit references, special cases handed to
the CUDA functions, and the built-in systems in their "_fm" form against the CPU checker (which uses the C math library)."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ulps(got, ref_mp):
    import mpmath as mp
    out = 0.0
    for g, r in zip(got, ref_mp):
        rd = float(r)
        u = np.spacing(abs(rd)) if rd != 0 else 5e-324
        out = max(out, float(abs(mp.mpf(float(g)) - r) / mp.mpf(float(u))))
    return out


def test_sincos_sincospi_pow_accuracy_in_ulp(pkg):
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    rng = np.random.default_rng(11)
    n = 4000
    # sincos: small, moderate and large (still fast-path) arguments
    for scale, bound in ((3.0, 1.6), (1.0e3, 1.6), (1.0e6, 1.6)):
        x = rng.uniform(-scale, scale, n)
        s, c, rs, rc = pkg.fast_math_probe(0, x)
        es, ec = ulps(s, [mp.sin(mp.mpf(float(v))) for v in x]), ulps(c, [mp.cos(mp.mpf(float(v))) for v in x])
        print(f"[fast_math] sincos |x| <= {scale:g}: sin {es:.2f} ulp, cos {ec:.2f} ulp (CUDA: {ulps(rs, [mp.sin(mp.mpf(float(v))) for v in x]):.2f}, {ulps(rc, [mp.cos(mp.mpf(float(v))) for v in x]):.2f})")
        assert es <= bound and ec <= bound
    x = rng.uniform(-64.0, 64.0, n)
    s, c, rs, rc = pkg.fast_math_probe(1, x)
    es = ulps(s, [mp.sin(mp.pi * mp.mpf(float(v))) for v in x])
    ec = ulps(c, [mp.cos(mp.pi * mp.mpf(float(v))) for v in x])
    print(f"[fast_math] sincospi: sin {es:.2f} ulp, cos {ec:.2f} ulp")
    assert es <= 1.8 and ec <= 1.8
    # pow: the ranges of the bubble models (x = R0/R from 1e-2 to 1e2, y = 3 gamma) and wide ones
    for (xl, xh, yl, yh) in ((0.5, 2.0, 1.0, 5.0), (1e-2, 1e2, 4.2, 4.2), (1e-3, 1e3, -6.0, 6.0), (0.9, 1.1, 0.1, 30.0), (1e-30, 1e30, -3.0, 3.0)):
        x = np.exp(rng.uniform(np.log(xl), np.log(xh), n))
        y = rng.uniform(yl, yh, n)
        p, _, rp, _ = pkg.fast_math_probe(2, x, y)
        ref = [mp.power(mp.mpf(float(a)), mp.mpf(float(b))) for a, b in zip(x, y)]
        e, ecuda = ulps(p, ref), ulps(rp, ref)
        print(f"[fast_math] pow x in [{xl:g}, {xh:g}], y in [{yl:g}, {yh:g}]: {e:.2f} ulp (CUDA: {ecuda:.2f})")
        assert e <= 3.0


def test_bundled_subroutine_accuracy_in_ulp(pkg):
    """fm::pow_sincospi_sincos (one subroutine, three interleaved evaluations): the same bounds as the separate functions."""
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    rng = np.random.default_rng(12)
    n = 4000
    for (xl, xh, yl, yh) in ((0.5, 2.0, 1.0, 5.0), (1e-2, 1e2, 4.2, 4.2), (1e-3, 1e3, -6.0, 6.0)):
        x = np.exp(rng.uniform(np.log(xl), np.log(xh), n))
        y = rng.uniform(yl, yh, n)
        p, _, rp, _ = pkg.fast_math_probe(3, x, y)
        ref = [mp.power(mp.mpf(float(a)), mp.mpf(float(b))) for a, b in zip(x, y)]
        e = ulps(p, ref)
        print(f"[fast_math] bundle pow x in [{xl:g}, {xh:g}], y in [{yl:g}, {yh:g}]: {e:.2f} ulp (CUDA: {ulps(rp, ref):.2f})")
        assert e <= 3.0
    x = rng.uniform(-64.0, 64.0, n)
    y = rng.uniform(-1.0e3, 1.0e3, n)
    s, c, _, _ = pkg.fast_math_probe(4, x, y)
    es, ec = ulps(s, [mp.sin(mp.pi * mp.mpf(float(v))) for v in x]), ulps(c, [mp.cos(mp.pi * mp.mpf(float(v))) for v in x])
    s2, c2, _, _ = pkg.fast_math_probe(5, x, y)
    es2, ec2 = ulps(s2, [mp.sin(mp.mpf(float(v))) for v in y]), ulps(c2, [mp.cos(mp.mpf(float(v))) for v in y])
    print(f"[fast_math] bundle sincospi: {es:.2f} / {ec:.2f} ulp, sincos: {es2:.2f} / {ec2:.2f} ulp")
    assert max(es, ec) <= 1.8 and max(es2, ec2) <= 1.6
    # any argument outside its fast range: all five results are the CUDA functions'
    x = np.array([np.inf, 3e9, np.nan, 0.25, 0.25])
    y = np.array([1.0, 1.0, 1.0, 1e7, np.inf])
    for f in (4, 5):
        a, b, ra, rb = pkg.fast_math_probe(f, x, y)
        assert np.array_equal(a, ra, equal_nan=True) and np.array_equal(b, rb, equal_nan=True)
    x = np.array([0.0, -2.0, np.inf, np.nan, 5e-324, 2.0, 2.0])
    y = np.array([2.0, 2.0, 0.5, 1.0, 2.0, 1e6, -1e6])
    p, _, rp, _ = pkg.fast_math_probe(3, x, y)
    assert np.array_equal(p, rp, equal_nan=True)


def test_two_reciprocals_side_by_side(pkg):
    """fm::rcp2: <= 1 ulp from the correctly rounded quotient; zero, Inf, NaN or a subnormal divisor hand both to the division."""
    rng = np.random.default_rng(13)
    n = 4000
    x = np.exp(rng.uniform(np.log(1e-200), np.log(1e200), n)) * rng.choice([-1.0, 1.0], n)
    y = np.exp(rng.uniform(np.log(1e-3), np.log(1e3), n))
    ru, rv, du, dv = pkg.fast_math_probe(6, x, y)
    eu = np.abs(ru - du) / np.spacing(np.abs(du))
    ev = np.abs(rv - dv) / np.spacing(np.abs(dv))
    print(f"[fast_math] rcp2: {eu.max():.2f} / {ev.max():.2f} ulp from the rounded quotient")
    assert eu.max() <= 1.0 and ev.max() <= 1.0
    x = np.array([0.0, np.inf, np.nan, 5e-324, 2.0, -0.0])
    y = np.array([3.0, 3.0, 3.0, 3.0, 0.0, 1e-320])
    ru, rv, du, dv = pkg.fast_math_probe(6, x, y)
    assert np.array_equal(ru, du, equal_nan=True) and np.array_equal(rv, dv, equal_nan=True)


def test_cuda_sincos_is_cuda_sin_and_cos(pkg):
    """The CUDA-library forms of the built-in systems call ::sincos(x) where the reference's definitions call sin(x) and cos(x): on
    the device the three functions must agree bit for bit (they share one reduction and one pair of kernels)."""
    rng = np.random.default_rng(5)
    for scale in (1.0, 10.0, 1e3, 1e5, 1e8, 1e15, 1e300):
        x = rng.uniform(-scale, scale, 400_000)
        s, c, s2, c2 = pkg.fast_math_probe(7, x)
        assert np.array_equal(s, s2) and np.array_equal(c, c2), scale
    x = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 1e-310, np.pi, np.pi / 2, 105414357.0, 1e22])
    s, c, s2, c2 = pkg.fast_math_probe(7, x)
    assert np.array_equal(s, s2, equal_nan=True) and np.array_equal(c, c2, equal_nan=True) and np.array_equal(np.signbit(s), np.signbit(s2))


def test_special_cases_are_the_cuda_functions(pkg):
    x = np.array([0.0, -0.0, -2.0, -8.0, np.inf, np.nan, 1.0, 2.0, 5e-324, 1e-310, 1e300, 1e-300, 3.0, 0.5, np.inf, 2.0, 2.0])
    y = np.array([2.0, -1.0, 2.0, 1.0 / 3.0, 0.5, 1.0, np.nan, np.inf, 2.0, 0.5, 2.0, 2.0, 0.0, -np.inf, -1.0, 1e6, -1e6])
    p, _, rp, _ = pkg.fast_math_probe(2, x, y)
    assert np.array_equal(p, rp, equal_nan=True)
    x = np.array([np.inf, -np.inf, np.nan, 1e7, -3e9, 1e300, 0.0, -0.0])
    for f in (0, 1):
        a, b, ra, rb = pkg.fast_math_probe(f, x)
        assert np.array_equal(a[:6], ra[:6], equal_nan=True) and np.array_equal(b[:6], rb[:6], equal_nan=True)
        assert a[6] == 0.0 and b[6] == 1.0 and a[7] == 0.0 and b[7] == 1.0


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1.0)


def test_kellermiksis_with_fast_math_against_the_checker(pkg, workloads, oracle):
    """"kellermiksis_dense_fm" (fast_math.cuh) against the CPU checker on the C math library: the same tolerances as the CUDA-libm
    form (endpoints 1e-6, step counts equal on >= 90 % of the instances)."""
    from cases import subsample
    std = subsample(workloads.keller_miksis_t5(NT=4096, ndo=200), 96)
    fm = subsample(workloads.keller_miksis_t5(NT=4096, ndo=200, fast_math=True), 96)
    ref, ratt, racc = oracle.run_oracle(std, phases=2)
    out, att, acc = pkg.run_workload(fm, phases=2)
    cuda, catt, _ = pkg.run_workload(std, phases=2)
    assert np.array_equal(out.ActualTime, ref.ActualTime)
    e = rel_err(out.ActualState, ref.ActualState).max()
    print(f"[fast_math] Keller-Miksis fm vs checker {e:.2e}, CUDA libm vs checker {rel_err(cuda.ActualState, ref.ActualState).max():.2e}; same steps {(att == ratt).mean():.3f}")
    assert e < 1e-6
    assert (att == ratt).mean() >= 0.9 and np.abs(att - ratt).max() <= 8
    assert rel_err(out.Accessories, ref.Accessories).max() < 1e-6


def test_bubble_ensemble_with_fast_math_against_the_checker(pkg, oracle):
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    std = WC.bubble_ensemble(NS=13, UPS=64, amplitude=(0.2, 2.5))
    fm = WC.bubble_ensemble(NS=13, UPS=64, amplitude=(0.2, 2.5), fast_math=True)
    ref, ratt, racc = oracle.run_coupled_oracle(std, phases=1, host_threads=8)
    for tc in (1, 0):
        out, att, acc = pkg.run_coupled_workload(fm, phases=1, tensor_core_coupling=tc)
        assert np.array_equal(out.ActualTime, ref.ActualTime)
        e = rel_err(out.ActualState, ref.ActualState).max()
        print(f"[fast_math] bubble ensemble fm (tensor cores {tc}) vs checker {e:.2e}; step difference {np.abs(att - ratt).max()}")
        assert e < 1e-8 and np.abs(att - ratt).max() <= 1
