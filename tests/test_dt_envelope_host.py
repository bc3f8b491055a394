"""The product's streaming 1-D envelope (partsbaseddetector_b200/csrc/dt_envelope.cuh: eager emission, 8-entry ring, backing
store for deep pops, look-ahead table entries, reciprocal quotients) compiled for the host and compared bit for bit with the
oracle's restatement of DistanceTransform::computeRow (reference include/DistanceTransform.hpp:152-182).  The same header is
what dt_pass runs on the device; the GPU parity tests check the device compilation."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from conftest import ROOT

_lib = None


def envlib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "tests", "dt_envelope_host.cpp")
        hdr = os.path.join(ROOT, "partsbaseddetector_b200", "csrc", "dt_envelope.cuh")
        out = os.path.join(ROOT, "tests", "libdt_envelope_host.so")
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src])
        _lib = C.CDLL(out)
        f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
        u16p = np.ctypeslib.ndpointer(np.uint16, flags="C")
        _lib.envh_dt1d.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, f32p, u16p, C.POINTER(C.c_longlong), C.c_int]
        _lib.envh_dt1d_parallel.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, f32p, u16p,
                                            np.ctypeslib.ndpointer(np.int64, flags="C")]
        _lib.envh_quotient_fast.argtypes = [C.c_double, C.c_double]
        _lib.envh_quotient_fast.restype = C.c_float
        _lib.envh_quotient_exact.argtypes = [C.c_double, C.c_double]
        _lib.envh_quotient_exact.restype = C.c_float
    return _lib


def run_both(src, w_sq, w_lin, os_, maxn=None):
    """Direct emission and emission through the write-back window (what dt_pass does) against the oracle; returns the number of
    stores of the direct variant."""
    nl, N = src.shape
    maxn = maxn or N
    L = oracle_lib.lib()
    counts = []
    for window in (0, 1, 2, 3, 4, 5):              # direct, write-back window, lagged scan with LAG = 4 / 1 / 12, certified fp32 break points
        dst = np.full((nl, N), np.nan, np.float32)
        ptr = np.full((nl, N), 0xFFFF, np.uint16)
        stores = (C.c_longlong * 2)(0, 0)
        assert envlib().envh_dt1d(np.ascontiguousarray(src), nl, N, w_sq, w_lin, os_, maxn, dst, ptr, stores, window) == 0
        for i in range(nl):
            rd, rp = np.empty(N, np.float32), np.empty(N, np.int32)
            L.orc_dt1d_f32(np.ascontiguousarray(src[i]), N, -float(np.float32(w_sq)), -float(np.float32(w_lin)), os_, rd, rp)
            assert np.array_equal(dst[i], rd), (window, i, N, os_)
            assert np.array_equal(ptr[i].astype(np.int32), rp), (window, i, N, os_)
        counts.append(stores[0])
        if window == 5:
            run_both.exact = stores[1]
    # the parallel-in-q schedule of the same algorithm (prototype for the next kernel generation)
    dst = np.full((nl, N), np.nan, np.float32)
    ptr = np.full((nl, N), 0xFFFF, np.uint16)
    stats = np.zeros(4, np.int64)
    assert envlib().envh_dt1d_parallel(np.ascontiguousarray(src), nl, N, w_sq, w_lin, os_, dst, ptr, stats) == 0
    for i in range(nl):
        rd, rp = np.empty(N, np.float32), np.empty(N, np.int32)
        L.orc_dt1d_f32(np.ascontiguousarray(src[i]), N, -float(np.float32(w_sq)), -float(np.float32(w_lin)), os_, rd, rp)
        assert np.array_equal(dst[i], rd) and np.array_equal(ptr[i].astype(np.int32), rp), ("parallel", i, N, os_)
    run_both.last_stats = stats
    assert min(counts[1:]) >= nl * N               # every variant stores every index at least once
    run_both.last_counts = counts
    return counts[0]


def gen(rng, kind, nl, N):
    if kind == "noise":
        return rng.standard_normal((nl, N)).astype(np.float32)
    if kind == "smooth":       # what real score maps look like: few pops
        x = rng.standard_normal((nl, N + 8)).cumsum(axis=1)
        return (0.05 * x[:, 8:] + 0.01 * rng.standard_normal((nl, N))).astype(np.float32)
    if kind == "spikes":       # long ramps, then a spike that pops far below the 8-entry ring (backing-store reloads)
        x = np.tile(np.linspace(0, -3, N, dtype=np.float32), (nl, 1))
        for i in range(nl):
            for p in rng.integers(10, max(11, N), size=max(1, N // 25)):
                if p < N:
                    x[i, p] += rng.uniform(2, 30)
        return x
    if kind == "ties":         # quantised values: exact ties between intersections
        return (rng.integers(-3, 4, size=(nl, N)) * 0.25).astype(np.float32)
    if kind == "convex":       # every sample pops its predecessor
        t = np.arange(N, dtype=np.float32)
        return np.tile(0.5 * (t - N / 2) ** 2 / 7, (nl, 1)).astype(np.float32) + rng.standard_normal((nl, N)).astype(np.float32) * 1e-3
    if kind == "flat":
        return np.full((nl, N), rng.standard_normal(), np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["noise", "smooth", "spikes", "ties", "convex", "flat"])
def test_stream_envelope_equals_oracle(kind):
    rng = np.random.default_rng(hash(kind) % 2 ** 31)
    for trial in range(60):
        N = int(rng.choice([1, 2, 3, 7, 8, 9, 16, 17, 33, 64, 158, 159, 160, 257, 500]))
        nl = int(rng.integers(1, 33))
        w_sq = float(np.float32(rng.uniform(0.005, 0.08)))
        w_lin = float(np.float32(rng.uniform(-0.03, 0.03)))
        os_ = int(rng.integers(-7, 8))
        maxn = N + int(rng.integers(0, 5))
        run_both(gen(rng, kind, nl, N), w_sq, w_lin, os_, maxn)


def test_person_model_ranges_long_lines():
    rng = np.random.default_rng(7)
    for N in (1024, 4096):
        run_both(gen(rng, "smooth", 4, N), 0.01, -0.02, 3)
        run_both(gen(rng, "spikes", 4, N), 0.02, 0.02, -3)
        run_both(gen(rng, "noise", 2, N), 0.015, 0.0, 5)


def test_large_anchor_and_weights():
    rng = np.random.default_rng(9)
    for os_ in (-40, 40, -200, 200):          # |anchor| beyond the line: every position owned by one end
        run_both(gen(rng, "noise", 8, 33), 0.05, 0.01, os_)
    run_both(gen(rng, "noise", 8, 64), 5.0, 0.0, 0)        # steep parabolas: every sample owns its own position
    run_both(gen(rng, "noise", 8, 64), 1e-4, 0.0, 1)       # nearly flat: one or two winners for the whole line


def test_eager_emission_store_overhead_is_small():
    # rewrites after pops stay bounded: every store is either a position's first or follows a pop
    rng = np.random.default_rng(3)
    n = run_both(gen(rng, "smooth", 32, 158), 0.012, 0.005, 1)
    assert 32 * 158 <= n <= 3 * 32 * 158
    x = np.tile(np.linspace(1, 0, 158, dtype=np.float32) ** 2, (4, 1))      # smooth and monotone: no pops, exactly one store per position
    assert run_both(x, 0.012, 0.0, 0) == 4 * 158


def test_certified_fp32_break_points_fall_back_when_they_must():
    """envelope_stream_cert: on score-map-like inputs the double path is the exception; with huge magnitudes nothing can be certified
    a large share of the intersections is the reference's double expression; dyadic inputs put break points exactly on integers."""
    rng = np.random.default_rng(29)
    x = gen(rng, "smooth", 32, 158)
    run_both(x, 0.012, 0.005, 1)
    assert run_both.exact < 0.02 * x.size
    small = run_both.exact
    run_both(x * 3e4, 0.012, 0.005, 1)                            # the bound grows with the magnitudes: many more fall back
    assert run_both.exact > 10 * max(small, 1)
    run_both((np.round(x * 8) / 8).astype(np.float32), 0.0625, 0.0, 0)
    assert run_both.exact > 0
    run_both(gen(rng, "spikes", 8, 500), 0.03125, 0.25, -2)       # deep pops through the backing store: entries recomputed exactly
    run_both(gen(rng, "convex", 8, 300), 0.02, -0.01, 3)


def test_markstein_quotient_equals_exact_division():
    rng = np.random.default_rng(11)
    L = envlib()
    num = rng.standard_normal(200000) * np.exp(rng.uniform(-20, 20, 200000))
    den = -2.0 * rng.uniform(0.005, 0.08, 200000).astype(np.float32).astype(np.float64) * rng.integers(1, 32, 200000)
    for a, b in zip(num.tolist(), den.tolist()):
        assert L.envh_quotient_fast(a, b) == L.envh_quotient_exact(a, b)
    # constructed float-boundary cases: quotients that are exactly a float midpoint, or a few double ulps off it
    for m in (1.0 + 2.0 ** -24, 1.5 + 2.0 ** -24, 3.0 + 3 * 2.0 ** -23 + 2.0 ** -23 / 2):
        for d in (-0.02, -0.0625, -0.11):
            for k in range(-4, 5):
                q = m + k * 2.0 ** -52
                n_ = q * d
                assert L.envh_quotient_fast(n_, d) == L.envh_quotient_exact(n_, d)
    for a, b in ((0.0, -0.02), (np.inf, -0.02), (-np.inf, -0.02), (1e-310, -0.02), (1e300, -1e-9)):
        assert L.envh_quotient_fast(a, b) == L.envh_quotient_exact(a, b)
    assert np.isnan(L.envh_quotient_fast(np.nan, -0.02))


def test_literal_form_equals_oracle_on_non_finite_lines():
    """envelope_literal (what dt_pass / dt_pass_win run for a line with a NaN or an infinity): bit-identical to the oracle's
    restatement of computeRow on lines with NaN, +inf and -inf samples at the ends and in the middle, and on ordinary lines."""
    rng = np.random.default_rng(606)
    L = oracle_lib.lib()
    for N in (1, 2, 5, 33, 160, 300):
        for os_ in (0, 3, -2, 7):
            src = (rng.standard_normal((32, N)) * 0.3).astype(np.float32)
            bad = [np.nan, np.inf, -np.inf]
            for i in range(1, 32):
                for _ in range(1 + i % 3):
                    src[i, rng.integers(0, N)] = bad[(i + _) % 3]
            if N > 2:
                src[5, 0] = np.nan; src[6, N - 1] = np.nan; src[7, 0] = np.inf; src[8, N - 1] = -np.inf
            for w_sq, w_lin in ((0.0156, -0.014), (0.05, 0.0), (0.3, 0.2)):
                dst = np.full((32, N), -7.0, np.float32)
                ptr = np.full((32, N), 0xFFFF, np.uint16)
                assert envlib().envh_dt1d(np.ascontiguousarray(src), 32, N, w_sq, w_lin, os_, N, dst, ptr, None, 6) == 0
                for i in range(32):
                    rd, rp = np.empty(N, np.float32), np.empty(N, np.int32)
                    L.orc_dt1d_f32(np.ascontiguousarray(src[i]), N, -float(np.float32(w_sq)), -float(np.float32(w_lin)), os_, rd, rp)
                    assert np.array_equal(dst[i], rd, equal_nan=True), (i, N, os_)
                    assert np.array_equal(ptr[i].astype(np.int32), rp), (i, N, os_)
