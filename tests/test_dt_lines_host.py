"""The product's parallel-in-q 1-D distance transform (partsbaseddetector_b200/csrc/dt_lines.cuh: phase A per sample, phase B per
pop site, phase C scatter-max + prefix maximum) run on the CPU by a 32-thread warp emulator and compared bit for bit with the
oracle's restatement of DistanceTransform::computeRow (reference include/DistanceTransform.hpp:152-182).  The same header is what
dt_lines / dt_cols_mix run on the device; the GPU parity tests check the device compilation."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from conftest import ROOT
from test_dt_envelope_host import gen

_lib = None


def dtl():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "tests", "dt_lines_host.cpp")
        hdrs = [os.path.join(ROOT, "partsbaseddetector_b200", "csrc", h) for h in ("dt_lines.cuh", "dt_envelope.cuh")]
        out = os.path.join(ROOT, "tests", "libdt_lines_host.so")
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(p) for p in [src] + hdrs):
            subprocess.check_call(["g++", "-O2", "-std=c++20", "-ffp-contract=off", "-pthread", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", out, src])
        _lib = C.CDLL(out)
        f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
        u16p = np.ctypeslib.ndpointer(np.uint16, flags="C")
        _lib.dtl_dt1d.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, f32p, u16p, C.POINTER(C.c_longlong)]
    return _lib


def check(src, w_sq, w_lin, os_, b, kreg):
    nl, N = src.shape
    L = oracle_lib.lib()
    dst = np.full((nl, N), np.nan, np.float32)
    ptr = np.full((nl, N), 0xFFFF, np.uint16)
    pops = (C.c_longlong * 2)(0, 0)
    assert dtl().dtl_dt1d(np.ascontiguousarray(src), nl, N, w_sq, w_lin, os_, b, kreg, dst, ptr, pops) == 0
    for i in range(nl):
        rd, rp = np.empty(N, np.float32), np.empty(N, np.int32)
        L.orc_dt1d_f32(np.ascontiguousarray(src[i]), N, -float(np.float32(w_sq)), -float(np.float32(w_lin)), os_, rd, rp)
        assert np.array_equal(dst[i], rd), (i, N, os_, b, kreg)
        assert np.array_equal(ptr[i].astype(np.int32), rp), (i, N, os_, b, kreg)
    check.exact = pops[1]
    return pops[0]


@pytest.mark.parametrize("kind", ["noise", "smooth", "spikes", "ties", "convex", "flat"])
def test_parallel_schedule_equals_oracle(kind):
    rng = np.random.default_rng(hash(kind) % 2 ** 31 + 5)
    for trial in range(14):
        N = int(rng.choice([1, 2, 3, 7, 31, 32, 33, 64, 65, 118, 158, 159, 255, 256, 257, 500]))
        nl = int(rng.integers(1, 20))
        w_sq = float(np.float32(rng.uniform(0.005, 0.08)))
        w_lin = float(np.float32(rng.uniform(-0.03, 0.03)))
        os_ = int(rng.integers(-7, 8))
        b = int(rng.choice([1, 2, 4, 8, 16, 32]))
        check(gen(rng, kind, nl, N), w_sq, w_lin, os_, b, 0)
        if N <= 256:
            check(gen(rng, kind, nl, N), w_sq, w_lin, os_, b, 8)


def test_long_lines_and_extreme_anchors():
    rng = np.random.default_rng(17)
    check(gen(rng, "smooth", 3, 1024), 0.01, -0.02, 3, 2, 0)
    check(gen(rng, "spikes", 2, 4096), 0.02, 0.02, -3, 1, 0)
    for os_ in (-40, 40, -200, 200):          # |anchor| beyond the line: every position owned by one end
        check(gen(rng, "noise", 8, 33), 0.05, 0.01, os_, 8, 8)
    check(gen(rng, "noise", 8, 64), 5.0, 0.0, 0, 4, 8)        # steep parabolas: every sample owns its own position
    check(gen(rng, "noise", 8, 64), 1e-4, 0.0, 1, 8, 0)       # nearly flat: one or two winners for the whole line


def test_fp32_certificates_fall_back_to_double_when_they_must():
    rng = np.random.default_rng(23)
    x = gen(rng, "smooth", 16, 158)
    check(x, 0.012, 0.005, 1, 8, 8)
    assert check.exact < 0.01 * x.size                          # score-map magnitudes: the double path is the exception
    check(x * 3e4, 0.012, 0.005, 1, 8, 8)                       # huge magnitudes: eps > 1/2, nothing can be certified
    assert check.exact >= x.size - 16 * 2
    check(x * 1e-6, 0.012, 0.005, 1, 8, 8)                      # nearly flat: break points at half-integers + tiny offsets, still certified
    q = np.round(x * 8) / 8
    check(q.astype(np.float32), 0.0625, 0.0, 0, 4, 8)           # dyadic weights and values: break points exactly on integers / equal
    assert check.exact > 0
    check(gen(rng, "ties", 8, 200), 0.03125, 0.0, -2, 8, 8)


def test_pop_sites_are_rare_on_smooth_maps():
    rng = np.random.default_rng(3)
    pops = check(gen(rng, "smooth", 32, 158), 0.012, 0.005, 1, 8, 8)
    assert pops < 0.5 * 32 * 158
    x = np.tile(np.linspace(1, 0, 158, dtype=np.float32) ** 2, (4, 1))      # smooth and monotone: no pops at all
    assert check(x, 0.012, 0.0, 0, 4, 8) == 0
