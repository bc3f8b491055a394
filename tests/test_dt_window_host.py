"""The windowed, certified 1-D distance transform (partsbaseddetector_b200/csrc/dt_window.cuh) compiled for the host: every line it
ACCEPTS must equal the oracle's restatement of DistanceTransform::computeRow (reference include/DistanceTransform.hpp:152-182) bit for
bit -- on smooth maps, on noise, and on quantised maps built to put break points exactly on integers and samples exactly level -- and
the lines it refuses (handed to the literal stack algorithm on the device) must stay few on smooth maps."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from conftest import ROOT

_lib = None


def wndlib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "tests", "dt_window_host.cpp")
        hdrs = [os.path.join(ROOT, "partsbaseddetector_b200", "csrc", h) for h in ("dt_window.cuh", "dt_envelope.cuh")]
        out = os.path.join(ROOT, "tests", "libdt_window_host.so")
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(p) for p in [src] + hdrs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src])
        _lib = C.CDLL(out)
        f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
        u16p = np.ctypeslib.ndpointer(np.uint16, flags="C")
        u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
        _lib.wnd_dt1d.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, f32p, u16p, u8p, C.POINTER(C.c_longlong)]
    return _lib


def run_window(src, w_sq, w_lin, os_, W):
    """returns (number of refused lines, tier-2 visits) after checking every accepted line against the oracle; None if the map is not eligible"""
    src = np.ascontiguousarray(src, np.float32)
    nl, N = src.shape
    dst = np.full((nl, N), np.nan, np.float32)
    ptr = np.full((nl, N), 0xFFFF, np.uint16)
    dirty = np.zeros(nl, np.uint8)
    t2 = C.c_longlong(0)
    rc = wndlib().wnd_dt1d(src, nl, N, w_sq, w_lin, os_, W, dst, ptr, dirty, C.byref(t2))
    if rc == -1:
        return None
    assert rc == 0
    assert not (dirty == 2).any()          # tier 1 carries the winner's sample along: it must be that sample, bit for bit
    L = oracle_lib.lib()
    for i in range(nl):
        if dirty[i]:
            continue
        rd, rp = np.empty(N, np.float32), np.empty(N, np.int32)
        L.orc_dt1d_f32(np.ascontiguousarray(src[i]), N, -float(np.float32(w_sq)), -float(np.float32(w_lin)), os_, rd, rp)
        assert np.array_equal(dst[i], rd), (i, N, os_, W)
        assert np.array_equal(ptr[i].astype(np.int32), rp), (i, N, os_, W)
    return int((dirty != 0).sum()), int(t2.value)


def smooth(rng, nl, N, amp=1.0, corr=6):
    k = np.ones(corr) / corr
    x = rng.standard_normal((nl, N + corr - 1)).astype(np.float64)
    return (amp * np.stack([np.convolve(r, k, "valid") for r in x]) * np.sqrt(corr)).astype(np.float32)


@pytest.mark.parametrize("W", [3, 5, 8])
def test_accepted_lines_equal_the_oracle_smooth_and_noise(W):
    rng = np.random.default_rng(5 + W)
    refused = total = 0
    for N in (1, 2, 3, 7, 20, 53, 160, 255, 256, 300):
        for os_ in (0, 1, -2, W - 1, W, -W):
            for w_sq, w_lin in ((0.0156, -0.014), (0.05, 0.0), (0.0101, 0.0037), (0.3, 0.2)):
                for src in (smooth(rng, 24, N, 0.3), smooth(rng, 8, N, 2.0, 2), rng.standard_normal((8, N)).astype(np.float32) * 0.5):
                    r = run_window(src, w_sq, w_lin, os_, W)
                    assert r is not None
                    refused += r[0]; total += src.shape[0]
    assert refused < total            # not everything refused: the fast path does something on these inputs


def test_smooth_maps_are_mostly_accepted():
    rng = np.random.default_rng(11)
    src = smooth(rng, 400, 160, 0.02, 8)           # neighbouring samples differ by less than the parabola's first step, as on real score maps
    for os_ in (1, -5, 5):
        d, t2 = run_window(src, 0.0156, -0.014, os_, 5)
        assert d <= 20, d              # < 5 % of the lines
        # tier 1 counts only the candidates within 2 samples of the position: the |os| - 2 positions at a line's end whose window
        # centre lies that far outside the line always go to tier 2
        assert t2 < 0.01 * src.size + src.shape[0] * max(abs(os_) - 2, 0)


@pytest.mark.parametrize("W", [3, 5])
def test_quantised_maps_ties_and_integer_break_points(W):
    """a = -2^-k, b = 0 or a multiple of a, samples on a coarse binary grid: intersections fall exactly on integers and half integers,
    neighbouring candidates tie exactly -- every such position must be refused, never guessed."""
    rng = np.random.default_rng(23 + W)
    tot_ref = tot = 0
    for N in (5, 33, 100, 257):
        for os_ in (0, 2, -1):
            for w_sq, w_lin in ((0.0625, 0.0), (0.125, -0.125), (0.03125, 0.0625), (0.5, 0.0)):
                for q in (0.0625, 0.25, 1.0):
                    src = (np.round(smooth(rng, 16, N, 0.6, 3) / q) * q).astype(np.float32)
                    r = run_window(src, w_sq, w_lin, os_, W)
                    assert r is not None
                    tot_ref += r[0]; tot += 16
                src = np.zeros((4, N), np.float32)                     # constant lines: every neighbour pair intersects at a half integer
                run_window(src, w_sq, w_lin, os_, W)
    assert tot_ref > 0                 # the construction does produce refusals


def test_large_values_nonfinite_and_ineligible_maps():
    rng = np.random.default_rng(3)
    N = 80
    src = smooth(rng, 12, N, 0.4)
    big = src * 1e6
    run_window(big, 0.0156, 0.0, 0, 5)
    huge = src.copy(); huge[:, 10] = 3e9
    d, _ = run_window(huge, 0.0156, 0.0, 0, 5)
    assert d == 12                     # beyond ylim: all refused
    bad = src.copy(); bad[0, 5] = np.nan; bad[1, 7] = np.inf; bad[2, 9] = -np.inf
    dst = np.zeros_like(bad); ptr = np.zeros(bad.shape, np.uint16); dirty = np.zeros(12, np.uint8)
    assert wndlib().wnd_dt1d(bad, 12, N, 0.0156, 0.0, 0, 5, dst, ptr, dirty, None) == 0
    assert dirty[0] and dirty[1] and dirty[2]
    assert run_window(src, -0.0156, 0.0, 0, 5) is None      # a > 0: the reference then builds a lower envelope
    assert run_window(src, 0.0, 0.1, 0, 5) is None          # a = 0
    assert run_window(src, 0.0156, 0.0, 6, 5) is None       # anchor beyond the window: the map keeps the stack kernel
    assert run_window(src, 0.0156, 0.0, -6, 5) is None
    assert run_window(src, 0.0156, 0.0, 5, 5) is not None


def test_near_tie_sweep_around_integer_break_points():
    """two-sample bumps whose intersection is walked across an integer in float-ulp steps: accepted results must follow the oracle on both sides"""
    N, W = 40, 5
    w_sq, w_lin = 0.0156, -0.014
    a, b = -float(np.float32(w_sq)), -float(np.float32(w_lin))
    base = np.full(N, -3.0, np.float32)
    lines = []
    for x0 in (10, 17, 30):
        for dx in (1, 2, 3):
            x1 = x0 + dx
            for target in (x0 + 1, x1, x1 + 1):
                # y1 - y0 such that the intersection sits at `target`
                dy = (target * 2 * a * dx) + b * dx - a * (x1 * x1 - x0 * x0)
                for k in range(-6, 7):
                    l = base.copy()
                    l[x0] = np.float32(1.0)
                    l[x1] = np.nextafter(np.float32(1.0 + dy), np.float32(np.inf if k > 0 else -np.inf)) if k else np.float32(1.0 + dy)
                    for _ in range(abs(k) - 1):
                        l[x1] = np.nextafter(l[x1], np.float32(np.inf if k > 0 else -np.inf))
                    lines.append(l)
    src = np.stack(lines)
    r = run_window(src, w_sq, w_lin, 0, W)
    assert r is not None


def test_fuzz_short_quantised_lines_local_replay():
    """many short lines on coarse value grids with 'round' parabolas: nearly every line has positions only the local replay (tier 3)
    can decide -- isolated ties, ties at the line ends, runs of ties (refused) -- and every accepted line must equal the oracle."""
    rng = np.random.default_rng(2024)
    refused = total = 0
    for it in range(160):
        N = int(rng.integers(1, 48))
        W = int(rng.choice([3, 5, 8]))
        os_ = int(rng.integers(-W, W + 1))
        w_sq = float(rng.choice([0.0625, 0.125, 0.03125, 0.25, 0.0156, 0.0101]))
        w_lin = float(rng.choice([0.0, 0.0625, -0.125, 0.014, -0.0037]))
        q = float(rng.choice([0.03125, 0.125, 0.5, 1e-4]))
        amp = float(rng.choice([0.05, 0.3, 1.5]))
        src = (np.round(smooth(rng, 96, N, amp, int(rng.integers(1, 6))) / q) * q).astype(np.float32)
        r = run_window(src, w_sq, w_lin, os_, W)
        assert r is not None
        refused += r[0]; total += 96
    assert 0 < refused < total


def test_long_score_scale_lines_are_accepted():
    """4096-sample lines at the scale of the model's response maps: the float spacing at positions up to 4095 is 16x that of a VGA
    line, so several positions per line are open after tier 1 and a few after tier 2 -- the local replay must take care of them."""
    rng = np.random.default_rng(77)
    src = (rng.standard_normal((24, 4096)) * 0.01).astype(np.float32)
    d, t2 = run_window(src, 0.015, 0.003, 2, 5)
    assert d <= 2, d
    assert t2 > 24                     # tier 1 does leave positions open on such lines
