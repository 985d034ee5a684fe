"""CPU: the DEVICE position core (kami_b200/csrc/chess.cuh) compiled for the host with g++ and
checked bit-for-bit against the oracle, so that the first GPU run starts from known-good rules.
The host build exists only inside tests/ (tests/hostcore); the product never loads it."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

import harness as H

HC_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcore")


@pytest.fixture(scope="module")
def hc():
    so = os.path.join(HC_DIR, "libhostcore.so")
    src = os.path.join(HC_DIR, "hostcore.cpp")
    core = os.path.join(H.ROOT, "kami_b200", "csrc", "chess.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.hc_init()
    L.hc_bootstrap.restype = C.c_float
    L.hc_bootstrap.argtypes = [C.c_void_p, C.c_float]
    return L


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_fast_legality_equals_make_move(hc):
    """move_is_legal (attack lookup under the post-move occupancy, csrc/chess.cuh) == the verdict of a real make-move on
    every pseudo-legal move of positions from seeded random games: checks, pins, en passant (incl. the discovered-check
    cases), castling, promotions and king steps next to sliders all occur."""
    rng = random.Random(11)
    tested = bad = 0
    n = C.c_int()
    for g in range(220):
        pos = np.zeros(80, np.uint8)
        hc.hc_start(vp(pos))
        for ply in range(160):
            bad += hc.hc_legality_diff(vp(pos), C.byref(n))
            tested += n.value
            buf = np.zeros(128, np.int32)
            k = hc.hc_legal_actions(vp(pos), buf.ctypes.data_as(C.POINTER(C.c_int)))
            if k == 0:
                break
            hc.hc_push(vp(pos), int(buf[rng.randrange(k)]))
    assert tested > 400000 and bad == 0, (tested, bad)


def test_descent_move_equals_make_move(hc):
    """descend_move (the latency-oriented position update of the tree descent, csrc/chess.cuh) == make_move<false, true> +
    set_full_key, all 80 bytes, on every legal move of positions from seeded random games (captures, en passant, castling
    with and without rights, promotions and the (Q3) queen-promotion quirk, double pushes all occur)."""
    rng = random.Random(12)
    tested = bad = 0
    n = C.c_int()
    for g in range(200):
        pos = np.zeros(80, np.uint8)
        hc.hc_start(vp(pos))
        for ply in range(200):
            bad += hc.hc_descend_diff(vp(pos), C.byref(n))
            tested += n.value
            buf = np.zeros(128, np.int32)
            k = hc.hc_legal_actions(vp(pos), buf.ctypes.data_as(C.POINTER(C.c_int)))
            if k == 0:
                break
            hc.hc_push(vp(pos), int(buf[rng.randrange(k)]))
    assert tested > 400000 and bad == 0, (tested, bad)
    # and the table form of decode_action the descent uses: all 4 672 action codes, white and black to move
    pos = np.zeros(80, np.uint8)
    hc.hc_start(vp(pos))
    assert hc.hc_decode_tab_diff(vp(pos)) == 0
    buf = np.zeros(128, np.int32)
    hc.hc_legal_actions(vp(pos), buf.ctypes.data_as(C.POINTER(C.c_int)))
    hc.hc_push(vp(pos), int(buf[0]))
    assert hc.hc_decode_tab_diff(vp(pos)) == 0


def test_device_core_matches_oracle(hc):
    rng = random.Random(5)
    npos = 0
    for g in range(60):
        o = H.OracleEnv()
        pos = np.zeros(80, np.uint8)
        hc.hc_start(vp(pos))
        hist = []
        while True:
            assert np.array_equal(pos, o.export())
            t, v, r = o.terminal()
            hk = np.array(hist, np.uint64)
            pre = hc.hc_terminal_pre(vp(pos), hk.ctypes.data_as(C.POINTER(C.c_uint64)), len(hist))
            buf = np.zeros(128, np.int32)
            n = hc.hc_legal_actions(vp(pos), buf.ctypes.data_as(C.POINTER(C.c_int)))
            if t and r <= 3:
                assert pre == r
            else:
                assert pre == 0
                oa = o.actions()
                assert np.array_equal(buf[:n], oa)
            pl = np.zeros(1920, np.float32)
            hc.hc_planes(vp(pos), pl.ctypes.data_as(C.POINTER(C.c_float)))
            assert np.array_equal(pl, o.observe())
            assert hc.hc_eval(vp(pos)) == o.eval()
            assert hc.hc_bootstrap(vp(pos), 1600.0) == o.bootstrap(1600.0)
            if t:
                break
            for a in oa:
                m = o.decode(int(a))
                assert hc.hc_decode(vp(pos), int(a)) == m and hc.hc_encode(vp(pos), m) == a
            npos += 1
            a = int(oa[rng.randrange(len(oa))])
            hist.append(int(o.key()))
            o.push(a)
            assert hc.hc_push(vp(pos), a) == 1
    assert npos > 8000
