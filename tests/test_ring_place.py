"""Payload-ring placement (beta9_b200/csrc/ring_place.h, the function b9gpu.cu's pushes and the
rebalance append go through), compiled for the host. The ring model is the reference's fixed-capacity
`RingBuffer` (pkg/abstractions/common/ring_buffer.go:8-96) with refuse-when-full; what is checked here
is the physical side the reference does not have: a placed segment never overlaps pending bytes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_shim", "ring_place_shim.cpp")
SO = os.path.join(HERE, "host_shim", "libringplace.so")
DEP = os.path.join(os.path.dirname(HERE), "beta9_b200", "csrc", "ring_place.h")


@pytest.fixture(scope="module")
def script():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(DEP)):
        r = subprocess.run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lib = C.CDLL(SO)
    lib.b9_ring_place_script.argtypes = [C.c_uint64, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
    lib.b9_ring_place_script.restype = C.c_long

    def run(ring, ops):
        a = np.asarray(ops, dtype=np.int64)
        out = np.zeros(a.size, dtype=np.int64)
        rc = lib.b9_ring_place_script(ring, a.ctypes.data_as(C.POINTER(C.c_int64)), a.size, out.ctypes.data_as(C.POINTER(C.c_int64)))
        assert rc == a.size, f"overlap / out of range at step {-1 - rc}: ops={list(a[:max(0, -rc)])}"
        return list(out)
    return run


PUSH = lambda b: b + 1
POP = 0


def test_exact_fill_after_wrap_is_full_not_empty(script):
    # ring 1024; A(512) B(512); pop A; C(512) lands at 0 so wp == oldest == 512; D must be refused
    out = script(1024, [PUSH(512), PUSH(512), POP, PUSH(512), PUSH(256)])
    assert out == [0, 512, 0, 0, -1]
    # ... and fits again once B is gone (live = C at [0,512), free = [512, 1024))
    out = script(1024, [PUSH(512), PUSH(512), POP, PUSH(512), PUSH(256), POP, PUSH(256), PUSH(256), PUSH(1)])
    assert out == [0, 512, 0, 0, -1, 0, 512, 768, -1]


def test_exact_fill_without_wrap(script):
    out = script(1024, [PUSH(1024), PUSH(1), POP, PUSH(1024)])
    assert out == [0, -1, 0, 0]
    out = script(1024, [PUSH(512), PUSH(512), PUSH(1), POP, PUSH(512), PUSH(1)])
    assert out == [0, 512, -1, 0, 0, -1]


def test_zero_byte_segments_do_not_make_the_ring_full(script):
    # tasks with empty payloads: bytes == 0 segments are live but occupy nothing
    out = script(1024, [PUSH(0), PUSH(0), PUSH(1024), PUSH(0), PUSH(1), POP, POP, POP, PUSH(1024)])
    assert out == [0, 0, 0, 1024, -1, 0, 0, 0, 0]


def test_too_large(script):
    assert script(1024, [PUSH(1025)]) == [-1]
    assert script(1024, [PUSH(1024)]) == [0]


@pytest.mark.parametrize("seed", range(8))
def test_random_scripts_never_overlap(script, seed):
    rng = np.random.default_rng(seed)
    ring = int(rng.choice([1024, 2048, 4096]))
    ops = []
    for _ in range(4000):
        r = rng.random()
        if r < 0.45:
            ops.append(POP)
        else:
            # sizes that are multiples of the alignment (exact fills are easy to hit), a few odd ones, a few zeros
            k = rng.random()
            if k < 0.6:
                b = 256 * int(rng.integers(1, ring // 256 // 2 + 1))
            elif k < 0.9:
                b = int(rng.integers(1, ring // 2))
            else:
                b = 0
            ops.append(PUSH(b))
    out = script(ring, ops)
    # the ring is used: most pushes land, some are refused
    placed = sum(1 for o, x in zip(ops, out) if o != POP and x >= 0)
    refused = sum(1 for o, x in zip(ops, out) if o != POP and x == -1)
    assert placed > 500 and refused > 0
