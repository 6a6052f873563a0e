"""The micro-batching buffer behind b9_submit / b9_flush (beta9_b200/csrc/submit_buffer.h), compiled for the host: many
producer threads, one flusher, small arenas (so that they fill and switch constantly). The reference's counterpart is a
mutex-guarded RingBuffer per endpoint (pkg/abstractions/endpoint/buffer.go:139-195); the contract checked here is the
one a queue must keep: every submitted task comes out exactly once, whole, and in submission order per producer."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_shim", "submit_buffer_shim.cpp")
SO = os.path.join(HERE, "host_shim", "libsubmitbuffer.so")
DEP = os.path.join(os.path.dirname(HERE), "beta9_b200", "csrc", "submit_buffer.h")


@pytest.fixture(scope="module")
def run():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(DEP)):
        r = subprocess.run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", SO, SRC], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lib = C.CDLL(SO)
    lib.b9_submit_buffer_run.restype = C.c_long
    lib.b9_submit_buffer_run.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_ulonglong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib.b9_submit_buffer_run


@pytest.mark.parametrize("producers,per,cap_tasks,cap_bytes,us", [(8, 4000, 64, 16384, 50), (16, 2000, 1000, 1 << 20, 200), (3, 5000, 7, 2100, 20), (1, 3000, 16, 4096, 10)])
def test_nothing_lost_duplicated_or_torn(run, producers, per, cap_tasks, cap_bytes, us):
    total = producers * per
    ids = np.zeros(total * 16, np.uint8); lens = np.zeros(total, np.uint32); ok = np.zeros(total, np.uint8)
    n = run(producers, per, cap_tasks, cap_bytes, us, ids.ctypes.data, lens.ctypes.data, ok.ctypes.data)
    assert n == total
    assert ok.all()                                              # every payload byte and flag as submitted
    w = ids.reshape(-1, 16)[:, :8].copy().view(np.uint32).reshape(-1, 2)
    p, t = w[:, 0].astype(np.int64), w[:, 1].astype(np.int64)
    assert np.array_equal(lens, ((p * 37 + t * 13) % 700).astype(np.uint32))
    # exactly once, and FIFO per producer
    assert np.unique(p * per + t).size == total
    for k in range(producers):
        assert np.array_equal(t[p == k], np.arange(per))
