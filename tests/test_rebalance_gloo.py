"""world_size > 1 on CPU (gloo): the rebalance plan the NCCL path uses (b9_rebalance_plan, through the
C ABI) and the exchange protocol around it, with the data movement emulated by gloo point-to-point."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from beta9_b200.device_queue import rebalance_plan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def launch(world, skew, tmp_path):
    port = free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_rebalance_worker.py"), str(r), str(world), str(port), skew, str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    return [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]


@pytest.mark.parametrize("world,skew", [(2, "rank0_heavy"), (2, "last_empty"), (4, "rank0_heavy"), (3, "zipf")])
def test_rebalance_exchange_gloo(world, skew, tmp_path):
    res = launch(world, skew, tmp_path)
    # nothing lost, nothing duplicated: the multiset of (id, payload) is preserved
    def records(ids, lens, payload):
        off = np.concatenate([[0], np.cumsum(lens)])
        return sorted((ids[i].tobytes(), payload[off[i]:off[i + 1]].tobytes()) for i in range(len(lens)))
    before = sorted(sum((records(r["before_ids"], r["before_lens"], r["before_payload"]) for r in res), []))
    after = sorted(sum((records(r["ids"], r["lens"], r["payload"]) for r in res), []))
    assert before == after
    # byte balance: every rank within one maximal task of the ideal share
    total = sum(int(r["lens"].sum()) for r in res)
    biggest = max(int(r["before_lens"].max()) if len(r["before_lens"]) else 0 for r in res)
    for r in res:
        assert abs(int(r["lens"].sum()) - total / world) <= biggest + 1, (int(r["lens"].sum()), total / world, biggest)


def test_plan_properties_single_process():
    rng = np.random.default_rng(7)
    for world in (1, 2, 3, 8):
        lens = [rng.integers(0, 5000, size=int(rng.integers(0, 400))) for _ in range(world)]
        counts = np.array([len(l) for l in lens], np.uint64)
        nbytes = np.array([int(l.sum()) for l in lens], np.uint64)
        total = int(nbytes.sum())
        got = np.zeros(world, np.int64)
        for r in range(world):
            prefix = np.concatenate([[0], np.cumsum(lens[r])]).astype(np.uint64)
            lo, hi = rebalance_plan(world, r, counts, nbytes, prefix)
            assert int(lo[0]) == 0 and int(hi[-1]) == len(lens[r])
            assert np.all(hi[:-1] == lo[1:]) and np.all(lo <= hi)
            for d in range(world):
                got[d] += int(prefix[int(hi[d])] - prefix[int(lo[d])])
        assert int(got.sum()) == total
        if total:
            assert np.all(np.abs(got - total / world) <= 5000 + 1)


def test_plan_rejects_inconsistent_table():
    from beta9_b200 import _lib as L
    with pytest.raises(L.B9Error):
        rebalance_plan(2, 0, np.array([3, 1], np.uint64), np.array([10, 5], np.uint64), np.array([0, 5, 9], np.uint64))
