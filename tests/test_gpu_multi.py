"""Multi-GPU: NCCL byte-quantile rebalance of the pending ring, then the ordinary drain on every rank.
Needs >= 2 GPUs on the box (gpurun --gpus 2); skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

from beta9_b200 import synth
from oracle import coracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    from beta9_b200 import _lib as L
    return L.load().b9_device_count()


def run_world(world, tmp_path, handler, mode):
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gpu_multi_worker.py"), str(r), str(world), str(tmp_path), handler, mode],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    return [np.load(os.path.join(tmp_path, f"gpu_rank{r}.npz")) for r in range(world)]


def records(r):
    off = np.concatenate([[0], np.cumsum(r["lens"].astype(np.int64))])
    return [(r["ids"][i].tobytes(), (int(r["status"][i]), r["payload"][off[i]:off[i + 1]].tobytes() if r["has"][i] else None)) for i in range(len(r["status"]))]


@pytest.mark.parametrize("handler,mode", [("identity", "one_push"), ("crc32", "one_push"), ("identity", "many_pushes"), ("identity", "twice"), ("identity", "enospc")])
def test_rebalance_then_drain_matches_oracle(handler, mode, tmp_path):
    world = min(n_gpus(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    res = run_world(world, tmp_path, handler, mode)
    # expected: the oracle over the union of all pushed, non-cancelled tasks
    want = {}
    for r in res:
        o = coracle.run_batch(r["before_ids"], r["before_payload"], r["before_offsets"], handler)
        for i in range(len(r["flags"])):
            if not r["flags"][i]:
                want[r["before_ids"][i].tobytes()] = (int(o.status[i]), o.result(i))
    got = []
    for k, r in enumerate(res):
        got += records(r)
        if mode == "twice":
            got += records(np.load(os.path.join(tmp_path, f"gpu_rank{k}_half.npz")))
    if mode == "twice":                         # every task was pushed twice: it comes out exactly twice, with the same record
        assert len(got) == 2 * len(want)
        from collections import Counter
        cnt = Counter(k for k, _ in got)
        assert set(cnt.values()) == {2} and dict(got) == want
        return
    assert len(got) == len(want) and dict(got) == want
    if mode == "enospc":
        return
    # balance: pending payload bytes per rank within one task of the ideal share
    after = np.array([int(r["info"][7]) for r in res]); total = int(sum(int(r["info"][1]) for r in res))
    assert int(after.sum()) == total
    assert np.all(np.abs(after - total / world) <= 4200), after
    assert int(res[0]["info"][2]) > 0          # rank 0 was heavy: it sent something
