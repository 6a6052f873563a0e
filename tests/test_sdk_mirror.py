"""The host-side mirror of the SDK surface, written after the reference's own tests:
sdk/tests/test_task_queue.py:10-90 (init, local, put ok / put refused, direct call) and
sdk/tests/test_function.py:60-88 (.map yields one result per input). The device queue is mocked
here (CPU); tests/test_gpu_sdk.py runs the same surface on a B200."""
import json
import os
from unittest import mock
from unittest.mock import MagicMock

import numpy as np
import pytest

from beta9_b200.taskqueue import Task, TaskQueue, task_queue


def fake_queue(depth=0):
    q = MagicMock()
    q.depth.return_value = depth
    return q


def test_init():
    tq = TaskQueue(cpu=1, memory=128, image="python3.8", queue=fake_queue())
    assert tq.cpu == 1000 and tq.memory == 128 and tq.image == "python3.8"     # test_task_queue.py:16-19
    assert tq.max_pending_tasks == 100


def test_run_local():
    @TaskQueue(cpu=1, memory=128, queue=fake_queue())
    def test_func():
        return 1
    assert test_func.local() == 1                                               # test_task_queue.py:21-28


def test_put_ok_and_refused():
    q = fake_queue()

    @TaskQueue(cpu=1, memory=128, queue=q)
    def test_func():
        return 1
    t = test_func.put()
    assert isinstance(t, Task) and len(t.id) == 36
    ids, blob, offsets = q.push_batch.call_args[0]
    assert bytes(blob) == b'{"args": [], "kwargs": {}}' and list(offsets) == [0, 26] and ids.shape == (1, 16)
    t = test_func.put("x", k=1)
    assert bytes(q.push_batch.call_args[0][1]) == json.dumps({"args": ("x",), "kwargs": {"k": 1}}).encode()
    # admission limit (taskqueue.go:182-189): put answers False, nothing is pushed
    q2 = fake_queue(depth=100)

    @TaskQueue(queue=q2)
    def f2():
        return 1
    assert f2.put() is False and not q2.push_batch.called
    # runtime not prepared -> False (taskqueue.py:278-282)
    with mock.patch.object(TaskQueue, "prepare_runtime", return_value=False):
        assert test_func.put() is False


def test_direct_call():
    @TaskQueue(queue=fake_queue())
    def test_func():
        return 1
    os.environ.pop("CONTAINER_ID", None)
    with pytest.raises(NotImplementedError):
        test_func()
    with mock.patch.dict(os.environ, {"CONTAINER_ID": "1234"}):
        assert test_func() == 1                                                 # test_task_queue.py:86-88


def test_map_yields_one_result_per_input():
    q = fake_queue()
    pushed = {}

    def push(ids, blob, offsets, **kw):
        pushed["ids"] = ids.copy()
    q.push_batch.side_effect = push

    def drain(handler, max_tasks):
        ids = pushed["ids"]
        r = MagicMock()
        r.n = ids.shape[0]
        r.task_ids = ids
        r.status = np.zeros(r.n, np.uint8)
        r.result = lambda i: b"1998"
        return r
    q.drain.side_effect = drain

    @task_queue(queue=q, max_pending_tasks=10)
    def test_func(x):
        return 1998
    assert list(test_func.map([1, 2, 3])) == [1998, 1998, 1998]                # test_function.py:84-88
    assert q.push_batch.call_count == 1 and q.drain.call_count == 1


def test_map_formats_inputs_like_function_format_args():
    """function.py:246-251 `_format_args`: a tuple or a list IS the positional argument list, anything else is one
    argument; `.map()` never produces keyword arguments — not even for an (args-looking, dict) pair."""
    q = fake_queue()
    seen = {}

    def push(ids, blob, offsets, **kw):
        off = np.asarray(offsets, dtype=np.int64)
        raw = np.asarray(blob).tobytes()
        seen["payloads"] = [json.loads(raw[off[i]:off[i + 1]]) for i in range(off.size - 1)]
        seen["ids"] = ids.copy()
    q.push_batch.side_effect = push

    def drain(handler, max_tasks):
        r = MagicMock()
        r.n = seen["ids"].shape[0]; r.task_ids = seen["ids"]; r.status = np.zeros(r.n, np.uint8); r.result = lambda i: b"0"
        return r
    q.drain.side_effect = drain

    @task_queue(queue=q, max_pending_tasks=100)
    def f(*a):
        return 0
    inputs = [5, "s", (1, 2), [3, 4], ([1], {"k": 2}), {"d": 1}, None, (), []]
    list(f.map(inputs))
    assert seen["payloads"] == [
        {"args": [5], "kwargs": {}}, {"args": ["s"], "kwargs": {}}, {"args": [1, 2], "kwargs": {}}, {"args": [3, 4], "kwargs": {}},
        {"args": [[1], {"k": 2}], "kwargs": {}},            # two positional arguments, as the reference would pass them
        {"args": [{"d": 1}], "kwargs": {}}, {"args": [None], "kwargs": {}}, {"args": [], "kwargs": {}}, {"args": [], "kwargs": {}}]


def test_map_keeps_the_results_of_tasks_queued_before_it():
    """the queue is FIFO, so a map() behind pending put()s drains those first. Their results are kept for the
    next process_tasks(), and map() goes on draining until its own tasks have come back."""
    q = fake_queue()
    fifo = []                                   # (id bytes) in push order

    def push(ids, blob, offsets, **kw):
        fifo.extend(bytes(r) for r in np.asarray(ids).reshape(-1, 16))
    q.push_batch.side_effect = push

    def drain(handler, max_tasks):
        take = fifo[:2]                         # a small drain window: map() has to come back for more
        del fifo[:2]
        r = MagicMock()
        r.n = len(take)
        r.task_ids = np.frombuffer(b"".join(take), np.uint8).reshape(-1, 16) if take else np.zeros((0, 16), np.uint8)
        r.status = np.zeros(r.n, np.uint8)
        r.result = lambda i: json.dumps(take[i][:2].hex()).encode()
        return r
    q.drain.side_effect = drain

    @task_queue(queue=q, max_pending_tasks=100)
    def f(x):
        return x
    early = [f.put(i) for i in range(3)]
    out = list(f.map([10, 11, 12]))
    assert len(out) == 3 and all(o is not None for o in out)
    held = f.process_tasks()
    assert [t.id for t in held] == [t.id for t in early]           # nothing was dropped, FIFO order kept
    assert all(t.status == "COMPLETE" for t in held)
