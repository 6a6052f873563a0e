"""GPU tests of the host side of the C ABI: the pack step (b9_batch_push_v), producers and the drainer
running concurrently on one context (Dispatcher.Register's goroutine-safety contract,
pkg/task/dispatch.go:71-73), the claimed-task counter (pkg/repository/task_redis.go:58 TasksClaimed) and
the per-task duration of the completion record (taskqueue.proto:50)."""
import threading
import time

import numpy as np
import pytest

from beta9_b200 import synth
from oracle import coracle

pytestmark = pytest.mark.gpu


@pytest.fixture()
def dq():
    from beta9_b200.device_queue import DeviceQueue
    q = DeviceQueue(ring_bytes=1 << 30, ring_tasks=1 << 21, max_drain_tasks=1 << 21, max_result_bytes=1 << 30)
    yield q
    q.close()


def scattered(batch, seed=3):
    """The batch's payloads laid out in a pageable blob in a random order: (blob, pointers, lengths)."""
    n = batch.n
    lens = np.diff(batch.offsets).astype(np.uint32)
    perm = np.random.default_rng(seed).permutation(n)
    where = np.zeros(n, np.uint64)
    where[perm] = np.concatenate([[0], np.cumsum(lens[perm].astype(np.uint64))[:-1]])
    blob = np.empty(int(lens.sum()) + 1, np.uint8)
    for i in range(n):
        blob[int(where[i]):int(where[i]) + int(lens[i])] = batch.payload[int(batch.offsets[i]):int(batch.offsets[i + 1])]
    ptrs = (np.uint64(blob.ctypes.data) + where).astype(np.uint64)
    return blob, ptrs, lens


def test_push_scattered_equals_push_batch(dq):
    b = synth.concat([synth.strings_batch(6000, 64, adversarial_frac=0.1), synth.crc_batch(3000), synth.from_payloads([b"", b"{}", b'{"args": [""], "kwargs": {}}'])])
    blob, ptrs, lens = scattered(b)
    for rep in range(3):                                   # the two arenas alternate, the third push reuses the first
        dq.push_scattered(b.task_ids, ptrs, lens)
        assert dq.depth() == b.n
        r = dq.drain("identity", b.n)
        o = coracle.run_batch(b.task_ids, b.payload, b.offsets, "identity", nthreads=4)
        assert r.n == b.n and np.array_equal(r.task_ids, b.task_ids)
        assert np.array_equal(r.status, o.status) and np.array_equal(r.fifo_payload(), o.payload)
    del blob


def test_running_counter_and_task_duration(dq):
    b = synth.strings_batch(50_000, 64)
    dq.push_batch(b.task_ids, b.payload, b.offsets)
    assert dq.running() == 0 and dq.depth() == b.n
    dq.drain_launch("identity", 20_000, peek=True)
    assert dq.running() == 0                               # a peek claims nothing
    dq.drain_launch("identity", 20_000, wait=False)
    assert dq.running() == 20_000 and dq.depth() == b.n    # claimed, still counted as in flight
    r = dq.fetch()
    assert r.n == 20_000 and dq.running() == 0 and dq.depth() == b.n - 20_000
    assert 0.0 < r.task_duration < 1e-3
    r = dq.drain("identity")
    assert r.n == 30_000 and r.task_duration > 0.0


def test_expire_between_async_launch_and_fetch(dq):
    b = synth.strings_batch(40_000, 64)
    exp = np.zeros(b.n, np.int64); exp[::7] = 1000           # every 7th task expires at t = 1000 ns
    dq.push_batch(b.task_ids, b.payload, b.offsets, expires_unix_ns=exp)
    dq.drain_launch("identity", 10_000, wait=False)            # the window's tasks are claimed as they are
    n_exp = dq.expire(2000)                                     # ... and only then does the sweep run
    assert n_exp == len(range(0, b.n, 7))
    r = dq.fetch()
    assert r.n == 10_000                                        # the launch saw them ready
    rest = dq.drain("identity")
    live = np.ones(b.n, bool); live[::7] = False
    want = np.flatnonzero(live[10_000:]) + 10_000
    assert rest.n == want.size and np.array_equal(rest.task_ids, b.task_ids[want])
    # nothing cancelled is pending any more: the next window needs no count pre-pass (1 launch per drain)
    dq.push_batch(b.task_ids[:1000], b.payload[:int(b.offsets[1000])], b.offsets[:1001])
    k0 = dq.stats().kernel_launches
    assert dq.drain("identity").n == 1000
    assert dq.stats().kernel_launches - k0 == 1


def test_pusher_and_drainer_threads_run_concurrently(dq):
    """One thread pushes batches, another drains: every task comes out exactly once, in FIFO order, with the right
    bytes; and the two sides do not get in each other's way (the pipeline is not slower than the two sides back to back)."""
    b = synth.strings_batch(400_000, 256)
    o = coracle.run_batch(b.task_ids, b.payload, b.offsets, "identity", nthreads=8)
    want_lens = np.diff(o.offsets)
    rounds = 6
    pins = [(dq.pinned(b.n * 16), dq.pinned(b.payload.size), dq.pinned((b.n + 1) * 8)) for _ in range(2)]
    for pi, pp, po in pins:
        pi.array[:] = b.task_ids.reshape(-1); pp.array[:] = b.payload; po.view(np.uint64, b.n + 1)[:] = b.offsets

    def push(k):
        pi, pp, po = pins[k & 1]
        dq.push_batch(pi.array, pp.array, po.view(np.uint64, b.n + 1))       # synchronous: returns when the DMA is done

    def drain_one():
        r = dq.drain("identity", b.n)
        assert r.n == b.n and np.array_equal(r.task_ids, b.task_ids) and np.array_equal(r.lengths, want_lens)
        return r

    def serial_run():
        t0 = time.perf_counter()
        for k in range(rounds):
            push(k); drain_one()
        return time.perf_counter() - t0

    def piped_run():
        errors = []
        last = []
        pushed = threading.Semaphore(0)
        room = threading.Semaphore(2)              # at most two batches in the ring

        def producer():
            try:
                for k in range(rounds):
                    room.acquire()
                    push(k)
                    pushed.release()
            except Exception as e:                 # noqa: BLE001
                errors.append(e); pushed.release()

        def consumer():
            try:
                for k in range(rounds):
                    pushed.acquire()
                    if errors:
                        return
                    last.append(drain_one())
                    del last[:-1]
                    room.release()
            except Exception as e:                 # noqa: BLE001
                errors.append(e); room.release()

        t0 = time.perf_counter()
        tp, tc = threading.Thread(target=producer), threading.Thread(target=consumer)
        tp.start(); tc.start(); tp.join(120); tc.join(120)
        piped = time.perf_counter() - t0
        assert not errors, errors
        assert not tp.is_alive() and not tc.is_alive()
        assert dq.depth() == 0
        assert np.array_equal(last[-1].fifo_payload(), o.payload)      # (outside the timed region: it costs more than the run)
        return piped

    push(0); drain_one()
    # a shared box's host threads are noisy: the best of three attempts each
    serial = min(serial_run() for _ in range(3))
    piped = min(piped_run() for _ in range(3))
    print(f"serial {serial * 1e3:.1f} ms, two threads {piped * 1e3:.1f} ms")
    # What is asserted is that the two sides run beside each other WITHOUT getting in each other's way. How much shorter the
    # pipeline gets is printed, not asserted: most of a round here is this test's own Python (building and comparing the
    # results under the GIL), which no library can overlap — the overlap itself is measured by bench.py's e2e leg
    # (0.9+ of the link's two-way ceiling with one thread per side).
    assert piped < 1.25 * serial, (piped, serial)
    for t in pins:
        for p in t:
            p.free()


def test_submit_from_many_threads_then_flush(dq, monkeypatch):
    """b9_submit: one task per call from 8 threads (the reference's call pattern: one goroutine per put / per endpoint
    request), micro-batched into page-locked arenas; small arenas so that they fill, flush themselves and alternate."""
    import os
    os.environ["B9_SUBMIT_TASKS"] = "1000"; os.environ["B9_SUBMIT_BYTES"] = str(200_000)
    b = synth.strings_batch(24_000, 64, adversarial_frac=0.1)
    per = b.n // 8
    errors = []

    def producer(k):
        try:
            for i in range(k * per, (k + 1) * per):
                dq.submit(b.task_ids[i].tobytes(), b.task(i))
        except Exception as e:                 # noqa: BLE001
            errors.append(e)
    ths = [threading.Thread(target=producer, args=(k,)) for k in range(8)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(120)
    assert not errors, errors
    dq.flush()
    assert dq.buffered() == 0 and dq.depth() == b.n
    r = dq.drain("identity")
    o = coracle.run_batch(b.task_ids, b.payload, b.offsets, "identity", nthreads=4)
    assert r.n == b.n
    # arrival order is the order of the reservations: every producer's tasks in ITS order, every task exactly once, right bytes
    index = {b.task_ids[i].tobytes(): i for i in range(b.n)}
    got = [index[r.task_ids[j].tobytes()] for j in range(r.n)]
    assert sorted(got) == list(range(b.n))
    for k in range(8):
        mine = [i for i in got if k * per <= i < (k + 1) * per]
        assert mine == sorted(mine)
    for j in range(0, r.n, 37):
        i = got[j]
        assert r.status[j] == o.status[i] and r.result(j) == o.result(i)
    del os.environ["B9_SUBMIT_TASKS"], os.environ["B9_SUBMIT_BYTES"]
