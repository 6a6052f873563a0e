"""The SDK mirror on a real device: put / put_batch / process_tasks / map against the Python oracle."""
import uuid

import pytest

from oracle.pyoracle import loop

pytestmark = pytest.mark.gpu


def test_put_process_and_map_on_device():
    from beta9_b200.taskqueue import TaskQueue

    @TaskQueue(gpu_handler="identity", max_pending_tasks=1000)
    def echo(s):
        return s
    t1 = echo.put("hello")
    t2 = echo.put("wörld \"q\" 😀")
    t3 = echo.put("")
    assert echo.parent.queue.depth() == 3
    done = echo.process_tasks()
    assert [d.id for d in done] == [t1.id, t2.id, t3.id]                       # FIFO
    want = loop.run_task_loop([loop.sdk_put_payload("hello"), loop.sdk_put_payload("wörld \"q\" 😀"), loop.sdk_put_payload("")],
                              [uuid.UUID(t.id).bytes for t in (t1, t2, t3)], "identity")
    assert [(d.status, d.result_bytes) for d in done] == [(w.status, w.result) for w in want]
    assert done[0].result == echo.local("hello") and done[2].result is None
    inputs = [f"item-{i}" for i in range(500)]
    assert list(echo.map(inputs)) == inputs
    # admission: the 1001st pending task is refused
    assert len(echo.put_batch(inputs + inputs)) == 1000
    assert echo.put("one more") is False
    assert len(echo.process_tasks()) == 1000

    @TaskQueue(gpu_handler="crc32", queue=echo.parent.queue)
    def crc(s):
        import zlib
        return zlib.crc32(s.encode())
    assert list(crc.map(inputs[:50])) == [crc.local(s) for s in inputs[:50]]
    echo.parent.queue.close()
