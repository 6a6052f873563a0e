"""Result sink: the packed object of a drain against the reference's per-task path as restated in
oracle/pyoracle/resultsink.py — StoreTaskResult (pkg/task/dispatch.go:120-144, one PUT under "task/<id>/result" when
`in.Result != nil`) and addResultToTask (pkg/api/v1/task.go:295-325, JSON kept / anything else {"base64":...}).
The sink functions of the C ABI are pure host code: these tests need no GPU (tests/test_gpu_sink.py covers
b9_drain_fetch_object on the device)."""
import ctypes as C
import json
import random

import numpy as np
import pytest

from beta9_b200 import _lib as L
from beta9_b200 import synth
from oracle import coracle
from oracle.pyoracle import resultsink
from oracle.pyoracle.wire import format_uuid


def result_json(lib, data: bytes):
    buf = C.create_string_buffer(4 * len(data) // 3 + 64 + len(data))
    n = lib.b9_sink_result_json(data, len(data), buf, len(buf))
    assert n >= 0
    return None if n == 0 else buf.raw[:n]


CASES = [b"", b'"abc"', b'  {"a": 1e999}\n', b"\xff\x00", b"nul", b"null", b"true ", b"\tfalse", b"123 4", b"[1,2", b"[1,2]", b"[1,]", b"{}", b"{,}",
         b'{"a":1,}', b'{"a" : [ ] , "b":{ } }', b'"\\ud800"', b'"\\x"', b'"a\nb"', b"-", b"-0", b"01", b"1.", b"1.5e+3", b"1e", b".5", b"[[[[[[]]]]]]",
         b"[" * 10000 + b"]" * 10000, b"[" * 10001 + b"]" * 10001, b'{"a":"\xc3\xa9\xff"}', b" ", b"\n\n", b"0 ", b"[1 2]", b'{"a" 1}', b'{"a":}', b'{1:2}',
         b'"unterminated', b"tru", b"nulll", b'[1,2]x', b'\xef\xbb\xbf[1]']


def test_add_result_to_task_rule_known_cases():
    lib = L.load()
    for c in CASES:
        assert result_json(lib, c) == resultsink.add_result_to_task(c), c[:60]


@pytest.mark.parametrize("seed", range(4))
def test_add_result_to_task_rule_fuzz(seed):
    lib = L.load()
    rng = random.Random(seed)

    def rand_value(d=0):
        k = rng.random()
        if d > 3 or k < 0.3:
            return rng.choice([None, True, False, 0, -1, 3.5, 1e300, "", "a\"b\\c", "é ", "\ud83d", 12345678901234567890])
        if k < 0.65:
            return [rand_value(d + 1) for _ in range(rng.randint(0, 4))]
        return {rng.choice(["a", "b", "values", "k\n"]): rand_value(d + 1) for _ in range(rng.randint(0, 4))}

    for _ in range(1500):
        b = bytearray(json.dumps(rand_value(), ensure_ascii=rng.random() < 0.5, separators=rng.choice([(",", ":"), (", ", ": ")])).encode("utf-8", "surrogatepass"))
        if rng.random() < 0.6:
            for _k in range(rng.randint(1, 3)):
                pos = rng.randint(0, len(b))
                r = rng.random()
                if r < 0.3 and len(b):
                    del b[pos:pos + rng.randint(1, 2)]
                elif r < 0.6:
                    b[pos:pos] = rng.choice([b",", b"]", b"}", b'"', b"\\", b" ", b"\n", b"1", b"e", b".", b"-", b"\x00", b"\xff", b"[", b"{", b":"])
                else:
                    b[pos:pos] = rng.choice([b"  ", b"\t", b"null", b"1e5", b"\\u12G4"])
        b = bytes(b)
        assert result_json(lib, b) == resultsink.add_result_to_task(b), b


def make_results(handler="identity"):
    """Records of one 'drain' built from the oracle (ids, status, has, offsets, lengths, blob), FIFO-dense."""
    b = synth.concat([synth.strings_batch(300, 40, adversarial_frac=0.3), synth.from_payloads([b"", b"{}", b'{"args": [""], "kwargs": {}}', b'{"args": [0]}', b"not json"])])
    b.task_ids = synth.task_ids(b.n, seed=7)                 # (concat keeps the parts' ids: make them distinct)
    o = coracle.run_batch(b.task_ids, b.payload, b.offsets, handler)
    return b, o


def test_pack_get_find_match_the_per_task_store():
    lib = L.load()
    b, o = make_results()
    n = b.n
    lens = np.diff(o.offsets).astype(np.uint32)
    offs = np.ascontiguousarray(o.offsets[:-1], np.uint64)
    ids = np.ascontiguousarray(b.task_ids); status = np.ascontiguousarray(o.status); has = np.ascontiguousarray(o.has)
    blob = np.ascontiguousarray(o.payload)
    r = L.Results(ids.ctypes.data, status.ctypes.data, has.ctypes.data, offs.ctypes.data, lens.ctypes.data, blob.ctypes.data if blob.size else ids.ctypes.data,
                  n, int(blob.size), n, n, int(blob.size), 0, 1.5e-6, 0)
    need = lib.b9_sink_object_bytes(n, int(blob.size))
    obj = np.zeros(need, np.uint8)
    assert lib.b9_sink_pack(C.byref(r), obj.ctypes.data, need - 1) == L.B9_ENOSPC
    assert lib.b9_sink_pack(C.byref(r), obj.ctypes.data, need) == need
    # the reference: one PUT per task that produced result bytes
    store = {}
    for i in range(n):
        resultsink.store_task_result(store, b.task_ids[i].tobytes(), o.result(i))
    rec = L.SinkRecord()
    seen = 0
    for i in range(n):
        assert lib.b9_sink_get(obj.ctypes.data, need, i, C.byref(rec)) == 0
        tid = C.string_at(rec.task_id, 16)
        assert tid == b.task_ids[i].tobytes() and rec.status == o.status[i] and rec.index == i
        key = resultsink.task_result_path(format_uuid(tid))
        if rec.has_result:
            data = C.string_at(rec.data, rec.length)
            assert store[key] == data
            assert result_json(lib, data) == resultsink.add_result_to_task(store[key])
            seen += 1
        else:
            assert key not in store
    assert seen == len(store)
    # by id
    for i in (0, n // 2, n - 1):
        assert lib.b9_sink_find(obj.ctypes.data, need, b.task_ids[i].ctypes.data, C.byref(rec)) == 0 and rec.index == i
    assert lib.b9_sink_find(obj.ctypes.data, need, (C.c_uint8 * 16)(*([7] * 16)), C.byref(rec)) == L.B9_ENOENT
    assert lib.b9_sink_get(obj.ctypes.data, need, n, C.byref(rec)) == L.B9_EINVAL
    # a truncated or foreign buffer is refused
    assert lib.b9_sink_get(obj.ctypes.data, need - 1, 0, C.byref(rec)) == L.B9_EINVAL
    junk = np.zeros(256, np.uint8)
    assert lib.b9_sink_get(junk.ctypes.data, 256, 0, C.byref(rec)) == L.B9_EINVAL
