"""Both oracles against the committed golden fixtures (tests/golden/hot_path_golden.json)."""
import base64

import pytest

from oracle import coracle
from oracle.pyoracle import loop
from tests import golden_util as G

HANDLERS = ["identity", "crc32", "vadd_f32", "json_sum"]


@pytest.mark.parametrize("handler", HANDLERS)
def test_python_oracle_matches_golden(handler):
    g = G.load()
    for name, cases in g["groups"].items():
        b = G.group_batch(cases)
        res = loop.run_task_loop(b.tasks(), [bytes(x) for x in b.task_ids], handler, keep_wire=True,
                                 now_unix_ns=g["now_unix_ns"])
        for c, r in zip(cases, res):
            st, out = G.expected(c, handler)
            assert (G.STATUS_CODE[r.status], r.result) == (st, out), (name, c["payload"])
            assert (None if r.wire is None else base64.b64encode(r.wire).decode()) == c["wire"]


@pytest.mark.parametrize("handler", HANDLERS)
def test_c_oracle_matches_golden(handler):
    g = G.load()
    for name, cases in g["groups"].items():
        b = G.group_batch(cases)
        r = coracle.run_batch(b.task_ids, b.payload, b.offsets, handler, keep_wire=True, now_ns=g["now_unix_ns"])
        for i, c in enumerate(cases):
            if r.status[i] == coracle.UNSUPPORTED:
                continue
            st, out = G.expected(c, handler)
            assert (int(r.status[i]), r.result(i)) == (st, out), (name, c["payload"])
            assert r.wire_msg(i) == (base64.b64decode(c["wire"]) if c["wire"] else b"")
