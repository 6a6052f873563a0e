"""Soundness of the device json_sum fast path's acceptance rules (tests/json_coop_model.py mirrors
them): every payload the rules decide must be decided the same way by the oracle."""
import json

import numpy as np

from beta9_b200 import synth
from oracle import coracle
from tests.json_coop_model import decide

STRUCT = list(b'{}[],:" 0123456789-.eEtfn\\ax')


def _canonical_docs(rng, k):
    docs = []
    for i in range(k):
        kind = i % 6
        vals = rng.integers(0, 10**int(rng.integers(1, 16)), size=int(rng.integers(0, 12))).tolist()
        if kind == 0: d = {"id": i, "values": vals, "pad": "x" * int(rng.integers(0, 40))}
        elif kind == 1: d = {"values": vals}
        elif kind == 2: d = {"values": vals, "other": [1, 2, 3], "s": "a,b:[c]{d}", "values ": [7]}
        elif kind == 3: d = {"a": "values", "values": vals, "n": 0, "e": [], "z": [0]}
        elif kind == 4: d = {"values": [1], "k": "v", "values2": [5]}
        else: d = {"x": {"values": [9]}, "values": vals} if i % 12 == 5 else {"values": vals, "t": "]", "u": "[1, 2"}
        docs.append(d)
    return docs


def _large_payloads(rng, k):
    """Documents of 1-4 KiB (several 1 KiB segments on the device) and single-character mutants of them."""
    out = []
    for i in range(k):
        nvals = int(rng.integers(20, 400))
        vals = rng.integers(0, 10**int(rng.integers(1, 16)), size=nvals).tolist()
        pad = "z" * int(rng.integers(0, 1500))
        order = i % 3
        if order == 0: d = {"pad": pad, "values": vals, "id": i}
        elif order == 1: d = {"values": vals, "pad": pad, "tail": [1, 2, 3]}
        else: d = {"a": [7] * int(rng.integers(0, 200)), "pad": pad, "values": vals}
        p = bytearray(b'{"args": [' + json.dumps(d).encode() + b'], "kwargs": {}}')
        if len(p) > 4200:
            continue
        out.append(bytes(p))
        for _ in range(4):
            m = bytearray(p)
            pos = int(rng.integers(10, len(m) - 16))
            op = int(rng.integers(0, 3))
            ch = int(rng.choice(STRUCT))
            if op == 0: m[pos] = ch
            elif op == 1: del m[pos]
            else: m[pos:pos] = bytes([ch])
            out.append(bytes(m))
    return out


def _payloads(rng, k):
    out = []
    for d in _canonical_docs(rng, k):
        seps = (", ", ": ") if rng.random() < 0.7 else (",", ":")
        p = bytearray(b'{"args": [' + json.dumps(d, separators=seps).encode() + b'], "kwargs": {}}')
        out.append(bytes(p))
        for _ in range(6):      # single-character mutants inside the document
            m = bytearray(p)
            pos = int(rng.integers(10, len(m) - 16))
            op = int(rng.integers(0, 3))
            ch = int(rng.choice(STRUCT))
            if op == 0: m[pos] = ch
            elif op == 1: del m[pos]
            else: m[pos:pos] = bytes([ch])
            out.append(bytes(m))
    # duplicate keys: the later "values" wins
    out.append(b'{"args": [{"values": [1, 2], "values": [40]}], "kwargs": {}}')
    out.append(b'{"args": [{"values": [1, 2], "values": 3}], "kwargs": {}}')
    out.append(b'{"args": [{"values": []}], "kwargs": {}}')
    out.append(b'{"args": [{"values": [0, 0]}], "kwargs": {}}')
    out.append(b'{"args": [{"values": [123456789012345, 999999999999999]}], "kwargs": {}}')
    out.append(b'{"args": [{"values": [1234567890123456]}], "kwargs": {}}')
    out.append(b'{"args": [{}], "kwargs": {}}')
    return out


def test_model_decisions_agree_with_oracle():
    rng = np.random.default_rng(0xB9)
    payloads = _payloads(rng, 1500) + _large_payloads(rng, 250)
    b = synth.from_payloads(payloads)
    o = coracle.run_batch(b.task_ids, b.payload, b.offsets, "json_sum", nthreads=8)
    decided = 0
    for i, p in enumerate(payloads):
        s = decide(p)
        if s is None:
            continue
        decided += 1
        assert o.status[i] == 0, (p, s, int(o.status[i]))
        want = None if s == 0 else str(s).encode()
        assert o.result(i) == want, (p, s, o.result(i))
    # the rules must decide the canonical documents (otherwise the fast path is useless) ...
    assert decided > len(payloads) // 8
    # ... and the unmutated benchmark shape in particular
    jb = synth.json_batch(50)
    for i in range(jb.n):
        assert decide(jb.payload[int(jb.offsets[i]):int(jb.offsets[i + 1])].tobytes()) is not None
