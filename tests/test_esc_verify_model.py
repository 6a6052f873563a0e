"""The rule set of the device's cooperative "canonical escapes" check (tests/esc_verify_model.py mirrors
beta9_b200/csrc/drain2.cuh esc_verify_canonical) against the oracle: whenever the check says "the result
is the token itself (lone surrogates rewritten)", the reference loop — Go decode, Python json.loads,
identity, json.dumps (oracle/pyoracle/loop.py) — must produce exactly those bytes; and everything
json.dumps itself writes must be accepted (that is the 1 % share of BASELINE configs[1])."""
import json
import random

import pytest

from oracle.pyoracle.loop import COMPLETE, run_task_loop
from tests.esc_verify_model import esc16, sequential_escaped, verify

PRE = b'{"args": ["'
SUF = b'"], "kwargs": {}}'
TID = bytes(16)


def oracle_identity(body: bytes):
    r = run_task_loop([PRE + body + SUF], [TID], "identity")[0]
    return r.status, r.result


def test_esc16_is_the_sequential_definition():
    for carry in (0, 1):
        for bs in range(1 << 16):
            pos = [bool((bs >> j) & 1) for j in range(16)]
            # reference: a virtual escape start before position 0 when carry is set
            seq = sequential_escaped(([True] if carry else [False]) + pos + [False])
            want = sum(1 << j for j in range(16) if seq[j + 1])
            want_out = 1 if seq[17] else 0
            if carry and False:
                pass
            got, out = esc16(bs, carry)
            assert (got, out) == (want, want_out), (hex(bs), carry)


POOL = ['"', "\\", "<", ">", "&", "/", "\n", "\r", "\t", "\b", "\f", "\x01", "\x1f", "\x7f", "é", "ß", " ",
        " ", "�", "€", "\U0001f600", "\U00010348", "\ud83d", "\udc00", "\udbff", "\udfff"] + [chr(c) for c in range(0x20, 0x7F)]


def rand_string(rng, n, dense):
    if dense:
        return "".join(rng.choice(POOL[:26]) for _ in range(n))
    return "".join(rng.choice(POOL) for _ in range(n))


@pytest.mark.parametrize("seed", range(6))
def test_everything_json_dumps_writes_is_accepted_and_right(seed):
    rng = random.Random(seed)
    for _ in range(250):
        s = rand_string(rng, rng.choice([1, 5, 40, 256, 300, 700]), rng.random() < 0.3)
        body = json.dumps(s).encode()[1:-1]
        got = verify(body)
        assert got is not None, body
        st, res = oracle_identity(body)
        assert st == COMPLETE and res == b'"' + got + b'"', body


MUT = [b"\\", b'"', b"\\/", b"\\u0041", b"\\u00E9", b"\\u000a", b"\\ud83d", b"\\udc00", b"\\uD83D", b"\\u12", b"\\x", b"\xc3\xa9", b"\x7f",
       b"\x01", b"\\\\", b"\\\\\\", b"\\u007f", b"\\u0022", b"\\n", b"u", b"\\ud83d\\udc00", b"\\udc00\\ud83d", b"\xff", b"\\u005c"]


@pytest.mark.parametrize("seed", range(6))
def test_mutations_never_give_a_wrong_answer(seed):
    rng = random.Random(100 + seed)
    accepted = 0
    for _ in range(500):
        s = rand_string(rng, rng.choice([0, 3, 30, 120, 600]), rng.random() < 0.5)
        body = bytearray(json.dumps(s).encode()[1:-1])
        for _k in range(rng.randint(1, 4)):
            pos = rng.randint(0, len(body))
            if rng.random() < 0.3 and len(body):
                del body[pos:pos + rng.randint(1, 3)]
            else:
                body[pos:pos] = rng.choice(MUT)
        body = bytes(body)
        got = verify(body)
        if got is None:
            continue
        accepted += 1
        st, res = oracle_identity(body)
        if len(body) == 0:
            assert st == COMPLETE and res is None          # "" is falsy: the caller (not the check) knows
            continue
        assert st == COMPLETE and res == b'"' + got + b'"', body
    assert accepted > 20


def test_chunk_and_pass_boundaries():
    # escapes and backslash runs straddling the 16-byte lane chunks and the 512-byte passes
    for pad in list(range(0, 40)) + [500, 505, 509, 510, 511, 512, 1020, 1023]:
        for piece in (b"\\\\", b"\\\"", b"\\u00e9", b"\\ud83d\\udc00", b"\\ud83d", b"\\udc00", b"\\\\\\\\\\\\", b"\\\\" * 20, b"\\n"):
            body = b"a" * pad + piece + b"zz"
            got = verify(body)
            assert got is not None, body
            st, res = oracle_identity(body)
            assert st == COMPLETE and res == b'"' + got + b'"', (pad, piece)
        # an odd run at the very end escapes the frame's quote: not decided here
        assert verify(b"a" * pad + b"\\") is None
        assert verify(b"a" * pad + b"\\\\\\") is None
