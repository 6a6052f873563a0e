"""Unit facts of the restated Go encoding/json rules (oracle/pyoracle/gojson.py, wire.py).
These document the rules the oracle encodes; the boundary itself is PARITY UNPINNED (no Go here)."""
import json
import math

import pytest

from oracle.pyoracle import gojson as gj
from oracle.pyoracle.loop import run_task_loop, sdk_put_payload
from oracle.pyoracle.wire import QueueEnv, TaskMessage, TaskPolicy, build_task_message, format_uuid


@pytest.mark.parametrize("f,s", [
    (1.0, "1"), (0.1, "0.1"), (100.0, "100"), (1e20, "100000000000000000000"), (1e21, "1e+21"),
    (1.5e-7, "1.5e-7"), (1e-6, "0.000001"), (9.99e-7, "9.99e-7"), (123456789.125, "123456789.125"),
    (-0.0, "-0"), (0.0, "0"), (5e-324, "5e-324"), (1.7976931348623157e308, "1.7976931348623157e+308"),
    (2.0 ** 60, "1152921504606847000"), (-2.5, "-2.5"), (1e-10, "1e-10"), (123e-20, "1.23e-18"),
    (9007199254740993.0, "9007199254740992"),
])
def test_go_float_format(f, s):
    assert gj.go_format_float64(f) == s


def test_go_float_roundtrip_random():
    import random
    rnd = random.Random(0xB9)
    for _ in range(5000):
        f = rnd.uniform(-1, 1) * 10 ** rnd.randint(-30, 30)
        s = gj.go_format_float64(f)
        assert float(s) == f
        assert " " not in s and "E" not in s


def test_go_quote_rules():
    assert gj.go_quote('a<b>&"\\') == '"a\\u003cb\\u003e\\u0026\\"\\\\"'
    assert gj.go_quote("\b\f\n\r\t\x01\x1f\x7f") == '"\\b\\f\\n\\r\\t\\u0001\\u001f\x7f"'
    assert gj.go_quote("\u2028\u2029\u00e9\U0001f600") == '"\\u2028\\u2029\u00e9\U0001f600"'
    assert gj.go_quote("/") == '"/"'


def test_go_unmarshal_strings():
    u = gj.go_unmarshal
    assert u(b'"\\ud83d\\ude00"') == "\U0001f600"
    assert u(b'"\\ud83d"') == "\ufffd"
    assert u(b'"\\ud83dx"') == "\ufffdx"
    assert u(b'"\\ud83d\\u0041"') == "\ufffdA"
    assert u(b'"\\ude00"') == "\ufffd"
    assert u(b'"\\ud83d\\ud83d\\ude00"') == "\ufffd\U0001f600"
    assert u(b'"\xff\xfe"') == "\ufffd\ufffd"
    assert u(b'"\xc3\xa9"') == "\u00e9"
    assert u(b'"\xc0\xaf"') == "\ufffd\ufffd"            # overlong
    assert u(b'"\xed\xa0\x80"') == "\ufffd\ufffd\ufffd"  # UTF-8 encoded surrogate
    assert u(b'"\xe2\x82"') == "\ufffd\ufffd"            # truncated
    assert u(b'"\\/\\b\\f\\n\\r\\t\\"\\\\"') == '/\b\f\n\r\t"\\'
    assert u(b'"\\u0000"') == "\x00"


@pytest.mark.parametrize("bad", [
    b"", b" ", b"{", b'{"a":1,}', b"[1,]", b"01", b"1.", b".5", b"+1", b"1e", b"nul", b"NaN",
    b"Infinity", b'"\x01"', b'"\\x"', b'"\\u12g4"', b'"abc', b"{'a':1}", b"[1] x", b'{"a" 1}',
    b"-", b"--1", b'"\\u12"', b"tru", b'{"a":1 "b":2}', b"\xef\xbb\xbf{}",
])
def test_go_unmarshal_rejects(bad):
    with pytest.raises(gj.GoJSONError):
        gj.go_unmarshal(bad)


def test_go_unmarshal_accepts_whitespace_and_numbers():
    assert gj.go_unmarshal(b' \t\r\n[1, -0, 1e3, 1E-2, 0.5, 10] \n') == [1.0, -0.0, 1000.0, 0.01, 0.5, 10.0]
    assert math.copysign(1, gj.go_unmarshal(b"-0")) == -1
    assert gj.go_unmarshal(b"1e-400") == 0.0           # underflow is not an error
    with pytest.raises(gj.GoJSONError):
        gj.go_unmarshal(b"1e999")                       # overflow is


def test_task_payload_struct_rules():
    f = gj.go_unmarshal_task_payload
    assert f(b'{"args": ["x"], "kwargs": {}}') == (["x"], {})
    assert f(b"null") == (None, None)
    assert f(b"{}") == (None, None)
    assert f(b'{"args": null, "kwargs": null}') == (None, None)
    assert f(b'{"ARGS": [1], "Kwargs": {"a": 2}}') == ([1.0], {"a": 2.0})          # case-insensitive
    assert f('{"\u212awargs": {"a": 1}, "arg\u017f": [2]}'.encode()) == ([2.0], {"a": 1.0})  # Kelvin / long s
    assert f(b'{"args": [1], "args": [2, 3]}') == ([2.0, 3.0], None)              # last wins
    assert f(b'{"kwargs": {"a": 1}, "kwargs": {"b": 2}}') == (None, {"a": 1.0, "b": 2.0})   # map merges
    assert f(b'{"kwargs": {"a": 1}, "kwargs": null}') == (None, None)
    assert f(b'{"other": 1e999, "args": []}') == ([], None)                       # skipped, not converted
    for bad in (b"[]", b'"x"', b"3", b"true", b'{"args": {}}', b'{"args": "x"}', b'{"kwargs": []}',
                b'{"args": [1e999]}', b'{"kwargs": {"a": [1e999]}}'):
        with pytest.raises(gj.GoJSONError):
            f(bad)


def test_time_rfc3339nano():
    t = gj.go_time_rfc3339nano
    assert t(0) == "1970-01-01T00:00:00Z"
    assert t(1_789_970_992_573_161_412) == "2026-09-21T06:09:52.573161412Z"
    assert t(1_789_970_992_500_000_000) == "2026-09-21T06:09:52.5Z"
    assert t(951_782_400 * 10**9) == "2000-02-29T00:00:00Z"
    assert t(gj.GO_ZERO_TIME_UNIX_NS) == "0001-01-01T00:00:00Z"
    assert t(0, 330) == "1970-01-01T05:30:00+05:30"
    assert t(0, -480) == "1969-12-31T16:00:00-08:00"


def test_task_message_encode_layout():
    env = QueueEnv(workspace_name="w", stub_id="s")
    tm = build_task_message(env, "00000000-0000-4000-8000-000000000000", None, None, 10**18)
    assert tm.encode() == (b'{"task_id":"00000000-0000-4000-8000-000000000000","workspace_name":"w","stub_id":"s",'
                           b'"executor":"taskqueue","args":[],"kwargs":null,"policy":{"max_retries":3,"timeout":3600,'
                           b'"expires":"2001-09-09T03:46:40Z","ttl":7200},"retries":0,"timestamp":1000000000}')
    tm = TaskMessage(task_id="t", args=["<", 1.5, None, True, {"b": [], "a": {}}], kwargs={"z": 1.0, "\u00e9": "x", "a": "&"},
                     policy=TaskPolicy())
    assert tm.encode() == ('{"task_id":"t","workspace_name":"","stub_id":"","executor":"","args":["\\u003c",1.5,null,true,'
                           '{"a":{},"b":[]}],"kwargs":{"a":"\\u0026","z":1,"\u00e9":"x"},"policy":{"max_retries":3,'
                           '"timeout":3600,"expires":"0001-01-01T00:00:00Z","ttl":0},"retries":0,"timestamp":0}').encode()


def test_decode_base64_probe():
    tm = TaskMessage.decode(b'{"task_id":"t","args":["aGVsbG8=","hello","",5],"kwargs":null}')
    assert tm.args == [b"hello", "hello", b"", 5.0]     # "hello" is not valid base64 (len % 4)


def test_format_uuid():
    assert format_uuid(bytes(range(16))) == "00010203-0405-0607-0809-0a0b0c0d0e0f"


def test_loop_quirks():
    ids = [bytes([i]) * 16 for i in range(16)]
    cases = [
        (sdk_put_payload("x"), "identity", "COMPLETE", b'"x"'),
        (sdk_put_payload(""), "identity", "COMPLETE", None),            # falsy -> no result bytes
        (sdk_put_payload(0), "identity", "COMPLETE", None),
        (sdk_put_payload(-0.0), "identity", "COMPLETE", None),          # Go "-0" -> Python int 0
        (sdk_put_payload(1.0), "identity", "COMPLETE", b"1"),           # float64 1 -> "1" -> int
        (sdk_put_payload(1e21), "identity", "COMPLETE", b"1e+21"),
        (sdk_put_payload(0.1), "identity", "COMPLETE", b"0.1"),
        (sdk_put_payload([1, "a"]), "identity", "COMPLETE", b'[1, "a"]'),
        (sdk_put_payload({"b": 1, "a": 2}), "identity", "COMPLETE", b'{"a": 2, "b": 1}'),   # Go sorts keys
        (sdk_put_payload(), "identity", "ERROR", None),
        (sdk_put_payload("a", "b"), "identity", "ERROR", None),
        (sdk_put_payload(s="a"), "identity", "ERROR", None),
        (b'{"args": ["x"]', "identity", "REJECTED", None),
        (b'{"args": [NaN]}', "identity", "REJECTED", None),
        (sdk_put_payload("123456789"), "crc32", "COMPLETE", b"3421780262"),   # CRC-32 check value 0xCBF43926
        (sdk_put_payload(5), "crc32", "ERROR", None),
    ]
    for (p, h, st, res), tid in zip(cases, ids):
        r = run_task_loop([p], [tid], h)[0]
        assert (r.status, r.result) == (st, res), (p, h)
