"""The oracle's restatement of the reference's function path (oracle/pyoracle/funcloop.py) against what the reference's own,
unmodified `_call_remote` and `invoke_function` did (tests/golden/ref_function_golden.json, made by
tests/golden/make_ref_function_golden.py)."""
import base64
import json
import os

from oracle.pyoracle import funcloop
from tests.golden.make_ref_function_golden import inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_function_golden.json")


def load():
    return json.load(open(GOLDEN))["cases"]


def test_framing_is_the_references():
    cases = load()
    xs = inputs()
    assert len(xs) == len(cases)
    for x, c in zip(xs, cases):
        assert funcloop.frame_map_input(x) == base64.b64decode(c["args_pickle"]), c["input_repr"]
        blob, pickled = funcloop.gateway_args_blob(base64.b64decode(c["args_pickle"]))
        assert pickled and blob == base64.b64decode(c["args_pickle"])


def test_loop_is_the_references():
    for c in load():
        blob = base64.b64decode(c["args_pickle"])
        for h, want in c["results"].items():
            st, res = funcloop.run_function_task(blob, h)
            assert (st == funcloop.COMPLETE) == want["ok"], (c["input_repr"], h)
            assert res == (base64.b64decode(want["result_pickle"]) if want["ok"] else None), (c["input_repr"], h)


def test_read_result_rule():
    assert funcloop.read_result(None) is None and funcloop.read_result(b"") is None
    assert funcloop.read_result(funcloop.run_function_task(funcloop.frame_map_input("abc"), "identity")[1]) == "abc"
