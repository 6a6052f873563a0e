"""Runs the REFERENCE's own function path — `.map()` framing and the function runner — and freezes what it produced.

    python tests/golden/make_ref_function_golden.py      (needs /root/reference; run from the repo root)

Executed, unmodified, from /root/reference/sdk/src:
  producer  `_CallableWrapper._call_remote` (sdk/src/beta9/abstractions/function.py:198-232): cloudpickle.dumps of
            {"args": args, "kwargs": kwargs} into FunctionInvokeRequest.args, through `_format_args` + the map's
            per-input call (function.py:246-262); the stub records the request bytes
  runner    `invoke_function` (sdk/src/beta9/runner/function.py:236-283): FunctionGetArgs -> `_load_args`
            (cloudpickle, JSON fallback, :55-63) -> handler(*args, **kwargs) -> cloudpickle.dumps(result) ->
            FunctionSetResult; the stubs serve the args and record the result
Only the gRPC stubs are replaced. The handlers are oracle/pyoracle/handlers.py wrapped by the reference's function
`_CallableWrapper`. Output: tests/golden/ref_function_golden.json, checked on every CPU run by
tests/test_oracle_function.py (the oracle's restatement, oracle/pyoracle/funcloop.py) and on the GPU by
tests/test_gpu_function.py (the device's pickle-framed tasks)."""
import asyncio
import base64
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_ref_runner_golden import load_reference_runner  # noqa: E402  (same import harness: betterproto stand-in, watchdog mock)

HANDLERS = ["identity", "crc32"]


def inputs():
    """map() inputs: what `_format_args` sees (a tuple/list IS the argument list, anything else one argument)."""
    vals = ["", "a", "abc", "x" * 255, "y" * 256, "z" * 300, "é€\U0001f600", "\ud83d", "quote\"back\\slash\n", "w" * 70000,
            "args", "kwargs", 0, 5, 255, 256, 65535, 65536, 2**31 - 1, 2**31, 2**32 - 1, -1, 10**30, 1.5, None, True, False,
            b"bytes", b"\x80\x05\x95", ("two", "args"), ["a", "list"], (), [], {"d": 1}, ("k", {"callback_url": "x"}),
            [[1, 2]], ("nested", ("t",))]
    return vals


def main():
    load_reference_runner()
    os.environ["STUB_TYPE"] = "function"
    import beta9.abstractions.function as af
    import beta9.runner.function as rf
    import beta9.runner.common as rc
    import beta9.clients.function as cf
    from oracle.pyoracle import handlers as H
    out = {"source": "reference function path: abstractions/function.py:198-262 (framing), runner/function.py:236-283 (loop)", "cases": []}

    # ---- producer: the bytes `.map()` sends per input
    captured = []

    class InvokeStub:
        def function_invoke(self, req):
            captured.append(bytes(req.args))
            yield af.FunctionInvokeResponse(done=True, exit_code=0, result=b"", task_id="t")

    parent = af.Function.__new__(af.Function)
    parent.function_stub = InvokeStub()
    parent.stub_id = "stub"; parent.headless = False; parent.handler = "mod:f"
    w = af._CallableWrapper(H.identity, parent)
    for x in inputs():
        captured.clear()
        w._call_remote(*w._format_args(x))
        assert len(captured) == 1
        out["cases"].append({"input_repr": repr(x)[:80], "args_pickle": base64.b64encode(captured[0]).decode(), "results": {}})

    # ---- runner: what comes back for those bytes under each handler
    for h in HANDLERS:
        fw = af._CallableWrapper(getattr(H, h), parent)
        mod = types.ModuleType("ref_fn_handlers"); setattr(mod, h, fw)
        sys.modules["ref_fn_handlers"] = mod
        rc.config.handler = f"ref_fn_handlers:{h}"
        for c in out["cases"]:
            args_bytes = base64.b64decode(c["args_pickle"])
            got = {}

            class FnStub:
                def function_get_args(self, req):
                    return cf.FunctionGetArgsResponse(ok=True, args=args_bytes)

                def function_set_result(self, req):
                    got["result"] = bytes(req.result)
                    return cf.FunctionSetResultResponse(ok=True)

            ctx = rc.FunctionContext.new(config=rc.config, task_id="00000000-0000-4000-8000-000000000000", on_start_value=None)
            res = asyncio.run(rf.invoke_function(FnStub(), ctx, "00000000-0000-4000-8000-000000000000"))
            ok = res.exception is None
            c["results"][h] = {"ok": ok, "result_pickle": base64.b64encode(got["result"]).decode() if ok and "result" in got else None,
                               "exception": None if ok else type(res.exception).__name__}
    path = os.path.join(ROOT, "tests", "golden", "ref_function_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(path, len(out["cases"]), "inputs x", len(HANDLERS), "handlers through the reference function runner")


if __name__ == "__main__":
    main()
