"""CPU ORACLE (test infrastructure): the reference's FUNCTION path for one input — what `.map()` sends and what comes back.

Restates:

  frame    sdk/src/beta9/abstractions/function.py:198-205    cloudpickle.dumps({"args": args, "kwargs": kwargs})
           sdk/src/beta9/abstractions/function.py:246-262    map(): one call per input, `_format_args` first
  gateway  pkg/abstractions/function/task.go:84,104-108      a single []byte argument starting 80 05 95 IS the runner's
                                                             argument blob (otherwise the TaskPayload's JSON is)
  loop     sdk/src/beta9/runner/function.py:55-63,236-283    `_load_args` (cloudpickle, JSON fallback) ->
                                                             handler(*(args or []), **(kwargs or {})) with
                                                             `callback_url` popped -> cloudpickle.dumps(result)
  read     sdk/src/beta9/abstractions/function.py:228-232    `if not result: None` else cloudpickle.loads(result)

PINNED by running the reference's own code: tests/golden/make_ref_function_golden.py executes the unmodified
`_call_remote` and `invoke_function` with only the gRPC stubs replaced; tests/test_oracle_function.py holds this module to
what they did (tests/golden/ref_function_golden.json). cloudpickle is the pinned third-party arithmetic here
(cloudpickle 3.1.2 in this image; pickle protocol 5 of CPython 3.12): the oracle calls the very same library.
"""
from __future__ import annotations

import json
from typing import Any, Callable, List, Optional, Tuple

import cloudpickle

from .handlers import HANDLERS

CLOUDPICKLE_HEADER = b"\x80\x05\x95"          # pkg/abstractions/function/task.go:84

COMPLETE, ERROR = "COMPLETE", "ERROR"


def format_args(x: Any) -> List[Any]:
    """function.py:246-251."""
    if isinstance(x, tuple):
        return list(x)
    if not isinstance(x, list):
        return [x]
    return x


def frame_call(*args: Any, **kwargs: Any) -> bytes:
    """function.py:198-205: FunctionInvokeRequest.args."""
    return cloudpickle.dumps({"args": args, "kwargs": kwargs})


def frame_map_input(x: Any) -> bytes:
    return frame_call(*format_args(x))


def gateway_args_blob(args_field: bytes) -> Tuple[bytes, bool]:
    """function/task.go:94-108: (what FunctionGetArgs will hand to the runner, was it taken as a cloudpickle blob).
    The invoke handler wraps the request's bytes as the single argument of the TaskPayload."""
    if args_field.startswith(CLOUDPICKLE_HEADER):
        return bytes(args_field), True
    # not reachable from the SDK (it always pickles); Go would marshal TaskPayload{Args: [<base64 string>]}
    import base64
    return json.dumps({"args": [base64.b64encode(args_field).decode()], "kwargs": None}, separators=(",", ":")).encode(), False


def load_args(blob: bytes) -> dict:
    """runner/function.py:55-63."""
    try:
        return cloudpickle.loads(blob)
    except BaseException:
        return json.loads(blob.decode("utf-8"))     # json.JSONDecodeError -> InvalidFunctionArgumentsError -> ERROR


def run_function_task(blob: bytes, handler: "str | Callable[..., Any]") -> Tuple[str, Optional[bytes]]:
    """-> (status, FunctionSetResultRequest.result bytes or None)."""
    fn = HANDLERS[handler] if isinstance(handler, str) else handler
    try:
        payload = load_args(blob)
        args = payload.get("args") or []
        kwargs = payload.get("kwargs") or {}
        kwargs.pop("callback_url", None)
        result = fn(*args, **kwargs)
        return COMPLETE, cloudpickle.dumps(result)
    except BaseException:
        return ERROR, None


def read_result(result: Optional[bytes]) -> Any:
    """function.py:228-232."""
    if not result:
        return None
    return cloudpickle.loads(result)
