"""CPU ORACLE (test infrastructure): where a task's result bytes go after the runner, and how they come back.

Restates, per task:

  store   pkg/abstractions/taskqueue/taskqueue.go:394-399   `if in.Result != nil && workspace.StorageAvailable()` ->
          pkg/task/dispatch.go:120-144 StoreTaskResult:     one object-store PUT of the raw result bytes under
          pkg/task/dispatch.go:18-20   GetTaskResultPath:   "task/<task id>/result"
  read    pkg/api/v1/task.go:295-325   addResultToTask:     download; if len(result) > 0: json.Unmarshal(result, &t.Result)
                                                            (t.Result is a json.RawMessage, pkg/types/backend.go:346) and on an
                                                            error t.Result = {"base64":"<std base64 of the bytes>"}

json.Unmarshal into a RawMessage checks SYNTAX only (checkValid) and hands the value's own bytes — without the white space
around it — to RawMessage.UnmarshalJSON; no number is converted, so "1e999" is kept as it is (PARITY UNPINNED at this Go
rule: the reference has no test for addResultToTask). The batched sink (include/b9gpu.h b9_sink_*) must give, for every
task id, exactly what this per-task path gives.
"""
from __future__ import annotations

import base64
import sys
from typing import Dict, Optional

from .gojson import GoJSONError, _Parser
from .wire import format_uuid


def task_result_path(task_id: str) -> str:
    return f"task/{task_id}/result"


def store_task_result(store: Dict[str, bytes], task_id_raw16: bytes, result: Optional[bytes]) -> None:
    """TaskQueueComplete's hand-off: the runner's `result` (None = the field was not set: a falsy result or an error,
    runner/taskqueue.py:378) is uploaded only when present."""
    if result is not None:
        store[task_result_path(format_uuid(task_id_raw16))] = bytes(result)


def raw_message(data: bytes) -> bytes:
    """json.Unmarshal(data, &rawMessage): raises GoJSONError on a syntax error, else the value's literal bytes."""
    p = _Parser(bytes(data))
    p.skip_ws()
    if p.i >= p.n:
        raise p.err("unexpected end of JSON input")
    start = p.i
    limit = sys.getrecursionlimit()
    sys.setrecursionlimit(max(limit, 4 * _Parser.MAX_DEPTH + 1000))      # the restated parser recurses; Go's limit is 10000 levels
    try:
        p.parse_value(0)
    finally:
        sys.setrecursionlimit(limit)
    end = p.i
    p.skip_ws()
    if p.i != p.n:
        raise p.err("invalid character after top-level value")
    return bytes(data[start:end])


def add_result_to_task(stored: Optional[bytes]) -> Optional[bytes]:
    """t.Result after addResultToTask (None = left unset: nothing stored, or an empty object)."""
    if stored is None or len(stored) == 0:
        return None
    try:
        return raw_message(stored)
    except GoJSONError:
        return b'{"base64":"' + base64.b64encode(stored) + b'"}'
