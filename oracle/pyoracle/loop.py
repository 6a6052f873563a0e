"""CPU ORACLE (test infrastructure): the reference task loop, end to end, for one batch.

Restates, per task and strictly in the reference's order:

  put      sdk/src/beta9/abstractions/taskqueue.py:284-285   json.dumps({"args","kwargs"})
  decode   pkg/abstractions/taskqueue/taskqueue.go:213-214    json.Unmarshal -> TaskPayload
  send     pkg/task/dispatch.go:84-105                        TaskMessage fill + Encode
  push     pkg/abstractions/taskqueue/client.go:29-41         RPUSH   (in-memory FIFO here)
  pop      pkg/abstractions/taskqueue/client.go:43-96         LPOP; the popped bytes go to the runner
  loads    sdk/src/beta9/runner/taskqueue.py:196-201          json.loads(task_msg) -> Task
  call     sdk/src/beta9/runner/taskqueue.py:349-361          handler(*(args or []), **(kwargs or {}))
  result   sdk/src/beta9/runner/taskqueue.py:378              serialize_result(result) if result else None
           sdk/src/beta9/runner/common.py:484-489             json.dumps(result).encode("utf-8")

The Python halves call the very same stdlib functions the reference calls, and are PINNED against
the reference's own code: tests/golden/make_ref_runner_golden.py runs the unmodified
TaskQueueWorker.process_tasks / FunctionHandler / serialize_result / _CallableWrapper.put over the
golden wire records and tests/test_oracle_vs_ref_runner.py compares this module with what they did.
The Go halves are restated in gojson.py/wire.py (PARITY UNPINNED, see those headers). Redis, Postgres, gRPC and
the object store are omitted: they move bytes, they do not change them.
"""
from __future__ import annotations

import json
from collections import deque
from dataclasses import dataclass
from typing import Any, Callable, Iterable, List, Optional, Sequence, Tuple

from .gojson import GoJSONError, go_unmarshal_task_payload
from .handlers import HANDLERS
from .httpserialize import InvalidRequestPayload, serialize_http_payload
from .wire import QueueEnv, build_task_message, format_uuid

# Task status strings, sdk/src/beta9/type.py TaskStatus / pkg/types/backend.go TaskStatus*
COMPLETE = "COMPLETE"
ERROR = "ERROR"
RETRY = "RETRY"
# Not a reference status: `TaskQueuePut` answered Ok:false and no task was created
# (taskqueue.go:213-218). The batch interface has to report it per task.
REJECTED = "REJECTED"


@dataclass
class TaskResult:
    task_id: bytes              # raw 16-byte uuid
    status: str
    result: Optional[bytes]     # TaskQueueCompleteRequest.result (taskqueue.proto:55); None = unset
    wire: Optional[bytes] = None  # the bytes that sat in the Redis list


def sdk_put_payload(*args: Any, **kwargs: Any) -> bytes:
    """_CallableWrapper.put (sdk taskqueue.py:284-287): the TaskQueuePutRequest.payload bytes."""
    return json.dumps({"args": args, "kwargs": kwargs}).encode("utf-8")


def serialize_result(result: Any) -> Optional[bytes]:
    """runner/common.py:484-489."""
    try:
        return json.dumps(result).encode("utf-8")
    except Exception:
        return None


def run_task_loop(payloads: Sequence[bytes], task_ids: Sequence[bytes],
                  handler: "str | Callable[..., Any]", env: Optional[QueueEnv] = None,
                  now_unix_ns: int = 1_789_970_992_573_161_412,
                  keep_wire: bool = False, http_body: bool = False) -> List[TaskResult]:
    """Push the whole batch, then drain it FIFO through one runner (workers=1).

    http_body=False: `payloads` are TaskQueuePutRequest.payload bytes (the SDK's put), decoded by
    json.Unmarshal into TaskPayload (taskqueue.go:213-214). http_body=True: they are HTTP request
    bodies of the task-queue endpoint (taskqueue/http.go:38-78), turned into a TaskPayload by
    SerializeHttpPayload (pkg/task/serialize.go:16-101; no query string here); a body that is refused
    ("invalid request payload", HTTP 400) never becomes a task: REJECTED."""
    env = env or QueueEnv()
    fn = HANDLERS[handler] if isinstance(handler, str) else handler
    queue: deque = deque()
    results: List[Optional[TaskResult]] = [None] * len(payloads)
    # ---- producer side: one TaskQueuePut per payload
    for i, (p, tid) in enumerate(zip(payloads, task_ids)):
        try:
            if http_body:
                args, kwargs = serialize_http_payload(bytes(p))
            else:
                args, kwargs = go_unmarshal_task_payload(bytes(p))
        except (GoJSONError, InvalidRequestPayload):
            results[i] = TaskResult(bytes(tid), REJECTED, None)
            continue
        tm = build_task_message(env, format_uuid(tid), args, kwargs, now_unix_ns)
        queue.append((i, tm.encode()))          # Execute -> client.Push -> RPUSH
    # ---- consumer side: one runner process draining the list
    while queue:
        i, wire = queue.popleft()               # LPOP
        task = json.loads(wire)                 # runner/taskqueue.py:196
        assert task["task_id"] == format_uuid(task_ids[i])
        args = task["args"] or []               # :349-350
        kwargs = task["kwargs"] or {}
        status = COMPLETE
        result = None
        try:
            result = fn(*args, **kwargs)        # :353
        except BaseException:
            status = ERROR                      # :356 (retry_for is empty by default)
        out = serialize_result(result) if result else None   # :378
        results[i] = TaskResult(bytes(task_ids[i]), status, out, wire if keep_wire else None)
    return results  # type: ignore[return-value]
