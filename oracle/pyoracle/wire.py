"""CPU ORACLE (test infrastructure): the queue wire format `types.TaskMessage`.

Restates pkg/types/task.go:55-108,110-123 (struct, Encode, Decode, DefaultTaskPolicy,
TaskPolicy), the field population in pkg/task/dispatch.go:54-63,84-105 and the TTL/expires
defaulting in pkg/abstractions/taskqueue/taskqueue.go:191-196.

PARITY UNPINNED: no reference test pins Encode bytes; the vectors under
tests/golden/ are DERIVED from this restatement and must be re-validated against a Go 1.23
build when one is available.

The reference draws `task_id` (uuid.NewV4), `timestamp` (time.Now().Unix()) and
`policy.expires` (time.Now()+TTL) from the environment; the oracle takes them as inputs.
"""
from __future__ import annotations

import base64
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

from .gojson import (GO_ZERO_TIME_UNIX_NS, GoJSONError, go_marshal, go_quote,
                     go_time_rfc3339nano, go_unmarshal)

DEFAULT_TASK_QUEUE_TASK_TTL = 7200  # taskqueue.go `DefaultTaskQueueTaskTTL` (2h)
EXECUTOR_TASKQUEUE = "taskqueue"    # types.ExecutorTaskQueue, task.go:46


@dataclass
class TaskPolicy:
    """pkg/types/task.go:118-123; defaults :110-113."""
    max_retries: int = 3
    timeout: int = 3600
    expires_unix_ns: int = GO_ZERO_TIME_UNIX_NS
    expires_offset_min: int = 0
    ttl: int = 0

    def go_json(self) -> str:
        return ('{"max_retries":%d,"timeout":%d,"expires":"%s","ttl":%d}'
                % (self.max_retries, self.timeout,
                   go_time_rfc3339nano(self.expires_unix_ns, self.expires_offset_min), self.ttl))


@dataclass
class TaskMessage:
    """pkg/types/task.go:55-65 (field order = JSON order)."""
    task_id: str = ""
    workspace_name: str = ""
    stub_id: str = ""
    executor: str = ""
    args: Optional[List[Any]] = None
    kwargs: Optional[Dict[str, Any]] = None
    policy: TaskPolicy = field(default_factory=TaskPolicy)
    retries: int = 0
    timestamp: int = 0

    def encode(self) -> bytes:
        """TaskMessage.Encode (task.go:79-90): nil Args becomes `[]`, nil Kwargs stays `null`."""
        if self.args is None:
            self.args = []
        s = ('{"task_id":%s,"workspace_name":%s,"stub_id":%s,"executor":%s,"args":%s,"kwargs":%s,'
             '"policy":%s,"retries":%d,"timestamp":%d}'
             % (go_quote(self.task_id), go_quote(self.workspace_name), go_quote(self.stub_id),
                go_quote(self.executor), go_marshal(self.args), go_marshal(self.kwargs),
                self.policy.go_json(), self.retries, self.timestamp))
        return s.encode("utf-8")

    @classmethod
    def decode(cls, data: bytes) -> "TaskMessage":
        """TaskMessage.Decode (task.go:93-108) — including the base64 probe that turns every
        std-base64-decodable top-level string arg into []byte (kept here as Python `bytes`).
        Only the fields the hot path reads back are restored."""
        doc = go_unmarshal(data)
        if not isinstance(doc, dict):
            raise GoJSONError("TaskMessage.Decode: not an object")
        tm = cls()
        tm.task_id = doc.get("task_id") or ""
        tm.workspace_name = doc.get("workspace_name") or ""
        tm.stub_id = doc.get("stub_id") or ""
        tm.executor = doc.get("executor") or ""
        tm.args = doc.get("args")
        tm.kwargs = doc.get("kwargs")
        tm.retries = int(doc.get("retries") or 0)
        tm.timestamp = int(doc.get("timestamp") or 0)
        if tm.args:
            for i, a in enumerate(tm.args):
                if isinstance(a, str):
                    try:
                        # base64.StdEncoding.DecodeString: strict alphabet, padding required,
                        # '\r' and '\n' ignored
                        tm.args[i] = _go_std_b64decode(a)
                    except ValueError:
                        pass
        return tm


def _go_std_b64decode(s: str) -> bytes:
    cleaned = s.replace("\r", "").replace("\n", "")
    try:
        raw = cleaned.encode("ascii")
    except UnicodeEncodeError as e:
        raise ValueError("illegal base64 data") from e
    return base64.b64decode(raw, validate=True)


def format_uuid(raw16: bytes) -> str:
    """gofrs/uuid `UUID.String()`: canonical 8-4-4-4-12 lowercase hex of the 16 raw bytes."""
    h = bytes(raw16).hex()
    return f"{h[0:8]}-{h[8:12]}-{h[12:16]}-{h[16:20]}-{h[20:32]}"


@dataclass
class QueueEnv:
    """Per-queue constants the gateway fills into every message (dispatch.go:84-93,
    taskqueue.go:191-196)."""
    workspace_name: str = "ws-b200"
    stub_id: str = "7f1c2d3e-4a5b-4c6d-8e9f-0a1b2c3d4e5f"
    executor: str = EXECUTOR_TASKQUEUE
    max_retries: int = 3
    timeout: int = 3600
    ttl: int = DEFAULT_TASK_QUEUE_TASK_TTL


def build_task_message(env: QueueEnv, task_id: str, args, kwargs, now_unix_ns: int,
                       retries: int = 0) -> TaskMessage:
    """Dispatcher.Send field population (dispatch.go:84-93) with the put() policy
    (taskqueue.go:191-196): Expires = now + TTL; Timestamp = now in whole seconds."""
    ttl = env.ttl or DEFAULT_TASK_QUEUE_TASK_TTL
    pol = TaskPolicy(max_retries=env.max_retries, timeout=env.timeout,
                     expires_unix_ns=now_unix_ns + ttl * 10**9, ttl=ttl)
    return TaskMessage(task_id=task_id, workspace_name=env.workspace_name, stub_id=env.stub_id,
                       executor=env.executor, args=args, kwargs=kwargs, policy=pol,
                       retries=retries, timestamp=now_unix_ns // 10**9)
