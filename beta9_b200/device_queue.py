"""Python face of one b9_ctx: a device-resident task queue for one GPU.

Mirrors `taskQueueClient` (pkg/abstractions/taskqueue/client.go:29-106: Push / Pop / QueueLength)
batch-wise: `push_batch` = n x Push, `drain` = up to n x (Pop + runner loop + Complete hand-off).
All buffers are numpy arrays; results come back packed (see `DrainResult`).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _lib as L

HANDLERS = {"identity": L.H_IDENTITY, "echo": L.H_IDENTITY, "crc32": L.H_CRC32,
            "vadd_f32": L.H_VADD_F32, "json_sum": L.H_JSON_SUM}
STATUS_NAMES = {L.ST_COMPLETE: "COMPLETE", L.ST_ERROR: "ERROR", L.ST_RETRY: "RETRY",
                L.ST_REJECTED: "REJECTED", L.ST_UNSUPPORTED: "UNSUPPORTED"}


@dataclass
class DrainResult:
    task_ids: np.ndarray     # uint8 [n,16]
    status: np.ndarray       # uint8 [n]
    has_result: np.ndarray   # uint8 [n]
    offsets: np.ndarray      # uint64 [n]   start of record i's bytes in `payload`
    lengths: np.ndarray      # uint32 [n]
    payload: np.ndarray      # uint8 [n_bytes]  byte order = tile completion order (16-byte padded ranges), use offsets
    n_popped: int
    task_duration: float = 0.0   # seconds charged to every task of the drain (TaskQueueCompleteRequest.task_duration)

    @property
    def n(self) -> int:
        return int(self.status.shape[0])

    def result(self, i: int) -> Optional[bytes]:
        if not self.has_result[i]:
            return None
        o = int(self.offsets[i])
        return self.payload[o:o + int(self.lengths[i])].tobytes()

    def fifo_payload(self) -> np.ndarray:
        """The result bytes re-laid in record order (what `offsets[n+1]`-style consumers expect)."""
        if self.n == 0:
            return np.empty(0, np.uint8)
        ln = self.lengths.astype(np.int64)
        ends = np.cumsum(ln)
        idx = np.arange(int(ends[-1]), dtype=np.int64)
        rec = np.repeat(np.arange(self.n, dtype=np.int64), ln)
        src = self.offsets.astype(np.int64)[rec] + (idx - (ends - ln)[rec])
        return self.payload[src]

    def records(self) -> List[Tuple[bytes, str, Optional[bytes]]]:
        return [(self.task_ids[i].tobytes(), STATUS_NAMES[int(self.status[i])], self.result(i)) for i in range(self.n)]


class PinnedBuffer:
    """Page-locked host memory from b9_host_alloc, viewed as a numpy array."""

    def __init__(self, q: "DeviceQueue", nbytes: int):
        self._q = q
        self.nbytes = int(nbytes)
        self.ptr = q._lib.b9_host_alloc(q._ctx, self.nbytes)
        if not self.ptr:
            raise L.B9Error(L.B9_ENOMEM, L.last_error())
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(1, self.nbytes)).from_address(self.ptr))[:self.nbytes]

    def view(self, dtype, count: int, offset: int = 0) -> np.ndarray:
        return self.array[offset:offset + count * np.dtype(dtype).itemsize].view(dtype)

    def free(self) -> None:
        if self.ptr:
            self.array = None
            self._q._lib.b9_host_free(self._q._ctx, self.ptr)
            self.ptr = None


class DeviceQueue:
    def __init__(self, device: int = 0, ring_bytes: int = 0, ring_tasks: int = 0, max_drain_tasks: int = 0,
                 max_result_bytes: int = 0, max_task_bytes: int = 0):
        self._lib = L.load()
        o = L.Opts(C.sizeof(L.Opts), device, ring_bytes, ring_tasks, max_drain_tasks, max_result_bytes, max_task_bytes, 0)
        ctx = C.c_void_p()
        rc = self._lib.b9_ctx_create(C.byref(o), C.byref(ctx))
        if rc != L.B9_OK:
            raise L.B9Error(rc, L.last_error())
        self._ctx = ctx
        self.device = device
        self._res_bufs = None

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.b9_ctx_destroy(self._ctx)
            self._ctx = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise L.B9Error(int(rc), L.last_error(self._ctx))
        return int(rc)

    # ---- producer side (client.go:29-41 Push, batch-wise)
    def push_batch(self, task_ids: np.ndarray, payload: np.ndarray, offsets: np.ndarray,
                   timestamp_unix: Optional[np.ndarray] = None, expires_unix_ns: Optional[np.ndarray] = None,
                   retries: Optional[np.ndarray] = None, flags: Optional[np.ndarray] = None) -> None:
        ids = np.ascontiguousarray(task_ids, dtype=np.uint8)
        pl = np.ascontiguousarray(payload, dtype=np.uint8)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = off.shape[0] - 1
        assert ids.size == n * 16, "task_ids must be n x 16 bytes"
        keep = []

        def ptr(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            assert a.shape[0] == n
            keep.append(a)
            return a.ctypes.data
        meta = L.PushMeta(ptr(timestamp_unix, np.int64), ptr(expires_unix_ns, np.int64), ptr(retries, np.uint8), ptr(flags, np.uint8))
        self._check(self._lib.b9_batch_push(self._ctx, ids.ctypes.data, pl.ctypes.data if pl.size else ids.ctypes.data,
                                            off.ctypes.data, n, C.byref(meta)))

    def push_scattered(self, task_ids: np.ndarray, pointers: np.ndarray, lengths: np.ndarray) -> None:
        """b9_batch_push_v: `pointers` uint64[n] host addresses of the payloads (anywhere in pageable memory),
        `lengths` uint32[n]. The library gathers them into its page-locked arena (the pack step) and pushes that."""
        ids = np.ascontiguousarray(task_ids, dtype=np.uint8)
        ptrs = np.ascontiguousarray(pointers, dtype=np.uint64)
        lens = np.ascontiguousarray(lengths, dtype=np.uint32)
        n = lens.shape[0]
        assert ids.size == n * 16 and ptrs.shape[0] == n
        self._check(self._lib.b9_batch_push_v(self._ctx, ids.ctypes.data, ptrs.ctypes.data, lens.ctypes.data, n, None))

    def submit(self, task_id: bytes, payload: bytes, flags: int = 0) -> None:
        """b9_submit: one task (16 raw id bytes + payload) from any thread; batched under the hood, `flush()` pushes."""
        self._check(self._lib.b9_submit(self._ctx, task_id, payload, len(payload), flags))

    def flush(self) -> int:
        return self._check(self._lib.b9_flush(self._ctx))

    def buffered(self) -> int:
        return int(self._lib.b9_buffered(self._ctx))

    def depth(self) -> int:                       # client.go:99-106 QueueLength / task_redis.go:112-119 TasksInFlight
        return int(self._lib.b9_depth(self._ctx))

    def running(self) -> int:                     # task_redis.go:58 TasksClaimed / client.go:109 TasksRunning
        return int(self._lib.b9_running(self._ctx))

    def depth_bytes(self) -> int:
        return int(self._lib.b9_depth_bytes(self._ctx))

    def expire(self, now_unix_ns: int) -> int:    # dispatch.go:173-230 monitor, unclaimed branch
        return self._check(self._lib.b9_expire(self._ctx, now_unix_ns))

    # ---- consumer side
    def drain(self, handler: str, max_tasks: int = 1 << 22, cap_bytes: Optional[int] = None) -> DrainResult:
        h = HANDLERS[handler]
        n = self._check(self._lib.b9_drain_launch(self._ctx, h, max_tasks, 0))
        return self._fetch(n, cap_bytes)

    def drain_launch(self, handler: str, max_tasks: int = 1 << 22, peek: bool = False, wait: bool = True) -> int:
        """wait=False (B9_DRAIN_ASYNC): returns once the kernels are enqueued; `sync()` or `fetch()` completes it."""
        return self._check(self._lib.b9_drain_launch(self._ctx, HANDLERS[handler], max_tasks, (1 if peek else 0) | (0 if wait else 2)))

    def fetch(self, cap_bytes: Optional[int] = None) -> DrainResult:
        self.sync()                                  # completes a wait=False launch: its byte count sizes the buffers below
        st = self.stats()
        return self._fetch(None, cap_bytes if cap_bytes is not None else int(st.last_drain_out_bytes))

    def _fetch(self, n: Optional[int], cap_bytes: Optional[int]) -> DrainResult:
        st = self.stats()
        if cap_bytes is None:
            cap_bytes = int(st.last_drain_out_bytes)
        if n is None:
            n = int(st.last_drain_tiles) * 128      # an upper bound for either kernel generation
        ids = np.empty((max(n, 1), 16), np.uint8)
        status = np.empty(max(n, 1), np.uint8)
        has = np.empty(max(n, 1), np.uint8)
        off = np.zeros(max(n, 1), np.uint64)
        ln = np.zeros(max(n, 1), np.uint32)
        pl = np.empty(max(cap_bytes, 1), np.uint8)
        r = L.Results(ids.ctypes.data, status.ctypes.data, has.ctypes.data, off.ctypes.data, ln.ctypes.data, pl.ctypes.data,
                      max(n, 1), max(cap_bytes, 1), 0, 0, 0, 0, 0.0, 0)
        got = self._check(self._lib.b9_drain_fetch(self._ctx, C.byref(r)))
        return DrainResult(ids[:got], status[:got], has[:got], off[:got], ln[:got], pl[:int(r.n_bytes)], int(r.n_popped), float(r.task_duration))

    def drain_object(self, handler: str, max_tasks: int = 1 << 22) -> np.ndarray:
        """One drain whose records land as ONE result-sink object (b9_drain_fetch_object): header + index + blob."""
        self._check(self._lib.b9_drain_launch(self._ctx, HANDLERS[handler], max_tasks, 0))
        need = C.c_uint64(0)
        probe = np.zeros(128, np.uint8)
        rc = self._lib.b9_drain_fetch_object(self._ctx, probe.ctypes.data, 0, C.byref(need))      # B9_ENOSPC: tells the size, consumes nothing
        if rc != L.B9_ENOSPC:
            self._check(rc)
        obj = np.empty(int(need.value), np.uint8)
        self._check(self._lib.b9_drain_fetch_object(self._ctx, obj.ctypes.data, obj.size, C.byref(need)))
        return obj

    def drain_into(self, handler: str, max_tasks: int, res: "L.Results") -> int:
        """b9_drain straight into caller-owned (ideally pinned) buffers described by `res`."""
        return self._check(self._lib.b9_drain(self._ctx, HANDLERS[handler], max_tasks, C.byref(res)))

    def wire_encode(self, workspace_name: str, stub_id: str, max_tasks: int = 1 << 22, executor: str = "taskqueue",
                    max_retries: int = 3, timeout: int = 3600, ttl: int = 7200) -> DrainResult:
        """TaskMessage.Encode (pkg/types/task.go:79-90) of the pending tasks, on the device; they stay pending."""
        env = L.WireEnv(workspace_name.encode(), stub_id.encode(), executor.encode(), max_retries, timeout, ttl)
        n = self._check(self._lib.b9_wire_encode(self._ctx, C.byref(env), max_tasks))
        return self._fetch(n, None)

    # ---- multi-GPU (one DeviceQueue per rank / GPU)
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = L.load().b9_comm_unique_id(buf)
        if rc != 0:
            raise L.B9Error(rc, L.last_error())
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, world: int) -> None:
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self._lib.b9_comm_init(self._ctx, buf, rank, world))

    def rebalance(self) -> "L.RebalanceInfo":
        """Collective: byte-quantile rebalance of the pending ring over the communicator (NCCL all-to-all)."""
        info = L.RebalanceInfo()
        self._check(self._lib.b9_rebalance(self._ctx, C.byref(info)))
        return info

    def stats(self) -> "L.Stats":
        s = L.Stats()
        self._check(self._lib.b9_stats_get(self._ctx, C.byref(s)))
        return s

    def sync(self) -> None:
        self._check(self._lib.b9_sync(self._ctx))

    def pinned(self, nbytes: int) -> PinnedBuffer:
        return PinnedBuffer(self, nbytes)


def rebalance_plan(world: int, rank: int, counts, nbytes, prefix) -> Tuple[np.ndarray, np.ndarray]:
    """b9_rebalance_plan: local FIFO range [lo[d], hi[d]) destined for every rank d (pure host arithmetic)."""
    counts = np.ascontiguousarray(counts, np.uint64); nbytes = np.ascontiguousarray(nbytes, np.uint64)
    prefix = np.ascontiguousarray(prefix, np.uint64)
    lo = np.zeros(world, np.uint64); hi = np.zeros(world, np.uint64)
    rc = L.load().b9_rebalance_plan(world, rank, counts.ctypes.data, nbytes.ctypes.data, prefix.ctypes.data,
                                    prefix.shape[0] - 1, lo.ctypes.data, hi.ctypes.data)
    if rc != 0:
        raise L.B9Error(rc, L.last_error())
    return lo, hi


def task_queue_scale(queue_length: int, tasks_per_container: int, max_containers: int, max_replicas: int) -> Tuple[int, bool]:
    """taskQueueScaleFunc (pkg/abstractions/taskqueue/autoscaler.go:53-79) via the C ABI."""
    v = C.c_int(0)
    d = L.load().b9_task_queue_scale(queue_length, tasks_per_container, max_containers, max_replicas, C.byref(v))
    return int(d), bool(v.value)
