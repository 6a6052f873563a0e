#!/usr/bin/env python
"""bench.py — tasks/sec of the task fan-out hot path on N B200s (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: every pending task of the batch is popped,
deserialised, run through the handler and serialised (one persistent drain kernel per GPU).
Workload = BASELINE.json configs[1]: 1M x 256-character identity tasks per GPU (weak scaling:
every rank owns its own shard of the pending ring; no data-path collective).

  value      whole-job tasks/s with the batch already resident in HBM (kernel-resident number)
  e2e        the same through the C ABI with HOST buffers: b9_batch_push (H2D) + b9_drain (D2H)
  roofline   drain kernel: SURVEY.md §8(d) algorithmic bytes (578 B/task) / CUDA-event kernel time
  cpu_baseline  the oracle's C port of the reference loop on this box's host cores (rank 0, N=1)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_TASK = {"identity": 578.0}   # 256 in + 258 out + 2 x 32 B of index/id/header
L2_BYTES = 126e6


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tasks", type=int, default=1_000_000, help="tasks per GPU per step")
    ap.add_argument("--chars", type=int, default=256)
    ap.add_argument("--handler", default="identity")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cancelled", type=float, default=0.0, help="fraction of the resident tasks pushed with B9_TF_CANCELLED (the drain compacts them away: look-back path)")
    ap.add_argument("--skew", type=float, default=0.0, help="N>1: rank 0 pushes (1+skew)x the tasks and every step starts with the NCCL rebalance")
    ap.add_argument("--adversarial", type=float, default=0.01, help="share of tasks whose string needs escaping (SURVEY.md §8d: 1 %%)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def bind_to_gpu_numa_node(local_rank: int) -> str:
    """Run this rank on the CPUs next to its GPU (NVML's ideal affinity), BEFORE any pinned buffer is
    allocated: page-locked memory lands on the allocating thread's NUMA node, and a host<->device copy
    that crosses sockets is slower — it matters for the end-to-end number when 8 ranks share one host."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus near gpu {local_rank}"
    except Exception as e:                      # no NVML / no permission: keep the inherited affinity
        return f"unchanged ({type(e).__name__})"
    return "unchanged"


def ncu_traffic(args, n_tasks):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of this workload's drain kernel(s), from the
    committed ncu capture of the same command (profiles/ncu_traffic.json, written from
    scripts/gpu_profile_final.sh's reports). null when no capture of exactly this workload is committed."""
    try:
        table = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")))
    except (OSError, ValueError):
        return None
    key = f"{args.handler}:{n_tasks}:{args.chars}:{args.adversarial}" if args.handler == "identity" else f"{args.handler}:{n_tasks}"
    e = table.get(key)
    return None if e is None else e.get("dram_bytes")


def workload(args, rank):
    from beta9_b200 import synth
    n = args.tasks
    if args.skew > 0 and rank == 0 and int(os.environ.get("WORLD_SIZE", 1)) > 1:
        n = int(n * (1 + args.skew))
    if args.handler == "identity":
        return synth.strings_batch(n, args.chars, adversarial_frac=args.adversarial, seed=synth.SEED + 1000 * rank)
    if args.handler == "crc32":          # configs[2]: zipf 32..4096-char strings
        return synth.crc_batch(n, seed=synth.SEED + 1000 * rank)
    if args.handler == "vadd_f32":       # configs[3] payload shape: 2 x 32 fp32 as base64
        return synth.vadd_batch(n, seed=synth.SEED + 1000 * rank)
    if args.handler == "json_sum":       # configs[4] payload shape: 1 KB JSON documents
        return synth.json_batch(n, seed=synth.SEED + 1000 * rank)
    raise SystemExit(f"bench: unknown handler {args.handler}")


def run_reference(args):
    """--impl reference: the reference's CPU task loop on this box's host cores. The Go gateway and
    the Python runner cannot be built/imported here (no Go toolchain, SDK deps missing), so this is
    the oracle's C port of that loop (oracle/c/b9_oracle.c), multi-threaded over all cores; Redis,
    Postgres, gRPC and the object store are left out, which flatters the reference."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import coracle
    batch = workload(args, 0)
    cores = os.cpu_count() or 1
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        coracle.run_batch(batch.task_ids, batch.payload, batch.offsets, args.handler, nthreads=cores,
                          out_cap=int(batch.payload.size) + 64)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    v = batch.n * len(times) / total
    line = {
        "impl": "reference", "metric": "tasks_per_sec", "value": v, "unit": "tasks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"configs[1]: {batch.n} x {args.chars}-char identity tasks, reference CPU loop (C port), "
                               f"one bounded sample of {batch.n} tasks per step", "handler": args.handler},
        "cpu_baseline": {"value": v, "unit": "tasks/s", "cores": cores, "kind": "port",
                         "sample": f"{batch.n} tasks x {len(times)} steps, {cores} threads"},
        "e2e": {"value": v, "unit": "tasks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def _own_stdout():
    """stdout carries exactly ONE line, the JSON result: everything else a library prints there (NCCL's
    version banner, for one) is sent to stderr by pointing fd 1 at fd 2 for the lifetime of the run."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    _own_stdout()
    if args.impl == "reference":
        return run_reference(args)
    rank, local_rank, world = dist_env()
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the b200 arm has no CPU path")
    torch.cuda.set_device(local_rank)
    affinity = (bind_to_gpu_numa_node(local_rank) if world > 1 and not os.environ.get("B9_BENCH_NO_AFFINITY")
                else "inherited")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from beta9_b200 import _lib as L
    from beta9_b200.device_queue import DeviceQueue
    batch = workload(args, rank)
    n = batch.n
    in_bytes = int(batch.payload.size)
    dq = DeviceQueue(device=local_rank, ring_bytes=max(1 << 30, 4 * in_bytes), ring_tasks=max(1 << 21, 4 * n),
                     max_drain_tasks=max(1 << 21, n), max_result_bytes=max(1 << 30, 2 * in_bytes))

    rebalance = world > 1 and args.skew > 0
    if rebalance:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(DeviceQueue.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        dq.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ------------------------------------------------------------------ device-resident steps
    n_cancelled = 0
    if args.cancelled > 0:
        fl = (np.random.default_rng(5 + rank).random(batch.n) < args.cancelled).astype(np.uint8)
        n_cancelled = int(fl.sum())
        dq.push_batch(batch.task_ids, batch.payload, batch.offsets, flags=fl)
    else:
        dq.push_batch(batch.task_ids, batch.payload, batch.offsets)
    rebalance_info = None
    n_pushed = n
    if rebalance:
        # skewed ingest (rank 0 holds (1+skew)x): ONE NCCL all-to-all evens the pending bytes out before the drain
        barrier()
        r0 = time.perf_counter()
        info = dq.rebalance()
        barrier()
        r_ms = reduce_max(1e3 * (time.perf_counter() - r0))
        rebalance_info = {"ms": r_ms, "tasks_before_rank0": int(info.tasks_before) if rank == 0 else None,
                          "bytes_moved_total": reduce_sum(float(info.bytes_sent)), "skew": args.skew}
        n = dq.depth()                                   # my share after the exchange
    n_total = int(round(reduce_sum(float(n))))
    kernel_ms = []
    for _ in range(args.warmup):
        dq.drain_launch(args.handler, n, peek=True)
    sampler = ClockSampler(local_rank)
    launches0 = dq.stats().kernel_launches
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        # B9_DRAIN_ASYNC: the K steps are enqueued back to back, as a host that pipelines its drains would;
        # the barrier below (stream sync on every rank) closes the timed region
        got = dq.drain_launch(args.handler, n, peek=True, wait=False)
    dq.sync()
    burst_ms = float(dq.stats().last_drain_kernel_ms)        # CUDA events on the drain stream: first step's start -> last step's end
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop()
    launches = dq.stats().kernel_launches - launches0
    assert got == n, (got, n)
    for _ in range(min(args.steps, 5)):             # per-step kernel time (CUDA events around one step's kernels), outside the timed region
        dq.drain_launch(args.handler, n, peek=True)
        kernel_ms.append(dq.stats().last_drain_kernel_ms)
    out_bytes = int(dq.stats().last_drain_out_bytes)
    # timed on the device (CUDA events around the K steps), max over ranks; the host's wall clock over the same region,
    # barriers included, is reported beside it (a 4 ms region is at the mercy of one scheduling hiccup)
    wall_elapsed = reduce_max(t1 - t0)
    elapsed = reduce_max(burst_ms * 1e-3)
    value = n_total * args.steps / elapsed
    k_ms = reduce_max(statistics.mean(kernel_ms))
    # drop the resident batch
    dq.drain_launch(args.handler, n, peek=False)
    res = dq.fetch()
    assert res.n == n - (n_cancelled if not rebalance else 0) or rebalance, (res.n, n, n_cancelled)
    assert dq.depth() == 0
    if rebalance or n_cancelled:
        # the end-to-end leg below runs the un-skewed shard shape with every task live (the exchange / the
        # compaction of cancelled slots is timed above): take its reference records from one more drain
        res_n_after = n
        n = min(n_pushed, args.tasks)
        batch = batch.slice(0, n)
        in_bytes = int(batch.payload.size)
        dq.push_batch(batch.task_ids, batch.payload, batch.offsets)
        res = dq.drain(args.handler, n)
        out_bytes = int(res.payload.size)

    # ------------------------------------------------------------------ end to end through the C ABI, host buffers
    # Every step copies that step's inputs host->device from pinned memory and reads that step's result
    # records back device->host. Steps are software-pipelined the way a gateway would run them: the push of
    # batch k+1 is enqueued (b9_batch_push_async) before batch k is drained, so H2D(k+1) overlaps the kernel
    # and D2H(k) on the full-duplex link. Two input buffer sets alternate; results land in one pinned set.
    pins = []
    for _ in range(2):
        pi = dq.pinned(n * 16); pp = dq.pinned(in_bytes); po = dq.pinned((n + 1) * 8)
        pi.array[:] = batch.task_ids.reshape(-1); pp.array[:] = batch.payload
        po.view(np.uint64, n + 1)[:] = batch.offsets
        pins.append((pi, pp, po))
    cap_bytes = out_bytes + 4096
    o_ids = dq.pinned(n * 16); o_st = dq.pinned(n); o_has = dq.pinned(n); o_off = dq.pinned(n * 8); o_len = dq.pinned(n * 4); o_pl = dq.pinned(cap_bytes)
    resbuf = L.Results(o_ids.ptr, o_st.ptr, o_has.ptr, o_off.ptr, o_len.ptr, o_pl.ptr, n, cap_bytes, 0, 0, 0, 0)
    lib = L.load()
    hid = {"identity": 0, "crc32": 1, "vadd_f32": 2, "json_sum": 3}[args.handler]

    def e2e_push(k):
        pi, pp, po = pins[k & 1]
        rc = lib.b9_batch_push_async(dq._ctx, pi.ptr, pp.ptr, po.ptr, n, None)
        if rc != 0:
            raise SystemExit("push failed: " + L.last_error())

    def e2e_drain():
        r = lib.b9_drain(dq._ctx, hid, n, C.byref(resbuf))
        if r != n:
            raise SystemExit(f"drain returned {r}: " + L.last_error())

    def e2e_run(steps):
        e2e_push(0)
        for k in range(steps):
            if k + 1 < steps:
                e2e_push(k + 1)
            e2e_drain()

    e2e_run(3)
    s0 = dq.stats()
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.e2e_steps)
    barrier()
    t1 = time.perf_counter()
    s1 = dq.stats()
    e2e_elapsed = reduce_max(t1 - t0)
    e2e_value = int(round(reduce_sum(float(n)))) * args.e2e_steps / e2e_elapsed
    h2d = (s1.bytes_h2d - s0.bytes_h2d) // args.e2e_steps
    d2h = (s1.bytes_d2h - s0.bytes_d2h) // args.e2e_steps
    # the records that came back are the real ones
    e_off = o_off.view(np.uint64, n); e_len = o_len.view(np.uint32, n)
    for i in (0, n // 2, n - 1):
        assert bytes(o_pl.array[int(e_off[i]):int(e_off[i]) + int(e_len[i])]) == res.result(i)

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import coracle
        cores = os.cpu_count() or 1
        best = None
        for _ in range(3):
            c0 = time.perf_counter()
            o = coracle.run_batch(batch.task_ids, batch.payload, batch.offsets, args.handler, nthreads=cores, out_cap=in_bytes + 64)
            dt = time.perf_counter() - c0
            best = dt if best is None else min(best, dt)
        # while we are here: the device answers are the oracle's answers
        assert np.array_equal(o.payload, res.fifo_payload()) and np.array_equal(o.status, res.status)
        cpu = {"value": n / best, "unit": "tasks/s", "cores": cores, "kind": "port",
               "sample": f"all {n} tasks of the step, best of 3 passes, {cores} threads (oracle/c/b9_oracle.c)"}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        if args.handler in ALGO_BYTES_PER_TASK:
            algo = ALGO_BYTES_PER_TASK[args.handler] * n
        else:   # argument bytes as carried + result bytes + 2 x 32 B of index/id/header per task
            algo = float(in_bytes - 28 * n + out_bytes + 64 * n)
        achieved = algo / (k_ms * 1e-3) / 1e9
        line = {
            "metric": "tasks_per_sec", "value": value, "unit": "tasks/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "wall_ms_per_step": 1e3 * wall_elapsed / args.steps,
            "timing": "CUDA events on the drain stream around the K steps (enqueued back to back, B9_DRAIN_ASYNC), max over ranks; "
                      "wall_ms_per_step = host clock over the same region, barrier + stream sync on both sides",
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": (f"configs[1]: {n} x {args.chars}-char identity tasks per GPU ({100 * args.adversarial:g}% adversarial escapes), resident in HBM"
                                    if args.handler == "identity" else f"{args.handler}: {n} tasks per GPU, {in_bytes / n:.0f} payload bytes per task on average, resident in HBM"),
                       "handler": args.handler, "tasks_per_gpu": n,
                       "parallelism": (f"shard{world}" + ("+nccl_rebalance" if rebalance else "")) if world > 1 else "single",
                       "cpu_affinity": affinity, "cancelled_tasks_per_gpu": n_cancelled,
                       "l2": f"inputs {in_bytes / 1e6:.0f} MB + outputs {out_bytes / 1e6:.0f} MB per step exceed the 126 MB L2; no flush needed"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args, n), "kernel": f"b9::drain3_kernel<{args.handler}>" + (" + drain_slow_kernel" if args.handler == "identity" else ""),
                         "kernel_ms": k_ms, "algorithmic_bytes_per_task": algo / n, "peak_source": peak_src},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "tasks/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": args.e2e_steps, "ms_per_step": 1e3 * e2e_elapsed / args.e2e_steps,
                    "api": "b9_batch_push_async + b9_drain, pinned host buffers, push of step k+1 overlapped with drain of step k"},
            "gpu_launches": int(launches),
            "rebalance": rebalance_info,
            "clocks": clocks,
        }
        emit(line)
    for p in [x for t in pins for x in t] + [o_ids, o_st, o_has, o_off, o_len, o_pl]:
        p.free()
    dq.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
