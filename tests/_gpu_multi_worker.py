"""One rank of the multi-GPU rebalance test: push a skewed shard, b9_rebalance (NCCL), drain, dump."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    rank, world, out_dir, handler = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    mode = sys.argv[5] if len(sys.argv) > 5 else "one_push"
    from beta9_b200 import _lib as L
    from beta9_b200.device_queue import DeviceQueue
    from tests._rebalance_worker import shard_for
    # "enospc": the last rank's ring is too small for its fair share — the exchange must be refused on EVERY rank
    small = mode == "enospc" and rank == world - 1
    q = DeviceQueue(device=rank, ring_bytes=(1 << 16) if small else (1 << 28), ring_tasks=1 << 18, max_drain_tasks=1 << 18, max_result_bytes=1 << 28)
    idf = os.path.join(out_dir, "nccl_id.bin")
    if rank == 0:
        uid = DeviceQueue.comm_unique_id()
        with open(idf + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(idf + ".tmp", idf)
    else:
        t0 = time.time()
        while not os.path.exists(idf):
            if time.time() - t0 > 120:
                raise SystemExit("no NCCL id")
            time.sleep(0.05)
        uid = open(idf, "rb").read()
    q.comm_init(uid, rank, world)
    b = shard_for(rank, world, "rank0_heavy")
    # some cancelled tasks travel too
    flags = (np.arange(b.n) % 11 == 0).astype(np.uint8)

    def push(batch, fl):
        if mode == "many_pushes":      # the ranges that leave span several ring segments: staged through the arena, not sent from the ring
            k = max(1, batch.n // 7)
            for lo in range(0, batch.n, k):
                part = batch.slice(lo, min(batch.n, lo + k))
                q.push_batch(part.task_ids, part.payload, part.offsets, flags=fl[lo:lo + part.n])
        else:
            q.push_batch(batch.task_ids, batch.payload, batch.offsets, flags=fl)

    push(b, flags)
    if mode == "enospc":
        d0, b0 = q.depth(), q.depth_bytes()
        try:
            q.rebalance()
            refused = False
        except L.B9Error as e:
            refused = e.code == L.B9_ENOSPC
        assert refused, "the exchange went through although a rank cannot take its share"
        assert (q.depth(), q.depth_bytes()) == (d0, b0)          # nothing moved, nothing was lost
        r = q.drain(handler)                                      # ... and the local shard still drains
        info = L.RebalanceInfo()
        info.tasks_before = info.tasks_after = d0; info.bytes_before = info.bytes_after = b0
    else:
        info = q.rebalance()
        assert q.depth() == info.tasks_after and q.depth_bytes() == info.bytes_after
        if mode == "twice":            # the ring has aged: drain half, push again, exchange again
            half = q.drain(handler, max_tasks=q.depth() // 2)
            push(b, flags)
            info2 = q.rebalance()
            assert q.depth() == info2.tasks_after and q.depth_bytes() == info2.bytes_after
            rest = q.drain(handler)
            np.savez(os.path.join(out_dir, f"gpu_rank{rank}_half.npz"), ids=half.task_ids, status=half.status, has=half.has_result, lens=half.lengths, payload=half.fifo_payload())
            r = rest
        else:
            r = q.drain(handler)
    np.savez(os.path.join(out_dir, f"gpu_rank{rank}.npz"), ids=r.task_ids, status=r.status, has=r.has_result, lens=r.lengths,
             payload=r.fifo_payload(), before_ids=b.task_ids, before_offsets=b.offsets, before_payload=b.payload, flags=flags,
             info=np.array([info.tasks_before, info.bytes_before, info.tasks_sent, info.bytes_sent, info.tasks_received,
                            info.bytes_received, info.tasks_after, info.bytes_after], np.int64))
    q.close()


if __name__ == "__main__":
    main()
