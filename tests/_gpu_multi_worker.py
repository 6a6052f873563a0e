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
    from beta9_b200.device_queue import DeviceQueue
    from tests._rebalance_worker import shard_for
    q = DeviceQueue(device=rank, ring_bytes=1 << 28, ring_tasks=1 << 18, max_drain_tasks=1 << 18, max_result_bytes=1 << 28)
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
    q.push_batch(b.task_ids, b.payload, b.offsets, flags=flags)
    info = q.rebalance()
    assert q.depth() == info.tasks_after and q.depth_bytes() == info.bytes_after
    r = q.drain(handler)
    np.savez(os.path.join(out_dir, f"gpu_rank{rank}.npz"), ids=r.task_ids, status=r.status, has=r.has_result, lens=r.lengths,
             payload=r.fifo_payload(), before_ids=b.task_ids, before_offsets=b.offsets, before_payload=b.payload, flags=flags,
             info=np.array([info.tasks_before, info.bytes_before, info.tasks_sent, info.bytes_sent, info.tasks_received,
                            info.bytes_received, info.tasks_after, info.bytes_after], np.int64))
    q.close()


if __name__ == "__main__":
    main()
