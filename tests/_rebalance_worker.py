"""Worker for the world_size>1 CPU tests of the rebalance path (gloo): same plan function as the NCCL
path (b9_rebalance_plan through the C ABI), the exchange itself emulated with gloo point-to-point."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def shard_for(rank: int, world: int, skew: str):
    from beta9_b200 import synth
    base = 400
    if skew == "rank0_heavy":
        n = base * (3 * world if rank == 0 else 1)
    elif skew == "last_empty":
        n = 0 if rank == world - 1 else base
    elif skew == "zipf":
        n = base
    else:
        n = base
    if n == 0:
        return synth.Batch(np.zeros((0, 16), np.uint8), np.zeros(0, np.uint8), np.zeros(1, np.uint64), "empty")
    if skew == "zipf":
        b = synth.crc_batch(n, seed=100 + rank)
    else:
        b = synth.strings_batch(n, 48 + 16 * rank, adversarial_frac=0.1, seed=100 + rank)
    b.task_ids = synth.task_ids(n, seed=100 + rank)
    return b


def run(rank: int, world: int, port: int, skew: str, out_dir: str):
    import torch
    import torch.distributed as dist
    from beta9_b200.device_queue import rebalance_plan
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        b = shard_for(rank, world, skew)
        n = b.n
        prefix = (b.offsets - b.offsets[0]).astype(np.uint64)
        mine = torch.tensor([n, int(prefix[-1])], dtype=torch.int64)
        table = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(table, mine)                                   # step 1: (count, bytes) of every rank
        counts = np.array([int(t[0]) for t in table], np.uint64)
        nbytes = np.array([int(t[1]) for t in table], np.uint64)
        lo, hi = rebalance_plan(world, rank, counts, nbytes, prefix)    # step 2: same plan everywhere
        # the plan partitions my FIFO into contiguous, ordered ranges
        assert int(lo[0]) == 0 and int(hi[-1]) == n
        assert all(int(hi[d]) == int(lo[d + 1]) for d in range(world - 1))
        # step 3: send matrix row
        row = torch.tensor([[int(hi[d] - lo[d]), int(prefix[int(hi[d])] - prefix[int(lo[d])])] for d in range(world)], dtype=torch.int64)
        rows = [torch.zeros_like(row) for _ in range(world)]
        dist.all_gather(rows, row)
        # step 4: the exchange (ids + lengths + payload), point to point
        reqs, recv_bufs = [], {}
        for d in range(world):
            if d == rank:
                continue
            k, pb = int(row[d, 0]), int(row[d, 1])
            if k:
                a, z = int(lo[d]), int(hi[d])
                ids = torch.from_numpy(b.task_ids[a:z].reshape(-1).copy())
                lens = torch.from_numpy(np.diff(b.offsets[a:z + 1]).astype(np.int64))
                pay = torch.from_numpy(b.payload[int(b.offsets[a]):int(b.offsets[z])].copy()) if pb else torch.zeros(0, dtype=torch.uint8)
                reqs += [dist.isend(ids, d), dist.isend(lens, d)] + ([dist.isend(pay, d)] if pb else [])
            rk, rb = int(rows[d][rank, 0]), int(rows[d][rank, 1])
            if rk:
                recv_bufs[d] = (torch.zeros(rk * 16, dtype=torch.uint8), torch.zeros(rk, dtype=torch.int64), torch.zeros(rb, dtype=torch.uint8))
                reqs += [dist.irecv(recv_bufs[d][0], d), dist.irecv(recv_bufs[d][1], d)] + ([dist.irecv(recv_bufs[d][2], d)] if rb else [])
        for r in reqs:
            r.wait()
        # what I hold now: my kept range + what arrived
        a, z = int(lo[rank]), int(hi[rank])
        ids = [b.task_ids[a:z]]
        lens = [np.diff(b.offsets[a:z + 1]).astype(np.int64)]
        pays = [b.payload[int(b.offsets[a]):int(b.offsets[z])]]
        for d in sorted(recv_bufs):
            ids.append(recv_bufs[d][0].numpy().reshape(-1, 16)); lens.append(recv_bufs[d][1].numpy()); pays.append(recv_bufs[d][2].numpy())
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=np.concatenate(ids), lens=np.concatenate(lens), payload=np.concatenate(pays),
                 before_ids=b.task_ids, before_lens=np.diff(b.offsets).astype(np.int64), before_payload=b.payload)
        dist.barrier()
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5])
