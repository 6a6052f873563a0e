// Byte-quantile rebalance plan for the pending-task ring sharded over the GPUs of one box
// (SURVEY.md §8e). Pure host arithmetic, identical on every rank given the same all-gathered
// (count, bytes) table — exported through the C ABI so that the CPU (gloo) tests and the NCCL path
// run the very same function; on the GPU path ONE device thread runs it against the device-resident byte prefix
// (rebalance_cut_kernel), so that only the 2 W cut points cross PCIe.
//
// Model: the pending tasks of all ranks form one global sequence ordered by (rank, local FIFO
// index). Rank d is to end up with the tasks whose first byte's global position falls in
// [d * B / W, (d + 1) * B / W), B = total pending payload bytes. Because a rank's tasks are
// contiguous in that sequence, what it sends to every peer is a contiguous range of its local FIFO.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define B9_PLAN_HD __host__ __device__
#else
#define B9_PLAN_HD
#endif

// prefix[i] = payload bytes of the caller's pending tasks 0..i-1 (prefix[0] = 0, prefix[n] = bytes[rank]).
// Fills send_lo/send_hi [world]: local task range [lo, hi) destined for each rank (empty ranges have lo == hi).
// Returns 0, or -1 if the table is inconsistent with `prefix`.
B9_PLAN_HD static inline int b9_plan_ranges(uint32_t world, uint32_t rank, const uint64_t* counts, const uint64_t* bytes,
                                 const uint64_t* prefix, uint64_t n, uint64_t* send_lo, uint64_t* send_hi) {
    if (world == 0 || rank >= world || counts[rank] != n || prefix[n] != bytes[rank]) return -1;
    uint64_t total = 0, base = 0;
    for (uint32_t r = 0; r < world; ++r) { if (r < rank) base += bytes[r]; total += bytes[r]; }
    // boundary[d] = first global byte position owned by rank d; ceil so that the split is exact in integers
    uint64_t lo = 0;
    for (uint32_t d = 0; d < world; ++d) {
        const unsigned __int128 num = (unsigned __int128)total * (d + 1);
        const uint64_t next_boundary = (d + 1 == world) ? UINT64_MAX : (uint64_t)((num + world - 1) / world);
        // tasks with base + prefix[i] < next_boundary (and not taken by an earlier d) go to d: binary search
        uint64_t a = lo, b = n;
        while (a < b) {
            const uint64_t m = a + (b - a) / 2;
            if (base + prefix[m] < next_boundary) a = m + 1; else b = m;
        }
        send_lo[d] = lo; send_hi[d] = a;
        lo = a;
    }
    // zero-byte tasks at the very end all have position == total and fall to the last rank: covered by UINT64_MAX
    return 0;
}
