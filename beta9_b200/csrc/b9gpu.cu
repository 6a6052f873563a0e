// libb9gpu.so — host side of the C ABI declared in include/b9gpu.h.
//
// One b9_ctx = one GPU: the pending-task ring lives in HBM (payload bytes + SoA slot arrays),
// pushes DMA the caller's batch straight into it, drains run one persistent kernel over a FIFO
// window of the ring and DMA the dense result records back. The semantic model of the ring is
// the reference's FIFO Redis list (pkg/abstractions/taskqueue/client.go:29-96: RPUSH at the tail,
// LPOP at the head) with `RingBuffer`-style fixed capacity (pkg/abstractions/common/ring_buffer.go).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <deque>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

#include "../../include/b9gpu.h"
#include "drain_kernel.cuh"
#include "drain2.cuh"
#include "rebalance_plan.h"
#include "ring_place.h"
#include "sink.h"
#include "submit_buffer.h"
#include "wire_encode.cuh"

#include <dlfcn.h>

using namespace b9;

namespace {

thread_local std::string g_last_error;   // per calling thread, as the header promises

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(B9_EIO, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

uint64_t pow2_ceil(uint64_t v) { uint64_t p = 1; while (p < v) p <<= 1; return p; }
void nccl_comm_destroy(void* comm);   // defined with the NCCL loader below

struct Segment {             // one push = one contiguous byte range of the payload ring
    uint64_t first_task;     // logical index of its first task
    uint32_t n;
    uint64_t phys_start, bytes;
    cudaEvent_t ready;           // recorded on the ingest stream once the batch is fully on the device
    uint32_t max_len = 0;        // longest task of the batch (0 = not known: tasks that arrived in a rebalance)
};

constexpr uint64_t RING_SLACK = 256;   // readable bytes past the ring end (vector loads may over-read)
constexpr uint64_t SEG_ALIGN  = B9_SEG_ALIGN;

// A few persistent threads for the host-side pack (b9_batch_push_v): the calling thread is worker 0.
struct PackPool {
    std::vector<std::thread> th;
    std::mutex m; std::condition_variable cv, cv_done;
    std::function<void(int)> job;
    int nthreads = 1, generation = 0, pending = 0; bool stop = false;
    explicit PackPool(int n) : nthreads(n < 1 ? 1 : n) {
        for (int i = 1; i < nthreads; ++i) th.emplace_back([this, i] {
            int seen = 0;
            for (;;) {
                std::function<void(int)> f;
                { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return stop || generation != seen; }); if (stop) return; seen = generation; f = job; }
                f(i);
                { std::lock_guard<std::mutex> lk(m); if (--pending == 0) cv_done.notify_one(); }
            }
        });
    }
    void run(const std::function<void(int)>& f) {
        { std::lock_guard<std::mutex> lk(m); job = f; pending = nthreads - 1; ++generation; }
        cv.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(m); cv_done.wait(lk, [&] { return pending == 0; });
    }
    ~PackPool() { { std::lock_guard<std::mutex> lk(m); stop = true; } cv.notify_all(); for (auto& t : th) t.join(); }
};

// page-locked staging of one packed batch (b9_batch_push_v): two of them alternate, so that batch k+1 is gathered
// while batch k is still on the wire
struct PackArena {
    uint8_t* payload = nullptr; uint64_t cap_bytes = 0;
    uint64_t* offsets = nullptr; uint8_t* ids = nullptr; uint32_t cap_tasks = 0;
    cudaEvent_t free_ev = nullptr; bool in_flight = false;
};

}  // namespace

struct b9_ctx {
    int device = 0;
    int sm_count = 0;
    uint32_t stage_bytes_override = 0; // B9_STAGE_BYTES: force the v2 stage-buffer size
    cudaStream_t stream = nullptr;             // drain kernels + D2H
    cudaStream_t stream_in = nullptr;          // H2D + ingest kernel (so that pushes overlap drains: PCIe is full duplex)
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr;   // drain stream (under out_mu)
    cudaEvent_t ev_pa = nullptr, ev_pb = nullptr;                                    // ingest stream (under in_mu)
    // Three locks, always taken in this order: in_mu (push submission: H2D + ingest kernel are enqueued in FIFO order),
    // out_mu (drain-side submission and the waiting-results state: launch, fetch, expire, wire records, rebalance), mu
    // (ring bookkeeping + stats; held for pointer arithmetic only, never across a CUDA synchronisation). Producers
    // therefore never wait for a drain's kernels or its D2H copy, and the drainer never waits for a push's DMA
    // (Dispatcher.Register's goroutine-safety contract, pkg/task/dispatch.go:71-73).
    std::mutex in_mu, out_mu, mu;
    std::mutex pack_mu; PackPool* pack_pool = nullptr; PackArena pack[2]; int pack_next = 0;
    // micro-batching of single-task submissions (b9_submit / b9_flush): two page-locked arenas, lock-free appends
    std::mutex sb_mu; SubmitBuffer* sb = nullptr; cudaEvent_t sb_free[2] = {nullptr, nullptr}; bool sb_in_flight[2] = {false, false};
    int sb_unpushed = -1; uint32_t sb_unpushed_n = 0;   // an arena whose push was refused (ring full): pushed again by the next flush
    std::atomic<uint64_t> running{0};           // tasks a drain has claimed (launched, not peeking) and not yet committed by a fetch

    // ---- pending ring (device)
    uint64_t ring_bytes = 0; uint32_t ring_tasks = 0, slot_mask = 0, max_task_bytes = 0;
    uint8_t*  d_payload = nullptr;
    uint64_t* d_off = nullptr; uint64_t* d_hdr = nullptr; uint4* d_ids = nullptr;
    int64_t*  d_ts = nullptr;  int64_t* d_exp = nullptr;
    // push staging (device): raw offsets + optional metadata of the batch being ingested
    uint64_t* d_in_off = nullptr; int64_t* d_in_ts = nullptr; int64_t* d_in_exp = nullptr;
    uint8_t*  d_in_retries = nullptr; uint8_t* d_in_flags = nullptr;
    // ring bookkeeping (host)
    uint64_t head_task = 0, tail_task = 0;     // logical, monotonically increasing
    uint64_t write_pos = 0;                    // physical byte position of the next segment
    uint64_t pending_bytes = 0;
    std::deque<Segment> segs;

    // ---- drain staging (device)
    uint32_t max_drain_tasks = 0; uint64_t max_result_bytes = 0;
    uint8_t* d_out_payload = nullptr; uint64_t* d_out_off = nullptr; uint4* d_out_ids = nullptr;
    uint8_t* d_out_status = nullptr; uint8_t* d_out_has = nullptr; uint32_t* d_out_len = nullptr;
    SlowItem* d_slow = nullptr;                 // identity: work list of deferred tasks (kernel tail)
    uint32_t drain_epoch = 0;                   // tag of the current launch's work-list items
    WireEnv* d_wire_env = nullptr;
    uint32_t* d_crc_shift = nullptr;            // crc32: zero-byte shift tables
    uint64_t cancelled_pending = 0;            // pending tasks carrying B9_TF_CANCELLED (pushed so, or expired)
    DrainCtl* d_ctl = nullptr; uint64_t* d_tile_state = nullptr;
    bool burst_open = false; cudaEvent_t ev_burst = nullptr;   // B9_DRAIN_ASYNC launches since the last completion: timed as one burst
    bool res_async = false;                    // the last launch returned before its kernels finished: finalize on fetch / sync
    uint32_t res_async_n = 0; uint64_t res_async_in_bytes = 0;
    DrainCtl* h_ctl = nullptr;                 // pinned
    uint8_t* d_xchg_send = nullptr; uint8_t* d_xchg_recv = nullptr; uint64_t xchg_send_cap = 0, xchg_recv_cap = 0;   // b9_rebalance staging, grow-only
    uint64_t* d_prefix = nullptr; unsigned long long* d_scan_sums = nullptr;   // b9_rebalance: byte prefix of the pending tasks, on the device
    uint64_t* d_rtab = nullptr; uint64_t* h_rtab = nullptr; size_t rtab_words = 0;   // b9_rebalance: all-gather tables + cut points (device / pinned), allocated once
    unsigned long long* d_count = nullptr; unsigned long long* h_count = nullptr;
    // results waiting on the device for b9_drain_fetch
    bool have_results = false, res_peek = false;
    uint32_t res_n = 0, res_popped = 0; uint64_t res_bytes = 0, res_in_bytes = 0; float res_kernel_ms = 0.f;
    bool expired_since_launch = false, recount_cancelled = false;   // (b9_expire between a launch and its fetch: recount, see b9_drain_fetch)

    // ---- multi-GPU (NCCL, loaded with dlopen: the library has no link-time dependency on it)
    void* nccl_comm = nullptr; int comm_rank = 0, comm_world = 1;

    b9_stats stats{};
};

namespace {

void free_segments(b9_ctx* c) {
    while (!c->segs.empty() && c->segs.front().first_task + c->segs.front().n <= c->head_task) {
        c->event_pool.push_back(c->segs.front().ready);
        c->segs.pop_front();
    }
    if (c->segs.empty()) c->write_pos = 0;
}

// Physical payload offset of logical task `idx` of segment `sg` (idx == first_task + n gives the segment's
// end). Only partial-segment operations need it (a drain that stops inside a batch, the rebalance), so
// the host keeps no per-task index: the slot ring on the device is the index. Synchronous, 8 bytes.
int task_phys_off(b9_ctx* c, const Segment& sg, uint64_t idx, uint64_t* out) {
    if (idx == sg.first_task) { *out = sg.phys_start; return B9_OK; }
    if (idx == sg.first_task + sg.n) { *out = sg.phys_start + sg.bytes; return B9_OK; }
    CU(cudaStreamWaitEvent(c->stream, sg.ready, 0));
    CU(cudaMemcpyAsync(&c->h_ctl->total, c->d_off + (uint32_t)(idx & c->slot_mask), sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *out = c->h_ctl->total;
    return B9_OK;
}

// where can a segment of `bytes` go? returns false when the ring cannot take it now (ring_place.h)
bool place_segment(b9_ctx* c, uint64_t bytes, uint64_t* start) {
    uint64_t live = 0;
    for (const Segment& sg : c->segs) live += sg.bytes;
    return b9_ring_place(c->ring_bytes, !c->segs.empty(), c->segs.empty() ? 0 : c->segs.front().phys_start, c->write_pos, live, bytes, start) != 0;
}

// every warp owns a slice of dynamic shared memory = control block + one stage buffer.
template <int H> uint32_t drain3_warp_stride(uint32_t in_cap) {
    const size_t ctl = (sizeof(D3Warp<D3Cfg<H>::T>) + 127u) & ~(size_t)127u;
    return (uint32_t)(ctl + (((size_t)in_cap + 64u + 127u) & ~(size_t)127u));
}
template <int H> cudaError_t launch_drain3(DrainArgs a, uint64_t avg_task_bytes, uint32_t max_task_len, uint32_t cap_override, int sm_count, cudaStream_t s, int* grid_out) {
    constexpr int T = D3Cfg<H>::T;
    static const bool tight = !(getenv("B9_STAGE_TIGHT") && atoi(getenv("B9_STAGE_TIGHT")) == 0);      // (=0: the A/B switch)
    uint32_t in_cap = 0, stride = 0; size_t smem = 0; int per_sm = 0;
    // stage buffer per warp sized for the window's average warp-tile: tt tasks x 1.1 + 768 B, within [1 KiB, 48 KiB].
    // Tiles that do not fit are processed straight from global memory (same code, generic loads).
    // (crc32's configuration has long-tailed sizes - zipf strings - and gets a third of a tile of headroom: more would
    // cost resident warps, and its byte loop lives on those - 16 CTAs per SM fit with this, see B9_CRC_MINB)
    auto size_for = [&](uint32_t tt) -> cudaError_t {
        const uint64_t want = (H == B9_H_CRC32) ? avg_task_bytes * tt * 27 / 20 + 768 : avg_task_bytes * tt * 11 / 10 + 768;
        in_cap = (uint32_t)std::min<uint64_t>(48u << 10, std::max<uint64_t>(1u << 10, want));
        // no tile of the window is longer than tt of its longest tasks (+ the alignment of its first byte): with uniform task
        // sizes (configs[3]: every task 372 bytes) the headroom above buys nothing and costs a resident CTA per SM
        if (tight && H != B9_H_CRC32 && max_task_len) in_cap = (uint32_t)std::min<uint64_t>(in_cap, std::max<uint64_t>(1u << 10, (uint64_t)max_task_len * tt + 32u));
        in_cap = (in_cap + 127u) & ~127u;
        if (cap_override) in_cap = cap_override;
        stride = drain3_warp_stride<H>(in_cap);
        smem = (size_t)stride * D3_WARPS;
        cudaError_t e = cudaFuncSetAttribute(drain3_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, drain3_kernel<H>, D3_WARPS * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        return cudaSuccess;
    };
    uint32_t tt = (uint32_t)T;
    cudaError_t e = size_for(tt);
    if (e != cudaSuccess) return e;
    // json_sum works on a document with the whole warp: a window of about one 8-document tile per resident warp ends in a
    // ragged last round (configs[4]: 25 k documents per GPU over ~2960 warps); tiles of 4 balance it - 0.087 against
    // 0.100 ms - while larger windows prefer 8 (0.218 against 0.227 ms at 100 k, profiles/r2_j2_*)
    if (H == B9_H_JSON_SUM && T >= 8 && (uint64_t)(a.n_tasks + T - 1) / T < 3ull * (uint64_t)per_sm * (uint64_t)sm_count * D3_WARPS) {
        tt = 4u;
        e = size_for(tt);
        if (e != cudaSuccess) return e;
    }
    a.tile_tasks = tt;
    a.n_tiles = (a.n_tasks + tt - 1) / tt;
    const uint32_t ctas_needed = (a.n_tiles + D3_WARPS - 1) / D3_WARPS;
    const int grid = (int)std::min<uint32_t>(ctas_needed, (uint32_t)(per_sm * sm_count));
    *grid_out = grid;
    // identity with nothing cancelled in the window has a uniform tile cost (its expensive tasks are deferred):
    // all but the last round of a worker's tiles are assigned statically, measured 0.1915 -> 0.182 ms; the same
    // switch made vadd_f32 slower (0.196 -> 0.212 ms), so it stays on pure work stealing like crc32 / json_sum.
    // B9_STATIC_ROUNDS=0 turns it off.
    static const bool allow_static = !(getenv("B9_STATIC_ROUNDS") && atoi(getenv("B9_STATIC_ROUNDS")) == 0);
    if (a.count_mode) {                                                    // cancelled slots in the window: ready-count prefix per warp-tile
        const uint32_t blocks = (a.n_tasks + TC_SLOTS - 1u) / TC_SLOTS;
        tile_count_kernel<<<blocks, TC_SLOTS, 0, s>>>(a.hdr, a.slot_mask, a.first_task, a.n_tasks, tt, (uint32_t*)a.tile_base, (uint32_t*)a.block_base);
        tile_scan_kernel<<<1, 1024, 0, s>>>((uint32_t*)a.block_base, blocks);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (allow_static && H == B9_H_IDENTITY) {
        // B9_DYNAMIC_ROUNDS: how many of a worker's last rounds come from the ticket counter. One is enough when every
        // tile costs the same; with 1 % of escape-dense strings a tile that holds one takes ~25 % longer (the warp checks its
        // escapes), a worker's static share can hold a dozen of them, and one round of slack does not absorb that: 3 rounds
        // measured 6.47 against 6.29 G tasks/s on the mix and 7.46 against 7.47 on clean input (profiles/r2_d1_*, r2_d2_*)
        static const uint64_t dyn = getenv("B9_DYNAMIC_ROUNDS") ? (uint64_t)std::max(1, atoi(getenv("B9_DYNAMIC_ROUNDS"))) : 3u;
        const uint64_t rounds = a.n_tiles / ((uint64_t)grid * D3_WARPS);
        a.static_rounds = rounds > dyn ? (uint32_t)(rounds - dyn) : 0u;
    }
    drain3_kernel<H><<<grid, D3_WARPS * 32, smem, s>>>(a, in_cap, stride);
    return cudaGetLastError();
}

__global__ void count_cancelled_kernel(const uint64_t* __restrict__ hdr, uint32_t slot_mask, uint64_t first_task, uint32_t n, unsigned long long* __restrict__ count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool hit = i < n && (hdr_flags(hdr[(uint32_t)((first_task + i) & slot_mask)]) & 1u);
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31u) == 0u && m) atomicAdd(count, (unsigned long long)__popc(m));
}

// cancelled_pending := the cancelled slots among the pending tasks, counted on the device. Caller holds c->out_mu (not c->mu).
int recount_cancelled_locked_out(b9_ctx* c) {
    uint64_t head = 0, depth = 0;
    cudaStream_t s = c->stream;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        head = c->head_task; depth = c->tail_task - c->head_task;
        for (const Segment& sg : c->segs) CU(cudaStreamWaitEvent(s, sg.ready, 0));
    }
    CU(cudaMemsetAsync(c->d_count, 0, sizeof(unsigned long long), s));
    if (depth) { count_cancelled_kernel<<<(uint32_t)((depth + 255) / 256), 256, 0, s>>>(c->d_hdr, c->slot_mask, head, (uint32_t)depth, c->d_count); CU(cudaGetLastError()); }
    CU(cudaMemcpyAsync(c->h_count, c->d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    std::lock_guard<std::mutex> lk(c->mu);
    // tasks pushed (with B9_TF_CANCELLED) after the snapshot are in cancelled_pending already but not in the count: keep them
    const uint64_t depth_now = c->tail_task - c->head_task;
    if (depth_now == depth) c->cancelled_pending = *c->h_count;
    else c->cancelled_pending = std::max<uint64_t>(c->cancelled_pending, *c->h_count);   // (an over-count only costs the count pre-pass)
    c->recount_cancelled = false;
    if (depth) c->stats.kernel_launches++;
    return B9_OK;
}

}  // namespace

extern "C" {

uint32_t b9_abi_version(void) { return B9_ABI_VERSION; }

int b9_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* b9_last_error(const b9_ctx*) { return g_last_error.c_str(); }

static const char* const HANDLER_NAMES[B9_H_COUNT_] = {"identity", "crc32", "vadd_f32", "json_sum"};
const char* b9_handler_name(int h) { return (h >= 0 && h < B9_H_COUNT_) ? HANDLER_NAMES[h] : nullptr; }
int b9_handler_id(const char* name) {
    if (!name) return B9_EINVAL;
    if (!strcmp(name, "echo")) return B9_H_IDENTITY;
    for (int h = 0; h < B9_H_COUNT_; ++h) if (!strcmp(name, HANDLER_NAMES[h])) return h;
    return B9_ENOSYS;
}

int b9_ctx_create(const b9_opts* opts, b9_ctx** out) {
    if (!out) return fail(B9_EINVAL, "b9_ctx_create: out is NULL");
    *out = nullptr;
    b9_opts o{};
    if (opts) memcpy(&o, opts, std::min<size_t>(sizeof o, opts->struct_size ? opts->struct_size : sizeof o));
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(B9_ENODEV, "b9_ctx_create: no CUDA device is visible; libb9gpu has no CPU path");
    }
    if (o.device < 0 || o.device >= ndev) return fail(B9_EINVAL, "b9_ctx_create: device %d of %d", o.device, ndev);
    b9_ctx* c = new (std::nothrow) b9_ctx();
    if (!c) return fail(B9_ENOMEM, "b9_ctx_create: host allocation failed");
    c->device = o.device;
    c->ring_bytes = pow2_ceil(o.ring_bytes ? o.ring_bytes : (1ull << 30));
    c->ring_tasks = (uint32_t)pow2_ceil(o.ring_tasks ? o.ring_tasks : (4u << 20));
    c->slot_mask = c->ring_tasks - 1;
    c->max_drain_tasks = o.max_drain_tasks ? o.max_drain_tasks : (2u << 20);
    if (c->max_drain_tasks >= (1u << 24)) c->max_drain_tasks = (1u << 24) - 1;
    if (c->max_drain_tasks > c->ring_tasks) c->max_drain_tasks = c->ring_tasks;
    c->max_result_bytes = o.max_result_bytes ? o.max_result_bytes : (1ull << 30);
    if (c->max_result_bytes >= (1ull << 38)) c->max_result_bytes = (1ull << 38) - 1;
    c->max_task_bytes = o.max_task_bytes ? std::min<uint32_t>(o.max_task_bytes, (1u << 30) - 1u) : (1u << 20);   // (two bits of a 32-bit length word carry flags)

#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    int rc_ = fail(e_ == cudaErrorMemoryAllocation ? B9_ENOMEM : B9_EIO, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    b9_ctx_destroy(c); return rc_; } } while (0)
    CUC(cudaSetDevice(c->device));
    cudaDeviceProp prop{};
    CUC(cudaGetDeviceProperties(&prop, c->device));
    c->sm_count = prop.multiProcessorCount;
    CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&c->stream_in, cudaStreamNonBlocking));
    CUC(cudaEventCreate(&c->ev_burst));
    CUC(cudaEventCreate(&c->ev_a)); CUC(cudaEventCreate(&c->ev_b)); CUC(cudaEventCreate(&c->ev_c)); CUC(cudaEventCreate(&c->ev_d));
    CUC(cudaEventCreate(&c->ev_pa)); CUC(cudaEventCreate(&c->ev_pb));
    const uint32_t rt = c->ring_tasks, md = c->max_drain_tasks;
    CUC(cudaMalloc(&c->d_payload, c->ring_bytes + RING_SLACK));
    CUC(cudaMemset(c->d_payload + c->ring_bytes, 0, RING_SLACK));
    CUC(cudaMalloc(&c->d_off, rt * sizeof(uint64_t)));
    CUC(cudaMalloc(&c->d_hdr, rt * sizeof(uint64_t)));
    CUC(cudaMalloc(&c->d_ids, rt * sizeof(uint4)));
    CUC(cudaMalloc(&c->d_ts, rt * sizeof(int64_t)));
    CUC(cudaMalloc(&c->d_exp, rt * sizeof(int64_t)));
    CUC(cudaMalloc(&c->d_in_off, ((size_t)rt + 1) * sizeof(uint64_t)));
    CUC(cudaMalloc(&c->d_in_ts, rt * sizeof(int64_t)));
    CUC(cudaMalloc(&c->d_in_exp, rt * sizeof(int64_t)));
    CUC(cudaMalloc(&c->d_in_retries, rt));
    CUC(cudaMalloc(&c->d_in_flags, rt));
    CUC(cudaMalloc(&c->d_out_payload, c->max_result_bytes + RING_SLACK));
    CUC(cudaMalloc(&c->d_out_off, ((size_t)md + 1) * sizeof(uint64_t)));
    CUC(cudaMalloc(&c->d_out_ids, (size_t)md * sizeof(uint4)));
    CUC(cudaMalloc(&c->d_out_status, md));
    CUC(cudaMalloc(&c->d_out_has, md));
    CUC(cudaMalloc(&c->d_out_len, (size_t)md * sizeof(uint32_t)));
    CUC(cudaMalloc(&c->d_slow, (size_t)md * sizeof(SlowItem)));
    CUC(cudaMemset(c->d_slow, 0, (size_t)md * sizeof(SlowItem)));
    CUC(cudaMalloc(&c->d_wire_env, sizeof(WireEnv)));
    {
        // Z_k = advance the reflected CRC-32 register over 2^k zero bytes, as 4 byte-indexed tables per k
        std::vector<uint32_t> tabs((size_t)CRC_SHIFT_LEVELS * 1024);
        uint32_t step[256];
        for (uint32_t i = 0; i < 256; ++i) { uint32_t v = i; for (int b = 0; b < 8; ++b) v = (v & 1u) ? (0xEDB88320u ^ (v >> 1)) : (v >> 1); step[i] = v; }
        auto apply = [&](const uint32_t* t, uint32_t v) { return t[v & 255] ^ t[256 + ((v >> 8) & 255)] ^ t[512 + ((v >> 16) & 255)] ^ t[768 + (v >> 24)]; };
        for (uint32_t j = 0; j < 4; ++j) for (uint32_t b = 0; b < 256; ++b) { uint32_t v = b << (8 * j); tabs[j * 256 + b] = step[v & 255] ^ (v >> 8); }   // k = 0: one zero byte
        for (int k = 1; k < CRC_SHIFT_LEVELS; ++k) {
            const uint32_t* prev = tabs.data() + (size_t)(k - 1) * 1024; uint32_t* cur = tabs.data() + (size_t)k * 1024;
            for (uint32_t j = 0; j < 4; ++j) for (uint32_t b = 0; b < 256; ++b) cur[j * 256 + b] = apply(prev, apply(prev, b << (8 * j)));
        }
        CUC(cudaMalloc(&c->d_crc_shift, tabs.size() * sizeof(uint32_t)));
        CUC(cudaMemcpy(c->d_crc_shift, tabs.data(), tabs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    CUC(cudaMalloc(&c->d_ctl, sizeof(DrainCtl)));
    CUC(cudaMalloc(&c->d_tile_state, ((size_t)md / D2_THREADS + 2) * sizeof(uint64_t)));
    CUC(cudaMalloc(&c->d_count, sizeof(unsigned long long)));
    CUC(cudaHostAlloc(&c->h_ctl, sizeof(DrainCtl), cudaHostAllocDefault));
    CUC(cudaHostAlloc(&c->h_count, sizeof(unsigned long long), cudaHostAllocDefault));
    c->stats.sm_count = (uint32_t)c->sm_count;
    if (const char* v = getenv("B9_STAGE_BYTES")) c->stage_bytes_override = ((uint32_t)atoi(v) + 1023u) & ~1023u;
#undef CUC
    *out = c;
    return B9_OK;
}

void b9_ctx_destroy(b9_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->stream_in) cudaStreamSynchronize(c->stream_in);
    if (c->nccl_comm) nccl_comm_destroy(c->nccl_comm);
    for (Segment& sg : c->segs) cudaEventDestroy(sg.ready);
    for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
    cudaFree(c->d_payload); cudaFree(c->d_off); cudaFree(c->d_hdr); cudaFree(c->d_ids); cudaFree(c->d_ts); cudaFree(c->d_exp);
    cudaFree(c->d_in_off); cudaFree(c->d_in_ts); cudaFree(c->d_in_exp); cudaFree(c->d_in_retries); cudaFree(c->d_in_flags);
    cudaFree(c->d_out_payload); cudaFree(c->d_out_off); cudaFree(c->d_out_ids); cudaFree(c->d_out_status); cudaFree(c->d_out_has); cudaFree(c->d_out_len); cudaFree(c->d_slow); cudaFree(c->d_wire_env); cudaFree(c->d_crc_shift);
    cudaFree(c->d_ctl); cudaFree(c->d_tile_state); cudaFree(c->d_count);
    if (c->h_ctl) cudaFreeHost(c->h_ctl);
    if (c->h_rtab) cudaFreeHost(c->h_rtab);
    cudaFree(c->d_rtab); cudaFree(c->d_prefix); cudaFree(c->d_scan_sums);
    cudaFree(c->d_xchg_send); cudaFree(c->d_xchg_recv);
    if (c->h_count) cudaFreeHost(c->h_count);
    if (c->ev_burst) cudaEventDestroy(c->ev_burst);
    if (c->ev_a) cudaEventDestroy(c->ev_a);
    if (c->ev_b) cudaEventDestroy(c->ev_b);
    if (c->ev_c) cudaEventDestroy(c->ev_c);
    if (c->ev_d) cudaEventDestroy(c->ev_d);
    if (c->ev_pa) cudaEventDestroy(c->ev_pa);
    if (c->ev_pb) cudaEventDestroy(c->ev_pb);
    delete c->pack_pool;
    if (c->sb) {
        for (int k = 0; k < 2; ++k) {
            SubmitArena& A = c->sb->arena[k];
            if (A.payload) cudaFreeHost(A.payload);
            if (A.offsets) cudaFreeHost(A.offsets);
            if (A.ids) cudaFreeHost(A.ids);
            if (A.flags) cudaFreeHost(A.flags);
            if (c->sb_free[k]) cudaEventDestroy(c->sb_free[k]);
        }
        delete c->sb;
    }
    for (PackArena& A : c->pack) {
        if (A.payload) cudaFreeHost(A.payload);
        if (A.offsets) cudaFreeHost(A.offsets);
        if (A.ids) cudaFreeHost(A.ids);
        if (A.free_ev) cudaEventDestroy(A.free_ev);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->stream_in) cudaStreamDestroy(c->stream_in);
    delete c;
}

void* b9_host_alloc(b9_ctx* c, uint64_t bytes) {
    if (!c) { fail(B9_EINVAL, "b9_host_alloc: ctx is NULL"); return nullptr; }
    void* p = nullptr;
    cudaSetDevice(c->device);
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { fail(B9_ENOMEM, "cudaHostAlloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e)); return nullptr; }
    return p;
}

void b9_host_free(b9_ctx* c, void* p) {
    if (!p) return;
    if (c) cudaSetDevice(c->device);
    cudaFreeHost(p);
}

// One push. Reservation and commit of the ring space are two short critical sections under c->mu; the DMA and the
// ingest kernel are enqueued between them under c->in_mu only, so a drain can run (and be synchronised) meanwhile.
// The batch becomes visible to drains (tail_task) only at the commit, after its `ready` event has been recorded.
// `after`: optional event recorded once the copies are enqueued (b9_batch_push_v's arena reuse).
static int push_impl(b9_ctx* c, const uint8_t* task_ids, const uint8_t* payload, const uint64_t* offsets, uint32_t n, const b9_push_meta* meta, bool wait,
                     cudaEvent_t after = nullptr) {
    if (!c) return fail(B9_EINVAL, "b9_batch_push: ctx is NULL");
    if (n == 0) return B9_OK;
    if (!task_ids || !payload || !offsets) return fail(B9_EINVAL, "b9_batch_push: NULL buffer");
    if (offsets[n] < offsets[0]) return fail(B9_EINVAL, "b9_batch_push: offsets not monotonic");
    const uint64_t bytes = offsets[n] - offsets[0];
    std::lock_guard<std::mutex> in_lk(c->in_mu);
    CU(cudaSetDevice(c->device));
    uint64_t start = 0, tail = 0;
    cudaEvent_t ready = nullptr;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        free_segments(c);
        if ((uint64_t)n > (uint64_t)c->ring_tasks - (c->tail_task - c->head_task))
            return fail(B9_ENOSPC, "b9_batch_push: %u tasks do not fit (%llu of %u slots pending)", n,
                        (unsigned long long)(c->tail_task - c->head_task), c->ring_tasks);
        if (!place_segment(c, bytes, &start))
            return fail(B9_ENOSPC, "b9_batch_push: %llu payload bytes do not fit the ring (%llu pending of %llu)",
                        (unsigned long long)bytes, (unsigned long long)c->pending_bytes, (unsigned long long)c->ring_bytes);
        if (!c->event_pool.empty()) { ready = c->event_pool.back(); c->event_pool.pop_back(); }
        tail = c->tail_task;
    }
    if (!ready) CU(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    auto give_back = [&]() { std::lock_guard<std::mutex> lk(c->mu); c->event_pool.push_back(ready); };
    cudaStream_t s = c->stream_in;
    CU(cudaEventRecord(c->ev_pa, s));
    // the payload DMA goes first (into ring space that stays free until this push is committed), so that
    // the O(n) host-side validation of the index below runs while the bytes are already on the wire
    if (bytes) CU(cudaMemcpyAsync(c->d_payload + start, payload + offsets[0], bytes, cudaMemcpyHostToDevice, s));
    uint64_t n_cancelled = 0;
    uint32_t seg_max_len = 0;
    {
        const uint32_t maxb = c->max_task_bytes;
        uint64_t bad = ~0ull, big = ~0ull, prev = offsets[0], longest = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const uint64_t cur = offsets[i + 1];
            if (cur < prev) { if (bad == ~0ull) bad = i; }
            else { if (cur - prev > maxb && big == ~0ull) big = i; longest = std::max(longest, cur - prev); }
            prev = cur;
        }
        seg_max_len = (uint32_t)std::min<uint64_t>(longest, maxb);
        if (bad != ~0ull || big != ~0ull) {
            give_back();
            cudaStreamSynchronize(s);            // the stray DMA must not outlive the caller's buffer
            if (bad != ~0ull) return fail(B9_EINVAL, "b9_batch_push: offsets not monotonic at task %llu", (unsigned long long)bad);
            return fail(B9_E2BIG, "b9_batch_push: task %llu is %llu bytes, max_task_bytes is %u", (unsigned long long)big,
                        (unsigned long long)(offsets[big + 1] - offsets[big]), maxb);
        }
        if (meta && meta->flags) for (uint32_t i = 0; i < n; ++i) n_cancelled += (meta->flags[i] & B9_TF_CANCELLED) ? 1 : 0;
    }
    CU(cudaMemcpyAsync(c->d_in_off, offsets, ((size_t)n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    // ids go straight into their ring slots (two pieces when the slot ring wraps)
    const uint32_t slot0 = (uint32_t)(tail & c->slot_mask);
    const uint32_t first = std::min<uint32_t>(n, c->ring_tasks - slot0);
    CU(cudaMemcpyAsync(c->d_ids + slot0, task_ids, (size_t)first * 16, cudaMemcpyHostToDevice, s));
    if (first < n) CU(cudaMemcpyAsync(c->d_ids, task_ids + (size_t)first * 16, (size_t)(n - first) * 16, cudaMemcpyHostToDevice, s));
    const int64_t* ts = nullptr; const int64_t* ex = nullptr; const uint8_t* rt = nullptr; const uint8_t* fl = nullptr;
    uint64_t meta_bytes = 0;
    if (meta) {
        if (meta->timestamp_unix)  { CU(cudaMemcpyAsync(c->d_in_ts, meta->timestamp_unix, (size_t)n * 8, cudaMemcpyHostToDevice, s)); ts = c->d_in_ts; meta_bytes += (uint64_t)n * 8; }
        if (meta->expires_unix_ns) { CU(cudaMemcpyAsync(c->d_in_exp, meta->expires_unix_ns, (size_t)n * 8, cudaMemcpyHostToDevice, s)); ex = c->d_in_exp; meta_bytes += (uint64_t)n * 8; }
        if (meta->retries)         { CU(cudaMemcpyAsync(c->d_in_retries, meta->retries, n, cudaMemcpyHostToDevice, s)); rt = c->d_in_retries; meta_bytes += n; }
        if (meta->flags)           { CU(cudaMemcpyAsync(c->d_in_flags, meta->flags, n, cudaMemcpyHostToDevice, s)); fl = c->d_in_flags; meta_bytes += n; }
    }
    CU(cudaEventRecord(c->ev_pb, s));
    if (after) CU(cudaEventRecord(after, s));
    ingest_kernel<<<(n + 255) / 256, 256, 0, s>>>(c->d_in_off, n, start, tail, c->slot_mask, ts, ex, rt, fl,
                                                  c->d_off, c->d_hdr, c->d_ts, c->d_exp);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ready, s));
    {
        std::lock_guard<std::mutex> lk(c->mu);
        c->segs.push_back(Segment{tail, n, start, bytes, ready, seg_max_len});
        c->write_pos = start + b9_seg_span(bytes);
        c->tail_task += n;
        c->pending_bytes += bytes;
        c->cancelled_pending += n_cancelled;
        c->stats.kernel_launches++;
        c->stats.tasks_pushed += n;
        c->stats.bytes_h2d += bytes + ((uint64_t)n + 1) * 8 + (uint64_t)n * 16 + meta_bytes;
    }
    if (wait) {
        CU(cudaStreamSynchronize(s));            // the caller may reuse its buffers as soon as we return
        float ms = 0; cudaEventElapsedTime(&ms, c->ev_pa, c->ev_pb);
        std::lock_guard<std::mutex> lk(c->mu); c->stats.last_push_h2d_ms = ms;
    }
    return B9_OK;
}

int b9_batch_push(b9_ctx* c, const uint8_t* task_ids, const uint8_t* payload, const uint64_t* offsets, uint32_t n, const b9_push_meta* meta) {
    return push_impl(c, task_ids, payload, offsets, n, meta, true);
}

int b9_batch_push_async(b9_ctx* c, const uint8_t* task_ids, const uint8_t* payload, const uint64_t* offsets, uint32_t n, const b9_push_meta* meta) {
    return push_impl(c, task_ids, payload, offsets, n, meta, false);
}

// The pack step of the north star ("packs the whole pending batch into pinned host buffers"): n payloads scattered over
// the caller's (pageable) memory — Go's [][]byte — are gathered by a few threads into one of two page-locked arenas of
// the context, the n + 1 offsets are built on the way, and the arena is pushed asynchronously. Everything the call
// needs has been copied when it returns (cgo pointer rules); the arena is reused two pushes later, after its DMA.
int b9_batch_push_v(b9_ctx* c, const uint8_t* task_ids, const uint8_t* const* payloads, const uint32_t* lengths, uint32_t n, const b9_push_meta* meta) {
    if (!c) return fail(B9_EINVAL, "b9_batch_push_v: ctx is NULL");
    if (n == 0) return B9_OK;
    if (!task_ids || !payloads || !lengths) return fail(B9_EINVAL, "b9_batch_push_v: NULL buffer");
    std::lock_guard<std::mutex> pk(c->pack_mu);
    CU(cudaSetDevice(c->device));
    if (!c->pack_pool) {
        int nt = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
        if (const char* e = getenv("B9_PACK_THREADS")) nt = std::max(1, atoi(e));
        c->pack_pool = new (std::nothrow) PackPool(nt);
        if (!c->pack_pool) return fail(B9_ENOMEM, "b9_batch_push_v: thread pool");
    }
    const int T = c->pack_pool->nthreads;
    // pass 1: bytes per thread range (ranges of equal task count)
    std::vector<uint64_t> part((size_t)T + 1, 0);
    auto lo_of = [&](int t) { return (uint32_t)((uint64_t)n * (uint64_t)t / (uint64_t)T); };
    c->pack_pool->run([&](int t) {
        uint64_t sum = 0; const uint32_t lo = lo_of(t), hi = lo_of(t + 1);
        for (uint32_t i = lo; i < hi; ++i) sum += lengths[i];
        part[(size_t)t + 1] = sum;
    });
    for (int t = 0; t < T; ++t) part[(size_t)t + 1] += part[(size_t)t];
    const uint64_t total = part[(size_t)T];
    PackArena& A = c->pack[c->pack_next];
    c->pack_next ^= 1;
    if (A.in_flight) { CU(cudaEventSynchronize(A.free_ev)); A.in_flight = false; }   // its previous batch is on the device
    if (!A.free_ev) CU(cudaEventCreateWithFlags(&A.free_ev, cudaEventDisableTiming));
    if (A.cap_bytes < total + 64) {
        if (A.payload) cudaFreeHost(A.payload);
        A.payload = nullptr; A.cap_bytes = 0;
        const uint64_t want = (total + 64) + (total + 64) / 4;
        if (cudaHostAlloc(&A.payload, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return fail(B9_ENOMEM, "b9_batch_push_v: %llu bytes of page-locked memory", (unsigned long long)want); }
        A.cap_bytes = want;
    }
    if (A.cap_tasks < n) {
        if (A.offsets) cudaFreeHost(A.offsets);
        if (A.ids) cudaFreeHost(A.ids);
        A.offsets = nullptr; A.ids = nullptr; A.cap_tasks = 0;
        const uint32_t want = n + n / 4 + 16;
        if (cudaHostAlloc(&A.offsets, ((size_t)want + 1) * 8, cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc(&A.ids, (size_t)want * 16, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return fail(B9_ENOMEM, "b9_batch_push_v: page-locked index"); }
        A.cap_tasks = want;
    }
    // pass 2: gather
    c->pack_pool->run([&](int t) {
        const uint32_t lo = lo_of(t), hi = lo_of(t + 1);
        uint64_t o = part[(size_t)t];
        // the payloads sit anywhere in pageable memory: without a hint every task starts with a TLB miss and a chain of
        // cache misses the copy then waits for; the first lines of the task PF ahead are requested while this one is copied
        constexpr uint32_t PF = 8;
        for (uint32_t i = lo; i < hi; ++i) {
            if (i + PF < hi) {
                const uint8_t* q = payloads[i + PF];
                const uint32_t ql = lengths[i + PF];
                for (uint32_t b = 0; b < ql && b < 512u; b += 64u) __builtin_prefetch(q + b, 0, 0);
            }
            A.offsets[i] = o;
            const uint32_t l = lengths[i];
            if (l) memcpy(A.payload + o, payloads[i], l);
            o += l;
        }
        memcpy(A.ids + (size_t)lo * 16, task_ids + (size_t)lo * 16, (size_t)(hi - lo) * 16);
    });
    A.offsets[n] = total;
    const int rc = push_impl(c, A.ids, A.payload, A.offsets, n, meta, /*wait=*/false, A.free_ev);
    if (rc == B9_OK) A.in_flight = true;
    return rc;
}

// ---- single-task submissions from many threads, batched under the hood (submit_buffer.h) ------------------------------
static int sb_ensure(b9_ctx* c) {
    std::lock_guard<std::mutex> lk(c->sb_mu);
    if (c->sb) return B9_OK;
    CU(cudaSetDevice(c->device));
    uint64_t bytes = 64ull << 20; uint32_t tasks = 256u << 10;
    if (const char* e = getenv("B9_SUBMIT_BYTES")) bytes = (uint64_t)std::max<long long>(4096, atoll(e));
    if (const char* e = getenv("B9_SUBMIT_TASKS")) tasks = (uint32_t)std::max<long long>(16, atoll(e));
    tasks = std::min<uint32_t>(tasks, c->ring_tasks);
    bytes = std::min<uint64_t>(bytes, c->ring_bytes / 2);
    SubmitBuffer* sb = new (std::nothrow) SubmitBuffer();
    if (!sb) return fail(B9_ENOMEM, "b9_submit: buffer");
    for (int k = 0; k < 2; ++k) {
        SubmitArena& A = sb->arena[k];
        if (cudaHostAlloc(&A.payload, bytes, cudaHostAllocDefault) != cudaSuccess || cudaHostAlloc(&A.offsets, ((size_t)tasks + 1) * 8, cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc(&A.ids, (size_t)tasks * 16, cudaHostAllocDefault) != cudaSuccess || cudaHostAlloc(&A.flags, tasks, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return fail(B9_ENOMEM, "b9_submit: %llu bytes of page-locked memory per arena", (unsigned long long)bytes);
        }
        A.cap_bytes = bytes; A.cap_tasks = tasks;
        CU(cudaEventCreateWithFlags(&c->sb_free[k], cudaEventDisableTiming));
    }
    c->sb = sb;
    return B9_OK;
}

static int sb_push(b9_ctx* c, int k, uint32_t n) {
    SubmitArena& A = c->sb->arena[k];
    b9_push_meta meta{}; meta.flags = A.any_flags.load(std::memory_order_relaxed) ? A.flags : nullptr;
    const int rc = push_impl(c, A.ids, A.payload, A.offsets, n, meta.flags ? &meta : nullptr, /*wait=*/false, c->sb_free[k]);
    if (rc == B9_OK) { c->sb_in_flight[k] = true; c->sb_unpushed = -1; }
    else { c->sb_unpushed = k; c->sb_unpushed_n = n; }                       // nothing is lost: the next flush pushes it again
    return rc;
}

int64_t b9_flush(b9_ctx* c) {
    if (!c) return fail(B9_EINVAL, "b9_flush: ctx is NULL");
    if (!c->sb) return 0;
    SubmitBuffer* sb = c->sb;
    std::lock_guard<std::mutex> lk(sb->flush_mu);
    CU(cudaSetDevice(c->device));
    int64_t pushed = 0;
    if (c->sb_unpushed >= 0) {                                               // an earlier batch the ring refused
        const uint32_t n = c->sb_unpushed_n;
        const int rc = sb_push(c, c->sb_unpushed, n);
        if (rc) return rc;
        pushed += n;
    }
    const int cur = sb->active.load(std::memory_order_acquire), other = cur ^ 1;
    if (sb->buffered() == 0) return pushed;
    // the other arena takes the submissions from now on: its previous batch must be on the device
    if (c->sb_in_flight[other]) { CU(cudaEventSynchronize(c->sb_free[other])); c->sb_in_flight[other] = false; }
    SubmitBuffer::reset(sb->arena[other]);
    uint32_t n = 0; uint64_t bytes = 0;
    sb->seal(&n, &bytes);
    if (n == 0) return pushed;
    const int rc = sb_push(c, cur, n);
    if (rc) return rc;
    return pushed + n;
}

int b9_submit(b9_ctx* c, const uint8_t* task_id, const uint8_t* payload, uint32_t length, uint8_t flags) {
    if (!c || !task_id || (!payload && length)) return fail(B9_EINVAL, "b9_submit: NULL argument");
    if (length > c->max_task_bytes) return fail(B9_E2BIG, "b9_submit: task is %u bytes, max_task_bytes is %u", length, c->max_task_bytes);
    if (!c->sb) { const int rc = sb_ensure(c); if (rc) return rc; }
    for (int attempt = 0; attempt < 64; ++attempt) {
        const int r = c->sb->submit(task_id, payload, length, flags);
        if (r == SUBMIT_OK) return B9_OK;
        if (r == SUBMIT_TOO_BIG) return fail(B9_E2BIG, "b9_submit: task is %u bytes, the submission arena holds %llu", length, (unsigned long long)c->sb->arena[0].cap_bytes);
        const int64_t f = b9_flush(c);                                       // the arena is full: hand it to the device, go on in the other one
        if (f < 0) return (int)f;
    }
    return fail(B9_ENOSPC, "b9_submit: the submission arenas stay full");
}

uint64_t b9_buffered(b9_ctx* c) {
    if (!c || !c->sb) return 0;
    return (uint64_t)c->sb->buffered() + (c->sb_unpushed >= 0 ? c->sb_unpushed_n : 0u);
}

uint64_t b9_depth(b9_ctx* c) {
    if (!c) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    return c->tail_task - c->head_task;
}

uint64_t b9_depth_bytes(b9_ctx* c) {
    if (!c) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    return c->pending_bytes;
}

uint64_t b9_running(b9_ctx* c) { return c ? c->running.load(std::memory_order_relaxed) : 0; }

int64_t b9_expire(b9_ctx* c, int64_t now_unix_ns) {
    if (!c) return fail(B9_EINVAL, "b9_expire: ctx is NULL");
    std::lock_guard<std::mutex> out_lk(c->out_mu);
    CU(cudaSetDevice(c->device));
    uint64_t head = 0, depth = 0;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        head = c->head_task; depth = c->tail_task - c->head_task;
        if (!depth) return 0;
        for (const Segment& sg : c->segs) CU(cudaStreamWaitEvent(c->stream, sg.ready, 0));
    }
    CU(cudaMemsetAsync(c->d_count, 0, sizeof(unsigned long long), c->stream));
    expire_kernel<<<(uint32_t)((depth + 255) / 256), 256, 0, c->stream>>>(c->d_hdr, c->d_exp, c->slot_mask, head, (uint32_t)depth, now_unix_ns, c->d_count);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_count, c->d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    std::lock_guard<std::mutex> lk(c->mu);
    c->stats.kernel_launches++;
    c->cancelled_pending += *c->h_count;
    c->expired_since_launch = true;
    return (int64_t)*c->h_count;
}

// Waits for the last launch and publishes its outcome (records, bytes, overflow) to the ctx. Caller holds c->out_mu (not c->mu).
static int finish_launch(b9_ctx* c) {
    if (!c->res_async) return B9_OK;
    c->res_async = false;
    CU(cudaMemcpyAsync(c->h_ctl, c->d_ctl, sizeof(DrainCtl), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    // one launch: the time around its kernels; a burst of B9_DRAIN_ASYNC launches: from the first one's start to the last one's end
    float ms = 0; cudaEventElapsedTime(&ms, c->burst_open ? c->ev_burst : c->ev_a, c->ev_b);
    c->burst_open = false;
    c->res_kernel_ms = ms;
    if (c->h_ctl->overflow) {
        c->running.store(0, std::memory_order_relaxed);
        return fail(B9_ENOSPC, "b9_drain_launch: results exceed max_result_bytes (%llu); drain fewer tasks or enlarge the staging",
                    (unsigned long long)c->max_result_bytes);
    }
    c->res_bytes = c->h_ctl->bytes; c->res_n = c->h_ctl->total_cnt;
    c->res_popped = c->res_async_n;
    c->res_in_bytes = c->res_async_in_bytes;
    c->have_results = true;
    std::lock_guard<std::mutex> lk(c->mu);
    c->stats.last_drain_kernel_ms = ms;
    c->stats.last_drain_in_bytes = c->res_async_in_bytes;
    c->stats.last_drain_out_bytes = c->res_bytes;
    return B9_OK;
}

int64_t b9_drain_launch(b9_ctx* c, int handler, uint32_t max_tasks, int peek) {
    if (!c) return fail(B9_EINVAL, "b9_drain_launch: ctx is NULL");
    if (handler < 0 || handler >= B9_H_COUNT_) return fail(B9_ENOSYS, "b9_drain_launch: unknown handler %d", handler);
    std::lock_guard<std::mutex> out_lk(c->out_mu);
    CU(cudaSetDevice(c->device));
    // results of an earlier launch that were never fetched are dropped; their tasks are still
    // pending (a pop is committed by a successful fetch, never by a launch). (An open burst of async launches stays open.)
    c->have_results = false; c->res_async = false;
    c->running.store(0, std::memory_order_relaxed);
    c->res_n = 0; c->res_popped = 0; c->res_bytes = 0; c->res_in_bytes = 0; c->res_peek = (peek & B9_DRAIN_PEEK) != 0;
    cudaStream_t s = c->stream;
    // the window: a snapshot of the ring's head under the bookkeeping lock. Only drain-side calls (all under out_mu)
    // move the head or free segments, so the snapshot stays valid while the launch is prepared without c->mu.
    struct Piece { Segment sg; uint64_t a, b; };
    std::vector<Piece> pieces;
    uint64_t head = 0; uint32_t n = 0; uint32_t count_mode = 0;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        head = c->head_task;
        const uint64_t depth = c->tail_task - c->head_task;
        n = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(depth, max_tasks), c->max_drain_tasks);
        count_mode = c->cancelled_pending ? 1u : 0u;
        c->expired_since_launch = false;
        const uint64_t lo = head, hi = head + n;
        for (const Segment& sg : c->segs) {
            const uint64_t a = std::max<uint64_t>(lo, sg.first_task), b = std::min<uint64_t>(hi, sg.first_task + sg.n);
            if (a < b) pieces.push_back(Piece{sg, a, b});
        }
    }
    if (n == 0) { c->have_results = true; return 0; }
    uint64_t in_bytes = 0;                                                  // payload bytes of the window
    uint32_t win_max_len = 0; bool max_known = true;                        // its longest task, when every batch in it told
    for (const Piece& pc : pieces) { if (pc.sg.max_len) win_max_len = std::max(win_max_len, pc.sg.max_len); else max_known = false; }
    if (!max_known) win_max_len = 0;
    for (const Piece& pc : pieces) {
        CU(cudaStreamWaitEvent(s, pc.sg.ready, 0));                         // this batch's H2D + ingest must have landed
        uint64_t o0 = 0, o1 = 0;
        int rc = task_phys_off(c, pc.sg, pc.a, &o0); if (rc) return rc;
        rc = task_phys_off(c, pc.sg, pc.b, &o1); if (rc) return rc;
        in_bytes += o1 - o0;
    }
    DrainArgs a{};
    a.payload = c->d_payload; a.off = c->d_off; a.hdr = c->d_hdr; a.ids = c->d_ids; a.slot_mask = c->slot_mask;
    a.first_task = head; a.n_tasks = n; a.n_tiles = 0;                     // (set per handler by launch_drain3)
    a.out_payload = c->d_out_payload; a.out_cap = c->max_result_bytes; a.out_off = c->d_out_off; a.out_ids = c->d_out_ids;
    a.out_status = c->d_out_status; a.out_has = c->d_out_has; a.out_len = c->d_out_len; a.ctl = c->d_ctl; a.tile_base = (const uint32_t*)c->d_tile_state; a.block_base = (const uint32_t*)c->d_tile_state + ((size_t)c->max_drain_tasks / D2_THREADS + 2); a.handler = handler;
    a.count_mode = count_mode;
    a.slow = c->d_slow; a.crc_shift_tabs = c->d_crc_shift;
    a.static_rounds = 0;
    c->drain_epoch = (c->drain_epoch + 1u) & 0xFFFFFFu;
    if (c->drain_epoch == 0u) { c->drain_epoch = 1u; CU(cudaMemsetAsync(c->d_slow, 0, (size_t)c->max_drain_tasks * sizeof(SlowItem), s)); }   // the 24-bit tag wrapped: forget old tags
    a.epoch = c->drain_epoch;
    CU(cudaMemsetAsync(c->d_ctl, 0, sizeof(DrainCtl), s));
    int grid = 0;
    if (!c->burst_open) CU(cudaEventRecord(c->ev_burst, s));
    CU(cudaEventRecord(c->ev_a, s));
    cudaError_t le;
    switch (handler) {
    case B9_H_IDENTITY: le = launch_drain3<0>(a, in_bytes / n, win_max_len, c->stage_bytes_override, c->sm_count, s, &grid); break;
    case B9_H_CRC32:    le = launch_drain3<1>(a, in_bytes / n, win_max_len, c->stage_bytes_override, c->sm_count, s, &grid); break;
    case B9_H_VADD_F32: le = launch_drain3<2>(a, in_bytes / n, win_max_len, c->stage_bytes_override, c->sm_count, s, &grid); break;
    default:            le = launch_drain3<3>(a, in_bytes / n, win_max_len, c->stage_bytes_override, c->sm_count, s, &grid); break;
    }
    if (le != cudaSuccess) return fail(B9_EIO, "drain kernel launch failed: %s", cudaGetErrorString(le));
    CU(cudaEventRecord(c->ev_b, s));
    // (the control block is read back by finish_launch: a burst of B9_DRAIN_ASYNC launches then has no copy-engine
    // operation between one launch's kernel and the next one's)
    // (measured: 0.1463 against 0.1566 ms per step back to back, profiles/r2_k1_*)
    {
        std::lock_guard<std::mutex> lk(c->mu);
        c->stats.kernel_launches += a.count_mode ? 3 : 1;                  // (+ tile_count_kernel and tile_scan_kernel)
        c->stats.last_drain_tiles = (n + 127) / 128;                       // (reported in units of 128 tasks)
        c->stats.drains++;
    }
    c->res_async = true; c->res_async_n = n; c->res_async_in_bytes = in_bytes;
    if (!c->res_peek) c->running.store(n, std::memory_order_relaxed);
    if (peek & B9_DRAIN_ASYNC) { c->burst_open = true; return (int64_t)n; }   // records = n minus the cancelled slots: known after b9_sync / b9_drain_fetch
    int rc = finish_launch(c);
    return rc ? rc : (int64_t)c->res_n;
}

int64_t b9_drain_fetch(b9_ctx* c, b9_results* out) {
    if (!c || !out) return fail(B9_EINVAL, "b9_drain_fetch: NULL argument");
    std::lock_guard<std::mutex> out_lk(c->out_mu);
    CU(cudaSetDevice(c->device));
    { int rc = finish_launch(c); if (rc) return rc; }
    if (!c->have_results) return fail(B9_EINVAL, "b9_drain_fetch: no drain results are waiting");
    out->n_results = c->res_n; out->n_popped = c->res_popped; out->n_bytes = c->res_bytes; out->need_bytes = c->res_bytes;
    out->task_duration = 0.f;
    if (c->res_n > out->cap_tasks || c->res_bytes > out->cap_bytes)
        return fail(B9_ENOSPC, "b9_drain_fetch: need %u records / %llu bytes, caller gave %u / %llu", c->res_n,
                    (unsigned long long)c->res_bytes, out->cap_tasks, (unsigned long long)out->cap_bytes);
    cudaStream_t s = c->stream;
    const size_t n = c->res_n;
    CU(cudaEventRecord(c->ev_c, s));
    if (n) {
        if (!out->task_ids || !out->status || !out->has_result || !out->offsets || !out->lengths) return fail(B9_EINVAL, "b9_drain_fetch: NULL result array");
        CU(cudaMemcpyAsync(out->lengths, c->d_out_len, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(out->task_ids, c->d_out_ids, n * 16, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(out->status, c->d_out_status, n, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(out->has_result, c->d_out_has, n, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(out->offsets, c->d_out_off, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        if (c->res_bytes) {
            if (!out->payload) return fail(B9_EINVAL, "b9_drain_fetch: NULL payload buffer");
            CU(cudaMemcpyAsync(out->payload, c->d_out_payload, c->res_bytes, cudaMemcpyDeviceToHost, s));
        }
    }
    CU(cudaEventRecord(c->ev_d, s));
    CU(cudaStreamSynchronize(s));
    float ms = 0; cudaEventElapsedTime(&ms, c->ev_c, c->ev_d);
    // TaskQueueCompleteRequest.task_duration (taskqueue.proto:50): the runner reports the seconds it spent on a task; a
    // drain spends kernel + read-back time on res_popped tasks at once, so every task of the batch is charged its share
    if (c->res_popped) out->task_duration = (c->res_kernel_ms + ms) * 1e-3f / (float)c->res_popped;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        c->stats.last_drain_d2h_ms = ms;
        c->stats.bytes_d2h += c->res_bytes + n * (16 + 1 + 1 + 8 + 4);
        if (!c->res_peek) {
            c->head_task += c->res_popped; c->pending_bytes -= c->res_in_bytes; c->stats.tasks_drained += c->res_popped;
            c->cancelled_pending -= std::min<uint64_t>(c->cancelled_pending, (uint64_t)(c->res_popped - c->res_n));
            c->recount_cancelled = c->expired_since_launch;              // an expire ran between this drain's launch and its commit: its
                                                                          // count included tasks of the window that are gone now
            free_segments(c);
        }
    }
    if (!c->res_peek && c->recount_cancelled) { int rc = recount_cancelled_locked_out(c); if (rc) return rc; }
    c->running.store(0, std::memory_order_relaxed);
    c->have_results = false; c->res_async = false; c->burst_open = false;
    return (int64_t)n;
}

int64_t b9_drain(b9_ctx* c, int handler, uint32_t max_tasks, b9_results* out) {
    if (!c || !out) return fail(B9_EINVAL, "b9_drain: NULL argument");
    if (out->cap_tasks < max_tasks) max_tasks = out->cap_tasks;
    int64_t r = b9_drain_launch(c, handler, max_tasks, /*peek=*/0);
    if (r < 0) return r;
    return b9_drain_fetch(c, out);   // on B9_ENOSPC nothing is consumed and the records stay fetchable
}

// =========================================================================== wire records
static size_t host_go_quote(const char* s, uint8_t* o, size_t cap) {
    // encode.go appendString for an (assumed valid UTF-8) host string; HTML-safe
    size_t n = 0;
    auto put = [&](uint8_t c) { if (n < cap) o[n] = c; ++n; };
    static const char HX[] = "0123456789abcdef";
    put('"');
    for (const unsigned char* p = (const unsigned char*)s; *p; ++p) {
        unsigned char c = *p;
        switch (c) {
        case '"': put('\\'); put('"'); break;
        case '\\': put('\\'); put('\\'); break;
        case '\b': put('\\'); put('b'); break;
        case '\f': put('\\'); put('f'); break;
        case '\n': put('\\'); put('n'); break;
        case '\r': put('\\'); put('r'); break;
        case '\t': put('\\'); put('t'); break;
        default:
            if (c < 0x20 || c == '<' || c == '>' || c == '&') { put('\\'); put('u'); put('0'); put('0'); put(HX[c >> 4]); put(HX[c & 15]); }
            else if (c == 0xE2 && p[1] == 0x80 && (p[2] == 0xA8 || p[2] == 0xA9)) { put('\\'); put('u'); put('2'); put('0'); put('2'); put(p[2] == 0xA8 ? '8' : '9'); p += 2; }
            else put(c);
        }
    }
    put('"');
    return n;
}

int64_t b9_wire_encode(b9_ctx* c, const b9_wire_env* env, uint32_t max_tasks) {
    if (!c || !env || !env->workspace_name || !env->stub_id) return fail(B9_EINVAL, "b9_wire_encode: NULL argument");
    std::lock_guard<std::mutex> out_lk(c->out_mu);
    CU(cudaSetDevice(c->device));
    c->have_results = false; c->res_async = false; c->burst_open = false;
    c->running.store(0, std::memory_order_relaxed);
    uint64_t head = 0; uint32_t n = 0;
    cudaStream_t s = c->stream;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        head = c->head_task;
        const uint64_t depth = c->tail_task - c->head_task;
        n = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(depth, max_tasks), c->max_drain_tasks);
        for (const Segment& sg : c->segs) CU(cudaStreamWaitEvent(s, sg.ready, 0));
    }
    c->res_n = 0; c->res_popped = 0; c->res_bytes = 0; c->res_in_bytes = 0; c->res_peek = true;
    if (n == 0) { c->have_results = true; return 0; }
    WireEnv we; memset(&we, 0, sizeof we);
    {
        uint8_t* o = we.mid; size_t k = 0; const size_t cap = sizeof we.mid;
        auto lit = [&](const char* t) { for (; *t; ++t) { if (k < cap) o[k] = (uint8_t)*t; ++k; } };
        lit("\",\"workspace_name\":"); k += host_go_quote(env->workspace_name, o + std::min(k, cap), cap - std::min(k, cap));
        lit(",\"stub_id\":");          k += host_go_quote(env->stub_id, o + std::min(k, cap), cap - std::min(k, cap));
        lit(",\"executor\":");         k += host_go_quote(env->executor ? env->executor : "taskqueue", o + std::min(k, cap), cap - std::min(k, cap));
        lit(",\"args\":");
        if (k > cap) return fail(B9_E2BIG, "b9_wire_encode: workspace/stub names too long");
        we.mid_len = (uint32_t)k;
    }
    we.max_retries = env->max_retries; we.timeout = env->timeout; we.ttl = env->ttl;
    CU(cudaMemcpyAsync(c->d_wire_env, &we, sizeof we, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(c->d_ctl, 0, sizeof(DrainCtl), s));
    WireArgs a{};
    a.payload = c->d_payload; a.off = c->d_off; a.hdr = c->d_hdr; a.ids = c->d_ids; a.ts = c->d_ts; a.exp = c->d_exp;
    a.slot_mask = c->slot_mask; a.first_task = head; a.n_tasks = n;
    a.out_payload = c->d_out_payload; a.out_cap = c->max_result_bytes; a.out_off = c->d_out_off; a.out_len = c->d_out_len;
    a.out_ids = c->d_out_ids; a.out_status = c->d_out_status; a.out_has = c->d_out_has; a.ctl = c->d_ctl;
    CU(cudaEventRecord(c->ev_a, s));
    wire_encode_kernel<<<(n + 127) / 128, 128, 0, s>>>(a, c->d_wire_env);
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev_b, s));
    CU(cudaMemcpyAsync(c->h_ctl, c->d_ctl, sizeof(DrainCtl), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));     // (the stack copy of `we` is consumed by now as well)
    if (c->h_ctl->overflow) return fail(B9_ENOSPC, "b9_wire_encode: records exceed max_result_bytes (%llu)", (unsigned long long)c->max_result_bytes);
    c->res_bytes = c->h_ctl->bytes; c->res_n = n; c->res_popped = 0; c->have_results = true;
    float ms = 0; cudaEventElapsedTime(&ms, c->ev_a, c->ev_b); c->res_kernel_ms = ms;
    std::lock_guard<std::mutex> lk(c->mu);
    c->stats.kernel_launches++;
    c->stats.last_drain_out_bytes = c->res_bytes;
    c->stats.last_drain_kernel_ms = ms;
    return (int64_t)n;
}

// =========================================================================== multi-GPU: NCCL rebalance
}  // extern "C" (reopened below)

namespace {

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ struct B9NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
struct B9NcclId { char internal[128]; };
constexpr int NCCL_U8 = 1, NCCL_U64 = 5;        // ncclUint8, ncclUint64 (nccl.h ncclDataType_t)

NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return B9_OK;
    const char* path = getenv("B9_NCCL_LIB");
    void* h = dlopen(path ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h && !path) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(B9_EIO, "cannot load NCCL (%s); set B9_NCCL_LIB", dlerror());
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return fail(B9_EIO, "NCCL symbol %s missing", name)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = h;
    return B9_OK;
}
void nccl_comm_destroy(void* comm) { if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm); }
#define NC(call) do { int r_ = (call); if (r_ != 0) return fail(B9_EIO, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); } while (0)

// ---- rebalance, device side ---------------------------------------------------------------------------------------
// byte prefix of the pending tasks (prefix[0] = 0, prefix[i + 1] = bytes of tasks 0..i), from the slot ring's length words:
// three small kernels (block sums, scan of the block sums, write-out) — only the W cut points ever cross PCIe
constexpr uint32_t SCAN_BLOCK = 1024;        // tasks per block (256 threads x 4)
__global__ void __launch_bounds__(256) scan_block_sums_kernel(const uint64_t* __restrict__ hdr, uint32_t slot_mask, uint64_t first_task, uint32_t n,
                                                             unsigned long long* __restrict__ block_sum) {
    __shared__ unsigned long long s_w[8];
    const uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * 4u;
    unsigned long long v = 0;
    #pragma unroll
    for (uint32_t k = 0; k < 4; ++k) if (base + k < n) v += hdr_len(hdr[(uint32_t)((first_task + base + k) & slot_mask)]);
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31u) == 0u) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long t = 0; for (int w = 0; w < 8; ++w) t += s_w[w]; block_sum[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) scan_of_sums_kernel(unsigned long long* __restrict__ block_sum, uint32_t blocks) {   // in place, exclusive
    __shared__ unsigned long long s_w[32];
    const uint32_t per = (blocks + 1023u) / 1024u, lo = min(blocks, threadIdx.x * per), hi = min(blocks, lo + per);
    unsigned long long sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += block_sum[i];
    unsigned long long inc = sum;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc += t; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long x = s_w[lane], y = x;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, y, d); if ((int)lane >= d) y += t; }
        s_w[lane] = y - x;
    }
    __syncthreads();
    unsigned long long run = s_w[warp] + inc - sum;
    for (uint32_t i = lo; i < hi; ++i) { const unsigned long long t = block_sum[i]; block_sum[i] = run; run += t; }
}
__global__ void __launch_bounds__(256) scan_write_kernel(const uint64_t* __restrict__ hdr, uint32_t slot_mask, uint64_t first_task, uint32_t n,
                                                        const unsigned long long* __restrict__ block_base, uint64_t* __restrict__ prefix) {
    __shared__ unsigned long long s_w[8];
    const uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * 4u;
    unsigned long long l[4]; unsigned long long v = 0;
    #pragma unroll
    for (uint32_t k = 0; k < 4; ++k) { l[k] = (base + k < n) ? hdr_len(hdr[(uint32_t)((first_task + base + k) & slot_mask)]) : 0ull; v += l[k]; }
    unsigned long long inc = v;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc += t; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    unsigned long long before = block_base[blockIdx.x];
    for (uint32_t w = 0; w < warp; ++w) before += s_w[w];
    unsigned long long run = before + inc - v;
    if (blockIdx.x == 0 && threadIdx.x == 0) prefix[0] = 0;
    #pragma unroll
    for (uint32_t k = 0; k < 4; ++k) { run += l[k]; if (base + k < n) prefix[base + k + 1] = run; }
}

// What one rank learns in the planning phase, written by the cut kernel and read back in ONE copy.
struct RebalanceCuts {
    uint64_t lo[16], hi[16];             // my FIFO range for every destination rank
    uint64_t bytes[16];                  // payload bytes of that range
    uint64_t phys[16];                   // physical ring offset of its first task
    uint64_t contiguous[16];             // 1: the range is one contiguous byte range of the ring (send it from there)
    uint64_t keep_end_phys;              // physical offset of the first task that leaves to a HIGHER rank (= end of what I keep)
    uint64_t ok;                         // the plan succeeded
};
constexpr int REBALANCE_MAX_WORLD = 16;

// ONE thread: the same plan function the host exports (rebalance_plan.h) over the device-resident prefix, plus the physical
// facts the host needs to post the sends. tab = all-gathered (count, bytes, cancelled, 0) per rank; row_out = my row of the
// send matrix (tasks, bytes per destination), all-gathered next.
__global__ void rebalance_cut_kernel(const uint64_t* __restrict__ tab, uint32_t world, uint32_t rank, const uint64_t* __restrict__ prefix, uint64_t n,
                                     const uint64_t* __restrict__ off, const uint64_t* __restrict__ hdr, uint32_t slot_mask, uint64_t first_task,
                                     RebalanceCuts* __restrict__ cuts, uint64_t* __restrict__ row_out) {
    if (blockIdx.x || threadIdx.x) return;
    uint64_t counts[REBALANCE_MAX_WORLD], bytes[REBALANCE_MAX_WORLD];
    for (uint32_t r = 0; r < world; ++r) { counts[r] = tab[4 * r]; bytes[r] = tab[4 * r + 1]; }
    RebalanceCuts c;
    c.ok = b9_plan_ranges(world, rank, counts, bytes, prefix, n, c.lo, c.hi) == 0 ? 1ull : 0ull;
    for (uint32_t d = 0; d < world; ++d) {
        if (!c.ok) { c.lo[d] = c.hi[d] = 0; }
        const uint64_t lo = c.lo[d], hi = c.hi[d];
        c.bytes[d] = prefix[hi] - prefix[lo];
        c.phys[d] = 0; c.contiguous[d] = 1;
        if (hi > lo) {
            const uint32_t s0 = (uint32_t)((first_task + lo) & slot_mask), s1 = (uint32_t)((first_task + hi - 1) & slot_mask);
            c.phys[d] = off[s0];
            c.contiguous[d] = (off[s1] + hdr_len(hdr[s1]) - off[s0] == c.bytes[d] && off[s1] >= off[s0]) ? 1ull : 0ull;
        }
        row_out[2 * d] = hi - lo; row_out[2 * d + 1] = c.bytes[d];
    }
    const uint64_t keep_hi = c.hi[rank];
    c.keep_end_phys = keep_hi < n ? off[(uint32_t)((first_task + keep_hi) & slot_mask)] : 0;
    *cuts = c;
}

// slot words of the tasks that leave (ids, lengths as relative offsets, flags, retries, timestamps, expiries) into one
// meta message; with `out_payload` also their bytes (only for ranges that are not contiguous in the ring)
__global__ void gather_tasks_kernel(const uint8_t* __restrict__ payload, const uint64_t* __restrict__ off, const uint64_t* __restrict__ hdr,
                                    const uint4* __restrict__ ids, const int64_t* __restrict__ ts, const int64_t* __restrict__ exp,
                                    uint32_t slot_mask, uint64_t first_task, uint32_t n, const uint64_t* __restrict__ prefix /* device prefix, at the range's first task */,
                                    uint8_t* __restrict__ out_payload, uint64_t* __restrict__ o_rel, int64_t* __restrict__ o_ts, int64_t* __restrict__ o_exp,
                                    uint4* __restrict__ o_ids, uint8_t* __restrict__ o_flags, uint8_t* __restrict__ o_retries) {
    if (out_payload) {                                   // one warp per task
        const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
        if (w >= n) return;
        const uint32_t slot = (uint32_t)((first_task + w) & slot_mask);
        const uint64_t h = hdr[slot];
        const uint32_t len = hdr_len(h);
        const uint8_t* src = payload + off[slot];
        uint8_t* dst = out_payload + (prefix[w] - prefix[0]);
        for (uint32_t i = lane; i < len; i += 32) dst[i] = src[i];
        if (lane == 0) { o_rel[w] = prefix[w]; if (w + 1 == n) o_rel[n] = prefix[n]; o_ts[w] = ts[slot]; o_exp[w] = exp[slot]; o_ids[w] = ids[slot]; o_flags[w] = (uint8_t)hdr_flags(h); o_retries[w] = (uint8_t)(h >> 40); }
    } else {                                             // slot words only: one thread per task
        const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
        if (w >= n) return;
        const uint32_t slot = (uint32_t)((first_task + w) & slot_mask);
        const uint64_t h = hdr[slot];
        o_rel[w] = prefix[w]; if (w + 1 == n) o_rel[n] = prefix[n];
        o_ts[w] = ts[slot]; o_exp[w] = exp[slot]; o_ids[w] = ids[slot]; o_flags[w] = (uint8_t)hdr_flags(h); o_retries[w] = (uint8_t)(h >> 40);
    }
}

// layout of one peer's meta message for n tasks (every section 16-byte aligned)
struct MetaLayout { size_t rel, ts, exp, ids, flags, retries, total; };
MetaLayout meta_layout(uint64_t n) {
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    MetaLayout m; size_t o = 0;
    m.rel = o; o = al(o + (n + 1) * 8);
    m.ts = o; o = al(o + n * 8);
    m.exp = o; o = al(o + n * 8);
    m.ids = o; o = al(o + n * 16);
    m.flags = o; o = al(o + n);
    m.retries = o; o = al(o + n);
    m.total = o; return m;
}

// drop k tasks from the tail of the ring (they were handed to another rank); cut_phys = physical offset of the first dropped task
void drop_back(b9_ctx* c, uint64_t k, uint64_t cut_phys) {
    while (k && !c->segs.empty()) {
        Segment& sg = c->segs.back();
        if (sg.n <= k) {
            k -= sg.n; c->tail_task -= sg.n; c->pending_bytes -= sg.bytes; c->write_pos = sg.phys_start;
            c->event_pool.push_back(sg.ready); c->segs.pop_back();
        } else {
            const uint32_t keep = sg.n - (uint32_t)k;
            const uint64_t nb = cut_phys - sg.phys_start;               // (the cut lies inside this segment)
            c->pending_bytes -= sg.bytes - nb; c->tail_task -= k;
            sg.n = keep; sg.bytes = nb;
            c->write_pos = sg.phys_start + b9_seg_span(nb);
            k = 0;
        }
    }
    if (c->segs.empty()) c->write_pos = 0;
}

}  // namespace

extern "C" {

int b9_rebalance_plan(uint32_t world, uint32_t rank, const uint64_t* counts, const uint64_t* bytes, const uint64_t* prefix, uint64_t n,
                      uint64_t* send_lo, uint64_t* send_hi) {
    if (!counts || !bytes || !prefix || !send_lo || !send_hi) return fail(B9_EINVAL, "b9_rebalance_plan: NULL argument");
    if (b9_plan_ranges(world, rank, counts, bytes, prefix, n, send_lo, send_hi)) return fail(B9_EINVAL, "b9_rebalance_plan: table inconsistent with prefix");
    return B9_OK;
}

int b9_comm_unique_id(uint8_t* out128) {
    if (!out128) return fail(B9_EINVAL, "b9_comm_unique_id: NULL");
    int rc = load_nccl(); if (rc) return rc;
    B9NcclId id; memset(&id, 0, sizeof id);
    NC(g_nccl.GetUniqueId(&id));
    memcpy(out128, id.internal, 128);
    return B9_OK;
}

int b9_comm_init(b9_ctx* c, const uint8_t* id128, int rank, int world) {
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return fail(B9_EINVAL, "b9_comm_init: bad argument");
    if (world > REBALANCE_MAX_WORLD) return fail(B9_EINVAL, "b9_comm_init: world %d > %d", world, REBALANCE_MAX_WORLD);
    int rc = load_nccl(); if (rc) return rc;
    std::lock_guard<std::mutex> in_lk(c->in_mu);
    std::lock_guard<std::mutex> out_lk(c->out_mu);
    CU(cudaSetDevice(c->device));
    B9NcclId id; memcpy(id.internal, id128, 128);
    if (c->nccl_comm) { nccl_comm_destroy(c->nccl_comm); c->nccl_comm = nullptr; }   // re-init: the old communicator is not leaked
    void* comm = nullptr;
    NC(g_nccl.CommInitRank(&comm, world, id, rank));
    c->nccl_comm = comm; c->comm_rank = rank; c->comm_world = world;
    if (world > 1) {
        // everything the planning phase of b9_rebalance touches is allocated here, once: the byte prefix of a full slot ring,
        // the block sums of its scan, the all-gather tables and the page-locked block the cut points come back in
        if (!c->d_prefix) {
            CU(cudaMalloc(&c->d_prefix, ((size_t)c->ring_tasks + 2) * sizeof(uint64_t)));
            CU(cudaMalloc(&c->d_scan_sums, ((size_t)c->ring_tasks / SCAN_BLOCK + 2) * sizeof(unsigned long long)));
        }
        if (!c->d_rtab) {
            c->rtab_words = 4 * (size_t)REBALANCE_MAX_WORLD * 2 + 2 * (size_t)REBALANCE_MAX_WORLD * (REBALANCE_MAX_WORLD + 1) + (size_t)REBALANCE_MAX_WORLD + sizeof(RebalanceCuts) / 8 + 16;
            CU(cudaMalloc(&c->d_rtab, c->rtab_words * sizeof(uint64_t)));
            CU(cudaHostAlloc(&c->h_rtab, c->rtab_words * sizeof(uint64_t), cudaHostAllocDefault));
        }
        // ... and a first size for the two arenas the slot words (and non-contiguous payload ranges) are staged in
        const uint64_t first = std::max<uint64_t>((uint64_t)c->ring_tasks * 48 / 2, 16ull << 20);
        if (c->xchg_send_cap < first) { cudaFree(c->d_xchg_send); c->d_xchg_send = nullptr; c->xchg_send_cap = 0;
                                        if (cudaMalloc(&c->d_xchg_send, first) == cudaSuccess) c->xchg_send_cap = first; else cudaGetLastError(); }
        if (c->xchg_recv_cap < first) { cudaFree(c->d_xchg_recv); c->d_xchg_recv = nullptr; c->xchg_recv_cap = 0;
                                        if (cudaMalloc(&c->d_xchg_recv, first) == cudaSuccess) c->xchg_recv_cap = first; else cudaGetLastError(); }
    }
    // NCCL sets its point-to-point channels up lazily, on the first send/recv of every pair and of every
    // channel a message is spread over (seconds on an 8-GPU box): do that here, once, with an all-to-all
    // large enough to touch all of them (B9_COMM_WARMUP_BYTES per peer, default 4 MiB), plus the
    // all-gather b9_rebalance opens with, so that the first real exchange finds everything connected
    if (world > 1) {
        size_t per = 4u << 20;
        if (const char* e = getenv("B9_COMM_WARMUP_BYTES")) per = (size_t)std::max<long long>(16, atoll(e));
        uint8_t* d = nullptr;
        CU(cudaMalloc(&d, (size_t)world * per * 2));
        CU(cudaMemsetAsync(d, 0, (size_t)world * per * 2, c->stream));
        NC(g_nccl.AllGather(d, d + per, 16, NCCL_U64, comm, c->stream));      // (world * 128 bytes <= per)
        NC(g_nccl.GroupStart());
        for (int p = 0; p < world; ++p) {
            if (p == rank) continue;
            NC(g_nccl.Send(d + (size_t)p * per * 2, per, NCCL_U8, p, comm, c->stream));
            NC(g_nccl.Recv(d + (size_t)p * per * 2 + per, per, NCCL_U8, p, comm, c->stream));
        }
        NC(g_nccl.GroupEnd());
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(d);
    }
    return B9_OK;
}

// Collective: every rank of the communicator must call it. Moves pending tasks between the ranks' rings so that every
// rank holds ~1/W of the pending payload BYTES. Nothing per-task crosses PCIe:
//   1. device: byte prefix of my pending tasks from the slot ring's length words (three small scan kernels);
//   2. all-gather of (count, bytes, cancelled) per rank — on the device;
//   3. device: ONE thread runs the byte-quantile plan (rebalance_plan.h, the function the CPU tests run) against the
//      prefix and writes my cut points, the physical ring offsets at the cuts and my row of the send matrix;
//   4. all-gather of the matrix rows; ONE copy brings table + cuts + matrix to the host (first synchronise);
//   5. every rank decides whether what arrives fits its ring — slot count AND contiguous byte placement, with what leaves
//      still in place — and the ranks vote (all-gather of one flag, second synchronise): if any rank cannot take its
//      share, ALL return B9_ENOSPC with their rings untouched;
//   6. the exchange, one grouped ncclSend/ncclRecv: payload ranges that are contiguous in the ring (a batch pushed in one
//      piece is) leave straight from the ring, arriving payload lands straight in the ring space placed in 5; only the
//      slot words (41 B per task) are staged through the two arenas;
//   7. bookkeeping: what left is dropped from both ends, what arrived is appended source by source (ingest_kernel).
int b9_rebalance(b9_ctx* c, b9_rebalance_info* info) {
    if (!c) return fail(B9_EINVAL, "b9_rebalance: ctx is NULL");
    if (!c->nccl_comm) return fail(B9_EINVAL, "b9_rebalance: b9_comm_init was not called");
    std::lock_guard<std::mutex> in_lk(c->in_mu);       // the exchange rewrites both ends of the ring: no push, no drain meanwhile
    std::lock_guard<std::mutex> out_lk(c->out_mu);
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    const int W = c->comm_world, R = c->comm_rank;
    cudaStream_t s = c->stream;
    static const bool trace = getenv("B9_REBALANCE_TRACE") != nullptr;       // phase wall times on stderr
    auto t_last = std::chrono::steady_clock::now();
    auto phase = [&](const char* what) {
        if (!trace) return;
        cudaStreamSynchronize(s);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[b9_rebalance rank %d] %-28s %8.3f ms\n", R, what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    if (W == 1) {
        if (info) { memset(info, 0, sizeof *info); info->world = 1; info->tasks_before = info->tasks_after = c->tail_task - c->head_task; info->bytes_before = info->bytes_after = c->pending_bytes; }
        return B9_OK;
    }
    for (const Segment& sg : c->segs) CU(cudaStreamWaitEvent(s, sg.ready, 0));
    c->have_results = false; c->res_async = false; c->burst_open = false;
    c->running.store(0, std::memory_order_relaxed);
    free_segments(c);
    const uint64_t n = c->tail_task - c->head_task;
    const uint64_t my_bytes = c->pending_bytes;
    // ---- layout of the device / pinned table (64-bit words)
    uint64_t* const d_me = c->d_rtab;                                   // [4] count, bytes, cancelled, 0
    uint64_t* const d_tab = d_me + 4;                                   // [4 W] all-gathered
    uint64_t* const d_row = d_tab + 4 * REBALANCE_MAX_WORLD;            // [2 W] my matrix row
    uint64_t* const d_mat = d_row + 2 * REBALANCE_MAX_WORLD;            // [2 W W]
    uint64_t* const d_vote = d_mat + 2 * REBALANCE_MAX_WORLD * REBALANCE_MAX_WORLD;   // [1 + W]
    RebalanceCuts* const d_cuts = (RebalanceCuts*)(d_vote + 1 + REBALANCE_MAX_WORLD);
    uint64_t* const h = c->h_rtab;
    const size_t off_tab = 4, off_row = off_tab + 4 * REBALANCE_MAX_WORLD, off_mat = off_row + 2 * REBALANCE_MAX_WORLD,
                 off_vote = off_mat + 2 * REBALANCE_MAX_WORLD * REBALANCE_MAX_WORLD, off_cuts = off_vote + 1 + REBALANCE_MAX_WORLD;
    // ---- 1. byte prefix on the device
    const uint32_t blocks = (uint32_t)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
    if (n) {
        scan_block_sums_kernel<<<blocks, 256, 0, s>>>(c->d_hdr, c->slot_mask, c->head_task, (uint32_t)n, c->d_scan_sums);
        scan_of_sums_kernel<<<1, 1024, 0, s>>>(c->d_scan_sums, blocks);
        scan_write_kernel<<<blocks, 256, 0, s>>>(c->d_hdr, c->slot_mask, c->head_task, (uint32_t)n, c->d_scan_sums, c->d_prefix);
        CU(cudaGetLastError());
        c->stats.kernel_launches += 3;
    } else CU(cudaMemsetAsync(c->d_prefix, 0, sizeof(uint64_t), s));
    // ---- 2. all-gather (count, bytes, cancelled)
    h[0] = n; h[1] = my_bytes; h[2] = c->cancelled_pending; h[3] = 0;
    CU(cudaMemcpyAsync(d_me, h, 4 * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    NC(g_nccl.AllGather(d_me, d_tab, 4, NCCL_U64, c->nccl_comm, s));
    // ---- 3. cuts, 4. matrix
    rebalance_cut_kernel<<<1, 32, 0, s>>>(d_tab, (uint32_t)W, (uint32_t)R, c->d_prefix, n, c->d_off, c->d_hdr, c->slot_mask, c->head_task, d_cuts, d_row);
    CU(cudaGetLastError());
    c->stats.kernel_launches++;
    NC(g_nccl.AllGather(d_row, d_mat, (size_t)2 * W, NCCL_U64, c->nccl_comm, s));
    CU(cudaMemcpyAsync(h + off_tab, d_tab, (c->rtab_words - off_tab) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    phase("prefix + plan + matrix");
    const RebalanceCuts& cuts = *(const RebalanceCuts*)(h + off_cuts);
    const uint64_t* tab = h + off_tab;
    const uint64_t* mat = h + off_mat;
    auto M_tasks = [&](int src, int dst) { return mat[((size_t)src * W + dst) * 2]; };
    auto M_bytes = [&](int src, int dst) { return mat[((size_t)src * W + dst) * 2 + 1]; };
    if (!cuts.ok) return fail(B9_EIO, "b9_rebalance: ring bookkeeping inconsistent with the slot ring (%llu tasks, %llu bytes)", (unsigned long long)n, (unsigned long long)my_bytes);
    bool any_cancelled = false;
    for (int r = 0; r < W; ++r) any_cancelled |= tab[4 * r + 2] != 0;
    // ---- 5. does what arrives fit? (placement with everything that leaves still in place: a receive never lands on bytes a send reads)
    uint64_t sent_tasks = 0, sent_bytes = 0, recv_tasks = 0, recv_bytes = 0;
    for (int d = 0; d < W; ++d) if (d != R) { sent_tasks += cuts.hi[d] - cuts.lo[d]; sent_bytes += cuts.bytes[d]; recv_tasks += M_tasks(d, R); recv_bytes += M_bytes(d, R); }
    std::vector<uint64_t> r_start(W, 0);
    bool fits = n - sent_tasks + recv_tasks <= c->ring_tasks;
    {
        // a dry run of the appends on a copy of the placement state
        uint64_t wp = c->write_pos, live = 0; bool have = !c->segs.empty();
        const uint64_t oldest = have ? c->segs.front().phys_start : 0;
        for (const Segment& sg : c->segs) live += sg.bytes;
        for (int src = 0; src < W && fits; ++src) {
            if (src == R || !M_tasks(src, R)) continue;
            const uint64_t pb = M_bytes(src, R);
            uint64_t st = 0;
            if (!b9_ring_place(c->ring_bytes, have, have ? oldest : 0, wp, live, pb, &st)) { fits = false; break; }
            r_start[src] = st;
            have = true;                                               // (an empty ring's first arrival is placed at 0 and becomes the oldest segment: oldest = 0)
            wp = st + b9_seg_span(pb); live += pb;
        }
    }
    h[0] = fits ? 1 : 0;
    CU(cudaMemcpyAsync(d_vote, h, sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    NC(g_nccl.AllGather(d_vote, d_vote + 1, 1, NCCL_U64, c->nccl_comm, s));
    CU(cudaMemcpyAsync(h + off_vote, d_vote, (size_t)(1 + W) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (int r = 0; r < W; ++r)
        if (!h[off_vote + 1 + r])
            return fail(B9_ENOSPC, "b9_rebalance: rank %d cannot take its share (%llu tasks / %llu bytes arrive here); no rank changed its ring",
                        r, (unsigned long long)recv_tasks, (unsigned long long)recv_bytes);
    phase("feasibility vote");
    // ---- 6. stage the slot words of what leaves (and the payload of ranges that are not one piece of the ring)
    auto al = [](uint64_t x) { return (x + 255ull) & ~255ull; };
    uint64_t need_s = 0, need_r = 0;
    for (int d = 0; d < W; ++d) {
        if (d == R) continue;
        const uint64_t k = cuts.hi[d] - cuts.lo[d];
        if (k) need_s += al(meta_layout(k).total) + (cuts.contiguous[d] ? 0 : al(cuts.bytes[d] + 16));
        const uint64_t rk = M_tasks(d, R);
        if (rk) need_r += al(meta_layout(rk).total);
    }
    auto ensure = [&](uint8_t*& buf, uint64_t& cap, uint64_t need) -> cudaError_t {
        if (need <= cap) return cudaSuccess;
        if (buf) cudaFree(buf);
        buf = nullptr; cap = 0;
        const uint64_t want = need + need / 4;
        cudaError_t e = cudaMalloc(&buf, want);
        if (e == cudaSuccess) cap = want;
        return e;
    };
    CU(ensure(c->d_xchg_send, c->xchg_send_cap, need_s));
    CU(ensure(c->d_xchg_recv, c->xchg_recv_cap, need_r));
    std::vector<uint8_t*> s_meta(W, nullptr), r_meta(W, nullptr);
    std::vector<const uint8_t*> s_pay(W, nullptr);
    uint64_t so = 0, ro = 0;
    for (int d = 0; d < W; ++d) {
        if (d == R) continue;
        const uint64_t k = cuts.hi[d] - cuts.lo[d];
        if (k) {
            const MetaLayout ml = meta_layout(k);
            s_meta[d] = c->d_xchg_send + so; so += al(ml.total);
            uint8_t* staged = nullptr;
            if (!cuts.contiguous[d]) { staged = c->d_xchg_send + so; so += al(cuts.bytes[d] + 16); }
            s_pay[d] = staged ? staged : c->d_payload + cuts.phys[d];
            const uint32_t threads = staged ? (uint32_t)k * 32u : (uint32_t)k;
            gather_tasks_kernel<<<(threads + 255) / 256, 256, 0, s>>>(c->d_payload, c->d_off, c->d_hdr, c->d_ids, c->d_ts, c->d_exp, c->slot_mask,
                                                                       c->head_task + cuts.lo[d], (uint32_t)k, c->d_prefix + cuts.lo[d], staged,
                                                                       (uint64_t*)(s_meta[d] + ml.rel), (int64_t*)(s_meta[d] + ml.ts), (int64_t*)(s_meta[d] + ml.exp),
                                                                       (uint4*)(s_meta[d] + ml.ids), s_meta[d] + ml.flags, s_meta[d] + ml.retries);
            CU(cudaGetLastError());
            c->stats.kernel_launches++;
        }
        const uint64_t rk = M_tasks(d, R);
        if (rk) { r_meta[d] = c->d_xchg_recv + ro; ro += al(meta_layout(rk).total); }
    }
    // ---- the exchange: one grouped send/recv; arriving payload goes straight to its place in the ring
    NC(g_nccl.GroupStart());
    for (int d = 0; d < W; ++d) {
        if (d == R) continue;
        const uint64_t k = cuts.hi[d] - cuts.lo[d];
        if (k) {
            NC(g_nccl.Send(s_meta[d], meta_layout(k).total, NCCL_U8, d, c->nccl_comm, s));
            if (cuts.bytes[d]) NC(g_nccl.Send(s_pay[d], cuts.bytes[d], NCCL_U8, d, c->nccl_comm, s));
        }
        const uint64_t rk = M_tasks(d, R);
        if (rk) {
            NC(g_nccl.Recv(r_meta[d], meta_layout(rk).total, NCCL_U8, d, c->nccl_comm, s));
            if (M_bytes(d, R)) NC(g_nccl.Recv(c->d_payload + r_start[d], M_bytes(d, R), NCCL_U8, d, c->nccl_comm, s));
        }
    }
    NC(g_nccl.GroupEnd());
    phase("pack slot words + exchange");
    // ---- 7. what left: a prefix of my FIFO went to lower ranks, a suffix to higher ranks
    {
        const uint64_t front = cuts.lo[R], back = n - cuts.hi[R];
        uint64_t front_bytes = 0;
        for (int d = 0; d < R; ++d) front_bytes += cuts.bytes[d];
        if (back) drop_back(c, back, cuts.keep_end_phys);
        if (front) { c->pending_bytes -= front_bytes; c->head_task += front; free_segments(c); }
    }
    // what arrived is appended, source by source, as new segments at the places chosen in 5
    for (int src = 0; src < W; ++src) {
        const uint64_t rk = (src == R) ? 0 : M_tasks(src, R);
        if (!rk) continue;
        const MetaLayout ml = meta_layout(rk);
        const uint64_t pb = M_bytes(src, R);
        const uint64_t start = r_start[src];
        // slot words: ids may wrap in the slot ring
        const uint32_t slot0 = (uint32_t)(c->tail_task & c->slot_mask);
        const uint32_t first = (uint32_t)std::min<uint64_t>(rk, c->ring_tasks - slot0);
        CU(cudaMemcpyAsync(c->d_ids + slot0, r_meta[src] + ml.ids, (size_t)first * 16, cudaMemcpyDeviceToDevice, s));
        if (first < rk) CU(cudaMemcpyAsync(c->d_ids, r_meta[src] + ml.ids + (size_t)first * 16, (size_t)(rk - first) * 16, cudaMemcpyDeviceToDevice, s));
        ingest_kernel<<<(uint32_t)((rk + 255) / 256), 256, 0, s>>>((const uint64_t*)(r_meta[src] + ml.rel), (uint32_t)rk, start, c->tail_task, c->slot_mask,
                                                                  (const int64_t*)(r_meta[src] + ml.ts), (const int64_t*)(r_meta[src] + ml.exp),
                                                                  r_meta[src] + ml.retries, r_meta[src] + ml.flags, c->d_off, c->d_hdr, c->d_ts, c->d_exp);
        CU(cudaGetLastError());
        c->stats.kernel_launches++;
        cudaEvent_t ready;
        if (!c->event_pool.empty()) { ready = c->event_pool.back(); c->event_pool.pop_back(); }
        else CU(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        CU(cudaEventRecord(ready, s));
        c->segs.push_back(Segment{c->tail_task, (uint32_t)rk, start, pb, ready, 0u});
        c->write_pos = start + b9_seg_span(pb);
        c->tail_task += rk; c->pending_bytes += pb;
    }
    // ---- cancelled tasks may have moved either way: recount over the new window (only if some rank held any)
    if (any_cancelled) {
        const uint64_t depth = c->tail_task - c->head_task;
        CU(cudaMemsetAsync(c->d_count, 0, sizeof(unsigned long long), s));
        if (depth) { count_cancelled_kernel<<<(uint32_t)((depth + 255) / 256), 256, 0, s>>>(c->d_hdr, c->slot_mask, c->head_task, (uint32_t)depth, c->d_count); CU(cudaGetLastError()); c->stats.kernel_launches++; }
        CU(cudaMemcpyAsync(c->h_count, c->d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        c->cancelled_pending = *c->h_count;
    } else {
        CU(cudaStreamSynchronize(s));         // the caller's next call may be on any stream; the exchange is complete when we return
        c->cancelled_pending = 0;
    }
    phase("drop + append + recount");
    if (info) {
        info->world = (uint32_t)W; info->rank = (uint32_t)R;
        info->tasks_before = n; info->bytes_before = my_bytes;
        info->tasks_sent = sent_tasks; info->bytes_sent = sent_bytes; info->tasks_received = recv_tasks; info->bytes_received = recv_bytes;
        info->tasks_after = c->tail_task - c->head_task; info->bytes_after = c->pending_bytes;
    }
    return B9_OK;
}

// =========================================================================== result sink
uint64_t b9_sink_object_bytes(uint32_t n_records, uint64_t blob_bytes) { return b9sink::layout(n_records, blob_bytes).total; }

int64_t b9_sink_pack(const b9_results* r, uint8_t* obj, uint64_t cap) {
    if (!r || !obj) return fail(B9_EINVAL, "b9_sink_pack: NULL argument");
    const uint32_t n = r->n_results;
    const b9sink::Layout L = b9sink::layout(n, r->n_bytes);
    if (cap < L.total) return fail(B9_ENOSPC, "b9_sink_pack: object needs %llu bytes, caller gave %llu", (unsigned long long)L.total, (unsigned long long)cap);
    memset(obj, 0, L.blob);                                               // header + index, padding included: the object is deterministic
    b9sink::write_header(obj, n, r->n_bytes, r->task_duration, r->n_popped);
    if (n) {
        memcpy(obj + L.ids, r->task_ids, (size_t)n * 16); memcpy(obj + L.offsets, r->offsets, (size_t)n * 8);
        memcpy(obj + L.lengths, r->lengths, (size_t)n * 4); memcpy(obj + L.status, r->status, n); memcpy(obj + L.has, r->has_result, n);
    }
    if (r->n_bytes) memcpy(obj + L.blob, r->payload, r->n_bytes);
    return (int64_t)L.total;
}

// b9_drain_fetch with the records landing as ONE sink object in the caller's buffer: the device-to-host copies write the
// object's sections directly (pass page-locked memory for full PCIe bandwidth), the header is written last.
int64_t b9_drain_fetch_object(b9_ctx* c, uint8_t* obj, uint64_t cap, uint64_t* object_bytes) {
    if (!c || !obj) return fail(B9_EINVAL, "b9_drain_fetch_object: NULL argument");
    uint32_t n = 0; uint64_t bytes = 0;
    {
        std::lock_guard<std::mutex> out_lk(c->out_mu);
        CU(cudaSetDevice(c->device));
        { int rc = finish_launch(c); if (rc) return rc; }
        if (!c->have_results) return fail(B9_EINVAL, "b9_drain_fetch_object: no drain results are waiting");
        n = c->res_n; bytes = c->res_bytes;
    }
    const b9sink::Layout L = b9sink::layout(n, bytes);
    if (object_bytes) *object_bytes = L.total;
    if (cap < L.total) return fail(B9_ENOSPC, "b9_drain_fetch_object: object needs %llu bytes, caller gave %llu", (unsigned long long)L.total, (unsigned long long)cap);
    memset(obj, 0, sizeof(b9sink::Header));
    b9_results r; memset(&r, 0, sizeof r);
    r.task_ids = obj + L.ids; r.offsets = (uint64_t*)(obj + L.offsets); r.lengths = (uint32_t*)(obj + L.lengths);
    r.status = obj + L.status; r.has_result = obj + L.has; r.payload = obj + L.blob;
    r.cap_tasks = n ? n : 1; r.cap_bytes = bytes ? bytes : 1;
    const int64_t got = b9_drain_fetch(c, &r);
    if (got < 0) return got;
    // the gaps between the sections (alignment) are part of the object: make them deterministic
    auto zero_gap = [&](uint64_t from, uint64_t to) { if (to > from) memset(obj + from, 0, to - from); };
    zero_gap(L.ids + (uint64_t)n * 16, L.offsets); zero_gap(L.offsets + (uint64_t)n * 8, L.lengths); zero_gap(L.lengths + (uint64_t)n * 4, L.status);
    zero_gap(L.status + n, L.has); zero_gap(L.has + n, L.blob);
    b9sink::write_header(obj, n, bytes, r.task_duration, r.n_popped);
    return got;
}

int b9_sink_get(const uint8_t* obj, uint64_t size, uint32_t index, b9_sink_record* rec) {
    const b9sink::Header* h = b9sink::check(obj, size);
    if (!h || !rec) return fail(B9_EINVAL, "b9_sink_get: not a sink object");
    if (index >= h->n_records) return fail(B9_EINVAL, "b9_sink_get: record %u of %u", index, h->n_records);
    const uint64_t off = ((const uint64_t*)(obj + h->off_offsets))[index];
    const uint32_t len = ((const uint32_t*)(obj + h->off_lengths))[index];
    if (off + len > h->blob_bytes) return fail(B9_EINVAL, "b9_sink_get: record %u points outside the blob", index);
    rec->task_id = obj + h->off_ids + (uint64_t)index * 16;
    rec->status = obj[h->off_status + index]; rec->has_result = obj[h->off_has + index];
    rec->data = obj + h->off_blob + off; rec->length = len; rec->index = index;
    return B9_OK;
}

int b9_sink_find(const uint8_t* obj, uint64_t size, const uint8_t* task_id, b9_sink_record* rec) {
    const b9sink::Header* h = b9sink::check(obj, size);
    if (!h || !task_id || !rec) return fail(B9_EINVAL, "b9_sink_find: not a sink object");
    const uint8_t* ids = obj + h->off_ids;
    for (uint32_t i = 0; i < h->n_records; ++i)
        if (memcmp(ids + (uint64_t)i * 16, task_id, 16) == 0) return b9_sink_get(obj, size, i, rec);
    return B9_ENOENT;
}

int64_t b9_sink_result_json(const uint8_t* data, uint64_t length, uint8_t* out, uint64_t cap) {
    if (length == 0) return 0;                                             // `if len(result) > 0`: the field stays unset
    if (!data || !out) return fail(B9_EINVAL, "b9_sink_result_json: NULL argument");
    uint64_t vs = 0, ve = 0;
    if (b9sink::raw_message(data, length, &vs, &ve)) {
        if (ve - vs > cap) return fail(B9_ENOSPC, "b9_sink_result_json: %llu bytes needed", (unsigned long long)(ve - vs));
        memcpy(out, data + vs, ve - vs);
        return (int64_t)(ve - vs);
    }
    const uint64_t need = 11 + b9sink::base64_len(length) + 2;             // {"base64":"..."}
    if (need > cap) return fail(B9_ENOSPC, "b9_sink_result_json: %llu bytes needed", (unsigned long long)need);
    memcpy(out, "{\"base64\":\"", 11);
    b9sink::base64_std(data, length, out + 11);
    memcpy(out + 11 + b9sink::base64_len(length), "\"}", 2);
    return (int64_t)need;
}

int b9_stats_get(b9_ctx* c, b9_stats* out) {
    if (!c || !out) return fail(B9_EINVAL, "b9_stats_get: NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    *out = c->stats;
    return B9_OK;
}

int b9_sync(b9_ctx* c) {
    if (!c) return fail(B9_EINVAL, "b9_sync: ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream_in));
    CU(cudaStreamSynchronize(c->stream));
    std::lock_guard<std::mutex> out_lk(c->out_mu);
    return finish_launch(c);
}

int b9_task_queue_scale(int64_t queue_length, int64_t tasks_per_container, int64_t max_containers, int64_t max_replicas, int* valid) {
    if (valid) *valid = 1;
    if (queue_length == 0) return 0;
    if (queue_length == -1 || tasks_per_container <= 0) { if (valid) *valid = 0; return 0; }
    int64_t desired = queue_length / tasks_per_container + (queue_length % tasks_per_container > 0 ? 1 : 0);
    return (int)std::min<int64_t>(std::min<int64_t>(max_containers, max_replicas), desired);
}

}  // extern "C"
