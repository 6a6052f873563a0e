// The drain kernel: deserialise -> handler -> serialise for every pending task of a window of the
// device ring, in ONE persistent launch.
//
//   * persistent thread blocks (grid = SMs x resident CTAs) take tiles of TILE_TASKS consecutive
//     ring slots from a global ticket counter (work stealing: a CTA that finishes early simply
//     takes the next ticket; there is no relaunch and no static tile->CTA assignment);
//   * per tile: headers are read once into shared memory, ready tasks (not cancelled/expired —
//     the ones TaskQueuePop would hand out, taskqueue.go:243-271) are compacted with warp ballots
//     + a block scan, the tile's result-byte count is chained to its predecessors with a
//     decoupled look-back, and results are written densely in FIFO order;
//   * reference behaviour realised per task:
//       pop + decode   pkg/abstractions/taskqueue/client.go:43-96, taskqueue.go:213-214
//       loads + call   sdk/src/beta9/runner/taskqueue.py:196-201,349-361
//       result         sdk/src/beta9/runner/taskqueue.py:378, runner/common.py:484-489
#pragma once
#include <stdint.h>
#include "json_device.cuh"
#include "handlers_device.cuh"
#include "handler_seq.cuh"

namespace b9 {

constexpr int TILE_TASKS    = 128;
constexpr int DRAIN_THREADS = 256;
constexpr int DRAIN_WARPS   = DRAIN_THREADS / 32;

// ring header word: len:32 | flags:8 | retries:8 | reserved:16
__host__ __device__ __forceinline__ uint64_t hdr_pack(uint32_t len, uint8_t flags, uint8_t retries) {
    return (uint64_t)len | ((uint64_t)flags << 32) | ((uint64_t)retries << 40);
}
__host__ __device__ __forceinline__ uint32_t hdr_len(uint64_t h) { return (uint32_t)h; }
__host__ __device__ __forceinline__ uint32_t hdr_flags(uint64_t h) { return (uint32_t)(h >> 32) & 0xFFu; }
constexpr uint32_t B9_TF_HTTP_BODY_BIT = 0x02u;      // == B9_TF_HTTP_BODY (include/b9gpu.h)

// tile look-back word: status:2 | bytes:38 | count:24
constexpr uint64_t LB_AGG = 1ull << 62, LB_INC = 2ull << 62, LB_STATUS = 3ull << 62;
__device__ __forceinline__ uint64_t lb_pack(uint64_t bytes, uint32_t count) { return (bytes << 24) | count; }
__device__ __forceinline__ uint64_t lb_value(uint64_t w) { return w & ~LB_STATUS; }

struct DrainCtl {
    unsigned long long ticket;      // next tile to hand out
    unsigned long long total;       // v1: lb_pack(result bytes, result count) of the whole window
    unsigned long long bytes;       // v2: result-byte cursor (one atomicAdd per tile); final value = total bytes
    unsigned int overflow;          // result staging too small
    unsigned int total_cnt;         // v2: result records of the whole window
    unsigned int n_slow;            // v2 identity: tasks deferred to the second kernel (escapes, foreign framing)
    unsigned int slow_head;         // v2 identity: next deferred task to take
};

// a task the main identity kernel could not settle with its quick look
struct SlowItem {
    uint64_t goff;                  // physical ring offset of the payload
    uint32_t len;                   // payload length; bit 31 = the SDK's canonical frame is present
    uint32_t j;                     // index of the task's result record
};

struct DrainArgs {
    // ring (inputs)
    const uint8_t*  payload;        // ring bytes
    const uint64_t* off;            // [ring_tasks] physical byte offset of each task's payload
    const uint64_t* hdr;            // [ring_tasks]
    const uint4*    ids;            // [ring_tasks] raw UUID
    uint32_t slot_mask;             // ring_tasks - 1
    uint64_t first_task;            // logical index of the window's first task
    uint32_t n_tasks;
    uint32_t n_tiles;
    // outputs
    uint8_t*  out_payload; uint64_t out_cap;
    uint64_t* out_off;              // [n_tasks + 1]  start of each record's bytes
    uint32_t* out_len;              // [n_tasks]      length of each record's bytes
    uint4*    out_ids;              // [n_tasks]
    uint8_t*  out_status;           // [n_tasks]
    uint8_t*  out_has;              // [n_tasks]
    // control
    DrainCtl* ctl;
    uint64_t* tile_state;           // v1: [n_tiles] look-back words, zeroed before launch
    const uint32_t* tile_base;      // v2, count_mode: [n_tiles] ready tasks before each warp-tile inside its 256-slot block (tile_count_kernel)
    const uint32_t* block_base;     // v2, count_mode: [n / 256 + 1] ready tasks before each 256-slot block (tile_scan_kernel); last = window total
    int handler;
    uint32_t count_mode;            // v2: 0 = no pending task is cancelled (record index = task index), 1 = record index from tile_base
    SlowItem* slow;                 // v2 identity: [n_tasks] work list for the second kernel
    const uint32_t* crc_shift_tabs; // v2 crc32: [levels][4][256] "advance the CRC register over 2^k zero bytes" tables
    uint32_t static_rounds;         // v2: a worker's first static_rounds tiles are worker + q * workers, the rest come from the ticket counter
};

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) { return *(const volatile uint64_t*)p; }
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) { *(volatile uint64_t*)p = v; }

// warp-cooperative byte copy global -> global (arbitrary alignment on both sides).
// 16-byte stores on the destination, 4-byte loads + funnel shift on the source.
__device__ inline void warp_copy(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n, int lane) {
    uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
    if (head > n) head = n;
    if (lane < (int)head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    uint32_t nvec = n >> 4;
    if (nvec) {
        uint32_t sh = (uint32_t)((uintptr_t)src & 3u);
        const uint32_t* sw = (const uint32_t*)(src - sh);
        uint4* dv = (uint4*)dst;
        uint32_t bits = sh * 8;
        for (uint32_t v = lane; v < nvec; v += 32) {
            const uint32_t* s4 = sw + 4 * v;
            uint32_t w0 = s4[0], w1 = s4[1], w2 = s4[2], w3 = s4[3];   // plain loads: the source may be shared memory
            uint4 o;
            if (sh) {
                uint32_t w4 = s4[4];           // may read up to 3 bytes past the payload: the ring / stage buffer has slack
                o.x = __funnelshift_r(w0, w1, bits); o.y = __funnelshift_r(w1, w2, bits);
                o.z = __funnelshift_r(w2, w3, bits); o.w = __funnelshift_r(w3, w4, bits);
            } else { o.x = w0; o.y = w1; o.z = w2; o.w = w3; }
            dv[v] = o;
        }
    }
    uint32_t done = nvec << 4, tail = n - done;
    if (lane < (int)tail) dst[done + lane] = src[done + lane];
}

// ---- the SDK's canonical frame around one string argument ------------------------------------
// json.dumps({"args": (s,), "kwargs": {}})  ->  {"args": ["<body>"], "kwargs": {}}
__device__ __constant__ uint8_t FRAME_PRE[11] = {'{', '"', 'a', 'r', 'g', 's', '"', ':', ' ', '[', '"'};
__device__ __constant__ uint8_t FRAME_SUF[17] = {'"', ']', ',', ' ', '"', 'k', 'w', 'a', 'r', 'g', 's', '"', ':', ' ', '{', '}', '}'};
constexpr uint32_t FRAME_PRE_LEN = 11, FRAME_SUF_LEN = 17;

// Result of the warp-level look at one payload.
struct Quick {
    bool framed;        // canonical frame present and the body is "clean" (printable ASCII, no '"' or '\\')
    bool maybe_framed;  // frame bytes present but the body needs the escape-aware check
};

__device__ inline Quick quick_frame(const uint8_t* __restrict__ p, uint32_t len, int lane) {
    Quick q; q.framed = false; q.maybe_framed = false;
    if (len < FRAME_PRE_LEN + FRAME_SUF_LEN) return q;
    bool ok = true;
    if (lane < (int)FRAME_PRE_LEN) ok = p[lane] == FRAME_PRE[lane];
    else if (lane < (int)(FRAME_PRE_LEN + FRAME_SUF_LEN)) ok = p[len - FRAME_SUF_LEN + (lane - FRAME_PRE_LEN)] == FRAME_SUF[lane - FRAME_PRE_LEN];
    if (!__all_sync(0xffffffffu, ok)) return q;
    uint32_t b0 = FRAME_PRE_LEN, b1 = len - FRAME_SUF_LEN;
    bool special = false;
    for (uint32_t i = b0 + lane; i < b1; i += 32) {
        uint8_t c = p[i];
        special |= (c < 0x20) | (c >= 0x7F) | (c == '"') | (c == '\\');
    }
    if (__any_sync(0xffffffffu, special)) q.maybe_framed = true; else q.framed = true;
    return q;
}

template <int HANDLER>
__global__ void __launch_bounds__(DRAIN_THREADS, 2) drain_kernel(DrainArgs a) {
    __shared__ TaskRec s_rec[TILE_TASKS];
    __shared__ uint32_t s_excl_bytes[TILE_TASKS];   // exclusive prefix of out_len inside the tile
    __shared__ uint32_t s_excl_cnt[TILE_TASKS];     // exclusive prefix of ready inside the tile
    __shared__ uint32_t s_warp_bytes[DRAIN_WARPS], s_warp_cnt[DRAIN_WARPS];
    __shared__ unsigned long long s_tile;
    __shared__ uint64_t s_base;                     // lb_pack(exclusive bytes, exclusive count) of this tile

    __shared__ uint32_t s_crc_table[HANDLER == 1 ? 256 : 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (HANDLER == 1) { s_crc_table[tid & 255] = crc_table_entry(tid & 255); __syncthreads(); }

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(&a.ctl->ticket, 1ull);
        __syncthreads();
        const unsigned long long tile = s_tile;
        if (tile >= a.n_tiles) break;
        const uint32_t t0 = (uint32_t)tile * TILE_TASKS;
        const uint32_t nt = min((uint32_t)TILE_TASKS, a.n_tasks - t0);

        // ---------------- phase A: deserialise, run the handler's sizing pass --------------------
        if (HANDLER != 0) {
            // v1 of the non-identity handlers: one thread per task, start to finish
            for (uint32_t k = tid; k < nt; k += DRAIN_THREADS) {
                const uint32_t slot = (uint32_t)((a.first_task + t0 + k) & a.slot_mask);
                const uint64_t h = __ldg(a.hdr + slot);
                TaskRec rec; rec.ready = !(hdr_flags(h) & 1u); rec.status = 0; rec.has = 0; rec.mode = OM_NONE; rec.out_len = 0;
                rec.src_off = 0; rec.src_len = 0; rec.value = 0;
                if (rec.ready) {
                    const uint8_t* p = a.payload + __ldg(a.off + slot);
                    Parsed pr = parse_payload(p, hdr_len(h), (hdr_flags(h) & B9_TF_HTTP_BODY_BIT) != 0);
                    handler_phase_a(HANDLER, p, pr, rec, s_crc_table);
                }
                s_rec[k] = rec;
            }
        } else
        for (uint32_t k = warp; k < nt; k += DRAIN_WARPS) {
            const uint32_t slot = (uint32_t)((a.first_task + t0 + k) & a.slot_mask);
            const uint64_t h = __ldg(a.hdr + slot);
            const uint32_t len = hdr_len(h);
            const bool ready = !(hdr_flags(h) & 1u);
            TaskRec rec; rec.ready = ready; rec.status = 0; rec.has = 0; rec.mode = OM_NONE; rec.out_len = 0;
            rec.src_off = 0; rec.src_len = 0; rec.value = 0;
            if (ready) {
                const uint8_t* p = a.payload + __ldg(a.off + slot);
                const bool http = (hdr_flags(h) & B9_TF_HTTP_BODY_BIT) != 0;
                Quick q = quick_frame(p, len, lane);
                if (q.framed && HANDLER == 0 && !http) {
                    // args == [body], kwargs == {}; body is printable ASCII without '"' or '\\':
                    // json.dumps(body) is the token itself
                    uint32_t tok = len - FRAME_PRE_LEN - FRAME_SUF_LEN + 2;
                    if (tok > 2) { rec.has = 1; rec.mode = OM_COPY; rec.src_off = FRAME_PRE_LEN - 1; rec.src_len = tok; rec.out_len = tok; }
                } else {
                    if (lane == 0) {
                        Parsed pr = parse_payload(p, len, http);
                        handler_phase_a(HANDLER, p, pr, rec);
                    }
                }
            }
            if (lane == 0) s_rec[k] = rec;
        }
        __syncthreads();

        // ---------------- compaction + sizes: block scan over the tile ---------------------------
        uint32_t my_bytes = 0, my_cnt = 0;
        if (tid < (int)nt) { my_bytes = s_rec[tid].out_len; my_cnt = s_rec[tid].ready ? 1u : 0u; }
        // ready tasks are compacted by ballot; byte sizes by an inclusive warp scan
        const uint32_t ready_mask = __ballot_sync(0xffffffffu, my_cnt);
        uint32_t inc_bytes = my_bytes;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, inc_bytes, d); if (lane >= d) inc_bytes += v; }
        if (lane == 31) { s_warp_bytes[warp] = inc_bytes; s_warp_cnt[warp] = __popc(ready_mask); }
        __syncthreads();
        uint32_t wb = 0, wc = 0, tb = 0, tc = 0;
        #pragma unroll
        for (int w = 0; w < DRAIN_WARPS; ++w) {
            uint32_t b = s_warp_bytes[w], c = s_warp_cnt[w];
            if (w < warp) { wb += b; wc += c; }
            tb += b; tc += c;
        }
        if (tid < (int)nt) {
            s_excl_bytes[tid] = wb + inc_bytes - my_bytes;
            s_excl_cnt[tid] = wc + __popc(ready_mask & ((1u << lane) - 1u));
        }

        // ---------------- decoupled look-back for the tile's global base -------------------------
        if (warp == 0) {
            const uint64_t agg = lb_pack(tb, tc);
            uint64_t excl = 0;
            if (tile == 0) {
                if (lane == 0) st_volatile_u64(a.tile_state + 0, LB_INC | agg);
            } else {
                if (lane == 0) st_volatile_u64(a.tile_state + tile, LB_AGG | agg);
                long long look = (long long)tile - 1;
                for (;;) {
                    long long idx = look - lane;
                    uint64_t w = (idx >= 0) ? ld_volatile_u64(a.tile_state + idx) : LB_INC;
                    while (__any_sync(0xffffffffu, (w & LB_STATUS) == 0)) {
                        if ((w & LB_STATUS) == 0) w = ld_volatile_u64(a.tile_state + idx);
                    }
                    const uint32_t inc_mask = __ballot_sync(0xffffffffu, (w & LB_STATUS) == LB_INC);
                    uint64_t v = lb_value(w);
                    if (inc_mask) {
                        const int first = __ffs(inc_mask) - 1;
                        if (lane > first) v = 0;
                    }
                    #pragma unroll
                    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                    excl += v;
                    if (inc_mask) break;
                    look -= 32;
                }
                if (lane == 0) st_volatile_u64(a.tile_state + tile, LB_INC | (excl + agg));
            }
            if (lane == 0) {
                s_base = excl;
                if (tile == a.n_tiles - 1) {
                    const uint64_t tot = excl + agg;
                    a.ctl->total = tot;
                    a.out_off[(uint32_t)(tot & 0xFFFFFFu)] = tot >> 24;       // terminal offset
                }
            }
        }
        __syncthreads();
        const uint64_t base_bytes = s_base >> 24;
        const uint32_t base_cnt = (uint32_t)(s_base & 0xFFFFFFu);
        const bool fits = base_bytes + tb <= a.out_cap;
        if (!fits && tid == 0) a.ctl->overflow = 1u;

        // ---------------- phase B: serialise ------------------------------------------------------
        if (HANDLER != 0) {
            for (uint32_t k = tid; k < nt; k += DRAIN_THREADS) {
                const TaskRec rec = s_rec[k];
                if (!rec.ready) continue;
                const uint32_t slot = (uint32_t)((a.first_task + t0 + k) & a.slot_mask);
                const uint32_t j = base_cnt + s_excl_cnt[k];
                const uint64_t ob = base_bytes + s_excl_bytes[k];
                a.out_off[j] = ob; a.out_len[j] = rec.out_len; a.out_ids[j] = __ldg(a.ids + slot); a.out_status[j] = rec.status; a.out_has[j] = rec.has;
                if (!rec.has || !fits) continue;
                uint8_t* o = a.out_payload + ob;
                if (rec.mode == OM_VADD) vadd_write(a.payload + __ldg(a.off + slot), rec.src_off, rec.src_len, o);
                else {
                    long long v = rec.value; uint32_t l = rec.out_len;
                    if (v < 0) { *o++ = '-'; --l; v = -v; }
                    write_dec(o, (unsigned long long)v, l);
                }
            }
        } else
        for (uint32_t k = warp; k < nt; k += DRAIN_WARPS) {
            const TaskRec rec = s_rec[k];
            if (!rec.ready) continue;
            const uint32_t slot = (uint32_t)((a.first_task + t0 + k) & a.slot_mask);
            const uint32_t j = base_cnt + s_excl_cnt[k];
            const uint64_t ob = base_bytes + s_excl_bytes[k];
            if (lane == 0) {
                a.out_off[j] = ob;
                a.out_len[j] = rec.out_len;
                a.out_ids[j] = __ldg(a.ids + slot);
                a.out_status[j] = rec.status;
                a.out_has[j] = rec.has;
            }
            if (!rec.has || !fits) continue;
            const uint8_t* p = a.payload + __ldg(a.off + slot);
            uint8_t* o = a.out_payload + ob;
            switch (rec.mode) {
            case OM_COPY:
                warp_copy(o, p + rec.src_off, rec.src_len, lane);
                break;
            case OM_STR_ESC:
                if (lane == 0) {
                    uint32_t i = rec.src_off + 1, end = rec.src_off + rec.src_len - 1;
                    *o++ = '"';
                    while (i < end) o += py_emit(next_cp(p, i, end), o);
                    *o = '"';
                }
                break;
            case OM_U32_DEC: case OM_I64_DEC:
                if (lane == 0) {
                    long long v = rec.value; uint32_t l = rec.out_len;
                    if (v < 0) { *o++ = '-'; --l; v = -v; }
                    write_dec(o, (unsigned long long)v, l);
                }
                break;
            default: break;
            }
        }
        __syncthreads();   // s_rec / s_base are reused by the next tile
    }
}

// Marks expired pending tasks as cancelled (Dispatcher.monitor's unclaimed branch, dispatch.go:173-230).
__global__ void expire_kernel(uint64_t* __restrict__ hdr, const int64_t* __restrict__ expires, uint32_t slot_mask,
                              uint64_t first_task, uint32_t n, int64_t now_ns, unsigned long long* __restrict__ count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t slot = (uint32_t)((first_task + i) & slot_mask);
    int64_t e = expires[slot];
    uint64_t h = hdr[slot];
    if (e != 0 && e <= now_ns && !(hdr_flags(h) & 1u)) {
        hdr[slot] = h | (1ull << 32);
        atomicAdd(count, 1ull);
    }
}

// Turns a pushed batch's relative offsets into ring slots (off, hdr, cold metadata).
__global__ void ingest_kernel(const uint64_t* __restrict__ rel_off, uint32_t n, uint64_t seg_start, uint64_t first_task, uint32_t slot_mask,
                              const int64_t* __restrict__ ts, const int64_t* __restrict__ exp, const uint8_t* __restrict__ retries,
                              const uint8_t* __restrict__ flags, uint64_t* __restrict__ off, uint64_t* __restrict__ hdr,
                              int64_t* __restrict__ ring_ts, int64_t* __restrict__ ring_exp) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t slot = (uint32_t)((first_task + i) & slot_mask);
    uint64_t o0 = rel_off[0], o = rel_off[i], o1 = rel_off[i + 1];
    off[slot] = seg_start + (o - o0);
    hdr[slot] = hdr_pack((uint32_t)(o1 - o), flags ? flags[i] : 0, retries ? retries[i] : 0);
    ring_ts[slot] = ts ? ts[i] : 0;
    ring_exp[slot] = exp ? exp[i] : 0;
}

}  // namespace b9
