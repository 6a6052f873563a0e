// Micro-batching of single-task submissions from many threads.
//
// The reference admits ONE task per call, from one goroutine per request: `RedisTaskQueue.put`
// (pkg/abstractions/taskqueue/taskqueue.go:176-226) for gRPC / HTTP puts, and the endpoint's `RequestBuffer.ForwardRequest`
// (pkg/abstractions/endpoint/buffer.go:139-168), which pushes the request into a mutex-guarded `RingBuffer` that one
// processor goroutine pops one element at a time (buffer.go:170-195). Here the request threads append their payload
// straight into a page-locked arena — a lock-free reservation of (slot, byte range) by one compare-and-swap, then a memcpy
// into the reserved range — and a flush hands the arena to the device as ONE batch (b9_flush -> the ordinary push path).
// Two arenas alternate: submissions go on into the second while the first is on the wire.
//
// Pure C++ (atomics only, no CUDA): the library allocates the arenas page-locked, the CPU test (tests/host_shim/
// submit_buffer_shim.cpp) with malloc.
#pragma once
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <thread>

namespace b9 {

struct SubmitArena {
    uint8_t*  payload = nullptr; uint64_t cap_bytes = 0;
    uint64_t* offsets = nullptr;               // [cap_tasks + 1]
    uint8_t*  ids = nullptr;                   // [cap_tasks * 16]
    uint8_t*  flags = nullptr;                 // [cap_tasks]
    uint32_t  cap_tasks = 0;
    // bit 63 = sealed (a flush took the arena), bits 40..62 = tasks reserved, bits 0..39 = bytes reserved
    std::atomic<uint64_t> state{0};
    std::atomic<uint32_t> committed{0};        // reservations whose bytes are in place
    std::atomic<uint32_t> any_flags{0};        // OR of the tasks' flags (0: the flush need not send the flags array)
};

constexpr uint64_t SUBMIT_SEALED = 1ull << 63;
constexpr uint64_t SUBMIT_BYTES_MASK = (1ull << 40) - 1;
static inline uint32_t submit_count(uint64_t st) { return (uint32_t)((st & ~SUBMIT_SEALED) >> 40); }
static inline uint64_t submit_bytes(uint64_t st) { return st & SUBMIT_BYTES_MASK; }

enum { SUBMIT_OK = 0, SUBMIT_FULL = 1, SUBMIT_TOO_BIG = 2 };

struct SubmitBuffer {
    SubmitArena arena[2];
    std::atomic<int> active{0};
    std::mutex flush_mu;                       // one flusher at a time

    // Any thread. SUBMIT_FULL: the active arena cannot take this task — flush, then call again.
    int submit(const uint8_t* id16, const uint8_t* payload, uint32_t len, uint8_t flag) {
        for (;;) {
            SubmitArena& A = arena[active.load(std::memory_order_acquire)];
            uint64_t old = A.state.load(std::memory_order_acquire);
            if (old & SUBMIT_SEALED) { std::this_thread::yield(); continue; }       // a flush is switching arenas
            if ((uint64_t)len > A.cap_bytes) return SUBMIT_TOO_BIG;
            const uint32_t cnt = submit_count(old); const uint64_t bytes = submit_bytes(old);
            if (cnt >= A.cap_tasks || bytes + len > A.cap_bytes) return SUBMIT_FULL;
            if (!A.state.compare_exchange_weak(old, old + (1ull << 40) + len, std::memory_order_acq_rel)) continue;
            // slot `cnt` and bytes [bytes, bytes + len) are mine: submissions keep the order of their reservations (FIFO)
            A.offsets[cnt] = bytes;
            memcpy(A.ids + (size_t)cnt * 16, id16, 16);
            A.flags[cnt] = flag;
            if (flag) A.any_flags.fetch_or(flag, std::memory_order_relaxed);
            if (len) memcpy(A.payload + bytes, payload, len);
            A.committed.fetch_add(1, std::memory_order_release);
            return SUBMIT_OK;
        }
    }

    uint32_t buffered() const { return submit_count(arena[active.load(std::memory_order_acquire)].state.load(std::memory_order_acquire)); }

    // Caller holds flush_mu. Seals the active arena and makes the other one active (the caller has made sure it is free and
    // reset); waits for the in-flight memcpys of the sealed arena. Returns it with *n tasks / *bytes bytes, offsets[n] set.
    SubmitArena* seal(uint32_t* n, uint64_t* bytes) {
        const int cur = active.load(std::memory_order_acquire);
        SubmitArena& A = arena[cur];
        const uint64_t st = A.state.fetch_or(SUBMIT_SEALED, std::memory_order_acq_rel);
        *n = submit_count(st); *bytes = submit_bytes(st);
        active.store(cur ^ 1, std::memory_order_release);
        while (A.committed.load(std::memory_order_acquire) != *n) std::this_thread::yield();
        A.offsets[*n] = *bytes;
        return &A;
    }
    // Caller holds flush_mu; the arena's batch is on the device (or was empty): it may take submissions again.
    static void reset(SubmitArena& A) {
        A.committed.store(0, std::memory_order_relaxed);
        A.any_flags.store(0, std::memory_order_relaxed);
        A.state.store(0, std::memory_order_release);
    }
};

}  // namespace b9
