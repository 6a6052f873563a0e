// Host harness around beta9_b200/csrc/submit_buffer.h: P producer threads submit T tasks each while a flusher thread
// seals and collects arenas; returns what came out so that the test can check nothing is lost, duplicated or torn.
//   g++ -O2 -std=c++17 -shared -fPIC -pthread -o libsubmitbuffer.so submit_buffer_shim.cpp
#include <cstdlib>
#include <thread>
#include <vector>
#include "../../beta9_b200/csrc/submit_buffer.h"

using namespace b9;

extern "C" {

// Every task's payload is `len(p, t)` bytes of the value (p * 131 + t) & 0xFF, its id = {p, t} in the first 8 bytes.
// out_ids: [P * T * 16], out_lens: [P * T], out_ok: [P * T] (1 = payload bytes all right), in the order the flushes delivered.
// Returns the number of tasks delivered, or -1 on an internal inconsistency.
long b9_submit_buffer_run(int producers, int tasks_per_producer, unsigned cap_tasks, unsigned long long cap_bytes, int flush_every_us,
                          unsigned char* out_ids, unsigned* out_lens, unsigned char* out_ok) {
    SubmitBuffer sb;
    for (int k = 0; k < 2; ++k) {
        SubmitArena& A = sb.arena[k];
        A.payload = (uint8_t*)malloc(cap_bytes); A.offsets = (uint64_t*)malloc(((size_t)cap_tasks + 1) * 8);
        A.ids = (uint8_t*)malloc((size_t)cap_tasks * 16); A.flags = (uint8_t*)malloc(cap_tasks);
        A.cap_bytes = cap_bytes; A.cap_tasks = cap_tasks;
    }
    std::mutex out_mu; long delivered = 0; bool bad = false;
    auto flush = [&]() {                                                  // what b9_flush does, minus the device
        std::lock_guard<std::mutex> lk(sb.flush_mu);
        if (sb.buffered() == 0) return;
        const int cur = sb.active.load();
        SubmitBuffer::reset(sb.arena[cur ^ 1]);                            // (its previous batch was consumed synchronously below)
        uint32_t n = 0; uint64_t bytes = 0;
        SubmitArena* A = sb.seal(&n, &bytes);
        std::lock_guard<std::mutex> ok(out_mu);
        for (uint32_t i = 0; i < n; ++i) {
            const uint64_t o0 = A->offsets[i], o1 = A->offsets[i + 1];
            if (o1 < o0 || o1 > bytes) { bad = true; return; }
            memcpy(out_ids + (size_t)delivered * 16, A->ids + (size_t)i * 16, 16);
            out_lens[delivered] = (unsigned)(o1 - o0);
            const uint32_t p = *(const uint32_t*)(A->ids + (size_t)i * 16), t = *(const uint32_t*)(A->ids + (size_t)i * 16 + 4);
            const uint8_t v = (uint8_t)(p * 131u + t);
            unsigned char okb = A->flags[i] == (uint8_t)(t & 3u);
            for (uint64_t b = o0; b < o1; ++b) okb &= A->payload[b] == v;
            out_ok[delivered] = okb;
            ++delivered;
        }
    };
    std::atomic<int> live{producers};
    std::vector<std::thread> th;
    for (int p = 0; p < producers; ++p) th.emplace_back([&, p] {
        std::vector<uint8_t> buf(4096);
        for (int t = 0; t < tasks_per_producer; ++t) {
            uint8_t id[16] = {0};
            *(uint32_t*)id = (uint32_t)p; *(uint32_t*)(id + 4) = (uint32_t)t;
            const uint32_t len = (uint32_t)((p * 37 + t * 13) % 700);
            memset(buf.data(), (uint8_t)(p * 131u + t), len);
            for (;;) {
                const int r = sb.submit(id, buf.data(), len, (uint8_t)(t & 3));
                if (r == SUBMIT_OK) break;
                if (r == SUBMIT_TOO_BIG) { bad = true; break; }
                flush();                                                   // full: the submitting thread flushes, like b9_submit
            }
        }
        live.fetch_sub(1);
    });
    std::thread flusher([&] {
        while (live.load() > 0) { std::this_thread::sleep_for(std::chrono::microseconds(flush_every_us)); flush(); }
    });
    for (auto& t : th) t.join();
    flusher.join();
    flush();
    for (int k = 0; k < 2; ++k) { free(sb.arena[k].payload); free(sb.arena[k].offsets); free(sb.arena[k].ids); free(sb.arena[k].flags); }
    return bad ? -1 : delivered;
}

}  // extern "C"
