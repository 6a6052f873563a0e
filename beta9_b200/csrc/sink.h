// Result sink: one packed object per drain instead of one object-store PUT per task.
//
// Reference, per task: TaskQueueComplete uploads `in.Result` under "task/<id>/result"
// (pkg/abstractions/taskqueue/taskqueue.go:394-399 -> pkg/task/dispatch.go:18-20,120-144), and the REST read-back
// downloads it and applies `addResultToTask` (pkg/api/v1/task.go:295-325): valid JSON is kept as it is (json.RawMessage),
// anything else becomes {"base64":"..."}; an empty object leaves the field unset.
//
// Here: the drain's records land as ONE self-describing object — header, SoA index, result blob — written straight by the
// device-to-host copies of b9_drain_fetch_object (no per-task host work); the gateway uploads it once and keeps
// (object, record index) with the task. Pure host code, no CUDA in this header: the CPU tests call it through the library.
#pragma once
#include <stdint.h>
#include <string.h>

namespace b9sink {

constexpr uint64_t ALIGN = 64;
static inline uint64_t al(uint64_t v) { return (v + ALIGN - 1) & ~(ALIGN - 1); }

struct Header {                 // 128 bytes, little-endian
    char     magic[8];          // "B9SINK1\0"
    uint32_t n_records, version;
    uint64_t total_bytes, blob_bytes;
    uint64_t off_ids, off_offsets, off_lengths, off_status, off_has, off_blob;   // from the start of the object
    float    task_duration; uint32_t n_popped;
    uint8_t  reserved[40];
};
static_assert(sizeof(Header) == 128, "sink header layout");

struct Layout { uint64_t ids, offsets, lengths, status, has, blob, total; };
static inline Layout layout(uint64_t n, uint64_t blob_bytes) {
    Layout L; uint64_t o = sizeof(Header);
    L.ids = o; o = al(o + n * 16);
    L.offsets = o; o = al(o + n * 8);
    L.lengths = o; o = al(o + n * 4);
    L.status = o; o = al(o + n);
    L.has = o; o = al(o + n);
    L.blob = o; o = o + blob_bytes;
    L.total = o; return L;
}
static inline void write_header(uint8_t* obj, uint32_t n, uint64_t blob_bytes, float task_duration, uint32_t n_popped) {
    const Layout L = layout(n, blob_bytes);
    Header h; memset(&h, 0, sizeof h);
    memcpy(h.magic, "B9SINK1", 8);
    h.n_records = n; h.version = 1; h.total_bytes = L.total; h.blob_bytes = blob_bytes;
    h.off_ids = L.ids; h.off_offsets = L.offsets; h.off_lengths = L.lengths; h.off_status = L.status; h.off_has = L.has; h.off_blob = L.blob;
    h.task_duration = task_duration; h.n_popped = n_popped;
    memcpy(obj, &h, sizeof h);
}
// nullptr if the bytes are not a well-formed sink object of `size` bytes
static inline const Header* check(const uint8_t* obj, uint64_t size) {
    if (!obj || size < sizeof(Header)) return nullptr;
    const Header* h = (const Header*)obj;
    if (memcmp(h->magic, "B9SINK1", 8) != 0 || h->version != 1) return nullptr;
    const Layout L = layout(h->n_records, h->blob_bytes);
    if (h->total_bytes != L.total || size < L.total) return nullptr;
    if (h->off_ids != L.ids || h->off_offsets != L.offsets || h->off_lengths != L.lengths || h->off_status != L.status || h->off_has != L.has || h->off_blob != L.blob) return nullptr;
    return h;
}

// ---- json.Unmarshal(data, &rawMessage): Go's syntax check (encoding/json scanner.go checkValid), nothing converted.
// Returns true and the value's own span [*vs, *ve) (white space around it excluded) for valid JSON.
struct Scan {
    const uint8_t* d; uint64_t n, i;
    bool ws() { while (i < n && (d[i] == ' ' || d[i] == '\t' || d[i] == '\r' || d[i] == '\n')) ++i; return true; }
    static bool hex(uint8_t c) { return (c >= '0' && c <= '9') || ((c | 0x20) >= 'a' && (c | 0x20) <= 'f'); }
    bool string() {                     // d[i] == '"'
        ++i;
        while (i < n) {
            const uint8_t c = d[i];
            if (c == '"') { ++i; return true; }
            if (c < 0x20) return false;
            if (c == '\\') {
                if (i + 1 >= n) return false;
                const uint8_t e = d[i + 1];
                if (e == 'u') { if (i + 6 > n || !hex(d[i + 2]) || !hex(d[i + 3]) || !hex(d[i + 4]) || !hex(d[i + 5])) return false; i += 6; continue; }
                if (e == '"' || e == '\\' || e == '/' || e == 'b' || e == 'f' || e == 'n' || e == 'r' || e == 't') { i += 2; continue; }
                return false;
            }
            ++i;
        }
        return false;
    }
    bool number() {
        if (i < n && d[i] == '-') ++i;
        if (i >= n) return false;
        if (d[i] == '0') ++i;
        else if (d[i] >= '1' && d[i] <= '9') { while (i < n && d[i] >= '0' && d[i] <= '9') ++i; }
        else return false;
        if (i < n && d[i] == '.') { ++i; if (i >= n || d[i] < '0' || d[i] > '9') return false; while (i < n && d[i] >= '0' && d[i] <= '9') ++i; }
        if (i < n && (d[i] == 'e' || d[i] == 'E')) {
            ++i; if (i < n && (d[i] == '+' || d[i] == '-')) ++i;
            if (i >= n || d[i] < '0' || d[i] > '9') return false;
            while (i < n && d[i] >= '0' && d[i] <= '9') ++i;
        }
        return true;
    }
    bool lit(const char* t, uint64_t k) { if (i + k > n || memcmp(d + i, t, k) != 0) return false; i += k; return true; }
};
constexpr uint32_t MAX_DEPTH = 10000;   // scanner.go maxNestingDepth

static inline bool raw_message(const uint8_t* data, uint64_t n, uint64_t* vs, uint64_t* ve) {
    Scan s{data, n, 0};
    s.ws();
    if (s.i >= n) return false;
    *vs = s.i;
    // iterative descent: a bit per open container (1 = object), `expect` = what may come next
    static thread_local uint8_t stack[MAX_DEPTH / 8 + 2];
    uint32_t depth = 0;
    enum { VALUE, AFTER_VALUE, KEY_OR_END, KEY, VALUE_OR_END } st = VALUE;
    for (;;) {
        if (st == AFTER_VALUE && depth == 0) break;       // (before the white space that may follow the value)
        s.ws();
        if (s.i >= n) return false;
        const uint8_t c = data[s.i];
        switch (st) {
        case VALUE_OR_END:
            if (c == ']') { ++s.i; --depth; st = AFTER_VALUE; break; }
            /* fall through */
        case VALUE:
            if (c == '{' || c == '[') {
                if (depth + 1 > MAX_DEPTH) return false;
                const bool obj = c == '{';
                if (obj) stack[depth >> 3] |= (uint8_t)(1u << (depth & 7)); else stack[depth >> 3] &= (uint8_t)~(1u << (depth & 7));
                ++depth; ++s.i;
                st = obj ? KEY_OR_END : VALUE_OR_END;
            } else if (c == '"') { if (!s.string()) return false; st = AFTER_VALUE; }
            else if (c == '-' || (c >= '0' && c <= '9')) { if (!s.number()) return false; st = AFTER_VALUE; }
            else if (c == 't') { if (!s.lit("true", 4)) return false; st = AFTER_VALUE; }
            else if (c == 'f') { if (!s.lit("false", 5)) return false; st = AFTER_VALUE; }
            else if (c == 'n') { if (!s.lit("null", 4)) return false; st = AFTER_VALUE; }
            else return false;
            break;
        case KEY_OR_END:
            if (c == '}') { ++s.i; --depth; st = AFTER_VALUE; break; }
            /* fall through */
        case KEY:
            if (c != '"' || !s.string()) return false;
            s.ws();
            if (s.i >= n || data[s.i] != ':') return false;
            ++s.i; st = VALUE;
            break;
        case AFTER_VALUE: {
            const bool obj = (stack[(depth - 1) >> 3] >> ((depth - 1) & 7)) & 1u;
            if (c == ',') { ++s.i; st = obj ? KEY : VALUE; }
            else if (obj && c == '}') { ++s.i; --depth; }
            else if (!obj && c == ']') { ++s.i; --depth; }
            else return false;
            break;
        }
        }
    }
    *ve = s.i;
    s.ws();
    return s.i == n;
}

static inline uint64_t base64_len(uint64_t n) { return ((n + 2) / 3) * 4; }
static inline void base64_std(const uint8_t* p, uint64_t n, uint8_t* o) {
    static const char T[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    uint64_t i = 0;
    for (; i + 3 <= n; i += 3) { const uint32_t v = (p[i] << 16) | (p[i + 1] << 8) | p[i + 2]; *o++ = T[v >> 18]; *o++ = T[(v >> 12) & 63]; *o++ = T[(v >> 6) & 63]; *o++ = T[v & 63]; }
    if (n - i == 1) { const uint32_t v = p[i] << 16; *o++ = T[v >> 18]; *o++ = T[(v >> 12) & 63]; *o++ = '='; *o++ = '='; }
    else if (n - i == 2) { const uint32_t v = (p[i] << 16) | (p[i + 1] << 8); *o++ = T[v >> 18]; *o++ = T[(v >> 12) & 63]; *o++ = T[(v >> 6) & 63]; *o++ = '='; }
}

}  // namespace b9sink
