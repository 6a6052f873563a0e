/* Minimal C (C99) client of libb9gpu.so: push two SDK payloads and one HTTP body, drain them through the
 * identity handler, print the records. Doubles as the proof that include/b9gpu.h is a plain C header
 * (tests/test_abi_symbols.py compiles this file with `gcc -std=c99 -pedantic -Wall -Werror -fsyntax-only`).
 *
 *   gcc -std=c99 -Iinclude examples/c_api_demo.c -Lbeta9_b200 -lb9gpu -Wl,-rpath,$PWD/beta9_b200 -o c_api_demo
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b9gpu.h"

int main(void) {
    b9_opts opts;
    memset(&opts, 0, sizeof opts);
    opts.struct_size = (uint32_t)sizeof opts;
    b9_ctx *ctx = NULL;
    if (b9_ctx_create(&opts, &ctx) != B9_OK) {          /* B9_ENODEV without a GPU: there is no CPU path */
        fprintf(stderr, "b9_ctx_create: %s\n", b9_last_error(NULL));
        return 1;
    }
    const char *p0 = "{\"args\": [\"hello\"], \"kwargs\": {}}";      /* the SDK's put payload            */
    const char *p1 = "{\"args\": [\"caf\\u00e9 \\\"x\\\"\"], \"kwargs\": {}}";
    const char *p2 = "{\"args\":[\"from an HTTP body\"]}";           /* raw request body: B9_TF_HTTP_BODY */
    const char *ps[3];
    uint8_t ids[3 * 16], flags[3] = {0, 0, B9_TF_HTTP_BODY}, blob[512];
    uint64_t off[4] = {0, 0, 0, 0};
    b9_push_meta meta;
    int i;
    ps[0] = p0; ps[1] = p1; ps[2] = p2;
    for (i = 0; i < 3; ++i) {
        memset(ids + 16 * i, i + 1, 16);
        memcpy(blob + off[i], ps[i], strlen(ps[i]));
        off[i + 1] = off[i] + strlen(ps[i]);
    }
    memset(&meta, 0, sizeof meta);
    meta.flags = flags;
    if (b9_batch_push(ctx, ids, blob, off, 3, &meta) != B9_OK) { fprintf(stderr, "push: %s\n", b9_last_error(ctx)); return 1; }
    printf("pending: %llu\n", (unsigned long long)b9_depth(ctx));

    uint8_t out_ids[3 * 16], status[3], has[3], payload[1024];
    uint64_t offsets[3];
    uint32_t lengths[3];
    b9_results res;
    memset(&res, 0, sizeof res);
    res.task_ids = out_ids; res.status = status; res.has_result = has; res.offsets = offsets; res.lengths = lengths;
    res.payload = payload; res.cap_tasks = 3; res.cap_bytes = sizeof payload;
    {
        const int64_t n = b9_drain(ctx, b9_handler_id("identity"), 3, &res);
        if (n < 0) { fprintf(stderr, "drain: %s\n", b9_last_error(ctx)); return 1; }
        for (i = 0; i < (int)n; ++i)
            printf("task %02x status %u result %.*s\n", out_ids[16 * i], status[i], has[i] ? (int)lengths[i] : 6,
                   has[i] ? (const char *)payload + offsets[i] : "(none)");
    }
    b9_ctx_destroy(ctx);
    return 0;
}
