"""Model of the warp-wide tile copy's index arithmetic (beta9_b200/csrc/drain2.cuh d3_copy_tile): which copied task does an
aligned 16-byte vector of the tile's output start in (prefix-popcount over a bitmap with bit ceil(ex_r / 16) set), where do
its bytes sit in the stage buffer, and how is the vector around a task's end merged from two tasks. The CUDA is covered by
the GPU parity tests; this keeps the arithmetic honest on the CPU: every output byte is written exactly once and equals the
concatenation of the copied tokens."""
import random

import numpy as np
import pytest


def run_tile(lens_in, tok_off, c_lens, rng):
    soff = np.concatenate([[0], np.cumsum(lens_in)])[:-1]
    total = int(sum(lens_in))
    sbuf = rng.integers(0, 256, total + 96, dtype=np.uint8)
    ex = np.concatenate([[0], np.cumsum(c_lens)])[:-1]
    tb = int(sum(c_lens)); tb16 = (tb + 15) & ~15
    out = np.full(tb16, 0xEE, dtype=np.uint8); written = np.zeros(tb16, bool)
    expect = np.concatenate([sbuf[soff[k] + tok_off[k]:soff[k] + tok_off[k] + c_lens[k]] for k in range(32)])
    nzl = [k for k in range(32) if c_lens[k]]
    cnt = len(nzl); nblk = (tb + 511) >> 9
    bp = [[0, 0] for _ in range(32)]; ent = [None] * 33
    for r, k in enumerate(nzl):
        delta = int(soff[k] + tok_off[k] - ex[k])
        above = r + 1 < cnt
        dn = int(soff[nzl[r + 1]] + tok_off[nzl[r + 1]] - ex[nzl[r + 1]]) if above else delta
        ent[r] = (int(ex[k]), int(ex[k] + c_lens[k]) if above else 0xFFFFFFFF, delta, dn)
        if r:
            fv = (int(ex[k]) + 15) >> 4
            bp[fv >> 5][0] |= 1 << (fv & 31)
    run = 0
    for w in range(nblk):
        bp[w][1] = run; run += bin(bp[w][0]).count("1")
    for i in range(nblk):
        for lane in range(32):
            o = (i << 9) + (lane << 4)
            if o < tb:
                e = ent[bp[i][1] + bin(bp[i][0] & ((1 << (lane + 1)) - 1)).count("1")]
                if ((e[1] - o) & 0xFFFFFFFF) >= 16:
                    assert not written[o:o + 16].any()
                    out[o:o + 16] = sbuf[o + e[2]:o + e[2] + 16]; written[o:o + 16] = True
    for lane in range(32):
        if lane + 1 < cnt:
            e = ent[lane]; keep = e[1] & 15
            if keep:
                o = e[1] - keep
                a = sbuf[o + e[2]:o + e[2] + 16]; b = sbuf[o + e[3]:o + e[3] + 16]
                assert not written[o:o + 16].any()
                out[o:o + 16] = np.where(np.arange(16) < keep, a, b); written[o:o + 16] = True
    assert written.all()
    assert np.array_equal(out[:tb], expect)


@pytest.mark.parametrize("seed", range(3))
def test_every_byte_written_once_and_right(seed):
    rnd = random.Random(seed); rng = np.random.default_rng(seed)
    done = 0
    while done < 400:
        mode = rnd.random()
        lens, toff, cl = [], [], []
        for _k in range(32):
            L = 284 if mode < 0.3 else rnd.choice([284, 284, 300, 470, 44, 60, 700])
            c = L - 26 if rnd.random() < 0.9 else 0
            if mode > 0.8 and rnd.random() < 0.3:
                c = rnd.randint(16, L - 26) if L - 26 >= 16 else 0
            if c and c < 16:
                c = 0
            lens.append(L); toff.append(10); cl.append(c)
        if not 0 < sum(cl) <= 16384:
            continue
        run_tile(lens, toff, cl, rng)
        done += 1
