"""Python model of the device's warp-cooperative "canonical escapes" check (beta9_b200/csrc/drain2.cuh
esc_verify_canonical): a framed string body whose escapes are exactly the ones json.dumps writes comes
back from identity as the very same token, except that lone surrogate escapes become \\ufffd (Go's
decoder, oracle/pyoracle/gojson.py:212-264). The model mirrors the device's data flow — 32 lanes x 16
bytes per pass, per-lane backslash / quote bit masks, the odd-backslash-run carry resolved across
lanes by ballots, then one step per escape start — so that the rule set can be fuzzed against the
oracle on a CPU (tests/test_esc_verify_model.py); the GPU tests then cover the CUDA transcription.

verify(body) -> None (not decided here: the sequential / general path decides)
             or  bytes   (the patched body: json.dumps(identity(s)) == b'"' + patched + b'"')
"""
from __future__ import annotations

from typing import List, Optional

EVEN = 0x5555


def esc16(bs: int, carry_in: int, cnt: int = 16):
    """16 positions, the first `cnt` of them real bytes (the rest is filler, never a backslash): which are
    escaped (follow an escape-start backslash), and do the real bytes end on an unmatched escape start
    (carry out). simdjson's find_escaped on a 16-bit word."""
    bs &= ~carry_in & 0xFFFF
    follows = ((bs << 1) | carry_in) & 0xFFFF
    odd_starts = bs & ~EVEN & ~follows & 0xFFFF
    seq_even = odd_starts + bs                      # 17 bits
    carry_out = (seq_even >> 16) & 1
    invert = (seq_even << 1) & 0xFFFF
    escaped = (EVEN ^ invert) & follows & 0xFFFF
    if cnt < 16:
        carry_out = (escaped >> cnt) & 1            # an escape start on the last real byte "escapes" the filler
    return escaped, carry_out


def _hex4(b: bytes) -> int:
    try:
        if len(b) != 4 or any(c not in b"0123456789abcdefABCDEF" for c in b):
            return -1
        return int(b, 16)
    except ValueError:
        return -1


def _bs_run_before(body: bytes, q: int) -> int:
    r = 0
    while r < q and body[q - 1 - r] == 0x5C:
        r += 1
    return r


def _is_u_escape_in(body: bytes, q: int, lo: int, hi: int) -> bool:
    n = len(body)
    if q < 0 or q + 6 > n or body[q] != 0x5C or body[q + 1] != ord("u"):
        return False
    h = _hex4(body[q + 2:q + 6])
    return lo <= h <= hi


def verify(body: bytes) -> Optional[bytes]:
    n = len(body)
    out = bytearray(body)
    carry = 0
    for base0 in range(0, n, 512):
        # ---- per lane: masks of its 16 bytes
        bs_m: List[int] = []
        qt_m: List[int] = []
        cnts: List[int] = []
        bad = False
        for lane in range(32):
            base = base0 + 16 * lane
            chunk = body[base:base + 16] if base < n else b""
            bs = qt = 0
            for j, c in enumerate(chunk):
                if c == 0x5C:
                    bs |= 1 << j
                if c == 0x22:
                    qt |= 1 << j
                if c < 0x20 or c >= 0x7F:
                    bad = True                       # raw control byte (invalid) / DEL or non-ASCII (json.dumps escapes them)
            bs_m.append(bs)
            qt_m.append(qt)
            cnts.append(len(chunk))
        if bad:
            return None
        # ---- carry into every lane: each lane's chunk is const0 / const1 / passes its carry-in on (negated or not)
        out0 = [esc16(bs_m[l], 0, cnts[l])[1] for l in range(32)]
        out1 = [esc16(bs_m[l], 1, cnts[l])[1] for l in range(32)]
        K = sum(1 << l for l in range(32) if out0[l] != out1[l])      # lanes whose carry-out depends on their carry-in
        V0 = sum(1 << l for l in range(32) if out0[l])
        cin = []
        for lane in range(33):
            lt = (1 << lane) - 1
            below = ~K & lt & 0xFFFFFFFF
            if below:
                j = below.bit_length() - 1
                rng = lt & ~((1 << j) - 1)
                c = bin(V0 & rng).count("1") & 1
            else:
                c = (bin(V0 & lt).count("1") & 1) ^ carry
            cin.append(c)
        patches: List[int] = []
        for lane in range(32):
            base = base0 + 16 * lane
            escaped, co = esc16(bs_m[lane], cin[lane], cnts[lane])
            assert co == (out1[lane] if cin[lane] else out0[lane])
            if qt_m[lane] & ~escaped:
                return None                           # a raw quote inside the body
            bsl = bs_m[lane] & ~(cin[lane] & 1)
            starts = bsl & ~escaped
            while starts:
                j = (starts & -starts).bit_length() - 1
                starts &= starts - 1
                i = base + j
                t = body[i + 1] if i + 1 < n else 0x22  # (the closing quote follows the body)
                if t == ord("u"):
                    if i + 6 > n:
                        return None
                    hx = body[i + 2:i + 6]
                    h = _hex4(hx)
                    if h < 0 or any(c in b"ABCDEF" for c in hx):
                        return None
                    if h < 0x20:
                        if h in (8, 9, 10, 12, 13):
                            return None               # json.dumps writes \b \t \n \f \r
                    elif h < 0x7F:
                        return None                   # json.dumps writes the character itself (or \" \\)
                    elif 0xD800 <= h < 0xE000:
                        if h < 0xDC00:
                            lone = not _is_u_escape_in(body, i + 6, 0xDC00, 0xDFFF)
                        else:
                            lone = not (i >= 6 and _is_u_escape_in(body, i - 6, 0xD800, 0xDBFF) and not (_bs_run_before(body, i - 6) & 1))
                        if lone:
                            patches.append(i)
                elif t not in b'"\\bfnrt':
                    return None                       # "\/" changes length, anything else is not JSON
        carry = cin[32]
        for i in patches:
            out[i + 2:i + 6] = b"fffd"
    if carry:
        return None                                   # the body ends inside an escape: the frame's quote is not the closing one
    return bytes(out)


def sequential_escaped(bs_positions: List[bool]) -> List[bool]:
    """reference definition for esc16's test: walk left to right."""
    esc = [False] * len(bs_positions)
    i = 0
    pending = False
    for i, b in enumerate(bs_positions):
        if pending:
            esc[i] = True
            pending = False
        elif b:
            pending = True
    return esc
