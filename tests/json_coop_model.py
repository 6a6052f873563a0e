"""Python model (test infrastructure) of the rule set `json_sum_coop` applies in
beta9_b200/csrc/drain2.cuh: the same class masks and look-behind rules, with Python integers as
bit masks over the whole document instead of one 32-bit word per lane. It exists so that the
ACCEPTANCE logic of the device fast path can be fuzzed against the oracle on a CPU: whenever the
model decides a payload, the oracle must agree with the decision."""
from __future__ import annotations

PRE = b'{"args": ['
SUF = b'], "kwargs": {}}'
MAX_DOC = 4096


def _prefix_xor(m: int, n: int) -> int:
    out, par = 0, 0
    for i in range(n):
        par ^= (m >> i) & 1
        out |= par << i
    return out


def decide(payload: bytes):
    """None when the fast path does not decide the payload, else the integer sum."""
    if len(payload) < len(PRE) + len(SUF) + 2 or not payload.startswith(PRE) or not payload.endswith(SUF):
        return None
    D = payload[len(PRE):len(payload) - len(SUF)]
    n = len(D)
    if n > MAX_DOC:
        return None
    full = (1 << n) - 1

    def mask(pred):
        m = 0
        for i, c in enumerate(D):
            if pred(c):
                m |= 1 << i
        return m
    BAD = mask(lambda c: c < 0x20 or c >= 0x7F or c == 0x5C)
    if BAD:
        return None
    Q = mask(lambda c: c == 0x22)
    DG = mask(lambda c: 0x30 <= c <= 0x39)
    ZR = mask(lambda c: c == 0x30)
    CM, CL, SP = mask(lambda c: c == 0x2C), mask(lambda c: c == 0x3A), mask(lambda c: c == 0x20)
    OB, CB, LB, RB = mask(lambda c: c == 0x5B), mask(lambda c: c == 0x5D), mask(lambda c: c == 0x7B), mask(lambda c: c == 0x7D)
    OTH = full & ~(Q | DG | CM | CL | SP | OB | CB | LB | RB)
    qinc = _prefix_xor(Q, n)
    OPENQ, CLOSEQ = Q & qinc, Q & ~qinc
    out = ~(qinc & ~Q) & full
    v = OTH & out
    dg, zr, cm, cl, sp = DG & out, ZR & out, CM & out, CL & out, SP & out
    ob, cb, lb, rb = OB & out, CB & out, LB & out, RB & out
    br = ob | cb
    binc = _prefix_xor(br, n)
    v |= (ob & ~binc) | (cb & binc)
    arr = binc & ~ob & full
    v |= cl & arr
    sep = (cm | cl) & ~arr
    sinc = _prefix_xor(sep, n)
    v |= (cl & ~sinc) | (cm & ~arr & sinc)
    quotes_odd, br_odd, sep_odd = bin(Q).count("1") & 1, bin(br).count("1") & 1, bin(sep).count("1") & 1

    def P(m, k=1):
        return (m << k) & full
    a_closeq, a_dg, a_cm, a_cl, a_sp, a_ob, a_cb, a_zr, a_lb = P(CLOSEQ), P(dg), P(cm), P(cl), P(sp), P(ob), P(cb), P(zr), P(lb)
    t_cm, t_cl = a_cm | (a_sp & P(cm, 2)), a_cl | (a_sp & P(cl, 2))
    ds = dg & ~a_dg
    v |= sp & ~(a_cm | a_cl)
    v |= a_closeq & ~(cl | cm | rb)
    v |= a_dg & ~dg & ~(cm | cb | rb)
    v |= a_cb & ~(cm | rb)
    v |= dg & a_zr & ~P(dg, 2)
    v |= OPENQ & (arr | ~(a_lb | t_cm | t_cl))
    v |= cl & ~a_closeq
    v |= cm & ~arr & ~(a_closeq | a_dg | a_cb)
    v |= cm & arr & ~a_dg
    v |= ob & ~t_cl
    v |= cb & ~(a_ob | a_dg)
    v |= ds & ((arr & ~(a_ob | t_cm)) | (~arr & ~t_cl))
    v |= lb ^ 1
    v |= rb ^ (1 << (n - 1))
    v |= rb & ~(a_lb | a_closeq | a_dg | a_cb)
    if rb and not (rb & a_lb) and not sep_odd:
        v |= 1
    if (v & full) or quotes_odd or br_odd:
        return None
    cand = cl & a_closeq & P(OPENQ, 8)
    best = -1
    for j in range(n):
        if (cand >> j) & 1 and D[j - 7:j - 1] == b"values":
            best = j
    if best < 0:
        return None
    vs = best + 1
    if D[vs] == 0x20:
        vs += 1
    if D[vs] != 0x5B:
        return None
    ve = D.index(b"]", vs)      # the first ']' after vs is outside any string: strings cannot start inside an array
    total, i = 0, 0
    while i < n:
        if (dg >> i) & 1:
            j = i
            while (dg >> j) & 1:
                j += 1
            if j - i > 15:
                return None
            if vs < i < ve:
                total += int(D[i:j])
            i = j
        else:
            i += 1
    return total
