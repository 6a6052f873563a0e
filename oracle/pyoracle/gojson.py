"""CPU ORACLE (test infrastructure, never the product path): Go `encoding/json` rules.

PARITY UNPINNED at this boundary: the reference holds no golden bytes for
`TaskMessage.Encode`/`Decode` and no Go toolchain exists in the build image, so
this file restates Go 1.23's *documented* `encoding/json` behaviour (go.mod:3 pins go 1.23).
Call sites it stands in for:

  * `json.Unmarshal(in.Payload, &payload)`   pkg/abstractions/taskqueue/taskqueue.go:213-214
  * `json.Marshal(tm)` in `TaskMessage.Encode` pkg/types/task.go:79-90
  * `json.Unmarshal` in `TaskMessage.Decode`  pkg/types/task.go:93-108

Rules restated (Go 1.23 encoding/json):
  decode: strict RFC 8259 grammar validated over the WHOLE input first (`checkValid`); numbers
    -> float64 (`strconv.ParseFloat`, overflow is an error, underflow is not); strings: escapes
    `\\" \\\\ \\/ \\b \\f \\n \\r \\t \\uXXXX`, surrogate pairs joined, lone surrogates -> U+FFFD,
    invalid UTF-8 -> U+FFFD per offending byte, raw control bytes < 0x20 are a syntax error;
    objects into `map[string]any`: later duplicate keys win.
  encode: struct fields in declaration order; map keys sorted bytewise; no whitespace;
    `<`,`>`,`&` -> `\\u003c`,`\\u003e`,`\\u0026`; U+2028/2029 escaped; `\\b \\f \\n \\r \\t` short
    forms (since Go 1.22), other controls `\\u00XX`; invalid UTF-8 -> `\\ufffd`; DEL and all other
    non-ASCII raw; float64 in %f form with shortest digits unless |x| < 1e-6 or >= 1e21, then
    %e with the exponent cleaned (`e-07` -> `e-7`); nil map -> `null`; time.Time -> RFC3339Nano.
"""
from __future__ import annotations

from decimal import Decimal
from typing import Any, Dict, List, Optional, Tuple

__all__ = [
    "GoJSONError", "go_unmarshal", "go_unmarshal_task_payload", "go_marshal",
    "go_format_float64", "go_quote", "go_time_rfc3339nano", "go_fold_equal",
]


class GoJSONError(ValueError):
    """`json.Unmarshal` returned a non-nil error (SyntaxError or UnmarshalTypeError)."""


_WS = b" \t\r\n"
_HEX = b"0123456789abcdefABCDEF"


class _Parser:
    """Single-pass strict JSON parser producing Go `interface{}` values.

    Go validates the whole document before decoding; for a recursive-descent parser that accepts
    exactly the RFC 8259 grammar, the set of accepted inputs is identical, so validation and
    decoding are fused here.
    """

    MAX_DEPTH = 10000  # Go: 10000 open containers are fine, the 10001st is "exceeded max depth" (scanner.go pushParseState)

    def __init__(self, data: bytes):
        self.d = data
        self.i = 0
        # number literals ParseFloat refused (ErrRange), in source order. decode.go keeps the FIRST conversion error
        # (saveError) and goes on; Unmarshal returns it at the end even if a later duplicate key overwrote the value.
        self.overflows: List[str] = []
        self.n = len(data)

    def err(self, msg: str) -> GoJSONError:
        return GoJSONError(f"invalid JSON at byte {self.i}: {msg}")

    def skip_ws(self) -> None:
        d, n = self.d, self.n
        i = self.i
        while i < n and d[i] in _WS:
            i += 1
        self.i = i

    def parse_document(self) -> Any:
        self.skip_ws()
        if self.i >= self.n:
            raise self.err("unexpected end of JSON input")
        v = self.parse_value(0)
        self.skip_ws()
        if self.i != self.n:
            raise self.err("invalid character after top-level value")
        return v

    def parse_value(self, depth: int) -> Any:
        if self.i >= self.n:
            raise self.err("unexpected end of JSON input")
        c = self.d[self.i]
        if c == 0x7B or c == 0x5B:  # { [   (`depth` containers are open around this one)
            if depth + 1 > self.MAX_DEPTH:
                raise self.err("exceeded max depth")
            return self.parse_object(depth) if c == 0x7B else self.parse_array(depth)
        if c == 0x22:
            return self.parse_string()
        if c == 0x2D or 0x30 <= c <= 0x39:
            return self.parse_number()
        if self.d.startswith(b"true", self.i):
            self.i += 4
            return True
        if self.d.startswith(b"false", self.i):
            self.i += 5
            return False
        if self.d.startswith(b"null", self.i):
            self.i += 4
            return None
        raise self.err("invalid character looking for beginning of value")

    def parse_object(self, depth: int) -> Dict[str, Any]:
        self.i += 1
        out: Dict[str, Any] = {}
        self.skip_ws()
        if self.i < self.n and self.d[self.i] == 0x7D:
            self.i += 1
            return out
        while True:
            self.skip_ws()
            if self.i >= self.n or self.d[self.i] != 0x22:
                raise self.err("invalid character looking for beginning of object key string")
            k = self.parse_string()
            self.skip_ws()
            if self.i >= self.n or self.d[self.i] != 0x3A:
                raise self.err("invalid character after object key")
            self.i += 1
            self.skip_ws()
            out[k] = self.parse_value(depth + 1)  # later duplicates win (decode.go objectInterface)
            self.skip_ws()
            if self.i >= self.n:
                raise self.err("unexpected end of JSON input")
            c = self.d[self.i]
            self.i += 1
            if c == 0x2C:
                continue
            if c == 0x7D:
                return out
            raise self.err("invalid character after object key:value pair")

    def parse_array(self, depth: int) -> List[Any]:
        self.i += 1
        out: List[Any] = []
        self.skip_ws()
        if self.i < self.n and self.d[self.i] == 0x5D:
            self.i += 1
            return out
        while True:
            self.skip_ws()
            out.append(self.parse_value(depth + 1))
            self.skip_ws()
            if self.i >= self.n:
                raise self.err("unexpected end of JSON input")
            c = self.d[self.i]
            self.i += 1
            if c == 0x2C:
                continue
            if c == 0x5D:
                return out
            raise self.err("invalid character after array element")

    def parse_number(self) -> float:
        d, n = self.d, self.n
        s = self.i
        i = s
        if i < n and d[i] == 0x2D:
            i += 1
        if i >= n:
            self.i = i
            raise self.err("unexpected end of JSON input")
        if d[i] == 0x30:
            i += 1
        elif 0x31 <= d[i] <= 0x39:
            while i < n and 0x30 <= d[i] <= 0x39:
                i += 1
        else:
            self.i = i
            raise self.err("invalid character in numeric literal")
        if i < n and d[i] == 0x2E:
            i += 1
            if i >= n or not (0x30 <= d[i] <= 0x39):
                self.i = i
                raise self.err("invalid character after decimal point in numeric literal")
            while i < n and 0x30 <= d[i] <= 0x39:
                i += 1
        if i < n and d[i] in b"eE":
            i += 1
            if i < n and d[i] in b"+-":
                i += 1
            if i >= n or not (0x30 <= d[i] <= 0x39):
                self.i = i
                raise self.err("invalid character in exponent of numeric literal")
            while i < n and 0x30 <= d[i] <= 0x39:
                i += 1
        self.i = i
        text = d[s:i].decode("ascii")
        f = float(text)  # correctly rounded, like strconv.ParseFloat(s, 64)
        if f in (float("inf"), float("-inf")):
            # decode.go convertNumber: ParseFloat ErrRange -> UnmarshalTypeError{"number " + s}.
            # Only raised when the literal is actually stored (values under ignored struct keys
            # are skipped without conversion), so hand back a marker and let the caller decide.
            self.overflows.append(text)
            return _Overflow(text)
        return f

    def _u4(self, at: int) -> int:
        """decode.go getu4: value of `\\uXXXX` starting at `at`, or -1."""
        d = self.d
        if at + 6 > self.n or d[at] != 0x5C or d[at + 1] != 0x75:
            return -1
        v = 0
        for c in d[at + 2:at + 6]:
            if c not in _HEX:
                return -1
            v = v * 16 + int(chr(c), 16)
        return v

    def parse_string(self) -> str:
        d, n = self.d, self.n
        i = self.i + 1
        out: List[str] = []
        while True:
            if i >= n:
                self.i = i
                raise self.err("unexpected end of JSON input")
            c = d[i]
            if c == 0x22:
                self.i = i + 1
                return "".join(out)
            if c < 0x20:
                self.i = i
                raise self.err("invalid character in string literal")
            if c == 0x5C:
                if i + 1 >= n:
                    self.i = i
                    raise self.err("unexpected end of JSON input")
                e = d[i + 1]
                simple = {0x22: '"', 0x5C: "\\", 0x2F: "/", 0x62: "\b", 0x66: "\f",
                          0x6E: "\n", 0x72: "\r", 0x74: "\t"}
                if e in simple:
                    out.append(simple[e])
                    i += 2
                    continue
                if e != 0x75:
                    self.i = i
                    raise self.err("invalid character in string escape code")
                rr = self._u4(i)
                if rr < 0:
                    self.i = i
                    raise self.err("invalid character in \\u hexadecimal character escape")
                i += 6
                if 0xD800 <= rr <= 0xDFFF:
                    rr1 = self._u4(i)
                    if 0xD800 <= rr <= 0xDBFF and 0xDC00 <= rr1 <= 0xDFFF:
                        # valid pair (utf16.DecodeRune)
                        out.append(chr(0x10000 + ((rr - 0xD800) << 10) + (rr1 - 0xDC00)))
                        # the second escape must itself be syntactically valid, which _u4 ensured
                        i += 6
                        continue
                    rr = 0xFFFD  # lone surrogate: only the first escape is consumed
                out.append(chr(rr))
                continue
            if c < 0x80:
                out.append(chr(c))
                i += 1
                continue
            cp, size = _decode_rune(d, i, n)
            out.append(chr(cp))
            i += size


class _Overflow:
    """A syntactically valid number literal that ParseFloat rejects with ErrRange."""

    def __init__(self, text: str):
        self.text = text


def _check_overflow(v: Any) -> None:
    if isinstance(v, _Overflow):
        raise GoJSONError(f"json: cannot unmarshal number {v.text} into Go value of type float64")
    if isinstance(v, list):
        for x in v:
            _check_overflow(x)
    elif isinstance(v, dict):
        for x in v.values():
            _check_overflow(x)


def _decode_rune(d: bytes, i: int, n: int) -> Tuple[int, int]:
    """Go utf8.DecodeRune: (rune, width); invalid encodings give (U+FFFD, 1)."""
    c0 = d[i]
    if c0 < 0x80:
        return c0, 1
    if 0xC2 <= c0 <= 0xDF:
        if i + 1 < n and 0x80 <= d[i + 1] <= 0xBF:
            return ((c0 & 0x1F) << 6) | (d[i + 1] & 0x3F), 2
        return 0xFFFD, 1
    if 0xE0 <= c0 <= 0xEF:
        lo, hi = 0x80, 0xBF
        if c0 == 0xE0:
            lo = 0xA0
        elif c0 == 0xED:
            hi = 0x9F
        if i + 2 < n and lo <= d[i + 1] <= hi and 0x80 <= d[i + 2] <= 0xBF:
            return ((c0 & 0x0F) << 12) | ((d[i + 1] & 0x3F) << 6) | (d[i + 2] & 0x3F), 3
        return 0xFFFD, 1
    if 0xF0 <= c0 <= 0xF4:
        lo, hi = 0x80, 0xBF
        if c0 == 0xF0:
            lo = 0x90
        elif c0 == 0xF4:
            hi = 0x8F
        if (i + 3 < n and lo <= d[i + 1] <= hi and 0x80 <= d[i + 2] <= 0xBF
                and 0x80 <= d[i + 3] <= 0xBF):
            return (((c0 & 0x07) << 18) | ((d[i + 1] & 0x3F) << 12)
                    | ((d[i + 2] & 0x3F) << 6) | (d[i + 3] & 0x3F)), 4
        return 0xFFFD, 1
    return 0xFFFD, 1


def go_unmarshal(data: bytes) -> Any:
    """`json.Unmarshal(data, &v)` with `v interface{}`: dict / list / str / float / bool / None."""
    p = _Parser(bytes(data))
    v = p.parse_document()
    if p.overflows:                 # anywhere in the document, overwritten by a duplicate key or not (decode.go saveError)
        raise GoJSONError(f"json: cannot unmarshal number {p.overflows[0]} into Go value of type float64")
    _check_overflow(v)
    return v


def go_fold_equal(key: str, field: str) -> bool:
    """encoding/json field-name matching: exact, else case-insensitive under Go's `foldName`
    (ASCII case folding plus the two non-ASCII runes that fold into ASCII letters:
    U+212A KELVIN SIGN -> k, U+017F LONG S -> s)."""
    def fold(s: str) -> str:
        return s.replace("K", "k").replace("ſ", "s").upper()
    # str.upper() maps some non-ASCII runes (e.g. U+00DF) to ASCII pairs; Go's simple folding does
    # not, so restrict the comparison to keys that are ASCII after the two special cases.
    k = key.replace("K", "k").replace("ſ", "s")
    if not k.isascii():
        return False
    return k.upper() == field.upper()


def go_unmarshal_task_payload(data: bytes) -> Tuple[Optional[List[Any]], Optional[Dict[str, Any]]]:
    """`json.Unmarshal(in.Payload, &types.TaskPayload{})` (taskqueue.go:213-214; struct
    pkg/types/task.go:13-16).  Returns (Args, Kwargs) where None is Go's nil.

    Raises GoJSONError when Go's Unmarshal returns an error (=> `TaskQueuePutResponse{Ok:false}`).
    """
    doc = _Parser(bytes(data)).parse_document()
    if doc is None:        # top-level null: Unmarshal is a no-op
        return None, None
    if not isinstance(doc, dict):
        raise GoJSONError("json: cannot unmarshal non-object into Go value of type types.TaskPayload")
    # Duplicate struct keys need source order, which `dict` lost; re-scan the top level.
    args: Optional[List[Any]] = None
    kwargs: Optional[Dict[str, Any]] = None
    type_err: Optional[str] = None
    for key, val, overflowed in _top_level_pairs(bytes(data)):
        # exact match wins over folded match, but both names are distinct under folding, so a key
        # can match at most one field
        is_args = key == "args" or (key != "kwargs" and go_fold_equal(key, "args"))
        is_kwargs = not is_args and (key == "kwargs" or go_fold_equal(key, "kwargs"))
        if (is_args or is_kwargs) and overflowed and isinstance(val, (list, dict)):
            # a number inside a DECODED value failed to convert: the error is saved and returned at the end, whatever a
            # later duplicate of the key does (decode.go literalInterface -> saveError). Values under unknown keys, and
            # values of the wrong kind for the field (skipped after the type error), are never converted.
            type_err = type_err or f"json: cannot unmarshal number {overflowed} into Go value of type float64"
        if is_args:
            if val is None:
                args = None
            elif isinstance(val, list):
                args = val            # slice is re-filled from scratch
            else:
                type_err = type_err or "json: cannot unmarshal into Go struct field TaskPayload.args of type []interface {}"
        elif is_kwargs:
            if val is None:
                kwargs = None
            elif isinstance(val, dict):
                if kwargs is None:
                    kwargs = {}
                kwargs.update(val)    # decoding into a non-nil map keeps existing entries
            else:
                type_err = type_err or "json: cannot unmarshal into Go struct field TaskPayload.kwargs of type map[string]interface {}"
        # unknown keys are ignored
    if type_err:
        raise GoJSONError(type_err)
    _check_overflow(args)
    _check_overflow(kwargs)
    return args, kwargs


def _top_level_pairs(data: bytes):
    """(key, value, first overflowing number literal inside the value or None) of the top-level object, in source
    order, duplicates kept."""
    p = _Parser(data)
    p.skip_ws()
    assert p.d[p.i] == 0x7B
    p.i += 1
    p.skip_ws()
    if p.d[p.i] == 0x7D:
        return
    while True:
        p.skip_ws()
        k = p.parse_string()
        p.skip_ws()
        p.i += 1  # ':'
        p.skip_ws()
        seen = len(p.overflows)
        v = p.parse_value(1)
        yield k, v, (p.overflows[seen] if len(p.overflows) > seen else None)
        p.skip_ws()
        c = p.d[p.i]
        p.i += 1
        if c == 0x7D:
            return


# ----------------------------------------------------------------------------- encode

def _shortest_digits(f: float) -> Tuple[str, int]:
    """Shortest round-trip decimal digits of |f| and the decimal-point position dp such that
    |f| = 0.d1d2d3... x 10^dp (strconv's `decimal` convention). CPython's repr and Go's
    strconv both emit the shortest digit string that round-trips, ties to the closest."""
    t = Decimal(repr(abs(f))).as_tuple()
    digits = "".join(map(str, t.digits)).lstrip("0")
    exp = t.exponent
    stripped = digits.rstrip("0")
    exp += len(digits) - len(stripped)
    digits = stripped
    if not digits:
        return "0", 1
    return digits, len(digits) + exp


def go_format_float64(f: float) -> str:
    """encode.go floatEncoder (bits=64)."""
    if f != f or f in (float("inf"), float("-inf")):
        raise GoJSONError("json: unsupported value: NaN/Inf")
    neg = (f < 0) or (f == 0 and str(f).startswith("-"))
    a = abs(f)
    digits, dp = _shortest_digits(f)
    use_e = a != 0 and (a < 1e-6 or a >= 1e21)
    if a == 0:
        s = "0"
    elif not use_e:
        nd = len(digits)
        if dp > 0:
            ip = digits[:dp] + "0" * max(0, dp - nd)
            fp = digits[dp:] if nd > dp else ""
        else:
            ip = "0"
            fp = "0" * (-dp) + digits
        s = ip + ("." + fp if fp else "")
    else:
        e = dp - 1
        mant = digits[0] + ("." + digits[1:] if len(digits) > 1 else "")
        ae = abs(e)
        # strconv %e writes at least two exponent digits; encode.go then turns e-09 into e-9
        es = f"{ae:02d}"
        s = mant + "e" + ("-" if e < 0 else "+") + es
        if len(s) >= 4 and s[-4] == "e" and s[-3] == "-" and s[-2] == "0":
            s = s[:-2] + s[-1]
    return ("-" if neg else "") + s


_GO_SHORT = {0x08: "\\b", 0x0C: "\\f", 0x0A: "\\n", 0x0D: "\\r", 0x09: "\\t",
             0x22: '\\"', 0x5C: "\\\\"}


def go_quote(s: str, escape_html: bool = True) -> str:
    """encode.go appendString for a Go string holding the UTF-8 of `s`.
    Lone surrogates cannot be encoded as UTF-8; Go strings reaching here never contain them
    (the decoder replaced them), but if `s` does they are treated as invalid UTF-8 -> \\ufffd."""
    out = ['"']
    for ch in s:
        c = ord(ch)
        if c < 0x80:
            if c in _GO_SHORT:
                out.append(_GO_SHORT[c])
            elif c < 0x20:
                out.append("\\u00%02x" % c)
            elif escape_html and ch in "<>&":
                out.append("\\u00%02x" % c)
            else:
                out.append(ch)
        elif c in (0x2028, 0x2029):
            out.append("\\u%04x" % c)
        elif 0xD800 <= c <= 0xDFFF:
            out.append("\\ufffd")
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def go_marshal(v: Any) -> str:
    """`json.Marshal` of a value built from interface{} kinds (nil/bool/float64/string/
    []interface{}/map[string]interface{}) plus Python ints for Go integer fields."""
    if v is None:
        return "null"
    if v is True:
        return "true"
    if v is False:
        return "false"
    if isinstance(v, int):
        return str(v)
    if isinstance(v, float):
        return go_format_float64(v)
    if isinstance(v, str):
        return go_quote(v)
    if isinstance(v, (list, tuple)):
        return "[" + ",".join(go_marshal(x) for x in v) + "]"
    if isinstance(v, dict):
        # encode.go mapEncoder: keys sorted by strings.Compare of the key string = bytewise UTF-8
        items = sorted(v.items(), key=lambda kv: kv[0].encode("utf-8", "surrogatepass"))
        return "{" + ",".join(go_quote(k) + ":" + go_marshal(x) for k, x in items) + "}"
    raise TypeError(f"go_marshal: unsupported type {type(v)!r}")


_DAYS_BEFORE = [0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334]


def _civil_from_days(z: int) -> Tuple[int, int, int]:
    """days since 1970-01-01 -> proleptic Gregorian (y, m, d)."""
    z += 719468
    era = (z if z >= 0 else z - 146096) // 146097
    doe = z - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
    y = yoe + era * 400
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    mp = (5 * doy + 2) // 153
    d = doy - (153 * mp + 2) // 5 + 1
    m = mp + 3 if mp < 10 else mp - 9
    return (y + 1 if m <= 2 else y), m, d


def go_time_rfc3339nano(unix_ns: int, offset_min: int = 0) -> str:
    """time.Time.MarshalJSON body: RFC3339Nano ("2006-01-02T15:04:05.999999999Z07:00") of the
    instant `unix_ns` shown at UTC offset `offset_min` (0 -> "Z"). Trailing zeros of the
    fraction are dropped; a zero fraction is omitted."""
    local_ns = unix_ns + offset_min * 60 * 10**9
    secs, ns = divmod(local_ns, 10**9)
    days, sod = divmod(secs, 86400)
    y, m, d = _civil_from_days(days)
    if not (0 <= y <= 9999):
        raise GoJSONError("Time.MarshalJSON: year outside of range [0,9999]")
    hh, rem = divmod(sod, 3600)
    mm, ss = divmod(rem, 60)
    s = f"{y:04d}-{m:02d}-{d:02d}T{hh:02d}:{mm:02d}:{ss:02d}"
    if ns:
        s += "." + f"{ns:09d}".rstrip("0")
    if offset_min == 0:
        s += "Z"
    else:
        sign = "+" if offset_min > 0 else "-"
        a = abs(offset_min)
        s += f"{sign}{a // 60:02d}:{a % 60:02d}"
    return s


GO_ZERO_TIME_UNIX_NS = -62135596800 * 10**9  # time.Time{} = 0001-01-01T00:00:00Z
