"""CPU ORACLE (test infrastructure): HTTP body + query -> TaskPayload.

Restates pkg/task/serialize.go:16-101 (`SerializeHttpPayload`, `parseRequestPayload`,
`parseRequestArgs`, `tryParseNumeric`). jsoniter's `ConfigCompatibleWithStandardLibrary`
(serialize.go:14) decodes like encoding/json for the value kinds involved, so the body goes
through gojson.go_unmarshal. PINNED by the reference's own 15 known answers,
pkg/task/serialize_test.go:39-173 (tests/test_oracle_reference_answers.py).
"""
from __future__ import annotations

from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple

from .gojson import GoJSONError, go_unmarshal


class InvalidRequestPayload(ValueError):
    """serialize.go:24 `errors.New("invalid request payload")`."""


def try_parse_numeric(s: str) -> Any:
    """serialize.go:95-100 — strconv.ParseFloat(s, 64) or the string itself.
    ParseFloat accepts decimal/hex floats, "inf"/"infinity"/"nan" (any case) and '_' only with a
    base prefix; Python's float() additionally accepts surrounding whitespace and '_' between
    digits, which ParseFloat rejects."""
    t = s
    if not t or t != t.strip() or "_" in t:
        return s
    try:
        if t.lower().lstrip("+-").startswith("0x"):
            return float.fromhex(t)
        return float(t)
    except (ValueError, OverflowError):
        return s


def serialize_http_payload(body: bytes, query: Optional[Mapping[str, Sequence[str]]] = None
                           ) -> Tuple[Optional[List[Any]], Dict[str, Any]]:
    """Returns (Args, Kwargs); Args None = Go nil."""
    payload: Dict[str, Any] = {}
    if body.strip(b" \t\r\n"):
        try:
            doc = go_unmarshal(body)
        except GoJSONError as e:
            raise InvalidRequestPayload("invalid request payload") from e
        if doc is not None and not isinstance(doc, dict):
            raise InvalidRequestPayload("invalid request payload")
        payload = doc or {}
    # an empty body makes decoder.Decode return io.EOF, which serialize.go:22-25 tolerates
    args: Optional[List[Any]] = None
    kwargs: Dict[str, Any] = {}
    if payload:
        a = payload.get("args")
        if isinstance(a, list):                 # serialize.go:48-51
            args = a
            del payload["args"]
        k = payload.get("kwargs")
        if isinstance(k, dict):                 # :54-57
            kwargs = k
            del payload["kwargs"]
        elif payload:                           # :57-60
            kwargs = payload
    if query:
        for key, values in query.items():       # :63-93
            if len(values) == 1:
                kwargs[key] = try_parse_numeric(values[0])
                continue
            conv = [try_parse_numeric(v) for v in values]
            if all(isinstance(c, float) for c in conv):
                kwargs[key] = conv
            else:
                kwargs[key] = list(values)
    return args, kwargs
