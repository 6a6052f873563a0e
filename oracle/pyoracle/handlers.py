"""CPU ORACLE (test infrastructure): the user handlers of BASELINE.json's configs, written the
way a beta9 user would write them under `@task_queue` and invoked exactly as
`FunctionHandler.__call__` does (sdk/src/beta9/runner/common.py:297-305):
`handler(*args, **kwargs)`; any exception -> TaskStatus.ERROR (runner/taskqueue.py:352-361).

The parameters are positional-only so that the calling convention is part of the definition:
a payload carrying kwargs raises TypeError (-> ERROR) instead of binding by name.
"""
from __future__ import annotations

import base64
import zlib

import numpy as np


def identity(s, /):
    """configs[0] "echo" and configs[1] "identity": return the argument unchanged."""
    return s


echo = identity


def crc32(s, /):
    """configs[2]: IEEE 802.3 CRC-32 of the UTF-8 bytes (zlib.crc32)."""
    return zlib.crc32(s.encode())


def vadd_f32(x, /):
    """configs[3]: x = std-base64 of little-endian fp32 a||b; returns std-base64 of a+b (fp32)."""
    raw = base64.b64decode(x, validate=True)
    if len(raw) % 8:
        raise ValueError("payload must hold two fp32 vectors of equal length")
    v = np.frombuffer(raw, dtype="<f4")
    n = v.size // 2
    with np.errstate(all="ignore"):
        c = (v[:n] + v[n:]).astype("<f4")      # IEEE-754 binary32 add, round-to-nearest-even
    return base64.b64encode(c.tobytes()).decode("ascii")


def json_sum(obj, /):
    """configs[4]: parse+reduce — sum of obj["values"]."""
    return sum(obj["values"])


HANDLERS = {
    "identity": identity,
    "echo": echo,
    "crc32": crc32,
    "vadd_f32": vadd_f32,
    "json_sum": json_sum,
}
