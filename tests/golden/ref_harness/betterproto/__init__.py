"""Minimal stand-in for python-betterproto (third-party, absent from this image): just enough of its
dataclass field helpers for the reference SDK's GENERATED message classes to import and hold values.
No wire encoding: the harness never serialises protobuf, it only reads the fields the reference's
own runner code sets."""
import dataclasses, enum
PLACEHOLDER = None
class Message:
    def __post_init__(self): pass
    def to_dict(self, *a, **k): return dataclasses.asdict(self)
class Enum(enum.IntEnum):
    @classmethod
    def from_string(cls, name): return cls[name]
def _f(default=None, factory=None):
    return dataclasses.field(default_factory=factory) if factory else dataclasses.field(default=default)
def string_field(n, **k): return _f("")
def bytes_field(n, **k): return _f(b"")
def bool_field(n, **k): return _f(False)
def int32_field(n, **k): return _f(0)
def int64_field(n, **k): return _f(0)
def uint32_field(n, **k): return _f(0)
def uint64_field(n, **k): return _f(0)
def float_field(n, **k): return _f(0.0)
def double_field(n, **k): return _f(0.0)
def enum_field(n, **k): return _f(0)
def message_field(n, **k): return _f(None)
def map_field(n, *a, **k): return _f(factory=dict)
def which_one_of(msg, group): return ("", None)
TYPE_STRING = "string"; TYPE_BYTES = "bytes"; TYPE_MESSAGE = "message"; TYPE_INT32="int32"; TYPE_INT64="int64"; TYPE_UINT32="uint32"; TYPE_UINT64="uint64"; TYPE_BOOL="bool"; TYPE_ENUM="enum"; TYPE_FLOAT="float"; TYPE_DOUBLE="double"
class Casing: SNAKE = 1; CAMEL = 0
