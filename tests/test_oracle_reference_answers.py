"""The oracle against every known answer the reference's own tests hold for this path
(SURVEY.md §8c): these are the PINNED parts of the oracle."""
import pytest

from oracle import coracle
from oracle.pyoracle.autoscaler import task_queue_scale
from oracle.pyoracle.httpserialize import InvalidRequestPayload, serialize_http_payload
from oracle.pyoracle.ringbuffer import RingBuffer

# pkg/task/serialize_test.go:39-173 — (name, body, query, want_args, want_kwargs, want_err)
SERIALIZE_CASES = [
    ("complete empty body", "", None, None, {}, False),
    ("empty json object", "{}", None, None, {}, False),
    ("kwargs only", '{"mykwarg": 1, "mykwarg2": 2}', None, None, {"mykwarg": 1.0, "mykwarg2": 2.0}, False),
    ("args and kwargs", '{"args": [1, 2, 3], "mykwarg": "value"}', None, [1.0, 2.0, 3.0], {"mykwarg": "value"}, False),
    ("args only", '{"args": [1, 2, 3]}', None, [1.0, 2.0, 3.0], {}, False),
    ("explicit kwargs and args", '{"kwargs": {"test": 1}, "args": [1, 2, 3]}', None, [1.0, 2.0, 3.0], {"test": 1.0}, False),
    ("malformed json", '{"args": [1, 2, 3}', None, None, None, True),
    ("nested kwargs", '{"kwargs": {"nestedList": [1, 2, 3], "nestedMap": {"key": "value"}}}', None, None,
     {"nestedList": [1.0, 2.0, 3.0], "nestedMap": {"key": "value"}}, False),
    ("list of strings", "{}", {"listOfStrings": ["a", "b", "c"]}, None, {"listOfStrings": ["a", "b", "c"]}, False),
    ("single number", "{}", {"sleep": ["100"]}, None, {"sleep": 100.0}, False),
    ("single number no body", "", {"sleep": ["100"]}, None, {"sleep": 100.0}, False),
    ("list of ints", "", {"sleep": ["100", "200", "300"]}, None, {"sleep": [100.0, 200.0, 300.0]}, False),
    ("list of floats", "", {"sleep": ["100.1", "200.2", "300.3"]}, None, {"sleep": [100.1, 200.2, 300.3]}, False),
    ("list of mixed ints and floats", "", {"sleep": ["100", "200.2", "300"]}, None, {"sleep": [100.0, 200.2, 300.0]}, False),
    ("mix of strings and numbers", "", {"sleep": ["Today", "200.2", "300"]}, None, {"sleep": ["Today", "200.2", "300"]}, False),
]


@pytest.mark.parametrize("name,body,query,want_args,want_kwargs,want_err", SERIALIZE_CASES,
                         ids=[c[0] for c in SERIALIZE_CASES])
def test_serialize_http_payload(name, body, query, want_args, want_kwargs, want_err):
    if want_err:
        with pytest.raises(InvalidRequestPayload):
            serialize_http_payload(body.encode(), query)
        return
    args, kwargs = serialize_http_payload(body.encode(), query)
    assert args == want_args
    assert kwargs == want_kwargs
    # reflect.DeepEqual distinguishes float64 from string: check the leaf types too
    for k, v in want_kwargs.items():
        assert type(kwargs[k]) is type(v)


# pkg/abstractions/taskqueue/autoscaler_test.go:34-123 — (q, tpc, max_containers, max_replicas, desired, valid)
AUTOSCALER_CASES = [
    (10, 1, 1, 10, 1, True),
    (0, 1, 1, 10, 0, True),
    (-1, 1, 1, 10, 0, False),
    (3, 1, 3, 10, 3, True),
    (4, 1, 3, 10, 3, True),
    (11, 5, 5, 10, 3, True),
    (10, 5, 5, 10, 2, True),
]


@pytest.mark.parametrize("q,tpc,mc,mr,desired,valid", AUTOSCALER_CASES)
def test_task_queue_scale_func(q, tpc, mc, mr, desired, valid):
    assert task_queue_scale(q, tpc, mc, mr) == (desired, valid)
    assert coracle.task_queue_scale(q, tpc, mc, mr) == (desired, valid)


# pkg/abstractions/common/ring_buffer_test.go:7-169
def test_ring_push_and_pop():
    rb = RingBuffer(3)
    for v in (1, 2, 3):
        rb.push(v)
    assert len(rb) == 3
    assert [rb.pop() for _ in range(3)] == [(1, True), (2, True), (3, True)]
    assert rb.pop()[1] is False


def test_ring_priority_push():
    rb = RingBuffer(4)
    for v in (1, 2, 3):
        rb.push(v)
    rb.push(0, True)
    assert [rb.pop()[0] for _ in range(4)] == [0, 1, 2, 3]


def test_ring_full_buffer():
    rb = RingBuffer(3)
    for v in (1, 2, 3, 4):
        rb.push(v)
    assert len(rb) == 3
    assert [rb.pop()[0] for _ in range(3)] == [2, 3, 4]
    assert rb.pop()[1] is False


def test_ring_priority_push_on_full():
    rb = RingBuffer(3)
    for v in (1, 2, 3):
        rb.push(v)
    rb.push(0, True)
    assert len(rb) == 3
    assert [rb.pop()[0] for _ in range(3)] == [0, 2, 3]


def test_ring_overwrite_stats():
    rb = RingBuffer(2)
    assert rb.capacity() == 2
    assert rb.push(1) is False
    assert rb.push(2) is False
    assert rb.push(3) is True
    assert rb.overwrites == 1
    assert rb.push(0, True) is True
    assert rb.overwrites == 2
