"""CPU ORACLE (test infrastructure): queue-depth scale function.

Restates pkg/abstractions/taskqueue/autoscaler.go:53-79 (`taskQueueScaleFunc`). PINNED by the
reference's seven known answers, pkg/abstractions/taskqueue/autoscaler_test.go:34-123.
"""
from __future__ import annotations

from typing import Tuple


def task_queue_scale(queue_length: int, tasks_per_container: int, max_containers: int,
                     max_replicas: int) -> Tuple[int, bool]:
    """-> (DesiredContainers, ResultValid)."""
    if queue_length == 0:
        return 0, True
    if queue_length == -1:
        return 0, False
    desired = queue_length // tasks_per_container
    if queue_length % tasks_per_container > 0:
        desired += 1
    return int(min(float(min(max_containers, max_replicas)), float(desired))), True
