"""HANDLER=ref_handlers:<name> for tests/golden/make_ref_runner_golden.py: the configs' user functions
(oracle/pyoracle/handlers.py) wrapped by the REFERENCE's own `_CallableWrapper`
(sdk/src/beta9/abstractions/taskqueue.py:204-295), as `@task_queue(...)` would leave them in a user's
module. `TaskQueue.__init__` is not run (it talks to the gateway); the runner only reads
`parent.retry_for` (runner/taskqueue.py:358)."""
import beta9.abstractions.taskqueue as _at
from oracle.pyoracle import handlers as _h


def _wrap(func):
    parent = _at.TaskQueue.__new__(_at.TaskQueue)
    parent.retry_for = []
    return _at._CallableWrapper(func, parent)


identity = _wrap(_h.identity)
crc32 = _wrap(_h.crc32)
vadd_f32 = _wrap(_h.vadd_f32)
json_sum = _wrap(_h.json_sum)
