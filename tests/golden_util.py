import base64
import json
import os

import numpy as np

from beta9_b200 import synth

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hot_path_golden.json")
STATUS_CODE = {"COMPLETE": 0, "ERROR": 1, "REJECTED": 3}


def load():
    with open(PATH) as f:
        return json.load(f)


def group_batch(cases):
    payloads = [base64.b64decode(c["payload"]) for c in cases]
    b = synth.from_payloads(payloads)
    b.task_ids = np.frombuffer(b"".join(bytes.fromhex(c["task_id"]) for c in cases), np.uint8).reshape(-1, 16).copy()
    return b


def expected(case, handler):
    st, res = case["results"][handler]
    return STATUS_CODE[st], (None if res is None else base64.b64decode(res))
