"""Regenerates tests/golden/hot_path_golden.json — run from the repo root:

    python tests/golden/make_golden.py

DERIVED vectors: produced by oracle/pyoracle (json.dumps / json.loads / zlib / base64 are the
reference's own stdlib calls; the Go encoding/json half is a restatement — PARITY UNPINNED,
see oracle/pyoracle/gojson.py). Re-validate against a Go 1.23 build of pkg/types/task.go when one
is available. The reference repo itself is not read by this script.
"""
import base64
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from beta9_b200 import synth  # noqa: E402
from oracle.pyoracle import loop  # noqa: E402
from tests.test_oracle_c_vs_py import HANDCRAFTED  # noqa: E402

HANDLERS = ["identity", "crc32", "vadd_f32", "json_sum"]


def main():
    groups = {
        "handcrafted": synth.from_payloads(HANDCRAFTED),
        "echo64": synth.strings_batch(48, 64, adversarial_frac=0.34),
        "identity256": synth.strings_batch(24, 256, adversarial_frac=0.25, seed=1),
        "crc_zipf": synth.crc_batch(24),
        "vadd": synth.vadd_batch(8),
        "json1k": synth.json_batch(8),
    }
    out = {"now_unix_ns": 1_789_970_992_573_161_412, "workspace": "ws-b200",
           "stub": "7f1c2d3e-4a5b-4c6d-8e9f-0a1b2c3d4e5f", "groups": {}}
    for name, b in groups.items():
        ids = [bytes(x) for x in b.task_ids]
        cases = []
        per_handler = {h: loop.run_task_loop(b.tasks(), ids, h, keep_wire=True) for h in HANDLERS}
        for i in range(b.n):
            c = {"task_id": ids[i].hex(), "payload": base64.b64encode(b.task(i)).decode()}
            w = per_handler["identity"][i].wire
            c["wire"] = base64.b64encode(w).decode() if w is not None else None
            c["results"] = {h: [per_handler[h][i].status,
                                None if per_handler[h][i].result is None
                                else base64.b64encode(per_handler[h][i].result).decode()] for h in HANDLERS}
            cases.append(c)
        out["groups"][name] = cases
    path = os.path.join(ROOT, "tests", "golden", "hot_path_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(path, sum(len(v) for v in out["groups"].values()), "cases")


if __name__ == "__main__":
    main()
