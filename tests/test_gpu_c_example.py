"""examples/c_api_demo.c built with gcc and run on the GPU: the C ABI from plain C, without Python in between."""
import os
import shutil
import subprocess

import pytest

from beta9_b200 import build as B

pytestmark = pytest.mark.gpu


def test_c_example_prints_the_three_records(tmp_path):
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler on this box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = B.build()
    exe = str(tmp_path / "c_api_demo")
    r = subprocess.run([cc, "-std=c99", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "c_api_demo.c"),
                        "-L" + os.path.dirname(so), "-lb9gpu", "-Wl,-rpath," + os.path.dirname(so), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "pending: 3"
    assert lines[1] == 'task 01 status 0 result "hello"'
    assert lines[2] == 'task 02 status 0 result "caf\\u00e9 \\"x\\""'
    assert lines[3] == 'task 03 status 0 result "from an HTTP body"'
