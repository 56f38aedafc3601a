"""End-to-end: the C++ sample programs over the drop-in `snn` facade print byte-for-byte the JSON
the reference's own sample programs print (md5 of stdout)."""
import hashlib
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("name", ["brunel", "brunel+", "vogels"])
def test_sample_stdout_md5(golden, name):
    exe = ROOT / "samples" / "build" / name
    if not exe.exists():
        subprocess.run(["make", "-s", "-C", str(ROOT / "samples")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, check=True, timeout=600).stdout
    assert hashlib.md5(out).hexdigest() == golden["sample_stdout_md5"][name]
