"""End-to-end: the C++ sample programs over the drop-in `snn` facade print byte-for-byte the JSON
the reference's own sample programs print (md5 of stdout)."""
import hashlib
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("name", ["brunel", "brunel+", "vogels", "ping_pong", "external_input"])
def test_sample_stdout_md5(golden, name):
    exe = ROOT / "samples" / "build" / name
    if not exe.exists():
        subprocess.run(["make", "-s", "-C", str(ROOT / "samples")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, check=True, timeout=600).stdout
    assert hashlib.md5(out).hexdigest() == golden["sample_stdout_md5"][name]


def test_sssp_distances(golden):
    """samples/sssp over the facade: DeliverFromTo synapses (deliver() reads the source neuron), per-synapse and
    per-population init hooks; the distances the compiled reference computes, and the sample's own asserts."""
    exe = ROOT / "samples" / "build" / "sssp"
    if not exe.exists():
        subprocess.run(["make", "-s", "-C", str(ROOT / "samples")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, check=True, timeout=600).stdout.split()
    assert [int(x) for x in out] == golden["sssp_distances"]


def test_facade_spec_program():
    """tests/cpp/facade_spec.cu: the reference's own unit tests of delivery and per-population update
    (test/detail/synapse_population.cpp:28-81, test/detail/neuron_population.cpp:146-170) through the public API."""
    exe = ROOT / "tests" / "cpp" / "build" / "facade_spec"
    if not exe.exists():
        subprocess.run(["make", "-s", "-C", str(ROOT / "tests" / "cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, timeout=600)
    assert r.returncode == 0, r.stdout.decode() + r.stderr.decode()


def test_multi_rank_spec_program():
    """tests/cpp/multi_rank_spec.cu: per-synapse init hooks (over fixed_probability and adj_list connections) and a host-fed
    population on two ranks reproduce the one-rank run: weights, spikes and neuron state."""
    exe = ROOT / "tests" / "cpp" / "build" / "multi_rank_spec"
    if not exe.exists():
        subprocess.run(["make", "-s", "-C", str(ROOT / "tests" / "cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, timeout=600)
    assert r.returncode == 0, r.stdout.decode() + r.stderr.decode()
