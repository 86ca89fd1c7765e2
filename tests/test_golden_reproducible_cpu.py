"""CPU (-m "not gpu"), container side: the committed fixtures under tests/golden ARE what the real reference produces.
Every generator script (each imports the reference from /root/reference, asserts oracle == reference and writes its .npz)
is re-run and must leave its fixture byte-identical. Skipped where the reference tree does not exist (the GPU box).
Not in the list: the full-size generator (minutes of CPU) and the legacy-fusion one, whose recorded gradient NORMS of the
reference's multi-threaded CPU backward move by ~1e-8 between runs (the tests that read them use tolerances)."""
import hashlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")

SCRIPTS = {"make_golden.py": ["w2c_small.npz"], "make_golden_cobevt.py": ["cobevt_small.npz"],
           "make_golden_v2xvit.py": ["v2xvit_small.npz"], "make_golden_labels.py": ["labels.npz"],
           "make_golden_postprocess.py": ["postprocess.npz"], "make_golden_lss.py": ["lss_small.npz"],
           "make_golden_bevencode.py": ["bevencode_small.npz"]}


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def test_generators_reproduce_the_committed_fixtures(tmp_path):
    gold = os.path.join(ROOT, "tests", "golden")
    before = {f: _sha(os.path.join(gold, f)) for files in SCRIPTS.values() for f in files}
    os.makedirs(tmp_path / "debug", exist_ok=True)           # the reference writes debug images relative to the cwd
    env = dict(os.environ, OMP_NUM_THREADS="2", MKL_NUM_THREADS="2")     # seven generators side by side
    procs = {s: subprocess.Popen([sys.executable, os.path.join(ROOT, "scripts", s)], cwd=str(tmp_path), env=env,
                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, stdin=subprocess.DEVNULL, text=True)
             for s in SCRIPTS}
    for s, p in procs.items():
        out, _ = p.communicate(timeout=900)
        assert p.returncode == 0, (s, out[-1500:])
    after = {f: _sha(os.path.join(gold, f)) for f in before}
    assert after == before, [f for f in before if after[f] != before[f]]
