"""CPU (-m "not gpu"), container side: the committed fixtures under tests/golden ARE what the real reference produces.
Every generator script (each imports the reference from /root/reference, asserts oracle == reference and writes its .npz)
is re-run and must reproduce its fixture (byte-identical here; on another CPU model the same arrays to 1e-5). Skipped where the reference tree does not exist (the GPU box).
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


def _same_arrays(a_path, b_path, atol=1e-5):
    """fallback when the bytes differ (another CPU model can move the reference's fp32 results by an ulp): same keys, dtypes,
    shapes; integers / strings equal, floats within atol"""
    import numpy as np
    a, b = np.load(a_path), np.load(b_path)
    if sorted(a.files) != sorted(b.files):
        return "keys differ"
    for k in a.files:
        if a[k].dtype != b[k].dtype or a[k].shape != b[k].shape:
            return k + ": dtype / shape"
        if a[k].dtype.kind == "f":
            if a[k].size and float(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)).max()) > atol:
                return k + ": values"
        elif not np.array_equal(a[k], b[k]):
            return k + ": values"
    return None


def test_generators_reproduce_the_committed_fixtures(tmp_path):
    import shutil

    gold = os.path.join(ROOT, "tests", "golden")
    names = [f for files in SCRIPTS.values() for f in files]
    before = {f: _sha(os.path.join(gold, f)) for f in names}
    keep = tmp_path / "committed"
    os.makedirs(keep)
    for f in names:
        shutil.copy2(os.path.join(gold, f), keep / f)
    os.makedirs(tmp_path / "debug", exist_ok=True)           # the reference writes debug images relative to the cwd
    env = dict(os.environ, OMP_NUM_THREADS="2", MKL_NUM_THREADS="2")     # seven generators side by side
    try:
        procs = {s: subprocess.Popen([sys.executable, os.path.join(ROOT, "scripts", s)], cwd=str(tmp_path), env=env,
                                     stdout=subprocess.PIPE, stderr=subprocess.STDOUT, stdin=subprocess.DEVNULL, text=True)
                 for s in SCRIPTS}
        for s, p in procs.items():
            out, _ = p.communicate(timeout=900)
            assert p.returncode == 0, (s, out[-1500:])
        problems = {}
        for f in names:
            if _sha(os.path.join(gold, f)) != before[f]:
                why = _same_arrays(os.path.join(gold, f), str(keep / f))
                if why:
                    problems[f] = why
        assert not problems, problems
    finally:
        for f in names:                                      # the working tree keeps the committed bytes either way
            if _sha(os.path.join(gold, f)) != before[f]:
                shutil.copy2(keep / f, os.path.join(gold, f))
