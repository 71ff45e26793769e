"""GPU (-m gpu): the bounds-asserting debug builds (compute-sanitizer is closed on the
GPU pool).  `beatheritage_b200.build.build(bounds=True)` -- run by __graft_entry__.build() -- compiles
libbhmel_bounds.so (-DBHMEL_BOUNDS) and libbhstem_bounds.so (-DBHSTEM_BOUNDS), in which every shared- and
global-memory index of the staging, transpose, power-buffer, mel, store and TMEM-epilogue paths is an
assert that prints the expression and traps.  tools/bounds_cases.py drives ragged / unaligned / gather /
pitched / tiny-stem cases through them in a separate process (a trap kills the CUDA context)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_bounds_asserting_builds_run_the_edge_cases_clean():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from beatheritage_b200 import build as _build
    from tests.conftest import ROOT
    mel_lib, stem_lib = _build.build_bounds()          # no-op when the in-tree files are current
    env = dict(os.environ, BHMEL_LIB=mel_lib, BHSTEM_LIB=stem_lib)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bounds_cases.py")], env=env, capture_output=True,
                         text=True, timeout=900)
    tail = res.stdout[-3000:] + res.stderr[-3000:]
    assert "BOUNDS violated" not in tail, tail
    assert res.returncode == 0, tail
    assert "bounds ok" in res.stdout, tail


def test_the_shipped_libraries_carry_no_bounds_code():
    """The asserts are compiled out of the product: the release libraries hold none of their strings."""
    from beatheritage_b200 import build as _build
    for lib in (_build.LIB, _build.LIB_STEM):
        blob = open(lib, "rb").read()
        assert b"BOUNDS violated" not in blob, lib
    for lib in (_build.LIB_BOUNDS, _build.LIB_STEM_BOUNDS):
        if os.path.exists(lib):
            assert b"BOUNDS violated" in open(lib, "rb").read(), lib
