"""CPU: the C-ABI library builds, loads and exports every symbol include/bhmel.h declares.
No compute calls here (no GPU in the build container)."""
import ctypes
import os
import re

import pytest

from beatheritage_b200 import _lib, build
from tests.conftest import ROOT

HEADER = os.path.join(ROOT, "include", "bhmel.h")


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bhmel_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_the_binding_binds():
    assert declared_functions() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"libbhmel.so does not export {name}"


def test_version_and_kernel_info(lib):
    assert lib.bhmel_version() == 101
    smem, threads, tile = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    lib.bhmel_kernel_info(ctypes.byref(smem), ctypes.byref(threads), ctypes.byref(tile))
    assert threads.value == 512 and tile.value == 32
    assert 0 < smem.value <= 227 * 1024          # fits one CTA per SM on sm_100


def test_create_rejects_bad_parameters_before_touching_cuda(lib):
    out = ctypes.c_void_p()
    bad = [
        dict(n_fft=2048),                 # only 1024/128 are compiled in
        dict(hop_length=160),
        dict(n_mels=0),
        dict(n_mels=5000),
        dict(pad_mode=7),
        dict(f_min=9000.0),
    ]
    for kw in bad:
        base = dict(sample_rate=16000, n_fft=1024, hop_length=128, n_mels=80, f_min=20.0, f_max=8000.0,
                    pad_mode=1, log_scale=1)
        base.update(kw)
        prm = _lib.BhmelParams(base["sample_rate"], base["n_fft"], base["hop_length"], base["n_mels"],
                               base["f_min"], base["f_max"], base["pad_mode"], base["log_scale"], None, None)
        rc = lib.bhmel_create(ctypes.byref(prm), ctypes.byref(out))
        assert rc == _lib.EINVAL, kw
        assert out.value is None
        assert len(lib.bhmel_last_error()) > 0
    assert lib.bhmel_create(None, ctypes.byref(out)) == _lib.EINVAL


def test_num_frames(lib):
    assert lib.bhmel_num_frames(None, 524160) == 4096
    assert lib.bhmel_num_frames(None, 160000) == 1251
    assert lib.bhmel_num_frames(None, 513) == 5


def test_sass_is_sm100_and_uses_tma_bulk_copy():
    """The built library must carry sm_100a SASS with the TMA bulk copy (UBLKCP) in it."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([exe, "-sass", build.LIB], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass and "LDGSTS" in sass and "SYNCS" in sass
    assert "bhmel_logmel_kernel" in sass
