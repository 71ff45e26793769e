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


def test_host_chunk_plan_covers_the_batch_exactly_once(lib):
    """bhmel_host_chunk_plan is pure host arithmetic: the chunks of a host-buffer call tile the batch, the taper
    at both ends is symmetric, no chunk is larger than the steady-state one, invalid arguments give 0."""
    from beatheritage_b200 import MelSpectrogram
    for pcm in (False, True):
        for B in (1, 2, 3, 5, 6, 7, 8, 15, 16, 46, 47, 48, 49, 100, 255, 256, 257, 1090, 24576):
            for N in (513, 160000, 524160, 524161, 5_000_000):
                plan = MelSpectrogram.host_chunk_plan(B, N, pcm16=pcm)
                assert sum(plan) == B and min(plan) >= 1
                big = max(plan)
                if len(plan) > 8:
                    k = next(i for i, r in enumerate(plan) if r == big)
                    assert plan[:k] == plan[::-1][:k]                   # symmetric taper
                    assert all(plan[i] < plan[i + 1] for i in range(k))
    assert lib.bhmel_host_chunk_plan(0, 100, 0, None, 0) == 0
    assert lib.bhmel_host_chunk_plan(4, 0, 0, None, 0) == 0
    assert lib.bhmel_host_chunk_plan(4, 100, 7, None, 0) == 0
    # the 256-window bench batch: 8 fp32 / 32 int16 windows per steady-state chunk
    assert max(MelSpectrogram.host_chunk_plan(256, 524160)) == 8
    assert max(MelSpectrogram.host_chunk_plan(256, 524160, pcm16=True)) == 32


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


# ------------------------------------------------------------------ libbhstem.so
STEM_HEADER = os.path.join(ROOT, "include", "bhstem.h")


def stem_declared_functions():
    src = re.sub(r"/\*.*?\*/", "", open(STEM_HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(bhstem_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def stem_lib():
    from beatheritage_b200 import _stem_lib
    build.build_stem()
    return _stem_lib.lib()


def test_stem_header_binding_and_exports_agree(stem_lib):
    from beatheritage_b200 import _stem_lib
    assert stem_declared_functions() == sorted(_stem_lib.SIGNATURES)
    for name in stem_declared_functions():
        assert hasattr(stem_lib, name), f"libbhstem.so does not export {name}"
    assert stem_lib.bhstem_version() == 2


def test_stem_create_rejects_bad_parameters_before_touching_cuda(stem_lib):
    out = ctypes.c_void_p()
    fp = ctypes.POINTER(ctypes.c_float)
    w = (ctypes.c_float * 16)()
    wp = ctypes.cast(w, fp)
    for c_in, d in ((0, 768), (465, 768), (464, 100), (464, 0)):
        assert stem_lib.bhstem_create(c_in, d, wp, wp, wp, wp, ctypes.byref(out)) == 1
        assert out.value is None and len(stem_lib.bhstem_last_error()) > 0
    assert stem_lib.bhstem_create(464, 768, None, wp, wp, wp, ctypes.byref(out)) == 1
    assert stem_lib.bhstem_forward(None, None, 1, 2, None, None, None) == 1


def test_stem_sass_uses_tcgen05_tmem_and_tma():
    """The stem must be a Blackwell-native kernel: UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA)."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", build.build_stem()], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass            # no legacy mma.sync path
    # the epilogue's GELU runs on packed fp32 pairs, and the TMEM hand-back is a relaxed arrival: the only
    # MEMBARs left in the CTA-pair kernels belong to the two cluster barriers (start / end of the kernel)
    for mnemonic in ("FFMA2", "FMUL2", "FADD2"):
        assert mnemonic in sass, mnemonic
    kernels = sass.split("Function : ")[1:]
    pair = [k for k in kernels if "pair_kernel" in k.split("\n", 1)[0]]
    assert len(pair) >= 3                      # <8, 6, 3>, <8, 8, 2>, <16, 6, 2>
    for k in pair:
        assert k.count("MEMBAR.ALL.CTA") <= 2, k.split("\n", 1)[0]
    assert any("cond_bias_kernel" in k.split("\n", 1)[0] for k in kernels)
