import ctypes
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


# ------------------------------------------------------------------ golden fixtures
PSET_ARGS = {   # log_scale, n_mels, f_min, f_max, pad_mode  (sr 16000, n_fft 1024, hop 128 everywhere)
    "P0": (True, 80, 20, 8000, "reflect"),
    "P1": (False, 388, 0, 8000, "constant"),
    "T5": (False, 512, 0, 8000, "constant"),
    "P128": (True, 128, 20, 8000, "reflect"),
    "P0C": (True, 80, 20, 8000, "constant"),
}


def golden_case_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "case_*.npz")))


def load_params(pset):
    z = np.load(os.path.join(GOLDEN, f"params_{pset}.npz"))
    return np.ascontiguousarray(z["window"]), np.ascontiguousarray(z["fb"])


def regenerate_input(case):
    """Rebuild the input of a golden case from its recipe (or the stored x)."""
    from tests.golden import signals
    if "x" in case.files:
        return np.ascontiguousarray(case["x"], dtype=np.float32)
    rec = eval(str(case["recipe"]))   # a dict literal written by make_golden.py
    k = rec["kind"]
    if k == "noise":
        x = signals.noise(rec["B"], rec["N"], rec["seed"])
    elif k == "sine":
        x = signals.sine(rec["B"], rec["N"], rec["freq"], rec["amp"])
    elif k == "zeros":
        x = np.zeros((rec["B"], rec["N"]), np.float32)
    elif k == "impulse":
        x = signals.impulse(rec["N"], rec["pos"])
    elif k == "music":
        x = signals.music(rec["B"] * rec["N"], rec["seed"]).reshape(rec["B"], rec["N"])
    elif k == "noise_zero_tail":
        x = signals.noise(rec["B"], rec["N"], rec["seed"])
        x[:, rec["real_hop_frames"] * 128:] = 0.0
    else:
        raise KeyError(k)
    import hashlib
    sha = np.frombuffer(hashlib.sha256(x.tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, case["x_sha"]), "regenerated input differs from the one the fixture was made from"
    return x


def load_case(name):
    case = np.load(os.path.join(GOLDEN, f"case_{name}.npz"), allow_pickle=False)
    pset = str(case["pset"])
    frames = case["frames"] if "frames" in case.files else None
    return case, pset, frames


def parity_error(y, y_ref, log_scale):
    """Max abs error in the log1p domain (the domain the 1e-3 bar is stated in)."""
    y = np.asarray(y, np.float64)
    y_ref = np.asarray(y_ref, np.float64)
    if not log_scale:
        y, y_ref = np.log1p(np.maximum(y, 0)), np.log1p(np.maximum(y_ref, 0))
    return float(np.abs(y - y_ref).max())


# ------------------------------------------------------------------ CPU lane emulator of the kernel
@pytest.fixture(scope="session")
def emu_lib():
    build = os.path.join(ROOT, "tests", "emu", "_build")
    os.makedirs(build, exist_ok=True)
    so = os.path.join(build, "libbhmel_emu.so")
    from beatheritage_b200 import build as bbuild
    gen = bbuild.generate()
    src = os.path.join(ROOT, "tests", "emu", "emu.cpp")
    csrc = os.path.join(ROOT, "beatheritage_b200", "csrc")
    deps = [src, gen, os.path.join(csrc, "bhmel_tables.h"), os.path.join(csrc, "mel_static_gen.h"),
            os.path.join(csrc, "bhmel_fb_baked.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src],
                       check=True)
    lib = ctypes.CDLL(so)
    fp = ctypes.POINTER(ctypes.c_float)
    ll = ctypes.c_longlong
    lib.bhmel_emu_forward.argtypes = [fp, ll, ll, ll, ll, ll, ctypes.c_int, fp, fp, ctypes.c_double,
                                      ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
    lib.bhmel_emu_forward.restype = ctypes.c_int

    def run(x, n_mels, fb=None, window=None, f_min=20.0, f_max=8000.0, reflect=True, log=True,
            gather=None, rounds=False, static_mel=False):
        """gather = (first_offset, stride, W, window_len) treats x as a 1-D song.  rounds / static_mel pick
        the mel stage of the independent-warp kernel / the generated-code stage of a baked filterbank
        (static_mel=True: the direct form, as BHMEL_OPT_STATIC_MEL = 1; static_mel=2: P0's hybrid form)."""
        x = np.ascontiguousarray(x, np.float32)
        if gather is None:
            B, N = x.shape
            stride, row0, n_total = N, 0, 2 ** 62
        else:
            row0, stride, B, N = gather
            n_total = x.size
        y = np.zeros((B, N // 128 + 1, n_mels), np.float32)
        fbp = fb.ctypes.data_as(fp) if fb is not None else None
        wp = window.ctypes.data_as(fp) if window is not None else None
        rc = lib.bhmel_emu_forward(x.ctypes.data_as(fp), B, N, stride, row0, n_total, n_mels, fbp, wp,
                                   float(f_min), float(f_max), 16000, int(reflect), int(log),
                                   (2 if rounds else 0) | (4 if static_mel else 0) | (8 if static_mel == 2 else 0),
                                   y.ctypes.data_as(fp))
        assert rc == 0
        return y
    return run
