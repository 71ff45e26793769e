"""Generates tests/golden/*.npz by running the UNMODIFIED reference module
(/root/reference/osuT5/osuT5/model/spectrogram.py, implementation="torchaudio") on CPU fp32.

Run in the build container only:  python tests/golden/make_golden.py
Inputs are regenerated from seeds by tests/golden/signals.py; fixtures store the reference's
buffers (window, fb), its outputs (all frames, or a listed subset of frames for full
524 160-sample windows to keep the files small) and, for tiny cases, the input too.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402
from oracle import mel_oracle  # noqa: E402
from tests.golden import signals  # noqa: E402

# (name) -> ctor args after `implementation`: log_scale, sr, n_fft, n_mels, hop, f_min, f_max, pad_mode
PARAM_SETS = {
    "P0": (True, 16000, 1024, 80, 128, 20, 8000, "reflect"),      # configs/model/whisper_small_v2.yaml:16-21
    "P1": (False, 16000, 1024, 388, 128, 0, 8000, "constant"),    # configs/model/default.yaml:18-27 (torchaudio arithmetic)
    "T5": (False, 16000, 1024, 512, 128, 0, 8000, "constant"),    # configs/model/t5_small.yaml:9-10
    "P128": (True, 16000, 1024, 128, 128, 20, 8000, "reflect"),   # configs/train/tiny_dist22.yaml:10-12
    "P0C": (True, 16000, 1024, 80, 128, 20, 8000, "constant"),    # osuT5/dataloading.py:81-90 omits pad_mode
}

WINDOW = 524160


def frame_subset(T: int, extra=()):
    idx = set(range(0, T, 16)) | set(range(0, 8)) | set(range(T - 8, T)) | set(extra)
    return np.array(sorted(i for i in idx if 0 <= i < T), dtype=np.int64)


def cases():
    """yield (case_name, param_set, recipe dict, x [B,N] float32, frames or None)"""
    for N in (513, 1000, 1024, 1025, 4133):
        yield f"noise_N{N}", "P0", dict(kind="noise", B=2, N=N, seed=100 + N), signals.noise(2, N, 100 + N), None
    yield "c1_noise_10s", "P0", dict(kind="noise", B=1, N=160000, seed=0), signals.noise(1, 160000, 0), None
    yield "sine440", "P0", dict(kind="sine", B=1, N=16000, freq=440.0, amp=0.5), signals.sine(1, 16000), None
    yield "sine_fullscale_3k", "P0", dict(kind="sine", B=1, N=8192, freq=3000.0, amp=1.0), signals.sine(1, 8192, 3000.0, 1.0), None
    yield "zeros", "P0", dict(kind="zeros", B=2, N=4096), np.zeros((2, 4096), np.float32), None
    yield "impulse0", "P0", dict(kind="impulse", N=2048, pos=0), signals.impulse(2048, 0), None
    yield "impulse_last", "P0", dict(kind="impulse", N=2048, pos=2047), signals.impulse(2048, 2047), None
    for B in (6, 16, 46):
        yield f"batch_B{B}", "P0", dict(kind="noise", B=B, N=2048, seed=200 + B), signals.noise(B, 2048, 200 + B), None
    # full model-context windows: music-like song, 2 windows, subset of frames
    song = signals.music(2 * WINDOW, seed=1)
    x = song.reshape(2, WINDOW)
    yield "music_2win", "P0", dict(kind="music", B=2, N=WINDOW, seed=1), x, frame_subset(4096)
    # last training window: 2026 of 4095 hop-frames real, zero tail (ors_dataset.py:577-586)
    xt = signals.noise(1, WINDOW, 7)
    xt[0, 2026 * 128:] = 0.0
    yield "zero_tail", "P0", dict(kind="noise_zero_tail", B=1, N=WINDOW, seed=7, real_hop_frames=2026), xt, \
        frame_subset(4096, extra=range(2010, 2046))
    yield "noise_N524161", "P0", dict(kind="noise", B=1, N=WINDOW + 1, seed=8), signals.noise(1, WINDOW + 1, 8), \
        frame_subset(4096)
    # secondary parameter sets
    yield "p1_noise", "P1", dict(kind="noise", B=2, N=20000, seed=300), signals.noise(2, 20000, 300), None
    yield "t5_noise", "T5", dict(kind="noise", B=1, N=8000, seed=301), signals.noise(1, 8000, 301), None
    yield "p128_noise", "P128", dict(kind="noise", B=1, N=8000, seed=302), signals.noise(1, 8000, 302), None
    yield "p0c_noise", "P0C", dict(kind="noise", B=2, N=5000, seed=303), signals.noise(2, 5000, 303), None
    yield "p1_music", "P1", dict(kind="music", B=1, N=40000, seed=5), signals.music(40000, 5)[None], None


def main():
    Ref = ref_loader.load_reference_class()
    torch.set_num_threads(os.cpu_count())
    mods = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, args in PARAM_SETS.items():
            mods[name] = Ref("torchaudio", *args).eval()
            sd = mods[name].state_dict()
            np.savez(os.path.join(HERE, f"params_{name}.npz"),
                     window=sd["transform.spectrogram.window"].numpy(),
                     fb=sd["transform.mel_scale.fb"].numpy(),
                     args=np.array([str(a) for a in args]))
    total = 0
    for case, pset, recipe, x, frames in cases():
        with torch.no_grad():
            y = mods[pset](torch.from_numpy(x)).contiguous().numpy()
            y64 = mods[pset].double()(torch.from_numpy(x).double()).contiguous().numpy()
            mods[pset].float()
        assert y.dtype == np.float32 and y.shape[:2] == (x.shape[0], x.shape[1] // 128 + 1)
        payload = dict(pset=pset, recipe=str(recipe), shape=np.array(y.shape),
                       x_sha=np.frombuffer(__import__("hashlib").sha256(x.tobytes()).digest(), dtype=np.uint8))
        if frames is not None:
            payload["frames"] = frames
            y, y64 = y[:, frames], y64[:, frames]
        payload["y"] = y
        if y.size <= 20000:     # fp64 run of the same module (the arbiter); the fp64 oracle
            payload["y64"] = y64  # reproduces it to 1e-12, so big cases recompute it instead
        if x.size <= 8192:
            payload["x"] = x
        path = os.path.join(HERE, f"case_{case}.npz")
        np.savez_compressed(path, **payload)
        total += os.path.getsize(path)
        # report how far the numpy oracle is from the reference on this case (sanity, not a gate)
        p = PARAM_SETS[pset]
        yo = mel_oracle.mel_forward(x, n_mels=p[3], f_min=p[5], f_max=p[6], pad_mode=p[7], log_scale=p[0],
                                    dtype=np.float64, fb=sd_fb(mods[pset]), window=sd_win(mods[pset]))
        yo = yo if frames is None else yo[:, frames]
        print(f"{case:18s} {pset:5s} x{tuple(x.shape)} -> y{tuple(y.shape)} "
              f"max|oracle64-ref64|={np.abs(yo - y64).max():.3e} "
              f"max|ref32-ref64|={np.abs(y - y64).max():.3e}")
    print(f"wrote fixtures, {total / 1e6:.2f} MB")


def sd_fb(m):
    return m.state_dict()["transform.mel_scale.fb"].double().numpy()


def sd_win(m):
    return m.state_dict()["transform.spectrogram.window"].double().numpy()


if __name__ == "__main__":
    main()
