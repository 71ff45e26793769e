"""Generates tests/golden/segment_cases.json by running the UNMODIFIED reference
`Preprocessor.segment` (/root/reference/osuT5/osuT5/inference/preprocessor.py:41-102) here.

The module's two imports that cannot be satisfied in this container (`config.InferenceConfig`, a
hydra dataclass, and `..dataset.data_utils`, which needs pydub / slider) are stubbed with exactly
the two names preprocessor.py uses from them (`InferenceConfig` as an annotation only,
`MILISECONDS_PER_SECOND = 1000`, data_utils.py:16); the class body itself is executed as shipped.

Run in the build container only:  python tests/golden/make_segment_golden.py
Every case stores the window start offsets (recovered from sample VALUES: the song is
1, 2, 3, ... so a window's first non-padding value names its offset; padding is 0), the number of
leading / trailing zeros of each window, `sequence_times` and `song_length`.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/osuT5/osuT5/inference/preprocessor.py"


def load_reference_preprocessor():
    cfg = types.ModuleType("config")
    cfg.InferenceConfig = object
    sys.modules.setdefault("config", cfg)
    for name in ("refpkg", "refpkg.inference", "refpkg.dataset"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    du = types.ModuleType("refpkg.dataset.data_utils")
    du.MILISECONDS_PER_SECOND = 1000          # data_utils.py:16
    du.load_audio_file = None                 # only used by Preprocessor.load (not exercised)
    sys.modules["refpkg.dataset.data_utils"] = du
    spec = importlib.util.spec_from_file_location("refpkg.inference.preprocessor", REF)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refpkg.inference.preprocessor"] = mod
    spec.loader.exec_module(mod)
    return mod.Preprocessor


def make_args(src_seq_len, lookback, lookahead, start_time, end_time, hop=128, sr=16000):
    data = SimpleNamespace(src_seq_len=src_seq_len, hop_length=hop, sample_rate=sr, normalize_audio=True)
    return SimpleNamespace(train=SimpleNamespace(data=data), lookback=lookback, lookahead=lookahead,
                           start_time=start_time, end_time=end_time)


def case_list():
    rng = np.random.default_rng(20261019)
    out = []
    # (n_samples, src_seq_len, lookback, lookahead, parallel, start_time, end_time, begin_pad, end_pad)
    out.append((2_880_000, 4096, 0.5, 0.4, False, None, None, 0, 0))       # C2 sequential: 46 windows
    out.append((2_880_000, 4096, 0.5, 0.4, True, None, None, 0, 0))        # C2 parallel: 6 windows
    out.append((2_880_000, 4096, 0.5, 0.4, False, 30_000, 90_000, 0, 0))
    out.append((2_880_000, 4096, 0.5, 0.4, False, 1_000_000, None, 0, 0))  # start past the end: keep the last
    out.append((2_880_000, 4096, 0.5, 0.4, False, None, -5, 0, 0))         # end before the start: keep the first
    out.append((2_880_000, 4096, 0.5, 0.4, False, 0, 0, 0, 0))
    out.append((100_000, 4096, 0.5, 0.4, False, None, None, 0, 0))         # shorter than one window
    out.append((524_160, 4096, 0.5, 0.4, False, None, None, 0, 0))         # exactly one window
    out.append((524_161, 4096, 0.5, 0.4, False, 100, 200, 0, 0))
    out.append((1_000_000, 1024, 0.3, 0.3, False, 12_345, 40_000, 4096, 1000))
    out.append((1_000_000, 2048, 0.0, 0.0, False, 5_000, 50_000, 0, 0))
    out.append((700_001, 512, 0.5, 0.4, True, 3_000, 30_000, 17, 0))
    for _ in range(40):
        ssl = int(rng.choice([512, 1024, 2048, 4096]))
        lb = float(rng.choice([0.0, 0.25, 0.3, 0.5]))
        la = float(rng.choice([0.0, 0.2, 0.3, 0.4]))
        n = int(rng.integers(1, 3_000_000))
        dur = n / 16.0
        st = None if rng.random() < 0.3 else float(rng.uniform(-0.2, 1.3) * dur)
        en = None if rng.random() < 0.3 else float(rng.uniform(-0.2, 1.3) * dur)
        if rng.random() < 0.5 and st is not None:
            st = int(st)
        bp = int(rng.integers(0, 3) * rng.integers(0, 5000))
        ep = int(rng.integers(0, 3) * rng.integers(0, 5000))
        out.append((n, ssl, lb, la, bool(rng.random() < 0.25), st, en, bp, ep))
    return out


def main():
    Preprocessor = load_reference_preprocessor()
    cases = []
    for (n, ssl, lb, la, par, st, en, bp, ep) in case_list():
        pre = Preprocessor(make_args(ssl, lb, la, st, en), parallel=par)
        samples = np.arange(1, n + 1, dtype=np.float32)        # exact in float32 up to 2^24
        assert n < (1 << 24)
        seqs, times, song_length = pre.segment(samples, begin_pad=bp, end_pad=ep)
        seqs = seqs.numpy()
        starts, lead, trail = [], [], []
        for w in seqs:
            nz = np.flatnonzero(w)
            if len(nz) == 0:
                starts.append(None); lead.append(len(w)); trail.append(0)
                continue
            first = int(nz[0])
            # offset of the window inside the begin-padded song: value v sits at index bp + v - 1
            starts.append(bp + int(w[first]) - 1 - first)
            lead.append(first)
            trail.append(len(w) - 1 - int(nz[-1]))
        cases.append(dict(n_samples=n, src_seq_len=ssl, lookback=lb, lookahead=la, parallel=par, start_time=st,
                          end_time=en, begin_pad=bp, end_pad=ep, samples_per_sequence=int(pre.samples_per_sequence),
                          sequence_stride=int(pre.sequence_stride), n_windows=int(seqs.shape[0]), starts=starts,
                          leading_zeros=lead, trailing_zeros=trail, sequence_times=[int(t) for t in times.tolist()],
                          song_length=float(song_length)))
    path = os.path.join(HERE, "segment_cases.json")
    with open(path, "w") as f:
        json.dump(dict(generator="tests/golden/make_segment_golden.py", reference=REF, cases=cases), f)
    print(f"wrote {len(cases)} cases -> {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
