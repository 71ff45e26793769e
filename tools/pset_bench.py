"""Per-parameter-set timing of the frontend kernel (256 windows per launch, device resident):
P0 (baked filterbank -> generated mel stage), P0 with the generic stage, P128, P1 (388 mels), T5 (512 mels).
One JSON line per set; `BHMEL_LIB=build/libbhmel_x.so` selects an A/B build.

    python tools/pset_bench.py [--batch 256] [--reps 30] [--sets P0,P1,...]
"""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from beatheritage_b200 import MelSpectrogram  # noqa: E402

SETS = {   # name -> (ctor args, static mel allowed)
    "P0": (("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect"), True),
    "P0_generic": (("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect"), False),
    "P0_hybrid": (("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect"), 2),
    "M64": (("torchaudio", True, 16000, 1024, 64, 128, 0, 8000, "reflect"), True),
    "M64_pairs": (("torchaudio", True, 16000, 1024, 64, 128, 0, 8000, "reflect"), False),
    "P128": (("torchaudio", True, 16000, 1024, 128, 128, 20, 8000, "reflect"), True),
    "P1": (("torchaudio", False, 16000, 1024, 388, 128, 0, 8000, "constant"), True),
    "P1_generic": (("torchaudio", False, 16000, 1024, 388, 128, 0, 8000, "constant"), False),
    "T5_generic": (("torchaudio", False, 16000, 1024, 512, 128, 0, 8000, "constant"), False),
    "P128_generic": (("torchaudio", True, 16000, 1024, 128, 128, 20, 8000, "reflect"), False),
    "T5": (("torchaudio", False, 16000, 1024, 512, 128, 0, 8000, "constant"), True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--samples", type=int, default=524160)
    ap.add_argument("--sets", default=",".join(SETS))
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1234)
    x = torch.rand(a.batch, a.samples, device=dev, generator=g).mul_(2).sub_(1)
    T = a.samples // 128 + 1
    for name in a.sets.split(","):
        args, static = SETS[name]
        mel = MelSpectrogram(*args).to(dev)
        if static is not True:
            mel.set_static_mel(int(static))
        y = torch.empty(a.batch, T, args[4], device=dev)
        for _ in range(5):
            mel.forward_into(x, y)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
        ev[0].record()
        for i in range(a.reps):
            mel.forward_into(x, y)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps)]
        ms = statistics.median(ts)
        bytes_ = a.batch * (4 * a.samples + 4 * T * args[4])
        print(json.dumps({"set": name, "n_mels": args[4], "ms_median": round(ms, 4), "ms_min": round(min(ts), 4),
                          "audio_s_per_s": round(a.batch * a.samples / 16000 / (ms / 1e3)),
                          "algorithmic_GBps": round(bytes_ / ms / 1e6, 1), "checksum": float(y.double().sum())}), flush=True)
        del mel, y


if __name__ == "__main__":
    main()
