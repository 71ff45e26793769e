"""Profiling target: a few launches of the bench workload (256 windows per launch, P0)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from beatheritage_b200 import MelSpectrogram  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--samples", type=int, default=524160)
ap.add_argument("--mels", type=int, default=80)
ap.add_argument("--variant", default=None)
ap.add_argument("--no-static-mel", action="store_true")
ap.add_argument("--static-mel", type=int, default=None, help="BHMEL_OPT_STATIC_MEL value (0..2)")
ap.add_argument("--pset", default=None, choices=["P0", "P128", "P1", "T5"], help="a reference parameter set (overrides --mels)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
PSETS = {"P0": (True, 80, 20, "reflect"), "P128": (True, 128, 20, "reflect"), "P1": (False, 388, 0, "constant"),
         "T5": (False, 512, 0, "constant")}
if a.pset:
    log, n_mels, f_min, pad = PSETS[a.pset]
    mel = MelSpectrogram("torchaudio", log, 16000, 1024, n_mels, 128, f_min, 8000, pad).to(dev)
else:
    mel = MelSpectrogram("torchaudio", True, 16000, 1024, a.mels, 128, 20, 8000, "reflect").to(dev)
if a.variant:
    mel.set_kernel_variant(a.variant)
if a.no_static_mel:
    mel.set_static_mel(False)
if a.static_mel is not None:
    mel.set_static_mel(a.static_mel)
g = torch.Generator(device=dev).manual_seed(1234)
x = torch.rand(a.batch, a.samples, device=dev, generator=g).mul_(2).sub_(1)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
ev[0].record()
for i in range(a.reps):
    y = mel(x)
    ev[i + 1].record()
torch.cuda.synchronize()
import hashlib  # noqa: E402
digest = hashlib.sha1(y.cpu().numpy().tobytes()).hexdigest()[:16]
print("ms per launch:", [round(ev[i].elapsed_time(ev[i + 1]), 4) for i in range(a.reps)], "sum", float(y.sum()), "sha1", digest)
