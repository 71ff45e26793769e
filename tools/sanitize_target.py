"""Small target for compute-sanitizer: every kernel variant, both staging paths, edge tiles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from beatheritage_b200 import MelSpectrogram  # noqa: E402

dev = torch.device("cuda", 0)
for n_mels, pad in ((80, "reflect"), (388, "constant")):
    mel = MelSpectrogram("torchaudio", True, 16000, 1024, n_mels, 128, 20, 8000, pad).to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.rand(3, 40000 + 13, device=dev, generator=g) * 2 - 1
    outs = []
    for variant in ("ws", "barrier", "warp"):
        mel.set_kernel_variant(variant)
        for bulk in (True, False):
            mel.set_bulk_copy(bulk)
            outs.append(mel(x[:, :40000]))          # aligned rows
            outs.append(mel(x[:, 1:40001]))         # unaligned rows
    song = x[0]
    outs.append(mel.forward_gather(song, 3, 397, 5, 31 * 128))
    torch.cuda.synchronize()
    ref = outs[0]
    assert all(torch.equal(o, ref) for o in outs[0:12:2]), "variants disagree"
print("sanitize target ok")
