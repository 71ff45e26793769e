"""Runs the frontend and the conv stem through their BOUNDS-ASSERTING debug libraries
(beatheritage_b200/libbhmel_bounds.so / libbhstem_bounds.so: -DBHMEL_BOUNDS / -DBHSTEM_BOUNDS turn every
shared- and global-memory index of the staging, transpose, power-buffer, mel, store and TMEM-epilogue
paths into a trap-on-violation assert).  compute-sanitizer is closed on the GPU pool, so this is the
memory-safety net: ragged lengths, unaligned rows, odd-stride gathers, every mel stage (hybrid, direct,
generic), pitched / bf16 outputs, all three kernel schedules, tiny and ragged stems.  A violation prints
the failed expression from the kernel and the process dies with a CUDA error.

    BHMEL_LIB=.../libbhmel_bounds.so BHSTEM_LIB=.../libbhstem_bounds.so python tools/bounds_cases.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

assert "bounds" in os.environ.get("BHMEL_LIB", "") and "bounds" in os.environ.get("BHSTEM_LIB", ""), \
    "point BHMEL_LIB / BHSTEM_LIB at the bounds-asserting builds"
from beatheritage_b200 import MelSpectrogram  # noqa: E402
from beatheritage_b200.conv_stem import ConvStem  # noqa: E402
from oracle import conv_stem_oracle, mel_oracle  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.random.default_rng(5)
SETS = {"P0": ("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect"),
        "P0C": ("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "constant"),
        "P128": ("torchaudio", True, 16000, 1024, 128, 128, 20, 8000, "reflect"),
        "P1": ("torchaudio", False, 16000, 1024, 388, 128, 0, 8000, "constant"),
        "T5": ("torchaudio", False, 16000, 1024, 512, 128, 0, 8000, "constant"),
        "M33": ("torchaudio", True, 16000, 1024, 33, 128, 0, 8000, "reflect")}
n_cases = 0


def check(m, x, y, args):
    ref = mel_oracle.mel_forward(x, fb=m.transform.mel_scale.fb.cpu().numpy(), window=m.transform.spectrogram.window.cpu().numpy(),
                                 pad_mode=args[8], log_scale=args[1], dtype=np.float64)
    got = y.float().cpu().numpy().astype(np.float64)
    err = np.abs(got - ref).max() if args[1] else np.abs(np.log1p(got) - np.log1p(ref)).max()
    assert err < 2e-5, err


for name, args in SETS.items():
    m = MelSpectrogram(*args).to(dev)
    lo = 513 if args[8] == "reflect" else 1
    for variant in ("ws", "barrier", "warp"):
        m.set_kernel_variant(variant)
        for static in ((1, 0, 2) if variant == "ws" else (1,)):
            m.set_static_mel(static)
            for (B, N) in ((1, lo), (3, 1000 if lo == 1 else 1025), (2, 4133), (5, 40001), (1, 524160), (149, 4096)):
                x = (rng.random((B, N), dtype=np.float32) * 2 - 1)
                y = m(torch.from_numpy(x).to(dev))
                torch.cuda.synchronize()
                if B * N < 300000:
                    check(m, x, y, args)
                n_cases += 1
    m.set_kernel_variant("ws")
    m.set_static_mel(1)
    # unaligned rows + row stride, gather at the reference's odd stride, zero tail past the song's end
    base = torch.rand(4 * 70003 + 16, device=dev) * 2 - 1
    for off in (0, 1, 2, 3):
        xv = base[off:off + 4 * 70003].view(4, 70003)[:, :66000]
        y = m(xv)
        torch.cuda.synchronize()
        check(m, xv.cpu().numpy(), y, args)
        n_cases += 1
    song = torch.rand(700001, device=dev) * 2 - 1
    for (first, stride, W, wlen) in ((0, 52415, 12, 262144), (3, 131040, 5, 524160), (17, 999, 40, 4000)):
        m.forward_gather(song, first, stride, W, wlen)
        torch.cuda.synchronize()
        n_cases += 1
    # typed / pitched outputs: bf16 and f32 into a wider buffer, aligned and unaligned channel offsets
    xt = torch.rand(3, 33000, device=dev) * 2 - 1
    T = 33000 // 128 + 1
    for dt in (torch.float32, torch.bfloat16):
        for choff in (0, 4, 3):
            wide = torch.zeros(3, T, args[4] + 12, dtype=dt, device=dev)
            m.forward_into(xt, wide, channel_offset=choff)
            torch.cuda.synchronize()
            assert bool((wide[..., :choff] == 0).all()) and bool((wide[..., choff + args[4]:] == 0).all())
            n_cases += 1
    # the assembled encoder input, both layouts
    cond = torch.randn(3, 8, device=dev)
    for cf in (False, True):
        m.forward_encoder_input(xt, [cond], dtype=torch.bfloat16, channels_first=cf)
        torch.cuda.synchronize()
        n_cases += 1
    # a perturbed filterbank: generic stage of the default schedule
    with torch.no_grad():
        m.transform.mel_scale.fb[100, min(36, args[4] - 1)] += 0.25
    x = (rng.random((2, 9000), dtype=np.float32) * 2 - 1)
    check(m, x, m(torch.from_numpy(x).to(dev)), args)
    n_cases += 1
    del m

# ---- conv stem: whisper-small dims, ragged rows, whisper-tiny width, the smallest problem
torch.manual_seed(0)
for (B, T, c_in, d) in ((2, 256, 464, 768), (1, 200, 464, 768), (3, 64, 80, 384), (1, 2, 8, 128), (1, 4096, 464, 768)):
    stem = ConvStem(c_in, d)
    with torch.no_grad():
        for prm in stem.parameters():
            prm.copy_(prm.to(torch.bfloat16).float())
    stem = stem.to(dev)
    x = (torch.randn(B, T, c_in) * 1.5).to(torch.bfloat16)
    y = stem(x.to(dev))
    torch.cuda.synchronize()
    if B * T <= 1024:
        want = conv_stem_oracle.conv_stem(x, stem.conv1.weight.cpu(), stem.conv1.bias.cpu(), stem.conv2.weight.cpu(), stem.conv2.bias.cpu())
        diff = (y.float().cpu() - want).abs()
        assert bool((diff <= want.abs() * 2.0 ** -6 + 4e-3).all()), float(diff.max())
    n_cases += 1
# ---- split conv stem (folded conditioning channels, 8 and 16 epilogue warps): the reference's dims, ragged rows,
# every frame an edge, the 128-column kernel
for (B, T, c_in, d, n_var) in ((2, 256, 464, 768, 80), (1, 200, 464, 768, 80), (1, 2, 16, 128, 8), (3, 64, 80, 384, 16),
                               (17, 258, 464, 768, 80)):
    stem = ConvStem(c_in, d)
    with torch.no_grad():
        for prm in stem.parameters():
            prm.copy_(prm.to(torch.bfloat16).float())
    stem = stem.to(dev)
    x = (torch.randn(B, T, c_in) * 1.5).to(torch.bfloat16)
    x[:, :, n_var:] = x[:, :1, n_var:]
    for epi in (16, 8):
        stem.set_epilogue_warps(8, 8, epi)
        y = stem.forward_split(x[:, :, :n_var].contiguous().to(dev), x[:, 0, n_var:].contiguous().to(dev))
        torch.cuda.synchronize()
        if B * T <= 1024:
            want = conv_stem_oracle.conv_stem(x, stem.conv1.weight.cpu(), stem.conv1.bias.cpu(), stem.conv2.weight.cpu(), stem.conv2.bias.cpu())
            diff = (y.float().cpu() - want).abs()
            assert bool((diff <= want.abs() * 2.0 ** -6 + 4e-3).all()), float(diff.max())
        n_cases += 1
print(f"bounds ok: {n_cases} cases through libbhmel_bounds.so / libbhstem_bounds.so, no assertion fired")
