"""GPU (-m gpu): parity of the CUDA path against the oracle and the reference fixtures, called
through the reference-facing module and the C ABI.  Tolerance: max abs error in the log1p
domain <= 1e-3 (north star), and we additionally hold the kernel to 2e-5 (fp32 FFT noise)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import mel_oracle, torch_port
from tests.conftest import PSET_ARGS, golden_case_names, load_case, load_params, parity_error, regenerate_input
from tests.golden import signals

pytestmark = pytest.mark.gpu

TOL = 1e-3
TARGET = 2e-5
WINDOW = 524160
DEFAULT_VARIANT = "ws"


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def mods(dev):
    from beatheritage_b200 import MelSpectrogram
    out = {}
    for pset, (log, n_mels, f_min, f_max, pad) in PSET_ARGS.items():
        m = MelSpectrogram("torchaudio", log, 16000, 1024, n_mels, 128, f_min, f_max, pad).to(dev)
        out[pset] = m
    return out


def run(mod, x, dev):
    y = mod(torch.from_numpy(np.ascontiguousarray(x)).to(dev))
    torch.cuda.synchronize()
    return y.cpu().numpy()


def test_native_library_is_loaded_and_kernel_launches(mods, dev):
    from beatheritage_b200 import _lib
    m = mods["P0"]
    before = m.launch_count()
    y = m(torch.zeros(1, 2048, device=dev))
    torch.cuda.synchronize()
    assert m.launch_count() == before + 1
    assert y.shape == (1, 17, 80) and y.dtype == torch.float32 and y.is_contiguous()
    assert torch.all(y == 0)                     # zero input -> exactly 0.0 (log1p(0)), SURVEY.md 0.1
    maps = open("/proc/self/maps").read()
    assert "libbhmel.so" in maps
    assert _lib.lib().bhmel_version() == 101


@pytest.mark.parametrize("name", golden_case_names())
def test_golden_fixtures(mods, dev, name):
    case, pset, frames = load_case(name)
    log = PSET_ARGS[pset][0]
    x = regenerate_input(case)
    y = run(mods[pset], x, dev)
    assert y.shape == tuple(case["shape"])
    if frames is not None:
        y = y[:, frames]
    err = parity_error(y, case["y"], log)
    assert err < TOL
    assert err < TARGET, f"{name}: {err}"


@pytest.mark.parametrize("pset,shape", [("P0", (3, 40000)), ("P0", (2, 524160)), ("P0", (5, 513)), ("P0C", (2, 100)),
                                        ("P1", (2, 30000)), ("T5", (1, 20000)), ("P128", (3, 9999))])
def test_kernel_variants_are_bit_identical(mods, dev, pset, shape):
    """The three schedules (barrier, independent warps, warp-specialised with the generic mel stage)
    run the same arithmetic bit for bit; the warp-specialised kernel's direct mel stages (the default for
    every baked reference filterbank) sum in a different order and stay within a few ulp."""
    m = mods[pset]
    x = signals.noise(shape[0], shape[1], 77 + shape[1])
    m.set_kernel_variant("barrier")
    y_bar = run(m, x, dev)
    m.set_kernel_variant("warp")
    y_iw = run(m, x, dev)
    m.set_kernel_variant("ws")
    y_ws_default = run(m, x, dev)
    m.set_static_mel(False)
    y_ws = run(m, x, dev)
    m.set_static_mel(True)
    m.set_kernel_variant(DEFAULT_VARIANT)
    assert np.array_equal(y_iw, y_bar)
    assert np.array_equal(y_ws, y_bar)
    if PSET_ARGS[pset][0]:
        assert np.abs(y_ws_default - y_bar).max() <= 2e-6
    else:
        assert np.all(np.abs(y_ws_default - y_bar) <= 4e-6 * np.abs(y_bar))
    log, n_mels, f_min, f_max, pad = PSET_ARGS[pset]
    window, fb = load_params(pset)
    ref = mel_oracle.mel_forward(x, fb=fb, window=window, pad_mode=pad, log_scale=log, dtype=np.float64)
    assert parity_error(y_ws, ref, log) < TARGET


@pytest.mark.parametrize("log", [True, False])
@pytest.mark.parametrize("shape", [(1, 513), (3, 40000), (2, 524160), (7, 4097)])
def test_static_mel_stage_is_bit_identical_to_generic(dev, log, shape):
    """P0's filterbank is recognised as the baked table.  Its HYBRID generated stage
    (BHMEL_OPT_STATIC_MEL = 2) must not differ from the generic stage in a single bit; its default
    direct form stays within a few ulp.  A filterbank that differs in one weight must fall back to
    the generic stage by itself."""
    from beatheritage_b200 import MelSpectrogram
    m = MelSpectrogram("torchaudio", log, 16000, 1024, 80, 128, 20, 8000, "reflect").to(dev)
    x = signals.noise(shape[0], shape[1], 5 + shape[1])
    y_direct = run(m, x, dev)
    m.set_static_mel(2)
    y_static = run(m, x, dev)
    m.set_static_mel(False)
    y_generic = run(m, x, dev)
    m.set_static_mel(True)
    assert np.array_equal(y_static.view(np.uint32), y_generic.view(np.uint32))
    if log:
        assert np.abs(y_direct - y_generic).max() <= 2e-6
    else:
        assert np.all(np.abs(y_direct - y_generic) <= 4e-6 * np.abs(y_generic))
    window, fb = load_params("P0")
    ref = mel_oracle.mel_forward(x, fb=fb, window=window, pad_mode="reflect", log_scale=log, dtype=np.float64)
    assert parity_error(y_static, ref, log) < TARGET
    # perturb one weight: no longer the baked table -> generic stage, result follows the new weight
    fb2 = fb.copy()
    fb2[100, 36] *= 1.5
    with torch.no_grad():
        m.transform.mel_scale.fb.copy_(torch.from_numpy(fb2))
    y2 = run(m, x, dev)
    ref2 = mel_oracle.mel_forward(x, fb=fb2, window=window, pad_mode="reflect", log_scale=log, dtype=np.float64)
    assert parity_error(y2, ref2, log) < TARGET


@pytest.mark.parametrize("pset,args", [
    ("P128", ("torchaudio", True, 16000, 1024, 128, 128, 20, 8000, "reflect")),
    ("P1", ("torchaudio", False, 16000, 1024, 388, 128, 0, 8000, "constant")),
    ("T5", ("torchaudio", False, 16000, 1024, 512, 128, 0, 8000, "constant")),
    ("T5log", ("torchaudio", True, 16000, 1024, 512, 128, 0, 8000, "reflect")),
])
@pytest.mark.parametrize("shape", [(1, 513), (3, 40000), (2, 262016)])
def test_direct_mel_stages_agree_with_generic_and_oracle(dev, pset, args, shape):
    """The other reference filterbanks (128 / 388 / 512 mels) are recognised too and take their direct
    generated stage (one or two chains per filter, vector stores from registers).  Against the generic
    stage: a few ulp of the linear mel value, all-zero filters (20 of P1's, 54 of T5's) exactly 0; the
    result meets the oracle; bfloat16 / pitched output and an unaligned output (falls back to the
    generic stage) behave like forward + cast."""
    from beatheritage_b200 import MelSpectrogram
    m = MelSpectrogram(*args).to(dev)
    x = signals.noise(shape[0], shape[1], 11 + shape[1])
    y_static = run(m, x, dev)
    m.set_static_mel(False)
    y_generic = run(m, x, dev)
    m.set_static_mel(True)
    log = args[1]
    if log:   # log domain: a few float32 ulp of values <= ~12
        assert np.abs(y_static - y_generic).max() <= 2e-6
    else:
        assert np.all(np.abs(y_static - y_generic) <= 4e-6 * np.abs(y_generic))
    window, fb = load_params(pset.replace("log", ""))
    ref = mel_oracle.mel_forward(x, fb=fb, window=window, pad_mode=args[8], log_scale=log, dtype=np.float64)
    assert parity_error(y_static, ref, log) < TARGET
    empty = np.flatnonzero(~(fb != 0).any(axis=0))
    assert np.all(y_static[..., empty] == 0.0)
    # typed / pitched output through the direct stores: [B, T, n_mels + 8] bf16, mel channels first
    xt = torch.from_numpy(x).to(dev)
    T, M = x.shape[1] // 128 + 1, fb.shape[1]
    wide = torch.full((x.shape[0], T, M + 8), 7.0, dtype=torch.bfloat16, device=dev)
    m.forward_into(xt, wide)
    torch.cuda.synchronize()
    assert torch.equal(wide[..., :M].cpu(), torch.from_numpy(y_static).to(torch.bfloat16))
    assert bool((wide[..., M:] == 7.0).all())
    # channel offsets 1 / 3: rows are no longer 8-byte aligned -> element stores, the very same values
    for choff, dt in ((1, torch.bfloat16), (3, torch.float32)):
        wide2 = torch.zeros((x.shape[0], T, M + 8), dtype=dt, device=dev)
        m.forward_into(xt, wide2, channel_offset=choff)
        torch.cuda.synchronize()
        assert torch.equal(wide2[..., choff:M + choff].cpu(), torch.from_numpy(y_static).to(dt))
        assert bool((wide2[..., :choff] == 0).all()) and bool((wide2[..., M + choff:] == 0).all())


def test_p0_direct_and_hybrid_forms(dev):
    """P0's default is its direct form; BHMEL_OPT_STATIC_MEL = 2 selects the hybrid form: same tolerance
    against the oracle, silence exactly 0 in both."""
    from beatheritage_b200 import MelSpectrogram
    m = MelSpectrogram("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect").to(dev)
    x = signals.noise(3, 70000, 21)
    y_direct = run(m, x, dev)
    m.set_static_mel(2)
    y_hybrid = run(m, x, dev)
    window, fb = load_params("P0")
    ref = mel_oracle.mel_forward(x, fb=fb, window=window, dtype=np.float64)
    assert parity_error(y_direct, ref, True) < TARGET and parity_error(y_hybrid, ref, True) < TARGET
    assert np.abs(y_direct - y_hybrid).max() < 2e-6
    for mode in (1, 2):
        m.set_static_mel(mode)
        assert np.all(run(m, np.zeros((1, 4096), np.float32), dev) == 0.0)


@pytest.mark.parametrize("offset", [0, 1, 2, 3])
def test_unaligned_rows_use_aligned_down_tma(mods, dev, offset):
    """Rows starting 0..3 elements off a 16-byte boundary: interior tiles stay on the TMA path via
    the aligned-down copy (+delta); results must equal the cp.async path bit for bit."""
    m = mods["P0"]
    base = torch.from_numpy(signals.noise(1, 3 * 70000 + 8, 21)[0]).to(dev)
    x = base[offset:offset + 3 * 70000].view(3, 70000)     # row stride 70000 (16-byte multiple), shifted base
    m.set_bulk_copy(True)
    y_tma = m(x)
    m.set_bulk_copy(False)
    y_cp = m(x)
    m.set_bulk_copy(True)
    torch.cuda.synchronize()
    assert torch.equal(y_tma, y_cp)
    assert torch.equal(y_tma, m(x.clone()))                 # an aligned copy of the same data
    # odd row stride: every row has a different misalignment
    big = torch.from_numpy(signals.noise(1, 5 * 65537 + 8, 22)[0]).to(dev)
    xs = big[1:1 + 5 * 65537].as_strided((5, 65536), (65537, 1))
    assert torch.equal(m(xs), m(xs.contiguous()))


def test_kernel_variants_gather_unaligned(mods, dev):
    m = mods["P0"]
    song = torch.from_numpy(signals.noise(1, 700001, 5)[0]).to(dev)
    outs = {}
    for variant in ("barrier", "warp", "ws"):
        m.set_kernel_variant(variant)
        if variant == "ws":                     # pair tables: bit-identical to the other schedules
            m.set_static_mel(False)
        outs[variant] = m.forward_gather(song[1:], 3, 52415, 9, 262144).cpu().numpy()
        m.set_static_mel(True)
    outs["ws_default"] = m.forward_gather(song[1:], 3, 52415, 9, 262144).cpu().numpy()   # direct mel stage
    m.set_kernel_variant(DEFAULT_VARIANT)
    assert np.array_equal(outs["warp"], outs["barrier"])
    assert np.array_equal(outs["ws"], outs["barrier"])
    assert np.abs(outs["ws_default"] - outs["barrier"]).max() <= 2e-6
    window, fb = load_params("P0")
    seq = np.stack([np.pad(song[1:].cpu().numpy(), (0, 600000))[3 + w * 52415: 3 + w * 52415 + 262144] for w in range(9)])
    ref = mel_oracle.mel_forward(seq, fb=fb, window=window, dtype=np.float64)
    assert parity_error(outs["warp"], ref, True) < TARGET


@pytest.mark.parametrize("n,offset", [(700001, 0), (700001, 1), (12345, 3), (7, 0), (524160 * 3, 5)])
def test_peak_scale_of_a_resident_int16_song(mods, dev, n, offset):
    """bhmel_peak_scale_pcm16 == float32(1) / max|pcm| as NumPy computes the reference's factor
    (data_utils.py:94-96), for aligned and unaligned device pointers, including -32768."""
    m = mods["P0"]
    rng = np.random.default_rng(n + offset)
    pcm = rng.integers(-20000, 20000, size=n + offset, dtype=np.int16)
    pcm[offset + int(rng.integers(n))] = -32768 if n % 2 else 23456
    t = torch.from_numpy(pcm).to(dev)[offset:]
    scale = m.peak_scale(t).cpu().numpy()
    want = np.float32(1.0) / np.max(np.abs(pcm[offset:].astype(np.float32)))
    assert scale.dtype == np.float32 and scale[0] == want
    zeros = torch.zeros(1000, dtype=torch.int16, device=dev)
    assert np.isinf(m.peak_scale(zeros).cpu().numpy()[0])          # the reference divides by zero here too


def test_peak_scale_is_reentrant_across_threads_and_streams_of_one_handle(mods, dev):
    """two host threads share ONE module (one handle) and reduce different songs on their own
    streams; the reduction lives in each call's own output word, so neither disturbs the other."""
    import threading
    m = mods["P0"]
    rng = np.random.default_rng(77)
    songs, wants = [], []
    for peak in (1234, 31000):
        pcm = rng.integers(-peak, peak, size=3_000_001, dtype=np.int16)
        pcm[int(rng.integers(len(pcm)))] = peak
        songs.append(torch.from_numpy(pcm).to(dev))
        wants.append(np.float32(1.0) / np.float32(peak))
    m.peak_scale(songs[0])
    torch.cuda.synchronize()
    errors = []

    def worker(song, want):
        try:
            s = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(s):
                got = [m.peak_scale(song) for _ in range(200)]
            s.synchronize()
            bad = sum(float(g.cpu()[0]) != float(want) for g in got)
            if bad:
                errors.append(f"{bad} wrong scales")
        except Exception as e:   # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=a) for a in zip(songs, wants)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert errors == []


def test_gather_from_a_resident_int16_song_matches_the_reference_loader(mods, dev):
    """int16 song on the device -> on-device peak -> fused convert + gather  ==  the reference's host
    path: astype(float32), samples *= 1/max|samples| (data_utils.py:94-96), segment, forward."""
    m = mods["P0"]
    rng = np.random.default_rng(11)
    pcm = (signals.music(700001, seed=4) * 9000).astype(np.int16)
    pcm[1234] = 30000
    samples = pcm.astype(np.float32)
    samples *= np.float32(1.0) / np.max(np.abs(samples))
    song_i16 = torch.from_numpy(pcm).to(dev)
    song_f32 = torch.from_numpy(samples).to(dev)
    for first, stride, W, wlen in ((0, 52415, 9, 262144), (3, 131040, 5, 524160), (17, 999, 40, 4000)):
        y_ref = m.forward_gather(song_f32, first, stride, W, wlen)
        y_auto = m.forward_gather(song_i16, first, stride, W, wlen, normalize=True)
        y_given = m.forward_gather(song_i16, first, stride, W, wlen, normalize=m.peak_scale(song_i16))
        assert torch.equal(y_auto, y_ref) and torch.equal(y_given, y_ref)
    y_raw = m.forward_gather(song_i16, 0, 52415, 3, 262144)                  # normalize=False: scale 1.0
    assert torch.equal(y_raw, m.forward_gather(song_i16.to(torch.float32), 0, 52415, 3, 262144))
    with pytest.raises(RuntimeError):
        m.forward_gather(song_f32, 0, 52415, 3, 262144, normalize=True)      # float songs carry no PCM scale


@pytest.mark.parametrize("bulk", [True, False])
def test_bulk_and_cp_async_staging_agree(mods, dev, bulk):
    m = mods["P0"]
    x = signals.noise(3, 40000, 42)
    m.set_bulk_copy(True)
    y_bulk = run(m, x, dev)
    m.set_bulk_copy(bulk)
    y = run(m, x, dev)
    m.set_bulk_copy(True)
    assert np.array_equal(y, y_bulk)
    window, fb = load_params("P0")
    ref = mel_oracle.mel_forward(x, fb=fb, window=window, dtype=np.float64)
    assert parity_error(y, ref, True) < TARGET


@pytest.mark.parametrize("N", [513, 640, 1151, 4096, 4097, 12345, 131071])
@pytest.mark.parametrize("pset", ["P0", "P0C"])
def test_ragged_lengths(mods, dev, N, pset):
    log, n_mels, f_min, f_max, pad = PSET_ARGS[pset]
    window, fb = load_params(pset)
    x = signals.noise(3, N, 5000 + N)
    y = run(mods[pset], x, dev)
    ref = mel_oracle.mel_forward(x, fb=fb, window=window, pad_mode=pad, dtype=np.float64)
    assert y.shape == ref.shape
    assert parity_error(y, ref, log) < TARGET


def test_short_inputs_constant_pad(mods, dev):
    window, fb = load_params("P0C")
    for N in (1, 100, 128, 512):
        x = signals.noise(2, N, N)
        y = run(mods["P0C"], x, dev)
        ref = mel_oracle.mel_forward(x, fb=fb, window=window, pad_mode="constant", dtype=np.float64)
        assert parity_error(y, ref, True) < TARGET


def test_strided_and_unaligned_rows(mods, dev):
    """Row stride > N, and a base pointer that is not 16-byte aligned (forces the cp.async path)."""
    m = mods["P0"]
    window, fb = load_params("P0")
    big = torch.from_numpy(signals.noise(4, 9001, 3)).to(dev)
    view = big[:, 1:8193]                       # stride 9001, offset 1 element
    y = m(view)
    torch.cuda.synchronize()
    ref = mel_oracle.mel_forward(view.cpu().numpy(), fb=fb, window=window, dtype=np.float64)
    assert parity_error(y.cpu().numpy(), ref, True) < TARGET
    yt = m(big.t().contiguous().t()[:, :4096])  # non-unit inner stride -> made contiguous by the wrapper
    torch.cuda.synchronize()
    ref = mel_oracle.mel_forward(big[:, :4096].cpu().numpy(), fb=fb, window=window, dtype=np.float64)
    assert parity_error(yt.cpu().numpy(), ref, True) < TARGET


def test_full_context_batch_c2(mods, dev):
    """C2: a 3-min music-like song, 46 overlapped windows [46, 524160].  Checked
    against the torch CPU port on every window, all frames."""
    m = mods["P0"]
    window, fb = load_params("P0")
    song = signals.music(2_880_000, seed=1)
    w, s = mel_oracle.segment_params()
    seq = mel_oracle.segment(song, w, s)
    assert seq.shape == (46, WINDOW)
    y = run(m, seq, dev)
    assert y.shape == (46, 4096, 80)
    port = torch_port.TorchPortMel()
    port.fb.copy_(torch.from_numpy(fb))
    port.window.copy_(torch.from_numpy(window))
    ref = port(torch.from_numpy(seq)).numpy()
    assert parity_error(y, ref, True) < TARGET
    # fused segmentation must give bit-identical frames without the [46, 524160] batch
    yg = m.forward_gather(torch.from_numpy(song).to(dev), 0, s, 46, w)
    torch.cuda.synchronize()
    assert np.array_equal(yg.cpu().numpy(), y)
    # parallel (non-overlapped) segmentation: 6 windows
    seq6 = mel_oracle.segment(song, w, w)
    y6 = m.forward_gather(torch.from_numpy(song).to(dev), 0, w, 6, w)
    torch.cuda.synchronize()
    assert np.array_equal(y6.cpu().numpy(), run(m, seq6, dev))


def test_forward_gather_on_a_trimmed_plan_equals_the_reference_trimmed_windows(mods, dev):
    """a10: start_time / end_time trimming (preprocessor.py:75-90).  forward_gather on the trimmed plan
    must equal the frontend applied to the windows the reference keeps, bit for bit."""
    from beatheritage_b200 import segment as seg
    m = mods["P0"]
    song = signals.music(1_500_000, seed=4)
    for (st, en, bp, ep) in [(20_000, 60_000, 0, 0), (None, 30_000, 4096, 100), (70_000, None, 0, 0), (1e9, None, 0, 0)]:
        plan = seg.segment_plan(len(song), 1024, 128, 0.5, 0.4, False, 16000, st, en, bp, ep)
        seq = mel_oracle.segment(song, plan.window_len, plan.stride, bp, ep)
        first, kept, times = mel_oracle.segment_times_and_trim(len(seq), plan.window_len, plan.stride, 16000, 0.5, 0.4, st, en)
        assert (first, kept) == (plan.first_window, plan.n_windows) and times.tolist() == list(plan.sequence_times)
        want = run(m, seq[first:first + kept], dev)
        resident = torch.from_numpy(np.pad(song, [bp, ep])).to(dev)        # segment()'s begin / end padding
        got = m.forward_gather(resident, plan.first_offset, plan.stride, plan.n_windows, plan.window_len)
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy(), want)


def test_forward_host_matches_forward(mods, dev):
    m = mods["P0"]
    x = torch.from_numpy(signals.noise(37, 65536, 99))
    y_dev = m(x.to(dev)).cpu()
    xp = x.pin_memory()
    y_host = m.forward_host(xp)
    assert torch.equal(y_host, y_dev)
    y_pageable = m.forward_host(x)            # pageable memory also works (slower)
    assert torch.equal(y_pageable, y_dev)


def test_forward_host_pcm16_and_bf16(mods, dev):
    """N2 + N1 on the host entry: int16 PCM rows scaled on the device (the reference's
    astype(float32) * 1/peak, data_utils.py:95-97) and a bfloat16 result."""
    m = mods["P0"]
    rng = np.random.default_rng(5)
    pcm = rng.integers(-20000, 20000, size=(9, 70000), dtype=np.int16)
    pcm[3] = 0                                                      # a silent row
    peak = np.abs(pcm.astype(np.float32)).max(axis=1)
    scales = np.where(peak > 0, np.float32(1.0) / np.maximum(peak, 1), np.float32(0)).astype(np.float32)
    x_ref = pcm.astype(np.float32) * scales[:, None]                # what load_audio_file produces
    y_ref = m(torch.from_numpy(x_ref).to(dev)).cpu()
    y = m.forward_host(torch.from_numpy(pcm), scales=torch.from_numpy(scales))
    assert torch.equal(y, y_ref)
    y16 = m.forward_host(torch.from_numpy(pcm).pin_memory(), scales=torch.from_numpy(scales), out_dtype=torch.bfloat16)
    assert y16.dtype == torch.bfloat16 and torch.equal(y16, y_ref.to(torch.bfloat16))
    yf16 = m.forward_host(torch.from_numpy(x_ref), out_dtype=torch.bfloat16)
    assert torch.equal(yf16, y_ref.to(torch.bfloat16))
    window, fb = load_params("P0")
    ref = mel_oracle.mel_forward(x_ref[:2], fb=fb, window=window, dtype=np.float64)
    assert parity_error(y[:2].numpy(), ref, True) < TARGET


def test_forward_host_tapered_chunk_plan(mods, dev):
    """Batches large enough for the host entry's tapered chunk schedule (1/8, 1/4, 1/2 chunks at both ends,
    bhmel_host_chunk_plan): every row lands where the device-resident forward puts it, f32 and int16 ingest."""
    from beatheritage_b200 import MelSpectrogram
    m = mods["P0"]
    B, N = 53, 524161
    plan = MelSpectrogram.host_chunk_plan(B, N)
    assert plan[:3] == [1, 2, 4] and plan[-3:] == [4, 2, 1] and sum(plan) == B
    g = torch.Generator().manual_seed(17)
    x = (torch.rand(B, N, generator=g) * 2 - 1).pin_memory()
    y_dev = m(x.to(dev)).cpu()
    assert torch.equal(m.forward_host(x), y_dev)
    del x, y_dev
    B, N = 27, 4_000_003
    plan = MelSpectrogram.host_chunk_plan(B, N, pcm16=True)
    assert plan[0] < max(plan) and plan[-1] < max(plan) and sum(plan) == B
    pcm = torch.randint(-30000, 30000, (B, N), dtype=torch.int16, generator=g)
    scales = torch.full((B,), 1.0 / 30000.0)
    y_ref = m((pcm.to(torch.float32) * scales[:, None]).to(dev)).to(torch.bfloat16).cpu()
    y = m.forward_host(pcm.pin_memory(), scales=scales, out_dtype=torch.bfloat16)
    assert torch.equal(y, y_ref)


def test_properties_at_full_size(mods, dev):
    """Size-independent properties on a [64, 524160] batch (no oracle needed at this size):
    batch independence, determinism, power scaling (non-log set), zero rows, time shift."""
    m = mods["P0"]
    g = torch.Generator(device=dev).manual_seed(1234)
    x = torch.rand(64, WINDOW, device=dev, generator=g) * 2 - 1
    x[5].zero_()
    y = m(x)
    y2 = m(x)
    assert torch.equal(y, y2)                                   # deterministic
    assert torch.all(y[5] == 0)                                 # silent window -> exact zeros
    sub = m(x[10:13].clone())
    assert torch.equal(sub, y[10:13])                           # rows are independent
    # shifting a row by one hop shifts the interior frames by one
    xs = torch.roll(x[7:8], shifts=-128, dims=1)
    ys = m(xs)
    assert torch.allclose(ys[0, 8:4000], y[7, 9:4001], atol=2e-5)
    # power scaling on the non-log parameter set: mel(a x) = a^2 mel(x)
    p1 = mods["P1"]
    xa = x[:4, :65536].contiguous()
    ya, yb = p1(xa), p1(xa * 0.5)
    assert torch.allclose(yb * 4, ya, rtol=1e-5, atol=1e-6)
    # spot-check 3 windows of the big batch against the CPU port
    window, fb = load_params("P0")
    port = torch_port.TorchPortMel()
    port.fb.copy_(torch.from_numpy(fb))
    port.window.copy_(torch.from_numpy(window))
    idx = [0, 31, 63]
    ref = port(x[idx].cpu()).numpy()
    assert parity_error(y[idx].cpu().numpy(), ref, True) < TARGET


@pytest.mark.parametrize("variant", ["ws", "barrier", "warp"])
def test_forward_into_typed_pitched_output(mods, dev, variant):
    """N1: bf16 / wider-record output equals forward() followed by .to(dtype) + cat, bit for bit."""
    m = mods["P0"]
    m.set_kernel_variant(variant)
    x = torch.from_numpy(signals.noise(3, 20000, 12)).to(dev)
    y = m(x)                                                   # [3, 157, 80] f32
    T = y.shape[1]
    cond = torch.randn(3, T, 384, device=dev)
    # encoder input [B, T, 464] in bf16, mel channels first (modeling_mapperatorinator.py:352, 369-370)
    ref = torch.cat([y.to(torch.bfloat16), cond.to(torch.bfloat16)], dim=-1)
    buf = torch.empty(3, T, 464, device=dev, dtype=torch.bfloat16)
    buf[:, :, 80:] = cond.to(torch.bfloat16)
    m.forward_into(x, buf, channel_offset=0)
    assert torch.equal(buf, ref)
    # float32, mel channels in the middle of the record
    buf32 = torch.zeros(3, T, 100, device=dev)
    m.forward_into(x, buf32, channel_offset=7)
    assert torch.equal(buf32[:, :, 7:87], y) and torch.all(buf32[:, :, :7] == 0) and torch.all(buf32[:, :, 87:] == 0)
    # dense bf16 == .to(bfloat16)
    dense = torch.empty(3, T, 80, device=dev, dtype=torch.bfloat16)
    assert torch.equal(m.forward_into(x, dense), y.to(torch.bfloat16))
    m.set_kernel_variant(DEFAULT_VARIANT)
    with pytest.raises(RuntimeError):
        m.forward_into(x, torch.empty(3, T, 60, device=dev))


@pytest.mark.parametrize("channels_first", [False, True])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("pset,shape,dims", [("P0", (3, 40000), (128, 128, 64, 64)), ("P0", (2, 524160), (384,)),
                                             ("P128", (2, 9999), (5, 3)), ("P0", (1, 4097), ())])
def test_encoder_input_assembly_matches_the_reference_ops(mods, dev, pset, shape, dims, dtype, channels_first):
    """bhmel_forward_encoder_input == frames.to(dtype) + expand + concatenate (+ swapaxes) of
    modeling_mapperatorinator.py:351-352, 368-376, bit for bit, for both layouts and dtypes,
    vectorisable (80 + 384) and odd (128 + 5 + 3) channel counts, and no conditioning at all."""
    m = mods[pset]
    x = torch.from_numpy(signals.noise(shape[0], shape[1], 3 + shape[1])).to(dev)
    g = torch.Generator(device="cpu").manual_seed(5)
    conds = [torch.randn(shape[0], d, generator=g).to(dev) for d in dims]
    frames = m(x).to(dtype)
    T = frames.shape[1]
    want = torch.concatenate([frames] + [c.to(dtype).unsqueeze(1).expand(-1, T, -1) for c in conds], dim=-1)
    if channels_first:
        want = torch.swapaxes(want, 1, 2).contiguous()
    got = m.forward_encoder_input(x, conds, dtype=dtype, channels_first=channels_first)
    assert got.shape == want.shape and got.dtype == dtype and got.is_contiguous()
    assert torch.equal(got, want)


def test_handle_owned_scratch_paths_of_the_c_abi(mods, dev):
    """The Python mirror hands the library caller-owned scratch buffers; the C entries also work with
    scratch = NULL (handle-owned, grown on demand) and give the same bits."""
    import ctypes
    from beatheritage_b200 import _lib
    m = mods["P0"]
    lib = _lib.lib()
    x = torch.from_numpy(signals.noise(2, 30000, 8)).to(dev)
    cond = torch.randn(2, 24, device=dev).to(torch.bfloat16)
    want = m.forward_encoder_input(x, [cond], channels_first=True)
    h = m._handle_for(dev)
    for _ in range(2):                                              # second call reuses the grown buffer
        got = torch.empty_like(want)
        desc = _lib.BhmelEncoderInputDesc(got.data_ptr(), _lib.OUT_BF16, _lib.LAYOUT_BCT, cond.data_ptr(), 24, None)
        _lib.check(lib.bhmel_forward_encoder_input(h, x.data_ptr(), 2, 30000, 30000, ctypes.byref(desc),
                                                   torch.cuda.current_stream(dev).cuda_stream))
        assert torch.equal(got, want)
    pcm = (signals.music(200001, seed=2) * 8000).astype(np.int16)
    song = torch.from_numpy(pcm).to(dev)
    want_g = m.forward_gather(song, 5, 20011, 7, 65536, normalize=True)
    scale = m.peak_scale(song)
    got_g = torch.empty_like(want_g)
    _lib.check(lib.bhmel_forward_gather_pcm16(h, song.data_ptr(), song.numel(), scale.data_ptr(), 5, 20011, 7, 65536,
                                              got_g.data_ptr(), None, torch.cuda.current_stream(dev).cuda_stream))
    assert torch.equal(got_g, want_g)


def test_state_dict_reload_rebuilds_device_tables(dev):
    from beatheritage_b200 import MelSpectrogram
    m = MelSpectrogram("torchaudio", False, 16000, 1024, 80, 128, 20, 8000, "reflect").to(dev)
    x = torch.from_numpy(signals.noise(1, 8192, 1)).to(dev)
    y1 = m(x)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    sd["transform.mel_scale.fb"] *= 3.0
    m.load_state_dict(sd, strict=True)
    y2 = m(x)
    assert torch.allclose(y2, 3.0 * y1, rtol=1e-6)


def test_dense_filterbank_from_state_dict(dev):
    from beatheritage_b200 import MelSpectrogram
    m = MelSpectrogram("torchaudio", False, 16000, 1024, 80, 128, 20, 8000, "reflect").to(dev)
    rng = np.random.default_rng(3)
    fb = rng.random((513, 80), dtype=np.float32)
    fb[:, 3] = 0.0
    with torch.no_grad():
        m.transform.mel_scale.fb.copy_(torch.from_numpy(fb))
    x = signals.noise(1, 5000, 9)
    y = run(m, x, dev)
    window, _ = load_params("P0")
    ref = mel_oracle.mel_forward(x, fb=fb, window=window, log_scale=False, dtype=np.float64)
    assert np.all(y[..., 3] == 0)
    assert np.abs(y - ref).max() / ref.max() < 1e-5


@pytest.mark.parametrize("args,shape", [
    (("nnAudio", False, 16000, 1024, 388, 128, 0, 8000, "constant"), (2, 40000)),    # configs/model/default.yaml:19-27
    (("nnAudio", True, 16000, 1024, 128, 128, 20, 8000, "reflect"), (3, 9999)),
    (("nnAudio", False, 16000, 1024, 512, 128, 0, 8000, "constant"), (1, 524160)),   # 47 all-zero filters
])
def test_nnaudio_published_arithmetic(dev, monkeypatch, args, shape):
    """N4, parity UNPINNED against nnAudio itself: the module in nnAudio mode against
    the restatement of nnAudio's published conv-STFT + Slaney basis (oracle/nnaudio_oracle.py), and a
    state-dict round trip in nnAudio's buffer layout."""
    from beatheritage_b200 import MelSpectrogram
    from oracle import nnaudio_oracle
    monkeypatch.setattr(MelSpectrogram, "nnaudio_arithmetic", "published")
    _, log, _, _, n_mels, _, f_min, f_max, pad = args
    m = MelSpectrogram(*args).to(dev)
    x = signals.noise(shape[0], shape[1], 31 + n_mels)
    y = run(m, x, dev)
    x_ref, y_cmp = (x, y) if shape[1] < 100000 else (x[:, :40000], y[:, :300])    # conv form is O(T * 513 * 1024)
    ref = nnaudio_oracle.mel_forward(x_ref, n_mels=n_mels, f_min=f_min, f_max=f_max, pad_mode=pad, log_scale=log)
    ref = ref[:, :y_cmp.shape[1]]
    assert y.shape == (shape[0], shape[1] // 128 + 1, n_mels)
    assert parity_error(y_cmp, ref, log) < TARGET
    assert np.abs(y_cmp - ref).max() <= 1e-4 * np.abs(ref).max()                # SURVEY.md 8c bar for non-log sets
    dead = ~nnaudio_oracle.mel_basis(16000, 1024, n_mels, f_min, f_max).any(axis=1)
    assert np.all(y[..., dead] == 0)
    m2 = MelSpectrogram(*args).to(dev)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    sd["transform.mel_basis"] *= 2.0
    m2.load_state_dict(sd, strict=True)
    y2 = run(m2, x, dev)
    if log:
        assert np.allclose(np.expm1(y2.astype(np.float64)), 2.0 * np.expm1(y.astype(np.float64)), rtol=5e-5, atol=5e-5)
    else:
        assert np.array_equal(y2, 2.0 * y)


def test_runs_on_a_side_stream_and_under_autocast(mods, dev):
    m = mods["P0"]
    x = torch.from_numpy(signals.noise(2, 30000, 8)).to(dev)
    ref = m(x)
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s), torch.autocast("cuda", dtype=torch.bfloat16), torch.no_grad():
        y = m(x)
        assert y.dtype == torch.float32            # the frontend ignores autocast (always fp32)
    s.synchronize()
    assert torch.equal(y, ref)
    assert torch.equal(m(x.double()), ref)         # non-fp32 input is converted


def test_concurrent_threads_and_streams(mods, dev):
    """The library is re-entrant: two host threads, each on its own stream and with its own module
    (different filterbanks), interleave launches and get the same results as when run alone
    (the reference calls the model from a non-main batch thread, inference/server.py:176)."""
    import threading
    from beatheritage_b200 import MelSpectrogram
    m_a = mods["P0"]
    m_b = MelSpectrogram("torchaudio", False, 16000, 1024, 128, 128, 0, 8000, "constant").to(dev)
    xa = torch.from_numpy(signals.noise(4, 50000, 1)).to(dev)
    xb = torch.from_numpy(signals.noise(5, 30001, 2)).to(dev)
    ref_a, ref_b = m_a(xa).clone(), m_b(xb).clone()
    torch.cuda.synchronize()
    errors = []

    def worker(m, x, ref):
        try:
            s = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(s):
                for _ in range(30):
                    y = m(x)
                    if not torch.equal(y, ref):
                        errors.append("mismatch")
            s.synchronize()
        except Exception as e:   # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=a) for a in ((m_a, xa, ref_a), (m_b, xb, ref_b))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert errors == []


def test_many_windows_and_many_mels(dev):
    """Upper ends of the shape space: 1 090 windows of a short context in one launch, and a
    1 024-filter bank (the maximum the tables support; served by the barrier kernel when the
    warp-specialised kernel's descriptor table is too small is NOT needed -- all variants take it)."""
    from beatheritage_b200 import MelSpectrogram
    m = MelSpectrogram("torchaudio", True, 16000, 1024, 1024, 128, 0, 8000, "reflect").to(dev)
    x = signals.noise(2, 6000, 3)
    ys = {}
    for variant in ("ws", "barrier", "warp"):
        m.set_kernel_variant(variant)
        ys[variant] = run(m, x, dev)
    assert np.array_equal(ys["ws"], ys["barrier"]) and np.array_equal(ys["warp"], ys["barrier"])
    ref = mel_oracle.mel_forward(x, fb=m.transform.mel_scale.fb.cpu().numpy(),
                                 window=m.transform.spectrogram.window.cpu().numpy(), dtype=np.float64)
    assert parity_error(ys["ws"], ref, True) < TARGET
    m0 = MelSpectrogram("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect").to(dev)
    xw = torch.from_numpy(signals.noise(1090, 8192, 4)).to(dev)
    y = m0(xw)
    assert torch.equal(y[777:778], m0(xw[777:778].clone()))


def test_c_abi_errors_on_device(dev):
    from beatheritage_b200 import _lib
    lib = _lib.lib()
    prm = _lib.BhmelParams(16000, 1024, 128, 80, 20.0, 8000.0, _lib.PAD_REFLECT, 1, None, None)
    h = ctypes.c_void_p()
    _lib.check(lib.bhmel_create(ctypes.byref(prm), ctypes.byref(h)))
    x = torch.zeros(2, 512, device=dev)
    y = torch.zeros(2, 5, 80, device=dev)
    assert lib.bhmel_forward(h, x.data_ptr(), 2, 512, 512, y.data_ptr(), None) == _lib.ESHAPE
    assert b"reflect" in lib.bhmel_last_error()
    assert lib.bhmel_forward(h, x.data_ptr(), 0, 4096, 4096, y.data_ptr(), None) == _lib.ESHAPE
    assert lib.bhmel_forward(h, None, 2, 4096, 4096, y.data_ptr(), None) == _lib.EINVAL
    assert lib.bhmel_forward(h, x.data_ptr(), 2, 4096, 100, y.data_ptr(), None) == _lib.EINVAL
    # the library's own C++ filterbank/window builders (no override) stay inside the bar
    xs = torch.from_numpy(signals.noise(1, 6000, 5)).to(dev)
    ys = torch.empty(1, 47, 80, device=dev)
    _lib.check(lib.bhmel_forward(h, xs.data_ptr(), 1, 6000, 6000, ys.data_ptr(), None))
    torch.cuda.synchronize()
    ref = mel_oracle.mel_forward(xs.cpu().numpy(), dtype=np.float64)
    assert parity_error(ys.cpu().numpy(), ref, True) < 1e-4
    lib.bhmel_destroy(h)


def test_fuzz_three_schedules(dev):
    """A few seconds of tools/fuzz_variants.py: random shapes / strides / pad modes / filterbanks,
    module and gather mode; the three kernel schedules must agree bit for bit."""
    import os
    import subprocess
    import sys
    from tests.conftest import ROOT
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_variants.py"), "--seconds", "6", "--seed", "3"],
                         capture_output=True, text=True, timeout=200)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "fuzz ok" in res.stdout


def test_plain_c_host_program(dev, tmp_path):
    """examples/c_abi_demo.c: a C program (no Python, no torch) drives the library through the C ABI."""
    import os
    import shutil
    import subprocess
    from tests.conftest import ROOT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.join(ROOT, "beatheritage_b200")
    subprocess.run([nvcc, "-x", "cu", os.path.join(ROOT, "examples", "c_abi_demo.c"), "-I", os.path.join(ROOT, "include"),
                    "-L", libdir, "-lbhmel", "-Xlinker", f"-rpath={libdir}", "-o", exe], check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "frames per window: 513" in res.stdout


def test_torch_compile_does_not_graph_break(mods, dev):
    m = mods["P0"]

    class Wrap(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.spectrogram = m

        def forward(self, frames):
            return self.spectrogram(frames) * 2.0 + 1.0

    w = Wrap()
    x = torch.from_numpy(signals.noise(2, 16384, 4)).to(dev)
    try:
        cw = torch.compile(w, fullgraph=True)
        out = cw(x)
    except Exception as e:   # inductor needs a working host compiler/triton; not the subject here
        pytest.skip(f"torch.compile unavailable in this environment: {type(e).__name__}")
    assert torch.allclose(out, w(x), atol=1e-6)


def test_programmatic_dependent_launch_keeps_stream_order(mods, dev):
    """BHMEL_OPT_PDL (default on): the warp-specialised kernel may become resident while the previous kernel
    in the stream drains, but reads no sample and writes no output before that kernel has completed.  Chains
    of adjacent launches with read-after-write (the next input is made from this output by the previous
    launch's own output buffer), write-after-write (one output buffer) and write-after-read hazards give the
    same bits with and without it."""
    m = mods["P0"]
    B, N = 3, 66000
    T = N // 128 + 1
    x0 = torch.from_numpy(signals.noise(B, N, 4242)).to(dev)
    out = torch.empty(B, T, 80, device=dev)

    def chain():
        x = x0.clone()
        got = []
        for i in range(6):
            m.forward_into(x, out)                                   # same output buffer every time
            got.append(out.clone())
            # adjacent launch reads what the previous one wrote: rebuild an input from the output in place
            x.view(-1)[: out.numel()].copy_(out.view(-1)).mul_(0.1).sub_(0.3)
            m.forward_into(x, out)
            got.append(out.clone())
        # and with nothing between the launches at all
        for i in range(6):
            m.forward_into(x0 if i % 2 == 0 else x, out)
        got.append(out.clone())
        torch.cuda.synchronize()
        return got

    with_pdl = chain()
    m.set_pdl(False)
    try:
        without = chain()
    finally:
        m.set_pdl(True)
    for a, b in zip(with_pdl, without):
        assert torch.equal(a, b)


def test_non_finite_and_extreme_samples(mods, dev):
    """What a NaN / Inf / huge sample does (the reference just propagates IEEE arithmetic through stft, the
    filterbank matmul and log1p): every frame whose 1024-sample span holds the bad sample is non-finite in the
    reference and here; frames that do not touch it keep their oracle values.  Two frames share one complex
    FFT here, so a bad frame's partner frame (the one just before or after the affected run) may turn
    non-finite as well -- at most one extra frame at each end, documented in INTEGRATION.md.  Samples so large
    that the power overflows give +inf on both sides (log1p(inf) = inf), never NaN."""
    import warnings
    m = mods["P0"]
    N = 40000
    T = N // 128 + 1
    window, fb = load_params("P0")
    for bad, pos in ((np.nan, 20000), (np.inf, 13), (-np.inf, 39990), (np.nan, 128 * 100 + 512)):
        x = signals.noise(2, N, 5).copy()
        x[1, pos] = bad
        with warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore")
            ref = mel_oracle.mel_forward(x, fb=fb, window=window, dtype=np.float64)
        got = run(m, x, dev)
        assert np.isfinite(got[0]).all() and parity_error(got[:1], ref[:1], True) < TARGET      # the clean row
        ref_bad = ~np.isfinite(ref[1]).all(axis=1)            # frames the reference poisons
        got_bad = ~np.isfinite(got[1]).all(axis=1)
        assert ref_bad.sum() >= 4
        assert np.all(got_bad[ref_bad]), "a frame the reference poisons is finite here"
        # every mel of a poisoned frame is non-finite on both sides
        assert (~np.isfinite(got[1][ref_bad])).all() and (~np.isfinite(ref[1][ref_bad])).all()
        extra = np.flatnonzero(got_bad & ~ref_bad)
        first, last = np.flatnonzero(ref_bad)[[0, -1]]
        assert len(extra) <= 2 and all(t in (first - 1, last + 1) for t in extra), extra
        ok = ~got_bad
        assert parity_error(got[1][ok][None], ref[1][ok][None], True) < TARGET
    # overflow: a mel value beyond float32 is +inf (log1p(inf) = inf), never NaN; values short of the limit stay finite
    x = signals.noise(1, N, 6).copy()
    x[0, 20000] = 3e19
    got = run(m, x, dev)
    with np.errstate(all="ignore"):
        lin64 = np.expm1(mel_oracle.mel_forward(x, fb=fb, window=window, dtype=np.float64))
    over = lin64[0] > 4 * float(np.finfo(np.float32).max)           # clearly past the float32 range
    under = lin64[0] < 0.25 * float(np.finfo(np.float32).max)
    assert over.sum() > 100 and not np.isnan(got).any()
    assert np.isposinf(got[0][over]).all()
    assert np.isfinite(got[0][under]).all()
    assert T == got.shape[1]


def test_one_call_larger_than_int32_element_counts(mods, dev):
    """Maximum sizes: ONE call over 7 000 model-context windows -- 3.67e9 input samples and 2.29e9 output
    values, both past 2^31 (the input also past 2^32 bytes by a wide margin) -- must index with 64 bits
    everywhere.  Rows on both sides of the 2^31-element boundaries (input row 4097, output row 6553), the
    first and the last row equal the same rows computed alone."""
    free, _ = torch.cuda.mem_get_info(dev)
    B, N = 7000, 524160
    need = B * N * 4 + B * 4096 * 80 * 4
    if free < need * 1.15:
        pytest.skip(f"needs {need / 2**30:.1f} GiB of device memory")
    m = mods["P0"]
    g = torch.Generator(device=dev).manual_seed(99)
    x = torch.empty(B, N, device=dev)
    for b0 in range(0, B, 500):                         # filled in slices: torch's RNG kernels index with 32 bits
        x[b0:b0 + 500].uniform_(-1, 1, generator=g)
    y = m(x)
    torch.cuda.synchronize()
    assert y.shape == (B, 4096, 80) and y.numel() > 2 ** 31 and x.numel() > 2 ** 31
    for r in (0, 1, 4095, 4096, 4097, 4098, 6552, 6553, 6554, B - 2, B - 1):
        alone = m(x[r:r + 1].clone())
        assert torch.equal(y[r], alone[0]), f"row {r}"
    assert bool(torch.isfinite(y[::97]).all())
    del x, y
    torch.cuda.empty_cache()
