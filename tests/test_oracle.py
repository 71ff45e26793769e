"""CPU: pins the oracle (oracle/mel_oracle.py, oracle/torch_port.py) to the golden fixtures that
tests/golden/make_golden.py produced by running the UNMODIFIED reference module, and -- when the
reference tree is present (build container) -- to the live reference."""
import os

import numpy as np
import pytest
import torch

from oracle import mel_oracle, ref_loader, torch_port
from tests.conftest import PSET_ARGS, golden_case_names, load_case, load_params, parity_error, regenerate_input

SMALL = [n for n in golden_case_names() if n not in ("music_2win", "zero_tail", "noise_N524161")]
BIG = ["music_2win", "zero_tail", "noise_N524161"]


@pytest.mark.parametrize("name", SMALL + BIG)
def test_fp64_oracle_matches_reference_fixtures(name):
    case, pset, frames = load_case(name)
    log, n_mels, f_min, f_max, pad = PSET_ARGS[pset]
    window, fb = load_params(pset)
    x = regenerate_input(case)
    y = mel_oracle.mel_forward(x, n_mels=n_mels, f_min=f_min, f_max=f_max, pad_mode=pad, log_scale=log,
                               dtype=np.float64, fb=fb, window=window)
    assert y.shape == tuple(case["shape"])
    if frames is not None:
        y = y[:, frames]
    if "y64" in case.files:   # the same module run in fp64: the oracle restates it to rounding error
        np.testing.assert_allclose(y, case["y64"], rtol=1e-10, atol=1e-10)
    # fp32 reference output vs fp64 oracle: the reference's own rounding noise, well under the 1e-3 bar
    assert parity_error(case["y"], y, log) < 2e-5


@pytest.mark.parametrize("name", SMALL)
def test_fp32_oracle_and_torch_port_match_fixtures(name):
    case, pset, frames = load_case(name)
    log, n_mels, f_min, f_max, pad = PSET_ARGS[pset]
    window, fb = load_params(pset)
    x = regenerate_input(case)
    y32 = mel_oracle.mel_forward(x, n_mels=n_mels, f_min=f_min, f_max=f_max, pad_mode=pad, log_scale=log,
                                 dtype=np.float32, fb=fb, window=window)
    assert parity_error(y32, case["y"], log) < 2e-5
    port = torch_port.TorchPortMel(log, 16000, 1024, n_mels, 128, f_min, f_max, pad)
    port.fb.copy_(torch.from_numpy(fb))
    port.window.copy_(torch.from_numpy(window))
    yp = port(torch.from_numpy(x)).numpy()
    assert yp.shape == tuple(case["shape"])
    assert parity_error(yp, case["y"], log) < 2e-5


@pytest.mark.parametrize("pset", sorted(PSET_ARGS))
def test_filterbank_and_window_restatement(pset):
    log, n_mels, f_min, f_max, pad = PSET_ARGS[pset]
    window, fb = load_params(pset)
    fb32 = mel_oracle.melscale_fbanks(513, float(f_min), float(f_max), n_mels, 16000, np.float32)
    fb64 = mel_oracle.melscale_fbanks(513, float(f_min), float(f_max), n_mels, 16000, np.float64)
    assert fb32.shape == fb.shape == (513, n_mels)
    # fp32 restatement: bit-exact for P0; elsewhere a few ulps of torch's fp32 pow leak through
    assert np.abs(fb32 - fb).max() < 1e-4
    assert np.abs(fb64 - fb).max() < 1e-4
    if pset == "P0":
        assert np.array_equal(fb32, fb)
        nz = fb != 0
        assert nz.sum() == 1003 and nz.sum(0).min() == 3 and nz.sum(0).max() == 33
    assert np.abs(mel_oracle.hann_window(1024) - window).max() < 2e-7


def test_reflect_needs_more_than_half_window():
    with pytest.raises(RuntimeError):
        mel_oracle.mel_forward(np.zeros((1, 512), np.float32))
    assert mel_oracle.mel_forward(np.zeros((1, 513), np.float32)).shape == (1, 5, 80)
    assert mel_oracle.mel_forward(np.zeros((1, 100), np.float32), pad_mode="constant").shape == (1, 1, 80)


def test_zero_input_gives_exact_zero():
    y = mel_oracle.mel_forward(np.zeros((2, 2048), np.float32))
    assert np.all(y == 0.0)


def test_segment_matches_reference_numbers():
    # 3-min song -> 46 windows at stride 52 415 (6 when parallel); 1 h -> 1 090 / 110
    w, s = mel_oracle.segment_params()
    assert (w, s) == (524160, 52415)
    assert mel_oracle.segment_params(parallel=True) == (524160, 524160)
    song = np.arange(2_880_000, dtype=np.float32)
    seq = mel_oracle.segment(song, w, s)
    assert seq.shape == (46, 524160)
    assert seq[3, 0] == 3 * 52415 and seq[-1, -1] == 0.0     # right padding is zeros
    assert mel_oracle.segment(song, w, w).shape == (6, 524160)
    assert mel_oracle.segment(np.zeros(1000, np.float32), w, s).shape == (1, 524160)


def _segment_cases():
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "segment_cases.json")) as f:
        return json.load(f)["cases"]


def test_segment_trimming_matches_the_reference_preprocessor():
    """sequence_times and the start_time / end_time trimming (preprocessor.py:72-90) against outputs of
    the unmodified reference Preprocessor.segment (tests/golden/make_segment_golden.py, 52 cases)."""
    cases = _segment_cases()
    assert len(cases) >= 50
    for c in cases:
        w, s = mel_oracle.segment_params(c["src_seq_len"], 128, c["lookback"], c["lookahead"], c["parallel"])
        assert (w, s) == (c["samples_per_sequence"], c["sequence_stride"])
        n_total = c["n_samples"] + c["begin_pad"] + c["end_pad"]
        padded = w if n_total < w else n_total + (-(n_total - w)) % s
        first, kept, times = mel_oracle.segment_times_and_trim((padded - w) // s + 1, w, s, 16000, c["lookback"],
                                                               c["lookahead"], c["start_time"], c["end_time"])
        assert kept == c["n_windows"] and times.dtype == np.int32 and times.tolist() == c["sequence_times"]
        for i, (g, lead) in enumerate(zip(c["starts"], c["leading_zeros"])):
            if g is not None:
                assert (first + i) * s == g, (c, i)
    # the restated windows themselves, trimmed, on a small case the oracle materialises
    c = next(c for c in cases if c["n_samples"] == 1_000_000 and c["src_seq_len"] == 1024)
    song = np.arange(1, c["n_samples"] + 1, dtype=np.float32)
    seq = mel_oracle.segment(song, c["samples_per_sequence"], c["sequence_stride"], c["begin_pad"], c["end_pad"])
    first, kept, _ = mel_oracle.segment_times_and_trim(len(seq), c["samples_per_sequence"], c["sequence_stride"], 16000,
                                                       c["lookback"], c["lookahead"], c["start_time"], c["end_time"])
    seq = seq[first:first + kept]
    assert len(seq) == c["n_windows"]
    for w_, g in zip(seq, c["starts"]):
        nz = np.flatnonzero(w_)
        assert c["begin_pad"] + int(w_[nz[0]]) - 1 - int(nz[0]) == g


def test_sequence_times_keep_the_float32_rounding_of_torch_arange():
    # 1-hour song: 1 090 windows, times beyond 2^20 ms lose their 1/16 ms fractions in float32
    import torch
    w, s = mel_oracle.segment_params()
    _, kept, times = mel_oracle.segment_times_and_trim(1090, w, s)
    ms = s * 1000 / 16000
    want = torch.arange(0, 1090 * ms, ms).to(torch.int32).numpy()     # the reference's expression, verbatim
    assert kept == 1090 and np.array_equal(times, want)
    assert not np.array_equal(times, np.floor(np.arange(1090) * ms).astype(np.int32))   # the rounding is real


def test_dataset_windows_match_reference_numbers():
    # 180 s song: 22 501 hop-frames -> 6 windows, the last with 2 026 real hop-frames
    song = np.ones(2_880_000, np.float32)
    win = mel_oracle.dataset_windows(song)
    assert win.shape == (6, 524160)
    # ... 2 026 hop-frames of which the last one is the extra all-zero hop _get_frames appends
    real = np.count_nonzero(win[-1].reshape(4095, 128).any(axis=1))
    assert real == 2025
    # an already hop-aligned song still gains one zero hop-frame (ors_dataset.py:256)
    assert mel_oracle.dataset_windows(np.ones(128 * 10, np.float32), src_seq_len=5).shape == (3, 512)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_live_reference_agrees_with_oracle_and_fixtures():
    Ref = ref_loader.load_reference_class()
    ref = Ref("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect")
    window, fb = load_params("P0")
    sd = ref.state_dict()
    assert list(sd) == ["transform.spectrogram.window", "transform.mel_scale.fb"]
    assert np.array_equal(sd["transform.mel_scale.fb"].numpy(), fb)
    rng = np.random.default_rng(11)
    x = (rng.random((3, 6000), dtype=np.float32) * 2 - 1)
    y_ref = ref(torch.from_numpy(x)).numpy()
    y = mel_oracle.mel_forward(x, fb=fb, window=window, dtype=np.float64)
    assert parity_error(y_ref, y, True) < 2e-5


# ---------------------------------------------------------------------------------------------
# N4: the nnAudio-arithmetic restatement.  Parity UNPINNED against nnAudio itself
# (see oracle/nnaudio_oracle.py); these tests hold it to the two things that can be checked here.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_mels,f_min", [(80, 0.0), (128, 20.0), (388, 0.0), (512, 0.0)])
def test_nnaudio_mel_basis_matches_torchaudios_slaney_filterbank(n_mels, f_min):
    import warnings

    import torchaudio
    from oracle import nnaudio_oracle
    mb = nnaudio_oracle.mel_basis(16000, 1024, n_mels, f_min, 8000.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                # torchaudio warns about all-zero filters (388 / 512 mels)
        ta = torchaudio.functional.melscale_fbanks(513, f_min, 8000.0, n_mels, 16000, norm="slaney",
                                                   mel_scale="slaney").numpy()
    assert mb.shape == (n_mels, 513) and mb.dtype == np.float32
    # torchaudio does the same construction in fp32 tensor arithmetic: a few fp32 ulps of the peak apart
    assert np.abs(mb.T - ta).max() <= 4e-5 * ta.max()
    assert np.array_equal(mb.T > 0, ta > 0) or np.abs(mb.T - ta)[(mb.T > 0) != (ta > 0)].max() < 1e-6


def test_nnaudio_conv_stft_equals_the_pinned_fft_form():
    from oracle import nnaudio_oracle
    from tests.golden import signals
    x = signals.noise(2, 6000, 5)
    mb = nnaudio_oracle.mel_basis(16000, 1024, 388, 0.0, 8000.0)
    for pad, log in (("constant", False), ("reflect", True)):
        conv = nnaudio_oracle.mel_forward(x, pad_mode=pad, log_scale=log)
        fft = mel_oracle.mel_forward(x, fb=mb.T, pad_mode=pad, log_scale=log, dtype=np.float64)
        assert conv.shape == fft.shape == (2, 47, 388)
        # fp32-rounded sin/cos kernels vs the exact DFT: ~1e-7 relative
        assert np.abs(conv - fft).max() <= 1e-6 * np.abs(fft).max()
    wsin, wcos, window = nnaudio_oracle.fourier_kernels()
    assert wsin.shape == wcos.shape == (513, 1, 1024)
    assert np.array_equal(wcos[0, 0], window) and not wsin[0].any()


def test_stem_gelu_formula_matches_torch_for_every_bf16_input():
    """N3: the conv stem's epilogue evaluates GELU with Abramowitz-Stegun 7.1.26 instead of erff
    (csrc/bhstem.cu::conv_gelu).  Its input is always a bf16 value, so the claim is checked exhaustively:
    the fp32 model of the formula, rounded to bf16, equals torch's fp32 erf GELU for every finite bf16
    input below 1e30 except in the tail x <= -3.5 (|diff| <= 4e-6).  The GPU side of the same check is
    tests/test_gpu_stem.py::test_gelu_is_checked_for_every_bf16_input."""
    from tests.test_gpu_stem import gelu_model
    bits = torch.arange(65536, dtype=torch.int32)
    x = (bits << 16).view(torch.float32)
    x = torch.where(torch.isfinite(x) & (x.abs() < 1e30), x, torch.zeros(()))
    want = torch.nn.functional.gelu(x).to(torch.bfloat16).float()
    with np.errstate(all="ignore"):
        model = torch.from_numpy(gelu_model(x.numpy())).to(torch.bfloat16).float()
    differs = model != want
    assert not bool((differs & (x > -3.5)).any())
    assert int(differs.sum()) <= 32 and float((model - want).abs().max()) <= 4e-6


@pytest.mark.parametrize("B,T,c_in,d,n_var", [(2, 37, 464, 96, 80), (1, 2, 16, 8, 8), (3, 5, 24, 16, 16)])
def test_split_conv1_algebra_equals_conv1_over_the_concatenated_input(B, T, c_in, d, n_var):
    """N1 + N3 (bhstem_forward_split): folding the time-constant conditioning channels into a per-window bias --
    with one tap dropped at each end of the window, where conv1 reads its zero padding -- is conv1 over the
    reference's concatenated input (modeling_mapperatorinator.py:368-370), exactly.  float64, CPU only."""
    from oracle import conv_stem_oracle as cso
    g = torch.Generator().manual_seed(B * 100 + T)
    w1 = torch.randn(d, c_in, 3, generator=g) * 0.05
    b1 = torch.randn(d, generator=g) * 0.1
    frames = (torch.randn(B, T, n_var, generator=g) * 1.5).to(torch.bfloat16)
    cond = (torch.randn(B, c_in - n_var, generator=g) * 1.5).to(torch.bfloat16)
    full = torch.cat([frames, cond.unsqueeze(1).expand(-1, T, -1)], dim=-1).double().swapaxes(1, 2)
    w = w1.to(torch.bfloat16).double()
    want = torch.nn.functional.conv1d(full, w, b1.to(torch.bfloat16).double(), padding=1).swapaxes(1, 2)
    got = cso.split_conv1_preactivation(frames, cond, w1, b1)
    assert got.shape == want.shape == (B, T, d)
    assert float((got - want).abs().max()) < 1e-11
    fb = cso.folded_bias(cond, w1, b1, n_var)
    assert fb.shape == (B, 3, d)
    if T > 2:      # interior frames really see all three taps, the edges one fewer
        assert float((fb[:, 0] - fb[:, 1]).abs().max()) > 1e-3 and float((fb[:, 0] - fb[:, 2]).abs().max()) > 1e-3
