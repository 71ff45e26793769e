"""CPU: the lane emulator (tests/emu/emu.cpp) runs the kernel's exact algorithm -- same generated
FFT passes, tables and index arithmetic -- on the host.  These tests prove the algorithm against
the reference fixtures and the fp64 oracle without a GPU; the -m gpu tests then only have to
prove the CUDA mechanics."""
import numpy as np
import pytest

from oracle import mel_oracle
from tests.conftest import PSET_ARGS, golden_case_names, load_case, load_params, parity_error, regenerate_input
from tests.golden import signals

TOL = 1e-3       # north-star bar: max abs log-mel error vs the reference
TARGET = 2e-5    # what an fp32 FFT actually achieves


@pytest.mark.parametrize("name", golden_case_names())
def test_emulator_matches_reference_fixtures(emu_lib, name):
    case, pset, frames = load_case(name)
    log, n_mels, f_min, f_max, pad = PSET_ARGS[pset]
    window, fb = load_params(pset)
    x = regenerate_input(case)
    y = emu_lib(x, n_mels, fb=fb, window=window, reflect=(pad == "reflect"), log=log)
    assert y.shape == tuple(case["shape"])
    if frames is not None:
        y = y[:, frames]
    err = parity_error(y, case["y"], log)
    assert err < TOL
    assert err < TARGET, f"{name}: {err}"


def test_emulator_internal_tables_match_torchaudio_buffers(emu_lib):
    """With fb/window left to the library's own C++ builders the result still sits inside the bar."""
    x = signals.noise(2, 6000, 5)
    window, fb = load_params("P0")
    y_ext = emu_lib(x, 80, fb=fb, window=window)
    y_int = emu_lib(x, 80)
    assert parity_error(y_int, y_ext, True) < 1e-4


@pytest.mark.parametrize("N", [513, 640, 700, 1151, 4096, 4097, 8191])
@pytest.mark.parametrize("pad", ["reflect", "constant"])
def test_emulator_ragged_lengths_vs_oracle(emu_lib, N, pad):
    x = signals.noise(3, N, 1000 + N)
    window, fb = load_params("P0")
    y = emu_lib(x, 80, fb=fb, window=window, reflect=(pad == "reflect"))
    ref = mel_oracle.mel_forward(x, fb=fb, window=window, pad_mode=pad, dtype=np.float64)
    assert y.shape == ref.shape == (3, N // 128 + 1, 80)
    assert parity_error(y, ref, True) < TARGET


def test_emulator_short_constant_inputs(emu_lib):
    window, fb = load_params("P0")
    for N in (1, 100, 128, 512):
        x = signals.noise(2, N, N)
        y = emu_lib(x, 80, fb=fb, window=window, reflect=False)
        ref = mel_oracle.mel_forward(x, fb=fb, window=window, pad_mode="constant", dtype=np.float64)
        assert parity_error(y, ref, True) < TARGET


def test_emulator_gather_equals_materialised_windows(emu_lib):
    """Fused segmentation (row r = song[first + r*stride ...], zeros past the end) must equal the
    frontend applied to Preprocessor.segment's materialised windows."""
    window, fb = load_params("P0")
    song = signals.noise(1, 9000, 77)[0]
    wlen, stride = 31 * 128, 397          # stride not hop-aligned, like 52 415
    seq = mel_oracle.segment(song, wlen, stride)
    y_mat = emu_lib(seq, 80, fb=fb, window=window)
    y_gat = emu_lib(song, 80, fb=fb, window=window, gather=(0, stride, seq.shape[0], wlen))
    assert np.array_equal(y_mat, y_gat)
    ref = mel_oracle.mel_forward(seq, fb=fb, window=window, dtype=np.float64)
    assert parity_error(y_gat, ref, True) < TARGET


def test_emulator_dense_filterbank(emu_lib):
    """Any [513, M] matrix must work (a loaded state dict may carry a non-triangular fb)."""
    rng = np.random.default_rng(3)
    fb = rng.random((513, 7), dtype=np.float32)
    fb[:, 3] = 0.0                 # an all-zero filter
    fb[:100, 5] = 0.0
    window, _ = load_params("P0")
    x = signals.noise(1, 3000, 9)
    y = emu_lib(x, 7, fb=np.ascontiguousarray(fb), window=window, log=False)
    ref = mel_oracle.mel_forward(x, fb=fb, window=window, log_scale=False, dtype=np.float64)
    assert np.all(y[..., 3] == 0)
    assert np.abs(y - ref).max() / ref.max() < 1e-5


@pytest.mark.parametrize("pset", sorted(PSET_ARGS))
def test_round_tables_equal_pair_tables(emu_lib, pset):
    """The two mel table layouts (paired bands for the barrier kernel, lane-interleaved rounds for
    the independent-warp kernel) must give bit-identical projections."""
    log, n_mels, f_min, f_max, pad = PSET_ARGS[pset]
    window, fb = load_params(pset)
    x = signals.noise(2, 5000, 31)
    a = emu_lib(x, n_mels, fb=fb, window=window, reflect=(pad == "reflect"), log=log)
    b = emu_lib(x, n_mels, fb=fb, window=window, reflect=(pad == "reflect"), log=log, rounds=True)
    assert np.array_equal(a, b)


def test_round_tables_dense_and_odd_filter_counts(emu_lib):
    rng = np.random.default_rng(4)
    window, _ = load_params("P0")
    x = signals.noise(1, 3000, 2)
    for n_mels in (1, 7, 17, 33):
        fb = np.ascontiguousarray(rng.random((513, n_mels), dtype=np.float32))
        fb[rng.random((513, n_mels)) < 0.5] = 0.0
        a = emu_lib(x, n_mels, fb=fb, window=window, log=False)
        b = emu_lib(x, n_mels, fb=fb, window=window, log=False, rounds=True)
        ref = mel_oracle.mel_forward(x, fb=fb, window=window, log_scale=False, dtype=np.float64)
        assert np.array_equal(a, b)
        assert np.abs(b - ref).max() / ref.max() < 1e-5


@pytest.mark.parametrize("log", [True, False])
def test_static_mel_stage_is_bit_identical_to_the_generic_one(emu_lib, log):
    """The generated straight-line mel code for the baked P0 filterbank (mel_static_gen.h: the hybrid
    form -- filters below kStaticP0Filters generated, pair tables for the rest) reproduces the generic
    pair-table stage bit for bit: same chains, same order, zero weights dropped."""
    window, fb = load_params("P0")
    for x in (signals.noise(2, 9000, 11), signals.music(6000, seed=3)[None, :], 40.0 * signals.noise(1, 3000, 5),
              np.zeros((1, 2000), np.float32)):
        y_gen = emu_lib(x, 80, fb=fb, window=window, log=log)
        y_st = emu_lib(x, 80, fb=fb, window=window, log=log, static_mel=2)
        assert np.array_equal(y_st.view(np.uint32), y_gen.view(np.uint32))


@pytest.mark.parametrize("pset,log,mode", [("P128", True, True), ("P1", False, True), ("T5", False, True), ("T5", True, True),
                                           ("P0", True, True), ("P0", False, True)])
def test_direct_mel_stages_agree_with_the_generic_one_and_the_oracle(emu_lib, pset, log, mode):
    """The direct forms (the default of every baked reference filterbank) sum every filter in one or two
    chains instead of four, so they are not bit-identical to the generic stage: they must agree with
    it to a few ulp, keep all-zero filters and silence exactly 0, and meet
    the fp64 oracle like every other path."""
    window, fb = load_params(pset)
    n_mels = fb.shape[1]
    empty = np.flatnonzero(~(fb != 0).any(axis=0))
    for x in (signals.noise(2, 9000, 11), signals.music(6000, seed=3)[None, :], 40.0 * signals.noise(1, 3000, 5),
              np.zeros((1, 2000), np.float32)):
        y_gen = emu_lib(x, n_mels, fb=fb, window=window, log=log)
        y_dir = emu_lib(x, n_mels, fb=fb, window=window, log=log, static_mel=mode)
        if log:   # log domain: a few float32 ulp of values <= ~12
            assert np.abs(y_dir - y_gen).max() <= 2e-6
        else:
            assert np.all(np.abs(y_dir - y_gen) <= 4e-6 * np.abs(y_gen))
        assert np.all(y_dir[..., empty] == 0.0)
        if not x.any():
            assert np.all(y_dir == 0.0)
        ref = mel_oracle.mel_forward(x, fb=fb, window=window, log_scale=log, dtype=np.float64)
        err = np.abs(y_dir - ref).max() if log else np.abs(np.log1p(y_dir.astype(np.float64)) - np.log1p(ref)).max()
        assert err < 2e-5
