"""CPU: host-side mirror of the reference interface (no kernel launches)."""
import copy
import inspect
import pickle

import numpy as np
import pytest
import torch

from beatheritage_b200 import MelSpectrogram
from beatheritage_b200 import segment as seg
from beatheritage_b200.spectrogram import melscale_fbanks_htk
from oracle import mel_oracle, ref_loader
from tests.conftest import PSET_ARGS, load_params

P0 = ("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect")


def test_constructor_signature_matches_reference():
    sig = inspect.signature(MelSpectrogram.__init__)
    names = list(sig.parameters)[1:]
    assert names == ["implementation", "log_scale", "sample_rate", "n_ftt", "n_mels", "hop_length", "f_min",
                     "f_max", "pad_mode"]                     # reference spectrogram.py:8-19, positional order matters
    defaults = [sig.parameters[n].default for n in names]
    assert defaults == ["nnAudio", False, 16000, 2048, 512, 128, 0, 8000, "constant"]
    if ref_loader.available():
        ref_sig = inspect.signature(ref_loader.load_reference_class().__init__)
        assert list(ref_sig.parameters) == list(sig.parameters)
        assert [p.default for p in ref_sig.parameters.values()] == [p.default for p in sig.parameters.values()]


def test_positional_and_keyword_construction_like_the_call_sites():
    m = MelSpectrogram(*P0)                                   # modeling_mapperatorinator.py:56-66
    assert (m.n_mels, m.pad_mode, m.log_scale) == (80, "reflect", True)
    m2 = MelSpectrogram("torchaudio", True, 16000, 1024, 80, 128, f_min=20, f_max=8000)   # dataloading.py:81-90
    assert m2.pad_mode == "constant"


def test_state_dict_keys_and_buffers_match_reference():
    m = MelSpectrogram(*P0)
    sd = m.state_dict()
    assert list(sd) == ["transform.spectrogram.window", "transform.mel_scale.fb"]
    assert sd["transform.spectrogram.window"].shape == (1024,)
    assert sd["transform.mel_scale.fb"].shape == (513, 80)
    assert len(list(m.parameters())) == 0
    for pset, (log, n_mels, f_min, f_max, pad) in PSET_ARGS.items():
        window, fb = load_params(pset)            # the reference module's own buffers
        mm = MelSpectrogram("torchaudio", log, 16000, 1024, n_mels, 128, f_min, f_max, pad)
        assert np.array_equal(mm.transform.spectrogram.window.numpy(), window)
        assert np.array_equal(mm.transform.mel_scale.fb.numpy(), fb), pset


def test_strict_load_of_a_reference_state_dict():
    window, fb = load_params("P0")
    ref_sd = {"transform.spectrogram.window": torch.from_numpy(window), "transform.mel_scale.fb": torch.from_numpy(fb * 2)}
    m = MelSpectrogram(*P0)
    stamp = m._buffer_stamp()
    m.load_state_dict(ref_sd, strict=True)                   # inference.py:478-480 loads strictly
    assert torch.equal(m.transform.mel_scale.fb, torch.from_numpy(fb * 2))
    assert m._buffer_stamp() != stamp                         # device tables will be rebuilt

    class Model(torch.nn.Module):                             # attribute name carries "spectrogram" (inference.py:486-489)
        def __init__(self):
            super().__init__()
            self.spectrogram = MelSpectrogram(*P0)
    keys = list(Model().state_dict())
    assert keys == ["spectrogram.transform.spectrogram.window", "spectrogram.transform.mel_scale.fb"]


def test_errors_mirror_the_reference():
    with pytest.raises(AssertionError):
        MelSpectrogram("librosa")                             # spectrogram.py:35
    with pytest.raises(NotImplementedError):
        MelSpectrogram()                                      # nnAudio arithmetic is unpinned: opt-in only
    with pytest.raises(ValueError):
        MelSpectrogram("torchaudio", True, 16000, 2048)
    m = MelSpectrogram(*P0)
    with pytest.raises(RuntimeError):
        m(torch.zeros(4, 3, 2048))                            # not [batch, samples]
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 512))                                # reflect pad needs N > 512 (F.pad raises)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(2, 4096))                               # no CPU fallback by design


def test_module_survives_copy_and_pickle_without_carrying_handles():
    m = MelSpectrogram(*P0)
    m._handles[0] = 12345
    c = copy.deepcopy(m)
    assert c._handles == {} and c._key != m._key
    p = pickle.loads(pickle.dumps(m))
    assert p._handles == {} and p.n_mels == 80
    m._handles.clear()


def test_fake_kernel_gives_shapes_for_compile():
    from torch._subclasses.fake_tensor import FakeTensorMode
    m = MelSpectrogram(*P0)
    with FakeTensorMode():
        x = torch.empty(6, 524160, device="cuda")
        y = torch.ops.beatheritage_b200.mel_forward(x, m._key, 80)
        assert tuple(y.shape) == (6, 4096, 80) and y.dtype == torch.float32


def test_torch_fb_builder_equals_oracle_restatement_for_p0():
    fb = melscale_fbanks_htk(513, 20.0, 8000.0, 80, 16000).numpy()
    assert np.array_equal(fb, mel_oracle.melscale_fbanks(513, 20.0, 8000.0, 80, 16000, np.float32))


def test_segment_plan_matches_oracle_segment():
    for n, kw in [(2_880_000, {}), (2_880_000, dict(parallel=True)), (57_600_000, {}), (1000, {}),
                  (524160, {}), (524161, {}), (600000, dict(lookback=0.25, lookahead=0.25))]:
        plan = seg.segment_plan(n, **kw)
        w, s = mel_oracle.segment_params(lookback=kw.get("lookback", 0.5), lookahead=kw.get("lookahead", 0.4),
                                         parallel=kw.get("parallel", False))
        assert (plan.window_len, plan.stride) == (w, s)
        if n <= 3_000_000:
            assert plan.n_windows == mel_oracle.segment(np.zeros(n, np.float32), w, s).shape[0]
    assert seg.segment_plan(2_880_000).n_windows == 46
    assert seg.segment_plan(57_600_000).n_windows == 1090
    assert seg.segment_plan(57_600_000, parallel=True).n_windows == 110


def test_segment_plan_trimming_matches_reference_cases_and_oracle():
    """a10: sequence_times / start_time / end_time (preprocessor.py:72-90): the plan against the
    reference's own outputs (tests/golden/segment_cases.json) and, on random songs, the oracle."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "segment_cases.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        plan = seg.segment_plan(c["n_samples"], c["src_seq_len"], 128, c["lookback"], c["lookahead"], c["parallel"], 16000,
                                c["start_time"], c["end_time"], c["begin_pad"], c["end_pad"])
        assert plan.n_windows == c["n_windows"] and list(plan.sequence_times) == c["sequence_times"]
        assert plan.song_length_ms == c["song_length"]
        for i, g in enumerate(c["starts"]):
            if g is not None:
                assert plan.first_offset + i * plan.stride == g
    rng = np.random.default_rng(3)
    for _ in range(300):
        n = int(rng.integers(1, 5_000_000))
        ssl = int(rng.choice([512, 1024, 2048, 4096]))
        lb, la = float(rng.choice([0.0, 0.25, 0.5])), float(rng.choice([0.0, 0.2, 0.4]))
        dur = n / 16.0
        st = None if rng.random() < 0.3 else float(rng.uniform(-0.3, 1.4) * dur)
        en = None if rng.random() < 0.3 else float(rng.uniform(-0.3, 1.4) * dur)
        plan = seg.segment_plan(n, ssl, 128, lb, la, False, 16000, st, en)
        w, s = mel_oracle.segment_params(ssl, 128, lb, la)
        untrimmed = seg.segment_plan(n, ssl, 128, lb, la).n_windows
        first, kept, times = mel_oracle.segment_times_and_trim(untrimmed, w, s, 16000, lb, la, st, en)
        assert (plan.first_window, plan.n_windows, list(plan.sequence_times)) == (first, kept, times.tolist())
        assert plan.n_windows >= 1 and plan.first_offset == first * s


def test_dataset_window_plan_matches_oracle():
    for n in (2_880_000, 2_880_001, 1280, 100):
        plan = seg.dataset_window_plan(n)
        assert plan.n_windows == mel_oracle.dataset_windows(np.zeros(n, np.float32)).shape[0]
    assert seg.dataset_window_plan(2_880_000).n_windows == 6


def test_shard_range_partitions_everything_once():
    for n, ws in [(4096, 8), (10, 4), (3, 8), (0, 2)]:
        seen = sorted(i for r in range(ws) for i in seg.shard_range(n, r, ws))
        assert seen == list(range(n))
    with pytest.raises(ValueError):
        seg.shard_range(10, 4, 4)


def _baked_tables():
    """name -> dense uint32 [513, n_mels] table restated from csrc/bhmel_fb_baked.h (start / count / bits)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "beatheritage_b200", "csrc", "bhmel_fb_baked.h")).read()

    def arr(kind, name):
        body = re.search(rf"kBaked{name}{kind}\[\] = \{{(.*?)\}};", txt, re.S).group(1)
        return [int(v.rstrip("u"), 0) for v in body.replace("\n", " ").split(",") if v.strip()]

    out = {}
    for name, ident in re.findall(r'\{"(\w+)", (\d+), kBaked', txt):
        n_mels = int(re.search(rf"kBaked{name}Mels = (\d+)", txt).group(1))
        start, count, bits = arr("Start", name), arr("Count", name), arr("Bits", name)
        assert len(bits) == int(re.search(rf"kBaked{name}Nnz = (\d+)", txt).group(1)) == sum(count)
        t = np.zeros((513, n_mels), np.uint32)
        pos = 0
        for m in range(n_mels):
            for j in range(count[m]):
                t[start[m] + j, m] = bits[pos]
                pos += 1
        out[name] = (int(ident), t)
    return out


def test_baked_filterbank_header_holds_the_reference_tables():
    """csrc/bhmel_fb_baked.h (committed; selects the statically scheduled mel stage at run time and feeds
    its code generator) holds exactly the reference's `mel_scale.fb` buffers of every parameter set its
    configs use (the golden `params_*.npz` come from the unmodified reference module)."""
    tables = _baked_tables()
    assert {k: v[0] for k, v in tables.items()} == {"P0": 1, "P128": 2, "P1": 3, "T5": 4}
    ctor = {"P0": (20.0, 80), "P128": (20.0, 128), "P1": (0.0, 388), "T5": (0.0, 512)}
    for name, (_, t) in tables.items():
        _, fb = load_params(name)
        ref_bits = fb.view(np.uint32).copy()
        ref_bits[fb == 0] = 0          # the reference's P1 / T5 tables carry one negative zero (fb[0, 0] = -0.0)
        assert np.array_equal(t, ref_bits), name
        f_min, n_mels = ctor[name]
        mine = melscale_fbanks_htk(513, f_min, 8000.0, n_mels, 16000).numpy()
        assert np.array_equal(mine.view(np.uint32), fb.view(np.uint32))      # the host mirror keeps even that bit
    assert int((tables["P0"][1] != 0).sum()) == 1003


def test_static_mel_generator_covers_each_weight_once(tmp_path):
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "mel_static_gen.h"
    subprocess.run([sys.executable, os.path.join(root, "beatheritage_b200", "csrc", "gen_mel_static.py"), "-o", str(out)],
                   check=True)
    txt = out.read_text()
    n_static = int(re.search(r"kStaticP0Filters = (\d+)", txt).group(1))
    n_warps = int(re.search(r"kStaticP0Warps = (\d+)", txt).group(1))
    assert 0 < n_static < 80 and 0 < n_warps < 8
    tables = _baked_tables()
    for name, (_, t) in tables.items():
        fb = t.view(np.float32)
        n_mels = fb.shape[1]
        all_bits = sorted(int(b) for b in t[fb != 0])
        # direct form (every set): every non-zero weight once, every aligned group of four filters stored once
        body = re.search(rf"void mel_direct_{name}\(.*?\n}}\n", txt, re.S).group(0)
        got_bits = [int(h, 16) for h in re.findall(r"__uint_as_float\(0x([0-9a-f]+)u\)", body)]
        # fully covered blocks run packed: four weights from the constant table kWpk<name>, every entry used once
        tab = re.search(rf"kWpk{name}\[(\d+)\] = \{{(.*?)\n\}};", txt, re.S)
        if tab:
            vals = np.array([float(v.rstrip("f")) for v in re.findall(r"[-+0-9.e]+f", tab.group(2))], dtype=np.float32)
            assert len(vals) == 4 * int(tab.group(1))
            used = sorted(int(i) for i in re.findall(rf"kWpk{name}\[(\d+)\]\);", body))
            assert used == list(range(int(tab.group(1))))
            got_bits += [int(b) for b in vals.view(np.uint32)]
        assert sorted(got_bits) == all_bits, name
        staged = re.findall(r"mel_stage4\(srow, (\d+), v(\d+), v(\d+), v(\d+), v(\d+)\);", body)
        assert sorted(int(g[1]) for g in staged) == list(range(0, n_mels, 4))            # every aligned group of four, once
        assert all([int(g[1]) + j for j in range(4)] == [int(v) for v in g[1:]] and int(g[0]) % 4 == 0 and int(g[0]) < 32
                   for g in staged)
        runs = re.search(rf"kRun{name}\[\d+\]\[2\] = \{{(.*?)\}};", txt).group(1)
        runs = [(int(a), int(b)) for a, b in re.findall(r"\{(\d+), (\d+)\}", runs)]
        assert sorted(m for m0, n in runs for m in range(m0, m0 + n)) == list(range(n_mels))   # parts tile the filters
        assert all(n <= 32 and n % 4 == 0 and m0 % 4 == 0 for m0, n in runs)
        assert len(re.findall(r"const float v\d+ = ", body)) == n_mels
    # P0's hybrid form: the static warps cover every non-zero weight of filters below n_static once
    body = re.search(r"void mel_static_P0\(.*?\n}\n", txt, re.S).group(0)
    t = tables["P0"][1]
    fb = t.view(np.float32)
    want = sorted(int(b) for b in t[:, :n_static][fb[:, :n_static] != 0])
    assert sorted(int(h, 16) for h in re.findall(r"__uint_as_float\(0x([0-9a-f]+)u\)", body)) == want
    assert sorted(int(c) for c in re.findall(r"orow\[(\d+)\] = ", body)) == list(range(n_static))


# ---------------------------------------------------------------------------------------------
# N4: implementation="nnAudio" (reference spectrogram.py:50-61), opt-in, parity unpinned
# ---------------------------------------------------------------------------------------------
NNAUDIO_P1 = ("nnAudio", False, 16000, 1024, 388, 128, 0, 8000, "constant")   # configs/model/default.yaml:19-27


@pytest.fixture
def nnaudio_published(monkeypatch):
    monkeypatch.setattr(MelSpectrogram, "nnaudio_arithmetic", "published")


def test_nnaudio_is_opt_in_and_the_older_flag_still_maps_to_torchaudio(monkeypatch):
    with pytest.raises(NotImplementedError, match="nnaudio_arithmetic"):
        MelSpectrogram(*NNAUDIO_P1)
    monkeypatch.setattr(MelSpectrogram, "allow_nnaudio_as_torchaudio", True)
    m = MelSpectrogram(*NNAUDIO_P1)
    assert list(m.state_dict()) == ["transform.spectrogram.window", "transform.mel_scale.fb"]
    _, fb = load_params("P1")
    assert torch.equal(m.transform.mel_scale.fb, torch.from_numpy(fb))


def test_nnaudio_published_filterbank_and_state_dict_layout(nnaudio_published):
    from beatheritage_b200.spectrogram import melscale_fbanks_slaney
    from oracle import nnaudio_oracle
    m = MelSpectrogram(*NNAUDIO_P1)
    basis = nnaudio_oracle.mel_basis(16000, 1024, 388, 0.0, 8000.0)
    assert np.array_equal(m.transform.mel_scale.fb.numpy(), basis.T)             # two independent codings agree
    assert np.array_equal(melscale_fbanks_slaney(513, 20.0, 8000.0, 80, 16000).numpy(),
                          nnaudio_oracle.mel_basis(16000, 1024, 80, 20.0, 8000.0).T)
    sd = m.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "transform.mel_basis": (388, 513), "transform.stft.wsin": (513, 1, 1024),
        "transform.stft.wcos": (513, 1, 1024), "transform.stft.window_mask": (1, 1024, 1)}
    wsin, wcos, window = nnaudio_oracle.fourier_kernels()
    assert np.array_equal(sd["transform.mel_basis"].numpy(), basis)
    assert np.abs(sd["transform.stft.wsin"].numpy() - wsin).max() < 1e-6
    assert np.abs(sd["transform.stft.wcos"].numpy() - wcos).max() < 1e-6
    assert np.array_equal(sd["transform.stft.window_mask"].numpy().reshape(-1), window)


def test_nnaudio_checkpoint_buffers_load_strictly_and_take_precedence(nnaudio_published):
    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.spectrogram = MelSpectrogram(*NNAUDIO_P1)
    src = Model()
    sd = src.state_dict()
    assert list(sd) == ["spectrogram.transform.mel_basis", "spectrogram.transform.stft.wsin",
                        "spectrogram.transform.stft.wcos", "spectrogram.transform.stft.window_mask"]
    # a checkpoint with its own tables: scaled basis, a different (still plain-DFT) window, inverse kernels
    win = torch.hamming_window(1024)
    ang = 2 * np.pi * torch.arange(513, dtype=torch.float64)[:, None] * torch.arange(1024, dtype=torch.float64) / 1024
    sd["spectrogram.transform.mel_basis"] = sd["spectrogram.transform.mel_basis"] * 3
    sd["spectrogram.transform.stft.wcos"] = (torch.cos(ang).float() * win)[:, None]
    sd["spectrogram.transform.stft.wsin"] = (torch.sin(ang).float() * win)[:, None]
    sd["spectrogram.transform.stft.window_mask"] = win.reshape(1, -1, 1)
    sd["spectrogram.transform.stft.kernel_sin_inv"] = torch.zeros(513, 1024, 1)
    dst = Model()
    stamp = dst.spectrogram._buffer_stamp()
    res = dst.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(dst.spectrogram.transform.mel_scale.fb, src.spectrogram.transform.mel_scale.fb * 3)
    assert torch.equal(dst.spectrogram.transform.spectrogram.window, win)
    assert dst.spectrogram._buffer_stamp() != stamp                       # device tables will be rebuilt
    # our own layout still loads (e.g. a state dict saved before switching the module to nnAudio names)
    dst.load_state_dict({"spectrogram.transform.spectrogram.window": torch.hann_window(1024),
                         "spectrogram.transform.mel_scale.fb": src.spectrogram.transform.mel_scale.fb}, strict=True)
    assert torch.equal(dst.spectrogram.transform.spectrogram.window, torch.hann_window(1024))
    # a trained STFT (kernels that are not window * sin/cos) is refused, not silently approximated
    bad = dict(sd)
    bad["spectrogram.transform.stft.wsin"] = bad["spectrogram.transform.stft.wsin"] + 0.01
    with pytest.raises(RuntimeError, match="trained STFT"):
        Model().load_state_dict(bad, strict=True)
    with pytest.raises(RuntimeError, match="mel_basis"):
        Model().load_state_dict({"spectrogram.transform.mel_basis": torch.zeros(80, 513)}, strict=False)


def test_nnaudio_module_copies_and_pickles(nnaudio_published):
    m = MelSpectrogram(*NNAUDIO_P1)
    for c in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert list(c.state_dict()) == list(m.state_dict())
        c.load_state_dict(m.state_dict(), strict=True)


# ---------------------------------------------------------------------------------------------
# N3: host side of the conv stem (no kernel launches)
# ---------------------------------------------------------------------------------------------
def test_conv_stem_mirrors_the_encoders_parameters_and_has_no_cpu_path():
    from beatheritage_b200.conv_stem import ConvStem

    class Encoder(torch.nn.Module):                       # modeling_ropewhisper.py:1135-1136
        def __init__(self):
            super().__init__()
            self.conv1 = torch.nn.Conv1d(464, 768, kernel_size=3, padding=1)
            self.conv2 = torch.nn.Conv1d(768, 768, kernel_size=3, stride=2, padding=1)
            self.layer_norm = torch.nn.LayerNorm(768)
    enc = Encoder()
    stem = ConvStem.from_encoder(enc)
    assert list(stem.state_dict()) == ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias"]
    for k, v in stem.state_dict().items():
        assert torch.equal(v, enc.state_dict()[k])
    assert stem.conv2.stride == (2,) and stem.conv1.padding == (1,) and stem.conv1.kernel_size == (3,)
    stamp = stem._param_stamp()
    stem.load_state_dict({k: v * 2 for k, v in stem.state_dict().items()})
    assert stem._param_stamp() != stamp                   # a reload repacks the device weights at the next call
    with pytest.raises(ValueError):
        ConvStem(465, 768)                                # TMA needs 16-byte rows
    with pytest.raises(ValueError):
        ConvStem(464, 100)
    with pytest.raises(RuntimeError, match="no CPU path"):
        stem(torch.zeros(1, 64, 464, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError, match="channels-last"):
        stem(torch.zeros(1, 464, 64, dtype=torch.bfloat16))
    p = pickle.loads(pickle.dumps(stem))
    assert p._handles == {} and list(p.state_dict()) == list(stem.state_dict())


def test_slaney_bank_matches_the_librosa_compatible_bank_of_transformers():
    """N4, second independent pin of the Slaney / area-normalised bank nnAudio documents (`htk=False, norm=1`,
    librosa's filterbank): `transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney")` -- the
    librosa-compatible implementation Whisper's feature extractor uses -- agrees with
    melscale_fbanks_slaney to float32 rounding for every n_mels of the reference configs.  (nnAudio itself
    stays un-installable here: the row remains "parity unpinned".)"""
    au = pytest.importorskip("transformers.audio_utils")
    import warnings
    from beatheritage_b200.spectrogram import melscale_fbanks_slaney
    for n_mels in (80, 128, 388, 512):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            hf = au.mel_filter_bank(num_frequency_bins=513, num_mel_filters=n_mels, min_frequency=0.0, max_frequency=8000.0,
                                    sampling_rate=16000, norm="slaney", mel_scale="slaney")
        mine = melscale_fbanks_slaney(513, 0.0, 8000.0, n_mels, 16000).numpy()
        assert hf.shape == mine.shape
        assert np.abs(hf - mine).max() <= 4e-8 * max(1.0, np.abs(hf).max() / 0.1)
        assert np.array_equal(hf != 0, mine != 0) or np.abs(hf[(hf != 0) != (mine != 0)]).max() < 1e-9



def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU only): one JSON line with the contract's keys; it runs the reference's
    own module when oracle/_ref holds the copy build() makes, else the torch port, and says which."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "log-mel audio-sec/sec" and line["unit"] == "audio-s/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    have_copy = os.path.exists(os.path.join(root, "oracle", "_ref", "spectrogram.py"))
    assert cb["kind"] == ("reference" if have_copy else "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert "workload" in line["config"]
