"""GPU (-m gpu): the tcgen05 conv stem (include/bhstem.h) against the CPU oracle and
against the reference's own torch modules in bf16 on the GPU.

Tolerance: both sides round to bf16 at the same points (twice per stage: the convolution's output and
GELU's), so results agree except where the fp32 summation order flips one of those roundings -- a
flipped pre-activation (1 ulp of x) moves gelu(x) by up to 1.13 ulp(x), which is 2-3 ulps of the
smaller gelu(x), before that is rounded itself: |diff| <= 2^-6 |want| + 4e-3 everywhere.  One stage on
identical inputs: fewer than 2 % of the elements differ at all; the two-stage stem, where conv2 sees
the flipped hidden values: fewer than 10 %."""
import numpy as np
import pytest
import torch

from oracle import conv_stem_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def make_stem(c_in, d, dev, seed=0):
    from beatheritage_b200.conv_stem import ConvStem
    torch.manual_seed(seed)
    stem = ConvStem(c_in, d)
    with torch.no_grad():                        # bf16-representable parameters, like a model cast with .to(bfloat16)
        for p in stem.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    return stem.to(dev)


def make_input(B, T, c_in, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, c_in, generator=g) * 1.5       # log-mel + embeddings are O(1)
    return x.to(torch.bfloat16)


def assert_close(got, want, what, max_frac=0.02):
    got, want = got.float().cpu(), want.float().cpu()
    assert got.shape == want.shape, what
    diff = (got - want).abs()
    bound = want.abs() * 2.0 ** -6 + 4e-3
    worst = float((diff - bound).max())
    assert worst <= 0, f"{what}: exceeds 2^-6 |want| + 4e-3 by {worst} (max diff {float(diff.max())})"
    frac = float((diff > 0).float().mean())
    assert frac < max_frac, f"{what}: {frac:.4f} of the elements differ"


@pytest.mark.parametrize("B,T,c_in,d", [
    (2, 256, 464, 768),        # whisper-small dims (C5): 29 channel blocks of 16, BN = 256
    (1, 200, 464, 768),        # ragged: partial row tile, TMA zero fill below the last row
    (3, 64, 80, 384),          # whisper-tiny width: BN = 128, 80 channels = 1 full + 1 quarter block
    (1, 2, 8, 128),            # smallest legal problem
])
def test_each_stage_matches_the_oracle(dev, B, T, c_in, d):
    stem = make_stem(c_in, d, dev, seed=c_in + d)
    x = make_input(B, T, c_in, seed=T)
    want_y, want_h = conv_stem_oracle.conv_stem(x, stem.conv1.weight, stem.conv1.bias, stem.conv2.weight,
                                                stem.conv2.bias, return_hidden=True)
    before = stem.launch_count()
    h = stem.forward_stage(1, x.to(dev))
    torch.cuda.synchronize()
    assert stem.launch_count() == before + 1
    assert h.shape == (B, T, d) and h.dtype == torch.bfloat16
    assert_close(h, want_h, "gelu(conv1)")
    y2 = stem.forward_stage(2, want_h.to(torch.bfloat16).to(dev))      # stage 2 on the oracle's hidden: isolates conv2
    assert_close(y2, want_y, "gelu(conv2) on the oracle's hidden")
    y = stem(x.to(dev))
    torch.cuda.synchronize()
    assert y.shape == (B, T // 2, d) and y.is_contiguous()
    assert_close(y, want_y, "stem", 0.10)
    assert "libbhstem.so" in open("/proc/self/maps").read()


def test_full_context_matches_the_reference_modules_on_the_gpu(dev):
    """B = 6 windows x 4096 frames x 464 channels (C2 parallel batch, C5 dims): against torch's own
    Conv1d + gelu in bf16 on the GPU (the modules the reference runs), and the oracle on one window."""
    stem = make_stem(464, 768, dev, seed=7)
    x = make_input(6, 4096, 464, seed=11).to(dev)
    y = stem(x)
    ref = stem.to(torch.bfloat16)
    with torch.no_grad():
        r = torch.nn.functional.gelu(ref.conv1(x.swapaxes(1, 2)))
        r = torch.nn.functional.gelu(ref.conv2(r)).permute(0, 2, 1)
    torch.cuda.synchronize()
    assert_close(y, r, "stem vs torch bf16 modules on the GPU", 1.0)     # cuDNN rounds at other points
    want = conv_stem_oracle.conv_stem(x[4:5].cpu(), ref.conv1.weight, ref.conv1.bias, ref.conv2.weight, ref.conv2.bias)
    assert_close(y[4:5], want, "stem vs oracle, window 4", 0.10)
    assert_close(r[4:5], want, "torch bf16 modules vs oracle (pins the oracle)", 1.0)


def test_linearity_free_properties(dev):
    """Size-independent properties: zero input gives gelu(bias) rows; windows are independent of their
    batch neighbours; an all-zero stem gives exactly 0."""
    stem = make_stem(464, 768, dev, seed=3)
    x = make_input(3, 512, 464, seed=5).to(dev)
    y = stem(x)
    y1 = stem(x[1:2])
    assert torch.equal(y[1:2], y1)
    z = stem.forward_stage(1, torch.zeros(1, 128, 464, dtype=torch.bfloat16, device=dev))
    want = torch.nn.functional.gelu(stem.conv1.bias.to(torch.bfloat16).float()).to(torch.bfloat16)
    assert torch.equal(z[0, 5], want) and torch.equal(z[0, 0], want) and torch.equal(z[0, 127], want)
    with torch.no_grad():
        for p in stem.parameters():
            p.zero_()
    assert torch.count_nonzero(stem(x)) == 0      # parameters edited in place: the handle repacks


def test_errors(dev):
    from beatheritage_b200.conv_stem import ConvStem
    with pytest.raises(ValueError):
        ConvStem(465, 768)
    stem = make_stem(80, 128, dev)
    with pytest.raises(RuntimeError, match="no CPU path"):
        stem(torch.zeros(1, 64, 80, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError, match="bfloat16"):
        stem(torch.zeros(1, 64, 80, device=dev))
    with pytest.raises(RuntimeError, match="even"):
        stem(torch.zeros(1, 63, 80, dtype=torch.bfloat16, device=dev))
    with pytest.raises(RuntimeError, match="channels-last"):
        stem(torch.zeros(1, 80, 64, dtype=torch.bfloat16, device=dev))


def gelu_model(x):
    """numpy fp32 restatement of the kernel's GELU (csrc/bhstem.cu::conv_gelu: Abramowitz-Stegun 7.1.26),
    with exact 1/x and 2^x where the GPU uses MUFU approximations."""
    f = np.float32
    u = np.abs(x)
    t = (f(1) / (f(0.3275911 * 0.70710678118654752440) * u + f(1))).astype(f)
    e = np.exp2((u * u * f(-0.5 * 1.4426950408889634)).astype(f)).astype(f)
    poly = t * f(1.061405429) + f(-1.453152027)
    poly = poly * t + f(1.421413741)
    poly = poly * t + f(-0.284496736)
    poly = poly * t + f(0.254829592)
    erf_abs = f(1) - poly * t * e
    h = f(0.5) * x
    return (np.abs(h).astype(np.float64) * erf_abs.astype(np.float64) + h.astype(np.float64)).astype(f)   # one fma


def test_gelu_is_checked_for_every_bf16_input(dev):
    """GELU in the stem is a pure function bf16 -> bf16 (the conv output is rounded first).  An identity
    convolution (centre tap = I, bias 0) makes gelu(conv1(x)) = gelu(x) exactly, so ALL finite bf16 values
    below 1e30 go through the kernel's epilogue and are compared with torch's fp32 erf GELU on the same GPU:
    identical except in the tail x <= -3.5, where 1 + erf cancels in both and |diff| <= 1e-5."""
    from beatheritage_b200.conv_stem import ConvStem
    bits = torch.arange(65536, dtype=torch.int32)
    x = (bits << 16).view(torch.float32)
    x = torch.where(torch.isfinite(x) & (x.abs() < 1e30), x, torch.zeros(()))
    stem = ConvStem(128, 128)
    with torch.no_grad():
        for p in stem.parameters():
            p.zero_()
        stem.conv1.weight[:, :, 1] = torch.eye(128)
    stem = stem.to(dev)
    xb = x.to(torch.bfloat16).reshape(1, 512, 128).to(dev)
    got = stem.forward_stage(1, xb).reshape(-1).float().cpu()
    want = torch.nn.functional.gelu(xb.float()).to(torch.bfloat16).reshape(-1).float().cpu()
    with np.errstate(all="ignore"):
        model = torch.from_numpy(gelu_model(x.numpy())).to(torch.bfloat16).float()
    differs = got != want
    assert not bool((differs & (x > -3.5)).any()), x[differs & (x > -3.5)][:10]
    assert float((got - want).abs().max()) <= 1e-5
    assert int(differs.sum()) <= 64
    # the CPU model of the same formula agrees with the GPU up to the MUFU approximations
    assert int((got != model).sum()) <= 64 and float((got - model).abs().max()) <= 1e-5


def test_c5_slice_song_to_stem_output_matches_the_reference_op_chain(dev):
    """C5 as a parity test, not only a measurement: one model-context window of a
    music-like song through  OUR frontend -> fused encoder-input assembly (80 mel + 384 conditioning
    channels, channels last, bf16) -> tcgen05 conv stem  against the REFERENCE's op chain on the same
    GPU: torchaudio MelSpectrogram + log1p + permute (spectrogram.py:38-49, 79-82), .to(bf16), expand +
    cat of the conditioning channels, swapaxes (modeling_mapperatorinator.py:351-376), then
    gelu(conv1) / gelu(conv2) / permute with torch's bf16 Conv1d (modeling_ropewhisper.py:1206-1209)."""
    torchaudio = pytest.importorskip("torchaudio")
    from beatheritage_b200 import MelSpectrogram
    from tests.golden import signals
    mel = MelSpectrogram("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect").to(dev)
    ta = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=1024, n_mels=80, hop_length=128, center=True,
                                              f_min=20, f_max=8000, pad_mode="reflect").to(dev)
    stem = make_stem(464, 768, dev, seed=5)
    x = torch.from_numpy(signals.music(2 * 524160, seed=9).reshape(2, 524160)).to(dev)
    torch.manual_seed(1)
    cond = torch.randn(2, 384, device=dev)
    with torch.no_grad():
        # reference chain
        frames = torch.log1p(ta(x)).permute(0, 2, 1)                                   # spectrogram.py:79-82
        frames = frames.to(torch.bfloat16)                                             # :352
        emb = torch.cat([frames, cond.to(torch.bfloat16).unsqueeze(1).expand(-1, frames.shape[1], -1)], dim=-1)
        ref_in = emb.swapaxes(1, 2)                                                    # :375-376  [B, 464, T]
        conv1 = stem.conv1.to(torch.bfloat16)
        conv2 = stem.conv2.to(torch.bfloat16)
        want = torch.nn.functional.gelu(conv2(torch.nn.functional.gelu(conv1(ref_in)))).permute(0, 2, 1)
        conv1.float(), conv2.float()
        # ours
        enc_in = mel.forward_encoder_input(x, [cond], dtype=torch.bfloat16, channels_first=False)
        got = stem(enc_in)
    torch.cuda.synchronize()
    # the two front ends differ by ~1e-6 before the bf16 rounding, so a handful of mel values round the
    # other way; everything downstream is the stem's own tolerance
    mel_diff = (enc_in[..., :80].float() - frames.float()).abs()
    assert float(mel_diff.max()) <= 2.0 ** -7 * float(frames.float().abs().max())      # at most one bf16 ulp
    assert float((mel_diff > 0).float().mean()) < 0.01
    assert torch.equal(enc_in[..., 80:], emb[..., 80:])
    # (a flipped mel value reaches 3 x 768 hidden values and every output under them, so most outputs move
    # by an fp32 hair and many bf16 roundings flip: only the magnitude bound is meaningful here)
    assert_close(got, want, "song -> stem output (C5 slice)", max_frac=1.01)


@pytest.mark.parametrize("B,T", [(2, 512), (1, 200), (3, 4096)])
def test_kernel_variants_agree(dev, B, T):
    """The CTA-pair kernel (default for d_model % 256 == 0: tcgen05.mma.cta_group::2, half a weight tile per
    CTA) and the one-CTA shared-tap kernel accumulate the same products in the same order: identical
    results, ragged row tiles included.  The one-box-per-tap kernel sums tap-major instead of channel-block
    major, so it may flip a bf16 rounding here and there: the stem's usual tolerance."""
    stem = make_stem(464, 768, dev, seed=3)
    x = make_input(B, T, 464, seed=B * 7 + T).to(dev)
    outs = {}
    for variant in ("cta_pairs", "shared_taps", "tap_boxes"):
        stem.set_variant(variant)
        outs[variant] = stem(x).clone()
    stem.set_variant("cta_pairs")
    stem.set_deep_a_ring(0)                      # 2 activation + 8 weight stages instead of the default 3 + 6
    outs["cta_pairs_2a8w"] = stem(x).clone()
    stem.set_deep_a_ring(3)
    torch.cuda.synchronize()
    assert torch.equal(outs["cta_pairs"], outs["shared_taps"]) and torch.equal(outs["cta_pairs"], outs["cta_pairs_2a8w"])
    assert_close(outs["tap_boxes"], outs["shared_taps"], "tap_boxes vs shared_taps", max_frac=0.10)
    assert bool(torch.isfinite(outs["cta_pairs"].float()).all())


def test_programmatic_dependent_launch_keeps_stream_order(dev):
    """BHSTEM_OPT_PDL (default on): conv2 may become resident while conv1 drains, and the next call's conv1
    while this call's conv2 does, but no kernel touches global memory before its predecessor has completed.
    A chain of calls that reuse ONE hidden / output buffer and feed each result into the next input
    (read-after-write, write-after-read and write-after-write hazards between adjacent launches) gives the
    same bits with and without it."""
    stem = make_stem(464, 768, dev, seed=5)
    x0 = make_input(2, 1024, 464, seed=11).to(dev)
    hidden = torch.empty(2, 1024, 768, dtype=torch.bfloat16, device=dev)
    out = torch.empty(2, 512, 768, dtype=torch.bfloat16, device=dev)

    def chain():
        x = x0.clone()
        results = []
        for _ in range(6):
            y = stem(x, hidden=hidden, out=out)              # [2, 512, 768], same buffers every time
            results.append(y.clone())
            # next input depends on this output (same stream): first 464 channels of y, frames repeated twice
            x = y[:, :, :464].repeat_interleave(2, dim=1).contiguous()
        torch.cuda.synchronize()
        return results

    with_pdl = chain()
    stem.set_pdl(False)
    without = chain()
    stem.set_pdl(True)
    for a, b in zip(with_pdl, without):
        assert torch.equal(a, b)
    # back-to-back launches with nothing in between (the case the overlap is for)
    ys = [stem(x0, hidden=hidden, out=out).clone() for _ in range(8)]
    torch.cuda.synchronize()
    for y in ys[1:]:
        assert torch.equal(y, ys[0])


def test_cuda_graph_capture_of_the_serving_chain(dev):
    """The serving loop of one model-context window -- frontend + encoder-input assembly + conv stem, four
    launches, two of them programmatic dependents -- captured ONCE in a CUDA graph and replayed on new audio in
    the same buffers gives the bits of the eager calls: the C ABI makes no allocation, no synchronisation and
    no host-side state change per call, so it is capturable as it stands."""
    from beatheritage_b200 import MelSpectrogram
    from tests.golden import signals
    mel = MelSpectrogram("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect").to(dev)
    stem = make_stem(464, 768, dev, seed=5)
    torch.manual_seed(2)
    cond = torch.randn(1, 384, device=dev)
    x = torch.zeros(1, 524160, device=dev)
    songs = [torch.from_numpy(signals.music(524160, seed=s).reshape(1, 524160)).to(dev) for s in (21, 22, 23)]
    hidden = torch.empty(1, 4096, 768, dtype=torch.bfloat16, device=dev)
    out = torch.empty(1, 2048, 768, dtype=torch.bfloat16, device=dev)

    def chain():
        enc_in = mel.forward_encoder_input(x, [cond], dtype=torch.bfloat16, channels_first=False)
        return stem(enc_in, hidden=hidden, out=out)

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):                 # handles, scratch and lazy state exist before the capture
        x.copy_(songs[0])
        for _ in range(3):
            chain()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    before = (mel.launch_count(), stem.launch_count())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y = chain()
    assert y.data_ptr() == out.data_ptr()
    for song in songs:
        x.copy_(song)
        graph.replay()
        replayed = out.clone()
        eager = chain().clone()
        torch.cuda.synchronize()
        assert torch.equal(replayed, eager)
        assert bool(torch.isfinite(replayed.float()).all()) and float(replayed.float().abs().max()) > 0
    assert mel.launch_count() > before[0] and stem.launch_count() > before[1]


def test_one_call_larger_than_int32_element_counts(dev):
    """Maximum sizes: 700 windows in one call -- the hidden activations are 2.2e9 bf16 values (past 2^31
    elements, 4.4 GB), the input 1.33e9, the output 1.1e9 -- rows around the boundary and at both ends equal the
    same windows computed alone (the CTA-pair kernel is deterministic and batch-independent)."""
    B, T, C, D = 700, 4096, 464, 768
    need = 2 * (B * T * C + B * T * D + B * (T // 2) * D)
    free, _ = torch.cuda.mem_get_info(dev)
    if free < need * 1.2:
        pytest.skip(f"needs {need / 2**30:.1f} GiB of device memory")
    stem = make_stem(C, D, dev, seed=8)
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.empty(B, T, C, dtype=torch.bfloat16, device=dev)
    for b0 in range(0, B, 100):
        x[b0:b0 + 100] = (torch.randn(min(100, B - b0), T, C, device=dev, generator=g) * 1.5).to(torch.bfloat16)
    hidden = torch.empty(B, T, D, dtype=torch.bfloat16, device=dev)
    assert hidden.numel() > 2 ** 31
    y = stem(x, hidden=hidden)
    torch.cuda.synchronize()
    for r in (0, 1, 340, 341, 342, 682, 683, B - 1):       # 2^31 / (4096 * 768) = 682.67: hidden row 682 straddles it
        alone = stem(x[r:r + 1].contiguous())
        assert torch.equal(y[r], alone[0]), f"window {r}"
    assert bool(torch.isfinite(y[::53].float()).all())
    del x, y, hidden
    torch.cuda.empty_cache()


# ------------------------------------------------------------------ split conv1 (bhstem_forward_split)
def _cat_input(frames, cond):
    """The reference's concatenation (modeling_mapperatorinator.py:368-370)."""
    return torch.cat([frames, cond.unsqueeze(1).expand(-1, frames.shape[1], -1)], dim=-1)


@pytest.mark.parametrize("B,T,c_in,d,n_var", [
    (2, 256, 464, 768, 80),        # the reference's dims: 80 mel + 3 x 128 conditioning channels
    (1, 200, 464, 768, 80),        # ragged row tile: the last frame sits inside a partial tile
    (3, 258, 464, 768, 80),        # the last frame alone in its CTA-pair tile
    (3, 64, 80, 384, 16),          # BN = 128 kernel, one 16-channel block
    (1, 2, 16, 128, 8),            # smallest: every frame is an edge frame
    (2, 130, 208, 256, 128),       # 128 varying channels: two full 64-channel blocks (lean issue path)
])
def test_split_conv1_matches_the_oracle_on_the_concatenated_input(dev, B, T, c_in, d, n_var):
    stem = make_stem(c_in, d, dev, seed=c_in + d + 1)
    x = make_input(B, T, c_in, seed=T + 1)
    frames, cond = x[:, :, :n_var].contiguous(), x[:, 0, n_var:].contiguous()
    full = _cat_input(frames, cond)
    want_y, want_h = conv_stem_oracle.conv_stem(full, stem.conv1.weight, stem.conv1.bias, stem.conv2.weight,
                                                stem.conv2.bias, return_hidden=True)
    hidden = torch.empty(B, T, d, dtype=torch.bfloat16, device=dev)
    before = stem.launch_count()
    y = stem.forward_split(frames.to(dev), cond.to(dev), hidden=hidden)
    torch.cuda.synchronize()
    assert stem.launch_count() == before + 3
    assert y.shape == (B, T // 2, d) and y.dtype == torch.bfloat16 and y.is_contiguous()
    assert_close(hidden, want_h, "gelu(split conv1)")
    # the two edge frames see one tap of zero padding each: check them on their own as well
    assert_close(hidden[:, 0], want_h[:, 0], "gelu(split conv1), first frame", 0.05)
    assert_close(hidden[:, T - 1], want_h[:, T - 1], "gelu(split conv1), last frame", 0.05)
    assert_close(y, want_y, "split stem", 0.10)
    # and against this library's own convolution over the materialised input
    hidden_full = stem.forward_stage(1, full.to(dev))
    assert_close(hidden, hidden_full, "split conv1 vs conv1 over the concatenated input")
    assert "libbhstem.so" in open("/proc/self/maps").read()


def test_split_conv1_edge_frames_drop_exactly_one_tap(dev):
    """Zero frames isolate the folded bias: interior frames carry bias + S0 + S1 + S2, frame 0 loses tap 0,
    frame T-1 loses tap 2 (the zero padding).  Checked against fp64 sums of the same bf16 products."""
    stem = make_stem(464, 768, dev, seed=21)
    B, T, n_var = 2, 384, 80
    cond = make_input(B, 1, 464 - n_var, seed=22)[:, 0].contiguous()
    hidden = torch.empty(B, T, 768, dtype=torch.bfloat16, device=dev)
    bias3 = torch.empty(B, 3, 768, device=dev)
    stem.forward_split(torch.zeros(B, T, n_var, dtype=torch.bfloat16, device=dev), cond.to(dev), hidden=hidden,
                       bias_scratch=bias3)
    want3 = conv_stem_oracle.folded_bias(cond, stem.conv1.weight, stem.conv1.bias, n_var)
    assert float((bias3.cpu().double() - want3).abs().max()) < 2e-5           # fp32 sums of 384 exact products per tap
    w = stem.conv1.weight.detach().cpu().double()[:, n_var:, :]               # [D, n_cond, 3]
    s = torch.einsum("nct,bc->btn", w, cond.double())                          # [B, 3, D]
    bias = stem.conv1.bias.detach().cpu().double()
    gelu = lambda v: torch.nn.functional.gelu(v.to(torch.bfloat16).float()).to(torch.bfloat16)
    want_mid = gelu((bias + s.sum(1)).float())
    want_first = gelu((bias + s[:, 1] + s[:, 2]).float())
    want_last = gelu((bias + s[:, 0] + s[:, 1]).float())
    got = hidden.cpu()
    assert_close(got[:, 0], want_first, "first frame", 0.05)
    assert_close(got[:, T - 1], want_last, "last frame", 0.05)
    for t in (1, 2, 127, 128, 255, 256, T - 2):
        assert_close(got[:, t], want_mid, f"interior frame {t}", 0.05)
    assert torch.equal(got[:, 1], got[:, 200]) and torch.equal(got[:, 1], got[:, T - 2])
    assert float((want_first.float() - want_mid.float()).abs().max()) > 0.05   # the edges really differ


def test_split_stem_on_the_frontend_output_matches_the_materialised_chain(dev):
    """C5 slice at full context: frontend (bf16 frames, dense) -> split stem against frontend -> encoder
    input [B, 4096, 464] -> stem, 6 windows."""
    from beatheritage_b200 import MelSpectrogram
    mel = MelSpectrogram("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect").to(dev)
    stem = make_stem(464, 768, dev, seed=31)
    g = torch.Generator().manual_seed(32)
    samples = (torch.rand(6, 524160, generator=g) * 2 - 1).to(dev)
    conds = [(torch.randn(6, 128, generator=g) * 0.5).to(dev) for _ in range(3)]
    enc_in = mel.forward_encoder_input(samples, conds, dtype=torch.bfloat16)
    want = stem(enc_in)
    frames = torch.empty(6, 4096, 80, dtype=torch.bfloat16, device=dev)
    mel.forward_into(samples, frames)
    assert torch.equal(frames, enc_in[:, :, :80])
    cond = torch.cat([c.to(torch.bfloat16) for c in conds], dim=1)
    got = stem.forward_split(frames, cond)
    torch.cuda.synchronize()
    assert_close(got, want, "split stem vs stem over the encoder input", 0.10)
    one = conv_stem_oracle.conv_stem(enc_in[2:3].cpu(), stem.conv1.weight, stem.conv1.bias, stem.conv2.weight,
                                     stem.conv2.bias)
    assert_close(got[2:3], one, "split stem vs oracle, window 2", 0.10)


def test_split_errors(dev):
    import ctypes
    from beatheritage_b200 import _stem_lib
    stem = make_stem(464, 768, dev, seed=41)
    frames = torch.zeros(1, 64, 80, dtype=torch.bfloat16, device=dev)
    cond = torch.zeros(1, 384, dtype=torch.bfloat16, device=dev)
    with pytest.raises(RuntimeError):
        stem.forward_split(frames, cond[:, :380])                       # channel counts do not add up
    with pytest.raises(RuntimeError):
        stem.forward_split(frames.float(), cond)
    with pytest.raises(RuntimeError):
        stem.forward_split(frames[:, :63], cond)                        # odd T
    with pytest.raises(RuntimeError):
        stem.forward_split(frames.cpu(), cond.cpu())
    lib = _stem_lib.lib()
    h = stem._handle_for(dev)
    buf = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
    rc = lib.bhstem_forward_split(h, frames.data_ptr(), cond.data_ptr(), 1, 64, buf.data_ptr(), buf.data_ptr(),
                                  buf.data_ptr(), None)
    assert rc == 1 and b"prepare_split" in lib.bhstem_last_error()     # not prepared yet
    assert lib.bhstem_prepare_split(h, 84) == 1 and lib.bhstem_prepare_split(h, 464) == 1
    stem.forward_split(frames, cond)
    assert lib.bhstem_prepare_split(h, 80) == 0                         # idempotent
    assert lib.bhstem_prepare_split(h, 88) == 1                         # but one n_var per handle
    with pytest.raises(RuntimeError):
        stem.forward_split(torch.zeros(1, 64, 88, dtype=torch.bfloat16, device=dev),
                           torch.zeros(1, 376, dtype=torch.bfloat16, device=dev))


def test_cuda_graph_capture_of_the_split_serving_chain(dev):
    """The split form of the serving chain -- frontend writing bf16 frames, folded bias, split conv1, conv2:
    four launches -- captured once and replayed on new audio gives the eager bits, and stays within the bf16
    tolerance of the chain over the materialised encoder input."""
    from beatheritage_b200 import MelSpectrogram
    from tests.golden import signals
    mel = MelSpectrogram("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect").to(dev)
    stem = make_stem(464, 768, dev, seed=6)
    torch.manual_seed(3)
    cond = torch.randn(1, 384, device=dev)
    cond16 = cond.to(torch.bfloat16)
    x = torch.zeros(1, 524160, device=dev)
    songs = [torch.from_numpy(signals.music(524160, seed=s).reshape(1, 524160)).to(dev) for s in (31, 32)]
    frames = torch.empty(1, 4096, 80, dtype=torch.bfloat16, device=dev)
    hidden = torch.empty(1, 4096, 768, dtype=torch.bfloat16, device=dev)
    out = torch.empty(1, 2048, 768, dtype=torch.bfloat16, device=dev)

    def chain():
        return stem.forward_split(mel.forward_into(x, frames), cond16, hidden=hidden, out=out)

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        x.copy_(songs[0])
        for _ in range(3):
            chain()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y = chain()
    assert y.data_ptr() == out.data_ptr()
    for song in songs:
        x.copy_(song)
        graph.replay()
        replayed = out.clone()
        eager = chain().clone()
        full = stem(mel.forward_encoder_input(x, [cond], dtype=torch.bfloat16, channels_first=False)).clone()
        torch.cuda.synchronize()
        assert torch.equal(replayed, eager)
        assert_close(replayed, full, "split chain vs chain over the encoder input", 0.10)


def test_split_stem_is_reentrant_across_threads_and_streams(dev):
    """One ConvStem (one handle) driven by two host threads on their own streams with different conditioning
    vectors: the folded bias lives in a per-call scratch, the handle is read-only after bhstem_prepare_split."""
    import threading
    stem = make_stem(464, 768, dev, seed=51)
    xs = [make_input(2, 512, 464, seed=s) for s in (52, 53)]
    frames = [x[:, :, :80].contiguous().to(dev) for x in xs]
    conds = [x[:, 0, 80:].contiguous().to(dev) for x in xs]
    refs = [stem.forward_split(f, c).clone() for f, c in zip(frames, conds)]
    torch.cuda.synchronize()
    assert not torch.equal(refs[0], refs[1])
    errors = []

    def worker(f, c, ref):
        try:
            s = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(s):
                for _ in range(40):
                    if not torch.equal(stem.forward_split(f, c), ref):
                        errors.append("mismatch")
            s.synchronize()
        except Exception as e:   # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=a) for a in zip(frames, conds, refs)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert errors == []


def test_small_batch_tile_width_gives_the_same_bits(dev):
    """One window leaves most SMs without a 256-column tile, so conv2 (and anything smaller) runs 128-column tiles
    by default (BHSTEM_OPT_SMALL_BATCH_TILES): same products in the same order, identical results."""
    stem = make_stem(464, 768, dev, seed=61)
    x = make_input(1, 4096, 464, seed=62).to(dev)
    frames, cond = x[:, :, :80].contiguous(), x[:, 0, 80:].contiguous()
    got = (stem(x).clone(), stem.forward_split(frames, cond).clone(), stem(x[:, :200]).clone())
    stem.set_small_batch_tiles(False)
    want = (stem(x).clone(), stem.forward_split(frames, cond).clone(), stem(x[:, :200]).clone())
    torch.cuda.synchronize()
    for g, w in zip(got, want):
        assert torch.equal(g, w)


@pytest.mark.parametrize("B,T", [(2, 512), (1, 200), (3, 4096)])
def test_kernel_variants_agree_on_the_split_stem(dev, B, T):
    """Every schedule runs the split conv1 with the same per-window bias addressing: the CTA-pair kernel with 16 and
    8 epilogue warps and the one-CTA shared-tap kernel give identical bits, the one-box-per-tap kernel (tap-major
    summation) the stem's usual tolerance."""
    stem = make_stem(464, 768, dev, seed=4)
    x = make_input(B, T, 464, seed=B * 11 + T)
    frames, cond = x[:, :, :80].contiguous().to(dev), x[:, 0, 80:].contiguous().to(dev)
    outs = {}
    for variant in ("cta_pairs", "shared_taps", "tap_boxes"):
        stem.set_variant(variant)
        outs[variant] = stem.forward_split(frames, cond).clone()
    stem.set_variant("cta_pairs")
    stem.set_epilogue_warps(8, 8, 8)
    outs["cta_pairs_8"] = stem.forward_split(frames, cond).clone()
    torch.cuda.synchronize()
    assert torch.equal(outs["cta_pairs"], outs["shared_taps"]) and torch.equal(outs["cta_pairs"], outs["cta_pairs_8"])
    assert_close(outs["tap_boxes"], outs["shared_taps"], "tap_boxes vs shared_taps (split)", max_frac=0.10)


def test_plain_c_host_program_drives_the_split_serving_chain(dev, tmp_path):
    """examples/c_abi_stem_demo.c: a C program (no Python, no torch) runs frontend -> stem through both C ABIs in
    the full and the split form and checks that they agree within the stem's tolerance."""
    import os
    import shutil
    import subprocess
    from tests.conftest import ROOT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    from beatheritage_b200 import build as _build
    _build.build()
    exe = str(tmp_path / "c_abi_stem_demo")
    libdir = os.path.join(ROOT, "beatheritage_b200")
    subprocess.run([nvcc, "-x", "cu", os.path.join(ROOT, "examples", "c_abi_stem_demo.c"), "-I", os.path.join(ROOT, "include"),
                    "-L", libdir, "-lbhmel", "-lbhstem", "-Xlinker", f"-rpath={libdir}", "-o", exe], check=True,
                   capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "frames per window: 512" in res.stdout and "split vs full" in res.stdout
