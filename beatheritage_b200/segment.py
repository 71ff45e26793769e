"""Window producers either side of the frontend, as index arithmetic instead of copies.

`SegmentPlan` restates what Preprocessor.__init__/segment compute (reference
osuT5/osuT5/inference/preprocessor.py:12-25, 41-102): the model-context window length
(src_seq_len - 1) * hop, the stride int(window * (1 - lookback - lookahead)) -- with its float
truncation -- the right padding that makes the strided windows tile the song, `sequence_times`
(float32 arange truncated to int32) and the start_time / end_time trimming with its keep-one rules.  The plan is
what `MelSpectrogram.forward_gather` / `bhmel_forward_gather` consume, so the [W, window]
batch (10x duplicated audio in sequential mode) is never materialised.

`dataset_window_plan` restates the training-side framing (reference
osuT5/osuT5/dataset/ors_dataset.py:243-262 `_get_frames`, :303-308 window starts,
:563-590 `_pad_frame_sequence`).

`shard_range` is the static partition used when songs/windows are spread over the GPUs of one
box (no collective on the data path).
"""
from __future__ import annotations

import bisect
import math
import struct
from dataclasses import dataclass

HOP = 128


@dataclass(frozen=True)
class SegmentPlan:
    window_len: int        # samples per sequence
    stride: int            # samples between window starts
    first_offset: int      # start of the first KEPT window inside the (begin-padded) song
    n_windows: int         # windows kept (after start_time / end_time trimming)
    padded_len: int        # song length after segment()'s right padding (zeros)
    sequence_times: tuple = ()     # int32 start time in ms of every kept window (preprocessor.py:72-73)
    first_window: int = 0          # index of the first kept window among the untrimmed ones
    song_length_ms: float = 0.0    # len(samples) / sample_rate * 1000 before any padding (preprocessor.py:57)

    @property
    def frames_per_window(self) -> int:
        return self.window_len // HOP + 1


def _f32(x: float) -> float:
    """Round a Python float to float32 (what a float32 torch tensor element holds)."""
    return struct.unpack("f", struct.pack("f", x))[0]


def segment_plan(n_samples: int, src_seq_len: int = 4096, hop_length: int = HOP, lookback: float = 0.5,
                 lookahead: float = 0.4, parallel: bool = False, sample_rate: int = 16000,
                 start_time: float | None = None, end_time: float | None = None,
                 begin_pad: int = 0, end_pad: int = 0) -> SegmentPlan:
    """Everything Preprocessor.segment returns, as indices: which windows exist, which survive the
    start_time / end_time trimming (keep-one rules included), their start times and the song length.
    `n_samples` is the raw song; `begin_pad` / `end_pad` are segment()'s own arguments -- the caller of
    forward_gather pads (or offsets into) the resident song accordingly."""
    window = (src_seq_len - 1) * hop_length                       # preprocessor.py:14-17
    stride = int(window * (1 - lookback - lookahead))             # preprocessor.py:18
    if parallel:
        stride = window                                           # preprocessor.py:20-21
    song_length_ms = n_samples / sample_rate * 1000               # preprocessor.py:57
    n_total = n_samples + begin_pad + end_pad                     # preprocessor.py:58
    if n_total < window:                                          # preprocessor.py:60-62
        padded = window
    else:                                                         # preprocessor.py:63-66
        rem = (n_total - window) % stride
        padded = n_total + (0 if rem == 0 else stride - rem)
    # window(): as_strided rows 0 .. padded - window, every `stride`-th (preprocessor.py:94-98)
    n_windows = (padded - window) // stride + 1
    ms_per_stride = stride * 1000 / sample_rate                   # preprocessor.py:22
    ms_per_sequence = window * 1000 / sample_rate                 # preprocessor.py:23
    # torch.arange(0, W * ms, ms).to(int32): float32 elements, truncated (preprocessor.py:72-73)
    n_times = math.ceil((n_windows * ms_per_stride) / ms_per_stride)
    times = [int(_f32(i * ms_per_stride)) for i in range(n_times)]
    first = 0
    if start_time is not None:                                    # preprocessor.py:75-82
        first = bisect.bisect_right(times, start_time - (1 - lookahead) * ms_per_sequence)
        if first == len(times):
            first -= 1
        times = times[first:]
    if end_time is not None:                                      # preprocessor.py:83-90
        kept = bisect.bisect_left(times, end_time - lookback * ms_per_sequence)
        if kept == 0:
            kept = 1
        times = times[:kept]
    return SegmentPlan(window, stride, first * stride, len(times), padded, tuple(times), first, song_length_ms)


def dataset_window_plan(n_samples: int, src_seq_len: int = 4096, hop_length: int = HOP, offset: int = 0,
                        gen_start_frame: int = 0) -> SegmentPlan:
    """Training windows: the song is padded to a hop multiple (a whole extra hop when already
    aligned), cut into (src_seq_len - 1)-hop-frame windows starting at hop-frame `offset`, the last
    one zero padded.  Returned as a gather plan (stride == window_len)."""
    fsl = src_seq_len - 1
    n_frames = (n_samples + (hop_length - n_samples % hop_length)) // hop_length     # ors_dataset.py:256-257
    starts = range(offset, n_frames - gen_start_frame, fsl)                           # ors_dataset.py:307
    window = fsl * hop_length
    return SegmentPlan(window, window, offset * hop_length, len(starts), n_frames * hop_length)


def shard_range(n_items: int, rank: int, world_size: int) -> range:
    """Round-robin (by item index) static shard: item i belongs to rank i % world_size.  Mirrors
    the reference's per-worker slicing of the track list (osuT5/osuT5/utils/model_utils.py:256-269)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return range(rank, n_items, world_size)
