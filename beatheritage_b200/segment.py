"""Window producers either side of the frontend, as index arithmetic instead of copies.

`SegmentPlan` restates what Preprocessor.__init__/segment compute (reference
osuT5/osuT5/inference/preprocessor.py:12-21, 58-71, 94-102): the model-context window length
(src_seq_len - 1) * hop, the stride int(window * (1 - lookback - lookahead)) -- with its float
truncation -- and the right padding that makes the strided windows tile the song.  The plan is
what `MelSpectrogram.forward_gather` / `bhmel_forward_gather` consume, so the [W, window]
batch (10x duplicated audio in sequential mode) is never materialised.

`dataset_window_plan` restates the training-side framing (reference
osuT5/osuT5/dataset/ors_dataset.py:243-262 `_get_frames`, :303-308 window starts,
:563-590 `_pad_frame_sequence`).

`shard_range` is the static partition used when songs/windows are spread over the GPUs of one
box (no collective on the data path).
"""
from __future__ import annotations

from dataclasses import dataclass

HOP = 128


@dataclass(frozen=True)
class SegmentPlan:
    window_len: int        # samples per sequence
    stride: int            # samples between window starts
    first_offset: int      # start of window 0 inside the (begin-padded) song
    n_windows: int
    padded_len: int        # song length after segment()'s right padding (zeros)

    @property
    def frames_per_window(self) -> int:
        return self.window_len // HOP + 1


def segment_plan(n_samples: int, src_seq_len: int = 4096, hop_length: int = HOP, lookback: float = 0.5,
                 lookahead: float = 0.4, parallel: bool = False) -> SegmentPlan:
    """Plan for a song of n_samples (after any begin/end padding the caller applied)."""
    window = (src_seq_len - 1) * hop_length                       # preprocessor.py:14-17
    stride = int(window * (1 - lookback - lookahead))             # preprocessor.py:18
    if parallel:
        stride = window                                           # preprocessor.py:20-21
    if n_samples < window:                                        # preprocessor.py:61-63
        padded = window
    else:                                                         # preprocessor.py:64-67
        rem = (n_samples - window) % stride
        padded = n_samples + (0 if rem == 0 else stride - rem)
    # window(): as_strided rows 0 .. padded - window, every `stride`-th (preprocessor.py:94-98)
    n_windows = (padded - window) // stride + 1
    return SegmentPlan(window, stride, 0, n_windows, padded)


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
