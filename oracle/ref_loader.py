"""Loads the UNMODIFIED reference module by path -- build-container only.

/root/reference does not exist on the GPU box; only `tests/golden/make_golden.py` and the
CPU-side oracle pinning tests (which skip when the tree is absent) may call load_reference_class.
bench.py's CPU legs use load_shipped_reference_class (the copy build() leaves in oracle/_ref).
"""
from __future__ import annotations

import importlib.util
import os

REF_FILE = "/root/reference/osuT5/osuT5/model/spectrogram.py"
# the same file, copied unmodified by __graft_entry__.build() (git-ignored; travels to the GPU box)
SHIPPED_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "spectrogram.py")


def available() -> bool:
    return os.path.exists(REF_FILE)


def load_shipped_reference_class():
    """The reference's own module from oracle/_ref (bench.py's reference arm / cpu_baseline on the GPU
    box, where /root/reference does not exist).  Raises when the copy is absent."""
    spec = importlib.util.spec_from_file_location("_bh_ref_spectrogram_shipped", SHIPPED_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MelSpectrogram


def load_reference_class():
    spec = importlib.util.spec_from_file_location("_bh_ref_spectrogram", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MelSpectrogram
