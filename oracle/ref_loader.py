"""Loads the UNMODIFIED reference module by path -- build-container only.

/root/reference does not exist on the GPU box; only `tests/golden/make_golden.py` and the
CPU-side oracle pinning tests (which skip when the tree is absent) may call this.
"""
from __future__ import annotations

import importlib.util
import os

REF_FILE = "/root/reference/osuT5/osuT5/model/spectrogram.py"


def available() -> bool:
    return os.path.exists(REF_FILE)


def load_reference_class():
    spec = importlib.util.spec_from_file_location("_bh_ref_spectrogram", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MelSpectrogram
