"""sac_cot_b200 — B200-native SAC-COT registration hot path (hand-written sm_100a CUDA behind
the C ABI of include/sac_cot.h).  No CPU fallback: importing works anywhere, but every compute
entry point raises unless sac_cot_b200/lib/libsaccot.so is built and a B200 is visible."""
from . import _abi, synth  # noqa: F401
from .api import Group, Registrar, Result, SacCotError, load_library, register, register_batch  # noqa: F401

__all__ = ["Group", "Registrar", "Result", "SacCotError", "load_library", "register", "register_batch", "synth"]
