"""Tensor-core operand precision of the 3x3x3 convolutions.

"parity" (default): every bf16 operand is split hi+lo and three tcgen05 MMAs (hi*hi, hi*lo, lo*hi)
    accumulate in fp32 — ~16 mantissa bits, meets the 1e-3 / 99.9 % contract (SURVEY.md §0 surprise #2).
"fast": single bf16 MMA; ~4e-2 relative logit error at random init (stated, measured in tests).
"""
_MODE = "parity"


def set_precision(mode):
    global _MODE
    if mode not in ("parity", "fast"):
        raise ValueError("precision must be 'parity' or 'fast'")
    _MODE = mode


def get_precision():
    return _MODE


def planes():
    return 2 if _MODE == "parity" else 1
