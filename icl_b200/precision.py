"""Tensor-core operand precision of the 3x3x3 convolutions.

"parity" (default): every bf16 operand is split hi+lo and three tcgen05 MMAs (hi*hi, hi*lo, lo*hi)
    accumulate in fp32 — ~16 mantissa bits, meets the 1e-3 / 99.9 % contract (SURVEY.md §0 surprise #2).
"fast": single bf16 MMA; ~4e-2 relative logit error at random init (stated, measured in tests).
"fp32": no tensor cores — every convolution runs the fp32 CUDA-core kernels (conv3d_direct.cu).  Slow; it exists so the
    parity tests can separate "the algorithm is the reference's" (fp32 arithmetic, ~1e-7 rounding) from the ~4e-6 operand
    rounding of "parity", which moves a few pre-activations across zero and so flips (Leaky)ReLU derivatives.
"""
_MODE = "parity"


def set_precision(mode):
    global _MODE
    if mode not in ("parity", "fast", "fp32"):
        raise ValueError("precision must be 'parity', 'fast' or 'fp32'")
    _MODE = mode


def get_precision():
    return _MODE


def planes():
    return 1 if _MODE == "fast" else 2


def tensor_cores():
    return _MODE != "fp32"
