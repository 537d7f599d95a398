"""Stand-in for monai.utils (MONAI 1.0.1): ensure_tuple_rep / optional_import."""


def ensure_tuple_rep(t, dim):
    return tuple(t) if isinstance(t, (list, tuple)) else (t,) * dim


def optional_import(*a, **k):
    return None, False
