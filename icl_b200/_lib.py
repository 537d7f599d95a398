"""ctypes binding of the C-ABI in include/icl_b200.h (no torch C++ ABI dependency).

The product path has NO fallback: if libicl_b200.so is missing or a call fails, this raises.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libicl_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "icl_b200.h")

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "icl_b200: CUDA extension %s not built (run `python -m icl_b200.build`); there is no CPU fallback" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.icl_last_error.restype = ctypes.c_char_p
        _lib.icl_launch_count.restype = ctypes.c_ulonglong
        _declare(_lib)
    return _lib


_CT = {
    "int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float, "double": ctypes.c_double,
    "unsigned long long": ctypes.c_ulonglong,
}


_RET = {}


def header_prototypes():
    """Parse include/icl_b200.h -> {name: [(ctype, argname), ...]} (pointers map to c_void_p)."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(unsigned long long|long long|int|const char\*)\s+(icl_\w+)\s*\(([^)]*)\)\s*;", txt):
        name, args = m.group(2), m.group(3).strip()
        _RET[name] = m.group(1)
        sig = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    sig.append((ctypes.c_void_p, a.split("*")[-1].strip()))
                else:
                    parts = a.rsplit(" ", 1)
                    sig.append((_CT[parts[0].replace("const ", "")], parts[1]))
        protos[name] = sig
    return protos


def _declare(l):
    for name, sig in header_prototypes().items():
        fn = getattr(l, name)
        fn.argtypes = [t for t, _ in sig]
        if name not in ("icl_last_error", "icl_launch_count"):
            fn.restype = ctypes.c_longlong if _RET.get(name) == "long long" else ctypes.c_int


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_PROFILE = None  # list of (name, start_event, end_event, gflop, mbytes) while profiling


def call(name, *args, gflop=0.0, mbytes=0.0, tag=""):
    """Invoke an int-returning entry point on the current stream; raise on failure.
    gflop / mbytes: algorithmic work of this launch, recorded only while profiling (bench.py roofline)."""
    l = lib()
    if _PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(l, name)(*args, stream())
        e1.record()
        _PROFILE.append((name, e0, e1, gflop, mbytes, tag))
    else:
        rc = getattr(l, name)(*args, stream())
    if rc != 0:
        raise RuntimeError("icl_b200.%s failed (%d): %s" % (name, rc, l.icl_last_error().decode()))


def profile_start():
    global _PROFILE
    _PROFILE = []


def profile_stop(n_steps=1):
    """Aggregate per entry point: CUDA-event time on the launching stream, launches, algorithmic GFLOP / MB."""
    global _PROFILE
    rec, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    agg = {}
    detail = {}
    for name, e0, e1, gf, mb, tag in rec:
        t = e0.elapsed_time(e1)
        for d, key in ((agg, name), (detail, name + ("[%s]" % tag if tag else ""))):
            a = d.setdefault(key, [0.0, 0, 0.0, 0.0])
            a[0] += t
            a[1] += 1
            a[2] += gf
            a[3] += mb
    total = sum(a[0] for a in agg.values()) or 1.0
    ks = []
    for name, (ms, n, gf, mb) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        ks.append({"name": name, "ms_per_step": ms / n_steps, "launches_per_step": n / n_steps, "share": ms / total,
                   "ms_per_launch": ms / n, "gflop_per_launch": gf / n, "mbytes_per_launch": mb / n,
                   "tflops": (gf / ms) if ms and gf else None, "gbs": (mb / ms) if ms and mb else None})
    det = [{"name": k, "ms_per_step": v[0] / n_steps, "launches_per_step": v[1] / n_steps, "gflop_per_launch": v[2] / v[1],
            "mbytes_per_launch": v[3] / v[1]} for k, v in sorted(detail.items(), key=lambda kv: -kv[1][0])]
    return {"kernels": ks, "kernel_ms_per_step": total / n_steps, "detail": det}


def launch_count():
    return int(lib().icl_launch_count())
