"""Side streams ("lanes") for the independent branches of the ICL heads.

The three scales of an InherentConsistent pass (6^3 / 12^3 / 24^3 tokens), the query chain that links them, and the uscl
pass are mostly independent of each other and of the upper decoder levels of the backbone (up_concat2 / up_concat1 / final
work on 48^3 / 96^3 and never feed the heads).  Issued on one stream they are ~650 kernels of 2-10 us that the GPU runs one
after the other; on lanes they run next to each other and next to the backbone's tensor-core kernels.  Values do not change:
every kernel sees the same inputs and reductions keep their fixed order; two calls of the same BatchNorm module always share a
lane, so the running-statistics updates keep the reference's order.

Inside a CUDA-graph capture the lanes fork from / join the capturing stream through events, so the captured step is a DAG with
the same edges.  Autograd runs each backward node on the stream of its forward op and inserts the cross-stream waits itself.

ICL_HEAD_LANES=0 disables the lanes (everything on the current stream)."""
import contextlib
import os

import torch

_STREAMS = {}


def enabled(device):
    return device.type == "cuda" and os.environ.get("ICL_HEAD_LANES", "1") != "0"


def get(device, names, priority=0):
    """{name: stream} of cached side streams on `device` (+ "main": the current stream).  Lanes run at the default priority; the
    captured step's own stream is created with a higher one (icl_b200/graph.py), so the backbone chain — the critical path — is
    scheduled ahead of the lanes' kernels (measured: 12.19 ms vs 12.30 ms with the priorities the other way round)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    out = {"main": torch.cuda.current_stream(device)}
    for n in names:
        s = _STREAMS.get((idx, n))
        if s is None:
            if "ICL_LANE_PRIORITY" in os.environ:   # measurement knob
                priority = int(os.environ["ICL_LANE_PRIORITY"])
            s = _STREAMS[(idx, n)] = torch.cuda.Stream(device=device, priority=priority)
        out[n] = s
    return out


def on(stream):
    return contextlib.nullcontext() if stream is None else torch.cuda.stream(stream)


def handoff(dst, src, *tensors):
    """`dst` continues after everything enqueued on `src` so far; `tensors` (allocated on src) will be read on dst: tell the caching
    allocator, so that their memory is not handed out again on src while dst still reads it.  No-op without lanes."""
    if dst is None or src is None or dst == src:
        return
    dst.wait_stream(src)
    for t in tensors:
        if t is not None and t.is_cuda:
            t.record_stream(dst)


def join_all(device=None):
    """The current stream waits for everything enqueued so far on every lane of `device` (cheap: one event per lane).  Used where
    gradients produced on the lanes are read outside autograd's own stream bookkeeping (gradient buckets, the fused optimizer's
    factor lists).  Inside a CUDA-graph capture only lanes that are part of the capture are waited for."""
    if not torch.cuda.is_available():
        return
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    cur = torch.cuda.current_stream(idx)
    capturing = torch.cuda.is_current_stream_capturing()
    for (d, _), s in _STREAMS.items():
        if d != idx or s == cur:
            continue
        if capturing:
            with torch.cuda.stream(s):
                if not torch.cuda.is_current_stream_capturing():
                    continue
        cur.wait_stream(s)


def fan(device, prefix, thunks, tensors=()):
    """Runs the independent thunks on lanes prefix0, prefix1, ... forked from the current stream and joins them: [results].
    `tensors`: inputs produced elsewhere that the thunks read (announced to the allocator per lane).  Results that are tensors (or
    tuples of tensors) are announced for the current stream.  Sequential without lanes or with a single thunk."""
    if len(thunks) < 2 or not enabled(device):
        return [t() for t in thunks]
    L = get(device, ["%s%d" % (prefix, i) for i in range(len(thunks))])
    main = L["main"]
    out = []
    for i, t in enumerate(thunks):
        lane = L["%s%d" % (prefix, i)]
        handoff(lane, main, *tensors)
        with on(lane):
            r = t()
        out.append(r)
        handoff(main, lane, *(r if isinstance(r, (tuple, list)) else (r,)))
    return out
