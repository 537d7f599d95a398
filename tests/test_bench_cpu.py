"""CPU: the pure bookkeeping of bench.py (which roofline the dominant kernel is measured against, workload selection)."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

PK = {"tf_sus": 1392.7, "hbm": 6546.2, "src": "test"}


def _top(**kw):
    d = {"name": "icl_sgd_factored_apply", "gflop_per_launch": 12.2, "mbytes_per_launch": 3057.6, "ms_per_launch": 0.5325, "share": 0.13,
         "launches_per_step": 4.0}
    d.update(kw)
    return d


def test_fused_update_hbm_bound_at_small_rank():
    r = bench.roofline_of(_top(), PK, 3.0)   # R = 32: 16 B per parameter in 0.53 ms
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - 3057.6 / 0.5325 / 6546.2) < 1e-9
    assert r["traffic"] is None or r["traffic"]["algorithmic_bytes"] > 0


def test_fused_update_tensor_bound_at_large_rank():
    r = bench.roofline_of(_top(gflop_per_launch=391.4, ms_per_launch=0.9755), PK, 3.0)   # R = 1 024 (8 ranks)
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert abs(r["achieved"] - 391.4 / 0.9755) < 1e-9 and abs(r["mma_issue_frac"] - 3.0 * r["frac"]) < 1e-12
    assert r["traffic"] is None   # the DRAM capture on file is the R = 32 one


def test_conv_kernel_is_tensor_bound():
    r = bench.roofline_of({"name": "icl_conv3d_umma_walk_fwd", "gflop_per_launch": 48.9, "mbytes_per_launch": 0.0, "ms_per_launch": 0.223,
                           "share": 0.1, "launches_per_step": 3.0}, PK, 3.0)
    assert r["bound"] == "tensor" and 0.1 < r["frac"] < 0.2


def test_default_workload_follows_baseline_configs():
    a = types.SimpleNamespace(workload="")
    assert bench.default_workload(a, 1) == "cfg2" and bench.default_workload(a, 8) == "cfg3"
    assert bench.default_workload(types.SimpleNamespace(workload="infer"), 4) == "infer"
