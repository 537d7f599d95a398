"""CPU: host logic of icl_b200/lanes.py and of the two-node backbone wiring (no GPU: lanes are disabled on CPU tensors, every
helper must degrade to plain sequential execution)."""
import torch

from icl_b200 import lanes


def test_lanes_disabled_on_cpu():
    dev = torch.device("cpu")
    assert not lanes.enabled(dev)
    # no-ops without streams
    lanes.handoff(None, None, torch.zeros(1))
    with lanes.on(None):
        pass
    lanes.join_all()
    order = []
    out = lanes.fan(dev, "t", [lambda i=i: order.append(i) or i * 2 for i in range(3)], [torch.zeros(1)])
    assert out == [0, 2, 4] and order == [0, 1, 2]


def test_env_switch(monkeypatch):
    class FakeDev:
        type = "cuda"
    monkeypatch.setenv("ICL_HEAD_LANES", "0")
    assert not lanes.enabled(FakeDev())
    monkeypatch.setenv("ICL_HEAD_LANES", "1")
    assert lanes.enabled(FakeDev())


def test_backbone_param_split_matches_block_order():
    """BackbonePairLowFn takes the parameters of conv1..up_concat3 (28 tensors), BackbonePairTopFn those of up_concat2, up_concat1 and
    final (10): the split must follow _backbone_params() / param_names()."""
    from icl_b200.networks import backbone3d as bb
    names = bb.param_names()
    assert len(names) == 38 and bb.N_LOW == 28
    assert all(n.split(".conv")[0] in ("conv1", "conv2", "conv3", "conv4", "center", "up_concat4", "up_concat3") for n in names[:bb.N_LOW])
    assert [n.split(".")[0] for n in names[bb.N_LOW:]] == ["up_concat2"] * 4 + ["up_concat1"] * 4 + ["final"] * 2
    assert bb.LOW_BLOCKS + bb.TOP_BLOCKS == bb.PARAM_BLOCKS
