"""CNN restatement (oracle.forward) pinned against the only goldens the reference holds for it: the blob and parameter
SHAPES printed in materials/caffe_script.ipynb, plus the algorithmic FLOP count of BASELINE.md."""
import hashlib

import numpy as np

from oracle.forward import OracleNet, algorithmic_flops
from oracle.weights import make_weights, variable_shapes

# materials/caffe_script.ipynb cell 3 (N,C,H,W) -> NHWC spatial/channels at 368 input
BLOB_SHAPES = {
    "conv1": (184, 64), "pool1": (92, 64), "res2a": (92, 256), "res2b": (92, 256), "res2c": (92, 256),
    "res3a": (46, 512), "res3d": (46, 512), "res4a": (23, 1024), "res4f": (23, 1024), "res5a": (23, 1024),
    "res5b_branch2c_new": (23, 256), "res5c_branch2a_feat": (46, 212), "res5c_branch2b": (46, 128),
}


def test_variable_inventory():
    shapes = variable_shapes()
    assert len(shapes) == 109  # 51 x {weights,biases} + 3 kernels + 4 BN
    assert shapes["conv1/weights"] == (7, 7, 3, 64)
    assert shapes["res5c_branch1a/kernel"] == (4, 4, 63, 256)
    assert shapes["res5c_branch2a/kernel"] == (4, 4, 128, 256)
    assert shapes["res5c_branch2c/kernel"] == (1, 1, 128, 84)
    assert shapes["res5c_branch2b/weights"] == (3, 3, 212, 128)
    # "~14.6 M weights + ~22 k biases"; exact counts of the 55 kernels and of all 109 variables
    assert sum(int(np.prod(s)) for n, s in shapes.items() if n.endswith(("weights", "kernel"))) == 14_596_288
    assert sum(int(np.prod(s)) for s in shapes.values()) == 14_615_936


def test_weights_are_deterministic():
    a, b = make_weights("W0"), make_weights("W0")
    assert all(np.array_equal(a[k], b[k]) for k in a)
    h = hashlib.sha256(b"".join(np.ascontiguousarray(a[k]).tobytes() for k in sorted(a))).hexdigest()
    assert h == hashlib.sha256(b"".join(np.ascontiguousarray(b[k]).tobytes() for k in sorted(b))).hexdigest()
    w1 = make_weights("W1")
    assert np.array_equal(w1["conv1/weights"], a["conv1/weights"]) and np.abs(w1["conv1/biases"]).max() > 0


def test_flops_match_baseline():
    assert algorithmic_flops(368) == 23_830_290_432
    assert algorithmic_flops(448) == 35_317_481_472


def test_forward_shapes_and_dead_branch(w0):
    x = (np.random.default_rng(0).integers(0, 256, (1, 368, 368, 3)).astype(np.float32) / 255 - 0.4)
    outs, taps = OracleNet(w0)(x, want_taps=True)
    assert all(o.shape == (1, 46, 46, 21) and o.dtype == np.float32 for o in outs)
    for name, (hw, c) in BLOB_SHAPES.items():
        assert taps[name].shape == (1, hw, hw, c), name
    # res2c_branch2a is dead (src/vnect_model.py:56 feeds res2b_branch2a into res2c_branch2b)
    w = dict(w0)
    w["res2c_branch2a/weights"] = w["res2c_branch2a/weights"] * 0 + 7
    outs2 = OracleNet(w)(x)
    assert all(np.array_equal(a, b) for a, b in zip(outs, outs2))


def test_fp64_restatement_agrees(w0):
    import torch
    x = (np.random.default_rng(1).integers(0, 256, (1, 368, 368, 3)).astype(np.float32) / 255 - 0.4)
    o32 = OracleNet(w0)(x)
    o64 = OracleNet(w0, dtype=torch.float64)(x)
    for a, b in zip(o32, o64):
        assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-5
