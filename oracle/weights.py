"""Seeded random-init VNect weights in the reference's interchange format.

The reference loads a pickled ``{tf_variable_name: ndarray}`` dict (``src/vnect_model.py:219-236``) whose naming and
layouts come from ``src/caffe2pkl.py:57-76``:

* ``tc.layers.conv2d`` scope X  -> ``X/weights`` [kh, kw, Cin, Cout] and ``X/biases`` [Cout]
* ``tf.layers.*``               -> ``X/kernel`` (transposed convs are [kh, kw, Cout, Cin]; res5c_branch2c is [1,1,128,84])
* batch norm                    -> ``bn5c_branch2a/{gamma,beta,moving_mean,moving_variance}`` [128]

The real weights are not distributed with the reference, so parity is defined on these two seeded sets:

* ``W0``: what TF1 would initialise -- Xavier/Glorot uniform kernels, zero biases, identity batch norm.
* ``W1``: W0 kernels plus random biases and batch-norm statistics, so every parameter path is exercised.

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import numpy as np

# (scope, kernel, Cin, Cout) for every tc.layers.conv2d of src/vnect_model.py:27-211, in graph order.
CONV_SCOPES = [("conv1", 7, 3, 64)]


def _bottleneck(prefix, cin, mid, cout, proj, suffix=""):
    out = []
    if proj:
        out.append((f"{prefix}_branch1{suffix}", 1, cin, cout))
    out.append((f"{prefix}_branch2a{suffix}", 1, cin, mid))
    out.append((f"{prefix}_branch2b{suffix}", 3, mid, mid))
    out.append((f"{prefix}_branch2c{suffix}", 1, mid, cout))
    return out


CONV_SCOPES += _bottleneck("res2a", 64, 64, 256, True)
CONV_SCOPES += _bottleneck("res2b", 256, 64, 256, False)
CONV_SCOPES += _bottleneck("res2c", 256, 64, 256, False)
CONV_SCOPES += _bottleneck("res3a", 256, 128, 512, True)
for _b in "bcd":
    CONV_SCOPES += _bottleneck("res3" + _b, 512, 128, 512, False)
CONV_SCOPES += _bottleneck("res4a", 512, 256, 1024, True)
for _b in "bcdef":
    CONV_SCOPES += _bottleneck("res4" + _b, 1024, 256, 1024, False)
CONV_SCOPES += _bottleneck("res5a", 1024, 512, 1024, True, "_new")
CONV_SCOPES += [
    ("res5b_branch2a_new", 1, 1024, 256),
    ("res5b_branch2b_new", 3, 256, 128),
    ("res5b_branch2c_new", 1, 128, 256),
    ("res5c_branch2b", 3, 212, 128),
]
# tf.layers.* variables: (scope, shape)
KERNEL_SCOPES = [
    ("res5c_branch1a", (4, 4, 63, 256)),
    ("res5c_branch2a", (4, 4, 128, 256)),
    ("res5c_branch2c", (1, 1, 128, 84)),
]
BN_SCOPE = "bn5c_branch2a"


def variable_shapes():
    """{tf variable name: shape} for the 109 variables of the graph."""
    shapes = {}
    for scope, k, cin, cout in CONV_SCOPES:
        shapes[scope + "/weights"] = (k, k, cin, cout)
        shapes[scope + "/biases"] = (cout,)
    for scope, shp in KERNEL_SCOPES:
        shapes[scope + "/kernel"] = shp
    for v in ("gamma", "beta", "moving_mean", "moving_variance"):
        shapes[f"{BN_SCOPE}/{v}"] = (128,)
    return shapes


def _xavier(rng, shape):
    rf = int(np.prod(shape[:-2]))
    fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def make_weights(kind="W0", seed=0):
    """Return the weight dict (float32 arrays).  kind: 'W0' (TF-faithful init) or 'W1' (all parameters random)."""
    assert kind in ("W0", "W1")
    rng = np.random.default_rng(seed)
    w = {}
    for name, shp in variable_shapes().items():
        leaf = name.split("/")[1]
        if leaf in ("weights", "kernel"):
            w[name] = _xavier(rng, shp)
        elif leaf in ("biases", "beta", "moving_mean"):
            w[name] = np.zeros(shp, np.float32)
        else:  # gamma, moving_variance
            w[name] = np.ones(shp, np.float32)
    if kind == "W1":
        rng1 = np.random.default_rng(seed + 7919)
        for name, shp in variable_shapes().items():
            leaf = name.split("/")[1]
            if leaf == "biases":
                w[name] = rng1.uniform(-0.05, 0.05, shp).astype(np.float32)
            elif leaf in ("beta", "moving_mean"):
                w[name] = rng1.uniform(-0.1, 0.1, shp).astype(np.float32)
            elif leaf in ("gamma", "moving_variance"):
                w[name] = rng1.uniform(0.5, 1.5, shp).astype(np.float32)
    return w
