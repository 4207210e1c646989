"""Weight sources for the estimator.

The reference restores a TF checkpoint made from a pickled ``{tf_variable_name: ndarray}`` dict
(src/vnect_model.py:219-236, src/caffe2pkl.py:57-88, init_weights.py).  The trained weights are not distributed, so
beside that pickle format this module offers the seeded random initialisation TF1 itself would perform
(Xavier/Glorot-uniform kernels, zero biases, identity batch norm) -- the weights every parity number is quoted on.
"""
import os
import pickle

import numpy as np

DEFAULT_PICKLES = ("./models/caffe_model/params.pkl", "../models/caffe_model/params.pkl")  # init_weights.py:35-36
DEFAULT_CHECKPOINT_DIRS = ("./models/tf_model/", "../models/tf_model/")  # src/estimator.py:55-60


def _conv_scopes():
    s = [("conv1", 7, 3, 64)]

    def block(pre, cin, mid, cout, proj, suf=""):
        out = [(f"{pre}_branch1{suf}", 1, cin, cout)] if proj else []
        return out + [(f"{pre}_branch2a{suf}", 1, cin, mid), (f"{pre}_branch2b{suf}", 3, mid, mid),
                      (f"{pre}_branch2c{suf}", 1, mid, cout)]

    s += block("res2a", 64, 64, 256, True) + block("res2b", 256, 64, 256, False) + block("res2c", 256, 64, 256, False)
    s += block("res3a", 256, 128, 512, True)
    for b in "bcd":
        s += block("res3" + b, 512, 128, 512, False)
    s += block("res4a", 512, 256, 1024, True)
    for b in "bcdef":
        s += block("res4" + b, 1024, 256, 1024, False)
    s += block("res5a", 1024, 512, 1024, True, "_new")
    s += [("res5b_branch2a_new", 1, 1024, 256), ("res5b_branch2b_new", 3, 256, 128),
          ("res5b_branch2c_new", 1, 128, 256), ("res5c_branch2b", 3, 212, 128)]
    return s


def variable_shapes():
    """Names and shapes of the 109 variables of the reference graph, in graph order."""
    shapes = {}
    for scope, k, cin, cout in _conv_scopes():
        shapes[scope + "/weights"] = (k, k, cin, cout)
        shapes[scope + "/biases"] = (cout,)
    shapes["res5c_branch1a/kernel"] = (4, 4, 63, 256)
    shapes["res5c_branch2a/kernel"] = (4, 4, 128, 256)
    shapes["res5c_branch2c/kernel"] = (1, 1, 128, 84)
    for v in ("gamma", "beta", "moving_mean", "moving_variance"):
        shapes["bn5c_branch2a/" + v] = (128,)
    return shapes


def seeded_init(kind="W0", seed=0):
    """Seeded TF1-style initialisation.  'W0': Glorot-uniform kernels, zero biases, identity BN.  'W1': the same
    kernels with random biases / BN statistics (exercises every parameter path)."""
    if kind not in ("W0", "W1"):
        raise ValueError("kind must be 'W0' or 'W1'")
    rng = np.random.default_rng(seed)
    w = {}
    for name, shp in variable_shapes().items():
        leaf = name.split("/")[1]
        if leaf in ("weights", "kernel"):
            rf = int(np.prod(shp[:-2]))
            lim = np.sqrt(6.0 / (shp[-2] * rf + shp[-1] * rf))
            w[name] = rng.uniform(-lim, lim, size=shp).astype(np.float32)
        elif leaf in ("biases", "beta", "moving_mean"):
            w[name] = np.zeros(shp, np.float32)
        else:
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


def load_pickle(path):
    """The reference's params.pkl (src/caffe2pkl.py:83-88): {name: ndarray}."""
    with open(path, "rb") as f:
        d = pickle.load(f)
    return {k: np.asarray(v, dtype=np.float32) for k, v in d.items()}


def load_npz(path):
    """A numpy archive {name: ndarray} (np.savez of the params dict)."""
    with np.load(path, allow_pickle=False) as z:
        return {k: np.asarray(z[k], dtype=np.float32) for k in z.files}


def load_tf_checkpoint(path):
    """The TensorFlow-1.x checkpoint the reference restores (src/estimator.py:55-60): ``path`` is the directory holding
    the ``checkpoint`` state file, or a checkpoint prefix such as ./models/tf_model/vnect_tf."""
    from . import tf_checkpoint
    prefix = path
    if os.path.isdir(path):
        prefix = tf_checkpoint.latest_checkpoint(path)
        if prefix is None:
            raise FileNotFoundError("no 'checkpoint' state file in %s" % path)
    for suffix in (".index", ".meta"):
        if prefix.endswith(suffix):
            prefix = prefix[:-len(suffix)]
    return tf_checkpoint.read_checkpoint(prefix)


def save(path, wdict):
    """Write a weight dict in the format the suffix names: .pkl (the reference's params.pkl), .npz, or a TensorFlow
    checkpoint prefix (anything else)."""
    if path.endswith(".pkl"):
        with open(path, "wb") as f:
            pickle.dump({k: np.asarray(v) for k, v in wdict.items()}, f)
    elif path.endswith(".npz"):
        np.savez(path, **wdict)
    else:
        from . import tf_checkpoint
        tf_checkpoint.write_checkpoint(path, wdict)


def check_complete(wdict):
    """All 109 variables of the reference graph with their shapes (incl. the dead res2c_branch2a/*, which the
    checkpoint holds and the graph ignores); extra keys (saver bookkeeping) are dropped."""
    out = {}
    for name, shp in variable_shapes().items():
        if name not in wdict:
            raise KeyError("weights lack variable '%s'" % name)
        arr = np.asarray(wdict[name], dtype=np.float32)
        if tuple(arr.shape) != tuple(shp):
            raise KeyError("variable '%s' has shape %s, expected %s" % (name, tuple(arr.shape), tuple(shp)))
        out[name] = arr
    return out


def resolve(spec=None):
    """spec: dict | params.pkl | .npz | TensorFlow checkpoint (directory or prefix) | 'random:W0[:seed]' | None (env
    VNECT_B200_WEIGHTS, then the reference's default locations: the TF checkpoint of src/estimator.py:55-60, then the
    pickle of init_weights.py:35-36).  Raises FileNotFoundError when nothing is found, like the reference does when its
    checkpoint is absent."""
    if isinstance(spec, dict):
        return spec
    if spec is None:
        spec = os.environ.get("VNECT_B200_WEIGHTS")
    if spec is None:
        for d in DEFAULT_CHECKPOINT_DIRS:
            if os.path.isfile(os.path.join(d, "checkpoint")):
                spec = d
                break
    if spec is None:
        for p in DEFAULT_PICKLES:
            if os.path.isfile(p):
                spec = p
                break
    if spec is None:
        raise FileNotFoundError(
            "no VNect weights: pass weights=<dict | params.pkl | .npz | TF checkpoint | 'random:W0'> or set "
            "VNECT_B200_WEIGHTS (the reference's trained weights are not distributed; see models/*/README.md there)")
    spec = str(spec)
    if spec.startswith("random:"):
        parts = spec.split(":")
        return seeded_init(parts[1], int(parts[2]) if len(parts) > 2 else 0)
    if spec.endswith(".npz"):
        return load_npz(spec)
    if spec.endswith(".pkl"):
        return load_pickle(spec)
    if os.path.isdir(spec) or os.path.isfile(spec + ".index") or spec.endswith((".index", ".meta")):
        return load_tf_checkpoint(spec)
    return load_pickle(spec)
