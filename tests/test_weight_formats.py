"""Weight sources of the drop-in: the reference's params.pkl, a numpy archive and the
TensorFlow-1.x checkpoint of src/estimator.py:55-60, all 109 variables incl. the dead res2c_branch2a/* pair."""
import os
import struct

import numpy as np
import pytest

from vnect_b200 import tf_checkpoint, weights


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors
    assert tf_checkpoint.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert tf_checkpoint.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tf_checkpoint.crc32c(bytes(range(32))) == 0x46DD794E
    assert tf_checkpoint.crc32c(b"123456789") == 0xE3069283
    data = np.random.default_rng(0).integers(0, 256, 100003, dtype=np.uint8).tobytes()
    c = 0xFFFFFFFF
    for b in data[:4099]:
        c = tf_checkpoint._TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    assert tf_checkpoint.crc32c(data[:4099]) == c ^ 0xFFFFFFFF   # native helper == table-driven definition
    assert tf_checkpoint.mask_crc(0) == 0xa282ead8


@pytest.mark.parametrize("fmt", ["pkl", "npz", "ckpt_prefix", "ckpt_dir"])
def test_every_format_round_trips_all_109_variables(tmp_path, fmt):
    w = weights.seeded_init("W1", seed=3)
    assert len(w) == 109 and "res2c_branch2a/weights" in w
    if fmt == "pkl":
        path = str(tmp_path / "params.pkl")
        weights.save(path, w)
    elif fmt == "npz":
        path = str(tmp_path / "params.npz")
        weights.save(path, w)
    else:
        weights.save(str(tmp_path / "vnect_tf"), w)
        assert os.path.isfile(tmp_path / "vnect_tf.index") and os.path.isfile(tmp_path / "vnect_tf.data-00000-of-00001")
        assert tf_checkpoint.latest_checkpoint(str(tmp_path)) == str(tmp_path / "vnect_tf")
        path = str(tmp_path / "vnect_tf") if fmt == "ckpt_prefix" else str(tmp_path)
    got = weights.check_complete(weights.resolve(path))
    assert got.keys() == w.keys()
    for k in w:
        assert got[k].dtype == np.float32 and np.array_equal(got[k], w[k]), k


def test_checkpoint_index_structure_and_corruption(tmp_path):
    w = {"conv1/weights": np.arange(7 * 7 * 3 * 64, dtype=np.float32).reshape(7, 7, 3, 64),
         "conv1/biases": np.linspace(-1, 1, 64).astype(np.float32), "scalar": np.float32(3.5)}
    prefix = str(tmp_path / "m")
    tf_checkpoint.write_checkpoint(prefix, w)
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack_from("<Q", raw, len(raw) - 8)[0] == tf_checkpoint.TABLE_MAGIC and len(raw) > 48
    header, entries = tf_checkpoint.read_index(prefix + ".index")
    assert header == {"num_shards": 1, "endianness": 0}
    assert entries["conv1/weights"]["shape"] == (7, 7, 3, 64) and entries["conv1/weights"]["dtype"] == tf_checkpoint.DT_FLOAT
    assert entries["conv1/biases"]["offset"] == 0 and entries["conv1/weights"]["offset"] == 64 * 4   # sorted by name
    got = tf_checkpoint.read_checkpoint(prefix)
    assert got["scalar"].shape == () and got["scalar"] == np.float32(3.5)
    # a flipped bit in the tensor data or in the index is detected
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[100] ^= 0x10
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="checksum"):
        tf_checkpoint.read_checkpoint(prefix)
    assert np.array_equal(tf_checkpoint.read_checkpoint(prefix, verify=False)["conv1/weights"], w["conv1/weights"])
    idx = bytearray(raw)
    idx[10] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError):
        tf_checkpoint.read_index(prefix + ".index")
    open(prefix + ".index", "wb").write(b"not a table" * 10)
    with pytest.raises(ValueError, match="magic"):
        tf_checkpoint.read_index(prefix + ".index")


def test_incomplete_or_misshapen_weights_are_refused(tmp_path):
    w = weights.seeded_init("W0")
    bad = dict(w)
    del bad["res2c_branch2a/biases"]          # dead in the graph, but part of the reference's variable set
    with pytest.raises(KeyError):
        weights.check_complete(bad)
    bad = dict(w)
    bad["res4a_branch1/weights"] = bad["res4a_branch1/weights"][..., :512]
    with pytest.raises(KeyError):
        weights.check_complete(bad)
    extra = dict(w)
    extra["global_step"] = np.zeros((), np.float32)
    assert len(weights.check_complete(extra)) == 109


def test_default_locations_prefer_the_tf_checkpoint(tmp_path, monkeypatch):
    monkeypatch.delenv("VNECT_B200_WEIGHTS", raising=False)
    monkeypatch.chdir(tmp_path)
    os.makedirs(tmp_path / "models" / "tf_model")
    os.makedirs(tmp_path / "models" / "caffe_model")
    w = {k: v for k, v in list(weights.seeded_init("W0").items())[:4]}
    weights.save(str(tmp_path / "models" / "caffe_model" / "params.pkl"), {k: v + 1 for k, v in w.items()})
    got = weights.resolve(None)
    assert all(np.array_equal(got[k], w[k] + 1) for k in w)       # only the pickle exists (init_weights.py:35-36)
    weights.save(str(tmp_path / "models" / "tf_model" / "vnect_tf"), w)
    got = weights.resolve(None)
    assert all(np.array_equal(got[k], w[k]) for k in w)           # the checkpoint wins (src/estimator.py:55-60)
