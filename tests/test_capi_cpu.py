"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/vnect_b200.h
declares, the product never routes through the oracle, and the host logic (weights, sharding) behaves."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vnect_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(vnect_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from vnect_b200 import _capi
    lib = _capi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 18
    assert set(declared) == set(_capi.SIGNATURES), "ctypes table and header disagree"
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.vnect_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vnect_b200 import VNectEngine
    with pytest.raises(RuntimeError, match="no CUDA device"):
        VNectEngine("random:W0")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vnect_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    code = "import sys; import vnect_b200, vnect_b200.parallel; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_seeded_init_matches_oracle_weights():
    from oracle.weights import make_weights, variable_shapes
    from vnect_b200 import weights
    assert weights.variable_shapes() == variable_shapes()
    for kind in ("W0", "W1"):
        a, b = weights.seeded_init(kind), make_weights(kind)
        assert a.keys() == b.keys()
        assert all(np.array_equal(a[k], b[k]) for k in a)


def test_weight_resolution(tmp_path, monkeypatch):
    import pickle
    from vnect_b200 import weights
    monkeypatch.delenv("VNECT_B200_WEIGHTS", raising=False)
    monkeypatch.chdir(tmp_path)
    with pytest.raises(FileNotFoundError):
        weights.resolve(None)
    w = weights.resolve("random:W0")
    p = tmp_path / "params.pkl"
    with open(p, "wb") as f:
        pickle.dump({k: v for k, v in list(w.items())[:3]}, f)
    got = weights.resolve(str(p))
    assert len(got) == 3 and all(v.dtype == np.float32 for v in got.values())
    monkeypatch.setenv("VNECT_B200_WEIGHTS", "random:W1:3")
    assert np.abs(weights.resolve(None)["conv1/biases"]).max() > 0


def test_config_struct_layout_matches_header():
    from vnect_b200 import _capi
    # int32 x3, pad, double x4, int32 x5 (+pad) -- the C compiler's layout for vnect_config
    assert ctypes.sizeof(_capi.Config) == 72
    assert _capi.Config.scales.offset == 16 and _capi.Config.max_frames.offset == 48


def test_stream_sharding_is_a_partition():
    from vnect_b200 import parallel
    for n, w in ((256, 8), (256, 2), (7, 4), (64, 1)):
        seen = sorted(s for r in range(w) for s in parallel.owned_streams(n, r, w))
        assert seen == list(range(n))
    j2 = np.random.default_rng(0).uniform(0, 368, (5, 21, 2))
    j3 = np.random.default_rng(1).uniform(-500, 500, (5, 21, 3)).astype(np.float32)
    a, b = parallel.unpack_results(parallel.pack_results(j2, j3))
    assert np.array_equal(a, j2) and np.array_equal(b, j3)
