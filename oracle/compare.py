"""Comparison rules shared by the GPU parity tests and ``__graft_entry__.smoke()``.

TEST INFRASTRUCTURE (see oracle/__init__.py).

BASELINE.json's bar for the 2D joints is "argmax indices bit-exact except for documented ties".  The CUDA CNN computes
with fp16 operands (maps within 3e-3 of the fp32 restatement), so where the x8-upsampled heat-map has two samples
closer than that error the argmax may legitimately land on the other one.  ``explain_argmax_diffs`` therefore accepts a
difference only if the ORACLE's own upsampled map (utils.py:169-171, cv2) at the CUDA position is within
``rel_bound * max|hm|`` of the oracle's maximum, and returns one record per such tie so the caller can log it.
"""
import cv2
import numpy as np

REL_BOUND = 3e-3  # the asserted CNN error bound (normwise-max of the maps, tests/test_gpu_parity.py)


def explain_argmax_diffs(raw_gpu, ref_est, rel_bound=REL_BOUND, label=""):
    """raw_gpu: [21,2] integer (row, col) argmax of the CUDA path in box pixels; ref_est: an OracleEstimator that has
    just processed the same frame.  Raises AssertionError for a difference that is not a near-tie; returns a list of
    dicts (joint, gpu, ref, gap, bound) for the accepted ones."""
    hm = ref_est.last["hm_avg"]
    raw_ref = ref_est.last["joints_2d_raw"].astype(int)
    raw_gpu = np.asarray(raw_gpu).astype(int)
    bound = rel_bound * float(np.abs(hm).max())
    ties = []
    for j in range(hm.shape[2]):
        if tuple(raw_gpu[j]) == tuple(raw_ref[j]):
            continue
        up = cv2.resize(hm[:, :, j], (0, 0), fx=ref_est.hm_factor, fy=ref_est.hm_factor, interpolation=cv2.INTER_LINEAR)
        gap = float(up.max() - up[raw_gpu[j][0], raw_gpu[j][1]])
        rec = dict(joint=j, gpu=tuple(int(v) for v in raw_gpu[j]), ref=tuple(int(v) for v in raw_ref[j]), gap=gap,
                   bound=bound)
        assert gap <= bound, f"{label}: argmax of joint {j} differs and is NOT a near-tie: {rec}"
        ties.append(rec)
    return ties


def format_ties(ties, label=""):
    return "; ".join(f"{label} joint {t['joint']}: gpu {t['gpu']} vs ref {t['ref']}, oracle gap {t['gap']:.3e} <= bound "
                     f"{t['bound']:.3e}" for t in ties)


class StreamComparer:
    """Frame-by-frame comparison of one video stream (filters live) against the oracle.

    A near-tie moves one joint's raw argmax; from then on that joint's filter state -- and, if it is the root joint 14,
    every root-relative 3D joint -- legitimately differs from the oracle's, so the joint is 'tainted' for the rest of
    the stream.  Untainted joints must match: 2D bit-exact (same float64 filter arithmetic on the same integers),
    3D within ``mm`` (BASELINE.json: 1 mm).  Every accepted tie is printed."""

    def __init__(self, label="", mm=1.0, rel_bound=REL_BOUND):
        self.label, self.mm, self.rel_bound = label, mm, rel_bound
        self.tainted = set()
        self.ties = []

    def check(self, k, raw_gpu, j2_gpu, j3_gpu, ref_est, r2, r3, j2_shift=(0.0, 0.0)):
        """j2_gpu / r2 in the same coordinates (pass j2_shift if the caller already added a crop origin to r2)."""
        ties = explain_argmax_diffs(raw_gpu, ref_est, self.rel_bound, f"{self.label} frame {k}")
        if ties:
            print("near-tie accepted:", format_ties(ties, f"{self.label} frame {k}"))
        self.ties += ties
        self.tainted |= {t["joint"] for t in ties}
        ok = np.array([j not in self.tainted for j in range(21)])
        assert np.array_equal(np.asarray(j2_gpu)[ok], np.asarray(r2)[ok]), f"{self.label} frame {k}: untainted 2D joints differ"
        if 14 not in self.tainted:
            d = np.abs(np.asarray(j3_gpu, np.float64)[ok] - np.asarray(r3, np.float64)[ok]).max()
            assert d < self.mm, f"{self.label} frame {k}: 3D joints differ by {d:.3f} mm"
        return ties
