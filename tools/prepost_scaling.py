#!/usr/bin/env python3
"""pyramid_kernel / postprocess_kernel at larger batches: us per launch and GB/s by algorithmic bytes (SURVEY.md 8d:
406 272 B in + 2 x 1 083 392 B out per frame for the pyramid, 355 488 B of heat-map planes per frame for the
post-process), CUDA events around single launches (vnect_time_prepost)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vnect_b200 import VNectEngine  # noqa: E402
from vnect_b200.weights import seeded_init  # noqa: E402

for nf in (64, 256, 512):
    eng = VNectEngine(seeded_init("W0"), [1.0, 0.7], max_frames=nf, max_streams=nf)
    frames = np.stack([np.random.default_rng(1000 + i).integers(0, 256, (368, 368, 3), dtype=np.uint8) for i in range(min(nf, 64))])
    frames = np.concatenate([frames] * (nf // frames.shape[0]))
    ids = np.arange(nf)
    eng.estimate(frames, ids, np.full(nf, 4.0), np.full(nf, 4.004))   # real maps in place
    pre, post = eng.time_prepost(nf, reps=10)
    pre_b = nf * (406272 + 2 * 1083392)
    post_b = nf * 355488
    print(f"{nf} frames: pyramid {pre*1e3:.1f} us = {pre_b/pre/1e6:.0f} GB/s, post-process {post*1e3:.1f} us = {post_b/post/1e6:.0f} GB/s")
    eng.close()
