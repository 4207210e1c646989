#!/usr/bin/env python3
"""A small end-to-end run for compute-sanitizer (memcheck / racecheck): 2 frames, 2 scales, non-square input."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VNECT_B200_NO_GRAPH"] = "1"
from vnect_b200 import VNectEngine  # noqa: E402
from vnect_b200.weights import seeded_init  # noqa: E402

eng = VNectEngine(seeded_init("W1"), [1.0, 0.7], max_frames=2, max_streams=2, max_input=(300, 420))
frames = np.random.default_rng(0).integers(0, 256, (2, 300, 420, 3), dtype=np.uint8)
for k in range(2):
    j2, j3 = eng.estimate(frames, [0, 1], [1.0 + k, 1.0 + k], [1.01 + k, 1.01 + k])
eng.set_box(0, (10, 20, 200, 250))
eng.set_box(1, (0, 0, 420, 300))
j2, j3, boxes = eng.track(frames, [0, 1], [5.0, 5.0], [5.01, 5.01])
print("ok", float(np.abs(j3).max()), boxes.tolist())
eng.close()

# 5 frames x 2 scales = 10 forwards: past the small-batch plan, so the CTA-pair (cta_group::2) and 256-column kernels run
eng = VNectEngine(seeded_init("W1"), [1.0, 0.7], max_frames=5, max_streams=5)
frames = np.random.default_rng(1).integers(0, 256, (5, 368, 368, 3), dtype=np.uint8)
j2, j3 = eng.estimate(frames, list(range(5)), [1.0] * 5, [1.01] * 5)
print("ok pairs", float(np.abs(j3).max()), eng.launch_count())
eng.close()
