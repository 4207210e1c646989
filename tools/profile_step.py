#!/usr/bin/env python3
"""Minimal driver for ncu: W warm-up steps + K steps of the C2 workload (64 frames, 2 scales) on device-resident
frames.  One step = 55 kernel launches (pyramid, 51 conv GEMMs, pool, post-process...)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vnect_b200 import VNectEngine  # noqa: E402
from vnect_b200.weights import seeded_init  # noqa: E402

warm, steps, nf = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 64
eng = VNectEngine(seeded_init("W0"), [1.0, 0.7], max_frames=nf, max_streams=nf)
frames = torch.from_numpy(np.stack([np.random.default_rng(1000 + i).integers(0, 256, (368, 368, 3), dtype=np.uint8)
                                    for i in range(nf)])).cuda()
j2 = torch.empty((nf, 21, 2), dtype=torch.float64, device="cuda")
j3 = torch.empty((nf, 21, 3), dtype=torch.float32, device="cuda")
for k in range(warm + steps):
    t = 1000 + k / 30
    eng.estimate_device(frames.data_ptr(), nf, 368, 368, j2.data_ptr(), j3.data_ptr(), np.arange(nf), np.full(nf, t),
                        np.full(nf, t + 0.004))
eng.synchronize()
print("launches", eng.launch_count())
