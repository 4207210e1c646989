#!/usr/bin/env python3
"""Per-layer CUDA-event times of the batch-1 latency plan (max_frames = 1: 2 forwards, 64-column tiles, single CTAs)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from vnect_b200 import VNectEngine
from vnect_b200.weights import seeded_init
eng = VNectEngine(seeded_init("W0"), [1.0, 0.7], max_frames=1, max_streams=1)
tot, per = eng.time_forward(2, reps=20, per_layer=True)
print("forward(2) total ms", tot, "sum of layers", sum(per.values()))
for k, v in per.items():
    print(f"  {k:40s} {v*1e3:8.1f} us")
