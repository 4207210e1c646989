#!/usr/bin/env python3
"""Per-layer times at 32 and 128 forwards (same engine): what an L2-resident quarter batch would cost per layer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vnect_b200 import VNectEngine
from vnect_b200.weights import seeded_init
eng = VNectEngine(seeded_init("W0"), [1.0, 0.7], max_frames=64, max_streams=64)
t32, p32 = eng.time_forward(32, reps=5, per_layer=True)
t128, p128 = eng.time_forward(128, reps=5, per_layer=True)
print(f"forward 32: {t32*1e3:.1f} us, forward 128: {t128*1e3:.1f} us")
s32 = s128 = 0.0
for k in p128:
    if k.startswith("res4") or k.startswith("res5"):
        s32 += p32[k]; s128 += p128[k]
    print(f"  {k:44s} 32: {p32[k]*1e3:7.1f} us  x4 = {4*p32[k]*1e3:7.1f}   128: {p128[k]*1e3:7.1f} us")
print(f"res4 + res5 layers: 4 x t(32) = {4*s32*1e3:.1f} us vs t(128) = {s128*1e3:.1f} us")
