#!/usr/bin/env python3
"""The other BASELINE.json configs on one GPU (bench.py measures C2): C3 latency stream, C4 many streams over time,
C5 448x448 3-scale stress.  Prints one JSON object; results are copied to profiles/ by hand."""
import json
import os
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vnect_b200 import VNectEngine  # noqa: E402
from vnect_b200.weights import seeded_init  # noqa: E402

W = seeded_init("W0")
out = {}


def noise(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


# ---- C3: one 960x540 video stream, batch 1, scales [1.0, 0.7], filters + bbox tracker on the device, t_k = 1000 + k/25
eng = VNectEngine(W, [1.0, 0.7], max_frames=1, max_streams=1, max_input=(540, 960))
eng.set_box(0, (0, 0, 960, 540))
base = noise(7, 540, 960)
lat = []
for k in range(120):
    frame = np.ascontiguousarray(np.roll(base, (k, 2 * k), axis=(0, 1)))
    t0 = time.perf_counter()
    eng.track(frame, [0], [1000 + k / 25], [1000 + k / 25 + 0.004])
    lat.append((time.perf_counter() - t0) * 1e3)
lat = sorted(lat[20:])
out["C3_video_stream_960x540_batch1_tracked"] = {"latency_ms_p50": statistics.median(lat), "latency_ms_p95": lat[int(0.95 * (len(lat) - 1))],
                                                 "frames_per_s_single_stream": 1e3 / statistics.median(lat)}
eng.close()

# ---- C4 (per-GPU share): 128 streams x T = 32 steps, 368x368, 2 scales, pipelined host API
ns, T = 128, 32
eng = VNectEngine(W, [1.0, 0.7], max_frames=ns, max_streams=ns)
frames = torch.empty((ns, 368, 368, 3), dtype=torch.uint8).pin_memory()
fn = frames.numpy()
for s in range(ns):
    fn[s] = noise(2000 + s, 368, 368)
outs = [(torch.empty((ns, 21, 2), dtype=torch.float64).pin_memory().numpy(), torch.empty((ns, 21, 3), dtype=torch.float32).pin_memory().numpy()) for _ in range(2)]
ids = np.arange(ns)
for rep in range(2):
    eng.reset()
    t0 = time.perf_counter()
    for k in range(T):
        lane = k & 1
        if k >= 2:
            eng.wait(lane)
        tk = 1000 + k / 30
        eng.submit(lane, fn, ids, np.full(ns, tk), np.full(ns, tk + 0.004), out=outs[lane])
    eng.wait(0)
    eng.wait(1)
    dt = time.perf_counter() - t0
out["C4_128_streams_x_32_steps_per_gpu"] = {"frames_per_s": ns * T / dt, "ms_per_step": dt / T * 1e3}
eng.close()

# ---- C5: 448x448, scales [1, 0.85, 0.7], 128 frames per GPU per step (384 forwards)
nf = 128
eng = VNectEngine(W, [1, 0.85, 0.7], box_size=448, max_frames=nf, max_streams=nf)
frames = torch.empty((nf, 448, 448, 3), dtype=torch.uint8).pin_memory()
fn = frames.numpy()
for s in range(nf):
    fn[s] = noise(3000 + s, 448, 448)
outs = [(torch.empty((nf, 21, 2), dtype=torch.float64).pin_memory().numpy(), torch.empty((nf, 21, 3), dtype=torch.float32).pin_memory().numpy()) for _ in range(2)]
ids = np.arange(nf)
steps = 8
for rep in range(2):
    t0 = time.perf_counter()
    for k in range(steps):
        lane = k & 1
        if k >= 2:
            eng.wait(lane)
        tk = 10.0 * rep + 1 + k / 30
        eng.submit(lane, fn, ids, np.full(nf, tk), np.full(nf, tk + 0.004), out=outs[lane])
    eng.wait(0)
    eng.wait(1)
    dt = time.perf_counter() - t0
fwd_ms = eng.time_forward(nf * 3, reps=2)
out["C5_448x448_3scale_128_frames_per_gpu"] = {"frames_per_s_e2e": nf * steps / dt, "ms_per_step": dt / steps * 1e3,
                                               "forward_ms_384_images": fwd_ms,
                                               "tflops_algorithmic": 35317481472 * 384 / (fwd_ms * 1e-3) / 1e12}
eng.close()
print(json.dumps(out))
