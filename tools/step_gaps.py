#!/usr/bin/env python3
"""Where a bench step's time goes: 20 back-to-back steps (CUDA graph replay) against the forward, pre- and post-process
timed on their own in the same process, alternating so that clocks and temperature are shared."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vnect_b200 import VNectEngine  # noqa: E402
from vnect_b200.weights import seeded_init  # noqa: E402

nf = 64
eng = VNectEngine(seeded_init("W0"), [1.0, 0.7], max_frames=nf, max_streams=nf)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng.set_cuda_stream(stream.cuda_stream)
frames = torch.from_numpy(np.stack([np.random.default_rng(1000 + i).integers(0, 256, (368, 368, 3), dtype=np.uint8)
                                    for i in range(nf)])).cuda()
j2 = torch.empty((nf, 21, 2), dtype=torch.float64, device="cuda")
j3 = torch.empty((nf, 21, 3), dtype=torch.float32, device="cuda")
t = [1000.0]


def step():
    t[0] += 1 / 30
    eng.estimate_device(frames.data_ptr(), nf, 368, 368, j2.data_ptr(), j3.data_ptr(), np.arange(nf), np.full(nf, t[0]),
                        np.full(nf, t[0] + 0.004))


for _ in range(5):
    step()
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / 20
    fwd = eng.time_forward(2 * nf, reps=5, per_layer=False)
    pre, post = eng.time_prepost(nf, reps=5)
    t[0] += 10.0  # time_prepost advanced the streams' clocks on its own
    print(f"rep {rep}: step {ms_step*1e3:.1f} us = forward {fwd*1e3:.1f} + pre {pre*1e3:.1f} + post {post*1e3:.1f} "
          f"+ rest {(ms_step - fwd - pre - post)*1e3:.1f} us")
eng.close()
