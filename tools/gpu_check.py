#!/usr/bin/env python3
"""Developer diagnostics on a B200: layer-by-layer parity against the CPU oracle plus per-layer timings.
(Test infrastructure; the pytest -m gpu suite is the gate, this prints the detail.)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import prepost, synth  # noqa: E402
from oracle.forward import OracleNet  # noqa: E402
from oracle.weights import make_weights  # noqa: E402
from vnect_b200 import VNectEngine  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b) + 1e-30))


def main():
    what = sys.argv[1:] or ["pre", "fwd", "post", "e2e", "time"]
    scales = [1.0, 0.7]
    w = make_weights("W1")
    t = time.time()
    eng = VNectEngine(w, scales, max_frames=64, max_streams=64, max_input=(540, 960))
    print(f"engine up in {time.time()-t:.1f}s; steps={len(eng.step_names())} gemm_flops/fwd={eng.info('gemm_flops_per_forward'):.4g}")
    net = OracleNet(w)

    if "pre" in what:
        for (h, wd, seed) in [(368, 368, 1), (540, 960, 2), (300, 200, 3), (736, 736, 4), (368, 200, 5)]:
            img = np.random.default_rng(seed).integers(0, 256, (h, wd, 3), dtype=np.uint8)
            if h > 540 or wd > 960:
                e2 = VNectEngine(False, scales, max_frames=1, max_input=(h, wd))
                got, sc, off = e2.preprocess(img)
                e2.close()
            else:
                got, sc, off = eng.preprocess(img)
            ref, rsc, roff = prepost.gen_input_batch(img, 368, scales)
            ref16 = ref.astype(np.float16).astype(np.float32)
            print(f"pre {h}x{wd}: mismatches={int((got != ref16).sum())} of {got.size}; scaler {sc == rsc} offs {off == roff}")

    if "fwd" in what:
        x = np.stack([prepost.gen_input_batch(synth.frame_c2(i), 368, [1.0])[0][0] for i in range(2)])
        outs = eng.forward(x)
        (rhm, rxm, rym, rzm), taps = net(x, want_taps=True)
        order = ["pool1", "res2a_branch2a", "res2a_branch2b", "res2a", "res2b_branch2a", "res2b", "res2c",
                 "res3a", "res3b", "res3c", "res3d", "res4a", "res4b", "res4f", "res5a", "res5b_branch2c_new",
                 "res5c_branch2a_feat", "res5c_branch2b"]
        for name in order:
            got = eng.tap(name, 2)
            ref = taps[name]
            if name in ("res2b", "res2c", "res3d"):
                ref = ref[:, ::2, ::2, :]
            if name == "res5c_branch2a_feat":
                got = got[..., :212]
            print(f"  tap {name:24s} shape {got.shape} rel-L2 {rel(got, ref):.3e} max|d| {np.abs(got-ref).max():.3e} ref-std {ref.std():.3e}")
        for nm, g, r in zip("hm xm ym zm".split(), outs, (rhm, rxm, rym, rzm)):
            am = (g.reshape(2, -1, 21).argmax(1) == r.reshape(2, -1, 21).argmax(1)).mean()
            print(f"  out {nm}: rel-L2 {rel(g, r):.3e} normwise-max {np.abs(g-r).max()/np.abs(r).max():.3e} 46x46 argmax agree {am:.3f}")

    if "post" in what:
        for seed in range(4):
            maps = synth.synthetic_maps(500 + seed, 2, border_joints=(seed % 2 == 0))
            eng.reset()
            ts = [(100.0 + 0.033 * k, 100.004 + 0.033 * k) for k in range(3)]
            clock_vals = []
            ref = prepost.OracleEstimator(lambda b: cur["m"], scales, clock=lambda: clock_vals.pop(0))
            cur = {}
            for k, (t2, t3) in enumerate(ts):
                cur["m"] = synth.synthetic_maps(500 + seed + 10 * k, 2, border_joints=(seed % 2 == 0))
                clock_vals[:] = [t2, t3]
                r2, r3 = ref(np.zeros((368, 368, 3), np.uint8))
                j2, j3, raw = eng.postprocess(cur["m"], 1.0, (0, 0), [0], [t2], [t3])
                print(f"post seed {seed} frame {k}: raw argmax equal {np.array_equal(raw[0], ref.last['joints_2d_raw'])} "
                      f"j2 bit-equal {np.array_equal(j2[0], r2)} max|dj2| {np.abs(j2[0]-r2).max():.3e} "
                      f"j3 bit-equal {np.array_equal(j3[0], r3)} max|dj3| {np.abs(j3[0]-r3).max():.3e}")

    if "e2e" in what:
        eng.reset()
        clock_vals = []
        ref = prepost.OracleEstimator(net, scales, clock=lambda: clock_vals.pop(0))
        for k in range(3):
            img = synth.stream_frame(0, k)
            t2, t3 = 1000 + k / 30, 1000 + k / 30 + 0.004
            clock_vals[:] = [t2, t3]
            r2, r3 = ref(img)
            j2, j3 = eng.estimate(img, [0], [t2], [t3])
            print(f"e2e frame {k}: joints with different argmax {(np.abs(j2[0]-r2).max(1) > 1e-9).sum()} max|dj2| {np.abs(j2[0]-r2).max():.3f} px; max|dj3| {np.abs(j3[0]-r3).max():.4f} mm")
        frames = np.stack([synth.frame_c2(i) for i in range(8)])
        eng.reset()
        j2, j3 = eng.estimate(frames, np.arange(8), np.full(8, 5.0), np.full(8, 5.004))
        bad = 0
        for i in range(8):
            clock_vals[:] = [5.0, 5.004]
            ref = prepost.OracleEstimator(net, scales, clock=lambda: clock_vals.pop(0))
            r2, r3 = ref(frames[i])
            bad += int((np.abs(j2[i] - r2).max(1) > 1e-9).sum())
            print(f"e2e batch frame {i}: argmax diffs {(np.abs(j2[i]-r2).max(1) > 1e-9).sum()} max|dj3| {np.abs(j3[i]-r3).max():.4f} mm")

    if "time" in what:
        for n in (2, 16, 128):
            tot, per = eng.time_forward(n, reps=5, per_layer=True)
            fl = 23830290432 * n
            print(f"forward n={n}: {tot:.3f} ms  -> {n/tot*1e3:.0f} forwards/s, {fl/tot*1e-9:.1f} TFLOP/s algorithmic")
            if n == 128:
                for k, v in per.items():
                    print(f"    {k:24s} {v*1e3:9.1f} us")
        frames = np.stack([synth.frame_c2(i) for i in range(64)])
        eng.reset()
        eng.estimate(frames, np.arange(64), np.full(64, 4.0), np.full(64, 4.004))
        pre_ms, post_ms = eng.time_prepost(64, reps=10)
        print(f"pre-processing (pyramid) {pre_ms*1e3:.1f} us, post-process {post_ms*1e3:.1f} us per 64 two-scale frames")
        for it in range(3):
            t0 = time.time()
            eng.estimate(frames, np.arange(64), np.full(64, 5.0 + it), np.full(64, 5.004 + it))
            dt = time.time() - t0
            print(f"estimate(64 frames, host in/out): {dt*1e3:.2f} ms -> {64/dt:.0f} frames/s")
    print("launches:", eng.launch_count())
    eng.close()


if __name__ == "__main__":
    main()
