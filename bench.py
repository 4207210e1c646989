#!/usr/bin/env python3
"""Benchmark of the VNect per-frame hot path (BASELINE.json: frames/s @368x368, 2 scales, on 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path over one batch of 64 synthetic 368x368 BGR frames per GPU (config C2 of
SURVEY.md section 8d: scales [1.0, 0.7] => 128 CNN forwards), OneEuroFilter state live across steps (each frame slot is
a video stream).  For N > 1 launch with torchrun (one rank per GPU); streams are sharded, nothing is exchanged on the
compute path, and the per-step results are all-gathered with NCCL inside the timed region (weak scaling).

value  = frames/s with the frames already resident in HBM (device-timed with CUDA events on the launch stream).
e2e    = frames/s through the public host API (pinned host frames in, host joints out; H2D/D2H inside the timed region).
roofline = the implicit-GEMM convolution kernel family (the dominant kernel): algorithmic FLOPs of the CNN
            / summed device time of its launches, against the measured bf16 tensor peak.
cpu_baseline = the CPU restatement of the reference (oracle/: torch-CPU fp32 CNN + cv2/numpy pre/post) on a bounded
           sample, timed on this box's host cores.  `--impl reference` times only that.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BOX = 368
SCALES = [1.0, 0.7]
FRAMES_PER_GPU = 64
FLOPS_PER_FORWARD = 23_830_290_432  # BASELINE.md section 3 (reference graph, every conv at full resolution)
# res2c / res3d feed only stride-2 1x1 convs, so their 3x3 and last 1x1 are evaluated at even pixels only (identical
# results): 3/4 of those four convs' MACs are not executed.  The roofline uses EXECUTED flops.
SKIPPED_FLOPS = 2 * 3 * ((92 * 92 // 4) * (576 * 64 + 64 * 256) + (46 * 46 // 4) * (1152 * 128 + 128 * 512))
EXECUTED_FLOPS_PER_FORWARD = FLOPS_PER_FORWARD - SKIPPED_FLOPS
METRIC = "frames/sec @368x368 2-scale"
WORKLOAD = "C2: 64 synthetic 368x368 BGR frames per GPU per step, scales [1.0, 0.7] (128 CNN forwards), W0 seeded random-init weights, filters on"


def frame_c2(i, size=BOX):
    """C2 frame i: uniform-noise BGR image, seed 1000 + i."""
    return np.random.default_rng(1000 + i).integers(0, 256, (size, size, 3), dtype=np.uint8)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_arm(n_frames, threads=None):
    """The reference's CPU path on the oracle restatement (TensorFlow 1.x is not installable: SURVEY.md section 8c).
    Returns frames/s over n_frames two-scale frames (after one warm-up frame)."""
    import torch
    from oracle import prepost, synth
    from oracle.forward import OracleNet
    from oracle.weights import make_weights
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    net = OracleNet(make_weights("W0"))
    clock = {"t": 1000.0}

    def tick():
        clock["t"] += 0.004
        return clock["t"]
    est = prepost.OracleEstimator(net, SCALES, clock=tick)
    est(synth.frame_c2(0))
    t0 = time.perf_counter()
    for i in range(n_frames):
        est(synth.frame_c2(i))
    dt = time.perf_counter() - t0
    return n_frames / dt, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 8
    vals = []
    for _ in range(args.warmup):
        cpu_reference_arm(2)
    threads = os.cpu_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, threads = cpu_reference_arm(n)
        vals.append(v)
    wall = time.perf_counter() - t0
    v = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference (TensorFlow 1.x not installable); each step = %d frames" % n},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d two-scale 368x368 frames per step, oracle estimator (torch-CPU fp32 CNN + cv2/numpy pre/post)" % n},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from vnect_b200 import VNectEngine
    from vnect_b200.weights import seeded_init
    from vnect_b200 import parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun for N>1", file=sys.stderr)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "INFO"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=dev)
    nf = args.frames
    n_streams = nf * world

    eng = VNectEngine(seeded_init("W0"), SCALES, BOX, max_frames=nf, max_streams=nf, device=local)
    # a dedicated (capturable) torch stream is made current and handed to the engine: the engine's kernels, the
    # torch.cuda.Event timers and the NCCL gather all live on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_cuda_stream(stream.cuda_stream)

    # synthetic frames: stream s = global frame slot; pinned host copy for the e2e leg, device copy for `value`
    my_streams = parallel.owned_streams(n_streams, rank, world)
    host_frames = torch.empty((nf, BOX, BOX, 3), dtype=torch.uint8).pin_memory()
    hf = host_frames.numpy()
    for i, s in enumerate(my_streams):
        hf[i] = frame_c2(s)
    dev_frames = host_frames.to(dev)
    d_j2 = torch.empty((nf, 21, 2), dtype=torch.float64, device=dev)
    d_j3 = torch.empty((nf, 21, 3), dtype=torch.float32, device=dev)
    ids = np.arange(nf, dtype=np.int32)
    tclock = {"t": 1000.0}

    def stamps():
        tclock["t"] += 1.0 / 30
        return np.full(nf, tclock["t"]), np.full(nf, tclock["t"] + 0.004)

    def device_step():
        t2, t3 = stamps()
        eng.estimate_device(dev_frames.data_ptr(), nf, BOX, BOX, d_j2.data_ptr(), d_j3.data_ptr(), ids, t2, t3)
        if world > 1:
            packed = torch.cat([d_j2, d_j3.to(torch.float64)], dim=2)
            out = torch.empty((world * nf, 21, 5), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(out, packed)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- value: device-resident, CUDA-event timed
    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        device_step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = n_streams * args.steps / (ms_max * 1e-3)

    # ---------------------------------------------------------------- e2e: host frames in, host joints out
    # Public host API, pinned host frames -> pinned host joints, two submission lanes so the H2D copy of batch k+1
    # overlaps the kernels of batch k.  Every step copies its 26 MB of frames to the device and its joints back.
    outs = []
    for _ in range(2):
        a = torch.empty((nf, 21, 2), dtype=torch.float64).pin_memory()
        b = torch.empty((nf, 21, 3), dtype=torch.float32).pin_memory()
        outs.append((a.numpy(), b.numpy(), a, b))
    j2h, j3h = outs[0][0], outs[0][1]

    # cross-rank gather of a finished step's results: device-side, asynchronous, waited one step later so that it
    # never stalls the submission pipeline (the only collective on the path, 21 x 5 numbers per frame)
    g_in = [torch.empty((nf, 21, 5), dtype=torch.float64, device=dev) for _ in range(2)]
    g_out = [torch.empty((world * nf, 21, 5), dtype=torch.float64, device=dev) for _ in range(2)]
    g_work = [None, None]

    def gather(lane):
        if world == 1:
            return
        if g_work[lane] is not None:
            g_work[lane].wait()
        g_in[lane][:, :, :2].copy_(outs[lane][2], non_blocking=True)
        g_in[lane][:, :, 2:].copy_(outs[lane][3], non_blocking=True)
        g_work[lane] = dist.all_gather_into_tensor(g_out[lane], g_in[lane], async_op=True)

    def e2e_steps(k):
        for i in range(k):
            lane = i & 1
            if i >= 2:
                eng.wait(lane)
                gather(lane)
            t2, t3 = stamps()
            eng.submit(lane, hf, ids, t2, t3, out=(outs[lane][0], outs[lane][1]))
        for i in range(max(k - 2, 0), k):
            eng.wait(i & 1)
            gather(i & 1)
        for w in g_work:
            if w is not None:
                w.wait()

    e2e_steps(3)
    barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_streams * args.steps / float(t.item())
    clocks = sampler.stop() if rank == 0 else None  # sampled across both timed regions (value and e2e)

    # ---------------------------------------------------------------- roofline of the conv-GEMM kernel family
    barrier()
    fwd_ms, per = eng.time_forward(nf * len(SCALES), reps=3, per_layer=True)
    conv_ms = sum(per.values())  # every launch of one forward batch is an implicit-GEMM kernel (pool1 is fused)
    peaks, peak_src = measured_peaks()
    achieved = EXECUTED_FLOPS_PER_FORWARD * nf * len(SCALES) / (conv_ms * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    traffic, traffic_note = None, None
    try:  # DRAM traffic of the same launches from the committed ncu capture (profiles/r01_step_traffic.txt)
        with open(os.path.join(ROOT, "profiles", "conv_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("frames_per_step") == nf:
            traffic = tj["conv_family_dram_bytes_per_step"]
            traffic_note = "dram__bytes_read.sum + dram__bytes_write.sum summed over the family's launches of one step, " + tj["source"]
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "conv_gemm_kernel<*> (implicit-GEMM conv family, %d launches per forward batch)" % (len(per) - 1),
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": "%s MEASURED_PEAKS.json bf16_tflops_sustained (burst %.1f)" % (peak_src, peaks["bf16_tflops"]),
                "traffic": traffic, "traffic_note": traffic_note,
                "hbm_view": None if traffic is None else {
                    "achieved_gbs": traffic / (conv_ms * 1e-3) / 1e9, "peak_gbs": peaks["hbm_gbs"],
                    "frac": traffic / (conv_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "note": "the family mixes tensor-bound 3x3 convs with HBM-bound 1x1 expand convs; per-launch numbers in profiles/r01_step_traffic.txt"},
                "conv_ms_per_batch": conv_ms, "forward_ms_per_batch": fwd_ms,
                "flops_per_forward_executed": EXECUTED_FLOPS_PER_FORWARD, "flops_per_forward_reference": FLOPS_PER_FORWARD}

    # ---------------------------------------------------------------- batch-1 latency (C3-style: one stream, filters on)
    # measured on a latency plan (max_frames = 1, what the drop-in VNectEstimator builds): host frame in, host joints out
    lat = []
    if rank == 0:
        lat_eng = VNectEngine(seeded_init("W0"), SCALES, BOX, max_frames=1, max_streams=1, device=local)
        one = hf[:1]
        o2 = np.empty((1, 21, 2), np.float64)
        o3 = np.empty((1, 21, 3), np.float32)
        for k in range(60):
            tclock["t"] += 1.0 / 30
            t0 = time.perf_counter()
            lat_eng.estimate(one, [0], [tclock["t"]], [tclock["t"] + 0.004], out=(o2, o3))
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = sorted(lat[10:])
        lat_eng.close()

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world >= 1:
            n_cpu = 12
            v, threads = cpu_reference_arm(n_cpu)
            cpu = {"value": v, "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": "%d two-scale 368x368 frames of the same workload through the oracle estimator (torch-CPU fp32 CNN + cv2/numpy pre/post; TF1 not installable)" % n_cpu}
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": nf, "box_size": BOX, "scales": SCALES,
                       "precision": "fp16 operands, fp32 accumulate (TMEM)",
                       "l2": "per-step working set (~8 GB of activations) exceeds the 126 MB L2; no flush needed",
                       "sharding": "streams round-robin over ranks, results all-gathered (NCCL) each step" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(hf.nbytes) * world,
                    "d2h_bytes_per_step": int(j2h.nbytes + j3h.nbytes) * world},
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "latency_ms_p50_batch1": statistics.median(lat) if lat else None,
            "latency_ms_p95_batch1": lat[int(0.95 * (len(lat) - 1))] if lat else None,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
