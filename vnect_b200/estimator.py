"""Host side of the B200-native VNect hot path.

* ``VNectEstimator`` -- drop-in for the reference class (src/estimator.py:16-142): same constructor call, attributes,
  methods and return types, so ``run_pic.py`` / ``run_estimator.py`` / ``run_estimator_ps.py`` work unchanged once
  ``src/estimator.py`` re-exports it.
* ``VNectEngine`` -- the batched, multi-stream form of the same path (independent frames or video streams per call),
  which is what throughput on a B200 needs; it is what ``bench.py`` drives.

Both are thin ctypes callers of ``libvnect_b200.so`` (include/vnect_b200.h).  There is no CPU path.
"""
import ctypes as C
import time

import numpy as np

from . import _capi, weights as _weights

JOINTS = 21


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class VNectEngine:
    """Batched estimator: ``estimate(frames[n,H,W,3] uint8 BGR) -> (joints_2d [n,21,2] f64, joints_3d [n,21,3] f32)``.

    One engine = one GPU + one CUDA stream.  Frames of one call are independent (distinct ``stream_ids``); successive
    calls with the same stream id continue that stream's OneEuroFilter state (src/estimator.py:83-95).
    """

    def __init__(self, weights=None, scales=(1, 0.85, 0.7), box_size=368, max_frames=1, max_streams=None,
                 max_input=None, device=0, filters=True):
        self._lib = _capi.load_library()
        self._h = C.c_void_p()
        self.scales = [float(s) for s in scales]
        self.box_size = int(box_size)
        self.hm_size = self.box_size // 8
        self.max_frames = int(max_frames)
        self.max_streams = int(max_streams if max_streams is not None else max_frames)
        self.max_input = tuple(max_input) if max_input else (self.box_size, self.box_size)
        self.device = int(device)
        cfg = _capi.Config()
        cfg.device = self.device
        cfg.box_size = self.box_size
        cfg.n_scales = len(self.scales)
        if not 1 <= len(self.scales) <= _capi.MAX_SCALES:
            raise ValueError("1..4 scales supported")
        for i, s in enumerate(self.scales):
            cfg.scales[i] = s
        cfg.max_frames = self.max_frames
        cfg.max_streams = self.max_streams
        cfg.max_input_h, cfg.max_input_w = int(self.max_input[0]), int(self.max_input[1])
        cfg.filters = 1 if filters else 0
        rc = self._lib.vnect_create(C.byref(self._h), C.byref(cfg))
        try:
            self._check(rc)
            if weights is not False:  # False: pre/post-processing only, no CNN
                self._load(_weights.resolve(weights))
        except Exception:
            self.close()
            raise

    # ------------------------------------------------------------------------------------------------ plumbing
    def _check(self, rc):
        _capi.raise_for(self._lib, self._h, rc)

    def _load(self, wdict):
        """VNect.load_weights (src/vnect_model.py:219-236): every variable of the graph, by TF name."""
        for name in _weights.variable_shapes():
            if name not in wdict:
                raise KeyError(f"weights lack variable '{name}'")
            arr = np.ascontiguousarray(wdict[name], dtype=np.float32)
            shape = (C.c_int64 * arr.ndim)(*arr.shape)
            self._check(self._lib.vnect_set_weight(self._h, name.encode(), _ptr(arr), shape, arr.ndim))
        self._check(self._lib.vnect_finalize(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.vnect_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------ operators
    def forward(self, batch):
        """sess.run([split_2:0..3], {Placeholder:0: batch}) (src/estimator.py:100-104)."""
        batch = np.ascontiguousarray(batch, dtype=np.float32)
        n, s = batch.shape[0], self.box_size
        if batch.shape != (n, s, s, 3):
            raise ValueError(f"expected [n,{s},{s},3], got {batch.shape}")
        outs = [np.empty((n, self.hm_size, self.hm_size, JOINTS), np.float32) for _ in range(4)]
        self._check(self._lib.vnect_forward(self._h, _ptr(batch), n, *[_ptr(o) for o in outs]))
        return tuple(outs)

    @staticmethod
    def _frames(frames):
        frames = np.asarray(frames)
        if frames.ndim == 3:
            frames = frames[None]
        if frames.dtype != np.uint8 or frames.ndim != 4 or frames.shape[3] != 3:
            raise ValueError("frames must be uint8 [n,H,W,3] (BGR)")
        st = frames.strides
        if not (st[3] == 1 and st[2] == 3 and st[1] >= frames.shape[2] * 3 and st[0] >= 0):
            frames = np.ascontiguousarray(frames)
            st = frames.strides
        return frames, st

    def _meta(self, n, stream_ids, t2d, t3d):
        ids = np.arange(n, dtype=np.int32) if stream_ids is None else np.ascontiguousarray(stream_ids, dtype=np.int32)
        if t2d is None:
            t2d = np.full(n, time.time())
        if t3d is None:
            t3d = np.full(n, time.time())
        t2d = np.ascontiguousarray(np.broadcast_to(np.asarray(t2d, dtype=np.float64), (n,)))
        t3d = np.ascontiguousarray(np.broadcast_to(np.asarray(t3d, dtype=np.float64), (n,)))
        if ids.shape != (n,):
            raise ValueError("stream_ids must have one entry per frame")
        return ids, t2d, t3d

    def estimate(self, frames, stream_ids=None, t2d=None, t3d=None, out=None):
        """VNectEstimator.__call__ for n independent frames (host arrays in, host arrays out)."""
        frames, st = self._frames(frames)
        n, h, w = frames.shape[:3]
        ids, t2d, t3d = self._meta(n, stream_ids, t2d, t3d)
        j2, j3 = out if out is not None else (np.empty((n, JOINTS, 2), np.float64), np.empty((n, JOINTS, 3), np.float32))
        self._check(self._lib.vnect_estimate(self._h, _ptr(frames), n, h, w, st[1], st[0] if n > 1 else st[1] * h,
                                             _ptr(ids), _ptr(t2d), _ptr(t3d), _ptr(j2), _ptr(j3)))
        return j2, j3

    def submit(self, lane, frames, stream_ids=None, t2d=None, t3d=None, out=None):
        """Pipelined estimate(): enqueue a batch on lane 0/1 and return (joints_2d, joints_3d) arrays that are filled
        once wait(lane) returns.  `frames` and the result arrays must stay alive until then (pinned memory makes the
        copies truly asynchronous)."""
        frames, st = self._frames(frames)
        n, h, w = frames.shape[:3]
        ids, t2d, t3d = self._meta(n, stream_ids, t2d, t3d)
        j2, j3 = out if out is not None else (np.empty((n, JOINTS, 2), np.float64), np.empty((n, JOINTS, 3), np.float32))
        self._check(self._lib.vnect_submit(self._h, int(lane), _ptr(frames), n, h, w, st[1],
                                           st[0] if n > 1 else st[1] * h, _ptr(ids), _ptr(t2d), _ptr(t3d), _ptr(j2),
                                           _ptr(j3)))
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[int(lane)] = (frames, j2, j3)  # keep the buffers alive
        return j2, j3

    def wait(self, lane):
        self._check(self._lib.vnect_wait(self._h, int(lane)))
        return self.__dict__.get("_inflight", {}).pop(int(lane), (None, None, None))[1:]

    def set_box(self, stream_id, rect):
        """Seed the tracked crop box (x, y, w, h) of a stream (what HOGBox provides in run_estimator.py:66-83)."""
        x, y, w, h = (int(v) for v in rect)
        self._check(self._lib.vnect_track_set_box(self._h, int(stream_id), x, y, w, h))

    def get_box(self, stream_id):
        out = (C.c_int32 * 4)()
        self._check(self._lib.vnect_track_get_box(self._h, int(stream_id), out))
        return tuple(out)

    def track(self, frames, stream_ids=None, t2d=None, t3d=None):
        """One step of the reference's video loop (run_estimator.py:98-119) for n streams: crop each full frame by
        its stream's box, estimate, shift joints_2d to full-frame coordinates, update the box on the device.
        Returns (joints_2d [n,21,2], joints_3d [n,21,3], boxes_used [n,4])."""
        frames, st = self._frames(frames)
        n, h, w = frames.shape[:3]
        ids, t2d, t3d = self._meta(n, stream_ids, t2d, t3d)
        j2, j3 = np.empty((n, JOINTS, 2), np.float64), np.empty((n, JOINTS, 3), np.float32)
        boxes = np.empty((n, 4), np.int32)
        self._check(self._lib.vnect_track(self._h, _ptr(frames), n, h, w, st[1], st[0] if n > 1 else st[1] * h,
                                          _ptr(ids), _ptr(t2d), _ptr(t3d), _ptr(j2), _ptr(j3), _ptr(boxes)))
        return j2, j3, boxes

    def estimate_device(self, dev_frames_ptr, n, h, w, dev_j2_ptr, dev_j3_ptr, stream_ids=None, t2d=None, t3d=None,
                        pitch=None, frame_stride=None):
        """Same path with frames / results resident in device memory (raw device pointers, e.g. tensor.data_ptr())."""
        ids, t2d, t3d = self._meta(n, stream_ids, t2d, t3d)
        pitch = pitch or w * 3
        frame_stride = frame_stride or pitch * h
        self._check(self._lib.vnect_estimate_device(self._h, C.c_void_p(dev_frames_ptr), n, h, w, pitch, frame_stride,
                                                    _ptr(ids), _ptr(t2d), _ptr(t3d), C.c_void_p(dev_j2_ptr),
                                                    C.c_void_p(dev_j3_ptr)))

    def preprocess(self, frames):
        """gen_input_batch (src/estimator.py:70-81) -> (batch f32 [n*n_scales,S,S,3], scaler, [offset_x, offset_y])."""
        frames, st = self._frames(frames)
        n, h, w = frames.shape[:3]
        s = self.box_size
        out = np.empty((n * len(self.scales), s, s, 3), np.float32)
        meta = np.zeros(3, np.float64)
        self._check(self._lib.vnect_preprocess(self._h, _ptr(frames), n, h, w, st[1], st[0] if n > 1 else st[1] * h,
                                               _ptr(out), _ptr(meta)))
        return out, float(meta[0]), [int(meta[1]), int(meta[2])]

    def postprocess(self, maps, scaler=1.0, offsets=(0, 0), stream_ids=None, t2d=None, t3d=None):
        """src/estimator.py:105-142 on given (hm, xm, ym, zm) maps, each [n*n_scales, S/8, S/8, 21] float32.
        Returns (joints_2d, joints_3d, raw_argmax)."""
        maps = [np.ascontiguousarray(m, dtype=np.float32) for m in maps]
        n = maps[0].shape[0] // len(self.scales)
        ids, t2d, t3d = self._meta(n, stream_ids, t2d, t3d)
        j2, j3 = np.empty((n, JOINTS, 2), np.float64), np.empty((n, JOINTS, 3), np.float32)
        raw = np.empty((n, JOINTS, 2), np.int32)
        self._check(self._lib.vnect_postprocess(self._h, *[_ptr(m) for m in maps], n, _ptr(ids), _ptr(t2d), _ptr(t3d),
                                                float(scaler), int(offsets[0]), int(offsets[1]), _ptr(j2), _ptr(j3),
                                                _ptr(raw)))
        return j2, j3, raw

    def filter(self, values, dim, t, stream_id=0, is_f32=None):
        """One joint_filter step (src/estimator.py:83-95) on explicit values [21, dim].  ``is_f32`` says whether the
        caller's array is float32 like the reference's joints_3d (default: dim == 3); the result is float64-typed
        either way and float32-valued when it is."""
        is_f32 = (dim == 3) if is_f32 is None else bool(is_f32)
        v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1).copy()
        if v.size != JOINTS * dim:
            raise ValueError(f"expected [21, {dim}] values")
        self._check(self._lib.vnect_filter(self._h, int(stream_id), int(dim), 1 if is_f32 else 0, float(t), _ptr(v)))
        return v.reshape(JOINTS, dim)

    def joints2angles(self, joints_3d, stream_ids=None, t=None):
        """src/joints2angles.py:60-110 for n frames: [n,21,3] float32 -> [n,8] float64 radians, one-euro filtered per
        stream when clock readings ``t`` are given."""
        j = np.ascontiguousarray(joints_3d, dtype=np.float32)
        if j.ndim == 2:
            j = j[None]
        n = j.shape[0]
        ids = np.arange(n, dtype=np.int32) if stream_ids is None else np.ascontiguousarray(stream_ids, dtype=np.int32)
        tt = None if t is None else np.ascontiguousarray(np.broadcast_to(np.asarray(t, dtype=np.float64), (n,)))
        out = np.empty((n, 8), np.float64)
        self._check(self._lib.vnect_joints2angles(self._h, _ptr(j), n, _ptr(ids), _ptr(tt) if tt is not None else None,
                                                  _ptr(out)))
        return out

    def raw_argmax(self, n=1):
        """Unfiltered (row, col) argmax in box pixels of the last estimate / submit / track call, int32 [n,21,2]."""
        out = np.empty((n, JOINTS, 2), np.int32)
        self._check(self._lib.vnect_get_raw_argmax(self._h, int(n), _ptr(out)))
        return out

    def export_stream_state(self, stream_id=0):
        """Temporal state of a stream (filters, last clock readings, tracked box) as a float64 vector."""
        out = np.empty(_capi.STREAM_STATE_DOUBLES, np.float64)
        self._check(self._lib.vnect_export_stream_state(self._h, int(stream_id), _ptr(out)))
        return out

    def import_stream_state(self, state, stream_id=0):
        state = np.ascontiguousarray(state, dtype=np.float64)
        if state.shape != (_capi.STREAM_STATE_DOUBLES,):
            raise ValueError("stream state has the wrong size")
        self._check(self._lib.vnect_import_stream_state(self._h, int(stream_id), _ptr(state)))

    def tap(self, name, n=1):
        """Intermediate activation `name` of the last forward as float32 NHWC (fp16 on the device)."""
        dims = (C.c_int32 * 4)()
        self._check(self._lib.vnect_get_tap(self._h, name.encode(), n, None, 0, dims))
        out = np.empty(tuple(dims), np.float32)
        self._check(self._lib.vnect_get_tap(self._h, name.encode(), n, _ptr(out), out.size, dims))
        return out

    def check_finite(self, n=1):
        """{launch name: count of NaN / Inf / saturated fp16 values in its output} for the first n forwards of the last
        batch -- names the layer where un-normalised weights overflow fp16 (65504)."""
        names = self.step_names()
        counts = np.zeros(len(names), np.int64)
        self._check(self._lib.vnect_check_finite(self._h, int(n), _ptr(counts)))
        return dict(zip(names, counts.tolist()))

    def reset(self, stream_id=-1):
        self._check(self._lib.vnect_reset_stream(self._h, int(stream_id)))

    def set_cuda_stream(self, stream_ptr):
        self._check(self._lib.vnect_set_stream(self._h, C.c_void_p(stream_ptr)))

    def set_packed_results(self, dev_ptr):
        """Device pointer to float64 [max_frames,21,5]: later calls also write (row, col, x, y, z) there (0 = off)."""
        self._check(self._lib.vnect_set_packed_results(self._h, C.c_void_p(dev_ptr) if dev_ptr else None))

    def synchronize(self):
        self._check(self._lib.vnect_synchronize(self._h))

    def launch_count(self):
        return int(self._lib.vnect_launch_count(self._h))

    def info(self, key):
        return float(self._lib.vnect_info(self._h, key.encode()))

    def step_names(self):
        out, i = [], 0
        while True:
            s = self._lib.vnect_step_name(self._h, i)
            if s is None:
                return out
            out.append(s.decode())
            i += 1

    def time_forward(self, n, reps=5, per_layer=False):
        """Device time (ms) of one forward of n images already resident; optionally per launch."""
        total = C.c_float()
        names = self.step_names()
        per = np.zeros(len(names), np.float32) if per_layer else None
        self._check(self._lib.vnect_time_forward(self._h, n, reps, C.byref(total), _ptr(per) if per_layer else None))
        return (total.value, dict(zip(names, per.tolist()))) if per_layer else total.value


    def time_prepost(self, n_frames, reps=10):
        """Device time (ms) of the pre-processing kernels and of the post-process kernel for n_frames box-size frames."""
        pre, post = C.c_float(), C.c_float()
        self._check(self._lib.vnect_time_prepost(self._h, int(n_frames), int(reps), C.byref(pre), C.byref(post)))
        return pre.value, post.value


class VNectEstimator:
    """Drop-in for the reference ``VNectEstimator`` (src/estimator.py:16-142).

    ``VNectEstimator()`` then ``joints_2d, joints_3d = estimator(img_bgr)``: ``joints_2d`` float64 [21,2] (row, col) in
    input-image pixels, ``joints_3d`` float32 [21,3] in mm relative to joint 14; fresh writable host arrays each call.
    Optional keyword arguments (all absent in the reference) choose weights, scales, clock and device.
    """

    box_size = 368  # src/estimator.py:19
    hm_factor = 8  # :21
    joints_sum = 21  # :23
    joint_parents = [16, 15, 1, 2, 3, 1, 5, 6, 14, 8, 9, 14, 11, 12, 14, 14, 1, 4, 7, 10, 13]  # :25

    # Largest frame the device context is sized for up front.  The reference's video loop hands in a crop whose size
    # changes every frame (run_estimator.py:100, 110-119), so sizing by the first frame would rebuild mid-video.
    default_max_input = (1080, 1920)

    def __init__(self, weights=None, scales=None, box_size=None, device=0, clock=None, verbose=True, max_input=None):
        self._verbose = verbose
        if verbose:
            print('Initializing VNect Estimator...')
        self.scales = [1, 0.85, 0.7] if scales is None else list(scales)  # src/estimator.py:32
        if box_size is not None:
            self.box_size = int(box_size)
        self._weights = _weights.resolve(weights)
        self._device = device
        self._clock = clock or time.time
        self._engine = None
        self._engine_key = None
        self._max_input = tuple(max_input) if max_input else self.default_max_input
        self._ensure_engine((self.box_size, self.box_size))
        if verbose:
            print('VNect Estimator initialized.')

    def _ensure_engine(self, hw):
        """(Re)build the device context when `scales` was changed by the caller (the reference reads self.scales on
        every call, src/estimator.py:99) or a frame larger than `max_input` arrives.  The OneEuroFilter state belongs
        to the estimator object in the reference (src/estimator.py:46-53) and survives both: it is carried over."""
        need_h = max(hw[0], self.box_size, self._max_input[0])
        need_w = max(hw[1], self.box_size, self._max_input[1])
        key = (tuple(float(s) for s in self.scales), self.box_size)
        if self._engine is not None and key == self._engine_key and need_h <= self._engine.max_input[0] \
                and need_w <= self._engine.max_input[1]:
            return
        state = None
        if self._engine is not None:
            need_h = max(need_h, self._engine.max_input[0])
            need_w = max(need_w, self._engine.max_input[1])
            state = self._engine.export_stream_state(0)
            self._engine.close()
        self._max_input = (need_h, need_w)
        self._engine = VNectEngine(self._weights, self.scales, self.box_size, max_frames=1, max_streams=1,
                                   max_input=(need_h, need_w), device=self._device, filters=True)
        if state is not None:
            self._engine.import_stream_state(state, 0)
        self._engine_key = key

    @staticmethod
    def gen_input_batch(img_input, box_size, scales):
        """src/estimator.py:70-81; computed by the CUDA preprocessing kernels (values are the fp16 network input)."""
        img = np.asarray(img_input)
        eng = VNectEngine(False, scales, box_size, max_frames=1, max_input=(max(img.shape[0], box_size),
                                                                           max(img.shape[1], box_size)))
        try:
            return eng.preprocess(img)
        finally:
            eng.close()

    def joint_filter(self, joints, dim=2):
        """src/estimator.py:83-95: in-place one-euro filtering of [21, dim] joints at the current clock reading.  A
        float32 array (the reference's joints_3d) is filtered with float32 raw differences and float32 results, a
        float64 one (joints_2d) in float64."""
        is_f32 = getattr(joints, "dtype", None) == np.float32
        joints[...] = self._engine.filter(joints, dim, self._clock(), is_f32=is_f32)
        return joints

    def __call__(self, img_input):
        t0 = time.time()
        img = np.asarray(img_input)
        self._ensure_engine(img.shape[:2])
        # The reference reads the clock once per filter group, AFTER its sess.run (src/estimator.py:84, 132-135).  Here
        # pre-processing, CNN and post-processing are one device submission, so both readings are taken just before
        # it: every reading is earlier by about the CNN latency, the differences between frames -- all the filters
        # use -- are unchanged.
        t2d = self._clock()
        t3d = self._clock()
        j2, j3 = self._engine.estimate(img, stream_ids=[0], t2d=[t2d], t3d=[t3d])
        joints_2d, joints_3d = j2[0].copy(), j3[0].copy()
        if self._verbose:
            print('FPS: {:>2.2f}'.format(1 / (time.time() - t0)))
        return joints_2d, joints_3d


class Joints2Angles:
    """Drop-in for the reference class (src/joints2angles.py:23-57): ``angles = Joints2Angles()(joints_3d)`` returns the
    list [s0_l, s1_l, e0_l, e1_l, s0_r, s1_r, e0_r, e1_r] in radians and prints them, computed (and one-euro filtered) on
    the device.  ``engine`` may be shared with an estimator; otherwise a small CNN-less context is created."""

    def __init__(self, filter=True, engine=None, clock=None, verbose=True):
        if verbose:
            print('Initializing Joints2Angles...')
        self.filter = filter
        self._engine = engine or VNectEngine(False, [1.0], max_frames=1, max_streams=1)
        self._clock = clock or time.time
        self._verbose = verbose
        if verbose:
            print('Joints2Angles initialized.')

    def __call__(self, joints_3d):
        t = [self._clock()] if self.filter else None
        angles = [float(a) for a in self._engine.joints2angles(np.asarray(joints_3d)[None], [0], t)[0]]
        if self._verbose:
            print('%5.2f | %5.2f | %5.2f | %5.2f | %5.2f | %5.2f | %5.2f | %5.2f' % tuple(angles))
        return angles

    @staticmethod
    def joints2angles(joints_3d, engine=None):
        eng = engine or VNectEngine(False, [1.0], max_frames=1, max_streams=1)
        try:
            return tuple(float(a) for a in eng.joints2angles(np.asarray(joints_3d)[None])[0])
        finally:
            if engine is None:
                eng.close()
