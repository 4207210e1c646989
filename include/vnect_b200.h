/* vnect_b200 -- C ABI of the B200-native VNect per-frame hot path.
 *
 * The reference (XinArkh/VNect) has no FFI of its own: its seam is the Python class `VNectEstimator`
 * (src/estimator.py:16-142) on top of a TensorFlow session looked up by tensor name (src/estimator.py:62-66).
 * Each entry point below names the reference interface it replaces.  Plain pointers and sizes only; every buffer
 * passed in or out is owned by the caller, the handle owns device weights, workspaces and per-stream filter state.
 * All functions return 0 on success or a negative VNECT_E_* code; vnect_last_error() gives the message.
 * One handle = one device + one CUDA stream; a handle is not thread-safe (neither is the reference object).  A process
 * may hold handles on several devices: every entry point switches to its handle's device and restores the caller's.
 */
#ifndef VNECT_B200_H
#define VNECT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VNECT_OK 0
#define VNECT_E_INVALID (-1)      /* bad argument / wrong call order */
#define VNECT_E_CUDA (-2)         /* CUDA runtime or driver error, incl. no usable sm_100 device */
#define VNECT_E_WEIGHT (-3)       /* unknown variable name, wrong shape, or missing variable at finalize */
#define VNECT_E_ZERO_DT (-4)      /* repeated timestamp for a stream: the reference raises ZeroDivisionError
                                     (src/OneEuroFilter.py:66) */
#define VNECT_E_UNSUPPORTED (-5)
#define VNECT_E_NUMERIC (-6)      /* NaN / Inf in the CNN output maps: fp16 activations overflowed (un-normalised weights?) */

#define VNECT_JOINTS 21
#define VNECT_MAX_SCALES 4

typedef struct vnect_handle vnect_t;

typedef struct vnect_config {
  int32_t device;       /* CUDA device ordinal */
  int32_t box_size;     /* CNN input side, 368 in the reference (src/estimator.py:19); multiple of 16 */
  int32_t n_scales;     /* 1..VNECT_MAX_SCALES */
  double scales[VNECT_MAX_SCALES]; /* pyramid scales, each in (0, 1]; reference default {1, 0.85, 0.7}
                                      (src/estimator.py:32) */
  int32_t max_frames;   /* largest n_frames of one vnect_estimate call (forward batch = max_frames * n_scales) */
  int32_t max_streams;  /* number of independent temporal-filter slots (video streams) */
  int32_t max_input_h;  /* largest raw frame accepted by vnect_estimate; 0 = box_size */
  int32_t max_input_w;
  int32_t filters;      /* 1 = OneEuroFilter smoothing on (reference behaviour), 0 = raw joints */
} vnect_config;

/* replaces VNectEstimator.__init__ (src/estimator.py:27-68): allocates the device context; weights come next */
int vnect_create(vnect_t** out, const vnect_config* cfg);

/* replaces VNect.load_weights / assign_weights_from_dict (src/vnect_model.py:219-236): one call per TF variable, names
 * and layouts of src/caffe2pkl.py:57-76 ("<scope>/weights" HWIO, "<scope>/biases", "<scope>/kernel",
 * "bn5c_branch2a/{gamma,beta,moving_mean,moving_variance}").  Host pointer, copied.  res2c_branch2a/* is accepted
 * and ignored (dead in the reference graph, src/vnect_model.py:54-56). */
int vnect_set_weight(vnect_t* h, const char* tf_name, const float* data, const int64_t* shape, int32_t rank);

/* weight-file helper: CRC32C (Castagnoli) of a host buffer -- the per-tensor / per-block checksum of the TensorFlow
 * checkpoint the reference restores (src/estimator.py:55-60); used by vnect_b200/tf_checkpoint.py */
uint32_t vnect_crc32c(const void* data, uint64_t n);

/* folds batch norm, converts to fp16, packs for the tensor cores, builds the launch plan (what saver.restore +
 * graph import do in src/estimator.py:54-60) */
int vnect_finalize(vnect_t* h);

/* replaces sess.run([split_2:0..3], {Placeholder:0: batch}) (src/estimator.py:100-104):
 * nhwc float32 [n, S, S, 3] -> four float32 [n, S/8, S/8, 21] maps (heat-map, x, y, z).  Host pointers. */
int vnect_forward(vnect_t* h, const float* nhwc, int32_t n, float* hm, float* xm, float* ym, float* zm);

/* replaces VNectEstimator.__call__ (src/estimator.py:97-142) for n_frames independent frames of identical size:
 * bgr uint8 [n_frames][H][W][3] with row pitch `pitch` and frame stride `frame_stride` (bytes);
 * stream_ids[n_frames] select the filter slot of each frame (distinct within one call);
 * t2d / t3d [n_frames] are the two clock readings the reference takes per frame (src/estimator.py:84);
 * joints2d float64 [n_frames][21][2] = (row, col) in input-image pixels, joints3d float32 [n_frames][21][3] in mm,
 * root-relative.  All pointers are HOST pointers; copies are part of the call. */
int vnect_estimate(vnect_t* h, const uint8_t* bgr, int32_t n_frames, int32_t H, int32_t W, int64_t pitch,
                   int64_t frame_stride, const int32_t* stream_ids, const double* t2d, const double* t3d,
                   double* joints2d, float* joints3d);

/* pipelined form of vnect_estimate: enqueue one batch on submission lane 0 or 1 and return immediately; the H2D copy
 * of a lane overlaps the kernels of the other one.  bgr / joints2d / joints3d are HOST pointers that must stay valid
 * (pinned memory recommended) until vnect_wait(h, lane) returns.  Batches execute in submission order, so frames of one
 * stream may alternate between lanes.  vnect_estimate == vnect_submit(lane 0) + vnect_wait(lane 0). */
int vnect_submit(vnect_t* h, int32_t lane, const uint8_t* bgr, int32_t n_frames, int32_t H, int32_t W, int64_t pitch,
                 int64_t frame_stride, const int32_t* stream_ids, const double* t2d, const double* t3d, double* joints2d,
                 float* joints3d);
int vnect_wait(vnect_t* h, int32_t lane);

/* Tracked video streams: replaces the loop body of run_estimator.py:98-119 for n streams at once.  Every stream keeps a
 * crop box on the device; vnect_track crops each FULL frame by its stream's box (run_estimator.py:100), runs
 * VNectEstimator.__call__ on the crop, shifts joints_2d to full-frame coordinates (:104-105) and updates the box from the
 * joints (:110-119) -- no host round trip between frames.  vnect_track_set_box seeds a stream's box (what HOGBox does in
 * run_estimator.py:66-83).  frames: HOST uint8 [n][FH][FW][3]; joints2d are full-frame (row, col); boxes_used (optional)
 * int32 [n][4] = the (x, y, w, h) each frame was cropped with. */
int vnect_track_set_box(vnect_t* h, int32_t stream_id, int32_t x, int32_t y, int32_t w, int32_t hh);
int vnect_track_get_box(vnect_t* h, int32_t stream_id, int32_t* xywh);
int vnect_track(vnect_t* h, const uint8_t* frames, int32_t n_frames, int32_t FH, int32_t FW, int64_t pitch,
                int64_t frame_stride, const int32_t* stream_ids, const double* t2d, const double* t3d, double* joints2d,
                float* joints3d, int32_t* boxes_used);

/* same computation with the frames and the results resident in device memory (dev_bgr, dev_joints2d, dev_joints3d are
 * DEVICE pointers; stream_ids / t2d / t3d stay host arrays).  Asynchronous on the handle's stream. */
int vnect_estimate_device(vnect_t* h, const uint8_t* dev_bgr, int32_t n_frames, int32_t H, int32_t W, int64_t pitch,
                          int64_t frame_stride, const int32_t* stream_ids, const double* t2d, const double* t3d,
                          double* dev_joints2d, float* dev_joints3d);

/* replaces VNectEstimator.gen_input_batch (src/estimator.py:70-81): out float32 [n_frames*n_scales][S][S][3]
 * (fp16-rounded values of the network input), scaler_offsets float64 [3] = {scaler, offset_x, offset_y} */
int vnect_preprocess(vnect_t* h, const uint8_t* bgr, int32_t n_frames, int32_t H, int32_t W, int64_t pitch,
                     int64_t frame_stride, float* out_nhwc, double* scaler_offsets);

/* replaces src/estimator.py:105-142 fed with given maps (host float32 NHWC [n_frames*n_scales][S/8][S/8][21] each);
 * scaler / offset_x / offset_y as returned by gen_input_batch.  raw_argmax (optional, may be NULL): int32
 * [n_frames][21][2] unfiltered (row, col) in box pixels. */
int vnect_postprocess(vnect_t* h, const float* hm, const float* xm, const float* ym, const float* zm, int32_t n_frames,
                      const int32_t* stream_ids, const double* t2d, const double* t3d, double scaler, int32_t offset_x,
                      int32_t offset_y, double* joints2d, float* joints3d, int32_t* raw_argmax);

/* replaces VNectEstimator.joint_filter (src/estimator.py:83-95) on explicit values: one step of the 21*dim scalar
 * filters of a stream at clock reading t.  values: host float64 [21*dim], filtered in place.  values_are_f32 = 1 when
 * the caller's array is float32 (the reference's joints_3d, src/utils.py:186): the raw difference is then taken in
 * float32 and the results are rounded to float32; 0 for a float64 array (joints_2d).
 * Errors like the reference: a repeated timestamp is VNECT_E_ZERO_DT (ZeroDivisionError, src/OneEuroFilter.py:66), an
 * earlier one VNECT_E_INVALID (ValueError "alpha ... should be in (0.0, 1.0]", src/OneEuroFilter.py:21-22); both are
 * detected before the stream's state is touched.  The same checks guard vnect_estimate / vnect_submit / vnect_track. */
int vnect_filter(vnect_t* h, int32_t stream_id, int32_t dim, int32_t values_are_f32, double t, double* values);

/* replaces Joints2Angles.__call__ / joints2angles (src/joints2angles.py:44-110): the eight arm angles (radians:
 * s0_l, s1_l, e0_l, e1_l, s0_r, s1_r, e0_r, e1_r) of n frames from their 3D joints (host float32 [n][21][3], as
 * returned by vnect_estimate), smoothed per stream by the eight OneEuroFilters of joints2angles.py:35-42 when t
 * (host float64 [n] clock readings) is not NULL.  angles: host float64 [n][8]. */
int vnect_joints2angles(vnect_t* h, const float* joints3d, int32_t n, const int32_t* stream_ids, const double* t,
                        double* angles);

/* forget the temporal-filter state of one stream (a new VNectEstimator() in the reference); -1 = all streams */
int vnect_reset_stream(vnect_t* h, int32_t stream_id);

/* Temporal state of one stream as VNECT_STREAM_STATE_DOUBLES plain doubles: the 42 + 63 OneEuroFilter objects of
 * src/estimator.py:46-53 ({prev, s_x, s_dx, lasttime, freq, has_prev, has_time} each), the two last clock readings and
 * the tracked crop box.  Lets a caller move a stream to another handle / GPU or keep it across a rebuild of the device
 * context (the reference keeps one filter set for the lifetime of its estimator object). */
#define VNECT_STREAM_STATE_DOUBLES (7 * 21 * 5 + 6)
int vnect_export_stream_state(vnect_t* h, int32_t stream_id, double* state);
int vnect_import_stream_state(vnect_t* h, int32_t stream_id, const double* state);

/* Multi-GPU gather layout: when dev_packed (DEVICE pointer, float64 [max_frames][21][5]) is set, every later
 * estimate / submit / track call also writes (row, col, x, y, z) per joint there, straight from the post-process
 * kernel -- the buffer a caller hands to ncclAllGather.  NULL turns it off. */
int vnect_set_packed_results(vnect_t* h, void* dev_packed);

/* run on a caller-provided cudaStream_t (e.g. torch's current stream) instead of the handle's own */
int vnect_set_stream(vnect_t* h, void* cuda_stream);
int vnect_synchronize(vnect_t* h);

/* introspection for tests and benchmarks */
int vnect_get_tap(vnect_t* h, const char* name, int32_t n, float* out_nhwc, int64_t capacity_elems,
                  int32_t* dims4 /* n, H, W, C */);
/* unfiltered (row, col) argmax in box pixels of the frames of the most recent estimate / submit / track call
 * (what utils.extract_2d_joints returns, src/utils.py:153-175, before joint_filter): int32 [n_frames][21][2].
 * Waits for that call to complete.  Parity tests use it to prove that an argmax difference is a near-tie. */
int vnect_get_raw_argmax(vnect_t* h, int32_t n_frames, int32_t* raw_argmax);
/* fp16 safety net.  Every estimate / submit / track call counts NaN / Inf values met in the CNN's output maps and the
 * call that synchronises with that batch (vnect_wait, vnect_estimate, vnect_track, the next use of the lane) fails
 * with VNECT_E_NUMERIC instead of returning joints computed from garbage.  vnect_check_finite then scans the
 * activations of the last forward: counts[i] = values of launch i's output (first n forwards) that are NaN / Inf or
 * saturated (|x| >= 65504); counts has one entry per launch (vnect_step_name order). */
int vnect_check_finite(vnect_t* h, int32_t n, int64_t* counts);
int64_t vnect_launch_count(vnect_t* h);      /* kernels launched by this handle so far */
double vnect_info(vnect_t* h, const char* key); /* "flops_per_forward", "num_sms", "conv_launches_per_forward", ... */
/* device time of `reps` back-to-back forwards of n images already in device memory (CUDA events on the handle's
 * stream); per_layer_ms may be NULL or float[conv_launches_per_forward + 1] (pool is the last entry) */
int vnect_time_forward(vnect_t* h, int32_t n, int32_t reps, float* total_ms, float* per_layer_ms);
/* device time (ms, CUDA events on the handle's stream) of the pre-processing kernels and of the post-process kernel for
 * n_frames S x S frames: whatever the lane-0 frame buffer and the CNN output maps currently hold; the filter state
 * of streams 0..n_frames-1 advances by one step per repetition */
int vnect_time_prepost(vnect_t* h, int32_t n_frames, int32_t reps, float* pre_ms, float* post_ms);
const char* vnect_step_name(vnect_t* h, int32_t i); /* name of launch i of one forward, NULL past the end */

const char* vnect_last_error(vnect_t* h);
const char* vnect_version(void);
void vnect_destroy(vnect_t* h);

#ifdef __cplusplus
}
#endif
#endif
