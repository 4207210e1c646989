// C ABI of the B200-native VNect hot path (see include/vnect_b200.h): handle, weight ingest, plan, entry points.
#include "vnect_b200.h"

#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <tuple>
#include <map>
#include <algorithm>
#include <set>
#include <string>
#include <vector>

#include "conv_plan.cuh"
#include "prepost_kernels.cuh"

using namespace vnect;

namespace {

struct Act {
  __half* p = nullptr;
  int H = 0, W = 0, C = 0;
  int row_px = 0;       // row pitch in pixels (== W except for conv1's virtual layout)
  int64_t img_px = 0;   // image stride in pixels (== H*W except for conv1)
};

struct Step {
  int kind = 0;  // 0 = implicit-GEMM conv (conv_gemm_kernel), 3 = fused conv1 + pool1 (stem_roll_kernel)
  std::string name;
  ConvLaunch launch;
  StemPoolLaunch stem_pool;
};

struct HostVar {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

}  // namespace

struct vnect_handle {
  vnect_config cfg{};
  int S = 368, hs = 46, n_scales = 1, cap_fw = 1, num_sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  bool finalized = false;
  std::string err;
  long long launches = 0;
  std::map<std::string, HostVar> vars;
  std::vector<void*> allocs;
  std::map<std::string, Act> acts;
  std::vector<Step> steps;
  int conv_steps = 0;
  // stem input
  __half* x1 = nullptr;
  int stem_rpp = 0, stem_pitch = 0;
  // output maps, planar fp32 [cap_fw][84][hs][hs]
  float* maps = nullptr;
  // pre/post buffers
  size_t d_frames_bytes = 0;
  uint8_t* d_sq = nullptr;
  float* d_f32_in = nullptr;  // vnect_forward staging [cap_fw][S][S][3]
  ScaleTable* d_tables = nullptr;
  unsigned long long* d_post_trace = nullptr;  // VNECT_B200_POST_TRACE=1: phase timestamps of the last post-process launch
  std::vector<ScaleTable> h_tables;  // host copy: the post-process block's shared-memory plan is derived from it
  PyramidTable* d_pyr_tables = nullptr;
  uint32_t* d_norm_lut = nullptr;
  int pyr_src_rows = 1;         // most source rows a pyramid block stages
  bool x1_surround_ok = false;  // the constant surround of the shrunken scales is in place in every slot of x1
  FilterState *d_st2d = nullptr, *d_st3d = nullptr;
  double* d_j2_box = nullptr;
  float* d_j3_raw = nullptr;
  int* d_raw_argmax = nullptr;
  unsigned int* d_counter = nullptr;
  double* d_prep3 = nullptr;   // [max_frames][21][3][3]: see PostParams::prep3
  double* d_filter_scratch = nullptr;
  // Two submission lanes: each owns its input staging, per-call meta (pinned host + device) and result buffers, so
  // the H2D copy of batch k+1 (copy stream) overlaps the kernels of batch k (compute stream).
  struct Lane {
    uint8_t* d_frames = nullptr;
    int* d_stream_ids = nullptr;
    double *d_t2d = nullptr, *d_t3d = nullptr;
    double* d_out2d = nullptr;
    float* d_out3d = nullptr;
    int* h_stream_ids = nullptr;
    double *h_t2d = nullptr, *h_t3d = nullptr;
    void* h_meta = nullptr;               // pinned block behind h_stream_ids / h_t2d / h_t3d (one H2D copy per call)
    size_t meta_bytes = 0;
    unsigned int* h_nonfinite = nullptr;  // pinned + device-visible: set by the post-process when a batch of this lane met NaN / Inf maps
    cudaEvent_t copy_done = nullptr, done = nullptr;
    bool pending = false;
  } lanes[2];
  Lane* cur = &lanes[0];
  cudaStream_t copy_stream = nullptr;
  std::vector<unsigned int> seen_stamp;  // duplicate stream ids within a call (stage_frame_meta)
  unsigned int seen_epoch = 0;
  unsigned device_calls = 0;
  // tracked streams (run_estimator.py:100-119): per-stream crop box on the device, per-call frame geometry
  int4* d_boxes = nullptr;         // [max_streams] (x, y, w, h)
  FrameGeom* d_geoms = nullptr;    // [max_frames]
  int* d_boxes_used = nullptr;     // [max_frames][4]
  // CUDA graphs of the kernel sequence of one estimate call, keyed by everything baked into the kernel parameters.
  // First use of a key runs directly (warm-up), the second captures, later ones replay: ~50 launches -> 1.
  typedef std::tuple<int, int, int, int, long long, long long, const void*, const void*, const void*> GraphKey;
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    long long launches = 0;
    bool unusable = false;
    unsigned long long last_use = 0;
  };
  static constexpr size_t kMaxGraphs = 16;  // a video loop changes its crop size every frame: bound the cache (LRU)
  unsigned long long graph_clock = 0;
  std::map<GraphKey, GraphEntry> graphs;
  bool use_graphs = true;
  std::vector<double> last_t2d, last_t3d;  // host mirror of the filters' last timestamps (NaN = none yet)
  PyramidParams pyr{};
  double* d_packed = nullptr;  // caller-owned device buffer [max_frames][21][5] (vnect_set_packed_results) or null
  FilterState* d_st_ang = nullptr;      // [max_streams][8] one-euro filters of Joints2Angles (joints2angles.py:35-42)
  std::vector<double> last_t_ang;
  float* d_ang_in = nullptr;            // [max_frames][21][3]
  double* d_ang_t = nullptr;            // [max_frames]
  int* d_ang_ids = nullptr;
  double* d_ang_out = nullptr;          // [max_frames][8]
  unsigned int* d_nonfinite = nullptr;  // (frame, joint) blocks whose maps held NaN / Inf since the last check
  unsigned long long* d_scan = nullptr; // vnect_check_finite scratch
};

// Every entry point runs on the handle's device and leaves the caller's current device as it found it: function
// attributes, streams and allocations are per device, and a process may hold one handle per GPU.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
#define ON_DEVICE(h) DeviceGuard device_guard_((h)->cfg.device)

static int fail(vnect_t* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

#define CU(h, call)                                                                                       \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) return fail(h, VNECT_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template <typename T>
static int dev_alloc(vnect_t* h, T** p, size_t count, bool zero = true) {
  void* q = nullptr;
  CU(h, cudaMalloc(&q, count * sizeof(T) + 256));
  if (zero) CU(h, cudaMemset(q, 0, count * sizeof(T) + 256));
  h->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return VNECT_OK;
}

static inline int grid_for(int64_t total, int threads, int sms) {
  int64_t b = (total + threads - 1) / threads;
  int64_t cap = (int64_t)sms * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// ------------------------------------------------------------------------------------------------ weights
static const char* kBnVars[4] = {"gamma", "beta", "moving_mean", "moving_variance"};

struct ConvDef {
  std::string scope;
  int k, cin, cout;
};

static std::vector<ConvDef> conv_defs() {
  std::vector<ConvDef> d;
  d.reserve(64);
  auto add = [&](const std::string& s, int k, int ci, int co) { d.push_back({s, k, ci, co}); };
  auto block = [&](const std::string& pre, int cin, int mid, int cout, bool proj, const std::string& suf) {
    if (proj) add(pre + "_branch1" + suf, 1, cin, cout);
    add(pre + "_branch2a" + suf, 1, cin, mid);
    add(pre + "_branch2b" + suf, 3, mid, mid);
    add(pre + "_branch2c" + suf, 1, mid, cout);
  };
  add("conv1", 7, 3, 64);
  block("res2a", 64, 64, 256, true, "");
  block("res2b", 256, 64, 256, false, "");
  block("res2c", 256, 64, 256, false, "");
  block("res3a", 256, 128, 512, true, "");
  for (const char* b : {"res3b", "res3c", "res3d"}) block(b, 512, 128, 512, false, "");
  block("res4a", 512, 256, 1024, true, "");
  for (const char* b : {"res4b", "res4c", "res4d", "res4e", "res4f"}) block(b, 1024, 256, 1024, false, "");
  block("res5a", 1024, 512, 1024, true, "_new");
  add("res5b_branch2a_new", 1, 1024, 256);
  add("res5b_branch2b_new", 3, 256, 128);
  add("res5b_branch2c_new", 1, 128, 256);
  add("res5c_branch2b", 3, 212, 128);
  return d;
}

static bool expected_shape(const std::string& name, std::vector<int64_t>* shape) {
  const size_t slash = name.find('/');
  if (slash == std::string::npos) return false;
  const std::string scope = name.substr(0, slash), leaf = name.substr(slash + 1);
  if (scope == "bn5c_branch2a") {
    for (const char* v : kBnVars)
      if (leaf == v) {
        *shape = {128};
        return true;
      }
    return false;
  }
  if (leaf == "kernel") {
    if (scope == "res5c_branch1a") { *shape = {4, 4, 63, 256}; return true; }
    if (scope == "res5c_branch2a") { *shape = {4, 4, 128, 256}; return true; }
    if (scope == "res5c_branch2c") { *shape = {1, 1, 128, 84}; return true; }
    return false;
  }
  for (const ConvDef& c : conv_defs())
    if (scope == c.scope) {
      if (leaf == "weights") { *shape = {c.k, c.k, c.cin, c.cout}; return true; }
      if (leaf == "biases") { *shape = {c.cout}; return true; }
    }
  return false;
}

static std::vector<__half> to_half(const std::vector<float>& v) {
  std::vector<__half> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __float2half_rn(v[i]);
  return o;
}

// HWIO [k,k,cin,cout] -> [n_pad][k*k*cin_pad], K-major per output channel, zero padded
static std::vector<float> pack_conv(const HostVar& w, int k, int cin, int cout, int cin_pad, int n_pad) {
  std::vector<float> b((size_t)n_pad * k * k * cin_pad, 0.f);
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx)
      for (int ci = 0; ci < cin; ++ci)
        for (int co = 0; co < cout; ++co)
          b[(size_t)co * k * k * cin_pad + (size_t)(ky * k + kx) * cin_pad + ci] =
              w.data[(((size_t)ky * k + kx) * cin + ci) * cout + co];
  return b;
}

// stem: [64][7 rows][8 px][4 ch] (px 7 and ch 3 are zero)
static std::vector<float> pack_stem(const HostVar& w) {
  std::vector<float> b((size_t)64 * 224, 0.f);
  for (int ky = 0; ky < 7; ++ky)
    for (int kx = 0; kx < 7; ++kx)
      for (int c = 0; c < 3; ++c)
        for (int co = 0; co < 64; ++co) b[(size_t)co * 224 + ky * 32 + kx * 4 + c] = w.data[((ky * 7 + kx) * 3 + c) * 64 + co];
  return b;
}

// Both transposed convs + folded batch norm as 4 phase GEMMs: [4 phases][192 cols][4 taps][256 ch].
// cols 0-127 = res5c_branch2a * bn scale, 128-190 = res5c_branch1a, 191 = 0.  TF kernel layout [kh,kw,out,in].
static std::vector<float> pack_deconv(const HostVar& w2a, const HostVar& w1a, const std::vector<float>& bn_scale) {
  std::vector<float> b((size_t)4 * 192 * 1024, 0.f);
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    for (int a = 0; a < 2; ++a)
      for (int bb = 0; bb < 2; ++bb) {
        const int ky = py == 0 ? (a == 0 ? 1 : 3) : (a == 0 ? 0 : 2);
        const int kx = px == 0 ? (bb == 0 ? 1 : 3) : (bb == 0 ? 0 : 2);
        for (int col = 0; col < 191; ++col)
          for (int ci = 0; ci < 256; ++ci) {
            float v;
            if (col < 128)
              v = w2a.data[(((size_t)ky * 4 + kx) * 128 + col) * 256 + ci] * bn_scale[col];
            else
              v = w1a.data[(((size_t)ky * 4 + kx) * 63 + (col - 128)) * 256 + ci];
            b[((size_t)ph * 192 + col) * 1024 + (size_t)(a * 2 + bb) * 256 + ci] = v;
          }
      }
  }
  return b;
}

template <typename T>
static int upload(vnect_t* h, const std::vector<T>& v, T** out) {
  int rc = dev_alloc(h, out, v.size(), false);
  if (rc) return rc;
  CU(h, cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return VNECT_OK;
}

// ------------------------------------------------------------------------------------------------ plan
static int new_act(vnect_t* h, const std::string& name, int H, int W, int C) {
  Act a;
  a.H = H; a.W = W; a.C = C; a.row_px = W; a.img_px = (int64_t)H * W;
  int rc = dev_alloc(h, &a.p, (size_t)h->cap_fw * H * W * C, true);
  if (rc) return rc;
  h->acts[name] = a;
  return VNECT_OK;
}

struct ConvOpts {
  bool relu = true;
  std::string residual;  // act name or empty
  int in_stride = 1;     // 2: compute the conv at even input pixels only (output grid is half-size)
  int res_stride = 1;    // 2: the residual is a full-size tensor read at even pixels
  // fused block tail (block_tail.cuh): scope of the NEXT block's 1x1 reduce conv, computed from this conv's output
  // while it is still in shared memory; its output activation is created under that scope's name
  std::string chain_scope;
  int chain_cout = 0;
};

// The block tail is fused with the next block's reduce conv on the throughput plans (256-column tiles on CTA pairs);
// VNECT_B200_CHAIN=0 turns it off (A/B measurement)
static bool chain_enabled() {
  static const bool on = [] { const char* e = getenv("VNECT_B200_CHAIN"); return !(e && atoi(e) == 0); }();
  return on;
}

// uploads the chained conv's weights ([n2][cout] K-major fp16) and bias, creates its output activation
static int attach_chain(vnect_t* h, ConvSpec& s, const ConvOpts& o, int cout, int OH, int OW) {
  const HostVar& w = h->vars.at(o.chain_scope + "/weights");
  const HostVar& b = h->vars.at(o.chain_scope + "/biases");
  __half* dw = nullptr;
  float* db = nullptr;
  int rc = upload(h, to_half(pack_conv(w, 1, cout, o.chain_cout, cout, o.chain_cout)), &dw);
  if (rc) return rc;
  if ((rc = upload(h, b.data, &db))) return rc;
  if ((rc = new_act(h, o.chain_scope, OH, OW, o.chain_cout))) return rc;
  s.w2 = dw;
  s.bias2 = db;
  s.n2 = o.chain_cout;
  s.out2 = h->acts.at(o.chain_scope).p;
  return VNECT_OK;
}

// Measured at batch 128 (profiles/r02_gpu_check_tail.log): the fused tail beats the two kernels it replaces where the
// pair is HBM-bound -- res2a (224 vs 262 us), res2c (78 vs 101), res3b / res3c (137 vs 172) -- and loses where the
// second accumulator is 256 wide (single-buffered TMEM, res3d / res4* / res5a: 120 vs 107 us) or the main GEMM carries a
// folded projection with a 128-wide second conv (res3a: 165 vs 161).  VNECT_B200_CHAIN=all forces it everywhere.
static bool can_chain(const ConvSpec& s, const ConvOpts& o, int cout) {
  if (!chain_enabled() || o.chain_scope.empty() || s.block_n != kTailBlockN || s.cg != 2 || cout % kTailBlockN != 0) return false;
  static const bool all = [] { const char* e = getenv("VNECT_B200_CHAIN"); return e && strcmp(e, "all") == 0; }();
  if (all) return o.chain_cout == 64 || o.chain_cout == 128 || o.chain_cout == 256;
  return o.chain_cout == 64 || (o.chain_cout == 128 && s.residual != nullptr);
}

// Output-channel tile.  Throughput plans (many forwards) use the widest tile; latency plans (a handful of forwards, e.g.
// the drop-in estimator: one frame x n_scales) use 64-wide tiles so that 4x more CTAs share each layer's serial K loop.
static int pick_block_n(int n, int cap_fw) {
  if (cap_fw <= 8) return 64;
  return n >= 256 ? 256 : n >= 128 ? 128 : 64;
}

// CTA pairs (tcgen05 cta_group::2) pay off where the MMA's smem operand reads are the limit: N >= 128 tiles
// (measured 5-8 % on the tensor-bound layers, nothing at N = 64).  VNECT_B200_PAIRS=0 / 64 moves the threshold (A/B).
static int pick_cg(int block_n) {
  static const int min_n = [] {
    const char* e = getenv("VNECT_B200_PAIRS");
    return e ? (atoi(e) > 0 ? atoi(e) : 1 << 30) : 128;
  }();
  return block_n >= min_n ? 2 : 1;
}

static int add_conv(vnect_t* h, const std::string& scope, int k, const std::string& in, const std::string& out,
                    int cin, int cout, const ConvOpts& o) {
  const Act& ai = h->acts.at(in);
  const HostVar& w = h->vars.at(scope + "/weights");
  const HostVar& b = h->vars.at(scope + "/biases");
  const int cin_pad = ai.C;
  __half* dw = nullptr;
  float* db = nullptr;
  int rc = upload(h, to_half(pack_conv(w, k, cin, cout, cin_pad, cout)), &dw);
  if (rc) return rc;
  rc = upload(h, b.data, &db);
  if (rc) return rc;
  const int OH = ai.H / o.in_stride, OW = ai.W / o.in_stride;
  rc = new_act(h, out, OH, OW, cout);
  if (rc) return rc;
  ConvSpec s;
  s.kind = k == 1 ? CONV_1x1 : CONV_3x3;
  s.NB = h->cap_fw; s.H = OH; s.W = OW;
  s.in = ai.p; s.cin_pad = cin_pad; s.in_stride = o.in_stride;
  s.w = dw; s.n_pad = cout; s.n_valid = cout; s.block_n = pick_block_n(cout, h->cap_fw); s.cg = pick_cg(s.block_n);
  s.bias = db; s.relu_cols = o.relu ? cout : 0;
  if (!o.residual.empty()) {
    const Act& r = h->acts.at(o.residual);
    if (r.H != OH * o.res_stride || r.W != OW * o.res_stride || r.C != cout)
      return fail(h, VNECT_E_INVALID, "residual shape mismatch at %s", scope.c_str());
    s.residual = r.p; s.ldr = r.C; s.res_stride = o.res_stride;
  }
  // outputs leave through swizzled smem + TMA tensor stores; residual tiles are prefetched by TMA
  s.out = h->acts.at(out).p; s.ldc = cout;
  s.epi = s.residual ? EPI_TMA_RES : EPI_TMA;
  // 64-channel convs whose whole weight matrix fits 9 K blocks keep it resident in smem: the 3x3s of res2 are bound
  // by L2->SM traffic (158 -> 135 us), the 1x1 256->64 convs gain a few percent
  static const bool bres_on = [] { const char* e = getenv("VNECT_B200_BRES"); return !(e && atoi(e) == 0); }();
  if (bres_on && cout == 64 && s.block_n == 64 && s.cg == 1 && s.epi == EPI_TMA && k * k * cin_pad / 64 <= kMaxResidentKBlocks)
    s.b_resident = 1;
  // stride-1 3x3 convs on the larger grids read one halo patch per channel block instead of nine shifted boxes
  // (conv_gemm.cuh HALO); VNECT_B200_HALO=0 turns it off (A/B measurement)
  static const bool halo_on = [] { const char* e = getenv("VNECT_B200_HALO"); return !(e && atoi(e) == 0); }();
  if (halo_on && k == 3 && o.in_stride == 1 && s.epi == EPI_TMA && halo_tiling_ok(OH, OW) &&
      (s.block_n == 64 ? s.cg == 1 : (s.block_n == 128)))
    s.halo = 1;
  // the other stride-1 3x3 convs (23 x 23 at box size 368: five 23 x 5 tiles fill 529 of 640 GEMM rows) run on flat pixel
  // rows through an im2col tensor map: every row of a 128-row tile is a real pixel.  Same K order, so the same results.
  // VNECT_B200_IM2COL=0 turns it off (A/B measurement)
  static const bool im2col_on = [] { const char* e = getenv("VNECT_B200_IM2COL"); return !(e && atoi(e) == 0); }();
  // (stride-2 convs included: res3d's 3x3 at 46 -> 23, 31 -> 27 us; the resident-weight 64-channel ones gain nothing)
  if (im2col_on && k == 3 && !s.halo && !s.b_resident && cin_pad % 64 == 0) s.im2col = 1;
  Step st;
  st.kind = 0; st.name = scope;
  if (k == 1 && can_chain(s, o, cout)) {
    if ((rc = attach_chain(h, s, o, cout, OH, OW))) return rc;
    st.name = scope + ">" + o.chain_scope;
  }
  std::string err;
  if (!build_conv(s, h->num_sms, &st.launch, &err)) return fail(h, VNECT_E_CUDA, "%s: %s", scope.c_str(), err.c_str());
  h->steps.push_back(st);
  return VNECT_OK;
}

// Last conv of a projection block with the shortcut folded in (vnect_model.py:32-41, 64-73, 106-115, 168-177):
//   relu(W_2c * b + bias_2c + W_1 * x + bias_1) == relu([W_2c | W_1] * [b ; x] + (bias_2c + bias_1))
// one GEMM whose K runs over the 3x3's output and then over the block input; the branch1 tensor never exists.
static int add_proj_tail(vnect_t* h, const std::string& scope2c, const std::string& scope1, const std::string& in_b,
                         const std::string& in_x, const std::string& out, int mid, int cin, int cout,
                         const ConvOpts& o = ConvOpts()) {
  const Act& ab = h->acts.at(in_b);
  const Act& ax = h->acts.at(in_x);
  if (ab.H != ax.H || ab.W != ax.W || ab.C != mid || ax.C != cin)
    return fail(h, VNECT_E_INVALID, "projection fold shape mismatch at %s", scope2c.c_str());
  const HostVar& w2 = h->vars.at(scope2c + "/weights");
  const HostVar& w1 = h->vars.at(scope1 + "/weights");
  const HostVar& b2 = h->vars.at(scope2c + "/biases");
  const HostVar& b1 = h->vars.at(scope1 + "/biases");
  const int K = mid + cin;
  std::vector<float> w((size_t)cout * K), bias(cout);
  for (int co = 0; co < cout; ++co) {
    for (int ci = 0; ci < mid; ++ci) w[(size_t)co * K + ci] = w2.data[(size_t)ci * cout + co];
    for (int ci = 0; ci < cin; ++ci) w[(size_t)co * K + mid + ci] = w1.data[(size_t)ci * cout + co];
    bias[co] = b2.data[co] + b1.data[co];
  }
  __half* dw = nullptr;
  float* db = nullptr;
  int rc = upload(h, to_half(w), &dw);
  if (rc) return rc;
  if ((rc = upload(h, bias, &db))) return rc;
  if ((rc = new_act(h, out, ab.H, ab.W, cout))) return rc;
  ConvSpec s;
  s.kind = CONV_1x1;
  s.NB = h->cap_fw; s.H = ab.H; s.W = ab.W;
  s.in = ab.p; s.cin_pad = mid; s.in2 = ax.p; s.cin2_pad = cin;
  s.w = dw; s.n_pad = cout; s.n_valid = cout; s.block_n = pick_block_n(cout, h->cap_fw); s.cg = pick_cg(s.block_n);
  s.bias = db; s.relu_cols = cout;
  s.out = h->acts.at(out).p; s.ldc = cout; s.epi = EPI_TMA;
  Step st;
  st.kind = 0; st.name = scope2c + "+" + scope1;
  if (can_chain(s, o, cout)) {
    if ((rc = attach_chain(h, s, o, cout, ab.H, ab.W))) return rc;
    st.name += ">" + o.chain_scope;
  }
  std::string err;
  if (!build_conv(s, h->num_sms, &st.launch, &err)) return fail(h, VNECT_E_CUDA, "%s: %s", st.name.c_str(), err.c_str());
  h->steps.push_back(st);
  return VNECT_OK;
}

// bottleneck block (reference: src/vnect_model.py:31-177); `a_override` replaces the 3x3's input (the res2c wiring,
// :56).  even_only: the block's output feeds nothing but stride-2 1x1 convs (res2c -> res3a, res3d -> res4a), so
// the 3x3, the last 1x1 and the residual add are evaluated at even pixels only -- same values, a quarter of the work.
static int add_block(vnect_t* h, const std::string& pre, const std::string& in, int cin, int mid, int cout, bool proj,
                     const std::string& suf, bool even_only, const std::string& a_override = "",
                     const std::string& chain_scope = "", int chain_cout = 0) {
  int rc;
  ConvOpts relu;
  std::string a = a_override;
  if (a.empty()) {
    a = pre + "_branch2a" + suf;
    if (!h->acts.count(a)) {  // otherwise the previous block's fused tail has already computed it
      rc = add_conv(h, a, 1, in, a, cin, mid, relu);
      if (rc) return rc;
    }
  }
  ConvOpts mid3; mid3.in_stride = even_only ? 2 : 1;
  rc = add_conv(h, pre + "_branch2b" + suf, 3, a, pre + "_branch2b" + suf, mid, mid, mid3);
  if (rc) return rc;
  ConvOpts last; last.relu = true;
  last.chain_scope = chain_scope; last.chain_cout = chain_cout;
  if (proj)
    return add_proj_tail(h, pre + "_branch2c" + suf, pre + "_branch1" + suf, pre + "_branch2b" + suf, in, pre, mid, cin, cout, last);
  // the residual is read at the pixels this block is evaluated at: stride 2 into a full-size input, stride 1 when the
  // input itself is already an even-pixel tensor (res2c reading res2b)
  last.residual = in;
  last.res_stride = h->acts.at(in).H / h->acts.at(pre + "_branch2b" + suf).H;
  return add_conv(h, pre + "_branch2c" + suf, 1, pre + "_branch2b" + suf, pre, mid, cout, last);
}

static void set_batch(ConvLaunch& L, int nb, int sms) {
  ConvGemmParams& p = L.p;
  p.NB = nb;
  p.M = nb * p.H * p.W;
  p.num_m_tiles = p.mode == 0 ? (p.M + kBlockM - 1) / kBlockM : nb * p.tiles_x * p.tiles_y;
  L.grid = L.n2 ? tail_grid(p, sms) : conv_grid(p, L.cg, sms);
}

// cv2 INTER_LINEAR coordinate rule on the host, identical arithmetic to cv_linear_coord (double -> float)
static void host_linear_coord(int d, double inv_scale, int src, bool reset, int* i_out, float* f_out) {
  const double fd = ((double)d + 0.5) * inv_scale - 0.5;
  float f = (float)fd;
  int i = (int)std::floor(f);
  f = f - (float)i;
  if (reset) {
    if (i < 0) { i = 0; f = 0.f; }
    if (i >= src - 1) { i = src - 1; f = 0.f; }
  }
  *i_out = i;
  *f_out = f;
}

static inline int cv_round_host(double v) { return (int)std::nearbyint(v); }

static int build_tables(vnect_t* h) {
  std::vector<ScaleTable> t(h->n_scales);
  const int hs = h->hs;
  for (int s = 0; s < h->n_scales; ++s) {
    ScaleTable& T = t[s];
    memset(&T, 0, sizeof T);
    const double sc = h->cfg.scales[s];
    if (sc == 1.0) {
      T.identity = 1;
      continue;
    }
    // estimator.py:112-120: rescale = 1.0 / s; cv2.resize(fx = fy = rescale); centre crop of hs cells
    const double rescale = 1.0 / sc;
    const int R = cv_round_host(hs * rescale);
    const double inv = 1.0 / rescale;
    const int crop0 = R / 2 - hs / 2;
    for (int c = 0; c < hs; ++c) {
      int i;
      float f;
      host_linear_coord(c + crop0, inv, hs, true, &i, &f);
      T.i0[c] = (short)i;
      T.i1[c] = (short)std::min(i + 1, hs - 1);
      T.a0[c] = 1.f - f;
      T.a1[c] = f;
      host_linear_coord(c + crop0, inv, hs, false, &i, &f);
      T.j0[c] = (short)std::min(std::max(i, 0), hs - 1);
      T.j1[c] = (short)std::min(std::max(i + 1, 0), hs - 1);
      T.b0[c] = 1.f - f;
      T.b1[c] = f;
      int b0_bits, b1_bits;
      memcpy(&b0_bits, &T.b0[c], 4);
      memcpy(&b1_bits, &T.b1[c], 4);
      T.rowpk[c] = make_int4(T.j0[c] * hs, T.j1[c] * hs, b0_bits, b1_bits);
    }
    T.row_lo = hs - 1;
    T.row_hi = 0;
    for (int c = 0; c < hs; ++c) {
      T.row_lo = std::min<int>(T.row_lo, T.j0[c]);
      T.row_hi = std::max<int>(T.row_hi, T.j1[c]);
    }
  }
  h->h_tables = t;
  return upload(h, t, &h->d_tables);
}

static int alloc_prepost(vnect_t* h) {
  const int S = h->S, nb = h->cap_fw;
  int rc;
  // stem input: parity-split, zero-padded NHWC4 (zeros are written once here and never touched again)
  h->stem_rpp = S / 2 + 3;
  h->stem_pitch = (S + 6) * 4;
  // + slack: the last strip of the last image reads a few KB past its parity plane (stem_roll.cuh)
  if ((rc = dev_alloc(h, &h->x1, (size_t)nb * 2 * h->stem_rpp * h->stem_pitch + 16384, true))) return rc;
  if ((rc = dev_alloc(h, &h->maps, (size_t)nb * 84 * h->hs * h->hs, true))) return rc;
  const int mf = h->cfg.max_frames, ms = h->cfg.max_streams;
  h->d_frames_bytes = (size_t)mf * h->cfg.max_input_h * h->cfg.max_input_w * 3;
  CU(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (auto& L : h->lanes) {
    if ((rc = dev_alloc(h, &L.d_frames, h->d_frames_bytes))) return rc;
    {  // ids | t2d | t3d in one block on both sides: one H2D copy per call
      const size_t off_t2 = ((size_t)mf * sizeof(int) + 15) & ~size_t(15), off_t3 = off_t2 + (size_t)mf * sizeof(double);
      L.meta_bytes = off_t3 + (size_t)mf * sizeof(double);
      uint8_t* d_meta = nullptr;
      if ((rc = dev_alloc(h, &d_meta, L.meta_bytes))) return rc;
      L.d_stream_ids = reinterpret_cast<int*>(d_meta);
      L.d_t2d = reinterpret_cast<double*>(d_meta + off_t2);
      L.d_t3d = reinterpret_cast<double*>(d_meta + off_t3);
      CU(h, cudaMallocHost(&L.h_meta, L.meta_bytes));
      memset(L.h_meta, 0, L.meta_bytes);
      L.h_stream_ids = reinterpret_cast<int*>(L.h_meta);
      L.h_t2d = reinterpret_cast<double*>(static_cast<uint8_t*>(L.h_meta) + off_t2);
      L.h_t3d = reinterpret_cast<double*>(static_cast<uint8_t*>(L.h_meta) + off_t3);
    }
    if ((rc = dev_alloc(h, &L.d_out2d, (size_t)mf * kJoints * 2))) return rc;
    if ((rc = dev_alloc(h, &L.d_out3d, (size_t)mf * kJoints * 3))) return rc;
    CU(h, cudaHostAlloc(&L.h_nonfinite, sizeof(unsigned int), cudaHostAllocMapped));
    *L.h_nonfinite = 0;
    CU(h, cudaEventCreateWithFlags(&L.copy_done, cudaEventDisableTiming));
    CU(h, cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
  }
  if ((rc = dev_alloc(h, &h->d_sq, (size_t)mf * S * S * 3))) return rc;
  if ((rc = dev_alloc(h, &h->d_f32_in, (size_t)nb * S * S * 3))) return rc;
  if ((rc = build_tables(h))) return rc;
  if ((rc = dev_alloc(h, &h->d_st2d, (size_t)ms * kJoints * 2))) return rc;
  if ((rc = dev_alloc(h, &h->d_st3d, (size_t)ms * kJoints * 3))) return rc;
  if ((rc = dev_alloc(h, &h->d_j2_box, (size_t)mf * kJoints * 2))) return rc;
  if ((rc = dev_alloc(h, &h->d_j3_raw, (size_t)mf * kJoints * 3))) return rc;
  if ((rc = dev_alloc(h, &h->d_raw_argmax, (size_t)mf * kJoints * 2))) return rc;
  if ((rc = dev_alloc(h, &h->d_counter, mf))) return rc;
  if ((rc = dev_alloc(h, &h->d_prep3, (size_t)mf * kJoints * 9))) return rc;
  if (getenv("VNECT_B200_POST_TRACE") && (rc = dev_alloc(h, &h->d_post_trace, (size_t)mf * kJoints * 16))) return rc;
  if ((rc = dev_alloc(h, &h->d_filter_scratch, 64))) return rc;
  if ((rc = dev_alloc(h, &h->d_st_ang, (size_t)ms * 8))) return rc;
  if ((rc = dev_alloc(h, &h->d_ang_in, (size_t)mf * kJoints * 3))) return rc;
  if ((rc = dev_alloc(h, &h->d_ang_t, mf))) return rc;
  if ((rc = dev_alloc(h, &h->d_ang_ids, mf))) return rc;
  if ((rc = dev_alloc(h, &h->d_ang_out, (size_t)mf * 8))) return rc;
  h->last_t_ang.assign(ms, NAN);
  if ((rc = dev_alloc(h, &h->d_nonfinite, 1))) return rc;
  if ((rc = dev_alloc(h, &h->d_scan, 1))) return rc;
  if ((rc = dev_alloc(h, &h->d_boxes, ms))) return rc;
  if ((rc = dev_alloc(h, &h->d_geoms, mf))) return rc;
  if ((rc = dev_alloc(h, &h->d_boxes_used, (size_t)mf * 4))) return rc;
  h->last_t2d.assign(ms, NAN);
  h->last_t3d.assign(ms, NAN);
  PyramidParams& py = h->pyr;
  py.S = S; py.n_scales = h->n_scales;
  py.rows_per_parity = h->stem_rpp; py.row_pitch = h->stem_pitch;
  std::vector<PyramidTable> ptab(h->n_scales);
  for (int i = 0; i < h->n_scales; ++i) {
    const double sc = h->cfg.scales[i];
    py.R[i] = sc < 1.0 ? cv_round_host(S * sc) : S;  // estimator.py:77: only scales < 1 are resized
    py.pad0[i] = (S - py.R[i]) / 2;
    py.inv_scale[i] = 1.0 / sc;
    memset(&ptab[i], 0, sizeof(PyramidTable));
    for (int d = 0; d < py.R[i] && py.R[i] != S; ++d) {  // same arithmetic as cv_linear_coord / cv_coef on the device
      int ix, iy;
      float fx, fy;
      host_linear_coord(d, py.inv_scale[i], S, true, &ix, &fx);
      host_linear_coord(d, py.inv_scale[i], S, false, &iy, &fy);
      const unsigned int a0 = (unsigned int)std::nearbyint((1.f - fx) * 2048.f), a1 = (unsigned int)std::nearbyint(fx * 2048.f);
      ptab[i].xt[d].x = (unsigned int)(4 * ix) | ((unsigned int)(4 * std::min(ix + 1, S - 1)) << 16);
      ptab[i].xt[d].y = a0 | (a1 << 16);
      ptab[i].yt[d].x = (short)std::min(std::max(iy, 0), S - 1);
      ptab[i].yt[d].y = (short)std::min(std::max(iy + 1, 0), S - 1);
      ptab[i].yt[d].z = (short)std::nearbyint((1.f - fy) * 2048.f);
      ptab[i].yt[d].w = (short)std::nearbyint(fy * 2048.f);
    }
    py.r_magic[i] = (unsigned int)((1ull << 32) / (unsigned long long)py.R[i]) + 1u;
    // source rows one block of kPyrRows output rows samples
    for (int y0 = 0; y0 < py.R[i] && py.R[i] != S; y0 += 1) {
      const int y1 = std::min(y0 + kPyrRows, py.R[i]) - 1;
      h->pyr_src_rows = std::max(h->pyr_src_rows, ptab[i].yt[y1].y - ptab[i].yt[y0].x + 1);
    }
  }
  if ((rc = upload(h, ptab, &h->d_pyr_tables))) return rc;
  py.tables = h->d_pyr_tables;
  // float32(v) / 255 - 0.4 -> fp16 (estimator.py:81 in float32, then the operand precision of the stem)
  std::vector<uint32_t> lut(256);
  for (int v = 0; v < 256; ++v) {
    const __half hv = __float2half_rn((float)v / 255.f - 0.4f);
    unsigned short bits;
    memcpy(&bits, &hv, 2);
    lut[v] = bits;
  }
  if ((rc = upload(h, lut, &h->d_norm_lut))) return rc;
  py.lut = h->d_norm_lut;
  py.q_magic = (unsigned int)((1ull << 32) / (unsigned long long)(S / 4)) + 1u;
  py.full = 0;
  return VNECT_OK;
}

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

const char* vnect_version(void) { return "vnect_b200 0.1 (sm_100a; fp16 operands, fp32 accumulate)"; }

const char* vnect_last_error(vnect_t* h) { return h ? h->err.c_str() : "null handle"; }

int vnect_create(vnect_t** out, const vnect_config* cfg) {
  if (!out || !cfg) return VNECT_E_INVALID;
  *out = nullptr;
  vnect_t* h = new vnect_handle();
  *out = h;  // returned even on failure so the caller can read the error, then destroy
  h->cfg = *cfg;
  if (cfg->box_size < 64 || cfg->box_size % 16 != 0 || cfg->box_size / 8 > kMaxHm || cfg->box_size > kMaxBox)
    return fail(h, VNECT_E_INVALID, "box_size %d must be a multiple of 16 in [64, %d]", cfg->box_size, kMaxHm * 8);
  if (cfg->n_scales < 1 || cfg->n_scales > kMaxScales) return fail(h, VNECT_E_INVALID, "n_scales out of range");
  for (int i = 0; i < cfg->n_scales; ++i)
    if (!(cfg->scales[i] > 0.0 && cfg->scales[i] <= 1.0))
      return fail(h, VNECT_E_INVALID, "scale %g not in (0, 1]", cfg->scales[i]);
  if (cfg->max_frames < 1 || cfg->max_streams < 1) return fail(h, VNECT_E_INVALID, "max_frames/max_streams must be >= 1");
  h->S = cfg->box_size;
  h->hs = h->S / 8;
  h->n_scales = cfg->n_scales;
  h->cap_fw = cfg->max_frames * cfg->n_scales;
  if (h->cfg.max_input_h <= 0) h->cfg.max_input_h = h->S;
  if (h->cfg.max_input_w <= 0) h->cfg.max_input_w = h->S;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, VNECT_E_CUDA, "no CUDA device: this library has no CPU path");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(h, VNECT_E_INVALID, "device %d not in [0, %d)", cfg->device, ndev);
  ON_DEVICE(h);  // the caller's current device is restored on return
  cudaDeviceProp prop;
  CU(h, cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail(h, VNECT_E_CUDA, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
  h->num_sms = prop.multiProcessorCount;
  CU(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  h->use_graphs = getenv("VNECT_B200_NO_GRAPH") == nullptr;
  int rc = alloc_prepost(h);
  if (rc) return rc;
  return vnect_reset_stream(h, -1);
}

int vnect_set_weight(vnect_t* h, const char* tf_name, const float* data, const int64_t* shape, int32_t rank) {
  if (!h || !tf_name || !data || !shape) return fail(h, VNECT_E_INVALID, "null argument");
  if (h->finalized) return fail(h, VNECT_E_INVALID, "weights are frozen after vnect_finalize");
  std::vector<int64_t> want;
  if (!expected_shape(tf_name, &want)) return fail(h, VNECT_E_WEIGHT, "unknown variable '%s'", tf_name);
  if ((int)want.size() != rank) return fail(h, VNECT_E_WEIGHT, "'%s': rank %d, expected %zu", tf_name, rank, want.size());
  size_t n = 1;
  for (int i = 0; i < rank; ++i) {
    if (shape[i] != want[i]) return fail(h, VNECT_E_WEIGHT, "'%s': dim %d is %lld, expected %lld", tf_name, i, (long long)shape[i], (long long)want[i]);
    n *= (size_t)shape[i];
  }
  HostVar& v = h->vars[tf_name];
  v.data.assign(data, data + n);
  v.shape = want;
  return VNECT_OK;
}

int vnect_finalize(vnect_t* h) {
  if (!h) return VNECT_E_INVALID;
  if (h->finalized) return fail(h, VNECT_E_INVALID, "already finalized");
  ON_DEVICE(h);
  // every variable of the graph must be present (the reference restores all 109 from the checkpoint)
  for (const ConvDef& c : conv_defs())
    for (const char* leaf : {"/weights", "/biases"})
      if (!h->vars.count(c.scope + leaf)) return fail(h, VNECT_E_WEIGHT, "missing variable %s%s", c.scope.c_str(), leaf);
  for (const char* k : {"res5c_branch1a/kernel", "res5c_branch2a/kernel", "res5c_branch2c/kernel"})
    if (!h->vars.count(k)) return fail(h, VNECT_E_WEIGHT, "missing variable %s", k);
  for (const char* v : kBnVars)
    if (!h->vars.count(std::string("bn5c_branch2a/") + v)) return fail(h, VNECT_E_WEIGHT, "missing variable bn5c_branch2a/%s", v);

  const int S = h->S, nb = h->cap_fw;
  int rc;
  {  // conv1 + pool1 (vnect_model.py:27-29) fused: rolling raw-strip implicit GEMM + max-pool (stem_roll.cuh)
    std::vector<__half> wk = to_half(pack_stem(h->vars.at("conv1/weights"))), ws(wk.size());
    pack_stem_stacked(wk.data(), ws.data());
    std::vector<__half> wp((size_t)kPairWBytes);  // two ranks x kPairWBytes / 2 halves
    pack_stem_pair(wk.data(), wp.data());
    __half *dws = nullptr, *dwp = nullptr;
    float* db = nullptr;
    if ((rc = upload(h, ws, &dws))) return rc;
    if ((rc = upload(h, wp, &dwp))) return rc;
    if ((rc = upload(h, h->vars.at("conv1/biases").data, &db))) return rc;
    if ((rc = new_act(h, "pool1", S / 4, S / 4, 64))) return rc;
    Step st;
    st.kind = 3; st.name = "conv1+pool1";
    std::string err;
    if (!build_stem_pool(h->x1, S, h->stem_rpp, h->stem_pitch, dws, db, h->acts.at("pool1").p, nb, h->num_sms, &st.stem_pool, &err,
                         dwp, (size_t)nb * 2 * h->stem_rpp * h->stem_pitch + 16384))
      return fail(h, VNECT_E_CUDA, "conv1+pool1: %s", err.c_str());
    h->steps.push_back(st);
  }
  // each block's tail also computes the next block's 1x1 reduce conv (block_tail.cuh) where the plan allows
  if ((rc = add_block(h, "res2a", "pool1", 64, 64, 256, true, "", false, "", "res2b_branch2a", 64))) return rc;
  // res2b's output feeds nothing but the residual add of res2c: res2c's 3x3 reads res2b_branch2a (vnect_model.py:56) and
  // res2c_branch2a is dead.  res2c is evaluated at even pixels only (below), so res2b's 3x3, last 1x1 and add are too.
  if ((rc = add_block(h, "res2b", "res2a", 256, 64, 256, false, "", true))) return rc;
  // res2c: its output only feeds stride-2 1x1 convs -> even pixels only; the stride-2 reduce of res3a is then a plain
  // 1x1 conv on the compact tensor and can ride on res2c's tail
  if ((rc = add_block(h, "res2c", "res2b", 256, 64, 256, false, "", true, "res2b_branch2a", "res3a_branch2a", 128))) return rc;
  if ((rc = add_block(h, "res3a", "res2c", 256, 128, 512, true, "", false, "", "res3b_branch2a", 128))) return rc;
  if ((rc = add_block(h, "res3b", "res3a", 512, 128, 512, false, "", false, "", "res3c_branch2a", 128))) return rc;
  if ((rc = add_block(h, "res3c", "res3b", 512, 128, 512, false, "", false, "", "res3d_branch2a", 128))) return rc;
  if ((rc = add_block(h, "res3d", "res3c", 512, 128, 512, false, "", true, "", "res4a_branch2a", 256))) return rc;
  if ((rc = add_block(h, "res4a", "res3d", 512, 256, 1024, true, "", false, "", "res4b_branch2a", 256))) return rc;
  {
    const char* names[] = {"res4b", "res4c", "res4d", "res4e", "res4f"};
    std::string prev = "res4a";
    for (int i = 0; i < 5; ++i) {
      // res4f feeds res5a_branch2a_new (512 channels: too wide for the second accumulator) -> not chained
      const std::string next = i < 4 ? std::string(names[i + 1]) + "_branch2a" : std::string();
      if ((rc = add_block(h, names[i], prev, 1024, 256, 1024, false, "", false, "", next, next.empty() ? 0 : 256))) return rc;
      prev = names[i];
    }
  }
  if ((rc = add_block(h, "res5a", "res4f", 1024, 512, 1024, true, "_new", false, "", "res5b_branch2a_new", 256))) return rc;
  ConvOpts relu;
  if (!h->acts.count("res5b_branch2a_new") &&
      (rc = add_conv(h, "res5b_branch2a_new", 1, "res5a", "res5b_branch2a_new", 1024, 256, relu))) return rc;
  if ((rc = add_conv(h, "res5b_branch2b_new", 3, "res5b_branch2a_new", "res5b_branch2b_new", 256, 128, relu))) return rc;
  if ((rc = add_conv(h, "res5b_branch2c_new", 1, "res5b_branch2b_new", "res5b_branch2c_new", 128, 256, relu))) return rc;
  {  // res5c head: two transposed convs + BN + ReLU + bone lengths + concat (vnect_model.py:188-209), one kernel
    const HostVar& g = h->vars.at("bn5c_branch2a/gamma");
    const HostVar& be = h->vars.at("bn5c_branch2a/beta");
    const HostVar& mu = h->vars.at("bn5c_branch2a/moving_mean");
    const HostVar& var = h->vars.at("bn5c_branch2a/moving_variance");
    std::vector<float> scale(128), bias(192, 0.f);
    for (int c = 0; c < 128; ++c) {
      scale[c] = g.data[c] / std::sqrt(var.data[c] + 0.001f);  // tc.layers.batch_norm epsilon default
      bias[c] = be.data[c] - mu.data[c] * scale[c];
    }
    __half* dw = nullptr;
    float* db = nullptr;
    rc = upload(h, to_half(pack_deconv(h->vars.at("res5c_branch2a/kernel"), h->vars.at("res5c_branch1a/kernel"), scale)), &dw);
    if (rc) return rc;
    if ((rc = upload(h, bias, &db))) return rc;
    if ((rc = new_act(h, "res5c_branch2a_feat", S / 8, S / 8, 256))) return rc;
    const Act& ai = h->acts.at("res5b_branch2c_new");
    ConvSpec s;
    s.kind = CONV_DECONV4;
    s.NB = nb; s.H = ai.H; s.W = ai.W; s.in = ai.p; s.cin_pad = 256;
    s.w = dw; s.n_pad = 192; s.n_valid = 191; s.block_n = 192; s.bias = db; s.relu_cols = 128; s.cg = pick_cg(192);
    s.out = h->acts.at("res5c_branch2a_feat").p; s.ldc = 256; s.epi = EPI_DECONV_HEAD;
    {  // flat pixel rows through the im2col tensor map (see add_conv): 114 -> 103 us per 128 forwards
      const char* e = getenv("VNECT_B200_IM2COL");
      if (!(e && atoi(e) == 0)) s.im2col = 1;
    }
    Step st;
    st.kind = 0; st.name = "res5c_deconv_head";
    std::string err;
    if (!build_conv(s, h->num_sms, &st.launch, &err)) return fail(h, VNECT_E_CUDA, "deconv head: %s", err.c_str());
    h->steps.push_back(st);
  }
  if ((rc = add_conv(h, "res5c_branch2b", 3, "res5c_branch2a_feat", "res5c_branch2b", 212, 128, relu))) return rc;
  {  // res5c_branch2c: 1x1 128 -> 84, no bias, linear (vnect_model.py:213); channel-planar fp32 for the post-process
    const HostVar& w = h->vars.at("res5c_branch2c/kernel");
    __half* dw = nullptr;
    rc = upload(h, to_half(pack_conv(w, 1, 128, 84, 128, 96)), &dw);
    if (rc) return rc;
    const Act& ai = h->acts.at("res5c_branch2b");
    ConvSpec s;
    s.kind = CONV_1x1;
    s.NB = nb; s.H = ai.H; s.W = ai.W; s.in = ai.p; s.cin_pad = 128;
    s.w = dw; s.n_pad = 96; s.n_valid = 84; s.block_n = 96; s.bias = nullptr; s.relu_cols = 0; s.cg = pick_cg(96);
    s.out = h->maps; s.ldc = 0; s.epi = EPI_PLANAR_F32;
    Step st;
    st.kind = 0; st.name = "res5c_branch2c";
    std::string err;
    if (!build_conv(s, h->num_sms, &st.launch, &err)) return fail(h, VNECT_E_CUDA, "res5c_branch2c: %s", err.c_str());
    h->steps.push_back(st);
  }
  h->conv_steps = 0;
  h->conv_steps = (int)h->steps.size();
  // ping-pong tile order: layer i walks its tiles in the opposite direction of layer i-1, so it starts on the data
  // that is still in the 126 MB L2 (the pyramid kernel writes forwards in ascending order, hence the stem reverses)
  const char* order = getenv("VNECT_B200_TILE_ORDER");  // "forward" turns the ping-pong off (profiling A/B only)
  const bool pingpong = !(order && strcmp(order, "forward") == 0);
  for (size_t i = 0; i < h->steps.size(); ++i) {
    const int rev = (pingpong && i % 2 == 0) ? 1 : 0;
    if (h->steps[i].kind == 0) h->steps[i].launch.p.reverse = rev;
    else h->steps[i].stem_pool.r.reverse = rev;
  }

  h->vars.clear();  // host copies no longer needed
  h->finalized = true;
  CU(h, cudaDeviceSynchronize());
  return VNECT_OK;
}

}  // extern "C"

// run the CNN on the first n forwards of x1 (already filled), on the handle's stream
static int run_forward(vnect_t* h, int n, cudaEvent_t* layer_events = nullptr) {
  int ei = 0;
  for (Step& st : h->steps) {
    if (layer_events) CU(h, cudaEventRecord(layer_events[ei++], h->stream));
    if (st.kind == 0) {
      set_batch(st.launch, n, h->num_sms);
      CU(h, launch_conv(st.launch, h->stream));
    } else {
      stem_pool_set_batch(st.stem_pool, n, h->num_sms);
      CU(h, launch_stem_pool(st.stem_pool, h->stream));
    }
    ++h->launches;
  }
  if (layer_events) CU(h, cudaEventRecord(layer_events[ei++], h->stream));
  return VNECT_OK;
}

struct Geometry {
  double scaler;
  int dh, dw, off_x, off_y, mode;
  bool alias;  // frames already are S x S: the pyramid reads them directly
};

// utils.img_scale_squarify geometry (utils.py:82-120)
static Geometry squarify_geometry(int S, int H, int W) {
  Geometry g;
  g.scaler = (double)S / (double)std::max(H, W);
  g.dw = cv_round_host(W * g.scaler);
  g.dh = cv_round_host(H * g.scaler);
  g.off_x = g.off_y = 0;
  if (g.dh > g.dw) g.off_x = S / 2 - g.dw / 2;
  else g.off_y = S / 2 - g.dh / 2;
  g.mode = (1.0 / g.scaler == 2.0) ? 1 : 0;  // OpenCV turns exact 2x INTER_LINEAR decimation into INTER_AREA
  g.alias = (H == S && W == S);
  return g;
}

// device frames -> x1 (stem layout) for n_frames * n_scales forwards
static int launch_pyramid(vnect_t* h, const uint8_t* sq, int64_t sq_pitch, int64_t sq_stride, int n_frames, bool full_surround) {
  const int S = h->S;
  PyramidParams py = h->pyr;
  py.n_frames = n_frames; py.sq_pitch = sq_pitch; py.sq_frame_stride = sq_stride;
  py.full = full_surround ? 1 : 0;
  const size_t smem = (size_t)h->pyr_src_rows * S * 4;  // staged source rows, one word per pixel
  static unsigned long long done = 0;
  CU(h, ensure_dyn_smem(pyramid_kernel, 160 * 1024, &done));
  if (smem > 160 * 1024) return fail(h, VNECT_E_INVALID, "pyramid scale too small for the staged rows (%zu bytes)", smem);
  CU(h, launch_pdl(pyramid_kernel, dim3((S + kPyrRows - 1) / kPyrRows, n_frames * h->n_scales), dim3(kPyrThreads), smem,
                   h->stream, sq, h->x1, py));
  ++h->launches;
  return VNECT_OK;
}

// The surround of every shrunken scale is a per-slot constant of x1: written here for all slots, once, and again only
// after vnect_forward has put caller-supplied images into the buffer.  Never part of a captured graph.
static int ensure_surround(vnect_t* h) {
  if (h->x1_surround_ok) return VNECT_OK;
  const int S = h->S;
  int rc = launch_pyramid(h, h->d_sq, (int64_t)S * 3, (int64_t)S * S * 3, h->cfg.max_frames, true);
  if (rc) return rc;
  h->x1_surround_ok = true;
  return VNECT_OK;
}

static int run_preprocess(vnect_t* h, const uint8_t* dev_bgr, int n_frames, int H, int W, int64_t pitch,
                          int64_t frame_stride, const Geometry& g, bool tracked = false) {
  const int S = h->S;
  int rc0 = ensure_surround(h);
  if (rc0) return rc0;
  const uint8_t* sq = dev_bgr;
  int64_t sq_pitch = pitch, sq_stride = frame_stride;
  if (tracked) {
    track_geometry_kernel<<<(n_frames + 63) / 64, 64, 0, h->stream>>>(h->d_boxes, h->cur->d_stream_ids, n_frames, S, H, W,
                                                                      h->d_geoms, h->d_boxes_used);
    CU(h, cudaGetLastError());
    ++h->launches;
  }
  if (!g.alias || tracked) {
    SquarifyParams sp;
    sp.geoms = tracked ? h->d_geoms : nullptr;
    sp.n_frames = n_frames; sp.H = H; sp.W = W; sp.pitch = pitch; sp.frame_stride = frame_stride; sp.S = S;
    sp.dh = g.dh; sp.dw = g.dw; sp.off_x = g.off_x; sp.off_y = g.off_y; sp.inv_scale = 1.0 / g.scaler; sp.mode = g.mode;
    const int64_t total = (int64_t)n_frames * S * S;
    squarify_kernel<<<grid_for(total, 256, h->num_sms), 256, 0, h->stream>>>(dev_bgr, h->d_sq, sp);
    CU(h, cudaGetLastError());
    ++h->launches;
    sq = h->d_sq; sq_pitch = (int64_t)S * 3; sq_stride = (int64_t)S * S * 3;
  }
  if (int rc = launch_pyramid(h, sq, sq_pitch, sq_stride, n_frames, false)) return rc;
  return VNECT_OK;
}

// validates ids / timestamps on the host exactly where the reference would raise, then stages them to the device
// waits until the lane's previous submission has fully completed (its pinned meta and device buffers are reusable)
static int lane_acquire(vnect_t* h, int lane) {
  h->cur = &h->lanes[lane];
  if (h->cur->pending) {
    CU(h, cudaEventSynchronize(h->cur->done));
    h->cur->pending = false;
    if (*reinterpret_cast<volatile unsigned int*>(h->cur->h_nonfinite) != 0) {  // the batch that just completed produced NaN / Inf maps: fail loudly, once
      unsigned int n = 0;  // error path only: the count lives on the device
      cudaStreamSynchronize(h->stream);
      cudaMemcpy(&n, h->d_nonfinite, sizeof n, cudaMemcpyDeviceToHost);
      *h->cur->h_nonfinite = 0;
      cudaMemset(h->d_nonfinite, 0, sizeof(unsigned int));
      return fail(h, VNECT_E_NUMERIC, "non-finite values in the CNN output maps (%u joint blocks): fp16 activations "
                  "overflow at 65504 -- are the weights normalised? (vnect_check_finite names the first layer)", n);
    }
  }
  return VNECT_OK;
}

// The post-process itself raises the lane's host-visible flag (PostParams::nonfinite_flag, a mapped pinned word) when
// it meets NaN / Inf maps, so no device -> host copy of a counter sits in the stream after every batch.
static int snapshot_nonfinite(vnect_t*) { return VNECT_OK; }

static int stage_frame_meta(vnect_t* h, int n_frames, const int32_t* stream_ids, const double* t2d, const double* t3d,
                            cudaStream_t copy_on) {
  if (n_frames < 1 || n_frames > h->cfg.max_frames) return fail(h, VNECT_E_INVALID, "n_frames %d not in [1, %d]", n_frames, h->cfg.max_frames);
  h->seen_stamp.resize(h->cfg.max_streams, 0);
  const unsigned int stamp = ++h->seen_epoch;
  for (int i = 0; i < n_frames; ++i) {
    const int sid = stream_ids ? stream_ids[i] : i;
    if (sid < 0 || sid >= h->cfg.max_streams) return fail(h, VNECT_E_INVALID, "stream id %d not in [0, %d)", sid, h->cfg.max_streams);
    const bool dup = h->seen_stamp[sid] == stamp;
    h->seen_stamp[sid] = stamp;
    if (dup) return fail(h, VNECT_E_INVALID, "stream id %d appears twice in one call (frames of a stream are sequential)", sid);
    if (h->cfg.filters) {
      if (!t2d || !t3d) return fail(h, VNECT_E_INVALID, "timestamps required when filters are on");
      // OneEuroFilter.py:65-66: freq = 1.0 / (timestamp - lasttime) when both are truthy
      if (h->last_t2d[sid] == h->last_t2d[sid] && h->last_t2d[sid] != 0.0 && t2d[i] != 0.0 && t2d[i] == h->last_t2d[sid])
        return fail(h, VNECT_E_ZERO_DT, "float division by zero (stream %d: repeated 2D timestamp %.17g)", sid, t2d[i]);
      if (h->last_t3d[sid] == h->last_t3d[sid] && h->last_t3d[sid] != 0.0 && t3d[i] != 0.0 && t3d[i] == h->last_t3d[sid])
        return fail(h, VNECT_E_ZERO_DT, "float division by zero (stream %d: repeated 3D timestamp %.17g)", sid, t3d[i]);
      // an EARLIER timestamp makes freq negative and every alpha leave (0, 1]: LowPassFilter raises ValueError
      // (OneEuroFilter.py:21-22).  Refused here, before any filter state is touched.
      if (h->last_t2d[sid] == h->last_t2d[sid] && h->last_t2d[sid] != 0.0 && t2d[i] != 0.0 && t2d[i] < h->last_t2d[sid])
        return fail(h, VNECT_E_INVALID, "alpha should be in (0.0, 1.0] (stream %d: 2D timestamp %.17g earlier than the previous %.17g)", sid, t2d[i], h->last_t2d[sid]);
      if (h->last_t3d[sid] == h->last_t3d[sid] && h->last_t3d[sid] != 0.0 && t3d[i] != 0.0 && t3d[i] < h->last_t3d[sid])
        return fail(h, VNECT_E_INVALID, "alpha should be in (0.0, 1.0] (stream %d: 3D timestamp %.17g earlier than the previous %.17g)", sid, t3d[i], h->last_t3d[sid]);
    }
  }
  for (int i = 0; i < n_frames; ++i) {
    const int sid = stream_ids ? stream_ids[i] : i;
    h->cur->h_stream_ids[i] = sid;
    h->cur->h_t2d[i] = t2d ? t2d[i] : 0.0;
    h->cur->h_t3d[i] = t3d ? t3d[i] : 0.0;
    if (h->cfg.filters) {
      h->last_t2d[sid] = t2d[i];
      h->last_t3d[sid] = t3d[i];
    }
  }
  CU(h, cudaMemcpyAsync(h->cur->d_stream_ids, h->cur->h_meta, h->cur->meta_bytes, cudaMemcpyHostToDevice, copy_on));
  return VNECT_OK;
}

static int run_postprocess(vnect_t* h, int n_frames, double scaler, int off_x, int off_y, double* dev_out2d,
                           float* dev_out3d, bool tracked = false, bool guard = true) {
  PostParams p;
  p.geoms = tracked ? h->d_geoms : nullptr;
  p.n_frames = n_frames; p.n_scales = h->n_scales; p.hs = h->hs; p.S = h->S;
  p.maps = h->maps; p.tables = h->d_tables;
  p.stream_ids = h->cur->d_stream_ids; p.t2d = h->cur->d_t2d; p.t3d = h->cur->d_t3d;
  p.st2d = h->d_st2d; p.st3d = h->d_st3d;
  p.cfg2d = {30.0, 1.7, 0.3, 0.4};  // estimator.py:34-39
  p.cfg3d = {30.0, 0.8, 0.4, 0.4};  // estimator.py:40-45
  p.filters_on = h->cfg.filters;
  p.scaler = scaler; p.off_x = off_x; p.off_y = off_y;
  p.j2_box = h->d_j2_box; p.j3_raw = h->d_j3_raw; p.raw_argmax = h->d_raw_argmax;
  p.frame_counter = h->d_counter;
  p.prep3 = h->d_prep3;
  p.out2d = dev_out2d; p.out3d = dev_out3d;
  p.packed = h->d_packed;
  p.nonfinite = guard ? h->d_nonfinite : nullptr;  // caller-supplied maps (vnect_postprocess) are taken as they are
  p.nonfinite_flag = nullptr;
  if (guard) CU(h, cudaHostGetDevicePointer(reinterpret_cast<void**>(&p.nonfinite_flag), h->cur->h_nonfinite, 0));
  // Threads per (frame, joint) block = the cells of `rows` heat-map rows.  A full batch is bound by instruction issue
  // summed over all blocks (two rows: every lane busy at hs = 46); a few frames leave most SMs empty, so wider blocks
  // shorten each block's own chain instead.  VNECT_B200_POST_ROWS overrides for A/B runs.
  static const int rows_env = [] {
    const char* e = getenv("VNECT_B200_POST_ROWS");
    return e ? atoi(e) : 0;
  }();
  const int rows = rows_env > 0 ? rows_env : n_frames >= 48 ? 2 : n_frames >= 12 ? 4 : 8;
  const int threads = post_threads(h->hs, rows);
  const int rpp = threads / h->hs;
  post_smem_plan(p, h->h_tables.data(), (h->hs / 2 / rpp) * rpp);
  p.trace = h->d_post_trace;
  const size_t smem = (size_t)p.smem_floats * sizeof(float);
  static unsigned long long done[kMaxScales] = {0, 0, 0, 0};
  auto launch = [&](auto kern, unsigned long long* mask) -> cudaError_t {
    if (cudaError_t e = ensure_dyn_smem(kern, 96 * 1024, mask); e != cudaSuccess) return e;
    return launch_pdl(kern, dim3(n_frames * kJoints), dim3(threads), smem, h->stream, p);
  };
  // the usual pyramid (scale 1.0 first, smaller ones after) gets its identity mask as a compile-time constant
  static unsigned long long done1[kMaxScales] = {0, 0, 0, 0};
  if (p.identity_mask == 1) {
    switch (h->n_scales) {
      case 1: CU(h, launch(postprocess_kernel<1, 1>, &done1[0])); break;
      case 2: CU(h, launch(postprocess_kernel<2, 1>, &done1[1])); break;
      case 3: CU(h, launch(postprocess_kernel<3, 1>, &done1[2])); break;
      default: CU(h, launch(postprocess_kernel<4, 1>, &done1[3])); break;
    }
  } else {
    switch (h->n_scales) {
      case 1: CU(h, launch(postprocess_kernel<1, -1>, &done[0])); break;
      case 2: CU(h, launch(postprocess_kernel<2, -1>, &done[1])); break;
      case 3: CU(h, launch(postprocess_kernel<3, -1>, &done[2])); break;
      default: CU(h, launch(postprocess_kernel<4, -1>, &done[3])); break;
    }
  }
  ++h->launches;
  return VNECT_OK;
}

extern "C" {

int vnect_forward(vnect_t* h, const float* nhwc, int32_t n, float* hm, float* xm, float* ym, float* zm) {
  if (!h || !h->finalized) return fail(h, VNECT_E_INVALID, "handle not finalized");
  ON_DEVICE(h);
  if (!nhwc || !hm || !xm || !ym || !zm) return fail(h, VNECT_E_INVALID, "null buffer");
  if (n < 1 || n > h->cap_fw) return fail(h, VNECT_E_INVALID, "n %d not in [1, %d]", n, h->cap_fw);
  const int S = h->S, hs = h->hs;
  const size_t in_elems = (size_t)n * S * S * 3;
  CU(h, cudaMemcpyAsync(h->d_f32_in, nhwc, in_elems * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  h->x1_surround_ok = false;  // caller-supplied images replace whole slots of x1
  f32_to_stem_kernel<<<grid_for((int64_t)n * S * S, 256, h->num_sms), 256, 0, h->stream>>>(h->d_f32_in, h->x1, n, S, h->stem_rpp, h->stem_pitch);
  CU(h, cudaGetLastError());
  ++h->launches;
  int rc = run_forward(h, n);
  if (rc) return rc;
  std::vector<float> planar((size_t)n * 84 * hs * hs);
  CU(h, cudaMemcpyAsync(planar.data(), h->maps, planar.size() * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  float* outs[4] = {hm, xm, ym, zm};
  const size_t plane = (size_t)hs * hs;
  for (int i = 0; i < n; ++i)
    for (int m = 0; m < 4; ++m)
      for (int j = 0; j < kJoints; ++j) {
        const float* src = planar.data() + ((size_t)i * 84 + m * kJoints + j) * plane;
        float* dst = outs[m] + (size_t)i * plane * kJoints + j;
        for (size_t px = 0; px < plane; ++px) dst[px * kJoints] = src[px];
      }
  return VNECT_OK;
}

// pre-process -> CNN -> post-process for one batch, as a replayed CUDA graph when possible
static int run_pipeline(vnect_t* h, int lane, const uint8_t* dev_bgr, int n_frames, int H, int W, int64_t pitch,
                        int64_t frame_stride, double* out2d, float* out3d, bool tracked = false) {
  const Geometry g = squarify_geometry(h->S, H, W);
  if (int rc = ensure_surround(h)) return rc;
  auto direct = [&]() -> int {
    int rc;
    if ((rc = run_preprocess(h, dev_bgr, n_frames, H, W, pitch, frame_stride, g, tracked))) return rc;
    if ((rc = run_forward(h, n_frames * h->n_scales))) return rc;
    if ((rc = run_postprocess(h, n_frames, g.scaler, g.off_x, g.off_y, out2d, out3d, tracked))) return rc;
    if (tracked) {
      track_update_kernel<<<n_frames, 32, 0, h->stream>>>(out2d, h->cur->d_stream_ids, n_frames, H, W, h->d_boxes);
      CU(h, cudaGetLastError());
      ++h->launches;
    }
    return VNECT_OK;
  };
  if (!h->use_graphs) return direct();
  const vnect_handle::GraphKey key(lane + (tracked ? 2 : 0), n_frames, H, W, (long long)pitch, (long long)frame_stride, dev_bgr, out2d, out3d);
  auto it = h->graphs.find(key);
  if (it == h->graphs.end()) {
    if (h->graphs.size() >= vnect_handle::kMaxGraphs) {  // evict the least recently used entry
      auto victim = h->graphs.begin();
      for (auto g = h->graphs.begin(); g != h->graphs.end(); ++g)
        if (g->second.last_use < victim->second.last_use) victim = g;
      if (victim->second.exec) cudaGraphExecDestroy(victim->second.exec);
      h->graphs.erase(victim);
    }
    h->graphs[key].last_use = ++h->graph_clock;
    return direct();  // warm-up run: sets function attributes, exercises every launch configuration
  }
  vnect_handle::GraphEntry& e = it->second;
  e.last_use = ++h->graph_clock;
  if (e.unusable) return direct();
  if (!e.exec) {
    const long long before = h->launches;
    if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      e.unusable = true;  // e.g. the legacy default stream cannot be captured
      return direct();
    }
    const int rc = direct();
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
    e.launches = h->launches - before;
    h->launches = before;
    if (rc != VNECT_OK || ce != cudaSuccess || graph == nullptr) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      e.unusable = true;
      return rc != VNECT_OK ? rc : direct();
    }
    const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      cudaGetLastError();
      e.exec = nullptr;
      e.unusable = true;
      return direct();
    }
  }
  CU(h, cudaGraphLaunch(e.exec, h->stream));
  h->launches += e.launches;
  return VNECT_OK;
}

int vnect_estimate_device(vnect_t* h, const uint8_t* dev_bgr, int32_t n_frames, int32_t H, int32_t W, int64_t pitch,
                          int64_t frame_stride, const int32_t* stream_ids, const double* t2d, const double* t3d,
                          double* dev_joints2d, float* dev_joints3d) {
  if (!h || !h->finalized) return fail(h, VNECT_E_INVALID, "handle not finalized");
  ON_DEVICE(h);
  if (!dev_bgr || !dev_joints2d || !dev_joints3d) return fail(h, VNECT_E_INVALID, "null buffer");
  if (H < 2 || W < 2 || pitch < (int64_t)W * 3) return fail(h, VNECT_E_INVALID, "bad frame geometry %dx%d pitch %lld", H, W, (long long)pitch);
  // alternate lanes for the per-call meta so the host can run one call ahead of the GPU
  int rc = lane_acquire(h, (int)(h->device_calls++ & 1));
  if (rc) return rc;
  // the per-call meta travels on the copy stream, under the previous call's kernels
  if ((rc = stage_frame_meta(h, n_frames, stream_ids, t2d, t3d, h->copy_stream))) return rc;
  CU(h, cudaEventRecord(h->cur->copy_done, h->copy_stream));
  CU(h, cudaStreamWaitEvent(h->stream, h->cur->copy_done, 0));
  const int lane = (int)(h->cur - h->lanes);
  if ((rc = run_pipeline(h, lane, dev_bgr, n_frames, H, W, pitch, frame_stride, dev_joints2d, dev_joints3d))) return rc;
  if ((rc = snapshot_nonfinite(h))) return rc;
  CU(h, cudaEventRecord(h->cur->done, h->stream));
  h->cur->pending = true;
  return VNECT_OK;
}

int vnect_submit(vnect_t* h, int32_t lane, const uint8_t* bgr, int32_t n_frames, int32_t H, int32_t W, int64_t pitch,
                 int64_t frame_stride, const int32_t* stream_ids, const double* t2d, const double* t3d, double* joints2d,
                 float* joints3d) {
  if (!h || !h->finalized) return fail(h, VNECT_E_INVALID, "handle not finalized");
  ON_DEVICE(h);
  if (lane < 0 || lane > 1) return fail(h, VNECT_E_INVALID, "lane must be 0 or 1");
  if (!bgr || !joints2d || !joints3d) return fail(h, VNECT_E_INVALID, "null buffer");
  if (n_frames < 1 || n_frames > h->cfg.max_frames) return fail(h, VNECT_E_INVALID, "n_frames %d not in [1, %d]", n_frames, h->cfg.max_frames);
  if (H < 2 || W < 2 || H > h->cfg.max_input_h || W > h->cfg.max_input_w)
    return fail(h, VNECT_E_INVALID, "frame %dx%d outside [2, max_input %dx%d]", H, W, h->cfg.max_input_h, h->cfg.max_input_w);
  if (pitch < (int64_t)W * 3) return fail(h, VNECT_E_INVALID, "pitch smaller than a row");
  int rc = lane_acquire(h, lane);
  if (rc) return rc;
  // host -> device on the copy stream (tightly packed on the device), overlapping the previous lane's kernels
  if ((rc = stage_frame_meta(h, n_frames, stream_ids, t2d, t3d, h->copy_stream))) return rc;
  const int64_t dpitch = (int64_t)W * 3, dstride = dpitch * H;
  if (pitch == dpitch && frame_stride == dstride) {
    CU(h, cudaMemcpyAsync(h->cur->d_frames, bgr, (size_t)dstride * n_frames, cudaMemcpyHostToDevice, h->copy_stream));
  } else {
    for (int i = 0; i < n_frames; ++i)
      CU(h, cudaMemcpy2DAsync(h->cur->d_frames + (size_t)i * dstride, dpitch, bgr + (size_t)i * frame_stride, pitch, dpitch, H, cudaMemcpyHostToDevice, h->copy_stream));
  }
  CU(h, cudaEventRecord(h->cur->copy_done, h->copy_stream));
  CU(h, cudaStreamWaitEvent(h->stream, h->cur->copy_done, 0));
  if ((rc = run_pipeline(h, lane, h->cur->d_frames, n_frames, H, W, dpitch, dstride, h->cur->d_out2d, h->cur->d_out3d))) return rc;
  CU(h, cudaMemcpyAsync(joints2d, h->cur->d_out2d, (size_t)n_frames * kJoints * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaMemcpyAsync(joints3d, h->cur->d_out3d, (size_t)n_frames * kJoints * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if ((rc = snapshot_nonfinite(h))) return rc;
  CU(h, cudaEventRecord(h->cur->done, h->stream));
  h->cur->pending = true;
  return VNECT_OK;
}

int vnect_track_set_box(vnect_t* h, int32_t stream_id, int32_t x, int32_t y, int32_t w, int32_t hh) {
  if (!h || !h->d_boxes) return fail(h, VNECT_E_INVALID, "handle not created");
  ON_DEVICE(h);
  if (stream_id < 0 || stream_id >= h->cfg.max_streams) return fail(h, VNECT_E_INVALID, "stream id out of range");
  if (x < 0 || y < 0 || w < 2 || hh < 2) return fail(h, VNECT_E_INVALID, "box must have x, y >= 0 and w, h >= 2");
  const int4 b = make_int4(x, y, w, hh);
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, cudaMemcpy(h->d_boxes + stream_id, &b, sizeof b, cudaMemcpyHostToDevice));
  return VNECT_OK;
}

int vnect_track_get_box(vnect_t* h, int32_t stream_id, int32_t* xywh) {
  if (!h || !h->d_boxes || !xywh) return fail(h, VNECT_E_INVALID, "handle not created / null buffer");
  ON_DEVICE(h);
  if (stream_id < 0 || stream_id >= h->cfg.max_streams) return fail(h, VNECT_E_INVALID, "stream id out of range");
  CU(h, cudaStreamSynchronize(h->stream));
  int4 b;
  CU(h, cudaMemcpy(&b, h->d_boxes + stream_id, sizeof b, cudaMemcpyDeviceToHost));
  xywh[0] = b.x; xywh[1] = b.y; xywh[2] = b.z; xywh[3] = b.w;
  return VNECT_OK;
}

int vnect_track(vnect_t* h, const uint8_t* frames, int32_t n_frames, int32_t FH, int32_t FW, int64_t pitch,
                int64_t frame_stride, const int32_t* stream_ids, const double* t2d, const double* t3d, double* joints2d,
                float* joints3d, int32_t* boxes_used) {
  if (!h || !h->finalized) return fail(h, VNECT_E_INVALID, "handle not finalized");
  ON_DEVICE(h);
  if (!frames || !joints2d || !joints3d) return fail(h, VNECT_E_INVALID, "null buffer");
  if (n_frames < 1 || n_frames > h->cfg.max_frames) return fail(h, VNECT_E_INVALID, "n_frames %d not in [1, %d]", n_frames, h->cfg.max_frames);
  if (FH < 2 || FW < 2 || FH > h->cfg.max_input_h || FW > h->cfg.max_input_w)
    return fail(h, VNECT_E_INVALID, "frame %dx%d outside [2, max_input %dx%d]", FH, FW, h->cfg.max_input_h, h->cfg.max_input_w);
  if (pitch < (int64_t)FW * 3) return fail(h, VNECT_E_INVALID, "pitch smaller than a row");
  int rc = lane_acquire(h, 0);
  if (rc) return rc;
  if ((rc = stage_frame_meta(h, n_frames, stream_ids, t2d, t3d, h->copy_stream))) return rc;
  const int64_t dpitch = (int64_t)FW * 3, dstride = dpitch * FH;
  for (int i = 0; i < n_frames; ++i)
    CU(h, cudaMemcpy2DAsync(h->cur->d_frames + (size_t)i * dstride, dpitch, frames + (size_t)i * frame_stride, pitch, dpitch, FH, cudaMemcpyHostToDevice, h->copy_stream));
  CU(h, cudaEventRecord(h->cur->copy_done, h->copy_stream));
  CU(h, cudaStreamWaitEvent(h->stream, h->cur->copy_done, 0));
  if ((rc = run_pipeline(h, 0, h->cur->d_frames, n_frames, FH, FW, dpitch, dstride, h->cur->d_out2d, h->cur->d_out3d, true))) return rc;
  CU(h, cudaMemcpyAsync(joints2d, h->cur->d_out2d, (size_t)n_frames * kJoints * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaMemcpyAsync(joints3d, h->cur->d_out3d, (size_t)n_frames * kJoints * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if (boxes_used)
    CU(h, cudaMemcpyAsync(boxes_used, h->d_boxes_used, (size_t)n_frames * 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if ((rc = snapshot_nonfinite(h))) return rc;
  CU(h, cudaEventRecord(h->cur->done, h->stream));
  h->cur->pending = true;
  return lane_acquire(h, 0);  // synchronises and reports non-finite maps
}

int vnect_wait(vnect_t* h, int32_t lane) {
  if (!h || lane < 0 || lane > 1) return fail(h, VNECT_E_INVALID, "bad handle or lane");
  ON_DEVICE(h);
  return lane_acquire(h, lane);
}

int vnect_estimate(vnect_t* h, const uint8_t* bgr, int32_t n_frames, int32_t H, int32_t W, int64_t pitch,
                   int64_t frame_stride, const int32_t* stream_ids, const double* t2d, const double* t3d,
                   double* joints2d, float* joints3d) {
  int rc = vnect_submit(h, 0, bgr, n_frames, H, W, pitch, frame_stride, stream_ids, t2d, t3d, joints2d, joints3d);
  if (rc) return rc;
  return vnect_wait(h, 0);
}

int vnect_preprocess(vnect_t* h, const uint8_t* bgr, int32_t n_frames, int32_t H, int32_t W, int64_t pitch,
                     int64_t frame_stride, float* out_nhwc, double* scaler_offsets) {
  if (!h || !h->x1) return fail(h, VNECT_E_INVALID, "handle not created");
  ON_DEVICE(h);
  if (!bgr || !out_nhwc) return fail(h, VNECT_E_INVALID, "null buffer");
  if (n_frames < 1 || n_frames > h->cfg.max_frames) return fail(h, VNECT_E_INVALID, "n_frames out of range");
  if (H < 2 || W < 2 || H > h->cfg.max_input_h || W > h->cfg.max_input_w) return fail(h, VNECT_E_INVALID, "frame size outside max_input");
  const int64_t dpitch = (int64_t)W * 3, dstride = dpitch * H;
  if (lane_acquire(h, 0)) return VNECT_E_CUDA;
  for (int i = 0; i < n_frames; ++i)
    CU(h, cudaMemcpy2DAsync(h->cur->d_frames + (size_t)i * dstride, dpitch, bgr + (size_t)i * frame_stride, pitch, dpitch, H, cudaMemcpyHostToDevice, h->stream));
  const Geometry g = squarify_geometry(h->S, H, W);
  int rc = run_preprocess(h, h->cur->d_frames, n_frames, H, W, dpitch, dstride, g);
  if (rc) return rc;
  const int n = n_frames * h->n_scales, S = h->S;
  stem_to_f32_kernel<<<grid_for((int64_t)n * S * S, 256, h->num_sms), 256, 0, h->stream>>>(h->x1, h->d_f32_in, n, S, h->stem_rpp, h->stem_pitch);
  CU(h, cudaGetLastError());
  ++h->launches;
  CU(h, cudaMemcpyAsync(out_nhwc, h->d_f32_in, (size_t)n * S * S * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  if (scaler_offsets) {
    scaler_offsets[0] = g.scaler;
    scaler_offsets[1] = g.off_x;
    scaler_offsets[2] = g.off_y;
  }
  return VNECT_OK;
}

int vnect_postprocess(vnect_t* h, const float* hm, const float* xm, const float* ym, const float* zm, int32_t n_frames,
                      const int32_t* stream_ids, const double* t2d, const double* t3d, double scaler, int32_t offset_x,
                      int32_t offset_y, double* joints2d, float* joints3d, int32_t* raw_argmax) {
  if (!h || !h->x1) return fail(h, VNECT_E_INVALID, "handle not created");
  ON_DEVICE(h);
  if (!hm || !xm || !ym || !zm || !joints2d || !joints3d) return fail(h, VNECT_E_INVALID, "null buffer");
  int rc = lane_acquire(h, 0);
  if (rc) return rc;
  if ((rc = stage_frame_meta(h, n_frames, stream_ids, t2d, t3d, h->stream))) return rc;
  const int n = n_frames * h->n_scales, hs = h->hs;
  const size_t plane = (size_t)hs * hs;
  std::vector<float> planar((size_t)n * 84 * plane);
  const float* ins[4] = {hm, xm, ym, zm};
  for (int i = 0; i < n; ++i)
    for (int m = 0; m < 4; ++m)
      for (int j = 0; j < kJoints; ++j) {
        float* dst = planar.data() + ((size_t)i * 84 + m * kJoints + j) * plane;
        const float* src = ins[m] + (size_t)i * plane * kJoints + j;
        for (size_t px = 0; px < plane; ++px) dst[px] = src[px * kJoints];
      }
  CU(h, cudaMemcpyAsync(h->maps, planar.data(), planar.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  if ((rc = run_postprocess(h, n_frames, scaler, offset_x, offset_y, h->cur->d_out2d, h->cur->d_out3d, false, false))) return rc;
  CU(h, cudaMemcpyAsync(joints2d, h->cur->d_out2d, (size_t)n_frames * kJoints * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaMemcpyAsync(joints3d, h->cur->d_out3d, (size_t)n_frames * kJoints * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if (raw_argmax)
    CU(h, cudaMemcpyAsync(raw_argmax, h->d_raw_argmax, (size_t)n_frames * kJoints * 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return VNECT_OK;
}

int vnect_reset_stream(vnect_t* h, int32_t stream_id) {
  if (!h || !h->d_st2d) return fail(h, VNECT_E_INVALID, "handle not finalized");
  ON_DEVICE(h);
  const int ms = h->cfg.max_streams;
  if (stream_id < -1 || stream_id >= ms) return fail(h, VNECT_E_INVALID, "stream id out of range");
  const int lo = stream_id < 0 ? 0 : stream_id, hi = stream_id < 0 ? ms : stream_id + 1;
  std::vector<FilterState> s2((size_t)(hi - lo) * kJoints * 2), s3((size_t)(hi - lo) * kJoints * 3);
  FilterState z;
  memset(&z, 0, sizeof z);
  z.freq = 30.0;  // estimator.py:35,41
  for (auto& s : s2) s = z;
  for (auto& s : s3) s = z;
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, cudaMemcpy(h->d_st2d + (size_t)lo * kJoints * 2, s2.data(), s2.size() * sizeof(FilterState), cudaMemcpyHostToDevice));
  CU(h, cudaMemcpy(h->d_st3d + (size_t)lo * kJoints * 3, s3.data(), s3.size() * sizeof(FilterState), cudaMemcpyHostToDevice));
  {
    FilterState za = z;
    za.freq = 120.0;  // joints2angles.py:36
    std::vector<FilterState> sa((size_t)(hi - lo) * 8, za);
    CU(h, cudaMemcpy(h->d_st_ang + (size_t)lo * 8, sa.data(), sa.size() * sizeof(FilterState), cudaMemcpyHostToDevice));
    for (int i = lo; i < hi; ++i) h->last_t_ang[i] = NAN;
  }
  for (int i = lo; i < hi; ++i) h->last_t2d[i] = h->last_t3d[i] = NAN;
  // the tracked crop box restarts as the whole frame (run_estimator.py:68: rect = 0, 0, W_img, H_img); the geometry
  // kernel clips it to the actual frame size
  std::vector<int4> full((size_t)(hi - lo), make_int4(0, 0, 1 << 30, 1 << 30));
  CU(h, cudaMemcpy(h->d_boxes + lo, full.data(), full.size() * sizeof(int4), cudaMemcpyHostToDevice));
  return VNECT_OK;
}

int vnect_filter(vnect_t* h, int32_t stream_id, int32_t dim, int32_t values_are_f32, double t, double* values) {
  if (!h || !h->x1 || !values) return fail(h, VNECT_E_INVALID, "handle not created / null buffer");
  ON_DEVICE(h);
  if (stream_id < 0 || stream_id >= h->cfg.max_streams || (dim != 2 && dim != 3)) return fail(h, VNECT_E_INVALID, "bad stream id or dim");
  std::vector<double>& last = dim == 2 ? h->last_t2d : h->last_t3d;
  if (last[stream_id] == last[stream_id] && last[stream_id] != 0.0 && t != 0.0) {
    if (t == last[stream_id])
      return fail(h, VNECT_E_ZERO_DT, "float division by zero (stream %d: repeated timestamp %.17g)", stream_id, t);
    if (t < last[stream_id])
      return fail(h, VNECT_E_INVALID, "alpha should be in (0.0, 1.0] (stream %d: timestamp %.17g earlier than the previous %.17g)", stream_id, t, last[stream_id]);
  }
  last[stream_id] = t;
  const int n = kJoints * dim;
  CU(h, cudaMemcpyAsync(h->d_filter_scratch, values, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  FilterState* st = dim == 2 ? h->d_st2d + (size_t)stream_id * kJoints * 2 : h->d_st3d + (size_t)stream_id * kJoints * 3;
  const FilterCfg cfg = dim == 2 ? FilterCfg{30.0, 1.7, 0.3, 0.4} : FilterCfg{30.0, 0.8, 0.4, 0.4};
  joint_filter_kernel<<<1, 64, 0, h->stream>>>(st, cfg, h->d_filter_scratch, t, dim, values_are_f32 ? 1 : 0);
  CU(h, cudaGetLastError());
  ++h->launches;
  CU(h, cudaMemcpyAsync(values, h->d_filter_scratch, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return VNECT_OK;
}

// waits for every submitted batch: the per-call scratch (raw argmax) and the per-stream state are then at rest
static int quiesce(vnect_t* h) {
  for (auto& L : h->lanes)
    if (L.pending) {
      CU(h, cudaEventSynchronize(L.done));
      L.pending = false;
    }
  CU(h, cudaStreamSynchronize(h->stream));
  return VNECT_OK;
}

int vnect_joints2angles(vnect_t* h, const float* joints3d, int32_t n, const int32_t* stream_ids, const double* t,
                        double* angles) {
  if (!h || !h->d_st_ang || !joints3d || !angles) return fail(h, VNECT_E_INVALID, "handle not created / null buffer");
  ON_DEVICE(h);
  if (n < 1 || n > h->cfg.max_frames) return fail(h, VNECT_E_INVALID, "n %d not in [1, %d]", n, h->cfg.max_frames);
  std::vector<int> ids(n);
  for (int i = 0; i < n; ++i) {
    ids[i] = stream_ids ? stream_ids[i] : i;
    if (ids[i] < 0 || ids[i] >= h->cfg.max_streams) return fail(h, VNECT_E_INVALID, "stream id out of range");
    if (t) {  // same clock rules as the joint filters (OneEuroFilter.py:21-22, 66)
      const double last = h->last_t_ang[ids[i]];
      if (last == last && last != 0.0 && t[i] != 0.0) {
        if (t[i] == last) return fail(h, VNECT_E_ZERO_DT, "float division by zero (stream %d: repeated timestamp %.17g)", ids[i], t[i]);
        if (t[i] < last) return fail(h, VNECT_E_INVALID, "alpha should be in (0.0, 1.0] (stream %d: timestamp %.17g earlier than the previous %.17g)", ids[i], t[i], last);
      }
    }
  }
  if (t) for (int i = 0; i < n; ++i) h->last_t_ang[ids[i]] = t[i];
  CU(h, cudaMemcpyAsync(h->d_ang_in, joints3d, (size_t)n * kJoints * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_ang_ids, ids.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (t) CU(h, cudaMemcpyAsync(h->d_ang_t, t, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  AnglesParams ap;
  ap.joints3d = h->d_ang_in; ap.stream_ids = h->d_ang_ids; ap.t = t ? h->d_ang_t : nullptr;
  ap.st = h->d_st_ang; ap.angles = h->d_ang_out; ap.n = n;
  joints2angles_kernel<<<(n + 63) / 64, 64, 0, h->stream>>>(ap);
  CU(h, cudaGetLastError());
  ++h->launches;
  CU(h, cudaMemcpyAsync(angles, h->d_ang_out, (size_t)n * 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));  // ids is a host temporary
  return VNECT_OK;
}

int vnect_get_raw_argmax(vnect_t* h, int32_t n_frames, int32_t* raw_argmax) {
  if (!h || !h->d_raw_argmax || !raw_argmax) return fail(h, VNECT_E_INVALID, "handle not created / null buffer");
  ON_DEVICE(h);
  if (n_frames < 1 || n_frames > h->cfg.max_frames) return fail(h, VNECT_E_INVALID, "n_frames out of range");
  if (int rc = quiesce(h)) return rc;
  CU(h, cudaMemcpy(raw_argmax, h->d_raw_argmax, (size_t)n_frames * kJoints * 2 * sizeof(int), cudaMemcpyDeviceToHost));
  return VNECT_OK;
}

// Per-stream temporal state as plain doubles (VNECT_STREAM_STATE_DOUBLES of them): 105 filters x {prev, s_x, s_dx,
// lasttime, freq, has_prev, has_time} (2D filters first, [21][2], then 3D, [21][3]), then the host-side last timestamps
// (2D, 3D; NaN = none) and the tracked box (x, y, w, h).
int vnect_export_stream_state(vnect_t* h, int32_t stream_id, double* state) {
  if (!h || !h->d_st2d || !state) return fail(h, VNECT_E_INVALID, "handle not created / null buffer");
  ON_DEVICE(h);
  if (stream_id < 0 || stream_id >= h->cfg.max_streams) return fail(h, VNECT_E_INVALID, "stream id out of range");
  if (int rc = quiesce(h)) return rc;
  std::vector<FilterState> st(kJoints * 5);
  CU(h, cudaMemcpy(st.data(), h->d_st2d + (size_t)stream_id * kJoints * 2, kJoints * 2 * sizeof(FilterState), cudaMemcpyDeviceToHost));
  CU(h, cudaMemcpy(st.data() + kJoints * 2, h->d_st3d + (size_t)stream_id * kJoints * 3, kJoints * 3 * sizeof(FilterState), cudaMemcpyDeviceToHost));
  for (int i = 0; i < kJoints * 5; ++i) {
    double* o = state + 7 * i;
    o[0] = st[i].prev; o[1] = st[i].s_x; o[2] = st[i].s_dx; o[3] = st[i].lasttime; o[4] = st[i].freq;
    o[5] = st[i].has_prev; o[6] = st[i].has_time;
  }
  int4 b;
  CU(h, cudaMemcpy(&b, h->d_boxes + stream_id, sizeof b, cudaMemcpyDeviceToHost));
  double* tail = state + 7 * kJoints * 5;
  tail[0] = h->last_t2d[stream_id]; tail[1] = h->last_t3d[stream_id];
  tail[2] = b.x; tail[3] = b.y; tail[4] = b.z; tail[5] = b.w;
  return VNECT_OK;
}

int vnect_import_stream_state(vnect_t* h, int32_t stream_id, const double* state) {
  if (!h || !h->d_st2d || !state) return fail(h, VNECT_E_INVALID, "handle not created / null buffer");
  ON_DEVICE(h);
  if (stream_id < 0 || stream_id >= h->cfg.max_streams) return fail(h, VNECT_E_INVALID, "stream id out of range");
  std::vector<FilterState> st(kJoints * 5);
  for (int i = 0; i < kJoints * 5; ++i) {
    const double* o = state + 7 * i;
    memset(&st[i], 0, sizeof(FilterState));
    st[i].prev = o[0]; st[i].s_x = o[1]; st[i].s_dx = o[2]; st[i].lasttime = o[3]; st[i].freq = o[4];
    st[i].has_prev = o[5] != 0.0; st[i].has_time = o[6] != 0.0;
  }
  if (int rc = quiesce(h)) return rc;
  CU(h, cudaMemcpy(h->d_st2d + (size_t)stream_id * kJoints * 2, st.data(), kJoints * 2 * sizeof(FilterState), cudaMemcpyHostToDevice));
  CU(h, cudaMemcpy(h->d_st3d + (size_t)stream_id * kJoints * 3, st.data() + kJoints * 2, kJoints * 3 * sizeof(FilterState), cudaMemcpyHostToDevice));
  const double* tail = state + 7 * kJoints * 5;
  h->last_t2d[stream_id] = tail[0];
  h->last_t3d[stream_id] = tail[1];
  const int4 b = make_int4((int)tail[2], (int)tail[3], (int)tail[4], (int)tail[5]);
  CU(h, cudaMemcpy(h->d_boxes + stream_id, &b, sizeof b, cudaMemcpyHostToDevice));
  return VNECT_OK;
}

int vnect_set_stream(vnect_t* h, void* cuda_stream) {
  if (!h) return VNECT_E_INVALID;
  ON_DEVICE(h);
  if (h->own_stream && h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  h->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  h->own_stream = false;
  return VNECT_OK;
}

int vnect_set_packed_results(vnect_t* h, void* dev_packed) {
  if (!h) return VNECT_E_INVALID;
  ON_DEVICE(h);
  CU(h, cudaStreamSynchronize(h->stream));
  for (auto& kv : h->graphs)  // the pointer is baked into the captured post-process launch
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  h->graphs.clear();
  h->d_packed = reinterpret_cast<double*>(dev_packed);
  return VNECT_OK;
}

int vnect_synchronize(vnect_t* h) {
  if (!h) return VNECT_E_INVALID;
  ON_DEVICE(h);
  CU(h, cudaStreamSynchronize(h->stream));
  return VNECT_OK;
}

int vnect_get_tap(vnect_t* h, const char* name, int32_t n, float* out_nhwc, int64_t capacity_elems, int32_t* dims4) {
  if (!h || !h->finalized || !name) return fail(h, VNECT_E_INVALID, "handle not finalized");
  ON_DEVICE(h);
  auto it = h->acts.find(name);
  if (it == h->acts.end()) return fail(h, VNECT_E_INVALID, "no activation named '%s'", name);
  const Act& a = it->second;
  if (n < 1 || n > h->cap_fw) return fail(h, VNECT_E_INVALID, "n out of range");
  if (dims4) { dims4[0] = n; dims4[1] = a.H; dims4[2] = a.W; dims4[3] = a.C; }
  const size_t elems = (size_t)n * a.H * a.W * a.C;
  if (!out_nhwc) return VNECT_OK;
  if ((int64_t)elems > capacity_elems) return fail(h, VNECT_E_INVALID, "buffer too small for tap '%s'", name);
  std::vector<__half> tmp((size_t)n * a.img_px * a.C);
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, cudaMemcpy(tmp.data(), a.p, tmp.size() * sizeof(__half), cudaMemcpyDeviceToHost));
  size_t o = 0;
  for (int i = 0; i < n; ++i)
    for (int y = 0; y < a.H; ++y)
      for (int x = 0; x < a.W; ++x) {
        const __half* src = tmp.data() + ((size_t)i * a.img_px + (size_t)y * a.row_px + x) * a.C;
        for (int c = 0; c < a.C; ++c) out_nhwc[o++] = __half2float(src[c]);
      }
  return VNECT_OK;
}

// CRC32C (Castagnoli) of a host buffer, slicing-by-4: the checksum TensorFlow checkpoints carry per tensor and per
// index block (vnect_b200/tf_checkpoint.py verifies the 58 MB of weights with it).
uint32_t vnect_crc32c(const void* data, uint64_t n) {
  static uint32_t tab[4][256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 4; ++t) tab[t][i] = (tab[t - 1][i] >> 8) ^ tab[0][tab[t - 1][i] & 0xFF];
    init = true;
  }
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 4) {
    c ^= (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    c = tab[3][c & 0xFF] ^ tab[2][(c >> 8) & 0xFF] ^ tab[1][(c >> 16) & 0xFF] ^ tab[0][c >> 24];
    p += 4;
    n -= 4;
  }
  while (n--) c = tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

int vnect_check_finite(vnect_t* h, int32_t n, int64_t* counts) {
  if (!h || !h->finalized || !counts) return fail(h, VNECT_E_INVALID, "handle not finalized / null buffer");
  ON_DEVICE(h);
  if (n < 1 || n > h->cap_fw) return fail(h, VNECT_E_INVALID, "n out of range");
  CU(h, cudaStreamSynchronize(h->stream));
  int i = 0;
  for (const Step& st : h->steps) {
    // the activation a step writes carries the step's (first) scope name, fused steps are "a+b" / "a>b"
    std::string name = st.name.substr(0, st.name.find_first_of("+>"));
    if (name == "conv1") name = "pool1";
    auto it = h->acts.find(name);
    if (it == h->acts.end()) {  // block tails write the block's activation (scope without the branch suffix)
      it = h->acts.find(name.substr(0, name.find("_branch")));
    }
    unsigned long long c = 0;
    if (it != h->acts.end()) {
      const Act& a = it->second;
      CU(h, cudaMemsetAsync(h->d_scan, 0, sizeof(unsigned long long), h->stream));
      count_nonfinite_kernel<<<h->num_sms * 4, 256, 0, h->stream>>>(a.p, (size_t)n * a.img_px * a.C, h->d_scan);
      CU(h, cudaGetLastError());
      CU(h, cudaMemcpyAsync(&c, h->d_scan, sizeof c, cudaMemcpyDeviceToHost, h->stream));
      CU(h, cudaStreamSynchronize(h->stream));
    }
    counts[i++] = (int64_t)c;
  }
  return VNECT_OK;
}

int64_t vnect_launch_count(vnect_t* h) { return h ? h->launches : 0; }

const char* vnect_step_name(vnect_t* h, int32_t i) {
  if (!h || i < 0 || i >= (int)h->steps.size()) return nullptr;
  return h->steps[i].name.c_str();
}

double vnect_info(vnect_t* h, const char* key) {
  if (!h || !key) return NAN;
  const std::string k = key;
  if (k == "num_sms") return h->num_sms;
  if (k == "conv_launches_per_forward") return h->conv_steps;
  if (k == "launches_per_forward") return (double)h->steps.size();
  if (k == "gemm_flops_per_forward") {  // executed GEMM work incl. channel / tile padding, per image
    double f = 0;
    for (const Step& st : h->steps)
      if (st.kind == 0) f += st.launch.flops / h->cap_fw;
      else f += 2.0 * (h->S / 2) * (h->S / 2) * 64 * 224;  // conv1 with its K padded to 7 rows x 8 px x 4 ch
    return f;
  }
  if (k == "hm_size") return h->hs;
  if (k == "max_forwards") return h->cap_fw;
  return NAN;
}

int vnect_time_forward(vnect_t* h, int32_t n, int32_t reps, float* total_ms, float* per_layer_ms) {
  if (!h || !h->finalized) return fail(h, VNECT_E_INVALID, "handle not finalized");
  ON_DEVICE(h);
  if (n < 1 || n > h->cap_fw || reps < 1) return fail(h, VNECT_E_INVALID, "bad n/reps");
  const int ns = (int)h->steps.size();
  std::vector<cudaEvent_t> ev(ns + 1);
  for (auto& e : ev) CU(h, cudaEventCreate(&e));
  cudaEvent_t e0, e1;
  CU(h, cudaEventCreate(&e0));
  CU(h, cudaEventCreate(&e1));
  int rc = run_forward(h, n);  // warm-up
  if (rc) return rc;
  CU(h, cudaEventRecord(e0, h->stream));
  for (int r = 0; r < reps; ++r)
    if ((rc = run_forward(h, n))) return rc;
  CU(h, cudaEventRecord(e1, h->stream));
  CU(h, cudaEventSynchronize(e1));
  float ms = 0;
  CU(h, cudaEventElapsedTime(&ms, e0, e1));
  if (total_ms) *total_ms = ms / reps;
  if (per_layer_ms) {
    std::vector<double> acc(ns, 0.0);
    for (int r = 0; r < reps; ++r) {
      if ((rc = run_forward(h, n, ev.data()))) return rc;
      CU(h, cudaStreamSynchronize(h->stream));
      for (int i = 0; i < ns; ++i) {
        float t = 0;
        CU(h, cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
        acc[i] += t;
      }
    }
    for (int i = 0; i < ns; ++i) per_layer_ms[i] = (float)(acc[i] / reps);
  }
  for (auto& e : ev) cudaEventDestroy(e);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return VNECT_OK;
}

int vnect_time_prepost(vnect_t* h, int32_t n_frames, int32_t reps, float* pre_ms, float* post_ms) {
  if (!h || !h->finalized) return fail(h, VNECT_E_INVALID, "handle not finalized");
  ON_DEVICE(h);
  if (n_frames < 1 || n_frames > h->cfg.max_frames || reps < 1) return fail(h, VNECT_E_INVALID, "bad n_frames/reps");
  int rc = lane_acquire(h, 0);
  if (rc) return rc;
  cudaEvent_t e0, e1;
  CU(h, cudaEventCreate(&e0));
  CU(h, cudaEventCreate(&e1));
  const int S = h->S;
  const Geometry g = squarify_geometry(S, S, S);
  double pre = 0, post = 0;
  std::vector<int32_t> ids(n_frames);
  std::vector<double> t2(n_frames), t3(n_frames);
  for (int i = 0; i < n_frames; ++i) ids[i] = i % h->cfg.max_streams;
  for (int r = -1; r < reps; ++r) {  // r = -1: warm-up
    // fresh, strictly increasing timestamps every repetition (the filters divide by the time step)
    double base = 0;
    for (int i = 0; i < n_frames; ++i) {
      const double last = h->last_t2d[ids[i]] == h->last_t2d[ids[i]] ? h->last_t2d[ids[i]] : 1000.0;
      base = std::max(base, last);
    }
    for (int i = 0; i < n_frames; ++i) { t2[i] = base + 0.033; t3[i] = base + 0.037; }
    for (int i = 0; i < n_frames; ++i) h->last_t3d[ids[i]] = NAN;  // 3D clock readings may trail the 2D ones
    if ((rc = stage_frame_meta(h, n_frames, ids.data(), t2.data(), t3.data(), h->stream))) return rc;
    float ms = 0;
    CU(h, cudaEventRecord(e0, h->stream));
    if ((rc = run_preprocess(h, h->cur->d_frames, n_frames, S, S, (int64_t)S * 3, (int64_t)S * S * 3, g))) return rc;
    CU(h, cudaEventRecord(e1, h->stream));
    CU(h, cudaEventSynchronize(e1));
    CU(h, cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 0) pre += ms;
    CU(h, cudaEventRecord(e0, h->stream));
    if ((rc = run_postprocess(h, n_frames, g.scaler, g.off_x, g.off_y, h->cur->d_out2d, h->cur->d_out3d, false, false))) return rc;
    CU(h, cudaEventRecord(e1, h->stream));
    CU(h, cudaEventSynchronize(e1));
    CU(h, cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 0) post += ms;
  }
  if (h->d_post_trace) {  // developer diagnostics: where a (frame, joint) block of the last launch spent its time
    const int nb = n_frames * kJoints;
    std::vector<unsigned long long> tr((size_t)nb * 16);
    CU(h, cudaMemcpy(tr.data(), h->d_post_trace, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull, t1 = 0;
    for (int b = 0; b < nb; ++b) { t0 = std::min(t0, tr[(size_t)b * 16]); for (int k = 0; k <= 10; ++k) t1 = std::max(t1, tr[(size_t)b * 16 + k]); }
    static const char* names[] = {"start", "pdl_wait", "staged", "sum pass", "reduced", "marked", "exact+reduce", "2D filter", "gather", "fence+atomic", "tail"};
    fprintf(stderr, "post-process trace (%d blocks, kernel span %.2f us): phase end relative to the block's start, mean / max us; block start spread %.2f us\n",
            nb, (t1 - t0) * 1e-3, 0.0);
    double start_max = 0;
    for (int b = 0; b < nb; ++b) start_max = std::max(start_max, (tr[(size_t)b * 16] - t0) * 1e-3);
    fprintf(stderr, "  last block started %.2f us after the first\n", start_max);
    for (int b = 0; b < nb; ++b)
      if (tr[(size_t)b * 16 + 10] > tr[(size_t)b * 16]) {  // a block that ran the tail: SM cycles per ns over its lifetime
        fprintf(stderr, "  SM clock during the kernel: %.0f MHz\n",
                1e3 * (double)(tr[(size_t)b * 16 + 12] - tr[(size_t)b * 16 + 11]) / (double)(tr[(size_t)b * 16 + 10] - tr[(size_t)b * 16]));
        break;
      }
    {  // survivors per block and the spread of the blocks' own durations (start -> fence + counter)
      std::vector<double> dur;
      std::vector<unsigned long long> nqs;
      for (int b = 0; b < nb; ++b) {
        dur.push_back((tr[(size_t)b * 16 + 9] - tr[(size_t)b * 16]) * 1e-3);
        nqs.push_back(tr[(size_t)b * 16 + 13]);
      }
      std::sort(dur.begin(), dur.end());
      std::sort(nqs.begin(), nqs.end());
      fprintf(stderr, "  block duration p50 %.2f p90 %.2f p99 %.2f max %.2f us; surviving quads per block p50 %llu p90 %llu p99 %llu max %llu\n",
              dur[nb / 2], dur[nb * 9 / 10], dur[nb * 99 / 100], dur[nb - 1], nqs[nb / 2], nqs[nb * 9 / 10], nqs[nb * 99 / 100], nqs[nb - 1]);
    }
    {  // inside "exact+reduce": the exact cells are in shared memory (blocks on the few-quads path only)
      double sum = 0; int cnt = 0;
      for (int b = 0; b < nb; ++b) {
        const unsigned long long v = tr[(size_t)b * 16 + 14], s0 = tr[(size_t)b * 16 + 5];
        if (v >= s0 && v <= t1) { sum += (v - s0) * 1e-3; ++cnt; }
      }
      fprintf(stderr, "  exact cells ready %.2f us after the marking barrier (%d blocks)\n", cnt ? sum / cnt : 0.0, cnt);
    }
    for (int k = 1; k <= 10; ++k) {
      double sum = 0, mx = 0; int cnt = 0;
      for (int b = 0; b < nb; ++b) {
        const unsigned long long v = tr[(size_t)b * 16 + k], s0 = tr[(size_t)b * 16];
        if (v < s0 || v > t1) continue;  // stale slot (tail runs in one block per frame)
        if (k == 10 && v < tr[(size_t)b * 16 + 9]) continue;
        sum += (v - s0) * 1e-3; mx = std::max(mx, (v - s0) * 1e-3); ++cnt;
      }
      fprintf(stderr, "  %-14s %7.2f / %7.2f us  (%d blocks)\n", names[k], cnt ? sum / cnt : 0.0, mx, cnt);
    }
  }
  if (h->d_post_trace) {  // developer diagnostics: where a (frame, joint) block of the last launch spent its time
    const int nb = n_frames * kJoints;
    std::vector<unsigned long long> tr((size_t)nb * 16);
    CU(h, cudaMemcpy(tr.data(), h->d_post_trace, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull, t1 = 0;
    for (int b = 0; b < nb; ++b) { t0 = std::min(t0, tr[(size_t)b * 16]); for (int k = 0; k <= 10; ++k) t1 = std::max(t1, tr[(size_t)b * 16 + k]); }
    static const char* names[] = {"start", "pdl_wait", "staged", "sum pass", "reduced", "marked", "exact+reduce", "2D filter", "gather", "fence+atomic", "tail"};
    fprintf(stderr, "post-process trace (%d blocks, kernel span %.2f us): phase end relative to the block's start, mean / max us; block start spread %.2f us\n",
            nb, (t1 - t0) * 1e-3, 0.0);
    double start_max = 0;
    for (int b = 0; b < nb; ++b) start_max = std::max(start_max, (tr[(size_t)b * 16] - t0) * 1e-3);
    fprintf(stderr, "  last block started %.2f us after the first\n", start_max);
    for (int b = 0; b < nb; ++b)
      if (tr[(size_t)b * 16 + 10] > tr[(size_t)b * 16]) {  // a block that ran the tail: SM cycles per ns over its lifetime
        fprintf(stderr, "  SM clock during the kernel: %.0f MHz\n",
                1e3 * (double)(tr[(size_t)b * 16 + 12] - tr[(size_t)b * 16 + 11]) / (double)(tr[(size_t)b * 16 + 10] - tr[(size_t)b * 16]));
        break;
      }
    {  // survivors per block and the spread of the blocks' own durations (start -> fence + counter)
      std::vector<double> dur;
      std::vector<unsigned long long> nqs;
      for (int b = 0; b < nb; ++b) {
        dur.push_back((tr[(size_t)b * 16 + 9] - tr[(size_t)b * 16]) * 1e-3);
        nqs.push_back(tr[(size_t)b * 16 + 13]);
      }
      std::sort(dur.begin(), dur.end());
      std::sort(nqs.begin(), nqs.end());
      fprintf(stderr, "  block duration p50 %.2f p90 %.2f p99 %.2f max %.2f us; surviving quads per block p50 %llu p90 %llu p99 %llu max %llu\n",
              dur[nb / 2], dur[nb * 9 / 10], dur[nb * 99 / 100], dur[nb - 1], nqs[nb / 2], nqs[nb * 9 / 10], nqs[nb * 99 / 100], nqs[nb - 1]);
    }
    {  // inside "exact+reduce": the exact cells are in shared memory (blocks on the few-quads path only)
      double sum = 0; int cnt = 0;
      for (int b = 0; b < nb; ++b) {
        const unsigned long long v = tr[(size_t)b * 16 + 14], s0 = tr[(size_t)b * 16 + 5];
        if (v >= s0 && v <= t1) { sum += (v - s0) * 1e-3; ++cnt; }
      }
      fprintf(stderr, "  exact cells ready %.2f us after the marking barrier (%d blocks)\n", cnt ? sum / cnt : 0.0, cnt);
    }
    for (int k = 1; k <= 10; ++k) {
      double sum = 0, mx = 0; int cnt = 0;
      for (int b = 0; b < nb; ++b) {
        const unsigned long long v = tr[(size_t)b * 16 + k], s0 = tr[(size_t)b * 16];
        if (v < s0 || v > t1) continue;  // stale slot (tail runs in one block per frame)
        if (k == 10 && v < tr[(size_t)b * 16 + 9]) continue;
        sum += (v - s0) * 1e-3; mx = std::max(mx, (v - s0) * 1e-3); ++cnt;
      }
      fprintf(stderr, "  %-14s %7.2f / %7.2f us  (%d blocks)\n", names[k], cnt ? sum / cnt : 0.0, mx, cnt);
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (pre_ms) *pre_ms = (float)(pre / reps);
  if (post_ms) *post_ms = (float)(post / reps);
  return VNECT_OK;
}

void vnect_destroy(vnect_t* h) {
  if (!h) return;
  ON_DEVICE(h);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& kv : h->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (void* p : h->allocs) cudaFree(p);
  for (auto& L : h->lanes) {
    if (L.h_meta) cudaFreeHost(L.h_meta);
    if (L.h_nonfinite) cudaFreeHost(L.h_nonfinite);
    if (L.copy_done) cudaEventDestroy(L.copy_done);
    if (L.done) cudaEventDestroy(L.done);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

}  // extern "C"
