// conv1 (7x7/2, 3 -> 64, bias, ReLU) fused with pool1 (3x3/2 'SAME' max-pool), reference src/vnect_model.py:27-29,
// as a ROLLING implicit GEMM over input rows.
//
// A 128x64x16 MMA costs ~107 cycles here and ~72 of them are the read of its A operand from smem, so the 64-channel
// stem is bound by how often an A strip is read, not by FLOPs.  One padded input row r feeds up to four conv rows
// (y = (r - ky) / 2 for the row taps ky of r's parity), so instead of one N = 64 MMA per (conv row, ky) this kernel
// issues ONE MMA per input row whose B operand stacks the weights of those taps, [W6; W4; W2; W0] (N = 256) for even
// rows and [W5; W3; W1] (N = 192) for odd rows, and whose 64-column output groups are the accumulators of
// consecutive conv rows: TMEM is an 8-slot ring indexed by conv row.  Per conv row that is ~870 cycles of tensor
// pipe instead of 14 x 107, and nothing is computed twice between neighbouring pooled rows.
//
// A CTA's unit of work: (forward, x tile of 128 conv columns, segment of pooled rows).  Strips are the same raw
// no-swizzle K-major windows as in stem_gemm.cuh (consecutive conv columns are 16 bytes apart), 2112 bytes per input
// row, fetched with one bulk copy each.  Accumulators are only ever accumulated into: the epilogue warps zero a slot
// (tcgen05.st) right after draining it.  Each epilogue thread drains the same conv column of every row (bias, ReLU,
// fp16), keeps the vertical 3-max of the open pooling window in registers, and every second row the 8 epilogue
// warps exchange the column maxima through smem for the horizontal 3-max and write one pooled row straight to HBM.
#pragma once
#include "stem_gemm.cuh"

namespace vnect {

constexpr int kRollThreads = 128 + 256;      // 4 control warps + 8 epilogue / pooling warps
constexpr int kRollStripLoad = kBlockM * 16 + 64;
constexpr int kRollStripBytes = 2176;        // kRollStripLoad rounded up to a multiple of 128
constexpr int kRollStages = 16;
constexpr int kRollRowBufs = 2;         // column-maxima buffers (alternating pooling windows)
constexpr int kRollSlots = 8;                // 8 x 64 fp32 columns = all of TMEM
constexpr int kRollRowBytes = kBlockM * 128;
constexpr int kMaxXTiles = 4;
constexpr int kRollWEvenBytes = 256 * 64;  // [4 taps x 64 couts][32 K] fp16, 64B-swizzled K-major rows
constexpr int kRollWOddBytes = 192 * 64;   // [3 taps x 64 couts][32 K]
static_assert(kRollWEvenBytes + kRollWOddBytes == kStemWBytes, "stacked weights are a permutation of the canonical pack");

struct StemRollParams {
  const uint8_t* x1;
  int64_t plane_bytes;
  const uint8_t* w;     // stacked pack (pack_stem_stacked)
  const float* bias;
  __half* out;          // pooled NHWC [NB][PH][PW][64]
  int vw;               // virtual columns per conv row (S/2 + 3)
  int CH, CW, PH, PW;
  int n_xt;             // x tiles; tile i covers conv columns [xt_x0[i], +128) and writes pooled columns [xt_pb[i], xt_pe[i])
  int xt_x0[kMaxXTiles], xt_pb[kMaxXTiles], xt_pe[kMaxXTiles];
  int seg_rows;         // pooled rows per segment
  int segs_per_image;
  int num_items;        // forwards * segs_per_image * n_xt
  int reverse;          // 1: walk the items from last to first (see ConvGemmParams::reverse)
  unsigned long long* dbg;  // optional [4] cycle counters of CTA 0 (selftest only): wait-for-MMA, drain, pool, total
};

struct StemRollSmem {
  static constexpr int W_OFF = 0;
  static constexpr int STRIP_OFF = kStemWBytes;
  static constexpr int ROW_OFF = ((STRIP_OFF + kRollStages * kRollStripBytes + 1023) / 1024) * 1024;
  static constexpr int BAR_OFF = ROW_OFF + kRollRowBufs * kRollRowBytes;
  static constexpr int BYTES = BAR_OFF + 1024 + 1024;
};

struct RollItem {
  int img, xt, p0, n_pool, ya, n_rows, n_in;
};

__device__ __forceinline__ RollItem roll_decode(const StemRollParams& p, int it) {
  const int item = p.reverse ? p.num_items - 1 - it : it;
  const int per_img = p.n_xt * p.segs_per_image;
  RollItem r;
  r.img = item / per_img;
  const int rem = item - r.img * per_img;
  const int seg = rem / p.n_xt;
  r.xt = rem - seg * p.n_xt;
  r.p0 = seg * p.seg_rows;
  const int p1 = min(p.PH, r.p0 + p.seg_rows);
  r.n_pool = p1 - r.p0;
  r.ya = 2 * r.p0;
  const int yb = min(2 * p1 + 1, p.CH);  // pooled row p uses conv rows 2p .. 2p+2, clipped at the bottom edge
  r.n_rows = yb - r.ya;
  r.n_in = 2 * (r.n_rows - 1) + 7;       // padded input rows 2*ya .. 2*(yb-1)+6
  return r;
}

__device__ __forceinline__ void tmem_zero_32x32(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kRollThreads, 1) stem_roll_kernel(const __grid_constant__ StemRollParams p) {
  constexpr uint32_t TMEM_COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem + StemRollSmem::W_OFF;
  uint8_t* strips = smem + StemRollSmem::STRIP_OFF;
  uint8_t* rows = smem + StemRollSmem::ROW_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + StemRollSmem::BAR_OFF);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kRollStages;
  uint64_t* tmem_full = bars + 2 * kRollStages;
  uint64_t* tmem_empty = tmem_full + kRollSlots;
  uint64_t* w_bar = tmem_empty + kRollSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* bias_s = reinterpret_cast<float*>(bars + 64);  // 512 B into the 1 KB barrier block

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kRollStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kRollSlots; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);  // one arrive per epilogue warp
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  if (threadIdx.x >= 128 && threadIdx.x < 192) bias_s[threadIdx.x - 128] = p.bias[threadIdx.x - 128];
  pdl_launch_dependents();
  pdl_wait();  // bias (read above) is a constant; the input strips come from the previous kernel
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ strip loader: one bulk copy per padded input row
    const bool issuer = elect_one();
    if (issuer) {
      mbar_arrive_expect_tx(w_bar, kStemWBytes);
      bulk_load_1d(w_smem, p.w, kStemWBytes, w_bar);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const RollItem R = roll_decode(p, it);
      const uint8_t* base = p.x1 + static_cast<int64_t>(R.img) * 2 * p.plane_bytes + 16ll * p.xt_x0[R.xt];
      for (int ri = 0; ri < R.n_in; ++ri) {
        const int r = 2 * R.ya + ri;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (issuer) {
          mbar_arrive_expect_tx(&full_bar[stage], kRollStripLoad);
          bulk_load_1d(strips + stage * kRollStripBytes, base + (r & 1) * p.plane_bytes + 16ll * (r >> 1) * p.vw,
                       kRollStripLoad, &full_bar[stage]);
        }
        __syncwarp();
        if (++stage == kRollStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (warp-converged, one lane issues)
    const bool issuer = elect_one();
    mbar_wait(w_bar, 0);
    const uint32_t w_even = smem_u32(w_smem), w_odd = w_even + kRollWEvenBytes;
    int stage = 0;
    uint32_t phase = 0;
    int G0 = 0;  // conv rows this CTA has started before the current item: row G lives in TMEM slot G & 7
    long long m_full = 0, m_empty = 0;
    const long long m_begin = clock64();
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const RollItem R = roll_decode(p, it);
      for (int ri = 0; ri < R.n_in; ++ri) {
        const long long m0 = clock64();
        mbar_wait(&full_bar[stage], phase);
        m_full += clock64() - m0;
        tc_fence_after();
        const bool even = !(ri & 1);
        const int ly_hi = min(R.n_rows - 1, ri >> 1);
        const int ly_lo = ri >= 6 ? (ri - 5) >> 1 : 0;  // ceil((ri - 6) / 2)
        if (even && (ri >> 1) < R.n_rows) {  // first tap (ky = 0) of conv row ri/2: its slot must have been zeroed
          const int G = G0 + (ri >> 1);
          const long long m1 = clock64();
          mbar_wait(&tmem_empty[G & 7], (G >> 3) & 1);
          m_empty += clock64() - m1;
          tc_fence_after();
        }
        const uint64_t a_desc = make_noswz_desc(smem_u32(strips + stage * kRollStripBytes), 16, 128);
        for (int ly = ly_lo; ly <= ly_hi;) {
          const int slot = (G0 + ly) & 7;
          const int cnt = min(ly_hi - ly + 1, kRollSlots - slot);  // a run may not wrap around the ring
          const int ky = ri - 2 * ly;                              // tap of the run's first row; later rows: ky - 2, ...
          const int g = even ? (6 - ky) >> 1 : (5 - ky) >> 1;      // its 64-row group in the stacked weights
          // B rows are 64 B (one tap row = 32 K) in the 64B-swizzled layout: a no-swizzle B costs ~1.5 cycles per
          // row and MMA, the swizzled one ~0.5
          const uint64_t b_desc = make_kmajor_desc<64>((even ? w_even : w_odd) + g * 4096);
          const uint32_t idesc = make_idesc_f16(kBlockM, 64 * cnt, false);
          if (issuer) {
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(slot * 64);
            umma_f16(d_tmem, a_desc, b_desc, idesc, 1u);
            umma_f16(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);  // +32 B of A, +32 B along K of B
          }
          __syncwarp();
          ly += cnt;
        }
        if (issuer) {
          if (even && ri >= 6) umma_commit(&tmem_full[(G0 + ((ri - 6) >> 1)) & 7]);  // last tap (ky = 6) of that row
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kRollStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      G0 += R.n_rows;
    }
    if (p.dbg != nullptr && lane == 0 && blockIdx.x == 0) {
      p.dbg[4] = m_full; p.dbg[5] = m_empty; p.dbg[6] = clock64() - m_begin;
    }
  } else if (warp >= 4) {
    // ================================================================ 8 warps: drain + zero a slot, pool every 2nd row
    // A thread drains the same conv column of every row, so the vertical 3-max of a pooling window stays in
    // registers (the window's first two rows are kept packed); only the column maxima go through smem, once per
    // window, for the horizontal 3-max.  (Staging all three rows and reading 3x3 windows back cost 105 KB of smem
    // traffic per window on top of the MMAs' operand reads and made the epilogue the bottleneck.)
    const int ew = warp - 4;          // 0..7
    const int q4 = warp & 3;          // TMEM lane quarter this warp may access
    const int chalf = ew >> 2;        // which 32 of the 64 output channels
    const int r = q4 * 32 + lane;     // conv column (within the x tile) owned by this thread
    const int et = threadIdx.x - 128; // 0..255
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(chalf * 32);
    float bias_r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) bias_r[i] = bias_s[chalf * 32 + i];
    for (int s = 0; s < kRollSlots; ++s) tmem_zero_32x32(t_lane + s * 64);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int s = 0; s < kRollSlots; ++s) mbar_arrive(&tmem_empty[s]);
    int G0 = 0;
    uint32_t win = 0;  // pooling windows done by this CTA: column maxima alternate between two smem buffers
    long long t_wait = 0, t_drain = 0, t_pool = 0;
    const long long t_begin = clock64();
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const RollItem R = roll_decode(p, it);
      const int x0 = p.xt_x0[R.xt], pb = p.xt_pb[R.xt], pe = p.xt_pe[R.xt];
      __half2 row_a[16], row_b[16];  // the open window's first (even) and second (odd) conv row, this thread's column
      for (int ly = 0; ly < R.n_rows; ++ly) {
        const int G = G0 + ly;
        const long long c0 = clock64();
        mbar_wait(&tmem_full[G & 7], (G >> 3) & 1);
        const long long c1 = clock64();
        t_wait += c1 - c0;
        tc_fence_after();
        uint32_t v[32];
        const uint32_t taddr = t_lane + static_cast<uint32_t>((G & 7) * 64);
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        tmem_zero_32x32(taddr);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[G & 7]);
        __half2 cur[16];  // bias + ReLU + fp16 pack (cvt.rn.relu does the max(.,0) while packing)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t pk = pack_half2_relu(__uint_as_float(v[2 * j]) + bias_r[2 * j], __uint_as_float(v[2 * j + 1]) + bias_r[2 * j + 1]);
          cur[j] = *reinterpret_cast<const __half2*>(&pk);
        }
        const long long c2 = clock64();
        t_drain += c2 - c1;
        // a pooling window closes on every even row >= 2, and on the last row when the image edge clips it to 2 rows
        const bool full3 = ly >= 2 && !(ly & 1);
        const bool clipped = (ly == R.n_rows - 1) && (ly & 1);
        if (full3 || clipped) {
          const int lpy = full3 ? (ly >> 1) - 1 : (ly >> 1);
          uint8_t* vbuf = rows + (win & 1u) * kRollRowBytes;
          ++win;
          uint8_t* rowp = vbuf + static_cast<uint32_t>(r) * 128u;
          const uint32_t sw = static_cast<uint32_t>(r) & 7u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __half2 t = __hmax2(row_a[4 * j + e], cur[4 * j + e]);
              if (full3) t = __hmax2(t, row_b[4 * j + e]);
              oh[e] = t;
            }
            const uint32_t chunk = static_cast<uint32_t>(chalf * 4 + j);
            *reinterpret_cast<uint4*>(rowp + ((chunk ^ sw) << 4)) = o;
          }
          named_bar_sync(1, 256);  // the window's column maxima are in smem (the other buffer is still being read)
          const int units = (pe - pb) * 8;
          for (int u = et; u < units; u += 256) {
            const uint32_t pc = static_cast<uint32_t>(u & 7);
            const int px = pb + (u >> 3);
            const uint32_t c0 = static_cast<uint32_t>(2 * px - x0);
            const int nc = (2 * px + 2 < p.CW) ? 3 : 2;
            uint4 acc = *reinterpret_cast<const uint4*>(vbuf + c0 * 128u + ((pc ^ (c0 & 7u)) << 4));
            __half2* m = reinterpret_cast<__half2*>(&acc);
#pragma unroll
            for (int b = 1; b < 3; ++b) {
              if (b < nc) {
                const uint32_t col = c0 + b;
                const uint4 val = *reinterpret_cast<const uint4*>(vbuf + col * 128u + ((pc ^ (col & 7u)) << 4));
                const __half2* hv = reinterpret_cast<const __half2*>(&val);
                m[0] = __hmax2(m[0], hv[0]); m[1] = __hmax2(m[1], hv[1]);
                m[2] = __hmax2(m[2], hv[2]); m[3] = __hmax2(m[3], hv[3]);
              }
            }
            __half* o = p.out + ((static_cast<size_t>(R.img) * p.PH + (R.p0 + lpy)) * p.PW + px) * 64 + pc * 8;
            *reinterpret_cast<uint4*>(o) = acc;
          }
          t_pool += clock64() - c2;
        }
        if (ly & 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) row_b[j] = cur[j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) row_a[j] = cur[j];
        }
      }
      G0 += R.n_rows;
    }
    if (p.dbg != nullptr && et == 0 && blockIdx.x == 0) {
      p.dbg[0] = t_wait; p.dbg[1] = t_drain; p.dbg[2] = t_pool; p.dbg[3] = clock64() - t_begin;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace vnect
