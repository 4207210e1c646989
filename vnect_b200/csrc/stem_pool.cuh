// conv1 (7x7/2, 3 -> 64, bias, ReLU) FUSED with pool1 (3x3/2 'SAME' max-pool): reference src/vnect_model.py:27-29.
//
// Same raw-strip implicit GEMM as stem_gemm.cuh (A operand read in place through overlapping no-swizzle UMMA
// descriptors, weights resident in smem), but a CTA's unit of work is a BAND: conv rows 4q .. 4q+4 of one image
// = 5 x VW virtual pixels = 8 MMA tiles of 128 rows, whose fp32 accumulators occupy all 512 TMEM columns as an
// 8-slot ring.  The epilogue warps drain each tile (bias, ReLU, fp16) into a 128 KB smem band, then compute the two
// pooled rows 2q, 2q+1 from it and write them to HBM fully coalesced.  The 184 x 184 x 64 conv1 tensor never exists
// in memory: HBM traffic is the input strips (x1.25 for the shared band row) plus the 92 x 92 x 64 pooled output.
#pragma once
#include "stem_gemm.cuh"

namespace vnect {

constexpr int kBandTiles = 8;                       // 8 x 128 virtual pixels >= 5 conv rows of up to 204 columns
constexpr int kBandRows = 5;                        // conv rows per band (two pooled rows)
constexpr int kStemPoolThreads = 128 + 256;         // 4 control warps + 8 epilogue/pool warps
constexpr int kBandBytes = kBandTiles * kBlockM * 128;
// One strip per row tap covers ALL tiles of a band (they are consecutive virtual pixels): 7 bulk copies per band
// instead of 56 -- the single producer thread's issue rate (~200 cycles per copy) was the bottleneck with per-tile
// strips (measured: 360 us for 128 images).
constexpr int kBandStripBytes = kBandTiles * kBlockM * 16 + 64;
constexpr int kBandStages = 3;

struct StemPoolParams {
  const uint8_t* x1;
  int64_t plane_bytes;
  const uint8_t* w;
  const float* bias;
  __half* out;          // pooled NHWC [NB][PH][PW][64]
  int vw;               // virtual columns per conv row (S/2 + 3)
  int CH, CW;           // conv grid (S/2)
  int PH, PW;           // pooled grid (S/4)
  int ppb;              // pooled rows per band: 2 (5 conv rows) when 5*vw <= 1024, else 1 (3 conv rows)
  int band_tiles;       // ceil((2*ppb+1) * vw / 128) <= 8
  int bands_per_image;  // PH / ppb
  int num_items;        // images * bands_per_image
  int reverse;          // 1: walk the bands from last to first (see ConvGemmParams::reverse)
  unsigned long long* dbg;  // optional [4] cycle counters (selftest only): wait-for-MMA, drain, barrier, pool
};

struct StemPoolSmem {
  static constexpr int W_OFF = 0;
  static constexpr int STRIP_OFF = kStemWBytes;
  static constexpr int BAND_OFF = ((STRIP_OFF + kBandStages * kBandStripBytes + 1023) / 1024) * 1024;
  static constexpr int BAR_OFF = BAND_OFF + kBandBytes;
  static constexpr int BYTES = BAR_OFF + 1024 + 1024;
};

__global__ void __launch_bounds__(kStemPoolThreads, 1) stem_pool_kernel(const __grid_constant__ StemPoolParams p) {
  constexpr uint32_t IDESC = make_idesc_f16(kBlockM, 64, false);
  constexpr uint32_t TMEM_COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem + StemPoolSmem::W_OFF;
  uint8_t* strips = smem + StemPoolSmem::STRIP_OFF;
  uint8_t* band = smem + StemPoolSmem::BAND_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + StemPoolSmem::BAR_OFF);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kBandStages;
  uint64_t* tmem_full = bars + 2 * kBandStages;
  uint64_t* tmem_empty = tmem_full + kBandTiles;
  uint64_t* w_bar = tmem_empty + kBandTiles;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* bias_s = reinterpret_cast<float*>(w_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kBandStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kBandTiles; ++s) {
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
    // ================================================================ strip loader
    const bool issuer = elect_one();
    {
      if (issuer) {
        mbar_arrive_expect_tx(w_bar, kStemWBytes);
        bulk_load_1d(w_smem, p.w, kStemWBytes, w_bar);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
        const int item = p.reverse ? p.num_items - 1 - it : it;
        const int img = item / p.bands_per_image;
        const int q = item - img * p.bands_per_image;
        const uint8_t* img_base = p.x1 + static_cast<int64_t>(img) * 2 * p.plane_bytes;
        const int v_base = 2 * p.ppb * q * p.vw;
        const uint32_t bytes = static_cast<uint32_t>(p.band_tiles * kBlockM * 16 + 64);
        for (int ky = 0; ky < 7; ++ky) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (issuer) {
            mbar_arrive_expect_tx(&full_bar[stage], bytes);
            const uint8_t* src = img_base + (ky & 1) * p.plane_bytes + 16ll * (v_base + p.vw * (ky >> 1));
            bulk_load_1d(strips + stage * kBandStripBytes, src, bytes, &full_bar[stage]);
          }
          __syncwarp();
          if (++stage == kBandStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer: TMEM slot t <-> tile t of the band
    // (warp-converged loop, one elected lane issues: see the note in conv_gemm.cuh)
    const bool issuer = elect_one();
    {
      mbar_wait(w_bar, 0);
      int stage = 0;
      uint32_t phase = 0, item_par = 0;
      const uint64_t b_desc0 = make_noswz_desc(smem_u32(w_smem), 1024, 128);
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, item_par ^= 1) {
        for (int ky = 0; ky < 7; ++ky) {  // row tap outer, tile inner: all accumulators of the band are live in TMEM
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc0 = make_noswz_desc(smem_u32(strips + stage * kBandStripBytes), 16, 128);
          const uint64_t b_desc = b_desc0 + static_cast<uint64_t>(ky * 256);  // (ky*4 chunks * 1024 B) >> 4
          for (int t = 0; t < p.band_tiles; ++t) {
            if (ky == 0) {
              mbar_wait(&tmem_empty[t], item_par ^ 1);  // previous band's tile t has been drained
              tc_fence_after();
            }
            if (issuer) {
              const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(t * 64);
              const uint64_t a_desc = a_desc0 + static_cast<uint64_t>(t * 128);  // (t * 2048 B) >> 4
              umma_f16(d_tmem, a_desc, b_desc, IDESC, ky != 0 ? 1u : 0u);
              umma_f16(d_tmem, a_desc + 2, b_desc + 128, IDESC, 1u);  // +32 B of A, +2 K-chunks (2048 B) of B
              if (ky == 6) umma_commit(&tmem_full[t]);
            }
            __syncwarp();
          }
          if (issuer) umma_commit(&empty_bar[stage]);
          __syncwarp();
          if (++stage == kBandStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================================ 8 warps: drain tiles to the smem band, then pool
    const int ew = warp - 4;          // 0..7
    const int q4 = warp & 3;          // TMEM lane quarter this warp may read
    const int chalf = ew >> 2;        // which 32 of the 64 output channels
    const int r = q4 * 32 + lane;     // row within a tile
    const int et = threadIdx.x - 128; // 0..255
    float bias_r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) bias_r[i] = bias_s[chalf * 32 + i];
    // pooling role of this thread: one 16-byte channel chunk, a run of consecutive pooled columns
    const int pc = et & 7;
    const int run = (p.PW + 31) / 32;
    const int px_begin = (et >> 3) * run;
    const int px_end = min(px_begin + run, p.PW);
    uint32_t item_par = 0;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x, item_par ^= 1) {
      const int item = p.reverse ? p.num_items - 1 - it : it;
      const int img = item / p.bands_per_image;
      const int q = item - img * p.bands_per_image;
      long long c0 = clock64(), c1 = 0;
      for (int t = 0; t < p.band_tiles; ++t) {
        mbar_wait(&tmem_full[t], item_par);
        if (t == 0) c1 = clock64();
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(t * 64 + chalf * 32), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[t]);
        const uint32_t vrow = static_cast<uint32_t>(t * kBlockM + r);
        uint8_t* rowp = band + vrow * 128u;
        const uint32_t sw = vrow & 7u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;  // bias + ReLU + fp16 pack (cvt.rn.relu does the max(.,0) while packing)
          o.x = pack_half2_relu(__uint_as_float(v[8 * j + 0]) + bias_r[8 * j + 0], __uint_as_float(v[8 * j + 1]) + bias_r[8 * j + 1]);
          o.y = pack_half2_relu(__uint_as_float(v[8 * j + 2]) + bias_r[8 * j + 2], __uint_as_float(v[8 * j + 3]) + bias_r[8 * j + 3]);
          o.z = pack_half2_relu(__uint_as_float(v[8 * j + 4]) + bias_r[8 * j + 4], __uint_as_float(v[8 * j + 5]) + bias_r[8 * j + 5]);
          o.w = pack_half2_relu(__uint_as_float(v[8 * j + 6]) + bias_r[8 * j + 6], __uint_as_float(v[8 * j + 7]) + bias_r[8 * j + 7]);
          const uint32_t chunk = static_cast<uint32_t>(chalf * 4 + j);
          *reinterpret_cast<uint4*>(rowp + ((chunk ^ sw) << 4)) = o;
        }
      }
      const long long c2 = clock64();
      named_bar_sync(1, 256);  // the whole band is in smem
      const long long c3 = clock64();
      // ---- 3x3/2 max-pool of band rows (TF SAME pads (0,1): windows are clipped at the bottom / right edge).
      // Separable: vertical max of each conv column once, then the horizontal 3-max; the even column shared by two
      // neighbouring windows is carried in registers.
      for (int pr = 0; pr < p.ppb; ++pr) {
        const int py = p.ppb * q + pr;
        const int nrows = min(3, p.CH - 2 * py);
        const uint32_t row0 = static_cast<uint32_t>(2 * pr * p.vw);
        auto vmax = [&](int cx, __half2 (&m)[4]) {
          uint32_t vrow = row0 + static_cast<uint32_t>(cx);
          uint4 val = *reinterpret_cast<const uint4*>(band + vrow * 128u + ((static_cast<uint32_t>(pc) ^ (vrow & 7u)) << 4));
          const __half2* hv = reinterpret_cast<const __half2*>(&val);
          m[0] = hv[0]; m[1] = hv[1]; m[2] = hv[2]; m[3] = hv[3];
          for (int a = 1; a < nrows; ++a) {
            vrow += static_cast<uint32_t>(p.vw);
            val = *reinterpret_cast<const uint4*>(band + vrow * 128u + ((static_cast<uint32_t>(pc) ^ (vrow & 7u)) << 4));
            m[0] = __hmax2(m[0], hv[0]); m[1] = __hmax2(m[1], hv[1]);
            m[2] = __hmax2(m[2], hv[2]); m[3] = __hmax2(m[3], hv[3]);
          }
        };
        if (px_begin < px_end) {
          __half2 carry[4], c1[4], c2[4];
          vmax(2 * px_begin, carry);
          __half* o = p.out + ((static_cast<size_t>(img) * p.PH + py) * p.PW + px_begin) * 64 + pc * 8;
          for (int px = px_begin; px < px_end; ++px, o += 64) {
            vmax(2 * px + 1, c1);  // 2*px+1 <= CW-1 always
            __half2 m[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) m[e] = __hmax2(carry[e], c1[e]);
            if (2 * px + 2 < p.CW) {
              vmax(2 * px + 2, c2);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                m[e] = __hmax2(m[e], c2[e]);
                carry[e] = c2[e];
              }
            }
            *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(m);
          }
        }
      }
      if (p.dbg != nullptr && et == 0 && blockIdx.x == 0) {
        const long long c4 = clock64();
        atomicAdd(&p.dbg[0], (unsigned long long)(c1 - c0));
        atomicAdd(&p.dbg[1], (unsigned long long)(c2 - c1));
        atomicAdd(&p.dbg[2], (unsigned long long)(c3 - c2));
        atomicAdd(&p.dbg[3], (unsigned long long)(c4 - c3));
      }
      named_bar_sync(1, 256);  // band may be overwritten by the next item
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
