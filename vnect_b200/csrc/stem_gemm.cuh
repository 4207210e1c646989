// conv1 (7x7 stride 2, Cin = 3 -> 64, ReLU; reference src/vnect_model.py:27) as an implicit GEMM whose A operand is
// read IN PLACE from raw input strips: no im2col, no replication.
//
// Input ("stem layout", written by pyramid_kernel): [forward][row parity][rpp rows][row pitch] fp16, NHWC padded to
// 4 channels; the row pitch is (S+6)*4 halves = (S/2+3)*16 bytes, so with VW = S/2+3 "virtual" output columns per
// row the 8-pixel x 4-channel window of virtual output pixel v = oy*VW + ox for row tap ky starts at byte
//     plane(ky & 1) + 16 * (v + VW * (ky >> 1))
// i.e. consecutive GEMM rows are exactly 16 bytes apart.  That is the canonical no-swizzle K-major UMMA layout
// (8-row core matrices of 16-byte rows, SBO = 128 B between row groups) with overlapping K chunks (LBO = 16 B), so
// a contiguous strip of 128*16+48 bytes, fetched with ONE bulk copy per row tap, is a valid 128 x 32 A tile.
// The 3 virtual columns per row beyond S/2 produce junk outputs that are stored (the conv1 buffer has the same
// virtual pitch) and never read.  Weights (28 KB, pre-arranged in the canonical layout) stay resident in smem.
#pragma once
#include "conv_gemm.cuh"

namespace vnect {

constexpr int kStemStages = 16;
constexpr int kStemStripBytes = kBlockM * 16 + 64;  // 128 rows x 16 B + 48 B window tail, rounded to 16 B multiple
constexpr int kStemWBytes = 28 * 64 * 16;           // [28 K-chunks][64 couts][8 halves]

struct StemParams {
  const uint8_t* x1;          // stem-layout input
  int64_t plane_bytes;        // bytes of one parity plane (rpp * row pitch * 2)
  const uint8_t* w;           // packed weights, kStemWBytes
  const float* bias;          // [64]
  int vw;                     // virtual output columns per row (S/2 + 3)
  int tiles_per_image;        // ceil(S/2 * vw / 128)
  int num_tiles;              // images in this launch * tiles_per_image
  int img0;                   // first image of this launch (input side; the output buffer is a per-launch ring)
};

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// no-swizzle K-major descriptor: LBO = byte distance between the two 16-byte K chunks of one MMA, SBO = byte distance
// between 8-row groups
__device__ __forceinline__ uint64_t make_noswz_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= 1ull << 46;
  return d;
}

struct StemSmem {
  static constexpr int W_OFF = 0;
  static constexpr int STRIP_OFF = kStemWBytes;                                   // 28672
  static constexpr int OUT_OFF = ((STRIP_OFF + kStemStages * kStemStripBytes + 1023) / 1024) * 1024;
  static constexpr int BAR_OFF = OUT_OFF + kOutStages * kEpiChunkBytes;
  static constexpr int BYTES = BAR_OFF + 512 + 1024;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
stem_gemm_kernel(const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ StemParams p) {
  constexpr uint32_t IDESC = make_idesc_f16(kBlockM, 64, false);
  constexpr uint32_t TMEM_COLS = 128;  // two 64-column accumulators
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem + StemSmem::W_OFF;
  uint8_t* strips = smem + StemSmem::STRIP_OFF;
  uint8_t* out_stage = smem + StemSmem::OUT_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + StemSmem::BAR_OFF);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStemStages;
  uint64_t* tmem_full = bars + 2 * kStemStages;
  uint64_t* tmem_empty = bars + 2 * kStemStages + 2;
  uint64_t* w_bar = bars + 2 * kStemStages + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStemStages + 5);
  float* bias_s = reinterpret_cast<float*>(bars + 2 * kStemStages + 6);  // 64 floats (BAR region has 512 B)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < kStemStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  if (threadIdx.x >= 128 && threadIdx.x < 192) bias_s[threadIdx.x - 128] = p.bias[threadIdx.x - 128];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(w_bar, kStemWBytes);
      bulk_load_1d(w_smem, p.w, kStemWBytes, w_bar);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int img = tile / p.tiles_per_image;
        const int v0 = (tile - img * p.tiles_per_image) * kBlockM;
        const uint8_t* img_base = p.x1 + static_cast<int64_t>(img + p.img0) * 2 * p.plane_bytes;
        for (int ky = 0; ky < 7; ++ky) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], kStemStripBytes);
          const uint8_t* src = img_base + (ky & 1) * p.plane_bytes + 16ll * (v0 + p.vw * (ky >> 1));
          bulk_load_1d(strips + stage * kStemStripBytes, src, kStemStripBytes, &full_bar[stage]);
          if (++stage == kStemStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(w_bar, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t w_addr = smem_u32(w_smem);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 64);
        for (int ky = 0; ky < 7; ++ky) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(strips + stage * kStemStripBytes);
#pragma unroll
          for (int j = 0; j < 2; ++j) {  // K = 32 per row tap = two MMAs of K = 16 (two 16-byte chunks each)
            const uint64_t ad = make_noswz_desc(a_addr + 32 * j, 16, 128);
            const uint64_t bd = make_noswz_desc(w_addr + (ky * 4 + 2 * j) * 1024, 1024, 128);
            umma_f16(d_tmem, ad, bd, IDESC, (ky | j) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kStemStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const bool leader = (threadIdx.x == 128);
    const uint32_t row_off = static_cast<uint32_t>(r) * 128u;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    int acc = 0;
    uint32_t acc_phase = 0, ctr = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++ctr) {
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * 64);
      uint32_t v[64];
      tmem_ld_32x32(t_row, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld_32x32(t_row + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      uint8_t* ostage = out_stage + (ctr % kOutStages) * kEpiChunkBytes;
      if (leader) bulk_wait_group_read<kOutStages - 1>();
      named_bar_sync(1, 128);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + 8 * j);
        const float4 b1 = *reinterpret_cast<const float4*>(bias_s + 8 * j + 4);
        uint4 o;
        o.x = pack_half2(fmaxf(__uint_as_float(v[8 * j + 0]) + b0.x, 0.f), fmaxf(__uint_as_float(v[8 * j + 1]) + b0.y, 0.f));
        o.y = pack_half2(fmaxf(__uint_as_float(v[8 * j + 2]) + b0.z, 0.f), fmaxf(__uint_as_float(v[8 * j + 3]) + b0.w, 0.f));
        o.z = pack_half2(fmaxf(__uint_as_float(v[8 * j + 4]) + b1.x, 0.f), fmaxf(__uint_as_float(v[8 * j + 5]) + b1.y, 0.f));
        o.w = pack_half2(fmaxf(__uint_as_float(v[8 * j + 6]) + b1.z, 0.f), fmaxf(__uint_as_float(v[8 * j + 7]) + b1.w, 0.f));
        *reinterpret_cast<uint4*>(ostage + row_off + ((static_cast<uint32_t>(j) ^ sw) << 4)) = o;
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (leader) {
        tma_store_5d(&tmap_out, ostage, 0, tile * kBlockM, 0, 0, 0);  // output rows = global virtual pixel index
        bulk_commit_group();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (leader) bulk_wait_group<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace vnect
