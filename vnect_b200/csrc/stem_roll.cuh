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
// Input ("stem layout", written by pyramid_kernel): [forward][row parity][rpp rows][row pitch] fp16, NHWC padded to
// 4 channels; the row pitch is (S+6)*4 halves = (S/2+3)*16 bytes, so with VW = S/2+3 "virtual" conv columns per row the
// 8-pixel x 4-channel window of virtual conv pixel v = oy*VW + ox for row tap ky starts at byte
//     plane(ky & 1) + 16 * (v + VW * (ky >> 1))
// i.e. consecutive GEMM rows are exactly 16 bytes apart.  That is the canonical no-swizzle K-major UMMA layout (8-row
// core matrices of 16-byte rows, SBO = 128 B between row groups) with overlapping K chunks (LBO = 16 B), so a
// contiguous strip of 128*16+48 bytes is a valid 128 x 32 A tile: the A operand is read IN PLACE, no im2col.
//
// A CTA's unit of work: (forward, x tile of 128 conv columns, segment of pooled rows).  Strips are raw 2112-byte
// windows of one padded input row, fetched with one bulk copy each.  Accumulators are only ever accumulated into: the epilogue warps zero a slot
// (tcgen05.st) right after draining it.  Each epilogue thread drains the same conv column of every row (bias, ReLU,
// fp16), keeps the vertical 3-max of the open pooling window in registers, and every second row the 8 epilogue
// warps exchange the column maxima through smem for the horizontal 3-max and write one pooled row straight to HBM.
#pragma once
#include "conv_gemm.cuh"

namespace vnect {

constexpr int kStemWBytes = 28 * 64 * 16;  // 7 row taps x 32 K values x 64 couts, fp16

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int kRollThreads = 128 + 256;      // 4 control warps + 8 epilogue / pooling warps
constexpr int kRollStripLoad = kBlockM * 16 + 64;
constexpr int kRollStripBytes = 2176;        // kRollStripLoad rounded up to a multiple of 128
constexpr int kRollStages = 16;
constexpr int kRollRowBufs = 2;         // column-maxima buffers (alternating pooling windows)
constexpr int kRollSlots = 8;                // 8 x 64 fp32 columns = all of TMEM
constexpr int kRollRowBytes = kBlockM * 128;
constexpr int kMaxXTiles = 4;
constexpr int kRollMaxInputRows = 4 * 128 + 8;  // a segment of up to 128 pooled rows (box size 512)
constexpr int kRollWEvenBytes = 256 * 64;  // [4 taps x 64 couts][32 K] fp16, 64B-swizzled K-major rows
constexpr int kRollWOddBytes = 192 * 64;   // [3 taps x 64 couts][32 K]
static_assert(kRollWEvenBytes + kRollWOddBytes == kStemWBytes, "the two stacks hold all 7 row taps");

// CTA pairs (cta_group::2): the two CTAs of a cluster take the two x tiles of one (forward, segment), the leader issues
// M = 256 MMAs and each CTA supplies only HALF of the stacked-weight operand -- the MMA is bound by reading its
// operands from shared memory (12 KB per 128 x 256 x 16 step with a single CTA, 8 KB per CTA in a pair).  An MMA over
// `cnt` stacked taps starting at group g reads rows [64 g + 32 cnt rank, + 32 cnt) of the stack in CTA `rank`, from the
// SAME shared-memory offset in both CTAs, so every (g, cnt) gets its own 2 KB * cnt entry in a per-rank table:
//   even rows: g = 0..3, cnt = 1..4-g (20 units of 2 KB), odd rows: g = 0..2, cnt = 1..3-g (10 units).
constexpr int kPairUnitBytes = 32 * 64;                 // 32 weight rows of 32 K values
constexpr int kPairWEvenBytes = 20 * kPairUnitBytes;
constexpr int kPairWBytes = 30 * kPairUnitBytes;        // per rank
__host__ __device__ constexpr int pair_w_units(int n_groups, int g, int cnt) {  // entry offset in units, stack of n_groups
  int u = 0;
  for (int gg = 0; gg < g; ++gg) u += (n_groups - gg) * (n_groups - gg + 1) / 2;
  return u + cnt * (cnt - 1) / 2;
}

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
  int64_t x1_units;     // CTA pairs: 16-byte units the strip tensor map covers
  unsigned long long* dbg;  // optional [4] cycle counters of CTA 0 (selftest only): wait-for-MMA, drain, pool, total
};

struct StemRollSmem {
  static constexpr int W_OFF = 0;
  static constexpr int STRIP_OFF = kPairWBytes;  // the single-CTA kernel uses the first kStemWBytes of it
  static constexpr int ROW_OFF = ((STRIP_OFF + kRollStages * kRollStripBytes + 1023) / 1024) * 1024;
  static constexpr int BAR_OFF = ROW_OFF + kRollRowBufs * kRollRowBytes;
  static constexpr int CMD_OFF = BAR_OFF + 1024;                 // issue program of one item: 32 B per input row
  static constexpr int BYTES = CMD_OFF + kRollMaxInputRows * 32 + 1024;
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

// DBG = true adds cycle counters (selftest only); the shipped instantiation has none in its loops
template <bool DBG>
__device__ __forceinline__ long long roll_clock() {
  if constexpr (DBG) return clock64();
  else return 0;
}

template <bool DBG = false, int CG = 1>
__global__ void __launch_bounds__(kRollThreads, 1) stem_roll_kernel(const __grid_constant__ StemRollParams p,
                                                                    const __grid_constant__ CUtensorMap tmap_x) {
  constexpr uint32_t TMEM_COLS = 512;
  constexpr int kWBytes = CG == 2 ? kPairWBytes : kStemWBytes;
  const int cta_rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int worker = blockIdx.x / CG, n_workers = gridDim.x / CG, n_units = p.num_items / CG;
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
      mbar_init(&tmem_empty[s], 8 * CG);  // one arrive per epilogue warp (of both CTAs of a pair, on the leader's barrier)
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
    if constexpr (CG == 2) {
      tma_prefetch_desc(&tmap_x);
      // the weights are constants: fetched before the predecessor has finished, and resident in BOTH CTAs before the
      // cluster barrier below lets the leader issue anything that reads the peer's half
      mbar_arrive_expect_tx(w_bar, kWBytes);
      bulk_load_1d(w_smem, p.w + static_cast<size_t>(cta_rank) * kPairWBytes, kWBytes, w_bar);
    }
  }
  if (warp == 2) {
    if constexpr (CG == 2) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
    else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  if (threadIdx.x >= 128 && threadIdx.x < 192) bias_s[threadIdx.x - 128] = p.bias[threadIdx.x - 128];
  pdl_launch_dependents();
  pdl_wait();  // bias (read above) is a constant; the input strips come from the previous kernel
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) {
    if (warp == 0) mbar_wait(w_bar, 0);
    cluster_sync_all();  // the peer's barriers are initialised and its weights resident before anything signals / reads them
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ strip loader: one bulk copy per padded input row
    const bool issuer = elect_one();
    if constexpr (CG == 1) {
      if (issuer) {
        mbar_arrive_expect_tx(w_bar, kStemWBytes);
        bulk_load_1d(w_smem, p.w, kStemWBytes, w_bar);
      }
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int u = worker; u < n_units; u += n_workers) {
      const RollItem R = roll_decode(p, u * CG + cta_rank);
      const int64_t base_off = static_cast<int64_t>(R.img) * 2 * p.plane_bytes + 16ll * p.xt_x0[R.xt];
      for (int ri = 0; ri < R.n_in; ++ri) {
        const int r = 2 * R.ya + ri;
        const int64_t off = base_off + (r & 1) * p.plane_bytes + 16ll * (r >> 1) * p.vw;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (issuer) {
          if constexpr (CG == 2) {  // both CTAs' strips are counted on the leader's barrier
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kRollStripLoad);
            tma_load_2d_pair(strips + stage * kRollStripBytes, &tmap_x, &full_bar[stage], 0, static_cast<int>(off >> 4));
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], kRollStripLoad);
            bulk_load_1d(strips + stage * kRollStripBytes, p.x1 + off, kRollStripLoad, &full_bar[stage]);
          }
        }
        __syncwarp();
        if (++stage == kRollStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    if constexpr (CG == 2) {
      // drain: the leader's multicast commits still arrive on this CTA's empty barriers after its last load; the CTA
      // may not retire before every slot has been released
      for (int s2 = 0; s2 < kRollStages; ++s2) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (++stage == kRollStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ================================================================ MMA issuer (warp-converged, one lane issues)
    // Which accumulators an input row feeds, where a run wraps around the TMEM ring, which slot has to be waited for
    // and which row completes is ~100 dependent scalar instructions per input row; done inline, that arithmetic (not
    // the tensor pipe) set the pace.  So the 32 lanes first write the whole item's issue program into a smem table,
    // one input row per lane, and the issue loop only loads a 32-byte command and fires.
    const bool issuer = elect_one();
    if constexpr (CG == 1) mbar_wait(w_bar, 0);
    const uint32_t w_even = (smem_u32(w_smem) & 0x3FFFF) >> 4, w_odd = w_even + ((CG == 2 ? kPairWEvenBytes : kRollWEvenBytes) >> 4);
    const uint32_t a_lo0 = ((smem_u32(strips) & 0x3FFFF) >> 4) | (1u << 16);  // no-swizzle K-major: LBO = 16 B
    constexpr uint32_t kADescHi = (128u >> 4) | (1u << 14);                    // SBO = 128 B, descriptor version 1
    constexpr uint32_t kBDescHi = (512u >> 4) | (1u << 14) | (4u << 29);       // SBO = 512 B, version 1, 64B swizzle
    constexpr uint32_t kNone = 0xFFFFFFFFu;
    uint4* cmds = reinterpret_cast<uint4*>(smem + StemRollSmem::CMD_OFF);
    int stage = 0;
    uint32_t phase = 0;
    int G0 = 0;  // conv rows this CTA has started before the current item: row G lives in TMEM slot G & 7
    long long m_full = 0, m_empty = 0;
    const long long m_begin = roll_clock<DBG>();
    for (int u = worker; u < n_units; u += n_workers) {
      const RollItem R = roll_decode(p, u * CG + cta_rank);
      for (int ri = lane; ri < R.n_in; ri += 32) {
        const bool even = !(ri & 1);
        const int ly_hi = min(R.n_rows - 1, ri >> 1);
        const int ly_lo = ri >= 6 ? (ri - 5) >> 1 : 0;  // ceil((ri - 6) / 2)
        const int total = ly_hi - ly_lo + 1;            // conv rows this input row feeds (1..4)
        const int slot = (G0 + ly_lo) & 7;
        const int cnt1 = min(total, kRollSlots - slot);  // a run may not wrap around the ring
        const int ky = ri - 2 * ly_lo;                    // tap of the first row; later rows: ky - 2, ...
        const int g = even ? (6 - ky) >> 1 : (5 - ky) >> 1;  // its 64-row group in the stacked weights
        const uint32_t wb = (even ? w_even : w_odd) | (1u << 16);
        const int cnt2 = total - cnt1 > 0 ? total - cnt1 : 1;
        uint4 c0, c1;
        c0.x = static_cast<uint32_t>(slot * 64);                    // run 1: TMEM column offset
        c0.y = make_idesc_f16(kBlockM * CG, 64 * cnt1, false);
        c0.w = cnt1 < total ? 0u : kNone;                            // run 2 starts at slot 0
        c1.x = make_idesc_f16(kBlockM * CG, 64 * cnt2, false);
        if constexpr (CG == 2) {  // per-(group, count) entries of the pair table
          const int ng = even ? 4 : 3;
          c0.z = wb + static_cast<uint32_t>(pair_w_units(ng, g, cnt1) * (kPairUnitBytes >> 4));
          c1.y = wb + static_cast<uint32_t>(pair_w_units(ng, min(g + cnt1, ng - 1), cnt2) * (kPairUnitBytes >> 4));
        } else {
          c0.z = wb + static_cast<uint32_t>(g * (4096 >> 4));        // low word of the B descriptor
          c1.y = wb + static_cast<uint32_t>((g + cnt1) * (4096 >> 4));
        }
        // first tap (ky = 0) of conv row ri/2: its slot must have been drained and zeroed
        const int Gt = G0 + (ri >> 1);
        c1.z = (even && (ri >> 1) < R.n_rows) ? static_cast<uint32_t>((Gt & 7) | (((Gt >> 3) & 1) << 8)) : kNone;
        // last tap (ky = 6) of conv row (ri - 6) / 2: the row is complete
        c1.w = (even && ri >= 6) ? static_cast<uint32_t>((G0 + ((ri - 6) >> 1)) & 7) : kNone;
        cmds[2 * ri] = c0;
        cmds[2 * ri + 1] = c1;
      }
      __syncwarp();
      for (int ri = 0; ri < R.n_in; ++ri) {
        const uint4 c0 = cmds[2 * ri], c1 = cmds[2 * ri + 1];
        const long long m0 = roll_clock<DBG>();
        mbar_wait(&full_bar[stage], phase);
        m_full += roll_clock<DBG>() - m0;
        if (c1.z != kNone) {
          const long long m1 = roll_clock<DBG>();
          mbar_wait(&tmem_empty[c1.z & 7], (c1.z >> 8) & 1);
          m_empty += roll_clock<DBG>() - m1;
        }
        tc_fence_after();
        if (issuer) {
          const uint64_t a_desc = (static_cast<uint64_t>(kADescHi) << 32) | (a_lo0 + static_cast<uint32_t>(stage * (kRollStripBytes >> 4)));
          const uint64_t b1 = (static_cast<uint64_t>(kBDescHi) << 32) | c0.z;
          if constexpr (CG == 2) {
            umma_f16_pair(tmem_base + c0.x, a_desc, b1, c0.y, 1u);
            umma_f16_pair(tmem_base + c0.x, a_desc + 2, b1 + 2, c0.y, 1u);
          } else {
            umma_f16(tmem_base + c0.x, a_desc, b1, c0.y, 1u);
            umma_f16(tmem_base + c0.x, a_desc + 2, b1 + 2, c0.y, 1u);  // +32 B of A, +32 B along K of B
          }
          if (c0.w != kNone) {
            const uint64_t b2 = (static_cast<uint64_t>(kBDescHi) << 32) | c1.y;
            if constexpr (CG == 2) {
              umma_f16_pair(tmem_base, a_desc, b2, c1.x, 1u);
              umma_f16_pair(tmem_base, a_desc + 2, b2 + 2, c1.x, 1u);
            } else {
              umma_f16(tmem_base, a_desc, b2, c1.x, 1u);
              umma_f16(tmem_base, a_desc + 2, b2 + 2, c1.x, 1u);
            }
          }
          if constexpr (CG == 2) {  // both CTAs' epilogues / loaders
            if (c1.w != kNone) umma_commit_pair(&tmem_full[c1.w]);
            umma_commit_pair(&empty_bar[stage]);
          } else {
            if (c1.w != kNone) umma_commit(&tmem_full[c1.w]);
            umma_commit(&empty_bar[stage]);
          }
        }
        __syncwarp();
        if (++stage == kRollStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      __syncwarp();  // the table is rewritten for the next item
      G0 += R.n_rows;
    }
    if (DBG && p.dbg != nullptr && lane == 0 && blockIdx.x == 0) {
      p.dbg[4] = m_full; p.dbg[5] = m_empty; p.dbg[6] = roll_clock<DBG>() - m_begin;
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
      for (int s = 0; s < kRollSlots; ++s) {
        if constexpr (CG == 2) mbar_arrive_leader(&tmem_empty[s]);
        else mbar_arrive(&tmem_empty[s]);
      }
    int G0 = 0;
    uint32_t win = 0;  // pooling windows done by this CTA: column maxima alternate between two smem buffers
    long long t_wait = 0, t_drain = 0, t_pool = 0;
    const long long t_begin = roll_clock<DBG>();
    for (int u = worker; u < n_units; u += n_workers) {
      const RollItem R = roll_decode(p, u * CG + cta_rank);
      const int x0 = p.xt_x0[R.xt], pb = p.xt_pb[R.xt], pe = p.xt_pe[R.xt];
      __half2 row_a[16], row_b[16];  // the open window's first (even) and second (odd) conv row, this thread's column
      for (int ly = 0; ly < R.n_rows; ++ly) {
        const int G = G0 + ly;
        const long long c0 = roll_clock<DBG>();
        mbar_wait(&tmem_full[G & 7], (G >> 3) & 1);
        const long long c1 = roll_clock<DBG>();
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
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_leader(&tmem_empty[G & 7]);
          else mbar_arrive(&tmem_empty[G & 7]);
        }
        __half2 cur[16];  // bias + ReLU + fp16 pack (cvt.rn.relu does the max(.,0) while packing)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t pk = pack_half2_relu(__uint_as_float(v[2 * j]) + bias_r[2 * j], __uint_as_float(v[2 * j + 1]) + bias_r[2 * j + 1]);
          cur[j] = *reinterpret_cast<const __half2*>(&pk);
        }
        const long long c2 = roll_clock<DBG>();
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
          t_pool += roll_clock<DBG>() - c2;
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
    if (DBG && p.dbg != nullptr && et == 0 && blockIdx.x == 0) {
      p.dbg[0] = t_wait; p.dbg[1] = t_drain; p.dbg[2] = t_pool; p.dbg[3] = roll_clock<DBG>() - t_begin;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // neither CTA may retire while its partner can still signal it
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace vnect
