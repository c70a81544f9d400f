// Halo-reuse 3x3 convolution for the narrow, full-resolution layers (sm_100a, tcgen05 + TMA).
//
// Why a second conv kernel.  rv_conv_tc.cu re-loads the activation tile once per tap (9x) and the weight tile
// once per 128 output pixels.  Measured on B200 that kernel is bound by the L2->SM operand stream (~45-55 B/cycle
// per SM): for Cout = 96 every tap needs 24 KB of A + 24 KB of B for 384 cycles of MMA, i.e. 125 B/cycle, so the
// 96-channel 1024^2 layers of the Qwen VAE (28 % of its FLOPs) ran at 0.52 PFLOP/s while the 512-channel layers hit
// 1.4.  Here each input row is loaded ONCE per strip and each weight tap once per TWO output rows:
//
//   * a CTA walks a strip: 128 output columns x R output rows of one image, two rows (M = 2 x 128) per step;
//   * input rows live in a 6-slot shared-memory ring, each slot = the row's 130 pixels (1-pixel halo both sides)
//     as K-major SWIZZLE_128B [130][64ch] (+ SWIZZLE_64B [130][32ch] when Cin % 64 == 32), written by TMA; image
//     borders are TMA out-of-bounds zero fill;
//   * the A operand of tap (dy,dx) for output row y is the slot of input row y+dy-1 addressed from its row dx on:
//     the UMMA swizzle is a function of the absolute shared-memory address, so a descriptor may start at any row
//     (verified in scripts/experiments/umma_shift_test.cu) -- no im2col, no per-tap copies;
//   * weights stream per tap through a 3-slot ring and are consumed by both output rows (24 MMAs per tap);
//   * accumulators: 2 rows x N columns, double buffered in TMEM (4N <= 512 columns);
//   * epilogue (8 warps): the lean path of rv_tc_common.cuh incl. the fused RMS-norm + SiLU second output.
//
// L2->SM traffic per two output rows (Cin = Cout = 96): 2 x 25 KB of activations + 9 x 18 KB of weights = 212 KB for
// 5184 MMA cycles = 41 B/cycle (was 125).
//
// PAIR = true (cta_group::2): with both operands in shared memory an M = 128, N = 96 MMA reads 4 KB of A + 3 KB of B for
// 48 tensor-pipe cycles -- 146 B/cycle against the SM's 128 B/cycle operand path, so the single-CTA form tops out near
// 0.75 PFLOP/s.  Two CTAs on adjacent strips of identical shape run in lockstep as one M = 256 MMA: each SM reads its own
// 128 pixel rows of A but only HALF of the weight tap (the tap is split between the two shared memories, and its L2
// traffic halves too).  Protocol as in conv_tc2_kernel: both CTAs' TMA loads complete on the LEADER's full barriers, only
// the leader's MMA warp issues, tcgen05.commit multicasts every release to the same barrier in both CTAs, the peer's
// epilogue warps release accumulators on the leader's barrier through a remote mbarrier arrive.
#include <cstring>
#include <mutex>

#include "rv_tc_common.cuh"

namespace rv {

constexpr int HL_RING = 6;         // input-row slots; 5 for Cin = 128 (two 64-channel parts per row: six would not fit)
constexpr int hl_ring_slots(int nk128) { return nk128 >= 2 ? 5 : HL_RING; }
constexpr int HL_BRING = 3;       // weight-tap ring slots of the single-CTA form
constexpr int HL_BRING_MAX = 8;   // CTA pairs hold half a tap per slot: the same bytes give a deeper ring
constexpr int HL_EPI_WARPS = 8;
constexpr int HL_THREADS = 128 + 32 * HL_EPI_WARPS;  // row-TMA, MMA (row 0), weight-TMA, MMA (row 1), 8 epilogue warps
constexpr int HL_PIX = 130;                          // 128 output columns + halo
// Row slots are packed at 128-byte granularity (TMA's destination alignment), not at the 1024-byte swizzle period: TMA and
// UMMA both derive the swizzle XOR from the ABSOLUTE shared-memory address, so a slot may start at any 128-byte row (the
// same property that lets the A descriptor start at row dx).  The 10 KB this saves over six slots pays for the staging boxes.
constexpr uint32_t HL_R128_BYTES = 130 * 128;        // [130 pixels][64 ch] SWIZZLE_128B
constexpr uint32_t HL_R64_BYTES = 130 * 64;          // [130 pixels][32 ch] SWIZZLE_64B
constexpr uint32_t HL_SMEM_MAX = 227 * 1024 - 4096;  // dynamic budget next to < 4 KB of static shared memory

struct HaloParams {
  int n_img, h, w;
  int cin, bn;                 // bn = cout (multiple of 16, <= 128)
  int nk128, has64;            // K blocks per tap: nk128 x 64 channels (SW128) + optionally 32 channels (SW64)
  int col_blocks, strips_per_col, strip_rows, total_strips;
  uint32_t row_slot_bytes, r64_off, row_tx_bytes;
  uint32_t b_slot_bytes, b64_off, b_tx_bytes;
  uint32_t bring_off;          // offset of the weight ring behind the row ring
  int bring_slots;             // weight-tap ring depth
  EpiParams e;
};

struct StripCoord {
  int img, x0, ys, rows;  // rows: even number of output rows walked (may run past h by one masked row)
};
// PAIR: strips are ordered column-block fastest so that strips 2k and 2k+1 (the two CTAs of a pair) share sy, hence shape
template <bool PAIR>
__device__ __forceinline__ StripCoord decode_strip(const HaloParams& p, int s) {
  StripCoord c;
  const int ncol = p.n_img * p.col_blocks;
  const int sy = PAIR ? s / ncol : s % p.strips_per_col;
  const int t = PAIR ? s - sy * ncol : s / p.strips_per_col;
  c.x0 = (t % p.col_blocks) * 128;
  c.img = t / p.col_blocks;
  c.ys = sy * p.strip_rows;
  int rows = p.h - c.ys;
  if (rows > p.strip_rows) rows = p.strip_rows;
  c.rows = (rows + 1) & ~1;
  return c;
}

template <int NK128, int HAS64, bool PAIR>  // K blocks per tap: NK128 x 64 channels (+ 32 channels); PAIR: cta_group::2
__global__ void __launch_bounds__(HL_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a128, const __grid_constant__ CUtensorMap map_a64,
                 const __grid_constant__ CUtensorMap map_b128, const __grid_constant__ CUtensorMap map_b64,
                 const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_rowfull[HL_RING];
  __shared__ __align__(8) uint64_t bar_rowempty[HL_RING];
  __shared__ __align__(8) uint64_t bar_bfull[HL_BRING_MAX];
  __shared__ __align__(8) uint64_t bar_bempty[HL_BRING_MAX];
  __shared__ __align__(8) uint64_t bar_accfull[2];
  __shared__ __align__(8) uint64_t bar_accempty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[128];
  __shared__ __align__(16) float s_gamma[128];
  // [tile row][column half][pixel]: fused-norm partial sums.  One copy serves both accumulator buffers: the two warps that
  // exchange through a slot meet at the row's named barrier, and meet again (other row) before either reuses the slot.
  __shared__ float s_ss[2][2][128];

  // A tile keeps four input rows live (y-1 .. y+2) and frees two when it completes.  With six slots both rows of the next tile
  // prefetch during the current one; with five (Cin = 128) the second one loads right after the free -- it is needed only by
  // the last third of output row 1's taps, ~2 us into a 5.8 us tile.
  constexpr uint32_t RING = (uint32_t)hl_ring_slots(NK128);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bring = ring + p.bring_off;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  // work items: strips (single) or pairs of same-shape strips (2u, 2u+1); CTA `rank` of a pair takes strip 2u + rank
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int total_units = PAIR ? (p.total_strips >> 1) : p.total_strips;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < (int)RING; ++s) {
      mbar_init(smem_u32(&bar_rowfull[s]), 1);
      mbar_init(smem_u32(&bar_rowempty[s]), 2);  // released by both MMA issuers
    }
    for (int s = 0; s < p.bring_slots; ++s) {
      mbar_init(smem_u32(&bar_bfull[s]), 1);
      mbar_init(smem_u32(&bar_bempty[s]), 2);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bar_accfull[a]), 2);
      mbar_init(smem_u32(&bar_accempty[a]), PAIR ? 2 * HL_EPI_WARPS : HL_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a128) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b128) : "memory");
  }
  if (warp == 1) {
    __syncwarp();
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                   "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                   "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (p.e.fast && p.e.bias_mode == 1)
    for (int i = threadIdx.x; i < p.e.cout; i += HL_THREADS) s_bias[i] = p.e.bias[i];
  if (p.e.norm_gamma)
    for (int i = threadIdx.x; i < p.e.cout; i += HL_THREADS) s_gamma[i] = p.e.norm_gamma[i] * (p.e.norm_silu ? 0.5f : 1.0f);  // SiLU's 1/2 rides in gamma
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const uint32_t rowfull0 = smem_u32(&bar_rowfull[0]), rowempty0 = smem_u32(&bar_rowempty[0]);
  const uint32_t bfull0 = smem_u32(&bar_bfull[0]), bempty0 = smem_u32(&bar_bempty[0]);
  const uint32_t accfull0 = smem_u32(&bar_accfull[0]), accempty0 = smem_u32(&bar_accempty[0]);
  // the leader's barriers as cluster addresses (PAIR): TMA completions and the peer's accumulator releases land there
  const uint32_t lead_rowfull0 = PAIR ? mapa_rank(rowfull0, 0) : rowfull0;
  const uint32_t lead_bfull0 = PAIR ? mapa_rank(bfull0, 0) : bfull0;
  const uint32_t lead_accempty0 = PAIR ? mapa_rank(accempty0, 0) : accempty0;
  const uint32_t bn_cta = PAIR ? (uint32_t)p.bn >> 1 : (uint32_t)p.bn;  // weight rows held by this CTA

  if (warp == 0) {
    // ------------------------------ input-row producer ------------------------------
    uint32_t g = 0;  // rows loaded so far (ring position)
    for (int u = unit0; u < total_units; u += unit_step) {
      const StripCoord c = decode_strip<PAIR>(p, PAIR ? 2 * u + (int)rank : u);
      for (int y = c.ys - 1; y <= c.ys + c.rows; ++y) {
        const uint32_t slot = g % RING, par = (g / RING) & 1u;
        mbar_wait(rowempty0 + 8u * slot, par ^ 1u);
        if (elect_one()) {
          const uint32_t dst = ring + slot * p.row_slot_bytes;
          if (PAIR) {
            const uint32_t full = lead_rowfull0 + 8u * slot;
            if (leader) mbar_arrive_expect_tx(rowfull0 + 8u * slot, 2u * p.row_tx_bytes);
            for (int kb = 0; kb < p.nk128; ++kb) tma2_load_4d(dst + kb * HL_R128_BYTES, &map_a128, full, kb * 64, c.x0 - 1, y, c.img);
            if (p.has64) tma2_load_4d(dst + p.r64_off, &map_a64, full, p.nk128 * 64, c.x0 - 1, y, c.img);
          } else {
            const uint32_t full = rowfull0 + 8u * slot;
            mbar_arrive_expect_tx(full, p.row_tx_bytes);
            for (int kb = 0; kb < p.nk128; ++kb) tma_load_4d(dst + kb * HL_R128_BYTES, &map_a128, full, kb * 64, c.x0 - 1, y, c.img);
            if (p.has64) tma_load_4d(dst + p.r64_off, &map_a64, full, p.nk128 * 64, c.x0 - 1, y, c.img);
          }
        }
        __syncwarp();
        ++g;
      }
    }
  } else if (warp == 2) {
    // ------------------------------ weight-tap producer ------------------------------
    uint32_t t = 0;
    for (int u = unit0; u < total_units; u += unit_step) {
      const StripCoord c = decode_strip<PAIR>(p, PAIR ? 2 * u + (int)rank : u);
      const int ntiles = c.rows >> 1;
      for (int j = 0; j < ntiles; ++j) {
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t slot = t % (uint32_t)p.bring_slots, par = (t / (uint32_t)p.bring_slots) & 1u;
          mbar_wait(bempty0 + 8u * slot, par ^ 1u);
          if (elect_one()) {
            const uint32_t dst = bring + slot * p.b_slot_bytes;
            if (PAIR) {  // this CTA's half of the tap's Cout rows; completion on the leader's barrier
              const uint32_t full = lead_bfull0 + 8u * slot;
              const int row0 = (int)(rank * bn_cta);
              if (leader) mbar_arrive_expect_tx(bfull0 + 8u * slot, 2u * p.b_tx_bytes);
              for (int kb = 0; kb < p.nk128; ++kb)
                tma2_load_2d(dst + kb * bn_cta * 128u, &map_b128, full, tap * p.cin + kb * 64, row0);
              if (p.has64) tma2_load_2d(dst + p.b64_off, &map_b64, full, tap * p.cin + p.nk128 * 64, row0);
            } else {
              const uint32_t full = bfull0 + 8u * slot;
              mbar_arrive_expect_tx(full, p.b_tx_bytes);
              for (int kb = 0; kb < p.nk128; ++kb)
                tma_load_2d(dst + kb * (uint32_t)p.bn * 128u, &map_b128, full, tap * p.cin + kb * 64, 0);
              if (p.has64) tma_load_2d(dst + p.b64_off, &map_b64, full, tap * p.cin + p.nk128 * 64, 0);
            }
          }
          __syncwarp();
          ++t;
        }
      }
    }
  } else if ((warp == 1 || warp == 3) && leader) {
    // ------------------------------ MMA issuers (PAIR: the leader CTA only) ------------------------------
    // Two issuing warps, one per output row of the tile (warp 1: row 0 / accumulator d0, warp 3: row 1 / d1): a tcgen05.mma costs
    // its issuing thread a fixed ~46 cycles on top of the N/2 the tensor pipe needs, so with N = 96 one
    // issuer leaves the pipe half idle; two independent instruction streams overlap that cost.  Every barrier an issuer releases
    // (weight slot, accumulator, input rows) therefore counts two arrivals.
    const int my_r = warp == 3 ? 1 : 0;
    // Everything that can be hoisted is: the four live row-slot addresses per tile, the weight-slot address
    // (advanced incrementally), descriptor high words; the 9 taps are unrolled so (dy, dx) are immediates.
    // (Uniform-datapath integer ops cost ~10 cycles each when dependent: an un-hoisted loop body of ~77 of them
    // per 6 MMAs made this warp, not the tensor pipe, the bound.)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
    const uint64_t hi128 = make_smem_desc(0u, 1024u, 2u);
    const uint64_t hi64 = make_smem_desc(0u, 512u, 4u);
    const uint32_t bn = (uint32_t)p.bn;
    auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
      if (PAIR) umma2_bf16(d, a, b, idesc, acc);
      else umma_bf16(d, a, b, idesc, acc);
    };
    auto commit = [&](uint32_t bar) {
      if (PAIR) umma2_commit_both(bar);
      else umma_commit(bar);
    };
    const uint32_t b64_lo = p.b64_off >> 4, r64_lo = p.r64_off >> 4;
    uint32_t g0 = 0, tc = 0;
    uint32_t bslot = 0, bpar = 0;
    for (int u = unit0; u < total_units; u += unit_step) {
      const StripCoord c = decode_strip<PAIR>(p, PAIR ? 2 * u : u);
      const int ntiles = c.rows >> 1;
      for (int j = 0; j < ntiles; ++j) {
        const uint32_t buf = tc & 1u;
        mbar_wait(accempty0 + 8u * buf, ((tc >> 1) & 1u) ^ 1u);
        uint32_t a_lo[4], rslot[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t g = g0 + 2u * j + i;
          rslot[i] = g % RING;
          a_lo[i] = ((ring + rslot[i] * p.row_slot_bytes) & 0x3FFFFu) >> 4;
          if ((j == 0 || i >= 2) && i >= my_r && i <= my_r + 2) mbar_wait(rowfull0 + 8u * rslot[i], (g / RING) & 1u);
        }
        tc_fence_after();
        // the three input rows this issuer's output row reads (static indices: no local-memory array)
        const uint32_t a3[3] = {my_r ? a_lo[1] : a_lo[0], my_r ? a_lo[2] : a_lo[1], my_r ? a_lo[3] : a_lo[2]};
        const uint32_t d_tmem = tmem_base + buf * 2u * bn + (my_r ? bn : 0u);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3, dx = tap - 3 * dy;
          mbar_wait(bfull0 + 8u * bslot, bpar);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_lo = ((bring + bslot * p.b_slot_bytes) & 0x3FFFFu) >> 4;
            {
              const uint32_t ar = a3[dy];
#pragma unroll
              for (int kb = 0; kb < NK128; ++kb) {
                const uint64_t ad = hi128 | (uint64_t)(ar + kb * (HL_R128_BYTES >> 4) + dx * 8u);
                const uint64_t bd = hi128 | (uint64_t)(b_lo + kb * bn_cta * 8u);
                mma(d_tmem, ad, bd, (tap == 0 && kb == 0) ? 0u : 1u);
                mma(d_tmem, ad + 2u, bd + 2u, 1u);
                mma(d_tmem, ad + 4u, bd + 4u, 1u);
                mma(d_tmem, ad + 6u, bd + 6u, 1u);
              }
              if (HAS64) {
                const uint64_t ad = hi64 | (uint64_t)(ar + r64_lo + dx * 4u);
                const uint64_t bd = hi64 | (uint64_t)(b_lo + b64_lo);
                mma(d_tmem, ad, bd, (tap == 0 && NK128 == 0) ? 0u : 1u);
                mma(d_tmem, ad + 2u, bd + 2u, 1u);
              }
            }
            commit(bempty0 + 8u * bslot);
            if (tap == 8) {
              commit(accfull0 + 8u * buf);
              // rows y-1 and y of this tile are dead now; the strip's last tile frees its remaining two as well
              commit(rowempty0 + 8u * rslot[0]);
              commit(rowempty0 + 8u * rslot[1]);
              if (j == ntiles - 1) {
                commit(rowempty0 + 8u * rslot[2]);
                commit(rowempty0 + 8u * rslot[3]);
              }
            }
          }
          __syncwarp();
          if (++bslot == (uint32_t)p.bring_slots) {
            bslot = 0;
            bpar ^= 1u;
          }
        }
        ++tc;
      }
      g0 += (uint32_t)c.rows + 2u;
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int nsplit = (p.bn % 32 == 0) ? 2 : 1;
    const int cb = nsplit == 2 ? half * (p.bn >> 1) : 0;
    const int ce = nsplit == 2 ? cb + (p.bn >> 1) : (half == 0 ? p.bn : 0);
    const float* sbias = p.e.bias_mode == 1 ? s_bias : nullptr;
    {
      uint32_t tc = 0;
      for (int u = unit0; u < total_units; u += unit_step) {
        const StripCoord c = decode_strip<PAIR>(p, PAIR ? 2 * u + (int)rank : u);
        const int ntiles = c.rows >> 1;
        const int x = c.x0 + row;
        for (int j = 0; j < ntiles; ++j) {
          const uint32_t buf = tc & 1u;
          mbar_wait(accfull0 + 8u * buf, (tc >> 1) & 1u);
          tc_fence_after();
          for (int r = 0; r < 2; ++r) {
            const int y = c.ys + 2 * j + r;
            const bool valid = x < p.w && y < p.h;
            const int64_t pix = ((int64_t)c.img * p.h + y) * p.w + x;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 2u * (uint32_t)p.bn + (uint32_t)r * (uint32_t)p.bn;
            if (!p.e.fast)  // NCHW / fp32 / affine / clamp outputs (conv_out): generic epilogue
              epilogue_pixel(p.e, taddr, cb, ce, 0, valid, c.img, y, x, pix);
            else if (p.e.residual)
              epilogue_pixel_fast<true>(p.e, sbias, s_gamma, taddr, cb, ce, 0, valid, pix, &s_ss[r][0][0], row, half, nsplit, 1 + q);
            else
              epilogue_pixel_fast<false>(p.e, sbias, s_gamma, taddr, cb, ce, 0, valid, pix, &s_ss[r][0][0], row, half, nsplit, 1 + q);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(lead_accempty0 + 8u * buf);
            else mbar_arrive(accempty0 + 8u * buf);
          }
          ++tc;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // nobody frees TMEM or exits while the peer may still signal / read
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
int tc_encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box, CUtensorMapSwizzle sw);  // rv_conv_tc.cu
void fill_epi(EpiParams* e, const rv_conv_desc* d, const float* bias, const void* residual, void* y, const NormFuse* nf);
bool epi_fast_ok(const EpiParams& e, const rv_conv_desc* d, const float* bias, const NormFuse* nf, int covered_cols);

static std::mutex g_halo_mu;
static bool g_halo_attr[64] = {false};

// CTA pairs (cta_group::2): the weight tap splits into two halves of whole 8-row swizzle groups and the strips pair up whenever
// the image x column-block count is even.  (The pair form lost to the single-CTA one as long as the peer's accumulator release
// was a cluster-scope RELEASE arrive -- every epilogue warp drained its global loads / stores before each arrive; with the plain
// arrive it is 11-14 % faster on the 96-channel 1024^2 layers: DESIGN.md 4, item 13.)
static bool halo_pair_ok(const rv_conv_desc* d, int bn) {
  static const bool want_pair = getenv("RGBAVAE_HALO_NO_PAIR") == nullptr;
  return want_pair && bn % 32 == 0 && ((d->n * ((d->w + 127) / 128)) % 2 == 0);
}

// Can this convolution run on the halo kernel?  (3x3 stride-1, Cout <= 128, operands fit the rings.)
bool halo_eligible(const rv_conv_desc* d, const EpiParams& e) {
  if (d->ksize != 3 || d->stride != 1 || d->upsample || d->pad_lo != 1 || d->taps_1d) return false;
  if (!(d->cin == 64 || d->cin == 96 || d->cin == 128) || d->cout > 128) return false;
  if (e.norm_gamma && !e.fast) return false;
  if (d->bias_mode == 2 || d->w < 64 || d->h < 2) return false;
  const int nk128 = d->cin / 64, has64 = (d->cin % 64) ? 1 : 0;
  const uint32_t row_slot = nk128 * HL_R128_BYTES + has64 * HL_R64_BYTES;
  const int bn = (d->cout + 15) / 16 * 16;
  // the CTA-pair form holds half a weight tap per slot (what makes Cin = 128 fit at all)
  const uint32_t rows_cta = halo_pair_ok(d, bn) ? (uint32_t)bn / 2u : (uint32_t)bn;
  const uint32_t b_slot = (rows_cta * (uint32_t)d->cin * 2u + 1023u) & ~1023u;
  return (((uint32_t)hl_ring_slots(nk128) * row_slot + 1023u) & ~1023u) + HL_BRING * b_slot + 1024u <= HL_SMEM_MAX;
}

int launch_halo(const rv_conv_desc* d, const void* x, const void* w, int64_t w_ld, const float* bias, const void* residual,
                void* y, cudaStream_t st, const NormFuse* nf) {
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = d->n;
  p.h = d->h;
  p.w = d->w;
  p.cin = d->cin;
  p.bn = (d->cout + 15) / 16 * 16;
  p.nk128 = d->cin / 64;
  p.has64 = (d->cin % 64) ? 1 : 0;
  p.row_slot_bytes = p.nk128 * HL_R128_BYTES + p.has64 * HL_R64_BYTES;
  p.r64_off = p.nk128 * HL_R128_BYTES;
  p.row_tx_bytes = (uint32_t)HL_PIX * (uint32_t)(p.nk128 * 128 + p.has64 * 64);
  p.col_blocks = (d->w + 127) / 128;
  const bool pair = halo_pair_ok(d, p.bn);
  const uint32_t bn_cta = pair ? (uint32_t)p.bn / 2u : (uint32_t)p.bn;
  p.b64_off = (uint32_t)p.nk128 * bn_cta * 128u;
  p.b_tx_bytes = bn_cta * (uint32_t)d->cin * 2u;  // per CTA; full box bytes (rows past cout are zero-filled)
  p.b_slot_bytes = (p.b_tx_bytes + 1023u) & ~1023u;
  p.bring_slots = HL_BRING;
  fill_epi(&p.e, d, bias, residual, y, nf);
  p.e.fast = epi_fast_ok(p.e, d, bias, nf, p.bn) ? 1 : 0;
  p.bring_off = ((uint32_t)hl_ring_slots(p.nk128) * p.row_slot_bytes + 1023u) & ~1023u;
  if (pair) {  // half-tap slots: the single-CTA ring's bytes (and the rest of the budget) give a deeper weight ring
    p.bring_slots = (int)((HL_SMEM_MAX - 1024u - p.bring_off) / p.b_slot_bytes);
    if (p.bring_slots > HL_BRING_MAX) p.bring_slots = HL_BRING_MAX;
  }
  // strips: ~12 per SM, an even number of rows each, at least 8
  int64_t cols = (int64_t)d->n * p.col_blocks;
  int rows = (int)(((int64_t)d->h * cols + (int64_t)num_sms() * 12 - 1) / ((int64_t)num_sms() * 12));
  rows = (rows + 1) & ~1;
  if (rows < 8) rows = 8;
  if (rows > d->h) rows = (d->h + 1) & ~1;
  p.strip_rows = rows;
  p.strips_per_col = (d->h + rows - 1) / rows;
  p.total_strips = (int)(cols * p.strips_per_col);

  CUtensorMap ma128, ma64, mb128, mb64;
  const uint64_t pitch_b = (uint64_t)d->x_cstride * 2u;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
    cuuint64_t str[3] = {pitch_b, pitch_b * d->w, pitch_b * d->w * d->h};
    cuuint32_t box[4] = {64, (cuuint32_t)HL_PIX, 1, 1};
    if (int rc = tc_encode_map(&ma128, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    cuuint32_t box64[4] = {32, (cuuint32_t)HL_PIX, 1, 1};
    if (int rc = tc_encode_map(&ma64, x, 4, dims, str, box64, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)(9 * d->cin), (cuuint64_t)d->cout};
    cuuint64_t str[1] = {(cuuint64_t)w_ld * 2u};
    cuuint32_t box[2] = {64, (cuuint32_t)bn_cta};
    if (int rc = tc_encode_map(&mb128, w, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    cuuint32_t box64[2] = {32, (cuuint32_t)bn_cta};
    if (int rc = tc_encode_map(&mb64, w, 2, dims, str, box64, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }
  const size_t smem = (size_t)p.bring_off + (size_t)p.bring_slots * p.b_slot_bytes + 1024;
  {
    std::lock_guard<std::mutex> lk(g_halo_mu);
    int dev = 0;
    RV_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !g_halo_attr[dev]) {
      RV_CUDA(cudaFuncSetAttribute(conv_halo_kernel<1, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HL_SMEM_MAX));
      RV_CUDA(cudaFuncSetAttribute(conv_halo_kernel<1, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HL_SMEM_MAX));
      RV_CUDA(cudaFuncSetAttribute(conv_halo_kernel<2, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HL_SMEM_MAX));
      RV_CUDA(cudaFuncSetAttribute(conv_halo_kernel<1, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HL_SMEM_MAX));
      RV_CUDA(cudaFuncSetAttribute(conv_halo_kernel<1, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HL_SMEM_MAX));
      RV_CUDA(cudaFuncSetAttribute(conv_halo_kernel<2, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HL_SMEM_MAX));
      g_halo_attr[dev] = true;
    }
  }
  const double flops = 2.0 * (double)d->n * d->oh * d->ow * d->cout * d->cin * 9.0;
  LaunchScope scope(CAT_CONV_TC, st, flops);
  if (pair) {
    int pairs = p.total_strips / 2;
    if (pairs > num_sms() / 2) pairs = num_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(HL_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (p.nk128 == 1 && !p.has64) RV_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<1, 0, true>, ma128, ma64, mb128, mb64, p));
    else if (p.nk128 == 1 && p.has64) RV_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<1, 1, true>, ma128, ma64, mb128, mb64, p));
    else RV_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<2, 0, true>, ma128, ma64, mb128, mb64, p));
  } else {
    int grid = p.total_strips < num_sms() ? p.total_strips : num_sms();
    if (p.nk128 == 1 && !p.has64) conv_halo_kernel<1, 0, false><<<grid, HL_THREADS, smem, st>>>(ma128, ma64, mb128, mb64, p);
    else if (p.nk128 == 1 && p.has64) conv_halo_kernel<1, 1, false><<<grid, HL_THREADS, smem, st>>>(ma128, ma64, mb128, mb64, p);
    else conv_halo_kernel<2, 0, false><<<grid, HL_THREADS, smem, st>>>(ma128, ma64, mb128, mb64, p);
  }
  RV_LAUNCH_CHECK();
  return 0;
}

}  // namespace rv
