// tcgen05 / TMEM implicit-GEMM convolution fed by TMA  (sm_100a only).
//
// Replaces the cuDNN / cuBLAS calls behind every convolution and projection of diffusers'
// Encoder / Decoder (reference call sites: vae.encode / vae.decode at src/models/rgba_vae.py:277,279,
// src/training/rgba_vae_stage.py:449,452, src/models/flux_kontext_textalpha.py:331,497).
//
// Formulation.  Activations are NHWC bf16.  One CTA tile is 128 output pixels (a bh x bw
// spatial patch of one image, bh*bw = 128) x BN output channels.  The K loop runs over
// (tap, 16/32/64-channel block): for every tap a TMA box load of the patch shifted by the tap
// offset lands in shared memory as a K-major, hardware-swizzled [128][BK] operand -- the
// zero padding of the convolution is TMA's out-of-bounds fill, so there is no im2col buffer
// and no halo logic.  Weights are a plain 2-D K-major matrix [cout][taps*cin].
//   * 3x3 stride 1 pad 1         : 9 taps, box offset (dy-1, dx-1) on the (C, W, H, N) map
//   * 3x3 stride 2 pad (0,1,0,1) : 9 taps on a 5-D (2C, W/2, 2, H/2, N) parity view of x
//   * nearest x2 upsample + 3x3  : four 2x2 phase convolutions on the SOURCE grid (2.25x
//                                  fewer MACs than convolving the upsampled tensor)
//   * 1x1 conv / GEMM            : 1 tap
//
// Warp roles (192 threads, 1 CTA / SM, persistent over tiles):
//   warp 0     TMA producer      (one lane)   smem ring of `stages` {A,B} slots, mbarrier full/empty
//   warp 1     tcgen05.mma issue (one lane)   fp32 accumulators in TMEM, double-buffered (2 x 256 cols)
//   warps 2-9  epilogue          tcgen05.ld -> alpha, bias, residual, scale/shift, clamp -> global
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "rv_tc_common.cuh"

namespace rv {

int check_conv_desc(const rv_conv_desc* d);  // rv_conv_direct.cu

constexpr int TC_EPI_WARPS = 8;                      // two per TMEM lane quarter, each owning half of the tile's columns
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // + TMA warp + MMA warp
constexpr int TC_SMEM_BIAS = 2048;                   // per-channel bias staged in shared memory up to this many channels
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_MAX_TAPS = 9;
constexpr uint32_t TC_SMEM_BUDGET = 200 * 1024;

struct TcParams {
  // tiling of the (image, y, x) space the A operand is gathered from
  int n_img, th, tw;        // extents of the tiled pixel grid (output grid; source grid for upsample)
  int bw, bh;               // tile shape, bw*bh == 128
  int tiles_x, tiles_y;
  int m_tiles, n_tiles;
  int bn, bk;
  int kc_per_tap, ntaps;
  int mode;                 // 0: (C,W,H,N) map   1: stride-2 parity view (2C,W/2,2,H/2,N)
  int cin_pitch;            // pixel pitch of x in elements (parity view: offset of the odd column)
  int cin;
  int w_k_base;             // column offset into the weight matrix (upsample phase)
  int tap_k;                // weight columns per tap (cin, or cin padded to 64 for the stride-2 form)
  signed char tap_dx[TC_MAX_TAPS], tap_dy[TC_MAX_TAPS], tap_wp[TC_MAX_TAPS], tap_hp[TC_MAX_TAPS];
  // output mapping: out pixel = tile pixel * os + oo
  int osy, osx, ooy, oox;
  EpiParams e;
  int stages;
  uint32_t a_bytes, stage_bytes, tx_bytes;
  uint32_t sbo_bytes, layout_type;
};

struct TileCoord {
  int img, y0, x0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int tile) {
  TileCoord t;
  int mt = tile / p.n_tiles;
  t.n0 = (tile - mt * p.n_tiles) * p.bn;
  int tx = mt % p.tiles_x;
  int r = mt / p.tiles_x;
  int ty = r % p.tiles_y;
  t.img = r / p.tiles_y;
  t.x0 = tx * p.bw;
  t.y0 = ty * p.bh;
  return t;
}

template <int KSTEPS>  // BK / 16: 4 (128-byte rows), 2 (64-byte) or 1 (32-byte)
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[TC_SMEM_BIAS];
  __shared__ __align__(16) float s_gamma[256];
  __shared__ float s_ss[2][2][128];  // [accumulator buffer][column half][tile row]: fused-norm partial sums

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int num_kb = p.ntaps * p.kc_per_tap;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bar_acc_full[a]), 1);
      mbar_init(smem_u32(&bar_acc_empty[a]), TC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
  }
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const bool bias_smem = p.e.fast && p.e.bias_mode == 1;
  if (bias_smem)
    for (int i = threadIdx.x; i < p.e.cout; i += TC_THREADS) s_bias[i] = p.e.bias[i];
  if (p.e.norm_gamma)
    for (int i = threadIdx.x; i < p.e.cout; i += TC_THREADS) s_gamma[i] = p.e.norm_gamma[i] * (p.e.norm_silu ? 0.5f : 1.0f);  // SiLU's 1/2 rides in gamma
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // Both single-issuer roles keep their whole warp converged in the loop and elect one lane only around the
  // asynchronous instructions.  (Running the loop under `if (lane == 0)` made ptxas wrap every UTCHMMA / UTMALDG in
  // an ELECT + BRA.U.ANY uniformity loop fed through R2UR moves: ~130 dependent scalar instructions per k-block,
  // i.e. ~215 cycles per MMA regardless of N -- the issue loop, not the tensor pipe, bounded every layer.)
  const uint32_t bar_full0 = smem_u32(&bar_full[0]);
  const uint32_t bar_empty0 = smem_u32(&bar_empty[0]);
  const uint32_t bar_accf0 = smem_u32(&bar_acc_full[0]);
  const uint32_t bar_acce0 = smem_u32(&bar_acc_empty[0]);
  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    uint32_t stage = 0, phase = 0;
    const uint32_t nstages = (uint32_t)p.stages;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const int ax = t.x0 + p.tap_dx[tap], ay = t.y0 + p.tap_dy[tap];
        const int kcol = p.w_k_base + tap * p.tap_k;
        const int c5 = p.tap_wp[tap] * p.cin_pitch, hp = p.tap_hp[tap];
        for (int kc = 0; kc < p.kc_per_tap; ++kc) {
          mbar_wait(bar_empty0 + 8u * stage, phase ^ 1u);
          if (elect_one()) {
            const uint32_t full = bar_full0 + 8u * stage;
            const uint32_t a_dst = smem_base + stage * p.stage_bytes;
            mbar_arrive_expect_tx(full, p.tx_bytes);
            if (p.mode == 0) tma_load_4d(a_dst, &map_a, full, kc * p.bk, ax, ay, t.img);
            else tma_load_5d(a_dst, &map_a, full, c5 + kc * p.bk, ax, hp, ay, t.img);
            tma_load_2d(a_dst + p.a_bytes, &map_b, full, kcol + kc * p.bk, t.n0);
          }
          __syncwarp();
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    // cute::UMMA::InstrDescriptor: D fp32 (bit 4), A bf16 (bit 7), B bf16 (bit 10), both K-major,
    // N>>3 at bit 17, M>>4 at bit 24.
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t desc_hi = make_smem_desc(0u, p.sbo_bytes, p.layout_type);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    const uint32_t nstages = (uint32_t)p.stages;
    // operand addresses advance incrementally (no multiplies / modulo in the issue loop)
    const uint32_t a_lo0 = (smem_base & 0x3FFFFu) >> 4, stage_lo = p.stage_bytes >> 4, ab_lo = p.a_bytes >> 4;
    uint32_t a_lo = a_lo0, full_bar = bar_full0, empty_bar = bar_empty0;
    const int last_kb = num_kb - 1;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(bar_acce0 + 8u * acc, acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256u;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = desc_hi | (uint64_t)a_lo;
          const uint64_t b_desc = desc_hi | (uint64_t)(a_lo + ab_lo);
          // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr>>4) field
          umma_bf16(d_tmem, a_desc, b_desc, idesc, (uint32_t)(kb != 0));
          if (KSTEPS > 1) umma_bf16(d_tmem, a_desc + 2u, b_desc + 2u, idesc, 1u);
          if (KSTEPS > 2) {
            umma_bf16(d_tmem, a_desc + 4u, b_desc + 4u, idesc, 1u);
            umma_bf16(d_tmem, a_desc + 6u, b_desc + 6u, idesc, 1u);
          }
          umma_commit(empty_bar);
          if (kb == last_kb) umma_commit(bar_accf0 + 8u * acc);
        }
        __syncwarp();
        a_lo += stage_lo;
        full_bar += 8u;
        empty_bar += 8u;
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1u;
          a_lo = a_lo0;
          full_bar = bar_full0;
          empty_bar = bar_empty0;
        }
      }
      acc ^= 1u;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int q = warp & 3;                // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;      // which half of the tile's columns this warp owns
    const int row = q * 32 + lane;
    const int ly = row / p.bw, lx = row - ly * p.bw;
    // column split between the two warps of a quarter (needs halves that are multiples of 16)
    const int nsplit = (p.bn % 32 == 0) ? 2 : 1;
    const int cb = nsplit == 2 ? half * (p.bn >> 1) : 0;
    const int ce = nsplit == 2 ? cb + (p.bn >> 1) : (half == 0 ? p.bn : 0);
    const float* sbias = bias_smem ? s_bias : nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int ty = t.y0 + ly, tx = t.x0 + lx;
      const bool valid = ty < p.th && tx < p.tw;
      const int oy = ty * p.osy + p.ooy, ox = tx * p.osx + p.oox;
      const int64_t pix = ((int64_t)t.img * p.e.out_h + oy) * p.e.out_w + ox;
      mbar_wait(bar_accf0 + 8u * (uint32_t)acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
      if (p.e.fast == 1) {
        if (p.e.residual)
          epilogue_pixel_fast<true>(p.e, sbias, s_gamma, taddr, cb, ce, t.n0, valid, pix, &s_ss[acc][0][0], row, half,
                                    nsplit, 1 + q);
        else
          epilogue_pixel_fast<false>(p.e, sbias, s_gamma, taddr, cb, ce, t.n0, valid, pix, &s_ss[acc][0][0], row, half,
                                     nsplit, 1 + q);
      } else {
        epilogue_pixel(p.e, taddr, cb, ce, t.n0, valid, t.img, oy, ox, pix);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce0 + 8u * (uint32_t)acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2).  Why: with cta_group::1 and SS operands every M = 128 MMA reads 4 KB of A plus
// 32*N bytes of B from shared memory per K = 16 step, and the measured MMA time tracks (4096 + 32 N) / 64 cycles
// for every N -- the SM's operand read path, not the tensor pipe, is the ceiling (67 % of peak at N = 256, 43 % at
// N = 96).  A pair of CTAs on adjacent M tiles issues ONE M = 256 MMA: each SM still reads its own 128 rows of A but
// only HALF of B (the pair's B tile is split between the two shared memories), and the weight traffic through L2
// halves as well.  Roles per CTA are as in conv_tc_kernel; differences:
//   * both CTAs' TMA loads complete on the LEADER's full barrier (arrival count 1: the leader's producer posts the
//     expected bytes of both CTAs);
//   * only the leader's MMA warp issues; tcgen05.commit multicasts the arrive to the same barrier in both CTAs
//     (stage-empty and accumulator-full);
//   * the peer's epilogue warps release the accumulator on the leader's barrier through a remote mbarrier arrive;
//   * TMEM is allocated / freed with the cta_group::2 forms; cluster barriers fence set-up and tear-down.
// ---------------------------------------------------------------------------------------
// RS: 0, or the row-statistic mode (1 / 2) of rv_gemm_rowstat's lean epilogue, or 3: the GroupNorm statistics of the output
// (rv_conv2d_tc_gnstats) -- their own instantiations, so that the 64 registers of prefetched multiplicand / the group sums
// never weigh on the convolutions' epilogue
template <int KSTEPS, int RS = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[TC_SMEM_BIAS];
  __shared__ __align__(16) float s_gamma[256];
  __shared__ float s_ss[2][2][128];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int m_pairs = (p.m_tiles + 1) >> 1;
  const int total_pt = m_pairs * p.n_tiles;  // pair tiles: two adjacent M tiles x one N tile
  const int npairs = (int)gridDim.x >> 1, pair0 = (int)blockIdx.x >> 1;
  const int num_kb = p.ntaps * p.kc_per_tap;
  const int bn_half = p.bn >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bar_acc_full[a]), 1);
      mbar_init(smem_u32(&bar_acc_empty[a]), 2 * TC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
  }
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  const bool bias_smem = p.e.fast && p.e.bias_mode == 1;
  if (bias_smem)
    for (int i = threadIdx.x; i < p.e.cout; i += TC_THREADS) s_bias[i] = p.e.bias[i];
  if (p.e.norm_gamma)
    for (int i = threadIdx.x; i < p.e.cout; i += TC_THREADS) s_gamma[i] = p.e.norm_gamma[i] * (p.e.norm_silu ? 0.5f : 1.0f);  // SiLU's 1/2 rides in gamma
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const uint32_t bar_full0 = smem_u32(&bar_full[0]);
  const uint32_t bar_empty0 = smem_u32(&bar_empty[0]);
  const uint32_t bar_accf0 = smem_u32(&bar_acc_full[0]);
  const uint32_t bar_acce0 = smem_u32(&bar_acc_empty[0]);
  const uint32_t lead_full0 = mapa_rank(bar_full0, 0);   // the leader's barriers, as cluster addresses
  const uint32_t lead_acce0 = mapa_rank(bar_acce0, 0);

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    uint32_t stage = 0, phase = 0;
    const uint32_t nstages = (uint32_t)p.stages;
    for (int pt = pair0; pt < total_pt; pt += npairs) {
      const int mp = pt / p.n_tiles, nt = pt - mp * p.n_tiles;
      const TileCoord t = decode_tile(p, (2 * mp + (int)rank) * p.n_tiles + nt);
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const int ax = t.x0 + p.tap_dx[tap], ay = t.y0 + p.tap_dy[tap];
        const int kcol = p.w_k_base + tap * p.tap_k;
        const int c5 = p.tap_wp[tap] * p.cin_pitch, hp = p.tap_hp[tap];
        for (int kc = 0; kc < p.kc_per_tap; ++kc) {
          mbar_wait(bar_empty0 + 8u * stage, phase ^ 1u);
          if (elect_one()) {
            const uint32_t full = lead_full0 + 8u * stage;
            const uint32_t a_dst = smem_base + stage * p.stage_bytes;
            if (leader) mbar_arrive_expect_tx(bar_full0 + 8u * stage, 2u * p.tx_bytes);
            if (p.mode == 0) tma2_load_4d(a_dst, &map_a, full, kc * p.bk, ax, ay, t.img);
            else tma2_load_5d(a_dst, &map_a, full, c5 + kc * p.bk, ax, hp, ay, t.img);
            tma2_load_2d(a_dst + p.a_bytes, &map_b, full, kcol + kc * p.bk, t.n0 + (int)rank * bn_half);
          }
          __syncwarp();
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (leader) {
      // M = 256 across the pair: m_dim field = 256 >> 4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((256u >> 4) << 24);
      const uint64_t desc_hi = make_smem_desc(0u, p.sbo_bytes, p.layout_type);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      const uint32_t nstages = (uint32_t)p.stages;
      const uint32_t a_lo0 = (smem_base & 0x3FFFFu) >> 4, stage_lo = p.stage_bytes >> 4, ab_lo = p.a_bytes >> 4;
      uint32_t a_lo = a_lo0, full_bar = bar_full0, empty_bar = bar_empty0;
      const int last_kb = num_kb - 1;
      for (int pt = pair0; pt < total_pt; pt += npairs) {
        mbar_wait(bar_acce0 + 8u * acc, acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256u;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t a_desc = desc_hi | (uint64_t)a_lo;
            const uint64_t b_desc = desc_hi | (uint64_t)(a_lo + ab_lo);
            umma2_bf16(d_tmem, a_desc, b_desc, idesc, (uint32_t)(kb != 0));
            if (KSTEPS > 1) umma2_bf16(d_tmem, a_desc + 2u, b_desc + 2u, idesc, 1u);
            if (KSTEPS > 2) {
              umma2_bf16(d_tmem, a_desc + 4u, b_desc + 4u, idesc, 1u);
              umma2_bf16(d_tmem, a_desc + 6u, b_desc + 6u, idesc, 1u);
            }
            umma2_commit_both(empty_bar);
            if (kb == last_kb) umma2_commit_both(bar_accf0 + 8u * acc);
          }
          __syncwarp();
          a_lo += stage_lo;
          full_bar += 8u;
          empty_bar += 8u;
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1u;
            a_lo = a_lo0;
            full_bar = bar_full0;
            empty_bar = bar_empty0;
          }
        }
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------ epilogue (both CTAs, own 128 rows) ------------------------------
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int ly = row / p.bw, lx = row - ly * p.bw;
    const int nsplit = (p.bn % 32 == 0) ? 2 : 1;
    const int cb = nsplit == 2 ? half * (p.bn >> 1) : 0;
    const int ce = nsplit == 2 ? cb + (p.bn >> 1) : (half == 0 ? p.bn : 0);
    const float* sbias = bias_smem ? s_bias : nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int pt = pair0; pt < total_pt; pt += npairs) {
      const int mp = pt / p.n_tiles, nt = pt - mp * p.n_tiles;
      const int m_tile = 2 * mp + (int)rank;
      const TileCoord t = decode_tile(p, m_tile * p.n_tiles + nt);
      const int ty = t.y0 + ly, tx = t.x0 + lx;
      const bool valid = m_tile < p.m_tiles && ty < p.th && tx < p.tw;
      const int oy = ty * p.osy + p.ooy, ox = tx * p.osx + p.oox;
      const int64_t pix = ((int64_t)t.img * p.e.out_h + oy) * p.e.out_w + ox;
      RowstatPrefetch pf;
      if (RS == 1 || RS == 2) rowstat_prefetch<(RS == 2 ? 2 : 1)>(p.e, cb, ce, t.n0, valid, pix, pf);
      mbar_wait(bar_accf0 + 8u * (uint32_t)acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
      if (RS == 1 || RS == 2) {
        epilogue_pixel_rowstat<(RS == 2 ? 2 : 1)>(p.e, taddr, cb, ce, t.n0, valid, pix, pf);
      } else if (RS == 3) {
        // this warp's 32 floats of this tile: [tile][half * 4 + q][lane]
        float* part = p.e.gn_part + ((int64_t)(m_tile * p.n_tiles + nt) * 8 + half * 4 + q) * 32;
        const bool res = p.e.residual != nullptr;
        if (m_tile < p.m_tiles) {
          if (p.e.gn_gs == 8) {
            if (res) epilogue_pixel_gnstats<8, 4, true>(p.e, sbias, taddr, cb, t.n0, valid, pix, part, lane);
            else epilogue_pixel_gnstats<8, 4, false>(p.e, sbias, taddr, cb, t.n0, valid, pix, part, lane);
          } else {
            if (res) epilogue_pixel_gnstats<16, 4, true>(p.e, sbias, taddr, cb, t.n0, valid, pix, part, lane);
            else epilogue_pixel_gnstats<16, 4, false>(p.e, sbias, taddr, cb, t.n0, valid, pix, part, lane);
          }
        }
      } else if (p.e.fast == 1) {
        if (p.e.residual)
          epilogue_pixel_fast<true>(p.e, sbias, s_gamma, taddr, cb, ce, t.n0, valid, pix, &s_ss[acc][0][0], row, half, nsplit,
                                    1 + q);
        else
          epilogue_pixel_fast<false>(p.e, sbias, s_gamma, taddr, cb, ce, t.n0, valid, pix, &s_ss[acc][0][0], row, half, nsplit,
                                     1 + q);
      } else {
        epilogue_pixel(p.e, taddr, cb, ce, t.n0, valid, t.img, oy, ox, pix);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(lead_acce0 + 8u * (uint32_t)acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody frees TMEM or exits while the peer may still signal / read
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// Host side: tensor maps, tiling choice, launch
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::mutex g_init_mu;
static bool g_attr_set[64] = {false};

int tc_ensure_init() {
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    RV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    RV_CHECK_ARG(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = (EncodeTiledFn)fn;
  }
  int dev = 0;
  RV_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !g_attr_set[dev]) {
    cudaDeviceProp prop;
    RV_CUDA(cudaGetDeviceProperties(&prop, dev));
    RV_CHECK_ARG(prop.major == 10, "librgbavae needs an sm_100 GPU (found sm_%d%d); there is no fallback path", prop.major,
                 prop.minor);
    RV_CUDA(cudaFuncSetAttribute(conv_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TC_SMEM_BUDGET + 2048)));
    RV_CUDA(cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TC_SMEM_BUDGET + 2048)));
    RV_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TC_SMEM_BUDGET + 2048)));
    RV_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TC_SMEM_BUDGET + 2048)));
    RV_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TC_SMEM_BUDGET + 2048)));
    RV_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TC_SMEM_BUDGET + 2048)));
    RV_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TC_SMEM_BUDGET + 2048)));
    RV_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TC_SMEM_BUDGET + 2048)));
    g_attr_set[dev] = true;
  }
  return 0;
}

int tc_encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box, CUtensorMapSwizzle sw) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d, rank %d, dims %llu,%llu,%llu box %u,%u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0), box[0],
              box[1], rank > 2 ? box[2] : 0);
    return 1;
  }
  return 0;
}

// Tile shape (bw x bh = 128) wasting the fewest computed-but-unused pixels; ties go to the wider tile.
static void choose_tile(int th, int tw, int* bw_out, int* bh_out) {
  int64_t best = -1;
  for (int bw = 128; bw >= 8; bw >>= 1) {
    int bh = 128 / bw;
    int64_t area = (int64_t)((tw + bw - 1) / bw) * bw * ((th + bh - 1) / bh) * bh;
    if (best < 0 || area < best) {
      best = area;
      *bw_out = bw;
      *bh_out = bh;
    }
  }
}

static void choose_bn(int cout, int* bn_out, int* n_tiles_out) {
  int c16 = (cout + 15) / 16 * 16;
  if (c16 <= 256) {
    *bn_out = c16;
    *n_tiles_out = 1;
    return;
  }
  int t0 = (c16 + 255) / 256;
  int best_bn = 256, best_t = t0, best_waste = 1 << 30;
  for (int t = t0; t <= t0 + 4; ++t) {
    int bn = ((c16 + t - 1) / t + 15) / 16 * 16;
    int waste = bn * t - c16;
    if (waste < best_waste) {
      best_waste = waste;
      best_bn = bn;
      best_t = t;
    }
  }
  *bn_out = best_bn;
  *n_tiles_out = best_t;
}

void fill_epi(EpiParams* e, const rv_conv_desc* d, const float* bias, const void* residual, void* y, const NormFuse* nf) {
  e->out_h = d->oh;
  e->out_w = d->ow;
  e->cout = d->cout;
  e->y_cstride = d->y_cstride;
  e->y_nchw = d->y_nchw;
  e->y_f32 = d->y_dtype == RV_F32;
  e->bias_mode = d->bias_mode;
  e->clamp = d->clamp;
  e->alpha = d->alpha;
  e->out_scale = d->out_scale;
  e->out_shift = d->out_shift;
  e->clamp_lo = d->clamp_lo;
  e->clamp_hi = d->clamp_hi;
  e->bias = bias;
  e->residual = (const __nv_bfloat16*)residual;
  e->y = y;
  e->norm_gamma = nf ? nf->gamma_scaled : nullptr;
  e->y_act = nf ? (__nv_bfloat16*)nf->y_act : nullptr;
  e->norm_silu = nf ? nf->silu : 0;
  e->fast = 0;  // decided by the launcher once the N tiling is known
  e->rowstat = nullptr;
  e->rowstat_mode = 0;
  e->gn_part = nullptr;
  e->gn_gs = 0;
  e->vec_ok = (d->y_cstride % 8 == 0) && (!y || (uintptr_t)y % 16 == 0) && (!residual || (uintptr_t)residual % 16 == 0) &&
              (!nf || (uintptr_t)nf->y_act % 16 == 0);
}

// preconditions of the lean epilogue (epilogue_pixel_fast)
bool epi_fast_ok(const EpiParams& e, const rv_conv_desc* d, const float* bias, const NormFuse* nf, int covered_cols) {
  return (d->bias_mode == 0 || d->cout <= TC_SMEM_BIAS) && e.vec_ok && !d->y_nchw && d->y_dtype == RV_BF16 &&
         d->cout % 16 == 0 && covered_cols == d->cout && d->bias_mode != 2 && d->alpha == 1.0f && d->out_scale == 1.0f &&
         d->out_shift == 0.0f && !d->clamp && (d->bias_mode == 0 || (uintptr_t)bias % 16 == 0) &&
         (!nf || (uintptr_t)nf->gamma_scaled % 16 == 0);
}

bool halo_eligible(const rv_conv_desc* d, const EpiParams& e);  // rv_conv_halo.cu
int launch_halo(const rv_conv_desc* d, const void* x, const void* w, int64_t w_ld, const float* bias, const void* residual,
                void* y, cudaStream_t st, const NormFuse* nf);

static int launch_tc(const rv_conv_desc* d, const void* x, const void* w, int64_t w_ld, const float* bias,
                     const void* residual, void* y, cudaStream_t st, int phase /* -1: not upsample */,
                     const NormFuse* nf = nullptr, const RowStat* rs = nullptr, float* gn_part = nullptr, int gn_gs = 0) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = d->n;
  p.cin = d->cin;
  p.cin_pitch = d->x_cstride;
  // 128-byte operand rows move ~2x faster through TMA than 64-byte ones (measured: the row rate, not the
  // byte rate, bounds small-N layers), so channel counts above 64 always use BK=64: the last block of a tap
  // is partly out of bounds in x (zero-filled), which cancels whatever weight columns it is paired with.
  // (not for the stride-2 parity view, whose folded channel axis has real data past cin)
  // The stride-2 parity view has real data past cin on its folded channel axis (the odd column's pixel), so a 64-wide block
  // that overhangs cin needs ZERO WEIGHTS under the overhang instead: the caller packs such a conv (Qwen's 96 -> 96
  // down-sampler) with every tap's K range padded to a multiple of 64 (w_ld = 9 * 128), recognised here by the row pitch.
  int tap_k = d->cin;
  if (phase < 0 && d->stride == 2 && d->ksize == 3 && !d->taps_1d && w_ld % 9 == 0 && (w_ld / 9) % 64 == 0 && w_ld / 9 >= d->cin &&
      w_ld / 9 < d->cin + 64)
    tap_k = (int)(w_ld / 9);
  p.bk = (d->cin % 64 == 0 || (d->cin > 64 && (d->stride == 1 || tap_k % 64 == 0))) ? 64 : (d->cin % 32 == 0 ? 32 : 16);
  p.kc_per_tap = (d->cin + p.bk - 1) / p.bk;
  p.tap_k = tap_k;
  const CUtensorMapSwizzle sw =
      p.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  p.layout_type = p.bk == 64 ? 2u : (p.bk == 32 ? 4u : 6u);
  p.sbo_bytes = 8u * (uint32_t)p.bk * 2u;
  p.osy = p.osx = 1;
  int k_extent;
  if (phase >= 0) {
    // nearest x2 + 3x3 as four 2x2 phase convolutions on the source grid
    const int py = phase >> 1, px = phase & 1;
    p.th = d->h;
    p.tw = d->w;
    p.ntaps = 4;
    for (int t = 0; t < 4; ++t) {
      p.tap_dy[t] = (signed char)((t >> 1) - 1 + py);
      p.tap_dx[t] = (signed char)((t & 1) - 1 + px);
    }
    p.w_k_base = phase * 4 * d->cin;
    p.osy = p.osx = 2;
    p.ooy = py;
    p.oox = px;
    k_extent = 16 * d->cin;
  } else {
    p.th = d->oh;
    p.tw = d->ow;
    p.ntaps = d->taps_1d ? d->ksize : d->ksize * d->ksize;
    for (int t = 0; t < p.ntaps; ++t) {
      int dy = t / d->ksize, dx = t % d->ksize;
      if (d->taps_1d) {
        p.tap_dy[t] = (signed char)(t - d->pad_lo);
        p.tap_dx[t] = 0;
      } else if (d->stride == 2) {
        p.tap_dy[t] = (signed char)(dy >> 1);
        p.tap_hp[t] = (signed char)(dy & 1);
        p.tap_dx[t] = (signed char)(dx >> 1);
        p.tap_wp[t] = (signed char)(dx & 1);
      } else {
        p.tap_dy[t] = (signed char)(dy - d->pad_lo);
        p.tap_dx[t] = (signed char)(dx - d->pad_lo);
      }
    }
    k_extent = p.ntaps * tap_k;
  }
  p.mode = d->stride == 2 ? 1 : 0;
  choose_tile(p.th, p.tw, &p.bw, &p.bh);
  p.tiles_x = (p.tw + p.bw - 1) / p.bw;
  p.tiles_y = (p.th + p.bh - 1) / p.bh;
  p.m_tiles = d->n * p.tiles_x * p.tiles_y;
  choose_bn(d->cout, &p.bn, &p.n_tiles);
  RV_CHECK_ARG(!nf || p.n_tiles == 1, "conv_tc: fused norm needs all %d output channels in one tile (<= 256)", d->cout);
  fill_epi(&p.e, d, bias, residual, y, nf);
  p.e.fast = epi_fast_ok(p.e, d, bias, nf, p.bn * p.n_tiles);
  if (rs) {  // per-row statistic: generic epilogue only
    p.e.rowstat = rs->stat;
    p.e.rowstat_mode = rs->mode;
    // lean form: whole 32-column steps per epilogue warp, 32-byte aligned bf16 rows, nothing but alpha and the statistic
    const bool lean = p.e.vec_ok && !d->y_nchw && d->y_dtype == RV_BF16 && d->bias_mode == 0 && d->out_scale == 1.0f &&
                      d->out_shift == 0.0f && !d->clamp && p.bn % 64 == 0 && p.bn * p.n_tiles == d->cout &&
                      d->y_cstride % 16 == 0 && (uintptr_t)y % 32 == 0 && (!residual || (uintptr_t)residual % 32 == 0);
    p.e.fast = lean && p.bk == 64 ? 2 : 0;  // (confirmed below: the lean form lives in the CTA-pair kernel only)
  }
  RV_CHECK_ARG(!nf || p.e.fast, "conv_tc: fused norm needs cout %% 16 == 0, aligned NHWC bf16 tensors, per-channel bias, no affine");
  // CTA pairs (cta_group::2) whenever the B tile splits into two swizzle-aligned halves
  static const bool no_pair = getenv("RGBAVAE_DISABLE_PAIR") != nullptr;
  const bool pair = !no_pair && p.bn % 32 == 0 && p.bk >= 32 && p.m_tiles >= 2;
  p.a_bytes = 128u * (uint32_t)p.bk * 2u;
  const uint32_t b_rows = pair ? (uint32_t)p.bn / 2u : (uint32_t)p.bn;
  const uint32_t b_bytes = b_rows * (uint32_t)p.bk * 2u;
  p.tx_bytes = p.a_bytes + b_bytes;
  p.stage_bytes = (p.a_bytes + b_bytes + 1023u) & ~1023u;
  if (p.a_bytes % 1024u) {  // bk == 16: keep B 1024-aligned too
    p.a_bytes = (p.a_bytes + 1023u) & ~1023u;
    p.stage_bytes = (p.a_bytes + b_bytes + 1023u) & ~1023u;
  }
  p.stages = (int)(TC_SMEM_BUDGET / p.stage_bytes);
  if (p.stages > TC_MAX_STAGES) p.stages = TC_MAX_STAGES;
  RV_CHECK_ARG(p.stages >= 2, "conv_tc: tile does not fit shared memory");

  CUtensorMap map_a, map_b;
  const uint64_t pitch_b = (uint64_t)d->x_cstride * 2u;
  if (p.mode == 0) {
    cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
    cuuint64_t str[3] = {pitch_b, pitch_b * d->w, pitch_b * d->w * d->h};
    cuuint32_t box[4] = {(cuuint32_t)p.bk, (cuuint32_t)p.bw, (cuuint32_t)p.bh, 1};
    if (int rc = tc_encode_map(&map_a, x, 4, dims, str, box, sw)) return rc;
  } else {
    RV_CHECK_ARG(d->h % 2 == 0 && d->w % 2 == 0, "conv_tc: stride-2 conv needs even input size");
    cuuint64_t dims[5] = {(cuuint64_t)(2 * d->x_cstride), (cuuint64_t)(d->w / 2), 2, (cuuint64_t)(d->h / 2), (cuuint64_t)d->n};
    cuuint64_t str[4] = {2 * pitch_b, pitch_b * d->w, 2 * pitch_b * d->w, pitch_b * d->w * d->h};
    cuuint32_t box[5] = {(cuuint32_t)p.bk, (cuuint32_t)p.bw, 1, (cuuint32_t)p.bh, 1};
    if (int rc = tc_encode_map(&map_a, x, 5, dims, str, box, sw)) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)k_extent, (cuuint64_t)d->cout};
    cuuint64_t str[1] = {(cuuint64_t)w_ld * 2u};
    cuuint32_t box[2] = {(cuuint32_t)p.bk, b_rows};
    if (int rc = tc_encode_map(&map_b, w, 2, dims, str, box, sw)) return rc;
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  const double flops = 2.0 * (double)d->n * d->oh * d->ow * d->cout * d->cin * d->ksize * (d->taps_1d ? 1 : d->ksize) / (phase >= 0 ? 4.0 : 1.0);
  LaunchScope scope(phase >= 0 ? CAT_CONV_UPS : CAT_CONV_TC, st, flops);
  if (!pair && p.e.fast == 2) p.e.fast = 0;
  if (gn_part) {
    RV_CHECK_ARG(pair && p.bk == 64 && p.e.fast == 1 && !nf && !rs, "conv_tc gnstats: layer not eligible");
    p.e.gn_part = gn_part;
    p.e.gn_gs = gn_gs;
  }
  if (pair) {
    const int total_pt = ((p.m_tiles + 1) / 2) * p.n_tiles;
    const int max_pairs = num_sms() / 2;
    const int grid = 2 * (total_pt < max_pairs ? total_pt : max_pairs);
    if (gn_part) conv_tc2_kernel<4, 3><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, p);
    else if (p.e.fast == 2 && p.e.rowstat_mode == 1) conv_tc2_kernel<4, 1><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, p);
    else if (p.e.fast == 2) conv_tc2_kernel<4, 2><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, p);
    else if (p.bk == 64) conv_tc2_kernel<4><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, p);
    else conv_tc2_kernel<2><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, p);
  } else {
    const int total_tiles = p.m_tiles * p.n_tiles;
    const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
    if (p.bk == 64) conv_tc_kernel<4><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, p);
    else if (p.bk == 32) conv_tc_kernel<2><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, p);
    else conv_tc_kernel<1><<<grid, TC_THREADS, smem, st>>>(map_a, map_b, p);
  }
  RV_LAUNCH_CHECK();
  return 0;
}

// Sums the per-tile partials conv_tc2_kernel<4, 3> left (epilogue_pixel_gnstats) into the [n][groups][2] fp64 (sum, sum of
// squares) rv_groupnorm_silu reads.  Block = (group, sample); fixed summation order.
__global__ void __launch_bounds__(512) gn_finish_kernel(const float* __restrict__ part, double* __restrict__ stats, int tpi,
                                                       int n_tiles, int bn, int gs, int lane_step, int groups, int nseg,
                                                       int64_t seg_stride) {
  __shared__ double red[2][16];
  const int G = blockIdx.x, n = blockIdx.y;
  const int ch0 = G * gs;
  const int nt = ch0 / bn, col = ch0 - nt * bn;
  const int half = col / (bn / 2);
  const int g_local = (col - half * (bn / 2)) / gs;
  const int lane_s = 2 * g_local * lane_step, lane_q = (2 * g_local + 1) * lane_step;
  // 512 threads x 4 independent loads in flight each: the (tile, lane quarter) partials of one (sample, group) are 128-byte
  // strided 8-byte pairs -- a latency-bound gather (128 threads walking it one load at a time took 26 us per conv)
  double sa = 0.0, sq = 0.0;
  const int items = tpi * 4;
  for (int seg = 0; seg < nseg; ++seg) {  // the four phase launches of an up-sampling conv leave a segment each
    const float* base = part + seg * seg_stride + ((int64_t)n * tpi * n_tiles + nt) * 8 * 32 + half * 4 * 32;
    for (int i0 = threadIdx.x; i0 < items; i0 += 512 * 4) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 512;
        const bool ok = i < items;
        const float* w = base + ((int64_t)(i >> 2) * n_tiles * 8 + (i & 3)) * 32;
        a[u] = ok ? w[lane_s] : 0.f;
        b[u] = ok ? w[lane_q] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        sa += (double)a[u];
        sq += (double)b[u];
      }
    }
  }
  for (int o = 16; o >= 1; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = sa;
    red[1][threadIdx.x >> 5] = sq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tq = 0.0;
    for (int k = 0; k < 16; ++k) {
      ta += red[0][k];
      tq += red[1][k];
    }
    stats[((int64_t)n * groups + G) * 2] = ta;
    stats[((int64_t)n * groups + G) * 2 + 1] = tq;
  }
}

// tiling of a layer as launch_tc would choose it, and whether the statistics epilogue covers it; 0 or the scratch bytes
static int64_t gnstats_geometry(const rv_conv_desc* d, int groups, int* tpi, int* n_tiles, int* bn, int* gs, int* lane_step,
                                int64_t* seg_floats) {
  if (groups != 32 || d->y_nchw || d->y_dtype != RV_BF16 || d->x_dtype != RV_BF16 || d->cout % groups) return 0;
  if (d->upsample && (d->ksize != 3 || d->stride != 1 || d->cin % 64)) return 0;
  if (d->taps_1d || d->bias_mode != 1 || d->alpha != 1.0f || d->out_scale != 1.0f || d->out_shift != 0.0f || d->clamp) return 0;
  const int bk = (d->cin % 64 == 0 || (d->cin > 64 && d->stride == 1)) ? 64 : 0;
  if (bk != 64 || d->cout % 16 || d->y_cstride != d->cout) return 0;
  int bw, bh, b, t;
  const int gh = d->upsample ? d->h : d->oh, gw = d->upsample ? d->w : d->ow;  // the tiled grid (source grid for the phase convs)
  choose_tile(gh, gw, &bw, &bh);
  const int tx = (gw + bw - 1) / bw, ty = (gh + bh - 1) / bh;
  choose_bn(d->cout, &b, &t);
  if (b * t != d->cout || b % 64 || d->n * tx * ty < 2) return 0;
  const int g = d->cout / groups;
  const int nst = b / 2 / 32;  // 32-column steps per epilogue warp
  // 256 / 512 output channels.  (128: measured a small net loss -- N = 128 layers are paced by their epilogue, 256 -> 128 @1024^2
  // 2.44 -> 2.73 ms for 0.23 ms of statistics pass saved -- so those keep the stand-alone pass.)
  if (!((g == 8 && nst == 4) || (g == 16 && nst == 4))) return 0;
  {  // the halo kernel takes the narrow full-resolution layers: its epilogue has no statistics form
    EpiParams e;
    fill_epi(&e, d, (const float*)16, nullptr, (void*)16, nullptr);
    e.fast = 1;
    static const bool no_halo = getenv("RGBAVAE_DISABLE_HALO") != nullptr;
    if (!no_halo && halo_eligible(d, e)) return 0;
  }
  *tpi = tx * ty;
  *n_tiles = t;
  *bn = b;
  *gs = g;
  *lane_step = 32 / (2 * nst * (32 / g));
  const int64_t m_tiles = (int64_t)d->n * tx * ty;
  *seg_floats = ((m_tiles + 1) / 2 * 2) * t * 8 * 32;
  return *seg_floats * (d->upsample ? 4 : 1) * (int64_t)sizeof(float);
}

// fp32 [cout][cin][k][k] -> bf16 [cout][taps][cin]  (upsample: [cout][4 phases][4 taps][cin], folded)
__global__ void pack_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin, int ksize,
                               int upsample) {
  const int taps = ksize * ksize;
  const int slots = upsample ? 16 : taps;
  const int64_t total = (int64_t)cout * slots * cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cin);
    const int64_t r = i / cin;
    const int s = (int)(r % slots);
    const int co = (int)(r / slots);
    const float* src = w + ((int64_t)co * cin + c) * taps;
    float v;
    if (!upsample) {
      v = src[s];
    } else {
      const int phase = s >> 2, t = s & 3;
      const int py = phase >> 1, px = phase & 1, ty = t >> 1, tx = t & 1;
      // rows of the 3x3 kernel that land on source row (ty - 1 + py) for output parity py
      const int y_lo = py == 0 ? (ty == 0 ? 0 : 1) : (ty == 0 ? 0 : 2);
      const int y_hi = py == 0 ? (ty == 0 ? 0 : 2) : (ty == 0 ? 1 : 2);
      const int x_lo = px == 0 ? (tx == 0 ? 0 : 1) : (tx == 0 ? 0 : 2);
      const int x_hi = px == 0 ? (tx == 0 ? 0 : 2) : (tx == 0 ? 1 : 2);
      v = 0.f;
      for (int dy = y_lo; dy <= y_hi; ++dy)
        for (int dx = x_lo; dx <= x_hi; ++dx) v += src[dy * 3 + dx];
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace rv

extern "C" {

int rv_init(void) { return rv::tc_ensure_init(); }

static int conv2d_tc_impl(const rv_conv_desc* d, const void* x, const void* w_packed, int64_t w_ld, const float* bias,
                          const void* residual, void* y, void* stream, const rv::NormFuse* nf, const rv::RowStat* rs = nullptr) {
  if (int rc = rv::check_conv_desc(d)) return rc;
  if (int rc = rv::tc_ensure_init()) return rc;
  RV_CHECK_ARG(x && w_packed && (y || nf), "conv_tc: null tensor");
  RV_CHECK_ARG(d->x_dtype == RV_BF16 && !d->x_nchw, "conv_tc: x must be NHWC bf16");
  RV_CHECK_ARG(d->cin % 16 == 0, "conv_tc: cin (%d) must be a multiple of 16", d->cin);
  RV_CHECK_ARG(d->x_cstride % 8 == 0 && w_ld % 8 == 0, "conv_tc: x_cstride and w_ld must be multiples of 8");
  RV_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)w_packed % 16 == 0), "conv_tc: x / w must be 16-byte aligned");
  RV_CHECK_ARG(d->in_scale == 1.0f && d->in_shift == 0.0f, "conv_tc: in_scale/in_shift are direct-path only");
  RV_CHECK_ARG(d->bias_mode == 0 || bias, "conv_tc: bias_mode set but bias is null");
  RV_CHECK_ARG(!residual || (!d->y_nchw && d->y_dtype == RV_BF16), "conv_tc: residual needs an NHWC bf16 output");
  RV_CHECK_ARG(d->bias_mode != 1 || (uintptr_t)bias % 16 == 0, "conv_tc: bias must be 16-byte aligned");
  RV_CHECK_ARG(!(d->upsample && d->ksize != 3), "conv_tc: upsample requires ksize 3");
  RV_CHECK_ARG(!nf || (nf->gamma_scaled && nf->y_act && !d->y_nchw && d->y_dtype == RV_BF16),
               "conv_tc: fused norm needs gamma, y_act and an NHWC bf16 layout");
  cudaStream_t st = (cudaStream_t)stream;
  if (rs) {
    RV_CHECK_ARG(d->ksize == 1 && d->stride == 1 && !d->upsample && !nf && !d->y_nchw, "gemm_rowstat: plain GEMMs (1x1, NHWC) only");
    RV_CHECK_ARG(rs->stat && (rs->mode == 1 || rs->mode == 2), "gemm_rowstat: mode must be 1 (exp2) or 2 (multiply) with a statistic");
    RV_CHECK_ARG((rs->mode == 2) == (residual != nullptr), "gemm_rowstat: mode 2 needs the multiplicand, mode 1 takes none");
    return rv::launch_tc(d, x, w_packed, w_ld, bias, residual, y, st, -1, nullptr, rs);
  }
  if (d->upsample) {
    for (int phase = 0; phase < 4; ++phase)
      if (int rc = rv::launch_tc(d, x, w_packed, w_ld, bias, residual, y, st, phase, nf)) return rc;
    return 0;
  }
  {
    // narrow full-resolution 3x3 layers: halo-reuse kernel (rv_conv_halo.cu)
    static const bool no_halo = getenv("RGBAVAE_DISABLE_HALO") != nullptr;
    rv::EpiParams e;
    rv::fill_epi(&e, d, bias, residual, y, nf);
    e.fast = rv::epi_fast_ok(e, d, bias, nf, d->cout);
    if (!no_halo && rv::halo_eligible(d, e)) return rv::launch_halo(d, x, w_packed, w_ld, bias, residual, y, st, nf);
  }
  return rv::launch_tc(d, x, w_packed, w_ld, bias, residual, y, st, -1, nf);
}

int rv_conv2d_tc(const rv_conv_desc* d, const void* x, const void* w_packed, int64_t w_ld, const float* bias,
                 const void* residual, void* y, void* stream) {
  return conv2d_tc_impl(d, x, w_packed, w_ld, bias, residual, y, stream, nullptr);
}

int rv_conv2d_tc_norm(const rv_conv_desc* d, const void* x, const void* w_packed, int64_t w_ld, const float* bias,
                      const void* residual, void* y, void* y_act, const float* gamma_scaled, int apply_silu, void* stream) {
  rv::NormFuse nf{gamma_scaled, y_act, apply_silu};
  return conv2d_tc_impl(d, x, w_packed, w_ld, bias, residual, y, stream, &nf);
}

int64_t rv_conv2d_tc_gnstats_scratch_bytes(const rv_conv_desc* d, int groups) {
  if (!d || rv::check_conv_desc(d)) return 0;
  int tpi, n_tiles, bn, gs, lane_step;
  int64_t seg;
  return rv::gnstats_geometry(d, groups, &tpi, &n_tiles, &bn, &gs, &lane_step, &seg);
}

int rv_conv2d_tc_gnstats(const rv_conv_desc* d, const void* x, const void* w_packed, int64_t w_ld, const float* bias,
                         const void* residual, void* y, int groups, double* stats, void* scratch, int64_t scratch_bytes,
                         void* stream) {
  if (int rc = rv::check_conv_desc(d)) return rc;
  if (int rc = rv::tc_ensure_init()) return rc;
  int tpi, n_tiles, bn, gs, lane_step;
  int64_t seg = 0;
  const int64_t need = rv::gnstats_geometry(d, groups, &tpi, &n_tiles, &bn, &gs, &lane_step, &seg);
  RV_CHECK_ARG(need > 0, "conv2d_tc_gnstats: this layer has no statistics epilogue (rv_conv2d_tc_gnstats_scratch_bytes returned 0)");
  RV_CHECK_ARG(x && w_packed && y && bias && stats && scratch && scratch_bytes >= need && (uintptr_t)scratch % 128 == 0,
               "conv2d_tc_gnstats: null tensor or scratch smaller than %lld bytes", (long long)need);
  RV_CHECK_ARG(d->cin % 16 == 0 && d->x_cstride % 8 == 0 && w_ld % 8 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)w_packed % 16 == 0) &&
                   (uintptr_t)bias % 16 == 0 && d->in_scale == 1.0f && d->in_shift == 0.0f,
               "conv2d_tc_gnstats: operand alignment / layout as for rv_conv2d_tc");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->upsample) {  // nearest x2 + 3x3 as four phase convs on the source grid: a segment of partials per phase
    for (int phase = 0; phase < 4; ++phase)
      if (int rc = rv::launch_tc(d, x, w_packed, w_ld, bias, residual, y, st, phase, nullptr, nullptr, (float*)scratch + phase * seg, gs))
        return rc;
  } else if (int rc = rv::launch_tc(d, x, w_packed, w_ld, bias, residual, y, st, -1, nullptr, nullptr, (float*)scratch, gs)) {
    return rc;
  }
  rv::LaunchScope scope(rv::CAT_NORM, st, (double)need);
  rv::gn_finish_kernel<<<dim3((unsigned)groups, (unsigned)d->n), 512, 0, st>>>((const float*)scratch, stats, tpi, n_tiles, bn, gs, lane_step,
                                                                            groups, d->upsample ? 4 : 1, seg);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_gemm_rowstat(const rv_conv_desc* d, const void* x, const void* w, int64_t w_ld, const float* rowstat, int mode,
                    const void* mul_in, void* y, void* stream) {
  rv::RowStat rs{rowstat, mode};
  return conv2d_tc_impl(d, x, w, w_ld, nullptr, mul_in, y, stream, nullptr, &rs);
}

int rv_pack_conv_weights(const float* w, int cout, int cin, int ksize, int upsample, void* out_bf16, int64_t* w_ld,
                         void* stream) {
  RV_CHECK_ARG(w && out_bf16 && cout > 0 && cin > 0 && (ksize == 1 || ksize == 3), "pack: bad argument");
  RV_CHECK_ARG(!upsample || ksize == 3, "pack: upsample folding needs a 3x3 kernel");
  const int slots = upsample ? 16 : ksize * ksize;
  if (w_ld) *w_ld = (int64_t)slots * cin;
  const int64_t total = (int64_t)cout * slots * cin;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  rv::LaunchScope scope(rv::CAT_LAYOUT, (cudaStream_t)stream, 6.0 * total);
  rv::pack_tc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)out_bf16, cout, cin, ksize, upsample);
  RV_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
