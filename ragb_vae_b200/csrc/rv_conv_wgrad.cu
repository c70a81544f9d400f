// Weight gradient of the 3x3 (stride 1, pad 1) and 1x1 convolutions on the tensor cores (sm_100a):
//   dW[co][tap][ci] = sum over (n, y, x) of dY[n][y][x][co] * X[n][y+dy-pad][x+dx-pad][ci]
// i.e. per tap a GEMM D[co x ci] = A[co x P] * B[P x ci] whose reduction dimension is the PIXEL index.  In NHWC both
// operands have the reduction index as their slow dimension, so they are "MN-major" UMMA operands: a TMA box of
// (64 channels x KP pixels) lands in shared memory as KP rows of 128 swizzled bytes, which is exactly the canonical
// MN-major SWIZZLE_128B layout (64-element MN blocks LBO apart, 8-row K groups 1024 B apart).  The tap shift of X is a
// TMA coordinate offset; image borders are out-of-bounds zero fill.
//
// One CTA = (kernel row ty, 128-row Cout tile, Cin tile, pixel split).  The ksize taps of one kernel row differ only by a
// one-pixel shift of X, i.e. by one 128-byte row of the X box, so the CTA loads ONE dY box and ONE (KP + ksize - 1)-pixel
// X box per K block and issues the MMAs of all ksize taps from them (descriptor start address + tx * 128 B) into ksize
// TMEM accumulators: operand traffic per MMA is a third of the tap-per-CTA form.  It streams its image rows through a
// TMA ring and adds the tiles to dW with fp32 atomics (split-K over pixels; the bias gradient = column sums of dY rides along as
// one more N = 16 MMA against a tile of ones); dW is addressed through three strides so
// the gradient lands directly in the parameter's own layout.  Warps: 0 TMA, 1 / 6 / 7 MMA (one per tap), 2-5 epilogue.
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "rv_tc_common.cuh"

namespace rv {

constexpr int WG_THREADS = 288;        // 0 TMA, 1 / 6 / 7 MMA issuers (one per tap of the kernel row), 8 bias issuer, 2-5 epilogue
constexpr int WG_KP = 64;            // pixels per K block (4 MMAs of K = 16: the issue loop is written out)
static_assert(WG_KP == 64, "conv_wgrad_kernel issues exactly four K = 16 MMAs per block");
constexpr int WG_MAX_STAGES = 8;
constexpr uint32_t WG_A_BOX = WG_KP * 128;  // one (64 channels x KP pixels) dY box
constexpr uint32_t WG_B_BOX = 9 * 1024;     // one (64 channels x (KP + 2) pixels) X box, padded to the swizzle atom

struct WgParams {
  int n_img, h, w;
  int cin, cout, cin_valid, cout_valid, ksize, pad;
  int co_tiles, ci_tiles, bn, nboxes_b;  // bn: Cin columns per tile (multiple of 16, ksize * bn <= 512)
  int splits, rows_per_split, x_blocks;
  float* dw;      // element (co, ci, tap) at dw[co * s_co + ci * s_ci + tap * s_tap], accumulated
  float* dbias;   // optional [cout_valid], accumulated: the column sums of dY ride along as one extra N = 16 MMA per K block
                  // against a constant tile of ones (first kernel row / first Cin tile only) -- no second pass over dY
  uint32_t ones_off;
  long long s_co, s_ci, s_tap;
  int stages;
  uint32_t stage_bytes, tx_bytes, tmem_cols;
  int tapn;       // 1 (ksize 3, bn = 128): ONE N = 192 MMA covers the three taps of a 64-channel block of X -- the taps are the same
                  // box one pixel row (128 B) apart, so they are consecutive "MN blocks" of a descriptor whose LBO is 128 B.
                  // Two issuers (one per channel block) instead of three (one per tap): 8 MMAs reading 10 KB of operands each
                  // per K block instead of 12 reading 8 KB -- the shared-memory port is what paces these MMAs.
                  // Accumulator columns: [block][tap][64 channels].
};

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                  const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // unit decode
  // unit decode, pixel split SLOWEST: the CTAs that run together are the kernel rows / Cout / Cin tiles of the same image rows, so
  // each dY and X row comes from DRAM once and from L2 for the others (split-fastest order re-read both tensors: ncu 0.82 GB
  // of DRAM reads per 96-channel 512^2 launch against 0.40 GB algorithmic)
  int u = blockIdx.x;
  const int ty = u % p.ksize;
  u /= p.ksize;
  const int ci_t = u % p.ci_tiles;
  u /= p.ci_tiles;
  const int co_t = u % p.co_tiles;
  const int split = u / p.co_tiles;
  const int dy = ty - p.pad;
  const int co0 = co_t * 128, ci0 = ci_t * p.bn;
  const int total_rows = p.n_img * p.h;
  const int row_lo = split * p.rows_per_split;
  const int row_hi = min(total_rows, row_lo + p.rows_per_split);
  const int num_kb = max(0, row_hi - row_lo) * p.x_blocks;

  const bool do_bias = p.dbias != nullptr && ty == 0 && ci_t == 0;
  const int n_issue = p.tapn ? 2 : p.ksize;  // tap issuers (tapn: channel-block issuers)
  const uint32_t issuers = (uint32_t)n_issue + (do_bias ? 1u : 0u);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), issuers);  // released by every issuer
    }
    mbar_init(smem_u32(&bar_done), issuers);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (do_bias) {  // [KP pixels][64 ch] of bf16 1.0: an all-ones tile looks the same under any swizzle
    for (uint32_t i = threadIdx.x; i < WG_A_BOX / 16u; i += WG_THREADS)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_base + p.ones_off + i * 16u), "r"(0x3F803F80u) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]), done = smem_u32(&bar_done);
  const uint32_t a_bytes = 2u * WG_A_BOX;

  if (warp == 0) {
    uint32_t stage = 0, phase = 0;
    for (int row = row_lo; row < row_hi; ++row) {
      const int n = row / p.h, y = row - n * p.h;
      for (int xb = 0; xb < p.x_blocks; ++xb) {
        mbar_wait(empty0 + 8u * stage, phase ^ 1u);
        if (elect_one()) {
          const uint32_t full = full0 + 8u * stage;
          const uint32_t dst = smem_base + stage * p.stage_bytes;
          mbar_arrive_expect_tx(full, p.tx_bytes);
          tma_load_4d(dst, &map_dy, full, co0, xb * WG_KP, y, n);
          tma_load_4d(dst + WG_A_BOX, &map_dy, full, co0 + 64, xb * WG_KP, y, n);
          for (int b = 0; b < p.nboxes_b; ++b)
            tma_load_4d(dst + a_bytes + b * WG_B_BOX, &map_x, full, ci0 + 64 * b, xb * WG_KP - p.pad, y + dy, n);
        }
        __syncwarp();
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 || warp >= 6) {
    // One issuing warp per tap of the kernel row (warp 1: tx = 0, warps 6 / 7: tx = 1 / 2), each into its own accumulator: a
    // tcgen05.mma costs its issuing thread a fixed ~46 cycles on top of the N/2 the tensor pipe needs, so
    // one issuer left the pipe 54 % idle at N = 96; independent instruction streams overlap that cost.
    const int tx = warp == 1 ? 0 : warp - 5;   // 3: the bias issuer
    const bool bias_warp = tx == 3;
    if (tx < n_issue || (bias_warp && do_bias)) {
      // D fp32, A/B bf16, both MN-major (bits 15, 16), N = bn (tapn: 192 = three taps of one 64-channel block), M = 128
      const uint32_t n_mma = p.tapn ? 192u : (uint32_t)p.bn;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n_mma >> 3) << 17) | ((128u >> 4) << 24);
      // MN-major SW128 descriptors: LBO = distance between 64-element MN blocks (one box; tapn: one pixel row, i.e. the next
      // tap of the same box), SBO = 8 K-rows = 1024 B
      uint64_t hi_a = 0, hi_b = 0;
      hi_a |= (uint64_t)(WG_A_BOX >> 4) << 16;
      hi_b |= (uint64_t)((p.tapn ? 128u : WG_B_BOX) >> 4) << 16;
      hi_a |= (uint64_t)(1024u >> 4) << 32;
      hi_b |= (uint64_t)(1024u >> 4) << 32;
      hi_a |= ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      hi_b |= ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint32_t d_tmem = tmem_base + (uint32_t)(p.tapn ? tx * 192 : tx * p.bn);
      // bias gradient (its own issuer, so that its ~46 cycles per MMA overlap the taps'): D[co][0..15] += dY^T . ones
      const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t d_bias = tmem_base + (uint32_t)(p.tapn ? 384 : p.ksize * p.bn);
      const uint64_t ones_d = hi_a | (uint64_t)(((smem_base + p.ones_off) & 0x3FFFFu) >> 4);
      const uint32_t stage_lo = p.stage_bytes >> 4;
      const uint32_t a_lo0 = (smem_base & 0x3FFFFu) >> 4;
      // tap tx starts tx pixel rows (128 B) into the X box; tapn: channel block tx is the tx-th box
      const uint32_t b_off = (a_bytes >> 4) + (p.tapn ? (uint32_t)tx * (WG_B_BOX >> 4) : (uint32_t)tx * 8u);
      uint32_t stage = 0, phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full0 + 8u * stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a_lo0 + stage * stage_lo;
          const uint64_t ad = hi_a | (uint64_t)a_lo, bd = hi_b | (uint64_t)(a_lo + b_off);
          // 16 pixels = two 8-row groups = 2048 bytes further down the box
          if (bias_warp) {
            umma_bf16(d_bias, ad, ones_d, idesc1, kb ? 1u : 0u);
            umma_bf16(d_bias, ad + 128u, ones_d + 128u, idesc1, 1u);
            umma_bf16(d_bias, ad + 256u, ones_d + 256u, idesc1, 1u);
            umma_bf16(d_bias, ad + 384u, ones_d + 384u, idesc1, 1u);
          } else {
            umma_bf16(d_tmem, ad, bd, idesc, kb ? 1u : 0u);
            umma_bf16(d_tmem, ad + 128u, bd + 128u, idesc, 1u);
            umma_bf16(d_tmem, ad + 256u, bd + 256u, idesc, 1u);
            umma_bf16(d_tmem, ad + 384u, bd + 384u, idesc, 1u);
          }
          umma_commit(empty0 + 8u * stage);
          if (kb == num_kb - 1) umma_commit(done);
        }
        __syncwarp();
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 2 && warp < 6 && num_kb > 0) {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(done, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int tx = 0; tx < p.ksize; ++tx) {
      float* out = p.dw + (long long)co * p.s_co + (long long)(ty * p.ksize + tx) * p.s_tap;
      for (int c0 = 0; c0 < p.bn; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)(p.tapn ? (c0 >> 6) * 192 + tx * 64 + (c0 & 63) : tx * p.bn + c0), r);
        tmem_ld_wait();
        if (co < p.cout_valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (ci0 + c0 + j < p.cin_valid) atomicAdd(out + (long long)(ci0 + c0 + j) * p.s_ci, __uint_as_float(r[j]));
        }
      }
    }
    if (do_bias) {
      uint32_t r[16];
      tmem_ld16(taddr + (uint32_t)(p.tapn ? 384 : p.ksize * p.bn), r);
      tmem_ld_wait();
      if (co < p.cout_valid) atomicAdd(p.dbias + co, __uint_as_float(r[0]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// per-channel sum over pixels (bias gradient): x [pixels][c] bf16 -> out[ch] += sum, ch < c_valid.
// A thread owns one 16-byte chunk (8 channels) of the pixel row and walks pixels; the block's partial sums meet in shared
// memory, so each block issues one atomicAdd per channel.  c % 8 == 0.
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int64_t pixels, int c,
                                                    int c_valid) {
  using V = Vec16<__nv_bfloat16>;
  extern __shared__ float sacc[];  // [c]
  const int cv = c >> 3;
  for (int i = threadIdx.x; i < c; i += 256) sacc[i] = 0.f;
  __syncthreads();
  const int lanes = 256 / cv * cv;  // threads beyond the last whole pixel row idle
  if ((int)threadIdx.x < lanes) {
    const int vec = threadIdx.x % cv, prow = threadIdx.x / cv, rows_per_iter = lanes / cv;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * rows_per_iter + prow; r < pixels; r += (int64_t)gridDim.x * rows_per_iter) {
      V v;
      v.load(x + r * c + vec * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v.get(j);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sacc[vec * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c_valid; i += 256) atomicAdd(out + i, sacc[i]);
}

int tc_encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box, CUtensorMapSwizzle sw);
int tc_ensure_init();
static std::mutex g_wg_mu;
static bool g_wg_attr[64] = {false};

}  // namespace rv

extern "C" int rv_conv2d_wgrad(const void* x, const void* dy, float* dw, int64_t dw_co_stride, int64_t dw_ci_stride,
                               int64_t dw_tap_stride, float* dbias, int n, int h, int w, int cin, int cout, int cin_valid,
                               int cout_valid, int ksize, int pad, void* stream) {
  using namespace rv;
  if (int rc = tc_ensure_init()) return rc;
  RV_CHECK_ARG(x && dy && dw && n > 0 && h > 0 && w > 0, "conv2d_wgrad: bad argument");
  RV_CHECK_ARG(ksize == 1 || ksize == 3, "conv2d_wgrad: ksize must be 1 or 3");
  RV_CHECK_ARG(cin % 16 == 0 && cout % 8 == 0, "conv2d_wgrad: cin %% 16 and cout %% 8 must be 0 (got %d, %d)", cin, cout);
  RV_CHECK_ARG(cin_valid > 0 && cin_valid <= cin && cout_valid > 0 && cout_valid <= cout, "conv2d_wgrad: bad valid channel counts");
  RV_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)dy % 16 == 0), "conv2d_wgrad: tensors must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = n; p.h = h; p.w = w; p.cin = cin; p.cout = cout; p.ksize = ksize; p.pad = pad;
  p.cin_valid = cin_valid; p.cout_valid = cout_valid;
  p.co_tiles = (cout + 127) / 128;
  const int c16 = (cin + 15) / 16 * 16;
  const int bn_max = 128;  // ksize accumulators of bn columns share the 512 TMEM columns
  p.ci_tiles = (c16 + bn_max - 1) / bn_max;
  p.bn = ((c16 + p.ci_tiles - 1) / p.ci_tiles + 15) / 16 * 16;
  p.nboxes_b = (p.bn + 63) / 64;
  p.x_blocks = (w + WG_KP - 1) / WG_KP;
  p.tmem_cols = 32;
  // The bias gradient (column sums of dY) rides along as one more N = 16 MMA against a tile of ones in the CTAs of the first
  // kernel row / first Cin tile.  Measured: a win where those are a small share of the CTAs (384-channel layers: 9+ tiles per
  // split, 0.97 -> 1.10 PFLOP/s incl. the bias), a loss where they are a third of them (96 / 192 channels: the extra work
  // unbalances the grid) -- there a separate streaming pass over dY (colsum_kernel) is cheaper.
  const bool fuse_bias = dbias != nullptr && ksize * p.co_tiles * p.ci_tiles >= 9;
  static const bool no_tapn = getenv("RGBAVAE_WGRAD_NO_TAPN") != nullptr;
  p.tapn = (!no_tapn && ksize == 3 && p.bn == 128) ? 1 : 0;
  while ((int)p.tmem_cols < ksize * p.bn + (fuse_bias ? 16 : 0)) p.tmem_cols <<= 1;
  const int tiles = ksize * p.co_tiles * p.ci_tiles;
  const int total_rows = n * h;
  int splits = (num_sms() * 2) / tiles;
  if (splits > total_rows) splits = total_rows;
  if (splits < 1) splits = 1;
  p.rows_per_split = (total_rows + splits - 1) / splits;
  p.splits = (total_rows + p.rows_per_split - 1) / p.rows_per_split;
  p.dw = dw;
  p.s_co = dw_co_stride; p.s_ci = dw_ci_stride; p.s_tap = dw_tap_stride;
  const uint32_t b_rows = (uint32_t)(WG_KP + ksize - 1);
  p.tx_bytes = 2u * WG_A_BOX + (uint32_t)p.nboxes_b * b_rows * 128u;
  p.stage_bytes = 2u * WG_A_BOX + (uint32_t)p.nboxes_b * WG_B_BOX;
  p.stages = (int)((200u * 1024u - (fuse_bias ? WG_A_BOX : 0u)) / p.stage_bytes);
  if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
  p.dbias = fuse_bias ? dbias : nullptr;
  p.ones_off = (uint32_t)p.stages * p.stage_bytes;
  CUtensorMap mdy, mx;
  {
    cuuint64_t dims[4] = {(cuuint64_t)cout, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t str[3] = {(cuuint64_t)cout * 2u, (cuuint64_t)cout * 2u * w, (cuuint64_t)cout * 2u * w * h};
    cuuint32_t box[4] = {64, (cuuint32_t)WG_KP, 1, 1};
    if (int rc = tc_encode_map(&mdy, dy, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t str[3] = {(cuuint64_t)cin * 2u, (cuuint64_t)cin * 2u * w, (cuuint64_t)cin * 2u * w * h};
    cuuint32_t box[4] = {64, b_rows, 1, 1};
    if (int rc = tc_encode_map(&mx, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + (fuse_bias ? WG_A_BOX : 0u) + 1024;
  {
    std::lock_guard<std::mutex> lk(g_wg_mu);
    int dev = 0;
    RV_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !g_wg_attr[dev]) {
      RV_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
      g_wg_attr[dev] = true;
    }
  }
  {
    LaunchScope scope(CAT_CONV_TC, st, 2.0 * (double)n * h * w * cout * cin * ksize * ksize);
    conv_wgrad_kernel<<<tiles * p.splits, WG_THREADS, smem, st>>>(mdy, mx, p);
    RV_LAUNCH_CHECK();
  }
  if (dbias && !fuse_bias) {
    const int64_t pixels = (int64_t)n * h * w;
    LaunchScope scope(CAT_NORM, st, (double)pixels * cout * 2.0);
    RV_CHECK_ARG(cout <= 8192, "conv2d_wgrad: bias gradient supports up to 8192 output channels");
    const int cv = cout / 8;
    const int rows_per_iter = cv <= 256 ? 256 / cv : 0;
    RV_CHECK_ARG(rows_per_iter > 0, "conv2d_wgrad: bias gradient needs cout <= 2048");
    int64_t blocks = (pixels + (int64_t)rows_per_iter * 16 - 1) / ((int64_t)rows_per_iter * 16);
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    if (blocks < 1) blocks = 1;
    colsum_kernel<<<(unsigned)blocks, 256, (size_t)cout * sizeof(float), st>>>((const __nv_bfloat16*)dy, dbias, pixels, cout, cout_valid);
    RV_LAUNCH_CHECK();
  }
  return 0;
}
