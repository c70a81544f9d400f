// conv_out of the decoder: 3x3, Cin = 96 (Qwen) / 128 (Flux) -> Cout <= 5 (RGBA = 4), full resolution, NCHW output with the
// range map + clamp of RgbaVAE.forward fused (src/models/rgba_vae.py:279-280).  An HBM-bound layer (192 B read per pixel, 8 B
// written, 6.9 kFLOP): it must move at memory speed, but as an ordinary implicit GEMM with N = 16 (Cout padded) it needs 54
// tcgen05.mma per 128 pixels, each ~46 cycles whatever N is -- 1.0 ms at 8 x 1024^2, 25 % of HBM speed.
//
// Here the kernel ROW index rides in N instead of in the tap loop:
//     Q[yin][x][dy, co] = sum_{dx, c} X[yin][x + dx - 1][c] * W[co][c][dy][dx]      (3 taps x K = Cin, N = 3 * Cout <= 16)
//     out[y][x][co]     = Q[y-1][x][0, co] + Q[y][x][1, co] + Q[y+1][x][2, co] + bias[co]
// so every INPUT row is multiplied once (18 MMAs of N = 16 per 128 pixels for Cin = 96 instead of 54), and the sum over dy is
// three 4-column TMEM reads by the thread that owns the pixel -- no cross-lane traffic, no intermediate in HBM.
//
//   * input rows: the halo kernel's ring (rv_conv_halo.cu): each row's 130 pixels are TMA-loaded once as K-major SWIZZLE_128B
//     [130][64 ch] (+ SWIZZLE_64B [130][32 ch]); tap dx is the same slot addressed from row dx on (absolute-address swizzle);
//   * weights: all three taps ([3][16 rows = dy * Cout + co][Cin], 9-12 KB) are loaded ONCE per CTA and stay in shared memory;
//   * Q rows: a ring of 8 accumulators of 16 TMEM columns; the epilogue for output row y waits for Q[y+1], reads the three
//     column groups, and frees Q[y-1];
//   * warps: 0 row TMA, 1 MMA issuer (+ TMEM alloc, weight load), 4-7 epilogue (one thread per pixel of the 128-column strip).
#include <cstring>
#include <mutex>

#include "rv_tc_common.cuh"

namespace rv {

constexpr int CO_RING = 6;                  // input-row slots
constexpr int CO_QRING = 8;                 // Q-row accumulators (16 TMEM columns each)
constexpr int CO_N = 16;                    // MMA N: 3 * cout rows of the tap matrix, zero padded
constexpr int CO_THREADS = 256;
constexpr int CO_PIX = 130;                 // 128 output columns + halo
constexpr uint32_t CO_R128_BYTES = 130 * 128;
constexpr uint32_t CO_R64_BYTES = 130 * 64;
constexpr uint32_t CO_SMEM_MAX = 227 * 1024 - 2048;

struct ConvOutParams {
  int n_img, h, w, cin, cout;
  int nk128, has64;
  int col_blocks, strips_per_col, strip_rows, total_strips;
  uint32_t row_slot_bytes, r64_off, row_tx_bytes;
  uint32_t w_tap_bytes, w64_off, w_tx_bytes, ring_off;
  int y_f32, clamp;
  float out_scale, out_shift, clamp_lo, clamp_hi;
  const float* bias;  // [cout] or null
  void* y;            // NCHW [n][cout][h][w]
};

struct CoStrip {
  int img, x0, ys, rows;
};
__device__ __forceinline__ CoStrip co_strip(const ConvOutParams& p, int s) {
  CoStrip c;
  const int sy = s % p.strips_per_col;
  const int t = s / p.strips_per_col;
  c.x0 = (t % p.col_blocks) * 128;
  c.img = t / p.col_blocks;
  c.ys = sy * p.strip_rows;
  c.rows = p.h - c.ys;
  if (c.rows > p.strip_rows) c.rows = p.strip_rows;
  return c;
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}

template <int NK128, int HAS64>
__global__ void __launch_bounds__(CO_THREADS, 1)
conv_out_kernel(const __grid_constant__ CUtensorMap map_a128, const __grid_constant__ CUtensorMap map_a64,
                const __grid_constant__ CUtensorMap map_b128, const __grid_constant__ CUtensorMap map_b64,
                const __grid_constant__ ConvOutParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_rowfull[CO_RING];
  __shared__ __align__(8) uint64_t bar_rowempty[CO_RING];
  __shared__ __align__(8) uint64_t bar_qfull[CO_QRING];
  __shared__ __align__(8) uint64_t bar_qempty[CO_QRING];
  __shared__ __align__(8) uint64_t bar_w;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t wbase = (smem_u32(smem_raw) + 1023u) & ~1023u;  // weights first (1024-aligned), then the row ring
  const uint32_t ring = wbase + p.ring_off;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < CO_RING; ++s) {
      mbar_init(smem_u32(&bar_rowfull[s]), 1);
      mbar_init(smem_u32(&bar_rowempty[s]), 1);
    }
    for (int s = 0; s < CO_QRING; ++s) {
      mbar_init(smem_u32(&bar_qfull[s]), 1);
      mbar_init(smem_u32(&bar_qempty[s]), 4);  // the four epilogue warps
    }
    mbar_init(smem_u32(&bar_w), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a128) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b128) : "memory");
  }
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)(CO_QRING * CO_N))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t rowfull0 = smem_u32(&bar_rowfull[0]), rowempty0 = smem_u32(&bar_rowempty[0]);
  const uint32_t qfull0 = smem_u32(&bar_qfull[0]), qempty0 = smem_u32(&bar_qempty[0]);
  const uint32_t wbar = smem_u32(&bar_w);

  if (warp == 0) {
    // ------------------------------ input-row producer ------------------------------
    uint32_t g = 0;
    for (int s = blockIdx.x; s < p.total_strips; s += gridDim.x) {
      const CoStrip c = co_strip(p, s);
      for (int y = c.ys - 1; y <= c.ys + c.rows; ++y) {
        const uint32_t slot = g % CO_RING, par = (g / CO_RING) & 1u;
        mbar_wait(rowempty0 + 8u * slot, par ^ 1u);
        if (elect_one()) {
          const uint32_t dst = ring + slot * p.row_slot_bytes;
          const uint32_t full = rowfull0 + 8u * slot;
          mbar_arrive_expect_tx(full, p.row_tx_bytes);
          for (int kb = 0; kb < p.nk128; ++kb) tma_load_4d(dst + kb * CO_R128_BYTES, &map_a128, full, kb * 64, c.x0 - 1, y, c.img);
          if (p.has64) tma_load_4d(dst + p.r64_off, &map_a64, full, p.nk128 * 64, c.x0 - 1, y, c.img);
        }
        __syncwarp();
        ++g;
      }
    }
  } else if (warp == 1) {
    // ------------------------------ weights (once) + MMA issuer ------------------------------
    if (elect_one()) {
      mbar_arrive_expect_tx(wbar, p.w_tx_bytes);
      for (int tap = 0; tap < 3; ++tap) {
        for (int kb = 0; kb < p.nk128; ++kb) tma_load_2d(wbase + tap * p.w_tap_bytes + kb * 2048u, &map_b128, wbar, kb * 64, tap * CO_N);
        if (p.has64) tma_load_2d(wbase + tap * p.w_tap_bytes + p.w64_off, &map_b64, wbar, p.nk128 * 64, tap * CO_N);
      }
    }
    __syncwarp();
    mbar_wait(wbar, 0u);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CO_N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t hi128 = make_smem_desc(0u, 1024u, 2u);
    const uint64_t hi64 = make_smem_desc(0u, 512u, 4u);
    const uint32_t w_lo = (wbase & 0x3FFFFu) >> 4;
    const uint32_t r64_lo = p.r64_off >> 4, w64_lo = p.w64_off >> 4, wtap_lo = p.w_tap_bytes >> 4;
    uint32_t g = 0, qc = 0;
    for (int s = blockIdx.x; s < p.total_strips; s += gridDim.x) {
      const CoStrip c = co_strip(p, s);
      for (int i = 0; i < c.rows + 2; ++i) {
        const uint32_t slot = g % CO_RING, qs = qc % CO_QRING;
        mbar_wait(qempty0 + 8u * qs, ((qc / CO_QRING) & 1u) ^ 1u);
        mbar_wait(rowfull0 + 8u * slot, (g / CO_RING) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = ((ring + slot * p.row_slot_bytes) & 0x3FFFFu) >> 4;
          const uint32_t d_tmem = tmem_base + qs * CO_N;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const uint32_t b_lo = w_lo + dx * wtap_lo;
#pragma unroll
            for (int kb = 0; kb < NK128; ++kb) {
              const uint64_t ad = hi128 | (uint64_t)(a_lo + kb * (CO_R128_BYTES >> 4) + dx * 8u);
              const uint64_t bd = hi128 | (uint64_t)(b_lo + kb * 128u);
              umma_bf16(d_tmem, ad, bd, idesc, (dx == 0 && kb == 0) ? 0u : 1u);
              umma_bf16(d_tmem, ad + 2u, bd + 2u, idesc, 1u);
              umma_bf16(d_tmem, ad + 4u, bd + 4u, idesc, 1u);
              umma_bf16(d_tmem, ad + 6u, bd + 6u, idesc, 1u);
            }
            if (HAS64) {
              const uint64_t ad = hi64 | (uint64_t)(a_lo + r64_lo + dx * 4u);
              const uint64_t bd = hi64 | (uint64_t)(b_lo + w64_lo);
              umma_bf16(d_tmem, ad, bd, idesc, (dx == 0 && NK128 == 0) ? 0u : 1u);
              umma_bf16(d_tmem, ad + 2u, bd + 2u, idesc, 1u);
            }
          }
          umma_commit(qfull0 + 8u * qs);
          umma_commit(rowempty0 + 8u * slot);
        }
        __syncwarp();
        ++g;
        ++qc;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue: one thread per pixel of the strip's 128 columns ------------------------------
    const int q = warp & 3;
    const int col = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int cout = p.cout;
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    float bias4 = 0.f;
    if (p.bias) {
      for (int co = 0; co < 4 && co < cout; ++co) bias[co] = p.bias[co];
      if (cout > 4) bias4 = p.bias[4];
    }
    const int64_t plane = (int64_t)p.h * p.w;
    uint32_t qc = 0;  // Q row index of the strip's first input row (ys - 1)
    for (int s = blockIdx.x; s < p.total_strips; s += gridDim.x) {
      const CoStrip c = co_strip(p, s);
      const int x = c.x0 + col;
      for (int k = 0; k < c.rows; ++k) {
        // output row ys + k = Q[k] (dy = 0) + Q[k+1] (dy = 1) + Q[k+2] (dy = 2); the newest of the three completes last
        const uint32_t q2 = qc + (uint32_t)k + 2u;
        mbar_wait(qfull0 + 8u * (q2 % CO_QRING), (q2 / CO_QRING) & 1u);
        tc_fence_after();
        float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (cout == 4) {  // RGBA: the three row groups are the column quads 0-3 / 4-7 / 8-11 (static register indexing)
          uint32_t r[3][4];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) tmem_ld4(lane_base + ((qc + (uint32_t)k + (uint32_t)dy) % CO_QRING) * CO_N + 4u * dy, r[dy]);
          tmem_ld_wait();
#pragma unroll
          for (int co = 0; co < 4; ++co) v[co] = __uint_as_float(r[0][co]) + __uint_as_float(r[1][co]) + __uint_as_float(r[2][co]);
        } else {  // any other cout <= 5: one column at a time
          for (int dy = 0; dy < 3; ++dy) {
            const uint32_t qs = (qc + (uint32_t)k + (uint32_t)dy) % CO_QRING;
            for (int co = 0; co < cout; ++co) {
              uint32_t r1;
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r1) : "r"(lane_base + qs * CO_N + (uint32_t)(dy * cout + co)));
              tmem_ld_wait();
              const float f = __uint_as_float(r1);
              if (co == 0) v[0] += f;
              else if (co == 1) v[1] += f;
              else if (co == 2) v[2] += f;
              else if (co == 3) v[3] += f;
              else v[4] += f;
            }
          }
        }
        const int y = c.ys + k;
        if (x < p.w) {
          const int64_t base = ((int64_t)c.img * cout) * plane + (int64_t)y * p.w + x;
#pragma unroll
          for (int co = 0; co < 5; ++co) {
            if (co < cout) {
              float o = fmaf(v[co] + (co < 4 ? bias[co] : bias4), p.out_scale, p.out_shift);
              if (p.clamp) o = fminf(fmaxf(o, p.clamp_lo), p.clamp_hi);
              if (p.y_f32) reinterpret_cast<float*>(p.y)[base + co * plane] = o;
              else reinterpret_cast<__nv_bfloat16*>(p.y)[base + co * plane] = __float2bfloat16_rn(o);
            }
          }
        }
        // Q[k] is dead now; the strip's last output row also frees Q[rows] and Q[rows + 1]
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(qempty0 + 8u * ((qc + (uint32_t)k) % CO_QRING));
          if (k == c.rows - 1) {
            mbar_arrive(qempty0 + 8u * ((qc + (uint32_t)k + 1u) % CO_QRING));
            mbar_arrive(qempty0 + 8u * ((qc + (uint32_t)k + 2u) % CO_QRING));
          }
        }
      }
      qc += (uint32_t)c.rows + 2u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(CO_QRING * CO_N)) : "memory");
  }
}

int tc_encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box, CUtensorMapSwizzle sw);  // rv_conv_tc.cu
int tc_ensure_init();

static std::mutex g_co_mu;
static bool g_co_attr[64] = {false};

}  // namespace rv

extern "C" {

int rv_conv_out(const rv_conv_desc* d, const void* x, const void* w_taps, const float* bias, void* y, void* stream) {
  using namespace rv;
  RV_CHECK_ARG(d && x && w_taps && y, "conv_out: null argument");
  RV_CHECK_ARG(d->ksize == 3 && d->stride == 1 && !d->upsample && d->pad_lo == 1 && !d->taps_1d, "conv_out: 3x3 / stride 1 / pad 1 only");
  RV_CHECK_ARG((d->cin == 64 || d->cin == 96 || d->cin == 128) && d->x_dtype == RV_BF16 && !d->x_nchw && d->x_cstride % 8 == 0,
               "conv_out: NHWC bf16 input with 64 / 96 / 128 channels (got %d)", d->cin);
  RV_CHECK_ARG(d->cout >= 1 && 3 * d->cout <= CO_N, "conv_out: at most %d output channels (got %d)", CO_N / 3, d->cout);
  RV_CHECK_ARG(d->y_nchw && d->bias_mode != 2 && d->alpha == 1.0f, "conv_out: NCHW output, per-channel bias, alpha 1");
  RV_CHECK_ARG(d->w >= 64 && d->h >= 1 && d->oh == d->h && d->ow == d->w, "conv_out: image too small (%d x %d)", d->h, d->w);
  if (int rc = tc_ensure_init()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  ConvOutParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = d->n;
  p.h = d->h;
  p.w = d->w;
  p.cin = d->cin;
  p.cout = d->cout;
  p.nk128 = d->cin / 64;
  p.has64 = (d->cin % 64) ? 1 : 0;
  p.row_slot_bytes = p.nk128 * CO_R128_BYTES + p.has64 * CO_R64_BYTES;
  p.r64_off = p.nk128 * CO_R128_BYTES;
  p.row_tx_bytes = (uint32_t)CO_PIX * (uint32_t)(p.nk128 * 128 + p.has64 * 64);
  p.w_tap_bytes = (uint32_t)p.nk128 * 2048u + (uint32_t)p.has64 * 1024u;  // [16 rows][128 B] per 64 channels (+ [16][64 B])
  p.w64_off = (uint32_t)p.nk128 * 2048u;
  p.w_tx_bytes = 3u * p.w_tap_bytes;
  p.ring_off = (3u * p.w_tap_bytes + 1023u) & ~1023u;
  p.y_f32 = d->y_dtype == RV_F32;
  p.clamp = d->clamp;
  p.out_scale = d->out_scale;
  p.out_shift = d->out_shift;
  p.clamp_lo = d->clamp_lo;
  p.clamp_hi = d->clamp_hi;
  p.bias = d->bias_mode == 1 ? bias : nullptr;
  p.y = y;
  p.col_blocks = (d->w + 127) / 128;
  const int64_t cols = (int64_t)d->n * p.col_blocks;
  int rows = (int)(((int64_t)d->h * cols + (int64_t)num_sms() * 8 - 1) / ((int64_t)num_sms() * 8));  // ~8 strips per SM
  if (rows < 16) rows = 16;
  if (rows > d->h) rows = d->h;
  p.strip_rows = rows;
  p.strips_per_col = (d->h + rows - 1) / rows;
  p.total_strips = (int)(cols * p.strips_per_col);
  const size_t smem = (size_t)p.ring_off + (size_t)CO_RING * p.row_slot_bytes + 1024;
  RV_CHECK_ARG(smem <= CO_SMEM_MAX, "conv_out: operands do not fit shared memory (%zu bytes)", smem);

  CUtensorMap ma128, ma64, mb128, mb64;
  const uint64_t pitch_b = (uint64_t)d->x_cstride * 2u;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
    cuuint64_t str[3] = {pitch_b, pitch_b * d->w, pitch_b * d->w * d->h};
    cuuint32_t box[4] = {64, (cuuint32_t)CO_PIX, 1, 1};
    if (int rc = tc_encode_map(&ma128, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    cuuint32_t box64[4] = {32, (cuuint32_t)CO_PIX, 1, 1};
    if (int rc = tc_encode_map(&ma64, x, 4, dims, str, box64, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }
  {  // w_taps: bf16 [3 taps * 16 rows][cin]
    cuuint64_t dims[2] = {(cuuint64_t)d->cin, (cuuint64_t)(3 * CO_N)};
    cuuint64_t str[1] = {(cuuint64_t)d->cin * 2u};
    cuuint32_t box[2] = {64, (cuuint32_t)CO_N};
    if (int rc = tc_encode_map(&mb128, w_taps, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    cuuint32_t box64[2] = {32, (cuuint32_t)CO_N};
    if (int rc = tc_encode_map(&mb64, w_taps, 2, dims, str, box64, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }
  {
    std::lock_guard<std::mutex> lk(g_co_mu);
    int dev = 0;
    RV_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !g_co_attr[dev]) {
      RV_CUDA(cudaFuncSetAttribute(conv_out_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CO_SMEM_MAX));
      RV_CUDA(cudaFuncSetAttribute(conv_out_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CO_SMEM_MAX));
      RV_CUDA(cudaFuncSetAttribute(conv_out_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CO_SMEM_MAX));
      g_co_attr[dev] = true;
    }
  }
  const double flops = 2.0 * (double)d->n * d->oh * d->ow * d->cout * d->cin * 9.0;
  LaunchScope scope(CAT_CONV_TC, st, flops);
  const int grid = p.total_strips < num_sms() ? p.total_strips : num_sms();
  if (p.nk128 == 1 && !p.has64) conv_out_kernel<1, 0><<<grid, CO_THREADS, smem, st>>>(ma128, ma64, mb128, mb64, p);
  else if (p.nk128 == 1 && p.has64) conv_out_kernel<1, 1><<<grid, CO_THREADS, smem, st>>>(ma128, ma64, mb128, mb64, p);
  else conv_out_kernel<2, 0><<<grid, CO_THREADS, smem, st>>>(ma128, ma64, mb128, mb64, p);
  RV_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
