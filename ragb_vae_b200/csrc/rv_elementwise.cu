// HBM-bound kernels of the RGBA-VAE path: norm+SiLU (RMS and GroupNorm), softmax rows,
// NCHW<->NHWC plumbing and the posterior reparameterisation.  All are streaming kernels:
// 16-byte vector accesses, fp32 arithmetic, warp-shuffle + shared-memory reductions.
#include "rv_common.cuh"

namespace rv {

static inline int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

// ------------------------------------------------------------------------------------------
// RMS norm + SiLU (QwenImageRMS_norm + F.silu inside QwenImageResidualBlock / norm_out;
// oracle/vae_oracle.py QwenRMSNorm).  One 16-byte chunk per thread, flat over [pixels][c];
// the per-pixel sum of squares is reduced over the chunks of a pixel: xor-shuffle inside each
// lane quad (chunks-per-pixel is a multiple of 4), then a short shared-memory gather.
// ------------------------------------------------------------------------------------------
template <typename T, bool SILU, int U>
__global__ void __launch_bounds__(512) rmsnorm_silu_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                          T* __restrict__ y, int64_t total_chunks, int cpp,
                                                          float sqrt_c, int iters) {
  using V = Vec16<T>;
  extern __shared__ float part[];  // [2][U][blockDim/4]
  const int tid = threadIdx.x;
  const int nq = blockDim.x >> 2;
  const int col = tid % cpp;
  const int pix = tid / cpp;
  const int qpp = cpp >> 2;
  float g[V::N];
#pragma unroll
  for (int j = 0; j < V::N; ++j) g[j] = gamma[col * V::N + j] * sqrt_c;

  for (int it = 0; it < iters; ++it) {
    const int64_t base = ((int64_t)blockIdx.x * iters + it) * (int64_t)(blockDim.x * U);
    float* pbuf = part + (it & 1) * U * nq;
    V v[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = base + (int64_t)u * blockDim.x + tid;
      ok[u] = i < total_chunks;
      if (ok[u]) v[u].load(x + i * V::N);
      else v[u].zero();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < V::N; ++j) {
        float f = v[u].get(j);
        ss = fmaf(f, f, ss);
      }
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);
      ss += __shfl_xor_sync(0xffffffffu, ss, 2);
      if ((tid & 3) == 0) pbuf[u * nq + (tid >> 2)] = ss;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float tot = 0.f;
      const float* pp = pbuf + u * nq + pix * qpp;
      for (int j = 0; j < qpp; ++j) tot += pp[j];
      float rinv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
      if (ok[u]) {
        V o;
#pragma unroll
        for (int j = 0; j < V::N; ++j) {
          float f = v[u].get(j) * rinv * g[j];
          o.set(j, SILU ? silu_t<T>(f) : f);
        }
        int64_t i = base + (int64_t)u * blockDim.x + tid;
        o.store(y + i * V::N);
      }
    }
  }
}


// Fast path: a pixel's channels are split over L lanes (L = 4/8/16/32, a power of two) with three
// 16-byte chunks per lane, so the sum of squares is a pure xor-shuffle reduction: no shared memory,
// no block barrier, 3*PIX independent 16-byte loads in flight per thread.
template <typename T, bool SILU, int PIX>
__global__ void __launch_bounds__(256) rmsnorm_silu_warp_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                               T* __restrict__ y, int64_t pixels, int lanes_per_pixel,
                                                               float sqrt_c) {
  using V = Vec16<T>;
  constexpr int CPL = 3;
  const int L = lanes_per_pixel;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (L - 1);          // lane within the pixel group
  const int grp = lane / L;                // pixel group within the warp
  const int ppw = 32 / L;                  // pixels per warp per step
  const int64_t c = (int64_t)L * CPL * V::N;
  // bf16 + SiLU: the 1/2 of silu(f) = h tanh(h) + h, h = f / 2, rides in gamma (one multiply less per element)
  constexpr bool HALF = SILU && sizeof(T) == 2;
  float g[CPL][V::N];
#pragma unroll
  for (int k = 0; k < CPL; ++k)
#pragma unroll
    for (int j = 0; j < V::N; ++j) g[k][j] = gamma[(k * L + sub) * V::N + j] * sqrt_c * (HALF ? 0.5f : 1.0f);
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t p0 = warp_global * ppw * PIX; p0 < pixels; p0 += warps_total * ppw * PIX) {
    V v[PIX][CPL];
    bool ok[PIX];
#pragma unroll
    for (int u = 0; u < PIX; ++u) {
      const int64_t pix = p0 + (int64_t)u * ppw + grp;
      ok[u] = pix < pixels;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        if (ok[u]) v[u][k].load(x + pix * c + (int64_t)(k * L + sub) * V::N);
        else v[u][k].zero();
      }
    }
#pragma unroll
    for (int u = 0; u < PIX; ++u) {
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < CPL; ++k)
#pragma unroll
        for (int j = 0; j < V::N; ++j) {
          float f = v[u][k].get(j);
          ss = fmaf(f, f, ss);
        }
      for (int o = 1; o < L; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rinv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
      if (ok[u]) {
        const int64_t pix = p0 + (int64_t)u * ppw + grp;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          V o;
#pragma unroll
          for (int j = 0; j < V::N; ++j) {
            float f = v[u][k].get(j) * (rinv * g[k][j]);
            o.set(j, HALF ? silu_from_half(f) : (SILU ? silu_t<T>(f) : f));
          }
          o.store(y + pix * c + (int64_t)(k * L + sub) * V::N);
        }
      }
    }
  }
}

template <typename T>
static int launch_rmsnorm(const void* x, const float* gamma, void* y, int64_t pixels, int c, int apply_silu,
                          cudaStream_t st) {
  constexpr int VN = Vec16<T>::N;
  constexpr int U = 2;
  RV_CHECK_ARG(c % (4 * VN) == 0, "rmsnorm: channels (%d) must be a multiple of %d", c, 4 * VN);
  int cpp = c / VN;
  if (cpp % 3 == 0) {
    const int lanes = cpp / 3;
    if (lanes >= 1 && lanes <= 32 && (lanes & (lanes - 1)) == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0)) {
      constexpr int PIX = 2;
      const int ppw = 32 / lanes;
      const int64_t warps_needed = (pixels + (int64_t)ppw * PIX - 1) / ((int64_t)ppw * PIX);
      int64_t blocks = (warps_needed + 7) / 8;
      // a few blocks per SM that LOOP (two are resident at ~96 registers): with one block per 32 pixels a 100 MB tensor was
      // 4096 one-iteration blocks (cap 32 per SM): 53 us against 45 us with 6 per SM; the large tensors gain 1-2 %
      const int64_t cap = (int64_t)num_sms() * 6;
      if (blocks > cap) blocks = cap;
      if (blocks < 1) blocks = 1;
      LaunchScope scope(CAT_NORM, st, 2.0 * (double)pixels * c * sizeof(T));
      if (apply_silu)
        rmsnorm_silu_warp_kernel<T, true, PIX><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, gamma, (T*)y, pixels, lanes,
                                                                               sqrtf((float)c));
      else
        rmsnorm_silu_warp_kernel<T, false, PIX><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, gamma, (T*)y, pixels, lanes,
                                                                                sqrtf((float)c));
      RV_LAUNCH_CHECK();
      return 0;
    }
  }
  int l = cpp / gcd_i(cpp, 32) * 32;  // lcm(cpp, 32)
  RV_CHECK_ARG(l <= 512, "rmsnorm: unsupported channel count %d", c);
  int block = l * (256 / l > 0 ? 256 / l : 1);
  int64_t total_chunks = pixels * cpp;
  int64_t per_iter = (int64_t)block * U;
  int64_t need = (total_chunks + per_iter - 1) / per_iter;
  int64_t target_blocks = (int64_t)num_sms() * 16;
  int iters = (int)((need + target_blocks - 1) / target_blocks);
  if (iters < 1) iters = 1;
  int64_t blocks = (need + iters - 1) / iters;
  if (blocks < 1) blocks = 1;
  size_t smem = 2 * U * (block / 4) * sizeof(float);
  LaunchScope scope(CAT_NORM, st, 2.0 * (double)pixels * c * sizeof(T));
  if (apply_silu)
    rmsnorm_silu_kernel<T, true, U><<<(unsigned)blocks, block, smem, st>>>((const T*)x, gamma, (T*)y, total_chunks, cpp,
                                                                        sqrtf((float)c), iters);
  else
    rmsnorm_silu_kernel<T, false, U><<<(unsigned)blocks, block, smem, st>>>((const T*)x, gamma, (T*)y, total_chunks,
                                                                         cpp, sqrtf((float)c), iters);
  RV_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// GroupNorm(32, eps=1e-6) + SiLU (diffusers ResnetBlock2D.norm1/norm2, Attention.group_norm,
// conv_norm_out).  Pass 1 accumulates (sum, sumsq) per (sample, group) in fp64 through a
// shared-memory stage; pass 2 is the streaming apply.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(512) groupnorm_stats_kernel(const T* __restrict__ x, double* __restrict__ stats,
                                                             int64_t hw, int c, int groups, int cpp,
                                                             int64_t rows_per_block) {
  using V = Vec16<T>;
  extern __shared__ float sred[];  // [threads][2 * V::N]: every thread's partial sums, folded below without atomics
  const int tid = threadIdx.x;
  const int n = blockIdx.y;
  const int col = tid % cpp;
  const int rows_per_iter = blockDim.x / cpp;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(hw, r0 + rows_per_block);
  const T* xb = x + (int64_t)n * hw * c;
  float s[V::N], q[V::N];
#pragma unroll
  for (int j = 0; j < V::N; ++j) s[j] = q[j] = 0.f;
#pragma unroll 4
  for (int64_t r = r0 + tid / cpp; r < r1; r += rows_per_iter) {
    V v;
    v.load(xb + r * c + col * V::N);
#pragma unroll
    for (int j = 0; j < V::N; ++j) {
      float f = v.get(j);
      s[j] += f;
      q[j] = fmaf(f, f, q[j]);
    }
  }
  // Block reduction through plain shared memory: thread (g, which) adds the group's channels over the block's rows in fp64 and
  // issues ONE global atomic.  (fp64 atomicAdd on shared memory is a compare-and-swap loop: with 16 rows x 4 channels of
  // every group colliding on one address it cost as much as the streaming loop -- ncu: 3.46 TB/s for a read-only pass.)
  const int cg = c / groups;
  float* my = sred + tid * (2 * V::N);
#pragma unroll
  for (int j = 0; j < V::N; ++j) {
    my[j] = s[j];
    my[V::N + j] = q[j];
  }
  __syncthreads();
  for (int i = tid; i < groups * 2; i += blockDim.x) {
    const int g = i >> 1, which = i & 1;
    double acc = 0.0;
    for (int k = 0; k < cg; ++k) {
      const int ch = g * cg + k;
      const float* src = sred + (ch / V::N) * (2 * V::N) + which * V::N + (ch % V::N);
      for (int r = 0; r < rows_per_iter; ++r) acc += (double)src[(int64_t)r * cpp * (2 * V::N)];
    }
    atomicAdd(&stats[(int64_t)n * groups * 2 + i], acc);
  }
}

template <typename T, bool SILU>
__global__ void __launch_bounds__(512) groupnorm_apply_kernel(const T* __restrict__ x, const double* __restrict__ stats,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, T* __restrict__ y,
                                                             int64_t hw, int c, int groups, int cpp, float eps,
                                                             int64_t rows_per_block) {
  using V = Vec16<T>;
  extern __shared__ float smr[];  // [groups][2]: mean, rstd
  const int tid = threadIdx.x;
  const int n = blockIdx.y;
  const double cnt = (double)hw * (double)(c / groups);
  for (int g = tid; g < groups; g += blockDim.x) {
    double m = stats[((int64_t)n * groups + g) * 2] / cnt;
    double var = stats[((int64_t)n * groups + g) * 2 + 1] / cnt - m * m;
    if (var < 0.0) var = 0.0;
    smr[2 * g] = (float)m;
    smr[2 * g + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int col = tid % cpp;
  const int rows_per_iter = blockDim.x / cpp;
  const int cg = c / groups;
  float a[V::N], b[V::N];  // y = x*a + b
#pragma unroll
  for (int j = 0; j < V::N; ++j) {
    int ch = col * V::N + j;
    int g = ch / cg;
    float ga = gamma[ch], be = beta[ch];
    a[j] = smr[2 * g + 1] * ga;
    b[j] = be - smr[2 * g] * smr[2 * g + 1] * ga;
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(hw, r0 + rows_per_block);
  const T* xb = x + (int64_t)n * hw * c;
  T* yb = y + (int64_t)n * hw * c;
#pragma unroll 4
  for (int64_t r = r0 + tid / cpp; r < r1; r += rows_per_iter) {
    V v, o;
    v.load(xb + r * c + col * V::N);
#pragma unroll
    for (int j = 0; j < V::N; ++j) {
      float f = fmaf(v.get(j), a[j], b[j]);
      o.set(j, SILU ? silu_t<T>(f) : f);
    }
    o.store(yb + r * c + col * V::N);
  }
}

struct GnGeom {
  int cpp, block;
  int64_t rows_per_block;
  unsigned blocks_x;
};

template <typename T>
static int gn_geometry(int n, int64_t hw, int c, int groups, GnGeom* g) {
  constexpr int VN = Vec16<T>::N;
  RV_CHECK_ARG(groups > 0 && c % groups == 0 && c % VN == 0, "groupnorm: bad channel/group count %d/%d", c, groups);
  g->cpp = c / VN;
  RV_CHECK_ARG(g->cpp <= 512, "groupnorm: too many channels (%d)", c);
  int rows = 256 / g->cpp;
  if (rows < 1) rows = 1;
  g->block = rows * g->cpp;
  // Up to 64 rows per thread, fewer for small tensors so that one sample still makes ~2 blocks per SM (a 128^2 x 512 tensor
  // cut into 64-row-per-thread blocks was 64 blocks per sample: 88 us for 67 MB).  The partition depends on hw and c only, so
  // a sample's statistics (and therefore its output bits) do not depend on the batch it is in.
  (void)n;
  int64_t rpt = hw / ((int64_t)rows * 296);
  rpt = rpt > 64 ? 64 : (rpt < 8 ? 8 : rpt / 8 * 8);
  int64_t rpb = (int64_t)rows * rpt;
  g->rows_per_block = rpb;
  g->blocks_x = (unsigned)((hw + rpb - 1) / rpb);
  return 0;
}

// ------------------------------------------------------------------------------------------
// Row softmax (the SDPA softmax of the single-head mid-block attention; scores come from the
// QK^T GEMM already scaled by 1/sqrt(C)).  One block per row; pass 1 keeps an online
// (max, sum) pair, pass 2 re-reads the row (L2-resident) and writes probabilities.
// ------------------------------------------------------------------------------------------
// fp32 probabilities (the c1 parity mode) use expf; bf16 probabilities the fast intrinsic.
template <typename TP> __device__ __forceinline__ float sm_exp(float x);
template <> __device__ __forceinline__ float sm_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ float sm_exp<__nv_bfloat16>(float x) { return __expf(x); }

template <typename TP>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, TP* __restrict__ p,
                                                          int64_t cols, int64_t ld_s, int64_t ld_p) {
  __shared__ float sm_m[8], sm_l[8];
  const float* row = s + (int64_t)blockIdx.x * ld_s;
  TP* out = p + (int64_t)blockIdx.x * ld_p;
  const int tid = threadIdx.x;
  float m = -INFINITY, l = 0.f;
  for (int64_t i = (int64_t)tid * 4; i < cols; i += 256 * 4) {
    float4 v = *reinterpret_cast<const float4*>(row + i);
    float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    float mn = fmaxf(m, mx);
    l = l * sm_exp<TP>(m - mn) + sm_exp<TP>(v.x - mn) + sm_exp<TP>(v.y - mn) + sm_exp<TP>(v.z - mn) + sm_exp<TP>(v.w - mn);
    m = mn;
  }
  // combine (m, l) across the block
  float wm = warp_max(m);
  l *= (m == -INFINITY) ? 0.f : sm_exp<TP>(m - wm);
  l = warp_sum(l);
  if ((tid & 31) == 0) {
    sm_m[tid >> 5] = wm;
    sm_l[tid >> 5] = l;
  }
  __syncthreads();
  float bm = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) bm = fmaxf(bm, sm_m[i]);
  float bl = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) bl += (sm_m[i] == -INFINITY) ? 0.f : sm_l[i] * sm_exp<TP>(sm_m[i] - bm);
  const float inv = 1.0f / bl;
  for (int64_t i = (int64_t)tid * 4; i < cols; i += 256 * 4) {
    float4 v = *reinterpret_cast<const float4*>(row + i);
    stf(out + i + 0, sm_exp<TP>(v.x - bm) * inv);
    stf(out + i + 1, sm_exp<TP>(v.y - bm) * inv);
    stf(out + i + 2, sm_exp<TP>(v.z - bm) * inv);
    stf(out + i + 3, sm_exp<TP>(v.w - bm) * inv);
  }
}

// ------------------------------------------------------------------------------------------
// NCHW <-> NHWC (boundary tensors only: 4-channel images, 16/32-channel latents)
// ------------------------------------------------------------------------------------------
template <typename TX, typename TY>
__global__ void nchw_to_nhwc_kernel(const TX* __restrict__ x, TY* __restrict__ y, int c, int64_t hw, int c_pad,
                                    float scale, float shift) {
  const int n = blockIdx.y;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    TY* o = y + ((int64_t)n * hw + i) * c_pad;
    for (int ch = 0; ch < c_pad; ++ch) {
      float v = 0.f;
      if (ch < c) v = ldf(x + ((int64_t)n * c + ch) * hw + i) * scale + shift;
      stf(o + ch, v);
    }
  }
}

// Same, for pixel rows that are whole 16-byte chunks (the 16-channel padded stems): the per-channel plane reads stay
// coalesced across the warp, the pixel's row leaves as 16-byte stores instead of c_pad scalar ones.
template <typename TX, typename TY, int CHUNKS>
__global__ void __launch_bounds__(256) nchw_to_nhwc_vec_kernel(const TX* __restrict__ x, TY* __restrict__ y, int c, int64_t hw,
                                                              float scale, float shift) {
  using V = Vec16<TY>;
  constexpr int CP = CHUNKS * V::N;
  const int n = blockIdx.y;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    float v[CP];
#pragma unroll
    for (int ch = 0; ch < CP; ++ch) v[ch] = ch < c ? ldf(x + ((int64_t)n * c + ch) * hw + i) * scale + shift : 0.f;
    TY* o = y + ((int64_t)n * hw + i) * CP;
    V out[CHUNKS];
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k)
#pragma unroll
      for (int j = 0; j < V::N; ++j) out[k].set(j, v[k * V::N + j]);
    if (sizeof(TY) == 2 && CHUNKS % 2 == 0) {  // bf16 rows of 32 / 64 bytes: whole-sector stores
#pragma unroll
      for (int k = 0; k < CHUNKS; k += 2)
        store32(o + k * V::N, *reinterpret_cast<const uint4*>(&out[k].v), *reinterpret_cast<const uint4*>(&out[k + 1].v));
    } else {
#pragma unroll
      for (int k = 0; k < CHUNKS; ++k) out[k].store(o + k * V::N);
    }
  }
}

// Few-channel stem loader: 16 bf16 channels per pixel = the pixel's three horizontal neighbours (dx-major, channel-minor),
// zero outside the image and past 3c.  One thread per pixel, plane reads coalesced across the warp (the +-1 neighbours
// hit L1), two 16-byte stores.
template <typename TX>
__global__ void __launch_bounds__(256) nchw_to_nhwc_hpack_kernel(const TX* __restrict__ x, __nv_bfloat16* __restrict__ y, int c,
                                                                int h, int w, float scale, float shift) {
  using V = Vec16<__nv_bfloat16>;
  const int n = blockIdx.z, py = blockIdx.y;
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= w) return;
  const int64_t hw = (int64_t)h * w;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
#pragma unroll
  for (int dx = 0; dx < 3; ++dx) {
    const int sx = px + dx - 1;
    if (sx < 0 || sx >= w) continue;
    for (int ch = 0; ch < c; ++ch) v[dx * c + ch] = ldf(x + ((int64_t)n * c + ch) * hw + (int64_t)py * w + sx) * scale + shift;
  }
  __nv_bfloat16* o = y + (((int64_t)n * h + py) * w + px) * 16;
  V out[2];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) out[k].set(j, v[k * 8 + j]);
  if ((reinterpret_cast<uintptr_t>(y) & 31u) == 0) store32(o, out[0].v, out[1].v);
  else {
    out[0].store(o);
    out[1].store(o + 8);
  }
}

template <typename TX, typename TY>
__global__ void nhwc_to_nchw_kernel(const TX* __restrict__ x, TY* __restrict__ y, int c, int64_t hw, int x_cstride) {
  const int n = blockIdx.y;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    const TX* src = x + ((int64_t)n * hw + i) * x_cstride;
    for (int ch = 0; ch < c; ++ch) stf(y + ((int64_t)n * c + ch) * hw + i, ldf(src + ch));
  }
}


// ------------------------------------------------------------------------------------------
// DiagonalGaussianDistribution.sample with a supplied noise tensor (+ optional KL)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) reparam_kernel(const T* __restrict__ moments, const T* __restrict__ noise,
                                                     T* __restrict__ z, float* __restrict__ kl, int zc, int64_t hw,
                                                     float z_shift, float z_scale) {
  __shared__ float red[8];
  const int n = blockIdx.y;
  const int64_t per = (int64_t)zc * hw;
  const T* mean = moments + (int64_t)n * 2 * per;
  const T* logv = mean + per;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
    float mu = ldf(mean + i);
    float lv = fminf(fmaxf(ldf(logv + i), -30.f), 20.f);
    float sd = expf(0.5f * lv);
    if (z) {
      float e = ldf(noise + (int64_t)n * per + i);
      stf(z + (int64_t)n * per + i, (fmaf(sd, e, mu) - z_shift) * z_scale);
    }
    acc += mu * mu + expf(lv) - 1.0f - lv;
  }
  if (kl) {
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
      v = warp_sum(v);
      if (threadIdx.x == 0) atomicAdd(kl + n, 0.5f * v);
    }
  }
}

}  // namespace rv

extern "C" {

int rv_rmsnorm_silu(const void* x, const float* gamma, void* y, int64_t pixels, int c, int dtype, int apply_silu,
                    void* stream) {
  RV_CHECK_ARG(x && gamma && y && pixels > 0 && c > 0, "rmsnorm: bad argument");
  if (dtype == RV_F32) return rv::launch_rmsnorm<float>(x, gamma, y, pixels, c, apply_silu, (cudaStream_t)stream);
  if (dtype == RV_BF16)
    return rv::launch_rmsnorm<__nv_bfloat16>(x, gamma, y, pixels, c, apply_silu, (cudaStream_t)stream);
  RV_CHECK_ARG(false, "rmsnorm: bad dtype %d", dtype);
}

int rv_groupnorm_stats(const void* x, double* stats, int n, int64_t hw, int c, int groups, int dtype, void* stream) {
  RV_CHECK_ARG(x && stats && n > 0 && hw > 0, "groupnorm_stats: bad argument");
  RV_CHECK_ARG(dtype == RV_F32 || dtype == RV_BF16, "groupnorm_stats: bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  RV_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)n * groups, st));
  rv::GnGeom g;
  size_t es = dtype == RV_F32 ? 4 : 2;
  rv::LaunchScope scope(rv::CAT_NORM, st, (double)n * hw * c * es);
  if (dtype == RV_F32) {
    if (int rc = rv::gn_geometry<float>(n, hw, c, groups, &g)) return rc;
    rv::groupnorm_stats_kernel<float><<<dim3(g.blocks_x, n), g.block, sizeof(float) * 8 * g.block, st>>>(
        (const float*)x, stats, hw, c, groups, g.cpp, g.rows_per_block);
  } else {
    if (int rc = rv::gn_geometry<__nv_bfloat16>(n, hw, c, groups, &g)) return rc;
    rv::groupnorm_stats_kernel<__nv_bfloat16><<<dim3(g.blocks_x, n), g.block, sizeof(float) * 16 * g.block, st>>>(
        (const __nv_bfloat16*)x, stats, hw, c, groups, g.cpp, g.rows_per_block);
  }
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_groupnorm_silu(const void* x, const double* stats, const float* gamma, const float* beta, void* y, int n,
                      int64_t hw, int c, int groups, float eps, int dtype, int apply_silu, void* stream) {
  RV_CHECK_ARG(x && stats && gamma && beta && y && n > 0 && hw > 0, "groupnorm_silu: bad argument");
  RV_CHECK_ARG(dtype == RV_F32 || dtype == RV_BF16, "groupnorm_silu: bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  rv::GnGeom g;
  size_t es = dtype == RV_F32 ? 4 : 2;
  rv::LaunchScope scope(rv::CAT_NORM, st, 2.0 * (double)n * hw * c * es);
  size_t smem = sizeof(float) * 2 * groups;
#define RV_GN_LAUNCH(T, S)                                                                              \
  rv::groupnorm_apply_kernel<T, S><<<dim3(g.blocks_x, n), g.block, smem, st>>>(                          \
      (const T*)x, stats, gamma, beta, (T*)y, hw, c, groups, g.cpp, eps, g.rows_per_block)
  if (dtype == RV_F32) {
    if (int rc = rv::gn_geometry<float>(n, hw, c, groups, &g)) return rc;
    if (apply_silu) RV_GN_LAUNCH(float, true);
    else RV_GN_LAUNCH(float, false);
  } else {
    if (int rc = rv::gn_geometry<__nv_bfloat16>(n, hw, c, groups, &g)) return rc;
    if (apply_silu) RV_GN_LAUNCH(__nv_bfloat16, true);
    else RV_GN_LAUNCH(__nv_bfloat16, false);
  }
#undef RV_GN_LAUNCH
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_softmax_rows(const float* s, void* p, int64_t rows, int64_t cols, int64_t ld_s, int64_t ld_p, int dtype,
                    void* stream) {
  RV_CHECK_ARG(s && p && rows > 0 && cols > 0, "softmax: bad argument");
  RV_CHECK_ARG(cols % 4 == 0 && ld_s % 4 == 0 && ld_s >= cols && ld_p >= cols, "softmax: cols/ld must be multiples of 4");
  RV_CHECK_ARG(rows < (1ll << 31), "softmax: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  rv::LaunchScope scope(rv::CAT_SOFTMAX, st, (double)rows * cols * (4.0 + (dtype == RV_F32 ? 4.0 : 2.0)));
  if (dtype == RV_F32)
    rv::softmax_rows_kernel<float><<<(unsigned)rows, 256, 0, st>>>(s, (float*)p, cols, ld_s, ld_p);
  else if (dtype == RV_BF16)
    rv::softmax_rows_kernel<__nv_bfloat16><<<(unsigned)rows, 256, 0, st>>>(s, (__nv_bfloat16*)p, cols, ld_s, ld_p);
  else
    RV_CHECK_ARG(false, "softmax: bad dtype %d", dtype);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_nchw_to_nhwc(const void* x, void* y, int n, int c, int64_t hw, int c_pad, int x_dtype, int y_dtype, float scale,
                    float shift, void* stream) {
  RV_CHECK_ARG(x && y && n > 0 && c > 0 && hw > 0 && c_pad >= c, "nchw_to_nhwc: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned bx = (unsigned)((hw + 255) / 256);
  if (bx > 4096) bx = 4096;
  dim3 grid(bx, n);
  rv::LaunchScope scope(rv::CAT_LAYOUT, st,
                        (double)n * hw * (c * (x_dtype == RV_F32 ? 4.0 : 2.0) + c_pad * (y_dtype == RV_F32 ? 4.0 : 2.0)));
  if (y_dtype == RV_BF16 && (c_pad == 16 || c_pad == 32) && ((uintptr_t)y % 32 == 0)) {
    // 16-byte stores (the channel-padded stems and the 32-channel moments)
#define RV_VEC(TX, CH) \
  rv::nchw_to_nhwc_vec_kernel<TX, __nv_bfloat16, CH><<<grid, 256, 0, st>>>((const TX*)x, (__nv_bfloat16*)y, c, hw, scale, shift)
    if (x_dtype == RV_F32) {
      if (c_pad == 16) RV_VEC(float, 2);
      else RV_VEC(float, 4);
    } else if (x_dtype == RV_BF16) {
      if (c_pad == 16) RV_VEC(__nv_bfloat16, 2);
      else RV_VEC(__nv_bfloat16, 4);
    } else {
      RV_CHECK_ARG(false, "nchw_to_nhwc: bad dtype");
    }
#undef RV_VEC
    RV_LAUNCH_CHECK();
    return 0;
  }
  if (x_dtype == RV_F32 && y_dtype == RV_F32)
    rv::nchw_to_nhwc_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, c, hw, c_pad, scale, shift);
  else if (x_dtype == RV_F32 && y_dtype == RV_BF16)
    rv::nchw_to_nhwc_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)x, (__nv_bfloat16*)y, c, hw, c_pad,
                                                                       scale, shift);
  else if (x_dtype == RV_BF16 && y_dtype == RV_F32)
    rv::nchw_to_nhwc_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (float*)y, c, hw, c_pad,
                                                                       scale, shift);
  else if (x_dtype == RV_BF16 && y_dtype == RV_BF16)
    rv::nchw_to_nhwc_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, c, hw, c_pad, scale, shift);
  else
    RV_CHECK_ARG(false, "nchw_to_nhwc: bad dtype");
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_nchw_to_nhwc_hpack(const void* x, void* y, int n, int c, int h, int w, int x_dtype, float scale, float shift, void* stream) {
  RV_CHECK_ARG(x && y && n > 0 && c > 0 && 3 * c <= 16 && h > 0 && w > 0, "nchw_to_nhwc_hpack: bad argument (3*c <= 16)");
  RV_CHECK_ARG((uintptr_t)y % 16 == 0, "nchw_to_nhwc_hpack: y must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((w + 255) / 256), (unsigned)h, (unsigned)n);
  rv::LaunchScope scope(rv::CAT_LAYOUT, st, (double)n * h * w * (c * (x_dtype == RV_F32 ? 4.0 : 2.0) + 32.0));
  if (x_dtype == RV_F32)
    rv::nchw_to_nhwc_hpack_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (__nv_bfloat16*)y, c, h, w, scale, shift);
  else if (x_dtype == RV_BF16)
    rv::nchw_to_nhwc_hpack_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, c, h, w, scale, shift);
  else
    RV_CHECK_ARG(false, "nchw_to_nhwc_hpack: bad dtype");
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_nhwc_to_nchw(const void* x, void* y, int n, int c, int64_t hw, int x_cstride, int x_dtype, int y_dtype,
                    void* stream) {
  RV_CHECK_ARG(x && y && n > 0 && c > 0 && hw > 0 && x_cstride >= c, "nhwc_to_nchw: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned bx = (unsigned)((hw + 255) / 256);
  if (bx > 4096) bx = 4096;
  dim3 grid(bx, n);
  rv::LaunchScope scope(rv::CAT_LAYOUT, st,
                        (double)n * hw * c * ((x_dtype == RV_F32 ? 4.0 : 2.0) + (y_dtype == RV_F32 ? 4.0 : 2.0)));
  if (x_dtype == RV_F32 && y_dtype == RV_F32)
    rv::nhwc_to_nchw_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, c, hw, x_cstride);
  else if (x_dtype == RV_F32 && y_dtype == RV_BF16)
    rv::nhwc_to_nchw_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)x, (__nv_bfloat16*)y, c, hw,
                                                                       x_cstride);
  else if (x_dtype == RV_BF16 && y_dtype == RV_F32)
    rv::nhwc_to_nchw_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (float*)y, c, hw,
                                                                       x_cstride);
  else if (x_dtype == RV_BF16 && y_dtype == RV_BF16)
    rv::nhwc_to_nchw_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x,
                                                                               (__nv_bfloat16*)y, c, hw, x_cstride);
  else
    RV_CHECK_ARG(false, "nhwc_to_nchw: bad dtype");
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_reparam(const void* moments, const void* noise, void* z, float* kl_out, int n, int zc, int64_t hw, int dtype,
               float z_shift, float z_scale, void* stream) {
  RV_CHECK_ARG(moments && n > 0 && zc > 0 && hw > 0, "reparam: bad argument");
  RV_CHECK_ARG((z == nullptr) == (noise == nullptr), "reparam: z and noise must be given together");
  RV_CHECK_ARG(z || kl_out, "reparam: nothing to compute");
  cudaStream_t st = (cudaStream_t)stream;
  if (kl_out) RV_CUDA(cudaMemsetAsync(kl_out, 0, sizeof(float) * n, st));
  int64_t per = (int64_t)zc * hw;
  unsigned bx = (unsigned)((per + 255) / 256);
  if (bx > 1024) bx = 1024;
  dim3 grid(bx, n);
  size_t es = dtype == RV_F32 ? 4 : 2;
  rv::LaunchScope scope(rv::CAT_REPARAM, st, (double)n * per * es * (z ? 4.0 : 2.0));
  if (dtype == RV_F32)
    rv::reparam_kernel<float><<<grid, 256, 0, st>>>((const float*)moments, (const float*)noise, (float*)z, kl_out, zc, hw,
                                                   z_shift, z_scale);
  else if (dtype == RV_BF16)
    rv::reparam_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)moments, (const __nv_bfloat16*)noise,
                                                           (__nv_bfloat16*)z, kl_out, zc, hw, z_shift, z_scale);
  else
    RV_CHECK_ARG(false, "reparam: bad dtype %d", dtype);
  RV_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
