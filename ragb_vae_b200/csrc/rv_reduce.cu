// Fused reductions over RGBA image pairs (NCHW planes): the AlphaVAE reconstruction loss and
// the composite-over-background + PSNR + alpha-MAE validation metrics.  Both read each input
// exactly once (8*B*H*W elements of traffic), accumulate in fp32 per thread, fp64 across the
// block, and finish in a small second kernel so results are deterministic.
#include "rv_common.cuh"

namespace rv {

constexpr int RB_PIX_PER_BLOCK = 4096;
constexpr int RB_MAX = 512;
constexpr int MAX_BG = 4;

static inline int reduce_blocks(int64_t hw) {
  int64_t b = (hw + RB_PIX_PER_BLOCK - 1) / RB_PIX_PER_BLOCK;
  if (b < 1) b = 1;
  if (b > RB_MAX) b = RB_MAX;
  return (int)b;
}

// Block-level fp64 sum of NV per-thread fp32 values; result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum_d(const float (&v)[NV], double (&out)[NV]) {
  __shared__ double red[NV][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double d = warp_sum_d((double)v[k]);
    if (lane == 0) red[k][wid] = d;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double d = 0.0;
    if (threadIdx.x == 0)
      for (int w = 0; w < 8; ++w) d += red[k][w];
    out[k] = d;
  }
}

struct LossParams {
  float eb[3], eb2[3];
  int naive;
};

// AlphaVaeLoss.reconstruction_loss (reference src/models/losses.py:67-83; oracle
// reconstruction_loss): d = t_rgb*at - p_rgb*ap, da = at - ap, l = d^2 - 2*Eb*d*da + Eb2*da^2.
// Raw 16-byte vectors are kept in registers (4 per tensor per vector index) and converted element by element, two
// independent vector indices per sweep: 16 loads in flight per thread at ~64 registers, so 8 blocks stay resident per SM.
template <typename T, int VEC>
struct PlaneVec {
  Vec16<T> v[4];
  float s[4];
  __device__ __forceinline__ void load(const T* base, int64_t hw, int64_t i, bool ok) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (VEC == 1) s[c] = ok ? ldf(base + c * hw + i) : 0.f;
      else if (ok) v[c].load(base + c * hw + i * VEC);
      else v[c].zero();
    }
  }
  __device__ __forceinline__ float get(int c, int j) const { return VEC == 1 ? s[c] : v[c].get(j); }
};

template <typename T, int VEC>
__global__ void __launch_bounds__(256, 4) recon_loss_kernel(const T* __restrict__ pred, const T* __restrict__ target,
                                                        double* __restrict__ partial, int64_t hw, LossParams lp) {
  const int n = blockIdx.y;
  const T* p = pred + (int64_t)n * 4 * hw;
  const T* t = target + (int64_t)n * 4 * hw;
  float acc[1] = {0.f};
  const int64_t nvec = hw / VEC;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int UN = 1;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < nvec; i0 += UN * stride) {
    PlaneVec<T, VEC> pv[UN], tv[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t i = i0 + u * stride;
      pv[u].load(p, hw, i, i < nvec);
      tv[u].load(t, hw, i, i < nvec);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        if (lp.naive) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float d = pv[u].get(c, j) - tv[u].get(c, j);
            acc[0] = fmaf(d, d, acc[0]);
          }
        } else {
          float at = (tv[u].get(3, j) + 1.0f) * 0.5f, ap = (pv[u].get(3, j) + 1.0f) * 0.5f;
          float da = at - ap;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            // separately rounded products (no fma contraction): identical inputs give exactly 0
            float d = __fsub_rn(__fmul_rn(tv[u].get(c, j), at), __fmul_rn(pv[u].get(c, j), ap));
            acc[0] += d * d - 2.0f * lp.eb[c] * d * da + lp.eb2[c] * da * da;
          }
        }
      }
    }
  }
  double out[1];
  block_sum_d<1>(acc, out);
  if (threadIdx.x == 0) partial[(int64_t)n * gridDim.x + blockIdx.x] = out[0];
}

__global__ void loss_finish_kernel(const double* __restrict__ partial, float* __restrict__ per_sample, int blocks) {
  const int n = blockIdx.x;
  double d = 0.0;
  for (int i = threadIdx.x; i < blocks; i += 32) d += partial[(int64_t)n * blocks + i];
  d = warp_sum_d(d);
  if (threadIdx.x == 0) per_sample[n] = (float)d;
}

struct PsnrParams {
  float bg[MAX_BG][3];
  int nbg;
};

// composite_over_background (src/models/rgba_vae.py:75-84) of recon and target for each
// background, squared error summed over (3,H,W) (compute_psnr, rgba_vae_stage.py:712-715) and
// |alpha_recon - alpha_target| (rgba_vae_stage.py:749-753).
template <typename T, int VEC>
__global__ void __launch_bounds__(256, 4) composite_psnr_kernel(const T* __restrict__ recon, const T* __restrict__ target,
                                                            double* __restrict__ partial, int64_t hw, PsnrParams pp) {
  const int n = blockIdx.y;
  const T* p = recon + (int64_t)n * 4 * hw;
  const T* t = target + (int64_t)n * 4 * hw;
  float acc[MAX_BG + 1];
#pragma unroll
  for (int k = 0; k <= MAX_BG; ++k) acc[k] = 0.f;
  const int64_t nvec = hw / VEC;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int UN = 1;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < nvec; i0 += UN * stride) {
    PlaneVec<T, VEC> pv[UN], tv[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t i = i0 + u * stride;
      pv[u].load(p, hw, i, i < nvec);
      tv[u].load(t, hw, i, i < nvec);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float ap = pv[u].get(3, j), at = tv[u].get(3, j);
        acc[MAX_BG] += fabsf(ap - at);
        float pr[3], tr[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          pr[c] = pv[u].get(c, j);
          tr[c] = tv[u].get(c, j);
        }
#pragma unroll
        for (int b = 0; b < MAX_BG; ++b) {
          if (b < pp.nbg) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float cp = pr[c] * ap + pp.bg[b][c] * (1.0f - ap);
              float ct = tr[c] * at + pp.bg[b][c] * (1.0f - at);
              float d = cp - ct;
              acc[b] = fmaf(d, d, acc[b]);
            }
          }
        }
      }
    }
  }
  double out[MAX_BG + 1];
  block_sum_d<MAX_BG + 1>(acc, out);
  if (threadIdx.x == 0) {
    double* dst = partial + ((int64_t)n * gridDim.x + blockIdx.x) * (MAX_BG + 1);
#pragma unroll
    for (int k = 0; k <= MAX_BG; ++k) dst[k] = out[k];
  }
}

__global__ void psnr_finish_kernel(const double* __restrict__ partial, float* __restrict__ out, int blocks, int nbg,
                                   double hw) {
  const int n = blockIdx.x;
  for (int k = 0; k <= nbg; ++k) {
    const int src = k < nbg ? k : MAX_BG;
    double d = 0.0;
    for (int i = threadIdx.x; i < blocks; i += 32) d += partial[((int64_t)n * blocks + i) * (MAX_BG + 1) + src];
    d = warp_sum_d(d);
    if (threadIdx.x == 0) {
      if (k < nbg) {
        double mse = d / (3.0 * hw);
        if (mse < 1e-8) mse = 1e-8;
        out[n * (nbg + 1) + k] = (float)(-10.0 * log10(mse));
      } else {
        out[n * (nbg + 1) + k] = (float)(d / hw);
      }
    }
  }
}

}  // namespace rv

extern "C" {

int rv_reduce_blocks(int64_t hw) { return rv::reduce_blocks(hw); }

int rv_recon_loss(const void* pred, const void* target, const float* eb_host, const float* eb2_host, int naive_mse,
                  float* per_sample, double* partial, int n, int64_t hw, int dtype, void* stream) {
  RV_CHECK_ARG(pred && target && per_sample && partial && n > 0 && hw > 0, "recon_loss: bad argument");
  RV_CHECK_ARG(naive_mse || (eb_host && eb2_host), "recon_loss: Eb / Eb2 missing");
  RV_CHECK_ARG(dtype == RV_F32 || dtype == RV_BF16, "recon_loss: bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  rv::LossParams lp;
  for (int c = 0; c < 3; ++c) {
    lp.eb[c] = eb_host ? eb_host[c] : 0.f;
    lp.eb2[c] = eb2_host ? eb2_host[c] : 0.f;
  }
  lp.naive = naive_mse;
  const int blocks = rv::reduce_blocks(hw);
  dim3 grid(blocks, n);
  const size_t es = dtype == RV_F32 ? 4 : 2;
  const int vec = 16 / (int)es;
  const bool vec_ok = hw % vec == 0 && ((uintptr_t)pred % 16 == 0) && ((uintptr_t)target % 16 == 0);
  {
    rv::LaunchScope scope(rv::CAT_LOSS, st, 8.0 * (double)n * hw * es);
    if (dtype == RV_F32) {
      if (vec_ok) rv::recon_loss_kernel<float, 4><<<grid, 256, 0, st>>>((const float*)pred, (const float*)target, partial, hw, lp);
      else rv::recon_loss_kernel<float, 1><<<grid, 256, 0, st>>>((const float*)pred, (const float*)target, partial, hw, lp);
    } else {
      if (vec_ok)
        rv::recon_loss_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>((const __nv_bfloat16*)pred,
                                                                      (const __nv_bfloat16*)target, partial, hw, lp);
      else
        rv::recon_loss_kernel<__nv_bfloat16, 1><<<grid, 256, 0, st>>>((const __nv_bfloat16*)pred,
                                                                      (const __nv_bfloat16*)target, partial, hw, lp);
    }
    RV_LAUNCH_CHECK();
  }
  {
    rv::LaunchScope scope(rv::CAT_LOSS, st, 0.0);
    rv::loss_finish_kernel<<<n, 32, 0, st>>>(partial, per_sample, blocks);
    RV_LAUNCH_CHECK();
  }
  return 0;
}

int rv_composite_psnr(const void* recon, const void* target, const float* bgs_host, int nbg, float* out,
                      double* partial, int n, int64_t hw, int dtype, void* stream) {
  RV_CHECK_ARG(recon && target && out && partial && n > 0 && hw > 0, "composite_psnr: bad argument");
  RV_CHECK_ARG(nbg >= 0 && nbg <= rv::MAX_BG && (nbg == 0 || bgs_host), "composite_psnr: 0..%d backgrounds", rv::MAX_BG);
  RV_CHECK_ARG(dtype == RV_F32 || dtype == RV_BF16, "composite_psnr: bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  rv::PsnrParams pp;
  pp.nbg = nbg;
  for (int b = 0; b < rv::MAX_BG; ++b)
    for (int c = 0; c < 3; ++c) pp.bg[b][c] = b < nbg ? bgs_host[b * 3 + c] : 0.f;
  const int blocks = rv::reduce_blocks(hw);
  dim3 grid(blocks, n);
  const size_t es = dtype == RV_F32 ? 4 : 2;
  const int vec = 16 / (int)es;
  const bool vec_ok = hw % vec == 0 && ((uintptr_t)recon % 16 == 0) && ((uintptr_t)target % 16 == 0);
  {
    rv::LaunchScope scope(rv::CAT_PSNR, st, 8.0 * (double)n * hw * es);
    if (dtype == RV_F32) {
      if (vec_ok)
        rv::composite_psnr_kernel<float, 4><<<grid, 256, 0, st>>>((const float*)recon, (const float*)target, partial, hw, pp);
      else
        rv::composite_psnr_kernel<float, 1><<<grid, 256, 0, st>>>((const float*)recon, (const float*)target, partial, hw, pp);
    } else {
      if (vec_ok)
        rv::composite_psnr_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>((const __nv_bfloat16*)recon,
                                                                          (const __nv_bfloat16*)target, partial, hw, pp);
      else
        rv::composite_psnr_kernel<__nv_bfloat16, 1><<<grid, 256, 0, st>>>((const __nv_bfloat16*)recon,
                                                                          (const __nv_bfloat16*)target, partial, hw, pp);
    }
    RV_LAUNCH_CHECK();
  }
  {
    rv::LaunchScope scope(rv::CAT_PSNR, st, 0.0);
    rv::psnr_finish_kernel<<<n, 32, 0, st>>>(partial, out, blocks, nbg, (double)hw);
    RV_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
