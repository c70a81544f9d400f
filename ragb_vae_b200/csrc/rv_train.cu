// Building blocks of the rgba_vae training step (reference src/training/rgba_vae_stage.py:433-523) that are not
// convolutions: backward of the AlphaVAE reconstruction loss, of the reparameterisation (+KL), of RMS-norm + SiLU,
// the global gradient norm, and the fused clip + AdamW update over flat parameter buffers.
// All HBM-bound streaming kernels; fp32 arithmetic.
#include "rv_common.cuh"

namespace rv {

struct LossBwdParams {
  float eb[3], eb2[3];
  int naive;
  float scale;  // upstream gradient x reduction factor (1/(B*3*HW) for reduce_mean, 1/B otherwise)
  float clamp_lo, clamp_hi;  // zero gradient where pred sits on the clamp (lo >= hi: no clamp)
};

// d loss / d pred for AlphaVaeLoss.reconstruction_loss (src/models/losses.py:67-83):
//   ap = (p3+1)/2, at = (t3+1)/2, d_c = t_c*at - p_c*ap, da = at - ap, l_c = d_c^2 - 2 Eb_c d_c da + Eb2_c da^2
template <typename T>
__global__ void __launch_bounds__(256) recon_loss_bwd_kernel(const T* __restrict__ pred, const T* __restrict__ target,
                                                            T* __restrict__ dpred, int64_t hw, LossBwdParams lp) {
  const int n = blockIdx.y;
  const T* p = pred + (int64_t)n * 4 * hw;
  const T* t = target + (int64_t)n * 4 * hw;
  T* g = dpred + (int64_t)n * 4 * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    float pv[4], tv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      pv[c] = ldf(p + c * hw + i);
      tv[c] = ldf(t + c * hw + i);
    }
    float m[4];  // gradient mask of the decoder's output clamp (AutoencoderKLQwenImage._decode clamps to [-1,1])
#pragma unroll
    for (int c = 0; c < 4; ++c) m[c] = (lp.clamp_lo < lp.clamp_hi && (pv[c] <= lp.clamp_lo || pv[c] >= lp.clamp_hi)) ? 0.f : lp.scale;
    if (lp.naive) {
#pragma unroll
      for (int c = 0; c < 4; ++c) stf(g + c * hw + i, 2.0f * (pv[c] - tv[c]) * m[c]);
    } else {
      const float at = (tv[3] + 1.0f) * 0.5f, ap = (pv[3] + 1.0f) * 0.5f;
      const float da = at - ap;
      float dalpha = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = tv[c] * at - pv[c] * ap;
        stf(g + c * hw + i, (2.0f * d - 2.0f * lp.eb[c] * da) * (-ap) * m[c]);
        dalpha += -2.0f * d * pv[c] + 2.0f * lp.eb[c] * (pv[c] * da + d) - 2.0f * lp.eb2[c] * da;
      }
      stf(g + 3 * hw + i, 0.5f * dalpha * m[3]);
    }
  }
}

// z = mean + exp(0.5*clamp(logvar,-30,20))*eps ; loss += kl_w * 0.5*sum(mean^2 + var - 1 - logvar)
// dmoments = [dz + kl_w*mean | (dz*eps*0.5*std + kl_w*0.5*(var-1)) inside the clamp range, else 0]
template <typename T>
__global__ void __launch_bounds__(256) reparam_bwd_kernel(const T* __restrict__ moments, const T* __restrict__ noise,
                                                         const T* __restrict__ dz, T* __restrict__ dmoments, int zc, int64_t hw,
                                                         float kl_w) {
  const int n = blockIdx.y;
  const int64_t per = (int64_t)zc * hw;
  const T* mean = moments + (int64_t)n * 2 * per;
  const T* logv = mean + per;
  T* dmean = dmoments + (int64_t)n * 2 * per;
  T* dlogv = dmean + per;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
    const float mu = ldf(mean + i), lv_raw = ldf(logv + i);
    const float lv = fminf(fmaxf(lv_raw, -30.f), 20.f);
    const float sd = expf(0.5f * lv);
    const float gz = dz ? ldf(dz + (int64_t)n * per + i) : 0.f;
    const float e = noise ? ldf(noise + (int64_t)n * per + i) : 0.f;
    stf(dmean + i, gz + kl_w * mu);
    const bool inside = lv_raw >= -30.f && lv_raw <= 20.f;
    stf(dlogv + i, inside ? (gz * e * 0.5f * sd + kl_w * 0.5f * (sd * sd - 1.0f)) : 0.f);
  }
}

// KL(N(mean, var) || N(rmean, rvar)) against a frozen reference posterior (rgba_vae_stage.py:489-508; diffusers
// DiagonalGaussianDistribution.kl(other)): per element 0.5 * ((m - rm)^2 / rv + v / rv - 1 - lv + rlv), log-variances clamped
// to [-30, 20] on both sides.  kl_out[n] += per-sample sum; dmoments (optional) = weight * d kl / d moments:
// d/dm = (m - rm) / rv, d/dlv = 0.5 * (v / rv - 1) inside the clamp range, else 0.
template <typename T>
__global__ void __launch_bounds__(256) kl_ref_kernel(const T* __restrict__ moments, const T* __restrict__ ref, float* __restrict__ kl_out,
                                                    T* __restrict__ dmoments, int zc, int64_t hw, float weight) {
  __shared__ float red[8];
  const int n = blockIdx.y;
  const int64_t per = (int64_t)zc * hw;
  const T* mean = moments + (int64_t)n * 2 * per;
  const T* logv = mean + per;
  const T* rmean = ref + (int64_t)n * 2 * per;
  const T* rlogv = rmean + per;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = ldf(mean + i), lv_raw = ldf(logv + i);
    const float rm = ldf(rmean + i);
    const float lv = fminf(fmaxf(lv_raw, -30.f), 20.f);
    const float rlv = fminf(fmaxf(ldf(rlogv + i), -30.f), 20.f);
    const float v = expf(lv), irv = expf(-rlv);
    const float d = m - rm;
    acc += 0.5f * (d * d * irv + v * irv - 1.0f - lv + rlv);
    if (dmoments) {
      T* dmean = dmoments + (int64_t)n * 2 * per;
      stf(dmean + i, weight * d * irv);
      const bool inside = lv_raw >= -30.f && lv_raw <= 20.f;
      stf(dmean + per + i, inside ? weight * 0.5f * (v * irv - 1.0f) : 0.f);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(kl_out + n, t);
  }
}

// Backward of y = act(x * r * g), r = 1/max(||x||_2, 1e-12), g = gamma*sqrt(C), act = SiLU or identity.
// Same work split as the forward warp kernel: a pixel's channels lie on L lanes (L = C / (3 * 16-byte chunk)), three
// 16-byte chunks per lane, so both per-pixel reductions are xor-shuffles.  A lane owns fixed channels, so dgamma
// accumulates in registers across pixels; it is reduced over the warp's pixel groups by shuffles, over the block's warps
// through shared memory, and leaves the block as ONE atomicAdd per channel, already scaled to d loss / d gamma.
template <typename T, bool SILU>
__global__ void __launch_bounds__(256) rmsnorm_silu_bwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma_scaled,
                                                              const T* __restrict__ dy, const T* __restrict__ add,
                                                              T* __restrict__ dx, float* __restrict__ dgamma, int64_t pixels,
                                                              int lanes_per_pixel, float dgamma_scale) {
  using V = Vec16<T>;
  constexpr int CPL = 3;
  extern __shared__ float sdg[];  // [8 warps][C]
  const int L = lanes_per_pixel;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int sub = lane & (L - 1), grp = lane / L, ppw = 32 / L;
  const int c = L * CPL * V::N;
  // gh = gamma_scaled / 2 is the only per-channel constant: h = xr * gh is the tanh argument of the SiLU derivative, the
  // per-pixel dot product sum(x * g * du) = 2 / r * sum(h * du) reuses it, and the output's g * r * du = gh * (2 r * du)
  // (12.4 ms of the c4 step ran at 63 % issue utilisation with 29 instructions per element: ncu, r02b_rmsbwd_raw.csv)
  float gh[CPL][V::N], dg[CPL][V::N];
#pragma unroll
  for (int k = 0; k < CPL; ++k)
#pragma unroll
    for (int j = 0; j < V::N; ++j) {
      gh[k][j] = 0.5f * gamma_scaled[(k * L + sub) * V::N + j];
      dg[k][j] = 0.f;
    }
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t p0 = warp_global * ppw; p0 < pixels; p0 += warps_total * ppw) {
    const int64_t pix = p0 + grp;
    const bool ok = pix < pixels;
    V xv[CPL], dv[CPL], av[CPL];  // av: gradient of the skip branch that meets this one (optional), added to dx
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      if (ok) {
        xv[k].load(x + pix * c + (int64_t)(k * L + sub) * V::N);
        dv[k].load(dy + pix * c + (int64_t)(k * L + sub) * V::N);
        if (add) av[k].load(add + pix * c + (int64_t)(k * L + sub) * V::N);
        else av[k].zero();
      } else {
        xv[k].zero();
        dv[k].zero();
        av[k].zero();
      }
    }
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < CPL; ++k)
#pragma unroll
      for (int j = 0; j < V::N; ++j) {
        const float f = xv[k].get(j);
        ss = fmaf(f, f, ss);
      }
    for (int o = 1; o < L; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float nrm = sqrtf(ss);
    const bool clamped = nrm < 1e-12f;
    const float r = 1.0f / fmaxf(nrm, 1e-12f);
    float du[CPL][V::N];
    float hd = 0.f;  // sum(h * du) = r / 2 * sum(x * g * du)
#pragma unroll
    for (int k = 0; k < CPL; ++k)
#pragma unroll
      for (int j = 0; j < V::N; ++j) {
        const float xr = xv[k].get(j) * r;
        const float h = xr * gh[k][j];
        float d = dv[k].get(j);
        if (SILU) {
          if (sizeof(T) == 2) {  // silu'(u) = (1 + t + h * (1 - t * t)) / 2, t = tanh(h), h = u / 2: one SFU op (see gn_dsilu)
            float t;
            asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
            const float dh = 0.5f * d;
            d = fmaf(dh, fmaf(h, fmaf(-t, t, 1.0f), t), dh);
          } else {
            const float u = 2.0f * h;
            const float sg = __fdividef(1.0f, 1.0f + __expf(-u));
            d *= sg * (1.0f + u * (1.0f - sg));
          }
        }
        du[k][j] = d;
        dg[k][j] = fmaf(xr, d, dg[k][j]);
        hd = fmaf(h, d, hd);
      }
    for (int o = 1; o < L; o <<= 1) hd += __shfl_xor_sync(0xffffffffu, hd, o);
    const float r2 = 2.0f * r;
    const float corr = clamped ? 0.f : hd * r2 * r;  // sum(x g du) * r^3
    if (ok) {
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        V o;
#pragma unroll
        for (int j = 0; j < V::N; ++j) o.set(j, fmaf(gh[k][j], r2 * du[k][j], fmaf(-xv[k].get(j), corr, av[k].get(j))));
        o.store(dx + pix * c + (int64_t)(k * L + sub) * V::N);
      }
    }
  }
  // warp: sum over the pixel groups (lanes with equal `sub`), then block, then one atomic per channel
#pragma unroll
  for (int k = 0; k < CPL; ++k)
#pragma unroll
    for (int j = 0; j < V::N; ++j) {
      float v = dg[k][j];
      for (int o = L; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (grp == 0) sdg[wib * c + (k * L + sub) * V::N + j] = v;
    }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += sdg[i * c + ch];
    atomicAdd(dgamma + ch, acc * dgamma_scale);
  }
}

// Squared L2 norm in two deterministic stages (fixed partition, fixed summation order): data-parallel replicas must
// compute bit-identical clip factors from bit-identical all-reduced gradients, or their weights drift apart.
__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ partial) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = g[i];
    acc += (double)v * (double)v;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) sqnorm_finish_kernel(const double* __restrict__ partial, int blocks, float* __restrict__ out) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < blocks; i += 256) acc += partial[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] += (float)red[0];
}

// device-resident step counter: state = [t, 1 - beta1^t, 1 - beta2^t]; one launch per optimizer step, ahead of adamw_kernel
__global__ void adamw_advance_kernel(float* __restrict__ state, float beta1, float beta2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float t = state[0] + 1.0f;
    state[0] = t;
    state[1] = 1.0f - powf(beta1, t);
    state[2] = 1.0f - powf(beta2, t);
  }
}

// torch.optim.AdamW semantics (decoupled weight decay, bias correction), gradients scaled by
// min(1, max_norm / (sqrt(*sqnorm) + 1e-6)) (clip_grad_norm_) and by grad_scale (1/world for an all-reduce SUM).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, __nv_bfloat16* __restrict__ p_bf16, int64_t n, float lr,
                                                   float beta1, float beta2, float eps, float wd, float bc1, float bc2,
                                                   float grad_scale, const float* __restrict__ sqnorm, float max_norm,
                                                   const float* __restrict__ bc_dev) {
  if (bc_dev != nullptr) {  // step counter kept on the device (CUDA-graph replay): state = [t, 1-b1^t, 1-b2^t]
    bc1 = bc_dev[1];
    bc2 = bc_dev[2];
  }
  float clip = 1.0f;
  if (sqnorm != nullptr && max_norm > 0.f) {
    const float nrm = sqrtf(*sqnorm) * grad_scale;
    clip = fminf(1.0f, max_norm / (nrm + 1e-6f));
  }
  const float gs = grad_scale * clip;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(pi);
  }
}


// ---- spatial helpers of the backward pass (NHWC bf16, 16-byte vectors along channels) --------------------------------
// mode 0: zero-insert x2   y[n][2h][2w][c]: y[2i][2j] = x[i][j], 0 elsewhere      (dY of a stride-2 conv -> full grid)
// mode 1: nearest x2       y[n][2h][2w][c]: y[i][j] = x[i/2][j/2]                  (forward input of the upsample conv)
// mode 2: 2x2 sum pool     y[n][h/2][w/2][c] = sum of the 2x2 block of x           (dX of the nearest upsample)
__global__ void __launch_bounds__(256) resample2x_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int h,
                                                        int w, int c, int mode) {
  using V = Vec16<__nv_bfloat16>;
  const int cv = c / 8;
  const int oh = mode == 2 ? h / 2 : 2 * h, ow = mode == 2 ? w / 2 : 2 * w;
  const int64_t total = (int64_t)n * oh * ow * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % cv);
    int64_t r = i / cv;
    const int ox = (int)(r % ow);
    r /= ow;
    const int oy = (int)(r % oh);
    const int img = (int)(r / oh);
    V o;
    if (mode == 0) {
      if ((oy & 1) || (ox & 1)) o.zero();
      else o.load(x + (((int64_t)img * h + (oy >> 1)) * w + (ox >> 1)) * c + k * 8);
    } else if (mode == 1) {
      o.load(x + (((int64_t)img * h + (oy >> 1)) * w + (ox >> 1)) * c + k * 8);
    } else {
      V a, b, d, e;
      const __nv_bfloat16* base = x + (((int64_t)img * h + 2 * oy) * w + 2 * ox) * c + k * 8;
      a.load(base);
      b.load(base + c);
      d.load(base + (int64_t)w * c);
      e.load(base + (int64_t)w * c + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) o.set(j, a.get(j) + b.get(j) + d.get(j) + e.get(j));
    }
    o.store(y + i * 8);
  }
}

// y = a + b (bf16, flat): gradient accumulation where two branches meet
__global__ void __launch_bounds__(256) add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                 __nv_bfloat16* __restrict__ y, int64_t nvec) {
  using V = Vec16<__nv_bfloat16>;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    V p, q, o;
    p.load(a + i * 8);
    q.load(b + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) o.set(j, p.get(j) + q.get(j));
    o.store(y + i * 8);
  }
}

// softmax backward on one block of rows: dS = P * (dP - sum_k(dP*P)) * scale, written as dS [rows][cols] and as its
// transpose dSt [cols][ldt] (column offset = first row of the block) for the GEMMs that need keys as the slow index.
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const __nv_bfloat16* __restrict__ p, const float* __restrict__ dp,
                                                         __nv_bfloat16* __restrict__ ds, __nv_bfloat16* __restrict__ dst, int64_t cols,
                                                         int64_t ldt, int64_t row0, float scale) {
  __shared__ float red[8];
  const int64_t row = blockIdx.x;
  const __nv_bfloat16* pr = p + row * cols;
  const float* dr = dp + row * cols;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < cols; i += 256) acc = fmaf(__bfloat162float(pr[i]), dr[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  for (int64_t i = threadIdx.x; i < cols; i += 256) {
    const float v = __bfloat162float(pr[i]) * (dr[i] - tot) * scale;
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    ds[row * cols + i] = b;
    if (dst) dst[i * ldt + row0 + row] = b;
  }
}

// out[row] = scale * <a[row], b[row]> (bf16 rows of `cols` elements): one warp per row, 16-byte loads
__global__ void __launch_bounds__(256) rowdot_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                    int64_t rows, int cols, int64_t ld_a, int64_t ld_b, float scale,
                                                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = blockIdx.x * 8ll + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * 8) {
    const uint4* ar = reinterpret_cast<const uint4*>(a + row * ld_a);
    const uint4* br = reinterpret_cast<const uint4*>(b + row * ld_b);
    float acc = 0.f;
    for (int i = lane; i < cols / 8; i += 32) {
      const uint4 u = ar[i], v = br[i];
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w}, vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc = fmaf(__uint_as_float(uw[j] << 16), __uint_as_float(vw[j] << 16), acc);
        acc = fmaf(__uint_as_float(uw[j] & 0xffff0000u), __uint_as_float(vw[j] & 0xffff0000u), acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[row] = acc * scale;
  }
}

// dgrad weights straight from the parameter: out[ci][tap'][co] = W[co][ci][taps-1-tap'], zero for co >= cout
__global__ void __launch_bounds__(256) pack_dgrad_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout,
                                                        int cin, int cout_pad, int taps, int64_t s_co, int64_t s_ci) {
  const int64_t total = (int64_t)cin * taps * cout_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout_pad);
    int64_t r = i / cout_pad;
    const int tap = (int)(r % taps);
    const int ci = (int)(r / taps);
    out[i] = co < cout ? w[co * s_co + ci * s_ci + (taps - 1 - tap)] : __float2bfloat16_rn(0.f);
  }
}

static inline dim3 train_grid(int64_t items, int n) {
  unsigned bx = (unsigned)((items + 255) / 256);
  if (bx > 2048) bx = 2048;
  if (bx < 1) bx = 1;
  return dim3(bx, (unsigned)n);
}

// ------------------------------------------------------------------------------------------
// Backward of GroupNorm(32) + SiLU (rv_groupnorm_stats / rv_groupnorm_silu; diffusers ResnetBlock2D.norm1/norm2,
// Attention.group_norm, conv_norm_out of the Flux AutoencoderKL).  With xh = (x - mean) * rstd, u = xh * gamma + beta,
// y = silu(u):   du = dy * silu'(u),  dgamma_c = sum du * xh,  dbeta_c = sum du,
//                dx = rstd * (du * gamma - S1 / m - xh * S2 / m),  S1 = sum_group du * gamma, S2 = sum_group du * gamma * xh.
// Pass 1 leaves per (sample, channel) the two sums (A = sum du, B = sum du * xh; fp32 per thread, fp64 across threads) --
// they give dbeta / dgamma AND, weighted by gamma over a group's channels, S1 / S2.  Pass 2 streams dx (+ the skip branch).
// Same block geometry as the forward kernels (the partition depends on H*W only).
// ------------------------------------------------------------------------------------------
// dy * silu'(u), u = xh * ga + be.  bf16 tensors: sigmoid through ONE tanh.approx (sg = (1 + t) / 2, t = tanh(u / 2)), so
//   silu'(u) = sg * (1 + u * (1 - sg)) = (1 + t + h * (1 - t * t)) / 2,  h = u / 2 = xh * gh + bh  (gh = ga / 2, bh = be / 2):
// one SFU op and five FMA-pipe ops, against expf + an IEEE divide (a ~10-instruction sequence) and seven more -- the two
// GroupNorm backward passes ran ALU-bound at 1.7-1.9 TB/s with those (is the same story in the forward).
// fp32 tensors (parity mode) keep the exact form.
// Both passes work from the raw x: h = x * ah + bh with ah = rstd * gamma / 2, bh = (beta - mean * rstd * gamma) / 2 -- two
// per-channel constants instead of four, and xh never materialises (pass 1 accumulates sum(du * x) and turns it into
// sum(du * xh) = rstd * sum(du * x) - mean * rstd * sum(du) in fp64 at the end; pass 2 folds xh into its own constants): the
// kernels drop from 127 to under 100 registers.
template <typename T>
__device__ __forceinline__ float gn_dsilu(float dy, float h) {
  if (sizeof(T) == 2) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    const float w = fmaf(-t, t, 1.0f);
    const float a = fmaf(h, w, t);
    const float dh = 0.5f * dy;
    return fmaf(dh, a, dh);
  }
  const float u = 2.0f * h;
  const float sg = 1.0f / (1.0f + expf(-u));
  return dy * sg * (1.0f + u * (1.0f - sg));
}

template <typename T, bool SILU>
__global__ void __launch_bounds__(256, 3) groupnorm_bwd_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                  const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, double* __restrict__ chan,
                                                                  int64_t hw, int c, int groups, int cpp, float eps,
                                                                  int64_t rows_per_block) {
  using V = Vec16<T>;
  extern __shared__ float sred[];  // [threads][2 * V::N]: every thread's partial sums, folded below without shared-memory atomics
  __shared__ double s_mr[64][2];   // per group: mean, rstd -- computed ONCE per block (fp64 divide / sqrt are subroutine calls:
                                   // done per thread and channel they were a sixth of the kernel's instructions)
  const int tid = threadIdx.x;
  const int n = blockIdx.y;
  const int col = tid % cpp;
  const int rows_per_iter = blockDim.x / cpp;
  const int cg = c / groups;
  const double cnt = (double)hw * (double)cg;
  for (int g = tid; g < groups; g += blockDim.x) {
    const double m = stats[((int64_t)n * groups + g) * 2] / cnt;
    double var = stats[((int64_t)n * groups + g) * 2 + 1] / cnt - m * m;
    if (var < 0.0) var = 0.0;
    s_mr[g][0] = m;
    s_mr[g][1] = 1.0 / sqrt(var + (double)eps);
  }
  __syncthreads();
  float ah[V::N], bh[V::N], a[V::N], b[V::N];  // a = sum du, b = sum du * x
#pragma unroll
  for (int j = 0; j < V::N; ++j) {
    const int ch = col * V::N + j, g = ch / cg;
    const float m = (float)s_mr[g][0], rs = (float)s_mr[g][1];
    const float ga = gamma[ch];
    ah[j] = 0.5f * rs * ga;
    bh[j] = 0.5f * (beta[ch] - m * rs * ga);
    a[j] = b[j] = 0.f;
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(hw, r0 + rows_per_block);
  const T* xb = x + (int64_t)n * hw * c + col * V::N;
  const T* gb = dy + (int64_t)n * hw * c + col * V::N;
  constexpr int U = 3;  // rows in flight per thread: 6 x 16-byte loads, three blocks per SM
  for (int64_t rb = r0 + tid / cpp; rb < r1; rb += (int64_t)U * rows_per_iter) {
    V vx[U], vg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * rows_per_iter;
      if (r < r1) {
        vx[u].load(xb + r * c);
        vg[u].load(gb + r * c);
      } else {
        vx[u].zero();
        vg[u].zero();
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < V::N; ++j) {
        const float xv = vx[u].get(j);
        const float du = SILU ? gn_dsilu<T>(vg[u].get(j), fmaf(xv, ah[j], bh[j])) : vg[u].get(j);
        a[j] += du;
        b[j] = fmaf(du, xv, b[j]);
      }
  }
  // one thread per channel adds the block's rows in fp64 (fp64 atomics on shared memory are compare-and-swap loops: 16 rows
  // colliding per address made this tail as long as the streaming loop) and issues the two global atomics
  float* my = sred + tid * (2 * V::N);
#pragma unroll
  for (int j = 0; j < V::N; ++j) {
    my[j] = a[j];
    my[V::N + j] = b[j];
  }
  __syncthreads();
  for (int ch = tid; ch < c; ch += blockDim.x) {
    const int g = ch / cg;
    const float* src = sred + (ch / V::N) * (2 * V::N) + (ch % V::N);
    double sa = 0.0, sb = 0.0;
    for (int r = 0; r < rows_per_iter; ++r) {
      sa += (double)src[(int64_t)r * cpp * (2 * V::N)];
      sb += (double)src[(int64_t)r * cpp * (2 * V::N) + V::N];
    }
    const double m = s_mr[g][0], rs = s_mr[g][1];
    atomicAdd(&chan[(int64_t)n * c * 2 + 2 * ch], sa);
    atomicAdd(&chan[(int64_t)n * c * 2 + 2 * ch + 1], rs * (sb - m * sa));  // sum du * xh
  }
}

template <typename T, bool SILU>
__global__ void __launch_bounds__(256, 3) groupnorm_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                 const T* __restrict__ add, const double* __restrict__ stats,
                                                                 const double* __restrict__ chan, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, T* __restrict__ dx,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t hw,
                                                                 int c, int groups, int cpp, float eps, int64_t rows_per_block) {
  using V = Vec16<T>;
  extern __shared__ float sgrp[];  // [groups][4]: S1 / m, S2 / m, mean, rstd
  const int tid = threadIdx.x;
  const int n = blockIdx.y;
  const int cg = c / groups;
  const double cnt = (double)hw * (double)cg;
  const double* ch_n = chan + (int64_t)n * c * 2;
  for (int g = tid; g < groups; g += blockDim.x) {
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < cg; ++k) {
      const int ch = g * cg + k;
      s1 += (double)gamma[ch] * ch_n[2 * ch];
      s2 += (double)gamma[ch] * ch_n[2 * ch + 1];
    }
    const double m = stats[((int64_t)n * groups + g) * 2] / cnt;
    double var = stats[((int64_t)n * groups + g) * 2 + 1] / cnt - m * m;
    if (var < 0.0) var = 0.0;
    sgrp[4 * g] = (float)(s1 / cnt);
    sgrp[4 * g + 1] = (float)(s2 / cnt);
    sgrp[4 * g + 2] = (float)m;
    sgrp[4 * g + 3] = (float)(1.0 / sqrt(var + (double)eps));
  }
  if (blockIdx.x == 0) {  // parameter gradients: one block per sample adds the sample's channel sums
    for (int ch = tid; ch < c; ch += blockDim.x) {
      atomicAdd(&dbeta[ch], (float)ch_n[2 * ch]);
      atomicAdd(&dgamma[ch], (float)ch_n[2 * ch + 1]);
    }
  }
  __syncthreads();
  const int col = tid % cpp;
  const int rows_per_iter = blockDim.x / cpp;
  // dx = rstd * (du * ga - m1 - xh * m2), xh = (x - mean) * rstd  =>  dx = du * rg - x * k2 - k1 with rg = rstd * ga,
  // k2 = rstd^2 * m2, k1 = rstd * m1 - mean * rstd^2 * m2
  float ah[V::N], bh[V::N], rg[V::N], k1[V::N], k2[V::N];
#pragma unroll
  for (int j = 0; j < V::N; ++j) {
    const int ch = col * V::N + j, g = ch / cg;
    const float m1 = sgrp[4 * g], m2 = sgrp[4 * g + 1], m = sgrp[4 * g + 2], rs = sgrp[4 * g + 3];
    const float ga = gamma[ch];
    ah[j] = 0.5f * rs * ga;
    bh[j] = 0.5f * (beta[ch] - m * rs * ga);
    rg[j] = rs * ga;
    k2[j] = rs * rs * m2;
    k1[j] = rs * m1 - m * k2[j];
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(hw, r0 + rows_per_block);
  const int64_t base = (int64_t)n * hw * c + col * V::N;
  constexpr int U = 2;  // rows in flight per thread: 4 (6 with the skip branch) x 16-byte loads, three blocks per SM
  for (int64_t rb = r0 + tid / cpp; rb < r1; rb += (int64_t)U * rows_per_iter) {
    V vx[U], vg[U], va[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * rows_per_iter;
      if (r < r1) {
        const int64_t off = base + r * c;
        vx[u].load(x + off);
        vg[u].load(dy + off);
        if (add) va[u].load(add + off);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + (int64_t)u * rows_per_iter;
      if (r < r1) {
        V o;
#pragma unroll
        for (int j = 0; j < V::N; ++j) {
          const float xv = vx[u].get(j);
          const float du = SILU ? gn_dsilu<T>(vg[u].get(j), fmaf(xv, ah[j], bh[j])) : vg[u].get(j);
          float d = fmaf(-xv, k2[j], fmaf(du, rg[j], -k1[j]));
          if (add) d += va[u].get(j);
          o.set(j, d);
        }
        o.store(dx + base + r * c);
      }
    }
  }
}

struct GnBwdGeom {
  int cpp, block;
  int64_t rows_per_block;
  unsigned blocks_x;
};
template <typename T>
static int gn_bwd_geometry(int64_t hw, int c, int groups, GnBwdGeom* g) {  // = gn_geometry of rv_elementwise.cu
  constexpr int VN = Vec16<T>::N;
  RV_CHECK_ARG(groups > 0 && groups <= 64 && c % groups == 0 && c % VN == 0, "groupnorm_bwd: bad channel/group count %d/%d", c, groups);
  g->cpp = c / VN;
  RV_CHECK_ARG(g->cpp <= 256, "groupnorm_bwd: too many channels (%d)", c);
  int rows = 256 / g->cpp;
  if (rows < 1) rows = 1;
  g->block = rows * g->cpp;
  int64_t rpt = hw / ((int64_t)rows * 296);  // as gn_geometry: fewer rows per thread for small tensors
  rpt = rpt > 64 ? 64 : (rpt < 8 ? 8 : rpt / 8 * 8);
  g->rows_per_block = (int64_t)rows * rpt;
  g->blocks_x = (unsigned)((hw + g->rows_per_block - 1) / g->rows_per_block);
  return 0;
}

template <typename T>
static int groupnorm_bwd_launch(const void* x, const void* dy, const void* add, const double* stats, const float* gamma,
                                const float* beta, void* dx, float* dgamma, float* dbeta, double* chan, int n, int64_t hw, int c,
                                int groups, float eps, int apply_silu, cudaStream_t st) {
  GnBwdGeom g;
  if (int rc = gn_bwd_geometry<T>(hw, c, groups, &g)) return rc;
  RV_CUDA(cudaMemsetAsync(chan, 0, sizeof(double) * 2 * (size_t)n * c, st));
  const double bytes = (double)n * hw * c * sizeof(T);
  {
    LaunchScope scope(CAT_NORM, st, 2.0 * bytes);
    const size_t smem = sizeof(float) * 2 * (16 / sizeof(T)) * g.block;
    if (apply_silu)
      groupnorm_bwd_reduce_kernel<T, true><<<dim3(g.blocks_x, n), g.block, smem, st>>>((const T*)x, (const T*)dy, stats, gamma, beta, chan, hw,
                                                                                     c, groups, g.cpp, eps, g.rows_per_block);
    else
      groupnorm_bwd_reduce_kernel<T, false><<<dim3(g.blocks_x, n), g.block, smem, st>>>((const T*)x, (const T*)dy, stats, gamma, beta, chan, hw,
                                                                                      c, groups, g.cpp, eps, g.rows_per_block);
    RV_LAUNCH_CHECK();
  }
  {
    LaunchScope scope(CAT_NORM, st, (add ? 4.0 : 3.0) * bytes);
    const size_t smem = sizeof(float) * 4 * groups;
    if (apply_silu)
      groupnorm_bwd_apply_kernel<T, true><<<dim3(g.blocks_x, n), g.block, smem, st>>>((const T*)x, (const T*)dy, (const T*)add, stats, chan, gamma,
                                                                                    beta, (T*)dx, dgamma, dbeta, hw, c, groups, g.cpp, eps,
                                                                                    g.rows_per_block);
    else
      groupnorm_bwd_apply_kernel<T, false><<<dim3(g.blocks_x, n), g.block, smem, st>>>((const T*)x, (const T*)dy, (const T*)add, stats, chan,
                                                                                     gamma, beta, (T*)dx, dgamma, dbeta, hw, c, groups, g.cpp,
                                                                                     eps, g.rows_per_block);
    RV_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace rv

extern "C" {

int rv_recon_loss_bwd(const void* pred, const void* target, const float* eb_host, const float* eb2_host, int naive_mse,
                      float grad_scale, float clamp_lo, float clamp_hi, void* dpred, int n, int64_t hw, int dtype, void* stream) {
  RV_CHECK_ARG(pred && target && dpred && n > 0 && hw > 0, "recon_loss_bwd: bad argument");
  RV_CHECK_ARG(naive_mse || (eb_host && eb2_host), "recon_loss_bwd: Eb / Eb2 missing");
  rv::LossBwdParams lp;
  for (int c = 0; c < 3; ++c) {
    lp.eb[c] = eb_host ? eb_host[c] : 0.f;
    lp.eb2[c] = eb2_host ? eb2_host[c] : 0.f;
  }
  lp.naive = naive_mse;
  lp.scale = grad_scale;
  lp.clamp_lo = clamp_lo;
  lp.clamp_hi = clamp_hi;
  cudaStream_t st = (cudaStream_t)stream;
  rv::LaunchScope scope(rv::CAT_LOSS, st, 12.0 * n * hw * (dtype == RV_F32 ? 4 : 2));
  if (dtype == RV_F32)
    rv::recon_loss_bwd_kernel<float><<<rv::train_grid(hw, n), 256, 0, st>>>((const float*)pred, (const float*)target, (float*)dpred, hw, lp);
  else if (dtype == RV_BF16)
    rv::recon_loss_bwd_kernel<__nv_bfloat16><<<rv::train_grid(hw, n), 256, 0, st>>>((const __nv_bfloat16*)pred, (const __nv_bfloat16*)target,
                                                                                (__nv_bfloat16*)dpred, hw, lp);
  else RV_CHECK_ARG(false, "recon_loss_bwd: bad dtype %d", dtype);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_reparam_bwd(const void* moments, const void* noise, const void* dz, void* dmoments, int n, int zc, int64_t hw, int dtype,
                   float kl_weight, void* stream) {
  RV_CHECK_ARG(moments && dmoments && n > 0 && zc > 0 && hw > 0, "reparam_bwd: bad argument");
  RV_CHECK_ARG((dz == nullptr) == (noise == nullptr), "reparam_bwd: dz and noise must be given together");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t per = (int64_t)zc * hw;
  rv::LaunchScope scope(rv::CAT_REPARAM, st, 6.0 * n * per * (dtype == RV_F32 ? 4 : 2));
  if (dtype == RV_F32)
    rv::reparam_bwd_kernel<float><<<rv::train_grid(per, n), 256, 0, st>>>((const float*)moments, (const float*)noise, (const float*)dz,
                                                                       (float*)dmoments, zc, hw, kl_weight);
  else if (dtype == RV_BF16)
    rv::reparam_bwd_kernel<__nv_bfloat16><<<rv::train_grid(per, n), 256, 0, st>>>((const __nv_bfloat16*)moments, (const __nv_bfloat16*)noise,
                                                                               (const __nv_bfloat16*)dz, (__nv_bfloat16*)dmoments, zc, hw,
                                                                               kl_weight);
  else RV_CHECK_ARG(false, "reparam_bwd: bad dtype %d", dtype);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_kl_ref(const void* moments, const void* ref_moments, float* kl_out, void* dmoments, int n, int zc, int64_t hw, int dtype,
              float weight, void* stream) {
  RV_CHECK_ARG(moments && ref_moments && kl_out && n > 0 && zc > 0 && hw > 0, "kl_ref: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t per = (int64_t)zc * hw;
  rv::LaunchScope scope(rv::CAT_REPARAM, st, (dmoments ? 6.0 : 4.0) * n * per * (dtype == RV_F32 ? 4 : 2));
  if (dtype == RV_F32)
    rv::kl_ref_kernel<float><<<rv::train_grid(per, n), 256, 0, st>>>((const float*)moments, (const float*)ref_moments, kl_out, (float*)dmoments,
                                                                  zc, hw, weight);
  else if (dtype == RV_BF16)
    rv::kl_ref_kernel<__nv_bfloat16><<<rv::train_grid(per, n), 256, 0, st>>>((const __nv_bfloat16*)moments, (const __nv_bfloat16*)ref_moments,
                                                                          kl_out, (__nv_bfloat16*)dmoments, zc, hw, weight);
  else RV_CHECK_ARG(false, "kl_ref: bad dtype %d", dtype);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_rmsnorm_silu_bwd(const void* x, const float* gamma_scaled, const void* dy, const void* add, void* dx, float* dgamma,
                        float dgamma_scale, int64_t pixels, int c, int dtype, int apply_silu, void* stream) {
  RV_CHECK_ARG(x && gamma_scaled && dy && dx && dgamma && pixels > 0, "rmsnorm_silu_bwd: bad argument");
  RV_CHECK_ARG(dtype == RV_F32 || dtype == RV_BF16, "rmsnorm_silu_bwd: bad dtype %d", dtype);
  const int per_chunk = dtype == RV_F32 ? 4 : 8;
  const int lanes = c / (3 * per_chunk);
  RV_CHECK_ARG(c % (3 * per_chunk) == 0 && lanes >= 1 && lanes <= 32 && (lanes & (lanes - 1)) == 0,
               "rmsnorm_silu_bwd: channels must be 3 * 2^k 16-byte chunks (96, 192, 384 ...), got %d", c);
  RV_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)dx % 16 == 0) && ((uintptr_t)add % 16 == 0),
               "rmsnorm_silu_bwd: unaligned tensor");
  cudaStream_t st = (cudaStream_t)stream;
  const int ppw = 32 / lanes;
  int64_t blocks = (pixels + 8 * ppw * 4 - 1) / (8 * ppw * 4);
  const int64_t cap = (int64_t)rv::num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const size_t smem = 8 * (size_t)c * sizeof(float);
  rv::LaunchScope scope(rv::CAT_NORM, st, 3.0 * pixels * c * (dtype == RV_F32 ? 4 : 2));
#define RV_NB(T, S) \
  rv::rmsnorm_silu_bwd_kernel<T, S><<<(unsigned)blocks, 256, smem, st>>>((const T*)x, gamma_scaled, (const T*)dy, (const T*)add, (T*)dx, dgamma, pixels, lanes, dgamma_scale)
  if (dtype == RV_F32) {
    if (apply_silu) RV_NB(float, true);
    else RV_NB(float, false);
  } else {
    if (apply_silu) RV_NB(__nv_bfloat16, true);
    else RV_NB(__nv_bfloat16, false);
  }
#undef RV_NB
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_pack_dgrad_weights(const void* w, int64_t w_co_stride, int64_t w_ci_stride, void* out, int cout, int cin, int cout_pad,
                          int ksize, void* stream) {
  RV_CHECK_ARG(w && out && cout > 0 && cin > 0 && cout_pad >= cout && (ksize == 1 || ksize == 3), "pack_dgrad_weights: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int taps = ksize * ksize;
  const int64_t total = (int64_t)cin * taps * cout_pad;
  int64_t blocks = (total + 255) / 256;
  if (blocks > rv::num_sms() * 8) blocks = rv::num_sms() * 8;
  rv::LaunchScope scope(rv::CAT_LAYOUT, st, 4.0 * total);
  rv::pack_dgrad_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)w, (__nv_bfloat16*)out, cout, cin, cout_pad, taps,
                                                         w_co_stride, w_ci_stride);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_resample2x(const void* x, void* y, int n, int h, int w, int c, int mode, void* stream) {
  RV_CHECK_ARG(x && y && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "resample2x: bad argument (c %% 8)");
  RV_CHECK_ARG(mode >= 0 && mode <= 2 && (mode != 2 || (h % 2 == 0 && w % 2 == 0)), "resample2x: bad mode / odd size");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t out_el = mode == 2 ? (int64_t)n * (h / 2) * (w / 2) * c : (int64_t)n * 4 * h * w * c;
  int64_t blocks = (out_el / 8 + 255) / 256;
  if (blocks > rv::num_sms() * 32) blocks = rv::num_sms() * 32;
  if (blocks < 1) blocks = 1;
  rv::LaunchScope scope(rv::CAT_LAYOUT, st, 2.0 * ((double)n * h * w * c + out_el));
  rv::resample2x_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w, c, mode);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_add_bf16(const void* a, const void* b, void* y, int64_t n, void* stream) {
  RV_CHECK_ARG(a && b && y && n > 0 && n % 8 == 0, "add_bf16: element count must be a positive multiple of 8");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = (n / 8 + 255) / 256;
  if (blocks > rv::num_sms() * 32) blocks = rv::num_sms() * 32;
  rv::LaunchScope scope(rv::CAT_LAYOUT, st, 6.0 * n);
  rv::add_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, n / 8);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_softmax_bwd(const void* p, const float* dp, void* ds, void* ds_t, int64_t rows, int64_t cols, int64_t ld_t, int64_t row0,
                   float scale, void* stream) {
  RV_CHECK_ARG(p && dp && ds && rows > 0 && cols > 0 && (!ds_t || ld_t >= row0 + rows), "softmax_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  rv::LaunchScope scope(rv::CAT_SOFTMAX, st, (double)rows * cols * 10.0);
  rv::softmax_bwd_kernel<<<(unsigned)rows, 256, 0, st>>>((const __nv_bfloat16*)p, dp, (__nv_bfloat16*)ds, (__nv_bfloat16*)ds_t, cols, ld_t,
                                                         row0, scale);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_rowdot(const void* a, const void* b, int64_t rows, int cols, int64_t ld_a, int64_t ld_b, float scale, float* out,
              void* stream) {
  RV_CHECK_ARG(a && b && out && rows > 0 && cols > 0 && cols % 8 == 0 && ld_a % 8 == 0 && ld_b % 8 == 0 &&
                   (uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0,
               "rowdot: bad argument (bf16 rows, 16-byte aligned, cols %% 8 == 0)");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  rv::LaunchScope scope(rv::CAT_SOFTMAX, st, (double)rows * cols * 4.0);
  rv::rowdot_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, rows, cols, ld_a, ld_b, scale,
                                                     out);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_grad_sqnorm_scratch_bytes(void) { return 2048 * (int)sizeof(double); }

int rv_grad_sqnorm(const float* g, int64_t n, float* out, void* scratch, void* stream) {
  RV_CHECK_ARG(g && out && scratch && n > 0, "grad_sqnorm: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 2048) blocks = 2048;
  {
    rv::LaunchScope scope(rv::CAT_LAYOUT, st, 4.0 * n);
    rv::sqnorm_kernel<<<(unsigned)blocks, 256, 0, st>>>(g, n, (double*)scratch);
    RV_LAUNCH_CHECK();
  }
  {
    rv::LaunchScope scope(rv::CAT_LAYOUT, st, 8.0 * blocks);
    rv::sqnorm_finish_kernel<<<1, 256, 0, st>>>((const double*)scratch, (int)blocks, out);
    RV_LAUNCH_CHECK();
  }
  return 0;
}

int rv_adamw_advance(float* state, float beta1, float beta2, void* stream) {
  RV_CHECK_ARG(state != nullptr, "adamw_advance: null state");
  cudaStream_t st = (cudaStream_t)stream;
  rv::LaunchScope scope(rv::CAT_LAYOUT, st, 12.0);
  rv::adamw_advance_kernel<<<1, 32, 0, st>>>(state, beta1, beta2);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, const float* state, float grad_scale, const float* sqnorm,
                  float max_norm, void* stream) {
  RV_CHECK_ARG(p && g && m && v && n > 0 && (step >= 1 || state != nullptr), "adamw_step: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const float bc1 = 1.0f - powf(beta1, (float)(step > 0 ? step : 1)), bc2 = 1.0f - powf(beta2, (float)(step > 0 ? step : 1));
  int64_t blocks = (n + 256 * 4 - 1) / (256 * 4);
  if (blocks > 4096) blocks = 4096;
  rv::LaunchScope scope(rv::CAT_LAYOUT, st, 28.0 * n);
  rv::adamw_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, (__nv_bfloat16*)p_bf16, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2,
                                                     grad_scale, sqnorm, max_norm, state);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_groupnorm_silu_bwd(const void* x, const double* stats, const float* gamma, const float* beta, const void* dy, const void* add,
                          void* dx, float* dgamma, float* dbeta, double* scratch, int n, int64_t hw, int c, int groups, float eps,
                          int dtype, int apply_silu, void* stream) {
  RV_CHECK_ARG(x && stats && gamma && beta && dy && dx && dgamma && dbeta && scratch && n > 0 && hw > 0, "groupnorm_silu_bwd: bad argument");
  RV_CHECK_ARG(dtype == RV_F32 || dtype == RV_BF16, "groupnorm_silu_bwd: bad dtype %d", dtype);
  RV_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)dx % 16 == 0) && ((uintptr_t)add % 16 == 0),
               "groupnorm_silu_bwd: unaligned tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RV_F32)
    return rv::groupnorm_bwd_launch<float>(x, dy, add, stats, gamma, beta, dx, dgamma, dbeta, scratch, n, hw, c, groups, eps, apply_silu, st);
  return rv::groupnorm_bwd_launch<__nv_bfloat16>(x, dy, add, stats, gamma, beta, dx, dgamma, dbeta, scratch, n, hw, c, groups, eps,
                                                 apply_silu, st);
}

}  // extern "C"
