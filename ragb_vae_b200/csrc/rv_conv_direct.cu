// CUDA-core implicit-GEMM convolution (fp32 accumulate), any layout / dtype / stride.
//
// Replaces the torch.nn.functional.conv2d calls diffusers issues inside vae.encode / vae.decode
// (reference call sites: src/models/rgba_vae.py:277,279) for the layers the tensor-core kernel
// does not take: the 4-channel edge layers, NCHW boundary tensors, and the whole network in the
// fp32 parity mode (config c1).  It is also the in-library cross-check of rv_conv2d_tc.
//
// Tiling: a block computes 64 output pixels x 64 output channels; K runs over (tap, cin) in
// chunks of 16 staged through shared memory; each of the 256 threads owns a 4x4 micro-tile.
#include "rv_common.cuh"

namespace rv {

constexpr int DM = 64, DN = 64, DK = 16;

struct DirectParams {
  rv_conv_desc d;
  int64_t m_total;  // n*oh*ow
  int heff, weff;   // input extent seen by the taps (doubled when upsample)
};

template <typename TX, bool X_NCHW>
__device__ __forceinline__ float load_x(const TX* x, const rv_conv_desc& d, int n, int iy, int ix, int c) {
  if (X_NCHW) return ldf(x + (((int64_t)n * d.cin + c) * d.h + iy) * d.w + ix);
  return ldf(x + (((int64_t)n * d.h + iy) * d.w + ix) * d.x_cstride + c);
}

template <typename TX, typename TY, bool X_NCHW>
__global__ void __launch_bounds__(256) conv_direct_kernel(const DirectParams p, const TX* __restrict__ x,
                                                         const float* __restrict__ w,
                                                         const float* __restrict__ bias,
                                                         const TY* __restrict__ residual, TY* __restrict__ y) {
  const rv_conv_desc& d = p.d;
  __shared__ float As[DK][DM + 4];
  __shared__ float Bs[DK][DN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * DM;
  const int n0 = blockIdx.y * DN;
  const int taps = d.ksize * d.ksize;
  const int ktot = taps * d.cin;

  // A-load mapping: NHWC -> threads adjacent along channels; NCHW -> adjacent along pixels.
  int a_p[4], a_c[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (X_NCHW) {
      a_p[i] = tid % DM;
      a_c[i] = tid / DM + 4 * i;
    } else {
      a_p[i] = tid / DK + 16 * i;
      a_c[i] = tid % DK;
    }
  }
  int a_n[4], a_y[4], a_x[4];
  bool a_ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + a_p[i];
    a_ok[i] = m < p.m_total;
    int64_t mm = a_ok[i] ? m : 0;
    a_x[i] = (int)(mm % d.ow) * d.stride - d.pad_lo;
    int64_t t = mm / d.ow;
    a_y[i] = (int)(t % d.oh) * d.stride - d.pad_lo;
    a_n[i] = (int)(t / d.oh);
  }
  // B-load mapping: threads adjacent along k (contiguous in the [cout][taps][cin] matrix).
  const int b_k = tid % DK;
  const int b_n = tid / DK;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;

  for (int tap = 0; tap < taps; ++tap) {
    const int dy = tap / d.ksize, dx = tap % d.ksize;
    for (int c0 = 0; c0 < d.cin; c0 += DK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        int c = c0 + a_c[i];
        int uy = a_y[i] + dy, ux = a_x[i] + dx;
        if (a_ok[i] && c < d.cin && uy >= 0 && uy < p.heff && ux >= 0 && ux < p.weff) {
          int iy = d.upsample ? (uy >> 1) : uy, ix = d.upsample ? (ux >> 1) : ux;
          v = load_x<TX, X_NCHW>(x, d, a_n[i], iy, ix, c) * d.in_scale + d.in_shift;
        }
        As[a_c[i]][a_p[i]] = v;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int n = n0 + b_n + 16 * i;
        int c = c0 + b_k;
        float v = 0.f;
        if (n < d.cout && c < d.cin) v = w[(int64_t)n * ktot + tap * d.cin + c];
        Bs[b_k][b_n + 16 * i] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < DK; ++k) {
        float4 a = *reinterpret_cast<const float4*>(&As[k][tm]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  const int64_t ohw = (int64_t)d.oh * d.ow;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + tm + i;
    if (m >= p.m_total) continue;
    int64_t n_img = m / ohw, pix = m % ohw;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = n0 + tn + j;
      if (co >= d.cout) continue;
      float v = acc[i][j] * d.alpha;
      if (d.bias_mode == 1) v += bias[co];
      else if (d.bias_mode == 2) v += bias[m];
      int64_t off = d.y_nchw ? ((n_img * d.cout + co) * ohw + pix) : (m * d.y_cstride + co);
      if (residual) v += ldf(residual + off);
      v = v * d.out_scale + d.out_shift;
      if (d.clamp) v = fminf(fmaxf(v, d.clamp_lo), d.clamp_hi);
      stf(y + off, v);
    }
  }
}

template <typename TX, typename TY>
static int launch_direct(const DirectParams& p, const void* x, const float* w, const float* bias,
                         const void* residual, void* y, cudaStream_t st) {
  dim3 grid((unsigned)((p.m_total + DM - 1) / DM), (unsigned)((p.d.cout + DN - 1) / DN));
  double flops = 2.0 * (double)p.m_total * p.d.cout * p.d.cin * p.d.ksize * p.d.ksize;
  LaunchScope scope(CAT_CONV_DIRECT, st, flops);
  if (p.d.x_nchw)
    conv_direct_kernel<TX, TY, true><<<grid, 256, 0, st>>>(p, (const TX*)x, w, bias, (const TY*)residual, (TY*)y);
  else
    conv_direct_kernel<TX, TY, false><<<grid, 256, 0, st>>>(p, (const TX*)x, w, bias, (const TY*)residual, (TY*)y);
  RV_LAUNCH_CHECK();
  return 0;
}

int check_conv_desc(const rv_conv_desc* d) {
  RV_CHECK_ARG(d != nullptr, "conv: null descriptor");
  RV_CHECK_ARG(d->n > 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0, "conv: non-positive dimension");
  RV_CHECK_ARG(d->ksize == 1 || d->ksize == 3, "conv: ksize must be 1 or 3 (got %d)", d->ksize);
  RV_CHECK_ARG(d->stride == 1 || d->stride == 2, "conv: stride must be 1 or 2 (got %d)", d->stride);
  RV_CHECK_ARG(!(d->upsample && d->stride != 1), "conv: upsample requires stride 1");
  int heff = d->upsample ? 2 * d->h : d->h, weff = d->upsample ? 2 * d->w : d->w;
  // 3x3 stride 1 is always 'same' (pad_lo + pad_hi = 2; pad_lo = 2 is the data gradient of the stride-2 conv)
  RV_CHECK_ARG(d->pad_lo >= 0 && d->pad_lo <= (d->ksize == 3 ? 2 : 0), "conv: bad pad_lo %d", d->pad_lo);
  int pad_hi = (d->ksize == 3 && d->stride == 1) ? 2 - d->pad_lo : (d->ksize == 3 ? 1 - d->pad_lo : 0);
  RV_CHECK_ARG(!d->taps_1d || (d->ksize == 3 && d->stride == 1 && !d->upsample && d->pad_lo == 1),
               "conv: taps_1d needs a stride-1 3-tap kernel with pad 1");
  int oh = (heff + d->pad_lo + pad_hi - d->ksize) / d->stride + 1;
  int ow = d->taps_1d ? weff : (weff + d->pad_lo + pad_hi - d->ksize) / d->stride + 1;
  RV_CHECK_ARG(oh == d->oh && ow == d->ow, "conv: output size %dx%d inconsistent with input (expected %dx%d)",
               d->oh, d->ow, oh, ow);
  RV_CHECK_ARG((d->x_dtype == RV_F32 || d->x_dtype == RV_BF16) && (d->y_dtype == RV_F32 || d->y_dtype == RV_BF16),
               "conv: bad dtype code");
  RV_CHECK_ARG(d->x_nchw || d->x_cstride >= d->cin, "conv: x_cstride < cin");
  RV_CHECK_ARG(d->y_nchw || d->y_cstride >= d->cout, "conv: y_cstride < cout");
  RV_CHECK_ARG(d->bias_mode >= 0 && d->bias_mode <= 2, "conv: bad bias_mode");
  return 0;
}

}  // namespace rv

extern "C" int rv_conv2d_direct(const rv_conv_desc* d, const void* x, const float* w, const float* bias,
                                const void* residual, void* y, void* stream) {
  if (int rc = rv::check_conv_desc(d)) return rc;
  RV_CHECK_ARG(x && w && y, "conv_direct: null tensor");
  RV_CHECK_ARG(!d->taps_1d, "conv_direct: taps_1d is a tensor-core path layout");
  RV_CHECK_ARG(d->bias_mode == 0 || bias, "conv_direct: bias_mode set but bias is null");
  rv::DirectParams p;
  p.d = *d;
  p.m_total = (int64_t)d->n * d->oh * d->ow;
  p.heff = d->upsample ? 2 * d->h : d->h;
  p.weff = d->upsample ? 2 * d->w : d->w;
  cudaStream_t st = (cudaStream_t)stream;
  if (d->x_dtype == RV_F32 && d->y_dtype == RV_F32) return rv::launch_direct<float, float>(p, x, w, bias, residual, y, st);
  if (d->x_dtype == RV_F32 && d->y_dtype == RV_BF16)
    return rv::launch_direct<float, __nv_bfloat16>(p, x, w, bias, residual, y, st);
  if (d->x_dtype == RV_BF16 && d->y_dtype == RV_F32)
    return rv::launch_direct<__nv_bfloat16, float>(p, x, w, bias, residual, y, st);
  return rv::launch_direct<__nv_bfloat16, __nv_bfloat16>(p, x, w, bias, residual, y, st);
}

namespace rv {
__global__ void pack_direct_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int taps) {
  int64_t total = (int64_t)cout * cin * taps;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % cin);
    int64_t r = i / cin;
    int t = (int)(r % taps);
    int co = (int)(r / taps);
    out[i] = w[((int64_t)co * cin + c) * taps + t];
  }
}
}  // namespace rv

extern "C" int rv_pack_conv_weights_direct(const float* w, int cout, int cin, int ksize, float* out, void* stream) {
  RV_CHECK_ARG(w && out && cout > 0 && cin > 0 && (ksize == 1 || ksize == 3), "pack_direct: bad argument");
  int taps = ksize * ksize;
  int64_t total = (int64_t)cout * cin * taps;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  rv::LaunchScope scope(rv::CAT_LAYOUT, (cudaStream_t)stream, 8.0 * total);
  rv::pack_direct_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, out, cout, cin, taps);
  RV_LAUNCH_CHECK();
  return 0;
}
