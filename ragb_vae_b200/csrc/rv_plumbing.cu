// The data-format steps either side of the VAE: triplet detail augmentation in front of the
// encoder, Flux latent 2x2 patchify / un-patchify, and uint8 RGBA ingest / egress.  All HBM-bound, one pass each.
#include "rv_common.cuh"

namespace rv {

// build_detail_augmented_triplet (reference src/training/rgba_vae_stage.py:606-625): target (B,4,H,W) in [-1,1] ->
// (3B,4,H,W) = [target | target*fg - bg with alpha 1 | target*fg + bg with alpha 1], fg=(1+a)/2, bg=(1-a)/2.
template <typename T>
__global__ void __launch_bounds__(256) triplet_kernel(const T* __restrict__ x, T* __restrict__ y, int b, int64_t hw) {
  const int n = blockIdx.y;
  const T* xp = x + (int64_t)n * 4 * hw;
  T* y0 = y + (int64_t)n * 4 * hw;
  T* y1 = y + (int64_t)(b + n) * 4 * hw;
  T* y2 = y + (int64_t)(2 * b + n) * 4 * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = ldf(xp + 3 * hw + i);
    const float fg = (1.0f + a) * 0.5f, bg = (1.0f - a) * 0.5f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = ldf(xp + c * hw + i);
      // the reference rounds target*fg to the tensor dtype before the +/- bg
      float m = v * fg;
      T mr;
      stf(&mr, m);
      m = ldf(&mr);
      stf(y0 + c * hw + i, v);
      stf(y1 + c * hw + i, m - bg);
      stf(y2 + c * hw + i, m + bg);
    }
    stf(y0 + 3 * hw + i, a);
    stf(y1 + 3 * hw + i, 1.0f);
    stf(y2 + 3 * hw + i, 1.0f);
  }
}

// RandomBackgroundBlend._blend_tensor (reference src/training/rgba_vae_stage.py:85-130), batched on the device: samples with
// mask[n] != 0 become rgb*a + colour[n]*(1-a) with alpha 1, the others are copied.  Each product and the sum are rounded to
// the tensor dtype like the reference's elementwise ops.
template <typename T>
__global__ void __launch_bounds__(256) background_blend_kernel(const T* __restrict__ x, const float* __restrict__ colors,
                                                              const unsigned char* __restrict__ mask, T* __restrict__ y, int64_t hw) {
  const int n = blockIdx.y;
  const T* xp = x + (int64_t)n * 4 * hw;
  T* yp = y + (int64_t)n * 4 * hw;
  const bool on = mask[n] != 0;
  float col[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    T cr;
    stf(&cr, colors[n * 3 + c]);  // the reference draws the colour in the tensor dtype
    col[c] = ldf(&cr);
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = ldf(xp + 3 * hw + i);
    T om;
    stf(&om, __fsub_rn(1.0f, a));
    const float one_minus = ldf(&om);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = ldf(xp + c * hw + i);
      if (on) {
        T p0, p1;
        stf(&p0, __fmul_rn(v, a));
        stf(&p1, __fmul_rn(col[c], one_minus));
        stf(yp + c * hw + i, __fadd_rn(ldf(&p0), ldf(&p1)));
      } else {
        stf(yp + c * hw + i, v);
      }
    }
    stf(yp + 3 * hw + i, on ? 1.0f : a);
  }
}

// FluxPipeline._pack_latents: (B,C,h,w) -> (B,(h/2)(w/2),4C), token = (h/2 index, w/2 index), feature = (c, dy, dx);
// optional affine (z - shift) * scale on the way (src/models/flux_kontext_textalpha.py:330-340).
template <typename T>
__global__ void __launch_bounds__(256) pack_latents_kernel(const T* __restrict__ x, T* __restrict__ y, int c, int h, int w,
                                                          float shift, float scale, int unpack) {
  const int n = blockIdx.y;
  const int64_t per = (int64_t)c * h * w;
  const int h2 = h / 2, w2 = w / 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
    // i enumerates the PACKED layout: ((ty*w2 + tx) * c + ch) * 4 + dy*2 + dx
    const int dx = (int)(i & 1), dy = (int)((i >> 1) & 1);
    int64_t r = i >> 2;
    const int ch = (int)(r % c);
    r /= c;
    const int tx = (int)(r % w2), ty = (int)(r / w2);
    const int64_t plain = ((int64_t)ch * h + (2 * ty + dy)) * w + (2 * tx + dx);
    if (!unpack) stf(y + n * per + i, (ldf(x + n * per + plain) - shift) * scale);
    else stf(y + n * per + plain, ldf(x + n * per + i) * scale + shift);
  }
  (void)h2;
}

// uint8 RGBA HWC (PIL layout, inference_rgba_flux.py:15-20) -> NCHW T: v/255 * scale + shift
template <typename T>
__global__ void __launch_bounds__(256) u8_to_nchw_kernel(const uint8_t* __restrict__ x, T* __restrict__ y, int64_t hw, float scale,
                                                        float shift) {
  const int n = blockIdx.y;
  const uchar4* xp = reinterpret_cast<const uchar4*>(x) + (int64_t)n * hw;
  T* yp = y + (int64_t)n * 4 * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    const uchar4 p = xp[i];
    stf(yp + i, (float)p.x / 255.0f * scale + shift);
    stf(yp + hw + i, (float)p.y / 255.0f * scale + shift);
    stf(yp + 2 * hw + i, (float)p.z / 255.0f * scale + shift);
    stf(yp + 3 * hw + i, (float)p.w / 255.0f * scale + shift);
  }
}

// NCHW T in [0,1] -> uint8 RGBA HWC: (clamp(v,0,1) * 255) truncated (inference_rgba_flux.py:23-26)
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_u8_kernel(const T* __restrict__ x, uint8_t* __restrict__ y, int64_t hw) {
  const int n = blockIdx.y;
  const T* xp = x + (int64_t)n * 4 * hw;
  uchar4* yp = reinterpret_cast<uchar4*>(y) + (int64_t)n * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    uchar4 p;
    p.x = (uint8_t)(fminf(fmaxf(ldf(xp + i), 0.f), 1.f) * 255.0f);
    p.y = (uint8_t)(fminf(fmaxf(ldf(xp + hw + i), 0.f), 1.f) * 255.0f);
    p.z = (uint8_t)(fminf(fmaxf(ldf(xp + 2 * hw + i), 0.f), 1.f) * 255.0f);
    p.w = (uint8_t)(fminf(fmaxf(ldf(xp + 3 * hw + i), 0.f), 1.f) * 255.0f);
    yp[i] = p;
  }
}

// diffusers blend_v / blend_h (AutoencoderKL tiling): the first `extent` rows (vertical) or columns of every plane of
// b are replaced by a linear ramp between the LAST `extent` rows / columns of a and b itself:
// b[k] = a[-extent + k] * (1 - k/extent) + b[k] * (k/extent).  a: [planes][ah][aw], b: [planes][bh][bw].
template <typename T>
__global__ void __launch_bounds__(256) blend_kernel(const T* __restrict__ a, T* __restrict__ b, int ah, int aw, int bh, int bw,
                                                   int extent, int vertical) {
  const int plane = blockIdx.y;
  const T* ap = a + (int64_t)plane * ah * aw;
  T* bp = b + (int64_t)plane * bh * bw;
  const int other = vertical ? min(aw, bw) : min(ah, bh);  // extent of the non-blended axis that both tiles share
  const int64_t total = (int64_t)extent * other;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int k, o;
    if (vertical) { k = (int)(i / other); o = (int)(i % other); }
    else { o = (int)(i / extent); k = (int)(i % extent); }
    const float wgt = (float)k / (float)extent;
    const int64_t ai = vertical ? ((int64_t)(ah - extent + k) * aw + o) : ((int64_t)o * aw + (aw - extent + k));
    const int64_t bi = vertical ? ((int64_t)k * bw + o) : ((int64_t)o * bw + k);
    stf(bp + bi, ldf(ap + ai) * (1.0f - wgt) + ldf(bp + bi) * wgt);
  }
}

static inline dim3 grid_for(int64_t items, int n) {
  unsigned bx = (unsigned)((items + 255) / 256);
  if (bx > 2048) bx = 2048;
  if (bx < 1) bx = 1;
  return dim3(bx, (unsigned)n);
}

}  // namespace rv

extern "C" {

int rv_triplet_augment(const void* target, void* out, int b, int64_t hw, int dtype, void* stream) {
  RV_CHECK_ARG(target && out && b > 0 && hw > 0, "triplet_augment: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = dtype == RV_F32 ? 4 : 2;
  rv::LaunchScope scope(rv::CAT_LAYOUT, st, 16.0 * b * hw * es);
  if (dtype == RV_F32) rv::triplet_kernel<float><<<rv::grid_for(hw, b), 256, 0, st>>>((const float*)target, (float*)out, b, hw);
  else if (dtype == RV_BF16)
    rv::triplet_kernel<__nv_bfloat16><<<rv::grid_for(hw, b), 256, 0, st>>>((const __nv_bfloat16*)target, (__nv_bfloat16*)out, b, hw);
  else RV_CHECK_ARG(false, "triplet_augment: bad dtype %d", dtype);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_background_blend(const void* x, const float* colors, const unsigned char* mask, void* y, int n, int64_t hw, int dtype,
                        void* stream) {
  RV_CHECK_ARG(x && colors && mask && y && n > 0 && hw > 0, "background_blend: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = dtype == RV_F32 ? 4 : 2;
  rv::LaunchScope scope(rv::CAT_LAYOUT, st, 8.0 * n * hw * es);
  if (dtype == RV_F32)
    rv::background_blend_kernel<float><<<rv::grid_for(hw, n), 256, 0, st>>>((const float*)x, colors, mask, (float*)y, hw);
  else if (dtype == RV_BF16)
    rv::background_blend_kernel<__nv_bfloat16><<<rv::grid_for(hw, n), 256, 0, st>>>((const __nv_bfloat16*)x, colors, mask,
                                                                                 (__nv_bfloat16*)y, hw);
  else RV_CHECK_ARG(false, "background_blend: bad dtype %d", dtype);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_pack_latents(const void* x, void* y, int n, int c, int h, int w, int dtype, float shift, float scale, int unpack,
                    void* stream) {
  RV_CHECK_ARG(x && y && n > 0 && c > 0 && h > 0 && w > 0, "pack_latents: bad argument");
  RV_CHECK_ARG(h % 2 == 0 && w % 2 == 0, "pack_latents: latent height and width must be even (got %dx%d)", h, w);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t per = (int64_t)c * h * w;
  const size_t es = dtype == RV_F32 ? 4 : 2;
  rv::LaunchScope scope(rv::CAT_LAYOUT, st, 2.0 * n * per * es);
  if (dtype == RV_F32)
    rv::pack_latents_kernel<float><<<rv::grid_for(per, n), 256, 0, st>>>((const float*)x, (float*)y, c, h, w, shift, scale, unpack);
  else if (dtype == RV_BF16)
    rv::pack_latents_kernel<__nv_bfloat16><<<rv::grid_for(per, n), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, c, h, w,
                                                                                shift, scale, unpack);
  else RV_CHECK_ARG(false, "pack_latents: bad dtype %d", dtype);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_blend_tiles(const void* a, void* b, int planes, int ah, int aw, int bh, int bw, int extent, int vertical, int dtype,
                   void* stream) {
  RV_CHECK_ARG(a && b && planes > 0 && ah > 0 && aw > 0 && bh > 0 && bw > 0, "blend_tiles: bad argument");
  const int lim = vertical ? (ah < bh ? ah : bh) : (aw < bw ? aw : bw);
  if (extent > lim) extent = lim;
  if (extent <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int other = vertical ? (aw < bw ? aw : bw) : (ah < bh ? ah : bh);
  rv::LaunchScope scope(rv::CAT_LAYOUT, st, 3.0 * planes * (double)extent * other * (dtype == RV_F32 ? 4 : 2));
  dim3 grid = rv::grid_for((int64_t)extent * other, planes);
  if (dtype == RV_F32) rv::blend_kernel<float><<<grid, 256, 0, st>>>((const float*)a, (float*)b, ah, aw, bh, bw, extent, vertical);
  else if (dtype == RV_BF16)
    rv::blend_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a, (__nv_bfloat16*)b, ah, aw, bh, bw, extent, vertical);
  else RV_CHECK_ARG(false, "blend_tiles: bad dtype %d", dtype);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_rgba_u8_to_nchw(const void* x_u8, void* y, int n, int64_t hw, int y_dtype, float scale, float shift, void* stream) {
  RV_CHECK_ARG(x_u8 && y && n > 0 && hw > 0 && ((uintptr_t)x_u8 % 4 == 0), "rgba_u8_to_nchw: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  rv::LaunchScope scope(rv::CAT_LAYOUT, st, (double)n * hw * (4.0 + 4.0 * (y_dtype == RV_F32 ? 4 : 2)));
  if (y_dtype == RV_F32) rv::u8_to_nchw_kernel<float><<<rv::grid_for(hw, n), 256, 0, st>>>((const uint8_t*)x_u8, (float*)y, hw, scale, shift);
  else if (y_dtype == RV_BF16)
    rv::u8_to_nchw_kernel<__nv_bfloat16><<<rv::grid_for(hw, n), 256, 0, st>>>((const uint8_t*)x_u8, (__nv_bfloat16*)y, hw, scale, shift);
  else RV_CHECK_ARG(false, "rgba_u8_to_nchw: bad dtype %d", y_dtype);
  RV_LAUNCH_CHECK();
  return 0;
}

int rv_nchw_to_rgba_u8(const void* x, void* y_u8, int n, int64_t hw, int x_dtype, void* stream) {
  RV_CHECK_ARG(x && y_u8 && n > 0 && hw > 0 && ((uintptr_t)y_u8 % 4 == 0), "nchw_to_rgba_u8: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  rv::LaunchScope scope(rv::CAT_LAYOUT, st, (double)n * hw * (4.0 + 4.0 * (x_dtype == RV_F32 ? 4 : 2)));
  if (x_dtype == RV_F32) rv::nchw_to_u8_kernel<float><<<rv::grid_for(hw, n), 256, 0, st>>>((const float*)x, (uint8_t*)y_u8, hw);
  else if (x_dtype == RV_BF16)
    rv::nchw_to_u8_kernel<__nv_bfloat16><<<rv::grid_for(hw, n), 256, 0, st>>>((const __nv_bfloat16*)x, (uint8_t*)y_u8, hw);
  else RV_CHECK_ARG(false, "nchw_to_rgba_u8: bad dtype %d", x_dtype);
  RV_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
