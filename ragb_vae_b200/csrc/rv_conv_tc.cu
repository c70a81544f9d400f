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
//   warps 2-5  epilogue          tcgen05.ld -> alpha, bias, residual, scale/shift, clamp -> global
#include <cuda.h>

#include <mutex>

#include "rv_common.cuh"

namespace rv {

int check_conv_desc(const rv_conv_desc* d);  // rv_conv_direct.cu

constexpr int TC_THREADS = 192;
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
  signed char tap_dx[TC_MAX_TAPS], tap_dy[TC_MAX_TAPS], tap_wp[TC_MAX_TAPS], tap_hp[TC_MAX_TAPS];
  // output mapping: out pixel = tile pixel * os + oo
  int out_h, out_w, osy, osx, ooy, oox;
  int cout, y_cstride, y_nchw, y_f32;
  int bias_mode, clamp;
  int vec_ok;                // NHWC y / residual rows are 16-byte aligned: vector epilogue
  float alpha, out_scale, out_shift, clamp_lo, clamp_hi;
  const float* bias;
  const __nv_bfloat16* residual;
  void* y;
  int stages;
  uint32_t a_bytes, stage_bytes, tx_bytes;
  uint32_t sbo_bytes, layout_type;
};

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Spin on the phase parity.  A wait that lasts ~2 s of SM clocks is a protocol bug: trap so the
// launch fails with an error instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((it & 1023u) == 1023u) {
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) {
        printf("rgbavae: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
               (int)threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by one thread for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Shared-memory matrix descriptor, K-major operand written by TMA with a 32/64/128-byte
// swizzle (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout [61,64)).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

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

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ uint32_t tmem_base_slot;

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
      mbar_init(smem_u32(&bar_acc_empty[a]), 4);
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int dx = p.tap_dx[tap], dy = p.tap_dy[tap];
          const int kcol = p.w_k_base + tap * p.cin;
          for (int kc = 0; kc < p.kc_per_tap; ++kc) {
            const uint32_t full = smem_u32(&bar_full[stage]);
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            mbar_arrive_expect_tx(full, p.tx_bytes);
            const uint32_t a_dst = smem_base + stage * p.stage_bytes;
            const uint32_t b_dst = a_dst + p.a_bytes;
            if (p.mode == 0)
              tma_load_4d(a_dst, &map_a, full, kc * p.bk, t.x0 + dx, t.y0 + dy, t.img);
            else
              tma_load_5d(a_dst, &map_a, full, p.tap_wp[tap] * p.cin_pitch + kc * p.bk, t.x0 + dx, p.tap_hp[tap],
                          t.y0 + dy, t.img);
            tma_load_2d(b_dst, &map_b, full, kcol + kc * p.bk, t.n0);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      // cute::UMMA::InstrDescriptor: D fp32 (bit 4), A bf16 (bit 7), B bf16 (bit 10), both K-major,
      // N>>3 at bit 17, M>>4 at bit 24.
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((128u >> 4) << 24);
      const int ksteps = p.bk >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&bar_acc_empty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * p.stage_bytes;
          const uint64_t a_desc = make_smem_desc(a_addr, p.sbo_bytes, p.layout_type);
          const uint64_t b_desc = make_smem_desc(a_addr + p.a_bytes, p.sbo_bytes, p.layout_type);
          for (int k = 0; k < ksteps; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr>>4) field
            umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(smem_u32(&bar_empty[stage]));
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(smem_u32(&bar_acc_full[acc]));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int ly = row / p.bw, lx = row - ly * p.bw;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int ty = t.y0 + ly, tx = t.x0 + lx;
      const bool valid = ty < p.th && tx < p.tw;
      const int oy = ty * p.osy + p.ooy, ox = tx * p.osx + p.oox;
      const int64_t pix = ((int64_t)t.img * p.out_h + oy) * p.out_w + ox;
      mbar_wait(smem_u32(&bar_acc_full[acc]), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
      for (int c0 = 0; c0 < p.bn; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        const int co0 = t.n0 + c0;
        if (valid && co0 < p.cout) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
          const bool full16 = p.vec_ok && co0 + 16 <= p.cout;
          if (p.bias_mode == 1) {
            if (co0 + 16 <= p.cout) {
              const float4* bp = reinterpret_cast<const float4*>(p.bias + co0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float4 b = __ldg(bp + j);
                v[4 * j] += b.x;
                v[4 * j + 1] += b.y;
                v[4 * j + 2] += b.z;
                v[4 * j + 3] += b.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (co0 + j < p.cout) v[j] += __ldg(p.bias + co0 + j);
            }
          } else if (p.bias_mode == 2) {
            const float b = __ldg(p.bias + pix);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += b;
          }
          if (p.residual) {
            const __nv_bfloat16* rp = p.residual + pix * p.y_cstride + co0;
            if (full16) {
              uint4 a = *reinterpret_cast<const uint4*>(rp);
              uint4 b = *reinterpret_cast<const uint4*>(rp + 8);
              const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[2 * j] += __uint_as_float(w[j] << 16);
                v[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (co0 + j < p.cout) v[j] += __bfloat162float(rp[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            v[j] = fmaf(v[j], p.out_scale, p.out_shift);
            if (p.clamp) v[j] = fminf(fmaxf(v[j], p.clamp_lo), p.clamp_hi);
          }
          if (p.y_nchw) {
            const int64_t plane = (int64_t)p.out_h * p.out_w;
            const int64_t base = ((int64_t)t.img * p.cout + co0) * plane + (int64_t)oy * p.out_w + ox;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (co0 + j < p.cout) {
                if (p.y_f32) reinterpret_cast<float*>(p.y)[base + j * plane] = v[j];
                else reinterpret_cast<__nv_bfloat16*>(p.y)[base + j * plane] = __float2bfloat16_rn(v[j]);
              }
            }
          } else if (p.y_f32) {
            float* yp = reinterpret_cast<float*>(p.y) + pix * p.y_cstride + co0;
            if (full16) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(yp + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (co0 + j < p.cout) yp[j] = v[j];
            }
          } else {
            __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + pix * p.y_cstride + co0;
            if (full16) {
              uint4 a, b;
              a.x = pack_bf16x2(v[0], v[1]);
              a.y = pack_bf16x2(v[2], v[3]);
              a.z = pack_bf16x2(v[4], v[5]);
              a.w = pack_bf16x2(v[6], v[7]);
              b.x = pack_bf16x2(v[8], v[9]);
              b.y = pack_bf16x2(v[10], v[11]);
              b.z = pack_bf16x2(v[12], v[13]);
              b.w = pack_bf16x2(v[14], v[15]);
              *reinterpret_cast<uint4*>(yp) = a;
              *reinterpret_cast<uint4*>(yp + 8) = b;
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (co0 + j < p.cout) yp[j] = __float2bfloat16_rn(v[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[acc]));
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
// Host side: tensor maps, tiling choice, launch
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::mutex g_init_mu;
static bool g_attr_set[64] = {false};

static int ensure_init() {
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
    RV_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TC_SMEM_BUDGET + 2048)));
    g_attr_set[dev] = true;
  }
  return 0;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
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

static int launch_tc(const rv_conv_desc* d, const void* x, const void* w, int64_t w_ld, const float* bias,
                     const void* residual, void* y, cudaStream_t st, int phase /* -1: not upsample */) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = d->n;
  p.cin = d->cin;
  p.cin_pitch = d->x_cstride;
  // 128-byte operand rows move ~2x faster through TMA than 64-byte ones (measured: the row rate, not the
  // byte rate, bounds small-N layers), so channel counts above 64 always use BK=64: the last block of a tap
  // is partly out of bounds in x (zero-filled), which cancels whatever weight columns it is paired with.
  // (not for the stride-2 parity view, whose folded channel axis has real data past cin)
  p.bk = (d->cin % 64 == 0 || (d->cin > 64 && d->stride == 1)) ? 64 : (d->cin % 32 == 0 ? 32 : 16);
  p.kc_per_tap = (d->cin + p.bk - 1) / p.bk;
  const CUtensorMapSwizzle sw =
      p.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  p.layout_type = p.bk == 64 ? 2u : (p.bk == 32 ? 4u : 6u);
  p.sbo_bytes = 8u * (uint32_t)p.bk * 2u;
  p.out_h = d->oh;
  p.out_w = d->ow;
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
    p.ntaps = d->ksize * d->ksize;
    for (int t = 0; t < p.ntaps; ++t) {
      int dy = t / d->ksize, dx = t % d->ksize;
      if (d->stride == 2) {
        p.tap_dy[t] = (signed char)(dy >> 1);
        p.tap_hp[t] = (signed char)(dy & 1);
        p.tap_dx[t] = (signed char)(dx >> 1);
        p.tap_wp[t] = (signed char)(dx & 1);
      } else {
        p.tap_dy[t] = (signed char)(dy - d->pad_lo);
        p.tap_dx[t] = (signed char)(dx - d->pad_lo);
      }
    }
    k_extent = p.ntaps * d->cin;
  }
  p.mode = d->stride == 2 ? 1 : 0;
  choose_tile(p.th, p.tw, &p.bw, &p.bh);
  p.tiles_x = (p.tw + p.bw - 1) / p.bw;
  p.tiles_y = (p.th + p.bh - 1) / p.bh;
  p.m_tiles = d->n * p.tiles_x * p.tiles_y;
  choose_bn(d->cout, &p.bn, &p.n_tiles);
  p.cout = d->cout;
  p.y_cstride = d->y_cstride;
  p.y_nchw = d->y_nchw;
  p.y_f32 = d->y_dtype == RV_F32;
  p.bias_mode = d->bias_mode;
  p.clamp = d->clamp;
  p.alpha = d->alpha;
  p.out_scale = d->out_scale;
  p.out_shift = d->out_shift;
  p.clamp_lo = d->clamp_lo;
  p.clamp_hi = d->clamp_hi;
  p.bias = bias;
  p.residual = (const __nv_bfloat16*)residual;
  p.vec_ok = (d->y_cstride % 8 == 0) && ((uintptr_t)y % 16 == 0) && (!residual || (uintptr_t)residual % 16 == 0);
  p.y = y;
  p.a_bytes = 128u * (uint32_t)p.bk * 2u;
  const uint32_t b_bytes = (uint32_t)p.bn * (uint32_t)p.bk * 2u;
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
    if (int rc = encode_map(&map_a, x, 4, dims, str, box, sw)) return rc;
  } else {
    RV_CHECK_ARG(d->h % 2 == 0 && d->w % 2 == 0, "conv_tc: stride-2 conv needs even input size");
    cuuint64_t dims[5] = {(cuuint64_t)(2 * d->x_cstride), (cuuint64_t)(d->w / 2), 2, (cuuint64_t)(d->h / 2), (cuuint64_t)d->n};
    cuuint64_t str[4] = {2 * pitch_b, pitch_b * d->w, 2 * pitch_b * d->w, pitch_b * d->w * d->h};
    cuuint32_t box[5] = {(cuuint32_t)p.bk, (cuuint32_t)p.bw, 1, (cuuint32_t)p.bh, 1};
    if (int rc = encode_map(&map_a, x, 5, dims, str, box, sw)) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)k_extent, (cuuint64_t)d->cout};
    cuuint64_t str[1] = {(cuuint64_t)w_ld * 2u};
    cuuint32_t box[2] = {(cuuint32_t)p.bk, (cuuint32_t)p.bn};
    if (int rc = encode_map(&map_b, w, 2, dims, str, box, sw)) return rc;
  }
  const int total_tiles = p.m_tiles * p.n_tiles;
  int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  const double flops = 2.0 * (double)d->n * d->oh * d->ow * d->cout * d->cin * d->ksize * d->ksize / (phase >= 0 ? 4.0 : 1.0);
  LaunchScope scope(CAT_CONV_TC, st, flops);
  conv_tc_kernel<<<grid, TC_THREADS, smem, st>>>(map_a, map_b, p);
  RV_LAUNCH_CHECK();
  return 0;
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

int rv_init(void) { return rv::ensure_init(); }

int rv_conv2d_tc(const rv_conv_desc* d, const void* x, const void* w_packed, int64_t w_ld, const float* bias,
                 const void* residual, void* y, void* stream) {
  if (int rc = rv::check_conv_desc(d)) return rc;
  if (int rc = rv::ensure_init()) return rc;
  RV_CHECK_ARG(x && w_packed && y, "conv_tc: null tensor");
  RV_CHECK_ARG(d->x_dtype == RV_BF16 && !d->x_nchw, "conv_tc: x must be NHWC bf16");
  RV_CHECK_ARG(d->cin % 16 == 0, "conv_tc: cin (%d) must be a multiple of 16", d->cin);
  RV_CHECK_ARG(d->x_cstride % 8 == 0 && w_ld % 8 == 0, "conv_tc: x_cstride and w_ld must be multiples of 8");
  RV_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)w_packed % 16 == 0), "conv_tc: x / w must be 16-byte aligned");
  RV_CHECK_ARG(d->in_scale == 1.0f && d->in_shift == 0.0f, "conv_tc: in_scale/in_shift are direct-path only");
  RV_CHECK_ARG(d->bias_mode == 0 || bias, "conv_tc: bias_mode set but bias is null");
  RV_CHECK_ARG(!residual || (!d->y_nchw && d->y_dtype == RV_BF16), "conv_tc: residual needs an NHWC bf16 output");
  RV_CHECK_ARG(d->bias_mode != 1 || (uintptr_t)bias % 16 == 0, "conv_tc: bias must be 16-byte aligned");
  RV_CHECK_ARG(!(d->upsample && d->ksize != 3), "conv_tc: upsample requires ksize 3");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->upsample) {
    for (int phase = 0; phase < 4; ++phase)
      if (int rc = rv::launch_tc(d, x, w_packed, w_ld, bias, residual, y, st, phase)) return rc;
    return 0;
  }
  return rv::launch_tc(d, x, w_packed, w_ld, bias, residual, y, st, -1);
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
