// Device-side building blocks shared by the tcgen05 kernels of librgbavae (rv_conv_tc.cu, rv_conv_halo.cu):
// mbarrier / TMA / tcgen05 PTX wrappers, UMMA descriptors, and the accumulator epilogue
// (alpha, bias, residual, affine, clamp, optional fused RMS-norm + SiLU second output).
#pragma once
#include <cuda.h>

#include "rv_common.cuh"

namespace rv {

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

// One lane of the (converged) warp; the others get 0.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred;
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


// ---------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) primitives: a cluster of two CTAs issues one M = 256 MMA whose B operand is split
// between the two CTAs' shared memories, so each SM reads only half of B per MMA.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA's layout) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a barrier that may live in the peer CTA.  Default semantics (.release at CTA scope), NOT .release.cluster: the
// barriers signalled this way hand over TMEM accumulators / operand slots, whose ordering comes from tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync; a cluster-scope release made every epilogue warp drain its outstanding global loads and
// stores first (MEMBAR.ALL + ERRBAR: 22 % of all stall samples of the CTA-pair halo kernel, profiles/r02_*).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion is signalled on a barrier that may live in the peer CTA (the leader's full barrier)
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2,
                                             int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
// D[tmem of both CTAs] (+)= A[each CTA's smem, 128 rows each] * B[N/2 rows in each CTA's smem]; issued by the leader CTA
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once the pair's MMAs so far have completed
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(bar)
      : "memory");
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int W>
__device__ __forceinline__ void tmem_ldw(uint32_t taddr, uint32_t (&r)[W]);
template <> __device__ __forceinline__ void tmem_ldw<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }
template <> __device__ __forceinline__ void tmem_ldw<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}


// ---------------------------------------------------------------------------------------
// Epilogue: one thread owns one output pixel (TMEM lane) and walks its channels 16 at a time.
// ---------------------------------------------------------------------------------------
struct EpiParams {
  int out_h, out_w;
  int cout, y_cstride, y_nchw, y_f32;
  int bias_mode, clamp;
  int vec_ok;                // NHWC y / residual / y_act rows are 16-byte aligned: vector accesses
  float alpha, out_scale, out_shift, clamp_lo, clamp_hi;
  const float* bias;
  const __nv_bfloat16* residual;
  void* y;                   // may be null when only the normalised output is wanted
  // fused QwenImageRMS_norm (+SiLU) of the output pixel, written as a second NHWC bf16 tensor:
  // y_act = act(v / max(||v||_2, 1e-12) * gamma_scaled), gamma_scaled = gamma * sqrt(cout)
  const float* norm_gamma;   // null: no fused norm
  __nv_bfloat16* y_act;
  int norm_silu;
  int fast;                  // 1: epilogue_pixel_fast preconditions hold, 2: epilogue_pixel_rowstat's (set by the host)
  // per-row statistic of a GEMM whose rows are attention queries (generic path only; rv_gemm_rowstat):
  //   1: y = exp2(alpha * acc - rowstat[row])              probabilities from the forward's log-sum-exp
  //   2: y = residual * (alpha * acc - rowstat[row])       dS = P * (dP - delta) * scale (residual holds P)
  const float* rowstat;
  int rowstat_mode;
  // GroupNorm statistics of the output (rv_conv2d_tc_gnstats): per-tile partial sums, [tile][8 epilogue warps][32] floats
  float* gn_part;
  int gn_gs;                 // channels per group
};

// optional per-row statistic (host-side description; see EpiParams::rowstat_mode)
struct RowStat {
  const float* stat;
  int mode;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// optional fused RMS-norm second output (host-side description)
struct NormFuse {
  const float* gamma_scaled;
  void* y_act;
  int silu;
};

// accumulator chunk -> alpha, bias, residual, affine, clamp
__device__ __forceinline__ void epi_values(const EpiParams& e, const uint32_t (&r)[16], int co0, int64_t pix, float (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * e.alpha;
  const bool full16 = e.vec_ok && co0 + 16 <= e.cout;
  if (e.bias_mode == 1) {
    if (co0 + 16 <= e.cout) {
      const float4* bp = reinterpret_cast<const float4*>(e.bias + co0);
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
        if (co0 + j < e.cout) v[j] += __ldg(e.bias + co0 + j);
    }
  } else if (e.bias_mode == 2) {
    const float b = __ldg(e.bias + pix);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += b;
  }
  if (e.rowstat_mode) {
    const float s = __ldg(e.rowstat + pix);
    if (e.rowstat_mode == 1) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = ex2_approx(v[j] - s);
    } else {
      const __nv_bfloat16* rp = e.residual + pix * e.y_cstride + co0;
      if (full16) {
        uint4 a = *reinterpret_cast<const uint4*>(rp);
        uint4 b = *reinterpret_cast<const uint4*>(rp + 8);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[2 * j] = __uint_as_float(w[j] << 16) * (v[2 * j] - s);
          v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u) * (v[2 * j + 1] - s);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (co0 + j < e.cout) v[j] = __bfloat162float(rp[j]) * (v[j] - s);
      }
    }
  } else if (e.residual) {
    const __nv_bfloat16* rp = e.residual + pix * e.y_cstride + co0;
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
        if (co0 + j < e.cout) v[j] += __bfloat162float(rp[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    v[j] = fmaf(v[j], e.out_scale, e.out_shift);
    if (e.clamp) v[j] = fminf(fmaxf(v[j], e.clamp_lo), e.clamp_hi);
  }
}

__device__ __forceinline__ void store_bf16_row16(__nv_bfloat16* yp, const float (&v)[16], bool full16, int valid_ch) {
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
    if ((reinterpret_cast<uintptr_t>(yp) & 31u) == 0) {
      store32(yp, a, b);
    } else {
      *reinterpret_cast<uint4*>(yp) = a;
      *reinterpret_cast<uint4*>(yp + 8) = b;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < valid_ch) yp[j] = __float2bfloat16_rn(v[j]);
  }
}

__device__ __forceinline__ void epi_store(const EpiParams& e, const float (&v)[16], int img, int oy, int ox, int64_t pix,
                                          int co0) {
  const bool full16 = e.vec_ok && co0 + 16 <= e.cout;
  if (e.y_nchw) {
    const int64_t plane = (int64_t)e.out_h * e.out_w;
    const int64_t base = ((int64_t)img * e.cout + co0) * plane + (int64_t)oy * e.out_w + ox;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (co0 + j < e.cout) {
        if (e.y_f32) reinterpret_cast<float*>(e.y)[base + j * plane] = v[j];
        else reinterpret_cast<__nv_bfloat16*>(e.y)[base + j * plane] = __float2bfloat16_rn(v[j]);
      }
    }
  } else if (e.y_f32) {
    float* yp = reinterpret_cast<float*>(e.y) + pix * e.y_cstride + co0;
    if (full16) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(yp + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (co0 + j < e.cout) yp[j] = v[j];
    }
  } else {
    store_bf16_row16(reinterpret_cast<__nv_bfloat16*>(e.y) + pix * e.y_cstride + co0, v, full16, e.cout - co0);
  }
}

// Whole-tile epilogue for the thread's pixel: `taddr` = TMEM address of (lane quarter, first accumulator column),
// `bn` accumulator columns holding channels [n0, n0+bn).  All 32 lanes must call (tcgen05.ld is warp-collective).
__device__ __forceinline__ void epilogue_pixel(const EpiParams& e, uint32_t taddr, int cb, int ce, int n0, bool valid,
                                               int img, int oy, int ox, int64_t pix) {
  {
    for (int c0 = cb; c0 < ce; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(taddr + (uint32_t)c0, r);
      tmem_ld_wait();
      const int co0 = n0 + c0;
      if (valid && co0 < e.cout) {
        float v[16];
        epi_values(e, r, co0, pix, v);
        epi_store(e, v, img, oy, ox, pix, co0);
      }
    }
  }
  // (the fused RMS norm exists only in the lean path: epilogue_pixel_fast)
}

// ---------------------------------------------------------------------------------------
// Lean epilogue for the hot case: NHWC bf16 output, 16-byte aligned rows, cout % 16 == 0, per-channel bias
// (or none), alpha = 1, no affine / clamp.  Every per-element runtime check of the generic path is gone, bias and
// gamma come from shared memory (a global __ldg per chunk cost a full L2 round trip with the L1 carve-out at its
// minimum), and the residual loads of a step are issued before its TMEM load so their latency overlaps.
// W = 16 or 32 accumulator columns per step.  `sbias` / `sgamma` are indexed by absolute channel.
// ---------------------------------------------------------------------------------------
template <int W, bool RES>
struct FastStep {
  uint4 res[RES ? W / 8 : 1];
  uint32_t r[W];

  __device__ __forceinline__ void load(const EpiParams& e, uint32_t taddr, int co0, int64_t pix, bool valid) {
    if (RES && valid) {
      const __nv_bfloat16* rp16 = e.residual + pix * e.y_cstride + co0;
      if ((reinterpret_cast<uintptr_t>(rp16) & 31u) == 0) {  // 32-byte loads: one request per sector
#pragma unroll
        for (int q = 0; q < W / 16; ++q)
          asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(res[2 * q].x), "=r"(res[2 * q].y), "=r"(res[2 * q].z), "=r"(res[2 * q].w), "=r"(res[2 * q + 1].x),
                         "=r"(res[2 * q + 1].y), "=r"(res[2 * q + 1].z), "=r"(res[2 * q + 1].w)
                       : "l"(rp16 + 16 * q));
      } else {
        const uint4* rp = reinterpret_cast<const uint4*>(rp16);
#pragma unroll
        for (int q = 0; q < W / 8; ++q) res[q] = rp[q];
      }
    }
    tmem_ldw<W>(taddr, r);
    tmem_ld_wait();
  }

  __device__ __forceinline__ void values(const float* sbias, int co0, float (&v)[W]) const {
    if (sbias) {
#pragma unroll
      for (int j = 0; j < W / 4; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(sbias + co0 + 4 * j);
        v[4 * j] = __uint_as_float(r[4 * j]) + b.x;
        v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
        v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
        v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = __uint_as_float(r[j]);
    }
    if (RES) {
#pragma unroll
      for (int q = 0; q < W / 8; ++q) {
        const uint32_t w[4] = {res[q].x, res[q].y, res[q].z, res[q].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[8 * q + 2 * j] += __uint_as_float(w[j] << 16);
          v[8 * q + 2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
        }
      }
    }
  }
};

template <int W>
__device__ __forceinline__ void fast_store(__nv_bfloat16* yp, const float (&v)[W]) {
  // 32-byte (whole-sector) stores when the row is 32-byte aligned, else 16-byte ones
  if ((reinterpret_cast<uintptr_t>(yp) & 31u) == 0) {
#pragma unroll
    for (int q = 0; q < W / 16; ++q) {
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(yp + 16 * q),
                   "r"(pack_bf16x2(v[16 * q], v[16 * q + 1])), "r"(pack_bf16x2(v[16 * q + 2], v[16 * q + 3])),
                   "r"(pack_bf16x2(v[16 * q + 4], v[16 * q + 5])), "r"(pack_bf16x2(v[16 * q + 6], v[16 * q + 7])),
                   "r"(pack_bf16x2(v[16 * q + 8], v[16 * q + 9])), "r"(pack_bf16x2(v[16 * q + 10], v[16 * q + 11])),
                   "r"(pack_bf16x2(v[16 * q + 12], v[16 * q + 13])), "r"(pack_bf16x2(v[16 * q + 14], v[16 * q + 15]))
                   : "memory");
    }
    return;
  }
  uint4* dst = reinterpret_cast<uint4*>(yp);
#pragma unroll
  for (int q = 0; q < W / 8; ++q) {
    uint4 a;
    a.x = pack_bf16x2(v[8 * q], v[8 * q + 1]);
    a.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
    a.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
    a.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
    dst[q] = a;
  }
}

template <int W, bool RES>
__device__ __forceinline__ void fast_plain_step(const EpiParams& e, const float* sbias, uint32_t taddr, int c0, int co0,
                                                bool valid, int64_t pix) {
  FastStep<W, RES> st;
  st.load(e, taddr + (uint32_t)c0, co0, pix, valid);
  if (valid) {
    float v[W];
    st.values(sbias, co0, v);
    fast_store<W>(reinterpret_cast<__nv_bfloat16*>(e.y) + pix * e.y_cstride + co0, v);
  }
}

template <int W, bool RES>
__device__ __forceinline__ float fast_norm_pass1(const EpiParams& e, const float* sbias, uint32_t taddr, int c0, int co0,
                                                 bool valid, int64_t pix) {
  FastStep<W, RES> st;
  st.load(e, taddr + (uint32_t)c0, co0, pix, valid);
  float ss = 0.f;
  if (valid) {
    float v[W];
    st.values(sbias, co0, v);
    if (e.y) fast_store<W>(reinterpret_cast<__nv_bfloat16*>(e.y) + pix * e.y_cstride + co0, v);
#pragma unroll
    for (int j = 0; j < W; ++j) ss = fmaf(v[j], v[j], ss);
  }
  return ss;
}

template <int W, bool RES>
__device__ __forceinline__ void fast_norm_pass2(const EpiParams& e, const float* sbias, const float* sgamma, uint32_t taddr,
                                                int c0, int co0, bool valid, int64_t pix, float rinv) {
  FastStep<W, RES> st;
  st.load(e, taddr + (uint32_t)c0, co0, pix, valid);
  if (valid) {
    float v[W];
    st.values(sbias, co0, v);
#pragma unroll
    for (int j = 0; j < W / 4; ++j) {
      const float4 g = *reinterpret_cast<const float4*>(sgamma + co0 + 4 * j);
      v[4 * j] *= rinv * g.x;
      v[4 * j + 1] *= rinv * g.y;
      v[4 * j + 2] *= rinv * g.z;
      v[4 * j + 3] *= rinv * g.w;
    }
    if (e.norm_silu) {  // sgamma carries the 1/2: v is h = f / 2 here and silu(f) = h tanh(h) + h
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = silu_from_half(v[j]);
    }
    fast_store<W>(e.y_act + pix * e.y_cstride + co0, v);
  }
}

// Thread's pixel, accumulator columns [cb, ce) of the tile (both multiples of 16) = channels n0+cb .. n0+ce.
// Fused norm: the pixel's sum of squares is completed across the `nsplit` warps that share the lane quarter through
// `ss_slot` (shared memory, [nsplit][128]) and the named barrier `bar_id`.
template <bool RES>
__device__ __forceinline__ void epilogue_pixel_fast(const EpiParams& e, const float* sbias, const float* sgamma,
                                                    uint32_t taddr, int cb, int ce, int n0, bool valid, int64_t pix,
                                                    float* ss_slot, int row, int half, int nsplit, int bar_id) {
  const int c32 = cb + ((ce - cb) & ~31);
  if (e.norm_gamma == nullptr) {
    for (int c0 = cb; c0 < c32; c0 += 32) fast_plain_step<32, RES>(e, sbias, taddr, c0, n0 + c0, valid, pix);
    if (c32 < ce) fast_plain_step<16, RES>(e, sbias, taddr, c32, n0 + c32, valid, pix);
    return;
  }
  float ss = 0.f;
  for (int c0 = cb; c0 < c32; c0 += 32) ss += fast_norm_pass1<32, RES>(e, sbias, taddr, c0, n0 + c0, valid, pix);
  if (c32 < ce) ss += fast_norm_pass1<16, RES>(e, sbias, taddr, c32, n0 + c32, valid, pix);
  if (nsplit > 1) {
    ss_slot[half * 128 + row] = ss;
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 * nsplit) : "memory");
    ss = 0.f;
    for (int hh = 0; hh < nsplit; ++hh) ss += ss_slot[hh * 128 + row];
  }
  const float rinv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  for (int c0 = cb; c0 < c32; c0 += 32) fast_norm_pass2<32, RES>(e, sbias, sgamma, taddr, c0, n0 + c0, valid, pix, rinv);
  if (c32 < ce) fast_norm_pass2<16, RES>(e, sbias, sgamma, taddr, c32, n0 + c32, valid, pix, rinv);
}

// Lean epilogue of rv_gemm_rowstat (EpiParams::rowstat_mode; bf16 NHWC output, 32-byte aligned rows, column ranges in
// multiples of 32, at most 128 columns per warp).  What does not depend on the accumulator -- the row's statistic and, in
// mode 2, the thread's whole slice of the multiplicand (up to 4 x 64 bytes) -- is fetched BEFORE the warp waits for the
// tile's MMAs, so its HBM latency hides behind them.
struct RowstatPrefetch {
  uint4 res[16];
  float s;
};

template <int MODE>
__device__ __forceinline__ void rowstat_prefetch(const EpiParams& e, int cb, int ce, int n0, bool valid, int64_t pix,
                                                 RowstatPrefetch& pf) {
  pf.s = valid ? __ldg(e.rowstat + pix) : 0.f;
  if (MODE == 2 && valid) {
    const __nv_bfloat16* rp = e.residual + pix * e.y_cstride + n0 + cb;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (cb + 32 * c < ce) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
          asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(pf.res[4 * c + 2 * q].x), "=r"(pf.res[4 * c + 2 * q].y), "=r"(pf.res[4 * c + 2 * q].z),
                         "=r"(pf.res[4 * c + 2 * q].w), "=r"(pf.res[4 * c + 2 * q + 1].x), "=r"(pf.res[4 * c + 2 * q + 1].y),
                         "=r"(pf.res[4 * c + 2 * q + 1].z), "=r"(pf.res[4 * c + 2 * q + 1].w)
                       : "l"(rp + 32 * c + 16 * q));
      }
    }
  }
}

template <int MODE>
__device__ __forceinline__ void epilogue_pixel_rowstat(const EpiParams& e, uint32_t taddr, int cb, int ce, int n0, bool valid,
                                                       int64_t pix, const RowstatPrefetch& pf) {
  const float s = pf.s, alpha = e.alpha;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int c0 = cb + 32 * c;
    if (c0 < ce) {
      uint32_t r[32];
      tmem_ld32(taddr + (uint32_t)c0, r);
      tmem_ld_wait();
      if (valid) {
        float v[32];
        if (MODE == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ex2_approx(fmaf(__uint_as_float(r[j]), alpha, -s));
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 u = pf.res[4 * c + q];
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[8 * q + 2 * j] = __uint_as_float(w[j] << 16) * fmaf(__uint_as_float(r[8 * q + 2 * j]), alpha, -s);
              v[8 * q + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u) * fmaf(__uint_as_float(r[8 * q + 2 * j + 1]), alpha, -s);
            }
          }
        }
        fast_store<32>(reinterpret_cast<__nv_bfloat16*>(e.y) + pix * e.y_cstride + n0 + c0, v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// GroupNorm statistics of a conv's OUTPUT from its epilogue (rv_conv2d_tc_gnstats): the lean plain epilogue, plus per
// (pixel, channel group) sum and sum of squares of the values it stores, folded over the warp's 32 pixels by a
// transpose-reduction (V values on every lane -> lane l holds the total of value l / (32 / V): 31 shuffles for V = 32
// instead of 160) and written as the warp's 32 floats of this tile.  A finishing kernel adds the tiles of a sample in a fixed
// order (fp64): deterministic, and independent of the batch the sample is in.
// GS = channels per group (C / 32); NST = 32-column steps of the warp's column range; V = 2 * NST * 32 / GS <= 32.
// vals[2 * g + stat] with g counting groups from the warp's first column.
// ---------------------------------------------------------------------------------------
template <int V>
__device__ __forceinline__ float warp_transpose_reduce(float (&vals)[V], int lane) {
  static_assert(V == 8 || V == 16 || V == 32, "V must be 8, 16 or 32");
  int n = V;
#pragma unroll
  for (int mask = 16; mask >= 1; mask >>= 1) {
    if (n > 1) {
      n >>= 1;
      const bool up = (lane & mask) != 0;
#pragma unroll
      for (int i = 0; i < V / 2; ++i) {
        if (i < n) {
          const float send = up ? vals[i] : vals[i + n];
          const float keep = up ? vals[i + n] : vals[i];
          vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
      }
    } else {
      vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], mask);
    }
  }
  return vals[0];
}

template <int GS, int NST, bool RES>
__device__ __forceinline__ void epilogue_pixel_gnstats(const EpiParams& e, const float* sbias, uint32_t taddr, int cb, int n0,
                                                       bool valid, int64_t pix, float* part, int lane) {
  constexpr int GP = 32 / GS;       // groups per 32-column step
  constexpr int V = 2 * NST * GP;
  float vals[V];
#pragma unroll
  for (int i = 0; i < V; ++i) vals[i] = 0.f;
#pragma unroll
  for (int st = 0; st < NST; ++st) {
    const int c0 = cb + 32 * st;
    FastStep<32, RES> fs;
    fs.load(e, taddr + (uint32_t)c0, n0 + c0, pix, valid);
    if (valid) {
      float v[32];
      fs.values(sbias, n0 + c0, v);
      fast_store<32>(reinterpret_cast<__nv_bfloat16*>(e.y) + pix * e.y_cstride + n0 + c0, v);
#pragma unroll
      for (int g = 0; g < GP; ++g) {
        float sa = 0.f, sq = 0.f;
#pragma unroll
        for (int k = 0; k < GS; ++k) {
          sa += v[g * GS + k];
          sq = fmaf(v[g * GS + k], v[g * GS + k], sq);
        }
        vals[2 * (st * GP + g)] = sa;
        vals[2 * (st * GP + g) + 1] = sq;
      }
    }
  }
  const float tot = warp_transpose_reduce<V>(vals, lane);
  part[lane] = tot;
}

}  // namespace rv
