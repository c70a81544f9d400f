// Fused reductions over RGBA image pairs (NCHW planes): the AlphaVAE reconstruction loss, the
// composite-over-background + PSNR + alpha-MAE validation metrics and the weighted terms of RgbaVAE.loss.
// Each reads both inputs exactly once (8*B*H*W elements of traffic), accumulates in fp32 per thread and in
// fp64 across threads and blocks, and finishes IN THE SAME LAUNCH: every block writes its fp64 partial, takes a
// ticket, and the block that draws the last ticket of its sample sums the partials in a fixed order
// (deterministic, independent of which block happens to be last) and writes the result.  The partition
// depends on H*W only, so a sample's bits do not depend on the batch it sits in.
#include "rv_common.cuh"

#include <map>
#include <mutex>

namespace rv {

constexpr int RB_PIX_PER_BLOCK = 8192;    // 4 sweeps of 256 threads x 8 bf16 (measured best of 2048 ... 65536 at B = 8 and B = 32)
constexpr int RB_MAX = 512;
constexpr int MAX_BG = 4;
constexpr int TICKET_SLOTS = 64;          // distinct streams that can use the single-launch path
constexpr int TICKET_BATCH = 1024;        // samples per launch on that path

// Ticket counters: zero at module load, and every launch leaves its counters at zero again (the finishing block
// resets its sample's counter).  One row per stream, so launches on different streams never share a counter and
// launches on one stream are ordered.
__device__ unsigned int g_tickets[TICKET_SLOTS][TICKET_BATCH];

static std::mutex g_slot_mu;
static std::map<std::pair<int, cudaStream_t>, int> g_slots;
static int g_next_slot[64];

// Row of g_tickets for (current device, stream); -1 when all rows are taken (the caller then uses two launches).
static int ticket_slot(cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_slot_mu);
  auto key = std::make_pair(dev, st);
  auto it = g_slots.find(key);
  if (it != g_slots.end()) return it->second;
  if (dev < 0 || dev >= 64 || g_next_slot[dev] >= TICKET_SLOTS) return -1;
  const int s = g_next_slot[dev]++;
  g_slots[key] = s;
  return s;
}

static inline int reduce_blocks(int64_t hw) {
  int64_t b = (hw + RB_PIX_PER_BLOCK - 1) / RB_PIX_PER_BLOCK;
  if (b < 1) b = 1;
  if (b > RB_MAX) b = RB_MAX;
  return (int)b;
}

// Block-level fp64 sum of NV per-thread values; result valid in thread 0.  Safe to call repeatedly.
template <int NV>
__device__ __forceinline__ void block_sum_d(const double (&v)[NV], double (&out)[NV]) {
  __shared__ double red[NV][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double d = warp_sum_d(v[k]);
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

// Raw 16-byte vectors are kept in registers (4 planes per tensor) and converted element by element.
template <typename T, int VEC>
struct PlaneVec {
  Vec16<T> v[4];
  float s[4];
  __device__ __forceinline__ void load(const T* base, int64_t hw, int64_t i) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (VEC == 1) s[c] = ldf(base + c * hw + i);
      else v[c].load(base + c * hw + i * VEC);
    }
  }
  __device__ __forceinline__ float get(int c, int j) const { return VEC == 1 ? s[c] : v[c].get(j); }
};

// ---- the three reductions as pixel functors --------------------------------------------------------------------
// Op::NV accumulators; pixel(acc, p, t) adds one pixel (p / t = the 4 channel values of the first / second tensor);
// finish(sum, out, hw) turns the sample's fp64 sums into its output row.

// AlphaVaeLoss.reconstruction_loss (reference src/models/losses.py:67-83): d = t_rgb*at - p_rgb*ap, da = at - ap,
// l = d^2 - 2*Eb*d*da + Eb2*da^2; naive: (p - t)^2 over the 4 channels.  Inputs in [-1, 1].
struct LossOp {
  static constexpr int NV = 1;
  float eb[3], eb2[3];
  int naive;
  __device__ __forceinline__ void pixel(float (&acc)[NV], const float (&p)[4], const float (&t)[4]) const {
    if (naive) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float d = p[c] - t[c];
        acc[0] = fmaf(d, d, acc[0]);
      }
    } else {
      const float at = (t[3] + 1.0f) * 0.5f, ap = (p[3] + 1.0f) * 0.5f, da = at - ap;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // separately rounded products (no fma contraction): identical inputs give exactly 0
        float d = __fsub_rn(__fmul_rn(t[c], at), __fmul_rn(p[c], ap));
        acc[0] += d * d - 2.0f * eb[c] * d * da + eb2[c] * da * da;
      }
    }
  }
  __device__ __forceinline__ void finish(const double (&s)[NV], float* out, double) const { out[0] = (float)s[0]; }
  __host__ __device__ int out_stride() const { return 1; }
};

// composite_over_background (src/models/rgba_vae.py:75-84) of recon and target for each background, squared error
// over (3,H,W) -> compute_psnr (rgba_vae_stage.py:712-715), and |alpha_recon - alpha_target| (rgba_vae_stage.py:749-753).
struct PsnrOp {
  static constexpr int NV = MAX_BG + 1;
  float bg[MAX_BG][3];
  int nbg;
  __device__ __forceinline__ void pixel(float (&acc)[NV], const float (&p)[4], const float (&t)[4]) const {
    const float ap = p[3], at = t[3];
    acc[MAX_BG] += fabsf(ap - at);
#pragma unroll
    for (int b = 0; b < MAX_BG; ++b) {
      if (b < nbg) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float cp = p[c] * ap + bg[b][c] * (1.0f - ap);
          float ct = t[c] * at + bg[b][c] * (1.0f - at);
          float d = cp - ct;
          acc[b] = fmaf(d, d, acc[b]);
        }
      }
    }
  }
  __device__ __forceinline__ void finish(const double (&s)[NV], float* out, double hw) const {
    for (int k = 0; k < nbg; ++k) {
      double mse = s[k] / (3.0 * hw);
      if (mse < 1e-8) mse = 1e-8;
      out[k] = (float)(-10.0 * log10(mse));
    }
    out[nbg] = (float)(s[MAX_BG] / hw);
  }
  __host__ __device__ int out_stride() const { return nbg + 1; }
};

// The per-sample sums RgbaVAE.loss combines (src/models/rgba_vae.py:283-316); recon / target in [0, 1]:
//  [0] AlphaVAE map on the rescaled (2x-1) pair   [1] (recon_rgb - target_rgb)^2   [2] white-composite squared error
//  [3] black-composite squared error               [4] (alpha_r - alpha_t)^2        [5] |alpha_r - alpha_t|
struct TermsOp {
  static constexpr int NV = 6;
  float eb[3], eb2[3];
  __device__ __forceinline__ void pixel(float (&acc)[NV], const float (&p)[4], const float (&t)[4]) const {
    const float ap = p[3], at = t[3], da = at - ap;  // ((2a-1)+1)/2 == a: the rescaled alpha is the [0,1] alpha
    acc[4] = fmaf(da, da, acc[4]);
    acc[5] += fabsf(da);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float ps = p[c] * 2.0f - 1.0f, ts = t[c] * 2.0f - 1.0f;
      float d = __fsub_rn(__fmul_rn(ts, at), __fmul_rn(ps, ap));
      acc[0] += d * d - 2.0f * eb[c] * d * da + eb2[c] * da * da;
      const float dn = p[c] - t[c];
      acc[1] = fmaf(dn, dn, acc[1]);
      const float dw = (p[c] * ap + (1.0f - ap)) - (t[c] * at + (1.0f - at));
      acc[2] = fmaf(dw, dw, acc[2]);
      const float db = p[c] * ap - t[c] * at;
      acc[3] = fmaf(db, db, acc[3]);
    }
  }
  __device__ __forceinline__ void finish(const double (&s)[NV], float* out, double) const {
#pragma unroll
    for (int k = 0; k < NV; ++k) out[k] = (float)s[k];
  }
  __host__ __device__ int out_stride() const { return NV; }
};

// grid = (blocks per sample, samples).  tickets == nullptr: partials only (finish_kernel follows).
template <typename T, int VEC, typename Op>
__global__ void __launch_bounds__(256, 4) pair_reduce_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                             double* __restrict__ partial, unsigned int* tickets,
                                                             float* __restrict__ out, int64_t hw, Op op) {
  constexpr int NV = Op::NV;
  const int n = blockIdx.y;
  const T* p = a + (int64_t)n * 4 * hw;
  const T* t = b + (int64_t)n * 4 * hw;
  float acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = 0.f;
  const int64_t nvec = hw / VEC;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += stride) {
    PlaneVec<T, VEC> pv, tv;
    pv.load(p, hw, i);
    tv.load(t, hw, i);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float pj[4], tj[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        pj[c] = pv.get(c, j);
        tj[c] = tv.get(c, j);
      }
      op.pixel(acc, pj, tj);
    }
  }
  double accd[NV], sum[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) accd[k] = (double)acc[k];
  block_sum_d<NV>(accd, sum);
  double* mine = partial + ((int64_t)n * gridDim.x + blockIdx.x) * NV;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) mine[k] = sum[k];
  }
  if (tickets == nullptr) return;
  __shared__ unsigned int s_ticket;
  if (threadIdx.x == 0) {
    __threadfence();  // partial visible device-wide before the ticket is drawn
    s_ticket = atomicAdd(&tickets[n], 1u);
  }
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  // last block of this sample: fixed-order sum of all partials (thread i takes blocks i, i+256, ...)
#pragma unroll
  for (int k = 0; k < NV; ++k) accd[k] = 0.0;
  const double* all = partial + (int64_t)n * gridDim.x * NV;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < NV; ++k) accd[k] += __ldcg(all + (int64_t)i * NV + k);
  }
  block_sum_d<NV>(accd, sum);
  if (threadIdx.x == 0) {
    op.finish(sum, out + (int64_t)n * op.out_stride(), (double)hw);
    tickets[n] = 0u;
  }
}

// Second launch of the fallback path (more than TICKET_SLOTS streams or TICKET_BATCH samples): same summation order.
template <typename Op>
__global__ void __launch_bounds__(256) finish_kernel(const double* __restrict__ partial, float* __restrict__ out, int blocks,
                                                    int64_t hw, Op op) {
  constexpr int NV = Op::NV;
  const int n = blockIdx.x;
  double accd[NV], sum[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) accd[k] = 0.0;
  const double* all = partial + (int64_t)n * blocks * NV;
  for (int i = threadIdx.x; i < blocks; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < NV; ++k) accd[k] += all[(int64_t)i * NV + k];
  }
  block_sum_d<NV>(accd, sum);
  if (threadIdx.x == 0) op.finish(sum, out + (int64_t)n * op.out_stride(), (double)hw);
}

template <typename Op>
static int launch_pair_reduce(int cat, const void* a, const void* b, double* partial, float* out, int n, int64_t hw,
                              int dtype, cudaStream_t st, const Op& op) {
  const int blocks = reduce_blocks(hw);
  dim3 grid(blocks, n);
  const size_t es = dtype == RV_F32 ? 4 : 2;
  const int vec = 16 / (int)es;
  const bool vec_ok = hw % vec == 0 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0);
  unsigned int* tickets = nullptr;
  const int slot = n <= TICKET_BATCH ? ticket_slot(st) : -1;
  if (slot >= 0) {
    void* base = nullptr;
    RV_CUDA(cudaGetSymbolAddress(&base, g_tickets));
    tickets = (unsigned int*)base + (size_t)slot * TICKET_BATCH;
  }
  {
    LaunchScope scope(cat, st, 8.0 * (double)n * hw * es);
    if (dtype == RV_F32) {
      if (vec_ok) pair_reduce_kernel<float, 4, Op><<<grid, 256, 0, st>>>((const float*)a, (const float*)b, partial, tickets, out, hw, op);
      else pair_reduce_kernel<float, 1, Op><<<grid, 256, 0, st>>>((const float*)a, (const float*)b, partial, tickets, out, hw, op);
    } else {
      if (vec_ok)
        pair_reduce_kernel<__nv_bfloat16, 8, Op><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, partial,
                                                                      tickets, out, hw, op);
      else
        pair_reduce_kernel<__nv_bfloat16, 1, Op><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, partial,
                                                                      tickets, out, hw, op);
    }
    RV_LAUNCH_CHECK();
  }
  if (tickets == nullptr) {
    LaunchScope scope(cat, st, 0.0);
    finish_kernel<Op><<<n, 256, 0, st>>>(partial, out, blocks, hw, op);
    RV_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace rv

extern "C" {

int rv_reduce_blocks(int64_t hw) { return rv::reduce_blocks(hw); }

int rv_recon_loss(const void* pred, const void* target, const float* eb_host, const float* eb2_host, int naive_mse,
                  float* per_sample, double* partial, int n, int64_t hw, int dtype, void* stream) {
  RV_CHECK_ARG(pred && target && per_sample && partial && n > 0 && hw > 0, "recon_loss: bad argument");
  RV_CHECK_ARG(naive_mse || (eb_host && eb2_host), "recon_loss: Eb / Eb2 missing");
  RV_CHECK_ARG(dtype == RV_F32 || dtype == RV_BF16, "recon_loss: bad dtype %d", dtype);
  rv::LossOp op;
  for (int c = 0; c < 3; ++c) {
    op.eb[c] = eb_host ? eb_host[c] : 0.f;
    op.eb2[c] = eb2_host ? eb2_host[c] : 0.f;
  }
  op.naive = naive_mse;
  return rv::launch_pair_reduce(rv::CAT_LOSS, pred, target, partial, per_sample, n, hw, dtype, (cudaStream_t)stream, op);
}

int rv_composite_psnr(const void* recon, const void* target, const float* bgs_host, int nbg, float* out,
                      double* partial, int n, int64_t hw, int dtype, void* stream) {
  RV_CHECK_ARG(recon && target && out && partial && n > 0 && hw > 0, "composite_psnr: bad argument");
  RV_CHECK_ARG(nbg >= 0 && nbg <= rv::MAX_BG && (nbg == 0 || bgs_host), "composite_psnr: 0..%d backgrounds", rv::MAX_BG);
  RV_CHECK_ARG(dtype == RV_F32 || dtype == RV_BF16, "composite_psnr: bad dtype %d", dtype);
  rv::PsnrOp op;
  op.nbg = nbg;
  for (int b = 0; b < rv::MAX_BG; ++b)
    for (int c = 0; c < 3; ++c) op.bg[b][c] = b < nbg ? bgs_host[b * 3 + c] : 0.f;
  return rv::launch_pair_reduce(rv::CAT_PSNR, recon, target, partial, out, n, hw, dtype, (cudaStream_t)stream, op);
}

int rv_rgba_loss_terms(const void* recon, const void* target, const float* eb_host, const float* eb2_host, float* out,
                       double* partial, int n, int64_t hw, int dtype, void* stream) {
  RV_CHECK_ARG(recon && target && out && partial && eb_host && eb2_host && n > 0 && hw > 0, "rgba_loss_terms: bad argument");
  RV_CHECK_ARG(dtype == RV_F32 || dtype == RV_BF16, "rgba_loss_terms: bad dtype %d", dtype);
  rv::TermsOp op;
  for (int c = 0; c < 3; ++c) {
    op.eb[c] = eb_host[c];
    op.eb2[c] = eb2_host[c];
  }
  return rv::launch_pair_reduce(rv::CAT_LOSS, recon, target, partial, out, n, hw, dtype, (cudaStream_t)stream, op);
}

}  // extern "C"
