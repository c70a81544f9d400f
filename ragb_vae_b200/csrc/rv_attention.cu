// Fused single-head attention for the VAE mid block (sm_100a): O = softmax(Q K^T / sqrt(d)) V, d = 384 (Qwen) or 512 (Flux).
//
// Replaces scaled_dot_product_attention inside diffusers' QwenImageAttentionBlock (mid block of
// vae.encode / vae.decode; reference call sites src/models/rgba_vae.py:277,279).  The unfused path
// (QK^T GEMM -> fp32 scores in HBM -> softmax kernel -> PV GEMM) moves ~3 GB per 16 384-token image;
// here scores and probabilities never leave the SM.
//
// One CTA = 128 queries of one image; it walks the keys in blocks of 128:
//   warp 0      TMA producer: Q once (6 x [128 x 64] SW128 chunks), then per key block the K chunks
//               (6 x 16 KB) and V^T chunks (6 x [128 d x 64 keys] = 16 KB) through one 5-slot ring,
//               in exactly the order the MMA warp consumes them
//   warp 1      tcgen05.mma issuer for S = Q K^T into TMEM columns [384, 512)
//   warp 18     tcgen05.mma issuer for O += P V into [0, 384): two issuing threads, because a tcgen05.mma costs its issuer a
//               fixed ~46 cycles on top of the N/2 the tensor pipe needs -- with one issuer the 48 N = 128
//               MMAs of a key block serialise that cost; S(j+1) and O(j) are independent accumulators, so the two streams overlap
//   warps 2-17  softmax (thread = query row; the four warps of a lane quarter split the 128 keys):
//               S -> registers, running max with lazy rescale (O is only rescaled when the max grows by
//               more than 2^8), P = exp2(...) written to shared memory as the bf16 K-major SW128 A
//               operand of the PV MMA; finally O / l -> bf16 global
// TMEM: O 384 fp32 columns + S 128 = 512.  Shared memory: Q 96 KB + ring 80 KB + P 32 KB.
//
// d = 512 (diffusers Attention of the Flux AutoencoderKL mid block): O alone would fill TMEM, so the key loop runs TWICE per
// query block, each pass producing 256 of the 512 output columns.  Q (8 chunks, 128 KB) stays resident, the ring shrinks to
// 3 slots.  Without a workspace the second pass recomputes S and the softmax (1.5x the MMA work of an ideal kernel).  With one
// (rv_attention_ws, SPILL = true) pass 1 also leaves every probability tile -- the 32 KB shared-memory image the PV MMA reads,
// plus the running maxima -- in a per-CTA slot of the workspace, and pass 2 is a pure P V stream: the tiles come back by TMA
// (the workspace seen as 128-byte rows, no swizzle: the image is already swizzled) through a 13-slot ring laid over the Q, ring
// and P regions, the O rescales of pass 1 are replayed at the (rare) blocks where a row's maximum moved, no S, no softmax:
// 64 instead of 96 MMAs per key block.  That form runs as CTA pairs (PAIR below) when tokens % 256 == 0.
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "rv_tc_common.cuh"

namespace rv {

constexpr int FA_BQ = 128, FA_BK = 128;
constexpr int FA_RING = 5;               // ring slots at d = 384 (3 at d = 512: Q takes 128 KB)
constexpr uint32_t FA_SLOT = 16384;      // ring slot: a K chunk [128 keys x 64 d] or a V^T chunk [128 d x 64 keys]
                                         // (CTA pairs: each CTA holds half of the 128 rows, 8 KB)
constexpr uint32_t FA_P_BYTES = 2 * 16384;
constexpr int FA_NSPLIT = 4;             // softmax warps per TMEM lane quarter (each owns 128/NSPLIT keys of a block)
constexpr int FA_SM_WARPS = 4 * FA_NSPLIT;
constexpr int FA_THREADS = 96 + 32 * FA_SM_WARPS;   // warps: 0 TMA, 1 MMA (S = Q K^T), 2..17 softmax, 18 MMA (O += P V)
constexpr int FA_PV_WARP = 2 + FA_SM_WARPS;
constexpr int FA_KCOLS = 128 / FA_NSPLIT;  // S columns per softmax thread
constexpr float FA_RESCALE_THRESHOLD = 8.0f;

struct FaParams {
  int tokens;          // per image, multiple of 128
  int n_img;
  int ld_out;          // row pitch of O in elements
  float scale_log2;    // 1/sqrt(d) * log2(e)
  __nv_bfloat16* out;  // [n_img*tokens][ld_out]
  float* lse;          // optional [n_img*tokens]: log2-domain log-sum-exp of the scaled scores, P = exp2(s * scale_log2 - lse)
  uint8_t* ws;         // SPILL: ws_slots slots of (tokens / 128) * FA_WS_BLOCK bytes
  int* ws_flags;       // SPILL: one int per slot, 0 = free (a CTA owns a slot from its first to its last instruction)
  int ws_slots;
};

constexpr int FA_RING2 = 13;                 // pass-2 ring of the SPILL form: Q (8) + ring (3) + P (2) slots of 16 KB
constexpr uint32_t FA_WS_BLOCK = 32768 + 512;  // workspace per key block: the P tile's shared-memory image + m_used[128]
constexpr int FA_WS_SLOTS = 160;             // > 148 co-resident CTAs (one per SM: 209 KB of shared memory each)
constexpr int FA_MAX_NB = 1024;              // SPILL: key blocks per image (rescale flags live in shared memory)

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// PAIR: two CTAs (256 queries of one image) issue M = 256 MMAs (cta_group::2) and split every K / V^T chunk between their shared
// memories: each CTA streams HALF the operand bytes, so the same ring bytes hold twice as many (8 KB) slots.  The kernel is
// paced by how far ahead of the MMAs its ring can run (measured: the d = 384 kernel with 3 instead of 5 slots drops from 1053
// to 774 TFLOP/s, and d = 512 has room for 3), which is what the pair form buys -- with the SAME number of slots it was no
// faster (933 vs 958 TFLOP/s, rounds 1-2).  Protocol as in conv_tc2_kernel: both CTAs' TMA loads complete on the LEADER's full
// barriers, only the leader issues MMAs, tcgen05.commit is multicast to both CTAs' barriers, the peer's softmax warps arrive
// remotely on the leader's sempty / pfull.
// DCH: 64-wide chunks of d; OPARTS: N = 128 parts of O per pass; NPASS: key-loop passes; RING: K / V^T ring slots
// (6, 3, 1, 5 for d = 384; 8, 2, 2, 3 for d = 512, 6 half-size slots in its pair form) -- compile-time so that the d = 384
// loops stay fully unrolled
template <bool PAIR, int DCH, int OPARTS, int NPASS, int RING, bool SPILL = false>
__global__ void __launch_bounds__(FA_THREADS, 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                  const __grid_constant__ CUtensorMap map_vt, const __grid_constant__ CUtensorMap map_ws,
                  const __grid_constant__ FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  // One operand ring, two consumers (the S and the O issuer): a slot's "full" signal goes to the barrier of the KIND of chunk it
  // holds, so each consumer's parity wait can only ever see its own loads (with one shared barrier per slot a consumer that is
  // a whole phase behind would alias phases k and k+2).
  __shared__ __align__(8) uint64_t bar_q, bar_fullk[RING], bar_fullv[RING], bar_empty[RING], bar_sfull, bar_sempty, bar_pfull,
      bar_pvdone;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_xmax[2][FA_NSPLIT][128];  // [block parity][column part][row]
  __shared__ float s_xsum[FA_NSPLIT][128];
  // SPILL form only
  __shared__ __align__(8) uint64_t bar_full2[SPILL ? FA_RING2 : 1], bar_empty2[SPILL ? FA_RING2 : 1], bar_p1done, bar_go, bar_resc,
      bar_p2done;
  __shared__ uint8_t s_resc[SPILL ? FA_MAX_NB : 1];  // block j of pass 1 rescaled O for some row
  __shared__ int s_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t q_bytes = (uint32_t)DCH * 16384u;
  constexpr uint32_t slot_bytes = PAIR ? FA_SLOT / 2 : FA_SLOT;
  constexpr int nshare = PAIR ? 2 : 1;
  const uint32_t q_smem = base, ring = base + q_bytes, p_smem = ring + (uint32_t)RING * slot_bytes;
  constexpr uint32_t nring = (uint32_t)RING;
  constexpr int ocols = OPARTS * 128 / FA_NSPLIT;  // O columns per softmax thread (rescale / output)
  const int nb = p.tokens / FA_BK;
  const int qblocks = p.tokens / FA_BQ;
  const int img = blockIdx.x / qblocks;
  const int q0 = (blockIdx.x - img * qblocks) * FA_BQ;   // consecutive blocks = the two CTAs of a pair
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    mbar_init(smem_u32(&bar_q), 1);
    for (int s = 0; s < RING; ++s) {
      mbar_init(smem_u32(&bar_fullk[s]), 1);
      mbar_init(smem_u32(&bar_fullv[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_sfull), 1);
    mbar_init(smem_u32(&bar_sempty), FA_SM_WARPS * nshare);
    mbar_init(smem_u32(&bar_pfull), FA_SM_WARPS * nshare);
    mbar_init(smem_u32(&bar_pvdone), 1);
    if (SPILL) {
      for (int s = 0; s < FA_RING2; ++s) {
        mbar_init(smem_u32(&bar_full2[s]), 1);
        mbar_init(smem_u32(&bar_empty2[s]), 1);
      }
      mbar_init(smem_u32(&bar_p1done), FA_SM_WARPS * nshare);
      mbar_init(smem_u32(&bar_go), 1);
      mbar_init(smem_u32(&bar_resc), FA_SM_WARPS * nshare);
      mbar_init(smem_u32(&bar_p2done), 1);
      int sl = (int)(blockIdx.x % (unsigned)p.ws_slots);
      while (atomicCAS(p.ws_flags + sl, 0, 1) != 0) sl = sl + 1 == p.ws_slots ? 0 : sl + 1;
      s_slot = sl;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (SPILL)
    for (int i = threadIdx.x; i < FA_MAX_NB; i += FA_THREADS) s_resc[i] = 0;
  if (warp == 1) {
    __syncwarp();
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t fullk0 = smem_u32(&bar_fullk[0]), fullv0 = smem_u32(&bar_fullv[0]), empty0 = smem_u32(&bar_empty[0]);
  const uint32_t sfull = smem_u32(&bar_sfull), sempty = smem_u32(&bar_sempty), pfull = smem_u32(&bar_pfull),
                 pvdone = smem_u32(&bar_pvdone), qbar = smem_u32(&bar_q);
  const uint32_t full20 = smem_u32(&bar_full2[0]), empty20 = smem_u32(&bar_empty2[0]), p1done = smem_u32(&bar_p1done),
                 gobar = smem_u32(&bar_go), rescbar = smem_u32(&bar_resc), p2done = smem_u32(&bar_p2done);
  constexpr int NP1 = SPILL ? 1 : NPASS;  // passes that compute S (the SPILL form's second pass replays stored P tiles)
  uint8_t* const ws = SPILL ? p.ws + (size_t)s_slot * (size_t)nb * FA_WS_BLOCK : nullptr;
  // barriers the MMA issuer (leader CTA) waits on, as cluster addresses, for signals that come from both CTAs
  const uint32_t lead_fullk0 = PAIR ? mapa_rank(fullk0, 0) : fullk0;
  const uint32_t lead_fullv0 = PAIR ? mapa_rank(fullv0, 0) : fullv0;
  const uint32_t lead_q = PAIR ? mapa_rank(qbar, 0) : qbar;
  const uint32_t lead_sempty = PAIR ? mapa_rank(sempty, 0) : sempty;
  const uint32_t lead_pfull = PAIR ? mapa_rank(pfull, 0) : pfull;
  const uint32_t lead_full20 = PAIR ? mapa_rank(full20, 0) : full20;
  const uint32_t lead_resc = PAIR ? mapa_rank(rescbar, 0) : rescbar;
  const uint32_t peer_p1done = PAIR ? mapa_rank(p1done, rank ^ 1u) : p1done;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      if (leader) mbar_arrive_expect_tx(qbar, q_bytes * nshare);
      for (int c = 0; c < DCH; ++c) {
        if (PAIR) tma2_load_2d(q_smem + c * 16384u, &map_q, lead_q, c * 64, img * p.tokens + q0);
        else tma_load_2d(q_smem + c * 16384u, &map_q, qbar, c * 64, img * p.tokens + q0);
      }
    }
    __syncwarp();
    uint32_t slot = 0, par = 0;
    // consumption order per pass: K(0), K(1), V(0), K(2), V(1), ..., K(nb-1), V(nb-2), V(nb-1)
    for (int pass = 0; pass < NP1; ++pass)
    for (int step = 0; step <= nb; ++step) {
      if (step < nb) {
        for (int c = 0; c < DCH; ++c) {
          mbar_wait(empty0 + 8u * slot, par ^ 1u);
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(fullk0 + 8u * slot, slot_bytes * nshare);
            if (PAIR)   // this CTA's 64 of the block's 128 keys
              tma2_load_2d(ring + slot * slot_bytes, &map_k, lead_fullk0 + 8u * slot, c * 64,
                           img * p.tokens + step * FA_BK + (int)rank * 64);
            else
              tma_load_2d(ring + slot * slot_bytes, &map_k, fullk0 + 8u * slot, c * 64, img * p.tokens + step * FA_BK);
          }
          __syncwarp();
          if (++slot == nring) { slot = 0; par ^= 1u; }
        }
      }
      if (step >= 1) {
        const int j = step - 1;
        const int vrow0 = pass * OPARTS * 128;  // first d-row (output column) of this pass
        for (int kc = 0; kc < 2; ++kc)
          for (int h = 0; h < OPARTS; ++h) {
            mbar_wait(empty0 + 8u * slot, par ^ 1u);
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(fullv0 + 8u * slot, slot_bytes * nshare);
              if (PAIR)  // this CTA's 64 of the 128 d-rows of the chunk
                tma2_load_3d(ring + slot * slot_bytes, &map_vt, lead_fullv0 + 8u * slot, j * FA_BK + kc * 64,
                             vrow0 + h * 128 + (int)rank * 64, img);
              else
                tma_load_3d(ring + slot * slot_bytes, &map_vt, fullv0 + 8u * slot, j * FA_BK + kc * 64, vrow0 + h * 128, img);
            }
            __syncwarp();
            if (++slot == nring) { slot = 0; par ^= 1u; }
          }
      }
    }
    if (SPILL) {
      // pass 2: per key block and 64-key half the stored P chunk, then the two V^T chunks of output columns [256, 512)
      mbar_wait(p1done, 0);  // every P tile is in the workspace (and fenced towards the async proxy); Q, ring and P are free
      uint32_t s2 = 0, par2 = 0;
      auto advance2 = [&]() { if (++s2 == (uint32_t)FA_RING2) { s2 = 0; par2 ^= 1u; } };
      for (int j = 0; j < nb; ++j)
        for (int kc = 0; kc < 2; ++kc) {
          mbar_wait(empty20 + 8u * s2, par2 ^ 1u);
          if (elect_one()) {
            // this CTA's own tile (its 128 rows of P): a [128 rows x 128 B] box of the workspace seen as 128-byte rows
            const int wrow = (int)(((size_t)(ws - p.ws) + (size_t)j * FA_WS_BLOCK + (size_t)kc * FA_SLOT) >> 7);
            if (leader) mbar_arrive_expect_tx(full20 + 8u * s2, FA_SLOT * nshare);
            if (PAIR) tma2_load_2d(base + s2 * FA_SLOT, &map_ws, lead_full20 + 8u * s2, 0, wrow);
            else tma_load_2d(base + s2 * FA_SLOT, &map_ws, full20 + 8u * s2, 0, wrow);
          }
          __syncwarp();
          advance2();
          if (PAIR) {
            // ONE slot per 64-key half: this CTA's 128 of the 256 d-rows [256, 512) as two 64-row boxes -- the pair's PV MMA is
            // N = 256 (8 instead of 16 MMAs per key block, and two ring slots per half instead of three)
            mbar_wait(empty20 + 8u * s2, par2 ^ 1u);
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(full20 + 8u * s2, FA_SLOT * nshare);
              for (int hh = 0; hh < 2; ++hh)
                tma2_load_3d(base + s2 * FA_SLOT + (uint32_t)hh * (FA_SLOT / 2), &map_vt, lead_full20 + 8u * s2, j * FA_BK + kc * 64,
                             OPARTS * 128 + (int)rank * 128 + hh * 64, img);
            }
            __syncwarp();
            advance2();
          } else {
            for (int h = 0; h < OPARTS; ++h) {
              mbar_wait(empty20 + 8u * s2, par2 ^ 1u);
              if (elect_one()) {
                mbar_arrive_expect_tx(full20 + 8u * s2, FA_SLOT);
                tma_load_3d(base + s2 * FA_SLOT, &map_vt, full20 + 8u * s2, j * FA_BK + kc * 64, OPARTS * 128 + h * 128, img);
              }
              __syncwarp();
              advance2();
            }
          }
        }
    }
  } else if (warp == 1 || warp == FA_PV_WARP) {
    // ------------------------------ MMA issuers ------------------------------
    // Both walk the producer's push sequence (per pass: K(0) | K(1) V(0) | K(2) V(1) | ... | V(nb-1)); each consumes its own
    // kind of ring slot and only counts past the other's.
    const bool is_qk = warp == 1;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
    const uint64_t hi = make_smem_desc(0u, 1024u, 2u);
    const uint32_t q_lo = (q_smem & 0x3FFFFu) >> 4, p_lo = (p_smem & 0x3FFFFu) >> 4, ring_lo = (ring & 0x3FFFFu) >> 4;
    const uint32_t s_tmem = tmem_base + 384u;
    uint32_t slot = 0, use_par = 0;  // bit s of use_par: parity of this issuer's next use of slot s
    auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t accf) {
      if (PAIR) umma2_bf16(d, a, b, idesc, accf);
      else umma_bf16(d, a, b, idesc, accf);
    };
    auto commit = [&](uint32_t bar) {
      if (PAIR) umma2_commit_both(bar);
      else umma_commit(bar);
    };
    auto skip = [&](int n) {
      for (int i = 0; i < n; ++i)
        if (++slot == nring) slot = 0;
    };
    if (leader) {
      if (is_qk) mbar_wait(qbar, 0);
      int g = 0;  // key blocks issued so far over all passes: every per-block barrier flips once per block
      for (int pass = 0; pass < NP1; ++pass)
        for (int step = 0; step <= nb; ++step) {
          if (step < nb) {
            if (is_qk) {
              if (g > 0) {  // S(previous block) is in the softmax warps' registers: the S columns may be overwritten
                mbar_wait(sempty, (uint32_t)((g - 1) & 1));
                tc_fence_after();
              }
              // S = Q K(step)^T : DCH chunks x 4 k-steps, N = 128
              for (int c = 0; c < DCH; ++c) {
                mbar_wait(fullk0 + 8u * slot, (use_par >> slot) & 1u);
                use_par ^= 1u << slot;
                tc_fence_after();
                if (elect_one()) {
                  const uint64_t ad = hi | (uint64_t)(q_lo + c * 1024u);
                  const uint64_t bd = hi | (uint64_t)(ring_lo + slot * (slot_bytes >> 4));
                  mma(s_tmem, ad, bd, c == 0 ? 0u : 1u);
                  mma(s_tmem, ad + 2u, bd + 2u, 1u);
                  mma(s_tmem, ad + 4u, bd + 4u, 1u);
                  mma(s_tmem, ad + 6u, bd + 6u, 1u);
                  commit(empty0 + 8u * slot);
                  if (c == DCH - 1) commit(sfull);
                }
                __syncwarp();
                if (++slot == nring) slot = 0;
              }
              ++g;
            } else {
              skip(DCH);
            }
          }
          if (step >= 1) {
            if (!is_qk) {
              const int j = step - 1;
              mbar_wait(pfull, (uint32_t)(g & 1));  // P(j) is in shared memory (and O has been rescaled if needed)
              tc_fence_after();
              for (int kc = 0; kc < 2; ++kc)
                for (int h = 0; h < OPARTS; ++h) {
                  mbar_wait(fullv0 + 8u * slot, (use_par >> slot) & 1u);
                  use_par ^= 1u << slot;
                  tc_fence_after();
                  if (elect_one()) {
                    const uint64_t ad = hi | (uint64_t)(p_lo + kc * 1024u);
                    const uint64_t bd = hi | (uint64_t)(ring_lo + slot * (slot_bytes >> 4));
                    const uint32_t d_tmem = tmem_base + (uint32_t)h * 128u;
                    mma(d_tmem, ad, bd, (j == 0 && kc == 0) ? 0u : 1u);
                    mma(d_tmem, ad + 2u, bd + 2u, 1u);
                    mma(d_tmem, ad + 4u, bd + 4u, 1u);
                    mma(d_tmem, ad + 6u, bd + 6u, 1u);
                    commit(empty0 + 8u * slot);
                    if (kc == 1 && h == OPARTS - 1) commit(pvdone);
                  }
                  __syncwarp();
                  if (++slot == nring) slot = 0;
                }
              ++g;
            } else {
              skip(2 * OPARTS);
            }
          }
        }
      if (SPILL && !is_qk) {
        // pass 2: O[:, 256:512) += P(j) V(j)[:, 256:512) from the stored tiles, into TMEM columns [256, 512)
        mbar_wait(p1done, 0);  // also orders the s_resc flags written in pass 1 (by both CTAs of a pair)
        if (PAIR) asm volatile("fence.acq_rel.cluster;" ::: "memory");
        tc_fence_after();
        const uint32_t base_lo = (base & 0x3FFFFu) >> 4;
        uint32_t s2 = 0, par2 = 0, rpar = 0;
        auto advance2 = [&]() { if (++s2 == (uint32_t)FA_RING2) { s2 = 0; par2 ^= 1u; } };
        for (int j = 0; j < nb; ++j) {
          if (s_resc[j]) {  // replay of a pass-1 rescale: the softmax warps scale O once PV(j-1) has completed
            if (elect_one()) commit(gobar);
            __syncwarp();
            mbar_wait(rescbar, rpar);
            rpar ^= 1u;
            tc_fence_after();
          }
          for (int kc = 0; kc < 2; ++kc) {
            const uint32_t ps = s2;
            mbar_wait(full20 + 8u * ps, par2);
            advance2();
            constexpr int H2 = PAIR ? 1 : OPARTS;  // pair form: one N = 256 MMA covers both 128-column parts
            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((256u >> 4) << 24);
            for (int h = 0; h < H2; ++h) {
              mbar_wait(full20 + 8u * s2, par2);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t ad = hi | (uint64_t)(base_lo + ps * (FA_SLOT >> 4));
                const uint64_t bd = hi | (uint64_t)(base_lo + s2 * (FA_SLOT >> 4));
                const uint32_t d_tmem = tmem_base + (uint32_t)(OPARTS + h) * 128u;
                const uint32_t acc0 = (j == 0 && kc == 0) ? 0u : 1u;
                if (PAIR) {
                  umma2_bf16(d_tmem, ad, bd, idesc2, acc0);
                  umma2_bf16(d_tmem, ad + 2u, bd + 2u, idesc2, 1u);
                  umma2_bf16(d_tmem, ad + 4u, bd + 4u, idesc2, 1u);
                  umma2_bf16(d_tmem, ad + 6u, bd + 6u, idesc2, 1u);
                } else {
                  mma(d_tmem, ad, bd, acc0);
                  mma(d_tmem, ad + 2u, bd + 2u, 1u);
                  mma(d_tmem, ad + 4u, bd + 4u, 1u);
                  mma(d_tmem, ad + 6u, bd + 6u, 1u);
                }
                commit(empty20 + 8u * s2);
                if (h == H2 - 1) commit(empty20 + 8u * ps);
                if (j == nb - 1 && kc == 1 && h == H2 - 1) commit(p2done);
              }
              __syncwarp();
              advance2();
            }
          }
        }
      }
    }
  } else if (warp >= 2 && warp < FA_PV_WARP) {
    // ------------------------------ softmax / correction / output ------------------------------
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t s_taddr = lane_addr + 384u + (uint32_t)part * FA_KCOLS;
    const uint32_t o_taddr = lane_addr + (uint32_t)(part * ocols);
    const int bar_id = 1 + q;
    // P row: key chunk (64 keys, 16 KB) part/2, 16-byte units (part%2)*4 .. +4 of the row; unit u lives at ((u ^ (row & 7)) * 16)
    const uint32_t p_row = p_smem + (uint32_t)(part >> 1) * 16384u + (uint32_t)row * 128u;
    const uint32_t u0 = (uint32_t)(part & 1) * 4u;
    const uint32_t sw = (uint32_t)(row & 7);
    for (int pass = 0, jj = 0; pass < NP1; ++pass) {
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < nb; ++j, ++jj) {
        mbar_wait(sfull, (uint32_t)(jj & 1));
        tc_fence_after();
        uint32_t s0[FA_KCOLS];
        tmem_ld32(s_taddr, s0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(lead_sempty); else mbar_arrive(sempty); }
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < FA_KCOLS; ++i) mx = fmaxf(mx, __uint_as_float(s0[i]));
        s_xmax[jj & 1][part][row] = mx;
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 * FA_NSPLIT) : "memory");
#pragma unroll
        for (int pp = 0; pp < FA_NSPLIT; ++pp) mx = fmaxf(mx, s_xmax[jj & 1][pp][row]);
        mx *= p.scale_log2;
        float alpha = 1.0f;
        bool need = false;
        if (j == 0) {
          m_used = mx;
        } else if (mx - m_used > FA_RESCALE_THRESHOLD) {
          alpha = fast_exp2(m_used - mx);
          m_used = mx;
          l *= alpha;
          need = true;
        }
        // probabilities (bf16) and their row sum
        uint32_t pk[FA_KCOLS / 2];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < FA_KCOLS / 2; ++i) {
          const float a = fast_exp2(fmaf(__uint_as_float(s0[2 * i]), p.scale_log2, -m_used));
          const float b = fast_exp2(fmaf(__uint_as_float(s0[2 * i + 1]), p.scale_log2, -m_used));
          pk[i] = pack_bf16x2(a, b);
          sum += a + b;
        }
        l += sum;
        if (SPILL) {
          // the tile's shared-memory image (same swizzled offsets) and the maximum it is relative to go to the workspace
          uint8_t* wt = ws + (size_t)j * FA_WS_BLOCK;
          uint8_t* wrow = wt + (size_t)(part >> 1) * 16384u + (size_t)row * 128u;
#pragma unroll
          for (int u = 0; u < FA_KCOLS / 8; ++u)
            *reinterpret_cast<uint4*>(wrow + (((u0 + (uint32_t)u) ^ sw) << 4)) =
                make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
          if (part == 0) reinterpret_cast<float*>(wt + 32768)[row] = m_used;
          if (need) {
            s_resc[j] = 1;
            if (PAIR)  // the pair's PV MMAs stop for a rescale of either CTA's rows
              asm volatile("st.shared::cluster.u8 [%0], %1;" ::"r"(mapa_rank(smem_u32(&s_resc[j]), rank ^ 1u)), "h"((unsigned short)1) : "memory");
          }
        }
        // the previous PV must be complete before P is overwritten or O is rescaled (block 0 of a later pass: already
        // waited for in the previous pass's output stage)
        if (j > 0) {
          mbar_wait(pvdone, (uint32_t)((jj - 1) & 1));
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {
            for (int c = 0; c < ocols / 32; ++c) {
              uint32_t o[32];
              tmem_ld32(o_taddr + (uint32_t)c * 32u, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st32(o_taddr + (uint32_t)c * 32u, o);
            }
            tmem_st_wait();
          }
        }
#pragma unroll
        for (int u = 0; u < FA_KCOLS / 8; ++u) {
          const uint32_t addr = p_row + (((u0 + (uint32_t)u) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * u]), "r"(pk[4 * u + 1]), "r"(pk[4 * u + 2]),
                       "r"(pk[4 * u + 3])
                       : "memory");
        }
        if (PAIR) asm volatile("fence.proxy.async;" ::: "memory");
        else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(lead_pfull); else mbar_arrive(pfull); }
      }
      // ---- output of this pass: O / l into columns [pass * oparts * 128, ...) ----
      s_xsum[part][row] = l;
      mbar_wait(pvdone, (uint32_t)((jj - 1) & 1));
      tc_fence_after();
      if (SPILL) {  // hand the stored tiles to the async proxy (pass 2's bulk loads) and free Q / ring / P for its ring
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        if (PAIR) asm volatile("fence.acq_rel.cluster;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(p1done);
          if (PAIR) mbar_arrive_cluster(peer_p1done);
        }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 * FA_NSPLIT) : "memory");
      float lt = 0.f;
#pragma unroll
      for (int pp = 0; pp < FA_NSPLIT; ++pp) lt += s_xsum[pp][row];
      const float inv = 1.0f / lt;
      if (p.lse && part == 0 && pass == 0) p.lse[(int64_t)img * p.tokens + q0 + row] = m_used + log2f(lt);
      __nv_bfloat16* orow = p.out + ((int64_t)img * p.tokens + q0 + row) * p.ld_out + pass * OPARTS * 128 + part * ocols;
      for (int c = 0; c < ocols / 32; ++c) {
        uint32_t o[32];
        tmem_ld32(o_taddr + (uint32_t)c * 32u, o);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(o[i]) * inv;
        fast_store<32>(orow + c * 32, v);
      }
      if (SPILL) {
        // pass 2 runs on the MMA side alone; here only the replay of pass 1's O rescales (blocks where a row's maximum moved)
        const uint32_t o2_taddr = o_taddr + (uint32_t)(OPARTS * 128);
        uint32_t gpar = 0;
        mbar_wait(p1done, 0);  // all softmax warps' (of both CTAs of a pair) flags and maxima are in place
        if (PAIR) asm volatile("fence.acq_rel.cluster;" ::: "memory");
        for (int j = 1; j < nb; ++j) {
          if (!s_resc[j]) continue;
          const float m0 = reinterpret_cast<const volatile float*>(ws + (size_t)(j - 1) * FA_WS_BLOCK + 32768)[row];
          const float m1 = reinterpret_cast<const volatile float*>(ws + (size_t)j * FA_WS_BLOCK + 32768)[row];
          const float alpha = fast_exp2(m0 - m1);
          mbar_wait(gobar, gpar);  // PV(j-1) of pass 2 has completed; PV(j) waits for bar_resc
          gpar ^= 1u;
          tc_fence_after();
          if (__any_sync(0xffffffffu, m0 != m1)) {
            for (int c = 0; c < ocols / 32; ++c) {
              uint32_t o[32];
              tmem_ld32(o2_taddr + (uint32_t)c * 32u, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st32(o2_taddr + (uint32_t)c * 32u, o);
            }
            tmem_st_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (PAIR) mbar_arrive_cluster(lead_resc); else mbar_arrive(rescbar); }
        }
        mbar_wait(p2done, 0);
        tc_fence_after();
        for (int c = 0; c < ocols / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(o2_taddr + (uint32_t)c * 32u, o);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(o[i]) * inv;
          fast_store<32>(orow + OPARTS * 128 + c * 32, v);
        }
      }
      if (pass + 1 < NP1) {  // the next pass rewrites s_xsum and (through the MMA warp, after this thread's next pfull arrive) O
        tc_fence_before();
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 * FA_NSPLIT) : "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // nobody frees TMEM or exits while the peer may still signal / read
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (SPILL && threadIdx.x == 0) {  // every bulk load of this CTA has been consumed: the workspace slot may change hands
    __threadfence();
    atomicExch(p.ws_flags + s_slot, 0);
  }
}

int tc_encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box, CUtensorMapSwizzle sw);  // rv_conv_tc.cu
int tc_ensure_init();

static std::mutex g_fa_mu;
static bool g_fa_attr[64] = {false};

}  // namespace rv

extern "C" int rv_attention(const void* q, const void* k, int64_t ld_qk, const void* vt, void* out, int64_t ld_out, int n_img,
                            int tokens, int d, void* stream) {
  return rv_attention_lse(q, k, ld_qk, vt, out, ld_out, nullptr, n_img, tokens, d, stream);
}

extern "C" int rv_attention_lse(const void* q, const void* k, int64_t ld_qk, const void* vt, void* out, int64_t ld_out, float* lse,
                                int n_img, int tokens, int d, void* stream) {
  return rv_attention_ws(q, k, ld_qk, vt, tokens, (int64_t)d * tokens, out, ld_out, lse, nullptr, 0, n_img, tokens, d, stream);
}

extern "C" int64_t rv_attention_workspace_bytes(int tokens, int d) {
  if (d != 512 || tokens <= 0 || tokens % 128 != 0 || tokens / 128 > rv::FA_MAX_NB) return 0;
  return 1024 + (int64_t)rv::FA_WS_SLOTS * (tokens / 128) * (int64_t)rv::FA_WS_BLOCK;
}

extern "C" int rv_attention_ws(const void* q, const void* k, int64_t ld_qk, const void* vt, int64_t ld_vt, int64_t vt_img_pitch,
                               void* out, int64_t ld_out, float* lse, void* workspace, int64_t workspace_bytes, int n_img,
                               int tokens, int d, void* stream) {
  using namespace rv;
  if (int rc = tc_ensure_init()) return rc;
  RV_CHECK_ARG(q && k && vt && out && n_img > 0 && tokens > 0, "attention: bad argument");
  RV_CHECK_ARG(d == 384 || d == 512, "attention: the fused kernel is built for d = 384 or 512 (got %d)", d);
  RV_CHECK_ARG(tokens % 128 == 0, "attention: tokens (%d) must be a multiple of 128", tokens);
  RV_CHECK_ARG(ld_qk % 8 == 0 && ld_out % 8 == 0 && ld_qk >= d && ld_out >= d, "attention: pitches must be multiples of 8 and >= d");
  RV_CHECK_ARG(ld_vt % 8 == 0 && vt_img_pitch % 8 == 0 && ld_vt >= tokens && vt_img_pitch >= tokens,
               "attention: V^T pitches (row %lld, image %lld) must be multiples of 8 and >= tokens", (long long)ld_vt, (long long)vt_img_pitch);
  RV_CHECK_ARG(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)vt % 16 == 0) && ((uintptr_t)out % 16 == 0),
               "attention: tensors must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  // d = 512 with a workspace: pass 2 replays the probability tiles pass 1 left there instead of recomputing them
  const int64_t ws_need = rv_attention_workspace_bytes(tokens, d);
  const bool spill = workspace != nullptr && ws_need > 0;
  RV_CHECK_ARG(!spill || (workspace_bytes >= ws_need && (uintptr_t)workspace % 128 == 0),
               "attention: workspace of %lld bytes (128-byte aligned) needed, got %lld", (long long)ws_need, (long long)workspace_bytes);
  // CTA pairs (two adjacent query blocks of one image) where the ring is the bound: the d = 512 SPILL form (3 full-size slots:
  // 678 -> 784 TFLOP/s as a pair with 6 half-size ones).  Not d = 384, whose 5 slots suffice (pair with 10: 984 vs 1058).
  const bool pair = tokens % 256 == 0 && spill;
  const cuuint32_t share = pair ? 2u : 1u;
  CUtensorMap mq, mk, mv, mw;
  {
    cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)((int64_t)n_img * tokens)};
    cuuint64_t str[1] = {(cuuint64_t)ld_qk * 2u};
    cuuint32_t box[2] = {64, 128};
    if (int rc = tc_encode_map(&mq, q, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    cuuint32_t boxk[2] = {64, 128u / share};
    if (int rc = tc_encode_map(&mk, k, 2, dims, str, boxk, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)tokens, (cuuint64_t)d, (cuuint64_t)n_img};
    cuuint64_t str[2] = {(cuuint64_t)ld_vt * 2u, (cuuint64_t)vt_img_pitch * 2u};
    cuuint32_t box[3] = {64, 128u / share, 1};
    if (int rc = tc_encode_map(&mv, vt, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  if (spill) {  // the workspace's tile area as 128-byte rows: a [128 x 128 B] box is one half (64 keys) of a stored P tile, verbatim
    cuuint64_t dims[2] = {64, (cuuint64_t)((ws_need - 1024) / 128)};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, 128};
    if (int rc = tc_encode_map(&mw, (const uint8_t*)workspace + 1024, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  } else {
    mw = mq;
  }
  const size_t smem = (size_t)(d / 64) * 16384 + (size_t)(d == 384 ? FA_RING : 3) * FA_SLOT + FA_P_BYTES + 1024;
  {
    std::lock_guard<std::mutex> lk(g_fa_mu);
    int dev = 0;
    RV_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !g_fa_attr[dev]) {
      const int smem_max = 6 * 16384 + FA_RING * (int)FA_SLOT + (int)FA_P_BYTES + 1024;  // d = 384: 214 016 B; d = 512 needs the same
      RV_CUDA(cudaFuncSetAttribute(flash_attn_kernel<false, 6, 3, 1, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
      RV_CUDA(cudaFuncSetAttribute(flash_attn_kernel<false, 8, 2, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
      RV_CUDA(cudaFuncSetAttribute(flash_attn_kernel<false, 8, 2, 2, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
      RV_CUDA(cudaFuncSetAttribute(flash_attn_kernel<true, 8, 2, 2, 6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
      g_fa_attr[dev] = true;
    }
  }
  FaParams p;
  p.tokens = tokens;
  p.n_img = n_img;
  p.ld_out = (int)ld_out;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)d);
  p.out = (__nv_bfloat16*)out;
  p.lse = lse;
  p.ws_flags = (int*)workspace;
  p.ws = spill ? (uint8_t*)workspace + 1024 : nullptr;
  p.ws_slots = FA_WS_SLOTS;
  const int grid = n_img * (tokens / FA_BQ);
  LaunchScope scope(CAT_ATTN, st, 4.0 * (double)n_img * tokens * tokens * d);
  if (pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(FA_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RV_CUDA(cudaLaunchKernelEx(&cfg, flash_attn_kernel<true, 8, 2, 2, 6, true>, mq, mk, mv, mw, p));
  } else if (d == 384) {
    flash_attn_kernel<false, 6, 3, 1, 5><<<grid, FA_THREADS, smem, st>>>(mq, mk, mv, mw, p);
  } else if (spill) {
    flash_attn_kernel<false, 8, 2, 2, 3, true><<<grid, FA_THREADS, smem, st>>>(mq, mk, mv, mw, p);
  } else {
    flash_attn_kernel<false, 8, 2, 2, 3><<<grid, FA_THREADS, smem, st>>>(mq, mk, mv, mw, p);
  }
  RV_LAUNCH_CHECK();
  return 0;
}
