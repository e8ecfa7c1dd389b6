// vz_attn_tc.cu -- CLIP ViT self-attention (577 tokens, 16 heads x 64, non-causal) on the 5th-gen
// tensor cores: S = Q K^T and O = P V are tcgen05.mma instructions with both accumulators in TMEM;
// Q / K / V tiles arrive by TMA (128-byte swizzle) straight from the packed qkv activation.
// Replaces HF CLIPAttention as called from vision_encoder/vision_encoder.py:101-105.
//
// PERSISTENT kernel, two CTAs per SM; a work item = 128 query rows of one (tile, head), items are
// walked round-robin.  Warps 0-3: softmax (thread = query row = TMEM lane); warp 4: one thread
// issues the MMAs; warp 5: one thread issues TMA.  The TMA and MMA threads run ahead ACROSS items
// (Q is double buffered, the K/V ring and the S buffers never drain), so the load latency of an
// item and the epilogue of the previous one overlap with softmax work instead of idling the SM.
// The 577 keys are walked in 10 blocks of 64 -- the last one holds a single valid key and is computed
// as a 16-key block -- with EVERYTHING double buffered -- two S accumulators and two P tiles in
// TMEM, a 4-stage K/V ring -- so Q K_{j+1}^T is issued before the softmax of block j starts and
// P_j V_j runs while the softmax of block j+1 computes:
//   S_j = Q K_j^T -> registers -> P_j = exp2((S_j - m) * scale) as bf16 back into TMEM -> O += P_j V_j
// (P is the TMEM A operand of the second MMA: it never touches shared memory)
// The online softmax rescales the accumulator LAZILY: O (in TMEM) is only multiplied by
// exp2(m_old - m_new) when the running maximum grew by more than 2^8, which is rare after the first
// block, so O normally stays untouched in TMEM until the epilogue divides by the row sum.  Softmax
// warps without a valid query row (rows >= 577 of the last query block) only keep the barriers going.
#include "vz_common.cuh"

#include <stdlib.h>

namespace vz {
namespace {

constexpr int VZ_ATTN_POLY_DEFAULT = 0;
constexpr int TOK = VZ_VIT_TOKENS;          // 577
constexpr int HD = 64;                      // head dim
constexpr int BQ = 128, BKV = 64;           // query rows per CTA, keys per block
constexpr int NKB = (TOK + BKV - 1) / BKV;  // 10 key blocks (the last one holds a single valid key)
constexpr int LAST_VALID = TOK - (NKB - 1) * BKV;  // 1 valid key in the last block ...
constexpr int LAST_N = 16;                          // ... which is computed as a 16-key block
static_assert(LAST_VALID >= 1 && LAST_VALID <= LAST_N, "last key block");
constexpr int Q_BYTES = BQ * 128;           // 128 rows x 64 bf16
constexpr int KV_BYTES = BKV * 128;         // 64 rows x 64 bf16
constexpr int KV_STAGES = 4;                // per ring: K and V have their own rings (K is consumed ~2 blocks before V)
constexpr int NQB = (TOK + BQ - 1) / BQ;    // 5 query blocks per (tile, head)
constexpr int OUT_SLAB = 32 * 128;          // per softmax warp: 32 output rows x 64 bf16, staged for the TMA store
constexpr int SMEM_Q = 0;                                     // 2 buffers (the next item's Q is prefetched)
constexpr int SMEM_K = 2 * Q_BYTES;                           // KV_STAGES x K
constexpr int SMEM_V = SMEM_K + KV_STAGES * KV_BYTES;         // KV_STAGES x V
constexpr int SMEM_OUT = SMEM_V + KV_STAGES * KV_BYTES;       // 4 x OUT_SLAB
constexpr int SMEM_BARS = SMEM_OUT + 4 * OUT_SLAB;
constexpr int SMEM_TOTAL = SMEM_BARS + 256;
static_assert(2 * (SMEM_TOTAL + 1024) <= 228 * 1024, "attention kernel must keep 2 CTAs per SM");
constexpr int THREADS = 192;                // 4 softmax warps + MMA warp + TMA warp
constexpr uint32_t TMEM_COLS = 256;         // S0: 0..63, S1: 64..127, O: 128..191, P0: 192..223, P1: 224..255
constexpr uint32_t TMEM_O = 128, TMEM_P = 192, P_COLS = BKV / 2;   // P: bf16 pairs, 32 columns per buffer
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for x <= ~8 on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f, |f| <= 0.5, cubic
// for 2^f (relative error < 1.5e-4, far below the bf16 rounding P gets anyway), exponent added as integer.
// Every PE-th exponential of a key block takes this route so the 16-lane MUFU pipe and the FMA
// pipe share the softmax (FA-4's trick); x is clamped at -125 so the result never leaves the normal range.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;          // 1.5 * 2^23: the integer part of x lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// (template parameter PE of the kernel: every PE-th exponential; 0 = all exponentials on the MUFU pipe)

__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// exp2 may run up to 2^8 above the value it would have with the exact running maximum before the
// accumulator is rescaled (FA-4 style lazy rescaling): keeps O untouched in TMEM most of the time.
constexpr float kRescaleLog2 = 8.0f;


__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

struct SoftmaxState {
  float m_used = -INFINITY;   // maximum the exponentials are currently taken against
  float l = 0.f;              // running row sum
  float sl2;                  // softmax scale * log2(e)
};

// One key block of the online softmax for query row r: NCOL score columns are read from TMEM, the
// first NVALID are real keys (the rest of the tile's last block belongs to the next tile).
// g = this CTA's running key-block counter (across work items): buffer b = g & 1, use = g >> 1;
// j = block index inside the item (the accumulator is only touched when j > 0).
template <int NCOL, int NVALID, int PE>
__device__ __forceinline__ void softmax_block(SoftmaxState& s, uint32_t g, int j, int r, uint32_t t_lane,
                                              uint32_t tmem_base, uint32_t tmem_o, uint64_t* bar_s_full,
                                              uint64_t* bar_s_free, uint64_t* bar_p_full, uint64_t* bar_pv_done) {
  static_assert(NCOL == 64 || NCOL == 16, "score columns per block");
  const int lane = threadIdx.x & 31;
  const uint32_t b = g & 1, use = g >> 1;
  mbar_wait(&bar_s_full[b], use & 1, 600 + b);
  tc_fence_after();
  uint32_t v[NCOL];
  if constexpr (NCOL == 64) {
#pragma unroll
    for (int c = 0; c < 2; ++c)
      tmem_ld_32x32b_x32(tmem_base + t_lane + b * BKV + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
  } else {
    tmem_ld_32x32b_x16(tmem_base + t_lane + b * BKV, v);
  }
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&bar_s_free[b]);
  constexpr int NCH = NVALID >= 4 ? 4 : NVALID;   // independent max / sum chains
  float bm4[NCH];
#pragma unroll
  for (int u = 0; u < NCH; ++u) bm4[u] = -INFINITY;
#pragma unroll
  for (int i = 0; i < NVALID; ++i) bm4[i % NCH] = fmaxf(bm4[i % NCH], __uint_as_float(v[i]));
  float bm = bm4[0];
#pragma unroll
  for (int u = 1; u < NCH; ++u) bm = fmaxf(bm, bm4[u]);
  // lazy rescale: only when the running maximum grows by more than 2^kRescaleLog2
  float alpha = 1.f;
  const bool need = (bm - s.m_used) * s.sl2 > kRescaleLog2;   // true on the first block (m_used = -inf)
  if (need) {
    alpha = ex2_approx((s.m_used - bm) * s.sl2);               // 0 on the first block
    s.m_used = bm;
    s.l *= alpha;
  }
  const bool any_need = __any_sync(0xffffffffu, need) && j > 0;
  const float m_sl2 = s.m_used * s.sl2;
  uint32_t pk[NCOL / 2];
  float ls4[NCH];
#pragma unroll
  for (int u = 0; u < NCH; ++u) ls4[u] = 0.f;
#pragma unroll
  for (int i = 0; i < NCOL; i += 2) {
    // keys beyond NVALID are masked: probability 0 without spending an exponential on them
    const float x0 = fmaf(__uint_as_float(v[i]), s.sl2, -m_sl2), x1 = fmaf(__uint_as_float(v[i + 1]), s.sl2, -m_sl2);
    constexpr int PE1 = PE > 0 ? PE : 1;
    const bool poly0 = PE > 0 && NCOL == 64 && (i % PE1) == PE1 - 1;      // compile-time after unrolling
    const bool poly1 = PE > 0 && NCOL == 64 && ((i + 1) % PE1) == PE1 - 1;
    const float p0 = i < NVALID ? (poly0 ? ex2_poly(x0) : ex2_approx(x0)) : 0.f;
    const float p1 = i + 1 < NVALID ? (poly1 ? ex2_poly(x1) : ex2_approx(x1)) : 0.f;
    ls4[(i >> 1) % NCH] += p0 + p1;
    pk[i >> 1] = pack_bf16x2(p0, p1);
  }
  float ls = ls4[0];
#pragma unroll
  for (int u = 1; u < NCH; ++u) ls += ls4[u];
  s.l += ls;
  if (any_need) {
    // every earlier P V must have retired before O is touched (MMAs retire in order)
    mbar_wait(&bar_pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1, 620);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_o + t_lane + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
      tmem_st_32x32b_x32(tmem_o + t_lane + c * 32, o);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  if (use > 0) mbar_wait(&bar_pv_done[b], (use - 1) & 1, 630 + b);   // P buffer b: previous tenant consumed
  // P[r][0..NCOL) stays in tensor memory: the A operand of P V (lane = row, two bf16 per 32-bit column)
  if constexpr (NCOL == 64) tmem_st_32x32b_x32(tmem_base + t_lane + TMEM_P + b * P_COLS, pk);
  else tmem_st_32x32b_x8(tmem_base + t_lane + TMEM_P + b * P_COLS, pk);
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&bar_p_full[b]);
}

// work item -> (query block, head, tile); consecutive items share K/V (L2 locality between co-resident CTAs)
__device__ __forceinline__ void item_coords(int item, int& qb, int& h, int& t) {
  qb = item % NQB;
  const int th = item / NQB;
  h = th % VZ_VIT_HEADS;
  t = th / VZ_VIT_HEADS;
}

template <int PE>
__global__ void __launch_bounds__(THREADS, 2)
vit_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmO, float scale, int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem + SMEM_Q;
  uint8_t* sK = smem + SMEM_K;
  uint8_t* sV = smem + SMEM_V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BARS);
  uint64_t* bar_q_full = bars;         // [2] Q of an item landed
  uint64_t* bar_q_empty = bars + 2;    // [2] every Q K^T of the item retired -> Q buffer reusable
  uint64_t* bar_k_full = bars + 4;                    // [KV_STAGES] K_j landed
  uint64_t* bar_k_empty = bar_k_full + KV_STAGES;     // [KV_STAGES] Q K_j^T retired -> slot reusable
  uint64_t* bar_v_full = bar_k_empty + KV_STAGES;     // [KV_STAGES] V_j landed
  uint64_t* bar_v_empty = bar_v_full + KV_STAGES;     // [KV_STAGES] P V_j retired -> slot reusable
  uint64_t* bar_s_full = bar_v_empty + KV_STAGES;     // [2] S buffer written by Q K^T
  uint64_t* bar_s_free = bar_s_full + 2;    // [2] S buffer copied to registers (4 warp arrivals)
  uint64_t* bar_p_full = bar_s_free + 2;    // [2] P buffer written (and O rescaled if needed) (4 warp arrivals)
  uint64_t* bar_pv_done = bar_p_full + 2;   // [2] P V retired -> P buffer reusable, O up to date
  uint64_t* bar_o_full = bar_pv_done + 2;   // every MMA of the item retired
  uint64_t* bar_o_free = bar_o_full + 1;    // O copied to registers (4 warp arrivals) -> next item may overwrite it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o_free + 1);
  static_assert((4 + 4 * KV_STAGES + 10) * 8 + 4 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the swizzled layouts need a 1024-byte aligned base
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_q_full[i], 1); mbar_init(&bar_q_empty[i], 1); }
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&bar_k_full[i], 1); mbar_init(&bar_k_empty[i], 1);
      mbar_init(&bar_v_full[i], 1); mbar_init(&bar_v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s_full[i], 1);
      mbar_init(&bar_s_free[i], 4);
      mbar_init(&bar_p_full[i], 4);
      mbar_init(&bar_pv_done[i], 1);
    }
    mbar_init(bar_o_full, 1);
    mbar_init(bar_o_free, 4);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + TMEM_O;
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t total = (uint32_t)my_items * NKB;   // key blocks this CTA walks, over all its items

  if (warp == 5) {
    // ======================= TMA producer: runs ahead across work items =======================
    // (warp-uniform like the MMA warp: every lane walks the loop, one elected lane issues.)
    // K and V travel through separate rings: a K slot is free again as soon as its Q K^T retired, long
    // before the V of the same block is consumed, so both streams keep their full prefetch distance.
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
    }
    uint32_t kg = 0, kj = 0, kst = 0, kn = 0; int kitem = blockIdx.x;   // next K block: index, block in item, slot, item
    uint32_t vg = 0, vj = 0, vst = 0;         int vitem = blockIdx.x;   // next V block
    while (kg < total || vg < total) {
      if (kg < total && __shfl_sync(0xffffffffu, (int)mbar_try_wait(&bar_k_empty[kst], ((kg / KV_STAGES) & 1) ^ 1), 0)) {
        int qb, h, t;
        item_coords(kitem, qb, h, t);
        if (kj == 0) {
          const uint32_t qs = kn & 1;
          mbar_wait(&bar_q_empty[qs], ((kn >> 1) & 1) ^ 1, 490 + qs);
          if (elect_one()) {
            mbar_arrive_expect_tx(&bar_q_full[qs], Q_BYTES);
            tma_load_2d(&tmQ, &bar_q_full[qs], sQ + qs * Q_BYTES, h * HD, t * TOK + qb * BQ);
          }
        }
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar_k_full[kst], KV_BYTES);
          tma_load_2d(&tmKV, &bar_k_full[kst], sK + kst * KV_BYTES, VZ_VIT_WIDTH + h * HD, t * TOK + kj * BKV);
        }
        __syncwarp();
        ++kg;
        if (++kst == KV_STAGES) kst = 0;
        if (++kj == NKB) { kj = 0; ++kn; kitem += gridDim.x; }
      }
      if (vg < total && __shfl_sync(0xffffffffu, (int)mbar_try_wait(&bar_v_empty[vst], ((vg / KV_STAGES) & 1) ^ 1), 0)) {
        int qb, h, t;
        item_coords(vitem, qb, h, t);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar_v_full[vst], KV_BYTES);
          tma_load_2d(&tmKV, &bar_v_full[vst], sV + vst * KV_BYTES, 2 * VZ_VIT_WIDTH + h * HD, t * TOK + vj * BKV);
        }
        __syncwarp();
        ++vg;
        if (++vst == KV_STAGES) vst = 0;
        if (++vj == NKB) { vj = 0; vitem += gridDim.x; }
      }
    }
  } else if (warp == 4) {
    // ======================= MMA issuer =======================
    // All 32 lanes walk the loop and the waits (warp-uniform control flow keeps the descriptors in
    // uniform registers and the instruction stream short -- the MMAs of one key block are only ~260
    // tensor-pipe cycles, so a slow issuing thread would be the critical path); one elected lane issues.
    constexpr uint32_t idesc_qk = umma_idesc_bf16_ex(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_qk_last = umma_idesc_bf16_ex(BQ, LAST_N, 0, 0);   // last block: 1 valid key
    constexpr uint32_t idesc_pv = umma_idesc_bf16_ex(BQ, HD, 0, 1);   // B = V is MN-major (dims contiguous)
    const uint64_t q_desc0 = umma_smem_desc_sw128(smem_u32(sQ));
    const uint64_t k_desc0 = umma_smem_desc_sw128(smem_u32(sK));
    const uint64_t v_desc0 = umma_smem_desc_sw128(smem_u32(sV));
    // S[g & 1] = Q K_j^T for the CTA's g-th key block (item n, block j of it; ring slot st)
    auto issue_qk = [&](uint32_t g, uint32_t n, uint32_t j, uint32_t st) {
      const uint32_t b = g & 1, use = g >> 1, qs = n & 1;
      if (j == 0) mbar_wait(&bar_q_full[qs], (n >> 1) & 1, 510 + qs);
      mbar_wait(&bar_k_full[st], (g / KV_STAGES) & 1, 520 + st);
      if (use > 0) mbar_wait(&bar_s_free[b], (use - 1) & 1, 530 + b);   // previous tenant is in registers
      tc_fence_after();
      if (elect_one()) {
        const uint64_t q_desc = q_desc0 + (uint64_t)(qs * (Q_BYTES >> 4));
        const uint64_t k_desc = k_desc0 + (uint64_t)(st * (KV_BYTES >> 4));
        const uint32_t idesc = j == NKB - 1 ? idesc_qk_last : idesc_qk;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem_base + b * BKV, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc,
                    k != 0 ? 1u : 0u);
        umma_commit(&bar_s_full[b]);
        umma_commit(&bar_k_empty[st]);
        if (j == NKB - 1) umma_commit(&bar_q_empty[qs]);   // the item's last use of Q
      }
      __syncwarp();
    };
    // Q K^T runs TWO blocks ahead of the softmax: S[g & 1] is rewritten with block g + 2 as soon as the
    // softmax has copied block g to registers, so the issue -> commit -> mbarrier -> waiter latency of a
    // block (~1000 cycles measured) hides behind almost two softmax blocks.  The events the loop waits
    // for alternate strictly (s_free(g), p_full(g), s_free(g + 1), ...), so blocking waits suffice.
    uint32_t n = 0, j = 0, st = 0;          // coordinates of block g
    uint32_t n2 = 0, j2 = 0, st2 = 0;       // ... and of the next block whose Q K^T is to be issued
    for (uint32_t a = 0; a < 2 && a < total; ++a) {
      issue_qk(a, n2, j2, st2);
      if (++j2 == NKB) { j2 = 0; ++n2; }
      if (++st2 == KV_STAGES) st2 = 0;
    }
    for (uint32_t g = 0; g < total; ++g) {
      const uint32_t b = g & 1, use = g >> 1;
      if (g + 2 < total) {
        issue_qk(g + 2, n2, j2, st2);   // waits until the softmax holds S_g in registers (s_free)
        if (++j2 == NKB) { j2 = 0; ++n2; }
        if (++st2 == KV_STAGES) st2 = 0;
      }
      mbar_wait(&bar_v_full[st], (g / KV_STAGES) & 1, 535 + st);
      mbar_wait(&bar_p_full[b], use & 1, 540 + b);
      if (j == 0 && n > 0) mbar_wait(bar_o_free, (n - 1) & 1, 550);   // previous item's O is in registers
      tc_fence_after();
      if (elect_one()) {
        const uint32_t p_tmem = tmem_base + TMEM_P + b * P_COLS;
        const uint64_t v_desc = v_desc0 + (uint64_t)(st * (KV_BYTES >> 4));
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          if (j == NKB - 1 && kk * 16 >= LAST_N) break;   // the last block only holds LAST_VALID keys
          // 16 keys = 2 x (8 rows x 128 B) of V, 8 columns of P
          umma_bf16_ts(tmem_o, p_tmem + kk * 8, v_desc + (uint64_t)(kk * (2048 >> 4)), idesc_pv,
                       (j > 0 || kk != 0) ? 1u : 0u);
        }
        umma_commit(&bar_pv_done[b]);
        umma_commit(&bar_v_empty[st]);
        if (j == NKB - 1) umma_commit(bar_o_full);
      }
      __syncwarp();
      if (++j == NKB) { j = 0; ++n; }
      if (++st == KV_STAGES) st = 0;
    }
  } else {
    // ======================= softmax warps: thread = query row = TMEM lane =======================
    const int r = warp * 32 + lane;
    const uint32_t t_lane = ((uint32_t)(warp * 32)) << 16;
    uint8_t* slab = smem + SMEM_OUT + warp * OUT_SLAB;   // this warp's 32 output rows, 128B-swizzled for the TMA store
    // epilogue of item (eq, eh, et) = the CTA's en-th: O / l -> bf16 -> swizzled smem slab -> TMA store (rows
    // beyond the tile's 577 are clipped by the tensor map)
    auto epilogue = [&](int eq, int eh, int et, uint32_t en, float l, bool active) {
      mbar_wait(bar_o_full, en & 1, 640);
      tc_fence_after();
      uint32_t o[64];
#pragma unroll
      for (int c = 0; c < 2; ++c)
        tmem_ld_32x32b_x32(tmem_o + t_lane + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&o[c * 32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_o_free);   // the next item's first P V may overwrite O now
        tma_store_wait_read();     // the previous store has drained this warp's slab
      }
      __syncwarp();
      if (active) {
        const float inv = 1.0f / l;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(slab + lane * 128 + ((i ^ (lane & 7)) << 4)) = w;
        }
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmO, slab, eh * HD, eq * BQ + warp * 32, et);
          tma_store_commit();
        }
      }
    };
    // The epilogue of an item is DEFERRED until the first key block of the next item has been handed to
    // the MMA warp: by then the item's last P V has long retired, so nobody waits for it (the next item's
    // first P V, the only instruction that needs O to be drained, waits on o_free instead).
    bool pend = false, pact = false;
    int pq = 0, ph = 0, pt = 0;
    float pl = 1.f;
    uint32_t g = 0, n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      int qb, h, t;
      item_coords(item, qb, h, t);
      const bool active = qb * BQ + warp * 32 < TOK;
      SoftmaxState stt;
      stt.sl2 = scale * kLog2e;
      // no valid query row in this warp (last query block): keep the barrier protocol going, skip the math
      auto idle_block = [&]() {
        const uint32_t b = g & 1, use = g >> 1;
        mbar_wait(&bar_s_full[b], use & 1, 600 + b);
        // same gate as the working warps: an arrival for block g must not land in p_full[b]'s previous phase
        if (use > 0) mbar_wait(&bar_pv_done[b], (use - 1) & 1, 610 + b);
        if (lane == 0) { mbar_arrive(&bar_s_free[b]); mbar_arrive(&bar_p_full[b]); }
        __syncwarp();
      };
      if (active) softmax_block<BKV, BKV, PE>(stt, g, 0, r, t_lane, tmem_base, tmem_o, bar_s_full, bar_s_free, bar_p_full,
                                          bar_pv_done);
      else idle_block();
      ++g;
      if (pend) epilogue(pq, ph, pt, n - 1, pl, pact);
      if (active) {
        for (int j = 1; j < NKB - 1; ++j, ++g)
          softmax_block<BKV, BKV, PE>(stt, g, j, r, t_lane, tmem_base, tmem_o, bar_s_full, bar_s_free, bar_p_full,
                                  bar_pv_done);
        softmax_block<LAST_N, LAST_VALID, PE>(stt, g, NKB - 1, r, t_lane, tmem_base, tmem_o, bar_s_full, bar_s_free,
                                          bar_p_full, bar_pv_done);
        ++g;
      } else {
        for (int j = 1; j < NKB; ++j, ++g) idle_block();
      }
      pend = true; pact = active; pq = qb; ph = h; pt = t; pl = stt.l;
    }
    if (pend) epilogue(pq, ph, pt, n - 1, pl, pact);
    if (lane == 0) tma_store_wait_all();    // the last store has landed before this CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

int vit_attn_tc_launch(const void* qkv, void* out, int T, cudaStream_t st) {
  CUtensorMap tmQ, tmKV, tmO;
  VZ_TRY(encode_tmap_2d_bf16(&tmQ, qkv, (long long)T * TOK, 3 * VZ_VIT_WIDTH, 3 * VZ_VIT_WIDTH, HD, BQ));
  VZ_TRY(encode_tmap_2d_bf16(&tmKV, qkv, (long long)T * TOK, 3 * VZ_VIT_WIDTH, 3 * VZ_VIT_WIDTH, HD, BKV));
  // output as [tile][577 rows][1024]: a 32-row store box that runs past a tile's last row is clipped by the TMA
  VZ_TRY(encode_tmap_3d_bf16(&tmO, out, TOK, VZ_VIT_WIDTH, VZ_VIT_WIDTH, 32, T, (long long)TOK * VZ_VIT_WIDTH));
  // share of the exponentials computed on the FMA pipe (VZ_ATTN_POLY = 0 | 4 | 8: none, every 4th, every 8th)
  static const int poly = []() { const char* e = getenv("VZ_ATTN_POLY"); return e ? atoi(e) : VZ_ATTN_POLY_DEFAULT; }();
  int dev = 0, num_sms = 0;
  VZ_CUDA_CHECK(cudaGetDevice(&dev));
  VZ_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  // persistent: two CTAs per SM walk the (tile, head, query block) items round-robin
  const int n_items = NQB * VZ_VIT_HEADS * T;
  const int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;
  // algorithmic FLOPs: QK^T and PV, 2 * 577 * 577 * 64 each, per (tile, head)
  ProfScope prof(VZ_PROF_VIT_ATTN, 4.0 * TOK * TOK * HD * VZ_VIT_HEADS * T, st);
  if (poly == 4) {
    VZ_ENSURE_DYN_SMEM(vit_attn_tc_kernel<4>, SMEM_TOTAL);
    vit_attn_tc_kernel<4><<<grid, THREADS, SMEM_TOTAL, st>>>(tmQ, tmKV, tmO, 0.125f, n_items);
  } else if (poly == 8) {
    VZ_ENSURE_DYN_SMEM(vit_attn_tc_kernel<8>, SMEM_TOTAL);
    vit_attn_tc_kernel<8><<<grid, THREADS, SMEM_TOTAL, st>>>(tmQ, tmKV, tmO, 0.125f, n_items);
  } else {
    VZ_ENSURE_DYN_SMEM(vit_attn_tc_kernel<0>, SMEM_TOTAL);
    vit_attn_tc_kernel<0><<<grid, THREADS, SMEM_TOTAL, st>>>(tmQ, tmKV, tmO, 0.125f, n_items);
  }
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

}  // namespace vz
