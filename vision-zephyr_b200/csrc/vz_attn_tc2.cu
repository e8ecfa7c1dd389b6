// vz_attn_tc2.cu -- CLIP ViT self-attention (577 tokens, 16 heads x 64), second tcgen05 form: the KEYS of every
// 64-key block are split between two softmax threads per query row, each with its OWN running maximum, row sum and
// output accumulator -- a split-KV softmax inside the CTA.  The first form (vz_attn_tc.cu) runs one softmax thread per
// row, i.e. two softmax warps per scheduler walking a serial ld -> max -> 64 x ex2 -> st chain per block (ncu: XU 56 %,
// issue 48 %); an earlier attempt at two threads per row exchanged block maxima through shared memory and lost.
// Here nothing is exchanged per block: thread A (warps 0-3) owns keys 0..31 of every block and accumulator O_a,
// thread B (warps 4-7) keys 32..63 and O_b; P_a V[0:32] and P_b V[32:64] are separate MMAs; the two partial results
// meet once per work item in the epilogue: O = (O_a 2^(m_a - m) + O_b 2^(m_b - m)) / (l_a 2^(m_a - m) + l_b 2^(m_b - m)).
// Tensor memory (256 columns per CTA, two CTAs per SM): S0 | S1 | O_a | O_b, with P written OVER S (the trick of
// vz_attn_causal.cu; Q K_{j+2}^T is issued behind P_j V_j, the tensor pipe executes in issue order).
// Everything else follows vz_attn_tc.cu: persistent CTAs over (tile, head, 128-query block) items, TMA from the packed
// qkv activation, K and V rings, lazy accumulator rescale, deferred epilogue, TMA stores clipped at the tile's 577 rows.
// The 577th key: the last block is computed as 16 keys with one valid one -- thread A's; thread B sits that block out.
#include "vz_common.cuh"

namespace vz {
namespace {

constexpr int TOK = VZ_VIT_TOKENS;          // 577
constexpr int HD = 64;
constexpr int BQ = 128, BKV = 64, HALF = 32;
constexpr int NKB = (TOK + BKV - 1) / BKV;  // 10
constexpr int LAST_VALID = TOK - (NKB - 1) * BKV;  // 1
constexpr int LAST_N = 16;
static_assert(LAST_VALID >= 1 && LAST_VALID <= LAST_N && LAST_N <= HALF, "last key block belongs to thread A");
constexpr int Q_BYTES = BQ * 128;
constexpr int KV_BYTES = BKV * 128;
constexpr int K_STAGES = 3, V_STAGES = 4;
constexpr int NQB = (TOK + BQ - 1) / BQ;    // 5
constexpr int OUT_SLAB = 32 * 128;          // per lane quarter: 32 output rows x 64 bf16
constexpr int SMEM_Q = 0;                                     // 2 buffers
constexpr int SMEM_K = 2 * Q_BYTES;
constexpr int SMEM_V = SMEM_K + K_STAGES * KV_BYTES;
constexpr int SMEM_OUT = SMEM_V + V_STAGES * KV_BYTES;        // 4 x OUT_SLAB
constexpr int SMEM_EXCH = SMEM_OUT + 4 * OUT_SLAB;            // float2 [2][128]: (m, l) of the two halves of every row
constexpr int SMEM_BARS = SMEM_EXCH + 2 * 128 * 8;
constexpr int SMEM_TOTAL = SMEM_BARS + 256;
static_assert(2 * (SMEM_TOTAL + 1024) <= 228 * 1024, "two CTAs per SM");
constexpr int SOFTMAX_WARPS = 8;
constexpr int THREADS = 32 * (SOFTMAX_WARPS + 2);   // + MMA warp + TMA warp
constexpr uint32_t TMEM_COLS = 256;         // S0: 0..63, S1: 64..127 (P_a over cols 0..15, P_b over 32..47), O_a: 128..191, O_b: 192..255
constexpr uint32_t TMEM_O = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleLog2 = 8.0f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// the two warps of a lane quarter (64 threads) meet; ids 1..4 (0 is __syncthreads)
__device__ __forceinline__ void pair_sync(int q) { asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory"); }

struct SoftmaxState {
  float m_used = -INFINITY;
  float l = 0.f;
  float sl2;
};

// One key block of this thread's half (NCOL score columns starting at column `col0` of S[g & 1], the first NVALID
// real keys); g = the CTA's running key-block counter, j = block index inside the item.
template <int NCOL, int NVALID>
__device__ __forceinline__ void softmax_half(SoftmaxState& s, uint32_t g, int j, uint32_t t_lane, uint32_t tmem_base,
                                             uint32_t tmem_ox, uint32_t col0, uint64_t* bar_s_full, uint64_t* bar_p_full,
                                             uint64_t* bar_pv_done) {
  static_assert(NCOL == 32 || NCOL == 16, "columns per half block");
  const int lane = threadIdx.x & 31;
  const uint32_t b = g & 1, use = g >> 1;
  mbar_wait(&bar_s_full[b], use & 1, 600 + b);
  tc_fence_after();
  uint32_t v[NCOL];
  if constexpr (NCOL == 32) tmem_ld_32x32b_x32(tmem_base + t_lane + b * BKV + col0, v);
  else tmem_ld_32x32b_x16(tmem_base + t_lane + b * BKV + col0, v);
  tmem_ld_wait();
  constexpr int NCH = NVALID >= 4 ? 4 : NVALID;
  float bm4[NCH];
#pragma unroll
  for (int u = 0; u < NCH; ++u) bm4[u] = -INFINITY;
#pragma unroll
  for (int i = 0; i < NVALID; ++i) bm4[i % NCH] = fmaxf(bm4[i % NCH], __uint_as_float(v[i]));
  float bm = bm4[0];
#pragma unroll
  for (int u = 1; u < NCH; ++u) bm = fmaxf(bm, bm4[u]);
  float alpha = 1.f;
  const bool need = (bm - s.m_used) * s.sl2 > kRescaleLog2;   // true on the first block (m_used = -inf)
  if (need) {
    alpha = ex2_approx((s.m_used - bm) * s.sl2);
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
    const float x0 = fmaf(__uint_as_float(v[i]), s.sl2, -m_sl2), x1 = fmaf(__uint_as_float(v[i + 1]), s.sl2, -m_sl2);
    const float p0 = i < NVALID ? ex2_approx(x0) : 0.f;
    const float p1 = i + 1 < NVALID ? ex2_approx(x1) : 0.f;
    ls4[(i >> 1) % NCH] += p0 + p1;
    pk[i >> 1] = pack_bf16x2(p0, p1);
  }
  float ls = ls4[0];
#pragma unroll
  for (int u = 1; u < NCH; ++u) ls += ls4[u];
  s.l += ls;
  if (any_need) {
    // every earlier P V must have retired before this half's accumulator is touched (MMAs retire in order)
    mbar_wait(&bar_pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1, 620);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_ox + t_lane + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
      tmem_st_32x32b_x32(tmem_ox + t_lane + c * 32, o);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  // this half's P goes over the first half of the score columns it has just read
  if constexpr (NCOL == 32) tmem_st_32x32b_x16(tmem_base + t_lane + b * BKV + col0, pk);
  else tmem_st_32x32b_x8(tmem_base + t_lane + b * BKV + col0, pk);
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&bar_p_full[b]);
}

__device__ __forceinline__ void item_coords(int item, int& qb, int& h, int& t) {
  qb = item % NQB;
  const int th = item / NQB;
  h = th % VZ_VIT_HEADS;
  t = th / VZ_VIT_HEADS;
}

__global__ void __launch_bounds__(THREADS, 2)
vit_attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmO, float scale, int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem + SMEM_Q;
  uint8_t* sK = smem + SMEM_K;
  uint8_t* sV = smem + SMEM_V;
  float2* exch = reinterpret_cast<float2*>(smem + SMEM_EXCH);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BARS);
  uint64_t* bar_q_full = bars;         // [2]
  uint64_t* bar_q_empty = bars + 2;    // [2]
  uint64_t* bar_k_full = bars + 4;                    // [K_STAGES]
  uint64_t* bar_k_empty = bar_k_full + K_STAGES;
  uint64_t* bar_v_full = bar_k_empty + K_STAGES;      // [V_STAGES]
  uint64_t* bar_v_empty = bar_v_full + V_STAGES;
  uint64_t* bar_s_full = bar_v_empty + V_STAGES;      // [2]
  uint64_t* bar_p_full = bar_s_full + 2;              // [2] 8 warp arrivals
  uint64_t* bar_pv_done = bar_p_full + 2;             // [2]
  uint64_t* bar_o_full = bar_pv_done + 2;
  uint64_t* bar_o_free = bar_o_full + 1;              // 8 warp arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o_free + 1);
  static_assert((4 + 2 * K_STAGES + 2 * V_STAGES + 8) * 8 + 4 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_q_full[i], 1); mbar_init(&bar_q_empty[i], 1); }
    for (int i = 0; i < K_STAGES; ++i) { mbar_init(&bar_k_full[i], 1); mbar_init(&bar_k_empty[i], 1); }
    for (int i = 0; i < V_STAGES; ++i) { mbar_init(&bar_v_full[i], 1); mbar_init(&bar_v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s_full[i], 1);
      mbar_init(&bar_p_full[i], SOFTMAX_WARPS);
      mbar_init(&bar_pv_done[i], 1);
    }
    mbar_init(bar_o_full, 1);
    mbar_init(bar_o_free, SOFTMAX_WARPS);
    fence_barrier_init();
  }
  if (warp == SOFTMAX_WARPS) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + TMEM_O;
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t total = (uint32_t)my_items * NKB;

  if (warp == SOFTMAX_WARPS + 1) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
    }
    uint32_t kg = 0, kj = 0, kst = 0, kn = 0; int kitem = blockIdx.x;
    uint32_t vg = 0, vj = 0, vst = 0;         int vitem = blockIdx.x;
    while (kg < total || vg < total) {
      if (kg < total) {
        bool ok = mbar_try_wait(&bar_k_empty[kst], ((kg / K_STAGES) & 1) ^ 1);
        const uint32_t qs = kn & 1;
        if (ok && kj == 0) ok = mbar_try_wait(&bar_q_empty[qs], ((kn >> 1) & 1) ^ 1);   // polled: never blocks the V stream
        if (__shfl_sync(0xffffffffu, (int)ok, 0)) {
          int qb, h, t;
          item_coords(kitem, qb, h, t);
          if (elect_one()) {
            if (kj == 0) {
              mbar_arrive_expect_tx(&bar_q_full[qs], Q_BYTES);
              tma_load_2d(&tmQ, &bar_q_full[qs], sQ + qs * Q_BYTES, h * HD, t * TOK + qb * BQ);
            }
            mbar_arrive_expect_tx(&bar_k_full[kst], KV_BYTES);
            tma_load_2d(&tmKV, &bar_k_full[kst], sK + kst * KV_BYTES, VZ_VIT_WIDTH + h * HD, t * TOK + kj * BKV);
          }
          __syncwarp();
          ++kg;
          if (++kst == K_STAGES) kst = 0;
          if (++kj == NKB) { kj = 0; ++kn; kitem += gridDim.x; }
        }
      }
      if (vg < total && __shfl_sync(0xffffffffu, (int)mbar_try_wait(&bar_v_empty[vst], ((vg / V_STAGES) & 1) ^ 1), 0)) {
        int qb, h, t;
        item_coords(vitem, qb, h, t);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar_v_full[vst], KV_BYTES);
          tma_load_2d(&tmKV, &bar_v_full[vst], sV + vst * KV_BYTES, 2 * VZ_VIT_WIDTH + h * HD, t * TOK + vj * BKV);
        }
        __syncwarp();
        ++vg;
        if (++vst == V_STAGES) vst = 0;
        if (++vj == NKB) { vj = 0; vitem += gridDim.x; }
      }
    }
  } else if (warp == SOFTMAX_WARPS) {
    // ======================= MMA issuer =======================
    constexpr uint32_t idesc_qk = umma_idesc_bf16_ex(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_qk_last = umma_idesc_bf16_ex(BQ, LAST_N, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16_ex(BQ, HD, 0, 1);   // B = V is MN-major
    const uint64_t q_desc0 = umma_smem_desc_sw128(smem_u32(sQ));
    const uint64_t k_desc0 = umma_smem_desc_sw128(smem_u32(sK));
    const uint64_t v_desc0 = umma_smem_desc_sw128(smem_u32(sV));
    // S[g & 1] = Q K_j^T for the CTA's g-th key block (item n, block j; K ring slot kst)
    auto issue_qk = [&](uint32_t g2, uint32_t n2, uint32_t j2) {
      const uint32_t b = g2 & 1, qs = n2 & 1, kst = g2 % K_STAGES;
      if (j2 == 0) mbar_wait(&bar_q_full[qs], (n2 >> 1) & 1, 510 + qs);
      mbar_wait(&bar_k_full[kst], (g2 / K_STAGES) & 1, 520 + kst);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t q_desc = q_desc0 + (uint64_t)(qs * (Q_BYTES >> 4));
        const uint64_t k_desc = k_desc0 + (uint64_t)(kst * (KV_BYTES >> 4));
        const uint32_t idesc = j2 == NKB - 1 ? idesc_qk_last : idesc_qk;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem_base + b * BKV, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc, k != 0 ? 1u : 0u);
        umma_commit(&bar_s_full[b]);
        umma_commit(&bar_k_empty[kst]);
        if (j2 == NKB - 1) umma_commit(&bar_q_empty[qs]);
      }
      __syncwarp();
    };
    uint32_t n = 0, j = 0, vst = 0;
    uint32_t n2 = 0, j2 = 0;
    for (uint32_t a = 0; a < 2 && a < total; ++a) {
      issue_qk(a, n2, j2);
      if (++j2 == NKB) { j2 = 0; ++n2; }
    }
    for (uint32_t g = 0; g < total; ++g) {
      const uint32_t b = g & 1, use = g >> 1;
      mbar_wait(&bar_v_full[vst], (g / V_STAGES) & 1, 535 + vst);
      mbar_wait(&bar_p_full[b], use & 1, 540 + b);
      if (j == 0 && n > 0) mbar_wait(bar_o_free, (n - 1) & 1, 550);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t p_tmem = tmem_base + b * BKV;
        const uint64_t v_desc = v_desc0 + (uint64_t)(vst * (KV_BYTES >> 4));
        const bool last = j == NKB - 1;
        // keys 0..31 of the block -> O_a (the last block only holds LAST_N keys, all of them thread A's)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          if (last && kk * 16 >= LAST_N) break;
          umma_bf16_ts(tmem_o, p_tmem + kk * 8, v_desc + (uint64_t)(kk * (2048 >> 4)), idesc_pv, (j > 0 || kk != 0) ? 1u : 0u);
        }
        // keys 32..63 -> O_b
        if (!last) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            umma_bf16_ts(tmem_o + HD, p_tmem + HALF + kk * 8, v_desc + (uint64_t)((2 + kk) * (2048 >> 4)), idesc_pv,
                         (j > 0 || kk != 0) ? 1u : 0u);
        }
        umma_commit(&bar_pv_done[b]);
        umma_commit(&bar_v_empty[vst]);
        if (last) umma_commit(bar_o_full);
      }
      __syncwarp();
      // the buffer's next tenant, behind P V in issue order
      if (g + 2 < total) {
        issue_qk(g + 2, n2, j2);
        if (++j2 == NKB) { j2 = 0; ++n2; }
      }
      if (++j == NKB) { j = 0; ++n; }
      if (++vst == V_STAGES) vst = 0;
    }
  } else {
    // ======================= softmax warps: two threads per query row =======================
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    const uint32_t tmem_ox = tmem_o + hf * HD;             // this half's accumulator
    const uint32_t col0 = hf * HALF;                       // ... and its score columns inside S[b]
    uint8_t* slab = smem + SMEM_OUT + q * OUT_SLAB;
    // epilogue of item (eq, eh, et) = the CTA's en-th: the halves exchange (m, l), each combines 32 output columns
    auto epilogue = [&](int eq, int eh, int et, uint32_t en, float m, float l, bool active) {
      exch[hf * 128 + r] = make_float2(m, l);
      if (hf == 0 && lane == 0) tma_store_wait_read();     // the previous store has drained this quarter's slab
      pair_sync(q);
      const float2 other = exch[(hf ^ 1) * 128 + r];
      mbar_wait(bar_o_full, en & 1, 640);
      tc_fence_after();
      uint32_t oa[32], ob[32];
      tmem_ld_32x32b_x32(tmem_o + t_lane + hf * 32, oa);          // columns hf*32 .. +31 of O_a
      tmem_ld_32x32b_x32(tmem_o + HD + t_lane + hf * 32, ob);     // ... and of O_b
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_o_free);
      if (active) {
        const float sl2 = scale * kLog2e;
        const float m_a = hf == 0 ? m : other.x, l_a = hf == 0 ? l : other.y;
        const float m_b = hf == 0 ? other.x : m, l_b = hf == 0 ? other.y : l;
        const float mm = fmaxf(m_a, m_b);
        const float wa = ex2_approx((m_a - mm) * sl2), wb = ex2_approx((m_b - mm) * sl2);
        const float inv = 1.0f / (l_a * wa + l_b * wb);
        const float ca = wa * inv, cb = wb * inv;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float f[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(oa[8 * i + u]) * ca + __uint_as_float(ob[8 * i + u]) * cb;
          uint4 w;
          w.x = pack_bf16x2(f[0], f[1]); w.y = pack_bf16x2(f[2], f[3]);
          w.z = pack_bf16x2(f[4], f[5]); w.w = pack_bf16x2(f[6], f[7]);
          // 16-byte chunk (hf * 4 + i) of row `lane`, 128-byte swizzle
          *reinterpret_cast<uint4*>(slab + lane * 128 + (((hf * 4 + i) ^ (lane & 7)) << 4)) = w;
        }
        fence_proxy_async_smem();
      }
      pair_sync(q);
      if (active && hf == 0 && lane == 0) {
        tma_store_3d(&tmO, slab, eh * HD, eq * BQ + q * 32, et);
        tma_store_commit();
      }
    };
    bool pend = false, pact = false;
    int pq = 0, ph = 0, pt = 0;
    float pm = 0.f, pl = 1.f;
    uint32_t g = 0, n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      int qb, h, t;
      item_coords(item, qb, h, t);
      const bool active = qb * BQ + q * 32 < TOK;
      SoftmaxState stt;
      stt.sl2 = scale * kLog2e;
      auto idle_block = [&]() {
        const uint32_t b = g & 1, use = g >> 1;
        mbar_wait(&bar_s_full[b], use & 1, 600 + b);
        if (lane == 0) mbar_arrive(&bar_p_full[b]);
        __syncwarp();
      };
      for (int j = 0; j < NKB; ++j, ++g) {
        if (!active) idle_block();
        else if (j < NKB - 1) softmax_half<HALF, HALF>(stt, g, j, t_lane, tmem_base, tmem_ox, col0, bar_s_full, bar_p_full, bar_pv_done);
        else if (hf == 0) softmax_half<LAST_N, LAST_VALID>(stt, g, j, t_lane, tmem_base, tmem_ox, col0, bar_s_full, bar_p_full, bar_pv_done);
        else idle_block();    // thread B has no key in the last block
        if (j == 0 && pend) { epilogue(pq, ph, pt, n - 1, pm, pl, pact); pend = false; }
      }
      pend = true; pact = active; pq = qb; ph = h; pt = t; pm = stt.m_used; pl = stt.l;
    }
    if (pend) epilogue(pq, ph, pt, n - 1, pm, pl, pact);
    if (hf == 0 && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SOFTMAX_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

int vit_attn_tc2_launch(const void* qkv, void* out, int T, cudaStream_t st) {
  CUtensorMap tmQ, tmKV, tmO;
  VZ_TRY(encode_tmap_2d_bf16(&tmQ, qkv, (long long)T * TOK, 3 * VZ_VIT_WIDTH, 3 * VZ_VIT_WIDTH, HD, BQ));
  VZ_TRY(encode_tmap_2d_bf16(&tmKV, qkv, (long long)T * TOK, 3 * VZ_VIT_WIDTH, 3 * VZ_VIT_WIDTH, HD, BKV));
  VZ_TRY(encode_tmap_3d_bf16(&tmO, out, TOK, VZ_VIT_WIDTH, VZ_VIT_WIDTH, 32, T, (long long)TOK * VZ_VIT_WIDTH));
  int dev = 0, num_sms = 0;
  VZ_CUDA_CHECK(cudaGetDevice(&dev));
  VZ_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  const int n_items = NQB * VZ_VIT_HEADS * T;
  const int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;
  ProfScope prof(VZ_PROF_VIT_ATTN, 4.0 * TOK * TOK * HD * VZ_VIT_HEADS * T, st);
  VZ_ENSURE_DYN_SMEM(vit_attn_tc2_kernel, SMEM_TOTAL);
  vit_attn_tc2_kernel<<<grid, THREADS, SMEM_TOTAL, st>>>(tmQ, tmKV, tmO, 0.125f, n_items);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

}  // namespace vz
