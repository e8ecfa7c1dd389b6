// vz_attn_causal.cu -- causal grouped-query self-attention of the LLM prefill (Mistral / Zephyr: 32 query heads,
// 8 key/value heads, head dim 128) on PACKED variable-length rows, on the 5th-gen tensor cores.
// Replaces the attention core HF Mistral runs behind language_model/vis_zephyr.py:86-98 (and the unpadded
// flash-attention call of train/zephyr_flash_attn_monkey_patch.py:86-136): q | k | v rows of all samples back to
// back, sample b = rows [cu[b], cu[b + 1]), query i of a sample attends to keys 0 .. i of the same sample.
//
// Same skeleton as vz_attn_tc.cu (the ViT kernel): PERSISTENT, two CTAs per SM; warps 0-3 softmax (thread = query
// row = TMEM lane), warp 4 issues the MMAs, warp 5 the TMA loads; S = Q K^T and O += P V are tcgen05.mma with S, P
// and O in tensor memory; lazy accumulator rescale; Q K^T issued two key blocks ahead of the softmax.
// What differs:
//   * work item = (128 query rows of one sample, one query head); its key blocks are the 64-key blocks up to the
//     diagonal, so items differ in length: the host sorts them by length and the CTAs take them in snake order
//     (round r: item r G + c, or r G + G - 1 - c on odd rounds) -- a static schedule every role can recompute;
//   * head dim 128 = two 128-byte swizzle atoms per row: Q, K and V tiles are pairs of TMA boxes, the K loop of
//     Q K^T walks eight 16-column steps over both atoms, and P V has N = 128 (V is the MN-major operand, its two
//     64-column atoms 8 KB apart);
//   * tensor memory is the scarce resource (256 columns per CTA for two CTAs per SM): S0 | S1 | O(128), and P_j
//     is written OVER the first half of S_j (the softmax holds S_j in registers by then); the next tenant of the
//     buffer, Q K_{j+2}^T, is issued right behind P_j V_j, and the tensor pipe executes in issue order;
//   * only the last two key blocks of an item straddle the diagonal and pay for masking; rows past the sample's
//     end are never stored;
//   * one Q buffer, K ring of 3, V ring of 2 (112 KB per CTA), output rows stored straight from registers.
#include "vz_common.cuh"

#include <algorithm>
#include <vector>

namespace vz {
namespace {

constexpr int HD = 128;
constexpr int BQ = 128, BKV = 64;
constexpr int ATOM_Q = BQ * 128;            // one 64-column swizzle atom of the Q tile (16 KB)
constexpr int ATOM_KV = BKV * 128;          // ... of a K or V block (8 KB)
constexpr int Q_BYTES = 2 * ATOM_Q;         // 32 KB
constexpr int KV_BYTES = 2 * ATOM_KV;       // 16 KB
constexpr int K_STAGES = 3, V_STAGES = 2;
constexpr int SMEM_Q = 0;
constexpr int SMEM_K = Q_BYTES;
constexpr int SMEM_V = SMEM_K + K_STAGES * KV_BYTES;
constexpr int SMEM_BARS = SMEM_V + V_STAGES * KV_BYTES;
constexpr int SMEM_TOTAL = SMEM_BARS + 256;
static_assert(2 * (SMEM_TOTAL + 1024) <= 228 * 1024, "two CTAs per SM");
constexpr int THREADS = 192;
constexpr uint32_t TMEM_COLS = 256;         // S0: 0..63 (P0 over 0..31), S1: 64..127 (P1 over 64..95), O: 128..255
constexpr uint32_t TMEM_O = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleLog2 = 8.0f;        // see vz_attn_tc.cu: lazy rescale threshold

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// item = {first query row (global), first row of the sample, visible keys of the tile's last row + 1, head | qt << 8}
struct ItemWalk {
  const int4* items;
  int n_items, c, G;
  int round;        // next round to fetch
  uint32_t n;       // items this walker has finished
  int4 it;
  int nkb, j;
  bool valid;
  __device__ __forceinline__ void fetch() {
    const int idx = round * G + ((round & 1) ? G - 1 - c : c);
    valid = idx < n_items;
    if (valid) {
      it = __ldg(items + idx);
      nkb = (it.z + BKV - 1) / BKV;
    }
    j = 0;
    ++round;
  }
  __device__ __forceinline__ ItemWalk(const int4* items_, int n_items_, int c_, int G_)
      : items(items_), n_items(n_items_), c(c_), G(G_), round(0), n(0) { fetch(); }
  __device__ __forceinline__ void next_block() {
    if (++j == nkb) { ++n; fetch(); }
  }
  __device__ __forceinline__ int head() const { return it.w & 0xff; }
};

struct SoftmaxState {
  float m_used = -INFINITY;
  float l = 0.f;
  float sl2;
};

// One 64-key block of the online softmax for query row `p` (position inside the sample); g = this CTA's running
// key-block counter (buffer b = g & 1), j = block index inside the item, key0 = first key of the block.
template <bool MASKED>
__device__ __forceinline__ void softmax_block(SoftmaxState& s, uint32_t g, int j, int p, int key0, uint32_t t_lane,
                                              uint32_t tmem_base, uint32_t tmem_o, uint64_t* bar_s_full,
                                              uint64_t* bar_p_full, uint64_t* bar_pv_done) {
  const int lane = threadIdx.x & 31;
  const uint32_t b = g & 1, use = g >> 1;
  mbar_wait(&bar_s_full[b], use & 1, 600 + b);
  tc_fence_after();
  uint32_t v[BKV];
#pragma unroll
  for (int c = 0; c < 2; ++c)
    tmem_ld_32x32b_x32(tmem_base + t_lane + b * BKV + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
  tmem_ld_wait();
  const int lim = p - key0;     // keys key0 .. key0 + lim are visible (lim < 0: none of this block)
  if (MASKED) {
#pragma unroll
    for (int i = 0; i < BKV; ++i)
      if (i > lim) v[i] = 0xff800000u;   // -inf
  }
  float bm4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < BKV; ++i) bm4[i & 3] = fmaxf(bm4[i & 3], __uint_as_float(v[i]));
  const float bm = fmaxf(fmaxf(bm4[0], bm4[1]), fmaxf(bm4[2], bm4[3]));
  float alpha = 1.f;
  // (a fully masked block has bm = -inf: the comparison is false, nothing changes)
  const bool need = (bm - s.m_used) * s.sl2 > kRescaleLog2;   // true on the first block (m_used = -inf, bm finite)
  if (need) {
    alpha = ex2_approx((s.m_used - bm) * s.sl2);
    s.m_used = bm;
    s.l *= alpha;
  }
  const bool any_need = __any_sync(0xffffffffu, need) && j > 0;
  const float m_sl2 = s.m_used * s.sl2;
  uint32_t pk[BKV / 2];
  float ls4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < BKV; i += 2) {
    const float x0 = fmaf(__uint_as_float(v[i]), s.sl2, -m_sl2), x1 = fmaf(__uint_as_float(v[i + 1]), s.sl2, -m_sl2);
    // masked scores are -inf: ex2(-inf) = 0 (and a row that has seen no key yet cannot occur: key 0 is always visible)
    const float p0 = (!MASKED || i <= lim) ? ex2_approx(x0) : 0.f;
    const float p1 = (!MASKED || i + 1 <= lim) ? ex2_approx(x1) : 0.f;
    ls4[(i >> 1) & 3] += p0 + p1;
    pk[i >> 1] = pack_bf16x2(p0, p1);
  }
  s.l += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
  if (any_need) {
    // every earlier P V must have retired before O is touched (MMAs retire in order)
    mbar_wait(&bar_pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1, 620);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < HD / 32; ++c) {
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
  // P_j goes over the first 32 columns of S_j (this thread holds its S row in registers; the buffer's next
  // tenant Q K_{j+2}^T is issued behind P_j V_j)
  tmem_st_32x32b_x32(tmem_base + t_lane + b * BKV, pk);
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&bar_p_full[b]);
}

__global__ void __launch_bounds__(THREADS, 2)
attn_causal_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   __nv_bfloat16* __restrict__ out, int ldo, const int4* __restrict__ items, int n_items,
                   int q_cols, int kv_cols, int group, float scale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem + SMEM_Q;
  uint8_t* sK = smem + SMEM_K;
  uint8_t* sV = smem + SMEM_V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BARS);
  uint64_t* bar_q_full = bars;                      // Q of an item landed
  uint64_t* bar_q_empty = bars + 1;                 // every Q K^T of the item retired
  uint64_t* bar_k_full = bars + 2;                  // [K_STAGES]
  uint64_t* bar_k_empty = bar_k_full + K_STAGES;    // [K_STAGES] Q K_j^T retired
  uint64_t* bar_v_full = bar_k_empty + K_STAGES;    // [V_STAGES]
  uint64_t* bar_v_empty = bar_v_full + V_STAGES;    // [V_STAGES] P V_j retired
  uint64_t* bar_s_full = bar_v_empty + V_STAGES;    // [2] S buffer written by Q K^T
  uint64_t* bar_p_full = bar_s_full + 2;            // [2] P written (and O rescaled if needed): 4 warp arrivals
  uint64_t* bar_pv_done = bar_p_full + 2;           // [2] P V retired
  uint64_t* bar_o_full = bar_pv_done + 2;           // every MMA of the item retired
  uint64_t* bar_o_free = bar_o_full + 1;            // O copied out (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o_free + 1);
  static_assert((2 + 2 * K_STAGES + 2 * V_STAGES + 8) * 8 + 4 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 0) {
    mbar_init(bar_q_full, 1);
    mbar_init(bar_q_empty, 1);
    for (int i = 0; i < K_STAGES; ++i) { mbar_init(&bar_k_full[i], 1); mbar_init(&bar_k_empty[i], 1); }
    for (int i = 0; i < V_STAGES; ++i) { mbar_init(&bar_v_full[i], 1); mbar_init(&bar_v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s_full[i], 1);
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
  const int c = blockIdx.x, G = gridDim.x;

  if (warp == 5) {
    // ======================= TMA producer: K (+ Q) and V streams, each as far ahead as its ring allows =======================
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
    }
    ItemWalk kw(items, n_items, c, G), vw(items, n_items, c, G);
    uint32_t kg = 0, vg = 0;
    while (kw.valid || vw.valid) {
      if (kw.valid) {
        const uint32_t kst = kg % K_STAGES;
        bool ok = mbar_try_wait(&bar_k_empty[kst], ((kg / K_STAGES) & 1) ^ 1);
        // the single Q buffer: the previous item's last Q K^T must have retired (never blocks the V stream)
        if (ok && kw.j == 0 && kw.n > 0) ok = mbar_try_wait(bar_q_empty, (kw.n - 1) & 1);
        if (__shfl_sync(0xffffffffu, (int)ok, 0)) {
          const int kvh = kw.head() / group;
          if (elect_one()) {
            if (kw.j == 0) {
              mbar_arrive_expect_tx(bar_q_full, Q_BYTES);
              tma_load_2d(&tmQ, bar_q_full, sQ, kw.head() * HD, kw.it.x);
              tma_load_2d(&tmQ, bar_q_full, sQ + ATOM_Q, kw.head() * HD + 64, kw.it.x);
            }
            mbar_arrive_expect_tx(&bar_k_full[kst], KV_BYTES);
            tma_load_2d(&tmKV, &bar_k_full[kst], sK + kst * KV_BYTES, q_cols + kvh * HD, kw.it.y + kw.j * BKV);
            tma_load_2d(&tmKV, &bar_k_full[kst], sK + kst * KV_BYTES + ATOM_KV, q_cols + kvh * HD + 64,
                        kw.it.y + kw.j * BKV);
          }
          __syncwarp();
          ++kg;
          kw.next_block();
        }
      }
      if (vw.valid) {
        const uint32_t vst = vg % V_STAGES;
        if (__shfl_sync(0xffffffffu, (int)mbar_try_wait(&bar_v_empty[vst], ((vg / V_STAGES) & 1) ^ 1), 0)) {
          const int kvh = vw.head() / group;
          if (elect_one()) {
            mbar_arrive_expect_tx(&bar_v_full[vst], KV_BYTES);
            tma_load_2d(&tmKV, &bar_v_full[vst], sV + vst * KV_BYTES, q_cols + kv_cols + kvh * HD, vw.it.y + vw.j * BKV);
            tma_load_2d(&tmKV, &bar_v_full[vst], sV + vst * KV_BYTES + ATOM_KV, q_cols + kv_cols + kvh * HD + 64,
                        vw.it.y + vw.j * BKV);
          }
          __syncwarp();
          ++vg;
          vw.next_block();
        }
      }
    }
  } else if (warp == 4) {
    // ======================= MMA issuer (warp-uniform, one elected lane issues) =======================
    constexpr uint32_t idesc_qk = umma_idesc_bf16_ex(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16_ex(BQ, HD, 0, 1);   // B = V is MN-major
    const uint64_t q_desc0 = umma_smem_desc_sw128(smem_u32(sQ));
    const uint64_t k_desc0 = umma_smem_desc_sw128(smem_u32(sK));
    // V: [64 keys][128 dims] as two 64-column atoms 8 KB apart (LBO); 16 keys = 2 KB
    const uint64_t v_desc0 = umma_smem_desc_sw128(smem_u32(sV)) + ((uint64_t)((ATOM_KV >> 4) - 1) << 16);
    ItemWalk cur(items, n_items, c, G), ahead(items, n_items, c, G);
    // S[g2 & 1] = Q K_j^T of the CTA's g2-th key block
    auto issue_qk = [&](uint32_t g2) {
      const uint32_t b = g2 & 1, kst = g2 % K_STAGES;
      if (ahead.j == 0) mbar_wait(bar_q_full, ahead.n & 1, 510);
      mbar_wait(&bar_k_full[kst], (g2 / K_STAGES) & 1, 520 + kst);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t k_desc = k_desc0 + (uint64_t)(kst * (KV_BYTES >> 4));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem_base + b * BKV, q_desc0 + (uint64_t)((k >> 2) * (ATOM_Q >> 4) + (k & 3) * 2),
                    k_desc + (uint64_t)((k >> 2) * (ATOM_KV >> 4) + (k & 3) * 2), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(&bar_s_full[b]);
        umma_commit(&bar_k_empty[kst]);
        if (ahead.j == ahead.nkb - 1) umma_commit(bar_q_empty);   // the item's last use of Q
      }
      __syncwarp();
      ahead.next_block();
    };
    uint32_t g = 0;
    for (uint32_t a = 0; a < 2 && ahead.valid; ++a) issue_qk(a);
    while (cur.valid) {
      const uint32_t b = g & 1, use = g >> 1, vst = g % V_STAGES;
      mbar_wait(&bar_v_full[vst], (g / V_STAGES) & 1, 535 + vst);
      mbar_wait(&bar_p_full[b], use & 1, 540 + b);
      if (cur.j == 0 && cur.n > 0) mbar_wait(bar_o_free, (cur.n - 1) & 1, 550);   // previous item's O is in registers
      tc_fence_after();
      if (elect_one()) {
        const uint32_t p_tmem = tmem_base + b * BKV;
        const uint64_t v_desc = v_desc0 + (uint64_t)(vst * (KV_BYTES >> 4));
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk)
          umma_bf16_ts(tmem_o, p_tmem + kk * 8, v_desc + (uint64_t)(kk * (2048 >> 4)), idesc_pv,
                       (cur.j > 0 || kk != 0) ? 1u : 0u);
        umma_commit(&bar_pv_done[b]);
        umma_commit(&bar_v_empty[vst]);
        if (cur.j == cur.nkb - 1) umma_commit(bar_o_full);
      }
      __syncwarp();
      // the buffer's next tenant, behind P V in issue order (which is execution order on the tensor pipe)
      if (ahead.valid) issue_qk(g + 2);
      cur.next_block();
      ++g;
    }
  } else {
    // ======================= softmax warps: thread = query row = TMEM lane =======================
    const int r = warp * 32 + lane;
    const uint32_t t_lane = ((uint32_t)(warp * 32)) << 16;
    // epilogue of the CTA's en-th item: O / l -> bf16 -> this row of the output
    auto epilogue = [&](const int4& it, uint32_t en, float l, bool active) {
      mbar_wait(bar_o_full, en & 1, 640);
      tc_fence_after();
      const int h = it.w & 0xff, qt = it.w >> 8;
      const bool row_ok = active && qt * BQ + r < it.z;
      const float inv = 1.0f / l;
      __nv_bfloat16* dst = out + (size_t)(it.x + r) * ldo + h * HD;
#pragma unroll 1
      for (int cch = 0; cch < HD / 32; ++cch) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(tmem_o + t_lane + cch * 32, o);
        tmem_ld_wait();
        if (cch == HD / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_o_free);   // the next item's first P V may overwrite O now
        }
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
            w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
            w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
            w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + cch * 32 + i * 8) = w;
          }
        }
      }
    };
    bool pend = false, pact = false;
    int4 pit = make_int4(0, 0, 0, 0);
    float pl = 1.f;
    uint32_t g = 0;
    ItemWalk w(items, n_items, c, G);
    while (w.valid) {
      const int4 it = w.it;
      const int qt = it.w >> 8, nkb = w.nkb;
      const uint32_t n = w.n;
      // rows of this warp that exist: the tile's rows are positions qt * 128 .. it.z - 1 of the sample
      const bool active = qt * BQ + warp * 32 < it.z;
      const int p = qt * BQ + r;
      SoftmaxState stt;
      stt.sl2 = scale * kLog2e;
      auto idle_block = [&]() {
        const uint32_t b = g & 1, use = g >> 1;
        mbar_wait(&bar_s_full[b], use & 1, 600 + b);
        if (lane == 0) mbar_arrive(&bar_p_full[b]);
        __syncwarp();
      };
      for (int j = 0; j < nkb; ++j, ++g) {
        if (!active) idle_block();
        else if (j >= 2 * qt) softmax_block<true>(stt, g, j, p, j * BKV, t_lane, tmem_base, tmem_o, bar_s_full, bar_p_full, bar_pv_done);
        else softmax_block<false>(stt, g, j, p, j * BKV, t_lane, tmem_base, tmem_o, bar_s_full, bar_p_full, bar_pv_done);
        if (j == 0 && pend) { epilogue(pit, n - 1, pl, pact); pend = false; }
        w.next_block();
      }
      pend = true; pact = active; pit = it; pl = stt.l;
    }
    if (pend) epilogue(pit, w.n - 1, pl, pact);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace
}  // namespace vz

// Host side: the work list of a batch.  lens_h[b] = rows of sample b (HOST array).  items_h receives int32 x 4 per
// item, longest first; returns the number of items (or the number needed when max_items is too small / items_h NULL).
extern "C" int vz_attn_causal_items(const int32_t* lens_h, int B, int n_heads, int32_t* items_h, int max_items,
                                    double* algo_flops) {
  if (!lens_h || B <= 0 || n_heads <= 0 || n_heads > 255) return VZ_ERR_BAD_ARG;
  struct Tile { int q_row0, kv_row0, n_keys, qt, nkb; };
  std::vector<Tile> tiles;
  long row0 = 0;
  double flops = 0;
  for (int b = 0; b < B; ++b) {
    const int S = lens_h[b];
    if (S < 0) return VZ_ERR_BAD_ARG;
    for (int qt = 0; qt * vz::BQ < S; ++qt) {
      const int n_keys = std::min((qt + 1) * vz::BQ, S);
      tiles.push_back({(int)(row0 + (long)qt * vz::BQ), (int)row0, n_keys, qt, (n_keys + vz::BKV - 1) / vz::BKV});
    }
    // algorithmic: every (query, visible key) pair once for Q K^T and once for P V, head dim 128
    flops += 2.0 * 2.0 * vz::HD * (0.5 * (double)S * (S + 1)) * n_heads;
    row0 += S;
    if (row0 > 0x7fffffffL) return VZ_ERR_UNSUPPORTED;
  }
  if (algo_flops) *algo_flops = flops;
  const long n = (long)tiles.size() * n_heads;
  if (n > 0x7fffffffL) return VZ_ERR_UNSUPPORTED;
  if (!items_h || n > max_items) return (int)n;
  std::stable_sort(tiles.begin(), tiles.end(), [](const Tile& a, const Tile& b) { return a.nkb > b.nkb; });
  long k = 0;
  for (const Tile& t : tiles)
    for (int h = 0; h < n_heads; ++h, ++k) {
      items_h[4 * k + 0] = t.q_row0;
      items_h[4 * k + 1] = t.kv_row0;
      items_h[4 * k + 2] = t.n_keys;
      items_h[4 * k + 3] = h | (t.qt << 8);
    }
  return (int)n;
}

extern "C" int vz_attn_causal(const void* qkv, int ld, int M, void* out, int ldo, const int32_t* items, int n_items,
                              int n_heads, int n_kv_heads, int head_dim, float scale, double algo_flops, void* stream) {
  if (!qkv || !out || !items || M <= 0 || n_items <= 0) return VZ_ERR_BAD_ARG;
  if (head_dim != vz::HD || n_heads <= 0 || n_heads > 255 || n_kv_heads <= 0 || n_heads % n_kv_heads) return VZ_ERR_UNSUPPORTED;
  const int q_cols = n_heads * vz::HD, kv_cols = n_kv_heads * vz::HD;
  if (ld < q_cols + 2 * kv_cols || (ld & 7) || (ldo & 7) || ldo < q_cols || !vz::aligned16(qkv) || !vz::aligned16(out) ||
      !vz::aligned16(items))
    return VZ_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap tmQ, tmKV;
  VZ_TRY(vz::encode_tmap_2d_bf16(&tmQ, qkv, M, q_cols + 2 * kv_cols, ld, 64, vz::BQ));
  VZ_TRY(vz::encode_tmap_2d_bf16(&tmKV, qkv, M, q_cols + 2 * kv_cols, ld, 64, vz::BKV));
  int dev = 0, num_sms = 0;
  VZ_CUDA_CHECK(cudaGetDevice(&dev));
  VZ_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;
  vz::ProfScope prof(VZ_PROF_LLM_ATTN, algo_flops, st);
  VZ_ENSURE_DYN_SMEM(vz::attn_causal_kernel, vz::SMEM_TOTAL);
  vz::attn_causal_kernel<<<grid, vz::THREADS, vz::SMEM_TOTAL, st>>>(
      tmQ, tmKV, reinterpret_cast<__nv_bfloat16*>(out), ldo, reinterpret_cast<const int4*>(items), n_items, q_cols,
      kv_cols, n_heads / n_kv_heads, scale);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}
