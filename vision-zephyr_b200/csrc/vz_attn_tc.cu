// vz_attn_tc.cu -- CLIP ViT self-attention (577 tokens, 16 heads x 64, non-causal) on the 5th-gen
// tensor cores: S = Q K^T and O = P V are tcgen05.mma instructions with both accumulators in TMEM;
// Q / K / V tiles arrive by TMA (128-byte swizzle) straight from the packed qkv activation.
// Replaces HF CLIPAttention as called from vision_encoder/vision_encoder.py:101-105.
//
// One CTA = 128 query rows of one (tile, head); two CTAs are resident per SM.  Warps 0-3: softmax
// (thread = query row = TMEM lane); warp 4: one thread issues the MMAs; warp 5: one thread issues TMA.
// The 577 keys are walked in 10 blocks of 64 (the last one masked) with EVERYTHING double buffered --
// two S accumulators in TMEM, two P tiles in smem, a 3-stage K/V ring -- so Q K_{j+1}^T is issued
// before the softmax of block j starts and P_j V_j runs while the softmax of block j+1 computes:
//   S_j = Q K_j^T -> registers -> P_j = exp2((S_j - m) * scale) as bf16 in swizzled smem -> O += P_j V_j
// The online softmax rescales the accumulator LAZILY: O (in TMEM) is only multiplied by
// exp2(m_old - m_new) when the running maximum grew by more than 2^8, which is rare after the first
// block, so O normally stays untouched in TMEM until the epilogue divides by the row sum.
#include "vz_common.cuh"

namespace vz {
namespace {

constexpr int TOK = VZ_VIT_TOKENS;          // 577
constexpr int HD = 64;                      // head dim
constexpr int BQ = 128, BKV = 64;           // query rows per CTA, keys per block
constexpr int NKB = (TOK + BKV - 1) / BKV;  // 10 key blocks (the last one holds a single valid key)
constexpr int Q_BYTES = BQ * 128;           // 128 rows x 64 bf16
constexpr int KV_BYTES = BKV * 128;         // 64 rows x 64 bf16
constexpr int P_BYTES = BQ * 128;           // 128 rows x 64 keys bf16 (one swizzle atom wide)
constexpr int KV_STAGES = 3;
constexpr int SMEM_Q = 0;
constexpr int SMEM_P = Q_BYTES;                               // 2 buffers
constexpr int SMEM_RING = SMEM_P + 2 * P_BYTES;               // 3 x (K, V)
constexpr int SMEM_BARS = SMEM_RING + KV_STAGES * 2 * KV_BYTES;
constexpr int SMEM_TOTAL = SMEM_BARS + 256;
static_assert(2 * (SMEM_TOTAL + 1024) <= 228 * 1024, "attention kernel must keep 2 CTAs per SM");
constexpr int THREADS = 192;                // 4 softmax warps + MMA warp + TMA warp
constexpr uint32_t TMEM_COLS = 256;         // S0: 0..63, S1: 64..127, O: 128..191
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// exp2 may run up to 2^8 above the value it would have with the exact running maximum before the
// accumulator is rescaled (FA-4 style lazy rescaling): keeps O untouched in TMEM most of the time.
constexpr float kRescaleLog2 = 8.0f;

__global__ void __launch_bounds__(THREADS, 2)
vit_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   __nv_bfloat16* __restrict__ out, float scale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem + SMEM_Q;
  uint8_t* sP = smem + SMEM_P;
  uint8_t* sRing = smem + SMEM_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BARS);
  uint64_t* bar_q = bars;              // Q landed
  uint64_t* bar_kv_full = bars + 1;    // [3] K_j, V_j landed
  uint64_t* bar_kv_empty = bars + 4;   // [3] P V_j retired -> slot reusable
  uint64_t* bar_s_full = bars + 7;     // [2] S buffer written by Q K^T
  uint64_t* bar_s_free = bars + 9;     // [2] S buffer copied to registers (4 warp arrivals)
  uint64_t* bar_p_full = bars + 11;    // [2] P buffer written (and O rescaled if needed) (4 warp arrivals)
  uint64_t* bar_pv_done = bars + 13;   // [2] P V retired -> P buffer reusable, O up to date
  uint64_t* bar_o_full = bars + 15;    // all MMAs retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int qb = blockIdx.x, h = blockIdx.y, t = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_base = t * TOK;  // first qkv row of this tile

  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the swizzled layouts need a 1024-byte aligned base
  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    for (int i = 0; i < KV_STAGES; ++i) { mbar_init(&bar_kv_full[i], 1); mbar_init(&bar_kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s_full[i], 1);
      mbar_init(&bar_s_free[i], 4);
      mbar_init(&bar_p_full[i], 4);
      mbar_init(&bar_pv_done[i], 1);
    }
    mbar_init(bar_o_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 5) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      const int qcol = h * HD, kcol = VZ_VIT_WIDTH + h * HD, vcol = 2 * VZ_VIT_WIDTH + h * HD;
      mbar_arrive_expect_tx(bar_q, Q_BYTES);
      tma_load_2d(&tmQ, bar_q, sQ, qcol, row_base + qb * BQ);
      for (int j = 0; j < NKB; ++j) {
        const uint32_t st = j % KV_STAGES, ph = (j / KV_STAGES) & 1;
        mbar_wait(&bar_kv_empty[st], ph ^ 1, 500 + st);
        uint8_t* dK = sRing + st * 2 * KV_BYTES;
        mbar_arrive_expect_tx(&bar_kv_full[st], 2 * KV_BYTES);
        tma_load_2d(&tmKV, &bar_kv_full[st], dK, kcol, row_base + j * BKV);
        tma_load_2d(&tmKV, &bar_kv_full[st], dK + KV_BYTES, vcol, row_base + j * BKV);
      }
    }
  } else if (warp == 4) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16_ex(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_ex(BQ, HD, 0, 1);   // B = V is MN-major (dims contiguous)
      mbar_wait(bar_q, 0, 510);
      const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(sQ));
      auto issue_qk = [&](int j) {          // S[j & 1] = Q K_j^T
        const uint32_t st = j % KV_STAGES, b = j & 1, use = j >> 1;
        mbar_wait(&bar_kv_full[st], (j / KV_STAGES) & 1, 520 + st);
        if (use > 0) mbar_wait(&bar_s_free[b], (use - 1) & 1, 530 + b);   // previous tenant is in registers
        tc_fence_after();
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(sRing + st * 2 * KV_BYTES));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem_base + b * BKV, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_qk,
                    k != 0 ? 1u : 0u);
        umma_commit(&bar_s_full[b]);
      };
      issue_qk(0);
      for (int j = 0; j < NKB; ++j) {
        if (j + 1 < NKB) issue_qk(j + 1);   // runs ahead of the softmax of block j
        const uint32_t st = j % KV_STAGES, b = j & 1, use = j >> 1;
        mbar_wait(&bar_p_full[b], use & 1, 540 + b);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(sP + b * P_BYTES);
        const uint32_t v_addr = smem_u32(sRing + st * 2 * KV_BYTES + KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          const uint64_t a_desc = umma_smem_desc_sw128(p_addr + kk * 32);
          const uint64_t b_desc = umma_smem_desc_sw128(v_addr + kk * 2048);  // 16 keys = 2 x (8 rows x 128 B)
          umma_bf16(tmem_o, a_desc, b_desc, idesc_pv, (j > 0 || kk != 0) ? 1u : 0u);
        }
        umma_commit(&bar_pv_done[b]);
        umma_commit(&bar_kv_empty[st]);
      }
      umma_commit(bar_o_full);
    }
  } else {
    // ======================= softmax warps: thread = query row = TMEM lane =======================
    const int r = warp * 32 + lane;
    const uint32_t t_lane = ((uint32_t)(warp * 32)) << 16;
    const float sl2 = scale * kLog2e;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < NKB; ++j) {
      const uint32_t b = j & 1, use = j >> 1;
      mbar_wait(&bar_s_full[b], use & 1, 600 + b);
      tc_fence_after();
      uint32_t v[64];
#pragma unroll
      for (int c = 0; c < 2; ++c)
        tmem_ld_32x32b_x32(tmem_base + t_lane + b * BKV + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_s_free[b]);
      // keys of the last block beyond the tile's 577 tokens belong to the next tile: mask them
      if (j == NKB - 1) {
        constexpr int nvalid = TOK - (NKB - 1) * BKV;   // 1 valid key in the last block
#pragma unroll
        for (int i = nvalid; i < 64; ++i) v[i] = 0xff800000u;  // -inf
      }
      float bm4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        bm4[0] = fmaxf(bm4[0], __uint_as_float(v[i]));
        bm4[1] = fmaxf(bm4[1], __uint_as_float(v[i + 1]));
        bm4[2] = fmaxf(bm4[2], __uint_as_float(v[i + 2]));
        bm4[3] = fmaxf(bm4[3], __uint_as_float(v[i + 3]));
      }
      const float bm = fmaxf(fmaxf(bm4[0], bm4[1]), fmaxf(bm4[2], bm4[3]));
      // lazy rescale: only when the running maximum grows by more than 2^kRescaleLog2
      float alpha = 1.f;
      const bool need = (bm - m_used) * sl2 > kRescaleLog2;   // true on the first block (m_used = -inf)
      if (need) {
        alpha = ex2_approx((m_used - bm) * sl2);               // 0 on the first block
        m_used = bm;
        l *= alpha;
      }
      const bool any_need = __any_sync(0xffffffffu, need) && j > 0;
      const float m_sl2 = m_used * sl2;
      uint32_t pk[32];
      float ls4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(v[i + 2 * u]), sl2, -m_sl2));
          const float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 2 * u + 1]), sl2, -m_sl2));
          ls4[u] += p0 + p1;
          pk[(i >> 1) + u] = pack_bf16x2(p0, p1);
        }
      }
      l += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      if (any_need) {
        // every earlier P V must have retired before O is touched (MMAs retire in order)
        mbar_wait(&bar_pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1, 620);
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
      // P[r][0..63] in the K-major 128B-swizzled UMMA layout (one atom): 8 chunks of row r
      uint8_t* pb = sP + b * P_BYTES;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 w = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
        *reinterpret_cast<uint4*>(pb + r * 128 + ((c ^ (r & 7)) << 4)) = w;
      }
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p_full[b]);
    }
    // ---- epilogue: O / l ----
    mbar_wait(bar_o_full, 0, 640);
    tc_fence_after();
    const int qrow = qb * BQ + r;
    const float inv = 1.0f / l;
    __nv_bfloat16* orow = out + (size_t)(row_base + qrow) * VZ_VIT_WIDTH + h * HD;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_o + t_lane + c * 32, o);
      tmem_ld_wait();
      if (qrow < TOK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = w;
        }
      }
    }
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
  CUtensorMap tmQ, tmKV;
  VZ_TRY(encode_tmap_2d_bf16(&tmQ, qkv, (long long)T * TOK, 3 * VZ_VIT_WIDTH, 3 * VZ_VIT_WIDTH, HD, BQ));
  VZ_TRY(encode_tmap_2d_bf16(&tmKV, qkv, (long long)T * TOK, 3 * VZ_VIT_WIDTH, 3 * VZ_VIT_WIDTH, HD, BKV));
  static bool attr_done = false;
  if (!attr_done) {
    VZ_CUDA_CHECK(cudaFuncSetAttribute(vit_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_done = true;
  }
  dim3 grid((TOK + BQ - 1) / BQ, VZ_VIT_HEADS, T);
  vit_attn_tc_kernel<<<grid, THREADS, SMEM_TOTAL, st>>>(tmQ, tmKV, reinterpret_cast<__nv_bfloat16*>(out), 0.125f);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

}  // namespace vz
