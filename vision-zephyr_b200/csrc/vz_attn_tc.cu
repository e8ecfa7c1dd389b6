// vz_attn_tc.cu -- CLIP ViT self-attention (577 tokens, 16 heads x 64, non-causal) on the 5th-gen
// tensor cores: S = Q K^T and O = P V are tcgen05.mma instructions with both accumulators in TMEM;
// Q / K / V tiles arrive by TMA (128-byte swizzle) straight from the packed qkv activation.
// Replaces HF CLIPAttention as called from vision_encoder/vision_encoder.py:101-105.
//
// One CTA = 128 query rows of one (tile, head); two CTAs are resident per SM so one CTA's softmax
// overlaps the other's MMAs.  Warps 0-7: softmax (two threads per query row, each owning 64 of the
// block's 128 keys); warp 8: one thread issues TMA and MMA.  One pass over the 5 key blocks (keys padded 577 -> 640 and masked):
//   S_j = Q K_j^T -> registers -> P_j = exp2((S_j - m) * scale) as bf16 in swizzled smem -> O += P_j V_j
// with an online softmax whose accumulator rescale is LAZY: O (in TMEM) is only multiplied by
// exp2(m_old - m_new) when the running maximum grew by more than 2^8, which is rare after the first
// block, so O normally stays untouched in TMEM until the epilogue divides by the row sum.
// Q K_{j+1}^T is issued as soon as S_j sits in registers, so it overlaps the exponentials of block j.
#include "vz_common.cuh"

namespace vz {
namespace {

constexpr int TOK = VZ_VIT_TOKENS;       // 577
constexpr int HD = 64;                   // head dim
constexpr int BQ = 128, BKV = 128;       // query rows per CTA, keys per block
constexpr int NKB = (TOK + BKV - 1) / BKV;  // 5 key blocks
constexpr int TILE_BYTES = 128 * 128;    // 128 rows x 64 bf16
constexpr int SMEM_Q = 0, SMEM_P = TILE_BYTES, SMEM_RING = 3 * TILE_BYTES;   // P = 2 tiles
constexpr int RING_STAGES = 2;
constexpr int SMEM_BARS = SMEM_RING + RING_STAGES * 2 * TILE_BYTES;
constexpr int SMEM_XCHG = SMEM_BARS + 128;            // 512 B: u16 [2 halves][128 rows] / float [128 rows]
constexpr int SMEM_TOTAL = SMEM_XCHG + 512;
// two CTAs per SM: 2 * (SMEM_TOTAL + 1 KB reserved) must fit the SM's 228 KB
static_assert(2 * (SMEM_TOTAL + 1024) <= 228 * 1024, "attention kernel must keep 2 CTAs per SM");
constexpr int THREADS = 288;        // 8 softmax warps (2 threads per query row) + 1 control warp
constexpr int SOFTMAX_THREADS = 256;
constexpr uint32_t TMEM_COLS = 256;      // S: columns 0..127, O: columns 128..191
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
vit_attn_tc_kernel(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* __restrict__ out, float scale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem + SMEM_Q;
  uint8_t* sP = smem + SMEM_P;
  uint8_t* sRing = smem + SMEM_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BARS);
  uint64_t* bar_q = bars;             // Q landed
  uint64_t* bar_full = bars + 1;      // [2] ring slot filled (TMA)
  uint64_t* bar_empty = bars + 3;     // [2] ring slot consumed (tcgen05.commit)
  uint64_t* bar_s_full = bars + 5;    // S ready in TMEM
  uint64_t* bar_s_free = bars + 6;    // S copied to registers (4 warp arrivals)
  uint64_t* bar_p_full = bars + 7;    // P written to smem, O rescaled if needed (4 warp arrivals)
  uint64_t* bar_pv_done = bars + 8;   // P V retired: P buffer and O reusable
  uint64_t* bar_o_full = bars + 9;    // all MMAs retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int qb = blockIdx.x, h = blockIdx.y, t = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_base = t * TOK;  // first qkv row of this tile

  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the swizzled layouts need a 1024-byte aligned base
  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    for (int i = 0; i < RING_STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    mbar_init(bar_s_full, 1);
    mbar_init(bar_s_free, 8);
    mbar_init(bar_p_full, 8);
    mbar_init(bar_pv_done, 1);
    mbar_init(bar_o_full, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

  if (warp == 8) {
    // ======================= control thread: TMA + MMA issue =======================
    if (lane == 0) {
      tma_prefetch_desc(&tm);
      const int qcol = h * HD, kcol = VZ_VIT_WIDTH + h * HD, vcol = 2 * VZ_VIT_WIDTH + h * HD;
      mbar_arrive_expect_tx(bar_q, TILE_BYTES);
      tma_load_2d(&tm, bar_q, sQ, qcol, row_base + qb * BQ);
      constexpr uint32_t idesc_qk = umma_idesc_bf16_ex(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_ex(BQ, HD, 0, 1);   // B = V is MN-major (dims contiguous)
      auto issue_load = [&](int j) {       // K_j and V_j into ring slot j % 2
        const uint32_t st = j % RING_STAGES, ph = (j / RING_STAGES) & 1;
        mbar_wait(&bar_empty[st], ph ^ 1, 500 + st);
        uint8_t* dK = sRing + st * 2 * TILE_BYTES;
        mbar_arrive_expect_tx(&bar_full[st], 2 * TILE_BYTES);
        tma_load_2d(&tm, &bar_full[st], dK, kcol, row_base + j * BKV);
        tma_load_2d(&tm, &bar_full[st], dK + TILE_BYTES, vcol, row_base + j * BKV);
      };
      auto issue_pv = [&](int j) {         // O (+)= P_j V_j, then release P, O and the ring slot
        const uint32_t st = j % RING_STAGES;
        mbar_wait(bar_p_full, j & 1, 540);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(sP), v_addr = smem_u32(sRing + st * 2 * TILE_BYTES + TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          const uint64_t a_desc = umma_smem_desc_sw128(p_addr + (kk >> 2) * TILE_BYTES + (kk & 3) * 32);
          const uint64_t b_desc = umma_smem_desc_sw128(v_addr + kk * 2048);  // 16 keys = 2 x (8 rows x 128 B)
          umma_bf16(tmem_o, a_desc, b_desc, idesc_pv, (j > 0 || kk != 0) ? 1u : 0u);
        }
        umma_commit(bar_pv_done);
        umma_commit(&bar_empty[st]);
      };
      issue_load(0);
      issue_load(1);
      mbar_wait(bar_q, 0, 510);
      const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(sQ));
      for (int j = 0; j < NKB; ++j) {
        const uint32_t st = j % RING_STAGES, ph = (j / RING_STAGES) & 1;
        mbar_wait(&bar_full[st], ph, 520 + st);
        if (j > 0) mbar_wait(bar_s_free, (j - 1) & 1, 530);  // S_{j-1} is in the softmax registers
        tc_fence_after();
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(sRing + st * 2 * TILE_BYTES));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem_s, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(bar_s_full);
        if (j > 0) {
          issue_pv(j - 1);                     // overlaps the softmax of block j
          if (j + 1 < NKB) issue_load(j + 1);  // slot (j+1)%2 == (j-1)%2 is free once P V_{j-1} retires
        }
      }
      issue_pv(NKB - 1);
      umma_commit(bar_o_full);
    }
  } else {
    // ======================= softmax warps: TWO threads per query row =======================
    // warp w: TMEM lane quarter q = w & 3 (rows 32q..32q+31), key half hf = w >> 2 (64 of the block's 128
    // keys, i.e. one 64-key swizzle atom of P, and 32 of O's 64 columns).  The two threads of a row
    // exchange their half maxima / sums through shared memory so both take identical rescale decisions.
    const int q4 = warp & 3, hf = warp >> 2;
    const int r = q4 * 32 + lane;                        // row inside the CTA's 128-query block
    const uint32_t t_lane = ((uint32_t)(q4 * 32)) << 16;
    uint16_t* xch16 = reinterpret_cast<uint16_t*>(smem + SMEM_XCHG);   // [half][row]
    float* xl = reinterpret_cast<float*>(smem + SMEM_XCHG);            // [row], epilogue only
    const float sl2 = scale * kLog2e;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < NKB; ++j) {
      mbar_wait(bar_s_full, j & 1, 600);
      tc_fence_after();
      uint32_t v[64];
#pragma unroll
      for (int c = 0; c < 2; ++c)
        tmem_ld_32x32b_x32(tmem_s + t_lane + hf * 64 + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
      tmem_ld_wait();
      // keys of the last block beyond the tile's 577 tokens belong to the next tile: mask them
      if (j == NKB - 1) {
        constexpr int nvalid = TOK - (NKB - 1) * BKV;   // 65 valid keys in the last block
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (hf * 64 + i >= nvalid) v[i] = 0xff800000u;  // -inf
      }
      float bm4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        bm4[0] = fmaxf(bm4[0], __uint_as_float(v[i]));
        bm4[1] = fmaxf(bm4[1], __uint_as_float(v[i + 1]));
        bm4[2] = fmaxf(bm4[2], __uint_as_float(v[i + 2]));
        bm4[3] = fmaxf(bm4[3], __uint_as_float(v[i + 3]));
      }
      const float hm = fmaxf(fmaxf(bm4[0], bm4[1]), fmaxf(bm4[2], bm4[3]));
      // The two threads of a row swap their half maxima as the upper 16 bits of the float; both then
      // use max(trunc(own), trunc(partner)), so they take IDENTICAL decisions (the reference point of
      // the exponentials only has to be common and close to the maximum, not the exact maximum).
      const uint32_t hb = __float_as_uint(hm) >> 16;
      xch16[hf * 128 + r] = (uint16_t)hb;
      asm volatile("bar.sync 1, %0;" ::"n"(SOFTMAX_THREADS) : "memory");
      const float bm = fmaxf(__uint_as_float(hb << 16), __uint_as_float((uint32_t)xch16[(hf ^ 1) * 128 + r] << 16));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_s_free);   // S_j is in registers AND the exchange slot is free again
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
      if (j > 0) mbar_wait(bar_pv_done, (j - 1) & 1, 620);   // P V_{j-1} retired: P and O are ours
      if (any_need) {
        tc_fence_after();
        uint32_t o[32];
        tmem_ld_32x32b_x32(tmem_o + t_lane + hf * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_32x32b_x32(tmem_o + t_lane + hf * 32, o);
        tmem_st_wait();
        tc_fence_before();
      }
      // this thread's 64 keys are exactly atom `hf` of P (K-major, 128B swizzle): 8 chunks of row r
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 w = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
        *reinterpret_cast<uint4*>(sP + hf * TILE_BYTES + r * 128 + ((c ^ (r & 7)) << 4)) = w;
      }
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p_full);
    }
    // ---- epilogue: O / l (row sum = both halves) ----
    asm volatile("bar.sync 1, %0;" ::"n"(SOFTMAX_THREADS) : "memory");   // last exchange fully consumed
    if (hf == 0) xl[r] = l;
    asm volatile("bar.sync 1, %0;" ::"n"(SOFTMAX_THREADS) : "memory");
    if (hf == 1) { l += xl[r]; xl[r] = l; }
    asm volatile("bar.sync 1, %0;" ::"n"(SOFTMAX_THREADS) : "memory");
    if (hf == 0) l = xl[r];
    mbar_wait(bar_o_full, 0, 630);
    tc_fence_after();
    const int qrow = qb * BQ + r;
    const float inv = 1.0f / l;
    __nv_bfloat16* orow = out + (size_t)(row_base + qrow) * VZ_VIT_WIDTH + h * HD + hf * 32;
    {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_o + t_lane + hf * 32, o);
      tmem_ld_wait();
      if (qrow < TOK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + i * 8) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

int vit_attn_tc_launch(const void* qkv, void* out, int T, cudaStream_t st) {
  CUtensorMap tm;
  VZ_TRY(encode_tmap_2d_bf16(&tm, qkv, (long long)T * TOK, 3 * VZ_VIT_WIDTH, 3 * VZ_VIT_WIDTH, HD, 128));
  static bool attr_done = false;
  if (!attr_done) {
    VZ_CUDA_CHECK(cudaFuncSetAttribute(vit_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_done = true;
  }
  dim3 grid((TOK + BQ - 1) / BQ, VZ_VIT_HEADS, T);
  vit_attn_tc_kernel<<<grid, THREADS, SMEM_TOTAL, st>>>(tm, reinterpret_cast<__nv_bfloat16*>(out), 0.125f);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

}  // namespace vz
