// testonly/vz_attn_legacy.cu -- the FIRST implementation of the CLIP ViT self-attention (flash-style on
// mma.sync m16n8k16, cp.async double buffering), kept as an independent cross-check of vz_attn_tc.cu for the
// tests (tests/test_gpu_attention.py).  Built into libvz_b200_testonly.so; NOT linked into libvz_b200.so, so
// the product path has no way to reach it.
#include "../vz_common.cuh"

namespace vz {
thread_local int g_last_cuda_error = 0;   // this library's own copy (vz_common.cuh declares it extern)
void count_launch() {}
namespace {

constexpr float kLog2e = 1.4426950408889634f;

// ==========================================================================================
// ViT attention
// ==========================================================================================
constexpr int VA_BQ = 64, VA_BK = 64, VA_D = 64, VA_THREADS = 128;

// element (row, col) of a [rows][64] bf16 tile with 16-byte chunks XOR-swizzled by row
__device__ __forceinline__ uint32_t sw64(int row, int col) {
  return (uint32_t)(row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1)));
}

__global__ void __launch_bounds__(VA_THREADS)
vit_attn_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int ntok,
                float scale) {
  __shared__ __align__(128) uint8_t sQ[VA_BQ * 128];
  __shared__ __align__(128) uint8_t sK[2][VA_BK * 128];
  __shared__ __align__(128) uint8_t sV[2][VA_BK * 128];

  const int qb = blockIdx.x, h = blockIdx.y, t = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int ld = 3 * VZ_VIT_WIDTH;
  const __nv_bfloat16* base = qkv + (size_t)t * ntok * ld + h * VA_D;
  const int nchunks = (ntok + VA_BK - 1) / VA_BK;

  // ---- async loads -------------------------------------------------------------------
  auto load_tile = [&](uint8_t* dst, int row0, int col_off) {
#pragma unroll
    for (int i = 0; i < (64 * 8) / VA_THREADS; ++i) {
      const int idx = tid + i * VA_THREADS;
      const int r = idx >> 3, c = idx & 7;
      const int row = row0 + r;
      const bool ok = row < ntok;
      const __nv_bfloat16* src = base + (size_t)(ok ? row : ntok - 1) * ld + col_off + c * 8;
      cp_async_16(dst + sw64(r, c * 8), src, ok);
    }
  };
  load_tile(sQ, qb * VA_BQ, 0);
  load_tile(sK[0], 0, VZ_VIT_WIDTH);
  load_tile(sV[0], 0, 2 * VZ_VIT_WIDTH);
  cp_async_commit();

  uint32_t qf[4][4];  // A fragments of this warp's 16 query rows, 4 k-steps
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sl2 = scale * kLog2e;

  for (int j = 0; j < nchunks; ++j) {
    const int buf = j & 1;
    if (j + 1 < nchunks) {
      load_tile(sK[buf ^ 1], (j + 1) * VA_BK, VZ_VIT_WIDTH);
      load_tile(sV[buf ^ 1], (j + 1) * VA_BK, 2 * VZ_VIT_WIDTH);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
      const uint32_t qbase = smem_u32(sQ);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldmatrix_x4(qf[ks], qbase + sw64(warp * 16 + (lane & 15), ks * 16 + (lane >> 4) * 8));
    }
    // ---- S = Q K^T -------------------------------------------------------------------
    float s[8][4];
    const uint32_t kbase = smem_u32(sK[buf]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {  // two k-steps per ldmatrix_x4
        uint32_t bfr[4];
        ldmatrix_x4(bfr, kbase + sw64(nt * 8 + (lane & 7), kp * 32 + (lane >> 3) * 8));
        const uint32_t b0[2] = {bfr[0], bfr[1]}, b1[2] = {bfr[2], bfr[3]};
        mma_bf16_16816(s[nt], qf[kp * 2], b0);
        mma_bf16_16816(s[nt], qf[kp * 2 + 1], b1);
      }
    }
    // ---- mask + online softmax ----------------------------------------------------------
    const int key0 = j * VA_BK;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int kc = key0 + nt * 8 + tq * 2;
      if (kc >= ntok) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (kc + 1 >= ntok) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      alpha[r] = exp2f((m_run[r] - m_new) * sl2);
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as A fragments for 4 k-steps of 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f((s[nt][0] - m_run[0]) * sl2);
      const float p1 = exp2f((s[nt][1] - m_run[0]) * sl2);
      const float p2 = exp2f((s[nt][2] - m_run[1]) * sl2);
      const float p3 = exp2f((s[nt][3] - m_run[1]) * sl2);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      const int kk = nt >> 1, hi = nt & 1;
      pf[kk][hi * 2 + 0] = pack_bf16x2(p0, p1);
      pf[kk][hi * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + rs[r];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      o[dt][0] *= alpha[0]; o[dt][1] *= alpha[0];
      o[dt][2] *= alpha[1]; o[dt][3] *= alpha[1];
    }
    // ---- O += P V --------------------------------------------------------------------
    const uint32_t vbase = smem_u32(sV[buf]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {  // two dim-tiles per ldmatrix_x4.trans
        uint32_t bfr[4];
        const int jm = lane >> 3;
        ldmatrix_x4_trans(bfr, vbase + sw64(kk * 16 + (jm & 1) * 8 + (lane & 7), (dp * 2 + (jm >> 1)) * 8));
        const uint32_t b0[2] = {bfr[0], bfr[1]}, b1[2] = {bfr[2], bfr[3]};
        mma_bf16_16816(o[dp * 2], pf[kk], b0);
        mma_bf16_16816(o[dp * 2 + 1], pf[kk], b1);
      }
    }
    __syncthreads();  // all warps done with buf before it is refilled
  }
  // ---- finalise --------------------------------------------------------------------------
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
  const int row0 = qb * VA_BQ + warp * 16 + g;
  __nv_bfloat16* obase = out + (size_t)t * ntok * VZ_VIT_WIDTH + h * VA_D;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int col = dt * 8 + tq * 2;
    if (row0 < ntok)
      *reinterpret_cast<uint32_t*>(obase + (size_t)row0 * VZ_VIT_WIDTH + col) =
          pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
    if (row0 + 8 < ntok)
      *reinterpret_cast<uint32_t*>(obase + (size_t)(row0 + 8) * VZ_VIT_WIDTH + col) =
          pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
  }
}

}  // namespace
}  // namespace vz

// qkv bf16 [T*577, 3072] (q | k | v, 16 heads x 64) -> out bf16 [T*577, 1024]
extern "C" int vz_test_vit_attention_legacy(const void* qkv, void* out, int T, void* stream) {
  using namespace vz;
  if (!qkv || !out || T <= 0) return VZ_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((VZ_VIT_TOKENS + VA_BQ - 1) / VA_BQ, VZ_VIT_HEADS, T);
  vit_attn_kernel<<<grid, VA_THREADS, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                               reinterpret_cast<__nv_bfloat16*>(out), VZ_VIT_TOKENS, 0.125f);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}
