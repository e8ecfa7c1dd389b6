// vz_llm.cu -- the row kernels of the LLM prefill behind the splice: rotary position
// embedding applied in place to the packed q | k | v rows, and the per-row statistics that let the first
// RMSNorm fold into the first GEMM.  Everything else of a decoder layer is vz_gemm_bf16 (RMSNorm and SwiGLU fused).
//
// Replaces, per decoder layer, HF MistralAttention's apply_rotary_pos_emb (the reference runs HF Mistral under
// language_model/vis_zephyr.py:86-98, optionally through train/zephyr_flash_attn_monkey_patch.py:86-136).
#include "vz_common.cuh"

namespace vz {
namespace {

// cos / sin of position * inv_freq, rounded to bf16 like MistralRotaryEmbedding.forward returns them
// (cos.to(dtype=x.dtype)); one table per prefill call, shared by all layers and heads.
__global__ void __launch_bounds__(256)
rope_table_kernel(const int32_t* __restrict__ pos, int M, const float* __restrict__ inv_freq, int half,
                  float2* __restrict__ cs) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)M * half) return;
  const int m = (int)(i / half), j = (int)(i - (long)m * half);
  const float ang = (float)pos[m] * inv_freq[j];
  float s, c;
  sincosf(ang, &s, &c);
  cs[i] = make_float2(__bfloat162float(__float2bfloat16_rn(c)), __bfloat162float(__float2bfloat16_rn(s)));
}

__device__ __forceinline__ void unpack8f(const uint4& w, float* f) {
  f[0] = bf16_lo(w.x); f[1] = bf16_hi(w.x); f[2] = bf16_lo(w.y); f[3] = bf16_hi(w.y);
  f[4] = bf16_lo(w.z); f[5] = bf16_hi(w.z); f[6] = bf16_lo(w.w); f[7] = bf16_hi(w.w);
}

// x[m, h * hd + j] for j < hd/2 pairs with x[m, h * hd + j + hd/2] (rotate_half):
//   lo' = lo cos - hi sin, hi' = hi cos + lo sin, each product and the sum rounded to bf16 as the bf16 tensor
//   expression (q * cos) + (rotate_half(q) * sin) rounds them.
// One thread = 8 consecutive j of one (row, head): two 16-byte loads, two 16-byte stores.
__global__ void __launch_bounds__(256)
rope_apply_kernel(__nv_bfloat16* __restrict__ x, int ld, int M, int n_heads, int hd, const float2* __restrict__ cs) {
  const int half = hd >> 1, per_head = half >> 3;
  const long items = (long)M * n_heads * per_head;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long)gridDim.x * blockDim.x) {
    const int m = (int)(i / (n_heads * per_head));
    const int r = (int)(i - (long)m * (n_heads * per_head));
    const int h = r / per_head, j0 = (r - h * per_head) * 8;
    __nv_bfloat16* base = x + (size_t)m * ld + h * hd + j0;
    const uint4 wlo = *reinterpret_cast<const uint4*>(base);
    const uint4 whi = *reinterpret_cast<const uint4*>(base + half);
    float lo[8], hi[8], olo[8], ohi[8];
    unpack8f(wlo, lo);
    unpack8f(whi, hi);
    const float4* t = reinterpret_cast<const float4*>(cs + (size_t)m * half + j0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 v = __ldg(t + k);   // (cos, sin) of j0 + 2k and j0 + 2k + 1
      const float c[2] = {v.x, v.z}, s[2] = {v.y, v.w};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = 2 * k + u;
        const float a = __bfloat162float(__float2bfloat16_rn(lo[e] * c[u]));
        const float b = __bfloat162float(__float2bfloat16_rn(-hi[e] * s[u]));
        const float d = __bfloat162float(__float2bfloat16_rn(hi[e] * c[u]));
        const float g = __bfloat162float(__float2bfloat16_rn(lo[e] * s[u]));
        olo[e] = a + b;
        ohi[e] = d + g;
      }
    }
    uint4 o;
    o.x = pack_bf16x2(olo[0], olo[1]); o.y = pack_bf16x2(olo[2], olo[3]);
    o.z = pack_bf16x2(olo[4], olo[5]); o.w = pack_bf16x2(olo[6], olo[7]);
    *reinterpret_cast<uint4*>(base) = o;
    o.x = pack_bf16x2(ohi[0], ohi[1]); o.y = pack_bf16x2(ohi[2], ohi[3]);
    o.z = pack_bf16x2(ohi[4], ohi[5]); o.w = pack_bf16x2(ohi[6], ohi[7]);
    *reinterpret_cast<uint4*>(base + half) = o;
  }
}

// Row copies between the padded [B, L] layout of the splice and the packed layout of the prefill:
// dst row map[i] <- src row i (scatter) or dst row i <- src row map[i] (gather); 16 bytes per thread.
__global__ void __launch_bounds__(256)
rows_move_kernel(const uint4* __restrict__ src, long lds16, uint4* __restrict__ dst, long ldd16,
                 const int32_t* __restrict__ map, int n_rows, int w16, int gather) {
  const long items = (long)n_rows * w16;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / w16), c = (int)(i - (long)r * w16);
    const long o = map[r];
    if (o < 0) continue;
    if (gather) dst[(size_t)r * ldd16 + c] = __ldg(src + (size_t)o * lds16 + c);
    else dst[(size_t)o * ldd16 + c] = __ldg(src + (size_t)r * lds16 + c);
  }
}

}  // namespace
}  // namespace vz

extern "C" int vz_rope_table(const int32_t* positions, int M, const float* inv_freq, int half_dim, float* cos_sin,
                             void* stream) {
  if (!positions || !inv_freq || !cos_sin || M <= 0 || half_dim <= 0) return VZ_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long n = (long)M * half_dim;
  vz::rope_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(positions, M, inv_freq, half_dim,
                                                                     reinterpret_cast<float2*>(cos_sin));
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

extern "C" int vz_rope_apply(void* x, int ld, int M, int n_heads, int head_dim, const float* cos_sin, void* stream) {
  if (!x || !cos_sin || M <= 0 || n_heads <= 0) return VZ_ERR_BAD_ARG;
  if (head_dim % 16 != 0 || (ld & 7) || !vz::aligned16(x) || !vz::aligned16(cos_sin)) return VZ_ERR_UNSUPPORTED;
  if ((long)n_heads * head_dim > ld) return VZ_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long items = (long)M * n_heads * (head_dim / 16);
  const long blocks = (items + 255) / 256;
  // algorithmic bytes: every rotated element read and written once
  vz::ProfScope prof(VZ_PROF_OTHER, (double)M * n_heads * head_dim * 4.0, st);
  vz::rope_apply_kernel<<<(unsigned)(blocks < 148L * 16 ? blocks : 148L * 16), 256, 0, st>>>(
      reinterpret_cast<__nv_bfloat16*>(x), ld, M, n_heads, head_dim, reinterpret_cast<const float2*>(cos_sin));
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

extern "C" int vz_row_stats(const void* x, int ldx, int M, int D, float* stats, void* stream) {
  if (!x || !stats || M <= 0 || D <= 0) return VZ_ERR_BAD_ARG;
  if ((ldx & 7) || !vz::aligned16(x)) return VZ_ERR_UNSUPPORTED;
  return vz::row_stats_launch(x, ldx, M, D, stats, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int vz_rows_move(const void* src, long long lds_bytes, void* dst, long long ldd_bytes, const int32_t* map,
                            int n_rows, int row_bytes, int gather, void* stream) {
  if (!src || !dst || !map || n_rows <= 0 || row_bytes <= 0) return VZ_ERR_BAD_ARG;
  if ((row_bytes & 15) || (lds_bytes & 15) || (ldd_bytes & 15) || !vz::aligned16(src) || !vz::aligned16(dst))
    return VZ_ERR_UNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long items = (long)n_rows * (row_bytes / 16);
  const long blocks = (items + 255) / 256;
  vz::ProfScope prof(VZ_PROF_OTHER, (double)n_rows * row_bytes * 2.0, st);
  vz::rows_move_kernel<<<(unsigned)(blocks < 148L * 16 ? blocks : 148L * 16), 256, 0, st>>>(
      reinterpret_cast<const uint4*>(src), lds_bytes / 16, reinterpret_cast<uint4*>(dst), ldd_bytes / 16, map, n_rows,
      row_bytes / 16, gather);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}
