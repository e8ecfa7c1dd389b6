// vz_rows.cu -- HBM-bound row kernels: LayerNorm, multi-layer feature fusion (+pre_norm),
// CLS-row initialisation, row gather and pixel_values -> patch rows (im2col).
// All are 128-bit vectorised, one row per CTA (or warp), fp32 statistics.
#include "vz_common.cuh"

namespace vz {
namespace {

// block-wide sum of one float per thread; red must hold >= 32 floats
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect red from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nwarps) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

__device__ __forceinline__ void unpack8(const uint4& w, float (&f)[8]) {
  f[0] = bf16_lo(w.x); f[1] = bf16_hi(w.x);
  f[2] = bf16_lo(w.y); f[3] = bf16_hi(w.y);
  f[4] = bf16_lo(w.z); f[5] = bf16_hi(w.z);
  f[6] = bf16_lo(w.w); f[7] = bf16_hi(w.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 w;
  w.x = pack_bf16x2(f[0], f[1]);
  w.y = pack_bf16x2(f[2], f[3]);
  w.z = pack_bf16x2(f[4], f[5]);
  w.w = pack_bf16x2(f[6], f[7]);
  return w;
}

// ------------------------------------------------------------------------------------------
// LayerNorm: one row per CTA, D/8 threads (each thread owns one 16-byte vector).
// Two-pass statistics held in registers (mean, then centred sum of squares) -- same numerics
// class as ATen's Welford kernel; eps inside the sqrt like nn.LayerNorm.
// row_map (optional): input row = row_map[m / rows_per_map] * rows_per_map + m % rows_per_map.
// ------------------------------------------------------------------------------------------
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                 const float* __restrict__ g, const float* __restrict__ b,
                                 __nv_bfloat16* __restrict__ out, int ldo, int D, float eps,
                                 const int32_t* __restrict__ row_map, int rows_per_map) {
  __shared__ float red[32];
  const int m = blockIdx.x;
  int src = m;
  if (row_map) src = row_map[m / rows_per_map] * rows_per_map + (m % rows_per_map);
  const int c = threadIdx.x * 8;
  float f[8];
  const bool active = c < D;
  if (active) {
    const uint4 w = *reinterpret_cast<const uint4*>(x + (size_t)src * ldx + c);
    unpack8(w, f);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 0.f;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i];
  const float mean = block_sum(s, red) / (float)D;
  float q = 0.f;
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = f[i] - mean; q += d * d; }
  }
  const float var = block_sum(q, red) / (float)D;
  const float rstd = rsqrtf(var + eps);
  if (active) {
    const float4 g0 = *reinterpret_cast<const float4*>(g + c), g1 = *reinterpret_cast<const float4*>(g + c + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(b + c), b1 = *reinterpret_cast<const float4*>(b + c + 4);
    float o[8];
    o[0] = (f[0] - mean) * rstd * g0.x + b0.x; o[1] = (f[1] - mean) * rstd * g0.y + b0.y;
    o[2] = (f[2] - mean) * rstd * g0.z + b0.z; o[3] = (f[3] - mean) * rstd * g0.w + b0.w;
    o[4] = (f[4] - mean) * rstd * g1.x + b1.x; o[5] = (f[5] - mean) * rstd * g1.y + b1.y;
    o[6] = (f[6] - mean) * rstd * g1.z + b1.z; o[7] = (f[7] - mean) * rstd * g1.w + b1.w;
    *reinterpret_cast<uint4*>(out + (size_t)m * ldo + c) = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------
// Fusion (+ optional QFormer.pre_norm) in two kernels: group_mean_kernel (one launch per group of five hidden
// states, as soon as the group is complete) and fuse_tail_kernel (concatenate on channels -> 5120, CLS row
// dropped (vision_encoder.py:68), LayerNorm(5120)).  128 threads per patch row.
// ------------------------------------------------------------------------------------------
struct Mean5Args {
  const __nv_bfloat16* hs[5];   // five consecutive hidden states, each [T*577,1024]
};

// mean of five consecutive hidden states (CLS dropped), rounded to bf16 like the reference's bf16 stack().mean()
// (gating_fusion.py:36-44), written to columns [1024 grp, 1024 grp + 1024) of means[T*576][4096].  Runs as soon as
// the group's last layer is done, so only six hidden states are alive at any time.
__global__ void __launch_bounds__(128)
group_mean_kernel(const Mean5Args a, __nv_bfloat16* __restrict__ means, int grp) {
  const int row = blockIdx.x;  // t*576 + p
  const int t = row / VZ_VIT_PATCHES, pidx = row - t * VZ_VIT_PATCHES;
  const size_t src = ((size_t)t * VZ_VIT_TOKENS + 1 + pidx) * VZ_VIT_WIDTH + threadIdx.x * 8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const uint4 w = *reinterpret_cast<const uint4*>(a.hs[j] + src);
    float f[8];
    unpack8(w, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += f[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] *= 0.2f;
  *reinterpret_cast<uint4*>(means + (size_t)row * (4 * VZ_VIT_WIDTH) + grp * VZ_VIT_WIDTH + threadIdx.x * 8) = pack8(acc);
}

// cat(4 group means, last hidden state) -> [T*576, 5120], with QFormer.pre_norm (LayerNorm(5120),
// multimodal_projector/builder.py:74) applied when gamma != NULL.  128 threads: thread owns channels
// [8*tid, 8*tid+8) of every group.
__global__ void __launch_bounds__(128)
fuse_tail_kernel(const __nv_bfloat16* __restrict__ means, const __nv_bfloat16* __restrict__ last,
                 const float* __restrict__ g, const float* __restrict__ b, __nv_bfloat16* __restrict__ out, float eps) {
  __shared__ float red[32];
  const int row = blockIdx.x;  // t*576 + p
  const int t = row / VZ_VIT_PATCHES, pidx = row - t * VZ_VIT_PATCHES;
  float v[5][8];
#pragma unroll
  for (int grp = 0; grp < 4; ++grp)
    unpack8(*reinterpret_cast<const uint4*>(means + (size_t)row * (4 * VZ_VIT_WIDTH) + grp * VZ_VIT_WIDTH + threadIdx.x * 8), v[grp]);
  unpack8(*reinterpret_cast<const uint4*>(last + ((size_t)t * VZ_VIT_TOKENS + 1 + pidx) * VZ_VIT_WIDTH + threadIdx.x * 8), v[4]);
  __nv_bfloat16* o = out + (size_t)row * VZ_FUSED_WIDTH + threadIdx.x * 8;
  if (g == nullptr) {
#pragma unroll
    for (int grp = 0; grp < 5; ++grp) *reinterpret_cast<uint4*>(o + grp * VZ_VIT_WIDTH) = pack8(v[grp]);
    return;
  }
  float s = 0.f;
#pragma unroll
  for (int grp = 0; grp < 5; ++grp)
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[grp][i];
  const float mean = block_sum(s, red) * (1.0f / VZ_FUSED_WIDTH);
  float q = 0.f;
#pragma unroll
  for (int grp = 0; grp < 5; ++grp)
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[grp][i] - mean; q += d * d; }
  const float var = block_sum(q, red) * (1.0f / VZ_FUSED_WIDTH);
  const float rstd = rsqrtf(var + eps);
#pragma unroll
  for (int grp = 0; grp < 5; ++grp) {
    const int c = grp * VZ_VIT_WIDTH + threadIdx.x * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + c)), g1 = __ldg(reinterpret_cast<const float4*>(g + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c)), b1 = __ldg(reinterpret_cast<const float4*>(b + c + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = (v[grp][i] - mean) * rstd * gg[i] + bb[i];
    *reinterpret_cast<uint4*>(o + grp * VZ_VIT_WIDTH) = pack8(r);
  }
}

// CLS rows of the embedding: emb[t*577] = class_embedding + position_embedding[0]
__global__ void cls_rows_kernel(const __nv_bfloat16* __restrict__ cls, const __nv_bfloat16* __restrict__ pos,
                                __nv_bfloat16* __restrict__ emb) {
  const int t = blockIdx.x;
  const int c = threadIdx.x * 8;
  float a[8], p[8], o[8];
  unpack8(*reinterpret_cast<const uint4*>(cls + c), a);
  unpack8(*reinterpret_cast<const uint4*>(pos + c), p);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = a[i] + p[i];
  *reinterpret_cast<uint4*>(emb + (size_t)t * VZ_VIT_TOKENS * VZ_VIT_WIDTH + c) = pack8(o);
}

// out[m] = in[row_map[m / rows_per] * rows_per + m % rows_per]  (16-byte vectors)
__global__ void gather_rows_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int vec_per_row,
                                   const int32_t* __restrict__ row_map, int rows_per) {
  const int m = blockIdx.x;
  // row_map == NULL broadcasts the first rows_per rows to every group
  const int src = (row_map ? row_map[m / rows_per] * rows_per : 0) + (m % rows_per);
  for (int i = threadIdx.x; i < vec_per_row; i += blockDim.x)
    out[(size_t)m * vec_per_row + i] = in[(size_t)src * vec_per_row + i];
}

// pixel_values [T,3,336,336] (f32 or bf16) -> patches bf16 [T*576, 592], k = c*196 + ky*14 + kx.
// One CTA per (patch row py, tile t): stage 3 x 14 x 336 pixels through shared memory so both the
// global reads (1344/672-byte rows) and the writes (24 x 1184 B contiguous) are coalesced.
template <bool F32>
__global__ void __launch_bounds__(256)
patchify_kernel(const void* __restrict__ pix, __nv_bfloat16* __restrict__ patches) {
  __shared__ __align__(16) __nv_bfloat16 sp[24 * VZ_PATCH_K];
  const int py = blockIdx.x, t = blockIdx.y;
  // zero the K padding (4 elements per patch row)
  for (int i = threadIdx.x; i < 24 * 4; i += blockDim.x) sp[(i >> 2) * VZ_PATCH_K + 588 + (i & 3)] = __float2bfloat16_rn(0.f);
  for (int i = threadIdx.x; i < 3 * 14 * 336; i += blockDim.x) {
    const int x = i % 336;
    const int ky = (i / 336) % 14;
    const int c = i / (336 * 14);
    const size_t gi = (((size_t)t * 3 + c) * 336 + (py * 14 + ky)) * 336 + x;
    __nv_bfloat16 v;
    if (F32) v = __float2bfloat16_rn(reinterpret_cast<const float*>(pix)[gi]);
    else v = reinterpret_cast<const __nv_bfloat16*>(pix)[gi];
    const int px = x / 14, kx = x - px * 14;
    sp[px * VZ_PATCH_K + c * 196 + ky * 14 + kx] = v;
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(patches + ((size_t)t * VZ_VIT_PATCHES + py * 24) * VZ_PATCH_K);
  const uint4* s4 = reinterpret_cast<const uint4*>(sp);
  for (int i = threadIdx.x; i < 24 * VZ_PATCH_K * 2 / 16; i += blockDim.x) dst[i] = s4[i];
}

// row softmax of fp32 scores (pre-scaled by `scale`) -> bf16 probabilities; one warp per row.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int rows, int n, float scale_log2e) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* sr = s + (size_t)row * n;
  float v[24];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < n ? sr[c] : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    v[i] = exp2f((v[i] - mx) * scale_log2e);
    sum += v[i];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __nv_bfloat16* pr = p + (size_t)row * n;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const int c = lane + 32 * i;
    if (c < n) pr[c] = __float2bfloat16_rn(v[i] * inv);
  }
}

// (sum, sum of squares) of every row -> stats[m][0..1]  (one partial per row; the debug GEMM path and
// the first ViT layer use it where no producing GEMM epilogue exists).  One warp per row.
__global__ void __launch_bounds__(256)
row_stats_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int M, int D, float* __restrict__ stats) {
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane * 8; c < D; c += 256) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(x + (size_t)m * ldx + c), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1 += f[i]; s2 = fmaf(f[i], f[i], s2); }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) { stats[(size_t)m * 2] = s1; stats[(size_t)m * 2 + 1] = s2; }
}

}  // namespace

int row_stats_launch(const void* x, int ldx, int M, int D, float* stats, cudaStream_t st) {
  if (D % 8) return VZ_ERR_UNSUPPORTED;
  row_stats_kernel<<<(M + 7) / 8, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, M, D, stats);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

int softmax_rows_launch(const float* s, void* p, int rows, int n, float scale, cudaStream_t st) {
  if (n > 24 * 32) return VZ_ERR_UNSUPPORTED;
  ProfScope prof(VZ_PROF_SOFTMAX, (double)rows * n * 6.0, st);   // f32 scores in, bf16 probabilities out
  softmax_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(s, reinterpret_cast<__nv_bfloat16*>(p), rows, n,
                                                      scale * 1.4426950408889634f);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

int layernorm_launch(const void* x, int ldx, const float* g, const float* b, void* out, int ldo,
                     int M, int D, float eps, const int32_t* row_map, int rows_per_map,
                     cudaStream_t st) {
  if (!x || !g || !b || !out || M <= 0) return VZ_ERR_BAD_ARG;
  if (D % 8 != 0 || D > 8192 || (ldx & 7) || (ldo & 7)) return VZ_ERR_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(out) || !aligned16(g) || !aligned16(b)) return VZ_ERR_BAD_ARG;
  int threads = ((D / 8 + 31) / 32) * 32;
  ProfScope prof(VZ_PROF_LAYERNORM, (double)M * D * 4.0, st);
  layernorm_kernel<<<M, threads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, g, b,
                                          reinterpret_cast<__nv_bfloat16*>(out), ldo, D, eps,
                                          row_map, rows_per_map > 0 ? rows_per_map : 1);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

int group_mean_launch(const void* const* hs5, int T, void* means, int group, cudaStream_t st) {
  Mean5Args a;
  for (int i = 0; i < 5; ++i) a.hs[i] = reinterpret_cast<const __nv_bfloat16*>(hs5[i]);
  // algorithmic bytes: five hidden-state rows of 1024 bf16 read + one mean slice written, per patch
  ProfScope prof(VZ_PROF_FUSE, (double)T * VZ_VIT_PATCHES * 6.0 * VZ_VIT_WIDTH * 2.0, st);
  group_mean_kernel<<<T * VZ_VIT_PATCHES, 128, 0, st>>>(a, reinterpret_cast<__nv_bfloat16*>(means), group);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

int fuse_tail_launch(const void* means, const void* last, int T, const float* g, const float* b, void* out,
                     cudaStream_t st) {
  // algorithmic bytes: 4 mean slices + the last hidden state read, one fused row of 5120 bf16 written, per patch
  ProfScope prof(VZ_PROF_FUSE, (double)T * VZ_VIT_PATCHES * (5.0 * VZ_VIT_WIDTH + VZ_FUSED_WIDTH) * 2.0, st);
  fuse_tail_kernel<<<T * VZ_VIT_PATCHES, 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(means),
                                                       reinterpret_cast<const __nv_bfloat16*>(last), g, b,
                                                       reinterpret_cast<__nv_bfloat16*>(out), 1e-5f);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

int cls_rows_launch(const void* cls, const void* pos, void* emb, int T, cudaStream_t st) {
  cls_rows_kernel<<<T, VZ_VIT_WIDTH / 8, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(cls),
                                                  reinterpret_cast<const __nv_bfloat16*>(pos),
                                                  reinterpret_cast<__nv_bfloat16*>(emb));
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

int gather_rows_launch(const void* in, void* out, int M, int row_bytes, const int32_t* row_map,
                       int rows_per, cudaStream_t st) {
  if (row_bytes % 16) return VZ_ERR_UNSUPPORTED;
  gather_rows_kernel<<<M, 256, 0, st>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out),
                                        row_bytes / 16, row_map, rows_per);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

}  // namespace vz

extern "C" int vz_layernorm_bf16(const void* x, int ldx, const float* gamma, const float* beta,
                                 void* out, int ldo, int M, int D, float eps, void* stream) {
  return vz::layernorm_launch(x, ldx, gamma, beta, out, ldo, M, D, eps, nullptr, 1,
                              reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int vz_patchify(const void* pixel_values, int src_is_f32, int T, void* patches,
                           void* stream) {
  if (!pixel_values || !patches || T <= 0) return VZ_ERR_BAD_ARG;
  if (!vz::aligned16(patches)) return VZ_ERR_BAD_ARG;
  dim3 grid(24, T);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (src_is_f32)
    vz::patchify_kernel<true><<<grid, 256, 0, st>>>(pixel_values, reinterpret_cast<__nv_bfloat16*>(patches));
  else
    vz::patchify_kernel<false><<<grid, 256, 0, st>>>(pixel_values, reinterpret_cast<__nv_bfloat16*>(patches));
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}
