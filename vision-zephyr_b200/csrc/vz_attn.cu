// vz_attn.cu -- the Q-Former's small attention kernel (the CLIP ViT attention is vz_attn_tc.cu; its first,
// mma.sync implementation lives in testonly/vz_attn_legacy.cu as a cross-check and is not part of libvz_b200.so).
//   qattn32_kernel  : the Q-Former's learned-query attention: 32 query rows x head_dim 512 against
//                     up to three key/value segments (multimodal_projector/builder.py:34-40):
//                       block-0 self-attention  : 32 query rows + the sample's text rows + the
//                                                  collapsed zero-padding row with multiplicity
//                       blocks 1..7 self-attn   : the tile's own 32 rows
//                       cross-attention         : the tile's 576 projected patch rows
// Flash-style (online softmax, fp32 statistics) on mma.sync m16n8k16 bf16 tiles with cp.async double-buffered,
// XOR-swizzled shared memory: a 32 x 32..2080 x 512 problem per (head, sample) is far below what a tcgen05
// tile (M >= 64 rows, TMEM round trips) pays off for; < 1 % of the step.
#include "vz_common.cuh"

namespace vz {
namespace {

constexpr float kLog2e = 1.4426950408889634f;

// ==========================================================================================
// Q-Former 32-query attention, head_dim 512
// ==========================================================================================
constexpr int QA_Q = 32, QA_D = 512, QA_BK = 32, QA_THREADS = 256;
constexpr int QA_ROW_BYTES = QA_D * 2;                      // 1024
constexpr int QA_TILE_BYTES = QA_BK * QA_ROW_BYTES;         // 32 KB
constexpr int QA_SMEM = QA_Q * QA_ROW_BYTES + 4 * QA_TILE_BYTES + QA_Q * QA_BK * 4 +
                        QA_Q * QA_BK * 2 + 4 * QA_Q * 4 + QA_BK * 4;

struct QAttnArgs {
  const __nv_bfloat16* q;  int q_rs;  int q_zrows;   // q row = z*q_zrows + r
  const __nv_bfloat16* k[3];
  const __nv_bfloat16* v[3];
  int rs[3];        // row stride (elements)
  int zrows[3];     // first row of segment for z = z*zrows (when not offset-driven)
  int count[3];     // fixed counts (seg 1 ignores it when off1 != NULL; seg 2 is the pad row)
  const int32_t* off1;  // [Z+1] row offsets of segment 1 (text rows) or NULL
  int L;            // batch-global text length; pad multiplicity = L - count1 (seg 2 only)
  int use_pad;      // segment 2 is the collapsed zero-padding key
  __nv_bfloat16* out; int ldo;
  float scale;
};

// element (row, col) of a [rows][512] bf16 tile, 16-byte chunks XOR-swizzled by row&7
__device__ __forceinline__ uint32_t sw512(int row, int col) {
  return (uint32_t)(row * QA_ROW_BYTES + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1)));
}

__global__ void __launch_bounds__(QA_THREADS, 1) qattn32_kernel(const QAttnArgs a) {
  extern __shared__ __align__(128) uint8_t qa_smem[];
  uint8_t* sQ = qa_smem;
  uint8_t* sKV = sQ + QA_Q * QA_ROW_BYTES;  // [2 stages][K,V][32 KB]
  float* sS = reinterpret_cast<float*>(sKV + 4 * QA_TILE_BYTES);         // [32][32]
  __nv_bfloat16* sP = reinterpret_cast<__nv_bfloat16*>(sS + QA_Q * QA_BK);  // [32][32]
  float* sM = reinterpret_cast<float*>(sP + QA_Q * QA_BK);               // [32] running max
  float* sL = sM + QA_Q;                                                 // [32] running sum
  float* sAlpha = sL + QA_Q;                                             // [32]
  float* sMult = sAlpha + QA_Q * 2;                                      // [32] per-key weight

  const int h = blockIdx.x, z = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;

  // ---- segment bookkeeping ---------------------------------------------------------------
  int cnt[3], row0[3];
  cnt[0] = a.count[0]; row0[0] = z * a.zrows[0];
  if (a.off1) { row0[1] = a.off1[z]; cnt[1] = a.off1[z + 1] - row0[1]; }
  else { cnt[1] = a.count[1]; row0[1] = z * a.zrows[1]; }
  float pad_mult = 1.f;
  cnt[2] = a.count[2]; row0[2] = z * a.zrows[2];
  if (a.use_pad) {
    const int m = a.L - cnt[1];
    cnt[2] = m > 0 ? 1 : 0;
    pad_mult = (float)m;
    row0[2] = 0;
  }
  const int total = cnt[0] + cnt[1] + cnt[2];
  const int nchunks = (total + QA_BK - 1) / QA_BK;
  const int hoff = h * QA_D;

  auto load_chunk = [&](int j, int stage) {
    uint8_t* dK = sKV + stage * 2 * QA_TILE_BYTES;
    uint8_t* dV = dK + QA_TILE_BYTES;
#pragma unroll
    for (int i = 0; i < (QA_BK * 64) / QA_THREADS; ++i) {
      const int idx = tid + i * QA_THREADS;
      const int r = idx >> 6, c = idx & 63;
      int kidx = j * QA_BK + r;
      const bool ok = kidx < total;
      if (!ok) kidx = total - 1;
      int seg = 0;
      if (kidx >= cnt[0]) { kidx -= cnt[0]; seg = 1; if (kidx >= cnt[1]) { kidx -= cnt[1]; seg = 2; } }
      const size_t off = (size_t)(row0[seg] + kidx) * a.rs[seg] + hoff + c * 8;
      cp_async_16(dK + sw512(r, c * 8), a.k[seg] + off, ok);
      cp_async_16(dV + sw512(r, c * 8), a.v[seg] + off, ok);
    }
  };

  // Q tile + first chunk
#pragma unroll
  for (int i = 0; i < (QA_Q * 64) / QA_THREADS; ++i) {
    const int idx = tid + i * QA_THREADS;
    const int r = idx >> 6, c = idx & 63;
    cp_async_16(sQ + sw512(r, c * 8), a.q + (size_t)(z * a.q_zrows + r) * a.q_rs + hoff + c * 8, true);
  }
  load_chunk(0, 0);
  cp_async_commit();
  if (tid < QA_Q) { sM[tid] = -INFINITY; sL[tid] = 0.f; }

  float o[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[mt][dt][i] = 0.f;

  const float sl2 = a.scale * kLog2e;
  const int s_mt = warp >> 2, s_nt = warp & 3;  // this warp's 16x8 tile of S

  for (int j = 0; j < nchunks; ++j) {
    const int stage = j & 1;
    if (j + 1 < nchunks) {
      load_chunk(j + 1, stage ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    // per-key multiplicity for this chunk (pad key counts L - S' times)
    if (tid < QA_BK) {
      const int kidx = j * QA_BK + tid;
      sMult[tid] = (a.use_pad && cnt[2] > 0 && kidx == total - 1) ? pad_mult : 1.f;
    }
    __syncthreads();
    const uint32_t qb = smem_u32(sQ);
    const uint32_t kb = smem_u32(sKV + stage * 2 * QA_TILE_BYTES);
    const uint32_t vb = kb + QA_TILE_BYTES;

    // ---- S tile (16 q x 8 keys) over K = 512 -------------------------------------------
    float sacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int kp = 0; kp < QA_D / 32; ++kp) {
      uint32_t af0[4], af1[4], bfr[4];
      ldmatrix_x4(af0, qb + sw512(s_mt * 16 + (lane & 15), kp * 32 + (lane >> 4) * 8));
      ldmatrix_x4(af1, qb + sw512(s_mt * 16 + (lane & 15), kp * 32 + 16 + (lane >> 4) * 8));
      ldmatrix_x4(bfr, kb + sw512(s_nt * 8 + (lane & 7), kp * 32 + (lane >> 3) * 8));
      const uint32_t b0[2] = {bfr[0], bfr[1]}, b1[2] = {bfr[2], bfr[3]};
      mma_bf16_16816(sacc, af0, b0);
      mma_bf16_16816(sacc, af1, b1);
    }
    {
      const int r = s_mt * 16 + g, c = s_nt * 8 + tq * 2;
      const int kc = j * QA_BK + c;
      sS[r * QA_BK + c] = (kc < total) ? sacc[0] : -INFINITY;
      sS[r * QA_BK + c + 1] = (kc + 1 < total) ? sacc[1] : -INFINITY;
      sS[(r + 8) * QA_BK + c] = (kc < total) ? sacc[2] : -INFINITY;
      sS[(r + 8) * QA_BK + c + 1] = (kc + 1 < total) ? sacc[3] : -INFINITY;
    }
    __syncthreads();
    // ---- online softmax: warp w owns rows 4w..4w+3, lane = key -------------------------
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int r = warp * 4 + rr;
      const float sv = sS[r * QA_BK + lane];
      const float mx = warp_max(sv);
      const float m_old = sM[r];
      const float m_new = fmaxf(m_old, mx);
      const float p = exp2f((sv - m_new) * sl2) * sMult[lane];
      const float sum = warp_sum(p);
      sP[r * QA_BK + lane] = __float2bfloat16_rn(p);
      if (lane == 0) {
        const float al = exp2f((m_old - m_new) * sl2);
        sAlpha[r] = al;
        sM[r] = m_new;
        sL[r] = sL[r] * al + sum;
      }
    }
    __syncthreads();
    // ---- O[32 x 64 dims of this warp] = alpha*O + P V ------------------------------------
    {
      float al[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) { al[mt][0] = sAlpha[mt * 16 + g]; al[mt][1] = sAlpha[mt * 16 + g + 8]; }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          o[mt][dt][0] *= al[mt][0]; o[mt][dt][1] *= al[mt][0];
          o[mt][dt][2] *= al[mt][1]; o[mt][dt][3] *= al[mt][1];
        }
      const uint32_t pb = smem_u32(sP);
      uint32_t pf[2][2][4];  // [mt][kk]
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
          ldmatrix_x4(pf[mt][kk], pb + (uint32_t)(((mt * 16 + (lane & 15)) * QA_BK + kk * 16 + (lane >> 4) * 8) * 2));
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t bfr[4];
          const int jm = lane >> 3;
          ldmatrix_x4_trans(bfr, vb + sw512(kk * 16 + (jm & 1) * 8 + (lane & 7),
                                            warp * 64 + (dp * 2 + (jm >> 1)) * 8));
          const uint32_t b0[2] = {bfr[0], bfr[1]}, b1[2] = {bfr[2], bfr[3]};
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma_bf16_16816(o[mt][dp * 2], pf[mt][kk], b0);
            mma_bf16_16816(o[mt][dp * 2 + 1], pf[mt][kk], b1);
          }
        }
      }
    }
    __syncthreads();  // stage buffers, sS/sP/sAlpha free for the next chunk
  }
  // ---- finalise --------------------------------------------------------------------------
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    const float i0 = 1.0f / sL[r0], i1 = 1.0f / sL[r1];
    __nv_bfloat16* o0 = a.out + (size_t)(z * QA_Q + r0) * a.ldo + hoff + warp * 64;
    __nv_bfloat16* o1 = a.out + (size_t)(z * QA_Q + r1) * a.ldo + hoff + warp * 64;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int col = dt * 8 + tq * 2;
      *reinterpret_cast<uint32_t*>(o0 + col) = pack_bf16x2(o[mt][dt][0] * i0, o[mt][dt][1] * i0);
      *reinterpret_cast<uint32_t*>(o1 + col) = pack_bf16x2(o[mt][dt][2] * i1, o[mt][dt][3] * i1);
    }
  }
}

}  // namespace

// mode 0: block-0 self-attention with text; 1: self-attention over own 32 rows; 2: cross-attention
int qattn_launch(int mode, const void* q, int q_rs, int q_zrows, const void* k0, const void* v0,
                 int rs0, int zrows0, int count0, const void* k1, const void* v1, int rs1,
                 const int32_t* off1, const void* kpad, const void* vpad, int L, void* out, int ldo,
                 int Z, cudaStream_t st) {
  QAttnArgs a;
  a.q = reinterpret_cast<const __nv_bfloat16*>(q); a.q_rs = q_rs; a.q_zrows = q_zrows;
  a.k[0] = reinterpret_cast<const __nv_bfloat16*>(k0); a.v[0] = reinterpret_cast<const __nv_bfloat16*>(v0);
  a.rs[0] = rs0; a.zrows[0] = zrows0; a.count[0] = count0;
  a.k[1] = reinterpret_cast<const __nv_bfloat16*>(k1 ? k1 : k0); a.v[1] = reinterpret_cast<const __nv_bfloat16*>(v1 ? v1 : v0);
  a.rs[1] = rs1; a.zrows[1] = 0; a.count[1] = 0; a.off1 = off1;
  a.k[2] = reinterpret_cast<const __nv_bfloat16*>(kpad ? kpad : k0); a.v[2] = reinterpret_cast<const __nv_bfloat16*>(vpad ? vpad : v0);
  a.rs[2] = rs1; a.zrows[2] = 0; a.count[2] = 0;
  a.L = L; a.use_pad = (mode == 0) ? 1 : 0;
  a.out = reinterpret_cast<__nv_bfloat16*>(out); a.ldo = ldo;
  a.scale = 0.044194173824159216f;  // 1/sqrt(512)
  VZ_ENSURE_DYN_SMEM(qattn32_kernel, QA_SMEM);
  dim3 grid(VZ_QF_HEADS, Z);
  ProfScope prof(VZ_PROF_QATTN, 0.0, st);
  qattn32_kernel<<<grid, QA_THREADS, QA_SMEM, st>>>(a);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

}  // namespace vz
