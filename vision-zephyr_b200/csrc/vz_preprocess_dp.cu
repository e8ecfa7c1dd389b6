// vz_preprocess_dp.cu -- the resampling form of subsystem (1): visual-prompt blend + canvas padding +
// Pillow LANCZOS / BICUBIC (horizontal pass, u8 intermediate, vertical pass) + CLIP normalise + patchify,
// with the tap loops on the 4-way byte dot product (dp4a) instead of one multiply-add per tap.
//
// Same arithmetic as vz_preprocess.cu (bit-exact vs Pillow's Resample.c: 22-bit fixed-point coefficients,
// int32 accumulator started at 1 << 21, >> 22, clip to u8 -- multi_scale_process.py:86-89,171-174 and
// mm_utils.py:59-63 are the reference call sites), reorganised so that four TAPS of one channel sit in one
// 32-bit register:
//   * a 23-bit signed coefficient c is split into byte limbs  c = c0 + 256 c1 + 65536 c2  (c0, c1 unsigned,
//     c2 signed).  sum_k c_k p_k = sum c0 p + 256 sum c1 p + 65536 sum c2 p, three dp4a per four taps;
//     the recombination wraps modulo 2^32 exactly like Pillow's own int32 accumulator, whose final value fits.
//   * horizontal pass: source rows are de-interleaved into R / G / B byte planes in shared memory (blend and
//     canvas padding happen on the way), every window is aligned DOWN to a multiple of four pixels and the
//     coefficient bytes are shifted to match (zero padded), so a thread's taps are aligned 32-bit words of a
//     plane; its 3 x G coefficient words stay in registers for all rows of the CTA.
//   * the u8 intermediate is stored as I[row / 4][x] = uint4 (R word, G word, B word, 0): one word = FOUR
//     VERTICALLY consecutive pixels of one channel, which is what the vertical pass needs for its own dp4a
//     (windows aligned down to four rows).  A horizontal-pass thread computes four rows of one column and
//     writes one uint4, coalesced in x; the vertical pass stages a band's window with 128-bit loads.
//   * coefficient limbs come from a host-built table (anyres.resample_table_dp), one uint4 per (output, group).
// The identity form (no resampling: the fixed 336 x 336 inputs of config 2) is pre_identity_kernel below.
#include "vz_common.cuh"

#include <stdlib.h>

namespace vz {
namespace {

constexpr int TILE = 336;
constexpr int BAND = 14;
constexpr int PREC = 22;
constexpr int MAX_PRIMS = 32;
constexpr int HX = 128;        // output columns of one horizontal-pass CTA
constexpr int HT = 256;        // its threads: 128 columns x 2 row quads
constexpr int VT = 352;        // vertical-pass threads: thread x < 336 owns output column x of the band

__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// a: four unsigned bytes (pixels), b: four SIGNED bytes (top coefficient limb)
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// limbs -> Pillow's clip8((ss + 2^21) >> 22); the sums wrap modulo 2^32, the true total fits in int32.
// Two multiply-adds recombine the limbs, one shift, one saturating convert clamps to [0, 255].
__device__ __forceinline__ uint32_t finish8(uint32_t a0, uint32_t a1, int a2) {
  const uint32_t t = (uint32_t)a2 * 65536u + (a1 * 256u + (a0 + (1u << (PREC - 1))));
  const int v = (int)t >> PREC;
  uint32_t d;
  asm("cvt.sat.u8.s32 %0, %1;" : "=r"(d) : "r"(v));
  return d;
}
// byte k of `word` := the low byte of v (k is a constant after unrolling: one PRMT)
__device__ __forceinline__ uint32_t put_byte_rt(uint32_t word, uint32_t v, int k) {
  return __byte_perm(word, v, k == 0 ? 0x3214 : k == 1 ? 0x3240 : k == 2 ? 0x3410 : 0x4210);
}

// Pillow AlphaComposite.c with an opaque destination
__device__ __forceinline__ int blend_over(int dst, int src, int alpha) {
  if (alpha == 0) return dst;
  const uint32_t t = (uint32_t)src * (uint32_t)(alpha * 128) + (uint32_t)dst * (uint32_t)((255 - alpha) * 128) + (0x80u << 7);
  return (int)((((t >> 8) + t) >> 8) >> 7);
}
// PIL ImageDraw.rectangle(outline, width) coverage (see vz_preprocess.cu)
__device__ __forceinline__ bool rect_covers(const vz_prim& p, int x, int y) {
  const int w = p.width;
  if (w <= 0) return false;
  const bool hl = (x >= p.x0 && x <= p.x1) && ((y >= p.y0 && y < p.y0 + w) || (y <= p.y1 && y > p.y1 - w));
  const int va = p.y0 + w, vb = p.y1 - w + 1;
  const bool in_v = (vb >= va) ? (y >= va && y < vb) : (y <= va && y > vb);
  const bool vl = in_v && ((x <= p.x1 && x > p.x1 - w) || (x >= p.x0 && x < p.x0 + w));
  return hl || vl;
}

// Host-built dp4a axis table (anyres.resample_table_dp), 32-bit words, every section 16-byte aligned:
//   [0] = G (groups of four taps per output index), [1] = n (output size), [2..3] = 0
//   abase[n4]  aligned start of each window (xmin & ~3), n4 = n rounded up to 4
//   ngrp[n4]   groups output i really needs ( ceil((count + (xmin & 3)) / 4) )
//   coef[n][G] uint4 = (c0 bytes, c1 bytes, c2 bytes, 0) of taps abase + 4g .. abase + 4g + 3 (zero padded)
struct DpTable {
  int G, n;
  const int32_t* abase;
  const int32_t* ngrp;
  const uint4* coef;
};
__device__ __forceinline__ DpTable dp_table(const int32_t* tables, int off) {
  const int32_t* t = tables + off;
  DpTable d;
  d.G = t[0]; d.n = t[1];
  const int n4 = (d.n + 3) & ~3;
  d.abase = t + 4;
  d.ngrp = d.abase + n4;
  d.coef = reinterpret_cast<const uint4*>(d.ngrp + n4);
  return d;
}

// The horizontal filter of one thread: column x, row quads q, q + 2, ...; GG groups of four taps (compile time).
template <int GG>
__device__ __forceinline__ void h_filter(const uint32_t* __restrict__ pl, const uint4* __restrict__ cop, int G, bool valid,
                                         int RB, int plane_words, int nrows, int q, uint4* __restrict__ inter, int rq0,
                                         int out_w, int x) {
  uint4 co[GG];
#pragma unroll
  for (int g = 0; g < GG; ++g) co[g] = (valid && g < G) ? __ldg(cop + g) : make_uint4(0u, 0u, 0u, 0u);
  const int nrq = (nrows + 3) >> 2;
  for (int rq = q; rq < nrq; rq += 2) {
    uint32_t o[3] = {0u, 0u, 0u};
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int r = rq * 4 + rr;
      if (r < nrows) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const uint32_t* pw = pl + (c * RB + r) * plane_words;
          uint32_t a0 = 0u, a1 = 0u;
          int a2 = 0;
#pragma unroll
          for (int g = 0; g < GG; ++g) {
            const uint32_t p = pw[g];
            a0 = dp4a_uu(p, co[g].x, a0);
            a1 = dp4a_uu(p, co[g].y, a1);
            a2 = dp4a_us(p, co[g].z, a2);
          }
          o[c] = put_byte_rt(o[c], finish8(a0, a1, a2), rr);
        }
      }
    }
    if (valid) inter[(size_t)(rq0 + rq) * out_w + x] = make_uint4(o[0], o[1], o[2], 0u);
  }
}

struct HArgs {
  const vz_image_desc* images;
  const vz_prim* prims;
  const vz_hview_desc* hviews;
  const int32_t* tables;
  uint32_t* scratch;
  int plane_words;    // 32-bit words per plane row
  int rows_per_cta;   // 8 or 16
};

// ------------------------------------------------------------------------------------------------
// Horizontal pass.  CTA = rows_per_cta canvas rows x 128 output columns of one (image, table) view.
// Nothing is staged: the de-interleave step reads the interleaved source straight from global memory
// (four aligned 32-bit loads per four pixels, re-aligned by funnel shifts), so a CTA needs only its three
// byte planes in shared memory and many CTAs share an SM.
// ------------------------------------------------------------------------------------------------
template <int GMAX>
__global__ void __launch_bounds__(HT) pre_h_dp_kernel(const HArgs a) {
  extern __shared__ __align__(16) uint8_t dp_smem[];
  const vz_hview_desc hv = a.hviews[blockIdx.z];
  const int RB = a.rows_per_cta;
  const int x0 = blockIdx.x * HX, y0 = blockIdx.y * RB, tid = threadIdx.x;
  if (x0 >= hv.out_w || y0 >= hv.rows) return;
  const vz_image_desc im = a.images[hv.image];
  const int nrows = min(RB, hv.rows - y0);
  const DpTable th = dp_table(a.tables, hv.tab_h_dp);
  const int G = th.G;                                                 // <= GMAX (checked by the launcher)
  uint32_t* s_pl = reinterpret_cast<uint32_t*>(dp_smem);             // [3][RB][plane_words]
  __shared__ vz_prim s_prims[MAX_PRIMS];
  const int n_prims = im.prim_count < MAX_PRIMS ? im.prim_count : MAX_PRIMS;
  for (int i = tid; i < n_prims; i += HT) s_prims[i] = a.prims[im.prim_begin + i];
  if (n_prims != 0) __syncthreads();   // (uniform) the blend below reads the instance list

  const int xl = tid & (HX - 1), q = tid >> 7;
  const int x = x0 + xl;
  const bool valid = x < hv.out_w;
  const int xlast = min(x0 + HX - 1, hv.out_w - 1);
  const int base = th.abase[x0];                                      // plane byte 0 = canvas column `base` (multiple of 4)
  const int nvec = ((th.abase[xlast] - base) >> 2) + th.ngrp[xlast];  // plane words that carry data
  const int wb = valid ? (th.abase[x] - base) >> 2 : 0;   // this thread's first plane word

  // ---- de-interleave (+ canvas padding, + visual prompts) into byte planes: one warp per row, four pixels per lane ----
  const uint32_t bgw = im.bg & 0xffffffu;
  const int warp = tid >> 5, lane = tid & 31;
  for (int r = warp; r < nrows; r += HT / 32) {
    const int yr = y0 + r - im.pad_y;
    const bool row_real = yr >= 0 && yr < im.H;
    const uint8_t* rowp = im.src + (size_t)(row_real ? yr : 0) * im.W * 3;
    for (int v = lane; v < nvec; v += 32) {
      const int xr0 = base + 4 * v - im.pad_x;
      uint32_t R, Gc, B;
      if (row_real && xr0 >= 0 && xr0 + 3 < im.W && n_prims == 0) {
        const uint8_t* p = rowp + (size_t)xr0 * 3;                       // 12 bytes from here, any alignment
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3) * 8u;
        const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
        const uint32_t w3 = sh ? __ldg(wp + 3) : 0u;                    // holds bytes of these pixels iff p is unaligned
        const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh), a2 = __funnelshift_r(w2, w3, sh);
        // a0 = R0 G0 B0 R1 | a1 = G1 B1 R2 G2 | a2 = B2 R3 G3 B3   (byte 0 first)
        R = __byte_perm(__byte_perm(a0, a1, 0x0630), a2, 0x5210);
        Gc = __byte_perm(__byte_perm(a0, a1, 0x0741), a2, 0x6210);
        B = __byte_perm(__byte_perm(a0, a1, 0x0052), a2, 0x7410);
      } else {
        R = Gc = B = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int xr = xr0 + i;
          int rr, gg, bb;
          if (row_real && xr >= 0 && xr < im.W) {
            const uint8_t* p = rowp + (size_t)xr * 3;
            rr = __ldg(p); gg = __ldg(p + 1); bb = __ldg(p + 2);
            for (int pi = 0; pi < n_prims; ++pi) {
              const vz_prim& pr = s_prims[pi];
              uint32_t ov;
              if (pr.type == VZ_PRIM_LAYER) {
                ov = __ldg(reinterpret_cast<const uint32_t*>(im.layers) + ((size_t)pr.layer * im.H + yr) * im.W + xr);
              } else {
                if (!rect_covers(pr, xr, yr)) continue;
                ov = pr.rgba;
              }
              const int al = (int)(ov >> 24);
              rr = blend_over(rr, (int)(ov & 0xff), al);
              gg = blend_over(gg, (int)((ov >> 8) & 0xff), al);
              bb = blend_over(bb, (int)((ov >> 16) & 0xff), al);
            }
          } else {   // canvas padding (expand2square, mm_utils.py:16-35)
            rr = (int)(bgw & 0xff); gg = (int)((bgw >> 8) & 0xff); bb = (int)((bgw >> 16) & 0xff);
          }
          R |= (uint32_t)rr << (8 * i); Gc |= (uint32_t)gg << (8 * i); B |= (uint32_t)bb << (8 * i);
        }
      }
      s_pl[(0 * RB + r) * a.plane_words + v] = R;
      s_pl[(1 * RB + r) * a.plane_words + v] = Gc;
      s_pl[(2 * RB + r) * a.plane_words + v] = B;
    }
  }
  __syncthreads();

  // ---- filter: thread = (column x, four consecutive rows), one intermediate uint4 (R, G, B words).  The loop
  // is instantiated for the view's group count rounded up within the kernel's class (a 1.5x downscale next to
  // a 3x one in the same launch should not pay for the longer filter) ----
  uint4* inter = reinterpret_cast<uint4*>(a.scratch + hv.offset);
  const uint4* cop = th.coef + (size_t)(valid ? x : 0) * G;
  const uint32_t* pl = s_pl + wb;
  if (GMAX > 4 && G <= (GMAX + 1) / 2 + (GMAX > 8 ? 2 : 1))
    h_filter<(GMAX + 1) / 2 + (GMAX > 8 ? 2 : 1)>(pl, cop, G, valid, RB, a.plane_words, nrows, q, inter, (y0 >> 2), hv.out_w, x);
  else
    h_filter<GMAX>(pl, cop, G, valid, RB, a.plane_words, nrows, q, inter, (y0 >> 2), hv.out_w, x);
}

struct VArgs {
  const vz_hview_desc* hviews;
  const vz_tile_desc* tiles;
  const int32_t* tables;
  const float* lut;
  const uint32_t* scratch;
  void* out;
  int out_mode;
  int gmax;           // largest G of any vertical table of the plan (sizes the coefficient slice)
  int win_groups;     // most row groups any band window spans (sizes the staged window)
};

// ------------------------------------------------------------------------------------------------
// Vertical pass + LUT + im2col of one 14-row band of one tile.  The band's window of the intermediate
// (win x 336 uint4, win = row groups the 14 rows touch, 9 for a 1.5x downscale) is staged ONCE in shared
// memory with coalesced 128-bit loads; thread x then runs 9 dp4a per (row, group) from shared memory, keeps
// the 14 x 3 result bytes in registers, and the staging area is reused for the im2col transpose.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VT) pre_v_dp_kernel(const VArgs a) {
  extern __shared__ __align__(16) uint8_t dp_smem[];
  const int band = blockIdx.x, t = blockIdx.y, tid = threadIdx.x;
  const vz_tile_desc td = a.tiles[t];
  const vz_hview_desc hv = a.hviews[td.hview];
  const DpTable tv = dp_table(a.tables, td.tab_v_dp);
  const int G = tv.G;
  const int stage_bytes = (a.out_mode == VZ_OUT_CHW_F32) ? 3 * BAND * TILE * 4 : 24 * VZ_PATCH_K * 2;
  const int win_bytes = a.win_groups * TILE * 16;
  uint8_t* s_stage = dp_smem;                                              // aliases the window (used after it)
  uint4* s_win = reinterpret_cast<uint4*>(dp_smem);                       // [win][336]
  uint4* s_vc = reinterpret_cast<uint4*>(dp_smem + (stage_bytes > win_bytes ? stage_bytes : win_bytes));   // [BAND][gmax]
  int32_t* s_vj0 = reinterpret_cast<int32_t*>(s_vc + BAND * a.gmax);
  int32_t* s_vng = s_vj0 + 16;
  const int ry0 = td.tile_y + band * BAND - td.off_y;   // resized-image row of band row 0
  if (tid < BAND) {
    const int ry = ry0 + tid;
    const bool ok = ry >= 0 && ry < td.out_h;
    s_vj0[tid] = ok ? tv.abase[ry] >> 2 : 0;
    s_vng[tid] = ok ? tv.ngrp[ry] : 0;
  }
  for (int i = tid; i < BAND * G; i += VT) {
    const int y = i / G, g = i - y * G;
    const int ry = ry0 + y;
    s_vc[y * a.gmax + g] = (ry >= 0 && ry < td.out_h) ? __ldg(tv.coef + (size_t)ry * G + g) : make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  int jlo = 0x7fffffff, jhi = 0;
#pragma unroll
  for (int y = 0; y < BAND; ++y) {
    const int ng = s_vng[y], j0 = s_vj0[y];
    if (ng > 0) { jlo = min(jlo, j0); jhi = max(jhi, j0 + ng); }
  }
  const int nwin = jhi > jlo ? jhi - jlo : 0;            // <= a.win_groups (host-computed bound)
  const uint4* inter = reinterpret_cast<const uint4*>(a.scratch + hv.offset);
  for (int i = tid; i < nwin * TILE; i += VT) {
    const int g = i / TILE, xx = i - g * TILE;
    const int rxx = td.tile_x + xx - td.off_x;
    s_win[i] = (rxx >= 0 && rxx < td.out_w) ? __ldg(inter + (size_t)(jlo + g) * hv.out_w + rxx) : make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const int x = tid < TILE ? tid : 0;
  const int rx = td.tile_x + x - td.off_x;
  const bool col_ok = tid < TILE && rx >= 0 && rx < td.out_w;
  uint32_t vr[4] = {0u, 0u, 0u, 0u}, vg[4] = {0u, 0u, 0u, 0u}, vb[4] = {0u, 0u, 0u, 0u};   // 14 result bytes per channel
#pragma unroll
  for (int y = 0; y < BAND; ++y) {
    const int ng = col_ok ? s_vng[y] : 0;
    const uint4* win = s_win + (size_t)(s_vj0[y] - jlo) * TILE + x;
    const uint4* cw = s_vc + y * a.gmax;
    uint32_t r0 = 0u, r1 = 0u, g0 = 0u, g1 = 0u, b0 = 0u, b1 = 0u;
    int r2 = 0, g2 = 0, b2 = 0;
    for (int g = 0; g < ng; ++g) {
      const uint4 c = cw[g];
      const uint4 p = win[(size_t)g * TILE];
      r0 = dp4a_uu(p.x, c.x, r0); r1 = dp4a_uu(p.x, c.y, r1); r2 = dp4a_us(p.x, c.z, r2);
      g0 = dp4a_uu(p.y, c.x, g0); g1 = dp4a_uu(p.y, c.y, g1); g2 = dp4a_us(p.y, c.z, g2);
      b0 = dp4a_uu(p.z, c.x, b0); b1 = dp4a_uu(p.z, c.y, b1); b2 = dp4a_us(p.z, c.z, b2);
    }
    if (ng > 0) {
      vr[y >> 2] |= finish8(r0, r1, r2) << (8 * (y & 3));
      vg[y >> 2] |= finish8(g0, g1, g2) << (8 * (y & 3));
      vb[y >> 2] |= finish8(b0, b1, b2) << (8 * (y & 3));
    }
  }
  __syncthreads();   // everybody is done with the window: the staging area takes its place
  if (a.out_mode == VZ_OUT_PATCHES_BF16) {
    __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(s_stage);
    for (int i = tid; i < 24 * 4; i += VT) sp[(i >> 2) * VZ_PATCH_K + 588 + (i & 3)] = __float2bfloat16_rn(0.f);
    if (tid < TILE) {
      const int px = x / 14, kx = x - px * 14;
#pragma unroll
      for (int y = 0; y < BAND; ++y) {
        __nv_bfloat16* q = sp + px * VZ_PATCH_K + y * 14 + kx;
        q[0] = __float2bfloat16_rn(a.lut[(vr[y >> 2] >> (8 * (y & 3))) & 0xff]);
        q[196] = __float2bfloat16_rn(a.lut[256 + ((vg[y >> 2] >> (8 * (y & 3))) & 0xff)]);
        q[392] = __float2bfloat16_rn(a.lut[512 + ((vb[y >> 2] >> (8 * (y & 3))) & 0xff)]);
      }
    }
  } else if (tid < TILE) {
    float* sf = reinterpret_cast<float*>(s_stage) + x;   // [3][BAND][336]
#pragma unroll
    for (int y = 0; y < BAND; ++y) {
      sf[y * TILE] = a.lut[(vr[y >> 2] >> (8 * (y & 3))) & 0xff];
      sf[(BAND + y) * TILE] = a.lut[256 + ((vg[y >> 2] >> (8 * (y & 3))) & 0xff)];
      sf[(2 * BAND + y) * TILE] = a.lut[512 + ((vb[y >> 2] >> (8 * (y & 3))) & 0xff)];
    }
  }
  __syncthreads();
  if (a.out_mode == VZ_OUT_PATCHES_BF16) {
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) +
                                          ((size_t)t * VZ_VIT_PATCHES + band * 24) * VZ_PATCH_K);
    const uint4* s4 = reinterpret_cast<const uint4*>(s_stage);
    for (int i = tid; i < 24 * VZ_PATCH_K * 2 / 16; i += VT) dst[i] = s4[i];
  } else {
    const float* sf = reinterpret_cast<const float*>(s_stage);
    float* o = reinterpret_cast<float*>(a.out);
    for (int i = tid; i < 3 * BAND * TILE / 4; i += VT) {
      const int e = i * 4;
      const int c = e / (BAND * TILE), rem = e - c * BAND * TILE;
      const int y = rem / TILE, xx = rem - y * TILE;
      *reinterpret_cast<float4*>(o + (((size_t)t * 3 + c) * TILE + band * BAND + y) * TILE + xx) =
          *reinterpret_cast<const float4*>(sf + e);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Identity form (BASELINE config 2: images that already are 336 x 336, one tile each): blend the visual
// prompts, normalise, patchify -- no resampling, so nothing is staged: a thread owns FOUR consecutive pixels
// of a row (three aligned 32-bit source loads, one 128-bit load per overlay layer), the 768-entry LUT sits
// in shared memory already converted to the output type, and the band leaves through 128-bit stores.
// ------------------------------------------------------------------------------------------------
constexpr int IT = 256;
constexpr int QUADS = BAND * (TILE / 4);   // 1176 four-pixel groups per band

struct IArgs {
  const vz_image_desc* images;
  const vz_prim* prims;
  const vz_tile_desc* tiles;
  const float* lut;
  void* out;
  int out_mode;
};

__global__ void __launch_bounds__(IT) pre_identity_kernel(const IArgs a) {
  extern __shared__ __align__(16) uint8_t dp_smem[];
  const int band = blockIdx.x, t = blockIdx.y, tid = threadIdx.x;
  const vz_image_desc im = a.images[a.tiles[t].image];
  const bool patches = a.out_mode == VZ_OUT_PATCHES_BF16;
  const int stage_bytes = patches ? 24 * VZ_PATCH_K * 2 : 3 * BAND * TILE * 4;
  uint8_t* s_stage = dp_smem;
  float* s_lutf = reinterpret_cast<float*>(dp_smem + stage_bytes);                 // [768] (chw mode)
  __nv_bfloat16* s_luth = reinterpret_cast<__nv_bfloat16*>(dp_smem + stage_bytes); // [768] (patch mode)
  __shared__ vz_prim s_prims[MAX_PRIMS];
  const int n_prims = im.prim_count < MAX_PRIMS ? im.prim_count : MAX_PRIMS;
  for (int i = tid; i < n_prims; i += IT) s_prims[i] = a.prims[im.prim_begin + i];
  for (int i = tid; i < 768; i += IT) {
    if (patches) s_luth[i] = __float2bfloat16_rn(a.lut[i]);
    else s_lutf[i] = a.lut[i];
  }
  if (patches) {
    __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(s_stage);
    for (int i = tid; i < 24 * 4; i += IT) sp[(i >> 2) * VZ_PATCH_K + 588 + (i & 3)] = __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int qd = tid; qd < QUADS; qd += IT) {
    const int ry = qd / (TILE / 4), xq = qd - ry * (TILE / 4);
    const int y = band * BAND + ry, x = 4 * xq;
    const uint32_t* sp32 = reinterpret_cast<const uint32_t*>(im.src + ((size_t)y * TILE + x) * 3);
    const uint32_t w0 = __ldg(sp32), w1 = __ldg(sp32 + 1), w2 = __ldg(sp32 + 2);
    // w0 = R0 G0 B0 R1 | w1 = G1 B1 R2 G2 | w2 = B2 R3 G3 B3
    int r[4], g[4], b[4];
    r[0] = w0 & 0xff; g[0] = (w0 >> 8) & 0xff; b[0] = (w0 >> 16) & 0xff; r[1] = w0 >> 24;
    g[1] = w1 & 0xff; b[1] = (w1 >> 8) & 0xff; r[2] = (w1 >> 16) & 0xff; g[2] = w1 >> 24;
    b[2] = w2 & 0xff; r[3] = (w2 >> 8) & 0xff; g[3] = (w2 >> 16) & 0xff; b[3] = w2 >> 24;
    for (int pi = 0; pi < n_prims; ++pi) {
      const vz_prim& p = s_prims[pi];
      uint32_t ov[4];
      if (p.type == VZ_PRIM_LAYER) {
        const uint4 o4 = __ldg(reinterpret_cast<const uint4*>(im.layers + (((size_t)p.layer * TILE + y) * TILE + x) * 4));
        ov[0] = o4.x; ov[1] = o4.y; ov[2] = o4.z; ov[3] = o4.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) ov[i] = rect_covers(p, x + i, y) ? p.rgba : 0u;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int al = (int)(ov[i] >> 24);
        if (al != 0) {
          r[i] = blend_over(r[i], (int)(ov[i] & 0xff), al);
          g[i] = blend_over(g[i], (int)((ov[i] >> 8) & 0xff), al);
          b[i] = blend_over(b[i], (int)((ov[i] >> 16) & 0xff), al);
        }
      }
    }
    if (patches) {
      __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(s_stage);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int xx = x + i, px = xx / 14, kx = xx - px * 14;
        __nv_bfloat16* q = sp + px * VZ_PATCH_K + ry * 14 + kx;
        q[0] = s_luth[r[i]];
        q[196] = s_luth[256 + g[i]];
        q[392] = s_luth[512 + b[i]];
      }
    } else {
      float* sf = reinterpret_cast<float*>(s_stage) + ry * TILE + x;   // [3][BAND][336], 16-byte aligned (x % 4 == 0)
      *reinterpret_cast<float4*>(sf) = make_float4(s_lutf[r[0]], s_lutf[r[1]], s_lutf[r[2]], s_lutf[r[3]]);
      *reinterpret_cast<float4*>(sf + BAND * TILE) =
          make_float4(s_lutf[256 + g[0]], s_lutf[256 + g[1]], s_lutf[256 + g[2]], s_lutf[256 + g[3]]);
      *reinterpret_cast<float4*>(sf + 2 * BAND * TILE) =
          make_float4(s_lutf[512 + b[0]], s_lutf[512 + b[1]], s_lutf[512 + b[2]], s_lutf[512 + b[3]]);
    }
  }
  __syncthreads();
  if (patches) {
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) +
                                          ((size_t)t * VZ_VIT_PATCHES + band * 24) * VZ_PATCH_K);
    const uint4* s4 = reinterpret_cast<const uint4*>(s_stage);
    for (int i = tid; i < 24 * VZ_PATCH_K * 2 / 16; i += IT) dst[i] = s4[i];
  } else {
    const float* sf = reinterpret_cast<const float*>(s_stage);
    float* o = reinterpret_cast<float*>(a.out);
    for (int i = tid; i < 3 * BAND * TILE / 4; i += IT) {
      const int e = i * 4;
      const int c = e / (BAND * TILE), rem = e - c * BAND * TILE;
      const int yy = rem / TILE, xx = rem - yy * TILE;
      *reinterpret_cast<float4*>(o + (((size_t)t * 3 + c) * TILE + band * BAND + yy) * TILE + xx) =
          *reinterpret_cast<const float4*>(sf + e);
    }
  }
}

template <int GMAX>
int launch_h(const HArgs& h, dim3 grid, size_t smem, cudaStream_t st) {
  VZ_ENSURE_DYN_SMEM(pre_h_dp_kernel<GMAX>, 200 * 1024);
  {
    ProfScope prof(VZ_PROF_PRE_H, 0.0, st);
    pre_h_dp_kernel<GMAX><<<grid, HT, smem, st>>>(h);
  }
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

}  // namespace
}  // namespace vz

extern "C" int vz_preprocess3(const vz_image_desc* images, int n_images, const vz_prim* prims, int n_prims,
                              const vz_hview_desc* hviews, int n_hviews, const vz_tile_desc* tiles, int n_tiles,
                              const int32_t* tables, const float* lut768, int out_mode, void* out, void* scratch,
                              long long scratch_words, int max_span_px, int max_rows, int max_out_w, int max_groups,
                              int max_band_groups, void* stream) {
  using namespace vz;
  if (!images || !hviews || !tiles || !tables || !lut768 || !out || !scratch) return VZ_ERR_BAD_ARG;
  if (n_images <= 0 || n_hviews <= 0 || n_tiles <= 0 || scratch_words <= 0) return VZ_ERR_BAD_ARG;
  if (n_prims > 0 && !prims) return VZ_ERR_BAD_ARG;
  if (out_mode != VZ_OUT_PATCHES_BF16 && out_mode != VZ_OUT_CHW_F32) return VZ_ERR_BAD_ARG;
  if (max_span_px <= 0 || max_rows <= 0 || max_out_w <= 0 || max_groups <= 0 || max_band_groups <= 0) return VZ_ERR_BAD_ARG;
  if (!aligned16(out) || !aligned16(scratch) || !aligned16(tables)) return VZ_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (max_groups > 16) return VZ_ERR_UNSUPPORTED;   // ksize > 58 (scale > 9.5): use vz_preprocess2
  const int G = max_groups <= 2 ? 2 : max_groups <= 4 ? 4 : max_groups <= 6 ? 6 : max_groups <= 8 ? 8 : max_groups <= 12 ? 12 : 16;
  // ---- vertical pass geometry first: the whole call is refused if its window cannot be staged ----
  const int stage_bytes = (out_mode == VZ_OUT_CHW_F32) ? 3 * BAND * TILE * 4 : 24 * VZ_PATCH_K * 2;
  const size_t win_bytes = (size_t)max_band_groups * TILE * 16;
  const size_t smem_v = (win_bytes > (size_t)stage_bytes ? win_bytes : (size_t)stage_bytes) + (size_t)BAND * max_groups * 16 + 128;
  if (smem_v > 200 * 1024) return VZ_ERR_UNSUPPORTED;
  // ---- horizontal pass ----
  HArgs h;
  h.images = images; h.prims = prims; h.hviews = hviews; h.tables = tables;
  h.scratch = reinterpret_cast<uint32_t*>(scratch);
  h.plane_words = ((max_span_px + 4 + 3) >> 2) + G + 1;   // window aligned down to 4 + groups a narrower column skips
  static const int rb_env = []() { const char* e = getenv("VZ_PRE_RB"); return e ? atoi(e) : 0; }();
  int rb = (rb_env == 8 || rb_env == 16 || rb_env == 32) ? rb_env : 32;   // measured on config 3: 126 / 114 / 110 us for 8 / 16 / 32 rows
  size_t smem_h = (size_t)3 * rb * h.plane_words * 4;
  while (smem_h > 96 * 1024 && rb > 8) { rb >>= 1; smem_h = (size_t)3 * rb * h.plane_words * 4; }
  if (smem_h > 200 * 1024) return VZ_ERR_UNSUPPORTED;
  h.rows_per_cta = rb;
  dim3 grid_h((max_out_w + HX - 1) / HX, (max_rows + rb - 1) / rb, n_hviews);
  int s;
  switch (G) {
    case 2: s = launch_h<2>(h, grid_h, smem_h, st); break;
    case 4: s = launch_h<4>(h, grid_h, smem_h, st); break;
    case 6: s = launch_h<6>(h, grid_h, smem_h, st); break;
    case 8: s = launch_h<8>(h, grid_h, smem_h, st); break;
    case 12: s = launch_h<12>(h, grid_h, smem_h, st); break;
    default: s = launch_h<16>(h, grid_h, smem_h, st); break;
  }
  VZ_TRY(s);
  // ---- vertical pass + normalise + patchify ----
  VArgs v;
  v.hviews = hviews; v.tiles = tiles; v.tables = tables; v.lut = lut768;
  v.scratch = reinterpret_cast<const uint32_t*>(scratch); v.out = out; v.out_mode = out_mode;
  v.gmax = max_groups; v.win_groups = max_band_groups;
  VZ_ENSURE_DYN_SMEM(pre_v_dp_kernel, 200 * 1024);
  dim3 grid_v(24, n_tiles);
  {
    ProfScope prof(VZ_PROF_PRE_V, 0.0, st);
    pre_v_dp_kernel<<<grid_v, VT, smem_v, st>>>(v);
  }
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

// Every tile is an identity view of a 336 x 336 image (no canvas, tile origin 0): the caller (preprocess.py)
// knows that from the geometry it built; image and layer base pointers must be 4- / 16-byte aligned.
extern "C" int vz_preprocess_identity(const vz_image_desc* images, int n_images, const vz_prim* prims, int n_prims,
                                      const vz_tile_desc* tiles, int n_tiles, const float* lut768, int out_mode,
                                      void* out, void* stream) {
  using namespace vz;
  if (!images || !tiles || !lut768 || !out || n_images <= 0 || n_tiles <= 0) return VZ_ERR_BAD_ARG;
  if (n_prims > 0 && !prims) return VZ_ERR_BAD_ARG;
  if (out_mode != VZ_OUT_PATCHES_BF16 && out_mode != VZ_OUT_CHW_F32) return VZ_ERR_BAD_ARG;
  if (!aligned16(out)) return VZ_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  IArgs a;
  a.images = images; a.prims = prims; a.tiles = tiles; a.lut = lut768; a.out = out; a.out_mode = out_mode;
  const size_t smem = (size_t)(out_mode == VZ_OUT_CHW_F32 ? 3 * BAND * TILE * 4 : 24 * VZ_PATCH_K * 2) + 768 * 4;
  VZ_ENSURE_DYN_SMEM(pre_identity_kernel, 64 * 1024);
  dim3 grid(24, n_tiles);
  {
    ProfScope prof(VZ_PROF_PRE_FUSED, 0.0, st);
    pre_identity_kernel<<<grid, IT, smem, st>>>(a);
  }
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}
