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
//   * the u8 intermediate is stored as I[channel][row / 4][x]: one word = FOUR VERTICALLY consecutive pixels
//     of one channel, which is what the vertical pass needs for its own dp4a (windows aligned down to four
//     rows).  A horizontal-pass thread computes four rows of one column and writes whole words, coalesced in x.
// The identity form (no resampling: the fixed 336 x 336 inputs of config 2) stays in vz_preprocess.cu.
#include "vz_common.cuh"

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
// limbs -> Pillow's clip8((ss + 2^21) >> 22); the sums wrap modulo 2^32, the true total fits in int32
__device__ __forceinline__ uint32_t finish8(uint32_t a0, uint32_t a1, int a2) {
  const uint32_t t = a0 + (a1 << 8) + ((uint32_t)a2 << 16) + (1u << (PREC - 1));
  int v = (int)t >> PREC;
  v = v < 0 ? 0 : (v > 255 ? 255 : v);
  return (uint32_t)v;
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

struct HArgs {
  const vz_image_desc* images;
  const vz_prim* prims;
  const vz_hview_desc* hviews;
  const int32_t* tables;
  uint32_t* scratch;
  int raw_stride;     // bytes per staged source row (multiple of 16)
  int plane_words;    // 32-bit words per plane row
  int rows_per_cta;   // multiple of 8
};

// ------------------------------------------------------------------------------------------------
// Horizontal pass.  CTA = rows_per_cta canvas rows x 128 output columns of one (image, table) view.
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
  const int32_t* th = a.tables + hv.tab_h;
  const int ksh = th[0];
  const int32_t* h_min = th + 2;
  const int32_t* h_cnt = h_min + hv.out_w;
  const int32_t* h_kk = h_cnt + hv.out_w;
  const int ngc = (ksh + 6) >> 2;            // groups of four taps any column of this view can need (<= GMAX)

  uint32_t* s_co = reinterpret_cast<uint32_t*>(dp_smem);                  // [GMAX][3][HX] limb words of this CTA's columns
  uint8_t* s_raw = dp_smem + (size_t)GMAX * 3 * HX * 4;                   // [RB][raw_stride] interleaved source bytes
  uint32_t* s_pl = reinterpret_cast<uint32_t*>(s_raw + (size_t)RB * a.raw_stride);   // [3][RB][plane_words]
  __shared__ vz_prim s_prims[MAX_PRIMS];
  __shared__ int s_phase[64];
  const int n_prims = im.prim_count < MAX_PRIMS ? im.prim_count : MAX_PRIMS;
  for (int i = tid; i < n_prims; i += HT) s_prims[i] = a.prims[im.prim_begin + i];

  const int xl = tid & (HX - 1), q = tid >> 7;
  const int x = x0 + xl;
  const bool valid = x < hv.out_w;
  const int xlast = min(x0 + HX - 1, hv.out_w - 1);
  const int sx0 = h_min[x0], sx1 = h_min[xlast] + h_cnt[xlast];     // canvas columns this CTA's outputs read
  const int base = sx0 & ~3;                                         // plane byte 0 = canvas column `base`
  const int nvec = (sx1 - base + 3) >> 2;                            // plane words that carry data

  // ---- coefficient limbs of this thread's column, shifted to its aligned window ---------------
  for (int i = tid; i < GMAX * 3 * HX; i += HT) s_co[i] = 0u;
  __syncthreads();
  int wb = 0;
  if (tid < HX && valid) {
    const int xm = h_min[x], cnt = h_cnt[x], sh = xm & 3;
    uint8_t* cb = reinterpret_cast<uint8_t*>(s_co);
    for (int k = 0; k < cnt; ++k) {
      const int c = h_kk[x * ksh + k];
      const int slot = k + sh, g = slot >> 2, b = slot & 3;
      cb[((g * 3 + 0) * HX + xl) * 4 + b] = (uint8_t)(c & 255);
      cb[((g * 3 + 1) * HX + xl) * 4 + b] = (uint8_t)((c >> 8) & 255);
      cb[((g * 3 + 2) * HX + xl) * 4 + b] = (uint8_t)((c >> 16) & 255);
    }
  }
  if (valid) wb = ((h_min[x] & ~3) - base) >> 2;

  // ---- stage the raw source rows (cp.async, 16-byte chunks at their global phase) -------------
  const uint8_t* img_end = im.src + (size_t)im.W * im.H * 3;
  const int rx0 = max(base - im.pad_x, 0), rx1 = min(sx1 - im.pad_x, im.W);   // real image columns of the window
  auto row_phase = [&](int sy) -> int {   // where canvas column `base` of this row sits in its raw buffer
    const int yr = sy - im.pad_y;
    if (yr < 0 || yr >= im.H || rx1 <= rx0) return 0;
    const uintptr_t g = reinterpret_cast<uintptr_t>(im.src + ((size_t)yr * im.W + rx0) * 3);
    return (int)((g - (uintptr_t)((rx0 + im.pad_x - base) * 3)) & 15);
  };
  if (tid < RB) s_phase[tid] = tid < nrows ? row_phase(y0 + tid) : 0;
  __syncthreads();
  for (int r = 0; r < nrows; ++r) {
    const int yr = y0 + r - im.pad_y;
    if (yr < 0 || yr >= im.H || rx1 <= rx0) continue;
    const uint8_t* srow = im.src + ((size_t)yr * im.W + rx0) * 3;
    const int a16 = (int)(reinterpret_cast<uintptr_t>(srow) & 15);
    const uint8_t* g0 = srow - a16;
    uint8_t* d0 = s_raw + (size_t)r * a.raw_stride + s_phase[r] + (rx0 + im.pad_x - base) * 3 - a16;   // 16-byte aligned
    const int nv16 = ((rx1 - rx0) * 3 + a16 + 15) >> 4;
    for (int i = tid; i < nv16; i += HT) {
      const uint8_t* g = g0 + 16 * i;
      const long left = img_end - g;
      cp_async_16_partial(d0 + 16 * i, g, left >= 16 ? 16 : (left > 0 ? (int)left : 0));
    }
  }
  cp_async_commit();
  // this thread's coefficient words -> registers (static indices) while the copies fly
  uint32_t co[GMAX][3];
#pragma unroll
  for (int g = 0; g < GMAX; ++g)
#pragma unroll
    for (int l = 0; l < 3; ++l) co[g][l] = s_co[(g * 3 + l) * HX + xl];
  cp_async_wait<0>();
  __syncthreads();

  // ---- de-interleave (+ canvas padding, + visual prompts) into byte planes, four pixels per thread ----
  const uint32_t bgw = im.bg & 0xffffffu;
  for (int it = tid; it < nrows * nvec; it += HT) {
    const int r = it / nvec, v = it - r * nvec;
    const int cx = base + 4 * v, yr = y0 + r - im.pad_y;
    const bool row_real = yr >= 0 && yr < im.H;
    const int xr0 = cx - im.pad_x;
    uint32_t R, G, B;
    if (row_real && xr0 >= 0 && xr0 + 3 < im.W && n_prims == 0) {
      const int b0 = s_phase[r] + 12 * v;
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_raw + (size_t)r * a.raw_stride) + (b0 >> 2);
      const uint32_t sh = (uint32_t)(b0 & 3) * 8u;
      const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
      const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh), a2 = __funnelshift_r(w2, w3, sh);
      // a0 = R0 G0 B0 R1 | a1 = G1 B1 R2 G2 | a2 = B2 R3 G3 B3   (byte 0 first)
      R = __byte_perm(__byte_perm(a0, a1, 0x0630), a2, 0x5210);
      G = __byte_perm(__byte_perm(a0, a1, 0x0741), a2, 0x6210);
      B = __byte_perm(__byte_perm(a0, a1, 0x0052), a2, 0x7410);
    } else {
      R = G = B = 0u;
      const uint8_t* rawrow = s_raw + (size_t)r * a.raw_stride + s_phase[r] + 12 * v;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int xr = xr0 + i;
        int rr, gg, bb;
        if (row_real && xr >= 0 && xr < im.W) {
          rr = rawrow[3 * i]; gg = rawrow[3 * i + 1]; bb = rawrow[3 * i + 2];
          for (int pi = 0; pi < n_prims; ++pi) {
            const vz_prim& p = s_prims[pi];
            uint32_t ov;
            if (p.type == VZ_PRIM_LAYER) {
              ov = __ldg(reinterpret_cast<const uint32_t*>(im.layers) + ((size_t)p.layer * im.H + yr) * im.W + xr);
            } else {
              if (!rect_covers(p, xr, yr)) continue;
              ov = p.rgba;
            }
            const int al = (int)(ov >> 24);
            rr = blend_over(rr, (int)(ov & 0xff), al);
            gg = blend_over(gg, (int)((ov >> 8) & 0xff), al);
            bb = blend_over(bb, (int)((ov >> 16) & 0xff), al);
          }
        } else {   // canvas padding (expand2square, mm_utils.py:16-35)
          rr = (int)(bgw & 0xff); gg = (int)((bgw >> 8) & 0xff); bb = (int)((bgw >> 16) & 0xff);
        }
        R |= (uint32_t)rr << (8 * i); G |= (uint32_t)gg << (8 * i); B |= (uint32_t)bb << (8 * i);
      }
    }
    s_pl[(0 * RB + r) * a.plane_words + v] = R;
    s_pl[(1 * RB + r) * a.plane_words + v] = G;
    s_pl[(2 * RB + r) * a.plane_words + v] = B;
  }
  __syncthreads();

  // ---- filter: thread = (column x, four consecutive rows), one intermediate word per channel ----
  const int RQ = (hv.rows + 3) >> 2;
  const int nrq = (nrows + 3) >> 2;
  for (int rq = q; rq < nrq; rq += 2) {
    uint32_t o[3] = {0u, 0u, 0u};
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int r = rq * 4 + rr;
      if (r < nrows) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const uint32_t* pw = s_pl + (c * RB + r) * a.plane_words + wb;
          uint32_t a0 = 0u, a1 = 0u;
          int a2 = 0;
#pragma unroll
          for (int g = 0; g < GMAX; ++g) {
            if (g < ngc) {
              const uint32_t p = pw[g];
              a0 = dp4a_uu(p, co[g][0], a0);
              a1 = dp4a_uu(p, co[g][1], a1);
              a2 = dp4a_us(p, co[g][2], a2);
            }
          }
          o[c] |= finish8(a0, a1, a2) << (8 * rr);
        }
      }
    }
    if (valid) {
      uint32_t* dst = a.scratch + hv.offset + (size_t)((y0 >> 2) + rq) * hv.out_w + x;
      dst[0] = o[0];
      dst[(size_t)RQ * hv.out_w] = o[1];
      dst[(size_t)2 * RQ * hv.out_w] = o[2];
    }
  }
}

struct VArgs {
  const vz_hview_desc* hviews;
  const vz_tile_desc* tiles;
  const int32_t* tables;
  const float* lut;
  const uint32_t* scratch;
  void* out;
  int out_mode;
  int gv;             // groups of four taps per output row the coefficient table is laid out for
};

// ------------------------------------------------------------------------------------------------
// Vertical pass + LUT + im2col of one 14-row band of one tile.  Thread x gathers, per band row, the words
// I[c][j0 + g][rx] of its column (coalesced in x, L1 / L2 resident) and runs 9 dp4a per group.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VT) pre_v_dp_kernel(const VArgs a) {
  extern __shared__ __align__(16) uint8_t dp_smem[];
  const int band = blockIdx.x, t = blockIdx.y, tid = threadIdx.x;
  const vz_tile_desc td = a.tiles[t];
  const vz_hview_desc hv = a.hviews[td.hview];
  const int32_t* tv = a.tables + td.tab_v;
  const int ksv = tv[0];
  const int32_t* v_min = tv + 2;
  const int32_t* v_cnt = v_min + td.out_h;
  const int32_t* v_kk = v_cnt + td.out_h;
  const int stage_bytes = (a.out_mode == VZ_OUT_CHW_F32) ? 3 * BAND * TILE * 4 : 24 * VZ_PATCH_K * 2;
  uint8_t* s_stage = dp_smem;
  uint32_t* s_vc = reinterpret_cast<uint32_t*>(dp_smem + stage_bytes);   // [BAND][gv][4]: limb words c0, c1, c2, 0
  int32_t* s_vj0 = reinterpret_cast<int32_t*>(s_vc + BAND * a.gv * 4);
  int32_t* s_vng = s_vj0 + BAND;
  const int ry0 = td.tile_y + band * BAND - td.off_y;   // resized-image row of band row 0
  for (int i = tid; i < BAND * a.gv * 4; i += VT) s_vc[i] = 0u;
  if (tid < BAND) {
    const int ry = ry0 + tid;
    const bool ok = ry >= 0 && ry < td.out_h;
    const int vm = ok ? v_min[ry] : 0, cnt = ok ? v_cnt[ry] : 0;
    s_vj0[tid] = vm >> 2;
    s_vng[tid] = cnt > 0 ? (cnt + (vm & 3) + 3) >> 2 : 0;
  }
  if (a.out_mode == VZ_OUT_PATCHES_BF16) {
    __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(s_stage);
    for (int i = tid; i < 24 * 4; i += VT) sp[(i >> 2) * VZ_PATCH_K + 588 + (i & 3)] = __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  {
    uint8_t* cb = reinterpret_cast<uint8_t*>(s_vc);
    for (int i = tid; i < BAND * ksv; i += VT) {
      const int y = i / ksv, k = i - y * ksv;
      const int ry = ry0 + y;
      if (ry < 0 || ry >= td.out_h || k >= v_cnt[ry]) continue;
      const int c = v_kk[ry * ksv + k];
      const int slot = k + (v_min[ry] & 3), g = slot >> 2, b = slot & 3;
      uint8_t* w = cb + (size_t)((y * a.gv + g) * 4) * 4 + b;
      w[0] = (uint8_t)(c & 255);
      w[4] = (uint8_t)((c >> 8) & 255);
      w[8] = (uint8_t)((c >> 16) & 255);
    }
  }
  __syncthreads();
  const int x = tid;
  const int rx = td.tile_x + x - td.off_x;
  const bool col_ok = tid < TILE && rx >= 0 && rx < td.out_w;
  const int RQ = (hv.rows + 3) >> 2;
  const size_t plane = (size_t)RQ * hv.out_w;
  const uint32_t* col = a.scratch + hv.offset + (col_ok ? rx : 0);
  const int px = x / 14, kx = x - px * 14;
#pragma unroll 2
  for (int y = 0; y < BAND; ++y) {
    const int ng = col_ok ? s_vng[y] : 0;
    const uint32_t* src = col + (size_t)s_vj0[y] * hv.out_w;
    const uint4* cw = reinterpret_cast<const uint4*>(s_vc + (size_t)y * a.gv * 4);
    uint32_t r0 = 0u, r1 = 0u, g0 = 0u, g1 = 0u, b0 = 0u, b1 = 0u;
    int r2 = 0, g2 = 0, b2 = 0;
    for (int g = 0; g < ng; ++g) {
      const uint4 c = cw[g];
      const uint32_t pr = __ldg(src + (size_t)g * hv.out_w);
      const uint32_t pg = __ldg(src + plane + (size_t)g * hv.out_w);
      const uint32_t pb = __ldg(src + 2 * plane + (size_t)g * hv.out_w);
      r0 = dp4a_uu(pr, c.x, r0); r1 = dp4a_uu(pr, c.y, r1); r2 = dp4a_us(pr, c.z, r2);
      g0 = dp4a_uu(pg, c.x, g0); g1 = dp4a_uu(pg, c.y, g1); g2 = dp4a_us(pg, c.z, g2);
      b0 = dp4a_uu(pb, c.x, b0); b1 = dp4a_uu(pb, c.y, b1); b2 = dp4a_us(pb, c.z, b2);
    }
    if (tid < TILE) {
      const int v0 = ng > 0 ? (int)finish8(r0, r1, r2) : 0, v1 = ng > 0 ? (int)finish8(g0, g1, g2) : 0,
                v2 = ng > 0 ? (int)finish8(b0, b1, b2) : 0;
      if (a.out_mode == VZ_OUT_PATCHES_BF16) {
        __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(s_stage) + px * VZ_PATCH_K + y * 14 + kx;
        sp[0] = __float2bfloat16_rn(a.lut[v0]);
        sp[196] = __float2bfloat16_rn(a.lut[256 + v1]);
        sp[392] = __float2bfloat16_rn(a.lut[512 + v2]);
      } else {
        float* sf = reinterpret_cast<float*>(s_stage) + y * TILE + x;   // [3][BAND][336]
        sf[0] = a.lut[v0];
        sf[BAND * TILE] = a.lut[256 + v1];
        sf[2 * BAND * TILE] = a.lut[512 + v2];
      }
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
                              long long scratch_words, int max_span_px, int max_rows, int max_out_w, int max_ksize,
                              void* stream) {
  using namespace vz;
  if (!images || !hviews || !tiles || !tables || !lut768 || !out || !scratch) return VZ_ERR_BAD_ARG;
  if (n_images <= 0 || n_hviews <= 0 || n_tiles <= 0 || scratch_words <= 0) return VZ_ERR_BAD_ARG;
  if (n_prims > 0 && !prims) return VZ_ERR_BAD_ARG;
  if (out_mode != VZ_OUT_PATCHES_BF16 && out_mode != VZ_OUT_CHW_F32) return VZ_ERR_BAD_ARG;
  if (max_span_px <= 0 || max_rows <= 0 || max_out_w <= 0 || max_ksize <= 0 || !aligned16(out) || !aligned16(scratch))
    return VZ_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int gmax = (max_ksize + 6) >> 2;     // groups of four taps after aligning a window down to 4
  if (gmax > 16) return VZ_ERR_UNSUPPORTED;  // ksize > 58 (scale > 9.5): use vz_preprocess2
  const int G = gmax <= 2 ? 2 : gmax <= 4 ? 4 : gmax <= 6 ? 6 : gmax <= 8 ? 8 : gmax <= 12 ? 12 : 16;
  // ---- horizontal pass ----
  HArgs h;
  h.images = images; h.prims = prims; h.hviews = hviews; h.tables = tables;
  h.scratch = reinterpret_cast<uint32_t*>(scratch);
  const int span = max_span_px + 4;                                   // + alignment of the window start
  h.raw_stride = ((span * 3 + 16 + 16 + 15) / 16) * 16;               // + phase + the 16 bytes the last funnel shift reads
  h.plane_words = ((span + 3) >> 2) + G + 1;                          // + groups a narrower column does not need
  int rb = 32;
  size_t smem_h = 0;
  for (;; rb >>= 1) {
    smem_h = (size_t)G * 3 * HX * 4 + (size_t)rb * h.raw_stride + (size_t)3 * rb * h.plane_words * 4;
    if (smem_h <= 200 * 1024 || rb == 8) break;
  }
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
  v.scratch = reinterpret_cast<const uint32_t*>(scratch); v.out = out; v.out_mode = out_mode; v.gv = gmax;
  const int stage_bytes = (out_mode == VZ_OUT_CHW_F32) ? 3 * BAND * TILE * 4 : 24 * VZ_PATCH_K * 2;
  const size_t smem_v = (size_t)stage_bytes + (size_t)BAND * gmax * 16 + 2 * BAND * 4 + 16;
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
