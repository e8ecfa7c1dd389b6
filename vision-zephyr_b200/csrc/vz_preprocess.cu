// vz_preprocess.cu -- fused visual-prompt alpha blend + anyres LANCZOS resize / centre pad /
// tile cut + CLIP normalise + patchify (subsystem (1) of the north star).
//
// Bit-exact restatement of the Pillow arithmetic the reference calls:
//   Image.alpha_composite          vip_processor/conversation_generator.py:143-146
//   image.resize(..., LANCZOS)     multi_scale_process.py:86-89, 171-174   (Resample.c: horizontal
//                                  pass first, u8 intermediate, 22-bit fixed-point coefficients)
//   Image.new + paste, crop        multi_scale_process.py:91-93, 109-113
//   CLIPImageProcessor.preprocess  multi_scale_process.py:178-181 (768-entry LUT from the oracle)
//
// One CTA = one 14-row band (one patch row) of one 336x336 output tile.  Source rows are streamed
// through a double-buffered shared-memory row buffer (blended on the way in); every thread owns one
// output PIXEL (3 channels) of the band's 336, runs the horizontal pass for it and folds the result
// straight into 14 x 3 fixed-point vertical accumulators held in registers, so neither the blended
// image nor the u8 intermediate ever touches HBM.  The band is then normalised through the LUT,
// transposed into im2col order in shared memory and written with 128-bit stores.
// The tap loops are bound by instruction issue, not by bytes (11-19 taps per axis): the horizontal
// pass walks 4 taps = 12 interleaved RGB bytes at a time -- three aligned 32-bit shared-memory loads
// re-aligned with funnel shifts, one 128-bit load for the four coefficients -- and the vertical pass
// is an unconditional 14 x 3 multiply-add against a zero-padded per-source-row coefficient table.
#include "vz_common.cuh"

namespace vz {
namespace {

constexpr int TILE = 336;
constexpr int BAND = 14;
constexpr int PP_THREADS = 352;     // 11 warps: thread x < 336 owns output pixel x of the band
constexpr int VT_STRIDE = 16;       // ints per row of the vertical coefficient table (14 used)
constexpr int NBUF = 4;             // source-row ring: rows sy+1 .. sy+3 are in flight (cp.async) while row sy is filtered
constexpr int PREC = 22;
constexpr int MAX_PRIMS = 32;  // visual-prompt instances per image (the reference draws 1-4)

struct PreArgs {
  const vz_image_desc* images;
  const vz_prim* prims;
  const vz_tile_desc* tiles;
  const int32_t* tables;
  const float* lut;
  void* out;
  int out_mode;
  int row_buf_bytes;  // per row buffer (>= 3*max W, multiple of 16)
  int max_ksize;
};

// byte i of a 32-bit word, zero extended (one PRMT)
template <int I>
__device__ __forceinline__ int byte_of(uint32_t w) { return (int)__byte_perm(w, 0u, 0x4440u | I); }

__device__ __forceinline__ int clip8(int v) {
  v >>= PREC;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// Pillow AlphaComposite.c with an opaque destination
__device__ __forceinline__ int blend_over(int dst, int src, int alpha) {
  if (alpha == 0) return dst;
  const uint32_t t = (uint32_t)src * (uint32_t)(alpha * 128) + (uint32_t)dst * (uint32_t)((255 - alpha) * 128) + (0x80u << 7);
  return (int)((((t >> 8) + t) >> 8) >> 7);
}

// PIL ImageDraw.rectangle(outline, width) coverage = union of the 4*width lines that Draw.c
// ImagingDrawRectangle (fill = 0) paints: hlines y0+i / y1-i over [x0,x1]; vlines x1-i / x0+i from
// y0+width towards y1-width+1 with the far end point excluded (line32, dx == 0).  width == 0 is
// filtered on the host (ImageDraw does not call the C routine then).
__device__ __forceinline__ bool rect_covers(const vz_prim& p, int x, int y) {
  const int w = p.width;
  if (w <= 0) return false;
  const bool hl = (x >= p.x0 && x <= p.x1) && ((y >= p.y0 && y < p.y0 + w) || (y <= p.y1 && y > p.y1 - w));
  const int va = p.y0 + w, vb = p.y1 - w + 1;
  const bool in_v = (vb >= va) ? (y >= va && y < vb) : (y <= va && y > vb);
  const bool vl = in_v && ((x <= p.x1 && x > p.x1 - w) || (x >= p.x0 && x < p.x0 + w));
  return hl || vl;
}

__global__ void __launch_bounds__(PP_THREADS, 2) preprocess_kernel(const PreArgs a) {
  extern __shared__ __align__(16) uint8_t pp_smem[];
  const int band = blockIdx.x, t = blockIdx.y, tid = threadIdx.x;
  const vz_tile_desc td = a.tiles[t];
  const vz_image_desc im = a.images[td.image];

  const int32_t* th = a.tables + td.tab_h;
  const int32_t* tv = a.tables + td.tab_v;
  const int ksh = th[0], ksv = tv[0];
  const int32_t* h_min = th + 2;
  const int32_t* h_cnt = h_min + td.out_w;
  const int32_t* h_kk = h_cnt + td.out_w;
  const int32_t* v_min = tv + 2;
  const int32_t* v_cnt = v_min + td.out_h;
  const int32_t* v_kk = v_cnt + td.out_h;

  const int stage_bytes = (a.out_mode == VZ_OUT_CHW_F32) ? 3 * BAND * TILE * 4 : 24 * VZ_PATCH_K * 2;
  const int ksh4 = (ksh + 3) >> 2;              // horizontal taps in groups of four
  const int vt_rows = 4 * a.max_ksize + 16;     // capacity of the vertical table (source rows of one band)
  uint8_t* s_stage = pp_smem;
  uint8_t* s_row = pp_smem + stage_bytes;
  int32_t* s_vtab = reinterpret_cast<int32_t*>(s_row + NBUF * a.row_buf_bytes);   // [vt_rows][16]: coefficient of source row i for band row y
  int32_t* s_vmin = s_vtab + vt_rows * VT_STRIDE;
  int32_t* s_vcnt = s_vmin + BAND;
  int4* s_hkk4 = reinterpret_cast<int4*>(s_vcnt + BAND + 4);  // [ksh4][336]: taps 4g..4g+3 of output column x (zero padded)

  // ---- instance list of this image (visual prompts) -> shared memory ---------------------------
  __shared__ vz_prim s_prims[MAX_PRIMS];
  const int n_prims = im.prim_count < MAX_PRIMS ? im.prim_count : MAX_PRIMS;
  for (int i = tid; i < n_prims; i += PP_THREADS) s_prims[i] = a.prims[im.prim_begin + i];

  // ---- vertical windows of the band's 14 output rows ------------------------------------------
  const int ry0 = td.tile_y + band * BAND - td.off_y;  // resized-image row of band row 0
  if (tid < BAND) {
    const int ry = ry0 + tid;
    const bool ok = ry >= 0 && ry < td.out_h;
    s_vmin[tid] = ok ? v_min[ry] : 0;
    s_vcnt[tid] = ok ? v_cnt[ry] : 0;
  }
  // ---- horizontal coefficients of the tile's 336 columns, four taps per 128-bit entry ----------
  for (int i = tid; i < ksh4 * TILE; i += PP_THREADS) {
    const int g = i / TILE, x = i - g * TILE;
    const int rx = td.tile_x + x - td.off_x;
    int c[4] = {0, 0, 0, 0};
    if (rx >= 0 && rx < td.out_w) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (4 * g + k < ksh) c[k] = h_kk[rx * ksh + 4 * g + k];
    }
    s_hkk4[i] = make_int4(c[0], c[1], c[2], c[3]);
  }
  // ---- this thread's output pixel -----------------------------------------------------------------
  const int x = tid < TILE ? tid : 0;
  int hx_min, hx_cnt;
  {
    const int rx = td.tile_x + x - td.off_x;
    const bool ok = tid < TILE && rx >= 0 && rx < td.out_w;
    hx_min = ok ? h_min[rx] : -1;
    hx_cnt = ok ? h_cnt[rx] : 0;
  }
  // source column window needed by this tile (uniform per CTA)
  int sx0, sx1;
  {
    int rxa = td.tile_x - td.off_x, rxb = td.tile_x + TILE - 1 - td.off_x;
    rxa = rxa < 0 ? 0 : rxa;
    rxb = rxb >= td.out_w ? td.out_w - 1 : rxb;
    if (rxa <= rxb) { sx0 = h_min[rxa]; sx1 = h_min[rxb] + h_cnt[rxb]; }
    else { sx0 = 0; sx1 = 0; }
  }
  if (hx_min < 0) hx_min = sx0;  // invalid column: all-zero taps, any in-buffer address will do
  __syncthreads();
  // source row window of the band
  int sy0 = 0, sy1 = 0;
  {
    int ya = 0, yb = BAND - 1;
    while (ya < BAND && s_vcnt[ya] == 0) ++ya;
    while (yb >= 0 && s_vcnt[yb] == 0) --yb;
    if (ya <= yb) { sy0 = s_vmin[ya]; sy1 = s_vmin[yb] + s_vcnt[yb]; }
  }
  if (sx1 <= sx0) sy1 = sy0;  // tile lies completely in the padding

  int acc[BAND][3];
#pragma unroll
  for (int y = 0; y < BAND; ++y)
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[y][c] = 1 << (PREC - 1);

  // ---- fast path: no resampling (identity tables, one tap per axis: the fixed-336 inputs of config 2) ----
  // Nothing is staged or walked row by row: the thread of output pixel x gathers its 14 source pixels
  // straight from global memory (all loads of a visual-prompt layer in flight together), blends the
  // instances in order and leaves the values in the accumulators in the fixed-point form the common
  // epilogue expects.  HBM-bound apart from the launch.
  if (ksh == 1 && ksv == 1) {
    const int rx = td.tile_x + x - td.off_x;
    const bool col_ok = tid < TILE && rx >= 0 && rx < td.out_w;
    const int xr = (col_ok ? h_min[rx] : 0) - im.pad_x;
    const int kh = col_ok ? h_kk[rx] : 0;
    uint32_t rgb[BAND];
    bool real[BAND];
    int yrs[BAND];
#pragma unroll
    for (int y = 0; y < BAND; ++y) {
      const int ry = ry0 + y;
      const bool row_ok = ry >= 0 && ry < td.out_h;
      const int yr = (row_ok ? v_min[ry] : 0) - im.pad_y;
      yrs[y] = yr;
      real[y] = col_ok && row_ok && xr >= 0 && xr < im.W && yr >= 0 && yr < im.H;
      uint32_t v = im.bg & 0xffffffu;
      if (real[y]) {
        const uint8_t* q = im.src + ((size_t)yr * im.W + xr) * 3;
        v = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16);
      }
      rgb[y] = v;
    }
    for (int pi = 0; pi < n_prims; ++pi) {
      const vz_prim p = s_prims[pi];
      uint32_t ov[BAND];
      if (p.type == VZ_PRIM_LAYER) {
        const uint32_t* lay = reinterpret_cast<const uint32_t*>(im.layers) + (size_t)p.layer * im.H * im.W;
#pragma unroll
        for (int y = 0; y < BAND; ++y) ov[y] = real[y] ? __ldg(lay + (size_t)yrs[y] * im.W + xr) : 0u;
      } else {
#pragma unroll
        for (int y = 0; y < BAND; ++y) ov[y] = (real[y] && rect_covers(p, xr, yrs[y])) ? p.rgba : 0u;
      }
#pragma unroll
      for (int y = 0; y < BAND; ++y) {
        const int al = (int)(ov[y] >> 24);   // alpha 0 leaves the pixel untouched (blend_over)
        const int r = blend_over((int)(rgb[y] & 0xff), (int)(ov[y] & 0xff), al);
        const int g = blend_over((int)((rgb[y] >> 8) & 0xff), (int)((ov[y] >> 8) & 0xff), al);
        const int bl = blend_over((int)((rgb[y] >> 16) & 0xff), (int)((ov[y] >> 16) & 0xff), al);
        rgb[y] = (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)bl << 16);
      }
    }
#pragma unroll
    for (int y = 0; y < BAND; ++y) {
      const int ry = ry0 + y;
      const int kv = (ry >= 0 && ry < td.out_h) ? v_kk[ry] : 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int hv = clip8((1 << (PREC - 1)) + (int)((rgb[y] >> (8 * c)) & 0xff) * kh);   // the two one-tap passes
        acc[y][c] = (1 << (PREC - 1)) + hv * kv;
      }
    }
    sy1 = sy0;   // no row loop
  }
  if (sy1 - sy0 > vt_rows) __trap();   // cannot happen: a band spans < 3.4 * ksize source rows (see vz_preprocess)
  // ---- vertical coefficient table: entry [i][y] = tap of source row sy0 + i in band row y, else 0 ----
  for (int i = tid; i < (sy1 - sy0) * VT_STRIDE; i += PP_THREADS) {
    const int r = i / VT_STRIDE, y = i - r * VT_STRIDE;
    int coef = 0;
    if (y < BAND) {
      const int k = sy0 + r - s_vmin[y];
      if (k >= 0 && k < s_vcnt[y]) coef = v_kk[(ry0 + y) * ksv + k];
    }
    s_vtab[i] = coef;
  }

  const uint8_t* img_end = im.src + (size_t)im.W * im.H * 3;
  // The tables address a virtual CANVAS; the image sits at (pad_x, pad_y) inside it (expand2square
  // padding: pad >= 0 and the rest of the canvas has the colour `bg`; centre crop: pad <= 0).
  // Real columns of this tile's window, in image coordinates:
  const int rx0 = max(sx0 - im.pad_x, 0), rx1 = min(sx1 - im.pad_x, im.W);
  const bool needs_fill = (sx0 - im.pad_x < 0) || (sx1 - im.pad_x > im.W) || (sy0 - im.pad_y < 0) ||
                          (sy1 - im.pad_y > im.H);
  // Rows land in smem at the same 16-byte phase they have in global memory, so the copy is made of
  // aligned 16-byte chunks; `row_phase(sy)` is where canvas column sx0 sits in the row's buffer.
  auto row_phase = [&](int sy) -> int {
    const int yr = sy - im.pad_y;
    if (yr < 0 || yr >= im.H || rx1 <= rx0) return 0;
    const uintptr_t g = reinterpret_cast<uintptr_t>(im.src + ((size_t)yr * im.W + rx0) * 3);
    return (int)((g - (uintptr_t)((rx0 + im.pad_x - sx0) * 3)) & 15);
  };
  // asynchronous copy of the raw bytes of canvas row sy into a ring slot (one commit group per row)
  auto issue_row = [&](int sy, uint8_t* dst) {
    const int yr = sy - im.pad_y;
    if (sy < sy1 && yr >= 0 && yr < im.H && rx1 > rx0) {
      const uint8_t* srow = im.src + ((size_t)yr * im.W + rx0) * 3;
      const int a16 = (int)(reinterpret_cast<uintptr_t>(srow) & 15);
      const uint8_t* g0 = srow - a16;
      uint8_t* d0 = dst + row_phase(sy) + (rx0 + im.pad_x - sx0) * 3 - a16;   // 16-byte aligned, >= dst
      const int nvec = ((rx1 - rx0) * 3 + a16 + 15) >> 4;
      for (int i = tid; i < nvec; i += PP_THREADS) {
        const uint8_t* g = g0 + 16 * i;
        const long left = img_end - g;     // bytes of the image from g on
        cp_async_16_partial(d0 + 16 * i, g, left >= 16 ? 16 : (left > 0 ? (int)left : 0));
      }
    }
    cp_async_commit();
  };
  // canvas pixels outside the image take the background colour (expand2square, mm_utils.py:16-35)
  auto fill_row = [&](int sy, uint8_t* dst) {
    const int ph = row_phase(sy);
    const int yr = sy - im.pad_y;
    const bool row_real = yr >= 0 && yr < im.H;
    const uint8_t b0 = (uint8_t)(im.bg & 0xff), b1 = (uint8_t)((im.bg >> 8) & 0xff), b2 = (uint8_t)((im.bg >> 16) & 0xff);
    for (int px = tid; px < sx1 - sx0; px += PP_THREADS) {
      const int xr = sx0 + px - im.pad_x;
      if (row_real && xr >= 0 && xr < im.W) continue;
      uint8_t* q = dst + ph + px * 3;
      q[0] = b0; q[1] = b1; q[2] = b2;
    }
  };
  // visual prompts: blend the instances over the row in place (one thread per pixel; the instance list
  // sits in shared memory and every RGBA overlay pixel is one aligned 32-bit load); image coordinates
  auto blend_row = [&](int sy, uint8_t* dst) {
    const int ph = row_phase(sy);
    const int yr = sy - im.pad_y;
    if (yr < 0 || yr >= im.H) return;
    for (int px = tid; px < sx1 - sx0; px += PP_THREADS) {
      const int xr = sx0 + px - im.pad_x;
      if (xr < 0 || xr >= im.W) continue;
      uint8_t* q = dst + ph + px * 3;
      int r = q[0], g = q[1], b = q[2];
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
        r = blend_over(r, (int)(ov & 0xff), al);
        g = blend_over(g, (int)((ov >> 8) & 0xff), al);
        b = blend_over(b, (int)((ov >> 16) & 0xff), al);
      }
      q[0] = (uint8_t)r; q[1] = (uint8_t)g; q[2] = (uint8_t)b;
    }
  };

  for (int i = 0; i < NBUF - 1; ++i) issue_row(sy0 + i, s_row + i * a.row_buf_bytes);
  const int4* hk = s_hkk4 + x;
  for (int sy = sy0; sy < sy1; ++sy) {
    uint8_t* cur = s_row + ((sy - sy0) % NBUF) * a.row_buf_bytes;
    cp_async_wait<NBUF - 2>();   // row sy has landed (for this thread's copies) ...
    __syncthreads();             // ... and for everyone's; row sy - 1 is no longer read by anybody
    issue_row(sy + NBUF - 1, s_row + ((sy + NBUF - 1 - sy0) % NBUF) * a.row_buf_bytes);
    if (needs_fill || im.prim_count != 0) {
      if (needs_fill) fill_row(sy, cur);
      if (im.prim_count != 0) blend_row(sy, cur);
      __syncthreads();
    }
    // horizontal pass for this thread's pixel: taps beyond the column's count have zero coefficients
    int hv0, hv1, hv2;
    {
      const int b0 = row_phase(sy) + (hx_min - sx0) * 3;          // first byte of tap 0 in the row buffer
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(cur) + (b0 >> 2);
      const uint32_t sh = (uint32_t)(b0 & 3) * 8u;
      int s0 = 1 << (PREC - 1), s1 = s0, s2 = s0;
      uint32_t w0 = wp[0];
      for (int g = 0; g < ksh4; ++g) {
        const uint32_t w1 = wp[3 * g + 1], w2 = wp[3 * g + 2], w3 = wp[3 * g + 3];
        const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh),
                       a2 = __funnelshift_r(w2, w3, sh);     // 12 bytes = 4 RGB pixels, byte-aligned to tap 4g
        const int4 c = hk[g * TILE];
        s0 += byte_of<0>(a0) * c.x + byte_of<3>(a0) * c.y + byte_of<2>(a1) * c.z + byte_of<1>(a2) * c.w;
        s1 += byte_of<1>(a0) * c.x + byte_of<0>(a1) * c.y + byte_of<3>(a1) * c.z + byte_of<2>(a2) * c.w;
        s2 += byte_of<2>(a0) * c.x + byte_of<1>(a1) * c.y + byte_of<0>(a2) * c.z + byte_of<3>(a2) * c.w;
        w0 = w3;
      }
      hv0 = clip8(s0); hv1 = clip8(s1); hv2 = clip8(s2);
    }
    // vertical pass: every band row takes this source row with its (possibly zero) coefficient
    {
      const int4* vt = reinterpret_cast<const int4*>(s_vtab + (sy - sy0) * VT_STRIDE);
      const int4 v0 = vt[0], v1 = vt[1], v2 = vt[2], v3 = vt[3];
      const int vk[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
#pragma unroll
      for (int y = 0; y < BAND; ++y) {
        acc[y][0] += hv0 * vk[y];
        acc[y][1] += hv1 * vk[y];
        acc[y][2] += hv2 * vk[y];
      }
    }
  }
  cp_async_wait<0>();

  // ---- normalise + transpose into the output layout -------------------------------------------
  if (a.out_mode == VZ_OUT_PATCHES_BF16) {
    __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(s_stage);
    for (int i = tid; i < 24 * 4; i += PP_THREADS) sp[(i >> 2) * VZ_PATCH_K + 588 + (i & 3)] = __float2bfloat16_rn(0.f);
    if (tid < TILE) {
      const int px = x / 14, kx = x - px * 14;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int y = 0; y < BAND; ++y) {
          const int v = (hx_cnt > 0 && s_vcnt[y] > 0) ? clip8(acc[y][c]) : 0;
          sp[px * VZ_PATCH_K + c * 196 + y * 14 + kx] = __float2bfloat16_rn(a.lut[c * 256 + v]);
        }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) +
                                          ((size_t)t * VZ_VIT_PATCHES + band * 24) * VZ_PATCH_K);
    const uint4* s4 = reinterpret_cast<const uint4*>(sp);
    for (int i = tid; i < 24 * VZ_PATCH_K * 2 / 16; i += PP_THREADS) dst[i] = s4[i];
  } else {
    float* sf = reinterpret_cast<float*>(s_stage);  // [3][BAND][336]
    if (tid < TILE) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int y = 0; y < BAND; ++y) {
          const int v = (hx_cnt > 0 && s_vcnt[y] > 0) ? clip8(acc[y][c]) : 0;
          sf[(c * BAND + y) * TILE + x] = a.lut[c * 256 + v];
        }
    }
    __syncthreads();
    float* o = reinterpret_cast<float*>(a.out);
    for (int i = tid; i < 3 * BAND * TILE / 4; i += PP_THREADS) {
      const int e = i * 4;
      const int c = e / (BAND * TILE), rem = e - c * BAND * TILE;
      const int y = rem / TILE, xx = rem - y * TILE;
      *reinterpret_cast<float4*>(o + (((size_t)t * 3 + c) * TILE + band * BAND + y) * TILE + xx) =
          *reinterpret_cast<const float4*>(sf + e);
    }
  }
}


// ================================================================================================
// Two-kernel form (vz_preprocess2): horizontal pass once per source row, vertical pass by gather.
// ================================================================================================
constexpr int HP_THREADS = 256;   // output columns of one horizontal-pass CTA
constexpr int HP_ROWS = 8;        // source rows of one horizontal-pass CTA (they share the coefficient slice)

struct HArgs {
  const vz_image_desc* images;
  const vz_prim* prims;
  const vz_hview_desc* hviews;
  const int32_t* tables;
  uint32_t* scratch;
  int row_buf_bytes;   // per source row (>= 3 * max span + tap overrun + phase, multiple of 16)
  int max_ksize;
};

// Horizontal pass (+ visual-prompt blend, + canvas padding) of HP_ROWS canvas rows x 256 output columns
// of one view: the rows' source windows are copied into shared memory with cp.async, blended in place,
// then every thread filters its output pixel (3 channels) with the 4-taps-per-step loop of the fused
// kernel and writes one RGBX word per row.
__global__ void __launch_bounds__(HP_THREADS) preprocess_h_kernel(const HArgs a) {
  extern __shared__ __align__(16) uint8_t pp_smem[];
  const vz_hview_desc hv = a.hviews[blockIdx.z];
  const int x0 = blockIdx.x * HP_THREADS, y0 = blockIdx.y * HP_ROWS, tid = threadIdx.x;
  if (x0 >= hv.out_w || y0 >= hv.rows) return;
  const vz_image_desc im = a.images[hv.image];
  const int nrows = min(HP_ROWS, hv.rows - y0);
  const int32_t* th = a.tables + hv.tab_h;
  const int ksh = th[0];
  const int32_t* h_min = th + 2;
  const int32_t* h_cnt = h_min + hv.out_w;
  const int32_t* h_kk = h_cnt + hv.out_w;
  const int ksh4 = (ksh + 3) >> 2;
  int4* s_hkk4 = reinterpret_cast<int4*>(pp_smem);                                   // [ksh4][256]
  uint8_t* s_rows = pp_smem + (size_t)((a.max_ksize + 3) >> 2) * HP_THREADS * 16;     // [HP_ROWS][row_buf_bytes]
  __shared__ vz_prim s_prims[MAX_PRIMS];
  const int n_prims = im.prim_count < MAX_PRIMS ? im.prim_count : MAX_PRIMS;
  for (int i = tid; i < n_prims; i += HP_THREADS) s_prims[i] = a.prims[im.prim_begin + i];

  const int x = x0 + tid;
  const bool valid = x < hv.out_w;
  const int xl = min(x0 + HP_THREADS - 1, hv.out_w - 1);
  const int sx0 = h_min[x0], sx1 = h_min[xl] + h_cnt[xl];      // canvas columns this CTA's outputs read
  const int hx_min = valid ? h_min[x] : sx0;
  for (int g = 0; g < ksh4; ++g) {
    int c[4] = {0, 0, 0, 0};
    if (valid) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (4 * g + k < ksh) c[k] = h_kk[x * ksh + 4 * g + k];
    }
    s_hkk4[g * HP_THREADS + tid] = make_int4(c[0], c[1], c[2], c[3]);   // read back by this thread only
  }
  const uint8_t* img_end = im.src + (size_t)im.W * im.H * 3;
  const int rx0 = max(sx0 - im.pad_x, 0), rx1 = min(sx1 - im.pad_x, im.W);
  const bool needs_fill = (sx0 - im.pad_x < 0) || (sx1 - im.pad_x > im.W) || (y0 - im.pad_y < 0) ||
                          (y0 + nrows - im.pad_y > im.H);
  auto row_phase = [&](int sy) -> int {
    const int yr = sy - im.pad_y;
    if (yr < 0 || yr >= im.H || rx1 <= rx0) return 0;
    const uintptr_t g = reinterpret_cast<uintptr_t>(im.src + ((size_t)yr * im.W + rx0) * 3);
    return (int)((g - (uintptr_t)((rx0 + im.pad_x - sx0) * 3)) & 15);
  };
  // where canvas column sx0 of each row sits in its buffer (computed once, read by everybody)
  __shared__ int s_phase[HP_ROWS];
  if (tid < HP_ROWS) s_phase[tid] = tid < nrows ? row_phase(y0 + tid) : 0;
  __syncthreads();
  // ---- all rows of the block in flight at once ----
  for (int r = 0; r < nrows; ++r) {
    const int sy = y0 + r, yr = sy - im.pad_y;
    if (yr < 0 || yr >= im.H || rx1 <= rx0) continue;
    const uint8_t* srow = im.src + ((size_t)yr * im.W + rx0) * 3;
    const int a16 = (int)(reinterpret_cast<uintptr_t>(srow) & 15);
    const uint8_t* g0 = srow - a16;
    uint8_t* d0 = s_rows + r * a.row_buf_bytes + s_phase[r] + (rx0 + im.pad_x - sx0) * 3 - a16;
    const int nvec = ((rx1 - rx0) * 3 + a16 + 15) >> 4;
    for (int i = tid; i < nvec; i += HP_THREADS) {
      const uint8_t* g = g0 + 16 * i;
      const long left = img_end - g;
      cp_async_16_partial(d0 + 16 * i, g, left >= 16 ? 16 : (left > 0 ? (int)left : 0));
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (needs_fill || n_prims != 0) {
    for (int r = 0; r < nrows; ++r) {
      const int sy = y0 + r, yr = sy - im.pad_y;
      const bool row_real = yr >= 0 && yr < im.H;
      uint8_t* dst = s_rows + r * a.row_buf_bytes + s_phase[r];
      for (int px = tid; px < sx1 - sx0; px += HP_THREADS) {
        const int xr = sx0 + px - im.pad_x;
        uint8_t* q = dst + px * 3;
        if (!(row_real && xr >= 0 && xr < im.W)) {      // canvas padding (expand2square)
          q[0] = (uint8_t)(im.bg & 0xff); q[1] = (uint8_t)((im.bg >> 8) & 0xff); q[2] = (uint8_t)((im.bg >> 16) & 0xff);
          continue;
        }
        if (n_prims == 0) continue;
        int rr = q[0], gg = q[1], bb = q[2];
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
        q[0] = (uint8_t)rr; q[1] = (uint8_t)gg; q[2] = (uint8_t)bb;
      }
    }
    __syncthreads();
  }
  // ---- horizontal filter ----
  const int4* hk = s_hkk4 + tid;
  uint32_t* orow = a.scratch + hv.offset + (size_t)y0 * hv.out_w + x;
  const int tap0 = (hx_min - sx0) * 3;
  for (int r = 0; r < nrows; ++r, orow += hv.out_w) {
    const uint8_t* cur = s_rows + r * a.row_buf_bytes;
    const int b0 = s_phase[r] + tap0;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(cur) + (b0 >> 2);
    const uint32_t sh = (uint32_t)(b0 & 3) * 8u;
    int s0 = 1 << (PREC - 1), s1 = s0, s2 = s0;
    uint32_t w0 = wp[0];
    for (int g = 0; g < ksh4; ++g) {
      const uint32_t w1 = wp[3 * g + 1], w2 = wp[3 * g + 2], w3 = wp[3 * g + 3];
      const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh), a2 = __funnelshift_r(w2, w3, sh);
      const int4 c = hk[g * HP_THREADS];
      s0 += byte_of<0>(a0) * c.x + byte_of<3>(a0) * c.y + byte_of<2>(a1) * c.z + byte_of<1>(a2) * c.w;
      s1 += byte_of<1>(a0) * c.x + byte_of<0>(a1) * c.y + byte_of<3>(a1) * c.z + byte_of<2>(a2) * c.w;
      s2 += byte_of<2>(a0) * c.x + byte_of<1>(a1) * c.y + byte_of<0>(a2) * c.z + byte_of<3>(a2) * c.w;
      w0 = w3;
    }
    if (valid) *orow = (uint32_t)clip8(s0) | ((uint32_t)clip8(s1) << 8) | ((uint32_t)clip8(s2) << 16);
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
  int max_ksize;
};

// Vertical pass + LUT + im2col of one 14-row band of one tile: the thread of output pixel x gathers, for
// each of the band's rows, exactly the taps of that row from the RGBX intermediate (coalesced 32-bit
// loads, L1 / L2 resident), no staging, no barriers inside the loop.
__global__ void __launch_bounds__(PP_THREADS) preprocess_v_kernel(const VArgs a) {
  extern __shared__ __align__(16) uint8_t pp_smem[];
  const int band = blockIdx.x, t = blockIdx.y, tid = threadIdx.x;
  const vz_tile_desc td = a.tiles[t];
  const vz_hview_desc hv = a.hviews[td.hview];
  const int32_t* tv = a.tables + td.tab_v;
  const int ksv = tv[0];
  const int32_t* v_min = tv + 2;
  const int32_t* v_cnt = v_min + td.out_h;
  const int32_t* v_kk = v_cnt + td.out_h;
  const int stage_bytes = (a.out_mode == VZ_OUT_CHW_F32) ? 3 * BAND * TILE * 4 : 24 * VZ_PATCH_K * 2;
  uint8_t* s_stage = pp_smem;
  int32_t* s_vkk = reinterpret_cast<int32_t*>(pp_smem + stage_bytes);   // [BAND][max_ksize]
  int32_t* s_vmin = s_vkk + BAND * a.max_ksize;
  int32_t* s_vcnt = s_vmin + BAND;
  const int ry0 = td.tile_y + band * BAND - td.off_y;
  if (tid < BAND) {
    const int ry = ry0 + tid;
    const bool ok = ry >= 0 && ry < td.out_h;
    s_vmin[tid] = ok ? v_min[ry] : 0;
    s_vcnt[tid] = ok ? v_cnt[ry] : 0;
  }
  for (int i = tid; i < BAND * ksv; i += PP_THREADS) {
    const int y = i / ksv, k = i - y * ksv;
    const int ry = ry0 + y;
    s_vkk[y * a.max_ksize + k] = (ry >= 0 && ry < td.out_h) ? v_kk[ry * ksv + k] : 0;
  }
  if (a.out_mode == VZ_OUT_PATCHES_BF16) {
    __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(s_stage);
    for (int i = tid; i < 24 * 4; i += PP_THREADS) sp[(i >> 2) * VZ_PATCH_K + 588 + (i & 3)] = __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  const int x = tid;
  const int rx = td.tile_x + x - td.off_x;
  const bool col_ok = tid < TILE && rx >= 0 && rx < td.out_w;
  const uint32_t* col = a.scratch + hv.offset + (col_ok ? rx : 0);
  const int px = x / 14, kx = x - px * 14;
#pragma unroll 2
  for (int y = 0; y < BAND; ++y) {
    const int cnt = col_ok ? s_vcnt[y] : 0;
    const uint32_t* src = col + (size_t)s_vmin[y] * hv.out_w;
    const int32_t* kk = s_vkk + y * a.max_ksize;
    int a0 = 1 << (PREC - 1), a1 = a0, a2 = a0;
    for (int k = 0; k < cnt; ++k) {
      const uint32_t p = __ldg(src + (size_t)k * hv.out_w);
      const int c = kk[k];
      a0 += byte_of<0>(p) * c;
      a1 += byte_of<1>(p) * c;
      a2 += byte_of<2>(p) * c;
    }
    if (tid < TILE) {
      const int v0 = cnt > 0 ? clip8(a0) : 0, v1 = cnt > 0 ? clip8(a1) : 0, v2 = cnt > 0 ? clip8(a2) : 0;
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
    for (int i = tid; i < 24 * VZ_PATCH_K * 2 / 16; i += PP_THREADS) dst[i] = s4[i];
  } else {
    const float* sf = reinterpret_cast<const float*>(s_stage);
    float* o = reinterpret_cast<float*>(a.out);
    for (int i = tid; i < 3 * BAND * TILE / 4; i += PP_THREADS) {
      const int e = i * 4;
      const int c = e / (BAND * TILE), rem = e - c * BAND * TILE;
      const int y = rem / TILE, xx = rem - y * TILE;
      *reinterpret_cast<float4*>(o + (((size_t)t * 3 + c) * TILE + band * BAND + y) * TILE + xx) =
          *reinterpret_cast<const float4*>(sf + e);
    }
  }
}

}  // namespace
}  // namespace vz

extern "C" int vz_preprocess(const vz_image_desc* images, int n_images, const vz_prim* prims,
                                int n_prims, const vz_tile_desc* tiles, int n_tiles,
                                const int32_t* tables, const float* lut768, int out_mode, void* out,
                                int max_src_w, int max_ksize, void* stream) {
  using namespace vz;
  if (!images || !tiles || !tables || !lut768 || !out || n_images <= 0 || n_tiles <= 0) return VZ_ERR_BAD_ARG;
  if (n_prims > 0 && !prims) return VZ_ERR_BAD_ARG;
  if (out_mode != VZ_OUT_PATCHES_BF16 && out_mode != VZ_OUT_CHW_F32) return VZ_ERR_BAD_ARG;
  if (max_src_w <= 0 || max_ksize <= 0 || !aligned16(out)) return VZ_ERR_BAD_ARG;
  PreArgs a;
  a.images = images; a.prims = prims; a.tiles = tiles; a.tables = tables; a.lut = lut768;
  a.out = out; a.out_mode = out_mode;
  a.row_buf_bytes = ((max_src_w * 3 + 3 * max_ksize + 32 + 15) / 16) * 16;  // + phase + tap overrun slack
  a.max_ksize = max_ksize;
  const int stage_bytes = (out_mode == VZ_OUT_CHW_F32) ? 3 * BAND * TILE * 4 : 24 * VZ_PATCH_K * 2;
  // a band of 14 output rows spans at most 14 * scale + ksize <= 3.4 * ksize source rows (ksize = 2 ceil(3 scale) + 1)
  const size_t vt_rows = 4 * (size_t)max_ksize + 16;
  const size_t smem = (size_t)stage_bytes + NBUF * (size_t)a.row_buf_bytes + vt_rows * VT_STRIDE * 4 + 2 * BAND * 4 + 16 +
                      (size_t)((max_ksize + 3) / 4) * TILE * 16;
  if (smem > 220 * 1024) return VZ_ERR_UNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VZ_ENSURE_DYN_SMEM(preprocess_kernel, 220 * 1024);
  dim3 grid(24, n_tiles);
  {
    ProfScope prof(VZ_PROF_PRE_FUSED, 0.0, st);
    preprocess_kernel<<<grid, PP_THREADS, smem, st>>>(a);
  }
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

extern "C" int vz_preprocess2(const vz_image_desc* images, int n_images, const vz_prim* prims, int n_prims,
                              const vz_hview_desc* hviews, int n_hviews, const vz_tile_desc* tiles, int n_tiles,
                              const int32_t* tables, const float* lut768, int out_mode, void* out, void* scratch,
                              long long scratch_pixels, int max_span_px, int max_rows, int max_out_w, int max_ksize,
                              void* stream) {
  using namespace vz;
  if (!images || !hviews || !tiles || !tables || !lut768 || !out || !scratch) return VZ_ERR_BAD_ARG;
  if (n_images <= 0 || n_hviews <= 0 || n_tiles <= 0 || scratch_pixels <= 0) return VZ_ERR_BAD_ARG;
  if (n_prims > 0 && !prims) return VZ_ERR_BAD_ARG;
  if (out_mode != VZ_OUT_PATCHES_BF16 && out_mode != VZ_OUT_CHW_F32) return VZ_ERR_BAD_ARG;
  if (max_span_px <= 0 || max_rows <= 0 || max_out_w <= 0 || max_ksize <= 0 || !aligned16(out) || !aligned16(scratch))
    return VZ_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // ---- horizontal pass ----
  HArgs h;
  h.images = images; h.prims = prims; h.hviews = hviews; h.tables = tables;
  h.scratch = reinterpret_cast<uint32_t*>(scratch);
  h.row_buf_bytes = ((max_span_px * 3 + 3 * max_ksize + 48 + 15) / 16) * 16;   // + phase + tap overrun slack
  h.max_ksize = max_ksize;
  const size_t smem_h = (size_t)((max_ksize + 3) / 4) * HP_THREADS * 16 + (size_t)HP_ROWS * h.row_buf_bytes;
  if (smem_h > 200 * 1024) return VZ_ERR_UNSUPPORTED;
  VZ_ENSURE_DYN_SMEM(preprocess_h_kernel, 200 * 1024);
  dim3 grid_h((max_out_w + HP_THREADS - 1) / HP_THREADS, (max_rows + HP_ROWS - 1) / HP_ROWS, n_hviews);
  {
    ProfScope prof(VZ_PROF_PRE_H, 0.0, st);
    preprocess_h_kernel<<<grid_h, HP_THREADS, smem_h, st>>>(h);
  }
  VZ_LAUNCH_CHECK();
  // ---- vertical pass + normalise + patchify ----
  VArgs v;
  v.hviews = hviews; v.tiles = tiles; v.tables = tables; v.lut = lut768;
  v.scratch = reinterpret_cast<const uint32_t*>(scratch); v.out = out; v.out_mode = out_mode; v.max_ksize = max_ksize;
  const int stage_bytes = (out_mode == VZ_OUT_CHW_F32) ? 3 * BAND * TILE * 4 : 24 * VZ_PATCH_K * 2;
  const size_t smem_v = (size_t)stage_bytes + (size_t)BAND * max_ksize * 4 + 2 * BAND * 4 + 16;
  if (smem_v > 200 * 1024) return VZ_ERR_UNSUPPORTED;
  VZ_ENSURE_DYN_SMEM(preprocess_v_kernel, 200 * 1024);
  dim3 grid_v(24, n_tiles);
  {
    ProfScope prof(VZ_PROF_PRE_V, 0.0, st);
    preprocess_v_kernel<<<grid_v, PP_THREADS, smem_v, st>>>(v);
  }
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}
