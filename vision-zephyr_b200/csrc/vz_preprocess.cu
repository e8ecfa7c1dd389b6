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
// through a double-buffered shared-memory row buffer (blended on the way in); every thread owns 4
// of the band's 1008 byte-columns, runs the horizontal pass for them and folds the result straight
// into 14 fixed-point vertical accumulators held in registers, so neither the blended image nor the
// u8 intermediate ever touches HBM.  The band is then normalised through the LUT, transposed into
// im2col order in shared memory and written with 128-bit stores.
#include "vz_common.cuh"

namespace vz {
namespace {

constexpr int TILE = 336;
constexpr int BAND = 14;
constexpr int NCOL = TILE * 3;  // 1008 byte-columns per band row
constexpr int PP_THREADS = 256;
constexpr int COLS_PER_THREAD = 4;  // 4*256 >= 1008
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

__global__ void __launch_bounds__(PP_THREADS) preprocess_kernel(const PreArgs a) {
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
  uint8_t* s_stage = pp_smem;
  uint8_t* s_row = pp_smem + stage_bytes;
  int32_t* s_vkk = reinterpret_cast<int32_t*>(s_row + 2 * a.row_buf_bytes);  // [BAND][ksv]
  int32_t* s_vmin = s_vkk + BAND * a.max_ksize;
  int32_t* s_vcnt = s_vmin + BAND;
  int32_t* s_hkk = s_vcnt + BAND + 4;  // [ksh][336]: tap k of output column x (zero outside the taps)

  // ---- instance list of this image (visual prompts) -> shared memory ---------------------------
  __shared__ vz_prim s_prims[MAX_PRIMS];
  const int n_prims = im.prim_count < MAX_PRIMS ? im.prim_count : MAX_PRIMS;
  for (int i = tid; i < n_prims; i += PP_THREADS) s_prims[i] = a.prims[im.prim_begin + i];

  // ---- vertical taps of the band's 14 output rows ------------------------------------------
  const int ry0 = td.tile_y + band * BAND - td.off_y;  // resized-image row of band row 0
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
  // ---- horizontal coefficients of the tile's 336 columns, tap-major (conflict-free reads) ---------
  for (int i = tid; i < ksh * TILE; i += PP_THREADS) {
    const int k = i / TILE, x = i - k * TILE;
    const int rx = td.tile_x + x - td.off_x;
    s_hkk[i] = (rx >= 0 && rx < td.out_w) ? h_kk[rx * ksh + k] : 0;
  }
  // ---- this thread's columns ------------------------------------------------------------------
  int hx_min[COLS_PER_THREAD], hx_cnt[COLS_PER_THREAD], hx_c[COLS_PER_THREAD], hx_x[COLS_PER_THREAD];
#pragma unroll
  for (int i = 0; i < COLS_PER_THREAD; ++i) {
    const int j = tid + i * PP_THREADS;
    const int x = (j < NCOL) ? j / 3 : 0;
    hx_x[i] = x;
    hx_c[i] = (j < NCOL) ? j - x * 3 : 0;
    const int rx = td.tile_x + x - td.off_x;
    const bool ok = (j < NCOL) && rx >= 0 && rx < td.out_w;
    hx_min[i] = ok ? h_min[rx] : -1;
    hx_cnt[i] = ok ? h_cnt[rx] : 0;
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
#pragma unroll
  for (int i = 0; i < COLS_PER_THREAD; ++i)
    if (hx_min[i] < 0) hx_min[i] = sx0;  // invalid column: all-zero taps, any in-buffer address will do
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

  int acc[BAND][COLS_PER_THREAD];
#pragma unroll
  for (int y = 0; y < BAND; ++y)
#pragma unroll
    for (int i = 0; i < COLS_PER_THREAD; ++i) acc[y][i] = 1 << (PREC - 1);

  const int nbytes = (sx1 - sx0) * 3;
  const uint8_t* img_end = im.src + (size_t)im.W * im.H * 3;
  // Rows land in smem at the same 16-byte phase they have in global memory, so the copy is made of
  // aligned 128-bit loads; `row_phase(sy)` is where the first needed byte sits in the buffer.
  auto row_phase = [&](int sy) -> int {
    return (int)(reinterpret_cast<uintptr_t>(im.src + ((size_t)sy * im.W + sx0) * 3) & 15);
  };
  auto load_row = [&](int sy, uint8_t* dst) {
    const uint8_t* srow = im.src + ((size_t)sy * im.W + sx0) * 3;
    const int ph = (int)(reinterpret_cast<uintptr_t>(srow) & 15);
    if (im.prim_count == 0) {
      const uint8_t* g0 = srow - ph;
      const int nvec = (nbytes + ph + 15) >> 4;
      for (int i = tid; i < nvec; i += PP_THREADS) {
        const uint8_t* g = g0 + 16 * i;
        if (g + 16 <= img_end) {
          *reinterpret_cast<uint4*>(dst + 16 * i) = __ldg(reinterpret_cast<const uint4*>(g));
        } else {
          for (int b = 0; b < 16; ++b) dst[16 * i + b] = (g + b < img_end) ? g[b] : (uint8_t)0;
        }
      }
      return;
    }
    // blend path: one thread per PIXEL; the instance list sits in shared memory and every RGBA
    // overlay pixel is one aligned 32-bit load
    const int npx = sx1 - sx0;
    for (int px = tid; px < npx; px += PP_THREADS) {
      const int x = sx0 + px;
      int r = srow[px * 3], g = srow[px * 3 + 1], b = srow[px * 3 + 2];
      for (int pi = 0; pi < n_prims; ++pi) {
        const vz_prim& p = s_prims[pi];
        uint32_t ov;
        if (p.type == VZ_PRIM_LAYER) {
          ov = __ldg(reinterpret_cast<const uint32_t*>(im.layers) + ((size_t)p.layer * im.H + sy) * im.W + x);
        } else {
          if (!rect_covers(p, x, sy)) continue;
          ov = p.rgba;
        }
        const int al = (int)(ov >> 24);
        r = blend_over(r, (int)(ov & 0xff), al);
        g = blend_over(g, (int)((ov >> 8) & 0xff), al);
        b = blend_over(b, (int)((ov >> 16) & 0xff), al);
      }
      dst[ph + px * 3] = (uint8_t)r;
      dst[ph + px * 3 + 1] = (uint8_t)g;
      dst[ph + px * 3 + 2] = (uint8_t)b;
    }
  };

  if (sy1 > sy0) load_row(sy0, s_row);
  __syncthreads();
  for (int sy = sy0; sy < sy1; ++sy) {
    const uint8_t* cur = s_row + ((sy - sy0) & 1) * a.row_buf_bytes;
    if (sy + 1 < sy1) load_row(sy + 1, s_row + ((sy + 1 - sy0) & 1) * a.row_buf_bytes);
    // horizontal pass: uniform tap loop (taps beyond a column's count have zero coefficients), four
    // independent accumulators per thread
    int hv[COLS_PER_THREAD];
    {
      const uint8_t* base = cur + row_phase(sy);
      const uint8_t* sp0 = base + (hx_min[0] - sx0) * 3 + hx_c[0];
      const uint8_t* sp1 = base + (hx_min[1] - sx0) * 3 + hx_c[1];
      const uint8_t* sp2 = base + (hx_min[2] - sx0) * 3 + hx_c[2];
      const uint8_t* sp3 = base + (hx_min[3] - sx0) * 3 + hx_c[3];
      const int32_t* c0 = s_hkk + hx_x[0];
      const int32_t* c1 = s_hkk + hx_x[1];
      const int32_t* c2 = s_hkk + hx_x[2];
      const int32_t* c3 = s_hkk + hx_x[3];
      int s0 = 1 << (PREC - 1), s1 = s0, s2 = s0, s3 = s0;
#pragma unroll 4
      for (int k = 0; k < ksh; ++k) {
        s0 += (int)sp0[k * 3] * c0[k * TILE];
        s1 += (int)sp1[k * 3] * c1[k * TILE];
        s2 += (int)sp2[k * 3] * c2[k * TILE];
        s3 += (int)sp3[k * 3] * c3[k * TILE];
      }
      hv[0] = clip8(s0); hv[1] = clip8(s1); hv[2] = clip8(s2); hv[3] = clip8(s3);
    }
#pragma unroll
    for (int y = 0; y < BAND; ++y) {
      const int k = sy - s_vmin[y];
      if (k >= 0 && k < s_vcnt[y]) {
        const int coef = s_vkk[y * a.max_ksize + k];
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD; ++i) acc[y][i] += hv[i] * coef;
      }
    }
    __syncthreads();
  }

  // ---- normalise + transpose into the output layout -------------------------------------------
  if (a.out_mode == VZ_OUT_PATCHES_BF16) {
    __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(s_stage);
    for (int i = tid; i < 24 * 4; i += PP_THREADS) sp[(i >> 2) * VZ_PATCH_K + 588 + (i & 3)] = __float2bfloat16_rn(0.f);
#pragma unroll
    for (int i = 0; i < COLS_PER_THREAD; ++i) {
      const int j = tid + i * PP_THREADS;
      if (j >= NCOL) continue;
      const int x = j / 3, c = hx_c[i];
      const int px = x / 14, kx = x - px * 14;
#pragma unroll
      for (int y = 0; y < BAND; ++y) {
        const int v = (hx_cnt[i] > 0 && s_vcnt[y] > 0) ? clip8(acc[y][i]) : 0;
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
#pragma unroll
    for (int i = 0; i < COLS_PER_THREAD; ++i) {
      const int j = tid + i * PP_THREADS;
      if (j >= NCOL) continue;
      const int x = j / 3, c = hx_c[i];
#pragma unroll
      for (int y = 0; y < BAND; ++y) {
        const int v = (hx_cnt[i] > 0 && s_vcnt[y] > 0) ? clip8(acc[y][i]) : 0;
        sf[(c * BAND + y) * TILE + x] = a.lut[c * 256 + v];
      }
    }
    __syncthreads();
    float* o = reinterpret_cast<float*>(a.out);
    for (int i = tid; i < 3 * BAND * TILE / 4; i += PP_THREADS) {
      const int e = i * 4;
      const int c = e / (BAND * TILE), rem = e - c * BAND * TILE;
      const int y = rem / TILE, x = rem - y * TILE;
      *reinterpret_cast<float4*>(o + (((size_t)t * 3 + c) * TILE + band * BAND + y) * TILE + x) =
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
  const size_t smem = (size_t)stage_bytes + 2 * (size_t)a.row_buf_bytes + (size_t)BAND * max_ksize * 4 + 2 * BAND * 4 +
                      16 + (size_t)max_ksize * TILE * 4;
  if (smem > 220 * 1024) return VZ_ERR_UNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static bool attr_done = false;
  if (!attr_done) {
    VZ_CUDA_CHECK(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_done = true;
  }
  dim3 grid(24, n_tiles);
  preprocess_kernel<<<grid, PP_THREADS, smem, st>>>(a);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}
