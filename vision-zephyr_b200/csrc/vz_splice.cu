// vz_splice.cu -- subsystem (4): anyres merge (unpad / image_newline) + token splice.
//
// Replaces the Python loops (with their .sum()/.tolist() host syncs) of
//   vis_zephyr_arch.py:157-195  text rows for the projector's conditioning
//   vis_zephyr_arch.py:214-305  strip padding, split at IMAGE_TOKEN_INDEX, interleave features
//   vis_zephyr_arch.py:396-473  _process_image_patches (flat / spatial / spatial_unpad)
//   vis_zephyr_arch.py:476-530  _pad_and_collate_multimodal_inputs
// with (a) ONE planning CTA that turns masks and image-token positions into destination rows via
// warp ballots + prefix sums, and (b) one bandwidth kernel that writes every output row (text row,
// visual row, newline row or padding) exactly once with 128-bit copies; the merge permutation is
// folded into the source-row index, so the merged feature list is never materialised.
#include "vz_common.cuh"

namespace vz {
namespace {

constexpr int PLAN_THREADS = 1024;
constexpr int PLAN_WARPS = PLAN_THREADS / 32;
constexpr int MAX_B_SMEM = 2048;

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PLAN_THREADS)
splice_plan_kernel(const int64_t* __restrict__ ids, const uint8_t* __restrict__ mask, int B, int S,
                   const vz_slot_desc* __restrict__ slots, int n_slots, int max_len,
                   int32_t* __restrict__ tok_dest, int32_t* __restrict__ slot_dest,
                   int32_t* __restrict__ lengths, int32_t* __restrict__ text_len,
                   int32_t* __restrict__ totals) {
  __shared__ int s_nimg[MAX_B_SMEM];   // image tokens kept per sample
  __shared__ int s_base[MAX_B_SMEM];   // first slot of the sample (exclusive prefix)
  __shared__ int s_red[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (int i = threadIdx.x; i < n_slots * 2; i += PLAN_THREADS) slot_dest[i] = -1;
  if (threadIdx.x < 4) s_red[threadIdx.x] = 0;

  // pass 1: per-sample counts (warp per sample)
  for (int b = warp; b < B; b += PLAN_WARPS) {
    int n_img = 0, n_txt_all = 0;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      const bool in = s < S;
      const int64_t id = in ? ids[(size_t)b * S + s] : 0;
      const bool keep = in && (mask ? mask[(size_t)b * S + s] != 0 : true);
      n_img += __popc(__ballot_sync(0xffffffffu, keep && id == VZ_IMAGE_TOKEN_INDEX));
      n_txt_all += __popc(__ballot_sync(0xffffffffu, in && id != VZ_IMAGE_TOKEN_INDEX));
    }
    if (lane == 0) { s_nimg[b] = n_img; text_len[b] = n_txt_all; }
  }
  __syncthreads();
  // slot base per sample: a sample without image token still consumes one slot
  // (vis_zephyr_arch.py:245-258); warp 0 scans the samples 32 at a time.
  if (warp == 0) {
    int running = 0, max_txt = 0, sum_txt = 0;
    for (int b0 = 0; b0 < B; b0 += 32) {
      const int b = b0 + lane;
      const int c = (b < B) ? (s_nimg[b] > 0 ? s_nimg[b] : 1) : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
      if (b < B) s_base[b] = running + incl - c;
      running += __shfl_sync(0xffffffffu, incl, 31);
      const int tl = (b < B) ? text_len[b] : 0;
      int mx = tl, sm = tl;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
      }
      max_txt = max(max_txt, mx);
      sum_txt += sm;
    }
    if (lane == 0) { totals[1] = max_txt; totals[2] = running; totals[3] = sum_txt; }
  }
  __syncthreads();
  // pass 2: destinations
  for (int b = warp; b < B; b += PLAN_WARPS) {
    int run = 0;      // rows emitted so far (text + visual)
    int img_seen = 0; // image tokens seen so far
    const int base = s_base[b];
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      const bool in = s < S;
      const int64_t id = in ? ids[(size_t)b * S + s] : 0;
      const bool keep = in && (mask ? mask[(size_t)b * S + s] != 0 : true);
      const bool is_img = keep && id == VZ_IMAGE_TOKEN_INDEX;
      const uint32_t bt = __ballot_sync(0xffffffffu, keep && !is_img);
      uint32_t bi = __ballot_sync(0xffffffffu, is_img);
      // rows contributed by image tokens at lower lanes (rare: loop over set bits)
      int vis_before = 0, vis_total = 0, imgs = 0;
      uint32_t w = bi;
      while (w) {
        const int l = __ffs(w) - 1;
        w &= w - 1;
        const int slot = base + img_seen + imgs;
        const int n = (slot < n_slots) ? slots[slot].n_rows : 0;
        if (lane == l) {
          // this lane is the image token: record where its slot starts
          const int start = run + __popc(bt & lt_mask) + vis_before;
          if (slot < n_slots) { slot_dest[slot * 2] = b; slot_dest[slot * 2 + 1] = start; }
        }
        if (l < lane) vis_before += n;
        vis_total += n;
        ++imgs;
      }
      int dest = -1;
      if (keep && !is_img) {
        dest = run + __popc(bt & lt_mask) + vis_before;
        if (max_len > 0 && dest >= max_len) dest = -1;
      }
      if (in) tok_dest[(size_t)b * S + s] = dest;
      run += __popc(bt) + vis_total;
      img_seen += imgs;
    }
    if (lane == 0) {
      int len = run;
      if (max_len > 0 && len > max_len) len = max_len;
      lengths[b] = len;
      atomicMax(&s_red[0], len);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) totals[0] = s_red[0];
}

// ------------------------------------------------------------------------------------------
// text rows for the projector conditioning
// ------------------------------------------------------------------------------------------
__global__ void text_off_kernel(const int32_t* __restrict__ text_len, int B, int32_t* __restrict__ text_off) {
  if (threadIdx.x == 0) {
    int r = 0;
    for (int b = 0; b < B; ++b) { text_off[b] = r; r += text_len[b]; }
    text_off[B] = r;
  }
}

// grid = (ceil(S/32), B), block = 256 (8 warps x 4 tokens). Each warp copies whole rows.
__global__ void __launch_bounds__(256)
text_gather_kernel(const int64_t* __restrict__ ids, int B, int S, const uint4* __restrict__ table,
                   int vec_per_row, const int32_t* __restrict__ text_off, uint4* __restrict__ out) {
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s_blk = blockIdx.x * 32;
  // rank of each token of this 32-token window among the sample's non-image tokens
  // = (# non-image tokens before the window) + popc within the window
  __shared__ int s_before;
  if (threadIdx.x < 32) {
    int cnt = 0;
    for (int s0 = 0; s0 < s_blk; s0 += 32) {
      const int64_t id = ids[(size_t)b * S + s0 + lane];
      cnt += __popc(__ballot_sync(0xffffffffu, id != VZ_IMAGE_TOKEN_INDEX));
    }
    if (lane == 0) s_before = cnt;
  }
  __syncthreads();
  const int s = s_blk + lane;
  const bool in = s < S;
  const int64_t my_id = in ? ids[(size_t)b * S + s] : VZ_IMAGE_TOKEN_INDEX;
  const uint32_t bal = __ballot_sync(0xffffffffu, in && my_id != VZ_IMAGE_TOKEN_INDEX);
  const int base = text_off[b] + s_before;
  for (int k = warp; k < 32; k += 8) {
    if (!((bal >> k) & 1u)) continue;
    const int64_t id = __shfl_sync(0xffffffffu, my_id, k);
    const int dst = base + __popc(bal & ((1u << k) - 1u));
    const uint4* src = table + (size_t)id * vec_per_row;
    uint4* d = out + (size_t)dst * vec_per_row;
    for (int i = lane; i < vec_per_row; i += 32) d[i] = __ldg(src + i);
  }
  // the single zero row that stands for every zero-padded text position
  if (blockIdx.x == 0 && b == 0) {
    uint4* d = out + (size_t)text_off[B] * vec_per_row;
    for (int i = threadIdx.x; i < vec_per_row; i += blockDim.x) d[i] = make_uint4(0, 0, 0, 0);
  }
}

// ------------------------------------------------------------------------------------------
// merge index: merged row r of a slot -> source row in the projector output, or -1 = newline
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int merged_source_row(const vz_slot_desc& sd, int r) {
  switch (sd.merge) {
    case VZ_MERGE_FLAT:
      return sd.row_base + r;
    case VZ_MERGE_SINGLE_NEWLINE:
      return r < sd.hw ? sd.row_base + r : -1;
    case VZ_MERGE_SPATIAL: {
      if (r < sd.hw) return sd.row_base + r;
      const int q = r - sd.hw;
      const int Wt = sd.n_w * sd.w;
      const int y = q / Wt, x = q - y * Wt;
      const int ty = y / sd.h, iy = y - ty * sd.h, tx = x / sd.w, ix = x - tx * sd.w;
      return sd.row_base + (1 + ty * sd.n_w + tx) * sd.hw + iy * sd.w + ix;
    }
    default: {  // VZ_MERGE_SPATIAL_UNPAD
      if (r < sd.hw) return sd.row_base + r;
      const int q = r - sd.hw;
      const int cw = sd.x1 - sd.x0 + 1;  // cropped width + newline column
      const int yy = q / cw, xx = q - yy * cw;
      if (xx == cw - 1) return -1;
      const int y = sd.y0 + yy, x = sd.x0 + xx;
      const int ty = y / sd.h, iy = y - ty * sd.h, tx = x / sd.w, ix = x - tx * sd.w;
      return sd.row_base + (1 + ty * sd.n_w + tx) * sd.hw + iy * sd.w + ix;
    }
  }
}

struct ScatterArgs {
  const int64_t* ids; const int64_t* labels; int B, S;
  const uint4* table; const uint4* vis; int ldv_vec; const uint4* newline; int vec_per_row;
  const vz_slot_desc* slots; int n_slots; const int32_t* slot_prefix;  // [n_slots+1] exclusive prefix of n_rows
  const int32_t* tok_dest; const int32_t* slot_dest; const int32_t* lengths;
  int Lout, pad_left;
  uint4* out_embeds; int64_t* out_labels; uint8_t* out_mask; int64_t* out_pos;
  int total_vis_rows;  // sum n_rows
  float2* row_stats;   // optional [B*Lout]: (0, sum of squares) of every output row (bf16 rows only)
};

__device__ __forceinline__ void copy_row(uint4* dst, const uint4* src, int n, int lane) {
  // 4 x 128-bit loads in flight per lane
  int i = lane;
  for (; i + 96 < n; i += 128) {
    const uint4 a = __ldg(src + i), b = __ldg(src + i + 32), c = __ldg(src + i + 64), d = __ldg(src + i + 96);
    dst[i] = a; dst[i + 32] = b; dst[i + 64] = c; dst[i + 96] = d;
  }
  for (; i < n; i += 32) dst[i] = __ldg(src + i);
}

// same copy, also returning the row's sum of squares (bf16 elements; identical on every lane): the statistic the
// first LLM layer's RMSNorm needs, taken while the row passes through the registers anyway
__device__ __forceinline__ float copy_row_sumsq(uint4* dst, const uint4* src, int n, int lane) {
  float s = 0.f;
  for (int i = lane; i < n; i += 32) {
    const uint4 w = __ldg(src + i);
    dst[i] = w;
    const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float lo = __uint_as_float(u[k] << 16), hi = __uint_as_float(u[k] & 0xffff0000u);
      s = fmaf(lo, lo, s);
      s = fmaf(hi, hi, s);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// One warp per work item; items = [B*S tokens] ++ [visual rows of all slots] ++ [B*Lout pad probes].
__global__ void __launch_bounds__(256) splice_scatter_kernel(const ScatterArgs a) {
  const int lane = threadIdx.x & 31;
  const long warp_global = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  const long n_tok = (long)a.B * a.S;
  const long n_vis = a.total_vis_rows;
  const long n_pad = (long)a.B * a.Lout;
  const long total = n_tok + n_vis + n_pad;
  const uint4 zero = make_uint4(0, 0, 0, 0);

  for (long item = warp_global; item < total; item += nwarps) {
    if (item < n_tok) {
      const int d = a.tok_dest[item];
      if (d < 0) continue;
      const int b = (int)(item / a.S);
      const int len = a.lengths[b];
      const int off = a.pad_left ? a.Lout - len : 0;
      const long orow = (long)b * a.Lout + off + d;
      if (a.row_stats) {
        const float ss = copy_row_sumsq(a.out_embeds + orow * a.vec_per_row, a.table + (size_t)a.ids[item] * a.vec_per_row,
                                        a.vec_per_row, lane);
        if (lane == 0) a.row_stats[orow] = make_float2(0.f, ss);
      } else {
        copy_row(a.out_embeds + orow * a.vec_per_row, a.table + (size_t)a.ids[item] * a.vec_per_row, a.vec_per_row, lane);
      }
      if (lane == 0) {
        a.out_labels[orow] = a.labels ? a.labels[item] : (int64_t)VZ_IGNORE_INDEX;
        a.out_mask[orow] = 1;
        a.out_pos[orow] = d;
      }
    } else if (item < n_tok + n_vis) {
      // find slot by binary search over the running row prefix
      const int v = (int)(item - n_tok);
      int lo = 0, hi = a.n_slots - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.slot_prefix[mid] <= v) lo = mid; else hi = mid - 1;
      }
      const int slot = lo;
      const int r = v - a.slot_prefix[slot];
      const int b = a.slot_dest[slot * 2];
      if (b < 0) continue;  // slot not consumed by any image token
      const int d = a.slot_dest[slot * 2 + 1] + r;
      const int len = a.lengths[b];
      if (d >= len) continue;  // truncated
      const vz_slot_desc sd = a.slots[slot];
      const int srow = merged_source_row(sd, r);
      const int off = a.pad_left ? a.Lout - len : 0;
      const long orow = (long)b * a.Lout + off + d;
      const uint4* src = srow >= 0 ? a.vis + (size_t)srow * a.ldv_vec : a.newline;
      if (a.row_stats) {
        const float ss = copy_row_sumsq(a.out_embeds + orow * a.vec_per_row, src, a.vec_per_row, lane);
        if (lane == 0) a.row_stats[orow] = make_float2(0.f, ss);
      } else {
        copy_row(a.out_embeds + orow * a.vec_per_row, src, a.vec_per_row, lane);
      }
      if (lane == 0) {
        a.out_labels[orow] = (int64_t)VZ_IGNORE_INDEX;
        a.out_mask[orow] = 1;
        a.out_pos[orow] = d;
      }
    } else {
      const long p = item - n_tok - n_vis;
      const int b = (int)(p / a.Lout), jrow = (int)(p - (long)b * a.Lout);
      const int len = a.lengths[b];
      const bool is_pad = a.pad_left ? (jrow < a.Lout - len) : (jrow >= len);
      if (!is_pad) continue;
      uint4* dst = a.out_embeds + p * a.vec_per_row;
      for (int i = lane; i < a.vec_per_row; i += 32) dst[i] = zero;
      if (lane == 0) {
        a.out_labels[p] = (int64_t)VZ_IGNORE_INDEX;
        a.out_mask[p] = 0;
        a.out_pos[p] = 0;
        if (a.row_stats) a.row_stats[p] = make_float2(0.f, 0.f);
      }
    }
  }
}

// merge only: out row (out_row_base[slot] + r) = merged row r of slot
__global__ void __launch_bounds__(256)
merge_rows_kernel(const uint4* __restrict__ vis, int ldv_vec, const uint4* __restrict__ newline, int vec_per_row,
                  const vz_slot_desc* __restrict__ slots, int n_slots, const int32_t* __restrict__ out_row_base,
                  uint4* __restrict__ out) {
  const int slot = blockIdx.y;
  const vz_slot_desc sd = slots[slot];
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < sd.n_rows; r += gridDim.x * 8) {
    const int srow = merged_source_row(sd, r);
    const uint4* src = srow >= 0 ? vis + (size_t)srow * ldv_vec : newline;
    copy_row(out + (size_t)(out_row_base[slot] + r) * vec_per_row, src, vec_per_row, lane);
  }
}

}  // namespace
}  // namespace vz

using namespace vz;

extern "C" int vz_splice_plan(const int64_t* input_ids, const uint8_t* mask, int B, int S,
                              const vz_slot_desc* slots, int n_slots, int max_len, int32_t* tok_dest,
                              int32_t* slot_dest, int32_t* lengths, int32_t* text_len, int32_t* totals,
                              void* stream) {
  if (!input_ids || !tok_dest || !slot_dest || !lengths || !text_len || !totals) return VZ_ERR_BAD_ARG;
  if (B <= 0 || S <= 0 || n_slots < 0 || (n_slots > 0 && !slots)) return VZ_ERR_BAD_ARG;
  if (B > MAX_B_SMEM) return VZ_ERR_UNSUPPORTED;
  splice_plan_kernel<<<1, PLAN_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      input_ids, mask, B, S, slots, n_slots, max_len, tok_dest, slot_dest, lengths, text_len, totals);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

extern "C" int vz_text_gather(const int64_t* input_ids, int B, int S, const void* embed_table, int D,
                              int elem_bytes, const int32_t* text_len, void* text_emb, int32_t* text_off,
                              void* stream) {
  if (!input_ids || !embed_table || !text_len || !text_emb || !text_off || B <= 0 || S <= 0) return VZ_ERR_BAD_ARG;
  const long row_bytes = (long)D * elem_bytes;
  if (row_bytes % 16 || !aligned16(embed_table) || !aligned16(text_emb)) return VZ_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  text_off_kernel<<<1, 32, 0, st>>>(text_len, B, text_off);
  VZ_LAUNCH_CHECK();
  dim3 grid((S + 31) / 32, B);
  ProfScope prof(VZ_PROF_TEXT_GATHER, 0.0, st);
  text_gather_kernel<<<grid, 256, 0, st>>>(input_ids, B, S, reinterpret_cast<const uint4*>(embed_table),
                                           (int)(row_bytes / 16), text_off, reinterpret_cast<uint4*>(text_emb));
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

extern "C" int vz_splice_scatter(const int64_t* input_ids, const int64_t* labels, int B, int S,
                                 const void* embed_table, const void* vis, int ldv, const void* image_newline,
                                 int D, int elem_bytes, const vz_slot_desc* slots, int n_slots,
                                 const int32_t* slot_prefix, int total_vis_rows, const int32_t* tok_dest,
                                 const int32_t* slot_dest, const int32_t* lengths, int Lout, int pad_left,
                                 void* out_embeds, int64_t* out_labels, uint8_t* out_mask, int64_t* out_pos,
                                 void* stream) {
  return vz_splice_scatter_rms(input_ids, labels, B, S, embed_table, vis, ldv, image_newline, D, elem_bytes, slots,
                               n_slots, slot_prefix, total_vis_rows, tok_dest, slot_dest, lengths, Lout, pad_left,
                               out_embeds, out_labels, out_mask, out_pos, nullptr, stream);
}

extern "C" int vz_splice_scatter_rms(const int64_t* input_ids, const int64_t* labels, int B, int S,
                                     const void* embed_table, const void* vis, int ldv, const void* image_newline,
                                     int D, int elem_bytes, const vz_slot_desc* slots, int n_slots,
                                     const int32_t* slot_prefix, int total_vis_rows, const int32_t* tok_dest,
                                     const int32_t* slot_dest, const int32_t* lengths, int Lout, int pad_left,
                                     void* out_embeds, int64_t* out_labels, uint8_t* out_mask, int64_t* out_pos,
                                     float* out_row_stats, void* stream) {
  if (out_row_stats && (elem_bytes != 2 || (reinterpret_cast<uintptr_t>(out_row_stats) & 7u))) return VZ_ERR_BAD_ARG;
  if (!input_ids || !embed_table || !tok_dest || !slot_dest || !lengths || !out_embeds || !out_labels ||
      !out_mask || !out_pos)
    return VZ_ERR_BAD_ARG;
  if (B <= 0 || S <= 0 || Lout <= 0 || n_slots < 0) return VZ_ERR_BAD_ARG;
  if (n_slots > 0 && (!slots || !slot_prefix || !vis)) return VZ_ERR_BAD_ARG;
  const long row_bytes = (long)D * elem_bytes;
  if (row_bytes % 16 || ((long)ldv * elem_bytes) % 16) return VZ_ERR_BAD_ARG;
  if (!aligned16(embed_table) || !aligned16(out_embeds) || (vis && !aligned16(vis)) ||
      (image_newline && !aligned16(image_newline)))
    return VZ_ERR_BAD_ARG;
  ScatterArgs a;
  a.ids = input_ids; a.labels = labels; a.B = B; a.S = S;
  a.table = reinterpret_cast<const uint4*>(embed_table);
  a.vis = reinterpret_cast<const uint4*>(vis);
  a.ldv_vec = (int)((long)ldv * elem_bytes / 16);
  a.newline = reinterpret_cast<const uint4*>(image_newline);
  a.vec_per_row = (int)(row_bytes / 16);
  a.slots = slots; a.n_slots = n_slots; a.slot_prefix = slot_prefix;
  a.tok_dest = tok_dest; a.slot_dest = slot_dest; a.lengths = lengths;
  a.Lout = Lout; a.pad_left = pad_left;
  a.out_embeds = reinterpret_cast<uint4*>(out_embeds);
  a.out_labels = out_labels; a.out_mask = out_mask; a.out_pos = out_pos;
  a.total_vis_rows = n_slots > 0 ? total_vis_rows : 0;
  a.row_stats = reinterpret_cast<float2*>(out_row_stats);
  const long items = (long)B * S + a.total_vis_rows + (long)B * Lout;
  int dev = 0, sms = 148;
  VZ_CUDA_CHECK(cudaGetDevice(&dev));
  VZ_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long blocks = (items + 7) / 8;
  const long cap = (long)sms * 8;  // 8 CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  ProfScope prof(VZ_PROF_SPLICE_SCATTER, 0.0, reinterpret_cast<cudaStream_t>(stream));
  splice_scatter_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

extern "C" int vz_merge_rows(const void* vis, int ldv, const void* image_newline, int D, int elem_bytes,
                             const vz_slot_desc* slots, int n_slots, const int32_t* out_row_base, void* out,
                             void* stream) {
  if (!vis || !slots || !out_row_base || !out || n_slots <= 0) return VZ_ERR_BAD_ARG;
  const long row_bytes = (long)D * elem_bytes;
  if (row_bytes % 16 || ((long)ldv * elem_bytes) % 16) return VZ_ERR_BAD_ARG;
  dim3 grid(64, n_slots);
  merge_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(vis), (int)((long)ldv * elem_bytes / 16),
      reinterpret_cast<const uint4*>(image_newline), (int)(row_bytes / 16), slots, n_slots, out_row_base,
      reinterpret_cast<uint4*>(out));
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}

// ------------------------------------------------------------------------------------------
// collate (next row of the scope table): DataCollatorForSupervisedDataset, train/train.py:657-707
// pad_sequence(input_ids, pad_token_id), pad_sequence(labels, IGNORE_INDEX), truncate to
// model_max_length, attention_mask = input_ids.ne(pad_token_id) -- from packed ragged rows.
// ------------------------------------------------------------------------------------------
namespace vz {
namespace {
__global__ void __launch_bounds__(256)
collate_kernel(const int64_t* __restrict__ flat_ids, const int64_t* __restrict__ flat_labels,
               const int32_t* __restrict__ offsets, int S_out, int64_t pad_id, int64_t* __restrict__ out_ids,
               int64_t* __restrict__ out_labels, uint8_t* __restrict__ out_mask) {
  const int b = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S_out) return;
  const int off = offsets[b], len = offsets[b + 1] - off;
  const bool in = s < len;
  const int64_t id = in ? flat_ids[off + s] : pad_id;
  const size_t o = (size_t)b * S_out + s;
  out_ids[o] = id;
  out_labels[o] = in ? flat_labels[off + s] : (int64_t)VZ_IGNORE_INDEX;
  out_mask[o] = id != pad_id;  // like the reference: a real token equal to the pad id is masked too
}
}  // namespace
}  // namespace vz

extern "C" int vz_collate(const int64_t* flat_ids, const int64_t* flat_labels, const int32_t* offsets, int B,
                          int S_out, int64_t pad_id, int64_t* out_ids, int64_t* out_labels, uint8_t* out_mask,
                          void* stream) {
  if (!flat_ids || !flat_labels || !offsets || !out_ids || !out_labels || !out_mask || B <= 0 || S_out <= 0)
    return VZ_ERR_BAD_ARG;
  dim3 grid((S_out + 255) / 256, B);
  vz::collate_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(flat_ids, flat_labels, offsets, S_out,
                                                                              pad_id, out_ids, out_labels, out_mask);
  VZ_LAUNCH_CHECK();
  return VZ_OK;
}
