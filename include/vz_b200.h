/*
 * vz_b200.h -- C ABI of the B200-native image -> LLM-embedding path of Vision-Zephyr.
 *
 * The reference (baohuyvanba/Vision-Zephyr) is pure Python: it has no FFI of its own.  The
 * drop-in boundary is therefore the four Python callables listed in SURVEY.md section 8(b); the
 * Python shim in vision-zephyr_b200/ keeps their signatures and calls the entry points below
 * through ctypes.  Every entry point names the reference code it replaces (file:line, relative
 * to the reference checkout).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - every pointer is a DEVICE pointer unless its name ends in _h (host).
 *   - every launch goes to the cudaStream_t passed last (as void*); no entry point allocates or
 *     synchronises, and every buffer an entry point writes is handed in by the caller (outputs, workspace,
 *     stream-K scratch).  Two calls may therefore run concurrently on two streams PROVIDED THEY ARE GIVEN
 *     DISJOINT WORKSPACES -- activations, LayerNorm statistics and the stream-K hand-over slots all live
 *     there.  The Python layer keys its workspaces by (device, stream) for exactly that reason
 *     (vision_tower.Workspace; reference threading note: vis_zephyr/serve/api.py:161-177).  Process-wide state
 *     inside the library is limited to monotonic tickets and caches that are safe to share: the launch
 *     counter, the stream-K epoch counter (an atomic ticket, unique per launch), the per-THREAD tensor-map
 *     cache, the per-device `max dynamic shared memory` flags, and the optional profiling hook.
 *   - return value: 0 = ok, <0 = vz_status; vz_status_string() names it.
 *   - bf16 tensors are row-major; "ld" arguments are leading dimensions in ELEMENTS.
 */
#ifndef VZ_B200_H
#define VZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  VZ_OK = 0,
  VZ_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, misaligned pointer/stride        */
  VZ_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels were written for                   */
  VZ_ERR_CUDA = -3,         /* a CUDA runtime/driver call failed (see vz_last_cuda_error)         */
  VZ_ERR_NO_DEVICE = -4,    /* no sm_100 device                                                   */
  VZ_ERR_WORKSPACE = -5     /* workspace too small                                                */
} vz_status;

/* constants.py:12-14 */
#define VZ_IGNORE_INDEX (-100)
#define VZ_IMAGE_TOKEN_INDEX (-200)

const char* vz_status_string(int status);
/* last cudaError_t seen by the calling thread inside this library (0 if none) */
int vz_last_cuda_error(void);
/* library/ABI version: major*100+minor */
int vz_version(void);

/* ------------------------------------------------------------------------------------------ */
/* GEMM core (tcgen05 / TMEM / TMA).  out[M,N] = epilogue(A[M,K] * W[N,K]^T)                   */
/* Replaces every torch.nn.Linear / F.linear / conv-as-GEMM call on the path                   */
/* (SURVEY.md 2.2 rows K4, K5, K7, K8).                                                        */
/* ------------------------------------------------------------------------------------------ */
/* VZ_ACT_SWIGLU: W holds gate / up rows interleaved in blocks of 64 (rows 128 j .. 128 j + 63 = gate rows
 * 64 j .., rows 128 j + 64 .. 128 j + 127 = the matching up rows); the epilogue writes
 * out[m, 64 j + i] = silu(gate) * up, i.e. N / 2 output columns (N % 128 == 0).  The MLP of the LLM behind the
 * splice (HF MistralMLP: down_proj(act_fn(gate_proj(x)) * up_proj(x))) without a [M, 2 I] round trip.      */
enum { VZ_ACT_NONE = 0, VZ_ACT_QUICK_GELU = 1, VZ_ACT_GELU_ERF = 2, VZ_ACT_SWIGLU = 3 };
enum {
  VZ_ROWS_PLAIN = 0,       /* out row = m, residual row = m                                       */
  VZ_ROWS_PATCH_EMBED = 1, /* out row = m + m/rows_per + 1, residual row = 1 + m % rows_per
                              (CLIP patch embedding + position embedding, rows_per = 576)         */
  VZ_ROWS_RES_MOD = 2      /* out row = m, residual row = m % rows_per (broadcast residual)       */
};

typedef struct {
  const void* A;        /* bf16 [M, lda]                       */
  const void* W;        /* bf16 [N, ldw]                       */
  void* out;            /* bf16 [*, ldo]                       */
  const float* bias;    /* f32 [N] or NULL                     */
  const void* residual; /* bf16 [*, ldr] or NULL (may alias out) */
  int M, N, K;
  int lda, ldw, ldo, ldr;
  int act;              /* VZ_ACT_*                            */
  int row_mode;         /* VZ_ROWS_*                           */
  int rows_per;         /* see row_mode                        */
  int force_simple;     /* 1 = debug path: plain CUDA-core GEMM (no tcgen05); for bring-up only */
  /* batched form (batch > 1): problem b uses A + b*a_bstride, W + b*w_bstride, out + b*o_bstride,
   * residual + b*r_bstride, bias + b*bias_bstride (all in ELEMENTS; 0 = shared).  M, N, K and
   * the leading dimensions are common to all problems.                                       */
  int batch;
  int out_f32;          /* 1 = write float32 instead of bf16 (no residual / row remap)          */
  long long a_bstride, w_bstride, o_bstride, r_bstride, bias_bstride;
  /* LayerNorm fused around the GEMM (ViT layers).  Producer: stats_out != NULL makes the epilogue write,
   * per output row and per (n-tile, column half), the partial (sum, sum of squares) of the row it
   * just produced into stats_out[M][stats_np][2] (stats_np = vz_gemm_stats_partials(M, N)).
   * Consumer: ln_stats != NULL treats A as the INPUT of a LayerNorm over its K columns whose gamma
   * is already folded into W (W' = W * gamma) and whose beta is folded into bias
   * (b' = b + W beta): out = rstd * (A W'^T - mean * ln_colsum) + b', ln_colsum[n] = sum_k W'[n,k].  */
  const float* ln_stats;
  const float* ln_colsum;
  int ln_np;
  float ln_eps;
  float* stats_out;
  int stats_np;
  /* Optional stream-K scratch (device memory, >= vz_gemm_sk_workspace_bytes(), 16-byte aligned, exclusive
   * to one stream at a time): lets the launcher split the k-blocks of the last, partially filled round of
   * tiles evenly over all SMs (fp32 partial sums handed over through this buffer and added in a fixed
   * order, so results stay deterministic).  NULL = whole tiles only.                                    */
  void* sk_ws;
  size_t sk_ws_bytes;
  /* 1 = W is given as [K, ldw] row-major with N contiguous (out = A * W instead of A * W^T); N % 64 == 0,
   * plain epilogue (bias only).  Used for P * f in the cross-attention: no transposed copy of f.        */
  int w_is_kn;
  /* 1 = the fused normalisation is an RMSNorm (HF MistralRMSNorm): rstd = rsqrt(sum x^2 / K + ln_eps), no mean;
   * only the second component of ln_stats is read, ln_colsum is not needed and bias may be NULL.          */
  int ln_rms;
} vz_gemm_args;
size_t vz_gemm_sk_workspace_bytes(void);

int vz_gemm_bf16(const vz_gemm_args* args, void* stream);
/* number of partial-statistics slots per row that a stats_out GEMM of this shape writes */
int vz_gemm_stats_partials(int M, int N);

/* Measurement hooks (bench.py / tests; no effect on results).
 * vz_kernel_launches: number of CUDA kernels this library has launched in the process.
 * vz_gemm_profile(1): record CUDA events around every tcgen05 GEMM launch (on its own stream);
 * vz_gemm_profile_read: synchronise them and return launches, summed ms and summed 2*M*N*K.     */
long long vz_kernel_launches(void);
int vz_gemm_profile(int enable);
int vz_gemm_profile_read(long long* launches, double* total_ms, double* total_flops);
/* The same for every kernel family: vz_profile(1) records CUDA events around every launch of the library
 * (each on its own stream); vz_profile_read(tag) synchronises them and returns the launches of that family,
 * their summed milliseconds and their summed algorithmic work (FLOPs for the tensor-core kernels, bytes for
 * the HBM-bound ones, 0 where the launcher cannot know it).  vz_gemm_profile* = tag VZ_PROF_GEMM.          */
enum {
  VZ_PROF_GEMM = 0, VZ_PROF_VIT_ATTN = 1, VZ_PROF_FUSE = 2, VZ_PROF_PRE_H = 3, VZ_PROF_PRE_V = 4,
  VZ_PROF_PRE_FUSED = 5, VZ_PROF_SPLICE_SCATTER = 6, VZ_PROF_QATTN = 7, VZ_PROF_SOFTMAX = 8,
  VZ_PROF_LAYERNORM = 9, VZ_PROF_TEXT_GATHER = 10, VZ_PROF_OTHER = 11, VZ_PROF_LLM_ATTN = 12, VZ_PROF_COUNT = 13
};
int vz_profile(int enable);
int vz_profile_read(int tag, long long* launches, double* total_ms, double* total_work);
const char* vz_profile_tag_name(int tag);

/* ------------------------------------------------------------------------------------------ */
/* Row kernels                                                                                 */
/* ------------------------------------------------------------------------------------------ */
/* LayerNorm over rows of width D (D % 8 == 0, D <= 8192), fp32 statistics, eps inside sqrt.
 * Replaces nn.LayerNorm in HF CLIP (pre_layrnorm, layer_norm1/2) and QFormer
 * (multimodal_projector/builder.py:15,18,27,68,70).                                           */
int vz_layernorm_bf16(const void* x, int ldx, const float* gamma, const float* beta, void* out,
                      int ldo, int M, int D, float eps, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* (1) Preprocess: visual-prompt alpha blend + anyres resize/pad/tile + normalise + patchify    */
/* Replaces vip_processor/conversation_generator.py:143-146 (alpha_composite),                  */
/* multi_scale_process.py:71-114,136-183 (LANCZOS resize, pad, tile cut, global view) and      */
/* CLIPImageProcessor.preprocess (rescale + normalise).                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const uint8_t* src;  /* RGB u8 [H, W, 3] (PIL / numpy HWC)                                      */
  const uint8_t* layers; /* RGBA u8 [n_layers, H, W, 4] host-rasterised overlays, or NULL         */
  int W, H;
  int prim_begin, prim_count; /* range into the primitive array, applied in order              */
  /* The resampling tables of the image's views address a virtual CANVAS in which the image sits at
   * (pad_x, pad_y): 0,0 = the image itself; > 0 = expand2square padding (mm_utils.py:16-35), canvas
   * pixels outside the image have the colour bg; < 0 = centre crop (mm_utils.py:65-74).
   * Visual prompts stay in image coordinates.                                                   */
  int pad_x, pad_y;
  uint32_t bg;         /* r | g<<8 | b<<16 */
  int reserved;
} vz_image_desc;

enum { VZ_PRIM_LAYER = 0, VZ_PRIM_RECT = 1 };
typedef struct {
  int type;            /* VZ_PRIM_LAYER: composite layers[layer]; VZ_PRIM_RECT: PIL rectangle outline */
  int layer;
  int x0, y0, x1, y1;  /* rectangle corners (inclusive), already int()-truncated like PIL        */
  int width;           /* outline width                                                          */
  uint32_t rgba;       /* r | g<<8 | b<<16 | a<<24                                               */
} vz_prim;

typedef struct {
  int image;           /* index into the image array                                              */
  int out_w, out_h;    /* LANCZOS target size of the source image for this view                   */
  int off_x, off_y;    /* where the resized image is pasted on the (black) canvas                 */
  int tile_x, tile_y;  /* origin of this 336x336 tile on the canvas                               */
  int tab_h, tab_v;    /* offsets (in int32 words) of the axis tables inside `tables`             */
  int hview;           /* vz_preprocess2 / 3: index of the horizontal view (image, tab_h) this tile reads */
  int tab_v_dp;        /* vz_preprocess3: word offset (multiple of 4) of the vertical dp4a table       */
  int reserved;
} vz_tile_desc;

/* vz_preprocess2: one entry per distinct (image, horizontal table) pair.  Its horizontally resampled
 * (and blended) rows live in the scratch buffer as RGBX u8x4 pixels [rows][out_w], first pixel at `offset`. */
typedef struct {
  int image;           /* index into the image array                                              */
  int tab_h;           /* word offset of the horizontal axis table                                */
  int out_w;           /* width after the horizontal pass                                         */
  int rows;            /* rows of the canvas = rows of the intermediate                           */
  long long offset;    /* first 32-bit word of this view's intermediate in `scratch` (multiple of 4) */
  int tab_h_dp;        /* vz_preprocess3: word offset (multiple of 4) of the horizontal dp4a table     */
  int reserved;
} vz_hview_desc;

/* dp4a axis table (vz_preprocess3), 32-bit words, every section 16-byte aligned, built on the host from the
 * table above (anyres.resample_table_dp): [0] = G groups of four taps per output, [1] = n outputs, [2..3] = 0;
 * abase[n4] = window start aligned down to 4 (n4 = n rounded up to 4); ngrp[n4] = groups output i needs;
 * coef[n][G] as uint4 = (low, middle, signed high) byte limbs of taps abase + 4g .. + 3 (zero padded), 0.     */

/* Axis table layout (int32 words), built on the host exactly like Pillow's precompute_coeffs
 * (Resample.c) for (in_size -> out_size):  [0]=ksize, [1]=out_size, then out_size words xmin,
 * then out_size words count, then out_size*ksize fixed-point (22-bit) coefficients.            */
enum { VZ_OUT_PATCHES_BF16 = 0, VZ_OUT_CHW_F32 = 1 };

/* out: VZ_OUT_PATCHES_BF16 -> bf16 [T*576, 592] (im2col rows in (c,ky,kx) order, K padded
 *      588->592 with zeros);  VZ_OUT_CHW_F32 -> f32 [T,3,336,336] (the reference layout).
 * lut768: f32 [3][256] normalisation table (generated by the oracle processor on a 0..255 ramp).
 * images / prims / tiles / tables / lut768 are DEVICE arrays; max_src_w = widest source image and
 * max_ksize = largest ksize among the tables (they size the kernel's shared-memory staging).    */
int vz_preprocess(const vz_image_desc* images, int n_images, const vz_prim* prims, int n_prims,
                  const vz_tile_desc* tiles, int n_tiles, const int32_t* tables,
                  const float* lut768, int out_mode, void* out, int max_src_w, int max_ksize,
                  void* stream);

/* Same result as vz_preprocess, as TWO kernels (the form used whenever something is resampled): the
 * horizontal pass (+ blend, + canvas padding) runs once per source row of every horizontal view into an
 * RGBX intermediate in `scratch` (it stays in L2), the vertical pass + normalise + patchify reads only
 * the taps it needs.  No source row is filtered twice (the fused kernel recomputes the horizontal
 * pass 1.43x because neighbouring 14-row bands overlap).  scratch: device memory, scratch_pixels x 4
 * bytes, >= the largest offset + rows * out_w of the views; max_span_px = the widest source window
 * any 256 consecutive output columns of any view need (host-computed from the tables).            */
int vz_preprocess2(const vz_image_desc* images, int n_images, const vz_prim* prims, int n_prims,
                   const vz_hview_desc* hviews, int n_hviews, const vz_tile_desc* tiles, int n_tiles,
                   const int32_t* tables, const float* lut768, int out_mode, void* out, void* scratch,
                   long long scratch_pixels, int max_span_px, int max_rows, int max_out_w, int max_ksize,
                   void* stream);

/* The default form whenever something is resampled: same interface and same bits as vz_preprocess2, with the
 * tap loops on the 4-way byte dot product (coefficients split into three byte limbs, windows aligned down to
 * four pixels / four rows; dp4a tables, see vz_hview_desc).  The intermediate in `scratch` is, per view,
 * [ceil(rows / 4)][out_w] uint4 = (R word, G word, B word, 0), one word = four vertically consecutive u8 pixels
 * of one channel, first word at `offset`; scratch_words >= the largest offset + 4 * ceil(rows / 4) * out_w.
 * max_span_px = the widest source window any 128 consecutive output columns of any view need; max_groups = the
 * largest G of the plan's dp4a tables; max_band_groups = the most 4-row groups the window of any 14-row output
 * band spans.  VZ_ERR_UNSUPPORTED when G > 16 (ksize > 58) or a band window does not fit shared memory
 * (scale > ~7): use vz_preprocess2 then.                                                                     */
int vz_preprocess3(const vz_image_desc* images, int n_images, const vz_prim* prims, int n_prims,
                   const vz_hview_desc* hviews, int n_hviews, const vz_tile_desc* tiles, int n_tiles,
                   const int32_t* tables, const float* lut768, int out_mode, void* out, void* scratch,
                   long long scratch_words, int max_span_px, int max_rows, int max_out_w, int max_groups,
                   int max_band_groups, void* stream);

/* The identity form (BASELINE config 2): every tile is the whole of a 336 x 336 image (no resampling, no
 * canvas): blend the visual prompts, normalise, patchify, four pixels per thread with 32- / 128-bit loads.
 * Same bits as vz_preprocess on such a plan.  Image base pointers must be 4-byte, layer pointers 16-byte aligned. */
int vz_preprocess_identity(const vz_image_desc* images, int n_images, const vz_prim* prims, int n_prims,
                           const vz_tile_desc* tiles, int n_tiles, const float* lut768, int out_mode, void* out,
                           void* stream);

/* f32/bf16 pixel_values [T,3,336,336] -> bf16 patches [T*576,592]; the API-compatible entry of
 * CLIPVisionTower.forward (vision_encoder/vision_encoder.py:80-117) when the caller already holds
 * reference-style pixel tensors.  src_is_f32: 1 = float32, 0 = bf16.                            */
int vz_patchify(const void* pixel_values, int src_is_f32, int T, void* patches, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* (2) CLIP ViT-L/14-336 encoder + multi-layer fusion                                           */
/* Replaces CLIPVisionTower.forward / feature_select (vision_encoder/vision_encoder.py:58-117)  */
/* and DenseChannelIntegrationFusion.forward (gating_fusion/gating_fusion.py:22-50).            */
/* ------------------------------------------------------------------------------------------ */
#define VZ_VIT_LAYERS 24
#define VZ_VIT_WIDTH 1024
#define VZ_VIT_TOKENS 577
#define VZ_VIT_PATCHES 576
#define VZ_VIT_HEADS 16
#define VZ_VIT_MLP 4096
#define VZ_PATCH_K 592
#define VZ_FUSED_WIDTH 5120

/* LayerNorm 1/2 of every encoder layer are FOLDED into the following Linear (see vz_gemm_args):
 * w_qkv = cat(q,k,v)_proj.weight * layer_norm1.weight, b_qkv = bias + W layer_norm1.bias,
 * s_qkv[n] = sum_k w_qkv[n,k]; likewise w_fc1 / b_fc1 / s_fc1 with layer_norm2.               */
typedef struct {
  const void* w_qkv;   /* bf16 [3072,1024] (gamma folded) */
  const float* b_qkv;  /* f32 [3072] (beta folded)        */
  const float* s_qkv;  /* f32 [3072] column sums          */
  const void* w_o;     /* bf16 [1024,1024] */
  const float* b_o;
  const void* w_fc1;   /* bf16 [4096,1024] (gamma folded) */
  const float* b_fc1;  /* f32 [4096] (beta folded)        */
  const float* s_fc1;  /* f32 [4096] column sums          */
  const void* w_fc2;   /* bf16 [1024,4096] */
  const float* b_fc2;
} vz_vit_layer;

typedef struct {
  const void* patch_w;    /* bf16 [1024,592]  conv weight flattened (c,ky,kx), zero padded */
  const void* class_emb;  /* bf16 [1024] */
  const void* pos_emb;    /* bf16 [577,1024] */
  const float *pre_ln_g, *pre_ln_b;
  vz_vit_layer layers[VZ_VIT_LAYERS];
} vz_vit_weights;

/* workspace of vz_vit_forward: a ring of six hidden-state buffers (the fusion takes each group's mean as soon
 * as the group is complete); keep_hidden = 1 sizes it for hidden_out != NULL (all 25 states kept, tests). */
size_t vz_vit_workspace_bytes(int T);
size_t vz_vit_workspace_bytes_ex(int T, int keep_hidden);

/* CLIP self-attention alone (HF CLIPAttention): qkv bf16 [T*577, 3072] (q | k | v, 16 heads x 64)
 * -> out bf16 [T*577, 1024].  impl: 1 or -1 = the tcgen05/TMEM kernel; 0 (the first, mma.sync implementation)
 * returns VZ_ERR_UNSUPPORTED: that kernel is test infrastructure now (libvz_b200_testonly.so).             */
int vz_vit_attention(const void* qkv, void* out, int T, int impl, void* stream);

/* patches: bf16 [T*576,592].  fused_out: bf16 [T*576,5120] =
 * cat(mean(h4..h8), mean(h9..h13), mean(h14..h18), mean(h19..h23), h24)[:,1:].
 * If norm_g/norm_b are non-NULL the QFormer.pre_norm LayerNorm(5120)
 * (multimodal_projector/builder.py:68,74) is applied in the same kernel.
 * hidden_out (optional, may be NULL): bf16 [25][T*577,1024] copy-out of all hidden states
 * (tests only; the workspace must then be sized with vz_vit_workspace_bytes_ex(T, 1)).           */
int vz_vit_forward(const vz_vit_weights* w, const void* patches, int T, void* fused_out,
                   const float* norm_g, const float* norm_b, void* hidden_out, void* workspace,
                   size_t workspace_bytes, int force_simple_gemm, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* (3) Q-Former projector                                                                      */
/* Replaces QFormer.forward / QFormerBlock.forward (multimodal_projector/builder.py:34-92).     */
/* ------------------------------------------------------------------------------------------ */
#define VZ_QF_BLOCKS 8
#define VZ_QF_QUERIES 32
#define VZ_QF_WIDTH 4096
#define VZ_QF_HEADS 8
#define VZ_QF_HEAD_DIM 512
#define VZ_QF_FFN 8192

/* norm1 / norm2 / norm3 of every block are FOLDED into the Linear that follows them, like the ViT's
 * layer norms (see vz_gemm_args): W' = W * gamma, b' = b + W beta, s[n] = sum_k W'[n,k]; the GEMM that
 * writes the residual stream emits the row statistics.  Exception: block 0's self-attention input is
 * the tile-invariant learned queries plus the per-sample text rows, a handful of rows that still go
 * through the LayerNorm kernel (n1_g, n1_b), so block 0 keeps a PLAIN sa_in_w / sa_in_b and a NULL s_sa_in. */
typedef struct {
  const float *n1_g, *n1_b; /* norm1 (used by block 0 only) */
  const void* sa_in_w;   /* bf16 [12288,4096]: plain in block 0, norm1-folded in blocks 1..7 */
  const float* sa_in_b;  /* f32 [12288] (beta-folded in blocks 1..7) */
  const float* s_sa_in;  /* f32 [12288] column sums of the folded weight; NULL in block 0 */
  const void* sa_out_w;  /* bf16 [4096,4096] */
  const float* sa_out_b;
  const void* ca_q_w;    /* bf16 [4096,4096] = q_proj_weight, norm2-folded */
  const float* ca_q_b;   /* f32 [4096] = in_proj_bias[0:4096] + Wq norm2.bias */
  const float* s_ca_q;   /* f32 [4096] column sums */
  const void* ca_kT_w;   /* bf16 [5120,4096] = k_proj_weight TRANSPOSED (row n, column h*512+d)   */
  const void* ca_v_w;    /* bf16 [4096,5120] = v_proj_weight                                        */
  const float* ca_in_b;  /* f32 [12288] (q,k,v biases; only the v part is read: the k bias cancels in the softmax) */
  const void* ca_out_w;  /* bf16 [4096,4096] */
  const float* ca_out_b;
  const void* ffn1_w;    /* bf16 [8192,4096], norm3-folded */
  const float* ffn1_b;   /* f32 [8192] (beta-folded) */
  const float* s_ffn1;   /* f32 [8192] column sums */
  const void* ffn2_w;    /* bf16 [4096,8192] */
  const float* ffn2_b;
} vz_qf_block;

typedef struct {
  const void* learned_queries; /* bf16 [32,4096] */
  const float *pre_g, *pre_b;  /* LayerNorm(5120) */
  const float *norm_g, *norm_b;/* LayerNorm(4096) */
  vz_qf_block blocks[VZ_QF_BLOCKS];
} vz_qf_weights;

size_t vz_qformer_workspace_bytes(int T, int n_samples, int text_rows);

/* feats: bf16 [T*576,5120] (already pre_norm'ed if feats_normed, else pre_norm is applied here).
 * Text conditioning (vis_zephyr_arch.py:157-195 + builder.py:76-87), dead rows removed:
 *   text_emb  bf16 [text_rows+1, 4096]: the non-image token embeddings of all samples packed
 *             back to back, followed by ONE all-zero row (the zero padding row of :181-186);
 *   text_off  int32 [n_samples+1] prefix offsets of each sample's rows in text_emb;
 *   L         batch-global max text length (quirk Q3: zero-pad rows take part in the softmax);
 *   tile_sample int32 [T] sample index of each tile.
 * text_emb == NULL reproduces QFormer.forward(features, text_embeddings=None).
 * out: bf16 [T*32, ldo] (ldo >= 4096; may point into the all-gather buffer).                   */
int vz_qformer_forward(const vz_qf_weights* w, const void* feats, int feats_normed, int T,
                       const void* text_emb, const int32_t* text_off, int text_rows,
                       int n_samples, int L, const int32_t* tile_sample, void* out, int ldo,
                       void* workspace, size_t workspace_bytes, int force_simple_gemm,
                       void* stream);

/* ------------------------------------------------------------------------------------------ */
/* (4) merge (anyres unpad / image_newline) + splice                                           */
/* Replaces vis_zephyr_arch.py:157-195 (text rows), :214-333 (splice), :396-473 (merge),        */
/* :476-530 (pad + collate).                                                                   */
/* ------------------------------------------------------------------------------------------ */
enum { VZ_MERGE_FLAT = 0, VZ_MERGE_SPATIAL = 1, VZ_MERGE_SPATIAL_UNPAD = 2, VZ_MERGE_SINGLE_NEWLINE = 3 };

/* One entry per image slot (one projected image = one slot, vis_zephyr_arch.py:283-296).      */
typedef struct {
  int row_base;   /* first row of this image in the projector output [sum T_i * hw, D]           */
  int n_rows;     /* rows this slot contributes after merging (host-computed, exact)             */
  int merge;      /* VZ_MERGE_*                                                                  */
  int hw;         /* rows per tile (h*w)                                                         */
  int h, w;       /* feature-map side lengths per tile                                           */
  int n_w, n_h;   /* anyres grid (calculate_grid_shape)                                          */
  int y0, y1, x0, x1; /* unpad_image crop on the [n_h*h, n_w*w] map (as written, quirk Q4)       */
} vz_slot_desc;

/* Plan pass: one CTA, warp-level prefix sums.  Outputs (all int32):
 *   tok_dest [B,S]   destination row of each kept non-image token inside its sample, -1 else
 *   slot_dest [n_slots*2] (sample, first destination row) of each consumed slot, -1 if unused
 *   lengths  [B]     spliced length per sample (after optional truncation)
 *   text_len [B]     count of ids != IMAGE_TOKEN_INDEX over the whole row (vis_zephyr_arch.py:168)
 *   totals   [4]     {Lmax, max text_len, slots consumed, sum text_len}
 * mask: uint8 [B,S] (attention_mask.bool()) or NULL (all ones).                               */
int vz_splice_plan(const int64_t* input_ids, const uint8_t* mask, int B, int S,
                   const vz_slot_desc* slots, int n_slots, int max_len, int32_t* tok_dest,
                   int32_t* slot_dest, int32_t* lengths, int32_t* text_len, int32_t* totals,
                   void* stream);

/* Gather pass for text conditioning: text_emb[text_off[b] + j] = embed[id] for the j-th
 * non-image token of sample b; row `sum text_len` is set to zero.  text_off int32 [B+1] is
 * written too.  elem_bytes = 2 (bf16/fp16) or 4 (f32); D*elem_bytes % 16 == 0.                 */
int vz_text_gather(const int64_t* input_ids, int B, int S, const void* embed_table, int D,
                   int elem_bytes, const int32_t* text_len, void* text_emb, int32_t* text_off,
                   void* stream);

/* Scatter pass.  vis: projector output rows [*, D] (ldv elements per row), image_newline [D] or
 * NULL, labels int64 [B,S] or NULL (=> IGNORE_INDEX everywhere).  Outputs: out_embeds [B,Lout,D],
 * out_labels int64 [B,Lout], out_mask uint8 [B,Lout], out_pos int64 [B,Lout]. Every output
 * element is written exactly once (padding included), so the buffers need no memset.
 * slot_prefix int32 [n_slots+1]: exclusive prefix sum of slots[].n_rows (host-known, uploaded with
 * the slot table); total_vis_rows = slot_prefix[n_slots].                                       */
int vz_splice_scatter(const int64_t* input_ids, const int64_t* labels, int B, int S,
                      const void* embed_table, const void* vis, int ldv, const void* image_newline,
                      int D, int elem_bytes, const vz_slot_desc* slots, int n_slots,
                      const int32_t* slot_prefix, int total_vis_rows, const int32_t* tok_dest, const int32_t* slot_dest, const int32_t* lengths,
                      int Lout, int pad_left, void* out_embeds, int64_t* out_labels,
                      uint8_t* out_mask, int64_t* out_pos, void* stream);

/* The same scatter, also handing the FIRST LLM LAYER its RMSNorm statistic ("first layer's
 * RMSNorm + QKV fused with the splice output", language_model/vis_zephyr.py:86-98): out_row_stats f32 [B*Lout][2]
 * = (0, sum of squares) of every output row, taken while the row passes through the registers (bf16 rows only).
 * Feeding it to vz_gemm_bf16 as ln_stats (ln_np = 1) with the RMSNorm gain folded into the stacked q/k/v weight makes
 * that GEMM compute rmsnorm(x) Wqkv^T directly: a zero sum makes the fused-LayerNorm epilogue's mean term vanish
 * and leaves rstd = rsqrt(sum of squares / K + eps).  out_row_stats == NULL: plain vz_splice_scatter.              */
int vz_splice_scatter_rms(const int64_t* input_ids, const int64_t* labels, int B, int S,
                          const void* embed_table, const void* vis, int ldv, const void* image_newline,
                          int D, int elem_bytes, const vz_slot_desc* slots, int n_slots,
                          const int32_t* slot_prefix, int total_vis_rows, const int32_t* tok_dest, const int32_t* slot_dest, const int32_t* lengths,
                          int Lout, int pad_left, void* out_embeds, int64_t* out_labels,
                          uint8_t* out_mask, int64_t* out_pos, float* out_row_stats, void* stream);

/* "Next" row: the collator in front of the splice.
 * DataCollatorForSupervisedDataset (train/train.py:657-707) from packed ragged rows:
 * out_ids[b,s] = s < len_b ? flat_ids[offsets[b]+s] : pad_id, labels padded with IGNORE_INDEX, both
 * truncated to S_out columns (S_out = min(max len, model_max_length), host-known),
 * out_mask = out_ids != pad_id.                                                               */
int vz_collate(const int64_t* flat_ids, const int64_t* flat_labels, const int32_t* offsets, int B, int S_out,
               int64_t pad_id, int64_t* out_ids, int64_t* out_labels, uint8_t* out_mask, void* stream);

/* Merge only (vis_zephyr_arch.py:396-473) for one slot list: out rows [sum n_rows, D].          */
int vz_merge_rows(const void* vis, int ldv, const void* image_newline, int D, int elem_bytes,
                  const vz_slot_desc* slots, int n_slots, const int32_t* out_row_base, void* out,
                  void* stream);

/* ------------------------------------------------------------------------------------------ */
/* (5) Row kernels of the LLM prefill behind the splice.  A Mistral decoder layer on     */
/* packed rows is four vz_gemm_bf16 calls (RMSNorm folded into q|k|v and into gate|up: ln_rms; SwiGLU in the     */
/* epilogue: VZ_ACT_SWIGLU; residual + next layer's row statistics in the o_proj / down_proj epilogues) around    */
/* these kernels and a causal attention core.  The reference runs HF Mistral here                                */
/* (language_model/vis_zephyr.py:86-98; train/zephyr_flash_attn_monkey_patch.py:86-136).                          */
/* ------------------------------------------------------------------------------------------ */
/* cos_sin f32 [M, half_dim, 2] = (cos, sin)(positions[m] * inv_freq[j]) rounded to bf16 values, as HF
 * MistralRotaryEmbedding.forward returns them; one table per call, shared by all layers.                     */
int vz_rope_table(const int32_t* positions, int M, const float* inv_freq, int half_dim, float* cos_sin, void* stream);
/* apply_rotary_pos_emb in place: x bf16 [M, ld], heads h = 0 .. n_heads-1 at columns h * head_dim (the q and k
 * heads of a packed q | k | v row), rotate_half pairing (j, j + head_dim / 2), bf16 rounding of every product. */
int vz_rope_apply(void* x, int ld, int M, int n_heads, int head_dim, const float* cos_sin, void* stream);
/* stats f32 [M, 2] = (sum, sum of squares) of every bf16 row x[m, 0:D]: the ln_stats of a vz_gemm_bf16 whose
 * A operand did not come out of a stats_out epilogue (first decoder layer without the scatter's statistics).   */
int vz_row_stats(const void* x, int ldx, int M, int D, float* stats, void* stream);
/* Row copies between the padded [B, L] layout of the splice and the packed rows of the prefill:
 * gather = 1: dst row i = src row map[i]; gather = 0: dst row map[i] = src row i; map[i] < 0 skips the row.
 * row_bytes and both leading dimensions in BYTES, multiples of 16.                                            */
int vz_rows_move(const void* src, long long lds_bytes, void* dst, long long ldd_bytes, const int32_t* map,
                 int n_rows, int row_bytes, int gather, void* stream);

/* Causal grouped-query attention of the prefill on packed rows (tcgen05 / TMEM; head_dim 128): qkv bf16 [M, ld] =
 * q heads | k heads | v heads per row (after vz_rope_apply), sample b = rows [cu[b], cu[b + 1]); out bf16 [M, ldo]
 * (n_heads * 128 columns) = softmax(scale q k^T, keys 0 .. i of the same sample) v.  The work list comes from
 * vz_attn_causal_items (HOST: lens_h[b] = rows of sample b -> int32 x 4 per (128-query tile, head), longest first;
 * returns the item count, or the count needed if items_h is NULL / max_items too small; algo_flops = 4 * 128 *
 * sum_b S_b (S_b + 1) / 2 * n_heads) and must be copied to the device by the caller.  Replaces the attention core of
 * HF MistralAttention.forward / train/zephyr_flash_attn_monkey_patch.py:86-136 (flash_attn_varlen on unpadded rows). */
int vz_attn_causal_items(const int32_t* lens_h, int B, int n_heads, int32_t* items_h, int max_items, double* algo_flops);
int vz_attn_causal(const void* qkv, int ld, int M, void* out, int ldo, const int32_t* items, int n_items, int n_heads,
                   int n_kv_heads, int head_dim, float scale, double algo_flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VZ_B200_H */
