// vz_gemm.cu -- persistent, warp-specialised bf16 GEMM for sm_100a:
//   out[M,N] = epilogue( A[M,K] * W[N,K]^T )          (both operands K-major, fp32 accumulate)
// TMA (cp.async.bulk.tensor, 128B swizzle) -> shared-memory ring -> tcgen05.mma (one issuing
// thread, 128xBN accumulator in TMEM, double buffered) -> tcgen05.ld epilogue with fused
// bias / quick-GELU / erf-GELU / residual / row remapping.
//
// This one kernel serves every Linear on the path (K4 patch embedding, K5 CLIP
// q/k/v/out/fc1/fc2, K7 the stacked cross-attention K/V projection, K8 the Q-Former linears).
#include "vz_common.cuh"

#include <unordered_map>

#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <vector>

namespace vz {

thread_local int g_last_cuda_error = 0;

static std::atomic<long long> g_launches{0};
static std::atomic<uint32_t> g_sk_epoch{0};   // stream-K hand-over flags carry the epoch of the launch that wrote them
constexpr size_t kSkFlagBytes = 8192;         // >= SMs * epilogue warps * 4 bytes
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Optional measurement hook (bench.py): CUDA events around every kernel launch of the library, on the
// launching stream, tagged by kernel family, plus the algorithmic work (FLOPs or bytes) of that launch.
// Off by default; a ProfScope costs one relaxed load when off.
struct Prof {
  std::mutex mu;
  std::atomic<bool> on{false};
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  std::vector<int> tag;
  std::vector<double> work;
};
static Prof g_prof;

ProfScope::ProfScope(int tag, double work, cudaStream_t stream) : e1(nullptr), st(stream) {
  if (!g_prof.on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof.mu);
  while (g_prof.used + 2 > g_prof.pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    g_prof.pool.push_back(e);
  }
  cudaEvent_t e0 = g_prof.pool[g_prof.used++];
  e1 = g_prof.pool[g_prof.used++];
  g_prof.tag.push_back(tag);
  g_prof.work.push_back(work);
  cudaEventRecord(e0, st);
}
ProfScope::~ProfScope() {
  if (e1) cudaEventRecord(e1, st);
}

namespace {

constexpr int BM = 128;       // UMMA M (cta_group::1)
constexpr int BK = 64;        // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;    // fixed for 16-bit inputs
constexpr int kEpiWarps = 8;  // two warps per TMEM lane quarter, interleaved 32-column chunks
constexpr int kThreads = 128 + 32 * kEpiWarps;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps4-11 epilogue
constexpr int EPI_STAGE_BYTES = 32 * 128;  // per epilogue warp: 32 rows x 32 values (bf16 or f32), chunk-swizzled

// TWO = 2-CTA form (cta_group::2): a pair of SMs computes a 256 x BN tile, each CTA stages its own 128
// A rows and HALF of the B tile, which cuts the shared-memory traffic per FLOP by a third.
// (Measured and dropped: TMA stores for the residual epilogues too.  They need residual landing buffers distinct
//  from the output staging, 8 KB per epilogue warp, i.e. one pipeline stage less in the 2-CTA form -- and five stages
//  cost more than the stores save: out-proj 45.2 -> 49.2 us, fc2 119.7 -> 127.0 us.)
template <int BN, bool TWO = false, bool RES = false>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (TWO ? BN / 2 : BN) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = TWO ? 6 : ((BN == 256) ? 4 : (BN == 192 ? 4 : 6));
  static constexpr int TMEM_COLS = (BN == 192) ? 512 : 2 * BN;  // two accumulator stages (power of two)
  static constexpr int BAR_BYTES = 256;
  static constexpr int EPI_WARP_BYTES = EPI_STAGE_BYTES;
  static constexpr int EPI_BYTES = kEpiWarps * EPI_WARP_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // + align slack
};

struct GemmDev {
  __nv_bfloat16* out;  // bf16, or float when out_f32
  const float* bias;
  const __nv_bfloat16* residual;
  int M, N, K;
  int ldo, ldr;
  int act, row_mode, rows_per;
  int num_m, num_n, num_k;
  int batch, out_f32;
  long long o_bstride, r_bstride, bias_bstride;  // elements between consecutive batches
  // LayerNorm fused around the GEMM (see vz_gemm_args): consumer side ...
  const float* ln_stats;   // [M][ln_np][2] partial (sum, sum of squares) of every A row, or NULL
  const float* ln_colsum;  // [N] sum_k W'[n,k]
  int ln_np;
  float ln_eps;
  int ln_rms;              // 1 = RMSNorm: no mean (rstd = rsqrt(sum x^2 / K + eps)), colsum / bias not needed
  // ... and producer side: partial row statistics of THIS GEMM's output
  float* stats_out;        // [M][stats_np][2] or NULL
  int stats_np;
  // stream-K tail (see WorkIter): the first dp_tiles tiles are walked whole, round-robin; the k-blocks of
  // the last sk_tiles tiles are cut into one contiguous range per worker
  int dp_tiles, sk_tiles;
  int tma_out;             // 1 = output chunks leave through TMA stores (plain rows, bf16): tmC is valid
  int group_n;             // n-blocks per rasterisation band: all m-tiles are walked per band, so A is re-read from
                           // DRAM once per band while the band's W rows (group_n * BN * K * 2 bytes) stay in L2
  float4* sk_ws;           // per (worker, CTA rank): one 128 x BN fp32 partial accumulator
  uint32_t* sk_flags;      // per (worker, CTA rank, epilogue warp): epoch of the partial it holds
  uint32_t sk_epoch;
};

// The work list of one worker (a CTA, or a CTA pair in the 2-CTA form).  Data-parallel part: tiles
// worker, worker + W, ... < dp_tiles, all k-blocks each.  Stream-K part: the sk_tiles * num_k k-blocks
// of the remaining tiles are split evenly; worker w owns units [U w / W, U (w + 1) / W), which cover the
// TAIL of one tile (k-blocks kb0 > 0: a partial sum, handed over through global memory), then whole
// tiles, then the HEAD of one tile (kb0 == 0, kb1 < num_k: this worker finishes the tile by adding the
// partial sums of the workers after it, in worker order -> deterministic).  Tails come first in every
// worker's list and never wait, so the hand-over cannot deadlock.
struct WorkItem { int tile, kb0, kb1; };
struct WorkIter {
  int W, dp_tiles, num_k, next_dp, u, u1;
  __device__ __forceinline__ WorkIter(int w, int W_, int dp_tiles_, int sk_tiles, int num_k_)
      : W(W_), dp_tiles(dp_tiles_), num_k(num_k_), next_dp(w) {
    const int U = sk_tiles * num_k_;
    u = (int)(((long long)U * w) / W_);
    u1 = (int)(((long long)U * (w + 1)) / W_);
  }
  __device__ __forceinline__ bool next(WorkItem& it) {
    if (next_dp < dp_tiles) { it.tile = next_dp; it.kb0 = 0; it.kb1 = num_k; next_dp += W; return true; }
    if (u < u1) {
      const int t = u / num_k, base = t * num_k;
      const int end = u1 < base + num_k ? u1 : base + num_k;
      it.tile = dp_tiles + t; it.kb0 = u - base; it.kb1 = end - base;
      u = end;
      return true;
    }
    return false;
  }
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int group_n, int& m_blk, int& n_blk, int& b) {
  const int per_batch = num_m * num_n;
  b = tile / per_batch;
  tile -= b * per_batch;
  const int per_band = num_m * group_n;
  const int band = tile / per_band;
  const int within = tile - band * per_band;
  const int n0 = band * group_n;
  const int gsz = min(group_n, num_n - n0);
  m_blk = within / gsz;
  n_blk = n0 + (within - m_blk * gsz);
}

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == VZ_ACT_QUICK_GELU) {
    // x * sigmoid(1.702 x) = 0.5 x (1 + tanh(0.851 x)): one MUFU op per element
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
  } else if (act == VZ_ACT_GELU_ERF) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
  } else if (act == VZ_ACT_SWIGLU) {
    // silu(x) = x * sigmoid(x) = 0.5 x (1 + tanh(0.5 x)); the caller multiplies by the "up" value
    const float hx = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(hx));
    return fmaf(hx, t, hx);
  }
  return x;
}

// Epilogue features are template parameters (ACT activation, RES residual, LN fused LayerNorm
// consumer, STATS fused LayerNorm producer, F32 fp32 output): the epilogue sits at the register
// limit of a 384-thread CTA, so each instantiation only carries the state it needs.
// BKN: the B operand is given as [K, N] row-major (N contiguous, "MN-major" for the tensor core): its
// stage is BN/64 TMA boxes of 64 k-rows x 64 columns (8 KB each, 128-byte rows, 8-row groups 1 KB apart).
template <int BN, int ACT, bool RES, bool LN, bool STATS, bool F32, bool TWO, bool BKN = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                         const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                         const GemmDev p) {
  using C = Cfg<BN, TWO, RES>;
  // 2-CTA form: `rank` is this CTA's position in its pair; tiles are 256 rows tall and the pair index
  // walks them.  1-CTA form: rank 0, every CTA is its own "pair".
  const uint32_t rank = TWO ? cluster_ctarank() : 0u;
  const int worker = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int TILE_M = TWO ? 2 * BM : BM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * C::A_BYTES;
  uint8_t* sEpi = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + C::EPI_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * C::STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_out) tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], TWO ? 2 : 1);   // 2-CTA: one arrive per CTA's producer (on the leader's barrier)
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], TWO ? 2 * kEpiWarps : kEpiWarps);  // one arrive per epilogue warp (of the pair)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (TWO) tmem_alloc_2cta(tmem_slot, C::TMEM_COLS);
    else tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();   // the peer's barriers must exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      WorkIter work(worker, n_workers, p.dp_tiles, p.sk_tiles, p.num_k);
      WorkItem it;
      while (work.next(it)) {
        int m_blk, n_blk, bz;
        tile_coords(it.tile, p.num_m, p.num_n, p.group_n, m_blk, n_blk, bz);
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
          if (TWO) {
            // both CTAs' bytes are counted on the leader's barrier
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
            else mbar_arrive_cluster(&full_bar[stage], 0);
            tma_load_3d_2cta(&tmA, &full_bar[stage], sA + stage * C::A_BYTES, kb * BK,
                             m_blk * TILE_M + (int)rank * BM, bz);
            if (BKN) {
#pragma unroll
              for (int i = 0; i < BN / 2 / 64; ++i)
                tma_load_3d_2cta(&tmB, &full_bar[stage], sB + stage * C::B_BYTES + i * 8192,
                                 n_blk * BN + (int)rank * (BN / 2) + i * 64, kb * BK, bz);
            } else {
              tma_load_3d_2cta(&tmB, &full_bar[stage], sB + stage * C::B_BYTES, kb * BK,
                               n_blk * BN + (int)rank * (BN / 2), bz);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            tma_load_3d(&tmA, &full_bar[stage], sA + stage * C::A_BYTES, kb * BK, m_blk * BM, bz);
            if (BKN) {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_3d(&tmB, &full_bar[stage], sB + stage * C::B_BYTES + i * 8192, n_blk * BN + i * 64,
                            kb * BK, bz);
            } else {
              tma_load_3d(&tmB, &full_bar[stage], sB + stage * C::B_BYTES, kb * BK, n_blk * BN, bz);
            }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    if (lane == 0 && rank == 0) {   // 2-CTA: only the leader issues; the MMA spans both SMs
      constexpr uint32_t idesc = umma_idesc_bf16_ex(TILE_M, BN, 0, BKN ? 1 : 0);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      WorkIter work(worker, n_workers, p.dp_tiles, p.sk_tiles, p.num_k);
      WorkItem it;
      while (work.next(it)) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 200 + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase, 300 + stage);
          tc_fence_after();
          const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(sA + stage * C::A_BYTES));
          // [K, N] operand: 64-column atoms 8 KB apart (LBO); one UMMA_K step = 16 k-rows = 2 KB
          const uint64_t b_desc = BKN ? (umma_smem_desc_sw128(smem_u32(sB + stage * C::B_BYTES)) +
                                         ((uint64_t)((8192 >> 4) - 1) << 16))
                                      : umma_smem_desc_sw128(smem_u32(sB + stage * C::B_BYTES));
          constexpr uint64_t b_step = BKN ? (2048 >> 4) : 2;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 32 B (16 bf16) along K inside the swizzle atom: +2 in the (addr>>4) field
            if (TWO) umma_bf16_2cta(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)k * b_step, idesc,
                                    (kb != it.kb0 || k != 0) ? 1u : 0u);
            else umma_bf16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)k * b_step, idesc,
                           (kb != it.kb0 || k != 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs) once these MMAs retire
          if (TWO) umma_commit_2cta(&empty_bar[stage], 3);
          else umma_commit(&empty_bar[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (TWO) umma_commit_2cta(&tfull_bar[acc], 3);
        else umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ============================ epilogue ================================
    // warp e: TMEM lane quarter q = e % 4 (== warp % 4, the hardware's lane-access rule), column
    // chunks c = e/4, e/4 + 2, ...  Each 32x32 chunk goes TMEM -> registers (row per lane) ->
    // per-warp swizzled smem tile -> 8 rows x 64 B per store instruction (coalesced); the residual
    // takes the same route in the opposite direction.
    const int e = warp - 4;
    const int q = e & 3;
    const int half = e >> 2;
    uint8_t* stg = sEpi + e * C::EPI_WARP_BYTES;
    // own-row addressing (lane = row) and cooperative addressing (4 lanes per row)
    const uint32_t own_off = (uint32_t)lane * 64u;
    const uint32_t own_sw = (uint32_t)((lane >> 1) & 3);
    const int co_r = lane >> 2, co_j = lane & 3;
    uint32_t acc = 0, acc_phase = 0;
    uint32_t cc = 0;  // chunk counter: selects which half of the warp's staging area is current
    uint32_t oc = 0;  // TMA stores issued by this warp: their staging halves alternate per STORE (the SwiGLU form skips
                      // every other chunk, so the chunk counter would hand two consecutive stores the same half)
    WorkIter work(worker, n_workers, p.dp_tiles, p.sk_tiles, p.num_k);
    WorkItem it;
    // stream-K hand-over slots: element (row 32 q + lane, column 32 c + 4 j .. + 3) of the CTA's partial
    // tile is float4 [((q * BN/32 + c) * 8 + j) * 32 + lane] -- the register layout, fully coalesced
    constexpr int SK_SLOT4 = BM * BN / 4;   // float4 per (worker, rank)
    constexpr int NR = TWO ? 2 : 1;         // CTAs per worker: slot / flag index = worker * NR + rank < number of SMs
    while (work.next(it)) {
      int m_blk, n_blk, bz;
      tile_coords(it.tile, p.num_m, p.num_n, p.group_n, m_blk, n_blk, bz);
      if (it.kb0 > 0) {
        // ---- tail of a tile that an earlier worker finishes: dump the raw accumulator, raise the flag ----
        mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
        float4* slot = p.sk_ws + (size_t)(worker * NR + (int)rank) * SK_SLOT4;
        // rows of this warp's lane quarter that exist (the Q-Former's 32-row problems leave three quarters of a
        // tile empty: nothing to hand over for them, and the finishing worker does not read them either)
        const bool live = m_blk * TILE_M + (int)rank * BM + q * 32 < p.M;
#pragma unroll 1
        for (int c = half; live && c < BN / 32; c += 2) {
          if (n_blk * BN + c * 32 >= p.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + c * 32, r);
          tmem_ld_wait();
          float4* dst = slot + (size_t)((q * (BN / 32) + c) * 8) * 32 + lane;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j * 32] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                      __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
        tc_fence_before();
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          st_release_u32(p.sk_flags + (worker * NR + (int)rank) * kEpiWarps + e, p.sk_epoch);
          if (TWO) mbar_arrive_cluster(&tempty_bar[acc], 0);
          else mbar_arrive(&tempty_bar[acc]);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      // head of a stream-K tile: the workers after this one hold the rest of its k-blocks
      int sk_first = 0, sk_last = -1;
      if (it.kb1 < p.num_k) {
        const long long U = (long long)p.sk_tiles * p.num_k;
        const long long tile_end = (long long)(it.tile - p.dp_tiles + 1) * p.num_k;
        sk_first = worker + 1;
        sk_last = worker;
        while (sk_last + 1 < n_workers && (U * (sk_last + 1)) / n_workers < tile_end) ++sk_last;
      }
      const int m_base = m_blk * TILE_M + (int)rank * BM + q * 32;
      if (m_base >= p.M) {
        // this warp's 32 rows lie beyond M: keep the accumulator hand-shake going, touch nothing else
        mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (TWO) mbar_arrive_cluster(&tempty_bar[acc], 0);
          else mbar_arrive(&tempty_bar[acc]);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      const float* bias_b = p.bias ? p.bias + (size_t)bz * p.bias_bstride : nullptr;
      const __nv_bfloat16* res_b = RES ? p.residual + (size_t)bz * p.r_bstride : nullptr;
      __nv_bfloat16* out_b = p.out + (size_t)bz * p.o_bstride;
      // fused LayerNorm (consumer): mean / rstd of this lane's own row from the producer's partials,
      // summed in a fixed order (deterministic)
      float ln_mu = 0.f, ln_rstd = 1.f;
      if (LN && m_base + lane < p.M) {
        const float2* ps = reinterpret_cast<const float2*>(p.ln_stats) + (size_t)(m_base + lane) * p.ln_np;
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < p.ln_np; ++i) { const float2 t = ps[i]; s1 += t.x; s2 += t.y; }
        const float inv_k = 1.0f / (float)p.K;
        ln_mu = p.ln_rms ? 0.f : s1 * inv_k;
        ln_rstd = rsqrtf(fmaxf(s2 * inv_k - ln_mu * ln_mu, 0.f) + p.ln_eps);
      }
      float st1 = 0.f, st2 = 0.f;  // producer side: partial statistics of this lane's output row
      // rows this lane touches in the cooperative phases: m_base + 8*i + co_r (32-bit element offsets)
      uint32_t co_out[4], co_res[4];
      bool co_ok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m_base + 8 * i + co_r;
        co_ok[i] = m < p.M;
        int out_row = m, res_row = m;
        if (p.row_mode == VZ_ROWS_PATCH_EMBED) {
          const int img = m / p.rows_per;
          out_row = m + img + 1;
          res_row = 1 + (m - img * p.rows_per);
        } else if (p.row_mode == VZ_ROWS_RES_MOD) {
          res_row = m % p.rows_per;
        }
        co_out[i] = (uint32_t)out_row * (uint32_t)p.ldo + (uint32_t)(co_j * 8);
        co_res[i] = (uint32_t)res_row * (uint32_t)p.ldr + (uint32_t)(co_j * 8);
      }
      // Residual chunks travel global -> smem with cp.async (no registers), one chunk ahead of their
      // use; the first chunk of a tile is requested BEFORE waiting for the accumulator, so its HBM/L2
      // latency hides behind the tile's MMAs.  Chunk cc lives in staging half (cc & 1).
      auto prefetch_residual = [&](int col0_, uint32_t which) {
        uint8_t* dst = stg + (which & 1u) * 2048u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = 8 * i + co_r;
          cp_async_16(dst + rr * 64 + ((co_j ^ ((rr >> 1) & 3)) << 4),
                      co_ok[i] ? (const void*)(res_b + co_res[i] + col0_) : (const void*)res_b, co_ok[i]);
        }
        cp_async_commit();
      };
      if (RES && n_blk * BN + half * 32 < p.N) prefetch_residual(n_blk * BN + half * 32, cc);
      mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      // stream-K: add the later workers' partial sums of chunk c_, in worker order
      auto add_partials = [&](int c_, float* v_, bool first) {
        for (int sw = sk_first; sw <= sk_last; ++sw) {
          const uint32_t* flag = p.sk_flags + (sw * NR + (int)rank) * kEpiWarps + e;
          if (first) {   // first chunk of the tile: the partial may still be on its way
            uint32_t polls = 0;
            while (ld_acquire_u32(flag) != p.sk_epoch) {
              if (++polls > (1u << 28)) { printf("vz: stream-K hand-over timeout worker=%d from=%d\n", worker, sw); __trap(); }
            }
          }
          const float4* src = p.sk_ws + (size_t)(sw * NR + (int)rank) * SK_SLOT4 + (size_t)((q * (BN / 32) + c_) * 8) * 32 + lane;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = __ldcg(src + j * 32);
            v_[4 * j] += t.x; v_[4 * j + 1] += t.y; v_[4 * j + 2] += t.z; v_[4 * j + 3] += t.w;
          }
        }
      };
      // fused normalisation + bias of the 32 columns starting at col_ (pre_ = the bias values if already in registers)
      auto norm_bias = [&](int col_, float* v_, const float4* pre_) {
        if (LN && p.ln_rms) {
          // RMSNorm(x) W^T == rstd * (x W'^T) (gain folded into W')
#pragma unroll
          for (int i = 0; i < 32; ++i) v_[i] *= ln_rstd;
          if (bias_b) {
            const float4* b4 = reinterpret_cast<const float4*>(bias_b + col_);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = __ldg(b4 + i);
              v_[4 * i + 0] += b.x; v_[4 * i + 1] += b.y; v_[4 * i + 2] += b.z; v_[4 * i + 3] += b.w;
            }
          }
        } else if (LN) {
          // LN(x) W^T + b  ==  rstd * (x W'^T - mu * colsum(W')) + b'   (gamma folded into W', beta into b')
          const float4* c4 = reinterpret_cast<const float4*>(p.ln_colsum + col_);
          const float4* b4 = reinterpret_cast<const float4*>(bias_b + col_);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 cs = __ldg(c4 + i), b = __ldg(b4 + i);
            v_[4 * i + 0] = fmaf(ln_rstd, fmaf(-ln_mu, cs.x, v_[4 * i + 0]), b.x);
            v_[4 * i + 1] = fmaf(ln_rstd, fmaf(-ln_mu, cs.y, v_[4 * i + 1]), b.y);
            v_[4 * i + 2] = fmaf(ln_rstd, fmaf(-ln_mu, cs.z, v_[4 * i + 2]), b.z);
            v_[4 * i + 3] = fmaf(ln_rstd, fmaf(-ln_mu, cs.w, v_[4 * i + 3]), b.w);
          }
        } else if (pre_) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v_[4 * i + 0] += pre_[i].x; v_[4 * i + 1] += pre_[i].y; v_[4 * i + 2] += pre_[i].z; v_[4 * i + 3] += pre_[i].w;
          }
        } else if (bias_b) {
          const float4* b4 = reinterpret_cast<const float4*>(bias_b + col_);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b = __ldg(b4 + i);
            v_[4 * i + 0] += b.x; v_[4 * i + 1] += b.y; v_[4 * i + 2] += b.z; v_[4 * i + 3] += b.w;
          }
        }
      };
      constexpr bool SWI = ACT == VZ_ACT_SWIGLU;
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2, ++cc) {
        // SwiGLU: every 128-column group of the tile is [gate 64 | up 64]; chunks 0, 1 of a group are consumed
        // together with chunks 2, 3 (same warp: same parity) and produce 64 output columns per group
        if (SWI && (c & 2)) continue;
        const int col0 = n_blk * BN + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        const int ocol0 = SWI ? n_blk * (BN / 2) + (c >> 2) * 64 + (c & 1) * 32 : col0;   // output column of the chunk
        uint8_t* stg = sEpi + e * C::EPI_WARP_BYTES + (cc & 1u) * 2048u;  // shadows the warp base on purpose
        uint8_t* stg_o = p.tma_out ? sEpi + e * C::EPI_WARP_BYTES + (oc & 1u) * 2048u : stg;   // output staging
        // plain bias: fetched BEFORE the accumulator / residual waits (ncu: the load -> add dependency was the
        // epilogue's largest single stall, and the epilogue is what bounds the K = 1024 GEMMs)
        float4 bb[8];
        const bool pre_bias = !LN && bias_b != nullptr;
        if (pre_bias) {
          const float4* b4 = reinterpret_cast<const float4*>(bias_b + col0);
#pragma unroll
          for (int i = 0; i < 8; ++i) bb[i] = __ldg(b4 + i);
        }
        float v[32];
        {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + c * 32, r);
          if (RES) {
            const bool more = (c + 2 < BN / 32) && (col0 + 64 < p.N);
            if (more) { prefetch_residual(col0 + 64, cc + 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        }
        add_partials(c, v, c == half);
        norm_bias(col0, v, pre_bias ? bb : nullptr);
        if (SWI) {
          float u[32];
          {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_row + (c + 2) * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) u[i] = __uint_as_float(r[i]);
          }
          add_partials(c + 2, u, false);
          norm_bias(col0 + 64, u, nullptr);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = act_apply(v[i], ACT) * u[i];
        } else if (ACT != VZ_ACT_NONE) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = act_apply(v[i], ACT);
        }
        if (F32) {
          // fp32 output (attention scores): 128 B per row, 8 chunks swizzled by row & 7; stores cover
          // 4 rows x 128 B per instruction
          uint8_t* stg = sEpi + e * C::EPI_WARP_BYTES;  // fp32 rows need the warp's whole 4 KB
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(stg + lane * 128 + (((uint32_t)i ^ (uint32_t)(lane & 7)) << 4)) =
                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          __syncwarp();
          float* outf = reinterpret_cast<float*>(p.out);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + (lane >> 3), jj = lane & 7;
            const int m = m_base + rr;
            const float4 w = *reinterpret_cast<const float4*>(stg + rr * 128 + ((jj ^ (rr & 7)) << 4));
            if (m < p.M)
              *reinterpret_cast<float4*>(outf + (size_t)bz * p.o_bstride + (size_t)m * p.ldo + col0 + jj * 4) = w;
          }
          __syncwarp();
          continue;
        }
        if (RES) {
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 w = *reinterpret_cast<const uint4*>(stg + own_off + (((uint32_t)i ^ own_sw) << 4));
            v[8 * i + 0] += bf16_lo(w.x); v[8 * i + 1] += bf16_hi(w.x);
            v[8 * i + 2] += bf16_lo(w.y); v[8 * i + 3] += bf16_hi(w.y);
            v[8 * i + 4] += bf16_lo(w.z); v[8 * i + 5] += bf16_hi(w.z);
            v[8 * i + 6] += bf16_lo(w.w); v[8 * i + 7] += bf16_hi(w.w);
          }
        }
        if (STATS) {
#pragma unroll
          for (int i = 0; i < 32; ++i) { st1 += v[i]; st2 = fmaf(v[i], v[i], st2); }
        }
        if (p.tma_out) {
          // the staging half's previous tenant (two chunks ago) must have been read by its store
          if (lane == 0) tma_store_wait_read1();
          __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
          w.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
          w.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
          w.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
          *reinterpret_cast<uint4*>(stg_o + own_off + (((uint32_t)i ^ own_sw) << 4)) = w;
        }
        if (p.tma_out) {
          // 32 rows x 64 B in the TMA's 64-byte swizzle (chunk ^ ((row >> 1) & 3)): one bulk store per chunk, rows
          // beyond M clipped by the tensor map
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmC, stg_o, ocol0, m_base, bz);
            tma_store_commit();
          }
          ++oc;
          continue;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = 8 * i + co_r;
          const uint4 w = *reinterpret_cast<const uint4*>(stg_o + rr * 64 + ((co_j ^ ((rr >> 1) & 3)) << 4));
          if (co_ok[i]) *reinterpret_cast<uint4*>(out_b + co_out[i] + ocol0) = w;
        }
        __syncwarp();  // staging tile is reused by the next chunk
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (TWO) mbar_arrive_cluster(&tempty_bar[acc], 0);   // the leader's MMA thread owns the wait
        else mbar_arrive(&tempty_bar[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      if (STATS && m_base + lane < p.M)
        reinterpret_cast<float2*>(p.stats_out)[(size_t)(m_base + lane) * p.stats_np + n_blk * 2 + half] =
            make_float2(st1, st2);
    }
    if (p.tma_out && lane == 0) tma_store_wait_all();    // the last stores have landed before this CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();   // neither CTA may leave (or free TMEM) while its peer still works
  if (warp == 2) {
    tc_fence_after();
    if (TWO) tmem_dealloc_2cta(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// Bring-up / cross-check path: plain CUDA-core tiled GEMM with the same epilogue semantics.
// Selected only by force_simple (tests use it to tell a tcgen05 bug from a model-graph bug).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gemm_bf16_simple_kernel(const __nv_bfloat16* __restrict__ A, int lda,
                        const __nv_bfloat16* __restrict__ W, int ldw, const GemmDev p) {
  __shared__ float sa[16][17];
  __shared__ float sw[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m = blockIdx.y * 16 + ty;
  const int n = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += 16) {
    const int ka = k0 + tx;
    sa[ty][tx] = (m < p.M && ka < p.K) ? __bfloat162float(A[(size_t)m * lda + ka]) : 0.f;
    const int wn = blockIdx.x * 16 + ty;
    sw[ty][tx] = (wn < p.N && ka < p.K) ? __bfloat162float(W[(size_t)wn * ldw + ka]) : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += sa[ty][k] * sw[tx][k];
    __syncthreads();
  }
  if (m >= p.M || n >= p.N) return;
  int out_row = m, res_row = m;
  if (p.row_mode == VZ_ROWS_PATCH_EMBED) {
    const int img = m / p.rows_per;
    out_row = m + img + 1;
    res_row = 1 + (m - img * p.rows_per);
  } else if (p.row_mode == VZ_ROWS_RES_MOD) {
    res_row = m % p.rows_per;
  }
  float v = acc;
  if (p.ln_stats) {
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < p.ln_np; ++i) {
      s1 += p.ln_stats[((size_t)m * p.ln_np + i) * 2];
      s2 += p.ln_stats[((size_t)m * p.ln_np + i) * 2 + 1];
    }
    const float mu = p.ln_rms ? 0.f : s1 / (float)p.K;
    const float rstd = rsqrtf(fmaxf(s2 / (float)p.K - mu * mu, 0.f) + p.ln_eps);
    v = rstd * (v - (p.ln_rms ? 0.f : mu * p.ln_colsum[n])) + (p.bias ? p.bias[n] : 0.f);
  } else if (p.bias) {
    v += p.bias[n];
  }
  v = act_apply(v, p.act);
  if (p.residual) v += __bfloat162float(p.residual[(size_t)res_row * p.ldr + n]);
  p.out[(size_t)out_row * p.ldo + n] = __float2bfloat16_rn(v);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

// 3-D bf16 tensor map: dim0 = K (contiguous), dim1 = rows, dim2 = batch; box = {64, box_rows, 1}.
// A descriptor is a pure function of (address, shape, strides, box), and a forward pass asks for the same few
// hundred of them on every call (weights and workspace slices do not move), so they are cached per calling
// thread: the single-image step spent more host time in cuTensorMapEncodeTiled (356 calls) than in launches.
struct TmapKey {
  const void* base;
  long long bstride;
  int rows, cols, ld, box_rows, batch;
  bool operator==(const TmapKey& o) const {
    return base == o.base && bstride == o.bstride && rows == o.rows && cols == o.cols && ld == o.ld &&
           box_rows == o.box_rows && batch == o.batch;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= (size_t)k.rows * 0xC2B2AE3D27D4EB4Full + (size_t)k.cols * 0x165667B19E3779F9ull + (size_t)k.ld * 31 +
         (size_t)k.box_rows * 131 + (size_t)k.batch * 1313 + (size_t)k.bstride * 0x27D4EB2F165667C5ull;
    return h ^ (h >> 29);
  }
};

int make_tmap(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int box_rows, int batch,
              long long bstride) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return VZ_ERR_CUDA;
  if (batch <= 1) { batch = 1; bstride = (long long)rows * ld; }
  static const bool use_cache = []() { const char* e = getenv("VZ_TMAP_CACHE"); return !(e && e[0] == '0'); }();
  thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  const TmapKey key{base, bstride, rows, cols, ld, box_rows, batch};
  if (use_cache) {
    auto it = cache.find(key);
    if (it != cache.end()) { *tm = it->second; return VZ_OK; }
  }
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bstride * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
    // a thread that has made no runtime call yet (PyTorch's autograd worker entering a backward) has no
    // current driver context: let the runtime bind the primary context of the current device, then retry
    cudaFree(nullptr);
    r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = 100000 + (int)r;
    return VZ_ERR_CUDA;
  }
  if (use_cache) {
    if (cache.size() > 8192) cache.clear();     // a caller that keeps changing buffers: start over
    cache.emplace(key, *tm);
  }
  return VZ_OK;
}

// Output map of the epilogue's TMA stores: bf16 [batch][rows][cols], box = 32 columns x 32 rows in the 64-byte swizzle
// (the staging layout of an epilogue warp); rows beyond `rows` are clipped by the hardware.
int make_tmap_out(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int batch, long long bstride) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return VZ_ERR_CUDA;
  if (batch <= 1) { batch = 1; bstride = (long long)rows * ld; }
  static const bool use_cache = []() { const char* e = getenv("VZ_TMAP_CACHE"); return !(e && e[0] == '0'); }();
  thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  const TmapKey key{base, bstride, rows, cols, ld, -32, batch};
  if (use_cache) {
    auto it = cache.find(key);
    if (it != cache.end()) { *tm = it->second; return VZ_OK; }
  }
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bstride * 2};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT && r != CUDA_ERROR_NOT_INITIALIZED) break;
    cudaFree(nullptr);   // see make_tmap: bind the primary context, retry once
  }
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = 100000 + (int)r;
    return VZ_ERR_CUDA;
  }
  if (use_cache) {
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, *tm);
  }
  return VZ_OK;
}

template <int BN, int ACT, bool RES, bool LN, bool STATS, bool F32, bool TWO, bool BKN = false>
int launch_tc(const vz_gemm_args& a, const GemmDev& p_in, int num_sms, cudaStream_t st) {
  using C = Cfg<BN, TWO, RES>;
  CUtensorMap tmA, tmB, tmC;
  VZ_TRY(make_tmap(&tmA, a.A, a.M, a.K, a.lda, BM, a.batch, a.a_bstride));
  if (BKN) VZ_TRY(make_tmap(&tmB, a.W, a.K, a.N, a.ldw, 64, a.batch, a.w_bstride));   // [K, N]: 64 x 64 boxes
  else VZ_TRY(make_tmap(&tmB, a.W, a.N, a.K, a.ldw, TWO ? BN / 2 : BN, a.batch, a.w_bstride));
  VZ_ENSURE_DYN_SMEM((gemm_bf16_tcgen05_kernel<BN, ACT, RES, LN, STATS, F32, TWO, BKN>), C::SMEM_BYTES);
  GemmDev p = p_in;
  // output chunks through TMA stores: plain rows, bf16, no residual (whose landing buffers share the staging, see Cfg)
  // Band width of the tile walk, measured with ncu on the LLM's GEMMs (8 970 rows; DRAM MB read / us):
  //   gate|up, K = 4096, 112 n-blocks: 6: 2094 / 1423, 8: 1613 / 1399, 12: 1288 / 1376, 16: 1061 / 1368, 20: 1191 / 1374,
  //   28: 1892 / 1417 (operands: 308 MB); down, K = 14336, 16 n-blocks: 4: 1569 / 752, 8: 1159 / 693, 16: 1321 / 701.
  // i.e. a band whose W rows take ~33 MB (a quarter of the L2), but never fewer than 8 n-blocks; bands evened out.
  static const int group_env = []() { const char* e = getenv("VZ_GEMM_GROUP_N"); return e ? atoi(e) : 0; }();
  {
    const long per_nblk = (long)BN * a.K * 2;
    long g = (33L << 20) / (per_nblk > 0 ? per_nblk : 1);
    g = g < 8 ? 8 : (g > 32 ? 32 : g);
    const long nb = (p.num_n + g - 1) / g;
    g = (p.num_n + nb - 1) / nb;
    p.group_n = group_env > 0 ? group_env : (int)(g < 1 ? 1 : g);
  }
  static const int tma_out_on = []() { const char* e = getenv("VZ_GEMM_TMA_OUT"); return e ? atoi(e) : 1; }();
  p.tma_out = (tma_out_on && a.row_mode == VZ_ROWS_PLAIN && !F32 && !RES) ? 1 : 0;
  if (p.tma_out) {
    const int n_out = ACT == VZ_ACT_SWIGLU ? a.N / 2 : a.N;
    VZ_TRY(make_tmap_out(&tmC, a.out, a.M, n_out, a.ldo, a.batch, a.o_bstride));
  } else {
    tmC = tmA;   // unused
  }
  const long tiles = (long)p.num_m * p.num_n * p.batch;
  const int workers = TWO ? num_sms / 2 : num_sms;
  p.dp_tiles = (int)tiles;
  p.sk_tiles = 0;
  p.sk_ws = nullptr; p.sk_flags = nullptr; p.sk_epoch = 0;
  // Stream-K tail: when the tile count is not a multiple of the worker count, the last (partial) round
  // leaves most SMs idle; instead the k-blocks of the last full round + the remainder are split evenly.
  // Worth it when the balanced schedule beats the rounded one by more than the hand-over.  Measured on single-tile
  // problems (tools/gemm_small.py): the hand-over costs ~6.7 us = ~22 k-blocks, not the 6 assumed in round 1 -- the
  // ViT out-proj of ONE tile (40 tiles of 16 k-blocks) ran 18.5 us split over 148 SMs against 15.3 us as whole tiles.
  // (Also measured: loading the next worker's partial while the current one is added does not shorten it.)
  static const int sk_on = []() { const char* e = getenv("VZ_GEMM_SK"); return e ? atoi(e) : 1; }();
  const size_t sk_need = kSkFlagBytes + (size_t)num_sms * BM * BN * sizeof(float);
  int sk_workers = workers;
  if (sk_on && a.sk_ws && a.sk_ws_bytes >= sk_need && aligned16(a.sk_ws) && tiles % workers != 0 && p.num_k >= 8 &&
      tiles * p.num_k >= workers) {
    const long full = tiles / workers;
    // fewer tiles than workers: at most three workers per tile -- the finishing worker adds its partners' partial tiles
    // one after the other (the single-tile cross-attention scores, 10 tiles of 80 k-blocks, were cut into ~15 pieces
    // each: 35 us per launch).  Single-image call with a cap of 2 / 3 / 4 / 6 / 8 / none: 3.22 / 3.11 / 3.15 / 3.25 /
    // 3.27 / 3.41 ms (VZ_GEMM_SK_MAXSPLIT, 0 = no cap).
    static const int max_split = []() { const char* e = getenv("VZ_GEMM_SK_MAXSPLIT"); return e ? atoi(e) : 3; }();
    if (full == 0 && max_split > 0 && tiles * max_split < sk_workers) sk_workers = (int)(tiles * max_split);
    const double t_dp = (double)(full + 1) * p.num_k;
    const double t_sk = (double)tiles * p.num_k / sk_workers + 22.0;
    if (t_sk < 0.96 * t_dp) {
      p.sk_tiles = full == 0 ? (int)tiles : (int)(workers + tiles % workers);
      p.dp_tiles = (int)tiles - p.sk_tiles;
      p.sk_flags = reinterpret_cast<uint32_t*>(a.sk_ws);
      p.sk_ws = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(a.sk_ws) + kSkFlagBytes);
      p.sk_epoch = g_sk_epoch.fetch_add(1, std::memory_order_relaxed) + 1;
      if (p.sk_epoch == 0) p.sk_epoch = g_sk_epoch.fetch_add(1, std::memory_order_relaxed) + 1;
    }
  }
  const int grid = (p.sk_tiles > 0 ? sk_workers : (tiles < workers ? (int)tiles : workers)) * (TWO ? 2 : 1);
  ProfScope prof(VZ_PROF_GEMM, 2.0 * a.M * (double)a.N * a.K * p.batch, st);
  if (TWO) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VZ_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, ACT, RES, LN, STATS, F32, TWO, BKN>, tmA, tmB, tmC, p));
    count_launch();
  } else {
    gemm_bf16_tcgen05_kernel<BN, ACT, RES, LN, STATS, F32, TWO, BKN><<<grid, kThreads, C::SMEM_BYTES, st>>>(tmA, tmB, tmC, p);
    VZ_LAUNCH_CHECK();
  }
  return VZ_OK;
}

// pick the epilogue instantiation for the requested feature combination
template <int BN, bool TWO = false>
int launch_bn(const vz_gemm_args& a, const GemmDev& p, int num_sms, cudaStream_t st) {
  const int act = a.act;
  const bool res = a.residual != nullptr, ln = a.ln_stats != nullptr, stt = a.stats_out != nullptr, f32 = a.out_f32 != 0;
  if (a.w_is_kn) {   // [K, N] weights: plain epilogue only (the cross-attention P f product)
    if (act || res || ln || stt || f32 || BN == 192) return VZ_ERR_UNSUPPORTED;
    if constexpr (BN != 192) return launch_tc<BN, 0, false, false, false, false, TWO, true>(a, p, num_sms, st);
  }
#define VZ_GO(ACT, RES, LN, ST, F32) return launch_tc<BN, ACT, RES, LN, ST, F32, TWO>(a, p, num_sms, st)
  if (f32) { if (act || res || ln || stt) return VZ_ERR_UNSUPPORTED; VZ_GO(0, false, false, false, true); }
  if (stt) { if (!res || act || ln) return VZ_ERR_UNSUPPORTED; VZ_GO(0, true, false, true, false); }
  if (act == VZ_ACT_SWIGLU) {   // gate / up pairs of the LLM's MLP, RMSNorm fused or not
    if (res || stt || f32 || BN == 192) return VZ_ERR_UNSUPPORTED;
    if constexpr (BN != 192) {
      if (ln) VZ_GO(3, false, true, false, false);
      VZ_GO(3, false, false, false, false);
    }
  }
  if (ln) {
    if (res) return VZ_ERR_UNSUPPORTED;
    if (act == VZ_ACT_NONE) VZ_GO(0, false, true, false, false);
    if (act == VZ_ACT_QUICK_GELU) VZ_GO(1, false, true, false, false);
    VZ_GO(2, false, true, false, false);
  }
  if (res) {
    if (act == VZ_ACT_NONE) VZ_GO(0, true, false, false, false);
    if (act == VZ_ACT_QUICK_GELU) VZ_GO(1, true, false, false, false);
    VZ_GO(2, true, false, false, false);
  }
  if (act == VZ_ACT_NONE) VZ_GO(0, false, false, false, false);
  if (act == VZ_ACT_QUICK_GELU) VZ_GO(1, false, false, false, false);
  VZ_GO(2, false, false, false, false);
#undef VZ_GO
}

}  // namespace

// 2-D bf16 tensor map with 128-byte swizzle (box_cols * 2 must be 128 bytes); used by the attention kernel
int encode_tmap_2d_bf16(CUtensorMap* tm, const void* base, long long rows, int cols, int ld, int box_cols,
                        int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return VZ_ERR_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
    cudaFree(nullptr);
    r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = 100000 + (int)r;
    return VZ_ERR_CUDA;
  }
  return VZ_OK;
}

// 3-D bf16 tensor map {64 columns, box_rows rows, 1 batch} with 128-byte swizzle (the GEMM's operand map; the
// attention kernel stores its output tiles through one)
int encode_tmap_3d_bf16(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int box_rows, int batch,
                        long long bstride) {
  return make_tmap(tm, base, rows, cols, ld, box_rows, batch, bstride);
}

namespace {
// tile width the launcher will use for a problem (shared with the orchestration, which sizes the
// LayerNorm partial-statistics buffers from it)
int pick_tile_n(int M, int N, int batch, int num_sms) {
  static const int forced_bn = []() { const char* e = getenv("VZ_GEMM_BN"); return e ? atoi(e) : 0; }();
  const long num_m = (M + BM - 1) / BM;
  const long tiles256 = num_m * ((N + 255) / 256) * batch;
  const long tiles192 = num_m * ((N + 191) / 192) * batch;
  const double cost256 = (double)((tiles256 + num_sms - 1) / num_sms);
  const double cost192 = 0.78 * (double)((tiles192 + num_sms - 1) / num_sms);
  int bn = 128;
  // (measured: with the 2-CTA form available, 256-wide tiles beat 128x192 on every shape of the path,
  //  so 192 is only reachable through VZ_GEMM_BN)
  (void)cost256; (void)cost192;
  if (N % 256 == 0 && tiles256 >= num_sms) bn = 256;
  // small problems: when 128-wide tiles need a second, mostly empty round and 192-wide ones fit in one (the ViT fc1 of
  // ONE tile: 160 vs 110 tiles on 148 SMs), the wider tile wins
  const long tiles128 = num_m * ((N + 127) / 128) * batch;
  if (bn == 128 && N >= 192 && tiles128 > num_sms && tiles192 <= num_sms) bn = 192;
  if (forced_bn == 256 && N % 256 == 0) bn = 256;
  if (forced_bn == 192 && N >= 192) bn = 192;
  if (forced_bn == 128) bn = 128;
  return bn;
}
}  // namespace

size_t gemm_sk_workspace_bytes() {
  int dev = 0, num_sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (num_sms < 148) num_sms = 148;
  return kSkFlagBytes + (size_t)num_sms * BM * 256 * sizeof(float);
}

int gemm_stats_partials(int M, int N) {
  int dev = 0, num_sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  const int bn = pick_tile_n(M, N, 1, num_sms);
  return ((N + bn - 1) / bn) * 2;
}

int gemm_launch(const vz_gemm_args& a, cudaStream_t st) {
  if (!a.A || !a.W || !a.out) return VZ_ERR_BAD_ARG;
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return VZ_ERR_BAD_ARG;
  if ((a.lda & 7) || (a.ldw & 7) || (a.ldo & 7) || (a.residual && (a.ldr & 7))) return VZ_ERR_BAD_ARG;
  if (!aligned16(a.A) || !aligned16(a.W) || !aligned16(a.out) || (a.residual && !aligned16(a.residual)) ||
      (a.bias && !aligned16(a.bias)))
    return VZ_ERR_BAD_ARG;
  if (a.N % 32 != 0 || a.K % 8 != 0) return VZ_ERR_UNSUPPORTED;
  if (a.w_is_kn && (a.N % 64 != 0 || a.force_simple)) return VZ_ERR_UNSUPPORTED;
  if (a.row_mode != VZ_ROWS_PLAIN && a.rows_per <= 0) return VZ_ERR_BAD_ARG;
  const int batch = a.batch > 1 ? a.batch : 1;
  if (batch > 1 && ((a.a_bstride & 7) || (a.w_bstride & 7) || (a.o_bstride & 7) || (a.r_bstride & 7) ||
                    (a.bias_bstride & 3) || a.a_bstride <= 0 || a.w_bstride <= 0))
    return VZ_ERR_BAD_ARG;
  if (a.out_f32 && (a.residual || a.row_mode != VZ_ROWS_PLAIN || (a.ldo & 3))) return VZ_ERR_UNSUPPORTED;
  if ((batch > 1 || a.out_f32) && a.force_simple) return VZ_ERR_UNSUPPORTED;

  GemmDev p;
  p.out = reinterpret_cast<__nv_bfloat16*>(a.out);
  p.bias = a.bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a.residual);
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.ldo = a.ldo; p.ldr = a.ldr;
  p.act = a.act; p.row_mode = a.row_mode; p.rows_per = a.rows_per;
  p.batch = batch; p.out_f32 = a.out_f32;
  p.o_bstride = batch > 1 ? a.o_bstride : 0; p.r_bstride = batch > 1 ? a.r_bstride : 0;
  p.bias_bstride = batch > 1 ? a.bias_bstride : 0;
  p.ln_stats = a.ln_stats; p.ln_colsum = a.ln_colsum; p.ln_np = a.ln_np; p.ln_eps = a.ln_eps;
  p.ln_rms = a.ln_rms != 0;
  p.stats_out = a.stats_out; p.stats_np = a.stats_np;
  if (a.ln_stats && a.ln_np <= 0) return VZ_ERR_BAD_ARG;
  if (a.ln_stats && !a.ln_rms && (!a.ln_colsum || !a.bias || !aligned16(a.ln_colsum))) return VZ_ERR_BAD_ARG;
  if (a.act < 0 || a.act > VZ_ACT_SWIGLU) return VZ_ERR_BAD_ARG;
  if (a.act == VZ_ACT_SWIGLU && (a.N % 128 != 0 || a.row_mode != VZ_ROWS_PLAIN || a.force_simple || a.w_is_kn))
    return VZ_ERR_UNSUPPORTED;
  if (a.stats_out && (a.stats_np <= 0 || a.out_f32 || batch > 1)) return VZ_ERR_BAD_ARG;
  // the epilogue addresses rows with 32-bit element offsets
  if ((double)(a.M + a.M / 2 + 2) * a.ldo >= 2147483647.0 || (a.residual && (double)(a.M + 2) * a.ldr >= 2147483647.0))
    return VZ_ERR_UNSUPPORTED;
  p.num_m = (a.M + BM - 1) / BM;
  p.num_k = (a.K + BK - 1) / BK;

  if (a.force_simple) {
    p.num_n = 0;
    if (a.stats_out) return VZ_ERR_UNSUPPORTED;  // the debug path gets its statistics from vz::row_stats_launch
    dim3 grid((a.N + 15) / 16, (a.M + 15) / 16);
    gemm_bf16_simple_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(a.A), a.lda,
                                                  reinterpret_cast<const __nv_bfloat16*>(a.W), a.ldw, p);
    VZ_LAUNCH_CHECK();
    return VZ_OK;
  }

  int dev = 0, num_sms = 0;
  VZ_CUDA_CHECK(cudaGetDevice(&dev));
  VZ_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  // Tile width.  128x256 tiles run the tensor pipe at full rate and amortise the A re-reads best;
  // measured on B200, 128x128 tiles reach only ~0.68 of that rate, so they are used only when N is
  // not a multiple of 256 or the problem has fewer 256-wide tiles than SMs.  The persistent grid works
  // in rounds of num_sms tiles; for mid-sized problems (the Q-Former's M = 32*T rows) 128x192 tiles
  // need fewer, cheaper rounds.  (VZ_GEMM_BN=256|192|128 overrides, for experiments.)
  int bn = pick_tile_n(a.M, a.N, batch, num_sms);
  if (bn == 192 && a.w_is_kn) bn = 128;   // the [K, N] operand form has no 192-wide instantiation
  if (bn == 256) {
    p.num_n = a.N / 256;
    // 2-CTA (cta_group::2) form for the large problems: 256-row pair tiles, a third less smem traffic
    static const int two_cta = []() { const char* e = getenv("VZ_GEMM_2CTA"); return e ? atoi(e) : 1; }();
    const long pair_tiles = (long)((a.M + 255) / 256) * p.num_n * batch;
    static const long two_min = []() { const char* e = getenv("VZ_GEMM_2CTA_MIN"); return e ? atol(e) : 74L; }();
    if (two_cta && a.M >= 256 && pair_tiles >= two_min) {
      p.num_m = (a.M + 255) / 256;
      return launch_bn<256, true>(a, p, num_sms, st);
    }
    return launch_bn<256>(a, p, num_sms, st);
  }
  if (bn == 192) {
    p.num_n = (a.N + 191) / 192;
    return launch_bn<192>(a, p, num_sms, st);
  }
  p.num_n = (a.N + 127) / 128;
  return launch_bn<128>(a, p, num_sms, st);
}

}  // namespace vz

extern "C" long long vz_kernel_launches(void) { return vz::g_launches.load(); }

extern "C" int vz_profile(int enable) {
  std::lock_guard<std::mutex> lk(vz::g_prof.mu);
  vz::g_prof.on.store(enable != 0);
  vz::g_prof.used = 0;
  vz::g_prof.tag.clear();
  vz::g_prof.work.clear();
  return VZ_OK;
}

// Synchronises the recorded events and returns launches, summed milliseconds and summed work of one tag.
extern "C" int vz_profile_read(int tag, long long* launches, double* total_ms, double* total_work) {
  std::lock_guard<std::mutex> lk(vz::g_prof.mu);
  double ms = 0, wk = 0;
  long long cnt = 0;
  const size_t n = vz::g_prof.used / 2;
  for (size_t i = 0; i < n; ++i) {
    if (vz::g_prof.tag[i] != tag) continue;
    float t = 0;
    cudaError_t e = cudaEventSynchronize(vz::g_prof.pool[2 * i + 1]);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, vz::g_prof.pool[2 * i], vz::g_prof.pool[2 * i + 1]);
    if (e != cudaSuccess) return vz::cuda_fail(e);
    ms += t;
    wk += vz::g_prof.work[i];
    ++cnt;
  }
  if (launches) *launches = cnt;
  if (total_ms) *total_ms = ms;
  if (total_work) *total_work = wk;
  return VZ_OK;
}

extern "C" const char* vz_profile_tag_name(int tag) {
  static const char* names[VZ_PROF_COUNT] = {"gemm_bf16_tcgen05", "vit_attn_tc", "fuse", "preprocess_h", "preprocess_v",
                                             "preprocess_fused", "splice_scatter", "qattn", "softmax_rows", "layernorm",
                                             "text_gather", "other", "llm_attn_causal"};
  return (tag >= 0 && tag < VZ_PROF_COUNT) ? names[tag] : "?";
}

extern "C" int vz_gemm_profile(int enable) { return vz_profile(enable); }
extern "C" int vz_gemm_profile_read(long long* launches, double* total_ms, double* total_flops) {
  return vz_profile_read(VZ_PROF_GEMM, launches, total_ms, total_flops);
}

extern "C" int vz_gemm_stats_partials(int M, int N) { return vz::gemm_stats_partials(M, N); }

extern "C" size_t vz_gemm_sk_workspace_bytes(void) { return vz::gemm_sk_workspace_bytes(); }

extern "C" int vz_gemm_bf16(const vz_gemm_args* args, void* stream) {
  if (!args) return VZ_ERR_BAD_ARG;
  return vz::gemm_launch(*args, reinterpret_cast<cudaStream_t>(stream));
}
