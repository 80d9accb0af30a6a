// Convolutions of the full-batch step as tcgen05 implicit GEMMs (sm_100a).
//
//   conv_gemm_kernel : out[128-pixel tile, N_TILE] = sum_taps sum_cblocks A * B^T, persistent CTAs (one per SM),
//                      A = NHWC activation boxes fetched by 4-D TMA with a per-tap spatial shift (halo / padding comes
//                      from TMA out-of-bounds zero fill), B = weight rows, both K-major SWIZZLE_128B tiles, fp32
//                      accumulation in double-buffered TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//                      Serves conv forward, dgrad (stride 1; the 4 phases of stride 2 as tap groups of one launch), 1x1
//                      convs and the im2col'ed stem.  A pipeline stage holds the hi and lo bf16 planes of both
//                      operands; hi*hi + hi*lo + lo*hi are "stacked" instructions on the same stage (ConvGemmCfg).
//                      352 threads: TMA producer warp, MMA issuer warp (warp-converged, elected lane), 8 epilogue warps.
//                      One launch serves `ng` microbatch groups: the tile schedule is cut into SUPER-TILES
//                      (group, N tile, row r) = the M tiles r, r + rows, ... of one group, so that the per-CTA partial
//                      BatchNorm statistics -- and therefore every result -- do not depend on how many groups share
//                      the launch; group g multiplies with ITS weight rows (the perturbed weights of the FD pass), and
//                      the last CTA of a (group, N tile) finalises that BatchNorm's mean / rstd (no grid barrier).
//   wgrad_kernel     : out[g][s][co, (tap,ci)] = sum_pixels dY[pixel, co] * X[pixel + tap, ci]; both operands are
//                      MN-major (the channel dimension is contiguous in NHWC), split-K over the 128-pixel blocks of a
//                      group, up to 8 (tap, ci-block) accumulators of 64 TMEM columns per CTA; haloed X boxes on
//                      32x32 / 16x16.  The output layout IS the flat gradient's ("native" [co][tap][ci]): without
//                      split-K the epilogue writes the gradient itself, with split-K reduce_multi_kernel sums the
//                      splits of ALL layers in one launch in a fixed order.
//
// Reference call sites replaced: torch.nn.Conv2d forward (fullbatch/models/resnets.py:69-73,206-210,285-291) and its
// autograd backward (fullbatch/training/training.py:82, fullbatch/models/modules.py:230).
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/fullbatch_b200.h"
#include "fb_common.cuh"

namespace fb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("FB_PDL");  // opt-in: measured gain on B200 is within noise
    return e && e[0] == '1';
  }();
  return on;
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s failed: %s", what, cudaGetErrorString(e));
  return static_cast<int>(e);
}

// ---------------------------------------------------------------------------------------------------------------
// TMA descriptor encoding through the driver entry point (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int encode(void* blob, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return FB_ERR_DRIVER;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  alignas(64) CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu, box %u %u)", int(r), rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return FB_ERR_DRIVER;
  }
  memcpy(blob, &m, sizeof(m));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// conv_gemm_kernel: persistent, warp-specialised (TMA producer / MMA issuer / 8 epilogue warps), double-buffered TMEM
// ---------------------------------------------------------------------------------------------------------------
struct alignas(64) ConvGemmKParams {
  CUtensorMap a_maps[FB_MAX_A_MAPS];
  CUtensorMap b_maps[FB_MAX_B_MAPS];
  fb_tap taps[FB_MAX_TAPS];
  int n_taps, cblocks;
  int tile_w, tile_h, tile_n;
  int grid_h, grid_n;
  // Schedule.  A super-tile u = ((mg * n_tapgroups + tg) * rows + r) * n_tiles + j is the set of M tiles
  // m = r, r + rows, ... (< mtg) of microbatch group mg for tap group tg and N tile j; CTA c runs the super-tiles
  // c, c + gridDim.x, ...  `rows` only depends on the problem of ONE group, so the partial BatchNorm statistics of a
  // (group, r, j) are the same sums whatever ng is.
  int mtg;      // M tiles per microbatch group
  int n_tiles;  // N tiles
  int rows;     // super-tile rows per (group, N tile) = partial statistics rows
  int ng;       // microbatch groups
  int n_tapgroups;
  int group_tap0[4], group_taps[4];
  long long group_off[4];
  int b_group_rows;  // weight rows per microbatch group (0: shared)
  int reverse;
  int cta_pair;
  int dbg;   // FB_CONV_DBG (timing experiments only, results are wrong): 1 no A loads, 2 no B loads, 4 no stores, 8 no epilogue
  int halo;  // 1: the taps come in triples (dh = -1, 0, 1 at one dw) that share ONE haloed A box of tile_h + 2 rows
  float* out;
  long long out_sn, out_sh, out_sw;
  int accumulate;
  int n_total;
  // BatchNorm statistics of the output (optional)
  float* stats;           // [ng][rows][2][n_total]
  unsigned int* tickets;  // [ng][n_tiles]
  float *bn_mean, *bn_rstd, *bn_batch;
  double bn_count;  // pixels per group
  float bn_eps;
};

constexpr int kTileM = 128;                        // pixels per CTA tile == UMMA M
constexpr int kBlockK = 64;                        // bf16 elements per K block == one 128-byte swizzle row
constexpr int kATileBytes = kTileM * kBlockK * 2;  // 16 KiB
// Warp roles of the tensor-core kernels (352 threads): 0 = TMA producer, 1 = MMA issuer, 2-5 and 7-10 = epilogue (warp 6
// idles: a warp may only read the TMEM lanes 32*(warp%4)..+31, so each lane quarter has TWO epilogue warps (groups 0
// and 1) that take alternate 16-column chunks: the accumulator drain, which is fully exposed for the last tile of a CTA,
// is twice as fast).  ONE issuing warp: two warps issuing alternate stages reach the pipe's 64 clk / MMA in a probe, but
// the tensor pipe does not retire MMAs of different warps in a fixed order, the fp32 accumulation order then varies from
// run to run and the step is no longer bit-reproducible (round-1 finding).
constexpr int kThreads = 352;
constexpr int kEpiWarps = 8;
constexpr int kEpiStageBytes = kEpiWarps * kEpiWarpFloats * 4;  // 8 epilogue warps x 32 x 20 floats
__device__ __forceinline__ bool is_epilogue_warp(int warp) { return (warp >= 2 && warp <= 5) || warp >= 7; }
__device__ __forceinline__ int epilogue_group(int warp) { return warp >= 7 ? 1 : 0; }
__device__ __forceinline__ void epilogue_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }  // the 8 epilogue warps
constexpr int kSmemBudget = 227 * 1024 - 2048 - kEpiStageBytes;

__device__ __forceinline__ void tile_origin(int tile, int tile_h, int tile_n, int grid_h, int& n0, int& h0) {
  if (tile_n == 1) {
    const int per_img = grid_h / tile_h;
    n0 = tile / per_img;
    h0 = (tile % per_img) * tile_h;
  } else {
    n0 = tile * tile_n;
    h0 = 0;
  }
}

struct SuperTile {
  int mg, tg, r, j;
};
// u enumerates (group, tap group, row [pair], N tile); `rows` rows per (group, tap group, N tile), of which a CTA pair
// takes two at a time (sched_rows = rows / 2, r = 2q + rank)
__device__ __forceinline__ SuperTile decode_super(const ConvGemmKParams& p, int u, int total, int sched_rows, int pair,
                                                  int rank) {
  if (p.reverse) u = total - 1 - u;
  SuperTile s;
  s.j = u % p.n_tiles;
  u /= p.n_tiles;
  s.r = (u % sched_rows) * pair + rank;
  u /= sched_rows;
  s.tg = u % p.n_tapgroups;
  s.mg = u / p.n_tapgroups;
  return s;
}

// Flush of the per-lane column statistics collected by the epilogue over one super-tile (BatchNorm statistics fused
// into the producing convolution).  acc[c][0..3] / acc[c][4..7] of lane l are the sums / sums of squares of columns
// c*16 + 4*(l%4) .. +3 over the rows the lane stored (rows = lane/4 mod 8).  Fixed order: shuffle tree over the 8 row
// groups, then the eight epilogue warps through the staging patch; one partial row stats[row][0 = sum | 1 = sq][channel].
// Ends with all epilogue warps synchronised and the staging patch free again.
template <int N_TILE>
__device__ __forceinline__ void flush_column_stats(float* epi_stage, float (&acc)[N_TILE / 32][8], int q, int eg, int lane,
                                                   float* stats, long long row, int n_total, int n_tile0) {
  static_assert(2 * N_TILE <= kEpiWarpFloats, "staging patch too small for the statistics");
  float* sm = epi_stage + (eg * 4 + q) * kEpiWarpFloats;  // [0 = sum | 1 = sq][N_TILE] of this warp
#pragma unroll
  for (int cc = 0; cc < N_TILE / 32; ++cc) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[cc][j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 4) {
        const int col = lane * 4 + (j & 3);
        sm[(j >> 2) * N_TILE + (2 * cc + eg) * 16 + col] = v;        // this warp's chunk
        sm[(j >> 2) * N_TILE + (2 * cc + (eg ^ 1)) * 16 + col] = 0.f;  // the other group's chunk
      }
      acc[cc][j] = 0.f;
    }
  }
  epilogue_bar();
  const int t = (eg * 4 + q) * 32 + lane;
  for (int idx = t; idx < 2 * N_TILE; idx += 256) {
    const int which = idx / N_TILE, col = idx % N_TILE;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kEpiWarps; ++w) v += epi_stage[w * kEpiWarpFloats + idx];
    stats[(row * 2 + which) * n_total + n_tile0 + col] = v;
  }
  __threadfence();  // the partial row must be visible before this CTA takes its ticket
  epilogue_bar();
}

// The last CTA of a (group, N tile) reduces the `rows` partial rows in a fixed order (row slices in parallel, then the
// slices in order) and publishes mean / rstd (and the batch statistics for the running-stat EMA) of N_TILE channels.
template <int N_TILE>
__device__ __forceinline__ void finalize_bn_stats(float* epi_stage, const ConvGemmKParams& p, int mg, int n_tile0,
                                                  int t) {
  constexpr int kItems = N_TILE / 2;     // (sum | sq) x float4 column
  constexpr int kSlices = 256 / kItems;  // row slices per item: 8 / 4 / 2 for N_TILE 64 / 128 / 256
  static_assert(256 * 4 * 8 <= kEpiStageBytes, "finalize scratch does not fit the staging patch");
  double* scratch = reinterpret_cast<double*>(epi_stage);  // [kSlices][kItems][4]
  {
    const int item = t % kItems, slice = t / kItems;
    const int which = item / (N_TILE / 4), c4 = item % (N_TILE / 4);
    const float* src = p.stats + ((long long)mg * p.rows * 2 + which) * p.n_total + n_tile0 + c4 * 4;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
    for (int row = slice; row < p.rows; row += kSlices) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (long long)row * 2 * p.n_total));
      a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
    }
    double* dst = scratch + (slice * kItems + item) * 4;
    dst[0] = a0; dst[1] = a1; dst[2] = a2; dst[3] = a3;
  }
  epilogue_bar();
  if (t < N_TILE) {
    const int it = t / 4, comp = t % 4;
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int s = 0; s < kSlices; ++s) {
      s1 += scratch[(s * kItems + it) * 4 + comp];
      s2 += scratch[(s * kItems + N_TILE / 4 + it) * 4 + comp];
    }
    const double m = s1 / p.bn_count;
    double var = s2 / p.bn_count - m * m;
    var = var < 0.0 ? 0.0 : var;
    const long long o = (long long)mg * p.n_total + n_tile0 + t;
    p.bn_mean[o] = float(m);
    p.bn_rstd[o] = float(1.0 / sqrt(var + double(p.bn_eps)));
    if (p.bn_batch) {
      const double unbiased = p.bn_count > 1.0 ? var * p.bn_count / (p.bn_count - 1.0) : var;
      p.bn_batch[(long long)mg * 2 * p.n_total + n_tile0 + t] = float(m);
      p.bn_batch[(long long)mg * 2 * p.n_total + p.n_total + n_tile0 + t] = float(unbiased);
    }
  }
  epilogue_bar();
}

template <int N_TILE, int PA, int PB, bool CTA2 = false>
struct ConvGemmCfg {
  // CTA2: a pair of CTAs (cluster of 2, cta_group::2) runs M = 256 instructions; each CTA stages its own 128 pixels of
  // A and HALF of the rows of every B (weight) tile, so the weight traffic from L2 per CTA halves (the kernels are bound
  // by the L2 -> SM operand bandwidth, not by the tensor pipe)
  static constexpr int kBBytes = N_TILE * kBlockK * 2 / (CTA2 ? 2 : 1);
  static constexpr int kStageBytes = PA * kATileBytes + PB * kBBytes;
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiStageBytes + 1024 + 256;
  // An SS-mode tcgen05.mma re-reads its 128x16 A tile from shared memory (~64 clocks) whatever N is, so N = 64
  // instructions run the tensor pipe at half rate, and the pipe queues only ~2 instructions, so short instructions
  // expose the issue overhead between pipeline stages.  With split weights the hi and lo B tiles are adjacent in the
  // stage: ONE instruction with N = 2*N_TILE computes A*[B_hi;B_lo]^T into two N_TILE-column halves that the epilogue
  // adds ("stacked" mode).  64-wide tiles: both A planes are stacked (N = 128, picks up the tiny lo*lo term);
  // 128-wide tiles: A_hi is stacked (N = 256) and A_lo multiplies B_hi only (N = 128).
  // A CTA pair splits the N dimension of an instruction between the two CTAs' shared memories, which scrambles the
  // column order of a stacked operand: pairs issue one instruction per operand-plane combination into natural columns.
  static constexpr bool kStack = (PB == 2 && N_TILE <= 128 && !CTA2);
  static constexpr int kUmmaN = kStack ? 2 * N_TILE : N_TILE;
  static constexpr int kTmemCols = 2 * kUmmaN;
  // instructions per K = 16 step: stacked -> one per A plane; otherwise (a0,b0), (a0,b1) if PB == 2, (a1,b0) if PA == 2
  static constexpr int kCombos = kStack ? PA : ((PA == 2 && PB == 2) ? 3 : (PA * PB));
  static_assert(kStages >= 2, "stage does not fit twice into shared memory");
};

// MODE 3: independent CTAs, 3x3 / stride 1 over tiles of whole image rows: the three taps of a filter COLUMN read row-
// shifted windows of one haloed A box (tile_h + 2 rows, fetched once per (dw, channel block) instead of three times: the
// kernels sit at the L2 throughput cap, and A is 1/2 .. 2/3 of their operand bytes).  The operand area is an A ring of
// two haloed boxes and a B ring of per-tap weight tiles; the A box of a triple rides on the full barrier of its first
// B tile, so the MMA role waits for B slots only.
// MODE 0: independent CTAs.  MODE 1: CTA pairs (cta_group::2, see ConvGemmCfg).  MODE 2: clusters of two CTAs that work
// on two rows of the same (group, tap group, N tile) in lockstep and SHARE the weight tiles: each CTA fetches half of
// every B tile and multicasts it into both shared memories (half the weight bytes from L2 per CTA); everything else,
// including the instruction sequence and hence every accumulator bit, is as in MODE 0.
template <int N_TILE, int PA, int PB, int MODE>
__global__ void __launch_bounds__(kThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmKParams p) {
  constexpr bool CTA2 = MODE == 1;
  constexpr bool MCAST = MODE == 2;
  constexpr bool CLUSTER = MODE == 1 || MODE == 2;
  constexpr bool HALO = MODE == 3;
  using Cfg = ConvGemmCfg<N_TILE, PA, PB, CTA2>;
  constexpr int STAGES = HALO ? 8 : Cfg::kStages;  // HALO: upper bound of the B ring (the barrier arrays)
  constexpr int kOperandBytes = HALO ? kSmemBudget : Cfg::kStages * Cfg::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by an OFFSET from the __shared__ array (a round trip through uintptr_t makes the compiler lose
  // the address space: every access to the staging patch then becomes a generic LD.E / ST.E instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* epi_stage = reinterpret_cast<float*>(smem + kOperandBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOperandBytes + kEpiStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;  // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint64_t* a_empty = acc_empty + 2;        // [2] (HALO: the haloed A boxes)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 2);
  // HALO geometry: bytes of one haloed box of one plane, byte shift of one image row, B ring slots that fit
  const int halo_bytes = HALO ? (p.tile_h + 2) * p.tile_w * 128 : 0;
  const int halo_row_bytes = p.tile_w * 128;
  int nb = HALO ? (kSmemBudget - 2 * PA * halo_bytes) / (PB * Cfg::kBBytes) : STAGES;
  nb = nb > STAGES ? STAGES : nb;
  uint8_t* b_ring = smem + 2 * PA * halo_bytes;
  volatile uint32_t* last_flag = tmem_slot + 1;  // "this CTA took the last ticket of its (group, N tile)"

  // warp index through a shuffle: the compiler then knows it is warp-uniform, and the producer / MMA roles below run
  // warp-converged with ONE elected lane issuing, so that their operands live in uniform registers and the unrolled
  // tcgen05.mma block compiles to back-to-back UTCHMMA (a `lane == 0` branch costs ~13 instructions per MMA and made
  // the single issuing thread, not the tensor pipe, the bottleneck)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  // CTA pair: rank 0 (the leader) owns the full / accumulator-empty barriers and issues the MMAs of both CTAs
  const uint32_t rank = CLUSTER ? cluster_ctarank() : 0u;

  if (threadIdx.x == 32) {  // descriptor fetch overlaps the barrier / TMEM set-up
#pragma unroll
    for (int pl = 0; pl < PA; ++pl) tma_prefetch_desc(&p.a_maps[p.taps[0].phase * PA + pl]);
#pragma unroll
    for (int pl = 0; pl < PB; ++pl) tma_prefetch_desc(&p.b_maps[pl]);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], MCAST ? 2 : 1);  // multicast: a stage is free when BOTH CTAs' MMAs have read it
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a_empty[b], 1);
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], CTA2 ? 2 * kEpiWarps : kEpiWarps);  // one arrival per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTA2) tmem_alloc2(tmem_slot, Cfg::kTmemCols); else tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  tc_fence_before();
  if (CLUSTER) cluster_sync_all(); else __syncthreads();  // the peer's barriers must be initialised before any use
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  griddep_wait();    // everything above overlapped the predecessor's tail
  griddep_launch();

  // Schedule: a pair works on two super-tiles that differ only in their row (r = 2q + rank): same group, tap group and
  // N tile, hence the same weight tiles and the same number of pipeline stages, in lockstep.
  const int sched_rows = CLUSTER ? p.rows / 2 : p.rows;
  const int total_super = p.ng * p.n_tapgroups * sched_rows * p.n_tiles;
  const int sched_first = CLUSTER ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int sched_step = CLUSTER ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (HALO && warp == 0) {
    // ---------------- TMA producer, haloed A boxes ----------------
    int s = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int u = sched_first; u < total_super; u += sched_step) {
      const SuperTile sp = decode_super(p, u, total_super, sched_rows, 1, 0);
      const int b_row = sp.j * N_TILE + sp.mg * p.b_group_rows;
      const int tap0 = p.group_tap0[sp.tg], tap1 = tap0 + p.group_taps[sp.tg];
      for (int m = sp.r; m < p.mtg; m += p.rows) {
        int n0, h0;
        tile_origin(sp.mg * p.mtg + m, p.tile_h, p.tile_n, p.grid_h, n0, h0);
        for (int t = tap0; t < tap1; t += 3) {
          const fb_tap tap = p.taps[t];  // dh = -1 of the triple
          for (int cb = 0; cb < p.cblocks; ++cb) {
            mbar_wait(&a_empty[as], aphase ^ 1, 5);
#pragma unroll 1
            for (int j = 0; j < 3; ++j) {
              const int b_k0 = p.taps[t + j].b_k0;
              mbar_wait(&empty_bar[s], phase ^ 1, 1);
              if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[s], ((p.dbg & 2) ? 0 : PB * Cfg::kBBytes) +
                                                        ((j == 0 && !(p.dbg & 1)) ? PA * halo_bytes : 0));
                if (j == 0 && !(p.dbg & 1)) {
#pragma unroll
                  for (int pl = 0; pl < PA; ++pl)
                    tma_load_4d(smem + (as * PA + pl) * halo_bytes, &p.a_maps[tap.phase * PA + pl], &full_bar[s],
                                cb * kBlockK, tap.dw, h0 - 1, n0);
                }
#pragma unroll
                for (int pl = 0; pl < ((p.dbg & 2) ? 0 : PB); ++pl)
                  tma_load_2d(b_ring + (s * PB + pl) * Cfg::kBBytes, &p.b_maps[pl], &full_bar[s], b_k0 + cb * kBlockK,
                              b_row);
              }
              __syncwarp();
              if (++s == nb) {
                s = 0;
                phase ^= 1;
              }
            }
            as ^= 1;
            aphase ^= (as == 0) ? 1u : 0u;
          }
        }
      }
    }
  } else if (warp == 0) {
    // ---------------- TMA producer ----------------
    int s = 0;
    uint32_t phase = 0;
    for (int u = sched_first; u < total_super; u += sched_step) {
      const SuperTile sp = decode_super(p, u, total_super, sched_rows, CLUSTER ? 2 : 1, (int)rank);
      // a pair splits every weight tile: this CTA fetches rows [rank * N_TILE/2, +N_TILE/2) of it
      const int b_row = sp.j * N_TILE + sp.mg * p.b_group_rows + (CLUSTER ? (int)rank * (N_TILE / 2) : 0);
      const int tap0 = p.group_tap0[sp.tg], tap1 = tap0 + p.group_taps[sp.tg];
      for (int m = sp.r; m < p.mtg; m += p.rows) {
        int n0, h0;
        tile_origin(sp.mg * p.mtg + m, p.tile_h, p.tile_n, p.grid_h, n0, h0);
        for (int t = tap0; t < tap1; ++t) {
          const fb_tap tap = p.taps[t];
          for (int cb = 0; cb < p.cblocks; ++cb) {
            mbar_wait(&empty_bar[s], phase ^ 1, 1);
            if (elect_one()) {
              uint8_t* st = smem + s * Cfg::kStageBytes;
              if (CTA2) {
                // both CTAs' bytes are counted on the LEADER's full barrier (it issues the MMAs of the pair)
                const uint32_t bar = mapa_shared(smem_u32(&full_bar[s]), 0);
                if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::kStageBytes);
#pragma unroll
                for (int pl = 0; pl < PA; ++pl)
                  tma_load_4d_pair(st + pl * kATileBytes, &p.a_maps[tap.phase * PA + pl], bar, cb * kBlockK, tap.dw,
                                   h0 + tap.dh, n0);
#pragma unroll
                for (int pl = 0; pl < PB; ++pl)
                  tma_load_2d_pair(st + PA * kATileBytes + pl * Cfg::kBBytes, &p.b_maps[pl], bar,
                                   tap.b_k0 + cb * kBlockK, b_row);
              } else if (MCAST) {
                // own A tile + the full B tile: this CTA's half and the peer's land on the barrier of this stage
                mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
#pragma unroll
                for (int pl = 0; pl < PA; ++pl)
                  tma_load_4d(st + pl * kATileBytes, &p.a_maps[tap.phase * PA + pl], &full_bar[s], cb * kBlockK, tap.dw,
                              h0 + tap.dh, n0);
#pragma unroll
                for (int pl = 0; pl < PB; ++pl)
                  tma_load_2d_mcast(st + PA * kATileBytes + pl * Cfg::kBBytes + rank * (Cfg::kBBytes / 2), &p.b_maps[pl],
                                    &full_bar[s], tap.b_k0 + cb * kBlockK, b_row, (uint16_t)3);
              } else {
                mbar_arrive_expect_tx(&full_bar[s], ((p.dbg & 1) ? 0 : PA * kATileBytes) + ((p.dbg & 2) ? 0 : PB * Cfg::kBBytes));
#pragma unroll
                for (int pl = 0; pl < ((p.dbg & 1) ? 0 : PA); ++pl)
                  tma_load_4d(st + pl * kATileBytes, &p.a_maps[tap.phase * PA + pl], &full_bar[s], cb * kBlockK, tap.dw,
                              h0 + tap.dh, n0);
#pragma unroll
                for (int pl = 0; pl < ((p.dbg & 2) ? 0 : PB); ++pl)
                  tma_load_2d(st + PA * kATileBytes + pl * Cfg::kBBytes, &p.b_maps[pl], &full_bar[s],
                              tap.b_k0 + cb * kBlockK, b_row);
              }
            }
            __syncwarp();
            if (++s == STAGES) {
              s = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 && (rank == 0 || !CTA2)) {
    // ---------------- MMA issuer (of a pair: the leader CTA only) ----------------
    // A tcgen05.mma blocks its issuing thread until the tensor pipe has taken it, and nothing that thread executes
    // between two MMAs overlaps with them (every instruction between two MMAs adds its full latency).  The role
    // therefore runs warp-converged with an elected lane, so that the descriptors live in uniform registers and the
    // unrolled block is back-to-back UTCHMMA, and TWO pipeline stages are waited for and issued per round.
    constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
    constexpr int kC = Cfg::kCombos;
    //   stacked, PA == 2: c0 = A_lo, c1 = A_hi (both against [B_hi;B_lo]); A_lo only needs B_hi, so it runs at
    //     N = N_TILE except in the first K block of a tile, where it must initialise both halves
    //   not stacked: (a1,b0) first if PA == 2, then (a0,b0), then (a0,b1) if PB == 2
    constexpr bool kNarrowLo = Cfg::kStack && PA == 2;
    auto ap_of = [](int c) { return (PA == 2 && c == 0) ? 1 : 0; };
    auto bp_of = [](int c) { return (!Cfg::kStack && PB == 2 && c == kC - 1) ? 1 : 0; };
    constexpr uint32_t idesc_full = make_idesc_bf16(CTA2 ? 2 * kTileM : kTileM, Cfg::kUmmaN, 0, 0);
    constexpr uint32_t idesc_half = make_idesc_bf16(CTA2 ? 2 * kTileM : kTileM, N_TILE, 0, 0);
    const uint32_t smem0 = smem_u32(smem);
    auto issue_stage = [&](uint32_t tmem_d, int st, int ki) {
      const uint32_t a_lo = smem_desc_lo(smem0 + st * Cfg::kStageBytes, 16);
      const uint32_t b_lo = a_lo + ((PA * kATileBytes) >> 4);
#pragma unroll
      for (int c = 0; c < kC; ++c) {
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint32_t da = a_lo + ((ap_of(c) * kATileBytes + k * 32) >> 4);
          const uint32_t db = b_lo + ((bp_of(c) * Cfg::kBBytes + k * 32) >> 4);
          const uint32_t idesc = (kNarrowLo && c == 0 && ki != 0) ? idesc_half : idesc_full;
          if (CTA2) tc_mma2_bf16_lohi(tmem_d, da, db, desc_hi, desc_hi, idesc, (ki | c | k) != 0);
          else tc_mma_bf16_lohi(tmem_d, da, db, desc_hi, desc_hi, idesc, (ki | c | k) != 0);
        }
      }
      if (CTA2) tc_commit2(&empty_bar[st]);  // pair: frees the stage in BOTH CTAs
      else if (MCAST) tc_commit_mcast(&empty_bar[st], (uint16_t)3);
      else tc_commit(&empty_bar[st]);
    };
    // HALO: B slot st against the window of A box `as` shifted by j rows; the box is released after its third tap
    const uint32_t b_ring0 = smem_u32(b_ring);
    auto issue_halo = [&](uint32_t tmem_d, int st, int ki, int as, int j) {
      const uint32_t a_lo = smem_desc_lo(smem0 + as * PA * halo_bytes + j * halo_row_bytes, 16);
      const uint32_t b_lo = smem_desc_lo(b_ring0 + st * PB * Cfg::kBBytes, 16);
#pragma unroll
      for (int c = 0; c < kC; ++c) {
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint32_t da = a_lo + ((ap_of(c) * halo_bytes + k * 32) >> 4);
          const uint32_t db = b_lo + ((bp_of(c) * Cfg::kBBytes + k * 32) >> 4);
          const uint32_t idesc = (kNarrowLo && c == 0 && ki != 0) ? idesc_half : idesc_full;
          tc_mma_bf16_lohi(tmem_d, da, db, desc_hi, desc_hi, idesc, (ki | c | k) != 0);
        }
      }
      tc_commit(&empty_bar[st]);
      if (j == 2) tc_commit(&a_empty[as]);
    };
    int s = 0;
    uint32_t phase = 0;
    int tile_i = 0;
    int as = 0, aj = 0;  // HALO: current A box and tap of its triple
    for (int u = sched_first; u < total_super; u += sched_step) {
      const SuperTile sp = decode_super(p, u, total_super, sched_rows, CLUSTER ? 2 : 1, (int)rank);
      const int k_iters = p.group_taps[sp.tg] * p.cblocks;
      for (int m = sp.r; m < p.mtg; m += p.rows, ++tile_i) {
        const int buf = tile_i & 1;
        const uint32_t tmem_d = tmem_base + buf * Cfg::kUmmaN;
        mbar_wait(&acc_empty[buf], ((tile_i >> 1) & 1) ^ 1, 4);  // epilogue has drained this accumulator
        if (HALO) {
          // rounds of two B slots while the ring has room for two more in flight, else of one
          const int per_round = nb >= 4 ? 2 : 1;
          for (int ki = 0; ki < k_iters; ki += per_round) {
            const bool two = per_round == 2 && ki + 1 < k_iters;
            int s1 = s + 1;
            uint32_t phase1 = phase;
            if (s1 == nb) {
              s1 = 0;
              phase1 ^= 1;
            }
            mbar_wait(&full_bar[s], phase, 2);
            if (two) mbar_wait(&full_bar[s1], phase1, 2);
            tc_fence_after();
            const int as1 = aj == 2 ? (as ^ 1) : as, aj1 = aj == 2 ? 0 : aj + 1;
            if (elect_one()) {
              issue_halo(tmem_d, s, ki, as, aj);
              if (two) issue_halo(tmem_d, s1, ki + 1, as1, aj1);
              if (ki + per_round >= k_iters) tc_commit(&acc_full[buf]);
            }
            __syncwarp();
            if (two) {
              as = aj1 == 2 ? (as1 ^ 1) : as1;
              aj = aj1 == 2 ? 0 : aj1 + 1;
              s = s1 + 1;
              phase = phase1;
              if (s == nb) {
                s = 0;
                phase ^= 1;
              }
            } else {
              as = as1;
              aj = aj1;
              s = s1;
              phase = phase1;
            }
          }
          continue;
        }
        for (int ki = 0; ki < k_iters; ki += 2) {
          const bool pair = ki + 1 < k_iters;
          int s1 = s + 1;
          uint32_t phase1 = phase;
          if (s1 == STAGES) {
            s1 = 0;
            phase1 ^= 1;
          }
          mbar_wait(&full_bar[s], phase, 2);
          if (pair) mbar_wait(&full_bar[s1], phase1, 2);
          tc_fence_after();
          if (elect_one()) {
            issue_stage(tmem_d, s, ki);
            if (pair) issue_stage(tmem_d, s1, ki + 1);
            if (ki + 2 >= k_iters) {
              if (CTA2) tc_commit2(&acc_full[buf]); else tc_commit(&acc_full[buf]);
            }
          }
          __syncwarp();
          if (pair) {
            s = s1 + 1;
            phase = phase1;
            if (s == STAGES) {
              s = 0;
              phase ^= 1;
            }
          } else {
            s = s1;
            phase = phase1;
          }
        }
      }
    }
  } else if (is_epilogue_warp(warp)) {
    // ---------------- epilogue: TMEM -> registers -> smem transpose -> coalesced global stores (fp32 NHWC) ----------
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int eg = epilogue_group(warp);  // 16-column chunks c with c % 2 == eg
    const int et = (eg * 4 + q) * 32 + lane;  // index among the 256 epilogue threads
    const int r = q * 32 + lane;
    const int w = r % p.tile_w;
    const int h = (r / p.tile_w) % p.tile_h;
    const int n = r / (p.tile_w * p.tile_h);
    float* stage = epi_stage + (eg * 4 + q) * kEpiWarpFloats;
    float col_acc[N_TILE / 32][8];
#pragma unroll
    for (int cc = 0; cc < N_TILE / 32; ++cc)
#pragma unroll
      for (int j = 0; j < 8; ++j) col_acc[cc][j] = 0.f;
    const bool do_stat = p.stats != nullptr;
    int tile_i = 0;
    for (int u = sched_first; u < total_super; u += sched_step) {
      const SuperTile sp = decode_super(p, u, total_super, sched_rows, CLUSTER ? 2 : 1, (int)rank);
      const int n_tile0 = sp.j * N_TILE;
      for (int m = sp.r; m < p.mtg; m += p.rows, ++tile_i) {
        int n0, h0;
        tile_origin(sp.mg * p.mtg + m, p.tile_h, p.tile_n, p.grid_h, n0, h0);
        const bool valid = (n0 + n) < p.grid_n && (h0 + h) < p.grid_h && !(p.dbg & 4);
        const long long row_off = p.group_off[sp.tg] + (long long)(n0 + n) * p.out_sn +
                                  (long long)(h0 + h) * p.out_sh + (long long)w * p.out_sw + n_tile0;
        const int buf = tile_i & 1;
        mbar_wait(&acc_full[buf], (tile_i >> 1) & 1, 3);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < ((p.dbg & 8) ? 0 : N_TILE / 32); ++cc) {  // unrolled: col_acc must stay in registers
          const int c = 2 * cc + eg;
          uint32_t v[16];
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + buf * Cfg::kUmmaN + c * 16;
          tmem_ld_32x16(taddr, v);
          if (Cfg::kStack) {
            uint32_t v2[16];
            tmem_ld_32x16(taddr + N_TILE, v2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
          } else {
            tmem_ld_wait();
          }
          warp_store_rows16(stage, v, p.out, row_off, valid, c * 16, p.accumulate != 0, lane, col_acc[cc], do_stat);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTA2) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[buf]), 0));  // the leader waits for both CTAs
          else mbar_arrive(&acc_empty[buf]);
        }
      }
      if (do_stat) {
        // partial row of this super-tile, then a ticket: the last CTA of the (group, N tile) finalises its BatchNorm
        flush_column_stats<N_TILE>(epi_stage, col_acc, q, eg, lane, p.stats, (long long)sp.mg * p.rows + sp.r,
                                   p.n_total, n_tile0);
        if (et == 0) {
          unsigned int* ticket = p.tickets + sp.mg * p.n_tiles + sp.j;
          const unsigned int prev = atomicAdd(ticket, 1u);
          const bool last = prev == (unsigned int)(p.rows - 1);
          if (last) *ticket = 0u;  // every arrival of this launch is in: ready for the next launch
          __threadfence();
          *last_flag = last ? 1u : 0u;
        }
        epilogue_bar();
        const bool last = *last_flag != 0u;
        if (last) finalize_bn_stats<N_TILE>(epi_stage, p, sp.mg, n_tile0, et);
        epilogue_bar();  // last_flag may be rewritten
      }
    }
  }
  tc_fence_before();
  if (CLUSTER) cluster_sync_all(); else __syncthreads();  // no CTA may leave while its peer can still signal it
  if (warp == 1) {
    if (CTA2) tmem_dealloc2(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// Super-tile rows per (group, N tile).  At most one row per CTA of a wave (148 / n_tiles); if the group has more M tiles
// than that, prefer a row count that DIVIDES them (every super-tile then holds the same number of tiles, and the
// super-tiles of all groups of a launch spread evenly over the persistent CTAs) unless it would idle a quarter of the
// SMs.  Depends on one group's problem only.
// With statistics (sched_k_iters > 0: K iterations of one tile) every super-tile ends with a statistics flush + ticket
// (~3 us of epilogue time, fully exposed when a tile is one or two K blocks: 1x1 convolutions), and a launch of
// policy_groups groups (a constant of the engine, never the ng of a launch) has policy_groups times more super-tiles
// than one group needs to fill the SMs.  The row count then minimises a small model of the launch's makespan on the
// persistent grid, in units of K iterations: rounds of super-tiles x (tiles per super-tile x (k_iters + 2) + 6).
// B200: ResNet-152 32x32 64->256 1x1 forward 289 -> 110 us, conv family -16 %; ResNet-18 conv family -3 %.
static int stats_rows(int m_tiles_per_group, int n_tiles, int policy_groups, int sched_k_iters) {
  if (n_tiles < 1) n_tiles = 1;
  int cap = kNumSMs / n_tiles;
  if (cap < 1) cap = 1;
  if (sched_k_iters <= 0 || policy_groups <= 1) {
    if (m_tiles_per_group <= cap) return m_tiles_per_group;
    for (int d = cap; 4 * d >= 3 * cap; --d)
      if (m_tiles_per_group % d == 0) return d;
    return cap;
  }
  const long long tile_cost = sched_k_iters + 2, flush_cost = 6;
  const int r_max = m_tiles_per_group < cap ? m_tiles_per_group : cap;
  int best = 1, best_class = 3;
  long long best_cost = -1;
  for (int r = r_max; r >= 1; --r) {
    const long long rounds = ((long long)policy_groups * r * n_tiles + kNumSMs - 1) / kNumSMs;
    const long long tiles = (m_tiles_per_group + r - 1) / r;
    const long long cost = rounds * (tiles * tile_cost + flush_cost);
    // ties: equal super-tiles that CTA pairs can share (divisor, even), then divisors, then more rows
    const int cls = (m_tiles_per_group % r == 0) ? ((r % 2 == 0) ? 0 : 1) : 2;
    if (best_cost < 0 || cost < best_cost || (cost == best_cost && cls < best_class)) {
      best = r;
      best_cost = cost;
      best_class = cls;
    }
  }
  return best;
}

static bool cta_pairs_enabled() {
  static const bool on = [] {
    const char* e = getenv("FB_CTA2");  // FB_CTA2=0: single-CTA tiles everywhere
    return !(e && e[0] == '0');
  }();
  return on;
}

template <int N_TILE, int PA, int PB, int MODE>
static int launch_conv_gemm_cluster(const ConvGemmKParams& kp, cudaStream_t stream) {
  using Cfg = ConvGemmCfg<N_TILE, PA, PB, MODE == 1>;
  static int max_clusters = -1;
  if (max_clusters < 0) {
    FB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<N_TILE, PA, PB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::kSmemBytes));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kNumSMs & ~1);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    FB_CUDA(cudaOccupancyMaxActiveClusters(&n, conv_gemm_kernel<N_TILE, PA, PB, MODE>, &cfg));
    max_clusters = n > 0 ? n : 1;
  }
  const int total = kp.ng * kp.n_tapgroups * (kp.rows / 2) * kp.n_tiles;
  int clusters = max_clusters < kNumSMs / 2 ? max_clusters : kNumSMs / 2;
  if (clusters > total) clusters = total;
  FB_CUDA(launch_cluster(conv_gemm_kernel<N_TILE, PA, PB, MODE>, dim3(2 * clusters), dim3(kThreads), Cfg::kSmemBytes,
                         stream, 2u, kp));
  return 0;
}

constexpr int kHaloSmemBytes = kSmemBudget + kEpiStageBytes + 1024 + 256;

template <int N_TILE, int PA, int PB>
static int launch_conv_gemm_halo(const ConvGemmKParams& kp, cudaStream_t stream) {
  if constexpr (N_TILE <= 128) {
    static bool configured = false;
    if (!configured) {
      FB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<N_TILE, PA, PB, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kHaloSmemBytes));
      configured = true;
    }
    const int halo_bytes = (kp.tile_h + 2) * kp.tile_w * 128;
    const int nb = (kSmemBudget - 2 * PA * halo_bytes) / (PB * ConvGemmCfg<N_TILE, PA, PB, false>::kBBytes);
    FB_REQUIRE(nb >= 2, "fb_conv_gemm: the haloed boxes leave no room for two weight tiles");
    const int total_super = kp.ng * kp.n_tapgroups * kp.rows * kp.n_tiles;
    const int grid = total_super < kNumSMs ? total_super : kNumSMs;
    FB_CUDA(launch_pdl(conv_gemm_kernel<N_TILE, PA, PB, 3>, dim3(grid), dim3(kThreads), kHaloSmemBytes, stream, kp));
    return 0;
  } else {
    set_error("fb_conv_gemm: haloed A boxes need n_tile <= 128");
    return FB_ERR_UNSUPPORTED;
  }
}

template <int N_TILE, int PA, int PB>
static int launch_conv_gemm(const ConvGemmKParams& kp, cudaStream_t stream) {
  if (kp.halo) return launch_conv_gemm_halo<N_TILE, PA, PB>(kp, stream);
  if (kp.cta_pair == 2) return launch_conv_gemm_cluster<N_TILE, PA, PB, 2>(kp, stream);
  if (kp.cta_pair) return launch_conv_gemm_cluster<N_TILE, PA, PB, 1>(kp, stream);
  using Cfg = ConvGemmCfg<N_TILE, PA, PB, false>;
  static bool configured = false;
  if (!configured) {
    FB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<N_TILE, PA, PB, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::kSmemBytes));
    configured = true;
  }
  const int total_super = kp.ng * kp.n_tapgroups * kp.rows * kp.n_tiles;
  const int grid = total_super < kNumSMs ? total_super : kNumSMs;
  FB_CUDA(launch_pdl(conv_gemm_kernel<N_TILE, PA, PB, 0>, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, stream, kp));
  return 0;
}

// CTA pairs need two rows of the same (group, tap group, N tile) with equally many tiles each
static bool pair_ok(int m_tiles_per_group, int n_tiles, int policy_groups, int sched_k_iters) {
  if (!cta_pairs_enabled() || m_tiles_per_group <= 0 || n_tiles <= 0) return false;
  const int rows = stats_rows(m_tiles_per_group, n_tiles, policy_groups, sched_k_iters);
  return rows % 2 == 0 && m_tiles_per_group % rows == 0;
}

template <int N_TILE>
static int dispatch_conv_gemm(const ConvGemmKParams& kp, int pa, int pb, cudaStream_t stream) {
  if (pa == 2 && pb == 2) return launch_conv_gemm<N_TILE, 2, 2>(kp, stream);
  if (pa == 1 && pb == 2) return launch_conv_gemm<N_TILE, 1, 2>(kp, stream);
  if (pa == 1 && pb == 1) return launch_conv_gemm<N_TILE, 1, 1>(kp, stream);
  set_error("fb_conv_gemm: unsupported operand planes (%d, %d)", pa, pb);
  return FB_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad_kernel
// ---------------------------------------------------------------------------------------------------------------
struct alignas(64) WgradKParams {
  CUtensorMap dy_map;
  CUtensorMap x_maps[FB_MAX_A_MAPS];
  fb_wgrad_tap taps[FB_MAX_WGRAD_TAPS];
  int n_taps, cblocks, planes, slots_per_cta;
  int cout, cin;
  int tile_w, tile_h, tile_n;
  int grid_h, grid_n;
  int splits, pbg;  // splits per group, 128-pixel blocks per group
  float* out;
  long long out_gstride, out_sstride;
  int halo;  // 1: a B stage is ONE haloed X box per plane ((tile_h + 2) rows) shared by the CTA's three dh taps
};

constexpr int kWgAStages = 2;
constexpr int kWgBStages = 8;              // 16 KiB X tiles; a ring stage is `planes` consecutive tiles
constexpr int kWgABytes = 2 * kATileBytes;  // two 64-channel chunks of dY: [chunk][128 pixels][64 co]
constexpr int kWgBBytes = kATileBytes;      // [128 pixels][64 ci]
constexpr int kWgTmemCols = 512;

__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ WgradKParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by an OFFSET from the __shared__ array (a round trip through uintptr_t makes the compiler lose
  // the address space: every access to the staging patch then becomes a generic LD.E / ST.E instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // operand region = dY ring (2 stages of one or two 64-channel chunks) followed by the X ring, which takes the rest
  constexpr int kWgOperandBytes = kWgAStages * kWgABytes + kWgBStages * kWgBBytes;
  const int a_stage_bytes = (p.cout > 64) ? kWgABytes : kATileBytes;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kWgAStages * a_stage_bytes;
  const int b_region_bytes = kWgOperandBytes - kWgAStages * a_stage_bytes;
  float* epi_stage = reinterpret_cast<float*>(smem + kWgOperandBytes);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + kWgOperandBytes + kEpiStageBytes);
  uint64_t* a_empty = a_full + kWgAStages;
  uint64_t* b_full = a_empty + kWgAStages;
  uint64_t* b_empty = b_full + kWgBStages;
  uint64_t* accum_bar = b_empty + kWgBStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform roles, see conv_gemm_kernel
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgAStages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kWgBStages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kWgTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  griddep_wait();
  griddep_launch();

  const int co0 = blockIdx.x * 128;
  const int slot0 = blockIdx.y * p.slots_per_cta;
  const int n_slots_total = p.n_taps * p.cblocks;
  const int n_slots = min(p.slots_per_cta, n_slots_total - slot0);
  const int mg = blockIdx.z / p.splits;    // microbatch group
  const int split = blockIdx.z % p.splits;  // split of the group's pixel blocks
  const int halo_box_bytes = (p.tile_h + 2) * p.tile_w * 128;  // haloed X box of one plane
  const int halo_row_bytes = p.tile_w * 128;                   // one image row = one dh shift
  int b_stages = b_region_bytes / (p.planes * (p.halo ? halo_box_bytes : kWgBBytes));
  b_stages = b_stages > kWgBStages ? kWgBStages : b_stages;  // the barrier arrays hold kWgBStages entries
  const int pb0 = mg * p.pbg + int((long long)split * p.pbg / p.splits);
  const int pb1 = mg * p.pbg + int((long long)(split + 1) * p.pbg / p.splits);
  const int npb = pb1 - pb0;
  const bool two_chunks = (co0 + 64) < p.cout;
  const int k_total = p.n_taps * p.cin;

  if (warp == 0) {
    // ---------------- TMA producer (warp-converged, one elected lane issues) ----------------
    int as = 0, bs = 0;
    uint32_t aphase = 0, bphase = 0;
    for (int pb = pb0; pb < pb1; ++pb) {
      int n0, h0;
      tile_origin(pb, p.tile_h, p.tile_n, p.grid_h, n0, h0);
      mbar_wait(&a_empty[as], aphase ^ 1, 11);
      if (elect_one()) {
        uint8_t* sa = smem_a + as * a_stage_bytes;
        mbar_arrive_expect_tx(&a_full[as], two_chunks ? kWgABytes : kATileBytes);
        tma_load_4d(sa, &p.dy_map, &a_full[as], co0, 0, h0, n0);
        if (two_chunks) tma_load_4d(sa + kATileBytes, &p.dy_map, &a_full[as], co0 + 64, 0, h0, n0);
      }
      __syncwarp();
      if (++as == kWgAStages) {
        as = 0;
        aphase ^= 1;
      }
      if (p.halo) {  // one haloed box per plane serves the three dh taps of this CTA
        const fb_wgrad_tap tap = p.taps[slot0 % p.n_taps];
        const int cb = slot0 / p.n_taps;
        mbar_wait(&b_empty[bs], bphase ^ 1, 12);
        if (elect_one()) {
          mbar_arrive_expect_tx(&b_full[bs], p.planes * halo_box_bytes);
          for (int pl = 0; pl < p.planes; ++pl)
            tma_load_4d(smem_b + (bs * p.planes + pl) * halo_box_bytes, &p.x_maps[pl], &b_full[bs], cb * kBlockK, tap.dw,
                        h0 - 1, n0);
        }
        __syncwarp();
        if (++bs == b_stages) {
          bs = 0;
          bphase ^= 1;
        }
        continue;
      }
      for (int j = 0; j < n_slots; ++j) {
        const int s = slot0 + j;
        const fb_wgrad_tap tap = p.taps[s % p.n_taps];
        const int cb = s / p.n_taps;
        mbar_wait(&b_empty[bs], bphase ^ 1, 12);
        if (elect_one()) {
          mbar_arrive_expect_tx(&b_full[bs], p.planes * kWgBBytes);
          for (int pl = 0; pl < p.planes; ++pl)
            tma_load_4d(smem_b + (bs * p.planes + pl) * kWgBBytes, &p.x_maps[tap.phase * p.planes + pl], &b_full[bs],
                        cb * kBlockK, tap.dw, h0 + tap.dh, n0);
        }
        __syncwarp();
        if (++bs == b_stages) {
          bs = 0;
          bphase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: one elected lane, see conv_gemm_kernel ------------
    // both operands MN-major; the hi and lo X tiles of a stage are adjacent, so ONE instruction with N = 64*planes
    // computes dY^T*[X_hi, X_lo] into two 64-column halves that the epilogue adds (an SS-mode MMA re-reads its
    // 128x16 A tile per instruction, N = 64 would run the tensor pipe at half rate).
    // K = 16 pixels = 16 rows of 128 bytes; MN chunks of 64 channels are kATileBytes apart (LBO), groups of 8 pixel
    // rows are 1024 bytes apart (SBO).
    const uint32_t idesc = make_idesc_bf16(128, 64 * p.planes, 1, 1);
    constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
    const uint32_t smem_a0 = smem_u32(smem_a), smem_b0 = smem_u32(smem_b);
    int bs = 0;
    uint32_t bphase = 0;
    for (int pb = 0; pb < npb; ++pb) {
      const int as = pb % kWgAStages;
      mbar_wait(&a_full[as], (pb / kWgAStages) & 1, 13);
      const uint32_t a_lo = smem_desc_lo(smem_a0 + as * a_stage_bytes, kATileBytes);
      if (p.halo) {
        // the three dh slots of a pixel block share ONE X stage (the hi / lo planes are one box apart, slot jj is the
        // view shifted by jj rows): one wait / elect / commit round per pixel block (24 MMAs)
        mbar_wait(&b_full[bs], bphase, 14);
        const uint32_t b0_lo = smem_desc_lo(smem_b0 + bs * p.planes * halo_box_bytes, halo_box_bytes);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) {
            const uint32_t b_lo = b0_lo + ((jj * halo_row_bytes) >> 4);
            const uint32_t tmem_d = tmem_base + jj * 64 * p.planes;
#pragma unroll
            for (int k = 0; k < kTileM / 16; ++k)
              tc_mma_bf16_lohi(tmem_d, a_lo + ((k * 2048) >> 4), b_lo + ((k * 2048) >> 4), desc_hi, desc_hi, idesc,
                               (pb != 0 || k != 0) ? 1u : 0u);
          }
          tc_commit(&b_empty[bs]);
          tc_commit(&a_empty[as]);
          if (pb == npb - 1) tc_commit(accum_bar);
        }
        __syncwarp();
        if (++bs == b_stages) {
          bs = 0;
          bphase ^= 1;
        }
        continue;
      }
      for (int j = 0; j < n_slots; ++j) {
        mbar_wait(&b_full[bs], bphase, 14);
        const uint32_t b_lo = smem_desc_lo(smem_b0 + bs * p.planes * kWgBBytes, kATileBytes);
        const uint32_t tmem_d = tmem_base + j * 64 * p.planes;
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kTileM / 16; ++k)
            tc_mma_bf16_lohi(tmem_d, a_lo + ((k * 2048) >> 4), b_lo + ((k * 2048) >> 4), desc_hi, desc_hi, idesc,
                             (pb != 0 || k != 0) ? 1u : 0u);
          tc_commit(&b_empty[bs]);
          if (j == n_slots - 1) tc_commit(&a_empty[as]);
          if (pb == npb - 1 && j == n_slots - 1) tc_commit(accum_bar);
        }
        __syncwarp();
        if (++bs == b_stages) {
          bs = 0;
          bphase ^= 1;
        }
      }
    }
  } else if (is_epilogue_warp(warp)) {
    const int q = warp & 3;
    const int eg = epilogue_group(warp);  // 16-column chunks c with c % 2 == eg
    const int co = co0 + q * 32 + lane;
    const bool valid = co < p.cout;
    const long long row_off = (long long)mg * p.out_gstride + (long long)split * p.out_sstride + (long long)co * k_total;
    mbar_wait(accum_bar, 0, 15);
    tc_fence_after();
    for (int j = 0; j < n_slots; ++j) {
      const int s = slot0 + j;
      const int tap = p.taps[s % p.n_taps].k_index;  // filter position = column block of the output matrix
      const int cb = s / p.n_taps;
#pragma unroll 1
      for (int c = eg; c < 4; c += 2) {
        uint32_t v[16];
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + j * 64 * p.planes + c * 16;
        tmem_ld_32x16(taddr, v);
        if (p.planes == 2) {
          uint32_t v2[16];
          tmem_ld_32x16(taddr + 64, v2);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] = __float_as_uint(__uint_as_float(v[t]) + __uint_as_float(v2[t]));
        } else {
          tmem_ld_wait();
        }
        warp_store_rows16(epi_stage + (eg * 4 + q) * kEpiWarpFloats, v, p.out, row_off, valid,
                          tap * p.cin + cb * kBlockK + c * 16, false, lane);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kWgTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
// reduce_multi: deterministic split-K reduction of every layer's weight gradient in one launch
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reduce_multi_kernel(const fb_reduce_entry* __restrict__ table, int n,
                                                           float* __restrict__ dst_base, long long dst_gstride) {
  griddep_wait();
  griddep_launch();
  int lo = 0, hi = n - 1;
  while (lo < hi) {  // last entry with block_start <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].block_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const fb_reduce_entry e = table[lo];
  const int g = blockIdx.y;
  const long long idx = ((long long)(blockIdx.x - e.block_start) * 256 + threadIdx.x) * e.vec;
  if (idx >= (long long)e.rows * e.cols) return;
  const int r = int(idx / e.cols), c = int(idx % e.cols);
  const float* src = e.src + (long long)g * e.src_gstride + (long long)r * e.src_ld + c;
  float* dst = dst_base + (long long)g * dst_gstride + e.dst_off + (long long)r * e.dst_ld + c;
  if (e.vec == 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    for (; s + 8 <= e.splits; s += 8) {  // 8 loads in flight, summed in the fixed order s = 0, 1, ...
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldcg(reinterpret_cast<const float4*>(src + (s + j) * e.src_sstride));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w;
      }
    }
    for (; s < e.splits; ++s) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(src + s * e.src_sstride));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(dst) = acc;
  } else {
    float acc = 0.f;
    for (int s = 0; s < e.splits; ++s) acc += __ldcg(src + s * e.src_sstride);
    *dst = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// weight prep: OIHW fp32 -> bf16 hi/lo GEMM operands
// ---------------------------------------------------------------------------------------------------------------
// block = 32 co x 32 ci x taps, 256 threads
__global__ void weight_prep_kernel(const float* __restrict__ w, int cout, int cin, int taps,
                                   __nv_bfloat16* __restrict__ wf_hi, __nv_bfloat16* __restrict__ wf_lo, long long ld_f,
                                   __nv_bfloat16* __restrict__ wd_hi, __nv_bfloat16* __restrict__ wd_lo,
                                   long long ld_d) {
  extern __shared__ float wtile[];  // [32 co][32 ci * taps + 1]
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  const int row_len = 32 * taps;
  const int pitch = row_len + 1;
  for (int i = threadIdx.x; i < 32 * row_len; i += blockDim.x) {
    const int co = i / row_len, r = i % row_len;
    wtile[co * pitch + r] = w[((long long)(co0 + co) * cin + ci0) * taps + r];
  }
  __syncthreads();
  // wf[co][tap][ci]
  for (int i = threadIdx.x; i < 32 * row_len; i += blockDim.x) {
    const int ci = i % 32, tap = (i / 32) % taps, co = i / (32 * taps);
    const float v = wtile[co * pitch + ci * taps + tap];
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    const long long o = (long long)(co0 + co) * ld_f + (long long)tap * cin + ci0 + ci;
    wf_hi[o] = hi;
    if (wf_lo) wf_lo[o] = lo;
  }
  if (wd_hi) {
    // wd[ci][tap][co]
    for (int i = threadIdx.x; i < 32 * row_len; i += blockDim.x) {
      const int co = i % 32, tap = (i / 32) % taps, ci = i / (32 * taps);
      const float v = wtile[co * pitch + ci * taps + tap];
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      const long long o = (long long)(ci0 + ci) * ld_d + (long long)tap * cout + co0 + co;
      wd_hi[o] = hi;
      if (wd_lo) wd_lo[o] = lo;
    }
  }
}

// small-cin (stem) variant: wf[co][k] = w[co][k], k = ci*taps + tap, row stride ld_f (padding columns stay untouched)
__global__ void weight_prep_direct_kernel(const float* __restrict__ w, int cout, int k, __nv_bfloat16* __restrict__ wf_hi,
                                          __nv_bfloat16* __restrict__ wf_lo, long long ld_f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * k) return;
  const int co = i / k, kk = i % k;
  __nv_bfloat16 hi, lo;
  split_bf16(w[i], hi, lo);
  wf_hi[co * ld_f + kk] = hi;
  if (wf_lo) wf_lo[co * ld_f + kk] = lo;
}

// ---------------------------------------------------------------------------------------------------------------
// all conv weights of the network in ONE launch (grid.y = microbatch group): table of per-layer descriptors, block ->
// layer by binary search.  With `grad` the operands are those of the perturbed point of the group,
//   w' = theta + step * (bs * grad_native + acc * pre_native),  step = scale * eps_n[g]
// (modules.py:217-226 / :279-286): theta' is never written for conv weights.
// ---------------------------------------------------------------------------------------------------------------
struct WprepPerturb {
  const float* grad;  // group's native gradient (nullptr: plain theta)
  const float* pre;   // native pre_grads (nullptr if acc == 0)
  float step, bs, acc;
};

__device__ __forceinline__ void weight_prep_tile(const float* __restrict__ w, int cout, int cin, int taps, int co0,
                                                 int ci0, __nv_bfloat16* __restrict__ wf_hi,
                                                 __nv_bfloat16* __restrict__ wf_lo, long long ld_f,
                                                 __nv_bfloat16* __restrict__ wd_hi, __nv_bfloat16* __restrict__ wd_lo,
                                                 long long ld_d, float* wtile, const WprepPerturb& pt) {
  const int row_len = 32 * taps;
  const int pitch = row_len + 1;
  for (int i = threadIdx.x; i < 32 * row_len; i += blockDim.x) {
    const int co = i / row_len, r = i % row_len;
    wtile[co * pitch + r] = w[((long long)(co0 + co) * cin + ci0) * taps + r];
  }
  if (pt.grad) {
    __syncthreads();
    // native gradient [co][tap][ci]: 32 consecutive ci per (co, tap) are contiguous
    for (int i = threadIdx.x; i < 32 * row_len; i += blockDim.x) {
      const int ci = i % 32, tap = (i / 32) % taps, co = i / (32 * taps);
      const long long o = ((long long)(co0 + co) * taps + tap) * cin + ci0 + ci;
      float v = pt.bs * pt.grad[o];
      if (pt.pre) v += pt.acc * pt.pre[o];
      float* dst = wtile + co * pitch + ci * taps + tap;
      *dst = *dst + pt.step * v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * row_len; i += blockDim.x) {  // two ci per thread -> 4-byte stores
    const int ci = (i % 16) * 2, tap = (i / 16) % taps, co = i / (16 * taps);
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(wtile[co * pitch + ci * taps + tap], h0, l0);
    split_bf16(wtile[co * pitch + (ci + 1) * taps + tap], h1, l1);
    const long long o = (long long)(co0 + co) * ld_f + (long long)tap * cin + ci0 + ci;
    *reinterpret_cast<__nv_bfloat162*>(wf_hi + o) = __halves2bfloat162(h0, h1);
    if (wf_lo) *reinterpret_cast<__nv_bfloat162*>(wf_lo + o) = __halves2bfloat162(l0, l1);
  }
  if (wd_hi) {
    for (int i = threadIdx.x; i < 16 * row_len; i += blockDim.x) {
      const int co = (i % 16) * 2, tap = (i / 16) % taps, ci = i / (16 * taps);
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(wtile[co * pitch + ci * taps + tap], h0, l0);
      split_bf16(wtile[(co + 1) * pitch + ci * taps + tap], h1, l1);
      const long long o = (long long)(ci0 + ci) * ld_d + (long long)tap * cout + co0 + co;
      *reinterpret_cast<__nv_bfloat162*>(wd_hi + o) = __halves2bfloat162(h0, h1);
      if (wd_lo) *reinterpret_cast<__nv_bfloat162*>(wd_lo + o) = __halves2bfloat162(l0, l1);
    }
  }
}

__global__ void __launch_bounds__(256) weight_prep_multi_kernel(const float* __restrict__ theta,
                                                                const fb_wprep_entry* __restrict__ table, int n,
                                                                const float* __restrict__ grad, long long grad_gstride,
                                                                const float* __restrict__ pre, float bs, float acc,
                                                                float scale, const float* __restrict__ scal,
                                                                int eps_base) {
  extern __shared__ float wtile[];
  griddep_wait();
  griddep_launch();
  int lo = 0, hi = n - 1;
  while (lo < hi) {  // last entry with block_start <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].block_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const fb_wprep_entry e = table[lo];
  const int g = blockIdx.y;
  const int b = blockIdx.x - e.block_start;
  const float* w = theta + e.w_offset;
  WprepPerturb pt;
  pt.grad = grad ? grad + (long long)g * grad_gstride + e.w_offset : nullptr;
  pt.pre = (pre && acc != 0.f) ? pre + e.w_offset : nullptr;
  pt.step = grad ? scale * scal[eps_base + g] : 0.f;
  pt.bs = bs;
  pt.acc = acc;
  __nv_bfloat16* wf_hi = static_cast<__nv_bfloat16*>(e.wf_hi) + (long long)g * e.wf_gstride;
  __nv_bfloat16* wf_lo = e.wf_lo ? static_cast<__nv_bfloat16*>(e.wf_lo) + (long long)g * e.wf_gstride : nullptr;
  if (e.cin % 32 != 0) {  // stem: wf[co][k] = w[co][k], k = ci*taps + tap (native layout == OIHW)
    const int k = e.cin * e.taps;
    for (int i = b * 256 + threadIdx.x; i < e.cout * k; i += e.n_blocks * 256) {
      float v = w[i];
      if (pt.grad) {
        float d = pt.bs * pt.grad[i];
        if (pt.pre) d += pt.acc * pt.pre[i];
        v = v + pt.step * d;
      }
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      wf_hi[(i / k) * e.ld_f + i % k] = h;
      if (wf_lo) wf_lo[(i / k) * e.ld_f + i % k] = l;
    }
    return;
  }
  __nv_bfloat16* wd_hi = e.wd_hi ? static_cast<__nv_bfloat16*>(e.wd_hi) + (long long)g * e.wd_gstride : nullptr;
  __nv_bfloat16* wd_lo = e.wd_lo ? static_cast<__nv_bfloat16*>(e.wd_lo) + (long long)g * e.wd_gstride : nullptr;
  const int ci_blocks = e.cin / 32;
  weight_prep_tile(w, e.cout, e.cin, e.taps, (b / ci_blocks) * 32, (b % ci_blocks) * 32, wf_hi, wf_lo, e.ld_f, wd_hi,
                   wd_lo, e.ld_d, wtile, pt);
}

}  // namespace fb

// =================================================================================================================
// C ABI
// =================================================================================================================
using namespace fb;

extern "C" int fb_version(void) { return 200; }

extern "C" int fb_last_error(char* buf, size_t n) {
  if (buf && n) {
    strncpy(buf, g_err, n - 1);
    buf[n - 1] = 0;
  }
  return static_cast<int>(strlen(g_err));
}

extern "C" int fb_tmap_encode_act4d(void* host_blob, const void* base, int c, int w, int h, int n, int64_t stride_w,
                                    int64_t stride_h, int64_t stride_n, int box_c, int box_w, int box_h, int box_n) {
  FB_REQUIRE(host_blob && base, "fb_tmap_encode_act4d: null pointer");
  FB_REQUIRE(box_c == 64, "fb_tmap_encode_act4d: box_c must be 64 (one 128-byte swizzle row), got %d", box_c);
  FB_REQUIRE(box_w >= 1 && box_w <= 256 && box_h >= 1 && box_h <= 256 && box_n >= 1 && box_n <= 256,
             "fb_tmap_encode_act4d: bad box %d %d %d", box_w, box_h, box_n);
  FB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (stride_w * 2) % 16 == 0 && (stride_h * 2) % 16 == 0 &&
                 (stride_n * 2) % 16 == 0,
             "fb_tmap_encode_act4d: base and strides must be 16-byte aligned");
  cuuint64_t dims[4] = {cuuint64_t(c), cuuint64_t(w), cuuint64_t(h), cuuint64_t(n)};
  cuuint64_t strides[3] = {cuuint64_t(stride_w * 2), cuuint64_t(stride_h * 2), cuuint64_t(stride_n * 2)};
  cuuint32_t box[4] = {cuuint32_t(box_c), cuuint32_t(box_w), cuuint32_t(box_h), cuuint32_t(box_n)};
  return encode(host_blob, base, 4, dims, strides, box);
}

extern "C" int fb_tmap_encode_mat2d(void* host_blob, const void* base, int k, int rows, int64_t ld, int box_k,
                                    int box_rows) {
  FB_REQUIRE(host_blob && base, "fb_tmap_encode_mat2d: null pointer");
  FB_REQUIRE(box_k == 64 && box_rows >= 1 && box_rows <= 256, "fb_tmap_encode_mat2d: bad box %d x %d", box_k, box_rows);
  FB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0,
             "fb_tmap_encode_mat2d: base and row stride must be 16-byte aligned");
  cuuint64_t dims[2] = {cuuint64_t(k), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(ld * 2)};
  cuuint32_t box[2] = {cuuint32_t(box_k), cuuint32_t(box_rows)};
  return encode(host_blob, base, 2, dims, strides, box);
}

// microbatch-group geometry shared by fb_conv_gemm / fb_conv_wgrad: returns the 128-pixel tiles per group, or -1
static int tiles_per_group(int tile_h, int tile_n, int grid_h, int grid_n, int mg_imgs, int ng) {
  if (ng <= 1) return (tile_n == 1) ? grid_n * (grid_h / tile_h) : (grid_n + tile_n - 1) / tile_n;
  if (mg_imgs <= 0 || grid_n != ng * mg_imgs) return -1;
  if (tile_n == 1) return mg_imgs * (grid_h / tile_h);
  return (mg_imgs % tile_n == 0) ? mg_imgs / tile_n : -1;
}

extern "C" int fb_conv_gemm(const fb_conv_gemm_args* a, void* stream) {
  FB_REQUIRE(a && a->host_a_maps && a->host_b_maps && a->out, "fb_conv_gemm: null pointer");
  FB_REQUIRE(a->a_planes >= 1 && a->a_planes <= 2 && a->b_planes >= 1 && a->b_planes <= 2 && a->n_phases >= 1 &&
                 a->n_phases * a->a_planes <= FB_MAX_A_MAPS,
             "fb_conv_gemm: plane / phase counts out of range (%d, %d, %d)", a->a_planes, a->b_planes, a->n_phases);
  FB_REQUIRE(a->n_taps >= 1 && a->n_taps <= FB_MAX_TAPS && a->cblocks >= 1, "fb_conv_gemm: bad tap count %d", a->n_taps);
  FB_REQUIRE(a->tile_w * a->tile_h * a->tile_n == 128, "fb_conv_gemm: tile %dx%dx%d is not 128 pixels", a->tile_w,
             a->tile_h, a->tile_n);
  FB_REQUIRE(a->tile_n == 1 ? (a->grid_h % a->tile_h == 0) : (a->tile_h == a->grid_h),
             "fb_conv_gemm: tile does not divide the pixel grid (grid_h %d, tile_h %d, tile_n %d)", a->grid_h, a->tile_h,
             a->tile_n);
  for (int i = 0; i < a->n_taps; ++i)
    FB_REQUIRE(a->taps[i].phase >= 0 && a->taps[i].phase < a->n_phases, "fb_conv_gemm: tap %d references a missing map",
               i);
  if (!(a->n_tile == 64 || a->n_tile == 128 || a->n_tile == 256) || a->n_total % a->n_tile != 0) {
    set_error("fb_conv_gemm: unsupported n_tile %d for n_total %d", a->n_tile, a->n_total);
    return FB_ERR_UNSUPPORTED;
  }
  FB_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && a->out_sn % 4 == 0 && a->out_sh % 4 == 0 &&
                 a->out_sw % 4 == 0,
             "fb_conv_gemm: output must be 16-byte aligned with strides multiple of 4");
  const int ng = a->ng > 0 ? a->ng : 1;
  FB_REQUIRE(ng <= FB_MAX_GROUPS, "fb_conv_gemm: at most %d groups", FB_MAX_GROUPS);
  const int mtg = tiles_per_group(a->tile_h, a->tile_n, a->grid_h, a->grid_n, a->mg_imgs, ng);
  if (mtg <= 0) {
    set_error("fb_conv_gemm: %d groups of %d images do not tile the grid of %d images (tile_n %d)", ng, a->mg_imgs,
              a->grid_n, a->tile_n);
    return FB_ERR_UNSUPPORTED;
  }
  ConvGemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  memcpy(kp.a_maps, a->host_a_maps, size_t(a->n_phases * a->a_planes) * FB_TMAP_BYTES);
  memcpy(kp.b_maps, a->host_b_maps, size_t(a->b_planes) * FB_TMAP_BYTES);
  memcpy(kp.taps, a->taps, sizeof(fb_tap) * a->n_taps);
  kp.n_taps = a->n_taps;
  kp.cblocks = a->cblocks;
  kp.tile_w = a->tile_w;
  kp.tile_h = a->tile_h;
  kp.tile_n = a->tile_n;
  kp.grid_h = a->grid_h;
  kp.grid_n = a->grid_n;
  kp.mtg = mtg;
  kp.n_tiles = a->n_total / a->n_tile;
  const int policy_groups = a->policy_groups > 0 ? a->policy_groups : 1;
  kp.rows = stats_rows(mtg, kp.n_tiles, policy_groups, a->sched_k_iters);
  kp.ng = ng;
  kp.b_group_rows = a->b_group_rows;
  kp.reverse = a->reverse ? 1 : 0;
  kp.cta_pair = a->cta_pair == 2 ? 2 : (a->cta_pair ? 1 : 0);
  FB_REQUIRE(!kp.cta_pair || pair_ok(mtg, kp.n_tiles, policy_groups, a->sched_k_iters),
             "fb_conv_gemm: this problem cannot run as CTA pairs");
  kp.halo = a->halo ? 1 : 0;
  {
    static const int dbg = [] {
      const char* e = getenv("FB_CONV_DBG");
      return e ? atoi(e) : 0;
    }();
    kp.dbg = dbg;
  }
  if (kp.halo) {
    FB_REQUIRE(!kp.cta_pair && a->tile_n == 1 && a->n_taps % 3 == 0 && a->n_tapgroups <= 1,
               "fb_conv_gemm: haloed A boxes need single-image tiles, taps in triples and independent CTAs");
    for (int t = 0; t < a->n_taps; t += 3)
      for (int j = 0; j < 3; ++j)
        FB_REQUIRE(a->taps[t + j].dh == j - 1 && a->taps[t + j].dw == a->taps[t].dw &&
                       a->taps[t + j].phase == a->taps[t].phase,
                   "fb_conv_gemm: tap triple %d is not (dh = -1, 0, 1) at one dw", t / 3);
  }
  kp.out = a->out;
  kp.out_sn = a->out_sn;
  kp.out_sh = a->out_sh;
  kp.out_sw = a->out_sw;
  kp.accumulate = a->accumulate;
  kp.n_total = a->n_total;
  FB_REQUIRE(a->n_tapgroups >= 0 && a->n_tapgroups <= 4, "fb_conv_gemm: n_tapgroups must be in 0..4");
  FB_REQUIRE(a->n_tapgroups <= 1 || !a->stats_ws, "fb_conv_gemm: tap groups cannot be combined with the statistics");
  kp.n_tapgroups = a->n_tapgroups > 0 ? a->n_tapgroups : 1;
  for (int g = 0; g < 4; ++g) {
    kp.group_tap0[g] = 0;
    kp.group_taps[g] = a->n_taps;
    kp.group_off[g] = 0;
  }
  if (a->n_tapgroups > 0)
    for (int g = 0; g < a->n_tapgroups; ++g) {
      FB_REQUIRE(a->tapgroups[g].tap0 >= 0 && a->tapgroups[g].n_taps >= 1 &&
                     a->tapgroups[g].tap0 + a->tapgroups[g].n_taps <= a->n_taps,
                 "fb_conv_gemm: tap group %d out of range", g);
      kp.group_tap0[g] = a->tapgroups[g].tap0;
      kp.group_taps[g] = a->tapgroups[g].n_taps;
      kp.group_off[g] = a->tapgroups[g].out_off;
    }
  if (a->stats_ws) {
    FB_REQUIRE(!a->accumulate, "fb_conv_gemm: the statistics need accumulate == 0");
    FB_REQUIRE(a->tickets && a->bn_mean && a->bn_rstd, "fb_conv_gemm: statistics need tickets, bn_mean and bn_rstd");
    kp.stats = a->stats_ws;
    kp.tickets = a->tickets;
    kp.bn_mean = a->bn_mean;
    kp.bn_rstd = a->bn_rstd;
    kp.bn_batch = a->bn_batch;
    kp.bn_eps = a->bn_eps;
    // pixels per group that really exist (a ragged single group: grid_n images)
    const long long imgs = (ng > 1) ? a->mg_imgs : a->grid_n;
    kp.bn_count = double(imgs) * double(a->grid_h) * double(a->tile_w);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->n_tile) {
    case 64: return dispatch_conv_gemm<64>(kp, a->a_planes, a->b_planes, st);
    case 128: return dispatch_conv_gemm<128>(kp, a->a_planes, a->b_planes, st);
    default: return dispatch_conv_gemm<256>(kp, a->a_planes, a->b_planes, st);
  }
}

extern "C" int fb_conv_stats_rows(int m_tiles_per_group, int n_tiles, int policy_groups, int sched_k_iters) {
  return stats_rows(m_tiles_per_group, n_tiles, policy_groups, sched_k_iters);
}
extern "C" int fb_conv_pair_ok(int m_tiles_per_group, int n_tiles, int policy_groups, int sched_k_iters) {
  return pair_ok(m_tiles_per_group, n_tiles, policy_groups, sched_k_iters) ? 1 : 0;
}

extern "C" int fb_conv_wgrad(const fb_wgrad_args* a, void* stream) {
  FB_REQUIRE(a && a->host_dy_map && a->host_x_maps && a->out, "fb_conv_wgrad: null pointer");
  FB_REQUIRE(a->planes >= 1 && a->planes <= 2 && a->n_x_maps >= a->planes && a->n_x_maps <= FB_MAX_A_MAPS,
             "fb_conv_wgrad: bad planes/maps (%d, %d)", a->planes, a->n_x_maps);
  FB_REQUIRE(a->n_taps >= 1 && a->n_taps <= FB_MAX_WGRAD_TAPS && a->cblocks >= 1 && a->cin == 64 * a->cblocks,
             "fb_conv_wgrad: bad taps/cblocks/cin (%d, %d, %d)", a->n_taps, a->cblocks, a->cin);
  FB_REQUIRE(a->slots_per_cta >= 1 && a->slots_per_cta * a->planes <= 8,
             "fb_conv_wgrad: slots_per_cta * planes must be in 1..8 (512 TMEM columns)");
  FB_REQUIRE(a->tile_w * a->tile_h * a->tile_n == 128, "fb_conv_wgrad: tile is not 128 pixels");
  FB_REQUIRE(a->tile_n == 1 ? (a->grid_h % a->tile_h == 0) : (a->tile_h == a->grid_h),
             "fb_conv_wgrad: tile does not divide the pixel grid");
  for (int i = 0; i < a->n_taps; ++i)
    FB_REQUIRE((a->taps[i].phase + 1) * a->planes <= a->n_x_maps, "fb_conv_wgrad: tap %d references a missing map", i);
  const int ng = a->ng > 0 ? a->ng : 1;
  FB_REQUIRE(ng <= FB_MAX_GROUPS, "fb_conv_wgrad: at most %d groups", FB_MAX_GROUPS);
  const int pbg = tiles_per_group(a->tile_h, a->tile_n, a->grid_h, a->grid_n, a->mg_imgs, ng);
  if (pbg <= 0) {
    set_error("fb_conv_wgrad: %d groups of %d images do not tile the grid of %d images (tile_n %d)", ng, a->mg_imgs,
              a->grid_n, a->tile_n);
    return FB_ERR_UNSUPPORTED;
  }
  FB_REQUIRE(a->splits >= 1 && a->splits <= pbg, "fb_conv_wgrad: splits %d must be in 1..%d", a->splits, pbg);
  FB_REQUIRE(ng * a->splits <= 65535, "fb_conv_wgrad: too many (group, split) pairs");
  FB_REQUIRE(a->cout % 64 == 0, "fb_conv_wgrad: cout must be a multiple of 64");
  FB_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && a->out_gstride % 4 == 0 && a->out_sstride % 4 == 0,
             "fb_conv_wgrad: output base and strides must be 16-byte aligned");
  WgradKParams kp;
  memset(&kp, 0, sizeof(kp));
  memcpy(&kp.dy_map, a->host_dy_map, FB_TMAP_BYTES);
  memcpy(kp.x_maps, a->host_x_maps, size_t(a->n_x_maps) * FB_TMAP_BYTES);
  memcpy(kp.taps, a->taps, sizeof(fb_wgrad_tap) * a->n_taps);
  kp.n_taps = a->n_taps;
  kp.cblocks = a->cblocks;
  kp.planes = a->planes;
  kp.slots_per_cta = a->slots_per_cta;
  kp.cout = a->cout;
  kp.cin = a->cin;
  kp.tile_w = a->tile_w;
  kp.tile_h = a->tile_h;
  kp.tile_n = a->tile_n;
  kp.grid_h = a->grid_h;
  kp.grid_n = a->grid_n;
  kp.splits = a->splits;
  kp.pbg = pbg;
  kp.out = a->out;
  kp.out_gstride = a->out_gstride;
  kp.out_sstride = a->out_sstride;
  kp.halo = a->halo ? 1 : 0;
  if (a->halo) {
    FB_REQUIRE(a->n_taps == 9 && a->slots_per_cta == 3 && a->tile_n == 1 && (a->tile_w * 128) % 1024 == 0,
               "fb_conv_wgrad: halo mode needs 9 taps, 3 slots per CTA and whole-row tiles of >= 1024 bytes");
    for (int t = 0; t < 9; t += 3)
      FB_REQUIRE(a->taps[t].dw == a->taps[t + 1].dw && a->taps[t].dw == a->taps[t + 2].dw && a->taps[t].dh == -1 &&
                     a->taps[t + 1].dh == 0 && a->taps[t + 2].dh == 1 && a->taps[t].phase == 0,
                 "fb_conv_wgrad: halo mode needs taps in triples (dh = -1, 0, 1) that share dw");
    FB_REQUIRE(a->planes * (a->tile_h + 2) * a->tile_w * 128 * 2 <= kWgBStages * kWgBBytes,
               "fb_conv_wgrad: two haloed stages do not fit");
  }
  for (int t = 0; t < a->n_taps; ++t)
    FB_REQUIRE(a->taps[t].k_index >= 0 && a->taps[t].k_index < a->n_taps, "fb_conv_wgrad: tap %d has a bad k_index", t);
  const int n_slots_total = a->n_taps * a->cblocks;
  dim3 grid((a->cout + 127) / 128, (n_slots_total + a->slots_per_cta - 1) / a->slots_per_cta, ng * a->splits);
  constexpr int smem = kWgAStages * kWgABytes + kWgBStages * kWgBBytes + kEpiStageBytes + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    FB_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  FB_CUDA(launch_pdl(wgrad_kernel, grid, dim3(kThreads), smem, static_cast<cudaStream_t>(stream), kp));
  return 0;
}

extern "C" int fb_reduce_multi(const fb_reduce_entry* table_dev, int n_entries, int total_blocks, float* dst_base,
                               int64_t dst_gstride, int ng, void* stream) {
  FB_REQUIRE(table_dev && dst_base && n_entries > 0 && total_blocks > 0 && ng >= 1 && ng <= FB_MAX_GROUPS,
             "fb_reduce_multi: bad arguments");
  FB_CUDA(launch_pdl(reduce_multi_kernel, dim3(total_blocks, ng), dim3(256), 0, static_cast<cudaStream_t>(stream),
                     table_dev, n_entries, dst_base, (long long)dst_gstride));
  return 0;
}

extern "C" int fb_weight_prep(const float* w_oihw, int cout, int cin, int taps, void* wf_hi, void* wf_lo, int64_t ld_f,
                              void* wd_hi, void* wd_lo, int64_t ld_d, void* stream) {
  FB_REQUIRE(w_oihw && wf_hi, "fb_weight_prep: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cin % 32 != 0) {
    FB_REQUIRE(wd_hi == nullptr, "fb_weight_prep: dgrad layout needs cin %% 32 == 0");
    const int total = cout * cin * taps;
    weight_prep_direct_kernel<<<(total + 255) / 256, 256, 0, st>>>(w_oihw, cout, cin * taps,
                                                                   static_cast<__nv_bfloat16*>(wf_hi),
                                                                   static_cast<__nv_bfloat16*>(wf_lo), ld_f);
  } else {
    FB_REQUIRE(cout % 32 == 0 && (taps == 1 || taps == 9), "fb_weight_prep: unsupported shape %d %d %d", cout, cin, taps);
    dim3 grid(cin / 32, cout / 32);
    const size_t smem = size_t(32) * (32 * taps + 1) * sizeof(float);
    weight_prep_kernel<<<grid, 256, smem, st>>>(w_oihw, cout, cin, taps, static_cast<__nv_bfloat16*>(wf_hi),
                                                static_cast<__nv_bfloat16*>(wf_lo), ld_f,
                                                static_cast<__nv_bfloat16*>(wd_hi), static_cast<__nv_bfloat16*>(wd_lo),
                                                ld_d);
  }
  FB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fb_weight_prep_multi(const float* theta, const fb_wprep_entry* table_dev, int n_entries, int total_blocks,
                                    const float* grad, int64_t grad_gstride, const float* pre, float bs, float acc,
                                    float scale, const float* scal, int eps_base, int ng, void* stream) {
  FB_REQUIRE(theta && table_dev && n_entries > 0 && total_blocks > 0, "fb_weight_prep_multi: bad arguments");
  FB_REQUIRE(ng >= 1 && ng <= FB_MAX_GROUPS, "fb_weight_prep_multi: ng out of range");
  FB_REQUIRE(grad || ng == 1, "fb_weight_prep_multi: several groups need a gradient (the plain weights are shared)");
  FB_REQUIRE(!grad || scal, "fb_weight_prep_multi: the perturbed point needs scal (eps_n per group)");
  const size_t smem = size_t(32) * (32 * 9 + 1) * sizeof(float);
  FB_CUDA(launch_pdl(weight_prep_multi_kernel, dim3(total_blocks, ng), dim3(256), smem,
                     static_cast<cudaStream_t>(stream), theta, table_dev, n_entries, grad, (long long)grad_gstride, pre,
                     bs, acc, scale, scal, eps_base));
  return 0;
}
