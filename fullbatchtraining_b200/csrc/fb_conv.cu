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
//   conv3x3_kernel   : opt-in haloed-box variant for 3x3 / stride 1 (row shifts are aligned views of one box).
//   wgrad_kernel     : partial[co, (tap,ci)] = sum_pixels dY[pixel, co] * X[pixel + tap, ci]; both operands are
//                      MN-major (the channel dimension is contiguous in NHWC), split-K over 128-pixel blocks,
//                      up to 8 (tap, ci-block) accumulators of 64 TMEM columns per CTA; haloed X boxes on 32x32 / 16x16.
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
  int m_tiles, n_tiles;
  // tap groups: tile t belongs to group t / (m_tiles * n_tiles); group g accumulates taps [tap0, tap0 + n_taps) and
  // writes at out + out_off (the four output phases of a stride-2 dgrad in one launch); n_groups == 1: all taps
  int n_groups;
  int group_tap0[4], group_taps[4];
  long long group_off[4];
  float* out;
  long long out_sn, out_sh, out_sw;
  int accumulate;
  float* stats;  // optional [gridDim.x / n_tiles][2][n_total]: per-CTA column sums / sums of squares of the output
  int n_total;
  // BatchNorm-backward statistics instead (fb_conv_gemm_args.bwd_y): sums of d*m and d*m*xhat of the stored gradient d
  const float* bwd_y;
  const __nv_bfloat16* bwd_mask;
  const float *bwd_mean, *bwd_rstd;
  // development switch (env FB_CONV_EXPERIMENT, results are WRONG when set): 1 = epilogue without global stores,
  // 2 = epilogue only hands the accumulator back, 4 = producer stops fetching after the first ring fill
  int experiment;
};

constexpr int kTileM = 128;                        // pixels per CTA tile == UMMA M
constexpr int kBlockK = 64;                        // bf16 elements per K block == one 128-byte swizzle row
constexpr int kATileBytes = kTileM * kBlockK * 2;  // 16 KiB
// Warp roles of the tensor-core kernels (352 threads): 0 = TMA producer, 1 (and 6 if kMmaIssuers == 2) = MMA issuer,
// 2-5 and 7-10 = epilogue.
// A warp may only read the TMEM lanes 32*(warp%4)..+31, so each lane quarter has TWO epilogue warps (groups 0 and 1)
// that take alternate 16-column chunks: the accumulator drain, which is fully exposed for the last tile of a CTA, is
// twice as fast.
// kMmaIssuers = 2 lets warps 1 and 6 issue alternate pipeline stages (a turn token orders the ISSUE; the probe reaches
// the pipe's 64 clk / MMA that way), but the tensor pipe does not retire MMAs of different warps in a fixed order: the
// fp32 accumulation order then varies from run to run (seen as non-bit-identical gradients), so the default is ONE issuer.
constexpr int kMmaIssuers = 1;
constexpr int kThreads = 352;
constexpr int kEpiWarps = 8;
constexpr int kEpiStageBytes = kEpiWarps * kEpiWarpFloats * 4;  // 8 epilogue warps x 32 x 20 floats
__device__ __forceinline__ bool is_epilogue_warp(int warp) { return (warp >= 2 && warp <= 5) || warp >= 7; }
__device__ __forceinline__ int epilogue_group(int warp) { return warp >= 7 ? 1 : 0; }
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

// Flush of the per-lane column statistics collected by the epilogue (BatchNorm statistics fused into the producing
// convolution).  acc[c][0..3] / acc[c][4..7] of lane l are the sums / sums of squares of columns
// c*16 + 4*(l%4) .. +3 over the rows the lane stored (rows = lane/4 mod 8).  Fixed order: shuffle tree over the 8 row
// groups, then the eight epilogue warps through the staging patch; one partial row stats[row][0 = sum | 1 = sq][channel].
template <int N_TILE>
__device__ __forceinline__ void flush_column_stats(float* epi_stage, float (&acc)[N_TILE / 32][8], int q, int eg, int lane,
                                                   float* stats, int row, int n_total, int n_tile0) {
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
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps only
  const int t = (eg * 4 + q) * 32 + lane;
  for (int idx = t; idx < 2 * N_TILE; idx += 256) {
    const int which = idx / N_TILE, col = idx % N_TILE;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kEpiWarps; ++w) v += epi_stage[w * kEpiWarpFloats + idx];
    stats[((long long)row * 2 + which) * n_total + n_tile0 + col] = v;
  }
}

template <int N_TILE, int PA, int PB>
struct ConvGemmCfg {
  static constexpr int kBBytes = N_TILE * kBlockK * 2;
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
  static constexpr bool kStack = (PB == 2 && N_TILE <= 128);
  static constexpr int kUmmaN = kStack ? 2 * N_TILE : N_TILE;
  static constexpr int kTmemCols = 2 * kUmmaN;
  // instructions per K = 16 step: stacked -> one per A plane; otherwise (a0,b0), (a0,b1) if PB == 2, (a1,b0) if PA == 2
  static constexpr int kCombos = kStack ? PA : ((PA == 2 && PB == 2) ? 3 : (PA * PB));
  static_assert(kStages >= 2, "stage does not fit twice into shared memory");
};

template <int N_TILE, int PA, int PB>
__global__ void __launch_bounds__(kThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmKParams p) {
  using Cfg = ConvGemmCfg<N_TILE, PA, PB>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes + kEpiStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;  // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint64_t* turn_bar = acc_empty + 2;       // [2] issue-order token of the two MMA warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(turn_bar + 2);

  // warp index through a shuffle: the compiler then knows it is warp-uniform, and the producer / MMA roles below run
  // warp-converged with ONE elected lane issuing, so that their operands live in uniform registers and the unrolled
  // tcgen05.mma block compiles to back-to-back UTCHMMA (a `lane == 0` branch costs ~13 instructions per MMA and made
  // the single issuing thread, not the tensor pipe, the bottleneck)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 32) {  // descriptor fetch overlaps the barrier / TMEM set-up
#pragma unroll
    for (int pl = 0; pl < PA; ++pl) tma_prefetch_desc(&p.a_maps[p.taps[0].phase * PA + pl]);
#pragma unroll
    for (int pl = 0; pl < PB; ++pl) tma_prefetch_desc(&p.b_maps[pl]);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpiWarps);  // one arrival per epilogue warp
      mbar_init(&turn_bar[b], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  griddep_wait();    // everything above overlapped the predecessor's tail
  griddep_launch();

  const int group_tiles = p.m_tiles * p.n_tiles;
  const int total_tiles = p.n_groups * group_tiles;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    int s = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int grp = tile / group_tiles, gt = tile % group_tiles;
      int n0, h0;
      tile_origin(gt / p.n_tiles, p.tile_h, p.tile_n, p.grid_h, n0, h0);
      const int n_tile0 = (gt % p.n_tiles) * N_TILE;
      for (int t = p.group_tap0[grp]; t < p.group_tap0[grp] + p.group_taps[grp]; ++t) {
        const fb_tap tap = p.taps[t];
        for (int cb = 0; cb < p.cblocks; ++cb, ++it) {
          mbar_wait(&empty_bar[s], phase ^ 1, 1);
          if (elect_one()) {
            uint8_t* st = smem + s * Cfg::kStageBytes;
            if ((p.experiment & 4) && it >= STAGES) {
              mbar_arrive(&full_bar[s]);
            } else {
              mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
#pragma unroll
              for (int pl = 0; pl < PA; ++pl)
                tma_load_4d(st + pl * kATileBytes, &p.a_maps[tap.phase * PA + pl], &full_bar[s], cb * kBlockK, tap.dw,
                            h0 + tap.dh, n0);
#pragma unroll
              for (int pl = 0; pl < PB; ++pl)
                tma_load_2d(st + PA * kATileBytes + pl * Cfg::kBBytes, &p.b_maps[pl], &full_bar[s],
                            tap.b_k0 + cb * kBlockK, n_tile0);
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
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    // A tcgen05.mma blocks its issuing thread until the tensor pipe has taken it, and nothing that thread executes
    // between two MMAs overlaps with them (tools/probes/mma_issue.cu: every instruction between two MMAs adds its full
    // latency; one issuing warp reaches 79-98 clk per M128xN128xK16 MMA with 8-4 MMAs per stage, two alternating warps
    // 64.0 = the pipe's rate).  The role therefore runs warp-converged with an elected lane, so that the descriptors
    // live in uniform registers and the unrolled block is back-to-back UTCHMMA.  With kMmaIssuers == 2 warps 1 and 6
    // alternate stages behind a turn token; that is NOT the default because the accumulation order is then no longer
    // reproducible (see kMmaIssuers).
    constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
    constexpr int kC = Cfg::kCombos;
    //   stacked, PA == 2: c0 = A_lo, c1 = A_hi (both against [B_hi;B_lo]); for 128-wide tiles A_lo only needs B_hi, so
    //     it runs at N = N_TILE except in the first K block of a tile, where it must initialise both halves
    //   not stacked: (a1,b0) first if PA == 2, then (a0,b0), then (a0,b1) if PB == 2
    constexpr bool kNarrowLo = Cfg::kStack && N_TILE == 128 && PA == 2;
    auto ap_of = [](int c) { return (PA == 2 && c == 0) ? 1 : 0; };
    auto bp_of = [](int c) { return (!Cfg::kStack && PB == 2 && c == kC - 1) ? 1 : 0; };
    constexpr uint32_t idesc_full = make_idesc_bf16(kTileM, Cfg::kUmmaN, 0, 0);
    constexpr uint32_t idesc_half = make_idesc_bf16(kTileM, N_TILE, 0, 0);
    const uint32_t smem0 = smem_u32(smem);
    // Single issuer (kMmaIssuers == 1, see above): everything the issuing warp executes between two MMAs is dead time
    // for the tensor pipe (~270 clocks per barrier wait + elect + descriptor set-up + commit round), so TWO pipeline
    // stages are waited for and issued per round whenever the tile has another stage left.
    static_assert(kMmaIssuers == 1, "conv_gemm_kernel issues from one warp; the two-issuer variant lives in the probes");
    (void)turn_bar;
    auto issue_stage = [&](uint32_t tmem_d, int st, int ki) {
      const uint32_t a_lo = smem_desc_lo(smem0 + st * Cfg::kStageBytes, 16);
      const uint32_t b_lo = a_lo + ((PA * kATileBytes) >> 4);
#pragma unroll
      for (int c = 0; c < kC; ++c) {
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)
          tc_mma_bf16_lohi(tmem_d, a_lo + ((ap_of(c) * kATileBytes + k * 32) >> 4),
                           b_lo + ((bp_of(c) * Cfg::kBBytes + k * 32) >> 4), desc_hi, desc_hi,
                           (kNarrowLo && c == 0 && ki != 0) ? idesc_half : idesc_full, (ki | c | k) != 0);
      }
      tc_commit(&empty_bar[st]);
    };
    int s = 0;
    uint32_t phase = 0;
    int tile_i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_i) {
      const int k_iters = p.group_taps[tile / group_tiles] * p.cblocks;
      const int buf = tile_i & 1;
      const uint32_t tmem_d = tmem_base + buf * Cfg::kUmmaN;
      mbar_wait(&acc_empty[buf], ((tile_i >> 1) & 1) ^ 1, 4);  // epilogue has drained this accumulator
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
          if (ki + 2 >= k_iters) tc_commit(&acc_full[buf]);
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
  } else if (is_epilogue_warp(warp)) {
    // ---------------- epilogue: TMEM -> registers -> smem transpose -> coalesced global stores (fp32 NHWC) ----------
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int eg = epilogue_group(warp);  // 16-column chunks c with c % 2 == eg
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
    int tile_i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_i) {
      const int grp = tile / group_tiles, gt = tile % group_tiles;
      int n0, h0;
      tile_origin(gt / p.n_tiles, p.tile_h, p.tile_n, p.grid_h, n0, h0);
      const int n_tile0 = (gt % p.n_tiles) * N_TILE;
      const bool valid = (n0 + n) < p.grid_n && (h0 + h) < p.grid_h;
      const long long row_off = p.group_off[grp] + (long long)(n0 + n) * p.out_sn + (long long)(h0 + h) * p.out_sh +
                                (long long)w * p.out_sw + n_tile0;
      const int buf = tile_i & 1;
      mbar_wait(&acc_full[buf], (tile_i >> 1) & 1, 3);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < N_TILE / 32; ++cc) {  // unrolled: col_acc must stay in registers
        if (p.experiment & 2) break;
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
        EpiBwd bwd;
        int stat_mode = p.stats ? kEpiStatFwd : kEpiStatNone;
        if (p.bwd_y) {
          stat_mode = kEpiStatBwd;
          bwd.y = p.bwd_y;
          bwd.mask = p.bwd_mask;
          const int col = n_tile0 + c * 16 + (lane & 3) * 4;
          const float4 mu = *reinterpret_cast<const float4*>(p.bwd_mean + col);
          const float4 rs = *reinterpret_cast<const float4*>(p.bwd_rstd + col);
          bwd.mu[0] = mu.x; bwd.mu[1] = mu.y; bwd.mu[2] = mu.z; bwd.mu[3] = mu.w;
          bwd.rs[0] = rs.x; bwd.rs[1] = rs.y; bwd.rs[2] = rs.z; bwd.rs[3] = rs.w;
        }
        warp_store_rows16(stage, v, p.out, row_off, valid && !(p.experiment & 1), c * 16, p.accumulate != 0, lane,
                          col_acc[cc], stat_mode, bwd);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    // host guarantees gridDim.x % n_tiles == 0 when stats are requested: this CTA only saw one N tile
    if (p.stats)
      flush_column_stats<N_TILE>(epi_stage, col_acc, q, eg, lane, p.stats, blockIdx.x / p.n_tiles, p.n_total,
                                 (blockIdx.x % p.n_tiles) * N_TILE);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int N_TILE, int PA, int PB>
static int launch_conv_gemm(const ConvGemmKParams& kp, cudaStream_t stream) {
  using Cfg = ConvGemmCfg<N_TILE, PA, PB>;
  static bool configured = false;
  if (!configured) {
    FB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<N_TILE, PA, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::kSmemBytes));
    configured = true;
  }
  const int tiles = kp.n_groups * kp.m_tiles * kp.n_tiles;
  int grid = tiles < kNumSMs ? tiles : kNumSMs;
  if (kp.stats) grid = (grid / kp.n_tiles) * kp.n_tiles;  // one N tile per CTA (see flush_column_stats)
  FB_CUDA(launch_pdl(conv_gemm_kernel<N_TILE, PA, PB>, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, stream, kp));
  return 0;
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
// conv3x3_kernel: 3x3 / stride 1 / pad 1 convolutions (forward and dgrad) on feature maps whose rows are >= 1024 bytes
// of one 64-channel block (W >= 8 pixels) and whose 256-pixel tiles are whole image rows.
//
// The generic kernel re-fetches the A box once per filter tap (9x).  Here the producer fetches, per column shift
// dw in {-1,0,1} and 64-channel block, ONE haloed box of (2*TH + 2) image rows (TH = 128 / W rows per 128-pixel half);
// the three row shifts dh are then 1024-byte-aligned VIEWS of that box (offset (dh+1) * W * 128 bytes), so the
// SWIZZLE_128B pattern is preserved and no data is moved.  A CTA tile is 256 pixels = two UMMA M=128 halves that share
// every weight tile.  Traffic per MMA-clock drops from ~125 to ~57 bytes (N_TILE 64) / ~37 bytes (N_TILE 128).
// ---------------------------------------------------------------------------------------------------------------
// Optional in-kernel cycle accounting of CTA 0 (FB_KERNEL_DEBUG=1): where do the producer / MMA / epilogue roles wait?
__device__ long long g_dbg[32];

struct alignas(64) Conv3x3KParams {
  // [plane]: imgs == 1: dims (C, W, H, N), box = 64 ch x W x (halves*TH+2) rows x 1 image
  //          imgs  > 1: dims (C, W, N, H), box = 64 ch x W x imgs images x (H+2) rows (image-interleaved slabs)
  CUtensorMap a_maps[2];
  CUtensorMap b_maps[2];  // [plane]: box = 64 x N_TILE weight rows
  int b_k0[3][3];         // [dw+1][dh+1] -> first K column of that tap in the weight matrix
  int cblocks;
  int w, h, n;            // feature map and images
  int th;                 // slabs per 128-pixel half; a slab = one image row of `imgs` consecutive images
  int imgs;               // images interleaved in a slab (1: a half is TH rows of one image; >1: a half is imgs whole images)
  int halves;             // 128-pixel halves per CTA tile (they share every weight tile)
  int n_tiles;
  float* out;
  long long out_sn, out_sh, out_sw;
  int accumulate;
  int debug;
  float* stats;  // optional per-CTA column statistics, see ConvGemmKParams
  int n_total;
};

#define FB_DBG_WAIT(slot, call)                      \
  do {                                               \
    if (dbg) {                                       \
      const long long _t = clock64();                \
      call;                                          \
      dbg_acc[slot & 1] += clock64() - _t;           \
    } else {                                         \
      call;                                          \
    }                                                \
  } while (0)

template <int N_TILE, int PA, int PB>
struct Conv3x3Cfg {
  static constexpr int kBBytes = N_TILE * kBlockK * 2;
  static constexpr int kBStageBytes = PB * kBBytes;
  static constexpr int kAStages = 2;
  static constexpr bool kStack = (PB == 2 && N_TILE == 64);  // see ConvGemmCfg
  static constexpr int kUmmaN = kStack ? 128 : N_TILE;
  static constexpr int kCombos = kStack ? PA : ((PA == 2 && PB == 2) ? 3 : (PA * PB));
  static constexpr int kTmemCols = 4 * kUmmaN;  // 2 halves x 2 buffers
};

template <int N_TILE, int PA, int PB>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_kernel(const __grid_constant__ Conv3x3KParams p, int a_box_bytes,
                                                         int b_stages) {
  using Cfg = Conv3x3Cfg<N_TILE, PA, PB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kAStages * PA * ((p.imgs == 1) ? 1 : p.halves) * a_box_bytes;
  float* epi_stage = reinterpret_cast<float*>(smem_b + b_stages * Cfg::kBStageBytes);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + b_stages * Cfg::kBStageBytes + kEpiStageBytes);
  uint64_t* a_empty = a_full + Cfg::kAStages;
  uint64_t* b_full = a_empty + Cfg::kAStages;
  uint64_t* b_empty = b_full + 8;
  uint64_t* acc_full = b_empty + 8;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* turn_bar = acc_empty + 2;  // [2] issue-order token of the two MMA warps (see conv_gemm_kernel)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(turn_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform roles, see conv_gemm_kernel
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kAStages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpiWarps);
      mbar_init(&turn_bar[b], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  griddep_wait();
  griddep_launch();

  // imgs == 1: a tile is halves*TH consecutive rows of one image, fetched as ONE haloed box;
  // imgs  > 1: a half is `imgs` whole images with rows interleaved ([h][img][w]); one haloed box per half
  const int slab_px = p.imgs * p.w;
  const int tiles_per_img = (p.imgs == 1) ? p.h / (p.halves * p.th) : 1;
  const int m_tiles = (p.imgs == 1) ? p.n * tiles_per_img : p.n / (p.halves * p.imgs);
  const int total_tiles = m_tiles * p.n_tiles;
  const int boxes = (p.imgs == 1) ? 1 : p.halves;
  const int row_bytes = slab_px * 128;      // one slab of one 64-channel block
  const int half_bytes = p.th * row_bytes;  // 128 pixels = 16 KiB
  const int half_stride = (p.imgs == 1) ? half_bytes : a_box_bytes;
  const int a_stage_bytes = PA * boxes * a_box_bytes;
  const bool dbg = p.debug && blockIdx.x == 0;
  const long long t_start = clock64();
  long long dbg_acc[2] = {0, 0};  // register accumulators (slot parity), flushed once per role
  long long dbg_issue = 0, dbg_acc_empty = 0;

  if (warp == 0) {
    // ---------------- TMA producer (warp-converged, one elected lane issues) ----------------
    int as = 0, bs = 0;
    uint32_t aphase = 0, bphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles;
      const int n0 = (p.imgs == 1) ? mt / tiles_per_img : mt * p.halves * p.imgs;
      const int h0 = (p.imgs == 1) ? (mt % tiles_per_img) * p.halves * p.th : 0;
      const int n_tile0 = (tile % p.n_tiles) * N_TILE;
      for (int dwi = 0; dwi < 3; ++dwi) {
        for (int cb = 0; cb < p.cblocks; ++cb) {
          FB_DBG_WAIT(0, mbar_wait(&a_empty[as], aphase ^ 1, 21));
          if (elect_one()) {
            mbar_arrive_expect_tx(&a_full[as], a_stage_bytes);
#pragma unroll
            for (int pl = 0; pl < PA; ++pl) {
              uint8_t* dst = smem_a + as * a_stage_bytes + pl * boxes * a_box_bytes;
              if (p.imgs == 1) {
                tma_load_4d(dst, &p.a_maps[pl], &a_full[as], cb * kBlockK, dwi - 1, h0 - 1, n0);
              } else {
                for (int b = 0; b < boxes; ++b)
                  tma_load_4d(dst + b * a_box_bytes, &p.a_maps[pl], &a_full[as], cb * kBlockK, dwi - 1,
                              n0 + b * p.imgs, -1);
              }
            }
          }
          __syncwarp();
          if (++as == Cfg::kAStages) {
            as = 0;
            aphase ^= 1;
          }
          for (int dhi = 0; dhi < 3; ++dhi) {
            FB_DBG_WAIT(1, mbar_wait(&b_empty[bs], bphase ^ 1, 22));
            if (elect_one()) {
              mbar_arrive_expect_tx(&b_full[bs], Cfg::kBStageBytes);
#pragma unroll
              for (int pl = 0; pl < PB; ++pl)
                tma_load_2d(smem_b + bs * Cfg::kBStageBytes + pl * Cfg::kBBytes, &p.b_maps[pl], &b_full[bs],
                            p.b_k0[dwi][dhi] + cb * kBlockK, n_tile0);
            }
            __syncwarp();
            if (++bs == b_stages) {
              bs = 0;
              bphase ^= 1;
            }
          }
        }
      }
    }
    if (dbg && lane == 0) {
      g_dbg[0] += dbg_acc[0];
      g_dbg[1] += dbg_acc[1];
      g_dbg[10] += clock64() - t_start;
    }
  } else if (warp == 1 || warp == 6) {
    // ---------------- MMA issuer(s): one elected lane per warp, see conv_gemm_kernel and kMmaIssuers ----------------
    constexpr uint32_t idesc = make_idesc_bf16(kTileM, Cfg::kUmmaN, 0, 0);
    constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
    const uint32_t smem_a0 = smem_u32(smem_a), smem_b0 = smem_u32(smem_b);
    const int me = (warp == 1) ? 0 : 1;
    const int spt = 9 * p.cblocks;  // tap stages per tile
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_stages = (me < kMmaIssuers) ? my_tiles * spt : 0;
    uint32_t tphase = 0;
    for (int it = me; it < total_stages; it += kMmaIssuers) {
      // the index arithmetic below is off the critical path: the other warp's MMAs are running meanwhile
      const int tile_i = it / spt, r = it % spt;
      const int dhi = r % 3;
      const int g = it / 3;  // (tile, dw, channel block) group = one A stage
      const int as = g % Cfg::kAStages;
      const uint32_t aphase = (g / Cfg::kAStages) & 1;
      const int bs = it % b_stages;
      const uint32_t bphase = (it / b_stages) & 1;
      const int buf = tile_i & 1;
      if (r == 0) {
        const long long _t = clock64();
        mbar_wait(&acc_empty[buf], ((tile_i >> 1) & 1) ^ 1, 24);
        dbg_acc_empty += clock64() - _t;
      }
      FB_DBG_WAIT(2, mbar_wait(&a_full[as], aphase, 23));
      FB_DBG_WAIT(3, mbar_wait(&b_full[bs], bphase, 25));
      const uint32_t tmem_d = tmem_base + buf * 2 * Cfg::kUmmaN;
      const uint32_t a_base = smem_a0 + as * a_stage_bytes + dhi * row_bytes;
      const uint32_t b_lo = smem_desc_lo(smem_b0 + bs * Cfg::kBStageBytes, 16);
      if (kMmaIssuers == 2 && it > 0) {
        mbar_wait(&turn_bar[me], tphase, 27);
        tphase ^= 1;
      }
      tc_fence_after();
      const long long t_issue = dbg ? clock64() : 0;
      if (elect_one()) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half >= p.halves) break;
          const uint32_t a_lo = smem_desc_lo(a_base + half * half_stride, 16);
#pragma unroll
          for (int c = 0; c < Cfg::kCombos; ++c) {
            const int ap = Cfg::kStack ? c : ((PA == 2 && c == Cfg::kCombos - 1) ? 1 : 0);
            const int bp = Cfg::kStack ? 0 : ((PB == 2 && c == 1) ? 1 : 0);
            const uint32_t a_pl = a_lo + ((ap * boxes * a_box_bytes) >> 4);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              tc_mma_bf16_lohi(tmem_d + half * Cfg::kUmmaN, a_pl + ((k * 32) >> 4),
                               b_lo + ((bp * Cfg::kBBytes + k * 32) >> 4), desc_hi, desc_hi, idesc,
                               (r == 0 && c == 0 && k == 0) ? 0u : 1u);
          }
        }
        if (kMmaIssuers == 2) mbar_arrive(&turn_bar[me ^ 1]);
        tc_commit(&b_empty[bs]);
        if (dhi == 2) tc_commit(&a_empty[as]);        // in-order pipe: covers the other warp's taps of this box
        if (r == spt - 1) tc_commit(&acc_full[buf]);
      }
      __syncwarp();
      if (dbg) dbg_issue += clock64() - t_issue;
    }
    if (dbg && lane == 0 && me == 0) {
      g_dbg[2] += dbg_acc[0];
      g_dbg[3] += dbg_acc[1];
      g_dbg[4] += dbg_acc_empty;
      g_dbg[5] += clock64() - t_start;  // MMA role total
      g_dbg[9] += my_tiles;
      g_dbg[11] += dbg_issue;
    }
  } else if (is_epilogue_warp(warp)) {
    const int q = warp & 3;
    const int eg = epilogue_group(warp);  // 16-column chunks c with c % 2 == eg
    const int r = q * 32 + lane;
    const int w = r % p.w;
    const int hr = r / slab_px;              // slab (image row) inside the half
    const int img = (r % slab_px) / p.w;     // image inside the slab
    const bool edbg = dbg && warp == 2 && lane == 0;
    float* stage = epi_stage + (eg * 4 + q) * kEpiWarpFloats;
    float col_acc[N_TILE / 32][8];
#pragma unroll
    for (int cc = 0; cc < N_TILE / 32; ++cc)
#pragma unroll
      for (int j = 0; j < 8; ++j) col_acc[cc][j] = 0.f;
    int tile_i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_i) {
      const int mt = tile / p.n_tiles;
      const int n0 = (p.imgs == 1) ? mt / tiles_per_img : mt * p.halves * p.imgs;
      const int h0 = (p.imgs == 1) ? (mt % tiles_per_img) * p.halves * p.th : 0;
      const int n_tile0 = (tile % p.n_tiles) * N_TILE;
      const int buf = tile_i & 1;
      {
        const long long _t = clock64();
        mbar_wait(&acc_full[buf], (tile_i >> 1) & 1, 26);
        dbg_acc[0] += clock64() - _t;
      }
      const long long t_epi = clock64();
      tc_fence_after();
#pragma unroll 1
      for (int half = 0; half < p.halves; ++half) {
        const int n_img = (p.imgs == 1) ? n0 : n0 + half * p.imgs + img;
        const int h_row = (p.imgs == 1) ? h0 + half * p.th + hr : hr;
        const long long row_off = (long long)n_img * p.out_sn + (long long)h_row * p.out_sh +
                                  (long long)w * p.out_sw + n_tile0;
#pragma unroll
        for (int cc = 0; cc < N_TILE / 32; ++cc) {  // unrolled: col_acc must stay in registers
          const int c = 2 * cc + eg;
          uint32_t v[16];
          const uint32_t taddr =
              tmem_base + (uint32_t(q * 32) << 16) + buf * 2 * Cfg::kUmmaN + half * Cfg::kUmmaN + c * 16;
          tmem_ld_32x16(taddr, v);
          if (Cfg::kStack) {
            uint32_t v2[16];
            tmem_ld_32x16(taddr + 64, v2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
          } else {
            tmem_ld_wait();
          }
          warp_store_rows16(stage, v, p.out, row_off, true, c * 16, p.accumulate != 0, lane, col_acc[cc],
                            p.stats != nullptr);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      dbg_acc[1] += clock64() - t_epi;
    }
    if (edbg) {
      g_dbg[6] += dbg_acc[0];
      g_dbg[7] += dbg_acc[1];
      g_dbg[8] += clock64() - t_start;
    }
    if (p.stats)
      flush_column_stats<N_TILE>(epi_stage, col_acc, q, eg, lane, p.stats, blockIdx.x / p.n_tiles, p.n_total,
                                 (blockIdx.x % p.n_tiles) * N_TILE);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int N_TILE, int PA, int PB>
static int launch_conv3x3(const Conv3x3KParams& kp, cudaStream_t stream) {
  using Cfg = Conv3x3Cfg<N_TILE, PA, PB>;
  const int boxes = (kp.imgs == 1) ? 1 : kp.halves;
  const int box_rows = (kp.imgs == 1) ? kp.halves * kp.th + 2 : kp.th + 2;
  const int a_box_bytes = box_rows * kp.imgs * kp.w * 128;
  const int a_bytes = Cfg::kAStages * PA * boxes * a_box_bytes;
  int b_stages = (kSmemBudget - 1024 - a_bytes) / Cfg::kBStageBytes;  // kSmemBudget already excludes the epilogue patch
  if (b_stages > 8) b_stages = 8;
  if (b_stages < 2) {
    set_error("fb_conv3x3: tile does not fit into shared memory (W %d, N_TILE %d)", kp.w, N_TILE);
    return FB_ERR_UNSUPPORTED;
  }
  const int smem = a_bytes + b_stages * Cfg::kBStageBytes + kEpiStageBytes + 1024 + 512;
  static int configured = 0;
  if (configured < smem) {
    FB_CUDA(cudaFuncSetAttribute(conv3x3_kernel<N_TILE, PA, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const int m_tiles = (kp.imgs == 1) ? kp.n * (kp.h / (kp.halves * kp.th)) : kp.n / (kp.halves * kp.imgs);
  const int tiles = m_tiles * kp.n_tiles;
  int grid = tiles < kNumSMs ? tiles : kNumSMs;
  if (kp.stats) grid = (grid / kp.n_tiles) * kp.n_tiles;
  FB_CUDA(launch_pdl(conv3x3_kernel<N_TILE, PA, PB>, dim3(grid), dim3(kThreads), smem, stream, kp, a_box_bytes, b_stages));
  return 0;
}

template <int N_TILE>
static int dispatch_conv3x3(const Conv3x3KParams& kp, int pa, int pb, cudaStream_t stream) {
  if (pa == 2 && pb == 2) return launch_conv3x3<N_TILE, 2, 2>(kp, stream);
  if (pa == 1 && pb == 2) return launch_conv3x3<N_TILE, 1, 2>(kp, stream);
  if (pa == 1 && pb == 1) return launch_conv3x3<N_TILE, 1, 1>(kp, stream);
  set_error("fb_conv3x3: unsupported operand planes (%d, %d)", pa, pb);
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
  int splits, n_pixblocks;
  float* partial;
  int debug;
  int halo;  // 1: a B stage is ONE haloed X box per plane ((tile_h + 2) rows) shared by the CTA's three dh taps
};

constexpr int kWgAStages = 2;
constexpr int kWgBStages = 8;              // 16 KiB X tiles; a ring stage is `planes` consecutive tiles
constexpr int kWgABytes = 2 * kATileBytes;  // two 64-channel chunks of dY: [chunk][128 pixels][64 co]
constexpr int kWgBBytes = kATileBytes;      // [128 pixels][64 ci]
constexpr int kWgTmemCols = 512;

__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ WgradKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
  uint64_t* turn_bar = accum_bar + 1;  // [2] issue-order token of the two MMA warps (see conv_gemm_kernel)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(turn_bar + 2);

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
    mbar_init(&turn_bar[0], 1);
    mbar_init(&turn_bar[1], 1);
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
  const int split = blockIdx.z;
  const int halo_box_bytes = (p.tile_h + 2) * p.tile_w * 128;  // haloed X box of one plane
  const int halo_row_bytes = p.tile_w * 128;                   // one image row = one dh shift
  int b_stages = b_region_bytes / (p.planes * (p.halo ? halo_box_bytes : kWgBBytes));
  b_stages = b_stages > kWgBStages ? kWgBStages : b_stages;  // the barrier arrays hold kWgBStages entries
  const int pb0 = int((long long)split * p.n_pixblocks / p.splits);
  const int pb1 = int((long long)(split + 1) * p.n_pixblocks / p.splits);
  const bool two_chunks = (co0 + 64) < p.cout;
  const int k_total = p.n_taps * p.cin;
  const bool dbg = p.debug && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  const long long t_start = clock64();
  long long dbg_acc[2] = {0, 0}, dbg_issue = 0;

  if (warp == 0) {
    // ---------------- TMA producer (warp-converged, one elected lane issues) ----------------
    int as = 0, bs = 0;
    uint32_t aphase = 0, bphase = 0;
    for (int pb = pb0; pb < pb1; ++pb) {
      int n0, h0;
      tile_origin(pb, p.tile_h, p.tile_n, p.grid_h, n0, h0);
      FB_DBG_WAIT(0, mbar_wait(&a_empty[as], aphase ^ 1, 11));
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
        FB_DBG_WAIT(1, mbar_wait(&b_empty[bs], bphase ^ 1, 12));
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
        FB_DBG_WAIT(1, mbar_wait(&b_empty[bs], bphase ^ 1, 12));
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
    if (dbg && lane == 0) {
      g_dbg[0] += dbg_acc[0];
      g_dbg[1] += dbg_acc[1];
      g_dbg[10] += clock64() - t_start;
    }
  } else if (warp == 1 || warp == 6) {
    // ---------------- MMA issuer(s): one elected lane per warp, see conv_gemm_kernel and kMmaIssuers ------------
    // both operands MN-major; the hi and lo X tiles of a stage are adjacent, so ONE instruction with N = 64*planes
    // computes dY^T*[X_hi, X_lo] into two 64-column halves that the epilogue adds (an SS-mode MMA re-reads its
    // 128x16 A tile per instruction, N = 64 would run the tensor pipe at half rate)
    const uint32_t idesc = make_idesc_bf16(128, 64 * p.planes, 1, 1);
    constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
    const uint32_t smem_a0 = smem_u32(smem_a), smem_b0 = smem_u32(smem_b);
    const int me = (warp == 1) ? 0 : 1;
    const int total_stages = (me < kMmaIssuers) ? (pb1 - pb0) * n_slots : 0;
    const int all_stages = (pb1 - pb0) * n_slots;
    int bs = me % b_stages;
    uint32_t bphase = (me / b_stages) & 1, tphase = 0;
    int j = me, pbi = 0;  // slot inside the pixel block, pixel block index relative to pb0
    while (j >= n_slots) {
      j -= n_slots;
      ++pbi;
    }
    if (p.halo && me == 0) {
      // halo mode: the three dh slots of a pixel block share ONE X stage -> one wait / elect / commit round per pixel
      // block (24 MMAs) instead of three (the issuing warp's per-round work is dead time for the tensor pipe)
      for (int pb = 0; pb < pb1 - pb0; ++pb) {
        const int as = pb % kWgAStages;
        FB_DBG_WAIT(0, mbar_wait(&a_full[as], (pb / kWgAStages) & 1, 13));
        FB_DBG_WAIT(1, mbar_wait(&b_full[bs], bphase, 14));
        const uint32_t a_lo = smem_desc_lo(smem_a0 + as * a_stage_bytes, kATileBytes);
        const uint32_t b0_lo = smem_desc_lo(smem_b0 + bs * p.planes * halo_box_bytes, halo_box_bytes);
        tc_fence_after();
        const long long t_issue = dbg ? clock64() : 0;
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
          if (pb == pb1 - pb0 - 1) tc_commit(accum_bar);
        }
        __syncwarp();
        if (dbg) dbg_issue += clock64() - t_issue;
        if (++bs == b_stages) {
          bs = 0;
          bphase ^= 1;
        }
      }
    }
    for (int it = me; it < (p.halo ? 0 : total_stages); it += kMmaIssuers) {
      const int as = pbi % kWgAStages;
      FB_DBG_WAIT(0, mbar_wait(&a_full[as], (pbi / kWgAStages) & 1, 13));
      FB_DBG_WAIT(1, mbar_wait(&b_full[bs], bphase, 14));
      // K = 16 pixels = 16 rows of 128 bytes; MN chunks of 64 channels are kATileBytes apart (LBO), groups of 8
      // pixel rows are 1024 bytes apart (SBO).  Halo mode: slot j is the view shifted by j rows of the stage's box and
      // the hi / lo planes are one box apart.
      const uint32_t a_lo = smem_desc_lo(smem_a0 + as * a_stage_bytes, kATileBytes);
      const uint32_t b_lo = p.halo ? smem_desc_lo(smem_b0 + bs * p.planes * halo_box_bytes + j * halo_row_bytes,
                                                  halo_box_bytes)
                                   : smem_desc_lo(smem_b0 + bs * p.planes * kWgBBytes, kATileBytes);
      const uint32_t tmem_d = tmem_base + j * 64 * p.planes;
      if (kMmaIssuers == 2 && it > 0) {
        mbar_wait(&turn_bar[me], tphase, 16);
        tphase ^= 1;
      }
      tc_fence_after();
      const long long t_issue = dbg ? clock64() : 0;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kTileM / 16; ++k)
          tc_mma_bf16_lohi(tmem_d, a_lo + ((k * 2048) >> 4), b_lo + ((k * 2048) >> 4), desc_hi, desc_hi, idesc,
                           (pbi != 0 || k != 0) ? 1u : 0u);
        if (kMmaIssuers == 2) mbar_arrive(&turn_bar[me ^ 1]);
        if (!p.halo || j == n_slots - 1) tc_commit(&b_empty[bs]);
        if (j == n_slots - 1) tc_commit(&a_empty[as]);
        if (it == all_stages - 1) tc_commit(accum_bar);
      }
      __syncwarp();
      if (dbg) dbg_issue += clock64() - t_issue;
      if (!p.halo) {
        bs += kMmaIssuers;
      } else if (j == n_slots - 1) {  // halo mode (single issuer): the B stage advances once per pixel block
        bs += 1;
      }
      while (bs >= b_stages) {
        bs -= b_stages;
        bphase ^= 1;
      }
      j += kMmaIssuers;
      while (j >= n_slots) {
        j -= n_slots;
        ++pbi;
      }
    }
    if (dbg && lane == 0 && me == 0) {
      g_dbg[2] += dbg_acc[0];
      g_dbg[3] += dbg_acc[1];
      g_dbg[5] += clock64() - t_start;
      g_dbg[9] += pb1 - pb0;
      g_dbg[11] += dbg_issue;
    }
  } else if (is_epilogue_warp(warp)) {
    const int q = warp & 3;
    const int eg = epilogue_group(warp);  // 16-column chunks c with c % 2 == eg
    const int co = co0 + q * 32 + lane;
    const bool valid = co < p.cout;
    const long long row_off = ((long long)split * p.cout + co) * k_total;
    const bool edbg = dbg && warp == 2 && lane == 0;
    mbar_wait(accum_bar, 0, 15);
    const long long t_epi = clock64();
    if (edbg) g_dbg[6] += t_epi - t_start;
    tc_fence_after();
    for (int j = 0; j < n_slots; ++j) {
      const int s = slot0 + j;
      const int tap = p.taps[s % p.n_taps].k_index;  // filter position = column block of the partial matrix
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
        warp_store_rows16(epi_stage + (eg * 4 + q) * kEpiWarpFloats, v, p.partial, row_off, valid,
                          tap * p.cin + cb * kBlockK + c * 16, false, lane);
      }
    }
    if (edbg) {
      g_dbg[7] += clock64() - t_epi;
      g_dbg[8] += clock64() - t_start;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kWgTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad finalize: deterministic split-K reduction + scatter to OIHW
// ---------------------------------------------------------------------------------------------------------------
// fixed-order sum over split-K partials with 8 loads in flight
__device__ __forceinline__ float sum_splits(const float* __restrict__ src, int splits, long long stride) {
  float acc = 0.f;
  int s = 0;
  for (; s + 8 <= splits; s += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = src[(s + j) * stride];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j];
  }
  for (; s < splits; ++s) acc += src[s * stride];
  return acc;
}

// block = (co, 32-wide ci block); blockDim = 32 * taps
__global__ void wgrad_finalize_kernel(const float* __restrict__ partial, int splits, int cout, int cin, int taps,
                                      int cin_stored, int mode, float* __restrict__ g) {
  __shared__ float tile[32 * 9];
  griddep_wait();
  griddep_launch();
  const int co = blockIdx.y;
  const int ci0 = blockIdx.x * 32;
  const int t = threadIdx.x;
  const int k_total = (mode == 0) ? taps * cin_stored : cin_stored;
  const long long split_stride = (long long)cout * k_total;
  if (mode == 0) {
    const int tap = t / 32, cil = t % 32;
    float acc = 0.f;
    if (ci0 + cil < cin) {
      const float* src = partial + (long long)co * k_total + tap * cin_stored + ci0 + cil;
      acc = sum_splits(src, splits, split_stride);
    }
    tile[cil * taps + tap] = acc;
    __syncthreads();
    const int ci = ci0 + t / taps;
    if (ci < cin) g[((long long)co * cin + ci0) * taps + t] = tile[t];
  } else {
    // columns already in (ci, tap) order: plain reduction of the first cin*taps columns
    const int col = ci0 * taps + t;
    if (col < cin * taps) {
      const float* src = partial + (long long)co * k_total + col;
      const float acc = sum_splits(src, splits, split_stride);
      g[(long long)co * cin * taps + col] = acc;
    }
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

// ---------------------------------------------------------------------------------------------------------------
// all conv weights of the network in ONE launch: table of per-layer descriptors, block -> layer by binary search
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void weight_prep_tile(const float* __restrict__ w, int cout, int cin, int taps, int co0,
                                                 int ci0, __nv_bfloat16* __restrict__ wf_hi,
                                                 __nv_bfloat16* __restrict__ wf_lo, long long ld_f,
                                                 __nv_bfloat16* __restrict__ wd_hi, __nv_bfloat16* __restrict__ wd_lo,
                                                 long long ld_d, float* wtile) {
  const int row_len = 32 * taps;
  const int pitch = row_len + 1;
  for (int i = threadIdx.x; i < 32 * row_len; i += blockDim.x) {
    const int co = i / row_len, r = i % row_len;
    wtile[co * pitch + r] = w[((long long)(co0 + co) * cin + ci0) * taps + r];
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
                                                                const fb_wprep_entry* __restrict__ table, int n) {
  extern __shared__ float wtile[];
  griddep_wait();
  griddep_launch();
  int lo = 0, hi = n - 1;
  while (lo < hi) {  // last entry with block_start <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].block_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const fb_wprep_entry e = table[lo];
  const int b = blockIdx.x - e.block_start;
  const float* w = theta + e.w_offset;
  if (e.cin % 32 != 0) {  // stem: wf[co][k] = w[co][k], k = ci*taps + tap
    const int k = e.cin * e.taps;
    __nv_bfloat16* wf_hi = static_cast<__nv_bfloat16*>(e.wf_hi);
    __nv_bfloat16* wf_lo = static_cast<__nv_bfloat16*>(e.wf_lo);
    for (int i = b * 256 + threadIdx.x; i < e.cout * k; i += e.n_blocks * 256) {
      __nv_bfloat16 h, l;
      split_bf16(w[i], h, l);
      wf_hi[(i / k) * e.ld_f + i % k] = h;
      if (wf_lo) wf_lo[(i / k) * e.ld_f + i % k] = l;
    }
    return;
  }
  const int ci_blocks = e.cin / 32;
  weight_prep_tile(w, e.cout, e.cin, e.taps, (b / ci_blocks) * 32, (b % ci_blocks) * 32,
                   static_cast<__nv_bfloat16*>(e.wf_hi), static_cast<__nv_bfloat16*>(e.wf_lo), e.ld_f,
                   static_cast<__nv_bfloat16*>(e.wd_hi), static_cast<__nv_bfloat16*>(e.wd_lo), e.ld_d, wtile);
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

}  // namespace fb

// =================================================================================================================
// C ABI
// =================================================================================================================
using namespace fb;

extern "C" int fb_version(void) { return 100; }

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
  kp.m_tiles = (a->tile_n == 1) ? a->grid_n * (a->grid_h / a->tile_h) : (a->grid_n + a->tile_n - 1) / a->tile_n;
  kp.n_tiles = a->n_total / a->n_tile;
  kp.out = a->out;
  kp.out_sn = a->out_sn;
  kp.out_sh = a->out_sh;
  kp.out_sw = a->out_sw;
  kp.accumulate = a->accumulate;
  kp.stats = a->stats_out;
  kp.n_total = a->n_total;
  kp.bwd_y = a->bwd_y;
  kp.bwd_mask = static_cast<const __nv_bfloat16*>(a->bwd_mask);
  kp.bwd_mean = a->bwd_mean;
  kp.bwd_rstd = a->bwd_rstd;
  FB_REQUIRE(!a->bwd_y || (a->stats_out && a->bwd_mean && a->bwd_rstd && a->n_groups <= 1),
             "fb_conv_gemm: bwd_y needs stats_out, bwd_mean, bwd_rstd and a single tap group");
  FB_REQUIRE(a->n_groups >= 0 && a->n_groups <= 4, "fb_conv_gemm: n_groups must be in 0..4");
  FB_REQUIRE(a->n_groups <= 1 || !a->stats_out, "fb_conv_gemm: tap groups cannot be combined with stats_out");
  kp.n_groups = a->n_groups > 0 ? a->n_groups : 1;
  for (int g = 0; g < 4; ++g) {
    kp.group_tap0[g] = 0;
    kp.group_taps[g] = a->n_taps;
    kp.group_off[g] = 0;
  }
  if (a->n_groups > 0)
    for (int g = 0; g < a->n_groups; ++g) {
      FB_REQUIRE(a->groups[g].tap0 >= 0 && a->groups[g].n_taps >= 1 &&
                     a->groups[g].tap0 + a->groups[g].n_taps <= a->n_taps,
                 "fb_conv_gemm: tap group %d out of range", g);
      kp.group_tap0[g] = a->groups[g].tap0;
      kp.group_taps[g] = a->groups[g].n_taps;
      kp.group_off[g] = a->groups[g].out_off;
    }
  {
    // development only (tools/conv_experiments.py): honoured only together with FB_KERNEL_DEBUG=1, read per call
    static const bool dev = [] {
      const char* d = getenv("FB_KERNEL_DEBUG");
      return d && d[0] == '1';
    }();
    const char* e = dev ? getenv("FB_CONV_EXPERIMENT") : nullptr;
    kp.experiment = e ? atoi(e) : 0;
  }
  FB_REQUIRE(!a->stats_out || !a->accumulate, "fb_conv_gemm: column statistics need accumulate == 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->n_tile) {
    case 64: return dispatch_conv_gemm<64>(kp, a->a_planes, a->b_planes, st);
    case 128: return dispatch_conv_gemm<128>(kp, a->a_planes, a->b_planes, st);
    default: return dispatch_conv_gemm<256>(kp, a->a_planes, a->b_planes, st);
  }
}

extern "C" int fb_conv3x3(const fb_conv3x3_args* a, void* stream) {
  FB_REQUIRE(a && a->host_a_maps && a->host_b_maps && a->out, "fb_conv3x3: null pointer");
  FB_REQUIRE(a->a_planes >= 1 && a->a_planes <= 2 && a->b_planes >= 1 && a->b_planes <= 2, "fb_conv3x3: bad planes");
  const int imgs = a->imgs > 0 ? a->imgs : 1;
  const int halves = a->halves > 0 ? a->halves : 2;
  FB_REQUIRE(halves == 1 || halves == 2, "fb_conv3x3: halves must be 1 or 2");
  const bool geom_ok =
      (imgs == 1) ? (a->w >= 8 && a->w <= 128 && 128 % a->w == 0 && a->h % (halves * (128 / a->w)) == 0)
                  : (imgs * a->w * a->h == 128 && (imgs * a->w) % 8 == 0 && a->n % (halves * imgs) == 0);
  if (!geom_ok) {
    set_error("fb_conv3x3: %d images of %dx%d (imgs %d, halves %d) not supported by the haloed tiling", a->n, a->h,
              a->w, imgs, halves);
    return FB_ERR_UNSUPPORTED;
  }
  if (!(a->n_tile == 64 || a->n_tile == 128) || a->n_total % a->n_tile != 0) {
    set_error("fb_conv3x3: unsupported n_tile %d for n_total %d", a->n_tile, a->n_total);
    return FB_ERR_UNSUPPORTED;
  }
  FB_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && a->out_sn % 4 == 0 && a->out_sh % 4 == 0 &&
                 a->out_sw % 4 == 0,
             "fb_conv3x3: output must be 16-byte aligned with strides multiple of 4");
  Conv3x3KParams kp;
  memset(&kp, 0, sizeof(kp));
  memcpy(kp.a_maps, a->host_a_maps, size_t(a->a_planes) * FB_TMAP_BYTES);
  memcpy(kp.b_maps, a->host_b_maps, size_t(a->b_planes) * FB_TMAP_BYTES);
  memcpy(kp.b_k0, a->b_k0, sizeof(kp.b_k0));
  kp.cblocks = a->cblocks;
  kp.w = a->w;
  kp.h = a->h;
  kp.n = a->n;
  kp.th = 128 / (imgs * a->w);
  kp.imgs = imgs;
  kp.halves = halves;
  kp.n_tiles = a->n_total / a->n_tile;
  kp.out = a->out;
  kp.out_sn = a->out_sn;
  kp.out_sh = a->out_sh;
  kp.out_sw = a->out_sw;
  kp.accumulate = a->accumulate;
  static const bool debug = [] {
    const char* e = getenv("FB_KERNEL_DEBUG");
    return e && e[0] == '1';
  }();
  kp.debug = debug ? 1 : 0;
  kp.stats = a->stats_out;
  kp.n_total = a->n_total;
  FB_REQUIRE(!a->stats_out || !a->accumulate, "fb_conv3x3: column statistics need accumulate == 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->n_tile == 64) return dispatch_conv3x3<64>(kp, a->a_planes, a->b_planes, st);
  return dispatch_conv3x3<128>(kp, a->a_planes, a->b_planes, st);
}

extern "C" int fb_conv_wgrad(const fb_wgrad_args* a, void* stream) {
  FB_REQUIRE(a && a->host_dy_map && a->host_x_maps && a->partial, "fb_conv_wgrad: null pointer");
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
  const int n_pixblocks =
      (a->tile_n == 1) ? a->grid_n * (a->grid_h / a->tile_h) : (a->grid_n + a->tile_n - 1) / a->tile_n;
  FB_REQUIRE(a->splits >= 1 && a->splits <= n_pixblocks, "fb_conv_wgrad: splits %d must be in 1..%d", a->splits,
             n_pixblocks);
  FB_REQUIRE(a->cout % 64 == 0, "fb_conv_wgrad: cout must be a multiple of 64");
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
  kp.n_pixblocks = n_pixblocks;
  kp.partial = a->partial;
  kp.halo = a->halo ? 1 : 0;
  if (a->halo) {
    FB_REQUIRE(kMmaIssuers == 1, "fb_conv_wgrad: halo mode needs the single-issuer build");
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
  {
    static const bool debug = [] {
      const char* e = getenv("FB_KERNEL_DEBUG");
      return e && e[0] == '1';
    }();
    kp.debug = debug ? 1 : 0;
  }
  const int n_slots_total = a->n_taps * a->cblocks;
  dim3 grid((a->cout + 127) / 128, (n_slots_total + a->slots_per_cta - 1) / a->slots_per_cta, a->splits);
  constexpr int smem = kWgAStages * kWgABytes + kWgBStages * kWgBBytes + kEpiStageBytes + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    FB_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  FB_CUDA(launch_pdl(wgrad_kernel, grid, dim3(kThreads), smem, static_cast<cudaStream_t>(stream), kp));
  return 0;
}

extern "C" int fb_wgrad_finalize(const float* partial, int splits, int cout, int cin, int taps, int cin_stored, int mode,
                                 float* g_oihw, void* stream) {
  FB_REQUIRE(partial && g_oihw && splits >= 1, "fb_wgrad_finalize: bad arguments");
  FB_REQUIRE(taps == 1 || taps == 9, "fb_wgrad_finalize: taps must be 1 or 9");
  FB_REQUIRE(mode == 1 || cin % 32 == 0, "fb_wgrad_finalize: cin must be a multiple of 32 in mode 0");
  dim3 grid((cin + 31) / 32, cout);
  FB_CUDA(launch_pdl(wgrad_finalize_kernel, grid, dim3(32 * taps), 0, static_cast<cudaStream_t>(stream), partial, splits,
                     cout, cin, taps, cin_stored, mode, g_oihw));
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
                                    void* stream) {
  FB_REQUIRE(theta && table_dev && n_entries > 0 && total_blocks > 0, "fb_weight_prep_multi: bad arguments");
  const size_t smem = size_t(32) * (32 * 9 + 1) * sizeof(float);
  FB_CUDA(launch_pdl(weight_prep_multi_kernel, dim3(total_blocks), dim3(256), smem, static_cast<cudaStream_t>(stream),
                     theta, table_dev, n_entries));
  return 0;
}

/* development aid: read (and clear) the in-kernel cycle counters of fb_conv3x3 (FB_KERNEL_DEBUG=1) */
extern "C" int fb_debug_counters(long long* host32, int clear) {
  FB_CUDA(cudaDeviceSynchronize());
  FB_CUDA(cudaMemcpyFromSymbol(host32, g_dbg, sizeof(long long) * 32));
  if (clear) {
    long long zero[32] = {0};
    FB_CUDA(cudaMemcpyToSymbol(g_dbg, zero, sizeof(zero)));
  }
  return 0;
}

extern "C" int fb_conv_stats_rows(int m_tiles, int n_tiles) {
  const int tiles = m_tiles * n_tiles;
  int grid = tiles < kNumSMs ? tiles : kNumSMs;
  grid = (grid / n_tiles) * n_tiles;
  return grid / n_tiles;
}
