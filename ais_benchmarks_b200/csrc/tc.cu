// Tensor-core (tcgen05 / TMEM) cross-term path for equal-covariance, equal-weight mixtures (the DM-AIS proposal
// mixture of reference sampling_methods/dm_ais.py:36-45,72-76 and the KDE of base.py:162-179) -- BASELINE.json
// configs[4] "tensor-core cross-term path".
//
//   log2( N-term ) :  t_ij = -||u_i - v_j||^2 = -||u_i||^2 + ( 2 u_i . v_j - ||v_j||^2 ),   u = a x,  v = a mu
//
// The bracket is one row of a GEMM  S = A B^T  with
//     A_i = [ tf32(2u_i), 1, 1, 1, 0.. ]            (128 points per row tile, built in shared memory per tile)
//     B_j = [ tf32(v_j),  w_h, w_m, w_l, 0.. ]       (w = -||v_j||^2 split in three TF32 terms; packed once)
// executed by tcgen05.mma kind::tf32 with FP32 accumulators in TMEM; the FP32 pipe only sees the epilogue:
// tcgen05.ld, ex2, sum (MUFU-bound instead of FMA-bound). A single TF32 product is enough because the tensor-core
// value is only a screen (see "Accuracy"); K = DP + 8 keeps the B stream (the L2 -> smem traffic that every CTA
// repeats for every 256 points) and the MMA count 2.6x below a 3xTF32 split at D = 32.
//
// CTA = 10 warps working on 256 points at a time (two M = 128 row tiles that share every B tile, which halves the L2
// traffic of the B stream): warps 0-7 build the two A tiles and run the epilogue (thread = point, TMEM lane = row),
// warp 8 streams B tiles with 1-D TMA bulk copies into an mbarrier ring, warp 9 owns TMEM and issues the MMAs.
// All 512 TMEM columns are used: 2 row tiles x 2 buffers x 128 columns, so the MMAs of component tile n+1 overlap the
// epilogue of tile n.
//
// Accuracy. The tensor-core value carries an ABSOLUTE error of up to ~0.3 in the exponent (TF32 inputs; even a 3xTF32
// split measured 2e-4, because the FP32 accumulators truncate at the magnitude of ||u||^2 ~ 100), far too much for the
// 1e-5 log-density bar if it were used for the terms that matter. It is therefore only a SCREEN: every pair whose
// screened exponent comes within kTcRefine (2^-30 of the reference term) is re-evaluated exactly, with the same centred
// FP32 form as the CUDA-core path, and only the exact value enters the sum; the screened values are used for the
// remaining mass (each term below 1e-9 of the reference term, in total below K * 2^-30, where a 25 % relative error
// is irrelevant). The reference term comes from a per-point HINT:
// the index of a component expected to dominate (for DM-AIS: the proposal the sample was drawn from), evaluated
// exactly up front. In 32-D the dominant set is the hinted component plus, rarely, a neighbour, so the exact work is
// ~0.1 % of the pairs and, because the samples of a tile share their proposal, it is not even divergent.
//
// Every wait is bounded: a protocol error sets an abort flag and the kernel drains instead of hanging.
#include <stdlib.h>

#include "common.cuh"

namespace aisb {

constexpr int kTcRows = 128;     // rows per MMA = UMMA M
constexpr int kTcRowTiles = 2;   // row tiles per CTA iteration (256 points share each B tile)
constexpr int kTcN = 128;        // components per B tile = UMMA N
constexpr int kTcEpiThreads = kTcRows * kTcRowTiles;
constexpr int kTcThreads = kTcEpiThreads + 64;
constexpr int kTcAccCols = 2 * kTcRowTiles * kTcN;  // 512: all of TMEM

__host__ __device__ constexpr int tc_kcols(int DP) { return DP + 8; }                // multiple of 8 for DP >= 8
__host__ __device__ constexpr int tc_tile_floats(int DP) { return tc_kcols(DP) * kTcN; }
__host__ __device__ constexpr int tc_stages(int DP) { return 4; }
__host__ __device__ constexpr size_t tc_smem_bytes(int DP) {  // A tiles are double-buffered across iterations
  return sizeof(float) * tc_tile_floats(DP) * (2 * kTcRowTiles + tc_stages(DP)) + 1024;
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ bool mbar_try_wait_ns(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return done != 0;
}

// bounded wait: ~2 s worst case, then the CTA-wide abort flag is raised
__device__ __forceinline__ bool tc_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
  for (uint32_t spin = 0; spin < (1u << 21); ++spin) {
    if (mbar_try_wait_ns(bar, parity, 1000u)) return true;
    if (*abort_flag) return false;
  }
  *abort_flag = 1;
  return false;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major, no-swizzle UMMA shared-memory descriptor. Tile layout: 16-byte K-chunk c of row r at c*(rows*16) + r*16,
// i.e. core matrices (8 rows x 16 B) are 128 B apart along M/N (SBO) and rows*16 B apart along K (LBO).
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr, uint32_t rows) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(((rows * 16u) >> 4) & 0x3FFFu) << 16;  // leading (K) byte offset
  d |= static_cast<uint64_t>((128u >> 4) & 0x3FFFu) << 32;          // stride (M/N) byte offset
  d |= static_cast<uint64_t>(1) << 46;                              // descriptor version (sm_100)
  return d;
}

// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = kTcN
__host__ __device__ constexpr uint32_t tc_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(kTcN >> 3) << 17) |
         (static_cast<uint32_t>(kTcRows >> 4) << 24);
}

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr float kTcRefine = 30.f;    // screened exponents within 2^-30 of the reference term are re-evaluated exactly
constexpr float kTcSkipMargin = 34.f;  // chunks entirely below 2^-(refine + 34) of the reference term are skipped
constexpr int kTcTrailer = 8;        // floats behind the image: operand magnitudes for the per-point screen error bound

// The screen's absolute error grows with the operand magnitudes (TF32 inputs: relative 2^-11 each, 2^-10 per product).
// Each point widens its refine window by twice its own worst-case bound E, so a term that matters is never left to
// the screen; for data far from the origin the kernel degrades towards exact evaluation instead of losing accuracy.
struct TcWindow {
  float refine;   // kTcRefine + 2E
  float skip;     // refine + kTcSkipMargin
  float floor_;   // 2^-refine: what the fast path adds for a column that the slow path then replaces
};
__device__ __forceinline__ TcWindow tc_window(float err_bound) {
  TcWindow w;
  w.refine = kTcRefine + 2.f * err_bound;
  w.skip = w.refine + kTcSkipMargin;
  w.floor_ = ex2(-w.refine);
  return w;
}

// exact log2-domain exponent of (point, component j): the centred FP32 form of the CUDA-core path on the packed
// ISO_SHARED row (b = -a mu), read from global memory (L1/L2 resident)
template <int DP>
struct TcExactIso {
  const float (&xr)[DP];
  float a_iso;
  const float* __restrict__ rows;
  __device__ __forceinline__ float operator()(int64_t j) const {
    const float* row = rows + j * DP;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const float dl = fmaf(xr[d], a_iso, __ldg(row + d));
      acc = fmaf(-dl, dl, acc);
    }
    return acc;
  }
};

// maximum of a 32-column chunk: four independent chains of three-input maxima (a single chain is 16 dependent FMNMX3,
// ~70 cycles of latency per chunk for a warp that has only one other warp on its scheduler to hide behind)
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float tc_chunk_max(const uint32_t (&r)[32]) {
  float g[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    g[q] = max3f(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1]), __uint_as_float(r[8 * q + 2]));
    g[q] = max3f(g[q], __uint_as_float(r[8 * q + 3]), __uint_as_float(r[8 * q + 4]));
    g[q] = max3f(g[q], __uint_as_float(r[8 * q + 5]), __uint_as_float(r[8 * q + 6]));
    g[q] = fmaxf(g[q], __uint_as_float(r[8 * q + 7]));
  }
  return fmaxf(max3f(g[0], g[1], g[2]), g[3]);
}

// fold 32 screened accumulator values (columns j0 .. j0+31, chunk maximum gr) into the partial sums, relative to the
// fixed reference exponent m:   screened term  e = ex2(v - c)   (v = t + ||u||^2, c = ||u||^2 + m).
// Fast path: four independent partial sums, no predicates. If any screened exponent comes within the refine window of
// the reference the rare slow path swaps the screened term of each such column for its exact value.
template <class EX>
__device__ __forceinline__ void tc_fold32(const uint32_t (&r)[32], float gr, float c, const TcWindow& win, float (&S)[4],
                                          float& m, const EX& exact, int64_t j0, int64_t K, int64_t j_hint,
                                          float m_hint) {
  // chunk maximum on the raw accumulators first (16 three-input FMNMX on the ALU pipe): a chunk whose largest term is
  // far below the reference cannot change a float32 sum (K * 2^-64 << 2^-24) and is dropped without touching the FMA
  // or MUFU pipes -- in 32-D that is almost every chunk (far proposals sit at 2^-280)
  if (!(gr - c > -win.skip)) {
    if (gr == gr) return;  // NaN falls through to the arithmetic path and propagates
  }
  float a[32];
  float g = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    a[i] = __uint_as_float(r[i]) - c;
    g = fmaxf(g, a[i]);
    S[i & 3] += ex2(fminf(a[i], -win.refine));  // a term that matters enters as 2^-refine until the slow path fixes it
  }
  if (g > -win.refine) {
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) mask |= (a[i] > -win.refine) ? (1u << i) : 0u;
    while (mask) {  // single call site: the exact evaluation stays inlined and the point stays in registers
      const int i = __ffs(mask) - 1;
      mask &= mask - 1;
      const int64_t j = j0 + i;
      if (j >= K) continue;
      S[0] -= win.floor_;  // what the fast path added for this column
      // a hinted component was evaluated up front: it IS the reference, its term is exactly 1
      const float t = j == j_hint ? m_hint : exact(j);
      if (t - m > 60.f) {  // a component far above the reference: move the reference (keeps ex2 finite)
        const float f = ex2(m - t);
        S[0] = S[0] * f + 1.f;
        S[1] *= f;
        S[2] *= f;
        S[3] *= f;
        m = t;
      } else {
        S[0] += ex2(t - m);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// B image: per tile of 128 components the exact shared-memory image [K/4 chunks][128 rows][4 floats].
// Input: the ISO_SHARED packed rows (b = -a mu). Padding components get w = -1e30.
// ---------------------------------------------------------------------------------------------
// Trailer (zeroed by the caller): [0] = max_j ||v_j|| (as ordered int bits); [1] = sum_j #{k : ||v_j - v_k||^2 < kTcRefine}
// (crowding: how many components, itself included, sit inside a component's exact-refinement window -- the work the
// refinement does per sample drawn from that component; only computed for K <= kTcCrowdMaxK, else left at 0 = unknown).
constexpr int64_t kTcCrowdMaxK = 8192;
__global__ void tc_pack_kernel(const float* __restrict__ rows, int64_t K, int D, int DP, float* __restrict__ image) {
  const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t KT = (K + kTcN - 1) / kTcN * kTcN;
  if (j >= KT) return;
  int* trailer = reinterpret_cast<int*>(image + (KT / kTcN) * static_cast<int64_t>(tc_kcols(DP)) * kTcN);
  const int kc = tc_kcols(DP);
  float* tile = image + (j / kTcN) * static_cast<int64_t>(kc) * kTcN;
  const int r = static_cast<int>(j % kTcN);
  auto put = [&](int k, float val) { tile[(k / 4) * (kTcN * 4) + r * 4 + (k % 4)] = val; };
  for (int k = 0; k < kc; ++k) put(k, 0.f);
  if (j >= K) {
    put(DP, -1.0e30f);
    return;
  }
  double w = 0.0;
  for (int d = 0; d < D; ++d) {
    const float v = -rows[j * DP + d];
    put(d, to_tf32(v));
    w -= static_cast<double>(v) * static_cast<double>(v);
  }
  const float wh = to_tf32(static_cast<float>(w));
  const float wm = to_tf32(static_cast<float>(w - wh));
  const float wl = to_tf32(static_cast<float>(w - wh - wm));
  put(DP, wh);
  put(DP + 1, wm);
  put(DP + 2, wl);
  atomicMax(trailer, __float_as_int(static_cast<float>(sqrt(-w)) * 1.0001f));  // non-negative floats order as ints
  if (K <= kTcCrowdMaxK) {
    int cnt = 0;
    for (int64_t k = 0; k < K; ++k) {
      float d2 = 0.f;
      for (int d = 0; d < D; ++d) {
        const float dl = rows[j * DP + d] - rows[k * DP + d];
        d2 = fmaf(dl, dl, d2);
      }
      cnt += d2 < kTcRefine ? 1 : 0;
    }
    atomicAdd(reinterpret_cast<float*>(trailer) + 1, static_cast<float>(cnt));
  }
}

// Epilogue loop shapes (config 5, un-adapted proposals, same box): 64-column loop rolled / accumulator-tile loop
// rolled 3.16 ms; 64-column loop fully unrolled (2 trips) 2.98 ms -- the TMEM load of the next half is hoisted over the
// fold of this one and 112 B fewer spills; two accumulator tiles per trip as well 2.95 ms (the buffer index and
// barrier parity become compile-time per copy). Crowded proposals (many exact refinements per sample, a regime the
// dispatcher keeps on the CUDA cores) lose ~10 % to the larger loop body.
#ifndef AISB_TC_EPI_UNROLL
#define AISB_TC_EPI_UNROLL 2
#endif
#ifndef AISB_TC_TILE_UNROLL
#define AISB_TC_TILE_UNROLL 2
#endif
constexpr int kTcTileUnroll = AISB_TC_TILE_UNROLL;   // accumulator tiles per trip of the epilogue's tile loop
constexpr int kTcEpiUnroll = AISB_TC_EPI_UNROLL;     // trips of the epilogue's 64-column loop unrolled (kTcN / 64 = 2 in all)

struct TcBars {
  uint64_t full[4], empty[4], tfull[2], tempty[2], a_ready;  // tfull/tempty index = accumulator buffer (both row tiles)
  uint32_t tmem_base;
  int abort_flag;
};

template <int DP>
__global__ void __launch_bounds__(kTcThreads, 1)
    tc_logq_kernel(const float* __restrict__ x, int64_t n, int D, float a_iso, float uniform_offset,
                   const float* __restrict__ rows, int64_t K, const int32_t* __restrict__ hint,
                   const float* __restrict__ image, int ntiles_n, float* __restrict__ out, int* __restrict__ err,
                   int dbg) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ TcBars bars;
  constexpr int KC = tc_kcols(DP);
  constexpr int NST = tc_stages(DP);
  constexpr uint32_t TILE_BYTES = sizeof(float) * KC * kTcN;
  float* sA = reinterpret_cast<float*>(tc_smem);            // 2 (iteration parity) x kTcRowTiles A tiles
  float* sB = sA + 2 * kTcRowTiles * KC * kTcRows;           // NST B stages

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  volatile int* abort_flag = &bars.abort_flag;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bars.full[s], 1);
      mbar_init(&bars.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars.tfull[b], 1);
      mbar_init(&bars.tempty[b], kTcEpiThreads);
    }
    mbar_init(&bars.a_ready, kTcEpiThreads);
    bars.abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == kTcEpiThreads / 32 + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)),
                 "r"(static_cast<uint32_t>(kTcAccCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;

  constexpr int kRowsPerIter = kTcRows * kTcRowTiles;
  const int64_t ntiles_m = (n + kRowsPerIter - 1) / kRowsPerIter;

  if (warp == kTcEpiThreads / 32) {
    // ===== B producer =====
    if (lane == 0) {
      uint32_t itn = 0;
      for (int64_t tm = blockIdx.x; tm < ntiles_m && !*abort_flag; tm += gridDim.x) {
        for (int tn = 0; tn < ntiles_n; ++tn, ++itn) {
          const uint32_t s = itn % NST, ph = (itn / NST) & 1u;
          if (!tc_wait(&bars.empty[s], ph ^ 1u, abort_flag)) break;
          mbar_arrive_expect_tx(&bars.full[s], TILE_BYTES);
          bulk_g2s(sB + s * (KC * kTcN), image + static_cast<int64_t>(tn) * KC * kTcN, TILE_BYTES, &bars.full[s]);
        }
      }
    }
  } else if (warp == kTcEpiThreads / 32 + 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = tc_idesc();
      uint32_t itn = 0, itm = 0;
      for (int64_t tm = blockIdx.x; tm < ntiles_m && !*abort_flag; tm += gridDim.x, ++itm) {
        if (!tc_wait(&bars.a_ready, itm & 1u, abort_flag)) break;
        tc_fence_after();
        for (int tn = 0; tn < ntiles_n; ++tn, ++itn) {
          const uint32_t s = itn % NST, ph = (itn / NST) & 1u;
          const uint32_t b = itn & 1u, bph = (itn >> 1) & 1u;
          if (!tc_wait(&bars.tempty[b], bph ^ 1u, abort_flag)) break;
          if (!tc_wait(&bars.full[s], ph, abort_flag)) break;
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB + s * (KC * kTcN));
#pragma unroll 1
          for (int rt = 0; rt < kTcRowTiles; ++rt) {
#pragma unroll 1
            for (int ks = 0; ks < KC / 8; ++ks) {
              const uint64_t adesc = tc_smem_desc(
                  a_addr + ((itm & 1u) * kTcRowTiles + rt) * (KC * kTcRows * 4) + ks * 2 * (kTcRows * 16), kTcRows);
              const uint64_t bdesc = tc_smem_desc(b_addr + ks * 2 * (kTcN * 16), kTcN);
              if (!(dbg & 2) || ks == 0) tc_mma(tmem_base + (rt * 2 + b) * kTcN, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
            }
          }
          tc_commit(&bars.empty[s]);   // the stage may be refilled once these MMAs have read it
          tc_commit(&bars.tfull[b]);   // accumulator ready for the epilogue
        }
      }
    }
  } else {
    // ===== A builder + epilogue: thread = point; row tile = warp / 4, TMEM lane = row within the tile =====
    const int rt = threadIdx.x / kTcRows;
    const int row = threadIdx.x % kTcRows;
    const bool vec = (D == DP) && (reinterpret_cast<uintptr_t>(x) % 16 == 0);

    auto load_point = [&](int64_t tm, float (&xv)[DP]) {
      const int64_t p = tm * kRowsPerIter + threadIdx.x;
      const int64_t pc = p < n ? p : n - 1;
      if (vec) {
        const float4* r = reinterpret_cast<const float4*>(x + pc * DP);
#pragma unroll
        for (int q = 0; q < DP / 4; ++q) {
          const float4 f = __ldg(r + q);
          xv[4 * q] = f.x;
          xv[4 * q + 1] = f.y;
          xv[4 * q + 2] = f.z;
          xv[4 * q + 3] = f.w;
        }
      } else {
#pragma unroll
        for (int d = 0; d < DP; ++d) xv[d] = d < D ? __ldg(x + pc * D + d) : 0.f;
      }
    };
    // A tile row [tf32(2u), 1, 1, 1, 0...] into the buffer of iteration parity `par`; returns ||u||^2
    auto build_a = [&](const float (&xv)[DP], uint32_t par) -> double {
      double nu = 0.0;
      float4* arow = reinterpret_cast<float4*>(sA + (par * kTcRowTiles + rt) * (KC * kTcRows)) + row;  // chunk c at + c*128
#pragma unroll
      for (int c = 0; c < DP / 4; ++c) {
        float hv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float u = a_iso * xv[4 * c + e];
          nu += static_cast<double>(u) * static_cast<double>(u);
          hv[e] = to_tf32(2.f * u);
        }
        arow[c * kTcRows] = make_float4(hv[0], hv[1], hv[2], hv[3]);
      }
      arow[(DP / 4) * kTcRows] = make_float4(1.f, 1.f, 1.f, 0.f);
      arow[(DP / 4 + 1) * kTcRows] = make_float4(0.f, 0.f, 0.f, 0.f);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&bars.a_ready);
      return nu;
    };

    uint32_t itn = 0, itm = 0;
    float xr[DP], xn[DP];
    double nu = 0.0, nu_next = 0.0;
    const float vmax = __ldg(image + static_cast<int64_t>(ntiles_n) * KC * kTcN);
    if (blockIdx.x < ntiles_m) {
      load_point(blockIdx.x, xr);
      nu = build_a(xr, 0u);
    }
    for (int64_t tm = blockIdx.x; tm < ntiles_m && !*abort_flag; tm += gridDim.x, ++itm) {
      const int64_t p = tm * kRowsPerIter + threadIdx.x;
      const int64_t pc = p < n ? p : n - 1;
      const bool has_next = tm + gridDim.x < ntiles_m;
      if (has_next) load_point(tm + gridDim.x, xn);  // lands while this iteration's tiles are being swept

      // ---- reference term: the hinted component, evaluated exactly (it is picked up again, exactly, by the sweep) ----
      int64_t h = hint[pc];
      h = h < 0 ? 0 : (h >= K ? K - 1 : h);
      const TcExactIso<DP> exact{xr, a_iso, rows};
      float m = exact(h);
      const float m_hint = m;             // exact exponent of the hinted component (reused when the sweep reaches it)
      if (!(m > -1.0e30f)) m = -1.0e30f;  // NaN / -inf inputs: keep the arithmetic finite, the result is -inf / NaN anyway
      float c = static_cast<float>(nu + static_cast<double>(m));
      float S[4] = {0.f, 0.f, 0.f, 0.f};
      // worst-case screen error of this point: sum_d |2 u_d v_jd| 2^-10 <= 2^-9 ||u|| max_j ||v_j||
      const TcWindow win = tc_window(0.001953125f * sqrtf(static_cast<float>(nu)) * vmax);

      // ---- epilogue: screened sum over the accumulator tiles, TMEM loads one chunk ahead of the arithmetic ----
      bool ok = true;
      const int build_at = ntiles_n > 1 ? 1 : 0;  // the next iteration's A tile is built in the shadow of this sweep
#pragma unroll kTcTileUnroll
      for (int tn = 0; tn < ntiles_n && ok; ++tn, ++itn) {
        const uint32_t b = itn & 1u, bph = (itn >> 1) & 1u;
        if (!tc_wait(&bars.tfull[b], bph, abort_flag)) {
          ok = false;
          break;
        }
        tc_fence_after();
        const uint32_t taddr = tmem_base + (rt * 2 + b) * kTcN + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const int64_t jbase = static_cast<int64_t>(tn) * kTcN;
        uint32_t r0[32], r1[32];
        tc_ld32_issue(taddr, r0);
        tc_ld_wait();
#pragma unroll kTcEpiUnroll
        for (int c0 = 0; c0 < kTcN; c0 += 64) {
          tc_ld32_issue(taddr + c0 + 32, r1);
          const float m0 = m;
          if (!(dbg & 1)) tc_fold32(r0, tc_chunk_max(r0), c, win, S, m, exact, jbase + c0, K, h, m_hint);
          if (m != m0) c = static_cast<float>(nu + static_cast<double>(m));
          tc_ld_wait();
          if (c0 + 64 < kTcN) tc_ld32_issue(taddr + c0 + 64, r0);
          const float m1 = m;
          if (!(dbg & 1)) tc_fold32(r1, tc_chunk_max(r1), c, win, S, m, exact, jbase + c0 + 32, K, h, m_hint);
          if (dbg & 1) S[0] += __uint_as_float(r0[3]) + __uint_as_float(r1[5]);
          if (m != m1) c = static_cast<float>(nu + static_cast<double>(m));
          tc_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(&bars.tempty[b]);
        // every MMA of the PREVIOUS iteration (which read the other A buffer) completed before this iteration's first
        // accumulator was signalled, so that buffer is free
        if (tn == build_at && has_next) nu_next = build_a(xn, (itm + 1u) & 1u);
      }
      if (ok && p < n) out[p] = (m + log2f((S[0] + S[1]) + (S[2] + S[3]))) * 0.69314718055994530942f + uniform_offset;
#pragma unroll
      for (int d = 0; d < DP; ++d) xr[d] = xn[d];
      nu = nu_next;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTcEpiThreads / 32 + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(kTcAccCols))
                 : "memory");
  }
  if (threadIdx.x == 0 && bars.abort_flag) atomicAdd(err, 1);
}

static int tc_debug_mode() {
  const char* e = getenv("AISB_TC_DEBUG");  // bit 0: skip the epilogue arithmetic; bit 1: issue one MMA per tile (timing experiments)
  return e ? atoi(e) : 0;
}

template <int DP>
static int launch_tc(const float* x, int64_t n, int D, const aisb_packed_mixture* mix, const int32_t* hint,
                     const float* image, float* out, int* err, cudaStream_t st) {
  const size_t smem = tc_smem_bytes(DP);
  auto kern = tc_logq_kernel<DP>;
  AISB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t ntiles_m = (n + kTcRows * kTcRowTiles - 1) / (kTcRows * kTcRowTiles);
  const int ntiles_n = static_cast<int>((mix->K + kTcN - 1) / kTcN);
  const int64_t grid = ntiles_m < sm_count() ? ntiles_m : sm_count();
  kern<<<static_cast<unsigned>(grid), kTcThreads, smem, st>>>(x, n, D, mix->iso_scale, mix->uniform_offset, mix->d_rows,
                                                             mix->K, hint, image, ntiles_n, out, err, tc_debug_mode());
  AISB_LAUNCH_CHECK("tc_logq_kernel");
  return AISB_OK;
}


}  // namespace aisb

using namespace aisb;

extern "C" int64_t aisb_tc_image_floats(int64_t K, int32_t D) {
  const int DP = aisb_padded_dims(D);
  if (K < 1 || DP < 8 || DP > 32) return 0;
  return (K + kTcN - 1) / kTcN * static_cast<int64_t>(tc_tile_floats(DP)) + kTcTrailer;
}

extern "C" int aisb_tc_pack_f32(const aisb_packed_mixture* mix, float* d_image, void* stream) {
  if (!mix || !mix->d_rows || !d_image || mix->cov_kind != AISB_COV_ISO_SHARED || mix->d_offsets ||
      aisb_tc_image_floats(mix->K, mix->D) == 0) {
    set_error("tc_pack: needs an ISO_SHARED mixture without offsets and 5 <= D <= 32");
    return AISB_ERR_UNSUPPORTED;
  }
  const int DP = aisb_padded_dims(mix->D);
  const int64_t KT = (mix->K + kTcN - 1) / kTcN * kTcN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AISB_CUDA_CHECK(cudaMemsetAsync(d_image + (KT / kTcN) * static_cast<int64_t>(tc_tile_floats(DP)), 0,
                                  sizeof(float) * kTcTrailer, st));
  tc_pack_kernel<<<static_cast<unsigned>((KT + 127) / 128), 128, 0, st>>>(mix->d_rows, mix->K, mix->D, DP, d_image);
  AISB_LAUNCH_CHECK("tc_pack_kernel");
  return AISB_OK;
}

namespace aisb {
int tc_logq_launch(const float* d_x, int64_t n, const aisb_packed_mixture* mix, const int32_t* d_hint,
                   const float* d_image, float* d_out, int32_t* d_err, cudaStream_t st) {
  if (!mix || mix->cov_kind != AISB_COV_ISO_SHARED || mix->d_offsets || aisb_tc_image_floats(mix->K, mix->D) == 0) {
    set_error("tc_logq: needs an ISO_SHARED mixture without offsets and 5 <= D <= 32");
    return AISB_ERR_UNSUPPORTED;
  }
  if (n < 0 || !d_image || !d_err || (n > 0 && (!d_x || !d_out || !d_hint))) {
    set_error("tc_logq: null pointer or negative n");
    return AISB_ERR_INVALID;
  }
  if (n == 0) return AISB_OK;
  switch (aisb_padded_dims(mix->D)) {
    case 8: return launch_tc<8>(d_x, n, mix->D, mix, d_hint, d_image, d_out, d_err, st);
    case 16: return launch_tc<16>(d_x, n, mix->D, mix, d_hint, d_image, d_out, d_err, st);
    case 32: return launch_tc<32>(d_x, n, mix->D, mix, d_hint, d_image, d_out, d_err, st);
    default: break;
  }
  set_error("tc_logq: unsupported D=%d", mix->D);
  return AISB_ERR_UNSUPPORTED;
}
}  // namespace aisb

extern "C" int aisb_tc_logq_f32(const float* d_x, int64_t n, const aisb_packed_mixture* mix, const int32_t* d_hint,
                                const float* d_image, float* d_out_logq, int32_t* d_err, void* stream) {
  return tc_logq_launch(d_x, n, mix, d_hint, d_image, d_out_logq, d_err, static_cast<cudaStream_t>(stream));
}
