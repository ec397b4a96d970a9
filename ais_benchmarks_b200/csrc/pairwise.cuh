// Pairwise (points x Gaussian components) online log-sum-exp: the arithmetic core of
// CMixtureModel.log_prob (reference distributions/mixture/CMixtureModel.py:43-57), the
// KDE evaluation (sampling_methods/base.py:156-179) and the DM weights (dm_ais.py:72-76).
//
// Work decomposition
//   * a CTA of 128 threads owns a tile of 128*E points; every thread keeps its E points in
//     registers for the whole kernel (E*DP floats; E = 4 up to DP = 32, 2 at DP = 64),
//   * component parameters stream through shared memory in chunks of J components, staged by
//     1-D TMA bulk copies (cp.async.bulk -> UBLKCP) into a 2-deep mbarrier ring,
//   * all lanes of a warp read the SAME component (LDS broadcast) and evaluate it against
//     their own points; JJ components are evaluated per online-LSE step so that the running
//     max / rescale costs one extra MUFU per JJ pairs,
//   * everything is carried in the log2 domain: t = c_k - ||A_k x + b_k||^2, accumulated as an
//     FFMA chain seeded with c_k, then S += ex2(t - m),
//   * ISO/DIAG kinds use the sm_100 packed FP32 instructions (FFMA2): two POINTS share one 64-bit
//     register pair per dimension and the component parameter enters as a scalar-broadcast operand
//     (FFMA2 Rd, Ra.F32x2, Rb.F32, Rc.F32), so one FFMA2 advances two (point, component) pairs:
//     half the issue slots for the same FMA-pipe work, which leaves issue bandwidth for the
//     MUFU / FMNMX / LDS instructions,
//   * the online-LSE update of group g runs while group g+1 is being evaluated (software pipeline:
//     the pending group's t values stay in registers), so MUFU.EX2 work interleaves with FFMA2 work
//     instead of arriving in bursts.
//
// Per (point, component) pair, ISO/DIAG kinds: DP FFMA2 issue slots (2*DP FMA-pipe lane-cycles)
// + 1 FMNMX(3) + 1 FADD + 1 MUFU.EX2 + 1 FADD + (2 MUFU-related)/JJ.
#pragma once
#include "common.cuh"

namespace aisb {

// ---------------------------------------------------------------------------------------------
// Compile-time tiling traits
// ---------------------------------------------------------------------------------------------
#ifndef AISB_E_SMALL
#define AISB_E_SMALL 4  // points per thread for DP <= 8
#endif
// 4 up to DP = 32 (two point pairs per thread: twice the independent FFMA2 chains and half the LDS per FFMA2 of one
// pair; 210-226 registers, no spills; measured at DP = 32 against 2 points per thread: DIAG K=64 +8 %, ISO K=1024 +2 %)
__host__ __device__ constexpr int points_per_thread(int DP) { return DP <= 8 ? AISB_E_SMALL : (DP <= 32 ? 4 : 2); }

#ifndef AISB_JJ_SMALL
#define AISB_JJ_SMALL 8  // components per online-LSE step for DP <= 8 (ISO/DIAG); 8 measured 1.6 % faster than 4
#endif

__host__ __device__ constexpr int comps_per_step(int kind, int DP) {
  return kind == AISB_COV_FULL ? (DP <= 8 ? 4 : (DP == 16 ? 2 : 1)) : (DP <= 8 ? AISB_JJ_SMALL : (DP <= 32 ? 4 : 2));
}

__host__ __device__ constexpr int pow2_floor(int v) {
  int p = 1;
  while (2 * p <= v) p *= 2;
  return p;
}

// components per shared-memory chunk: ~16 KB of rows, a power of two in [max(JJ,4), 256]
__host__ __device__ constexpr int chunk_comps(int kind, int DP) {
  const int jj = comps_per_step(kind, DP) < 4 ? 4 : comps_per_step(kind, DP);
  const int fit = pow2_floor(16384 / (row_floats(kind, DP) * 4) > 0 ? 16384 / (row_floats(kind, DP) * 4) : 1);
  return fit < jj ? jj : (fit > 256 ? 256 : fit);
}

__host__ __device__ constexpr int stage_floats(int kind, int DP, bool has_c) {
  return chunk_comps(kind, DP) * row_floats(kind, DP) + (has_c ? chunk_comps(kind, DP) : 0);
}

constexpr int kStages = 2;

__host__ __device__ constexpr size_t pass_smem_bytes(int kind, int DP, bool has_c) {
  return size_t(kStages) * stage_floats(kind, DP, has_c) * sizeof(float);
}

#ifdef __CUDACC__

// ---------------------------------------------------------------------------------------------
// Shared-memory vector loads (all lanes read the same address -> broadcast)
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void lds_vec(const float* __restrict__ p, float (&v)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 f = reinterpret_cast<const float4*>(p)[q];
      v[4 * q + 0] = f.x;
      v[4 * q + 1] = f.y;
      v[4 * q + 2] = f.z;
      v[4 * q + 3] = f.w;
    }
  } else if constexpr (N == 2) {
    const float2 f = *reinterpret_cast<const float2*>(p);
    v[0] = f.x;
    v[1] = f.y;
  } else {
#pragma unroll
    for (int q = 0; q < N; ++q) v[q] = p[q];
  }
}

// ---------------------------------------------------------------------------------------------
// One component against E register-resident points: t[e] = c - ||A x_e + b||^2   (log2 domain)
// ---------------------------------------------------------------------------------------------
template <int KIND, int DP, int E>
__device__ __forceinline__ void comp_eval(const float* __restrict__ row, float c, float a_iso,
                                          const float (&x)[E][DP], float (&t)[E]) {
  // scalar evaluation (used by the moment kernel)
  if constexpr (KIND == AISB_COV_ISO_SHARED) {
    float b[DP];
    lds_vec<DP>(row, b);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float acc = c;
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const float dl = fmaf(x[e][d], a_iso, b[d]);
        acc = fmaf(-dl, dl, acc);
      }
      t[e] = acc;
    }
  } else if constexpr (KIND == AISB_COV_DIAG) {
    float a[DP], b[DP];
    lds_vec<DP>(row, a);
    lds_vec<DP>(row + DP, b);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float acc = c;
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const float dl = fmaf(x[e][d], a[d], b[d]);
        acc = fmaf(-dl, dl, acc);
      }
      t[e] = acc;
    }
  } else {  // FULL: z = A x + b with A lower triangular (row-major packed), t = c - ||z||^2
    constexpr int TRI = DP * (DP + 1) / 2;
    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = c;
    int idx = 0;
#pragma unroll
    for (int r = 0; r < DP; ++r) {
      float z[E];
      const float br = row[TRI + r];
#pragma unroll
      for (int e = 0; e < E; ++e) z[e] = br;
#pragma unroll
      for (int cc = 0; cc <= r; ++cc) {
        const float a = row[idx++];
#pragma unroll
        for (int e = 0; e < E; ++e) z[e] = fmaf(a, x[e][cc], z[e]);
      }
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = fmaf(-z[e], z[e], acc[e]);
    }
#pragma unroll
    for (int e = 0; e < E; ++e) t[e] = acc[e];
  }
}

// ---------------------------------------------------------------------------------------------
// Online LSE over `cnt` staged components (cnt is a multiple of JJ)
// ---------------------------------------------------------------------------------------------
template <int KIND, int DP, int E, int JJ, bool HAS_C>
__device__ __forceinline__ void accumulate_chunk(const float* __restrict__ srows, const float* __restrict__ soffs,
                                                 int cnt, float a_iso, const float (&x)[E][DP], float (&m)[E],
                                                 float (&S)[E]) {
  constexpr int ROW = row_floats(KIND, DP);
#pragma unroll 1
  for (int j0 = 0; j0 < cnt; j0 += JJ) {
    float t[JJ][E];
    float c[JJ];
    if constexpr (HAS_C) {
      lds_vec<JJ>(soffs + j0, c);
    } else {
#pragma unroll
      for (int jj = 0; jj < JJ; ++jj) c[jj] = 0.f;
    }
#pragma unroll
    for (int jj = 0; jj < JJ; ++jj) comp_eval<KIND, DP, E>(srows + (j0 + jj) * ROW, c[jj], a_iso, x, t[jj]);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      float gm = t[0][e];
#pragma unroll
      for (int jj = 1; jj < JJ; ++jj) gm = fmaxf(gm, t[jj][e]);
      const float mn = fmaxf(m[e], gm);
      float s = S[e] * ex2(m[e] - mn);
#pragma unroll
      for (int jj = 0; jj < JJ; ++jj) s += ex2(t[jj][e] - mn);
      S[e] = s;
      m[e] = mn;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Packed (FFMA2) path for ISO / DIAG kinds. Points are held as pairs: x2[pp][d] = (x[2pp][d], x[2pp+1][d]); a
// component parameter is a scalar that the FFMA2 broadcasts to both halves (make_float2(b, b) compiles to the .F32
// operand form, no register pair is built).
// Row layout (include/ais_b200.h): ISO  b[0..DP)            DIAG  a[0..DP) then b[0..DP)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }
__device__ __forceinline__ float2 bc2(float v) { return make_float2(v, v); }

template <int KIND, int DP, int EP>
__device__ __forceinline__ void comp_eval2(const float* __restrict__ row, const float2 (&seed)[EP], float a_iso,
                                           const float2 (&x2)[EP][DP], float2 (&t2)[EP]) {
  float b[DP];
  if constexpr (KIND == AISB_COV_ISO_SHARED) {
    lds_vec<DP>(row, b);
#pragma unroll
    for (int pp = 0; pp < EP; ++pp) {
      float2 acc = seed[pp];
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const float2 dl = __ffma2_rn(x2[pp][d], bc2(a_iso), bc2(b[d]));
        acc = __ffma2_rn(neg2(dl), dl, acc);
      }
      t2[pp] = acc;
    }
  } else {
    float a[DP];
    lds_vec<DP>(row, a);
    lds_vec<DP>(row + DP, b);
#pragma unroll
    for (int pp = 0; pp < EP; ++pp) {
      float2 acc = seed[pp];
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const float2 dl = __ffma2_rn(x2[pp][d], bc2(a[d]), bc2(b[d]));
        acc = __ffma2_rn(neg2(dl), dl, acc);
      }
      t2[pp] = acc;
    }
  }
}

template <int KIND, int DP, int EP, int JJ, bool HAS_C>
__device__ __forceinline__ void eval_group2(const float* __restrict__ srows, const float* __restrict__ soffs, int j0,
                                            float a_iso, const float2 (&x2)[EP][DP], const float2 (&seed0)[EP],
                                            float2 (&t)[JJ][EP]) {
  constexpr int ROW = row_floats(KIND, DP);
  if constexpr (HAS_C) {
    float c[JJ];
    lds_vec<JJ>(soffs + j0, c);
#pragma unroll
    for (int jj = 0; jj < JJ; ++jj) {
      float2 seed[EP];
#pragma unroll
      for (int pp = 0; pp < EP; ++pp) seed[pp] = make_float2(c[jj], c[jj]);
      comp_eval2<KIND, DP, EP>(srows + (j0 + jj) * ROW, seed, a_iso, x2, t[jj]);
    }
  } else {
#pragma unroll
    for (int jj = 0; jj < JJ; ++jj) comp_eval2<KIND, DP, EP>(srows + (j0 + jj) * ROW, seed0, a_iso, x2, t[jj]);
  }
}

// fold one evaluated group into the running (m, S) of each point pair. Every FP32 operation here is PACKED
// (FADD2 / FMUL2): a scalar FADD between FFMA2s leaves half of the FMA datapath idle for its two cycles.
//
// Warp-uniform skip. A group whose largest term is 2^-margin below the running maximum of EVERY point of the warp
// cannot change a float32 sum (margin = 28 + log2 K: at most K such terms, together below 2^-28 of the sum, i.e. 1/16
// ulp), so its fold -- 2 MUFU, 2 FADD2 per pair, the whole LSE overhead -- is skipped after the FMNMX group maximum
// (ALU pipe) and one vote. The running maximum is a lower bound of the final one, so the test is conservative. It fires
// when the points of a warp and the components of a group are spatially coherent: large KDEs are evaluated in Morton
// order on both sides for exactly this reason (_device.py). On incoherent data the vote and its register pressure cost
// 2-5 %, so the test is a template flag (SKIP) that only the "ordered" entry point of the library switches on.
template <int EP, int JJ, bool SKIP = false>
__device__ __forceinline__ void lse_update2(const float2 (&t)[JJ][EP], float2 (&m2)[EP], float2 (&S2)[EP],
                                            float margin = INFINITY) {
#pragma unroll
  for (int pp = 0; pp < EP; ++pp) {
    float g0 = t[0][pp].x, g1 = t[0][pp].y;
#pragma unroll
    for (int jj = 1; jj < JJ; ++jj) {
      g0 = fmaxf(g0, t[jj][pp].x);
      g1 = fmaxf(g1, t[jj][pp].y);
    }
    if constexpr (SKIP) {
      // NaN maxima compare false and therefore take the full path
      const bool negligible = (g0 <= m2[pp].x - margin) && (g1 <= m2[pp].y - margin);
      if (__all_sync(0xffffffffu, negligible)) continue;
    }
    const float2 mn = make_float2(fmaxf(m2[pp].x, g0), fmaxf(m2[pp].y, g1));
    const float2 nmn = neg2(mn);
    const float2 dm = __fadd2_rn(m2[pp], nmn);
    float2 acc = __fmul2_rn(S2[pp], make_float2(ex2(dm.x), ex2(dm.y)));
#pragma unroll
    for (int jj = 0; jj < JJ; ++jj) {
      const float2 a = __fadd2_rn(t[jj][pp], nmn);
      acc = __fadd2_rn(acc, make_float2(ex2(a.x), ex2(a.y)));
    }
    S2[pp] = acc;
    m2[pp] = mn;
  }
}

// Software-pipelined chunk: on entry tP holds the previously evaluated, not yet folded group (all -inf at the very
// start: folding it is a no-op); on exit it holds this chunk's last group. cnt / JJ is even (cnt is a multiple of 32).
template <int KIND, int DP, int EP, int JJ, bool HAS_C, bool SKIP = false>
__device__ __forceinline__ void accumulate_chunk2(const float* __restrict__ srows, const float* __restrict__ soffs,
                                                  int cnt, float a_iso, const float2 (&x2)[EP][DP],
                                                  float2 (&tP)[JJ][EP], float2 (&m)[EP], float2 (&S)[EP],
                                                  float margin) {
  float2 zero[EP];
#pragma unroll
  for (int pp = 0; pp < EP; ++pp) zero[pp] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int j0 = 0; j0 < cnt; j0 += 2 * JJ) {
    float2 tQ[JJ][EP];
    eval_group2<KIND, DP, EP, JJ, HAS_C>(srows, soffs, j0, a_iso, x2, zero, tQ);
    lse_update2<EP, JJ, SKIP>(tP, m, S, margin);
    eval_group2<KIND, DP, EP, JJ, HAS_C>(srows, soffs, j0 + JJ, a_iso, x2, zero, tP);
    lse_update2<EP, JJ, SKIP>(tQ, m, S, margin);
  }
}

// ---------------------------------------------------------------------------------------------
// Stream components [k_begin, k_end) of a packed mixture through the smem ring, calling
// fn(srows, soffs, cnt, k0) on every staged chunk (cnt components starting at global index k0).
// `it` counts chunks consumed so far by this CTA (carries the mbarrier phase across passes).
// k_begin/k_end are multiples of AISB_COMP_ALIGN. All threads of the CTA must call this.
// ---------------------------------------------------------------------------------------------
// `stride` > 1 visits every stride-th chunk only (chunk c of this CTA starts at k_begin + c * stride * J): an
// INTERLEAVED split of the component range over CTAs, so that each of them sees a uniform sample of a spatially ordered
// mixture (its running maximum then approaches the global one as fast as an unsplit sweep's would).
// J: components per staged chunk (default chunk_comps: ~16 KB of rows; single-warp CTAs use smaller chunks).
template <int KIND, int DP, bool HAS_C, int J = chunk_comps(KIND, DP), typename Fn>
__device__ __forceinline__ void stream_components(const float* __restrict__ g_rows, const float* __restrict__ g_offs,
                                                  int64_t k_begin, int64_t k_end, float* smem, uint64_t* bars,
                                                  uint32_t& it, Fn&& fn, int stride = 1) {
  constexpr int ROW = row_floats(KIND, DP);
  constexpr int STAGE = J * ROW + (HAS_C ? J : 0);
  const int64_t step = static_cast<int64_t>(J) * stride;
  const int nchunks = k_end > k_begin ? static_cast<int>((k_end - k_begin + step - 1) / step) : 0;
  const uint32_t it0 = it;

  auto issue = [&](int c) {
    const int64_t k0 = k_begin + static_cast<int64_t>(c) * step;
    const int64_t rem = k_end - k0;
    const int cnt = rem < J ? static_cast<int>(rem) : J;
    const uint32_t stage = (it0 + c) % kStages;
    float* dst = smem + stage * STAGE;
    const uint32_t bytes_rows = static_cast<uint32_t>(cnt) * ROW * 4u;
    const uint32_t bytes_offs = HAS_C ? static_cast<uint32_t>(cnt) * 4u : 0u;
    mbar_arrive_expect_tx(&bars[stage], bytes_rows + bytes_offs);
    bulk_g2s(dst, g_rows + k0 * ROW, bytes_rows, &bars[stage]);
    if constexpr (HAS_C) bulk_g2s(dst + J * ROW, g_offs + k0, bytes_offs, &bars[stage]);
  };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s)
      if (s < nchunks) issue(s);
  }
#pragma unroll 1
  for (int c = 0; c < nchunks; ++c) {
    const uint32_t stage = (it0 + c) % kStages;
    const uint32_t parity = ((it0 + c) / kStages) & 1u;
    mbar_wait(&bars[stage], parity);
    const int64_t k0 = k_begin + static_cast<int64_t>(c) * step;
    const int64_t rem = k_end - k0;
    const int cnt = rem < J ? static_cast<int>(rem) : J;
    const float* s = smem + stage * STAGE;
    fn(s, s + J * ROW, cnt, k0);
    __syncthreads();  // every thread is done reading this stage
    if (threadIdx.x == 0 && c + kStages < nchunks) issue(c + kStages);
  }
  it = it0 + nchunks;
}

// ---------------------------------------------------------------------------------------------
// Two-level summation. A float32 running sum over 1e6 comparable terms carries a random-walk error of
// ~sqrt(1e6) * 2^-24 = 6e-5 (seen as 2e-5 on log-densities in dense regions at 1e6 centres). The per-thread sum S
// therefore only ever spans a few staged chunks (~1024 components); it is then merged, with the
// log-sum-exp rescale rule, into a float64 accumulator that lives in shared memory (one slot per thread and point,
// touched once per chunk, no registers) and is reset. Cost: two MUFU and a DFMA per point per chunk.
// ---------------------------------------------------------------------------------------------
template <int T>
struct OuterAccT {
  float M[4][T];   // reference exponent of the outer sum (log2 domain)
  double S[4][T];  // outer sum, relative to M
};
using OuterAcc = OuterAccT<kThreads>;

template <int E, class OA>
__device__ __forceinline__ void outer_init(OA& o) {
  static_assert(E <= 4, "OuterAcc holds at most 4 points per thread");
#pragma unroll
  for (int e = 0; e < E; ++e) {
    o.M[e][threadIdx.x] = kNegBig;
    o.S[e][threadIdx.x] = 0.0;
  }
}

// merge the inner (m_e, S_e) into the outer accumulator and restart the inner sum (its reference m_e is kept)
template <int E, class OA>
__device__ __forceinline__ void outer_fold(OA& o, const float (&m)[E], float (&S)[E]) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (S[e] != 0.f) {  // an empty inner sum must not move the outer reference (its own reference is not meaningful yet)
      const float Mo = o.M[e][threadIdx.x];
      const float Mn = fmaxf(Mo, m[e]);
      o.S[e][threadIdx.x] = o.S[e][threadIdx.x] * static_cast<double>(ex2(Mo - Mn)) +
                            static_cast<double>(S[e]) * static_cast<double>(ex2(m[e] - Mn));
      o.M[e][threadIdx.x] = Mn;
      S[e] = 0.f;
    }
  }
}

template <int E, class OA>
__device__ __forceinline__ void outer_finish(const OA& o, float (&m)[E], float (&S)[E]) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    m[e] = o.M[e][threadIdx.x];
    S[e] = static_cast<float>(o.S[e][threadIdx.x]);
  }
}

// Fold components [k_begin, k_end) into (m, S) of the E register-resident points (results replace m, S).
template <int KIND, int DP, int E, bool HAS_C, bool SKIP = false>
__device__ __forceinline__ void mixture_pass(const float* __restrict__ g_rows, const float* __restrict__ g_offs,
                                             int64_t k_begin, int64_t k_end, float a_iso, float* smem, uint64_t* bars,
                                             OuterAcc& outer, uint32_t& it, const float (&x)[E][DP], float (&m)[E],
                                             float (&S)[E], float margin = INFINITY) {
  constexpr int JJ = comps_per_step(KIND, DP);
  // the inner float32 sum spans at most kFoldEvery chunks (~1024 components: random-walk error sqrt(1024) 2^-24 = 2e-6)
  constexpr int kFoldEvery = chunk_comps(KIND, DP) >= 256 ? 4 : (chunk_comps(KIND, DP) >= 64 ? 16 : 32);
  int nchunk = 0;
  outer_init<E>(outer);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    m[e] = kNegBig;
    S[e] = 0.f;
  }
  if constexpr (KIND == AISB_COV_FULL) {
    stream_components<KIND, DP, HAS_C>(g_rows, g_offs, k_begin, k_end, smem, bars, it,
                                       [&](const float* srows, const float* soffs, int cnt, int64_t) {
                                         accumulate_chunk<KIND, DP, E, JJ, HAS_C>(srows, soffs, cnt, a_iso, x, m, S);
                                         if ((++nchunk % kFoldEvery) == 0) outer_fold<E>(outer, m, S);
                                       });
    outer_fold<E>(outer, m, S);
  } else {
    static_assert(E % 2 == 0, "packed path needs an even number of points per thread");
    constexpr int EP = E / 2;
    float2 x2[EP][DP];
#pragma unroll
    for (int pp = 0; pp < EP; ++pp)
#pragma unroll
      for (int d = 0; d < DP; ++d) x2[pp][d] = make_float2(x[2 * pp][d], x[2 * pp + 1][d]);
    float2 tP[JJ][EP], m2[EP], S2[EP];
#pragma unroll
    for (int jj = 0; jj < JJ; ++jj)
#pragma unroll
      for (int pp = 0; pp < EP; ++pp) tP[jj][pp] = make_float2(-INFINITY, -INFINITY);
    auto fold = [&]() {
#pragma unroll
      for (int pp = 0; pp < EP; ++pp) {
        m[2 * pp] = m2[pp].x;
        m[2 * pp + 1] = m2[pp].y;
        S[2 * pp] = S2[pp].x;
        S[2 * pp + 1] = S2[pp].y;
      }
      outer_fold<E>(outer, m, S);
#pragma unroll
      for (int pp = 0; pp < EP; ++pp) S2[pp] = make_float2(0.f, 0.f);
    };
#pragma unroll
    for (int pp = 0; pp < EP; ++pp) {
      m2[pp] = make_float2(kNegBig, kNegBig);
      S2[pp] = make_float2(0.f, 0.f);
    }
    stream_components<KIND, DP, HAS_C>(
        g_rows, g_offs, k_begin, k_end, smem, bars, it, [&](const float* srows, const float* soffs, int cnt, int64_t) {
          accumulate_chunk2<KIND, DP, EP, JJ, HAS_C, SKIP>(srows, soffs, cnt, a_iso, x2, tP, m2, S2, margin);
          if ((++nchunk % kFoldEvery) == 0) fold();
        });
    lse_update2<EP, JJ, SKIP>(tP, m2, S2, margin);
    fold();
  }
  outer_finish<E>(outer, m, S);
}


// ---------------------------------------------------------------------------------------------
// Screen-and-refine pass for a spatially ordered ISO_SHARED mixture without offsets (a large KDE evaluated in Z-order).
//
// The exact exponent of a pair, t = -||a x + b_j||^2, costs 2 DP FMA-lane-cycles (centred form: float32-safe). Its
// expanded form  ||a x + b_j||^2 = ||a x||^2 + 2 a (h_j + x . b_j),  h_j = ||b_j||^2 / (2 a),  costs DP: one FFMA2 chain
// seeded with h_j. The expanded value is NOT accurate enough to be summed (cancellation), but it is far more than
// accurate enough to decide that a pair cannot matter: a component whose exponent lies `marg` = 21 + log2 KP below the
// running maximum of a point contributes, together with all others like it, less than 2^-21 of the sum. So every pair
// is SCREENED with the cheap chain; a group of JJ components is evaluated exactly (and folded into the online
// log-sum-exp) only if it is not negligible for at least one of the 64 points a warp holds in one packed register set.
// The test per point is  s_ij = h_j + x_i . b_j >= thr_i,  thr_i = (marg + E_i - m_i - ||a x_i||^2) / (2 a), where m_i is
// the running maximum (a lower bound of the final one: conservative) and E_i = 2^-19 (||a x_i||^2 + max_j ||b_j||^2)
// bounds the rounding error of the float32 chain in exponent units -- data far from the origin costs time (the screen
// stops discriminating), never accuracy.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 min2(float2 a, float2 b) { return make_float2(fminf(a.x, b.x), fminf(a.y, b.y)); }
// three-input minimum (FMNMX3 on sm_100); NaN inputs are dropped like fminf drops them
__device__ __forceinline__ float min3f(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <int DP, int EP, int JJ, bool COUNT>
__device__ __forceinline__ void accumulate_chunk_screen(const float* __restrict__ srows, const float* __restrict__ shs,
                                                        int cnt, float a_iso, float ninv2a,
                                                        const float2 (&x2)[EP][DP], const float2 (&cthr)[EP],
                                                        float2 (&thr)[EP], float2 (&m2)[EP], float2 (&S2)[EP],
                                                        unsigned (&nrefined)) {
#pragma unroll 1
  for (int j0 = 0; j0 < cnt; j0 += JJ) {
    float h[JJ];
    lds_vec<JJ>(shs + j0, h);
    float2 smin[EP];
#pragma unroll
    for (int pp = 0; pp < EP; ++pp) smin[pp] = make_float2(INFINITY, INFINITY);
#pragma unroll
    for (int jj = 0; jj < JJ; ++jj) {
      float b[DP];
      lds_vec<DP>(srows + (j0 + jj) * DP, b);
#pragma unroll
      for (int pp = 0; pp < EP; ++pp) {
        float2 acc = bc2(h[jj]);
#pragma unroll
        for (int d = 0; d < DP; ++d) acc = __ffma2_rn(x2[pp][d], bc2(b[d]), acc);
        smin[pp] = min2(smin[pp], acc);
      }
    }
#pragma unroll
    for (int pp = 0; pp < EP; ++pp) {
      // NaN compares false: a NaN screen value takes the exact path
      const bool negligible = (smin[pp].x >= thr[pp].x) && (smin[pp].y >= thr[pp].y);
      if (__all_sync(0xffffffffu, negligible)) continue;
      if constexpr (COUNT) ++nrefined;
      // re-read the rows from shared memory here: keeping the group's JJ*DP parameters live across the branch (what
      // the compiler does on its own) costs more registers than the kernel can afford at 4 CTAs per SM
      asm volatile("" ::: "memory");
      float2 t[JJ];
#pragma unroll
      for (int jj = 0; jj < JJ; ++jj) {
        float b[DP];
        lds_vec<DP>(srows + (j0 + jj) * DP, b);
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          const float2 dl = __ffma2_rn(x2[pp][d], bc2(a_iso), bc2(b[d]));
          acc = __ffma2_rn(neg2(dl), dl, acc);
        }
        t[jj] = acc;
      }
      float g0 = t[0].x, g1 = t[0].y;
#pragma unroll
      for (int jj = 1; jj < JJ; ++jj) {
        g0 = fmaxf(g0, t[jj].x);
        g1 = fmaxf(g1, t[jj].y);
      }
      const float2 mn = make_float2(fmaxf(m2[pp].x, g0), fmaxf(m2[pp].y, g1));
      const float2 nmn = neg2(mn);
      const float2 dm = __fadd2_rn(m2[pp], nmn);
      float2 acc = __fmul2_rn(S2[pp], make_float2(ex2(dm.x), ex2(dm.y)));
#pragma unroll
      for (int jj = 0; jj < JJ; ++jj) {
        const float2 a = __fadd2_rn(t[jj], nmn);
        acc = __fadd2_rn(acc, make_float2(ex2(a.x), ex2(a.y)));
      }
      S2[pp] = acc;
      m2[pp] = mn;
      thr[pp] = make_float2(fmaf(mn.x, ninv2a, cthr[pp].x), fmaf(mn.y, ninv2a, cthr[pp].y));
    }
  }
}

// Fold the components {k_begin + c * stride * J .. } (every stride-th chunk from k_begin on, below k_end) into (m, S).
// g_hs: the screen vector h_j [KP] followed by its trailer (trailer[0] = max_j ||b_j||^2), see aisb_kde_screen_pack_f32.
// COUNT: nscreened / nrefined count this thread's (point-pair, group) units (identical for all lanes of a warp).
template <int DP, int E, int J, bool COUNT = false, class OA>
__device__ __forceinline__ void mixture_pass_screen(const float* __restrict__ g_rows, const float* __restrict__ g_hs,
                                                    int64_t KP, int64_t k_begin, int stride, float a_iso, float marg,
                                                    float* smem, uint64_t* bars, OA& outer, uint32_t& it,
                                                    const float (&x)[E][DP], float (&m)[E], float (&S)[E],
                                                    unsigned* nscreened = nullptr, unsigned* nrefined_out = nullptr) {
  constexpr int KIND = AISB_COV_ISO_SHARED;
  constexpr int JJ = comps_per_step(KIND, DP);
  constexpr int kFoldEvery = 1024 / J > 0 ? 1024 / J : 1;   // the float32 inner sum spans ~1024 components
  static_assert(E % 2 == 0, "packed path needs an even number of points per thread");
  static_assert(J % JJ == 0 && J % AISB_COMP_ALIGN == 0, "chunk size");
  constexpr int EP = E / 2;
  int nchunk = 0;
  unsigned nrefined = 0, nunits = 0;
  outer_init<E>(outer);
  const float bmax2 = __ldg(g_hs + KP);
  const float inv2a = 0.5f / a_iso;
  float2 x2[EP][DP], cthr[EP], thr[EP], m2[EP], S2[EP];
#pragma unroll
  for (int pp = 0; pp < EP; ++pp) {
    float u0 = 0.f, u1 = 0.f;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      x2[pp][d] = make_float2(x[2 * pp][d], x[2 * pp + 1][d]);
      u0 = fmaf(x[2 * pp][d], x[2 * pp][d], u0);
      u1 = fmaf(x[2 * pp + 1][d], x[2 * pp + 1][d], u1);
    }
    u0 *= a_iso * a_iso;
    u1 *= a_iso * a_iso;
    const float e0 = 1.9073486e-6f * (u0 + bmax2), e1 = 1.9073486e-6f * (u1 + bmax2);  // 2^-19 (...)
    cthr[pp] = make_float2((marg + e0 - u0) * inv2a, (marg + e1 - u1) * inv2a);
    m2[pp] = make_float2(kNegBig, kNegBig);
    S2[pp] = make_float2(0.f, 0.f);
    thr[pp] = make_float2(fmaf(kNegBig, -inv2a, cthr[pp].x), fmaf(kNegBig, -inv2a, cthr[pp].y));
  }
  auto fold = [&]() {
#pragma unroll
    for (int pp = 0; pp < EP; ++pp) {
      m[2 * pp] = m2[pp].x;
      m[2 * pp + 1] = m2[pp].y;
      S[2 * pp] = S2[pp].x;
      S[2 * pp + 1] = S2[pp].y;
    }
    outer_fold<E>(outer, m, S);
#pragma unroll
    for (int pp = 0; pp < EP; ++pp) S2[pp] = make_float2(0.f, 0.f);
  };
  stream_components<KIND, DP, true, J>(
      g_rows, g_hs, k_begin, KP, smem, bars, it,
      [&](const float* srows, const float* shs, int cnt, int64_t) {
        accumulate_chunk_screen<DP, EP, JJ, COUNT>(srows, shs, cnt, a_iso, -inv2a, x2, cthr, thr, m2, S2, nrefined);
        if constexpr (COUNT) nunits += static_cast<unsigned>(cnt / JJ) * EP;
        if ((++nchunk % kFoldEvery) == 0) fold();
      },
      stride);
  fold();
  outer_finish<E>(outer, m, S);
  if constexpr (COUNT) {
    *nscreened = nunits;
    *nrefined_out = nrefined;
  }
}

// ---------------------------------------------------------------------------------------------
// Fast pass for a spatially ordered ISO_SHARED mixture without offsets: RECENTRED EXPANDED FORM, no exact refinement.
//
// With both sides shifted by the centroid c of the warp's 32*E points (x' = x - c, b'_j = b_j + a c: distances do not
// change),   ||a x + b_j||^2 = ||a x'||^2 + 2 a s_ij,   s_ij = h'_j + x'_i . b'_j,   h'_j = ||b'_j||^2 / (2 a).
// The chain for s costs DP FMAs per pair -- half of the centred form -- and after the shift every magnitude in it is a
// LOCAL distance in kernel-bandwidth units (warp extent, distance to the components that matter) instead of a distance
// from the origin, so its float32 rounding error, ~2.4 * 2^-24 (||a x'||^2 + |t_max| + 1) on the log density (measured,
// scripts/kde_fast_error_model.py), stays far below the 1e-5 bar for warps whose points are close together in
// bandwidth units. The kernel estimates that error per point from quantities it has anyway and LISTS the points whose
// estimate exceeds 3e-6 max(1, |log p|); the caller re-evaluates exactly those with the centred kernel
// (logpdf_subset_kernel). On BASELINE configs[3] that is < 0.1 % of the points.
//
// The shift is applied to each staged chunk IN PLACE by the warp itself (each lane rewrites J/32 rows and their h'),
// ~2 % of the chunk's work. The online log-sum-exp runs on s directly: reference = running MINIMUM of s (the maximum
// of the exponent -||a x'||^2 - 2 a s), terms ex2(-2a (s - s_min)); a group of JJ components is folded only if, for at
// least one of the 64 points of a packed register set, some s is below s_min + marg / (2a).
// ---------------------------------------------------------------------------------------------
constexpr float kPosBig = 3.0e38f;

#ifndef AISB_FAST_UNROLL
#define AISB_FAST_UNROLL 2
#endif
// groups of JJ components per trip of the fast kernel's inner loop: 2 halves its address arithmetic, compare and branch
// (365.6 -> 356.4 ms at configs[3]); 4 is no faster and spills
constexpr int kFastUnroll = AISB_FAST_UNROLL;

template <int DP, int EP, int JJ, bool COUNT>
__device__ __forceinline__ void accumulate_chunk_fast(const float* __restrict__ srows, const float* __restrict__ sh,
                                                      int cnt, float n2a, float two_a, float dthr,
                                                      const float2 (&x2)[EP][DP], float2 (&ms)[EP], float2 (&cr)[EP],
                                                      float2 (&S2)[EP], float2 (&thr)[EP], const float2 (&gbest)[EP],
                                                      unsigned (&nfolded)) {
#pragma unroll kFastUnroll
  for (int j0 = 0; j0 < cnt; j0 += JJ) {
    float h[JJ];
    lds_vec<JJ>(sh + j0, h);
    float2 s[JJ][EP], smin[EP];
#pragma unroll
    for (int jj = 0; jj < JJ; ++jj) {
      float b[DP];
      lds_vec<DP>(srows + (j0 + jj) * DP, b);
#pragma unroll
      for (int pp = 0; pp < EP; ++pp) {
        float2 acc = bc2(h[jj]);
#pragma unroll
        for (int d = 0; d < DP; ++d) acc = __ffma2_rn(x2[pp][d], bc2(b[d]), acc);
        s[jj][pp] = acc;
      }
    }
    // group minimum with three-input FMNMX (every instruction that is not a packed FMA costs an issue slot the FMA
    // pipe could have used: profiles/r2_ncu_kde_fast_fullsize.txt)
#pragma unroll
    for (int pp = 0; pp < EP; ++pp) {
      float g0 = s[0][pp].x, g1 = s[0][pp].y;
#pragma unroll
      for (int jj = 1; jj + 1 < JJ; jj += 2) {
        g0 = min3f(g0, s[jj][pp].x, s[jj + 1][pp].x);
        g1 = min3f(g1, s[jj][pp].y, s[jj + 1][pp].y);
      }
      if constexpr (JJ % 2 == 0) {
        g0 = fminf(g0, s[JJ - 1][pp].x);
        g1 = fminf(g1, s[JJ - 1][pp].y);
      }
      smin[pp] = make_float2(g0, g1);
    }
#pragma unroll
    for (int pp = 0; pp < EP; ++pp) {
      // NaN compares false: NaN values are folded and propagate
      const bool negligible = (smin[pp].x >= thr[pp].x) && (smin[pp].y >= thr[pp].y);
      if (__all_sync(0xffffffffu, negligible)) continue;
      if constexpr (COUNT) ++nfolded;
      const float2 mn = min2(ms[pp], smin[pp]);
      // The sum is kept relative to the FLOAT c = fl(2a s_min), the very number the term arguments are formed with.
      // Rescaling by c_new - c_old is exactly 0 while the minimum stands (a factor derived from s_min itself, e.g.
      // -2a s_min_old + c_new, is the rounding residual of the product instead: 2^(3e-6) on every fold, compounding).
      const float2 c = make_float2(mn.x * two_a, mn.y * two_a);
      const float2 dm = __fadd2_rn(c, neg2(cr[pp]));                   // <= 0; -inf on the first fold (cr = +inf)
      float2 acc = __fmul2_rn(S2[pp], make_float2(ex2(dm.x), ex2(dm.y)));
#pragma unroll
      for (int jj = 0; jj < JJ; ++jj) {
        const float2 a = __ffma2_rn(s[jj][pp], bc2(n2a), c);           // -2a (s - s_min) <= 0
        acc = __fadd2_rn(acc, make_float2(ex2(a.x), ex2(a.y)));
      }
      S2[pp] = acc;
      ms[pp] = mn;
      cr[pp] = c;
      thr[pp] = make_float2(fminf(mn.x, gbest[pp].x) + dthr, fminf(mn.y, gbest[pp].y) + dthr);
    }
  }
}

// Sweep every stride-th chunk (J components) from k_begin on. Results: (m, S) in the exponent domain, as mixture_pass,
// and u2[e] = ||a x'_e||^2 (the caller's error estimate needs it). COUNT: nscreened / nfolded unit counters.
template <int DP, int E, int J, bool COUNT = false, class OA>
__device__ __forceinline__ void mixture_pass_fast(const float* __restrict__ g_rows, int64_t KP, int64_t k_begin,
                                                  int stride, float a_iso, float marg, float* smem, float* sh_h,
                                                  uint64_t* bars, OA& outer, uint32_t& it, const float (&x)[E][DP],
                                                  float (&m)[E], float (&S)[E], float (&u2)[E],
                                                  unsigned* nscreened = nullptr, unsigned* nfolded_out = nullptr,
                                                  int dbg = 0, uint32_t* __restrict__ gmin = nullptr,
                                                  const int64_t* pidx = nullptr, int64_t n = 0) {
  constexpr int KIND = AISB_COV_ISO_SHARED;
  constexpr int JJ = comps_per_step(KIND, DP);
  constexpr int kFoldEvery = 1024 / J > 0 ? 1024 / J : 1;   // the float32 inner sum spans ~1024 components
  static_assert(E % 2 == 0, "packed path needs an even number of points per thread");
  static_assert(J % JJ == 0 && J % 32 == 0, "chunk size");
  constexpr int EP = E / 2;
  const int lane = threadIdx.x & 31;
  int nchunk = 0;
  unsigned nfolded = 0, nunits = 0;
  outer_init<E>(outer);
  // centroid of the warp's 32*E points, and the shift a*c of the packed rows
  float ac[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) sum += x[e][d];
    sum = warp_sum(sum) * (1.f / (32 * E));
    if (!(fabsf(sum) < 1.0e30f)) sum = 0.f;     // NaN / inf coordinates: no shift (the result is NaN / -inf anyway)
    if (dbg & 2) sum = 0.f;
    ac[d] = sum;
  }
  const float two_a = 2.f * a_iso, n2a = -2.f * a_iso, inv2a = 0.5f / a_iso;
  const float dthr = (dbg & 1) ? INFINITY : marg * inv2a;
  // gbest: the smallest s any CTA working on the same points (another interleaved split of the components) has seen so
  // far, exchanged through `gmin` (one encoded float per point, initialised to 0xFFFFFFFF). It only tightens the SKIP
  // threshold -- every split then prunes almost as well as a single sweep would -- the sums keep their own reference.
  float2 x2[EP][DP], ms[EP], cr[EP], S2[EP], thr[EP], gbest[EP];
#pragma unroll
  for (int pp = 0; pp < EP; ++pp) {
    gbest[pp] = make_float2(kPosBig, kPosBig);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const float v0 = x[2 * pp][d] - ac[d], v1 = x[2 * pp + 1][d] - ac[d];
      x2[pp][d] = make_float2(v0, v1);
      q0 = fmaf(v0, v0, q0);
      q1 = fmaf(v1, v1, q1);
    }
    u2[2 * pp] = q0 * (a_iso * a_iso);
    u2[2 * pp + 1] = q1 * (a_iso * a_iso);
    ms[pp] = make_float2(kPosBig, kPosBig);
    cr[pp] = make_float2(INFINITY, INFINITY);
    S2[pp] = make_float2(0.f, 0.f);
    thr[pp] = make_float2(kPosBig, kPosBig);
  }
#pragma unroll
  for (int d = 0; d < DP; ++d) ac[d] *= a_iso;
  auto fold = [&]() {
#pragma unroll
    for (int pp = 0; pp < EP; ++pp) {
      // exponent-domain reference t_max = -||a x'||^2 - c: a fixed function of c, so every inner sum formed under the
      // same c meets the outer accumulator with the same reference (before the first fold c = +inf: the reference is
      // clamped, the inner sum is 0 and outer_fold ignores it)
      m[2 * pp] = fmaxf(-cr[pp].x - u2[2 * pp], kNegBig);
      m[2 * pp + 1] = fmaxf(-cr[pp].y - u2[2 * pp + 1], kNegBig);
      S[2 * pp] = S2[pp].x;
      S[2 * pp + 1] = S2[pp].y;
    }
    outer_fold<E>(outer, m, S);
#pragma unroll
    for (int pp = 0; pp < EP; ++pp) S2[pp] = make_float2(0.f, 0.f);
    if (gmin != nullptr) {   // publish this CTA's minima (fire-and-forget reductions)
#pragma unroll
      for (int pp = 0; pp < EP; ++pp) {
        if (pidx[2 * pp] < n) atomicMin(gmin + pidx[2 * pp], enc_ordered(ms[pp].x));
        if (pidx[2 * pp + 1] < n) atomicMin(gmin + pidx[2 * pp + 1], enc_ordered(ms[pp].y));
      }
    }
  };
  stream_components<KIND, DP, false, J>(
      g_rows, nullptr, k_begin, KP, smem, bars, it,
      [&](const float* srows, const float*, int cnt, int64_t) {
        // the other splits' minima: loads issued now, consumed after this chunk (their latency hides behind it)
        uint32_t gseen[E];
        if (gmin != nullptr) {
#pragma unroll
          for (int e = 0; e < E; ++e)
            gseen[e] = pidx[e] < n ? *reinterpret_cast<volatile uint32_t*>(gmin + pidx[e]) : 0xFFFFFFFFu;
        }
        // shift this chunk: b' = b + a c, h' = ||b'||^2 / (2a); every lane rewrites cnt / 32 rows
        float* rows = const_cast<float*>(srows);
#pragma unroll 1
        for (int j = lane; j < cnt; j += 32) {
          float b[DP];
          lds_vec<DP>(rows + j * DP, b);
          float q = 0.f;
#pragma unroll
          for (int d = 0; d < DP; ++d) {
            b[d] += ac[d];
            q = fmaf(b[d], b[d], q);
          }
          if constexpr (DP % 4 == 0) {
#pragma unroll
            for (int v = 0; v < DP / 4; ++v)
              reinterpret_cast<float4*>(rows + j * DP)[v] = make_float4(b[4 * v], b[4 * v + 1], b[4 * v + 2], b[4 * v + 3]);
          } else {
#pragma unroll
            for (int d = 0; d < DP; ++d) rows[j * DP + d] = b[d];
          }
          sh_h[j] = q * inv2a;
        }
        __syncwarp();
        accumulate_chunk_fast<DP, EP, JJ, COUNT>(srows, sh_h, cnt, n2a, two_a, dthr, x2, ms, cr, S2, thr, gbest,
                                                 nfolded);
        if constexpr (COUNT) nunits += static_cast<unsigned>(cnt / JJ) * EP;
        if (gmin != nullptr) {
#pragma unroll
          for (int pp = 0; pp < EP; ++pp) {
            // fminf drops NaN (the decoded initial pattern)
            gbest[pp] = make_float2(fminf(gbest[pp].x, dec_ordered(gseen[2 * pp])),
                                    fminf(gbest[pp].y, dec_ordered(gseen[2 * pp + 1])));
            thr[pp] = make_float2(fminf(ms[pp].x, gbest[pp].x) + dthr, fminf(ms[pp].y, gbest[pp].y) + dthr);
          }
        }
        if ((++nchunk % kFoldEvery) == 0) fold();
        // the next bulk copy (async proxy) overwrites rows this warp wrote through the generic proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      },
      stride);
  fold();
  outer_finish<E>(outer, m, S);
  if constexpr (COUNT) {
    *nscreened = nunits;
    *nfolded_out = nfolded;
  }
}

// ---------------------------------------------------------------------------------------------
// Point tile load: thread t owns points tile_base + e*128 + t (coalesced across the warp).
// Out-of-range points are clamped to n-1 (computed, never stored).
// ---------------------------------------------------------------------------------------------
template <int DP, int E>
__device__ __forceinline__ void load_points(const float* __restrict__ x, int64_t n, int D, int vec_ok,
                                            int64_t tile_base, float (&xr)[E][DP], int64_t (&pidx)[E]) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int64_t p = tile_base + e * kThreads + threadIdx.x;
    pidx[e] = p;
    const int64_t pc = p < n ? p : n - 1;
    bool done = false;
    if constexpr (DP >= 4) {
      if (D == DP && vec_ok) {
        const float4* r = reinterpret_cast<const float4*>(x + pc * DP);
#pragma unroll
        for (int q = 0; q < DP / 4; ++q) {
          const float4 f = __ldg(r + q);
          xr[e][4 * q + 0] = f.x;
          xr[e][4 * q + 1] = f.y;
          xr[e][4 * q + 2] = f.z;
          xr[e][4 * q + 3] = f.w;
        }
        done = true;
      }
    } else if constexpr (DP == 2) {
      if (D == 2 && vec_ok) {
        const float2 f = __ldg(reinterpret_cast<const float2*>(x + pc * 2));
        xr[e][0] = f.x;
        xr[e][1] = f.y;
        done = true;
      }
    }
    if (!done) {
      const float* r = x + pc * D;
#pragma unroll
      for (int d = 0; d < DP; ++d) xr[e][d] = d < D ? __ldg(r + d) : 0.f;
    }
  }
}

// Warp-contiguous variant for spatially ordered points: warp w owns the 32*E consecutive points
// tile_base + w*32*E + e*32 + lane, so a packed register set (two values of e) spans 64 neighbours of the ordering.
template <int DP, int E>
__device__ __forceinline__ void load_points_warp(const float* __restrict__ x, int64_t n, int64_t tile_base,
                                                 float (&xr)[E][DP], int64_t (&pidx)[E]) {
  static_assert(DP >= 4 || DP == 2, "vector loads");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // single-warp CTAs: warp == 0
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int64_t p = tile_base + warp * (32 * E) + e * 32 + lane;
    pidx[e] = p;
    const int64_t pc = p < n ? p : n - 1;
    if constexpr (DP >= 4) {
      const float4* r = reinterpret_cast<const float4*>(x + pc * DP);
#pragma unroll
      for (int q = 0; q < DP / 4; ++q) {
        const float4 f = __ldg(r + q);
        xr[e][4 * q + 0] = f.x;
        xr[e][4 * q + 1] = f.y;
        xr[e][4 * q + 2] = f.z;
        xr[e][4 * q + 3] = f.w;
      }
    } else {
      const float2 f = __ldg(reinterpret_cast<const float2*>(x + pc * 2));
      xr[e][0] = f.x;
      xr[e][1] = f.y;
    }
  }
}

__device__ __forceinline__ void ring_init(uint64_t* bars) {
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
}

// natural-log result from a log2-domain (m, S) pair; S == 0 -> -inf
__device__ __forceinline__ float lse_finish(float m, float S) { return (m + log2f(S)) * 0.69314718055994530942f; }

#endif  // __CUDACC__
}  // namespace aisb
