// Box-mixture ("haar") density, device-side sampling, and the FP32 / MUFU peak micro-benchmarks.
#include "common.cuh"

namespace aisb {

// ---------------------------------------------------------------------------------------------
// log sum_k w_k [x in box_k] / vol_k  over axis-aligned boxes.
// Replaces the per-sample, per-node Python scan of CTreePyramidSampling.prob/log_prob (haar kernel)
// (reference sampling_methods/tree_pyramid.py:587-603) and CMultivariateUniform.log_prob
// (distributions/parametric/CMultivariateUniform.py:38-47). One thread per point, boxes staged
// through shared memory in chunks.
// ---------------------------------------------------------------------------------------------
constexpr int kBoxChunk = 128;

template <int DP>
__global__ void __launch_bounds__(128) box_mixture_kernel(const float* __restrict__ x, int64_t n, int D,
                                                         const float* __restrict__ lo, const float* __restrict__ hi,
                                                         const float* __restrict__ lwv, int64_t K, int open_upper,
                                                         float* __restrict__ out) {
  __shared__ float s_lo[kBoxChunk * DP], s_hi[kBoxChunk * DP], s_w[kBoxChunk];
  const int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t pc = p < n ? p : n - 1;
  float xr[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) xr[d] = d < D ? x[pc * D + d] : 0.f;
  float m = kNegBig, S = 0.f;
  bool has_nan = false;
#pragma unroll
  for (int d = 0; d < DP; ++d) has_nan |= !(xr[d] == xr[d]);
  for (int64_t k0 = 0; k0 < K; k0 += kBoxChunk) {
    const int cnt = static_cast<int>(K - k0 < kBoxChunk ? K - k0 : kBoxChunk);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * DP; i += blockDim.x) {
      const int j = i / DP, d = i % DP;
      // padding dims: (-inf, +inf) always contains 0
      s_lo[i] = d < D ? lo[(k0 + j) * D + d] : -INFINITY;
      s_hi[i] = d < D ? hi[(k0 + j) * D + d] : INFINITY;
    }
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) s_w[i] = lwv[k0 + i];
    __syncthreads();
    for (int j = 0; j < cnt; ++j) {
      bool in = true;
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const float l = s_lo[j * DP + d], h = s_hi[j * DP + d];
        in = in && (l < xr[d]) && (open_upper ? xr[d] < h : xr[d] <= h);
      }
      const float t = s_w[j] * 1.4426950408889634f;
      if (in && t > -INFINITY) {
        const float mn = fmaxf(m, t);
        S = S * ex2(m - mn) + ex2(t - mn);
        m = mn;
      }
    }
  }
  if (p < n) out[p] = has_nan ? -INFINITY : (m + log2f(S)) * 0.69314718055994530942f;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al. 2011) + Box-Muller
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01_open(uint32_t v) { return (static_cast<float>(v >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__device__ __forceinline__ void normal4(uint64_t counter, uint64_t seed, float (&z)[4]) {
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32), 0u, 0u),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  const float r0 = sqrtf(-2.f * logf(u01_open(r.x))), r1 = sqrtf(-2.f * logf(u01_open(r.z)));
  float s0, c0, s1, c1;
  sincospif(2.f * u01_open(r.y), &s0, &c0);
  sincospif(2.f * u01_open(r.w), &s1, &c1);
  z[0] = r0 * c0;
  z[1] = r0 * s0;
  z[2] = r1 * c1;
  z[3] = r1 * s1;
}

// out[(k*per + j)*D + d] = means[k*D + d] + sd * z. The Philox counter of group q is q + offset. With d_state (two
// words in device memory: {stream position, key}) both come from there instead: a CUDA-graph replay must neither re-use
// the offsets nor keep the seed that were baked in at capture.
__global__ void sample_iso_kernel(const float* __restrict__ means, int64_t K, int D, int64_t per, float sd,
                                  uint64_t seed, uint64_t offset, const unsigned long long* __restrict__ d_state,
                                  float* __restrict__ out) {
  const int64_t total = K * per * D;
  const int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;  // group of 4 outputs
  if (q * 4 >= total) return;
  if (d_state != nullptr) {
    offset = d_state[0];
    seed = d_state[1];
  }
  float z[4];
  normal4(static_cast<uint64_t>(q) + offset, seed, z);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t e = q * 4 + i;
    if (e < total) {
      const int64_t row = e / D;
      const int d = static_cast<int>(e % D);
      out[e] = fmaf(sd, z[i], means[(row / per) * D + d]);
    }
  }
}

// Ancestral sampling from a Gaussian mixture with full covariances (CMixtureModel.sample, CMixtureModel.py:79-89, with
// CMultivariateNormal.sample, CMultivariateNormal.py:64-65, as the leaf): thread = draw;
//   out[i] = means[idx[i]] + L[idx[i]] z_i,   z_i ~ N(0, I) from Philox, L lower-triangular (row-major [K, D, D]).
template <int DP>
__global__ void __launch_bounds__(128) sample_mixture_kernel(const float* __restrict__ means,
                                                             const float* __restrict__ chol, int64_t K, int D,
                                                             const int64_t* __restrict__ idx, int64_t n, uint64_t seed,
                                                             uint64_t offset, float* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  int64_t k = idx[i];
  k = k < 0 ? 0 : (k >= K ? K - 1 : k);
  constexpr int G = (DP + 3) / 4;
  float z[G * 4];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float z4[4];
    normal4(static_cast<uint64_t>(i) * G + g + offset, seed, z4);
#pragma unroll
    for (int e = 0; e < 4; ++e) z[4 * g + e] = z4[e];
  }
  const float* L = chol + k * static_cast<int64_t>(D) * D;
  const float* mu = means + k * static_cast<int64_t>(D);
#pragma unroll
  for (int r = 0; r < DP; ++r) {
    if (r < D) {
      float acc = __ldg(mu + r);
#pragma unroll
      for (int c = 0; c <= r; ++c) acc = fmaf(__ldg(L + r * D + c), z[c], acc);
      out[i * D + r] = acc;
    }
  }
}

template <int DP>
static int launch_sample_mixture(const float* means, const float* chol, int64_t K, int D, const int64_t* idx, int64_t n,
                                 uint64_t seed, uint64_t offset, float* out, cudaStream_t st) {
  sample_mixture_kernel<DP><<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(means, chol, K, D, idx, n, seed,
                                                                                   offset, out);
  AISB_LAUNCH_CHECK("sample_mixture_kernel");
  return AISB_OK;
}

// Z-order (Morton) keys of a point set: `bits` most significant bits of every coordinate's position inside the
// bounding box [lo, lo + span], interleaved (bit b of coordinate d -> key bit b * D + d; bits * D <= 32). The key is
// stored with its top bit flipped so that SIGNED 32-bit order equals the unsigned order of the code (a 32-bit radix
// sort then gives the curve order). One read of the rows; replaces ~25 element-wise launches.
__global__ void __launch_bounds__(256) morton_keys_kernel(const float* __restrict__ x, int64_t n, int D, int bits,
                                                          const float* __restrict__ lo, const float* __restrict__ span,
                                                          int32_t* __restrict__ keys) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float scale = static_cast<float>(1 << bits);
  const float top = static_cast<float>((1 << bits) - 1);
  uint32_t code = 0;
  for (int d = 0; d < D; ++d) {
    float v = __fmul_rn(__fdiv_rn(__fsub_rn(x[i * D + d], lo[d]), fmaxf(span[d], 1e-30f)), scale);
    v = fminf(fmaxf(v, 0.f), top);   // NaN -> 0
    const uint32_t q = static_cast<uint32_t>(static_cast<int>(v));
    for (int b = 0; b < bits; ++b) code |= ((q >> b) & 1u) << (b * D + d);
  }
  keys[i] = static_cast<int32_t>(code ^ 0x80000000u);
}


// Bounding box of a point set in ONE pass (replaces two torch reductions + their glue per Z-order): per-dimension
// minimum / maximum through block reductions and atomicMin / atomicMax on order-preserving unsigned encodings of the
// floats (enc(f) = bits ^ 0x80000000 for f >= 0, ~bits otherwise; NaN coordinates are skipped). box[0..D) = encoded
// minima (initialised to 0xFFFFFFFF), box[D..2D) = encoded maxima (initialised to 0).
__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ x, int64_t n, int D,
                                                   uint32_t* __restrict__ box) {
  __shared__ uint32_t smin[AISB_MAX_DIMS], smax[AISB_MAX_DIMS];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    smin[d] = 0xFFFFFFFFu;
    smax[d] = 0u;
  }
  __syncthreads();
  const int64_t total = n * D;
  // flat, coalesced sweep; element e belongs to dimension e % D
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = __ldg(x + e);
    if (v == v) {
      const uint32_t c = enc_ordered(v);
      const int d = static_cast<int>(e % D);
      atomicMin(&smin[d], c);
      atomicMax(&smax[d], c);
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    atomicMin(&box[d], smin[d]);
    atomicMax(&box[D + d], smax[d]);
  }
}

// morton_keys_kernel with the box taken from bbox_kernel's encoded output
__global__ void __launch_bounds__(256) morton_keys_box_kernel(const float* __restrict__ x, int64_t n, int D, int bits,
                                                              const uint32_t* __restrict__ box,
                                                              int32_t* __restrict__ keys) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float scale = static_cast<float>(1 << bits);
  const float top = static_cast<float>((1 << bits) - 1);
  uint32_t code = 0;
  for (int d = 0; d < D; ++d) {
    const float lo = dec_ordered(box[d]);
    const float span = __fsub_rn(dec_ordered(box[D + d]), lo);
    float v = __fmul_rn(__fdiv_rn(__fsub_rn(x[i * D + d], lo), fmaxf(span, 1e-30f)), scale);
    v = fminf(fmaxf(v, 0.f), top);   // NaN -> 0
    const uint32_t q = static_cast<uint32_t>(static_cast<int>(v));
    for (int b = 0; b < bits; ++b) code |= ((q >> b) & 1u) << (b * D + d);
  }
  keys[i] = static_cast<int32_t>(code ^ 0x80000000u);
}

__global__ void counter_add_kernel(unsigned long long* ctr, unsigned long long inc) { *ctr += inc; }

__global__ void uniform_box_kernel(const float* __restrict__ lo, const float* __restrict__ hi, int D, int64_t n,
                                   uint64_t seed, uint64_t offset, float* __restrict__ out) {
  const int64_t total = n * D;
  const int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (q * 4 >= total) return;
  const uint64_t counter = static_cast<uint64_t>(q) + offset;
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32), 1u, 0u),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t e = q * 4 + i;
    if (e < total) {
      const int d = static_cast<int>(e % D);
      out[e] = fmaf(hi[d] - lo[d], u01_open(rv[i]), lo[d]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Peak micro-benchmarks: 8 independent dependent-chains per thread, no memory traffic
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) peak_fma_kernel(int64_t iters, float* sink, float mul, float add) {
  // packed FFMA2 chains (the form the pairwise kernels use): 8 chains x 8 rounds per iteration. Multiplier and addend
  // are kernel parameters, i.e. UNIFORM operands (FFMA2 R, R, UR, UR): with three vector-register operands the register
  // file, not the FMA pipe, limits the rate (measured 67.6 vs 74.0 TFLOP/s). 64 packed instructions = 128 FMAs per lane.
  float2 a[8];
  const float2 m = make_float2(mul, mul), c = make_float2(add, add);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(0.001f * (threadIdx.x + i), 0.002f * (threadIdx.x + i));
  for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], m, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  if (s == 123456.789f) sink[0] = s;  // never true; keeps the chains alive
}

__global__ void __launch_bounds__(256) peak_ex2_kernel(int64_t iters, float* sink) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.1f + 0.001f * (threadIdx.x + i);
  for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = ex2(-a[i]);  // contraction towards ~0.641; stays in (0.5, 1]
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123456.789f) sink[0] = s;
}

template <int DP>
static int launch_box(const float* x, int64_t n, int D, const float* lo, const float* hi, const float* lwv, int64_t K,
                      int open_upper, float* out, cudaStream_t st) {
  box_mixture_kernel<DP><<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(x, n, D, lo, hi, lwv, K, open_upper, out);
  AISB_LAUNCH_CHECK("box_mixture_kernel");
  return AISB_OK;
}

}  // namespace aisb

using namespace aisb;

extern "C" int aisb_box_mixture_logpdf_f32(const float* d_x, int64_t n, int32_t D, const float* d_lo,
                                           const float* d_hi, const float* d_logw_over_vol, int64_t K,
                                           int32_t open_upper, float* d_out_logp, void* stream) {
  const int DP = aisb_padded_dims(D);
  if (DP == 0 || DP > 32 || n < 0 || K < 1 || !d_lo || !d_hi || !d_logw_over_vol || (n > 0 && (!d_x || !d_out_logp))) {
    set_error("box_mixture_logpdf: invalid argument (n=%lld K=%lld D=%d; D <= 32)", static_cast<long long>(n),
              static_cast<long long>(K), D);
    return AISB_ERR_INVALID;
  }
  if (n == 0) return AISB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (DP) {
    case 1: return launch_box<1>(d_x, n, D, d_lo, d_hi, d_logw_over_vol, K, open_upper, d_out_logp, st);
    case 2: return launch_box<2>(d_x, n, D, d_lo, d_hi, d_logw_over_vol, K, open_upper, d_out_logp, st);
    case 4: return launch_box<4>(d_x, n, D, d_lo, d_hi, d_logw_over_vol, K, open_upper, d_out_logp, st);
    case 8: return launch_box<8>(d_x, n, D, d_lo, d_hi, d_logw_over_vol, K, open_upper, d_out_logp, st);
    case 16: return launch_box<16>(d_x, n, D, d_lo, d_hi, d_logw_over_vol, K, open_upper, d_out_logp, st);
    default: return launch_box<32>(d_x, n, D, d_lo, d_hi, d_logw_over_vol, K, open_upper, d_out_logp, st);
  }
}

extern "C" int aisb_sample_iso_f32(const float* d_means, int64_t K, int32_t D, int64_t per, double var, uint64_t seed,
                                   uint64_t offset, float* d_out, void* stream) {
  if (!d_means || !d_out || K < 1 || per < 0 || D < 1 || D > AISB_MAX_DIMS || !(var >= 0.0)) {
    set_error("sample_iso: invalid argument");
    return AISB_ERR_INVALID;
  }
  const int64_t total = K * per * D;
  if (total == 0) return AISB_OK;
  const int64_t groups = (total + 3) / 4;
  sample_iso_kernel<<<static_cast<unsigned>((groups + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_means, K, D, per, static_cast<float>(sqrt(var)), seed, offset, nullptr, d_out);
  AISB_LAUNCH_CHECK("sample_iso_kernel");
  return AISB_OK;
}

extern "C" int aisb_morton_keys_f32(const float* d_x, int64_t n, int32_t D, int32_t bits, const float* d_lo,
                                    const float* d_span, int32_t* d_keys, void* stream) {
  if (n < 0 || D < 1 || D > AISB_MAX_DIMS || bits < 1 || bits * D > 32 || !d_lo || !d_span ||
      (n > 0 && (!d_x || !d_keys))) {
    set_error("morton_keys: invalid argument (n=%lld D=%d bits=%d; bits * D must not exceed 32)",
              static_cast<long long>(n), D, bits);
    return AISB_ERR_INVALID;
  }
  if (n == 0) return AISB_OK;
  morton_keys_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_x, n, D, bits, d_lo, d_span, d_keys);
  AISB_LAUNCH_CHECK("morton_keys_kernel");
  return AISB_OK;
}


extern "C" int aisb_morton_keys_auto_f32(const float* d_x, int64_t n, int32_t D, int32_t bits, uint32_t* d_box,
                                         int32_t* d_keys, void* stream) {
  if (n < 0 || D < 1 || D > AISB_MAX_DIMS || bits < 1 || bits * D > 32 || !d_box || (n > 0 && (!d_x || !d_keys))) {
    set_error("morton_keys_auto: invalid argument (n=%lld D=%d bits=%d; bits * D must not exceed 32)",
              static_cast<long long>(n), D, bits);
    return AISB_ERR_INVALID;
  }
  if (n == 0) return AISB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AISB_CUDA_CHECK(cudaMemsetAsync(d_box, 0xFF, sizeof(uint32_t) * D, st));
  AISB_CUDA_CHECK(cudaMemsetAsync(d_box + D, 0, sizeof(uint32_t) * D, st));
  int64_t blocks = (n * D + 256 * 16 - 1) / (256 * 16);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  bbox_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(d_x, n, D, d_box);
  AISB_LAUNCH_CHECK("bbox_kernel");
  morton_keys_box_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(d_x, n, D, bits, d_box, d_keys);
  AISB_LAUNCH_CHECK("morton_keys_box_kernel");
  return AISB_OK;
}

// out[i, :] = x[idx[i], :]: one thread per 16-byte piece of a row when rows are whole float4s (D % 4 == 0 and aligned
// bases), else one thread per element. (torch's index kernel launches one block per ROW for this shape: 0.5 ms for the
// 1e6 x 8 points of configs[3] under ncu, on the serial path of every KDE evaluation.)
__global__ void gather_rows_vec_kernel(const float4* __restrict__ x, const int64_t* __restrict__ idx, int64_t m, int q,
                                       int64_t n_src, float4* __restrict__ out, int* __restrict__ err) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= m * q) return;
  const int64_t i = t / q;
  const int c = static_cast<int>(t - i * q);
  const int64_t r = idx[i];
  if (r < 0 || r >= n_src) {
    if (err) atomicAdd(err, 1);
    return;
  }
  out[t] = __ldg(x + r * q + c);
}

__global__ void gather_rows_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, int64_t m, int D,
                                   int64_t n_src, float* __restrict__ out, int* __restrict__ err) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= m * D) return;
  const int64_t i = t / D;
  const int d = static_cast<int>(t - i * D);
  const int64_t r = idx[i];
  if (r < 0 || r >= n_src) {
    if (err) atomicAdd(err, 1);
    return;
  }
  out[t] = __ldg(x + r * D + d);
}

extern "C" int aisb_gather_rows_f32(const float* d_x, int64_t n_src, int32_t D, const int64_t* d_idx, int64_t m,
                                    float* d_out, int* d_err, void* stream) {
  if (m < 0 || n_src < 0 || D < 1 || (m > 0 && (!d_x || !d_idx || !d_out))) {
    set_error("gather_rows: invalid argument (m=%lld n_src=%lld D=%d)", static_cast<long long>(m),
              static_cast<long long>(n_src), D);
    return AISB_ERR_INVALID;
  }
  if (m == 0) return AISB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = D % 4 == 0 && reinterpret_cast<uintptr_t>(d_x) % 16 == 0 && reinterpret_cast<uintptr_t>(d_out) % 16 == 0;
  if (vec) {
    const int q = D / 4;
    gather_rows_vec_kernel<<<static_cast<unsigned>((m * q + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(d_x), d_idx, m, q, n_src, reinterpret_cast<float4*>(d_out), d_err);
  } else {
    gather_rows_kernel<<<static_cast<unsigned>((m * D + 255) / 256), 256, 0, st>>>(d_x, d_idx, m, D, n_src, d_out, d_err);
  }
  AISB_LAUNCH_CHECK("gather_rows_kernel");
  return AISB_OK;
}

extern "C" int aisb_sample_iso_ctr_f32(const float* d_means, int64_t K, int32_t D, int64_t per, double var,
                                       uint64_t* d_state, float* d_out, void* stream) {
  uint64_t* d_counter = d_state;
  if (!d_means || !d_out || !d_state || K < 1 || per < 0 || D < 1 || D > AISB_MAX_DIMS || !(var >= 0.0)) {
    set_error("sample_iso_ctr: invalid argument");
    return AISB_ERR_INVALID;
  }
  const int64_t total = K * per * D;
  if (total == 0) return AISB_OK;
  const int64_t groups = (total + 3) / 4;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  sample_iso_kernel<<<static_cast<unsigned>((groups + 255) / 256), 256, 0, st>>>(
      d_means, K, D, per, static_cast<float>(sqrt(var)), 0, 0, reinterpret_cast<unsigned long long*>(d_state), d_out);
  AISB_LAUNCH_CHECK("sample_iso_kernel");
  counter_add_kernel<<<1, 1, 0, st>>>(reinterpret_cast<unsigned long long*>(d_counter),
                                      static_cast<unsigned long long>(groups + 1));
  AISB_LAUNCH_CHECK("counter_add_kernel");
  return AISB_OK;
}

extern "C" int aisb_sample_mixture_f32(const float* d_means, const float* d_chol, int64_t K, int32_t D,
                                       const int64_t* d_idx, int64_t n, uint64_t seed, uint64_t offset, float* d_out,
                                       void* stream) {
  if (!d_means || !d_chol || !d_idx || !d_out || K < 1 || n < 0 || D < 1 || D > AISB_MAX_DIMS) {
    set_error("sample_mixture: invalid argument");
    return AISB_ERR_INVALID;
  }
  if (n == 0) return AISB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (aisb_padded_dims(D)) {
    case 1: return launch_sample_mixture<1>(d_means, d_chol, K, D, d_idx, n, seed, offset, d_out, st);
    case 2: return launch_sample_mixture<2>(d_means, d_chol, K, D, d_idx, n, seed, offset, d_out, st);
    case 4: return launch_sample_mixture<4>(d_means, d_chol, K, D, d_idx, n, seed, offset, d_out, st);
    case 8: return launch_sample_mixture<8>(d_means, d_chol, K, D, d_idx, n, seed, offset, d_out, st);
    case 16: return launch_sample_mixture<16>(d_means, d_chol, K, D, d_idx, n, seed, offset, d_out, st);
    case 32: return launch_sample_mixture<32>(d_means, d_chol, K, D, d_idx, n, seed, offset, d_out, st);
    case 64: return launch_sample_mixture<64>(d_means, d_chol, K, D, d_idx, n, seed, offset, d_out, st);
    default: break;
  }
  set_error("sample_mixture: unsupported D=%d", D);
  return AISB_ERR_UNSUPPORTED;
}

extern "C" int aisb_uniform_box_f32(const float* d_lo, const float* d_hi, int32_t D, int64_t n, uint64_t seed,
                                    uint64_t offset, float* d_out, void* stream) {
  if (!d_lo || !d_hi || !d_out || n < 0 || D < 1 || D > AISB_MAX_DIMS) {
    set_error("uniform_box: invalid argument");
    return AISB_ERR_INVALID;
  }
  const int64_t total = n * D;
  if (total == 0) return AISB_OK;
  const int64_t groups = (total + 3) / 4;
  uniform_box_kernel<<<static_cast<unsigned>((groups + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_lo, d_hi, D, n, seed, offset, d_out);
  AISB_LAUNCH_CHECK("uniform_box_kernel");
  return AISB_OK;
}

extern "C" int aisb_peak_fma_f32(int64_t iters, float* d_sink, double* ops, void* stream) {
  if (iters < 1 || !d_sink) {
    set_error("peak: invalid argument");
    return AISB_ERR_INVALID;
  }
  const int blocks = sm_count() * 8;
  peak_fma_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(iters, d_sink, 0.999999f, 1.0e-7f);
  AISB_LAUNCH_CHECK("peak_fma_kernel");
  // 64 packed instructions per iteration, two FMAs per lane each
  if (ops) *ops = static_cast<double>(blocks) * 256.0 * static_cast<double>(iters) * 128.0;
  return AISB_OK;
}
extern "C" int aisb_peak_ex2_f32(int64_t iters, float* d_sink, double* ops, void* stream) {
  if (iters < 1 || !d_sink) {
    set_error("peak: invalid argument");
    return AISB_ERR_INVALID;
  }
  const int blocks = sm_count() * 8;
  peak_ex2_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(iters, d_sink);
  AISB_LAUNCH_CHECK("peak_ex2_kernel");
  if (ops) *ops = static_cast<double>(blocks) * 256.0 * static_cast<double>(iters) * 32.0;
  return AISB_OK;
}
