// GroupNorm(32)+SiLU and LayerNorm for pixel-major activations — HBM-bound kernels.
//
// GroupNorm (GroupNorm32, util.py:217-219; Normalize, attention.py:88-89):
//   pass 1 (gn_stats): grid (chunks, B).  A CTA streams `rows_per_chunk` pixels of one
//     image, every thread owning fixed channel PAIRS (so loads are fully coalesced 4-byte
//     /8-byte per lane and a pair never straddles a group: C/32 is even for every width
//     on this path), accumulates per-channel sum / sum-of-squares in registers, reduces
//     them per group through shared memory and writes ONE (sum, sumsq) partial per
//     (image, chunk, group) — deterministic, no atomics, no memset.
//   pass 2 (gn_apply): every CTA first folds the partials (in double) into per-channel
//     mean / rstd*gamma / beta tables in shared memory, then streams its rows with
//     16-byte accesses: y = act((x - mean) * rstd*gamma + beta).
//   Algorithmic bytes: read x twice (second read normally hits the 126 MB L2) + write y.
//
// LayerNorm (nn.LayerNorm, attention.py:263-265): one warp per row, the row kept in
//   registers, two-pass (mean, then centred variance) like torch.
#include "common.cuh"

namespace pd {

constexpr int GN_CHUNKS_MAX = 64;
constexpr int GN_GROUPS_MAX = 32;
constexpr int GN_THREADS = 256;
constexpr int GN_MAX_PAIRS_PER_THREAD = 6;  // C <= 2*256*6 = 3072

__host__ __device__ inline int gn_num_chunks(int HW) {
  // ~128 pixels per chunk, capped
  int c = (HW + 127) / 128;
  if (c < 1) c = 1;
  if (c > GN_CHUNKS_MAX) c = GN_CHUNKS_MAX;
  return c;
}

template <typename T> struct Pair;
template <> struct Pair<float> {
  __device__ __forceinline__ static float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
};
template <> struct Pair<bf16> {
  __device__ __forceinline__ static float2 ld(const bf16* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
  }
};

template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const T* __restrict__ x, int ldx, float* __restrict__ partial, int HW, int C, int groups,
                int chunks) {
  extern __shared__ float sm[];  // [2*C] per-channel sum, sumsq
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int rows_per = (HW + chunks - 1) / chunks;
  const int r0 = chunk * rows_per;
  const int r1 = min(HW, r0 + rows_per);
  const int pairs = C / 2;
  const T* base = x + (int64_t)b * HW * ldx;

  float s[GN_MAX_PAIRS_PER_THREAD][2], q[GN_MAX_PAIRS_PER_THREAD][2];
#pragma unroll
  for (int i = 0; i < GN_MAX_PAIRS_PER_THREAD; ++i) s[i][0] = s[i][1] = q[i][0] = q[i][1] = 0.f;

  for (int r = r0; r < r1; ++r) {
    const T* row = base + (int64_t)r * ldx;
#pragma unroll
    for (int i = 0; i < GN_MAX_PAIRS_PER_THREAD; ++i) {
      int j = threadIdx.x + i * GN_THREADS;
      if (j < pairs) {
        float2 v = Pair<T>::ld(row + 2 * j);
        s[i][0] += v.x; q[i][0] += v.x * v.x;
        s[i][1] += v.y; q[i][1] += v.y * v.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < GN_MAX_PAIRS_PER_THREAD; ++i) {
    int j = threadIdx.x + i * GN_THREADS;
    if (j < pairs) {
      sm[2 * j] = s[i][0]; sm[2 * j + 1] = s[i][1];
      sm[C + 2 * j] = q[i][0]; sm[C + 2 * j + 1] = q[i][1];
    }
  }
  __syncthreads();
  const int cpg = C / groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = warp; g < groups; g += GN_THREADS / 32) {
    float a = 0.f, bq = 0.f;
    for (int c = lane; c < cpg; c += 32) { a += sm[g * cpg + c]; bq += sm[C + g * cpg + c]; }
    a = warp_sum(a); bq = warp_sum(bq);
    if (lane == 0) {
      float* dst = partial + (((int64_t)b * GN_CHUNKS_MAX + chunk) * GN_GROUPS_MAX + g) * 2;
      dst[0] = a; dst[1] = bq;
    }
  }
}

template <typename T, typename TO, int V>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const T* __restrict__ x, int ldx, TO* __restrict__ out, int ldo,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ partial, int HW, int C, int groups, int chunks, float eps,
                int act, int row_blocks) {
  extern __shared__ float sm[];  // mean[C], a[C], beta[C], gstat[2*groups]
  float* s_mean = sm;
  float* s_a = sm + C;
  float* s_b = sm + 2 * C;
  float* s_g = sm + 3 * C;
  const int b = blockIdx.y;
  const int cpg = C / groups;
  if (threadIdx.x < groups) {
    double su = 0.0, sq = 0.0;
    for (int ch = 0; ch < chunks; ++ch) {
      const float* p = partial + (((int64_t)b * GN_CHUNKS_MAX + ch) * GN_GROUPS_MAX + threadIdx.x) * 2;
      su += (double)p[0]; sq += (double)p[1];
    }
    double n = (double)HW * (double)cpg;
    double mean = su / n;
    double var = sq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_g[2 * threadIdx.x] = (float)mean;
    s_g[2 * threadIdx.x + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    int g = c / cpg;
    s_mean[c] = s_g[2 * g];
    s_a[c] = s_g[2 * g + 1] * gamma[c];
    s_b[c] = beta[c];
  }
  __syncthreads();

  const int vpr = C / V;
  const int rows_per = (HW + row_blocks - 1) / row_blocks;
  const int r0 = blockIdx.x * rows_per;
  const int r1 = min(HW, r0 + rows_per);
  const int64_t nvec = (int64_t)(r1 - r0) * vpr;
  const T* xb = x + ((int64_t)b * HW + r0) * ldx;
  TO* ob = out + ((int64_t)b * HW + r0) * ldo;
  for (int64_t i = threadIdx.x; i < nvec; i += GN_THREADS) {
    int r = (int)(i / vpr);
    int c = (int)(i % vpr) * V;
    float f[V];
    if constexpr (sizeof(T) == 2 && V == 8) {
      unpack8(*reinterpret_cast<const bf16x8*>(xb + (int64_t)r * ldx + c), f);
    } else if constexpr (sizeof(T) == 4 && V == 4) {
      float4 t = *reinterpret_cast<const float4*>(xb + (int64_t)r * ldx + c);
      f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) f[k] = Dt<T>::ld(xb + (int64_t)r * ldx + c + k);
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float y = (f[k] - s_mean[c + k]) * s_a[c + k] + s_b[c + k];
      f[k] = act == PD_ACT_SILU ? (sizeof(TO) == 4 ? silu_acc(y) : silu_f(y)) : y;
    }
    if constexpr (sizeof(TO) == 2 && V == 8) {
      *reinterpret_cast<bf16x8*>(ob + (int64_t)r * ldo + c) = pack8(f);
    } else if constexpr (sizeof(TO) == 4 && V == 4) {
      *reinterpret_cast<float4*>(ob + (int64_t)r * ldo + c) = make_float4(f[0], f[1], f[2], f[3]);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) Dt<TO>::st(ob + (int64_t)r * ldo + c + k, f[k]);
    }
  }
}

// ---- LayerNorm -------------------------------------------------------------------------
constexpr int LN_MAX_PER_LANE = 48;  // C <= 1536

template <typename T>
__global__ void __launch_bounds__(256)
layer_norm_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ldo,
                  const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows, int C,
                  float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int per_lane = (C + 31) / 32;  // elements, strided by 32*V below
  (void)per_lane;
  constexpr int V = sizeof(T) == 2 ? 8 : 4;
  const int nvec = C / V;              // vectors per row (C % V == 0 enforced on host)
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += warps_total) {
    const T* xr = x + row * ldx;
    float f[LN_MAX_PER_LANE];
    int cnt = 0;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE / V; ++i) {
      int j = lane + i * 32;
      if (j < nvec) {
        if constexpr (sizeof(T) == 2) {
          unpack8(*reinterpret_cast<const bf16x8*>(xr + j * V), &f[i * V]);
        } else {
          float4 t = *reinterpret_cast<const float4*>(xr + j * V);
          f[i * V] = t.x; f[i * V + 1] = t.y; f[i * V + 2] = t.z; f[i * V + 3] = t.w;
        }
#pragma unroll
        for (int k = 0; k < V; ++k) sum += f[i * V + k];
        cnt = i + 1;
      }
    }
    (void)cnt;
    sum = warp_sum(sum);
    const float mean = sum / (float)C;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE / V; ++i) {
      int j = lane + i * 32;
      if (j < nvec) {
#pragma unroll
        for (int k = 0; k < V; ++k) { float d = f[i * V + k] - mean; var += d * d; }
      }
    }
    var = warp_sum(var) / (float)C;
    const float rstd = rsqrtf(var + eps);
    T* orow = out + row * ldo;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE / V; ++i) {
      int j = lane + i * 32;
      if (j < nvec) {
        float y[V];
#pragma unroll
        for (int k = 0; k < V; ++k)
          y[k] = (f[i * V + k] - mean) * rstd * gamma[j * V + k] + beta[j * V + k];
        if constexpr (sizeof(T) == 2) {
          *reinterpret_cast<bf16x8*>(orow + j * V) = pack8(y);
        } else {
          *reinterpret_cast<float4*>(orow + j * V) = make_float4(y[0], y[1], y[2], y[3]);
        }
      }
    }
  }
}

template <typename T, typename TO>
static int gn_launch(const void* x, int ldx, void* out, int ldo, const float* gamma, const float* beta,
                     float* partial, int B, int HW, int C, int groups, float eps, int act, cudaStream_t s) {
  const int chunks = gn_num_chunks(HW);
  gn_stats_kernel<T><<<dim3(chunks, B), GN_THREADS, 2 * C * sizeof(float), s>>>((const T*)x, ldx, partial, HW,
                                                                              C, groups, chunks);
  int rc = check_launch("gn_stats");
  if (rc) return rc;
  // apply: ~4 waves of CTAs over (row_blocks, B)
  int row_blocks = (num_sms() * 4 + B - 1) / B;
  int max_rb = (HW + 15) / 16;
  if (row_blocks > max_rb) row_blocks = max_rb;
  if (row_blocks < 1) row_blocks = 1;
  size_t smem = (3 * (size_t)C + 2 * groups) * sizeof(float);
  constexpr int V = sizeof(T) == 2 ? 8 : 4;
  bool vec = (C % V == 0) && (ldx % V == 0) && (ldo % V == 0) && ((uintptr_t)x % 16 == 0) &&
             ((uintptr_t)out % 16 == 0) && sizeof(T) == sizeof(TO);
  if (vec)
    gn_apply_kernel<T, TO, V><<<dim3(row_blocks, B), GN_THREADS, smem, s>>>(
        (const T*)x, ldx, (TO*)out, ldo, gamma, beta, partial, HW, C, groups, chunks, eps, act, row_blocks);
  else
    gn_apply_kernel<T, TO, 2><<<dim3(row_blocks, B), GN_THREADS, smem, s>>>(
        (const T*)x, ldx, (TO*)out, ldo, gamma, beta, partial, HW, C, groups, chunks, eps, act, row_blocks);
  return check_launch("gn_apply");
}

}  // namespace pd

using namespace pd;

extern "C" {

int64_t pd_group_norm_scratch_floats(int32_t B) {
  return (int64_t)B * GN_CHUNKS_MAX * GN_GROUPS_MAX * 2;
}

int pd_group_norm(const void* x, int32_t ldx, void* out, int32_t ldo, const float* gamma, const float* beta,
                  float* partial, int32_t B, int32_t HW, int32_t C, int32_t groups, float eps, int32_t act,
                  int32_t dtype, int32_t out_dtype, void* stream) {
  PD_REQUIRE(x && out && gamma && beta && partial, "pd_group_norm: null pointer");
  PD_REQUIRE(B > 0 && HW > 0 && C > 0 && groups > 0 && groups <= GN_GROUPS_MAX && C % groups == 0,
             "pd_group_norm: bad geometry B=%d HW=%d C=%d groups=%d", B, HW, C, groups);
  PD_REQUIRE((C / groups) % 2 == 0 && C <= 2 * GN_THREADS * GN_MAX_PAIRS_PER_THREAD && ldx % 2 == 0 &&
                 ldo % 2 == 0 && ldx >= C && ldo >= C,
             "pd_group_norm: need even channels-per-group, C <= %d, even pitches",
             2 * GN_THREADS * GN_MAX_PAIRS_PER_THREAD);
  PD_REQUIRE(B <= 65535, "pd_group_norm: B too large");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PD_F32 && out_dtype == PD_F32)
    return gn_launch<float, float>(x, ldx, out, ldo, gamma, beta, partial, B, HW, C, groups, eps, act, s);
  if (dtype == PD_BF16 && out_dtype == PD_BF16)
    return gn_launch<bf16, bf16>(x, ldx, out, ldo, gamma, beta, partial, B, HW, C, groups, eps, act, s);
  if (dtype == PD_F32 && out_dtype == PD_BF16)
    return gn_launch<float, bf16>(x, ldx, out, ldo, gamma, beta, partial, B, HW, C, groups, eps, act, s);
  PD_REQUIRE(false, "pd_group_norm: unsupported dtypes %d -> %d", dtype, out_dtype);
}

int pd_layer_norm(const void* x, int32_t ldx, void* out, int32_t ldo, const float* gamma, const float* beta,
                  int64_t rows, int32_t C, float eps, int32_t dtype, void* stream) {
  PD_REQUIRE(x && out && gamma && beta && rows > 0 && C > 0 && ldx >= C && ldo >= C, "pd_layer_norm: bad args");
  const int V = dtype == PD_BF16 ? 8 : 4;
  PD_REQUIRE(C % V == 0 && ldx % V == 0 && ldo % V == 0 && C <= 32 * LN_MAX_PER_LANE &&
                 ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0,
             "pd_layer_norm: C and pitches must be multiples of %d, C <= %d, 16B aligned", V,
             32 * LN_MAX_PER_LANE);
  cudaStream_t s = (cudaStream_t)stream;
  int64_t blocks = (rows + 7) / 8;
  int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == PD_F32)
    layer_norm_kernel<float><<<(int)blocks, 256, 0, s>>>((const float*)x, ldx, (float*)out, ldo, gamma, beta,
                                                         rows, C, eps);
  else if (dtype == PD_BF16)
    layer_norm_kernel<bf16><<<(int)blocks, 256, 0, s>>>((const bf16*)x, ldx, (bf16*)out, ldo, gamma, beta, rows,
                                                        C, eps);
  else
    PD_REQUIRE(false, "pd_layer_norm: bad dtype %d", dtype);
  return check_launch("pd_layer_norm");
}

}  // extern "C"
