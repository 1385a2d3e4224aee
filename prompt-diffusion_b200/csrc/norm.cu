// GroupNorm(32)+SiLU and LayerNorm for pixel-major activations — HBM-bound kernels.
//
// GroupNorm (GroupNorm32, util.py:217-219; Normalize, attention.py:88-89):
//   pass 1 (gn_stats): grid (chunks, B).  Threads form a (TY rows) x (TX channel-vectors) grid: a thread
//     owns ONE 16-byte channel vector (8 bf16 / 4 fp32) and walks rows r = ty, ty+TY, ... of its chunk, so
//     every warp load is contiguous, there is no index arithmetic in the loop and the 8 (4) per-channel
//     sum / sum-of-squares accumulators live in registers.  Per-channel totals are combined through shared
//     memory, reduced per group by one warp each, and ONE (sum, sumsq) partial per (image, chunk, group)
//     is written — deterministic, no global atomics, no memset.
//   pass 2 (gn_apply): same thread grid; each thread first folds the partials of its channels' groups (in
//     double) into per-channel mean / rstd*gamma / beta REGISTERS, then streams its rows:
//     y = act((x - mean) * (rstd*gamma) + beta).
//   Algorithmic bytes: read x twice (second read normally hits the 126 MB L2) + write y.
//
// LayerNorm (nn.LayerNorm, attention.py:263-265): one warp per row, the row kept in registers, two-pass
//   (mean, then centred variance) like torch; several rows in flight per warp for memory-level parallelism.
#include "common.cuh"

namespace pd {

constexpr int GN_CHUNKS_MAX = 64;
constexpr int GN_GROUPS_MAX = 32;
constexpr int GN_THREADS = 512;
#ifndef GN_ROWS_IN_FLIGHT
#define GN_ROWS_IN_FLIGHT 4       // 8: eight raw row vectors in flight in the statistics pass, four in the apply pass (bf16):
                                  // -14 % on the kernel alone at 16 x 4096 x 320, -0.05 ms per step in situ (noise level): not adopted
#endif
static bool g_gn_fused = true;    // pd_debug_group_norm_fused(0) falls back to the two-kernel form (A/B timing, tests)

__host__ __device__ inline int gn_num_chunks(int HW) {
  int c = (HW + 63) / 64;  // ~64 pixels per chunk, capped
  if (c < 1) c = 1;
  if (c > GN_CHUNKS_MAX) c = GN_CHUNKS_MAX;
  return c;
}

template <typename T> struct VecIO;
template <> struct VecIO<bf16> {
  static constexpr int V = 8;
  __device__ __forceinline__ static void ld(const bf16* p, float* f) { unpack8(*reinterpret_cast<const bf16x8*>(p), f); }
  __device__ __forceinline__ static void st(bf16* p, const float* f) { *reinterpret_cast<bf16x8*>(p) = pack8(f); }
};
template <> struct VecIO<float> {
  static constexpr int V = 4;
  __device__ __forceinline__ static void ld(const float* p, float* f) {
    float4 t = *reinterpret_cast<const float4*>(p);
    f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
  }
  __device__ __forceinline__ static void st(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};

// thread grid shared by both passes: TX = min(vectors per row, blockDim), TY = blockDim / TX
struct GnGrid { int vpr, tx_n, ty_n, passes; };
__host__ __device__ inline GnGrid gn_grid(int C, int V) {
  GnGrid g;
  g.vpr = C / V;
  g.tx_n = g.vpr < GN_THREADS ? g.vpr : GN_THREADS;
  g.ty_n = GN_THREADS / g.tx_n;
  g.passes = (g.vpr + g.tx_n - 1) / g.tx_n;  // > 1 only when a row has more than 512 vectors
  return g;
}

template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const T* __restrict__ x, int ldx, float* __restrict__ partial, int HW, int C, int groups,
                int chunks, float eps, float* __restrict__ stats, unsigned int* __restrict__ counters) {
  constexpr int V = VecIO<T>::V;
  extern __shared__ float sm[];  // [TY][2*C] per-row-lane, per-channel sum | sumsq (no atomics: deterministic)
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int rows_per = (HW + chunks - 1) / chunks;
  const int r0 = chunk * rows_per;
  const int r1 = min(HW, r0 + rows_per);
  const GnGrid g = gn_grid(C, V);
  const int tx = threadIdx.x % g.tx_n, ty = threadIdx.x / g.tx_n;
  const T* base = x + (int64_t)b * HW * ldx;
  griddep_wait();                  // (PDL launch: x comes from the previous kernel of the stream)
  if (ty < g.ty_n) {
    for (int ps = 0; ps < g.passes; ++ps) {
      const int j = tx + ps * g.tx_n;
      if (j >= g.vpr) break;
      float s[V], q[V];
#pragma unroll
      for (int k = 0; k < V; ++k) s[k] = q[k] = 0.f;
      const T* col = base + (int64_t)j * V;
      int r = r0 + ty;
      // two rows in flight per iteration
      for (; r + g.ty_n < r1; r += 2 * g.ty_n) {
        float f0[V], f1[V];
        VecIO<T>::ld(col + (int64_t)r * ldx, f0);
        VecIO<T>::ld(col + (int64_t)(r + g.ty_n) * ldx, f1);
#pragma unroll
        for (int k = 0; k < V; ++k) { s[k] += f0[k] + f1[k]; q[k] += f0[k] * f0[k] + f1[k] * f1[k]; }
      }
      if (r < r1) {
        float f0[V];
        VecIO<T>::ld(col + (int64_t)r * ldx, f0);
#pragma unroll
        for (int k = 0; k < V; ++k) { s[k] += f0[k]; q[k] += f0[k] * f0[k]; }
      }
      float* mine = sm + (size_t)ty * 2 * C;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        mine[j * V + k] = s[k];
        mine[C + j * V + k] = q[k];
      }
    }
  }
  __syncthreads();
  const int cpg = C / groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int gi = warp; gi < groups; gi += GN_THREADS / 32) {
    float a = 0.f, bq = 0.f;
    for (int c = lane; c < cpg; c += 32)
      for (int t = 0; t < g.ty_n; ++t) {          // fixed order
        a += sm[(size_t)t * 2 * C + gi * cpg + c];
        bq += sm[(size_t)t * 2 * C + C + gi * cpg + c];
      }
    a = warp_sum(a); bq = warp_sum(bq);
    if (lane == 0) {
      float* dst = partial + (((int64_t)b * GN_CHUNKS_MAX + chunk) * GN_GROUPS_MAX + gi) * 2;
      dst[0] = a; dst[1] = bq;
    }
  }
  // The last CTA of this image to finish folds the per-chunk partials (fixed order, in double) into
  // mean / rstd per group, so the apply pass starts from 64 floats instead of redoing the reduction.
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int prev = atomicAdd(&counters[b], 1u);
    s_last = (prev == (unsigned int)chunks - 1u) ? 1u : 0u;
    if (s_last) counters[b] = 0u;   // self-resetting: ready for the next launch
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (threadIdx.x < groups) {
      double su = 0.0, sq = 0.0;
      for (int ch = 0; ch < chunks; ++ch) {
        const volatile float* pp = partial + (((int64_t)b * GN_CHUNKS_MAX + ch) * GN_GROUPS_MAX + threadIdx.x) * 2;
        su += (double)pp[0]; sq += (double)pp[1];
      }
      const double n = (double)HW * (double)cpg;
      const double mean = su / n;
      double var = sq / n - mean * mean;
      if (var < 0.0) var = 0.0;
      stats[((int64_t)b * GN_GROUPS_MAX + threadIdx.x) * 2] = (float)mean;
      stats[((int64_t)b * GN_GROUPS_MAX + threadIdx.x) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
}

template <typename T, typename TO>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const T* __restrict__ x, int ldx, TO* __restrict__ out, int ldo,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ stats, int HW, int C, int groups, int act, int row_blocks) {
  constexpr int V = VecIO<T>::V;
  __shared__ float s_g[2 * GN_GROUPS_MAX];
  const int b = blockIdx.y;
  const int cpg = C / groups;
  griddep_wait();                  // (PDL launch: the statistics kernel in front has completed and flushed)
  if (threadIdx.x < 2 * groups) s_g[threadIdx.x] = stats[(int64_t)b * GN_GROUPS_MAX * 2 + threadIdx.x];
  __syncthreads();
  const GnGrid g = gn_grid(C, V);
  const int tx = threadIdx.x % g.tx_n, ty = threadIdx.x / g.tx_n;
  if (ty >= g.ty_n) return;
  const int rows_per = (HW + row_blocks - 1) / row_blocks;
  const int r0 = blockIdx.x * rows_per;
  const int r1 = min(HW, r0 + rows_per);
  for (int ps = 0; ps < g.passes; ++ps) {
    const int j = tx + ps * g.tx_n;
    if (j >= g.vpr) break;
    const int c0 = j * V;
    float mean[V], a[V], bt[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int gi = (c0 + k) / cpg;
      mean[k] = s_g[2 * gi];
      a[k] = s_g[2 * gi + 1] * gamma[c0 + k];
      bt[k] = beta[c0 + k];
    }
    const T* xc = x + (int64_t)b * HW * ldx + c0;
    TO* oc = out + (int64_t)b * HW * ldo + c0;
    int r = r0 + ty;
    for (; r + g.ty_n < r1; r += 2 * g.ty_n) {
      float f0[V], f1[V];
      VecIO<T>::ld(xc + (int64_t)r * ldx, f0);
      VecIO<T>::ld(xc + (int64_t)(r + g.ty_n) * ldx, f1);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float y0 = (f0[k] - mean[k]) * a[k] + bt[k];
        float y1 = (f1[k] - mean[k]) * a[k] + bt[k];
        if (act == PD_ACT_SILU) {
          y0 = sizeof(TO) == 4 ? silu_acc(y0) : silu_f(y0);
          y1 = sizeof(TO) == 4 ? silu_acc(y1) : silu_f(y1);
        }
        f0[k] = y0; f1[k] = y1;
      }
      if constexpr (sizeof(T) == sizeof(TO)) {
        VecIO<TO>::st(oc + (int64_t)r * ldo, f0);
        VecIO<TO>::st(oc + (int64_t)(r + g.ty_n) * ldo, f1);
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          Dt<TO>::st(oc + (int64_t)r * ldo + k, f0[k]);
          Dt<TO>::st(oc + (int64_t)(r + g.ty_n) * ldo + k, f1[k]);
        }
      }
    }
    if (r < r1) {
      float f0[V];
      VecIO<T>::ld(xc + (int64_t)r * ldx, f0);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float y0 = (f0[k] - mean[k]) * a[k] + bt[k];
        if (act == PD_ACT_SILU) y0 = sizeof(TO) == 4 ? silu_acc(y0) : silu_f(y0);
        f0[k] = y0;
      }
      if constexpr (sizeof(T) == sizeof(TO)) {
        VecIO<TO>::st(oc + (int64_t)r * ldo, f0);
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) Dt<TO>::st(oc + (int64_t)r * ldo + k, f0[k]);
      }
    }
  }
}

// ---- single-launch GroupNorm -------------------------------------------------------------------------------
// Same thread grid and arithmetic as gn_stats + gn_apply, but ONE cooperative launch: every CTA reduces its row
// chunk, publishes one (sum, sumsq) partial per group, meets the other CTAs of ITS image at a sense-reversing
// barrier in global memory (all CTAs are co-resident: cudaLaunchCooperativeKernel), folds the image's partials in
// a fixed order (deterministic) and normalises the chunk it has just read (L2-hot).  Halves the launches of the 88
// GroupNorms per step and removes the serial "last CTA folds" tail of the two-kernel form.
#ifndef GN_REVERSE_APPLY
#define GN_REVERSE_APPLY 0     // 1: apply pass walks its chunk backwards (L2 reuse): 59 -> 55 us at 16 x 4096 x 640 alone, +0.1 ms in the step (profiles/r02_gn_reverse.txt): off
#endif
template <typename T, typename TO>
__global__ void __launch_bounds__(GN_THREADS)
gn_fused_kernel(const T* __restrict__ x, int ldx, TO* __restrict__ out, int ldo, const float* __restrict__ gamma,
                const float* __restrict__ beta, float* __restrict__ partial, int HW, int C, int groups, int chunks,
                float eps, int act, unsigned int* __restrict__ sync_words) {
  constexpr int V = VecIO<T>::V;
  extern __shared__ float sm[];  // [TY][2*C] per-row-lane, per-channel sum | sumsq
  __shared__ float s_g[2 * GN_GROUPS_MAX];
  __shared__ unsigned int s_gen;
  const int chunk = blockIdx.x, b = blockIdx.y;
  unsigned int* count = sync_words + 2 * b;
  unsigned int* flag = sync_words + 2 * b + 1;
  if (threadIdx.x == 0) s_gen = *reinterpret_cast<volatile unsigned int*>(flag);   // read BEFORE arriving
  const int rows_per = (HW + chunks - 1) / chunks;
  const int r0 = chunk * rows_per;
  const int r1 = min(HW, r0 + rows_per);
  const GnGrid g = gn_grid(C, V);
  const int tx = threadIdx.x % g.tx_n, ty = threadIdx.x / g.tx_n;
  const T* base = x + (int64_t)b * HW * ldx;
  if (ty < g.ty_n) {
    for (int ps = 0; ps < g.passes; ++ps) {
      const int j = tx + ps * g.tx_n;
      if (j >= g.vpr) break;
      float s[V], q[V];
#pragma unroll
      for (int k = 0; k < V; ++k) s[k] = q[k] = 0.f;
      const T* col = base + (int64_t)j * V;
      int r = r0 + ty;
#if GN_ROWS_IN_FLIGHT >= 8
      if constexpr (sizeof(T) == 2) {
        for (; r + 7 * g.ty_n < r1; r += 8 * g.ty_n) {     // eight rows in flight (raw 16-byte vectors: 32 registers)
          uint4 raw[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) raw[u] = *reinterpret_cast<const uint4*>(col + (int64_t)(r + u * g.ty_n) * ldx);
#pragma unroll
          for (int u = 0; u < 8; u += 2) {
            float f0[V], f1[V];
            unpack8(*reinterpret_cast<const bf16x8*>(&raw[u]), f0);
            unpack8(*reinterpret_cast<const bf16x8*>(&raw[u + 1]), f1);
#pragma unroll
            for (int k = 0; k < V; ++k) { s[k] += f0[k] + f1[k]; q[k] += f0[k] * f0[k] + f1[k] * f1[k]; }
          }
        }
      }
#endif
      for (; r + 3 * g.ty_n < r1; r += 4 * g.ty_n) {       // four rows in flight
        float f0[V], f1[V], f2[V], f3[V];
        VecIO<T>::ld(col + (int64_t)r * ldx, f0);
        VecIO<T>::ld(col + (int64_t)(r + g.ty_n) * ldx, f1);
        VecIO<T>::ld(col + (int64_t)(r + 2 * g.ty_n) * ldx, f2);
        VecIO<T>::ld(col + (int64_t)(r + 3 * g.ty_n) * ldx, f3);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          s[k] += (f0[k] + f1[k]) + (f2[k] + f3[k]);
          q[k] += (f0[k] * f0[k] + f1[k] * f1[k]) + (f2[k] * f2[k] + f3[k] * f3[k]);
        }
      }
      for (; r < r1; r += g.ty_n) {
        float f0[V];
        VecIO<T>::ld(col + (int64_t)r * ldx, f0);
#pragma unroll
        for (int k = 0; k < V; ++k) { s[k] += f0[k]; q[k] += f0[k] * f0[k]; }
      }
      float* mine = sm + (size_t)ty * 2 * C;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        mine[j * V + k] = s[k];
        mine[C + j * V + k] = q[k];
      }
    }
  }
  __syncthreads();
  const int cpg = C / groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int gi = warp; gi < groups; gi += GN_THREADS / 32) {
    float a = 0.f, bq = 0.f;
    for (int c = lane; c < cpg; c += 32)
      for (int t = 0; t < g.ty_n; ++t) {          // fixed order
        a += sm[(size_t)t * 2 * C + gi * cpg + c];
        bq += sm[(size_t)t * 2 * C + C + gi * cpg + c];
      }
    a = warp_sum(a); bq = warp_sum(bq);
    if (lane == 0) {
      float* dst = partial + (((int64_t)b * GN_CHUNKS_MAX + chunk) * GN_GROUPS_MAX + gi) * 2;
      dst[0] = a; dst[1] = bq;
    }
  }
  // ---- barrier among the `chunks` CTAs of image b ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int gen = s_gen;
    const unsigned int prev = atomicAdd(count, 1u);
    if (prev == (unsigned int)chunks - 1u) {
      *count = 0u;                                 // self-resetting: ready for the next launch
      __threadfence();
      atomicExch(flag, gen + 1u);
    } else {
      const long long t0 = clock64();
      while (*reinterpret_cast<volatile unsigned int*>(flag) == gen) {
        __nanosleep(40);
        if (clock64() - t0 > 4000000000LL) { printf("pd_b200 gn_fused: barrier timeout image %d chunk %d\n", b, chunk); __trap(); }
      }
    }
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    double su = 0.0, sq = 0.0;
    for (int ch = 0; ch < chunks; ++ch) {          // fixed order, in double
      const volatile float* pp = partial + (((int64_t)b * GN_CHUNKS_MAX + ch) * GN_GROUPS_MAX + threadIdx.x) * 2;
      su += (double)pp[0]; sq += (double)pp[1];
    }
    const double n = (double)HW * (double)cpg;
    const double mean = su / n;
    double var = sq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_g[2 * threadIdx.x] = (float)mean;
    s_g[2 * threadIdx.x + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  if (ty >= g.ty_n) return;
  for (int ps = 0; ps < g.passes; ++ps) {
    const int j = tx + ps * g.tx_n;
    if (j >= g.vpr) break;
    const int c0 = j * V;
    float a[V], bt[V];                     // y = x * a + bt with a = rstd * gamma, bt = beta - mean * a
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int gi = (c0 + k) / cpg;
      a[k] = s_g[2 * gi + 1] * gamma[c0 + k];
      bt[k] = fmaf(-s_g[2 * gi], a[k], beta[c0 + k]);
    }
    const T* xc = base + c0;
    TO* oc = out + (int64_t)b * HW * ldo + c0;
    int r = r0 + ty;
#if GN_ROWS_IN_FLIGHT >= 8
    if constexpr (sizeof(T) == 2 && sizeof(TO) == 2) {
      for (; r + 3 * g.ty_n < r1; r += 4 * g.ty_n) {       // four rows in flight in the apply pass
        uint4 raw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) raw[u] = *reinterpret_cast<const uint4*>(xc + (int64_t)(r + u * g.ty_n) * ldx);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f0[V];
          unpack8(*reinterpret_cast<const bf16x8*>(&raw[u]), f0);
#pragma unroll
          for (int k = 0; k < V; ++k) {
            float y0 = fmaf(f0[k], a[k], bt[k]);
            if (act == PD_ACT_SILU) y0 = silu_f(y0);
            f0[k] = y0;
          }
          VecIO<TO>::st(oc + (int64_t)(r + u * g.ty_n) * ldo, f0);
        }
      }
    }
#endif
#if GN_REVERSE_APPLY
    // The apply pass walks the chunk BACKWARDS: the rows the statistics pass read last are the ones most likely still in
    // L2 (a 16 x 4096 x 640 tensor is 84 MB: walked in the same order, every row had been evicted by the time it was
    // read again — ncu: 8.5 % L2 hits).  Per-element arithmetic: the order changes no bits.
    const int r_first = r;
    const int n_it = r < r1 ? (r1 - r + 2 * g.ty_n - 1) / (2 * g.ty_n) : 0;
    for (int it = n_it - 1; it >= 0; --it) {
      r = r_first + it * 2 * g.ty_n;
#else
    for (; r < r1; r += 2 * g.ty_n) {
#endif
      const bool two = r + g.ty_n < r1;
      float f0[V], f1[V];
      VecIO<T>::ld(xc + (int64_t)r * ldx, f0);
      if (two) VecIO<T>::ld(xc + (int64_t)(r + g.ty_n) * ldx, f1);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float y0 = fmaf(f0[k], a[k], bt[k]);
        float y1 = two ? fmaf(f1[k], a[k], bt[k]) : 0.f;
        if (act == PD_ACT_SILU) {
          y0 = sizeof(TO) == 4 ? silu_acc(y0) : silu_f(y0);
          y1 = sizeof(TO) == 4 ? silu_acc(y1) : silu_f(y1);
        }
        f0[k] = y0; f1[k] = y1;
      }
      if constexpr (sizeof(T) == sizeof(TO)) {
        VecIO<TO>::st(oc + (int64_t)r * ldo, f0);
        if (two) VecIO<TO>::st(oc + (int64_t)(r + g.ty_n) * ldo, f1);
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          Dt<TO>::st(oc + (int64_t)r * ldo + k, f0[k]);
          if (two) Dt<TO>::st(oc + (int64_t)(r + g.ty_n) * ldo + k, f1[k]);
        }
      }
    }
  }
}


// ---- GroupNorm from producer statistics ------------------------------------------------------------------------
// The GEMM that produced x already emitted (sum, sum of squares) per (64-pixel record, channel) from its epilogue
// (pd_conv_params.gn_stats_out), so the tensor is streamed ONCE here: a tiny finalize kernel folds the records of
// every (image, group) in a fixed order (double accumulation -> deterministic), then the apply kernel normalises.
__global__ void __launch_bounds__(128)
gn_finalize_kernel(const float* __restrict__ colstats, int ld, int rpi, int C, int groups, int HW, float eps,
                   float* __restrict__ stats) {
  __shared__ double sh_s[128], sh_q[128];
  griddep_wait();
  const int g = blockIdx.x, b = blockIdx.y;
  const int cpg = C / groups;
  const int items = rpi * cpg;
  const float2* base = reinterpret_cast<const float2*>(colstats) + (int64_t)b * rpi * ld + g * cpg;
  double su = 0.0, sq = 0.0;
  for (int i = threadIdx.x; i < items; i += 128) {
    const int rec = i / cpg, c = i - rec * cpg;
    const float2 v = __ldg(base + (int64_t)rec * ld + c);
    su += (double)v.x; sq += (double)v.y;
  }
  sh_s[threadIdx.x] = su; sh_q[threadIdx.x] = sq;
  __syncthreads();
#pragma unroll
  for (int o = 64; o > 0; o >>= 1) {                 // fixed tree: the result does not depend on scheduling
    if (threadIdx.x < o) { sh_s[threadIdx.x] += sh_s[threadIdx.x + o]; sh_q[threadIdx.x] += sh_q[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double n = (double)HW * (double)cpg;
    const double mean = sh_s[0] / n;
    double var = sh_q[0] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[((int64_t)b * GN_GROUPS_MAX + g) * 2] = (float)mean;
    stats[((int64_t)b * GN_GROUPS_MAX + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// y = act(x * a + bt), a = rstd * gamma, bt = beta - mean * a; bf16 in, bf16 out, four rows in flight per thread.
// SPLIT: y is written as a (hi, lo) pair of bf16 tensors, hi = bf16(y) at column c and lo = bf16(y - hi) at column
// C + c: a consumer GEMM over the 2C columns sees y to ~16 mantissa bits (used in front of the UNet's `out` conv,
// whose operand rounding alone is a quarter of the bf16-mode eps error).
template <bool SPLIT>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_stream_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ out, int ldo, const float* __restrict__ gamma,
                       const float* __restrict__ beta, const float* __restrict__ stats, int HW, int C, int groups, int act,
                       int row_blocks) {
  constexpr int V = 8;
  __shared__ float s_g[2 * GN_GROUPS_MAX];
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const GnGrid g = gn_grid(C, V);
  const int tx = threadIdx.x % g.tx_n, ty = threadIdx.x / g.tx_n;
  griddep_wait();
  if (threadIdx.x < 2 * groups) s_g[threadIdx.x] = stats[(int64_t)b * GN_GROUPS_MAX * 2 + threadIdx.x];
  __syncthreads();
  if (ty >= g.ty_n) return;
  const int rows_per = (HW + row_blocks - 1) / row_blocks;
  const int r0 = blockIdx.x * rows_per;
  const int r1 = min(HW, r0 + rows_per);
  for (int ps = 0; ps < g.passes; ++ps) {
    const int j = tx + ps * g.tx_n;
    if (j >= g.vpr) break;
    const int c0 = j * V;
    float a[V], bt[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int gi = (c0 + k) / cpg;
      a[k] = s_g[2 * gi + 1] * gamma[c0 + k];
      bt[k] = fmaf(-s_g[2 * gi], a[k], beta[c0 + k]);
    }
    const bf16* xc = x + (int64_t)b * HW * ldx + c0;
    bf16* oc = out + (int64_t)b * HW * ldo + c0;
    auto emit = [&](int r, float (&f)[V]) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float y = fmaf(f[k], a[k], bt[k]);
        if (act == PD_ACT_SILU) y = silu_f(y);
        f[k] = y;
      }
      const bf16x8 hi = pack8(f);
      *reinterpret_cast<bf16x8*>(oc + (int64_t)r * ldo) = hi;
      if (SPLIT) {
        float h[V];
        unpack8(hi, h);
#pragma unroll
        for (int k = 0; k < V; ++k) h[k] = f[k] - h[k];
        *reinterpret_cast<bf16x8*>(oc + (int64_t)r * ldo + C) = pack8(h);
      }
    };
    int r = r0 + ty;
    for (; r + 3 * g.ty_n < r1; r += 4 * g.ty_n) {
      float f0[V], f1[V], f2[V], f3[V];
      unpack8(*reinterpret_cast<const bf16x8*>(xc + (int64_t)r * ldx), f0);
      unpack8(*reinterpret_cast<const bf16x8*>(xc + (int64_t)(r + g.ty_n) * ldx), f1);
      unpack8(*reinterpret_cast<const bf16x8*>(xc + (int64_t)(r + 2 * g.ty_n) * ldx), f2);
      unpack8(*reinterpret_cast<const bf16x8*>(xc + (int64_t)(r + 3 * g.ty_n) * ldx), f3);
      emit(r, f0); emit(r + g.ty_n, f1); emit(r + 2 * g.ty_n, f2); emit(r + 3 * g.ty_n, f3);
    }
    for (; r < r1; r += g.ty_n) {
      float f0[V];
      unpack8(*reinterpret_cast<const bf16x8*>(xc + (int64_t)r * ldx), f0);
      emit(r, f0);
    }
  }
}

// ---- LayerNorm -------------------------------------------------------------------------
// One warp per row; NV = vectors per lane (compile-time), so the row sits in registers and the
// loads of a row are all issued before the first use.  C in {320, 640, 1280} -> NV in {2, 3, 5} (bf16).
template <typename T, int NV>
__global__ void __launch_bounds__(256)
layer_norm_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ldo,
                  const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows, int C,
                  float eps) {
  constexpr int V = VecIO<T>::V;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int nvec = C / V;
  float gm[NV][V], bt[NV][V];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + i * 32;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      gm[i][k] = j < nvec ? gamma[j * V + k] : 0.f;
      bt[i][k] = j < nvec ? beta[j * V + k] : 0.f;
    }
  }
  const float invC = 1.0f / (float)C;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += warps_total) {
    const T* xr = x + row * ldx;
    float f[NV][V];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = lane + i * 32;
      if (j < nvec) {
        VecIO<T>::ld(xr + j * V, f[i]);
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) f[i][k] = 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int k = 0; k < V; ++k) sum += f[i][k];
    sum = warp_sum(sum);
    const float mean = sum * invC;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = lane + i * 32;
      if (j < nvec) {
#pragma unroll
        for (int k = 0; k < V; ++k) { float d = f[i][k] - mean; var += d * d; }
      }
    }
    var = warp_sum(var) * invC;
    const float rstd = rsqrtf(var + eps);
    T* orow = out + row * ldo;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = lane + i * 32;
      if (j < nvec) {
        float y[V];
#pragma unroll
        for (int k = 0; k < V; ++k) y[k] = (f[i][k] - mean) * rstd * gm[i][k] + bt[i][k];
        VecIO<T>::st(orow + j * V, y);
      }
    }
  }
}

// Grouped-lane variant: LPR lanes share one row and each lane holds NV vectors of it (C = LPR * NV * V), so a warp
// normalises 32 / LPR rows at once with every lane busy (C = 320 in bf16 is 40 vectors: the one-warp-per-row
// kernel above leaves 24 of 64 lane slots idle and has one 640-byte row in flight per warp).  gamma / beta live in
// shared memory, which keeps the register count low enough for 4+ CTAs per SM — the latency of an HBM-bound
// kernel is covered by bytes in flight, not by issue slots.
template <typename T, int LPR, int NV>
__global__ void __launch_bounds__(256)
layer_norm_grouped_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ldo,
                          const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows, int C,
                          float eps) {
  constexpr int V = VecIO<T>::V;
  constexpr int RPW = 32 / LPR;                       // rows per warp per iteration
  extern __shared__ float s_gb[];                     // gamma[C] | beta[C]
  for (int i = threadIdx.x; i < C; i += blockDim.x) { s_gb[i] = gamma[i]; s_gb[C + i] = beta[i]; }   // weights: static
  __syncthreads();
  griddep_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, li = lane % LPR;
  const float invC = 1.0f / (float)C;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5) * RPW;
  for (int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * RPW; row0 < rows; row0 += stride) {
    const int64_t row = row0 + sub;
    const bool ok = row < rows;
    const T* xr = x + (ok ? row : 0) * ldx;
    float f[NV][V];
#pragma unroll
    for (int i = 0; i < NV; ++i) VecIO<T>::ld(xr + (li + i * LPR) * V, f[i]);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int k = 0; k < V; ++k) sum += f[i][k];
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * invC;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int k = 0; k < V; ++k) { const float d = f[i][k] - mean; var += d * d; }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var * invC + eps);
    if (ok) {
      T* orow = out + row * ldo;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c0 = (li + i * LPR) * V;
        float y[V];
#pragma unroll
        for (int k = 0; k < V; ++k) y[k] = (f[i][k] - mean) * rstd * s_gb[c0 + k] + s_gb[C + c0 + k];
        VecIO<T>::st(orow + c0, y);
      }
    }
  }
}

// Statistics-only pass of the grouped-lane LayerNorm: stats[row] = (mean, rstd).  The normalisation itself is folded
// into the consuming GEMM (weights pre-scaled by gamma, epilogue rstd * (acc - mean * colsum) + bias'), so the
// normalised tensor is never written or re-read.
template <typename T, int LPR, int NV>
__global__ void __launch_bounds__(256)
layer_norm_stats_kernel(const T* __restrict__ x, int ldx, float2* __restrict__ stats, int64_t rows, int C, float eps) {
  constexpr int V = VecIO<T>::V;
  constexpr int RPW = 32 / LPR;
  griddep_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, li = lane % LPR;
  const float invC = 1.0f / (float)C;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5) * RPW;
  for (int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * RPW; row0 < rows; row0 += stride) {
    const int64_t row = row0 + sub;
    const bool ok = row < rows;
    const T* xr = x + (ok ? row : 0) * ldx;
    float f[NV][V];
#pragma unroll
    for (int i = 0; i < NV; ++i) VecIO<T>::ld(xr + (li + i * LPR) * V, f[i]);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int k = 0; k < V; ++k) sum += f[i][k];
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * invC;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int k = 0; k < V; ++k) { const float d = f[i][k] - mean; var += d * d; }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    if (ok && li == 0) stats[row] = make_float2(mean, rsqrtf(var * invC + eps));
  }
}

template <typename T>
static int ln_stats_launch(const void* x, int ldx, float* stats, int64_t rows, int C, float eps, cudaStream_t s) {
  constexpr int V = VecIO<T>::V;
  const int nvec = C / V;
#define PD_LNS(L, N)                                                                                              \
  if (nvec == (L) * (N)) {                                                                                        \
    int64_t blocks = (rows + 8 * (32 / (L)) - 1) / (8 * (32 / (L)));                                              \
    const int64_t cap = (int64_t)num_sms() * 8;                                                                   \
    if (blocks > cap) blocks = cap;                                                                               \
    cudaError_t le = launch_pdl(layer_norm_stats_kernel<T, L, N>, dim3((unsigned)blocks), dim3(256), 0, s, 1,     \
                                (const T*)x, ldx, (float2*)stats, rows, C, eps);                                  \
    if (le != cudaSuccess) { set_error("pd_layer_norm_stats: launch failed: %s", cudaGetErrorString(le)); return (int)le; } \
    return check_launch("pd_layer_norm_stats");                                                                   \
  }
  PD_LNS(8, 5) PD_LNS(16, 5) PD_LNS(32, 5) PD_LNS(32, 10) PD_LNS(8, 1) PD_LNS(16, 1) PD_LNS(32, 1) PD_LNS(32, 2) PD_LNS(32, 4)
#undef PD_LNS
  set_error("pd_layer_norm_stats: unsupported width C=%d", C);
  return PD_ERR_UNSUPPORTED;
}


// ---- register-resident GroupNorm for the small late-stage tensors (bf16) ---------------------------------------------------
// GroupNorm groups are independent, so a CTA that owns ALL pixels of `ng` whole groups of one image needs no other CTA:
// its slice (<= 12 16-byte vectors per thread, 512 threads) is loaded ONCE into registers with every load in flight, the statistics are
// reduced inside the CTA, and the same registers are normalised and stored.  One ordinary (PDL) launch, no cooperative
// launch, no global barrier, no second read — the cooperative kernel above costs ~15 us on a 16 x 64 x 1280 tensor whose
// data would move in 1 us.  Used for HW <= 256 when a split into <= #SM slices of >= 256-byte rows fits 12 vectors per thread.
constexpr int GS_THREADS = 512, GS_VT = 12;
static int g_gn_small = -1;       // -1: read PD_B200_GN_SMALL once (default on)

__global__ void __launch_bounds__(GS_THREADS, 1)
gn_small_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ out, int ldo, const float* __restrict__ gamma,
                const float* __restrict__ beta, int HW, int cpg, int ng, float eps, int act) {
  extern __shared__ float sm[];                 // [rpp][2][width] per-row-lane channel sums | sums of squares
  __shared__ float s_g[2 * GN_GROUPS_MAX];      // (mean, rstd) of this CTA's groups
  const int b = blockIdx.y;
  const int width = ng * cpg, c0 = blockIdx.x * width;
  const int nv = width >> 3;                    // 16-byte vectors per pixel row of the slice
  const int rpp = GS_THREADS / nv;              // pixel rows covered by one pass of the thread grid
  const int vcol = threadIdx.x % nv, row0 = threadIdx.x / nv;
  const bool active = row0 < rpp;
  const bf16* xp = x + (int64_t)b * HW * ldx + c0 + vcol * 8;
  griddep_wait();
  uint4 raw[GS_VT];
#pragma unroll
  for (int i = 0; i < GS_VT; ++i) {
    const int r = row0 + i * rpp;
    raw[i] = (active && r < HW) ? *reinterpret_cast<const uint4*>(xp + (int64_t)r * ldx) : make_uint4(0u, 0u, 0u, 0u);
  }
  float su[8], sq[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) su[k] = sq[k] = 0.f;
#pragma unroll
  for (int i = 0; i < GS_VT; ++i) {             // rows past HW hold zeros: they add nothing
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(&raw[i]), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { su[k] += f[k]; sq[k] = fmaf(f[k], f[k], sq[k]); }
  }
  if (active) {
    float* mine = sm + (size_t)row0 * 2 * width + vcol * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) { mine[k] = su[k]; mine[width + k] = sq[k]; }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int gi = warp; gi < ng; gi += GS_THREADS / 32) {
    double a = 0.0, q = 0.0;
    for (int c = lane; c < cpg; c += 32)
      for (int t = 0; t < rpp; ++t) {           // fixed order: deterministic
        a += (double)sm[(size_t)t * 2 * width + gi * cpg + c];
        q += (double)sm[(size_t)t * 2 * width + width + gi * cpg + c];
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (lane == 0) {
      const double n = (double)HW * (double)cpg;
      const double mean = a / n;
      double var = q / n - mean * mean;
      if (var < 0.0) var = 0.0;
      s_g[2 * gi] = (float)mean;
      s_g[2 * gi + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) griddep_launch();
  if (!active) return;
  float sa[8], sb[8];                           // y = x * sa + sb
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = vcol * 8 + k, gi = c / cpg;
    sa[k] = s_g[2 * gi + 1] * gamma[c0 + c];
    sb[k] = fmaf(-s_g[2 * gi], sa[k], beta[c0 + c]);
  }
  bf16* op = out + (int64_t)b * HW * ldo + c0 + vcol * 8;
#pragma unroll
  for (int i = 0; i < GS_VT; ++i) {
    const int r = row0 + i * rpp;
    if (r < HW) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(&raw[i]), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float y = fmaf(f[k], sa[k], sb[k]);
        if (act == PD_ACT_SILU) y = silu_f(y);
        f[k] = y;
      }
      *reinterpret_cast<bf16x8*>(op + (int64_t)r * ldo) = pack8(f);
    }
  }
}

// groups per CTA for the register-resident kernel, 0 if the tensor does not fit its budget
static int gn_small_groups(int B, int HW, int C, int groups) {
  if (g_gn_small < 0) {
    const char* e = getenv("PD_B200_GN_SMALL");
    g_gn_small = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  if (!g_gn_small) return 0;
  const int cpg = C / groups;
  int best = 0;
  for (int ng = 1; ng <= groups; ++ng) {
    if (groups % ng != 0 || (ng * cpg) % 8 != 0) continue;
    const int nv = ng * cpg / 8;
    if (nv > GS_THREADS) break;
    const int rpp = GS_THREADS / nv;
    const int npass = (HW + rpp - 1) / rpp;
    if (npass > GS_VT) break;                   // (npass grows with ng)
    const long long ctas = (long long)(groups / ng) * B;
    // measured (profiles/r02_gn_small.txt): wins where the slice rows are >= 256 bytes and the CTAs fit ONE wave of the
    // one-CTA-per-SM grid (16 x 256 x 1280: 18.6 -> 11.0 us, 16 x 64 x 1280 / 2560: 14.9 -> 8.8 us); narrow slices and
    // multi-wave grids lose to the cooperative kernel (16 x 1024 x 1280: 31 -> 83 us)
    if (ng * cpg * 2 < 256 || ctas > num_sms()) continue;
    best = ng;                                  // the coarsest split that still fits: most work per thread
  }
  if (HW > 256) return 0;
  return best;
}

template <typename T, typename TO>
static int gn_launch(const void* x, int ldx, void* out, int ldo, const float* gamma, const float* beta,
                     float* partial, int B, int HW, int C, int groups, float eps, int act, cudaStream_t s) {
  if constexpr (sizeof(T) == 2 && sizeof(TO) == 2) {
    const int ng = (ldx % 8 == 0 && ldo % 8 == 0 && ((uintptr_t)out % 16) == 0) ? gn_small_groups(B, HW, C, groups) : 0;
    if (ng > 0) {
      const int cpg = C / groups, width = ng * cpg, rpp = GS_THREADS / (width / 8);
      const size_t smem = (size_t)rpp * 2 * width * sizeof(float);      // <= 512 / nv * 2 * 8 nv * 4 = 32 KiB
      cudaError_t le = launch_pdl(gn_small_kernel, dim3((unsigned)(groups / ng), (unsigned)B), dim3(GS_THREADS), smem, s, 1,
                                  (const bf16*)x, ldx, (bf16*)out, ldo, gamma, beta, HW, cpg, ng, eps, act);
      if (le != cudaSuccess) { set_error("pd_group_norm: launch failed: %s", cudaGetErrorString(le)); return (int)le; }
      return check_launch("gn_small");
    }
  }
  int chunks = gn_num_chunks(HW);
  constexpr int Vv = VecIO<T>::V;
  const GnGrid gg = gn_grid(C, Vv);
  const size_t st_smem = (size_t)gg.ty_n * 2 * C * sizeof(float);   // <= 2*V*512*4 = 32 KiB
  // scratch layout: [B][64 chunks][32 groups][2] partials | [B][32][2] mean,rstd | [B] arrival counters (zeroed once)
  //                 | [B][2] barrier words of the fused kernel (zeroed once)
  float* stats = partial + (int64_t)B * GN_CHUNKS_MAX * GN_GROUPS_MAX * 2;
  unsigned int* counters = reinterpret_cast<unsigned int*>(stats + (int64_t)B * GN_GROUPS_MAX * 2);
  unsigned int* sync_words = counters + B;
  {
    // single cooperative launch when all CTAs can be co-resident (they meet at a barrier)
    static int per_sm = -1;
    if (per_sm < 0) {
      int n = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gn_fused_kernel<T, TO>, GN_THREADS, 40 * 1024) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
      }
      per_sm = n > 2 ? 2 : n;
    }
    const int capacity = per_sm * num_sms();
    int fchunks = B > 0 ? capacity / B : 0;
    // as many CTAs per image as can be co-resident, down to ~one pass of the thread grid over a chunk: the small
    // late-stage tensors (8x8, 16x16) are latency-bound and want parallelism, not long per-thread row walks
    int want = (HW + gg.ty_n - 1) / gg.ty_n;
    if (want > GN_CHUNKS_MAX) want = GN_CHUNKS_MAX;
    if (want < 1) want = 1;
    if (fchunks > want) fchunks = want;
    if (sizeof(T) == 4 && g_gn_fused && want <= capacity) {
      // fp32 mode (the accuracy anchor): the row partition — and with it the fp32 summation order — must not depend on
      // the batch size, so that a sharded run is bit-identical to the single-GPU one (parallel.sample_sharded).
      // Always `want` chunks per image; when the whole batch cannot be co-resident, launch it a few images at a time.
      const int ipl = capacity / want;                       // images per launch
      for (int b0 = 0; b0 < B; b0 += ipl) {
        const int nb = B - b0 < ipl ? B - b0 : ipl;
        const T* xp = (const T*)x + (int64_t)b0 * HW * ldx; TO* op = (TO*)out + (int64_t)b0 * HW * ldo;
        float* pp = partial + (int64_t)b0 * GN_CHUNKS_MAX * GN_GROUPS_MAX * 2;
        unsigned int* sw = sync_words + 2 * b0;
        void* args[] = {(void*)&xp, (void*)&ldx, (void*)&op, (void*)&ldo, (void*)&gamma, (void*)&beta, (void*)&pp,
                        (void*)&HW, (void*)&C, (void*)&groups, (void*)&want, (void*)&eps, (void*)&act, (void*)&sw};
        cudaError_t e = cudaLaunchCooperativeKernel((const void*)gn_fused_kernel<T, TO>, dim3(want, nb), dim3(GN_THREADS),
                                                    args, st_smem, s);
        if (e != cudaSuccess) { set_error("pd_group_norm: cooperative launch failed: %s", cudaGetErrorString(e)); return (int)e; }
      }
      return check_launch("gn_fused");
    }
    if (fchunks >= 1 && g_gn_fused) {
      const T* xp = (const T*)x; TO* op = (TO*)out;
      void* args[] = {(void*)&xp, (void*)&ldx, (void*)&op, (void*)&ldo, (void*)&gamma, (void*)&beta, (void*)&partial,
                      (void*)&HW, (void*)&C, (void*)&groups, (void*)&fchunks, (void*)&eps, (void*)&act, (void*)&sync_words};
      cudaError_t e = cudaLaunchCooperativeKernel((const void*)gn_fused_kernel<T, TO>, dim3(fchunks, B), dim3(GN_THREADS),
                                                  args, st_smem, s);
      if (e != cudaSuccess) { set_error("pd_group_norm: cooperative launch failed: %s", cudaGetErrorString(e)); return (int)e; }
      return check_launch("gn_fused");
    }
  }
  {
    cudaError_t le = launch_pdl(gn_stats_kernel<T>, dim3((unsigned)chunks, (unsigned)B), dim3(GN_THREADS), st_smem, s, 1,
                                (const T*)x, ldx, partial, HW, C, groups, chunks, eps, stats, counters);
    if (le != cudaSuccess) { set_error("pd_group_norm: launch failed: %s", cudaGetErrorString(le)); return (int)le; }
  }
  int rc = check_launch("gn_stats");
  if (rc) return rc;
  // apply: ~2 waves of 512-thread CTAs over (row_blocks, B)
  int row_blocks = (num_sms() * 8 + B - 1) / B;
  int max_rb = (HW + 31) / 32;
  if (row_blocks > max_rb) row_blocks = max_rb;
  if (row_blocks < 1) row_blocks = 1;
  {
    const float* stats_c = stats;
    cudaError_t le = launch_pdl(gn_apply_kernel<T, TO>, dim3((unsigned)row_blocks, (unsigned)B), dim3(GN_THREADS), 0, s, 1,
                                (const T*)x, ldx, (TO*)out, ldo, gamma, beta, stats_c, HW, C, groups, act, row_blocks);
    if (le != cudaSuccess) { set_error("pd_group_norm: launch failed: %s", cudaGetErrorString(le)); return (int)le; }
  }
  return check_launch("gn_apply");
}

template <typename T>
static int ln_launch(const void* x, int ldx, void* out, int ldo, const float* gamma, const float* beta,
                     int64_t rows, int C, float eps, cudaStream_t s) {
  constexpr int V = VecIO<T>::V;
  const int nvec = C / V;
  {
    // grouped-lane kernel for the widths of the path (5 or 10 vectors per lane)
    const size_t smem = (size_t)2 * C * sizeof(float);
#define PD_LNG(L, N)                                                                                              \
  if (nvec == (L) * (N)) {                                                                                        \
    int64_t blocks = (rows + 8 * (32 / (L)) - 1) / (8 * (32 / (L)));                                              \
    const int64_t cap = (int64_t)num_sms() * 8;                                                                   \
    if (blocks > cap) blocks = cap;                                                                               \
    cudaError_t le = launch_pdl(layer_norm_grouped_kernel<T, L, N>, dim3((unsigned)blocks), dim3(256), smem, s, 1, \
                                (const T*)x, ldx, (T*)out, ldo, gamma, beta, rows, C, eps);                       \
    if (le != cudaSuccess) { set_error("pd_layer_norm: launch failed: %s", cudaGetErrorString(le)); return (int)le; } \
    return check_launch("pd_layer_norm");                                                                         \
  }
    PD_LNG(8, 5) PD_LNG(16, 5) PD_LNG(32, 5) PD_LNG(32, 10)
#undef PD_LNG
  }
  const int nv = (nvec + 31) / 32;
  int64_t blocks = (rows + 7) / 8;
  int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
#define PD_LN(N)                                                                                              \
  if (nv <= N) {                                                                                              \
    layer_norm_kernel<T, N><<<(int)blocks, 256, 0, s>>>((const T*)x, ldx, (T*)out, ldo, gamma, beta, rows, C, \
                                                        eps);                                                 \
    return check_launch("pd_layer_norm");                                                                     \
  }
  PD_LN(1) PD_LN(2) PD_LN(3) PD_LN(5) PD_LN(6) PD_LN(10) PD_LN(12)
#undef PD_LN
  set_error("pd_layer_norm: C=%d too wide", C);
  return PD_ERR_UNSUPPORTED;
}

}  // namespace pd

using namespace pd;

extern "C" {

int64_t pd_group_norm_scratch_floats(int32_t B) {
  return (int64_t)B * GN_CHUNKS_MAX * GN_GROUPS_MAX * 2 + (int64_t)B * GN_GROUPS_MAX * 2 + B + 2 * (int64_t)B;
}

int pd_debug_group_norm_fused(int32_t on) { g_gn_fused = on != 0; return 0; }

int pd_group_norm(const void* x, int32_t ldx, void* out, int32_t ldo, const float* gamma, const float* beta,
                  float* partial, int32_t B, int32_t HW, int32_t C, int32_t groups, float eps, int32_t act,
                  int32_t dtype, int32_t out_dtype, void* stream) {
  PD_REQUIRE(x && out && gamma && beta && partial, "pd_group_norm: null pointer");
  PD_REQUIRE(B > 0 && HW > 0 && C > 0 && groups > 0 && groups <= GN_GROUPS_MAX && C % groups == 0,
             "pd_group_norm: bad geometry B=%d HW=%d C=%d groups=%d", B, HW, C, groups);
  const int V = dtype == PD_BF16 ? 8 : 4;
  const int oe = out_dtype == PD_BF16 ? 2 : 4, ie = dtype == PD_BF16 ? 2 : 4;
  PD_REQUIRE(C % V == 0 && ldx % V == 0 && ldx >= C && ldo >= C && ((uintptr_t)x % 16) == 0,
             "pd_group_norm: C and input pitch must be multiples of %d elements, input 16-byte aligned", V);
  PD_REQUIRE(ie != oe || (ldo % V == 0 && ((uintptr_t)out % 16) == 0),
             "pd_group_norm: output pitch must be a multiple of %d elements and 16-byte aligned", V);
  PD_REQUIRE((size_t)C * 2 * sizeof(float) <= 40 * 1024, "pd_group_norm: C=%d too wide (max 5120)", C);
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

int pd_group_norm_apply(const void* x, int32_t ldx, void* out, int32_t ldo, const float* gamma, const float* beta,
                        const float* colstats, int32_t stats_ld, int32_t recs_per_image, float* scratch, int32_t B,
                        int32_t HW, int32_t C, int32_t groups, float eps, int32_t act, int32_t split, void* stream) {
  PD_REQUIRE(x && out && gamma && beta && colstats && scratch, "pd_group_norm_apply: null pointer");
  PD_REQUIRE(B > 0 && HW > 0 && C > 0 && groups > 0 && groups <= GN_GROUPS_MAX && C % groups == 0 && B <= 65535,
             "pd_group_norm_apply: bad geometry B=%d HW=%d C=%d groups=%d", B, HW, C, groups);
  PD_REQUIRE(C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && ldx >= C && ldo >= (split ? 2 * C : C) &&
                 ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0,
             "pd_group_norm_apply: bf16 tensors with 16-byte aligned rows (C, pitches multiples of 8; ldo >= 2C when split)");
  PD_REQUIRE(HW % 64 == 0 && recs_per_image == HW / 64 && stats_ld >= C && ((uintptr_t)colstats % 8) == 0,
             "pd_group_norm_apply: one statistics record per 64 pixels (HW=%d, records=%d)", HW, recs_per_image);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = launch_pdl(gn_finalize_kernel, dim3((unsigned)groups, (unsigned)B), dim3(128), 0, s, 1, colstats,
                             (int)stats_ld, (int)recs_per_image, (int)C, (int)groups, (int)HW, eps, scratch);
  if (e != cudaSuccess) { set_error("pd_group_norm_apply: finalize launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  int rc = check_launch("gn_finalize");
  if (rc) return rc;
  // ~4 CTAs of 512 threads per SM over (row_blocks, B); at least one pass of the thread grid per CTA
  const GnGrid gg = gn_grid(C, 8);
  int row_blocks = (num_sms() * 4 + B - 1) / B;
  int max_rb = (HW + 4 * gg.ty_n - 1) / (4 * gg.ty_n);
  if (row_blocks > max_rb) row_blocks = max_rb;
  if (row_blocks < 1) row_blocks = 1;
  if (split)
    e = launch_pdl(gn_apply_stream_kernel<true>, dim3((unsigned)row_blocks, (unsigned)B), dim3(GN_THREADS), 0, s, 1,
                   (const bf16*)x, (int)ldx, (bf16*)out, (int)ldo, gamma, beta, (const float*)scratch, (int)HW, (int)C,
                   (int)groups, (int)act, row_blocks);
  else
    e = launch_pdl(gn_apply_stream_kernel<false>, dim3((unsigned)row_blocks, (unsigned)B), dim3(GN_THREADS), 0, s, 1,
                   (const bf16*)x, (int)ldx, (bf16*)out, (int)ldo, gamma, beta, (const float*)scratch, (int)HW, (int)C,
                   (int)groups, (int)act, row_blocks);
  if (e != cudaSuccess) { set_error("pd_group_norm_apply: apply launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  return check_launch("gn_apply_stream");
}

int pd_layer_norm(const void* x, int32_t ldx, void* out, int32_t ldo, const float* gamma, const float* beta,
                  int64_t rows, int32_t C, float eps, int32_t dtype, void* stream) {
  PD_REQUIRE(x && out && gamma && beta && rows > 0 && C > 0 && ldx >= C && ldo >= C, "pd_layer_norm: bad args");
  const int V = dtype == PD_BF16 ? 8 : 4;
  PD_REQUIRE(C % V == 0 && ldx % V == 0 && ldo % V == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0,
             "pd_layer_norm: C and pitches must be multiples of %d, 16B aligned", V);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PD_F32) return ln_launch<float>(x, ldx, out, ldo, gamma, beta, rows, C, eps, s);
  if (dtype == PD_BF16) return ln_launch<bf16>(x, ldx, out, ldo, gamma, beta, rows, C, eps, s);
  PD_REQUIRE(false, "pd_layer_norm: bad dtype %d", dtype);
}

int pd_layer_norm_stats(const void* x, int32_t ldx, float* stats, int64_t rows, int32_t C, float eps, int32_t dtype,
                        void* stream) {
  PD_REQUIRE(x && stats && rows > 0 && C > 0 && ldx >= C, "pd_layer_norm_stats: bad args");
  const int V = dtype == PD_BF16 ? 8 : 4;
  PD_REQUIRE(C % V == 0 && ldx % V == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)stats % 8) == 0,
             "pd_layer_norm_stats: C and pitch must be multiples of %d, x 16B aligned, stats 8B aligned", V);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PD_F32) return ln_stats_launch<float>(x, ldx, stats, rows, C, eps, s);
  if (dtype == PD_BF16) return ln_stats_launch<bf16>(x, ldx, stats, rows, C, eps, s);
  PD_REQUIRE(false, "pd_layer_norm_stats: bad dtype %d", dtype);
}

}  // extern "C"
