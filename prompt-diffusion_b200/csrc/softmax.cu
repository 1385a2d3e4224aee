// Row softmax over a materialised score matrix: the ONE attention of the first-stage decoder
// (AttnBlock.forward, ldm/modules/diffusionmodules/model.py:176-203: single head, d = 512 channels, h*w tokens).
// With d = 512 the score GEMM is already compute-dense, so that block runs as GEMM -> this kernel -> GEMM on the
// tcgen05 engine instead of a streaming-softmax kernel; the score matrix of one image (4096 x 4096 bf16 = 32 MB)
// stays in the 126 MB L2 between the three launches.
//
// One CTA per row, the row held in registers (<= 64 elements per thread), fp32 statistics:
// out[r, j] = exp2((x[r, j] - max_j) * scale * log2 e) / sum.  HBM/L2-bound: one read + one write of the matrix.
#include "common.cuh"

namespace pd {

constexpr int SM_THREADS = 256;
constexpr int SM_MAX_VEC = 8;     // 16-byte vectors per thread: cols <= 256 * 8 * (8 bf16 | 4 fp32)

template <typename T> struct SmVec;
template <> struct SmVec<bf16> {
  static constexpr int V = 8;
  __device__ __forceinline__ static void ld(const bf16* p, float* f) { unpack8(*reinterpret_cast<const bf16x8*>(p), f); }
  __device__ __forceinline__ static void st(bf16* p, const float* f) { *reinterpret_cast<bf16x8*>(p) = pack8(f); }
};
template <> struct SmVec<float> {
  static constexpr int V = 4;
  __device__ __forceinline__ static void ld(const float* p, float* f) {
    const float4 v = *reinterpret_cast<const float4*>(p); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ __forceinline__ static void st(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};

__device__ __forceinline__ float sm_block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();                 // red[] may still be read by the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < SM_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

template <typename T, int NV>   // NV: vectors per thread (compile time so the row stays in registers)
__global__ void __launch_bounds__(SM_THREADS)
softmax_rows_kernel(const T* __restrict__ x, int64_t ldx, T* __restrict__ out, int64_t ldo, int cols, float scale_log2) {
  constexpr int V = SmVec<T>::V;
  __shared__ float red[SM_THREADS / 32];
  griddep_wait();
  const int64_t row = blockIdx.x;
  const T* xr = x + row * ldx;
  T* orow = out + row * ldo;
  float f[NV][V];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * SM_THREADS + threadIdx.x) * V;
    if (c < cols) {
      SmVec<T>::ld(xr + c, f[i]);
#pragma unroll
      for (int e = 0; e < V; ++e) mx = fmaxf(mx, f[i][e]);
    }
  }
  mx = sm_block_reduce(mx, red, true);
  const float off = mx * scale_log2;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * SM_THREADS + threadIdx.x) * V;
    if (c < cols) {
#pragma unroll
      for (int e = 0; e < V; ++e) { f[i][e] = exp2f(fmaf(f[i][e], scale_log2, -off)); sum += f[i][e]; }
    }
  }
  sum = sm_block_reduce(sum, red, false);
  const float inv = 1.0f / sum;
  if (threadIdx.x == 0) griddep_launch();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * SM_THREADS + threadIdx.x) * V;
    if (c < cols) {
#pragma unroll
      for (int e = 0; e < V; ++e) f[i][e] *= inv;
      SmVec<T>::st(orow + c, f[i]);
    }
  }
}

template <typename T>
static int softmax_rows_t(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int cols, float scale,
                          cudaStream_t s) {
  constexpr int V = SmVec<T>::V;
  const int nv = (cols + SM_THREADS * V - 1) / (SM_THREADS * V);
  const float sl2 = scale * 1.4426950408889634f;
  cudaError_t le = cudaSuccess;
#define PD_SM(NVv)                                                                                                  \
  le = launch_pdl(softmax_rows_kernel<T, NVv>, dim3((unsigned)rows), dim3(SM_THREADS), 0, s, 1, (const T*)x, ldx,     \
                  (T*)out, ldo, cols, sl2)
  if (nv <= 1) PD_SM(1); else if (nv <= 2) PD_SM(2); else if (nv <= 4) PD_SM(4); else PD_SM(8);
#undef PD_SM
  if (le != cudaSuccess) { set_error("pd_softmax_rows: launch failed: %s", cudaGetErrorString(le)); return (int)le; }
  return check_launch("pd_softmax_rows");
}

}  // namespace pd

using namespace pd;

extern "C" int pd_softmax_rows(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int32_t cols, float scale,
                               int32_t dtype, void* stream) {
  PD_REQUIRE(x && out && rows > 0 && cols > 0 && ldx >= cols && ldo >= cols, "pd_softmax_rows: bad args");
  PD_REQUIRE(rows <= 0x7fffffff, "pd_softmax_rows: too many rows");
  PD_REQUIRE(dtype == PD_F32 || dtype == PD_BF16, "pd_softmax_rows: bad dtype %d", dtype);
  const int V = dtype == PD_BF16 ? 8 : 4;
  PD_REQUIRE(cols % V == 0 && ldx % V == 0 && ldo % V == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0,
             "pd_softmax_rows: cols / pitches must be multiples of %d elements and pointers 16B aligned", V);
  PD_REQUIRE(cols <= SM_THREADS * SM_MAX_VEC * V, "pd_softmax_rows: cols=%d too wide (max %d)", cols,
             SM_THREADS * SM_MAX_VEC * V);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PD_BF16) return softmax_rows_t<bf16>(x, ldx, out, ldo, rows, cols, scale, s);
  return softmax_rows_t<float>(x, ldx, out, ldo, rows, cols, scale, s);
}
