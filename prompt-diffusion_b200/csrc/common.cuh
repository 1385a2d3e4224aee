// Shared helpers for the pd_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pd_b200.h"

namespace pd {

typedef __nv_bfloat16 bf16;

// ---- error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);  // cudaPeekAtLastError -> code + message

#define PD_REQUIRE(cond, ...)                    \
  do {                                           \
    if (!(cond)) {                               \
      pd::set_error(__VA_ARGS__);                \
      return PD_ERR_BAD_ARG;                     \
    }                                            \
  } while (0)

// ---- dtype traits ------------------------------------------------------------
template <typename T> struct Dt;
template <> struct Dt<float> {
  static constexpr int code = PD_F32;
  __device__ __forceinline__ static float ld(const float* p) { return *p; }
  __device__ __forceinline__ static void st(float* p, float v) { *p = v; }
};
template <> struct Dt<bf16> {
  static constexpr int code = PD_BF16;
  __device__ __forceinline__ static float ld(const bf16* p) { return __bfloat162float(*p); }
  __device__ __forceinline__ static void st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// bf16-path SiLU (an IEEE divide + expf here made the bf16 GroupNorm pass ~55 % issue-bound, ncu: profiles/).
// silu(x) = x * sigmoid(x) = h + h * tanh(h), h = x / 2: ONE SFU op (tanh.approx, max rel. error 2^-11 — an order
// below bf16 rounding) and two FMA-pipe ops
__device__ __forceinline__ float silu_f(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// accurate variant for the fp32 mode (expf, IEEE divide)
__device__ __forceinline__ float silu_acc(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// erf by Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7 before the approximate exp / reciprocal): two SFU ops and
// seven FMAs instead of erff()'s ~25 instructions — the bf16 GEGLU pass is otherwise ALU-bound, not HBM-bound.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = x * 0.70710678118654752440f;
  const float az = fabsf(z);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, az, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = p * t * __expf(-az * az);          // 1 - erf(|z|)
  const float erf_abs = 1.0f - e;
  return 0.5f * x * (1.0f + copysignf(erf_abs, z));
}

// GELU for the bf16 GEGLU epilogue: x * Phi(x) with Phi(x) = 0.5 (1 + tanh(x (c0 + c1 x^2 + c2 x^4))), coefficients
// fitted (minimax-weighted least squares over |x| <= 6) to the exact erf form: max |error| 1.0e-4, plus tanh.approx's
// 2^-11 relative error (<= 2.5e-4 |x|) — both well under the bf16 rounding of the product that follows (2^-9 relative).
// Five FMA-pipe ops and ONE SFU op per value against ~16 + 2 for gelu_erf_fast: the GEGLU GEMM epilogue
// (attention.py:54-56) is issue-bound on exactly this arithmetic (8 epilogue warps for 128 x 256 accumulators).
__device__ __forceinline__ float gelu_tanh_fit(float x) {
  const float x2 = x * x;
  const float u = x * fmaf(fmaf(-3.93428114e-04f, x2, 3.73541892e-02f), x2, 7.96888895e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}
#ifndef PD_GEGLU_EXACT_ERF
#define PD_GEGLU_EXACT_ERF 0
#endif
__device__ __forceinline__ float gelu_epilogue(float x) {
#if PD_GEGLU_EXACT_ERF
  return gelu_erf_fast(x);
#else
  return gelu_tanh_fit(x);
#endif
}

// 8 x bf16 <-> 8 floats through one 16-byte access
struct alignas(16) bf16x8 { __nv_bfloat162 v[4]; };
__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// entry points implemented per engine (dispatched by pd_conv2d in conv_simt.cu)
int conv2d_simt(const pd_conv_params* p, cudaStream_t s);
int conv2d_tc(const pd_conv_params* p, cudaStream_t s);       // gemm_sm100.cu
bool conv2d_tc_supported(const pd_conv_params* p, const char** why);
int attention_simt(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                   int ldo, int B, int heads, int Nq, int Nk, int d, float scale, int dtype,
                   cudaStream_t s, int causal = 0);
int attention_mma(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                  int ldo, int B, int heads, int Nq, int Nk, int d, float scale, cudaStream_t s);

bool attention_short_supported(int dtype, int d, int Nk, int ldq, int ldk, int ldv, int ldo, const void* q,
                               const void* k, const void* v, const void* out);
int attention_short(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                    int B, int heads, int Nq, int Nk, int d, float scale, cudaStream_t s, int causal = 0);
bool attention_tc_supported(int dtype, int d, int ldq, int ldk, int ldv, int ldo, const void* q, const void* k,
                            const void* v, const void* out);
int attention_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                 int heads, int Nq, int Nk, int d, float scale, cudaStream_t s);

bool attention_tc3_supported(int d, int Nq, int Nk);            // attention_tc3.cu: three query groups, 128-key tiles, d <= 40
int attention_tc3(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                  int heads, int Nq, int Nk, int d, float scale, cudaStream_t s);
// attention_xtc.cu: persistent tcgen05 kernel for a short key sequence (Nk <= 128, d <= 128): the 77-key cross-attention
bool attention_xtc_supported(int dtype, int d, int Nk, int ldq, int ldk, int ldv, int ldo, const void* q, const void* k,
                             const void* v, const void* out);
int attention_xtc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                  int heads, int Nq, int Nk, int d, float scale, cudaStream_t s);
// attention_ptc.cu: persistent form of the streaming tcgen05 kernel (d <= 64): one CTA per SM walks (batch, head, query pair) units
bool attention_ptc_supported(int d, int Nq, int Nk, int B, int heads);
int attention_ptc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                  int heads, int Nq, int Nk, int d, float scale, cudaStream_t s);
bool attention_tc4_supported(int d, int Nq, int Nk);            // attention_tc4.cu: four query groups, 64-key tiles
int attention_tc4(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                  int heads, int Nq, int Nk, int d, float scale, cudaStream_t s);

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// Hot kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel's CTAs may be
// scheduled (and run their prologue: barrier init, TMEM allocation, descriptor prefetch) while the previous
// kernel drains, instead of paying a full launch gap ~500 times per denoising step.  Every such kernel calls
// griddep_wait() before it reads or writes memory another kernel of the stream may touch (it returns once all
// prerequisite grids have completed and flushed), and griddep_launch() when a CTA is about to finish.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef PD_PDL_TRIGGER
#define PD_PDL_TRIGGER 0
#endif
__device__ __forceinline__ void griddep_launch() {
#if PD_PDL_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
bool pdl_enabled();     // PD_B200_PDL=0 switches the launch attribute off (the device instructions become no-ops)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              unsigned cluster_x, Args... args) {
  cudaLaunchConfig_t lc = {};
  lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = s;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  lc.attrs = at; lc.numAttrs = n;
  return cudaLaunchKernelEx(&lc, kernel, static_cast<KArgs>(args)...);
}

}  // namespace pd
