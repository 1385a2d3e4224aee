// pd_attention dispatcher + the SIMT streaming-softmax attention engine.
//
// Replaces CrossAttention.forward's `einsum -> softmax -> einsum` (attention.py:171-193)
// without materialising the [B*heads, Nq, Nk] score tensor (8.6 GB per 64x64-level
// self-attention in the reference).  This engine is the fp32 mode (scores, softmax and
// P.V all in fp32 FFMA, expf) and the generic fallback for head dims the tensor-core
// engine does not cover.  One CTA = 64 queries of one (batch, head); keys/values stream
// through shared memory in tiles of 64 with the usual running max / running sum.
#include <cstdlib>

#include "common.cuh"

namespace pd {

constexpr int ABQ = 64, ABK = 64, ATHREADS = 256;

template <typename T, int NC>  // NC = ceil(d / 16); D = 16*NC padded head dim
__global__ void __launch_bounds__(ATHREADS)
attention_simt_kernel(const T* __restrict__ q, int ldq, const T* __restrict__ k, int ldk,
                      const T* __restrict__ v, int ldv, T* __restrict__ out, int ldo, int Nq, int Nk, int d,
                      float scale, int causal) {
  constexpr int D = 16 * NC;
  extern __shared__ float sm[];
  float* Qs = sm;                       // [ABQ][D+1]
  float* Ks = Qs + ABQ * (D + 1);       // [ABK][D+1]
  float* Vs = Ks + ABK * (D + 1);       // [ABK][D]
  float* Ps = Vs + ABK * D;             // [ABQ][ABK+1]

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int q0 = blockIdx.x * ABQ, h = blockIdx.y, b = blockIdx.z;
  const T* qb = q + ((int64_t)b * Nq) * ldq + h * d;
  const T* kb = k + ((int64_t)b * Nk) * ldk + h * d;
  const T* vb = v + ((int64_t)b * Nk) * ldv + h * d;

  for (int i = tid; i < ABQ * D; i += ATHREADS) {
    int r = i / D, c = i - r * D;
    Qs[r * (D + 1) + c] = (q0 + r < Nq && c < d) ? Dt<T>::ld(qb + (int64_t)(q0 + r) * ldq + c) : 0.f;
  }

  float m_run[4], l_run[4], o[4][NC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NC; ++j) o[i][j] = 0.f;
  }

  for (int k0 = 0; k0 < Nk; k0 += ABK) {
    __syncthreads();  // previous tile fully consumed (also orders the Q fill on iteration 0)
    for (int i = tid; i < ABK * D; i += ATHREADS) {
      int r = i / D, c = i - r * D;
      bool ok = (k0 + r < Nk) && (c < d);
      Ks[r * (D + 1) + c] = ok ? Dt<T>::ld(kb + (int64_t)(k0 + r) * ldk + c) : 0.f;
      Vs[r * D + c] = ok ? Dt<T>::ld(vb + (int64_t)(k0 + r) * ldv + c) : 0.f;
    }
    __syncthreads();

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int dd = 0; dd < D; ++dd) {
      float qa[4], ka[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qa[i] = Qs[(ty * 4 + i) * (D + 1) + dd];
#pragma unroll
      for (int j = 0; j < 4; ++j) ka[j] = Ks[(tx + 16 * j) * (D + 1) + dd];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qa[i], ka[j], s[i][j]);
    }

    float corr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = k0 + tx + 16 * j;
        // causal (CLIP text tower): query i sees keys <= i; key 0 is visible to every row, so m_run is finite from tile 0 on
        s[i][j] = (key < Nk && (!causal || key <= q0 + ty * 4 + i)) ? s[i][j] * scale : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[i], mx);  // finite: every tile holds >= 1 valid key
      corr[i] = expf(m_run[i] - m_new);         // exp(-inf) = 0 on the first tile
      m_run[i] = m_new;
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float pv = expf(s[i][j] - m_new);
        ps += pv;
        Ps[(ty * 4 + i) * (ABK + 1) + tx + 16 * j] = pv;
      }
      l_run[i] = l_run[i] * corr[i] + ps;
#pragma unroll
      for (int j = 0; j < NC; ++j) o[i][j] *= corr[i];
    }
    __syncwarp();  // a P row is produced and consumed by the same 16 lanes
#pragma unroll 4
    for (int c = 0; c < ABK; ++c) {
      float pa[4], va[NC];
#pragma unroll
      for (int i = 0; i < 4; ++i) pa[i] = Ps[(ty * 4 + i) * (ABK + 1) + c];
#pragma unroll
      for (int j = 0; j < NC; ++j) va[j] = Vs[c * D + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NC; ++j) o[i][j] = fmaf(pa[i], va[j], o[i][j]);
    }
    __syncwarp();
  }

  T* ob = out + ((int64_t)b * Nq) * ldo + h * d;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float l = l_run[i];
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
    const int r = q0 + ty * 4 + i;
    if (r < Nq) {
      const float inv = 1.0f / l;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        int c = tx + 16 * j;
        if (c < d) Dt<T>::st(ob + (int64_t)r * ldo + c, o[i][j] * inv);
      }
    }
  }
}

template <typename T, int NC>
static int launch_attn_simt(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                            int ldo, int B, int heads, int Nq, int Nk, int d, float scale, int causal, cudaStream_t s) {
  constexpr int D = 16 * NC;
  size_t smem = sizeof(float) * (ABQ * (D + 1) + ABK * (D + 1) + ABK * D + ABQ * (ABK + 1));
  auto kern = attention_simt_kernel<T, NC>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("attention_simt: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return (int)e;
  }
  dim3 grid((Nq + ABQ - 1) / ABQ, heads, B);
  kern<<<grid, ATHREADS, smem, s>>>((const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv, (T*)out, ldo, Nq, Nk,
                                    d, scale, causal);
  return check_launch("attention_simt");
}

template <typename T>
static int attn_simt_t(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                       int B, int heads, int Nq, int Nk, int d, float scale, int causal, cudaStream_t s) {
  int nc = (d + 15) / 16;
#define PD_CASE(N) \
  if (nc <= N) return launch_attn_simt<T, N>(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, causal, s)
  PD_CASE(1); PD_CASE(2); PD_CASE(3); PD_CASE(4); PD_CASE(5); PD_CASE(8); PD_CASE(10);
#undef PD_CASE
  set_error("pd_attention: head dim %d > 160 unsupported", d);
  return PD_ERR_UNSUPPORTED;
}

int attention_simt(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                   int B, int heads, int Nq, int Nk, int d, float scale, int dtype, cudaStream_t s, int causal) {
  if (dtype == PD_F32) return attn_simt_t<float>(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, causal, s);
  return attn_simt_t<bf16>(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, causal, s);
}

}  // namespace pd

using namespace pd;

extern "C" {

int pd_attention_ex(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv, void* out,
                    int32_t ldo, int32_t B, int32_t heads, int32_t Nq, int32_t Nk, int32_t d, float scale,
                    int32_t dtype, int32_t engine, void* stream) {
  PD_REQUIRE(q && k && v && out, "pd_attention: null pointer");
  PD_REQUIRE(B > 0 && heads > 0 && Nq > 0 && Nk > 0 && d > 0, "pd_attention: bad geometry");
  PD_REQUIRE(ldq >= heads * d && ldk >= heads * d && ldv >= heads * d && ldo >= heads * d,
             "pd_attention: pitch smaller than heads*d");
  PD_REQUIRE(dtype == PD_F32 || dtype == PD_BF16, "pd_attention: bad dtype %d", dtype);
  PD_REQUIRE(heads <= 65535 && B <= 65535, "pd_attention: grid too large");
  cudaStream_t s = (cudaStream_t)stream;
  const bool mma_ok = dtype == PD_BF16 && d % 8 == 0 && d <= 160 && ((d + 15) / 16 * 16 == 32 ||
                      (d + 15) / 16 * 16 == 48 || (d + 15) / 16 * 16 == 64 || (d + 15) / 16 * 16 == 80 ||
                      (d + 15) / 16 * 16 == 128 || (d + 15) / 16 * 16 == 160) &&
                      ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0 &&
                      ((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 &&
                      ((uintptr_t)out % 4) == 0;
  const bool tc_ok = attention_tc_supported(dtype, d, ldq, ldk, ldv, ldo, q, k, v, out);
  if (engine == 3 && !tc_ok) {
    set_error("pd_attention: tcgen05 engine needs sm_100, bf16, d <= 192 (multiple of 8), 16B-aligned tensors and pitches");
    return PD_ERR_UNSUPPORTED;
  }
  const bool short_ok = attention_short_supported(dtype, d, Nk, ldq, ldk, ldv, ldo, q, k, v, out);
  if (engine == 4 && !short_ok) {
    set_error("pd_attention: short-key engine needs bf16, Nk <= 128, d <= 80 (multiple of 8), 16B-aligned k/v");
    return PD_ERR_UNSUPPORTED;
  }
  static int auto_short = -1;          // PD_B200_ATTN_SHORT=0: auto never picks the short-key engine (A/B timing)
  if (auto_short < 0) { const char* e = getenv("PD_B200_ATTN_SHORT"); auto_short = (e && e[0] == '0') ? 0 : 1; }
  // engine 7 = the persistent tcgen05 short-key kernel (Nk <= 128, d <= 128); auto picks it for the 77-key cross-attention
  // once there are enough 256-query units to fill the machine (PD_B200_ATTN_XTC=0: never)
  const bool xtc_ok = attention_xtc_supported(dtype, d, Nk, ldq, ldk, ldv, ldo, q, k, v, out);
  if (engine == 7 && !xtc_ok) {
    set_error("pd_attention: the short-key tcgen05 engine needs sm_100, bf16, Nk <= 128, d <= 128 (multiple of 8), 16B-aligned tensors and pitches");
    return PD_ERR_UNSUPPORTED;
  }
  static int auto_xtc = -1;
  if (auto_xtc < 0) { const char* e = getenv("PD_B200_ATTN_XTC"); auto_xtc = (e && e[0] == '0') ? 0 : 1; }
  if (engine == 7 || (engine == 0 && xtc_ok && auto_xtc && (long long)B * heads * ((Nq + 255) / 256) >= 296))
    return attention_xtc(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);
  if (engine == 4 || (engine == 0 && short_ok && auto_short))
    return attention_short(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);
  // engine 5 = the four-group / 64-key-tile tcgen05 kernel (d <= 64); auto picks it for the long sequences it is built for
  if (engine == 5 && !(tc_ok && d <= 64)) {
    set_error("pd_attention: the four-group tcgen05 engine needs what engine 3 needs and d <= 64");
    return PD_ERR_UNSUPPORTED;
  }
  // engine 6 = the three-group / 128-key-tile tcgen05 kernel (d <= 40)
  if (engine == 6 && !(tc_ok && d <= 40)) {
    set_error("pd_attention: the three-group tcgen05 engine needs what engine 3 needs and d <= 40");
    return PD_ERR_UNSUPPORTED;
  }
  // engine 8 = the persistent two-group tcgen05 kernel (d <= 64): one CTA per SM walks the (batch, head, query pair) units
  if (engine == 8 && !(tc_ok && d <= 64)) {
    set_error("pd_attention: the persistent tcgen05 engine needs what engine 3 needs and d <= 64");
    return PD_ERR_UNSUPPORTED;
  }
  if (engine == 8 || (engine == 0 && tc_ok && !attention_tc3_supported(d, Nq, Nk) && attention_ptc_supported(d, Nq, Nk, B, heads)))
    return attention_ptc(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);
  if (engine == 6 || (engine == 0 && tc_ok && attention_tc3_supported(d, Nq, Nk)))
    return attention_tc3(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);
  if (engine == 5 || (engine == 0 && tc_ok && attention_tc4_supported(d, Nq, Nk)))
    return attention_tc4(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);
  if (engine == 3 || (engine == 0 && tc_ok)) return attention_tc(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);
  if (engine == 2 && !mma_ok) {
    set_error("pd_attention: tensor-core engine needs bf16, d in {32,40,48,64,80,128,160}, 16B-aligned q/k/v");
    return PD_ERR_UNSUPPORTED;
  }
  if (engine != 1 && mma_ok) return attention_mma(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);
  return attention_simt(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, dtype, s);
}

int pd_attention_causal(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv, void* out,
                        int32_t ldo, int32_t B, int32_t heads, int32_t N, int32_t d, float scale, int32_t dtype,
                        void* stream) {
  PD_REQUIRE(q && k && v && out, "pd_attention_causal: null pointer");
  PD_REQUIRE(B > 0 && heads > 0 && N > 0 && d > 0 && d <= 160, "pd_attention_causal: bad geometry");
  PD_REQUIRE(ldq >= heads * d && ldk >= heads * d && ldv >= heads * d && ldo >= heads * d,
             "pd_attention_causal: pitch smaller than heads*d");
  PD_REQUIRE(dtype == PD_F32 || dtype == PD_BF16, "pd_attention_causal: bad dtype %d", dtype);
  PD_REQUIRE(heads <= 65535 && B <= 65535, "pd_attention_causal: grid too large");
  cudaStream_t s = (cudaStream_t)stream;
  if (attention_short_supported(dtype, d, N, ldq, ldk, ldv, ldo, q, k, v, out))
    return attention_short(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, N, N, d, scale, s, 1);
  return attention_simt(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, N, N, d, scale, dtype, s, 1);
}

int pd_attention(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv, void* out,
                 int32_t ldo, int32_t B, int32_t heads, int32_t Nq, int32_t Nk, int32_t d, float scale,
                 int32_t dtype, void* stream) {
  return pd_attention_ex(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, dtype, 0, stream);
}

}  // extern "C"
