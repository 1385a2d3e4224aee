// bf16 attention over a SHORT key sequence (Nk <= 128): the 77-token cross-attention of every SpatialTransformer
// (CrossAttention.forward with context, ldm/modules/attention.py:163-194; 23 calls per apply_model).
//
// With 77 keys the op is 0.08 TFLOP per step but moves every query and output row once: it is bound by HBM and by
// per-CTA latency, not by the tensor pipe.  The streaming kernels (attention_tc.cu: one 132 KB / 512-TMEM-column CTA
// per SM, built for thousands of keys; attention_mma.cu: two 64-key tiles with a double buffer and online softmax)
// spend their time in set-up and serial load -> compute -> store chains (90 us at B16 h8 N4096 d40, 13 us of HBM
// traffic).  This kernel does the whole key range in one pass with light CTAs, so that many are resident per SM:
//
//   CTA = 128 queries (8 warps x 16 rows) of one (batch, head); K and V of that head (<= 128 x d) are staged once in
//   shared memory by cp.async while every thread fetches its Q operand fragments STRAIGHT from global memory into
//   registers (no Q staging, no second barrier).  S = Q K^T (m16n8k16, fp32 accumulate) stays in registers, the
//   softmax is a plain (not online) one — quad shuffles for the row max / sum — and P is re-used in place as the A
//   operand of P V.  fp32 statistics, bf16 operands, like the other tensor-core engines.
#include "common.cuh"

namespace pd {

constexpr int XBQ = 128, XTHREADS = 256;

__device__ __forceinline__ uint32_t xs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void xs_cp_async16(void* dst, const void* src, bool valid) {
  int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(xs_smem_u32(dst)), "l"(src), "r"(sz));
}
__device__ __forceinline__ void xs_ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void xs_ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void xs_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t xs_pack(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float xs_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// DP: head dim padded to a multiple of 16; NK16: key capacity in units of 16; CAUSAL: query i sees keys <= i (the
// CLIP text tower's self-attention, 77 tokens; a separate instantiation, the cross-attention code is unchanged)
#ifndef XS_MINB
#define XS_MINB 1
#endif
template <int DP, int NK16, bool CAUSAL>
__global__ void __launch_bounds__(XTHREADS, (DP <= 48 && NK16 <= 5) ? XS_MINB : 1)
attention_short_kernel(const bf16* __restrict__ q, int ldq, const bf16* __restrict__ k, int ldk,
                       const bf16* __restrict__ v, int ldv, bf16* __restrict__ out, int ldo, int Nq, int Nk, int d,
                       float scale_log2) {
  constexpr int LD = DP + 8;       // smem row pitch (elements): (DP+8)*2 B = odd multiple of 16 B -> conflict-free ldmatrix
  constexpr int KS = DP / 16;      // k16 steps of Q.K^T
  constexpr int ON = DP / 8;       // n8 tiles of O
  constexpr int NKP = NK16 * 16;   // key rows held in shared memory
  extern __shared__ __align__(16) unsigned char xs_smraw[];
  bf16* Ks = reinterpret_cast<bf16*>(xs_smraw);   // [NKP][LD]
  bf16* Vs = Ks + NKP * LD;                        // [NKP][LD]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * XBQ, h = blockIdx.y, b = blockIdx.z;
  const bf16* qb = q + ((int64_t)b * Nq) * ldq + h * d;
  const bf16* kb = k + ((int64_t)b * Nk) * ldk + h * d;
  const bf16* vb = v + ((int64_t)b * Nk) * ldv + h * d;
  const int cpr = d / 8;           // 16-byte chunks per row present in global memory
  griddep_wait();                  // q / k / v come from earlier kernels of the stream (PDL: set-up above overlapped them)

  // ---- K, V of this head -> shared memory (rows >= Nk arrive as zeros); pad columns [d, DP) zeroed by plain stores ----
  for (int i = tid; i < NKP * cpr; i += XTHREADS) {
    const int r = i / cpr, c = (i - r * cpr) * 8;
    const bool ok = r < Nk;
    const int64_t row = ok ? r : 0;
    xs_cp_async16(Ks + r * LD + c, kb + row * ldk + c, ok);
    xs_cp_async16(Vs + r * LD + c, vb + row * ldv + c, ok);
  }
  asm volatile("cp.async.commit_group;\n" ::);
  if (DP > d) {
    const int padc = DP - d;
    for (int i = tid; i < 2 * NKP * padc; i += XTHREADS) {
      const int r = i / padc, c = d + (i - r * padc);
      Ks[r * LD + c] = __float2bfloat16(0.f);      // Ks and Vs are contiguous: 2 x NKP rows
    }
  }

  // ---- Q operand fragments straight from global memory (m16n8k16 A layout: rows g, g+8; column pairs t, t+8) ----
  const int g = lane >> 2, t4 = lane & 3;
  const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;
  const bool ok0 = row0 < Nq, ok1 = row1 < Nq;
  const bf16* qr0 = qb + (int64_t)(ok0 ? row0 : 0) * ldq;
  const bf16* qr1 = qb + (int64_t)(ok1 ? row1 : 0) * ldq;
  uint32_t qa[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int c0 = ks * 16 + 2 * t4, c1 = c0 + 8;
#ifdef XS_NOQ
    qa[ks][0] = qa[ks][1] = qa[ks][2] = qa[ks][3] = 0x3c003c00u + lane; continue;
#endif
    qa[ks][0] = (ok0 && c0 < d) ? __ldg(reinterpret_cast<const uint32_t*>(qr0 + c0)) : 0u;
    qa[ks][1] = (ok1 && c0 < d) ? __ldg(reinterpret_cast<const uint32_t*>(qr1 + c0)) : 0u;
    qa[ks][2] = (ok0 && c1 < d) ? __ldg(reinterpret_cast<const uint32_t*>(qr0 + c1)) : 0u;
    qa[ks][3] = (ok1 && c1 < d) ? __ldg(reinterpret_cast<const uint32_t*>(qr1 + c1)) : 0u;
  }

  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();

  // per-lane ldmatrix address components
  const int mi = lane >> 3, lr = lane & 7;
  const int kb_row = (mi >> 1) * 8 + lr, kb_col = (mi & 1) * 8;            // K (B operand)

  // ---- S = Q K^T (16 x NKP per warp) -----------------------------------------------------------
  float s[2 * NK16][4];
#pragma unroll
  for (int j = 0; j < 2 * NK16; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int jp = 0; jp < NK16; ++jp) {  // pairs of n8 tiles (16 keys)
      uint32_t bq[4];
      xs_ldsm_x4(bq, xs_smem_u32(Ks + (jp * 16 + kb_row) * LD + ks * 16 + kb_col));
      xs_mma(s[2 * jp], qa[ks], bq[0], bq[1]);
      xs_mma(s[2 * jp + 1], qa[ks], bq[2], bq[3]);
    }
  }

  // ---- softmax over the whole key range (fp32, base 2) ---------------------------------------------
  const int kbase = 2 * t4;
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int j = 0; j < 2 * NK16; ++j) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int key = kbase + j * 8 + (e & 1);
      const bool vis = key < Nk && (!CAUSAL || key <= ((e >> 1) ? row1 : row0));
      const float val = vis ? s[j][e] * scale_log2 : -INFINITY;
      s[j][e] = val;
      mx[e >> 1] = fmaxf(mx[e >> 1], val);
    }
  }
  float l[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
  }
  uint32_t p[2 * NK16][2];
#pragma unroll
  for (int j = 0; j < 2 * NK16; ++j) {
    const float p0 = xs_ex2(s[j][0] - mx[0]), p1 = xs_ex2(s[j][1] - mx[0]);
    const float p2 = xs_ex2(s[j][2] - mx[1]), p3 = xs_ex2(s[j][3] - mx[1]);
    l[0] += p0 + p1;
    l[1] += p2 + p3;
    p[j][0] = xs_pack(p0, p1);
    p[j][1] = xs_pack(p2, p3);
  }

  // ---- O = P V -----------------------------------------------------------------------------------------
  float o[ON][4];
#pragma unroll
  for (int j = 0; j < ON; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < NK16; ++kk) {  // 16 keys per step
    const uint32_t a[4] = {p[2 * kk][0], p[2 * kk][1], p[2 * kk + 1][0], p[2 * kk + 1][1]};
#pragma unroll
    for (int jp = 0; jp < ON / 2; ++jp) {
      uint32_t bv[4];
      xs_ldsm_x4_trans(bv, xs_smem_u32(Vs + (kk * 16 + (mi & 1) * 8 + lr) * LD + jp * 16 + (mi >> 1) * 8));
      xs_mma(o[2 * jp], a, bv[0], bv[1]);
      xs_mma(o[2 * jp + 1], a, bv[2], bv[3]);
    }
  }

  // ---- normalise + store -------------------------------------------------------------------------------
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
  const float inv0 = 1.f / l[0], inv1 = 1.f / l[1];
  if (tid == 0) griddep_launch();
  bf16* ob = out + ((int64_t)b * Nq) * ldo + h * d;
#pragma unroll
  for (int j = 0; j < ON; ++j) {
    const int c = j * 8 + 2 * t4;
#ifdef XS_NOSTORE
    if (o[j][0] * inv0 != 12345.678f) continue;
#endif
    if (c < d) {
      if (ok0) *reinterpret_cast<uint32_t*>(ob + (int64_t)row0 * ldo + c) = xs_pack(o[j][0] * inv0, o[j][1] * inv0);
      if (ok1) *reinterpret_cast<uint32_t*>(ob + (int64_t)row1 * ldo + c) = xs_pack(o[j][2] * inv1, o[j][3] * inv1);
    }
  }
}

template <int DP, int NK16, bool CAUSAL>
static int launch_attn_short(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                             int ldo, int B, int heads, int Nq, int Nk, int d, float scale, cudaStream_t s) {
  constexpr int LD = DP + 8;
  const size_t smem = (size_t)2 * NK16 * 16 * LD * sizeof(bf16);
  auto kern = attention_short_kernel<DP, NK16, CAUSAL>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("attention_short: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  dim3 grid((Nq + XBQ - 1) / XBQ, heads, B);
  cudaError_t le = launch_pdl(kern, grid, dim3(XTHREADS), smem, s, 1, (const bf16*)q, ldq, (const bf16*)k, ldk,
                              (const bf16*)v, ldv, (bf16*)out, ldo, Nq, Nk, d, scale * 1.4426950408889634f);
  if (le != cudaSuccess) { set_error("attention_short: launch failed: %s", cudaGetErrorString(le)); return (int)le; }
  return check_launch("attention_short");
}

bool attention_short_supported(int dtype, int d, int Nk, int ldq, int ldk, int ldv, int ldo, const void* q,
                               const void* k, const void* v, const void* out) {
  return dtype == PD_BF16 && d % 8 == 0 && d >= 8 && d <= 80 && Nk <= 128 && ldq % 2 == 0 && ldk % 8 == 0 &&
         ldv % 8 == 0 && ldo % 2 == 0 && ((uintptr_t)q % 4) == 0 && ((uintptr_t)k % 16) == 0 &&
         ((uintptr_t)v % 16) == 0 && ((uintptr_t)out % 4) == 0;
}

int attention_short(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                    int B, int heads, int Nq, int Nk, int d, float scale, cudaStream_t s, int causal) {
  const bool wide = d > 48, longk = Nk > 80;
#define PD_XS(DPv, NKv)                                                                                                   \
  do {                                                                                                                    \
    if (causal) return launch_attn_short<DPv, NKv, true>(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s); \
    return launch_attn_short<DPv, NKv, false>(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);            \
  } while (0)
  if (!wide && !longk) PD_XS(48, 5);
  if (!wide && longk) PD_XS(48, 8);
  if (wide && !longk) PD_XS(80, 5);
  PD_XS(80, 8);
#undef PD_XS
}

}  // namespace pd
