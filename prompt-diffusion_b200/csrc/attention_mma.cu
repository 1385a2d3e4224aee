// bf16 streaming-softmax attention on warp-level tensor-core MMAs (m16n8k16, fp32 accum).
//
// Replaces CrossAttention.forward's einsum/softmax/einsum (attention.py:171-193) for the
// bf16 mode: self-attention over 4096..9216 tokens with head dims 40/80/160 and the
// 77-key cross-attention.  Scores never leave registers; K/V stream through shared memory
// with a cp.async double buffer.  One CTA = 64 queries (4 warps x 16 rows) of one
// (batch, head).  Head dim 40 is zero-padded to 48 in shared memory only.
#include "common.cuh"

namespace pd {

constexpr int MBQ = 64, MBK = 64, MTHREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(dst)), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int DP>  // padded head dim, multiple of 16
__global__ void __launch_bounds__(MTHREADS)
attention_mma_kernel(const bf16* __restrict__ q, int ldq, const bf16* __restrict__ k, int ldk,
                     const bf16* __restrict__ v, int ldv, bf16* __restrict__ out, int ldo, int Nq, int Nk,
                     int d, float scale_log2) {
  constexpr int LD = DP + 8;       // smem row pitch (elements): (DP+8)*2 B = odd multiple of 16 B
  constexpr int KS = DP / 16;      // k16 steps of Q.K^T
  constexpr int ON = DP / 8;       // n8 tiles of O
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);   // [MBQ][LD]
  bf16* Ks = Qs + MBQ * LD;                    // [2][MBK][LD]
  bf16* Vs = Ks + 2 * MBK * LD;                // [2][MBK][LD]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * MBQ, h = blockIdx.y, b = blockIdx.z;
  const bf16* qb = q + ((int64_t)b * Nq) * ldq + h * d;
  const bf16* kb = k + ((int64_t)b * Nk) * ldk + h * d;
  const bf16* vb = v + ((int64_t)b * Nk) * ldv + h * d;
  const int cpr = d / 8;           // 16-byte chunks per row actually present in global memory

  // zero the pad columns [d, DP) of every row once; cp.async never touches them
  if (DP > d) {
    const int padc = DP - d;
    for (int i = tid; i < 5 * 64 * padc; i += MTHREADS) {
      int r = i / padc, c = d + (i - r * padc);
      Qs[r * LD + c] = __float2bfloat16(0.f);  // Qs, Ks[2], Vs[2] are contiguous: 5 x 64 rows
    }
  }

  auto load_q = [&]() {
    for (int i = tid; i < MBQ * cpr; i += MTHREADS) {
      int r = i / cpr, c = (i - r * cpr) * 8;
      bool ok = q0 + r < Nq;
      cp_async16(Qs + r * LD + c, qb + (int64_t)(ok ? q0 + r : 0) * ldq + c, ok);
    }
  };
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * MBK;
    bf16* kd = Ks + buf * MBK * LD;
    bf16* vd = Vs + buf * MBK * LD;
    for (int i = tid; i < MBK * cpr; i += MTHREADS) {
      int r = i / cpr, c = (i - r * cpr) * 8;
      bool ok = k0 + r < Nk;
      int64_t row = ok ? k0 + r : 0;
      cp_async16(kd + r * LD + c, kb + row * ldk + c, ok);
      cp_async16(vd + r * LD + c, vb + row * ldv + c, ok);
    }
  };

  const int ntiles = (Nk + MBK - 1) / MBK;
  load_q();
  load_kv(0, 0);
  cp_async_commit();

  float o[ON][4];
#pragma unroll
  for (int j = 0; j < ON; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  // per-lane ldmatrix address components
  const int mi = lane >> 3, lr = lane & 7;
  const int a_row = warp * 16 + (mi & 1) * 8 + lr, a_col = (mi >> 1) * 8;   // Q (A operand) / V (trans)
  const int kb_row = (mi >> 1) * 8 + lr, kb_col = (mi & 1) * 8;            // K (B operand)

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      load_kv(t + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const bf16* kt = Ks + buf * MBK * LD;
    const bf16* vt = Vs + buf * MBK * LD;

    // ---- S = Q K^T (16 x 64 per warp) -------------------------------------------------
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t a[4];
      ldsm_x4(a, smem_u32(Qs + a_row * LD + ks * 16 + a_col));
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // pairs of n8 tiles (16 keys)
        uint32_t bq[4];
        ldsm_x4(bq, smem_u32(kt + (jp * 16 + kb_row) * LD + ks * 16 + kb_col));
        mma_bf16(s[2 * jp], a, bq[0], bq[1]);
        mma_bf16(s[2 * jp + 1], a, bq[2], bq[3]);
      }
    }

    // ---- online softmax ------------------------------------------------------------------
    const int kbase = t * MBK + 2 * (lane & 3);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int key = kbase + j * 8 + (e & 1);
        float val = key < Nk ? s[j][e] * scale_log2 : -INFINITY;
        s[j][e] = val;
        mx[e >> 1] = fmaxf(mx[e >> 1], val);
      }
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      float m_new = fmaxf(m_run[r], mx[r]);
      corr[r] = exp2f(m_run[r] - m_new);
      m_run[r] = m_new;
      l_run[r] *= corr[r];
    }
    uint32_t p[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p0 = exp2f(s[j][0] - m_run[0]), p1 = exp2f(s[j][1] - m_run[0]);
      float p2 = exp2f(s[j][2] - m_run[1]), p3 = exp2f(s[j][3] - m_run[1]);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      p[j][0] = pack_bf16(p0, p1);
      p[j][1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int j = 0; j < ON; ++j) {
      o[j][0] *= corr[0]; o[j][1] *= corr[0];
      o[j][2] *= corr[1]; o[j][3] *= corr[1];
    }

    // ---- O += P V ----------------------------------------------------------------------------
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 16 keys per step
      uint32_t a[4] = {p[2 * kk][0], p[2 * kk][1], p[2 * kk + 1][0], p[2 * kk + 1][1]};
#pragma unroll
      for (int jp = 0; jp < ON / 2; ++jp) {
        uint32_t bv[4];
        ldsm_x4_trans(bv, smem_u32(vt + (kk * 16 + (mi & 1) * 8 + lr) * LD + jp * 16 + (mi >> 1) * 8));
        mma_bf16(o[2 * jp], a, bv[0], bv[1]);
        mma_bf16(o[2 * jp + 1], a, bv[2], bv[3]);
      }
    }
    __syncthreads();  // everyone done with `buf` before the next-next tile's cp.async lands in it
  }

  // ---- normalise + store ---------------------------------------------------------------------
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int row0 = q0 + warp * 16 + (lane >> 2), row1 = row0 + 8;
  bf16* ob = out + ((int64_t)b * Nq) * ldo + h * d;
#pragma unroll
  for (int j = 0; j < ON; ++j) {
    const int c = j * 8 + 2 * (lane & 3);
    if (c < d) {
      if (row0 < Nq)
        *reinterpret_cast<uint32_t*>(ob + (int64_t)row0 * ldo + c) = pack_bf16(o[j][0] * inv0, o[j][1] * inv0);
      if (row1 < Nq)
        *reinterpret_cast<uint32_t*>(ob + (int64_t)row1 * ldo + c) = pack_bf16(o[j][2] * inv1, o[j][3] * inv1);
    }
  }
}

template <int DP>
static int launch_attn_mma(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                           int ldo, int B, int heads, int Nq, int Nk, int d, float scale, cudaStream_t s) {
  constexpr int LD = DP + 8;
  size_t smem = (size_t)5 * 64 * LD * sizeof(bf16);
  auto kern = attention_mma_kernel<DP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("attention_mma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return (int)e;
  }
  dim3 grid((Nq + MBQ - 1) / MBQ, heads, B);
  kern<<<grid, MTHREADS, smem, s>>>((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)out,
                                    ldo, Nq, Nk, d, scale * 1.4426950408889634f);
  return check_launch("attention_mma");
}

int attention_mma(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                  int B, int heads, int Nq, int Nk, int d, float scale, cudaStream_t s) {
  if (d % 8 != 0) {
    set_error("attention_mma: head dim %d not a multiple of 8", d);
    return PD_ERR_UNSUPPORTED;
  }
  int dp = (d + 15) / 16 * 16;
  switch (dp) {
    case 32: return launch_attn_mma<32>(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);
    case 48: return launch_attn_mma<48>(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);
    case 64: return launch_attn_mma<64>(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);
    case 80: return launch_attn_mma<80>(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);
    case 128: return launch_attn_mma<128>(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);
    case 160: return launch_attn_mma<160>(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, d, scale, s);
  }
  set_error("attention_mma: padded head dim %d unsupported", dp);
  return PD_ERR_UNSUPPORTED;
}

}  // namespace pd
