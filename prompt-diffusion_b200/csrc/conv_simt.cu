// pd_conv2d dispatcher + the SIMT (FFMA) implicit-GEMM engine + weight repack.
//
// The SIMT engine is the fp32 mode of the path (eps rel-L2 <= 1e-4 against the fp32
// reference needs fp32 storage and fp32 products — TF32/bf16 tensor-core inputs cannot
// give that) and also carries the few layers whose channel counts are too small for the
// 64-channel K blocks of the tcgen05 engine (hint-stack convs with 3/6/16/32 input
// channels, conv_in 4->320, the 320->4 `out` conv).  64x64 output tile per 256-thread CTA,
// 4x4 register micro-tile, K tiles of 16 channels of one filter tap, register-prefetch
// double buffering.  A-tile rows are gathered straight from the pixel-major activation
// (zero padding, stride 2 and the nearest-x2 upsample are index arithmetic), so no im2col
// buffer ever exists.
#include "common.cuh"

namespace pd {

constexpr int SBM = 64, SBN = 64, SBK = 16, STHREADS = 256;

template <typename T> struct Ld4;
template <> struct Ld4<float> {
  __device__ __forceinline__ static void ld(const float* p, float* f) {
    float4 t = *reinterpret_cast<const float4*>(p);
    f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
  }
  static constexpr int align_elems = 4;
};
template <> struct Ld4<bf16> {
  __device__ __forceinline__ static void ld(const bf16* p, float* f) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.x));
    float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
  static constexpr int align_elems = 4;
};

struct SimtArgs {
  const void* x; const void* x2; const void* w; const float* bias; const float* rowvec;
  const void* res; void* out;
  int B, H, W, C, C2, Cout, ksize, stride, upsample, Ho, Wo;
  int ldx, ldx2, ldr, ldo, ldrv, act;
  int Ktot, cpt0, nk0, nk1;   // channel tiles per tap (seg 0), k-tiles of seg 0 / seg 1
  int vec_a, vec_a2, vec_w;   // 4-element vector loads legal
  float alpha;
  int64_t M;
};

template <typename T, typename TO>
__global__ void __launch_bounds__(STHREADS)
conv_simt_kernel(const SimtArgs a) {
  __shared__ __align__(16) float As[2][SBK][SBM + 4];
  __shared__ __align__(16) float Bs[2][SBK][SBN + 4];

  const T* __restrict__ x = (const T*)a.x;
  const T* __restrict__ x2 = (const T*)a.x2;
  const T* __restrict__ w = (const T*)a.w;

  const int tid = threadIdx.x;
  const int lrow = tid >> 2, lq = tid & 3;          // loader mapping: 64 rows x 4 quads
  const int64_t m0 = (int64_t)blockIdx.x * SBM;
  const int n0 = blockIdx.y * SBN;

  // loader-row geometry (fixed for the whole K loop)
  const int64_t lm = m0 + lrow;
  const bool lm_ok = lm < a.M;
  int lb = 0, lyo = 0, lxo = 0;
  if (lm_ok) {
    lxo = (int)(lm % a.Wo);
    int64_t t = lm / a.Wo;
    lyo = (int)(t % a.Ho);
    lb = (int)(t / a.Ho);
  }
  const int pad = a.ksize >> 1;
  const int Hin = a.upsample ? 2 * a.H : a.H, Win = a.upsample ? 2 * a.W : a.W;
  const int ln = n0 + lrow;                          // weight row handled by this loader thread
  const bool ln_ok = ln < a.Cout;

  float ra[4], rb[4];
  auto load_tile = [&](int kt) {
    int c0, cmax, kbase;
    const T* src = nullptr;
    bool vec;
    if (kt < a.nk0) {
      int tap = kt / a.cpt0;
      c0 = (kt - tap * a.cpt0) * SBK;
      cmax = a.C;
      kbase = tap * a.C;
      vec = a.vec_a;
      if (lm_ok) {
        int dy = tap / a.ksize, dx = tap - dy * a.ksize;
        int yi = lyo * a.stride + dy - pad, xi = lxo * a.stride + dx - pad;
        if (yi >= 0 && yi < Hin && xi >= 0 && xi < Win) {
          if (a.upsample) { yi >>= 1; xi >>= 1; }
          src = x + (((int64_t)lb * a.H + yi) * a.W + xi) * a.ldx;
        }
      }
    } else {
      c0 = (kt - a.nk0) * SBK;
      cmax = a.C2;
      kbase = a.ksize * a.ksize * a.C;
      vec = a.vec_a2;
      if (lm_ok) src = x2 + lm * a.ldx2;
    }
    const int c = c0 + lq * 4;
    // A quad
    if (src != nullptr && vec && c + 3 < cmax) {
      Ld4<T>::ld(src + c, ra);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) ra[i] = (src != nullptr && c + i < cmax) ? Dt<T>::ld(src + c + i) : 0.f;
    }
    // B quad: weight row ln, k = kbase + c .. c+3 (zero beyond this segment's channel count)
    const T* wp = w + (int64_t)ln * a.Ktot + kbase + c;
    if (ln_ok && a.vec_w && c + 3 < cmax) {
      Ld4<T>::ld(wp, rb);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) rb[i] = (ln_ok && c + i < cmax) ? Dt<T>::ld(wp + i) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[buf][lq * 4 + i][lrow] = ra[i];
      Bs[buf][lq * 4 + i][lrow] = rb[i];
    }
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = a.nk0 + a.nk1;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile(kt + 1);
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      float4 av = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tile(buf ^ 1);
    __syncthreads();
  }

  // epilogue
  TO* __restrict__ out = (TO*)a.out;
  const TO* __restrict__ res = (const TO*)a.res;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
    const int b = (int)(m / ((int64_t)a.Ho * a.Wo));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.Cout) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[n];
      v *= a.alpha;
      if (a.rowvec) v += a.rowvec[(int64_t)b * a.ldrv + n];
      if (res) v += Dt<TO>::ld(res + m * a.ldr + n);
      if (a.act == PD_ACT_SILU) v = silu_acc(v);
      Dt<TO>::st(out + m * a.ldo + n, v);
    }
  }
}

int conv2d_simt(const pd_conv_params* p, cudaStream_t s) {
  SimtArgs a;
  a.x = p->x; a.x2 = p->C2 > 0 ? p->x2 : nullptr; a.w = p->w; a.bias = p->bias; a.rowvec = p->rowvec;
  a.res = p->res; a.out = p->out;
  a.B = p->B; a.H = p->H; a.W = p->W; a.C = p->C; a.C2 = p->C2; a.Cout = p->Cout;
  a.ksize = p->ksize; a.stride = p->stride; a.upsample = p->upsample;
  const int Hin = p->upsample ? 2 * p->H : p->H, Win = p->upsample ? 2 * p->W : p->W;
  const int pad = p->ksize / 2;
  a.Ho = (Hin + 2 * pad - p->ksize) / p->stride + 1;
  a.Wo = (Win + 2 * pad - p->ksize) / p->stride + 1;
  a.ldx = p->ldx; a.ldx2 = p->ldx2; a.ldr = p->ldr; a.ldo = p->ldo; a.ldrv = p->ldrv; a.act = p->act;
  a.alpha = p->alpha;
  a.Ktot = p->ksize * p->ksize * p->C + p->C2;
  a.cpt0 = (p->C + SBK - 1) / SBK;
  a.nk0 = p->ksize * p->ksize * a.cpt0;
  a.nk1 = (p->C2 + SBK - 1) / SBK;
  a.M = (int64_t)p->B * a.Ho * a.Wo;
  const size_t esz = p->dtype == PD_BF16 ? 2 : 4;
  const uintptr_t al = 4 * esz;  // bytes of a 4-element vector
  a.vec_a = (p->C % 4 == 0) && (p->ldx % 4 == 0) && ((uintptr_t)p->x % al == 0);
  a.vec_a2 = p->C2 > 0 && (p->C2 % 4 == 0) && (p->ldx2 % 4 == 0) && ((uintptr_t)p->x2 % al == 0);
  a.vec_w = (a.Ktot % 4 == 0) && (p->C % 4 == 0) && ((uintptr_t)p->w % al == 0);
  dim3 grid((unsigned)((a.M + SBM - 1) / SBM), (unsigned)((p->Cout + SBN - 1) / SBN));
  PD_REQUIRE(grid.y <= 65535, "pd_conv2d(simt): Cout too large");
  if (p->dtype == PD_F32 && p->out_dtype == PD_F32)
    conv_simt_kernel<float, float><<<grid, STHREADS, 0, s>>>(a);
  else if (p->dtype == PD_BF16 && p->out_dtype == PD_BF16)
    conv_simt_kernel<bf16, bf16><<<grid, STHREADS, 0, s>>>(a);
  else if (p->dtype == PD_BF16 && p->out_dtype == PD_F32)
    conv_simt_kernel<bf16, float><<<grid, STHREADS, 0, s>>>(a);
  else if (p->dtype == PD_F32 && p->out_dtype == PD_BF16)
    conv_simt_kernel<float, bf16><<<grid, STHREADS, 0, s>>>(a);
  else
    PD_REQUIRE(false, "pd_conv2d(simt): unsupported dtypes %d -> %d", p->dtype, p->out_dtype);
  return check_launch("conv_simt");
}

// ---- weight repack -----------------------------------------------------------------------
template <typename T>
__global__ void repack_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin,
                                     int taps, int cin_pad, int ldk, int k_offset) {
  int64_t total = (int64_t)cout * taps * cin_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % cin_pad);
    int64_t t = i / cin_pad;
    int tap = (int)(t % taps);
    int n = (int)(t / taps);
    float v = c < cin ? w[((int64_t)n * cin + c) * taps + tap] : 0.f;
    Dt<T>::st(out + (int64_t)n * ldk + k_offset + (int64_t)tap * cin_pad + c, v);
  }
}

}  // namespace pd

using namespace pd;

extern "C" {

int pd_conv2d(const pd_conv_params* p, void* stream) {
  PD_REQUIRE(p != nullptr, "pd_conv2d: null params");
  PD_REQUIRE(p->x && p->w && p->out, "pd_conv2d: null tensor");
  PD_REQUIRE(p->B > 0 && p->H > 0 && p->W > 0 && p->C > 0 && p->Cout > 0, "pd_conv2d: bad geometry");
  PD_REQUIRE(p->ksize == 1 || p->ksize == 3 || p->ksize == 2, "pd_conv2d: ksize must be 1, 3 or 2 (got %d)", p->ksize);
  const bool tc_only = p->ksize == 2 || p->ln_parts != nullptr || p->ln_parts_out != nullptr || p->gn_stats_out != nullptr ||
                       p->out_sx != 0 || p->out_sy != 0 || p->out_sb != 0;
  PD_REQUIRE(p->ksize != 2 || (p->stride == 1 && !p->upsample && (p->pad_y == 0 || p->pad_y == 1) && (p->pad_x == 0 || p->pad_x == 1)),
             "pd_conv2d: ksize 2 needs stride 1, no upsample flag and pad_y / pad_x in {0, 1}");
  PD_REQUIRE(p->stride == 1 || p->stride == 2, "pd_conv2d: stride must be 1 or 2 (got %d)", p->stride);
  PD_REQUIRE(!(p->upsample && p->stride != 1), "pd_conv2d: upsample with stride 2 is not a path op");
  PD_REQUIRE(p->C2 >= 0 && (p->C2 == 0 || p->x2 != nullptr), "pd_conv2d: C2 > 0 needs x2");
  PD_REQUIRE(p->act == PD_ACT_NONE || p->act == PD_ACT_SILU || p->act == PD_ACT_GEGLU, "pd_conv2d: bad act %d", p->act);
  PD_REQUIRE(p->ldx >= p->C && p->ldo >= (p->act == PD_ACT_GEGLU ? p->Cout / 2 : p->Cout) && (p->C2 == 0 || p->ldx2 >= p->C2) &&
                 (p->res == nullptr || p->ldr >= p->Cout) && (p->rowvec == nullptr || p->ldrv >= p->Cout),
             "pd_conv2d: pitch smaller than channel count");
  PD_REQUIRE(p->dtype == PD_F32 || p->dtype == PD_BF16, "pd_conv2d: bad dtype %d", p->dtype);
  PD_REQUIRE(p->out_dtype == PD_F32 || p->out_dtype == PD_BF16, "pd_conv2d: bad out_dtype %d", p->out_dtype);
  cudaStream_t s = (cudaStream_t)stream;
  if (p->act == PD_ACT_GEGLU && p->engine == PD_ENGINE_SIMT) {
    set_error("pd_conv2d: the GEGLU epilogue exists on the tcgen05 engine only (use pd_conv2d + pd_geglu)");
    return PD_ERR_UNSUPPORTED;
  }
  if (p->ln_stats != nullptr || p->ln_colsum != nullptr) {
    const char* why_ln = "";
    if (p->engine == PD_ENGINE_SIMT || !conv2d_tc_supported(p, &why_ln)) {
      set_error("pd_conv2d: the folded-LayerNorm epilogue exists on the tcgen05 engine only%s%s",
                p->engine == PD_ENGINE_SIMT ? "" : ": ", why_ln);
      return PD_ERR_UNSUPPORTED;
    }
  }
  if (p->w_blocked && p->engine == PD_ENGINE_SIMT) {
    set_error("pd_conv2d: k-block-major weights (w_blocked) are read by the tcgen05 engine only");
    return PD_ERR_UNSUPPORTED;
  }
  if (tc_only) {
    // phase convolutions, strided output and the statistics hand-off exist on the tcgen05 engine only: never a silent
    // fall-back to a kernel that would ignore them
    const char* why_tc = "";
    if (p->engine == PD_ENGINE_SIMT || !conv2d_tc_supported(p, &why_tc)) {
      set_error("pd_conv2d: ksize 2 / strided output / ln_parts / gn_stats_out need the tcgen05 engine%s%s",
                p->engine == PD_ENGINE_SIMT ? "" : ": ", why_tc);
      return PD_ERR_UNSUPPORTED;
    }
    return conv2d_tc(p, s);
  }
  if (p->engine == PD_ENGINE_SIMT) return conv2d_simt(p, s);
  const char* why = "";
  bool tc_ok = conv2d_tc_supported(p, &why);
  if (p->w_blocked && !tc_ok) {
    set_error("pd_conv2d: w_blocked weights but the tcgen05 engine cannot run this shape: %s", why);
    return PD_ERR_UNSUPPORTED;
  }
  if (p->act == PD_ACT_GEGLU && !tc_ok) {
    set_error("pd_conv2d: GEGLU epilogue: tcgen05 engine cannot run this shape: %s", why);
    return PD_ERR_UNSUPPORTED;
  }
  if (p->engine == PD_ENGINE_TC) {
    if (!tc_ok) {
      set_error("pd_conv2d: tcgen05 engine cannot run this shape: %s", why);
      return PD_ERR_UNSUPPORTED;
    }
    return conv2d_tc(p, s);
  }
  return tc_ok ? conv2d_tc(p, s) : conv2d_simt(p, s);
}

int pd_repack_conv_weight(const float* w_oihw, void* w_out, int32_t cout, int32_t cin, int32_t kh, int32_t kw,
                          int32_t cin_pad, int32_t ldk, int32_t k_offset, int32_t dtype, void* stream) {
  PD_REQUIRE(w_oihw && w_out && cout > 0 && cin > 0 && kh > 0 && kw > 0 && cin_pad >= cin &&
                 ldk >= k_offset + kh * kw * cin_pad && k_offset >= 0,
             "pd_repack_conv_weight: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t total = (int64_t)cout * kh * kw * cin_pad;
  int64_t blocks = (total + 255) / 256;
  int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (dtype == PD_F32)
    repack_weight_kernel<float><<<(int)blocks, 256, 0, s>>>(w_oihw, (float*)w_out, cout, cin, kh * kw, cin_pad,
                                                            ldk, k_offset);
  else if (dtype == PD_BF16)
    repack_weight_kernel<bf16><<<(int)blocks, 256, 0, s>>>(w_oihw, (bf16*)w_out, cout, cin, kh * kw, cin_pad,
                                                           ldk, k_offset);
  else
    PD_REQUIRE(false, "pd_repack_conv_weight: bad dtype %d", dtype);
  return check_launch("pd_repack_conv_weight");
}

}  // extern "C"
