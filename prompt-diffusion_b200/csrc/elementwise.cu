// Error state + HBM-bound elementwise kernels of the denoising path:
// layout bridges, GEGLU gate, SiLU, nearest-x2 upsample, timestep embedding and
// the fused CFG + DDIM update.  All are coalesced, vectorised (16 B per lane where
// the layout allows) grid-stride kernels sized in multiples of the SM count.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include <cstdlib>

#include "common.cuh"

namespace pd {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("PD_B200_PDL");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  count_launch();
  return 0;
}

static inline int grid_for(int64_t work_items, int threads, int max_waves = 8) {
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)num_sms() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---- NCHW fp32 -> pixel-major -------------------------------------------------
// One thread per output pixel-channel pair; reads are strided by H*W (small C: 3,4,6),
// writes are contiguous.  Only used at the API boundary (latents: 4 ch, hints: 3/6 ch).
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ out, int ldo, int B,
                                    int C, int HW, int accumulate) {
  int64_t total = (int64_t)B * HW * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t pix = i / C;
    int64_t b = pix / HW, p = pix % HW;
    float v = x[(b * C + c) * (int64_t)HW + p];
    if (accumulate) v += Dt<T>::ld(out + pix * ldo + c);
    Dt<T>::st(out + pix * ldo + c, v);
  }
}

// NCHW fp32 -> pixel-major bf16 in split precision: columns [0,C) = hi = bf16(x), [C,2C) = lo = bf16(x - hi),
// [2C,3C) = hi again.  Against weights laid out [w_hi | w_hi | w_lo] the tensor cores then form
// x_hi w_hi + x_lo w_hi + x_hi w_lo = x w up to 2^-16 relative: the 4-channel latent enters conv_in (both nets)
// at fp32-like precision for free — its K dimension is zero-padded to 64 channels per tap anyway.
__global__ void nchw_to_nhwc_split_kernel(const float* __restrict__ x, bf16* __restrict__ out, int ldo, int B, int C, int HW) {
  int64_t total = (int64_t)B * HW * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t pix = i / C;
    int64_t b = pix / HW, p = pix % HW;
    const float v = x[(b * C + c) * (int64_t)HW + p];
    const bf16 hi = __float2bfloat16_rn(v);
    const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    bf16* o = out + pix * ldo + c;
    o[0] = hi; o[C] = lo; o[2 * C] = hi;
  }
}

__global__ void add2d_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ y, int ldy, float* __restrict__ out,
                             int ldo, int64_t rows, int cols) {
  int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cols;
    int c = (int)(i % cols);
    out[r * ldo + c] = x[r * ldx + c] + y[r * ldy + c];
  }
}

// pixel-major -> NCHW fp32 through a 32x33 smem transpose tile (coalesced both sides).
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, int ldx, float* __restrict__ out, int C,
                                    int HW, float scale) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int p = p0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < HW && c < C) ? Dt<T>::ld(x + ((int64_t)b * HW + p) * ldx + c) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int c = c0 + r, p = p0 + threadIdx.x;
    if (c < C && p < HW) out[((int64_t)b * C + c) * HW + p] = tile[threadIdx.x][r] * scale;
  }
}

template <typename TI, typename TO>
__global__ void cast2d_kernel(const TI* __restrict__ x, int ldx, TO* __restrict__ out, int ldo,
                              int64_t rows, int cols) {
  int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cols;
    int c = (int)(i % cols);
    Dt<TO>::st(out + r * ldo + c, Dt<TI>::ld(x + r * ldx + c));
  }
}

// ---- GEGLU ----------------------------------------------------------------------
// One CTA owns rows_per_cta (4 or 8) consecutive rows; thread t owns the 16-byte column vectors t, t + blockDim, ... of every
// one of them, so the loop carries no index arithmetic (no divisions) and the loads of four rows (8 x 16 B per
// thread) are in flight before the first GELU is evaluated.
__global__ void __launch_bounds__(640)
geglu_bf16_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ out, int ldo, int64_t rows, int F,
                  int rows_per_cta) {
  griddep_wait();
  const int vpr = F / 8;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int nr = (int)min((int64_t)rows_per_cta, rows - r0);
  for (int j = threadIdx.x; j < vpr; j += blockDim.x) {
    const bf16* px = x + r0 * ldx + (int64_t)j * 8;
    bf16* po = out + r0 * ldo + (int64_t)j * 8;
    for (int rr = 0; rr < nr; rr += 4) {
      bf16x8 a[4], g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (rr + u < nr) {
          a[u] = *reinterpret_cast<const bf16x8*>(px + (int64_t)(rr + u) * ldx);
          g[u] = *reinterpret_cast<const bf16x8*>(px + (int64_t)(rr + u) * ldx + F);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (rr + u < nr) {
          float fa[8], fg[8];
          unpack8(a[u], fa);
          unpack8(g[u], fg);
#pragma unroll
          for (int k = 0; k < 8; ++k) fa[k] *= gelu_erf_fast(fg[k]);
          *reinterpret_cast<bf16x8*>(po + (int64_t)(rr + u) * ldo) = pack8(fa);
        }
      }
    }
  }
}
__global__ void geglu_f32_kernel(const float* __restrict__ x, int ldx, float* __restrict__ out, int ldo,
                                 int64_t rows, int F) {
  int64_t total = rows * F;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / F;
    int j = (int)(i % F);
    out[r * ldo + j] = x[r * ldx + j] * gelu_erf(x[r * ldx + F + j]);
  }
}

template <typename T>
__global__ void silu_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    Dt<T>::st(out + i, silu_acc(Dt<T>::ld(x + i)));
}

// ---- CLIP text tower pieces (FrozenCLIPEmbedder -> CLIPTextModel, ldm/modules/encoders/modules.py:117-128) ----------
// embeddings: out[r, :] = token_embedding[ids[r], :] + position_embedding[r % L, :]   (fp32 tables -> compute dtype)
template <typename T>
__global__ void embedding_lookup_kernel(const int64_t* __restrict__ ids, const float* __restrict__ tok,
                                        const float* __restrict__ pos, T* __restrict__ out, int ldo, int64_t rows, int L,
                                        int C, int vocab) {
  const int64_t n = rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    int64_t id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);           // out-of-range ids are clamped, never read out of bounds
    Dt<T>::st(out + r * ldo + c, tok[id * C + c] + pos[(r % L) * (int64_t)C + c]);
  }
}
// quick-GELU of the CLIP MLP: x * sigmoid(1.702 x)
template <typename T>
__global__ void quick_gelu_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ldo, int64_t rows, int C) {
  const int64_t n = rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const float v = Dt<T>::ld(x + r * ldx + c);
    Dt<T>::st(out + r * ldo + c, v / (1.0f + expf(-1.702f * v)));
  }
}

// ---- nearest x2 upsample ----------------------------------------------------------
// Each thread moves one 16-byte channel vector of one OUTPUT pixel (contiguous writes,
// reads hit L1/L2 four times per source vector).
template <typename T, int V>
__global__ void upsample2x_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ldo, int B,
                                  int H, int W, int C) {
  int vpr = C / V;
  int Ho = 2 * H, Wo = 2 * W;
  int64_t total = (int64_t)B * Ho * Wo * vpr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int j = (int)(i % vpr) * V;
    int64_t pix = i / vpr;
    int xo = (int)(pix % Wo);
    int64_t t = pix / Wo;
    int yo = (int)(t % Ho);
    int64_t b = t / Ho;
    const T* src = x + ((b * H + (yo >> 1)) * W + (xo >> 1)) * (int64_t)ldx + j;
    T* dst = out + pix * ldo + j;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
  }
}
template <typename T>
__global__ void upsample2x_scalar_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ldo,
                                         int B, int H, int W, int C) {
  int Ho = 2 * H, Wo = 2 * W;
  int64_t total = (int64_t)B * Ho * Wo * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t pix = i / C;
    int xo = (int)(pix % Wo);
    int64_t t = pix / Wo;
    int yo = (int)(t % Ho);
    int64_t b = t / Ho;
    out[pix * ldo + c] = x[((b * H + (yo >> 1)) * W + (xo >> 1)) * (int64_t)ldx + c];
  }
}

// ---- timestep embedding -------------------------------------------------------------
template <typename T>
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, const float* __restrict__ freqs,
                                          T* __restrict__ out, int ldo, int B, int dim, float neg_log_period) {
  int half = dim / 2;
  int total = B * half;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int b = i / half, j = i % half;
    // util.py:165-169: freqs = exp(-ln(max_period) * j / half) in fp32; args = t.float() * freqs
    // Only B*dim/2 elements: evaluate exp/cos/sin in double and round once, so each fp32
    // intermediate is the correctly rounded value the reference's fp32 ops aim for.
    float e = __fdiv_rn(__fmul_rn(neg_log_period, (float)j), (float)half);
    float freq = freqs != nullptr ? freqs[j] : (float)exp((double)e);
    float arg = __fmul_rn((float)t[b], freq);
    Dt<T>::st(out + (int64_t)b * ldo + j, (float)cos((double)arg));
    Dt<T>::st(out + (int64_t)b * ldo + half + j, (float)sin((double)arg));
    if ((dim & 1) && j == 0) Dt<T>::st(out + (int64_t)b * ldo + dim - 1, 0.f);
  }
}

// ---- fused CFG + DDIM update ----------------------------------------------------------
// cldm/ddim_hacked.py:193,211-233 in one pass: 3 reads (e_u, e_c, x) + up to 3 writes per
// element instead of ~12 elementwise launches.  Arithmetic order follows the reference so
// fp32 results are bit-comparable (IEEE sqrt / divide, no FMA contraction across the
// reference's separate ops).
__global__ void cfg_ddim_kernel(const float* __restrict__ eu, const float* __restrict__ ec,
                                const float* __restrict__ x, const float* __restrict__ noise,
                                const float* __restrict__ coef, float* __restrict__ x_prev,
                                float* __restrict__ pred_x0, float* __restrict__ e_out, int64_t n) {
  const float a_t = coef[0], a_prev = coef[1], sigma = coef[2], s1m = coef[3], scale = coef[4],
              temp = coef[5];
  const float sqrt_at = __fsqrt_rn(a_t);
  const float sqrt_aprev = __fsqrt_rn(a_prev);
  const float dir_c = __fsqrt_rn(__fsub_rn(__fsub_rn(1.0f, a_prev), __fmul_rn(sigma, sigma)));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float c = ec[i];
    float e = c;
    if (eu != nullptr) {
      float u = eu[i];
      e = __fadd_rn(u, __fmul_rn(scale, __fsub_rn(c, u)));
    }
    float xv = x[i];
    float p0 = __fdiv_rn(__fsub_rn(xv, __fmul_rn(s1m, e)), sqrt_at);
    float dir = __fmul_rn(dir_c, e);
    float nz = (noise != nullptr) ? __fmul_rn(__fmul_rn(sigma, noise[i]), temp) : 0.0f;
    float xp = __fadd_rn(__fadd_rn(__fmul_rn(sqrt_aprev, p0), dir), nz);
    x_prev[i] = xp;
    if (pred_x0 != nullptr) pred_x0[i] = p0;
    if (e_out != nullptr) e_out[i] = e;
  }
}

}  // namespace pd

using namespace pd;

extern "C" {

const char* pd_last_error(void) { return g_err; }
int pd_abi_version(void) { return 1; }
uint64_t pd_launch_count(void) { return g_launches.load(); }

int pd_device_is_sm100(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return prop.major == 10 ? 1 : 0;
}

int pd_nchw_to_nhwc(const float* x, void* out, int32_t ldo, int32_t B, int32_t C, int32_t H, int32_t W,
                    int32_t out_dtype, int32_t accumulate, void* stream) {
  PD_REQUIRE(x && out && B > 0 && C > 0 && H > 0 && W > 0 && ldo >= C, "pd_nchw_to_nhwc: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t total = (int64_t)B * C * H * W;
  int g = grid_for(total, 256);
  if (out_dtype == PD_F32)
    nchw_to_nhwc_kernel<float><<<g, 256, 0, s>>>(x, (float*)out, ldo, B, C, H * W, accumulate);
  else if (out_dtype == PD_BF16)
    nchw_to_nhwc_kernel<bf16><<<g, 256, 0, s>>>(x, (bf16*)out, ldo, B, C, H * W, accumulate);
  else
    PD_REQUIRE(false, "pd_nchw_to_nhwc: bad dtype %d", out_dtype);
  return check_launch("pd_nchw_to_nhwc");
}

int pd_nchw_to_nhwc_split(const float* x, void* out, int32_t ldo, int32_t B, int32_t C, int32_t H, int32_t W, void* stream) {
  PD_REQUIRE(x && out && B > 0 && C > 0 && H > 0 && W > 0 && ldo >= 3 * C, "pd_nchw_to_nhwc_split: bad args (ldo >= 3C)");
  cudaStream_t s = (cudaStream_t)stream;
  nchw_to_nhwc_split_kernel<<<grid_for((int64_t)B * C * H * W, 256), 256, 0, s>>>(x, (bf16*)out, ldo, B, C, H * W);
  return check_launch("pd_nchw_to_nhwc_split");
}

int pd_add2d(const float* x, int32_t ldx, const float* y, int32_t ldy, float* out, int32_t ldo, int64_t rows, int32_t cols,
             void* stream) {
  PD_REQUIRE(x && y && out && rows > 0 && cols > 0 && ldx >= cols && ldy >= cols && ldo >= cols, "pd_add2d: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  add2d_kernel<<<grid_for(rows * cols, 256), 256, 0, s>>>(x, ldx, y, ldy, out, ldo, rows, cols);
  return check_launch("pd_add2d");
}

int pd_nhwc_to_nchw(const void* x, int32_t ldx, float* out, int32_t B, int32_t C, int32_t H, int32_t W,
                    int32_t dtype, float scale, void* stream) {
  PD_REQUIRE(x && out && B > 0 && C > 0 && H > 0 && W > 0 && ldx >= C, "pd_nhwc_to_nchw: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int HW = H * W;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
  if (dtype == PD_F32)
    nhwc_to_nchw_kernel<float><<<grid, block, 0, s>>>((const float*)x, ldx, out, C, HW, scale);
  else if (dtype == PD_BF16)
    nhwc_to_nchw_kernel<bf16><<<grid, block, 0, s>>>((const bf16*)x, ldx, out, C, HW, scale);
  else
    PD_REQUIRE(false, "pd_nhwc_to_nchw: bad dtype %d", dtype);
  return check_launch("pd_nhwc_to_nchw");
}

int pd_cast2d(const void* x, int32_t ldx, int32_t dtype, void* out, int32_t ldo, int32_t out_dtype,
              int64_t rows, int32_t cols, void* stream) {
  PD_REQUIRE(x && out && rows > 0 && cols > 0 && ldx >= cols && ldo >= cols, "pd_cast2d: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int g = grid_for(rows * cols, 256);
  if (dtype == PD_F32 && out_dtype == PD_F32)
    cast2d_kernel<float, float><<<g, 256, 0, s>>>((const float*)x, ldx, (float*)out, ldo, rows, cols);
  else if (dtype == PD_F32 && out_dtype == PD_BF16)
    cast2d_kernel<float, bf16><<<g, 256, 0, s>>>((const float*)x, ldx, (bf16*)out, ldo, rows, cols);
  else if (dtype == PD_BF16 && out_dtype == PD_F32)
    cast2d_kernel<bf16, float><<<g, 256, 0, s>>>((const bf16*)x, ldx, (float*)out, ldo, rows, cols);
  else if (dtype == PD_BF16 && out_dtype == PD_BF16)
    cast2d_kernel<bf16, bf16><<<g, 256, 0, s>>>((const bf16*)x, ldx, (bf16*)out, ldo, rows, cols);
  else
    PD_REQUIRE(false, "pd_cast2d: bad dtypes %d %d", dtype, out_dtype);
  return check_launch("pd_cast2d");
}

int pd_geglu(const void* x, int32_t ldx, void* out, int32_t ldo, int64_t rows, int32_t F, int32_t dtype,
             void* stream) {
  PD_REQUIRE(x && out && rows > 0 && F > 0 && ldx >= 2 * F && ldo >= F, "pd_geglu: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PD_BF16) {
    PD_REQUIRE(F % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && ((uintptr_t)x % 16) == 0 &&
                   ((uintptr_t)out % 16) == 0,
               "pd_geglu(bf16): F, pitches must be multiples of 8 and pointers 16B aligned");
    const int vpr = F / 8;
    int threads = (vpr + 31) / 32 * 32;
    if (threads > 640) threads = 640;
    const int rpc = rows >= (int64_t)num_sms() * 64 ? 8 : 4;      // keep >= ~8 CTAs per SM on the small-M layers
    const int64_t blocks = (rows + rpc - 1) / rpc;
    PD_REQUIRE(blocks <= 0x7fffffff, "pd_geglu: too many rows");
    cudaError_t le = launch_pdl(geglu_bf16_kernel, dim3((unsigned)blocks), dim3(threads), 0, s, 1, (const bf16*)x, ldx,
                                (bf16*)out, ldo, rows, F, rpc);
    if (le != cudaSuccess) { set_error("pd_geglu: launch failed: %s", cudaGetErrorString(le)); return (int)le; }
  } else if (dtype == PD_F32) {
    geglu_f32_kernel<<<grid_for(rows * F, 256), 256, 0, s>>>((const float*)x, ldx, (float*)out, ldo, rows, F);
  } else {
    PD_REQUIRE(false, "pd_geglu: bad dtype %d", dtype);
  }
  return check_launch("pd_geglu");
}

int pd_silu(const void* x, void* out, int64_t n, int32_t dtype, void* stream) {
  PD_REQUIRE(x && out && n > 0, "pd_silu: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PD_F32)
    silu_kernel<float><<<grid_for(n, 256), 256, 0, s>>>((const float*)x, (float*)out, n);
  else if (dtype == PD_BF16)
    silu_kernel<bf16><<<grid_for(n, 256), 256, 0, s>>>((const bf16*)x, (bf16*)out, n);
  else
    PD_REQUIRE(false, "pd_silu: bad dtype %d", dtype);
  return check_launch("pd_silu");
}

int pd_embedding_lookup(const int64_t* ids, const float* tok, const float* pos, void* out, int32_t ldo, int64_t rows,
                        int32_t L, int32_t C, int32_t vocab, int32_t dtype, void* stream) {
  PD_REQUIRE(ids && tok && pos && out && rows > 0 && L > 0 && C > 0 && vocab > 0 && ldo >= C, "pd_embedding_lookup: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PD_F32)
    embedding_lookup_kernel<float><<<grid_for(rows * C, 256), 256, 0, s>>>(ids, tok, pos, (float*)out, ldo, rows, L, C, vocab);
  else if (dtype == PD_BF16)
    embedding_lookup_kernel<bf16><<<grid_for(rows * C, 256), 256, 0, s>>>(ids, tok, pos, (bf16*)out, ldo, rows, L, C, vocab);
  else
    PD_REQUIRE(false, "pd_embedding_lookup: bad dtype %d", dtype);
  return check_launch("pd_embedding_lookup");
}

int pd_quick_gelu(const void* x, int32_t ldx, void* out, int32_t ldo, int64_t rows, int32_t C, int32_t dtype, void* stream) {
  PD_REQUIRE(x && out && rows > 0 && C > 0 && ldx >= C && ldo >= C, "pd_quick_gelu: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PD_F32)
    quick_gelu_kernel<float><<<grid_for(rows * C, 256), 256, 0, s>>>((const float*)x, ldx, (float*)out, ldo, rows, C);
  else if (dtype == PD_BF16)
    quick_gelu_kernel<bf16><<<grid_for(rows * C, 256), 256, 0, s>>>((const bf16*)x, ldx, (bf16*)out, ldo, rows, C);
  else
    PD_REQUIRE(false, "pd_quick_gelu: bad dtype %d", dtype);
  return check_launch("pd_quick_gelu");
}

int pd_upsample2x(const void* x, int32_t ldx, void* out, int32_t ldo, int32_t B, int32_t H, int32_t W,
                  int32_t C, int32_t dtype, void* stream) {
  PD_REQUIRE(x && out && B > 0 && H > 0 && W > 0 && C > 0 && ldx >= C && ldo >= C, "pd_upsample2x: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t opix = (int64_t)B * 4 * H * W;
  bool al = ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0;
  if (dtype == PD_BF16) {
    if (al && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0)
      upsample2x_kernel<bf16, 8><<<grid_for(opix * (C / 8), 256), 256, 0, s>>>((const bf16*)x, ldx, (bf16*)out,
                                                                               ldo, B, H, W, C);
    else
      upsample2x_scalar_kernel<bf16><<<grid_for(opix * C, 256), 256, 0, s>>>((const bf16*)x, ldx, (bf16*)out,
                                                                             ldo, B, H, W, C);
  } else if (dtype == PD_F32) {
    if (al && C % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0)
      upsample2x_kernel<float, 4><<<grid_for(opix * (C / 4), 256), 256, 0, s>>>((const float*)x, ldx,
                                                                                (float*)out, ldo, B, H, W, C);
    else
      upsample2x_scalar_kernel<float><<<grid_for(opix * C, 256), 256, 0, s>>>((const float*)x, ldx,
                                                                              (float*)out, ldo, B, H, W, C);
  } else {
    PD_REQUIRE(false, "pd_upsample2x: bad dtype %d", dtype);
  }
  return check_launch("pd_upsample2x");
}

int pd_timestep_embedding(const int64_t* t, const float* freqs, void* out, int32_t ldo, int32_t B, int32_t dim,
                          float max_period, int32_t out_dtype, void* stream) {
  PD_REQUIRE(t && out && B > 0 && dim >= 2 && ldo >= dim, "pd_timestep_embedding: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int g = grid_for((int64_t)B * (dim / 2), 128);
  float nlp = (float)(-log((double)max_period));  // python: -math.log(max_period) -> fp32 scalar
  if (out_dtype == PD_F32)
    timestep_embedding_kernel<float><<<g, 128, 0, s>>>(t, freqs, (float*)out, ldo, B, dim, nlp);
  else if (out_dtype == PD_BF16)
    timestep_embedding_kernel<bf16><<<g, 128, 0, s>>>(t, freqs, (bf16*)out, ldo, B, dim, nlp);
  else
    PD_REQUIRE(false, "pd_timestep_embedding: bad dtype %d", out_dtype);
  return check_launch("pd_timestep_embedding");
}

int pd_cfg_ddim_step(const float* eps_uncond, const float* eps_cond, const float* x, const float* noise,
                     const float* coef, float* x_prev, float* pred_x0, float* e_t, int64_t n, void* stream) {
  PD_REQUIRE(eps_cond && x && coef && x_prev && n > 0, "pd_cfg_ddim_step: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  cfg_ddim_kernel<<<grid_for(n, 256), 256, 0, s>>>(eps_uncond, eps_cond, x, noise, coef, x_prev, pred_x0, e_t, n);
  return check_launch("pd_cfg_ddim_step");
}

}  // extern "C"
