// tcgen05 / TMEM / TMA streaming-softmax attention (bf16) for head dims <= 64 — the d = 40, 4096..9216-token
// self-attention of the 64x64 / 96x96 latent level and its 77-key cross-attention (CrossAttention.forward,
// ldm/modules/attention.py:171-193), which dominate the attention time of a denoising step.
//
// One CTA = 128 queries of one (batch, head); two CTAs are resident per SM (112 KiB smem, 256 TMEM columns each),
// so while one CTA's softmax warps run, the other CTA's MMAs use the tensor pipe.
//
//   warp 0   : TMA producer — Q once, then K/V tiles of 128 keys into a 2-stage ring.  The tensor maps view
//              q/k/v as (d, heads, tokens, batch); the 64-wide box over a 40-wide head makes TMA zero-fill
//              channels 40..63, so no padding pass and no padded copies exist.
//   warp 1   : MMA issuer — S[128x128] = Q K^T (both K-major, SWIZZLE_128B) into TMEM columns [0,128);
//              O[128xNP] += P V with P from shared memory (K-major) and the V tile used in place as the
//              MN-major B operand; O lives in TMEM columns [128, 128+NP).
//   warps 2-5: softmax — ONE THREAD PER QUERY ROW (TMEM lane == row): row max / sum need no shuffles.
//              Pass 1 reads S for the max, pass 2 re-reads S, exponentiates (exp2, log2e folded into the scale),
//              writes P (bf16, swizzled) to smem; O is rescaled in TMEM only when some row max moved.
//              Epilogue: O / l -> bf16 -> swizzled smem -> TMA store (clips columns >= d and rows >= Nq).
#include "tc_ptx.cuh"

namespace pd {

constexpr int FA_BQ = 128, FA_BK = 128, FA_THREADS = 192;
constexpr int FA_TILE_BYTES = 128 * 128;   // one [128 rows][64 bf16] SWIZZLE_128B tile
constexpr int FA_ALIGN_SLACK = 512;
constexpr int FA_SMEM_BYTES = 7 * FA_TILE_BYTES + 128 + FA_ALIGN_SLACK;
__device__ unsigned long long* g_fa_dbg = nullptr;   // optional phase timeline (block 0 only)

#define FA_DBG(slot, tile)                                                                       \
  do {                                                                                           \
    if (g_fa_dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tile) < 32) \
      g_fa_dbg[(slot) * 32 + (tile)] = gtimer();                                                 \
  } while (0)

struct FaArgs {
  int Nq, Nk, d;
  int kpad;            // d rounded up to 16 (K extent of Q K^T)
  int npad;            // d rounded up to 16 (N extent of P V)
  float scale_log2;
  uint32_t idesc_s_full, idesc_s_last, idesc_pv;
  int n_last_pad;      // S columns computed for the last K/V tile (multiple of 16)
  int n_last_valid;    // keys actually present in the last tile
  int ntiles;
};

__global__ void __launch_bounds__(FA_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                    const FaArgs a) {
  // No static shared memory: two CTAs must fit in one SM (2 x (7 x 16 KiB tiles + barriers + slack + 1 KiB
  // reserved) <= 228 KiB), so the barriers live behind the tiles in the dynamic allocation.
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  if (threadIdx.x == 0 && (smem - smem_raw) > FA_ALIGN_SLACK) {
    printf("attention_tc: dynamic smem base misaligned by %d bytes\n", (int)(smem - smem_raw));
    __trap();
  }
  unsigned char* q_s = smem;                              // 16 KiB  (reused as the O staging tile at the end)
  unsigned char* kv_s = smem + FA_TILE_BYTES;             // 2 stages x (K 16 KiB + V 16 KiB)
  unsigned char* p_s = smem + 5 * FA_TILE_BYTES;          // 2 x 16 KiB: keys 0..63 | keys 64..127
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * FA_TILE_BYTES);
  uint64_t& q_full = bars[0]; uint64_t& s_full = bars[1]; uint64_t& p_full = bars[2]; uint64_t& o_final = bars[3];
  uint64_t* kv_full = bars + 4; uint64_t* kv_empty = bars + 6;
  uint32_t& tmem_base_slot = *reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * FA_BQ, h = blockIdx.y, b = blockIdx.z;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); tma_prefetch_desc(&map_o);
    mbar_init(&q_full, 1); mbar_init(&s_full, 1); mbar_init(&p_full, 4); mbar_init(&o_final, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_base_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128u;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&q_full, FA_TILE_BYTES);
      tma_load_4d(q_s, &map_q, &q_full, 0, h, q0, b);
      for (int j = 0; j < a.ntiles; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((uint32_t)(j >> 1) & 1u) ^ 1u, 100 + st);
        unsigned char* ks = kv_s + st * 2 * FA_TILE_BYTES;
        mbar_expect_tx(&kv_full[st], 2 * FA_TILE_BYTES);
        tma_load_4d(ks, &map_k, &kv_full[st], 0, h, j * FA_BK, b);
        tma_load_4d(ks + FA_TILE_BYTES, &map_v, &kv_full[st], 0, h, j * FA_BK, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(&q_full, 0, 200);
      const uint32_t q_addr = s_u32(q_s), p_addr = s_u32(p_s);
      for (int j = 0; j < a.ntiles; ++j) {
        const int st = j & 1;
        const bool last = j == a.ntiles - 1;
        const uint32_t k_addr = s_u32(kv_s + st * 2 * FA_TILE_BYTES);
        const uint32_t v_addr = k_addr + FA_TILE_BYTES;
        mbar_wait(&kv_full[st], (uint32_t)(j >> 1) & 1u, 300 + st);
        tc_fence_after();
        FA_DBG(0, j);
        // S = Q K^T  (K extent = kpad; channels d..63 of both tiles are TMA zero fill)
        const uint64_t qd = make_smem_desc(q_addr), kd = make_smem_desc(k_addr);
        for (int k = 0; k < a.kpad / 16; ++k)
          umma_bf16(tmem_s, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), last ? a.idesc_s_last : a.idesc_s_full,
                    k != 0 ? 1u : 0u);
        umma_commit(&s_full);           // also implies: the previous tile's P V has retired
        // O += P V once the softmax warps have published P (and rescaled O)
        mbar_wait(&p_full, (uint32_t)j & 1u, 400);
        tc_fence_after();
        FA_DBG(1, j);
        const int ksteps = last ? (a.n_last_valid + 15) / 16 : FA_BK / 16;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t pd_ = make_smem_desc(p_addr + (uint32_t)(k >> 2) * FA_TILE_BYTES + (uint32_t)(k & 3) * 32u);
          const uint64_t vd = make_smem_desc_mn(v_addr + (uint32_t)k * 2048u, FA_TILE_BYTES, 1024);
          umma_bf16(tmem_o, pd_, vd, a.idesc_pv, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);     // K/V stage may be refilled once these MMAs retire
        if (last) umma_commit(&o_final);
      }
    }
  } else {
    // ---------------- softmax / correction / epilogue: thread == query row ----------------
    const int qd4 = warp & 3;
    const int r = qd4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd4 * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < a.ntiles; ++j) {
      const bool last = j == a.ntiles - 1;
      const int ncols = last ? a.n_last_pad : FA_BK;
      const int nvalid = last ? a.n_last_valid : FA_BK;
      mbar_wait(&s_full, (uint32_t)j & 1u, 500);
      tc_fence_after();
      if (warp == 2 && lane == 0) FA_DBG(2, j);
      // pass 1: row max
      float mx = -INFINITY;
      for (int c = 0; c < ncols; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_off + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (c + e < nvalid) mx = fmaxf(mx, __uint_as_float(v[e]));
      }
      if (warp == 2 && lane == 0) FA_DBG(3, j);
      const float m_new = fmaxf(m_run, mx * a.scale_log2);
      const float corr = exp2f(m_run - m_new);      // first tile: exp2(-inf) = 0
      // pass 2: p = exp2(s*scale - m), row sum, P -> smem (bf16, SWIZZLE_128B K-major)
      float lsum = 0.f;
      for (int c = 0; c < ncols; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_off + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int col = c + g * 8 + e;
            const float pv = exp2f(fmaf(__uint_as_float(v[g * 8 + e]), a.scale_log2, -m_new));
            p[e] = col < nvalid ? pv : 0.f;
            lsum += p[e];
          }
          const int kc = (c >> 3) + g;              // 16-byte chunk index along the key axis (0..15)
          unsigned char* dst = p_s + (kc >> 3) * FA_TILE_BYTES + r * 128 + (((kc & 7) ^ (r & 7)) << 4);
          *reinterpret_cast<bf16x8*>(dst) = pack8(p);
        }
      }
      if (warp == 2 && lane == 0) FA_DBG(4, j);
      l_run = l_run * corr + lsum;
      m_run = m_new;
      // rescale O (already complete for tiles < j: s_full of this tile was committed after their P V)
      if (j > 0 && __any_sync(0xffffffffu, corr != 1.0f)) {
        for (int c = 0; c < a.npad; c += 16) {
          uint32_t o[16];
          tmem_ld16(tmem_o + lane_off + (uint32_t)c, o);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * corr);
          tmem_st16(tmem_o + lane_off + (uint32_t)c, o);
        }
        tmem_st_wait();
      }
      fence_proxy_async();            // P (generic-proxy smem writes) -> visible to the tensor core's async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full);
      if (warp == 2 && lane == 0) FA_DBG(5, j);
    }
    // ---------------- epilogue ----------------
    mbar_wait(&o_final, 0, 600);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    for (int c = 0; c < a.npad; c += 16) {
      uint32_t o[16];
      tmem_ld16(tmem_o + lane_off + (uint32_t)c, o);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[g * 8 + e]) * inv;
        const int kc = (c >> 3) + g;
        *reinterpret_cast<bf16x8*>(q_s + r * 128 + ((kc ^ (r & 7)) << 4)) = pack8(f);   // Q tile is dead by now
      }
    }
    fence_proxy_async();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp == 2 && lane == 0) {
      tma_store_4d(&map_o, q_s, 0, h, q0, b);
      tma_store_commit();
      tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

bool attention_tc_supported(int dtype, int d, int ldq, int ldk, int ldv, int ldo, const void* q, const void* k,
                            const void* v, const void* out) {
  static int sm100 = -1;
  if (sm100 < 0) sm100 = pd_device_is_sm100();
  return sm100 && dtype == PD_BF16 && d % 8 == 0 && d >= 16 && d <= 64 && ldq % 8 == 0 && ldk % 8 == 0 &&
         ldv % 8 == 0 && ldo % 8 == 0 && ((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 &&
         ((uintptr_t)v % 16) == 0 && ((uintptr_t)out % 16) == 0;
}

int attention_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                 int heads, int Nq, int Nk, int d, float scale, cudaStream_t s) {
  FaArgs a;
  a.Nq = Nq; a.Nk = Nk; a.d = d;
  a.kpad = (d + 15) / 16 * 16;
  a.npad = (d + 15) / 16 * 16;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.ntiles = (Nk + FA_BK - 1) / FA_BK;
  a.n_last_valid = Nk - (a.ntiles - 1) * FA_BK;
  a.n_last_pad = (a.n_last_valid + 15) / 16 * 16;
  // kind::f16 instruction descriptor: fp32 accumulate, bf16 A/B, M = 128 (see gemm_sm100.cu)
  const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
  a.idesc_s_full = base | ((uint32_t)(FA_BK >> 3) << 17);
  a.idesc_s_last = base | ((uint32_t)(a.n_last_pad >> 3) << 17);
  a.idesc_pv = base | (1u << 16) | ((uint32_t)(a.npad >> 3) << 17);   // B (= V tile) is MN-major

  CUtensorMap mq, mk, mv, mo;
  const uint32_t box[4] = {64, 1, 128, 1};
  const uint32_t es[4] = {1, 1, 1, 1};
  struct { CUtensorMap* m; const void* p; int ld; int n; const char* nm; } t[4] = {
      {&mq, q, ldq, Nq, "attnQ"}, {&mk, k, ldk, Nk, "attnK"}, {&mv, v, ldv, Nk, "attnV"}, {&mo, out, ldo, Nq, "attnO"}};
  for (int i = 0; i < 4; ++i) {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)t[i].n, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)t[i].ld * 2, (uint64_t)t[i].n * t[i].ld * 2};
    int rc = encode_map(t[i].m, t[i].p, 4, dims, strides, box, es, t[i].nm);
    if (rc) return rc;
  }
  const size_t smem = FA_SMEM_BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("attention_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  dim3 grid((Nq + FA_BQ - 1) / FA_BQ, heads, B);
  attention_tc_kernel<<<grid, FA_THREADS, smem, s>>>(mq, mk, mv, mo, a);
  return check_launch("attention_tc");
}

}  // namespace pd

// debugging aid: device buffer of 6*32 uint64 receiving block (0,0,0)'s per-tile phase stamps; NULL = off
extern "C" int pd_debug_attention_timeline(void* dev_buf) {
  unsigned long long* p = (unsigned long long*)dev_buf;
  cudaError_t e = cudaMemcpyToSymbol(pd::g_fa_dbg, &p, sizeof(p));
  return e == cudaSuccess ? 0 : (int)e;
}
