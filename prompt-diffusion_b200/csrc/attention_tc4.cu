// tcgen05 / TMEM / TMA streaming-softmax attention, FOUR query groups per CTA, 64-key tiles — the d <= 64
// long-sequence case (d = 40 self-attention over 4096 / 9216 tokens, CrossAttention.forward,
// ldm/modules/attention.py:171-193), which is bound by the exponentials, not by the tensor pipe.
//
// Why a second kernel next to attention_tc.cu (two 128-query groups, 128-key tiles): there every softmax thread
// owns a query row and keeps 128 live scores, so only TWO softmax warps fit per SM sub-partition, and two
// in-order warps cannot keep the SFU busy (measured: 9.4 clk per exponential at 2 warps, 8.6 at 4, 7.8 = the pipe;
// the kernel ran at 61 % of the MUFU bound and every rearrangement of the same two warps landed on the same
// plateau — DESIGN.md 4.0).  Here the key tile is HALVED (64 live scores per thread -> ~100 registers) and the
// number of query groups DOUBLED, so four softmax warps share each sub-partition at the same work per key tile,
// with no row split, no shuffles and no shared-memory exchange:
//
//   one CTA per SM = 512 queries of one (batch, head) = four 128-query groups sharing every K/V tile, 640 threads:
//   warps 0-15 : softmax, warp w = group (w >> 2), TMEM lane quadrant (w & 3); one thread per query row
//   warp 16    : TMA producer (Q of all groups once, then K / V tiles of 64 keys into two rings)
//   warp 17    : MMA issuer.  Per key tile and group: O_g += P_g V (A = P_g from TMEM), then S_g = Q_g K^T of the
//                NEXT tile.  P_g aliases the first 32 columns of S_g; the in-order tensor pipe keeps Q K^T(j+1)
//                behind the P V(j) that reads it.  While one group's MMAs run, the other three exponentiate.
//   TMEM       : S_g at columns [64 g, 64 g + 64), O_g at [256 + 64 g, 256 + 64 g + KPAD)           (512 columns)
//   registers  : setmaxnreg — the control warp group drops to 56, the four softmax warp groups rise to 104
//
#include <cstdlib>

#include "tc_ptx.cuh"

namespace pd {

constexpr int F4_BQ = 128, F4_GROUPS = 4, F4_BK = 64, F4_THREADS = 640, F4_STAGES = 4;
constexpr int F4_Q_BYTES = 128 * 128;      // [128 rows][64 bf16] SWIZZLE_128B
constexpr int F4_KV_BYTES = 64 * 128;      // [64 keys][64 bf16]
constexpr float F4_RESCALE_THRESHOLD = 8.0f;   // log2 units
// Registers are a per-sub-partition resource (16384 each): five warps per sub-partition cap a thread at 96 at launch
// (an 18-warp CTA at 112 is refused: "too many resources").  setmaxnreg then moves registers inside the CTA's launch
// allocation: the control warp group gives back 128 x (96 - 56) = 5120, the four softmax warp groups take
// 512 x (104 - 96) = 4096.  ptxas -v: no spills in the softmax loop at 104.
constexpr int F4_REGS_CTRL = 56, F4_REGS_SOFTMAX = 104;
#ifndef F4_POLY
#define F4_POLY 0          // of every 8 element pairs, how many take the FMA-pipe exp2 (0..8)
#endif

extern unsigned long long* g_fa_dbg_host;     // attention_tc.cu: phase-timeline buffer set by pd_debug_attention_timeline

// phase stamps of block (0,0,0): slots 0-3 = MMA thread saw p_full[g], 4-7 = group g has S in registers, 8-11 = group g
// arrived on p_full (12 slots x 32 tiles of globaltimer ns)
#define F4_DBG(slot, tile)                                                                       \
  do {                                                                                           \
    if (a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tile) < 32) \
      a.dbg[(slot) * 32 + (tile)] = gtimer();                                                    \
  } while (0)

struct F4Args {
  unsigned long long* dbg;
  int Nq, Nk;
  float scale_log2;
  uint32_t idesc_s_full, idesc_s_last, idesc_pv;
  int n_last_valid;    // keys actually present in the last tile
  int ntiles;
};

template <int KP16>
__global__ void __launch_bounds__(F4_THREADS, 1)
attention_tc4_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                     const F4Args a) {
  static_assert(KP16 >= 1 && KP16 <= 4, "one 64-channel chunk per head");
  constexpr int KPAD = KP16 * 16;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* q_s = smem;                                        // [group]
  unsigned char* k_s = q_s + F4_GROUPS * F4_Q_BYTES;                // [stage]
  unsigned char* v_s = k_s + F4_STAGES * F4_KV_BYTES;               // [stage]
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + F4_STAGES * F4_KV_BYTES);
  uint64_t& q_full = bars[0];
  uint64_t* s_full = bars + 1;        // [4]  MMA -> softmax: S_g(j) complete (and O_g holds tiles < j)
  uint64_t* p_full = bars + 5;        // [4]  softmax -> MMA: P_g(j) in TMEM, O_g rescaled
  uint64_t* o_final = bars + 9;       // [4]
  uint64_t* k_full = bars + 13;       // [4]
  uint64_t* k_empty = bars + 17;      // [4]
  uint64_t* v_full = bars + 21;       // [4]
  uint64_t* v_empty = bars + 25;      // [4]
  uint32_t& tmem_base_slot = *reinterpret_cast<uint32_t*>(bars + 29);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (F4_GROUPS * F4_BQ), h = blockIdx.y, b = blockIdx.z;
  constexpr int W_TMA = 16, W_MMA = 17;

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); tma_prefetch_desc(&map_o);
    mbar_init(&q_full, 1);
    for (int g = 0; g < F4_GROUPS; ++g) { mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 4); mbar_init(&o_final[g], 1); }
    for (int i = 0; i < F4_STAGES; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_base_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();                        // the set-up above overlapped the previous kernel's tail (PDL)
  // every role reads the TMEM base into a register of ITS OWN branch: as one kernel-lifetime value ptxas parked it in
  // local memory (the softmax branch needs every register) and re-loaded it in front of each tcgen05.mma
  auto read_tmem_base = [&]() {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(s_u32(&tmem_base_slot)));
    return v;
  };

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F4_REGS_CTRL));
    if (warp == W_TMA) {
      if (lane == 0) {
        mbar_expect_tx(&q_full, F4_GROUPS * F4_Q_BYTES);
        for (int g = 0; g < F4_GROUPS; ++g) tma_load_4d(q_s + g * F4_Q_BYTES, &map_q, &q_full, 0, h, q0 + g * F4_BQ, b);
        int st = 0; uint32_t ph = 0;
        for (int j = 0; j < a.ntiles; ++j) {
          mbar_wait(&k_empty[st], ph ^ 1u, 100 + st);
          mbar_expect_tx(&k_full[st], F4_KV_BYTES);
          tma_load_4d(k_s + st * F4_KV_BYTES, &map_k, &k_full[st], 0, h, j * F4_BK, b);
          mbar_wait(&v_empty[st], ph ^ 1u, 110 + st);
          mbar_expect_tx(&v_full[st], F4_KV_BYTES);
          tma_load_4d(v_s + st * F4_KV_BYTES, &map_v, &v_full[st], 0, h, j * F4_BK, b);
          if (++st == F4_STAGES) { st = 0; ph ^= 1u; }
        }
      }
    } else if (warp == W_MMA) {
      // Warp-uniform loop, ONE ELECTED lane issues (elect.sync): inside an `if (lane == 0)` region ptxas wraps every
      // UTCHMMA / UTCBAR in an ELECT ... BRA.U.ANY retry loop — measured here at ~150 clk per MMA, 0.58 us per group and
      // key tile, which serialised the four groups behind this one thread (profiles/r02_attn4_timeline_a.txt).
      {
        const uint32_t tmem_base = read_tmem_base();
        const uint64_t qdesc0 = make_smem_desc(s_u32(q_s));
        const uint64_t kdesc0 = make_smem_desc(s_u32(k_s));
        const uint64_t vdesc0 = make_smem_desc_mn(s_u32(v_s), F4_KV_BYTES, 1024);
        constexpr uint64_t QT16 = F4_Q_BYTES >> 4, KVT16 = F4_KV_BYTES >> 4;
        // S_g = Q_g K^T : K extent KPAD (channels d..KPAD-1 of both operands are TMA zero fill)
        auto issue_qk = [&](int g, int st_k, uint32_t idesc) {
          const uint64_t qd = qdesc0 + (uint64_t)g * QT16, kd = kdesc0 + (uint64_t)st_k * KVT16;
#pragma unroll
          for (int k = 0; k < KP16; ++k)
            umma_bf16(tmem_base + (uint32_t)(g * 64), qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
          umma_commit(&s_full[g]);
        };
        // O_g += P_g V : A = P_g from TMEM (8 columns per 16-key step), B = V tile in place, MN-major
        auto issue_pv = [&](int g, int st_v, bool first, int ksteps) {
          const uint64_t vd = vdesc0 + (uint64_t)st_v * KVT16;
          const uint32_t pa = tmem_base + (uint32_t)(g * 64), oa = tmem_base + 256u + (uint32_t)(g * 64);
          for (int k = 0; k < ksteps; ++k)
            umma_bf16_ts(oa, pa + (uint32_t)(8 * k), vd + (uint64_t)(k * 128), a.idesc_pv, (k != 0 || !first) ? 1u : 0u);
        };
        mbar_wait(&q_full, 0, 200);
        mbar_wait(&k_full[0], 0, 300);
        tc_fence_after();
        const uint32_t id0 = a.ntiles == 1 ? a.idesc_s_last : a.idesc_s_full;
        if (elect_one()) {
          for (int g = 0; g < F4_GROUPS; ++g) issue_qk(g, 0, id0);
          umma_commit(&k_empty[0]);
        }
        __syncwarp();
        int st = 0; uint32_t ph = 0;          // ring position of tile j
        for (int j = 0; j < a.ntiles; ++j) {
          int stn = st + 1; uint32_t phn = ph;
          if (stn == F4_STAGES) { stn = 0; phn ^= 1u; }
          const bool more = j + 1 < a.ntiles;
          const uint32_t idn = (j + 2 == a.ntiles) ? a.idesc_s_last : a.idesc_s_full;
          const int ksteps = more ? F4_BK / 16 : (a.n_last_valid + 15) / 16;
          mbar_wait(&v_full[st], ph, 310 + st);
          if (more) mbar_wait(&k_full[stn], phn, 300 + stn);
          for (int g = 0; g < F4_GROUPS; ++g) {
            mbar_wait(&p_full[g], (uint32_t)j & 1u, 400 + g);
            tc_fence_after();
            if (lane == 0) F4_DBG(g, j);
            if (elect_one()) {
              issue_pv(g, st, j == 0, ksteps);
              if (g == F4_GROUPS - 1) umma_commit(&v_empty[st]);
              if (more) {
                issue_qk(g, stn, idn);
                if (g == F4_GROUPS - 1) umma_commit(&k_empty[stn]);
              } else {
                umma_commit(&o_final[g]);
              }
            }
            __syncwarp();
          }
          st = stn; ph = phn;
        }
      }
    }
  } else {
    // ---------------- softmax / correction / epilogue: thread == query row ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F4_REGS_SOFTMAX));
    const uint32_t tmem_base = read_tmem_base();
    const int g = warp >> 2;                       // query group
    const int qd4 = warp & 3;                      // TMEM lane quadrant this warp may touch
    const int r = qd4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd4 * 32) << 16;
    const uint32_t tmem_s = tmem_base + (uint32_t)(g * 64) + lane_off;      // P_g aliases S_g[:, 0:32)
    const uint32_t tmem_o = tmem_base + 256u + (uint32_t)(g * 64) + lane_off;
    const float sc = a.scale_log2;
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < a.ntiles; ++j) {
      const bool last = j == a.ntiles - 1;
      mbar_wait(&s_full[g], (uint32_t)j & 1u, 500 + g);
      tc_fence_after();
      uint32_t s[64];
      tmem_ld32p(tmem_s, s);
      tmem_ld32p(tmem_s + 32u, s + 32);
      tmem_ld_wait();
      if (qd4 == 0 && lane == 0) F4_DBG(4 + g, j);
      if (last && a.n_last_valid < F4_BK) {
        const int nv = a.n_last_valid;
#pragma unroll
        for (int e = 0; e < 64; ++e)
          if (e >= nv) s[e] = 0xff800000u;         // -inf: keys past Nk (stale / zero-filled columns)
      }
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]);
#pragma unroll
      for (int e = 2; e < 62; e += 4) {
        mx0 = fmax3(mx0, __uint_as_float(s[e]), __uint_as_float(s[e + 1]));
        mx1 = fmax3(mx1, __uint_as_float(s[e + 2]), __uint_as_float(s[e + 3]));
      }
      const float mx = fmax3(mx0, mx1, fmaxf(__uint_as_float(s[62]), __uint_as_float(s[63]))) * sc;
      // lazy reference update: keep the stale reference unless the row max grew by more than 2^THRESHOLD
      float corr = 1.0f;
      if (mx > m_ref + F4_RESCALE_THRESHOLD) {
        corr = ex2_approx(m_ref - mx);             // first tile: exp2(-inf) = 0
        m_ref = mx;
        l_run *= corr;
      }
      // s_full(j) was committed after P V(j-1) in issue order, so O_g holds every tile < j here
      if (j > 0 && __any_sync(0xffffffffu, corr != 1.0f)) {
#pragma unroll
        for (int c = 0; c < KPAD; c += 16) {
          uint32_t o[16];
          tmem_ld16(tmem_o + (uint32_t)c, o);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * corr);
          tmem_st16(tmem_o + (uint32_t)c, o);
        }
      }
      const float nm = -m_ref;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      // scale-and-subtract (packed FFMA2), exp2, row sum (packed FADD2), bf16 pack, P -> TMEM (over S's first columns)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const int i = c * 32 + e;
          float x0, x1, x2, x3;
          ffma2(x0, x1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), sc, sc, nm, nm);
          ffma2(x2, x3, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]), sc, sc, nm, nm);
          if (((e >> 1) & 7) >= 8 - F4_POLY) exp2_poly2(x0, x1); else { x0 = ex2_approx(x0); x1 = ex2_approx(x1); }
          if ((((e >> 1) + 1) & 7) >= 8 - F4_POLY) exp2_poly2(x2, x3); else { x2 = ex2_approx(x2); x3 = ex2_approx(x3); }
          fadd2(l0, l1, l0, l1, x0, x1);
          fadd2(l2, l3, l2, l3, x2, x3);
          pk[e >> 1] = pack_bf16x2(x0, x1);
          pk[(e >> 1) + 1] = pack_bf16x2(x2, x3);
        }
        tmem_st16p(tmem_s + (uint32_t)(c * 16), pk);   // P_g: 32 columns of bf16 pairs
      }
      tmem_st_wait();
      l_run += (l0 + l1) + (l2 + l3);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      if (qd4 == 0 && lane == 0) F4_DBG(8 + g, j);
    }
    // ---------------- epilogue ----------------
    if (warp == 0 && lane == 0) griddep_launch();
    mbar_wait(&o_final[g], 0, 600 + g);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    unsigned char* stage_o = q_s + g * F4_Q_BYTES;                 // Q_g is dead: every Q K^T has retired
#pragma unroll
    for (int c = 0; c < KPAD; c += 16) {
      uint32_t o[16];
      tmem_ld16(tmem_o + (uint32_t)c, o);
      tmem_ld_wait();
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[gg * 8 + e]) * inv;
        const int kc = (c >> 3) + gg;               // 16-byte chunk along the channel axis
        *reinterpret_cast<bf16x8*>(stage_o + r * 128 + (((kc & 7) ^ (r & 7)) << 4)) = pack8(f);
      }
    }
    fence_proxy_async();
    asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
    if (qd4 == 0 && lane == 0 && q0 + g * F4_BQ < a.Nq) {
      tma_store_4d(&map_o, stage_o, 0, h, q0 + g * F4_BQ, b);
      tma_store_commit();
      tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(read_tmem_base()), "r"(512));
  }
}

// d <= 64 with enough queries and keys to amortise a 512-query CTA: the long self-attention of the first level
// Measured on B200 (profiles/r02_attn_bench.txt, r02_attn4_timeline_a/b.txt, r02_attn_mma_modes.txt): B16 h8 N4096 d40
// took 1135 us as first written (single-lane MMA issue: 0.58 us per group and key tile on the issuing thread, the four
// groups ran strictly one after the other), 1069 us without the spilled TMEM base, and 761 us with the warp-uniform
// elected issue below — the same as the two-group kernel's 762 us; at 9216 tokens 7.57 vs 7.28 ms.  With 64-key tiles
// the kernel needs 6 + 8 tcgen05.mma per 128 keys and group instead of 3 + 8 and is bound by the tensor pipe's
// per-instruction cost.  Auto therefore does NOT select it (PD_B200_ATTN4=1 or engine 5 do).
static int g_tc4_on = -1;     // -1: read PD_B200_ATTN4 once (default off)
bool attention_tc4_supported(int d, int Nq, int Nk) {
  if (g_tc4_on < 0) {
    const char* e = getenv("PD_B200_ATTN4");
    g_tc4_on = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_tc4_on && d <= 64 && Nq >= 1024 && Nk >= 512;
}

int attention_tc4(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                  int heads, int Nq, int Nk, int d, float scale, cudaStream_t s) {
  F4Args a;
  a.dbg = g_fa_dbg_host;
  a.Nq = Nq; a.Nk = Nk;
  const int kpad = (d + 15) / 16 * 16;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.ntiles = (Nk + F4_BK - 1) / F4_BK;
  a.n_last_valid = Nk - (a.ntiles - 1) * F4_BK;
  const int n_last_pad = (a.n_last_valid + 15) / 16 * 16;
  // kind::f16 instruction descriptor: fp32 accumulate, bf16 A/B, M = 128 (see gemm_sm100.cu)
  const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
  a.idesc_s_full = base | ((uint32_t)(F4_BK >> 3) << 17);
  a.idesc_s_last = base | ((uint32_t)(n_last_pad >> 3) << 17);
  a.idesc_pv = base | (1u << 16) | ((uint32_t)(kpad >> 3) << 17);   // B (= V tile) is MN-major

  CUtensorMap mq, mk, mv, mo;
  const uint32_t es[4] = {1, 1, 1, 1};
  struct { CUtensorMap* m; const void* p; int ld; int n; uint32_t rows; const char* nm; } t[4] = {
      {&mq, q, ldq, Nq, 128, "attn4Q"}, {&mk, k, ldk, Nk, 64, "attn4K"}, {&mv, v, ldv, Nk, 64, "attn4V"}, {&mo, out, ldo, Nq, 128, "attn4O"}};
  for (int i = 0; i < 4; ++i) {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)t[i].n, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)t[i].ld * 2, (uint64_t)t[i].n * t[i].ld * 2};
    const uint32_t box[4] = {64, 1, t[i].rows, 1};
    int rc = encode_map(t[i].m, t[i].p, 4, dims, strides, box, es, t[i].nm);
    if (rc) return rc;
  }
  const size_t smem = (size_t)F4_GROUPS * F4_Q_BYTES + 2 * F4_STAGES * F4_KV_BYTES + 256 + 1024;
  dim3 grid((Nq + F4_GROUPS * F4_BQ - 1) / (F4_GROUPS * F4_BQ), heads, B);
#define F4_LAUNCH(KP)                                                                                              \
  case KP: {                                                                                                       \
    static bool attr_set[16] = {false};                                                                            \
    int dev = 0;                                                                                                   \
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) { set_error("attention_tc4: bad device"); return PD_ERR_NO_DEVICE; } \
    if (!attr_set[dev]) {                                                                                          \
      cudaError_t e = cudaFuncSetAttribute(attention_tc4_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           (int)smem);                                                             \
      if (e != cudaSuccess) { set_error("attention_tc4: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; } \
      attr_set[dev] = true;                                                                                        \
    }                                                                                                              \
    cudaError_t le = launch_pdl(attention_tc4_kernel<KP>, grid, dim3(F4_THREADS), smem, s, 1, mq, mk, mv, mo, a);  \
    if (le != cudaSuccess) { set_error("attention_tc4: launch failed: %s", cudaGetErrorString(le)); return (int)le; } \
  } break;
  switch (kpad / 16) {
    F4_LAUNCH(1) F4_LAUNCH(2) F4_LAUNCH(3) F4_LAUNCH(4)
    default: set_error("attention_tc4: unsupported head dim %d", d); return PD_ERR_UNSUPPORTED;
  }
#undef F4_LAUNCH
  return check_launch("attention_tc4");
}

}  // namespace pd

// A/B switch: 0 = auto never picks the four-group kernel (engine 5 still selects it explicitly)
extern "C" int pd_debug_attention_tc4(int32_t on) { pd::g_tc4_on = on != 0; return 0; }
