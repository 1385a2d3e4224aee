// PERSISTENT form of the tcgen05 streaming-softmax self-attention kernel (attention_tc.cu) for head dims <= 64 — the
// d = 40, 4096 .. 9216-token self-attention (CrossAttention.forward, ldm/modules/attention.py:171-193) that is a fifth of a
// denoising step.
//
// attention_tc.cu launches one CTA per 256 queries of a (batch, head): 2048 CTAs per call at config 2, each paying its own
// prologue (Q and first K tile through TMA, first Q K^T: ~2.5 us before the first exponential) and epilogue around ~50 us
// of useful work, one after the other on an SM.  Here ONE CTA per SM walks (batch, head, query pair) units dealt
// round-robin; the roles, the TMEM layout and the per-tile protocol are those of attention_tc.cu, and the unit boundary
// disappears from the critical path:
//   * Q lives in two slots: the producer loads Q of unit i + 1 (and keeps streaming K / V: the rings simply continue)
//     while unit i is still being processed;
//   * the MMA warp treats the tiles of all its units as ONE sequence: Q K^T of the FIRST tile of unit i + 1 is issued as
//     soon as the softmax threads have pulled the last S of unit i into registers — under unit i's last exponentials,
//     its last P V and its epilogue;
//   * O_g is handed back explicitly (o_free) once the epilogue has read it out of TMEM, the Q slot once the epilogue's
//     TMA store (staged in the dead Q tile, as before) has read it.
//
// TMEM (512 columns): S_a [0,128) | S_b [128,256) | P_a [256,320) | P_b [320,384) | O_a [384,448) | O_b [448,512)
#include "tc_ptx.cuh"

namespace pd {

constexpr int PT_BQ = 128, PT_BK = 128, PT_THREADS = 320, PT_STAGES = 3;
constexpr int PT_TILE_BYTES = 128 * 128;   // one [128 rows][64 bf16] SWIZZLE_128B tile
constexpr float PT_GROW_LIMIT = 1.8446744e19f;   // 2^64, see attention_tc.cu

struct PtArgs {
  int Nq, Nk, heads;
  int npairs;          // 256-query pairs per (batch, head)
  int units;           // B * heads * npairs
  float scale_log2;
  uint32_t idesc_s_full, idesc_s_last, idesc_pv;
  int n_last_valid;    // keys actually present in the last tile
  int ntiles;
};

template <int KP16>
__global__ void __launch_bounds__(PT_THREADS, 1)
attention_ptc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                     const PtArgs a) {
  static_assert(KP16 >= 1 && KP16 <= 4, "persistent kernel: head dim <= 64 (P has its own TMEM columns)");
  constexpr int KPAD = KP16 * 16;
  constexpr uint32_t COL_S = 0, COL_P = 256, COL_O = 384;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* q_s = smem;                                          // [slot][group]
  unsigned char* k_s = q_s + 2 * 2 * PT_TILE_BYTES;                   // [stage]
  unsigned char* v_s = k_s + PT_STAGES * PT_TILE_BYTES;               // [stage]
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + PT_STAGES * PT_TILE_BYTES);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2]  both groups' stores have read the slot
  uint64_t* s_full = bars + 4;        // [2]  MMA -> softmax: S_g(tile) complete
  uint64_t* p_full = bars + 6;        // [2]  softmax -> MMA: P_g(tile) in TMEM, O_g rescaled
  uint64_t* o_final = bars + 8;       // [2]  MMA -> epilogue: O_g of the unit complete
  uint64_t* o_free = bars + 10;       // [2]  epilogue -> MMA: O_g is in registers
  uint64_t* k_full = bars + 12;       // [3]
  uint64_t* k_empty = bars + 15;      // [3]
  uint64_t* v_full = bars + 18;       // [3]
  uint64_t* v_empty = bars + 21;      // [3]
  uint64_t* s_free = bars + 24;       // [2]  softmax -> MMA: S_g(tile) is in registers
  uint64_t* p_free = bars + 26;       // [2]  MMA -> softmax: P_g(tile) V(tile) retired (every tile but a unit's last)
  uint32_t& tmem_base_slot = *reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_TMA = 8, W_MMA = 9;
  // units are dealt round-robin: at any time the CTAs work on ~gridDim.x CONSECUTIVE units, i.e. on a handful of (batch,
  // head) pairs whose K / V tiles they share through L2 exactly as the one-CTA-per-unit kernel does (contiguous ranges
  // made every CTA stream its own head: 8.3 against 7.3 ms at 9216 tokens)
  const int u0 = (int)blockIdx.x, u1 = a.units, ustep = (int)gridDim.x;
  const int ntiles = a.ntiles;

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); tma_prefetch_desc(&map_o);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&q_full[g], 1); mbar_init(&q_empty[g], 2);
      mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 4); mbar_init(&o_final[g], 1); mbar_init(&o_free[g], 4);
      mbar_init(&s_free[g], 4); mbar_init(&p_free[g], 1);
    }
    for (int i = 0; i < PT_STAGES; ++i) {
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
  const uint32_t tmem_base = tmem_base_slot;
  griddep_wait();                        // the set-up above overlapped the previous kernel's tail (PDL)

  if (warp == W_TMA) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;       // ring position: the K / V rings run through all units
      for (int u = u0, i = 0; u < u1; u += ustep, ++i) {
        const int bh = u / a.npairs, pair = u - bh * a.npairs;
        const int b = bh / a.heads, h = bh - b * a.heads;
        const int slot = i & 1;
        mbar_wait(&q_empty[slot], ((uint32_t)(i >> 1) & 1u) ^ 1u, 90 + slot);
        mbar_expect_tx(&q_full[slot], 2 * PT_TILE_BYTES);
        for (int g = 0; g < 2; ++g)
          tma_load_4d(q_s + (slot * 2 + g) * PT_TILE_BYTES, &map_q, &q_full[slot], 0, h, (pair * 2 + g) * PT_BQ, b);
        for (int j = 0; j < ntiles; ++j) {
          mbar_wait(&k_empty[st], ph ^ 1u, 100 + st);
          mbar_expect_tx(&k_full[st], PT_TILE_BYTES);
          tma_load_4d(k_s + st * PT_TILE_BYTES, &map_k, &k_full[st], 0, h, j * PT_BK, b);
          mbar_wait(&v_empty[st], ph ^ 1u, 110 + st);
          mbar_expect_tx(&v_full[st], PT_TILE_BYTES);
          tma_load_4d(v_s + st * PT_TILE_BYTES, &map_v, &v_full[st], 0, h, j * PT_BK, b);
          if (++st == PT_STAGES) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // single-lane issue on purpose, as in attention_tc.cu (the slow issue keeps the two query groups out of phase)
    if (lane == 0 && u0 < u1) {
      const uint64_t qdesc0 = make_smem_desc(s_u32(q_s));
      const uint64_t kdesc0 = make_smem_desc(s_u32(k_s));
      const uint64_t vdesc0 = make_smem_desc_mn(s_u32(v_s), PT_TILE_BYTES, 1024);
      constexpr uint64_t TILE16 = PT_TILE_BYTES >> 4;
      const uint32_t tm_s = tmem_base + COL_S, tm_p = tmem_base + COL_P, tm_o = tmem_base + COL_O;
      auto issue_qk = [&](int g, int slot, int st_k, uint32_t idesc) {
        const uint64_t qd = qdesc0 + (uint64_t)(slot * 2 + g) * TILE16, kd = kdesc0 + (uint64_t)st_k * TILE16;
#pragma unroll
        for (int k = 0; k < KP16; ++k)
          umma_bf16(tm_s + (uint32_t)(g * 128), qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc, k != 0 ? 1u : 0u);
        umma_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, int st_v, bool first, bool last) {
        const uint64_t vd = vdesc0 + (uint64_t)st_v * TILE16;
        const uint32_t pa = tm_p + (uint32_t)g * 64u, oa = tm_o + (uint32_t)g * 64u;
        const int ksteps = last ? (a.n_last_valid + 15) / 16 : PT_BK / 16;
        for (int k = 0; k < ksteps; ++k)
          umma_bf16_ts(oa, pa + (uint32_t)(8 * k), vd + (uint64_t)(k * 128), a.idesc_pv, (k != 0 || !first) ? 1u : 0u);
      };
      const uint32_t id_first = ntiles == 1 ? a.idesc_s_last : a.idesc_s_full;
      // the first tile of the first unit
      mbar_wait(&q_full[0], 0, 200);
      mbar_wait(&k_full[0], 0, 300);
      tc_fence_after();
      issue_qk(0, 0, 0, id_first);
      issue_qk(1, 0, 0, id_first);
      umma_commit(&k_empty[0]);
      int st = 0; uint32_t ph = 0;          // ring position of the current tile
      uint32_t T = 0;                        // tiles completed so far (all units): parity source of the per-tile barriers
      for (int u = u0, i = 0; u < u1; u += ustep, ++i) {
        for (int j = 0; j < ntiles; ++j, ++T) {
          int stn = st + 1; uint32_t phn = ph;
          if (stn == PT_STAGES) { stn = 0; phn ^= 1u; }
          const bool last = j + 1 == ntiles;
          const bool succ = !last || u + ustep < u1;              // a next tile exists (possibly the first one of the next unit)
          if (succ) {
            // Q K^T of the NEXT tile of the sequence: needs its K tile, (first tile of a unit) its Q slot, and S_g of the
            // current tile in the softmax threads' registers
            const int nslot = last ? ((i + 1) & 1) : (i & 1);
            const int jn = last ? 0 : j + 1;
            const uint32_t idn = (jn + 1 == ntiles) ? a.idesc_s_last : a.idesc_s_full;
            if (last) mbar_wait(&q_full[nslot], (uint32_t)((i + 1) >> 1) & 1u, 200 + nslot);
            mbar_wait(&k_full[stn], phn, 300 + stn);
            for (int g = 0; g < 2; ++g) {
              mbar_wait(&s_free[g], T & 1u, 450 + g);
              tc_fence_after();
              issue_qk(g, nslot, stn, idn);
            }
            umma_commit(&k_empty[stn]);
          }
          mbar_wait(&v_full[st], ph, 310 + st);
          for (int g = 0; g < 2; ++g) {
            mbar_wait(&p_full[g], T & 1u, 400 + g);
            if (j == 0 && i > 0) mbar_wait(&o_free[g], (uint32_t)(i - 1) & 1u, 420 + g);   // O_g of the previous unit has left TMEM
            tc_fence_after();
            issue_pv(g, st, j == 0, last);
            umma_commit(last ? &o_final[g] : &p_free[g]);
          }
          umma_commit(&v_empty[st]);
          st = stn; ph = phn;
        }
      }
    }
  } else {
    // ---------------- softmax / correction / epilogue: thread == query row ----------------
    const int g = warp >> 2;                       // 0: warps 0-3, 1: warps 4-7
    const int qd4 = warp & 3;                      // TMEM lane quadrant this warp may touch
    const int r = qd4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd4 * 32) << 16;
    const uint32_t tmem_s = tmem_base + COL_S + (uint32_t)(g * 128) + lane_off;
    const uint32_t tmem_p = tmem_base + COL_P + (uint32_t)g * 64u + lane_off;
    const uint32_t tmem_o = tmem_base + COL_O + (uint32_t)g * 64u + lane_off;
    const float sc = a.scale_log2;
    uint32_t T = 0;                                // tiles completed so far (all units)
    uint32_t npf = 0;                              // p_free completions consumed so far
    for (int u = u0, i = 0; u < u1; u += ustep, ++i) {
      const int bh = u / a.npairs, pair = u - bh * a.npairs;
      const int b = bh / a.heads, h = bh - b * a.heads;
      const int slot = i & 1;
      const int q0 = (pair * 2 + g) * PT_BQ;
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < ntiles; ++j, ++T) {
        const bool last = j == ntiles - 1;
        mbar_wait(&s_full[g], T & 1u, 500 + g);
        tc_fence_after();
        uint32_t s[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32p(tmem_s + (uint32_t)(c * 32), s + c * 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[g]);    // S_g may be overwritten by Q K^T of the next tile
        if (last && a.n_last_valid < PT_BK) {
          const int nv = a.n_last_valid;
#pragma unroll
          for (int e = 0; e < 128; ++e)
            if (e >= nv) s[e] = 0xff800000u;       // -inf: keys past Nk (stale / zero-filled columns)
        }
        // reference for the exponentials: fixed at the unit's first tile, see attention_tc.cu
        auto row_max = [&]() {
          float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]);
#pragma unroll
          for (int e = 2; e < 126; e += 4) {
            mx0 = fmax3(mx0, __uint_as_float(s[e]), __uint_as_float(s[e + 1]));
            mx1 = fmax3(mx1, __uint_as_float(s[e + 2]), __uint_as_float(s[e + 3]));
          }
          return fmax3(mx0, mx1, fmaxf(__uint_as_float(s[126]), __uint_as_float(s[127]))) * sc;
        };
        if (j == 0) m_ref = fmaxf(row_max(), -1e30f);          // (a fully masked row keeps a finite reference)
        if (j > 0) {                                 // P_g(j-1) V(j-1) retired: P_g is free, O_g is complete
          mbar_wait(&p_free[g], npf & 1u, 550 + g);
          ++npf;
          tc_fence_after();
        }
        // (j == 0 of a later unit: the epilogue below waited for o_final, committed after the previous unit's last P V)
        float lt = 0.f;
        auto exp_tile = [&]() {
          const float nm = -m_ref;
          float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const int ii = c * 32 + e;
              float x0, x1, x2, x3;
              ffma2(x0, x1, __uint_as_float(s[ii]), __uint_as_float(s[ii + 1]), sc, sc, nm, nm);
              ffma2(x2, x3, __uint_as_float(s[ii + 2]), __uint_as_float(s[ii + 3]), sc, sc, nm, nm);
              x0 = ex2_approx(x0); x1 = ex2_approx(x1); x2 = ex2_approx(x2); x3 = ex2_approx(x3);
              fadd2(l0, l1, l0, l1, x0, x1);
              fadd2(l2, l3, l2, l3, x2, x3);
              pk[e >> 1] = pack_bf16x2(x0, x1);
              pk[(e >> 1) + 1] = pack_bf16x2(x2, x3);
            }
            tmem_st16p(tmem_p + (uint32_t)(c * 16), pk);   // P_g: 64 columns of bf16 pairs
          }
          lt = (l0 + l1) + (l2 + l3);
        };
        exp_tile();
        // overflow guard (warp-uniform: the rescale uses warp-collective tcgen05.ld / st)
        if (j > 0 && __any_sync(0xffffffffu, !(lt < PT_GROW_LIMIT))) {
          const float mx = row_max();
          float corr = 1.0f;
          if (mx > m_ref) { corr = ex2_approx(m_ref - mx); m_ref = mx; l_run *= corr; }
          tmem_st_wait();
#pragma unroll
          for (int c = 0; c < KPAD; c += 16) {
            uint32_t o[16];
            tmem_ld16(tmem_o + (uint32_t)c, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * corr);
            tmem_st16(tmem_o + (uint32_t)c, o);
          }
          exp_tile();                                  // P_g(j) rewritten against the new reference
        }
        l_run += lt;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
      }
      // ---------------- epilogue of the unit ----------------
      mbar_wait(&o_final[g], (uint32_t)i & 1u, 600 + g);
      tc_fence_after();
      const float inv = 1.0f / l_run;
      unsigned char* stage_o = q_s + (slot * 2 + g) * PT_TILE_BYTES;       // Q_g of this unit is dead: every Q K^T has retired
      uint32_t o[KPAD];
#pragma unroll
      for (int c = 0; c < KPAD; c += 16) tmem_ld16(tmem_o + (uint32_t)c, *reinterpret_cast<uint32_t(*)[16]>(&o[c]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[g]);      // the next unit's first P V may overwrite O_g
      const uint32_t stage_row = s_u32(stage_o) + (uint32_t)(r * 128 + ((r & 7) << 4));
#pragma unroll
      for (int kc = 0; kc < KPAD / 8; ++kc) {       // 16-byte chunk along the channel axis
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[kc * 8 + e]) * inv;
        sts_bf16x8(stage_row ^ (uint32_t)(kc << 4), pack8(f));
      }
      fence_proxy_async();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      if (qd4 == 0 && lane == 0) {
        if (q0 < a.Nq) tma_store_4d(&map_o, stage_o, 0, h, q0, b);
        tma_store_commit();
        tma_store_wait_read<0>();                  // ~1 us once per unit (50 us): the slot goes back to the producer
        mbar_arrive(&q_empty[slot]);
      }
    }
    if (warp == 0 && lane == 0) griddep_launch();
    if (qd4 == 0 && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// Measured on B200 (profiles/r02_attn_persistent.txt): B16 h8 N4096 d40 753 us against 760 us for the one-CTA-per-unit kernel,
// N9216 7.44 against 7.27 ms, N1024 66.7 against 71.6 us — the unit boundary was NOT what the streaming kernel loses its
// time to (the hardware starts the next CTA fast and its K / V tiles hit L2): the steady-state softmax pipeline is.  Kept
// as an explicit engine (8, bit-equal to engine 3); auto selects it only with PD_B200_ATTN_PERSIST=1.
static int g_ptc_on = -1;     // -1: read PD_B200_ATTN_PERSIST once (default off)
bool attention_ptc_supported(int d, int Nq, int Nk, int B, int heads) {
  if (g_ptc_on < 0) {
    const char* e = getenv("PD_B200_ATTN_PERSIST");
    g_ptc_on = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  // long key sequences with enough 256-query units to give every SM several
  return g_ptc_on && d <= 64 && Nk >= 512 && (long long)B * heads * ((Nq + 255) / 256) >= 592;
}

int attention_ptc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                  int heads, int Nq, int Nk, int d, float scale, cudaStream_t s) {
  PtArgs a;
  a.Nq = Nq; a.Nk = Nk; a.heads = heads;
  a.npairs = (Nq + 2 * PT_BQ - 1) / (2 * PT_BQ);
  const long long units = (long long)B * heads * a.npairs;
  if (units > 0x7fffffffLL || d > 64) { set_error("attention_ptc: unsupported geometry"); return PD_ERR_UNSUPPORTED; }
  a.units = (int)units;
  const int kpad = (d + 15) / 16 * 16;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.ntiles = (Nk + PT_BK - 1) / PT_BK;
  a.n_last_valid = Nk - (a.ntiles - 1) * PT_BK;
  const int n_last_pad = (a.n_last_valid + 15) / 16 * 16;
  const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
  a.idesc_s_full = base | ((uint32_t)(PT_BK >> 3) << 17);
  a.idesc_s_last = base | ((uint32_t)(n_last_pad >> 3) << 17);
  a.idesc_pv = base | (1u << 16) | ((uint32_t)(kpad >> 3) << 17);   // B (= V tile) is MN-major

  CUtensorMap mq, mk, mv, mo;
  const uint32_t box[4] = {64, 1, 128, 1};
  const uint32_t es[4] = {1, 1, 1, 1};
  struct { CUtensorMap* m; const void* p; int ld; int n; const char* nm; } t[4] = {
      {&mq, q, ldq, Nq, "pattnQ"}, {&mk, k, ldk, Nk, "pattnK"}, {&mv, v, ldv, Nk, "pattnV"}, {&mo, out, ldo, Nq, "pattnO"}};
  for (int i = 0; i < 4; ++i) {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)t[i].n, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)t[i].ld * 2, (uint64_t)t[i].n * t[i].ld * 2};
    int rc = encode_map(t[i].m, t[i].p, 4, dims, strides, box, es, t[i].nm);
    if (rc) return rc;
  }
  const size_t smem = (size_t)(4 + 2 * PT_STAGES) * PT_TILE_BYTES + 256 + 1024;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError(); sms = 148;
  }
  dim3 grid((unsigned)(a.units < sms ? a.units : sms));
#define PT_LAUNCH(KP)                                                                                              \
  case KP: {                                                                                                       \
    static bool attr_set = false;                                                                                  \
    if (!attr_set) {                                                                                               \
      cudaError_t e = cudaFuncSetAttribute(attention_ptc_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           (int)smem);                                                             \
      if (e != cudaSuccess) { set_error("attention_ptc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; } \
      attr_set = true;                                                                                             \
    }                                                                                                              \
    cudaError_t le = launch_pdl(attention_ptc_kernel<KP>, grid, dim3(PT_THREADS), smem, s, 1, mq, mk, mv, mo, a);  \
    if (le != cudaSuccess) { set_error("attention_ptc: launch failed: %s", cudaGetErrorString(le)); return (int)le; } \
  } break;
  switch (kpad / 16) {
    PT_LAUNCH(1) PT_LAUNCH(2) PT_LAUNCH(3) PT_LAUNCH(4)
    default: set_error("attention_ptc: unsupported head dim %d", d); return PD_ERR_UNSUPPORTED;
  }
#undef PT_LAUNCH
  return check_launch("attention_ptc");
}

}  // namespace pd

// A/B switch: 0 = pd_attention's auto selection never picks the persistent self-attention kernel (engine 8), 1 = default
extern "C" int pd_debug_attention_persistent(int32_t on) { pd::g_ptc_on = on != 0; return 0; }
