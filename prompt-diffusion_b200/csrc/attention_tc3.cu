// tcgen05 / TMEM / TMA streaming-softmax attention, THREE query groups per CTA, 128-key tiles, head dim <= 40 —
// the d = 40 self-attention over 4096 / 9216 tokens (CrossAttention.forward, ldm/modules/attention.py:171-193).
//
// What the round-2 timelines showed (profiles/r02_attn4_timeline_*.txt): at d = 40 the tensor pipe is a CO-limiter of
// the exponentials — every tcgen05.mma of this size costs ~90 clk whatever its N, a 128 x 128 score block needs
// 3 (Q K^T) + 8 (P V) of them = ~1000 clk, about what the SFU needs for its 16384 exponentials — so the two must
// overlap, which needs more than two independent softmax streams per SM, and 64-key tiles are out (they double the
// Q K^T count: attention_tc3.cu is tensor-bound).  Three groups with 128-key tiles fit TMEM only as
//   S_g (fp32, P_g aliased onto its first 64 columns) at [128 g, 128 g + 128),  O_g (40 columns) at [384 + 40 g, ...)
// = 504 of 512 columns, and the register file only as 12 softmax warps at 152 registers + a control warp group at 56
// (setmaxnreg inside a 512-thread x 128-register launch; 3 x 152 x 32 + 56 x 32 = 16384 per sub-partition, exactly).
//
//   warps 0-11 : softmax, warp w = group (w >> 2), TMEM lane quadrant (w & 3); one thread per query row
//   warp 12    : TMA producer (Q of all groups once, then K / V tiles of 128 keys into two rings)
//   warp 13    : MMA issuer, warp-uniform loop, one ELECTED lane issues.  Per key tile and group: O_g += P_g V
//                (A = P_g from TMEM, N = 40), then S_g = Q_g K^T of the NEXT tile; the in-order tensor pipe keeps
//                Q K^T(j+1) behind the P V(j) that reads the aliased columns.
//
#include <cstdlib>

#include "tc_ptx.cuh"

#ifndef F3_POLL_SLEEP
#define F3_POLL_SLEEP 0
#endif

namespace pd {

constexpr int F3_BQ = 128, F3_GROUPS = 3, F3_BK = 128, F3_THREADS = 512, F3_STAGES = 3;
constexpr int F3_Q_BYTES = 128 * 128;      // [128 rows][64 bf16] SWIZZLE_128B
constexpr int F3_KV_BYTES = 128 * 128;     // [128 keys][64 bf16]
// Registers are a per-sub-partition resource (16384 each).  Launch: 512 threads x 128.  setmaxnreg: the control warp
// group gives back 128 x (128 - 56) = 9216, the three softmax warp groups take 384 x (152 - 128) = 9216.
constexpr int F3_REGS_CTRL = 56, F3_REGS_SOFTMAX = 152;
#ifndef F3_STAGGER
#define F3_STAGGER 1200        // clk between the first exponentials of consecutive groups
#endif
constexpr int F3_OCOLS = 40;                 // O_g columns = N extent of P V (head dim rounded up to 8, <= 40)
constexpr float F3_GROW_LIMIT = 1.8446744e19f;   // 2^64, see attention_tc.cu
#ifndef F3_SWP
#define F3_SWP 0           // > 0: exponential section software-pipelined by hand in place (see attention_tc.cu, FA_SWP)
#endif
#ifndef F3_POLY
#define F3_POLY 0          // of every 8 element pairs, how many take the FMA-pipe exp2 (0..8)
#endif

extern unsigned long long* g_fa_dbg_host;     // attention_tc.cu: phase-timeline buffer set by pd_debug_attention_timeline

// phase stamps of block (0,0,0): slots 0-3 = MMA thread saw p_full[g], 4-7 = group g has S in registers, 8-11 = group g
// arrived on p_full (12 slots x 32 tiles of globaltimer ns)
#define F3_DBG(slot, tile)                                                                       \
  do {                                                                                           \
    if (a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tile) < 32) \
      a.dbg[(slot) * 32 + (tile)] = gtimer();                                                    \
  } while (0)

struct F3Args {
  unsigned long long* dbg;
  int Nq, Nk;
  float scale_log2;
  uint32_t idesc_s_full, idesc_s_last, idesc_pv;
  int n_last_valid;    // keys actually present in the last tile
  int ntiles;
};

template <int KP16>
__global__ void __launch_bounds__(F3_THREADS, 1)
attention_tc3_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                     const F3Args a) {
  static_assert(KP16 >= 1 && KP16 <= 3, "head dim <= 40 (Q K^T over at most 48 zero-padded channels)");
  constexpr int KPAD = KP16 * 16;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* q_s = smem;                                        // [group]
  unsigned char* k_s = q_s + F3_GROUPS * F3_Q_BYTES;                // [stage]
  unsigned char* v_s = k_s + F3_STAGES * F3_KV_BYTES;               // [stage]
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + F3_STAGES * F3_KV_BYTES);
  uint64_t& q_full = bars[0];
  uint64_t* s_full = bars + 1;        // [3]  MMA -> softmax: S_g(j) complete (and O_g holds tiles < j)
  uint64_t* p_full = bars + 5;        // [3]  softmax -> MMA: P_g(j) in TMEM, O_g rescaled
  uint64_t* o_final = bars + 9;       // [3]
  uint64_t* k_full = bars + 13;       // [3]
  uint64_t* k_empty = bars + 17;      // [3]
  uint64_t* v_full = bars + 21;       // [3]
  uint64_t* v_empty = bars + 25;      // [3]
  uint32_t& tmem_base_slot = *reinterpret_cast<uint32_t*>(bars + 29);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (F3_GROUPS * F3_BQ), h = blockIdx.y, b = blockIdx.z;
  constexpr int W_TMA = 12, W_MMA = 13;
  // query groups that hold at least one real row: a trailing CTA of a (batch, head) runs fewer groups instead of
  // exponentiating zero-filled rows
  const int n_act = min(F3_GROUPS, (a.Nq - q0 + F3_BQ - 1) / F3_BQ);

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); tma_prefetch_desc(&map_o);
    mbar_init(&q_full, 1);
    for (int g = 0; g < F3_GROUPS; ++g) { mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 4); mbar_init(&o_final[g], 1); }
    for (int i = 0; i < F3_STAGES; ++i) {
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

  if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F3_REGS_CTRL));
    if (warp == W_TMA) {
      if (lane == 0) {
        mbar_expect_tx(&q_full, F3_GROUPS * F3_Q_BYTES);
        for (int g = 0; g < F3_GROUPS; ++g) tma_load_4d(q_s + g * F3_Q_BYTES, &map_q, &q_full, 0, h, q0 + g * F3_BQ, b);
        int st = 0; uint32_t ph = 0;
        for (int j = 0; j < a.ntiles; ++j) {
          mbar_wait(&k_empty[st], ph ^ 1u, 100 + st);
          mbar_expect_tx(&k_full[st], F3_KV_BYTES);
          tma_load_4d(k_s + st * F3_KV_BYTES, &map_k, &k_full[st], 0, h, j * F3_BK, b);
          mbar_wait(&v_empty[st], ph ^ 1u, 110 + st);
          mbar_expect_tx(&v_full[st], F3_KV_BYTES);
          tma_load_4d(v_s + st * F3_KV_BYTES, &map_v, &v_full[st], 0, h, j * F3_BK, b);
          if (++st == F3_STAGES) { st = 0; ph ^= 1u; }
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
        const uint64_t vdesc0 = make_smem_desc_mn(s_u32(v_s), F3_KV_BYTES, 1024);
        constexpr uint64_t QT16 = F3_Q_BYTES >> 4, KVT16 = F3_KV_BYTES >> 4;
        // S_g = Q_g K^T : K extent KPAD (channels d..KPAD-1 of both operands are TMA zero fill)
        auto issue_qk = [&](int g, int st_k, uint32_t idesc) {
          const uint64_t qd = qdesc0 + (uint64_t)g * QT16, kd = kdesc0 + (uint64_t)st_k * KVT16;
#pragma unroll
          for (int k = 0; k < KP16; ++k)
            umma_bf16(tmem_base + (uint32_t)(g * 128), qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
          umma_commit(&s_full[g]);
        };
        // O_g += P_g V : A = P_g from TMEM (8 columns per 16-key step), B = V tile in place, MN-major
        auto issue_pv = [&](int g, int st_v, bool first, int ksteps) {
          const uint64_t vd = vdesc0 + (uint64_t)st_v * KVT16;
          const uint32_t pa = tmem_base + (uint32_t)(g * 128), oa = tmem_base + 384u + (uint32_t)(g * F3_OCOLS);
          for (int k = 0; k < ksteps; ++k)
            umma_bf16_ts(oa, pa + (uint32_t)(8 * k), vd + (uint64_t)(k * 128), a.idesc_pv, (k != 0 || !first) ? 1u : 0u);
        };
        mbar_wait(&q_full, 0, 200);
        mbar_wait(&k_full[0], 0, 300);
        tc_fence_after();
        const uint32_t id0 = a.ntiles == 1 ? a.idesc_s_last : a.idesc_s_full;
        if (elect_one()) {
          for (int g = 0; g < n_act; ++g) issue_qk(g, 0, id0);
          umma_commit(&k_empty[0]);
        }
        __syncwarp();
        // Groups are served IN ARRIVAL ORDER, not round-robin: each group's cycle (exponentials, then its P V + next
        // Q K^T round trip through this warp) then runs on its own clock, so the phase offsets the softmax warps start
        // with (F3_STAGGER) persist and one group's round trip hides under the other groups' exponentials.  Served in
        // fixed order the three groups lock step: all exponentiate together, then all wait together
        // (profiles/r02_attn3_timeline_a.txt: 1.85 us of exponentials + a 0.65 us bubble per key tile).
        int jg0 = 0, jg1 = 0, jg2 = 0;                // next key tile of each group
        int pv_cnt[F3_STAGES] = {0, 0, 0}, qk_cnt[F3_STAGES] = {0, 0, 0};   // groups that have issued on the stage's tile
        int remaining = n_act * a.ntiles;
        long long t_poll = clock64();
        while (remaining > 0) {
          bool served = false;
#pragma unroll
          for (int g = 0; g < F3_GROUPS; ++g) {
            const int jj = g == 0 ? jg0 : g == 1 ? jg1 : jg2;
            if (g >= n_act || jj >= a.ntiles) continue;
            const int st = jj % F3_STAGES, stn = (jj + 1) % F3_STAGES;
            const uint32_t ph = (uint32_t)(jj / F3_STAGES) & 1u, phn = (uint32_t)((jj + 1) / F3_STAGES) & 1u;
            const bool more = jj + 1 < a.ntiles;
            int ready = mbar_try_wait(&p_full[g], (uint32_t)jj & 1u) && mbar_try_wait(&v_full[st], ph) &&
                        (!more || mbar_try_wait(&k_full[stn], phn));
            ready = __shfl_sync(0xffffffffu, ready, 0);          // one verdict for the warp
            if (!ready) continue;
            tc_fence_after();
            if (lane == 0) F3_DBG(g, jj);
            const uint32_t idn = (jj + 2 == a.ntiles) ? a.idesc_s_last : a.idesc_s_full;
            const int ksteps = more ? F3_BK / 16 : (a.n_last_valid + 15) / 16;
            const bool v_done = ++pv_cnt[st] == n_act;
            if (v_done) pv_cnt[st] = 0;
            bool k_done = false;
            if (more) { k_done = ++qk_cnt[stn] == n_act; if (k_done) qk_cnt[stn] = 0; }
            if (elect_one()) {
              issue_pv(g, st, jj == 0, ksteps);
              if (v_done) umma_commit(&v_empty[st]);
              if (more) {
                issue_qk(g, stn, idn);
                if (k_done) umma_commit(&k_empty[stn]);
              } else {
                umma_commit(&o_final[g]);
              }
            }
            __syncwarp();
            if (g == 0) ++jg0; else if (g == 1) ++jg1; else ++jg2;
            --remaining;
            served = true;
            t_poll = clock64();
          }
#if F3_POLL_SLEEP > 0
          // nothing was ready: back off instead of spinning 32 lanes on three barriers (the spin costs issue slots of
          // the sub-partition this warp shares with three softmax warps, and power the step does not have to spare)
          if (!served) __nanosleep(F3_POLL_SLEEP);
#else
          (void)served;
#endif
          if (clock64() - t_poll > 4000000000LL) {
            if (lane == 0) printf("pd_b200 attention_tc3: MMA warp starved (block %d,%d,%d tiles %d %d %d of %d)\n", blockIdx.x,
                                  blockIdx.y, blockIdx.z, jg0, jg1, jg2, a.ntiles);
            __trap();
          }
        }
      }
    }
  } else {
    // ---------------- softmax / correction / epilogue: thread == query row ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F3_REGS_SOFTMAX));
    const uint32_t tmem_base = read_tmem_base();
    const int g = warp >> 2;                       // query group
    const int qd4 = warp & 3;                      // TMEM lane quadrant this warp may touch
    const int r = qd4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd4 * 32) << 16;
    const uint32_t tmem_s = tmem_base + (uint32_t)(g * 128) + lane_off;     // P_g aliases S_g[:, 0:64)
    const uint32_t tmem_o = tmem_base + 384u + (uint32_t)(g * F3_OCOLS) + lane_off;
    const float sc = a.scale_log2;
    float m_ref = -INFINITY, l_run = 0.f;
    if (g < n_act) {
    {
      // phase offset between the groups: a third of a group's cycle each (see the MMA warp)
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)g * F3_STAGGER) { }
    }
    for (int j = 0; j < a.ntiles; ++j) {
      const bool last = j == a.ntiles - 1;
      mbar_wait(&s_full[g], (uint32_t)j & 1u, 500 + g);
      tc_fence_after();
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32p(tmem_s + (uint32_t)(c * 32), s + c * 32);
      tmem_ld_wait();
      if (qd4 == 0 && lane == 0) F3_DBG(4 + g, j);
      if (last && a.n_last_valid < F3_BK) {
        const int nv = a.n_last_valid;
#pragma unroll
        for (int e = 0; e < 128; ++e)
          if (e >= nv) s[e] = 0xff800000u;         // -inf: keys past Nk (stale / zero-filled columns)
      }
      // reference fixed at tile 0's exact row maximum; later tiles skip the max pass (attention_tc.cu)
      auto row_max = [&]() {
        float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]);
#pragma unroll
        for (int e = 2; e < 126; e += 4) {
          mx0 = fmax3(mx0, __uint_as_float(s[e]), __uint_as_float(s[e + 1]));
          mx1 = fmax3(mx1, __uint_as_float(s[e + 2]), __uint_as_float(s[e + 3]));
        }
        return fmax3(mx0, mx1, fmaxf(__uint_as_float(s[126]), __uint_as_float(s[127]))) * sc;
      };
#if F3_SWP
      // hand-pipelined in-place section (attention_tc.cu, FA_SWP): the reference is settled before it — exact row maximum
      // on every tile, O / l rescaled only when it grew by more than 2^8
      {
        const float mx_tile = row_max();
        bool grow = false;
        if (j == 0) m_ref = fmaxf(mx_tile, -1e30f);
        else grow = mx_tile > m_ref + 8.0f;
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          float corr = 1.0f;
          if (grow) { corr = ex2_approx(m_ref - mx_tile); m_ref = mx_tile; l_run *= corr; }
#pragma unroll
          for (int c = 0; c < F3_OCOLS; c += 8) {
            uint32_t o[8];
            tmem_ld8(tmem_o + (uint32_t)c, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * corr);
            tmem_st8(tmem_o + (uint32_t)c, o);
          }
        }
      }
#else
      if (j == 0) m_ref = fmaxf(row_max(), -1e30f);
#endif
      float lt = 0.f;
      auto exp_tile = [&]() {
        const float nm = -m_ref;
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#if F3_SWP
        constexpr int D = F3_SWP;
        auto X = [&](int q) { ffma2_b32_v(s[4 * q], s[4 * q + 1], sc, nm); ffma2_b32_v(s[4 * q + 2], s[4 * q + 3], sc, nm); };
        auto E = [&](int q, int k) { if (q < 32) ex2_b32_v(s[4 * q + k]); };
#pragma unroll
        for (int q = 0; q <= D; ++q) X(q);
#pragma unroll
        for (int q = 0; q < D; ++q) { E(q, 0); E(q, 1); E(q, 2); E(q, 3); }
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int e = q + D, c32 = (q >> 3) * 32, k = (2 * q) & 15;
          E(e, 0);
          fadd2_b32_v(l0, l1, s[4 * q], s[4 * q + 1]);
          E(e, 1);
          const uint32_t p0 = pack_bf16x2_b32_v(s[4 * q], s[4 * q + 1]);
          E(e, 2);
          fadd2_b32_v(l2, l3, s[4 * q + 2], s[4 * q + 3]);
          E(e, 3);
          const uint32_t p1 = pack_bf16x2_b32_v(s[4 * q + 2], s[4 * q + 3]);
          s[c32 + k] = p0; s[c32 + k + 1] = p1;     // in place: the consecutive block tcgen05.st wants
          if (e + 1 < 32) X(e + 1);
          if ((q & 7) == 7) tmem_st16p(tmem_s + (uint32_t)((q >> 3) * 16), s + c32);
        }
        lt = (l0 + l1) + (l2 + l3);
        return;
#endif
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const int i = c * 32 + e;
            float x0, x1, x2, x3;
            ffma2(x0, x1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), sc, sc, nm, nm);
            ffma2(x2, x3, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]), sc, sc, nm, nm);
            if (((e >> 1) & 7) >= 8 - F3_POLY) exp2_poly2(x0, x1); else { x0 = ex2_approx(x0); x1 = ex2_approx(x1); }
            if ((((e >> 1) + 1) & 7) >= 8 - F3_POLY) exp2_poly2(x2, x3); else { x2 = ex2_approx(x2); x3 = ex2_approx(x3); }
            fadd2(l0, l1, l0, l1, x0, x1);
            fadd2(l2, l3, l2, l3, x2, x3);
            pk[e >> 1] = pack_bf16x2(x0, x1);
            pk[(e >> 1) + 1] = pack_bf16x2(x2, x3);
          }
          tmem_st16p(tmem_s + (uint32_t)(c * 16), pk);   // P_g: 64 columns of bf16 pairs over S_g's first columns
        }
        lt = (l0 + l1) + (l2 + l3);
      };
      exp_tile();
      // overflow guard (warp-uniform: the rescale uses warp-collective tcgen05.ld / st).  s_full(j) was committed after
      // P V(j-1) in issue order, so O_g holds every tile < j and no MMA touches it before p_full(j).
      if (!F3_SWP && j > 0 && __any_sync(0xffffffffu, !(lt < F3_GROW_LIMIT))) {
        const float mx = row_max();
        float corr = 1.0f;
        if (mx > m_ref) { corr = ex2_approx(m_ref - mx); m_ref = mx; l_run *= corr; }
        tmem_st_wait();
#pragma unroll
        for (int c = 0; c < F3_OCOLS; c += 8) {
          uint32_t o[8];
          tmem_ld8(tmem_o + (uint32_t)c, o);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * corr);
          tmem_st8(tmem_o + (uint32_t)c, o);
        }
        exp_tile();
      }
      tmem_st_wait();
      l_run += lt;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      if (qd4 == 0 && lane == 0) F3_DBG(8 + g, j);
    }
    // ---------------- epilogue ----------------
    if (warp == 0 && lane == 0) griddep_launch();
    mbar_wait(&o_final[g], 0, 600 + g);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    unsigned char* stage_o = q_s + g * F3_Q_BYTES;                 // Q_g is dead: every Q K^T has retired
#pragma unroll
    for (int c = 0; c < F3_OCOLS; c += 8) {
      uint32_t o[8];
      tmem_ld8(tmem_o + (uint32_t)c, o);
      tmem_ld_wait();
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[e]) * inv;
      const int kc = c >> 3;                        // 16-byte chunk along the channel axis
      *reinterpret_cast<bf16x8*>(stage_o + r * 128 + (((kc & 7) ^ (r & 7)) << 4)) = pack8(f);
    }
    fence_proxy_async();
    asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
    if (qd4 == 0 && lane == 0 && q0 + g * F3_BQ < a.Nq) {
      tma_store_4d(&map_o, stage_o, 0, h, q0 + g * F3_BQ, b);
      tma_store_commit();
      tma_store_wait_all();
    }
    }   // g < n_act
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(read_tmem_base()), "r"(512));
  }
}

// Measured on B200.  Kernel alone (profiles/r02_attn3_active.txt, r02_attn3_stagger.txt, timelines r02_attn3_timeline_*):
// B16 h8 N4096 d40 711 us against 764 us on the two-group kernel, B32 h8 N9216 d40 6.46 against 7.28 ms — also under
// sustained load between its neighbouring GEMMs (800 vs 844 us per qkv-GEMM + attention + out-GEMM iteration).
// Inside the whole denoising step, same box, CUDA-graph replay (profiles/r02_attn3_insitu_ab.txt): config 4 (9216 tokens)
// 125.3 against 128.5 ms per step (-2.5 %), but config 2 (4096 tokens) 23.9 / 24.5 against 23.5 / 24.1 ms (+0.3 ms): the
// step runs at the 1 kW power cap and the 16-warp CTA with its polling MMA warp buys its 7 % with more power than the
// step has to spare.  Auto therefore picked this kernel from 6144 queries up — until the two-group kernel learnt to take
// turns on the SFU with a hand-scheduled exponential section (attention_tc.cu: 648 us / 6.05 ms at the two shapes above,
// profiles/r02_attn_bench_hand.txt), which beats this kernel at every size.  Auto no longer picks it: PD_B200_ATTN3=1
// restores the 6144-query rule, PD_B200_ATTN3=2 selects it for every d <= 40 shape, engine 6 selects it explicitly.
static int g_tc3_on = -1;     // -1: read PD_B200_ATTN3 once (default off)
bool attention_tc3_supported(int d, int Nq, int Nk) {
  if (g_tc3_on < 0) {
    const char* e = getenv("PD_B200_ATTN3");
    g_tc3_on = (e != nullptr && e[0] == '1') ? 1 : (e != nullptr && e[0] == '2') ? 2 : 0;
  }
  return g_tc3_on && d <= 40 && Nq >= (g_tc3_on == 2 ? 768 : 6144) && Nk >= 256;
}

int attention_tc3(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                  int heads, int Nq, int Nk, int d, float scale, cudaStream_t s) {
  F3Args a;
  a.dbg = g_fa_dbg_host;
  a.Nq = Nq; a.Nk = Nk;
  const int kpad = (d + 15) / 16 * 16;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.ntiles = (Nk + F3_BK - 1) / F3_BK;
  a.n_last_valid = Nk - (a.ntiles - 1) * F3_BK;
  const int n_last_pad = (a.n_last_valid + 15) / 16 * 16;
  // kind::f16 instruction descriptor: fp32 accumulate, bf16 A/B, M = 128 (see gemm_sm100.cu)
  const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
  a.idesc_s_full = base | ((uint32_t)(F3_BK >> 3) << 17);
  a.idesc_s_last = base | ((uint32_t)(n_last_pad >> 3) << 17);
  a.idesc_pv = base | (1u << 16) | ((uint32_t)(F3_OCOLS >> 3) << 17);   // B (= V tile) is MN-major; N = 40 channels

  CUtensorMap mq, mk, mv, mo;
  const uint32_t es[4] = {1, 1, 1, 1};
  struct { CUtensorMap* m; const void* p; int ld; int n; uint32_t rows; const char* nm; } t[4] = {
      {&mq, q, ldq, Nq, 128, "attn3Q"}, {&mk, k, ldk, Nk, 128, "attn3K"}, {&mv, v, ldv, Nk, 128, "attn3V"}, {&mo, out, ldo, Nq, 128, "attn3O"}};
  for (int i = 0; i < 4; ++i) {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)t[i].n, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)t[i].ld * 2, (uint64_t)t[i].n * t[i].ld * 2};
    const uint32_t box[4] = {64, 1, t[i].rows, 1};
    int rc = encode_map(t[i].m, t[i].p, 4, dims, strides, box, es, t[i].nm);
    if (rc) return rc;
  }
  const size_t smem = (size_t)F3_GROUPS * F3_Q_BYTES + 2 * F3_STAGES * F3_KV_BYTES + 256 + 1024;
  dim3 grid((Nq + F3_GROUPS * F3_BQ - 1) / (F3_GROUPS * F3_BQ), heads, B);
#define F3_LAUNCH(KP)                                                                                              \
  case KP: {                                                                                                       \
    static bool attr_set[16] = {false};                                                                            \
    int dev = 0;                                                                                                   \
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) { set_error("attention_tc3: bad device"); return PD_ERR_NO_DEVICE; } \
    if (!attr_set[dev]) {                                                                                          \
      cudaError_t e = cudaFuncSetAttribute(attention_tc3_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           (int)smem);                                                             \
      if (e != cudaSuccess) { set_error("attention_tc3: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; } \
      attr_set[dev] = true;                                                                                        \
    }                                                                                                              \
    cudaError_t le = launch_pdl(attention_tc3_kernel<KP>, grid, dim3(F3_THREADS), smem, s, 1, mq, mk, mv, mo, a);  \
    if (le != cudaSuccess) { set_error("attention_tc3: launch failed: %s", cudaGetErrorString(le)); return (int)le; } \
  } break;
  switch (kpad / 16) {
    F3_LAUNCH(1) F3_LAUNCH(2) F3_LAUNCH(3)
    default: set_error("attention_tc3: unsupported head dim %d", d); return PD_ERR_UNSUPPORTED;
  }
#undef F3_LAUNCH
  return check_launch("attention_tc3");
}

}  // namespace pd

// A/B switch: 0 = auto never picks the three-group kernel (engine 6 still selects it explicitly)
extern "C" int pd_debug_attention_tc3(int32_t on) { pd::g_tc3_on = on; return 0; }   // 0 off, 1 auto (>= 6144 queries), 2 every d <= 40 shape
