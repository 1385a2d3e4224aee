// tcgen05 / TMEM / TMA streaming-softmax attention (bf16) for head dims <= 192 (one query group per CTA above 128) — the d = 40, 4096..9216-token
// self-attention of the 64x64 / 96x96 latent level, the d = 80 level below it and their 77-key cross-attention
// (CrossAttention.forward, ldm/modules/attention.py:171-193), which dominate the attention time of a step.
//
// At d = 40 the tensor pipe needs ~380 cycles per 128x128 score tile but the exponentials need >= 1024 cycles of
// MUFU, so the kernel is organised around keeping the softmax threads busy, not the MMA:
//
//   one CTA per SM = 256 queries of one (batch, head) = two 128-query groups (A, B) sharing every K/V tile
//   (halves the L2 -> smem traffic), 320 threads:
//   warp 8    : TMA producer — Q (both groups) once, then K and V tiles of 128 keys into two independent rings.
//               The tensor maps view q/k/v as (d, heads, tokens, batch); a 64-wide box over a 40-wide head makes
//               TMA zero-fill channels 40..63, so no padding pass and no padded copies exist.
//   warp 9    : MMA issuer (warp-uniform loop, one ELECTED lane issues) — S_g[128x128] = Q_g K^T (smem x smem) into
//               TMEM; O_g[128xd] += P_g V with the A operand P_g read FROM TMEM (bf16, written by the softmax threads)
//               and the V tile used in place as the MN-major B operand.  d <= 64: P has its own TMEM columns and
//               Q K^T of tile j + 1 is issued as soon as S(j) sits in registers; above, P lies over S and the in-order
//               tensor pipe (P V(j) issued before Q K^T(j+1)) protects it.
//   warps 0-3 : softmax of group A, warps 4-7: group B — ONE THREAD PER QUERY ROW (TMEM lane == row): the row
//               max / sum need no shuffles.  S is read from TMEM exactly once (128 registers).
//               The two groups take STRICT TURNS on the SFU (named barriers 3 / 4, FlashAttention-3's ping-pong): a
//               group exponentiates alone on its sub-partitions while the other one does its round trip (P -> TMEM,
//               arrive, P V + next Q K^T on the tensor pipe, S -> registers, row maximum).  The section itself is
//               hand-scheduled in place on the S registers (FA_SWP / FA_HAND below): rounds of eight scores, three
//               pairs through MUFU.EX2, one through an FMA-pipe polynomial whose stages sit in the issue slots between
//               the MUFU issues; packed FFMA2 scale-and-subtract two rounds ahead, packed FADD2 row sum and bf16 pack
//               one round behind, P back to TMEM with tcgen05.st from the registers it was packed into.
//               Reference: exact row maximum on every tile, O / l rescaled lazily — only when the maximum grew by
//               more than 2^8 (p <= 256 is exact after the final division by l).
//   epilogue  : O / l -> bf16 -> swizzled smem (the dead Q tile) -> TMA store (clips columns >= d, rows >= Nq).
//
// THIS FILE IS COMPILED WITH -Xptxas -O1 (build.sh): at -O3 ptxas re-schedules the hand-written section (every polynomial
// chain in front of the first MUFU.EX2): 712 instead of 648 us at B16 h8 N4096 d40.  History of the design and every
// measured alternative: DESIGN.md section 4.0.
//
#include "tc_ptx.cuh"

namespace pd {

constexpr int FA_BQ = 128, FA_GROUPS = 2, FA_BK = 128, FA_THREADS = 320, FA_MAX_STAGES = 3;
constexpr int FA_TILE_BYTES = 128 * 128;   // one [128 rows][64 bf16] SWIZZLE_128B tile
constexpr int FA_ALIGN_SLACK = 1024;
constexpr float FA_GROW_LIMIT = 1.8446744e19f;   // 2^64: a tile row-sum at or above it (or NaN) triggers the exact-max path
// ---- build switches: the defaults are the shipped kernel; the others are measured alternatives (profiles/r02_attn_*) ----
#ifndef FA_ELECT
#define FA_ELECT 1         // 1: the MMA warp runs its loop warp-uniformly and one ELECTED lane issues (plain predicated UTCHMMA instead of the ELECT ... BRA.U.ANY retry loop ptxas emits inside `if (lane == 0)`)
#endif
#if FA_ELECT
#define FA_ISSUE if (elect_one())
#define FA_SYNCW __syncwarp()
#define FA_LANE0 (lane == 0)
#else
#define FA_ISSUE
#define FA_SYNCW
#define FA_LANE0 true
#endif
#ifndef FA_PARK
#define FA_PARK 0          // 1: the TMA and MMA warps wait on their mbarriers with a suspend-time hint (parked by the hardware)
#endif
#if FA_PARK
#define FA_ROLE_WAIT mbar_wait_parked
#else
#define FA_ROLE_WAIT mbar_wait
#endif
#ifndef FA_HANDOFF
#define FA_HANDOFF 31      // (FA_PINGPONG with FA_SWP) group of four whose exponentials are the last ones before the SFU turn passes to the other query group
#endif
#ifndef FA_HAND
#define FA_HAND 1          // 1 (with FA_SWP, needs -Xptxas -O1 for this file): hand-scheduled section, one polynomial pair per three SFU pairs
#endif
#ifndef FA_HANDOFF_R
#define FA_HANDOFF_R 15    // (FA_HAND) round of eight scores behind whose exponentials the SFU turn passes to the other group
#endif
#ifndef FA_TURN_PER_SMSP
#define FA_TURN_PER_SMSP 0 // 1: one pair of turn barriers per sub-partition (64 threads: the two softmax warps that share it) instead of one pair for the CTA (256 threads): 639 -> 653 us, off
#endif
#if FA_TURN_PER_SMSP
#define FA_TURN_ID(turn) (3 + 2 * qd4 + (turn))
#define FA_TURN_N 64
#else
#define FA_TURN_ID(turn) (3 + (turn))
#define FA_TURN_N 256
#endif
#ifndef FA_THROTTLE
#define FA_THROTTLE 0      // (FA_HAND) 1: a data dependency from the consumers of round r - 1 to the exponentials of round r + 1 (bounds the SFU queue): 640 -> 665 us, off
#endif
#ifndef FA_CDIST
#define FA_CDIST 1         // (FA_HAND) the row sum / bf16 pack of round r - FA_CDIST sit between the exponentials of round r
#endif
#ifndef FA_SPLIT
#define FA_SPLIT 0         // > 0 (with FA_SWP): a never-taken branch every FA_SPLIT groups of the pipelined exponential section
#endif
#ifndef FA_PINGPONG
#define FA_PINGPONG 1      // 1: the two query groups take strict turns in the exponential section (FlashAttention-3 style named barriers)
#endif
#ifndef FA_SWP
#define FA_SWP 2           // > 0: exponential section software-pipelined by hand, MUFU.EX2 issued FA_SWP groups of four ahead of their consumers; measured (profiles/r02_attn_swp_variants.txt, B16 h8 N4096 d40): 0 -> 761 us, 1..4 -> 776 us: the consumer stall behind each MUFU pair is NOT what holds the section at 61 % of the SFU bound
#endif
#ifndef FA_POLY
#define FA_POLY 0          // (without FA_HAND) of every 8 groups of four, how many take the FMA-pipe exp2; at -O3 ptxas hoists the chains in front of the SFU work: neutral or slower
#endif
unsigned long long* g_fa_dbg_host = nullptr;   // optional phase timeline (block 0 only), passed by value (shared with attention_tc3 / tc4.cu)

#define FA_DBG(slot, tile)                                                                       \
  do {                                                                                           \
    if (a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tile) < 32) \
      a.dbg[(slot) * 32 + (tile)] = gtimer();                                                    \
  } while (0)

struct FaArgs {
  int Nq, Nk, d;
  int nd;              // 64-channel chunks per head (1 or 2)
  int stages;          // depth of the K ring and of the V ring
  float scale_log2;
  uint32_t idesc_s_full, idesc_s_last, idesc_pv;
  int n_last_pad;      // S columns computed for the last K/V tile (multiple of 16)
  int n_last_valid;    // keys actually present in the last tile
  int ntiles;
  unsigned long long* dbg;   // phase-timeline buffer or nullptr (a kernel argument: no global load on the hot path)
};

// KP16 = ceil(d / 16): K extent of Q K^T and N extent of P V in units of 16 (compile time so that the single
// MMA-issuing thread runs a branch-free, fully unrolled instruction stream — it shares an SM sub-partition with
// two softmax warps and every issue slot it wastes delays the tensor pipe).
//
// TMEM (512 columns), d <= 64:  S_a [0,128) | S_b [128,256) | P_a [256,320) | P_b [320,384) | O_a [384,448) | O_b [448,512)
//   P has its own columns, so Q K^T of tile j+1 is issued as soon as the softmax threads have pulled S(j) into
//   registers (s_free) and the softmax threads never wait for the tensor pipe.
// 64 < d <= 128 (ALIAS):        S_a [0,128) | S_b [128,256) | O_a [256,384) | O_b [384,512), P_g over S_g[:, 0:64)
//   P aliases S; in-order execution of the tensor pipe (PV_g(j) issued before QK_g(j+1)) protects it.
// 128 < d <= 192 (NG = 1):       S [0,128) | O [128, 128 + KPAD), P over S[:, 0:64): ONE 128-query group per CTA (two O tiles
//   of 160 columns do not fit next to two S tiles), three 64-channel chunks per Q / K / V tile, K and V single-buffered
//   (144 KiB of tiles).  The d = 160 levels (16x16: 256 tokens, 8x8: 64 tokens, and their 77-key cross-attention) are
//   two key tiles at most: latency-bound, the point is the tcgen05 data path instead of the mma.sync kernel.
template <int KP16, int NG>
__global__ void __launch_bounds__(FA_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                    const FaArgs a) {
  constexpr bool ALIAS = KP16 > 4;
#ifndef FA_SWP_ALIAS
#define FA_SWP_ALIAS 1     // 1: the hand-scheduled section also for 64 < d <= 128 (P over S, two groups): 77.1 -> 72.8 us at B16 h8 N1024 d80
#endif
  constexpr bool SWP_ON = FA_SWP > 0 && (!ALIAS || (FA_SWP_ALIAS && NG == 2));      // the hand-pipelined exponential section
  constexpr int ND = (KP16 + 3) / 4;
  constexpr int KPAD = KP16 * 16;
  static_assert(NG == 2 || (NG == 1 && KP16 > 8 && KP16 <= 12), "one query group only for 128 < d <= 192");
  static_assert(NG == 1 || KP16 <= 8, "two query groups need d <= 128");
  constexpr uint32_t COL_S = 0, COL_P = ALIAS ? 0 : 256, COL_O = NG == 1 ? 128 : ALIAS ? 256 : 384;
  constexpr uint32_t GSTRIDE_P = ALIAS ? 128 : 64, GSTRIDE_O = ALIAS ? 128 : 64;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stages = a.stages;
  unsigned char* q_s = smem;                                              // [group][chunk] tiles
  unsigned char* k_s = q_s + NG * ND * FA_TILE_BYTES;              // [stage][chunk]
  unsigned char* v_s = k_s + stages * ND * FA_TILE_BYTES;                 // [stage][chunk]
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + stages * ND * FA_TILE_BYTES);
  uint64_t& q_full = bars[0];
  uint64_t* s_full = bars + 1;        // [2]  MMA -> softmax: S_g(j) complete
  uint64_t* p_full = bars + 3;        // [2]  softmax -> MMA: P_g(j) in TMEM, O_g rescaled
  uint64_t* o_final = bars + 5;       // [2]
  uint64_t* k_full = bars + 7;        // [3]
  uint64_t* k_empty = bars + 10;      // [3]
  uint64_t* v_full = bars + 13;       // [3]
  uint64_t* v_empty = bars + 16;      // [3]
  uint64_t* s_free = bars + 19;       // [2]  softmax -> MMA: S_g(j) is in registers (!ALIAS only)
  uint64_t* p_free = bars + 21;       // [2]  MMA -> softmax: P_g(j) V(j) retired (!ALIAS only)
  uint32_t& tmem_base_slot = *reinterpret_cast<uint32_t*>(bars + 23);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (NG * FA_BQ), h = blockIdx.y, b = blockIdx.z;
  // Roles: warps 0-3 softmax A, 4-7 softmax B, 8 TMA, 9 MMA (the sub-partition arbiter favours high warp ids)
  constexpr int W_TMA = 8, W_MMA = 9;

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); tma_prefetch_desc(&map_o);
    mbar_init(&q_full, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 4); mbar_init(&o_final[g], 1);
      mbar_init(&s_free[g], 4); mbar_init(&p_free[g], 1);
    }
    for (int i = 0; i < FA_MAX_STAGES; ++i) {
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
      mbar_expect_tx(&q_full, NG * ND * FA_TILE_BYTES);
      for (int g = 0; g < NG; ++g)
        for (int c = 0; c < ND; ++c)
          tma_load_4d(q_s + (g * ND + c) * FA_TILE_BYTES, &map_q, &q_full, c * 64, h, q0 + g * FA_BQ, b);
      int st = 0; uint32_t ph = 0;
      for (int j = 0; j < a.ntiles; ++j) {
        FA_ROLE_WAIT(&k_empty[st], ph ^ 1u, 100 + st);
        mbar_expect_tx(&k_full[st], ND * FA_TILE_BYTES);
        for (int c = 0; c < ND; ++c)
          tma_load_4d(k_s + (st * ND + c) * FA_TILE_BYTES, &map_k, &k_full[st], c * 64, h, j * FA_BK, b);
        FA_ROLE_WAIT(&v_empty[st], ph ^ 1u, 110 + st);
        mbar_expect_tx(&v_full[st], ND * FA_TILE_BYTES);
        for (int c = 0; c < ND; ++c)
          tma_load_4d(v_s + (st * ND + c) * FA_TILE_BYTES, &map_v, &v_full[st], c * 64, h, j * FA_BK, b);
        if (++st == stages) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == W_MMA) {
    // Single-lane issue ON PURPOSE.  Inside `if (lane == 0)` ptxas wraps every UTCHMMA / UTCBAR in an ELECT ... BRA.U.ANY
    // retry loop (~150 clk per MMA); the warp-uniform elect.sync form used by the GEMM engine and by attention_tc3 / tc4
    // issues faster, but here the slow issue is what keeps the two query groups OUT of phase: measured at B16 h8 N4096
    // d40 (profiles/r02_attn_mma_modes.txt) 762 us as written, 788 us with elect.sync issue, 877 us with elect.sync
    // and the two groups' Q K^T / P V streams interleaved (both groups then exponentiate at the same time and wait at
    // the same time).
    if (FA_ELECT || lane == 0) {
      // descriptors: only the 14-bit start-address field changes between tiles, so every MMA operand is
      // base + a compile-time or per-stage constant in the low word
      const uint64_t qdesc0 = make_smem_desc(s_u32(q_s));
      const uint64_t kdesc0 = make_smem_desc(s_u32(k_s));
      const uint64_t vdesc0 = make_smem_desc_mn(s_u32(v_s), FA_TILE_BYTES, 1024);
      constexpr uint64_t TILE16 = FA_TILE_BYTES >> 4;
      const uint32_t tm_s = tmem_base + COL_S, tm_p = tmem_base + COL_P, tm_o = tmem_base + COL_O;
      // S_g = Q_g K^T : K extent KPAD (channels d..KPAD-1 of both operands are TMA zero fill)
      auto issue_qk = [&](int g, int st_k, uint32_t idesc) {
        const uint64_t qd = qdesc0 + (uint64_t)(g * ND) * TILE16, kd = kdesc0 + (uint64_t)(st_k * ND) * TILE16;
#pragma unroll
        for (int k = 0; k < KP16; ++k) {
          const uint64_t off = (uint64_t)(k >> 2) * TILE16 + (uint64_t)(k & 3) * 2u;
          umma_bf16(tm_s + (uint32_t)(g * 128), qd + off, kd + off, idesc, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[g]);
      };
      // O_g += P_g V : A = P_g from TMEM (8 columns per 16-key step), B = V tile in place, MN-major
      auto issue_pv = [&](int g, int st_v, bool first, bool last) {
        const uint64_t vd = vdesc0 + (uint64_t)(st_v * ND) * TILE16;
        const uint32_t pa = tm_p + (uint32_t)g * GSTRIDE_P, oa = tm_o + (uint32_t)g * GSTRIDE_O;
        if (!last) {
#pragma unroll
          for (int k = 0; k < FA_BK / 16; ++k)
            umma_bf16_ts(oa, pa + (uint32_t)(8 * k), vd + (uint64_t)(k * 128), a.idesc_pv, (k != 0 || !first) ? 1u : 0u);
        } else {
          const int ksteps = (a.n_last_valid + 15) / 16;
          for (int k = 0; k < ksteps; ++k)
            umma_bf16_ts(oa, pa + (uint32_t)(8 * k), vd + (uint64_t)(k * 128), a.idesc_pv, (k != 0 || !first) ? 1u : 0u);
        }
      };
      FA_ROLE_WAIT(&q_full, 0, 200);
      FA_ROLE_WAIT(&k_full[0], 0, 300);
      tc_fence_after();
      if (FA_LANE0) FA_DBG(0, 0);
      const uint32_t id0 = a.ntiles == 1 ? a.idesc_s_last : a.idesc_s_full;
      FA_ISSUE {
        for (int g = 0; g < NG; ++g) issue_qk(g, 0, id0);
        umma_commit(&k_empty[0]);
      }
      FA_SYNCW;
      int st = 0; uint32_t ph = 0;          // ring position of tile j
      for (int j = 0; j < a.ntiles; ++j) {
        int stn = st + 1; uint32_t phn = ph;
        if (stn == stages) { stn = 0; phn ^= 1u; }
        const bool more = j + 1 < a.ntiles, last = !more;
        const uint32_t idn = (j + 2 == a.ntiles) ? a.idesc_s_last : a.idesc_s_full;
        if constexpr (!ALIAS) {
          if (more) {
            FA_ROLE_WAIT(&k_full[stn], phn, 300 + stn);
            for (int g = 0; g < NG; ++g) {
              FA_ROLE_WAIT(&s_free[g], (uint32_t)j & 1u, 450 + g);
              tc_fence_after();
              FA_ISSUE {
                issue_qk(g, stn, idn);
                if (g == NG - 1) umma_commit(&k_empty[stn]);
              }
              FA_SYNCW;
              if (g == 0 && FA_LANE0) FA_DBG(0, j + 1);
            }
          }
          FA_ROLE_WAIT(&v_full[st], ph, 310 + st);
          for (int g = 0; g < NG; ++g) {
            FA_ROLE_WAIT(&p_full[g], (uint32_t)j & 1u, 400 + g);
            tc_fence_after();
            if (g == 0 && FA_LANE0) FA_DBG(1, j);
            FA_ISSUE {
              issue_pv(g, st, j == 0, last);
              umma_commit(last ? &o_final[g] : &p_free[g]);
              if (g == NG - 1) umma_commit(&v_empty[st]);
            }
            FA_SYNCW;
          }
        } else {
          FA_ROLE_WAIT(&v_full[st], ph, 310 + st);
          if (more) FA_ROLE_WAIT(&k_full[stn], phn, 300 + stn);
          for (int g = 0; g < NG; ++g) {
            FA_ROLE_WAIT(&p_full[g], (uint32_t)j & 1u, 400 + g);
            tc_fence_after();
            if (g == 0 && FA_LANE0) FA_DBG(1, j);
            FA_ISSUE {
              issue_pv(g, st, j == 0, last);
              if (g == NG - 1) umma_commit(&v_empty[st]);
              if (more) {
                issue_qk(g, stn, idn);
                if (g == NG - 1) umma_commit(&k_empty[stn]);
              } else {
                umma_commit(&o_final[g]);
              }
            }
            FA_SYNCW;
            if (more && g == 0 && FA_LANE0) FA_DBG(0, j + 1);
          }
        }
        st = stn; ph = phn;
      }
    }
  } else if ((warp >> 2) < NG) {
    // ---------------- softmax / correction / epilogue: thread == query row ----------------
    const int g = warp >> 2;                       // 0: warps 0-3, 1: warps 4-7 (idle when NG == 1)
    const int qd4 = warp & 3;                      // TMEM lane quadrant this warp may touch
    const int r = qd4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd4 * 32) << 16;
    const uint32_t tmem_s = tmem_base + COL_S + (uint32_t)(g * 128) + lane_off;
    const uint32_t tmem_p = tmem_base + COL_P + (uint32_t)g * GSTRIDE_P + lane_off;
    const uint32_t tmem_o = tmem_base + COL_O + (uint32_t)g * GSTRIDE_O + lane_off;
    const bool dbg = (warp & 3) == 0 && lane == 0;
    const int dslot = 2 + 4 * g;
    const float sc = a.scale_log2;
    float m_ref = -INFINITY, l_run = 0.f;
#if FA_PINGPONG
    if constexpr (NG == 2) { if (g == 1) asm volatile("bar.arrive %0, %1;" ::"r"(FA_TURN_ID(0)), "r"(FA_TURN_N) : "memory"); }   // group A goes first
#endif
    for (int j = 0; j < a.ntiles; ++j) {
      const bool last = j == a.ntiles - 1;
      mbar_wait(&s_full[g], (uint32_t)j & 1u, 500 + g);
      tc_fence_after();
      if (dbg) FA_DBG(dslot, j);
      uint32_t s[128];
      auto load_s = [&]() {
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32p(tmem_s + (uint32_t)(c * 32), s + c * 32);
        tmem_ld_wait();
        if (last && a.n_last_valid < FA_BK) {
          const int nv = a.n_last_valid;
#pragma unroll
          for (int e = 0; e < 128; ++e)
            if (e >= nv) s[e] = 0xff800000u;         // -inf: keys past Nk (stale / zero-filled columns)
        }
      };
      auto release_s = [&]() {                       // S_g may be overwritten by Q K^T of the next tile
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[g]);
      };
      load_s();
      if constexpr (!ALIAS) release_s();
      // ---- reference for the exponentials ------------------------------------------------------------------------
      // P is bf16 and l / O accumulate in fp32, all with 8 exponent bits, so the reference m_ref only has to keep
      // exp2(s * sc - m_ref) inside the fp32 range — it need not track the running row maximum.  Tile 0 fixes m_ref at
      // its exact row maximum; later tiles SKIP the max pass (64 FMNMX3 of the ~420 instructions a thread issues per
      // tile, on a section that is issue-bound, not SFU-bound) and exponentiate against the stale reference.  A row
      // whose scores outgrow it by more than 2^64 shows up in its tile row-sum; only then (never on ordinary data
      // after tile 0) the tile is redone behind an exact max + O / l rescale.
      auto row_max = [&]() {
        float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]);
#pragma unroll
        for (int e = 2; e < 126; e += 4) {
          mx0 = fmax3(mx0, __uint_as_float(s[e]), __uint_as_float(s[e + 1]));
          mx1 = fmax3(mx1, __uint_as_float(s[e + 2]), __uint_as_float(s[e + 3]));
        }
        return fmax3(mx0, mx1, fmaxf(__uint_as_float(s[126]), __uint_as_float(s[127]))) * sc;
      };
      // The hand-pipelined section (SWP_ON) exponentiates IN PLACE and S_g is already released, so a tile cannot be
      // redone: there the reference is settled BEFORE the section, with an exact row maximum on every tile (64 FMNMX3 on
      // the ALU pipe, issued while the other query group owns the SFU) and the lazy rescale of the online softmax —
      // O / l move to the new maximum only when it grew by more than 2^8 (p <= 256 is exact after the division by l).
      bool grow = false;
      float mx_tile = 0.f;
      if constexpr (SWP_ON) {
        mx_tile = row_max();
        if (j == 0) m_ref = fmaxf(mx_tile, -1e30f);
        else grow = mx_tile > m_ref + 8.0f;
      } else {
        if (j == 0) m_ref = fmaxf(row_max(), -1e30f);        // (a fully masked row keeps a finite reference)
      }
      if (dbg) FA_DBG(dslot + 1, j);
      if (j > 0) {
        if constexpr (!ALIAS) {                    // P_g(j-1) V(j-1) retired: P_g is free, O_g is complete
          mbar_wait(&p_free[g], (uint32_t)(j - 1) & 1u, 550 + g);
          tc_fence_after();
        }
        // (ALIAS: s_full(j) was committed after P V(j-1) in issue order, so O_g is complete here too)
      }
      if constexpr (SWP_ON) {
        if (j > 0 && __any_sync(0xffffffffu, grow)) {         // warp-uniform: the rescale is warp-collective tcgen05.ld / st
          float corr = 1.0f;
          if (grow) { corr = ex2_approx(m_ref - mx_tile); m_ref = mx_tile; l_run *= corr; }
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
      }
      float lt = 0.f;
      // scale-and-subtract (packed FFMA2), exp2, row sum (packed FADD2), bf16 pack, P -> TMEM.  FA_POLY of every 8
      // element pairs take the polynomial exp2 on the FMA pipe, the rest MUFU.EX2.
      auto exp_tile = [&]() {
        const float nm = -m_ref;
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        if constexpr (SWP_ON) {
        // Software pipeline over groups of four scores, in place on the S registers: the exponentials of group g + FA_SWP
        // are issued between the FADD2 / F2FP that consume group g and the FFMA2 of group g + FA_SWP + 1.  In the plain
        // loop below ptxas puts every FADD2 / F2FP right behind its own two MUFU.EX2, so the in-order warp idles for the
        // SFU latency once per pair (one warp alone: 16 clk per element against the SFU's 8).  FA_POLY of every 8 groups
        // take the FMA-pipe polynomial instead of the SFU.  The bf16 pairs are packed IN PLACE into the first 16
        // registers of each 32-score chunk (pair k of a chunk overwrites score k, consumed by pair k / 2 <= k), which is
        // the consecutive block tcgen05.st wants: no register shuffling in front of the store.
        constexpr int D = FA_SWP > 0 ? FA_SWP : 1;
        const int g_turn = g;                         // (the loops below reuse the name g for the group of four)
#if FA_HAND
        // HAND-SCHEDULED section (this file is compiled with -Xptxas -O1, which keeps the order written here; at -O3 ptxas
        // hoists every polynomial chain in front of the first MUFU.EX2 and the SFU then idles behind the FMA pipe).
        // Rounds of eight scores = four pairs: pairs 0..2 go through the SFU, pair 3 through the FMA-pipe polynomial, whose
        // six dependent stages sit one by one in the issue slots between the six MUFU.EX2 (8 clk of SFU each).  Round r
        // issues the exponentials of round r + 1, the scale-and-subtract of round r + 2 and consumes (row sum, bf16 pack)
        // round r: 8 exponentials per 48 SFU clk instead of 6.
        {
          auto X8 = [&](int r) {
#pragma unroll
            for (int k = 0; k < 8; k += 2) ffma2_b32_v(s[8 * r + k], s[8 * r + k + 1], sc, nm);
          };
          uint32_t t0 = 0, t1 = 0, n0 = 0, n1 = 0, q0 = 0, q1 = 0;
          uint32_t u0 = 0, u1 = 0, m0 = 0, m1 = 0, w0 = 0, w1 = 0;       // second polynomial chain (rounds with two of them)
          auto C = [&](int r, int pr, float& la, float& lb) {          // consume pair pr of round r
            const int i = 8 * r + 2 * pr, c32 = (r >> 2) * 32, k = (4 * r + pr) & 15;
            fadd2_b32_v(la, lb, s[i], s[i + 1]);
            const uint32_t pk = pack_bf16x2_b32_v(s[i], s[i + 1]);
            s[c32 + k] = pk;
          };
          // FA_HAND == 2: every third round sends pair 2 through the polynomial as well (a third of the exponentials off
          // the SFU: SFU and FMA pipe then carry about the same number of cycles per score)
          auto two_poly = [](int r) { return FA_HAND == 2 && (r % 3) == 2; };
          // exponentials of round r (MUFU pairs 0..2, polynomial pair 3), with the consumers of round r - 1 in between
          constexpr int CD = FA_CDIST;                  // rounds between a pair's exponentials and its consumers
          auto round = [&](int r, bool consume, bool scale_next) {
            const int b = 8 * r;
            const bool ex = r < 16, tp = ex && two_poly(r);
            if (ex) { ex2_b32_v(s[b]); poly_s1(s[b + 6], s[b + 7], t0, t1); }
            if (tp) poly_s1(s[b + 4], s[b + 5], u0, u1);
            if (ex) ex2_b32_v(s[b + 1]);
            if (consume) C(r - CD, 0, l0, l1);
            if (ex) { ex2_b32_v(s[b + 2]); poly_s2a(n0, n1, t0, t1); }
            if (tp) poly_s2a(m0, m1, u0, u1);
            if (consume) C(r - CD, 1, l2, l3);
            if (ex) { ex2_b32_v(s[b + 3]); poly_s2b(s[b + 6], s[b + 7], n0, n1); }
            if (tp) poly_s2b(s[b + 4], s[b + 5], m0, m1);
#if FA_THROTTLE
            // the scale-and-subtract of round r + 1 takes its addend through the row sum as it stands behind round r - 1's
            // consumers: the MUFU.EX2 of round r + 1 cannot be scheduled (or issued) ahead of them, so the SFU queue never
            // holds more than two rounds and the consumers find their operands ready
            if (scale_next && r + 1 < 16) {
              float d0, d1;
              if (consume) dep_on(d0, d1, l0, l1, nm); else { d0 = nm; d1 = nm; }
              ffma2_b32_v2(s[b + 8], s[b + 9], sc, d0, d1); ffma2_b32_v2(s[b + 10], s[b + 11], sc, d0, d1);
            }
#else
            if (scale_next && r + 1 < 16) { ffma2_b32_v(s[b + 8], s[b + 9], sc, nm); ffma2_b32_v(s[b + 10], s[b + 11], sc, nm); }
#endif
            if (ex && !tp) ex2_b32_v(s[b + 4]);
            if (ex) poly_s3(q0, q1, s[b + 6], s[b + 7]);
            if (tp) poly_s3(w0, w1, s[b + 4], s[b + 5]);
            if (consume) C(r - CD, 2, l0, l1);
            if (ex && !tp) ex2_b32_v(s[b + 5]);
            if (ex) poly_s45<4>(q0, q1, s[b + 6], s[b + 7]);
            if (tp) poly_s45<4>(w0, w1, s[b + 4], s[b + 5]);
#if FA_THROTTLE
            if (scale_next && r + 1 < 16) {
              float d0, d1;
              if (consume) dep_on(d0, d1, l2, l3, nm); else { d0 = nm; d1 = nm; }
              ffma2_b32_v2(s[b + 12], s[b + 13], sc, d0, d1); ffma2_b32_v2(s[b + 14], s[b + 15], sc, d0, d1);
            }
#else
            if (scale_next && r + 1 < 16) { ffma2_b32_v(s[b + 12], s[b + 13], sc, nm); ffma2_b32_v(s[b + 14], s[b + 15], sc, nm); }
#endif
            if (ex) poly_s45<5>(q0, q1, s[b + 6], s[b + 7]);
            if (tp) poly_s45<5>(w0, w1, s[b + 4], s[b + 5]);
            if (consume) C(r - CD, 3, l2, l3);
            if (ex) poly_s6(s[b + 6], s[b + 7], q0, q1, t0, t1);
            if (tp) poly_s6(s[b + 4], s[b + 5], w0, w1, u0, u1);
            if (consume && ((r - CD) & 3) == 3) tmem_st16p(tmem_p + (uint32_t)(((r - CD) >> 2) * 16), s + ((r - CD) >> 2) * 32);
          };
          X8(0);
#if FA_PINGPONG
          if constexpr (NG == 2) asm volatile("bar.sync %0, %1;" ::"r"(FA_TURN_ID(g_turn)), "r"(FA_TURN_N) : "memory");
#endif
#pragma unroll
          for (int r = 0; r < 16 + CD; ++r) {
            round(r, r >= CD, true);
#if FA_PINGPONG
            if constexpr (NG == 2) { if (r == FA_HANDOFF_R) asm volatile("bar.arrive %0, %1;" ::"r"(FA_TURN_ID(g_turn ^ 1)), "r"(FA_TURN_N) : "memory"); }
#endif
          }
          lt = (l0 + l1) + (l2 + l3);
          return;
        }
#endif
        auto is_poly = [](int g) {
          const int r = g & 7;
          return FA_POLY == 1 ? r == 3 : FA_POLY == 2 ? (r == 1 || r == 5) : FA_POLY == 3 ? (r == 1 || r == 4 || r == 6)
               : FA_POLY == 4 ? (r & 1) == 1 : FA_POLY >= 5 ? r != 0 && r != 3 && r != 6 : false;
        };
        auto X = [&](int g) { ffma2_b32_v(s[4 * g], s[4 * g + 1], sc, nm); ffma2_b32_v(s[4 * g + 2], s[4 * g + 3], sc, nm); };
        auto P = [&](int g, int h) {                  // polynomial exp2 of pair h of group g
          float x0 = __uint_as_float(s[4 * g + 2 * h]), x1 = __uint_as_float(s[4 * g + 2 * h + 1]);
          exp2_poly2(x0, x1);
          s[4 * g + 2 * h] = __float_as_uint(x0); s[4 * g + 2 * h + 1] = __float_as_uint(x1);
        };
        auto E = [&](int g, int k) {                  // k-th quarter of group g's exponentials
          if (g >= 32) return;
          if (!is_poly(g)) ex2_b32_v(s[4 * g + k]);
          else if (k == 0) P(g, 0);
          else if (k == 2) P(g, 1);
        };
#pragma unroll
        for (int g = 0; g <= D; ++g) X(g);
#if FA_PINGPONG
        // the group's turn on the SFU starts here (the FFMA2 above need no SFU) ...
        if constexpr (NG == 2) asm volatile("bar.sync %0, %1;" ::"r"(FA_TURN_ID(g_turn)), "r"(FA_TURN_N) : "memory");
#endif
#pragma unroll
        for (int g = 0; g < D; ++g) { E(g, 0); E(g, 1); E(g, 2); E(g, 3); }
#pragma unroll
        for (int g = 0; g < 32; ++g) {
          const int e = g + D;                       // group whose exponentials are issued in this round
          const int c32 = (g >> 3) * 32, k = (2 * g) & 15;
          E(e, 0);
          fadd2_b32_v(l0, l1, s[4 * g], s[4 * g + 1]);
          E(e, 1);
          const uint32_t p0 = pack_bf16x2_b32_v(s[4 * g], s[4 * g + 1]);
          E(e, 2);
          fadd2_b32_v(l2, l3, s[4 * g + 2], s[4 * g + 3]);
          E(e, 3);
#if FA_PINGPONG
          // ... and ends behind its last exponential: the remaining FADD2 / F2FP / tcgen05.st overlap the other group's start
          if constexpr (NG == 2) { if (e == FA_HANDOFF) asm volatile("bar.arrive %0, %1;" ::"r"(FA_TURN_ID(g_turn ^ 1)), "r"(FA_TURN_N) : "memory"); }
#endif
          const uint32_t p1 = pack_bf16x2_b32_v(s[4 * g + 2], s[4 * g + 3]);
          s[c32 + k] = p0; s[c32 + k + 1] = p1;
          if (e + 1 < 32) X(e + 1);
          if ((g & 7) == 7) tmem_st16p(tmem_p + (uint32_t)((g >> 3) * 16), s + c32);
#if FA_SPLIT > 0
          // basic-block boundary (a branch that is never taken): ptxas schedules inside a block, so the order written
          // here — SFU exponentials, their consumers and the polynomial chains in the same rounds — cannot be undone by
          // hoisting every polynomial chain in front of the first MUFU.EX2 or sinking every F2FP behind the last one
          if ((g % FA_SPLIT) == FA_SPLIT - 1 && g != 31) {
            if (a.Nq == -2 - g) asm volatile("trap;");   // a different impossible value each time: not foldable
          }
#endif
        }
        lt = (l0 + l1) + (l2 + l3);
        return;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const int i = c * 32 + e;
            float x0, x1, x2, x3;
            ffma2(x0, x1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), sc, sc, nm, nm);
            ffma2(x2, x3, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]), sc, sc, nm, nm);
            // pairs are numbered (e >> 1) and (e >> 1) + 1 within the 16-pair chunk; pair k mod 8 >= 8 - FA_POLY -> polynomial
            if (((e >> 1) & 7) >= 8 - FA_POLY) exp2_poly2(x0, x1); else { x0 = ex2_approx(x0); x1 = ex2_approx(x1); }
            if ((((e >> 1) + 1) & 7) >= 8 - FA_POLY) exp2_poly2(x2, x3); else { x2 = ex2_approx(x2); x3 = ex2_approx(x3); }
            fadd2(l0, l1, l0, l1, x0, x1);
            fadd2(l2, l3, l2, l3, x2, x3);
            pk[e >> 1] = pack_bf16x2(x0, x1);
            pk[(e >> 1) + 1] = pack_bf16x2(x2, x3);
          }
          tmem_st16p(tmem_p + (uint32_t)(c * 16), pk);   // P_g: 64 columns of bf16 pairs
        }
        lt = (l0 + l1) + (l2 + l3);
      };
#if FA_PINGPONG
      // strict alternation of the two groups' exponential sections (named barriers 3 / 4 over the 256 softmax threads):
      // a group exponentiates alone on its sub-partitions' SFUs while the other one does its TMEM / barrier round trip
      if constexpr (NG == 2 && !SWP_ON) asm volatile("bar.sync %0, %1;" ::"r"(FA_TURN_ID(g)), "r"(FA_TURN_N) : "memory");
#endif
      exp_tile();
#if FA_PINGPONG
      if constexpr (NG == 2 && !SWP_ON) asm volatile("bar.arrive %0, %1;" ::"r"(FA_TURN_ID(g ^ 1)), "r"(FA_TURN_N) : "memory");
#endif
      // overflow guard (warp-uniform: the rescale uses warp-collective tcgen05.ld / st)
      if (!SWP_ON && j > 0 && __any_sync(0xffffffffu, !(lt < FA_GROW_LIMIT))) {
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
      if (dbg) FA_DBG(dslot + 2, j);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      if (dbg) FA_DBG(dslot + 3, j);
    }
    // ---------------- epilogue ----------------
    if (warp == 0 && lane == 0) griddep_launch();
    mbar_wait(&o_final[g], 0, 600 + g);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    unsigned char* stage_o = q_s + (g * ND) * FA_TILE_BYTES;       // Q_g is dead: every Q K^T has retired
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
        *reinterpret_cast<bf16x8*>(stage_o + (kc >> 3) * FA_TILE_BYTES + r * 128 + (((kc & 7) ^ (r & 7)) << 4)) = pack8(f);
      }
    }
    fence_proxy_async();
    asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
    if (qd4 == 0 && lane == 0 && q0 + g * FA_BQ < a.Nq) {
      for (int c = 0; c < ND; ++c) tma_store_4d(&map_o, stage_o + c * FA_TILE_BYTES, c * 64, h, q0 + g * FA_BQ, b);
      tma_store_commit();
      tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

bool attention_tc_supported(int dtype, int d, int ldq, int ldk, int ldv, int ldo, const void* q, const void* k,
                            const void* v, const void* out) {
  static int sm100 = -1;
  if (sm100 < 0) sm100 = pd_device_is_sm100();
  return sm100 && dtype == PD_BF16 && d % 8 == 0 && d >= 16 && d <= 192 && ldq % 8 == 0 && ldk % 8 == 0 &&
         ldv % 8 == 0 && ldo % 8 == 0 && ((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 &&
         ((uintptr_t)v % 16) == 0 && ((uintptr_t)out % 16) == 0;
}

int attention_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                 int heads, int Nq, int Nk, int d, float scale, cudaStream_t s) {
  FaArgs a;
  a.Nq = Nq; a.Nk = Nk; a.d = d;
  a.nd = (d + 63) / 64;
  a.stages = a.nd == 1 ? 3 : a.nd == 2 ? 2 : 1;     // K ring and V ring depth (144 KiB of tiles at three chunks per head)
  const int ng = d > 128 ? 1 : FA_GROUPS;
  const int kpad = (d + 15) / 16 * 16;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.ntiles = (Nk + FA_BK - 1) / FA_BK;
  a.n_last_valid = Nk - (a.ntiles - 1) * FA_BK;
  a.n_last_pad = (a.n_last_valid + 15) / 16 * 16;
  a.dbg = g_fa_dbg_host;
  // kind::f16 instruction descriptor: fp32 accumulate, bf16 A/B, M = 128 (see gemm_sm100.cu)
  const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
  a.idesc_s_full = base | ((uint32_t)(FA_BK >> 3) << 17);
  a.idesc_s_last = base | ((uint32_t)(a.n_last_pad >> 3) << 17);
  a.idesc_pv = base | (1u << 16) | ((uint32_t)(kpad >> 3) << 17);   // B (= V tile) is MN-major

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
  const size_t smem = (size_t)(ng * a.nd + 2 * a.stages * a.nd) * FA_TILE_BYTES + 256 + FA_ALIGN_SLACK;
  dim3 grid((Nq + ng * FA_BQ - 1) / (ng * FA_BQ), heads, B);
#define FA_LAUNCH(KP, NGv)                                                                                         \
  case KP: {                                                                                                       \
    static bool attr_set = false;                                                                                  \
    if (!attr_set) {                                                                                               \
      cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<KP, NGv>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                           (int)smem);                                                             \
      if (e != cudaSuccess) { set_error("attention_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; } \
      attr_set = true;                                                                                             \
    }                                                                                                              \
    cudaError_t le = launch_pdl(attention_tc_kernel<KP, NGv>, grid, dim3(FA_THREADS), smem, s, 1, mq, mk, mv, mo, a);   \
    if (le != cudaSuccess) { set_error("attention_tc: launch failed: %s", cudaGetErrorString(le)); return (int)le; } \
  } break;
  switch (kpad / 16) {
    FA_LAUNCH(1, 2) FA_LAUNCH(2, 2) FA_LAUNCH(3, 2) FA_LAUNCH(4, 2) FA_LAUNCH(5, 2) FA_LAUNCH(6, 2) FA_LAUNCH(7, 2) FA_LAUNCH(8, 2)
    FA_LAUNCH(9, 1) FA_LAUNCH(10, 1) FA_LAUNCH(11, 1) FA_LAUNCH(12, 1)
    default: set_error("attention_tc: unsupported head dim %d", d); return PD_ERR_UNSUPPORTED;
  }
#undef FA_LAUNCH
  return check_launch("attention_tc");
}

}  // namespace pd

// debugging aid: device buffer of 10*32 uint64 receiving block (0,0,0)'s per-tile phase stamps; NULL = off
extern "C" int pd_debug_attention_timeline(void* dev_buf) {
  pd::g_fa_dbg_host = (unsigned long long*)dev_buf;
  return 0;
}
