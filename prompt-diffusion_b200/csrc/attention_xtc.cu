// tcgen05 / TMEM / TMA attention over a SHORT key sequence (Nk <= 128, d <= 128): the 77-token cross-attention of every
// SpatialTransformer (CrossAttention.forward with context, ldm/modules/attention.py:163-194) at the 64x64 / 32x32
// (96x96 / 48x48) levels — 14 calls per denoising step that move every query and output row once.
//
// One key tile means no online softmax and no rescale; what the op needs is to keep the loads, the two small MMAs, the
// exponentials and the stores of MANY query tiles in flight at once.  The streaming kernel (attention_tc.cu) runs one
// 256-query CTA per launch slot and pays its whole serial chain (Q/K/V load -> Q K^T -> softmax -> P V -> store, ~6.5 us)
// per CTA: 91 us at B16 h8 N4096 d40 for 13 us of HBM traffic; the mma.sync kernel (attention_short.cu) takes 56 us.
// This kernel is PERSISTENT and warp-specialised:
//
//   grid = one CTA per SM; the (batch, head, 256-query pair) units are cut into equal contiguous ranges.  A CTA loads
//   K and V of a (batch, head) ONCE (TMA, 128-key box: rows >= Nk arrive as zeros) and re-loads them only when its
//   range crosses into the next head; Q pairs stream through a ring of 2-3 slots.
//   warp 16   : TMA producer (K / V on a head change, Q pair of unit i + slots ahead)
//   warp 17   : MMA issuer — S_g = Q_g K^T (N = keys padded to 16) into TMEM, O_g = P_g V with P_g read from TMEM
//               (over S_g's first columns) and V used in place as the MN-major B operand.  Per unit and group:
//               P V_g(i), then Q K^T_g(i + 1): in-order execution of the tensor pipe protects the aliased P_g.
//   warps 0-3 / 4-7 : SOFTMAX of query groups A / B, one thread per query row: S -> registers, mask, row max, exp2, row
//               sum, P -> TMEM, 1 / l -> shared memory.
//   warps 8-11 / 12-15 : EPILOGUE of groups A / B: O_g / l -> bf16 -> swizzled staging IN THE Q SLOT of the unit (Q is
//               dead once S exists) -> TMA store; the slot goes back to the producer when both groups' stores have read
//               it.  Separate warps, so that the softmax of unit i + 1 runs while unit i is normalised and stored: with
//               one set of warps doing both, a unit cost its whole chain (S wait, softmax, P V, O wait, store: 4.2 us).
//   Registers move between the roles with setmaxnreg (XtRegs).
//   TMEM: S_a [0,128) | S_b [128,256) | O_a [256,384) | O_b [384,512); with <= 80 keys and d <= 80 each group has TWO S
//   buffers of n_pad columns (unit parity) in front of two O tiles of KPAD columns.
//
#include "tc_ptx.cuh"

namespace pd {

constexpr int XT_BQ = 128, XT_THREADS = 640, XT_TILE_BYTES = 128 * 128, XT_MAX_SLOTS = 4;
// setmaxnreg budget: the CTA owns 640 x 96 registers (what the kernel is compiled for), not the whole file:
// 8 x 32 x SOFTMAX + 8 x 32 x EPI + 4 x 32 x CTRL <= 61440, i.e. SOFTMAX + EPI + CTRL / 2 <= 240 (an `inc` that can never be
// satisfied blocks its warps forever).  NKC = 16-key chunks of S a softmax thread holds (5: up to 80 keys, 8: up to 128).
constexpr int XT_REGS_CTRL = 64;
template <int NKC> struct XtRegs { static constexpr int softmax = NKC <= 5 ? 128 : 144, epi = NKC <= 5 ? 72 : 64; };

struct XtArgs {
  int Nq, Nk, heads;
  int npairs;          // 256-query pairs per (batch, head)
  int units;           // B * heads * npairs
  int nslots;          // Q ring depth (4 when a head is one 64-channel chunk, else 2)
  int dbl;             // 1: two S buffers per group (TMEM has room: 4 * n_pad + 2 * KPAD <= 512): Q K^T of unit i + 1 is
                       // issued while the softmax of unit i runs, the MMA round trip leaves the softmax warps' chain
  int s_stride, o_col, o_stride;   // TMEM columns: S buffers at (g * 2 + parity) * s_stride (g * 128 without dbl), O_g at o_col + g * o_stride
  int n_pad;           // keys padded to 80 or 128 (NKC * 16): N extent of Q K^T, K extent of P V
  float scale_log2;
  uint32_t idesc_s, idesc_pv;
  unsigned long long* dbg;   // phase-timeline buffer [10][32] (block 0, first 32 units) or nullptr (pd_debug_attention_timeline)
};

extern unsigned long long* g_fa_dbg_host;
#define XT_DBG(slot, i)                                                                   \
  do {                                                                                    \
    if (a.dbg != nullptr && blockIdx.x == 0 && (i) < 32) a.dbg[(slot) * 32 + (i)] = gtimer(); \
  } while (0)

// KP16 = ceil(d / 16) (compile time: fully unrolled MMA issue and TMEM traffic); NKC = 16-key chunks held per softmax thread
template <int KP16, int NKC>
__global__ void __launch_bounds__(XT_THREADS, 1)
attention_xtc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                     const XtArgs a) {
  constexpr int ND = (KP16 + 3) / 4;          // 64-channel chunks per head
  constexpr int KPAD = KP16 * 16;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* k_s = smem;                                   // [chunk]
  unsigned char* v_s = k_s + ND * XT_TILE_BYTES;               // [chunk]
  unsigned char* q_s = v_s + ND * XT_TILE_BYTES;               // [slot][group][chunk]
  uint64_t* bars = reinterpret_cast<uint64_t*>(q_s + a.nslots * 2 * ND * XT_TILE_BYTES);
  uint64_t& kv_full = bars[0];
  uint64_t& kv_empty = bars[1];
  uint64_t* q_full = bars + 2;        // [4]
  uint64_t* q_empty = bars + 6;       // [4]  both groups' stores have read the slot
  uint64_t* s_full = bars + 10;       // [2 groups][2 buffers]  MMA -> softmax: S_g(i) complete
  uint64_t* p_full = bars + 14;       // [2]  softmax -> MMA: P_g(i) in TMEM
  uint64_t* o_full = bars + 16;       // [2]  MMA -> epilogue: O_g(i) complete
  uint64_t* o_free = bars + 18;       // [2]  epilogue -> MMA: O_g(i) is in registers
  uint64_t* l_full = bars + 20;       // [2]  softmax -> epilogue: 1 / l of unit i is in shared memory
  uint32_t& tmem_base_slot = *reinterpret_cast<uint32_t*>(bars + 22);
  float* inv_s = reinterpret_cast<float*>(bars + 24);     // [group][unit parity][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_TMA = 16, W_MMA = 17;
  const int u0 = (int)((long long)a.units * blockIdx.x / gridDim.x);
  const int u1 = (int)((long long)a.units * (blockIdx.x + 1) / gridDim.x);

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); tma_prefetch_desc(&map_o);
    mbar_init(&kv_full, 1); mbar_init(&kv_empty, 1);
    for (int i = 0; i < XT_MAX_SLOTS; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 2); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[2 * g], 1); mbar_init(&s_full[2 * g + 1], 1);
      mbar_init(&p_full[g], 4); mbar_init(&o_full[g], 1); mbar_init(&o_free[g], 4); mbar_init(&l_full[g], 4);
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

  if (warp >= 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(XT_REGS_CTRL));
  if (warp == W_TMA) {
    if (lane == 0) {
      int cur_bh = -1, nkv = 0;
      for (int u = u0; u < u1; ++u) {
        const int i = u - u0, bh = u / a.npairs, pair = u - bh * a.npairs;
        const int b = bh / a.heads, h = bh - b * a.heads;
        const int slot = i % a.nslots;
        mbar_wait(&q_empty[slot], ((uint32_t)(i / a.nslots) & 1u) ^ 1u, 110 + slot);
#ifdef XT_HALFQ               // (timing experiment, wrong results: only group 0's Q tile is fetched)
        mbar_expect_tx(&q_full[slot], ND * XT_TILE_BYTES);
        for (int c = 0; c < ND; ++c)
          tma_load_4d(q_s + ((slot * 2) * ND + c) * XT_TILE_BYTES, &map_q, &q_full[slot], c * 64, h, (pair * 2) * XT_BQ, b);
#else
        mbar_expect_tx(&q_full[slot], 2 * ND * XT_TILE_BYTES);
        for (int g = 0; g < 2; ++g)
          for (int c = 0; c < ND; ++c)
            tma_load_4d(q_s + ((slot * 2 + g) * ND + c) * XT_TILE_BYTES, &map_q, &q_full[slot], c * 64, h,
                        (pair * 2 + g) * XT_BQ, b);
#endif
        if (bh != cur_bh) {                  // (after the Q load: it does not depend on the K / V buffer)
          if (nkv > 0) mbar_wait(&kv_empty, (uint32_t)(nkv - 1) & 1u, 100);
          mbar_expect_tx(&kv_full, 2 * ND * XT_TILE_BYTES);
          for (int c = 0; c < ND; ++c) {
            tma_load_4d(k_s + c * XT_TILE_BYTES, &map_k, &kv_full, c * 64, h, 0, b);
            tma_load_4d(v_s + c * XT_TILE_BYTES, &map_v, &kv_full, c * 64, h, 0, b);
          }
          cur_bh = bh; ++nkv;
        }
      }
    }
  } else if (warp == W_MMA) {
    // warp-uniform loop, one ELECTED lane issues (inside `if (lane == 0)` ptxas wraps every UTCHMMA / UTCBAR in an
    // ELECT ... BRA.U.ANY retry loop, ~150 clk per MMA: with 16 MMAs per unit that alone was 1.3 us per unit here)
    if (u0 < u1) {
      const uint64_t qdesc0 = make_smem_desc(s_u32(q_s));
      const uint64_t kdesc0 = make_smem_desc(s_u32(k_s));
      const uint64_t vdesc0 = make_smem_desc_mn(s_u32(v_s), XT_TILE_BYTES, 1024);
      constexpr uint64_t TILE16 = XT_TILE_BYTES >> 4;
      const uint32_t tm_o = tmem_base + (uint32_t)a.o_col;
      const int ksteps_pv = a.n_pad >> 4;
      const int dbl = a.dbl;
      // S buffer of group g for unit index i
      auto s_col = [&](int g, int i) { return tmem_base + (uint32_t)(dbl ? (g * 2 + (i & 1)) * a.s_stride : g * 128); };
      auto issue_qk = [&](int g, int slot, int i) {
        const uint64_t qd = qdesc0 + (uint64_t)((slot * 2 + g) * ND) * TILE16;
        const uint32_t sd = s_col(g, i);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < KP16; ++k) {
            const uint64_t off = (uint64_t)(k >> 2) * TILE16 + (uint64_t)(k & 3) * 2u;
            umma_bf16(sd, qd + off, kdesc0 + off, a.idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit(&s_full[2 * g + (dbl ? (i & 1) : 0)]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g, int i) {
        const uint32_t pa = s_col(g, i);
        if (elect_one()) {
          for (int k = 0; k < ksteps_pv; ++k)
            umma_bf16_ts(tm_o + (uint32_t)(g * a.o_stride), pa + (uint32_t)(8 * k), vdesc0 + (uint64_t)(k * 128), a.idesc_pv,
                         k != 0 ? 1u : 0u);
          umma_commit(&o_full[g]);
        }
        __syncwarp();
      };
      int nkv = 0;
      mbar_wait(&kv_full, 0, 200); ++nkv;
      mbar_wait(&q_full[0], 0, 210);
      tc_fence_after();
      issue_qk(0, 0, 0);
      issue_qk(1, 0, 0);
      for (int u = u0; u < u1; ++u) {
        const int i = u - u0;
        const bool more = u + 1 < u1;
        const bool new_head = more && (u + 1) / a.npairs != u / a.npairs;
        const int nslot = (i + 1) % a.nslots;
        const uint32_t nq_par = (uint32_t)((i + 1) / a.nslots) & 1u;
        bool qk_done = false;
        if (lane == 0) XT_DBG(0, i);
        if (dbl && more && !new_head) {
          // two S buffers: Q K^T of the NEXT unit goes out before this unit's P is even awaited (its buffer held
          // S / P of unit i - 1, whose P V was issued an iteration ago: in-order execution of the tensor pipe).  The Q
          // slot it needs comes back from an epilogue that only waits for P Vs issued earlier: no cycle.
          mbar_wait(&q_full[nslot], nq_par, 210 + nslot);
          tc_fence_after();
          issue_qk(0, nslot, i + 1);
          issue_qk(1, nslot, i + 1);
          qk_done = true;
        }
        // (one S buffer) Q(i + 1) may still be on its way: never let that wait stand in front of a P V the epilogue is
        // waiting for — with two slots it would be a cycle
        const bool q_ready = !qk_done && more && !new_head && mbar_try_wait(&q_full[nslot], nq_par);
        const bool q_rdy = __shfl_sync(0xffffffffu, q_ready ? 1 : 0, 0) != 0;
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&p_full[g], (uint32_t)i & 1u, 220 + g);
          mbar_wait(&o_free[g], ((uint32_t)i & 1u) ^ 1u, 230 + g);     // O_g(i - 1) has left TMEM
          tc_fence_after();
          issue_pv(g, i);
          if (lane == 0) XT_DBG(1 + g, i);
          if (q_rdy) { tc_fence_after(); issue_qk(g, nslot, i + 1); }
        }
        if (!qk_done && more && !new_head && !q_rdy) {
          mbar_wait(&q_full[nslot], nq_par, 210 + nslot);
          tc_fence_after();
          issue_qk(0, nslot, i + 1);
          issue_qk(1, nslot, i + 1);
        }
        if (!more || new_head) {                                         // every MMA that reads this head's K / V is issued
          if (elect_one()) umma_commit(&kv_empty);
          __syncwarp();
        }
        if (new_head) {
          mbar_wait(&kv_full, (uint32_t)nkv & 1u, 200); ++nkv;
          mbar_wait(&q_full[nslot], nq_par, 210 + nslot);
          tc_fence_after();
          issue_qk(0, nslot, i + 1);
          issue_qk(1, nslot, i + 1);
        }
      }
    }
  } else if (warp < 8) {
    // ---------------- softmax: thread == query row ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(XtRegs<NKC>::softmax));
    const int g = warp >> 2;
    const int qd4 = warp & 3;
    const int r = qd4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd4 * 32) << 16;
    const float sc = a.scale_log2;
    for (int u = u0; u < u1; ++u) {
      const int i = u - u0;
      const uint32_t tmem_s = tmem_base + (uint32_t)(a.dbl ? (g * 2 + (i & 1)) * a.s_stride : g * 128) + lane_off;
      if (a.dbl) mbar_wait(&s_full[2 * g + (i & 1)], (uint32_t)(i >> 1) & 1u, 300 + g);
      else mbar_wait(&s_full[2 * g], (uint32_t)i & 1u, 300 + g);
      tc_fence_after();
      if (warp == 0 && lane == 0) XT_DBG(3, i);
      // NKC * 16 keys exactly (the host pads the key extent to 80 or 128: K / V rows past Nk are TMA zero fill): every
      // loop below is fully unrolled and predicate-free except the mask of the chunks that straddle Nk
      uint32_t s[NKC * 16];
#pragma unroll
      for (int c = 0; c < NKC; ++c) tmem_ld16(tmem_s + (uint32_t)(c * 16), *reinterpret_cast<uint32_t(*)[16]>(&s[c * 16]));
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < NKC; ++c) {
        if (a.Nk < (c + 1) * 16) {                   // warp-uniform: only the chunk(s) holding keys >= Nk
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (c * 16 + e >= a.Nk) s[c * 16 + e] = 0xff800000u;     // -inf
        }
      }
      // row maximum: four independent chains of 3-input max
      float m0 = __uint_as_float(s[0]), m1 = __uint_as_float(s[1]), m2 = __uint_as_float(s[2]), m3 = __uint_as_float(s[3]);
#pragma unroll
      for (int e = 4; e + 7 < NKC * 16; e += 8) {
        m0 = fmax3(m0, __uint_as_float(s[e]), __uint_as_float(s[e + 1]));
        m1 = fmax3(m1, __uint_as_float(s[e + 2]), __uint_as_float(s[e + 3]));
        m2 = fmax3(m2, __uint_as_float(s[e + 4]), __uint_as_float(s[e + 5]));
        m3 = fmax3(m3, __uint_as_float(s[e + 6]), __uint_as_float(s[e + 7]));
      }
      // (NKC * 16 - 4 is a multiple of 4 but not of 8: the last four elements)
      m0 = fmax3(m0, __uint_as_float(s[NKC * 16 - 4]), __uint_as_float(s[NKC * 16 - 3]));
      m1 = fmax3(m1, __uint_as_float(s[NKC * 16 - 2]), __uint_as_float(s[NKC * 16 - 1]));
      const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      const float nm = -fmaxf(mx * sc, -1e30f);
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int c = 0; c < NKC; ++c) {
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          float x0, x1, x2, x3;
          ffma2(x0, x1, __uint_as_float(s[c * 16 + e]), __uint_as_float(s[c * 16 + e + 1]), sc, sc, nm, nm);
          ffma2(x2, x3, __uint_as_float(s[c * 16 + e + 2]), __uint_as_float(s[c * 16 + e + 3]), sc, sc, nm, nm);
          x0 = ex2_approx(x0); x1 = ex2_approx(x1); x2 = ex2_approx(x2); x3 = ex2_approx(x3);
          fadd2(l0, l1, l0, l1, x0, x1);
          fadd2(l2, l3, l2, l3, x2, x3);
          pk[e >> 1] = pack_bf16x2(x0, x1);
          pk[(e >> 1) + 1] = pack_bf16x2(x2, x3);
        }
        tmem_st8(tmem_s + (uint32_t)(c * 8), pk);          // P_g over S_g: 8 columns of bf16 pairs per 16 keys
      }
      l0 = (l0 + l1) + (l2 + l3); l1 = 0.f;
      inv_s[(g * 2 + (i & 1)) * 128 + r] = 1.0f / (l0 + l1);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&p_full[g]); mbar_arrive(&l_full[g]); }
      if (warp == 0 && lane == 0) XT_DBG(4, i);
    }
  } else if (warp < 16) {
    // ---------------- epilogue: thread == query row ----------------
    // (the launch gives 96 registers per thread; 72 are enough here and the difference feeds the softmax warps)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(XtRegs<NKC>::epi));
    const int g = (warp - 8) >> 2;
    const int qd4 = warp & 3;
    const int r = qd4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd4 * 32) << 16;
    const uint32_t tmem_o = tmem_base + (uint32_t)(a.o_col + g * a.o_stride) + lane_off;
    const bool store_thread = qd4 == 0 && lane == 0;
    for (int u = u0; u < u1; ++u) {
      const int i = u - u0, bh = u / a.npairs, pair = u - bh * a.npairs;
      const int b = bh / a.heads, h = bh - b * a.heads;
      const int slot = i % a.nslots;
      const int q0 = (pair * 2 + g) * XT_BQ;
      // the previous unit's store has read ITS slot: hand that slot back to the producer
      // (four slots: one store may stay in flight — the slot of unit i - 2 comes back here, never a wait)
      if (store_thread) {
        if (a.nslots == 4) { if (i > 1) { tma_store_wait_read<1>(); mbar_arrive(&q_empty[(i - 2) % 4]); } }
        else if (i > 0) { tma_store_wait_read<0>(); mbar_arrive(&q_empty[(i - 1) % a.nslots]); }
      }
      mbar_wait(&l_full[g], (uint32_t)i & 1u, 320 + g);
      const float inv = inv_s[(g * 2 + (i & 1)) * 128 + r];
      if (warp == 8 && lane == 0) XT_DBG(5, i);
      mbar_wait(&o_full[g], (uint32_t)i & 1u, 310 + g);
      tc_fence_after();
      if (warp == 8 && lane == 0) XT_DBG(6, i);
      unsigned char* stage_o = q_s + ((slot * 2 + g) * ND) * XT_TILE_BYTES;     // Q_g(i) is dead: S_g(i) exists
      const uint32_t stage_row = s_u32(stage_o) + (uint32_t)(r * 128 + ((r & 7) << 4));
      // O leaves TMEM 32 columns per round trip (registers are scarce here by design)
#pragma unroll
      for (int cb = 0; cb < KPAD; cb += 32) {
        const int w = KPAD - cb < 32 ? KPAD - cb : 32;
        uint32_t o[32];
        tmem_ld16(tmem_o + (uint32_t)cb, *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
        if (w > 16) tmem_ld16(tmem_o + (uint32_t)(cb + 16), *reinterpret_cast<uint32_t(*)[16]>(&o[16]));
        tmem_ld_wait();
        if (cb + 32 >= KPAD) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_free[g]);                               // P V of the next unit may overwrite O_g
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {                // 16-byte chunk along the channel axis
          if (kk * 8 < w) {
            const int kc = (cb >> 3) + kk;
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[kk * 8 + e]) * inv;
            sts_bf16x8((stage_row + (uint32_t)((kc >> 3) * XT_TILE_BYTES)) ^ (uint32_t)((kc & 7) << 4), pack8(f));
          }
        }
      }
      fence_proxy_async();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      if (store_thread) {
#ifndef XT_NOSTORE            // (timing experiment: no output written)
        if (q0 < a.Nq)
          for (int c = 0; c < ND; ++c) tma_store_4d(&map_o, stage_o + c * XT_TILE_BYTES, c * 64, h, q0, b);
#endif
        tma_store_commit();
      }
      if (warp == 8 && lane == 0) XT_DBG(7, i);
    }
    if (warp == 8 && lane == 0) griddep_launch();
    if (store_thread) tma_store_wait_all();      // smem must outlive the bulk stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

bool attention_xtc_supported(int dtype, int d, int Nk, int ldq, int ldk, int ldv, int ldo, const void* q, const void* k,
                             const void* v, const void* out) {
  static int sm100 = -1;
  if (sm100 < 0) sm100 = pd_device_is_sm100();
  return sm100 && dtype == PD_BF16 && d % 8 == 0 && d >= 16 && d <= 128 && Nk >= 1 && Nk <= 128 && ldq % 8 == 0 &&
         ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && ((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 &&
         ((uintptr_t)v % 16) == 0 && ((uintptr_t)out % 16) == 0;
}

int attention_xtc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                  int heads, int Nq, int Nk, int d, float scale, cudaStream_t s) {
  XtArgs a;
  a.Nq = Nq; a.Nk = Nk; a.heads = heads;
  a.npairs = (Nq + 2 * XT_BQ - 1) / (2 * XT_BQ);
  const long long units = (long long)B * heads * a.npairs;
  if (units > 0x7fffffffLL) { set_error("attention_xtc: too many query tiles"); return PD_ERR_UNSUPPORTED; }
  a.units = (int)units;
  const int nd = (d + 63) / 64;
  a.nslots = nd == 1 ? 4 : 2;
  a.n_pad = Nk <= 80 ? 80 : 128;            // the softmax threads hold exactly 5 or 8 16-key chunks (template NKC)
  const int kpad = (d + 15) / 16 * 16;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.dbg = g_fa_dbg_host;
  a.dbl = (4 * a.n_pad + 2 * kpad <= 512) ? 1 : 0;
  a.s_stride = a.dbl ? a.n_pad : 128;
  a.o_col = a.dbl ? 4 * a.n_pad : 256;
  a.o_stride = a.dbl ? kpad : 128;
  const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
  a.idesc_s = base | ((uint32_t)(a.n_pad >> 3) << 17);
  a.idesc_pv = base | (1u << 16) | ((uint32_t)(kpad >> 3) << 17);   // B (= V tile) is MN-major

  CUtensorMap mq, mk, mv, mo;
  const uint32_t box[4] = {64, 1, 128, 1};
  const uint32_t es[4] = {1, 1, 1, 1};
  struct { CUtensorMap* m; const void* p; int ld; int n; const char* nm; } t[4] = {
      {&mq, q, ldq, Nq, "xattnQ"}, {&mk, k, ldk, Nk, "xattnK"}, {&mv, v, ldv, Nk, "xattnV"}, {&mo, out, ldo, Nq, "xattnO"}};
  for (int i = 0; i < 4; ++i) {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)t[i].n, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)t[i].ld * 2, (uint64_t)t[i].n * t[i].ld * 2};
    int rc = encode_map(t[i].m, t[i].p, 4, dims, strides, box, es, t[i].nm);
    if (rc) return rc;
  }
  const size_t smem = (size_t)(2 * nd + a.nslots * 2 * nd) * XT_TILE_BYTES + 256 + 2 * 2 * 128 * sizeof(float) + 1024;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError(); sms = 148;
  }
  dim3 grid((unsigned)(a.units < sms ? a.units : sms));
#define XT_LAUNCH(KP)                                                                                              \
  case KP: {                                                                                                       \
    auto kern = a.n_pad <= 80 ? attention_xtc_kernel<KP, 5> : attention_xtc_kernel<KP, 8>;                         \
    static bool attr_set[2] = {false, false};                                                                      \
    if (!attr_set[a.n_pad <= 80 ? 0 : 1]) {                                                                        \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
      if (e != cudaSuccess) { set_error("attention_xtc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; } \
      attr_set[a.n_pad <= 80 ? 0 : 1] = true;                                                                      \
    }                                                                                                              \
    cudaError_t le = launch_pdl(kern, grid, dim3(XT_THREADS), smem, s, 1, mq, mk, mv, mo, a);                      \
    if (le != cudaSuccess) { set_error("attention_xtc: launch failed: %s", cudaGetErrorString(le)); return (int)le; } \
  } break;
  switch (kpad / 16) {
    XT_LAUNCH(1) XT_LAUNCH(2) XT_LAUNCH(3) XT_LAUNCH(4) XT_LAUNCH(5) XT_LAUNCH(6) XT_LAUNCH(7) XT_LAUNCH(8)
    default: set_error("attention_xtc: unsupported head dim %d", d); return PD_ERR_UNSUPPORTED;
  }
#undef XT_LAUNCH
  return check_launch("attention_xtc");
}

}  // namespace pd
