// vi_stats_exact_px.cuh -- exact-mode statistics of the TOP levels: the literal float32 recurrence of
// IndexBuilder.cs:159-197 for one (range, 32 dimensions) per CTA, split over six specialised warps.
//
// Why: at the top of the tree there are only ranges x dims/32 chains-of-warps in the whole GPU and each is one serial
// dependency over up to N points, so the build time is (cycles per step) x N.  One warp issues at most one
// instruction every 2 cycles; the single-warp kernel (k_stats_big_exact, ~19 instructions per step) is therefore
// issue-bound at ~37-45 cycles per step although its dependency chain is 16-18.  Here the warp that owns the chain
// executes only the chain (LDS v, LDS.64 r, FADD, FMUL, FFMA, FADD per step + a STS.128 per four steps) and five other
// warps do the rest:
//
//   warp 5,6  loaders   (even / odd groups) row indexes + 16-byte cp.async row copies into the row ring,
//                       (RN(1/c), lo(1/c), c) tables; warps 4 and 7 exit at once, so that the chain warp has its
//                       scheduler (warp % 4) to itself
//   warp 0    chain     mean_k = mean_{k-1} + q0,  q0 = RN(d*r_hi + RN(d*r_lo))            -> mean ring
//   warp 1,2  verifiers (even / odd groups)  q1 = RN(q0 + (d - q0*c)*r_hi) == q0 ?  (welford_step_spec's check)
//   warp 3    variance  q_k = q_{k-1} + (v - mean_{k-1})*(v - mean_k), tiny-operand guard; owns the committed state
//
// Groups of 32 points flow loader -> chain -> verifier -> variance through shared-memory rings; progress counters
// (one writer each) are release stores / acquire loads at CTA scope (the release costs a MEMBAR that waits for the
// warp's stores in flight, so the chain publishes once per PX_K groups) and are polled.  A group a verifier or the
// guard rejects (about one in 10^4: the speculative quotient was not RN(d/c)) triggers a restart: the four compute
// warps meet at a named barrier, the variance warp -- whose (mean, q) is the state after the last group it committed
// -- redoes the groups up to the rejected one with welford_step_r (Markstein / IEEE division), and everybody resumes
// behind it; after 16 restarts (NaN-poisoned or denormal data) it finishes the range alone.  The result is
// bit-identical to the sequential recurrence.  Liveness: every wait polls the restart flag, every warp that can be
// waited on stays until the variance warp has committed the last full group, and a warp never waits for a group that
// needs its own unpublished output.
#pragma once
#include "vi_stats_exact.cuh"

constexpr int PX_NG = 24;          // row-ring groups: memory latency (~5 groups at 18 cycles/step) + pipeline depth
constexpr int PX_AS = 36;          // mean-ring lane stride in words (16-byte aligned, conflict-free 128-bit accesses)
constexpr int PX_LA = 8;           // chain warp: shared-memory loads issued this many steps ahead
constexpr int PX_K = 4;            // groups per publication (chain, variance)
constexpr u32 PX_NONE = 0xffffffffu;
constexpr unsigned PX_BACKOFF_NS = 100;
constexpr int PX_SD = 16;           // chain warp: mean-ring stores issued this many steps after the values exist
constexpr int PX_THREADS = 256;

template <int PX_NA>
struct PxShared
{
  float vring[PX_NG][EXU * 32];        // rows, chunk-swizzled like k_stats_big_exact<true>
  float aring[PX_NA][32 * PX_AS];      // means: [lane][step]
  CountRcp tab[PX_NG][EXU];            // per row of a group; .pad = 1.0, or -1.0: the group must take the safe path
  u32 pring[PX_NG][EXU];               // slot g % NG: row indexes of group g + NG
  float first[32];
  float rmean[32];                     // restart: mean after the last redone group
  u32 ld_ready[2], ch_ready, vf_done[2], va_done, restart, resume;  // [p]: next unfinished group of parity p
};

// progress counters: acquire loads (a plain LDS at CTA scope) and release stores (MEMBAR.ALL.CTA + STS)
__device__ __forceinline__ u32 px_ld(const u32* p)
{
  u32 v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((u32)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void px_st(u32* p, u32 v)
{
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void px_publish(u32* p, u32 v, int lane)
{
  __syncwarp();
  if (lane == 0) px_st(p, v);
}
__device__ unsigned long long g_px_dbg[8][4];  // VI_B200_TRACE: per warp role of CTA 0: waited A, waited B, total cycles

// polls until *ctr >= need (returns the value seen); PX_NONE when a restart was requested instead (warp-uniform)
__device__ __forceinline__ u32 px_poll(const u32* ctr, u32 need, const u32* restart, long long& waited)
{
  u32 v = px_ld(ctr);
  if (v >= need) return v;
  const long long t0 = clock64();
  for (;;)
  {
    if (restart != nullptr && px_ld(restart) != PX_NONE) return PX_NONE;
    v = px_ld(ctr);
    if (v >= need) break;
    __nanosleep(PX_BACKOFF_NS);  // a spinning warp's loads would queue in front of the chain warp's
  }
  waited += clock64() - t0;
  return v;
}
// the same with the last value cached: a producer that publishes several groups at once is polled less often
__device__ __forceinline__ bool px_wait(const u32* ctr, u32 need, const u32* restart, u32& seen, long long& waited)
{
  if (seen >= need) return true;
  const u32 v = px_poll(ctr, need, restart, waited);
  if (v == PX_NONE) return false;
  seen = v;
  return true;
}
__device__ __forceinline__ void px_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ int px_ring_off(int lane, int u) { return u * 32 + ((((lane >> 2) ^ (u & 7)) << 2) | (lane & 3)); }

template <int PX_NA>
__global__ void __launch_bounds__(PX_THREADS)
k_stats_big_exact_px(SegLevel sg, const u32* __restrict__ big_list, u32 nblk, const u32* __restrict__ perm,
                     const float* __restrict__ rows, int ld, int dims, float2* __restrict__ gstats)
{
  static_assert(PX_NA >= 3 * PX_K + 2 && PX_NG >= PX_NA + 4, "ring depths");
  extern __shared__ __align__(16) unsigned char px_smem[];
  PxShared<PX_NA>& sh = *reinterpret_cast<PxShared<PX_NA>*>(px_smem);
  const u32 slot = blockIdx.x / nblk;
  const int col0 = (int)(blockIdx.x % nblk) * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = col0 + lane;
  const u32 s = big_list[slot];
  const u32 S = sg.start[s], n = sg.count[s];
  const bool act = col < dims;
  const u32* pp = perm + S;
  const u32 nfull = (n - 1) / EXU;                  // groups the pipeline runs
  const u32 ngroups = (n - 1 + EXU - 1) / EXU;      // + a partial tail the variance warp finishes alone

  if (threadIdx.x == 0)
  {
    sh.ch_ready = sh.va_done = 0;
    sh.ld_ready[0] = sh.vf_done[0] = 0;
    sh.ld_ready[1] = sh.vf_done[1] = 1;
    sh.restart = PX_NONE;
    sh.resume = 0;
  }
  if (warp == 4 || warp == 7) return;
  long long wA = 0, wB = 0, wP = 0, wC = 0;
  const long long t_begin = clock64();
  auto trace = [&]()
  {
    if (blockIdx.x == 0 && lane == 0)
    {
      g_px_dbg[warp][0] = (unsigned long long)wA;
      g_px_dbg[warp][1] = (unsigned long long)wB;
      g_px_dbg[warp][2] = (unsigned long long)(clock64() - t_begin);
      g_px_dbg[warp][3] = (unsigned long long)(warp == 0 ? wC : wP);
    }
  };
  if (warp == 5)
  {
    const float* src = rows + (size_t)pp[0] * ld + col0;
    if (lane < 8 && col0 + 4 * lane + 4 <= ld)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((u32)__cvta_generic_to_shared(&sh.first[lane * 4])), "l"(src + 4 * lane)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  asm volatile("bar.sync 2, 192;" ::: "memory");

  if (warp >= 5)
  {
    // ---------------------------------- loader of the even / odd groups -----------------------------------------
    const u32 par = (u32)(warp - 5);
    auto issue = [&](u32 g)
    {
      const u32 j0 = 1u + g * EXU;
      u32 pv = 0;
      if (j0 + lane < n) pv = (g < PX_NG) ? pp[j0 + lane] : px_ld(&sh.pring[g % PX_NG][lane]);
      __syncwarp();
      {
        const u32 jn = j0 + PX_NG * EXU + lane;
        if (jn < n)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((u32)__cvta_generic_to_shared(&sh.pring[g % PX_NG][lane])), "l"(pp + jn)
                       : "memory");
      }
      if (j0 + lane < n)
      {
        const float* src = rows + (size_t)pv * ld + col0;
        const u32 dst = (u32)__cvta_generic_to_shared(&sh.vring[g % PX_NG][lane * 32]);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (col0 + 4 * k + 4 <= ld)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (u32)((k ^ (lane & 7)) << 4)), "l"(src + 4 * k)
                         : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      CountRcp t = count_rcp(j0 + lane + 1u);
      t.pad = __any_sync(0xffffffffu, count_all_ones(t.c)) ? -1.f : 1.f;
      sh.tab[g % PX_NG][lane] = t;
    };
    // issued / completed: next group (of this parity) to issue / to see complete
    u32 issued = par, completed = par, seen_va = 0;
    while (completed < ngroups)
    {
      while (issued < ngroups && issued < seen_va + PX_NG)
      {
        issue(issued);
        issued += 2;
      }
      const u32 outstanding = (issued - completed) >> 1;
      if (outstanding == 0)
      {
        px_wait(&sh.va_done, issued - PX_NG + 1, nullptr, seen_va, wA);
        continue;
      }
      const long long tw = clock64();
      switch (outstanding)  // wait for the oldest group in flight (cp.async.wait_group takes an immediate)
      {
#define PX_WAIT_CASE(N) case N + 1: asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); break;
        PX_WAIT_CASE(0) PX_WAIT_CASE(1) PX_WAIT_CASE(2) PX_WAIT_CASE(3) PX_WAIT_CASE(4) PX_WAIT_CASE(5)
        PX_WAIT_CASE(6) PX_WAIT_CASE(7) PX_WAIT_CASE(8) PX_WAIT_CASE(9) PX_WAIT_CASE(10)
#undef PX_WAIT_CASE
        static_assert(PX_NG / 2 - 1 == 11, "one case per outstanding group count");
        default: asm volatile("cp.async.wait_group %0;" ::"n"(PX_NG / 2 - 1) : "memory"); break;
      }
      wB += clock64() - tw;
      completed += 2;
      px_publish(&sh.ld_ready[par], completed, lane);
      if (issued < ngroups)
      {
        const u32 v = px_ld(&sh.va_done);  // refresh without blocking
        if (v > seen_va) seen_va = v;
      }
    }
    trace();
    return;
  }

  // -------------------------------------------- the four compute warps -------------------------------------------
  // g: next group of this warp (verifiers: of this warp's parity)
  u32 g = (warp == 2) ? 1u : 0u;
  float mean = sh.first[lane];  // chain: running mean; variance: committed mean
  float q = 0.f;                // variance warp only
  u32 seen_ch = 0, seen_va = 0, nrestart = 0;  // cached counter values
  for (;;)
  {
    bool restart = false;
    if (warp == 0)
    {
      // ---------------------------------------------- chain ------------------------------------------------------
      // Up to PX_K groups per iteration (one set of polls, one publication).  The shared-memory loads are volatile asm
      // so that they stay where they are written: PX_LA steps ahead of their use, across group boundaries (under the
      // other warps' traffic an LDS takes ~60 cycles; the compiler would sink the loads to ~3 steps = 50 cycles).
      const u32 vbase = (u32)__cvta_generic_to_shared(&sh.vring[0][0]) + (u32)px_ring_off(lane, 0) * 4u;
      const u32 tbase = (u32)__cvta_generic_to_shared(&sh.tab[0][0]);
      u32 voff[8];  // byte offset of this lane's dimension in the rows u = 0..7 (mod 8) of a group, minus row 0's
#pragma unroll
      for (int u = 0; u < 8; ++u) voff[u] = (u32)(px_ring_off(lane, u) - px_ring_off(lane, 0) - u * 32) * 4u;
      while (g < nfull)
      {
        u32 ng = min((u32)PX_K, nfull - g);  // groups of this iteration: all loaded, mean-ring space for all
        if (px_ld(&sh.restart) != PX_NONE) ng = 0;
        if (ng && ng > 1 && (px_ld(&sh.ld_ready[0]) < g + ng || px_ld(&sh.ld_ready[1]) < g + ng)) ng = 1;
        // (the mean ring keeps one group more than the variance warp has committed: a verifier reads the last mean of
        //  the group before its own)
        if (ng == 0 || px_poll(&sh.ld_ready[g & 1], g + 1, &sh.restart, wA) == PX_NONE ||
            (g + ng + 1 > PX_NA && !px_wait(&sh.va_done, g + ng + 1 - PX_NA, &sh.restart, seen_va, wB)))
        {
          restart = true;
          break;
        }
        const long long tc = clock64();
        float v[EXU], am[EXU];
        float2 k[EXU];
        u32 rs = g % PX_NG, rsn = (rs + 1 == PX_NG) ? 0u : rs + 1u;  // ring slots of the current / next group
        auto load = [&](u32 rslot, int u, int slot)  // row u of the group in ring slot rslot -> v[slot], k[slot]
        {
          const u32 va = vbase + rslot * (EXU * 32 * 4) + (u32)u * 128u + voff[u & 7];
          const u32 ta = tbase + rslot * (EXU * 16) + (u32)u * 16u;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[slot]) : "r"(va));
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(k[slot].x), "=f"(k[slot].y) : "r"(ta));
        };
#pragma unroll
        for (int u = 0; u < PX_LA; ++u) load(rs, u, u);
#pragma unroll 1
        for (u32 gi = 0; gi < ng; ++gi)
        {
          float4* ar = reinterpret_cast<float4*>(&sh.aring[(g + gi) % PX_NA][lane * PX_AS]);
#pragma unroll
          for (int u4 = 0; u4 < EXU; u4 += 4)
          {
#pragma unroll
            for (int e = 0; e < 4; ++e)
            {
              const int un = u4 + e + PX_LA;  // (the last group of the iteration prefetches rows nobody uses)
              if (un < EXU) load(rs, un, un);
              else load(rsn, un - EXU, un - EXU);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
            {
              const float d = __fsub_rn(v[u4 + e], mean);
              mean = __fadd_rn(mean, __fmaf_rn(d, k[u4 + e].x /* r_hi */, __fmul_rn(d, k[u4 + e].y /* r_lo */)));
              am[u4 + e] = mean;
            }
            // the means are stored PX_SD steps late: a store holds its source registers until the shared-memory
            // pipe has taken it, and the chain must not wait on that to reuse them
            if (u4 >= PX_SD) ar[(u4 - PX_SD) >> 2] = make_float4(am[u4 - PX_SD], am[u4 - PX_SD + 1], am[u4 - PX_SD + 2], am[u4 - PX_SD + 3]);
          }
#pragma unroll
          for (int u4 = EXU - PX_SD; u4 < EXU; u4 += 4) ar[u4 >> 2] = make_float4(am[u4], am[u4 + 1], am[u4 + 2], am[u4 + 3]);
          rs = rsn;
          rsn = (rs + 1 == PX_NG) ? 0u : rs + 1u;
        }
        g += ng;
        wC += clock64() - tc;
        px_publish(&sh.ch_ready, g, lane);
      }
    }
    else if (warp <= 2)
    {
      // -------------------------------------- verifier of the even / odd groups ----------------------------------
      const int par = warp - 1;
      while (g < nfull)
      {
        if (!px_wait(&sh.ch_ready, g + 1, &sh.restart, seen_ch, wA))
        {
          restart = true;
          break;
        }
        const float* vr = sh.vring[g % PX_NG];
        const CountRcp* tb = sh.tab[g % PX_NG];
        const float4* ar = reinterpret_cast<const float4*>(&sh.aring[g % PX_NA][lane * PX_AS]);
        bool neq = tb[0].pad < 0.f;
        float pa = (g == 0) ? sh.first[lane] : sh.aring[(g - 1) % PX_NA][lane * PX_AS + EXU - 1];
#pragma unroll
        for (int u4 = 0; u4 < EXU; u4 += 4)
        {
          const float4 a4 = ar[u4 >> 2];
          const float* ap = reinterpret_cast<const float*>(&a4);
#pragma unroll
          for (int e = 0; e < 4; ++e)
          {
            const CountRcp k = tb[u4 + e];
            const float d = __fsub_rn(vr[px_ring_off(lane, u4 + e)], pa);
            const float q0 = __fmaf_rn(d, k.r_hi, __fmul_rn(d, k.r_lo));
            const float q1 = __fmaf_rn(__fmaf_rn(-q0, k.c, d), k.r_hi, q0);
            neq |= !(q1 == q0);
            pa = ap[e];
          }
        }
        if (__any_sync(0xffffffffu, act && neq))
        {
          if (lane == 0) atomicMin(&sh.restart, g);
          restart = true;
          break;
        }
        g += 2;
        px_publish(&sh.vf_done[par], g, lane);
      }
    }
    else
    {
      // --------------------------------------------- variance ----------------------------------------------------
      while (g < nfull)
      {
        if (px_poll(&sh.vf_done[g & 1], g + 1, &sh.restart, wA) == PX_NONE)
        {
          restart = true;
          break;
        }
        const float* vr = sh.vring[g % PX_NG];
        const float4* ar = reinterpret_cast<const float4*>(&sh.aring[g % PX_NA][lane * PX_AS]);
        float pa = mean, qq = q;
        u32 umin1 = 0xffffffffu;
#pragma unroll
        for (int u4 = 0; u4 < EXU; u4 += 4)
        {
          const float4 a4 = ar[u4 >> 2];
          const float* ap = reinterpret_cast<const float*>(&a4);
#pragma unroll
          for (int e = 0; e < 4; ++e)
          {
            const float v = vr[px_ring_off(lane, u4 + e)];
            const float d = __fsub_rn(v, pa);
            umin1 = min(umin1, (__float_as_uint(d) << 1) - 1u);
            qq = __fadd_rn(qq, __fmul_rn(d, __fsub_rn(v, ap[e])));
            pa = ap[e];
          }
        }
        // a non-zero |d| below 2^-60: the check's remainder could lose bits (2 * bits(2^-60) - 1)
        if (__any_sync(0xffffffffu, act && umin1 < 2u * 0x21800000u - 1u))
        {
          if (lane == 0) atomicMin(&sh.restart, g);
          restart = true;
          break;
        }
        mean = pa;
        q = qq;
        ++g;
        if (g % PX_K == 0 || g == nfull) px_publish(&sh.va_done, g, lane);
      }
    }
    if (!restart)
    {
      // all full groups are through this warp; it is done once the variance warp has committed them all
      if (warp == 3 || px_wait(&sh.va_done, nfull, &sh.restart, seen_va, wB)) break;
    }
    // ---------------------------------------------- restart ------------------------------------------------------
    px_bar();  // the four warps have stopped; the variance warp holds the state after its group g - 1
    if (warp == 3)
    {
      u32 f = px_ld(&sh.restart);
      // data that keeps failing the check (NaN, denormals): stop speculating, this warp finishes the range alone
      if (++nrestart > 16u && nfull > 0) f = nfull - 1;
      for (; g <= f; ++g)
      {
        px_poll(&sh.ld_ready[g & 1], g + 1, nullptr, wB);
        const float* vr = sh.vring[g % PX_NG];
        const CountRcp* tb = sh.tab[g % PX_NG];
        for (int u = 0; u < EXU; ++u) welford_step_r(mean, q, vr[px_ring_off(lane, u)], tb[u].c, tb[u].r_hi);
        px_publish(&sh.va_done, g + 1, lane);  // (the loader reuses the slot)
      }
      sh.rmean[lane] = mean;
      if (g > 0) sh.aring[(g - 1) % PX_NA][lane * PX_AS + EXU - 1] = mean;  // what the verifier of group g starts from
      __syncwarp();
      if (lane == 0)
      {
        sh.resume = g;
        sh.ch_ready = g;
        sh.vf_done[g & 1] = g;
        sh.vf_done[(g & 1) ^ 1] = g + 1;
        sh.restart = PX_NONE;
        px_st(&sh.va_done, g);  // the loader may move on
      }
    }
    px_bar();
    g = px_ld(&sh.resume);
    if (warp == 1 || warp == 2) g += (g & 1u) != (u32)(warp - 1) ? 1u : 0u;  // first group of this verifier's parity
    mean = sh.rmean[lane];
    seen_ch = seen_va = 0;  // ch_ready may have moved back
  }

  if (warp == 3)
  {
    if (ngroups > nfull)  // partial tail: fewer than 32 points, safe steps
    {
      px_poll(&sh.ld_ready[nfull & 1], nfull + 1, nullptr, wB);
      const float* vr = sh.vring[nfull % PX_NG];
      const CountRcp* tb = sh.tab[nfull % PX_NG];
      const u32 m = (n - 1) - nfull * EXU;
      for (u32 u = 0; u < m; ++u) welford_step_r(mean, q, vr[px_ring_off(lane, (int)u)], tb[u].c, tb[u].r_hi);
    }
    if (act) gstats[(size_t)slot * dims + col] = make_float2(mean, q);
  }
  trace();
}
