// vi_stats_exact_px.cuh -- exact-mode statistics of the TOP levels: the literal float32 recurrence of
// IndexBuilder.cs:159-197 for one (range, 32 dimensions) per CTA, split over four specialised warps.
//
// Why: at the top of the tree there are only ranges x dims/32 chains-of-warps in the whole GPU and each is one serial
// dependency over up to N points, so the build time is (cycles per step) x N.  One warp issues at most one
// instruction every 2 cycles; the single-warp kernel (k_stats_big_exact, ~19 instructions per step) is therefore
// issue-bound at ~37-45 cycles per step although its dependency chain is 16.  Here the warp that owns the chain
// executes only the chain (6-7 instructions per step) and three other warps, on the SM's other schedulers, do the
// rest:
//
//   warp 3  loader    row indexes + 16-byte cp.async row copies into the row ring, (c, RN(1/c), lo(1/c)) tables
//   warp 0  chain     mean_k = mean_{k-1} + q0,  q0 = RN(d*r_hi + RN(d*r_lo))            -> mean ring
//   warp 1  verifier  q1 = RN(q0 + (d - q0*c)*r_hi) == q0 ?  (welford_step_spec's check, vi_stats_exact.cuh)
//   warp 2  variance  q_k = q_{k-1} + (v - mean_{k-1})*(v - mean_k), tiny-operand guard; owns the committed state
//
// Groups of 32 points flow loader -> chain -> verifier -> variance through shared-memory rings; progress counters
// (one writer each) are published with a CTA fence and polled.  A group the verifier or the guard rejects (about one
// in 10^4: the speculative quotient was not RN(d/c)) triggers a restart: the three compute warps meet at a named
// barrier, the variance warp -- whose (mean, q) is the state after the last group it committed -- redoes the groups up
// to the rejected one with welford_step_r (Markstein / IEEE division), and everybody resumes behind it.  The result is
// bit-identical to the sequential recurrence.
#pragma once
#include "vi_stats_exact.cuh"


constexpr int PX_NG = 12;          // row-ring groups: memory latency (~5 groups at 18 cycles/step) + pipeline depth
constexpr int PX_AS = 36;          // mean-ring lane stride in words (16-byte aligned, conflict-free 128-bit accesses)
constexpr int PX_LA = 8;           // chain warp: shared-memory loads issued this many steps ahead
constexpr u32 PX_NONE = 0xffffffffu;

template <int PX_NA>
struct PxShared
{
  float vring[PX_NG][EXU * 32];        // rows, chunk-swizzled like k_stats_big_exact<true>
  float aring[PX_NA][32 * PX_AS];      // means: [lane][step]
  CountRcp tab[PX_NG][EXU];            // per row of a group; .pad = 1.0, or -1.0: the group must take the safe path
  u32 pring[PX_NG][EXU];               // slot g % NG: row indexes of group g + NG
  float first[32];
  float rmean[32];                     // restart: mean after the last redone group
  u32 ld_ready, ch_ready, vf_done, va_done, restart, resume;
};

__device__ __forceinline__ u32 px_ld(const u32* p) { return *reinterpret_cast<const volatile u32*>(p); }
__device__ __forceinline__ void px_publish(u32* p, u32 v, int lane)
{
  __syncwarp();
  if (lane == 0)
  {
    __threadfence_block();
    *reinterpret_cast<volatile u32*>(p) = v;
  }
}
// waits until *ctr >= need; false when a restart was requested instead (warp-uniform)
__device__ __forceinline__ bool px_wait(const u32* ctr, u32 need, const u32* restart, u32& seen)
{
  if (seen >= need) return true;
  for (;;)
  {
    const u32 v = px_ld(ctr);
    if (v >= need)
    {
      seen = v;
      __threadfence_block();
      return true;
    }
    if (restart != nullptr && px_ld(restart) != PX_NONE) return false;
  }
}
__device__ __forceinline__ void px_bar() { asm volatile("bar.sync 1, 96;" ::: "memory"); }

__device__ __forceinline__ int px_ring_off(int lane, int u) { return u * 32 + ((((lane >> 2) ^ (u & 7)) << 2) | (lane & 3)); }

template <int PX_NA, bool FMA_CHAIN>
__global__ void __launch_bounds__(128)
k_stats_big_exact_px(SegLevel sg, const u32* __restrict__ big_list, u32 nblk, const u32* __restrict__ perm,
                     const float* __restrict__ rows, int ld, int dims, float2* __restrict__ gstats)
{
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
    sh.ld_ready = sh.ch_ready = sh.vf_done = sh.va_done = 0;
    sh.restart = PX_NONE;
    sh.resume = 0;
  }
  if (warp == 3)
  {
    const float* src = rows + (size_t)pp[0] * ld + col0;
    if (lane < 8 && col0 + 4 * lane + 4 <= ld)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((u32)__cvta_generic_to_shared(&sh.first[lane * 4])), "l"(src + 4 * lane)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();

  if (warp == 3)
  {
    // ------------------------------------------------ loader -----------------------------------------------------
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
      t.pad = __any_sync(0xffffffffu, count_all_ones(t.c)) ? -1.f : 1.f;  // 1.0 (the chain's multiplier); negative: safe path
      sh.tab[g % PX_NG][lane] = t;
    };
    u32 issued = 0, published = 0, seen_va = 0;
    while (published < ngroups)
    {
      while (issued < ngroups && issued < seen_va + PX_NG) issue(issued++);
      const u32 outstanding = issued - published;
      if (outstanding == 0)
      {
        px_wait(&sh.va_done, issued - PX_NG + 1, nullptr, seen_va);
        continue;
      }
      switch (outstanding)  // wait for the oldest group in flight (cp.async.wait_group takes an immediate)
      {
#define PX_WAIT_CASE(N) case N + 1: asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); break;
        PX_WAIT_CASE(0) PX_WAIT_CASE(1) PX_WAIT_CASE(2) PX_WAIT_CASE(3) PX_WAIT_CASE(4) PX_WAIT_CASE(5)
        PX_WAIT_CASE(6) PX_WAIT_CASE(7) PX_WAIT_CASE(8) PX_WAIT_CASE(9) PX_WAIT_CASE(10)
#undef PX_WAIT_CASE
        default: asm volatile("cp.async.wait_group %0;" ::"n"(PX_NG - 1) : "memory"); break;
      }
      ++published;
      px_publish(&sh.ld_ready, published, lane);
      if (issued < ngroups)
      {
        const u32 v = px_ld(&sh.va_done);  // refresh without blocking
        if (v > seen_va)
        {
          seen_va = v;
          __threadfence_block();
        }
      }
    }
    return;
  }

  // ------------------------------------------- the three compute warps -------------------------------------------
  u32 g = 0;
  float mean = sh.first[lane];  // chain: running mean; verifier: mean before the group; variance: committed mean
  float q = 0.f;                // variance warp only
  u32 seen_ld = 0, seen_ch = 0, seen_vf = 0, seen_va = 0, nrestart = 0;
  for (;;)
  {
    bool restart = false;
    if (warp == 0)
    {
      // ---------------------------------------------- chain ------------------------------------------------------
      while (g < nfull)
      {
        if (px_ld(&sh.restart) != PX_NONE || !px_wait(&sh.ld_ready, g + 1, &sh.restart, seen_ld) ||
            (g >= PX_NA && !px_wait(&sh.va_done, g - PX_NA + 1, &sh.restart, seen_va)))
        {
          restart = true;
          break;
        }
        const float* vr = sh.vring[g % PX_NG];
        const CountRcp* tb = sh.tab[g % PX_NG];
        float4* ar = reinterpret_cast<float4*>(&sh.aring[g % PX_NA][lane * PX_AS]);
        // loads run PX_LA steps ahead of the arithmetic, in program order before the mean-ring stores (the compiler
        // does not move a shared-memory load above a store it cannot disambiguate)
        float v[EXU];
        float4 k[EXU];
#pragma unroll
        for (int u = 0; u < PX_LA; ++u)
        {
          v[u] = vr[px_ring_off(lane, u)];
          k[u] = *reinterpret_cast<const float4*>(&tb[u]);
        }
#pragma unroll
        for (int u4 = 0; u4 < EXU; u4 += 4)
        {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (u4 + e + PX_LA < EXU)
            {
              v[u4 + e + PX_LA] = vr[px_ring_off(lane, u4 + e + PX_LA)];
              k[u4 + e + PX_LA] = *reinterpret_cast<const float4*>(&tb[u4 + e + PX_LA]);
            }
          float4 a4;
          float* ap = reinterpret_cast<float*>(&a4);
#pragma unroll
          for (int e = 0; e < 4; ++e)
          {
            if (FMA_CHAIN)
            {
              // FADD issues on the ALU pipe and a result crossing pipes costs a cycle: x*1 + y on the FMA pipe rounds
              // exactly like x + y (k.w holds a 1.0f the compiler cannot see)
              const float one = k[u4 + e].w;
              const float d = __fmaf_rn(mean, -one, v[u4 + e]);
              mean = __fmaf_rn(__fmaf_rn(d, k[u4 + e].x, __fmul_rn(d, k[u4 + e].y)), one, mean);
            }
            else
            {
              const float d = __fsub_rn(v[u4 + e], mean);
              mean = __fadd_rn(mean, __fmaf_rn(d, k[u4 + e].x /* r_hi */, __fmul_rn(d, k[u4 + e].y /* r_lo */)));
            }
            ap[e] = mean;
          }
          ar[u4 >> 2] = a4;
        }
        ++g;
        px_publish(&sh.ch_ready, g, lane);
      }
    }
    else if (warp == 1)
    {
      // --------------------------------------------- verifier ----------------------------------------------------
      while (g < nfull)
      {
        if (!px_wait(&sh.ch_ready, g + 1, &sh.restart, seen_ch))
        {
          restart = true;
          break;
        }
        const float* vr = sh.vring[g % PX_NG];
        const CountRcp* tb = sh.tab[g % PX_NG];
        const float4* ar = reinterpret_cast<const float4*>(&sh.aring[g % PX_NA][lane * PX_AS]);
        bool neq = tb[0].pad < 0.f;
        float pa = mean;
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
        mean = pa;
        ++g;
        px_publish(&sh.vf_done, g, lane);
      }
    }
    else
    {
      // --------------------------------------------- variance ----------------------------------------------------
      while (g < nfull)
      {
        if (!px_wait(&sh.vf_done, g + 1, &sh.restart, seen_vf))
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
        px_publish(&sh.va_done, g, lane);
      }
    }
    if (!restart)
    {
      // all full groups are through this warp; it is done once the variance warp has committed them all
      if (warp == 2 || px_wait(&sh.va_done, nfull, &sh.restart, seen_va)) break;
    }
    // ---------------------------------------------- restart ------------------------------------------------------
    px_bar();  // the three warps have stopped; the variance warp holds the state after group g - 1
    if (warp == 2)
    {
      u32 f = px_ld(&sh.restart);
      __threadfence_block();
      // data that keeps failing the check (NaN, denormals): stop speculating, this warp finishes the range alone
      if (++nrestart > 16u && nfull > 0) f = nfull - 1;
      for (; g <= f; ++g)
      {
        px_wait(&sh.ld_ready, g + 1, nullptr, seen_ld);
        const float* vr = sh.vring[g % PX_NG];
        const CountRcp* tb = sh.tab[g % PX_NG];
        for (int u = 0; u < EXU; ++u) welford_step_r(mean, q, vr[px_ring_off(lane, u)], tb[u].c, tb[u].r_hi);
        px_publish(&sh.va_done, g + 1, lane);  // (the loader reuses the slot)
      }
      sh.rmean[lane] = mean;
      __syncwarp();
      if (lane == 0)
      {
        sh.resume = g;
        sh.ch_ready = g;
        sh.vf_done = g;
        sh.restart = PX_NONE;
        __threadfence_block();
        *reinterpret_cast<volatile u32*>(&sh.va_done) = g;  // the loader may move on
      }
    }
    px_bar();
    g = px_ld(&sh.resume);
    mean = sh.rmean[lane];
    seen_ch = seen_vf = seen_va = 0;  // ch_ready / vf_done may have moved back
  }

  if (warp == 2)
  {
    if (ngroups > nfull)  // partial tail: fewer than 32 points, safe steps
    {
      u32 dummy = 0;
      px_wait(&sh.ld_ready, ngroups, nullptr, dummy);
      const float* vr = sh.vring[nfull % PX_NG];
      const CountRcp* tb = sh.tab[nfull % PX_NG];
      const u32 m = (n - 1) - nfull * EXU;
      for (u32 u = 0; u < m; ++u) welford_step_r(mean, q, vr[px_ring_off(lane, (int)u)], tb[u].c, tb[u].r_hi);
    }
    if (act) gstats[(size_t)slot * dims + col] = make_float2(mean, q);
  }
}
