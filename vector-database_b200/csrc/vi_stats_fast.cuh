// vi_stats_fast.cuh -- fast-mode statistics + split choice (replaces IndexBuilder.cs:55-88 for every open range).
//
// qfx (fixed point): every component is quantised once, xi = rint(x * 2^(QBITS-E)) with max|x| < 2^E and
// QBITS = 26, and the per-(range, dim) sums S1 = sum xi and S2 = sum xi^2 are EXACT integers, so any reduction
// order -- lanes, teams, CTAs, atomics, GPUs -- gives the same bits.  Split choice: K = n*S2 - S1^2 (exact,
// 128 bits), arg-max on even depths and arg-min on odd depths, lowest index on ties;
// Mid = float((double)S1 / n * 2^(E-QBITS)).  A range whose CHOSEN dimension has K < n^2 * 2^10 (stdev below 2^5
// quantisation steps: Mid's rounding error would exceed ~1.5 % of the spread) takes the reference's float32 statistics instead (welford_team).  The CPU statement of the
// same rules is oracle/vi_oracle.c mode 1.
//
// |xi| <= 2^26, xi^2 <= 2^52: a lane that accumulates fewer than 4096 rows keeps S1 and S2 in one 64-bit register
// pair each, so an element costs FMUL + F2I + 2 x IMAD.WIDE.  Wider totals only appear where partial sums meet
// (shared/global integer atomics on 32-bit limbs, k_finalize_big_fast).
//
// Work decomposition by range size n (the host launches only the classes present on a level):
//   n <  t_team           one TEAM of TS lanes per range   (k_stats_small_fast<TS,CH,FULL,false>)
//   t_team <= n < t_big   one WARP per range, its 32/TS teams stride the rows (k_stats_small_fast<..,true>)
//   n >= t_big            VI_CHUNK-row chunks, one CTA each, integer atomics into gacc (k_stats_big_fast) +
//                         one warp per range for the arg-max (k_finalize_big_fast);   t_big <= 4096
// A lane owns CH float4 column chunks of a row: 16-byte loads, a team reads one whole row per step.
// FULL: the row is exactly TS*CH float4 wide (no column guards, one pass): D = 96 -> <8,3>, D = 768 -> <32,6>.
#pragma once
#include "vi_stats_common.cuh"

constexpr int VI_QBITS = 26;
constexpr u32 VI_MAX_ROWS_PER_LANE = 1u << (64 - 2 * VI_QBITS);  // 4096

__device__ __forceinline__ void qfx_acc(i64& s1, u64& s2, float x, float k)
{
  const int xi = __float2int_rn(__fmul_rn(x, k));
  // mad.wide.s32: 32x32 -> 64-bit product added to a 64-bit accumulator in one instruction
  asm("mad.wide.s32 %0, %1, 1, %0;" : "+l"(s1) : "r"(xi));
  asm("mad.wide.s32 %0, %1, %1, %0;" : "+l"(s2) : "r"(xi));
}

__device__ __forceinline__ void qfx_acc4(i64* s1, u64* s2, const float4& x, float k)
{
  qfx_acc(s1[0], s2[0], x.x, k);
  qfx_acc(s1[1], s2[1], x.y, k);
  qfx_acc(s1[2], s2[2], x.z, k);
  qfx_acc(s1[3], s2[3], x.w, k);
}

struct Key128
{
  u64 hi, lo;
};

__device__ __forceinline__ bool key_lt(const Key128& a, const Key128& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
__device__ __forceinline__ bool key_eq(const Key128& a, const Key128& b) { return a.hi == b.hi && a.lo == b.lo; }

// K = n*S2 - S1^2 >= 0 (Cauchy-Schwarz), exact: n < 2^32, S2 = s2hi:s2lo < 2^96, |S1| < 2^63
__device__ __forceinline__ Key128 qfx_key(u32 n, i64 s1, u64 s2lo, u64 s2hi)
{
  const u64 plo = (u64)n * s2lo;
  const u64 phi = __umul64hi((u64)n, s2lo) + (u64)n * s2hi;
  const u64 a = s1 < 0 ? (u64)(-s1) : (u64)s1;
  const u64 qlo = a * a, qhi = __umul64hi(a, a);
  Key128 k;
  k.lo = plo - qlo;
  k.hi = phi - qhi - (plo < qlo ? 1ull : 0ull);
  return k;
}

// n^2 * 2^(2*VI_QFX_MIN_RES_BITS)
constexpr int VI_QFX_MIN_RES_BITS = 5;  // chosen dimension must spread over >= 2^5 quantisation steps (stdev)
__device__ __forceinline__ Key128 qfx_threshold(u32 n)
{
  const u64 n2 = (u64)n * (u64)n;
  Key128 t;
  t.lo = n2 << (2 * VI_QFX_MIN_RES_BITS);
  t.hi = n2 >> (64 - 2 * VI_QFX_MIN_RES_BITS);
  return t;
}

struct QfxBest
{
  Key128 key;
  i64 s1;
  int idx;  // INT_MAX = none
};

__device__ __forceinline__ bool qfx_better(bool mx, const Key128& k, int i, const Key128& bk, int bi)
{
  if (i == 0x7fffffff) return false;
  if (bi == 0x7fffffff) return true;
  if (!key_eq(k, bk)) return mx ? key_lt(bk, k) : key_lt(k, bk);
  return i < bi;  // lowest index wins ties (MaxBy keeps the first maximum, IndexBuilder.cs:77-79)
}

template <int W>
__device__ __forceinline__ QfxBest qfx_reduce(QfxBest b, bool mx, u32 mask)
{
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1)
  {
    QfxBest t;
    t.key.hi = __shfl_xor_sync(mask, b.key.hi, o);
    t.key.lo = __shfl_xor_sync(mask, b.key.lo, o);
    t.s1 = __shfl_xor_sync(mask, b.s1, o);
    t.idx = __shfl_xor_sync(mask, b.idx, o);
    if (qfx_better(mx, t.key, t.idx, b.key, b.idx)) b = t;
  }
  return b;
}

__device__ __forceinline__ float qfx_mid(i64 s1, u32 n, double qinv)
{
  return __double2float_rn(__dmul_rn(__ddiv_rn(__ll2double_rn(s1), (double)n), qinv));
}

// ---- bulk-async row ring of the warp-per-range kernel (RING) -----------------------------------------------------------
// The warp class (33 .. 511 points per range, levels 13-18 of 10M x 96) is latency-bound: a lane that fetches its
// share of a row into registers keeps 2 rows in flight (80 registers, profiles/r1_ncu_final.md).  With RING every row
// is ONE 1-D bulk copy (cp.async.bulk, the TMA engine without a tensor map: 384 contiguous bytes at D = 96) into a
// 32-slot shared-memory ring owned by the warp, completion on one mbarrier per slot (expect_tx = row bytes): the lane
// that holds a row's index issues its copy, 32 rows stay in flight per warp without holding a register, and the teams
// read their rows with LDS.128.  A slot is refilled with the row 32 positions further on as soon as the four teams have
// read the four rows of a step (__syncwarp), so the ring never drains inside a range.
constexpr int RING_SLOTS = 32;

__device__ __forceinline__ void mbar_init(u32 bar, u32 count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u32 bar, u32 parity)
{
  u32 ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity)
{
  u32 spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 24)) __trap();  // a copy that never lands is a bug: fail the launch instead of hanging the GPU
}

__host__ __device__ inline size_t ring_smem_bytes_per_warp(int ld) { return (size_t)RING_SLOTS * (size_t)ld * 4 + RING_SLOTS * 8; }

template <int TS, int CH, bool FULL, bool WPS, bool RING = false>
__global__ void __launch_bounds__(256, (CH == 1) ? 4 : ((TS * CH <= 32) ? (RING ? 2 : 3) : 1))
k_stats_small_fast(const LevelDev* __restrict__ lvp, SegLevel sg, u32 nmin, u32 nmax, const u32* __restrict__ perm,
                   const i64* __restrict__ pid, const float* __restrict__ rows, int ld, int dims, float qk, double qinv, int mx,
                   StatsOut out, u64* __restrict__ gacc, const u32* __restrict__ bl_parent, const u32* __restrict__ bl_sib,
                   u32 keep_thr)
{
  const u32 R = lvp->R;
  constexpr int NTW = 32 / TS;            // teams per warp
  constexpr int TPS = WPS ? NTW : 1;      // teams sharing one range
  constexpr int GS = WPS ? 32 : TS;       // lanes sharing one range
  const int lane = threadIdx.x & 31;
  const int tl = lane % TS;               // lane inside its team
  const int trow = WPS ? lane / TS : 0;   // first row (of every TPS) this team takes
  const int gl = WPS ? lane : tl;         // lane inside the group that shares the range
  const u32 gthread = blockIdx.x * 256u + threadIdx.x;
  const u32 s = gthread / GS;
  const u32 tmask = (TS == 32) ? 0xffffffffu : (((1u << TS) - 1u) << (lane - tl));
  const u32 gmask = WPS ? 0xffffffffu : tmask;
  u32 n = 0, S = 0;
  if (s < R)
  {
    n = sg.count[s];
    S = sg.start[s];
    if (n < nmin || n >= nmax) n = 0;
    if (WPS && n != 0 && gacc != nullptr)
    {
      // a range with a big-list slot whose sums are derived (parent - sibling) is k_finalize_big_fast's
      const u32 slot = sg.bslot[s];
      if (slot != 0xffffffffu && bl_parent[slot] != 0xffffffffu) n = 0;
    }
  }
  if (n == 0) return;  // n is uniform over the group
  const int C4 = FULL ? TS * CH : (ld >> 2);
  const u32* pp = perm + S;
  const i64 id0 = (gl < n) ? pid[S + gl] : 0;  // issued early: its latency hides behind the row loads
  const Key128 thr = qfx_threshold(n);
  QfxBest best;
  best.key.hi = 0;
  best.key.lo = 0;
  best.s1 = 0;
  best.idx = 0x7fffffff;

  for (int c0 = 0; c0 < C4; c0 += TS * CH)
  {
    i64 s1[CH * 4];
    u64 s2[CH * 4];
#pragma unroll
    for (int i = 0; i < CH * 4; ++i) { s1[i] = 0; s2[i] = 0; }

    u32 mine = (gl < n) ? pp[gl] : 0u;
    if constexpr (RING)
    {
      static_assert(!RING || (WPS && FULL && TS == 8), "the ring serves the warp-per-range kernel on rows of 8 * CH float4");
      extern __shared__ __align__(16) unsigned char ring_smem[];
      const u32 rowbytes = (u32)ld * 4u;
      unsigned char* wbase = ring_smem + (size_t)(threadIdx.x >> 5) * ring_smem_bytes_per_warp(ld);
      const u32 slot0 = (u32)__cvta_generic_to_shared(wbase);
      const u32 bar0 = slot0 + RING_SLOTS * rowbytes;
      mbar_init(bar0 + (u32)lane * 8u, 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      u32 nxt = (32u + (u32)lane < n) ? pp[32 + lane] : 0u;  // row indexes one and two blocks ahead
      if ((u32)lane < n)
      {
        mbar_expect_tx(bar0 + (u32)lane * 8u, rowbytes);
        bulk_g2s(slot0 + (u32)lane * rowbytes, rows + (size_t)mine * ld, rowbytes, bar0 + (u32)lane * 8u);
      }
      for (u32 jb = 0; jb < n; jb += 32)
      {
        const u32 nxt2 = (jb + 64u + (u32)lane < n) ? pp[jb + 64 + lane] : 0u;
        const u32 m = min(32u, n - jb);
        const u32 par = (jb >> 5) & 1u;
        for (u32 j0 = 0; j0 < m; j0 += 4)
        {
          const u32 jj = j0 + (u32)trow;
          if (jj < m)
          {
            mbar_wait(bar0 + jj * 8u, par);
            const float4* sp = reinterpret_cast<const float4*>(wbase + (size_t)jj * rowbytes);
            float4 x[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) x[k] = sp[k * TS + tl];
#pragma unroll
            for (int k = 0; k < CH; ++k) qfx_acc4(s1 + k * 4, s2 + k * 4, x[k], qk);
          }
          __syncwarp();  // the four rows of this step have been read (and used) by their teams: their slots are free
          if ((u32)lane >= j0 && (u32)lane < j0 + 4u && jb + 32u + (u32)lane < n)
          {
            mbar_expect_tx(bar0 + (u32)lane * 8u, rowbytes);
            bulk_g2s(slot0 + (u32)lane * rowbytes, rows + (size_t)nxt * ld, rowbytes, bar0 + (u32)lane * 8u);
          }
        }
        nxt = nxt2;
      }
    }
    else
    for (u32 jb = 0; jb < n; jb += GS)
    {
      const u32 nxt = (jb + GS + gl < n) ? pp[jb + GS + gl] : 0u;  // prefetch the next block of row indexes
      const u32 m = min((u32)GS, n - jb);
#pragma unroll (CH == 1 ? 8 : 2)
      for (u32 j0 = 0; j0 < m; j0 += TPS)
      {
        const u32 jj = j0 + trow;
        const bool valid = jj < m;
        const u32 r = __shfl_sync(gmask, mine, valid ? jj : 0u, GS);
        if (valid)
        {
          const float4* rp = reinterpret_cast<const float4*>(rows + (size_t)r * ld);
          float4 x[CH];
#pragma unroll
          for (int k = 0; k < CH; ++k)
          {
            const int c = c0 + k * TS + tl;
            x[k] = (FULL || c < C4) ? ldg_f4_stream(rp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int k = 0; k < CH; ++k) qfx_acc4(s1 + k * 4, s2 + k * 4, x[k], qk);
        }
      }
      mine = nxt;
    }
    if (TPS > 1)
    {
      // combine the teams of the warp (integer sums: order does not matter; n < 4096 so S2 < 2^64)
#pragma unroll
      for (int o = TS; o < 32; o <<= 1)
#pragma unroll
        for (int i = 0; i < CH * 4; ++i)
        {
          s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
          s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], o);
        }
    }
    // A summed range with a big-list slot keeps its sums if its own children may form a pair or its sibling is derived
    // from them (looked up here, after the row loop, so that nothing of it is live across the loop):
    // [dim][S1, S2 low 32 bits, S2 >> 32], the layout k_finalize_big_fast derives from
    u64* g = nullptr;
    if (WPS && gacc != nullptr && trow == 0)
    {
      const u32 slot = sg.bslot[s];
      if (slot != 0xffffffffu && (n >= keep_thr || bl_sib[slot] != 0xffffffffu))
        g = gacc + (size_t)slot * ((size_t)ld * 3 + 3);
    }
    if (WPS && g != nullptr)
    {
#pragma unroll
      for (int k = 0; k < CH; ++k)
      {
        const int c = c0 + k * TS + tl;
        if (FULL || c < C4)
        {
#pragma unroll
          for (int e = 0; e < 4; ++e)
          {
            u64* gd = g + (size_t)(c * 4 + e) * 3;
            gd[0] = (u64)s1[k * 4 + e];
            gd[1] = s2[k * 4 + e] & 0xffffffffull;
            gd[2] = s2[k * 4 + e] >> 32;
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < CH; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e)
      {
        const int d = (c0 + k * TS + tl) * 4 + e;
        if (d < dims)
        {
          const Key128 key = qfx_key(n, s1[k * 4 + e], s2[k * 4 + e], 0ull);
          if (qfx_better((mx & VI_MX_MAX) != 0, key, d, best.key, best.idx))
          {
            best.key = key;
            best.s1 = s1[k * 4 + e];
            best.idx = d;
          }
        }
      }
  }
  best = qfx_reduce<TS>(best, (mx & VI_MX_MAX) != 0, tmask);
  int dim = best.idx;
  float mid = qfx_mid(best.s1, n, qinv);
  bool null_dim = false;
  if (key_lt(best.key, thr))  // the chosen dimension is poorly resolved (uniform over the team after the reduction)
  {
    const ExBest eb = welford_team<TS, CH>(rows, ld, dims, pp, n, tl, tmask, (mx & VI_MX_MAX) != 0);
    dim = eb.idx;
    mid = eb.mean;
    null_dim = (mx & VI_MX_SQL) != 0 && eb.key == 0.f;  // Stdev = 0 (DDL.sql:193-194)
  }
  // id sum (Stats.IdN): 32-bit halves in two 64-bit accumulators hold the Int128 sum exactly
  u64 slo = (u32)id0;
  i64 shi = id0 >> 32;
  for (u32 j = gl + GS; j < n; j += GS)
  {
    const i64 id = pid[S + j];
    slo += (u32)id;
    shi += (id >> 32);
  }
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1)
  {
    slo += __shfl_xor_sync(gmask, slo, o);
    shi += __shfl_xor_sync(gmask, shi, o);
  }
  if (gl == 0)
  {
    write_split(sg, out, s, dim, mid, mean_id(slo, shi, n), null_dim, (mx & VI_MX_ROOT_HIGH) != 0);
    u64* g = nullptr;
    if (WPS && gacc != nullptr)
    {
      const u32 slot = sg.bslot[s];
      if (slot != 0xffffffffu && (n >= keep_thr || bl_sib[slot] != 0xffffffffu))
        g = gacc + (size_t)slot * ((size_t)ld * 3 + 3);
    }
    if (WPS && g != nullptr)
    {
      g[(size_t)ld * 3 + 0] = slo;
      g[(size_t)ld * 3 + 1] = (u64)shi;
      g[(size_t)ld * 3 + 2] = (u64)n;
    }
  }
}

// Arg-max over the combined sums of one big range, by one warp.  acc: [dim][S1, S2 limb0, S2 limb1] (shared or
// global), ids: the two id-sum words.
__device__ __forceinline__ void finalize_big_range(const SegLevel& sg, u32 s, u32 n, const u64* acc, const u64* ids,
                                                   int ld, int dims, double qinv, int mx, const StatsOut& out,
                                                   const float* __restrict__ rows, const u32* __restrict__ perm, int lane,
                                                   u32* __restrict__ no_fallback_err = nullptr)
{
  const Key128 thr = qfx_threshold(n);
  QfxBest best;
  best.key.hi = 0;
  best.key.lo = 0;
  best.s1 = 0;
  best.idx = 0x7fffffff;
  for (int d = lane; d < dims; d += 32)
  {
    const i64 s1 = (i64)acc[d * 3 + 0];
    // S2 = limb0 + limb1 * 2^32 (each limb is a sum of 32-bit pieces, < 2^64)
    const u64 l0 = acc[d * 3 + 1], l1 = acc[d * 3 + 2];
    const u64 s2lo = l0 + (l1 << 32);
    const u64 s2hi = (l1 >> 32) + ((s2lo < l0) ? 1ull : 0ull);
    const Key128 key = qfx_key(n, s1, s2lo, s2hi);
    if (qfx_better((mx & VI_MX_MAX) != 0, key, d, best.key, best.idx))
    {
      best.key = key;
      best.s1 = s1;
      best.idx = d;
    }
  }
  best = qfx_reduce<32>(best, (mx & VI_MX_MAX) != 0, 0xffffffffu);
  int dim = best.idx;
  float mid = qfx_mid(best.s1, n, qinv);
  bool null_dim = false;
  if (key_lt(best.key, thr))  // the chosen dimension is poorly resolved
  {
    if (no_fallback_err)
    {
      if (lane == 0) *no_fallback_err = 1u;
      return;
    }
    // many points that the quantisation cannot tell apart: reference arithmetic, one warp (rare, slow)
    const ExBest eb = welford_team<32, 1>(rows, ld, dims, perm + sg.start[s], n, lane, 0xffffffffu, (mx & VI_MX_MAX) != 0);
    dim = eb.idx;
    mid = eb.mean;
    null_dim = (mx & VI_MX_SQL) != 0 && eb.key == 0.f;  // Stdev = 0 (DDL.sql:193-194)
  }
  if (lane == 0)
    write_split(sg, out, s, dim, mid, mean_id(ids[0], (i64)ids[1], n), null_dim, (mx & VI_MX_ROOT_HIGH) != 0);
}

// One CTA owns one chunk (VI_CHUNK rows) of one big range.  A range that fits one chunk (and one column pass) is
// finished by its CTA straight from shared memory; otherwise the partial sums go to gacc with integer atomics
// (order-independent, so the result does not depend on scheduling) and k_finalize_big_fast picks the split.
// gacc per slot: [ld][S1, S2 limb0, S2 limb1] followed by the two id-sum words.
constexpr int BIG_NST = 4;  // cp.async ring depth of the pipelined chunk kernel (rows in flight per team)

// UNR: register-prefetch unroll depth, or 0 for the cp.async ring (dynamic shared memory BIG_NST*CH*256*16 bytes)
template <int TS, int CH, bool FULL, int UNR>
__global__ void __launch_bounds__(256, (CH == 1) ? 4 : ((TS * CH <= 32) ? 3 : 1))
k_stats_big_fast(const LevelDev* __restrict__ lvp, SegLevel sg, const u32* __restrict__ big_list,
                 const u32* __restrict__ chunk_first, const u32* __restrict__ perm, const i64* __restrict__ pid,
                 const float* __restrict__ rows, int ld, int dims, float qk, double qinv, int mx, StatsOut out,
                 u64* __restrict__ gacc, int allow_whole, u32 keep_thr, const u32* __restrict__ bl_sib)
{
  if (blockIdx.x >= lvp->chunks) return;  // the grid is a host-side bound
  const u32 nbig = lvp->nbig;
  constexpr int NT = 256 / TS;          // teams per CTA
  constexpr int PD = TS * CH * 4;       // dims per pass
  static_assert(VI_CHUNK / NT < VI_MAX_ROWS_PER_LANE, "a lane's 64-bit S2 accumulator would overflow");
  __shared__ u64 sacc[PD * 3 + 3];
  __shared__ u32 sperm[VI_CHUNK];
  const u32 bid = blockIdx.x;
  u32 lo = 0, hi = nbig;
  while (hi - lo > 1)
  {
    const u32 m = (lo + hi) >> 1;
    if (chunk_first[m] <= bid) lo = m; else hi = m;
  }
  const u32 slot = lo;
  const u32 s = big_list[slot];
  const u32 S = sg.start[s], n = sg.count[s];
  const u32 a = (bid - chunk_first[slot]) * VI_CHUNK;
  const u32 b = min(n, a + VI_CHUNK);
  const u32 m = b - a;
  const int tl = threadIdx.x % TS, team = threadIdx.x / TS;
  const int C4 = FULL ? TS * CH : (ld >> 2);
  // this CTA sees the whole range in one pass (never in the shared phase of a multi-rank build: the range's
  // other slices live on other ranks and the sums must meet in gacc)
  const bool whole = allow_whole && (n <= VI_CHUNK) && (C4 <= TS * CH);
  const size_t gstride = (size_t)ld * 3 + 3;
  u64* g = gacc + (size_t)slot * gstride;

  for (u32 i = threadIdx.x; i < m; i += 256) sperm[i] = perm[S + a + i];
  for (int i = threadIdx.x; i < PD * 3 + 3; i += 256) sacc[i] = 0;
  // id sums (Stats.IdN)
  {
    u64 slo = 0;
    i64 shi = 0;
    for (u32 j = a + threadIdx.x; j < b; j += 256)
    {
      const i64 id = pid[S + j];
      slo += (u32)id;
      shi += (id >> 32);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      slo += __shfl_xor_sync(0xffffffffu, slo, o);
      shi += __shfl_xor_sync(0xffffffffu, shi, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
    {
      atomicAdd(&sacc[PD * 3 + 0], slo);
      atomicAdd(&sacc[PD * 3 + 1], (u64)shi);
    }
  }

  for (int c0 = 0; c0 < C4; c0 += TS * CH)
  {
    i64 s1[CH * 4];
    u64 s2[CH * 4];
#pragma unroll
    for (int i = 0; i < CH * 4; ++i) { s1[i] = 0; s2[i] = 0; }
    if constexpr (UNR == 0)
    {
      // Rows staged through shared memory with cp.async: a thread copies exactly the CH x 16 bytes it consumes
      // itself (thread-private ring slots: no barrier), so BIG_NST rows per team stay in flight without holding
      // registers -- the register-prefetch variants below are latency-bound at 80 registers (profiles/).
      extern __shared__ float4 s_ring[];  // [BIG_NST][CH][256]
      auto issue = [&](u32 jj, int st)
      {
        if (jj < m)
        {
          const float4* rp = reinterpret_cast<const float4*>(rows + (size_t)sperm[jj] * ld);
#pragma unroll
          for (int k = 0; k < CH; ++k)
          {
            const int c = c0 + k * TS + tl;
            if (FULL || c < C4)
            {
              const u32 dst = (u32)__cvta_generic_to_shared(&s_ring[(st * CH + k) * 256 + threadIdx.x]);
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rp + c) : "memory");
            }
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
#pragma unroll
      for (int st = 0; st < BIG_NST - 1; ++st) issue(team + st * NT, st);
      int st = 0;
      for (u32 j = team; j < m; j += NT)
      {
        issue(j + (BIG_NST - 1) * NT, (st + BIG_NST - 1) % BIG_NST);
        asm volatile("cp.async.wait_group %0;" ::"n"(BIG_NST - 1) : "memory");
#pragma unroll
        for (int k = 0; k < CH; ++k)
        {
          const int c = c0 + k * TS + tl;
          const float4 x = (FULL || c < C4) ? s_ring[(st * CH + k) * 256 + threadIdx.x] : make_float4(0.f, 0.f, 0.f, 0.f);
          qfx_acc4(s1 + k * 4, s2 + k * 4, x, qk);
        }
        st = (st + 1) % BIG_NST;
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    else
    {
#pragma unroll UNR
      for (u32 j = team; j < m; j += NT)
      {
        const u32 r = sperm[j];
        const float4* rp = reinterpret_cast<const float4*>(rows + (size_t)r * ld);
        float4 x[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k)
        {
          const int c = c0 + k * TS + tl;
          x[k] = (FULL || c < C4) ? ldg_f4_stream(rp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < CH; ++k) qfx_acc4(s1 + k * 4, s2 + k * 4, x[k], qk);
      }
    }
#pragma unroll
    for (int k = 0; k < CH; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e)
      {
        const int dl = (k * TS + tl) * 4 + e;  // dim inside this pass
        if (FULL || (c0 + k * TS + tl) < C4)
        {
          atomicAdd(&sacc[dl * 3 + 0], (u64)s1[k * 4 + e]);
          atomicAdd(&sacc[dl * 3 + 1], s2[k * 4 + e] & 0xffffffffull);
          atomicAdd(&sacc[dl * 3 + 2], s2[k * 4 + e] >> 32);
        }
      }
    __syncthreads();
    if (whole)
    {
      // keep the sums: both children may be big (sibling derivation of the next level), or this range's own sibling
      // is derived from them at this level
      if (n >= keep_thr || (bl_sib != nullptr && bl_sib[slot] != 0xffffffffu))
      {
        const int pd = min(PD, C4 * 4);
        for (int i = threadIdx.x; i < pd * 3; i += 256) g[i] = sacc[i];
        if (threadIdx.x < 2) g[(size_t)ld * 3 + threadIdx.x] = sacc[PD * 3 + threadIdx.x];
        if (threadIdx.x == 2) g[(size_t)ld * 3 + 2] = (u64)n;
      }
      if (threadIdx.x < 32)
        finalize_big_range(sg, s, n, sacc, sacc + PD * 3, ld, dims, qinv, mx, out, rows, perm, threadIdx.x);
      return;
    }
    const int pass_dims = min(PD, (C4 - c0) * 4);
    for (int i = threadIdx.x; i < pass_dims * 3; i += 256)
    {
      const u64 v = sacc[i];
      if (v) atomicAdd(&g[(size_t)c0 * 12 + i], v);
      sacc[i] = 0;
    }
    __syncthreads();
  }
  if (threadIdx.x < 2) atomicAdd(&g[(size_t)ld * 3 + threadIdx.x], sacc[PD * 3 + threadIdx.x]);
  if (threadIdx.x == 2) atomicAdd(&g[(size_t)ld * 3 + 2], (u64)m);  // local point count of this chunk
}

// Sibling derivation + split choice of the big ranges, one warp per big-list slot.
// A derived range's sums = the parent's (previous level's gacc) - its sibling's (this level's): all words are exact
// integer sums (S1, the two S2 limbs, the id-sum halves, the count), so the difference is exact; they are stored in
// this level's gacc (the range's own children may be derived from them) and the split is chosen from them.
// Other ranges: arg-max over gacc, unless their single chunk CTA has already finished them.
__global__ void __launch_bounds__(256)
k_finalize_big_fast(const LevelDev* __restrict__ lvp, SegLevel sg, const u32* __restrict__ big_list, u64* __restrict__ gacc,
                    const u64* __restrict__ gacc_prev, int ld, int dims, double qinv, int mx, StatsOut out,
                    const float* __restrict__ rows, const u32* __restrict__ perm, int single_pass, int shared,
                    u32* __restrict__ err, const u32* __restrict__ bl_parent, const u32* __restrict__ bl_sib, u32 t_big)
{
  const u32 warp = (blockIdx.x * 256u + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= lvp->nbig) return;
  const u32 s = big_list[warp];
  const size_t gstride = (size_t)ld * 3 + 3;
  u64* g = gacc + (size_t)warp * gstride;
  const u32 par = bl_parent != nullptr ? bl_parent[warp] : 0xffffffffu;
  const bool derived = par != 0xffffffffu;
  if (derived)
  {
    const u64* gp = gacc_prev + (size_t)par * gstride;
    const u64* gs = gacc + (size_t)bl_sib[warp] * gstride;
    for (int d = lane; d < ld; d += 32)
    {
      g[d * 3 + 0] = gp[d * 3 + 0] - gs[d * 3 + 0];  // S1 (two's complement)
      // S2 = limb0 + limb1 * 2^32; the limbs are sums of 32-bit pieces of the lanes' partial sums, not a canonical
      // split, so the difference is taken on the 128-bit values and stored as (low 32 bits, the rest)
      const unsigned __int128 sp = (unsigned __int128)gp[d * 3 + 1] + ((unsigned __int128)gp[d * 3 + 2] << 32);
      const unsigned __int128 ss = (unsigned __int128)gs[d * 3 + 1] + ((unsigned __int128)gs[d * 3 + 2] << 32);
      const unsigned __int128 sd = sp - ss;
      g[d * 3 + 1] = (u64)(sd & 0xffffffffull);
      g[d * 3 + 2] = (u64)(sd >> 32);
    }
    if (lane < 3) g[(size_t)ld * 3 + lane] = gp[(size_t)ld * 3 + lane] - gs[(size_t)ld * 3 + lane];  // id sums, count
    __syncwarp();
  }
  // shared phase of a multi-rank build: n is the all-reduced (global) count and the float32 fallback, which needs
  // the range's rows in global order, is not available: a poorly resolved range is reported as an error
  const u32 n = shared ? (u32)g[(size_t)ld * 3 + 2] : sg.count[s];
  if (!shared && !derived && n < t_big) return;                     // finished by the warp-per-range kernel
  if (!shared && single_pass && n <= VI_CHUNK && !derived) return;  // finished by its chunk CTA
  finalize_big_range(sg, s, n, g, g + (size_t)ld * 3, ld, dims, qinv, mx, out, rows, perm, lane, shared ? err : nullptr);
}
