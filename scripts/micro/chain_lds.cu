// micro-benchmark: cost of shared-memory instructions inside the 4-deep dependent chain of ONE warp.
// 32-step groups, loads issued LA steps ahead of their use (like k_stats_big_exact_px's chain warp).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int LA = 8;

// MODE 0: no loads            1: v scalar LDS            2: v scalar + r LDS.64 (current kernel)
// MODE 3: v LDS.128/4 steps + r LDS.128/2 steps          4: MODE 3 + STS.128 of the means every 4 steps
// MODE 5: MODE 2 + STS.128 (current kernel incl. store)   6: v LDS.128/4 + r from registers (no table loads)
template <int MODE>
__global__ void k(float* out, const float* in, long long* cyc, int groups)
{
  __shared__ __align__(16) float sv[4][32 * 36];   // [slot][lane*36 + step]  (transposed) or [step*32 + lane]
  __shared__ __align__(16) float2 sr[4][32];
  __shared__ __align__(16) float sa[4][32 * 36];
  for (int i = threadIdx.x; i < 4 * 32 * 36; i += 32) (&sv[0][0])[i] = in[i & 1023];
  for (int i = threadIdx.x; i < 4 * 32; i += 32) (&sr[0][0])[i] = make_float2(in[i] * 1e-3f, in[i + 7] * 1e-9f);
  __syncthreads();
  const int lane = threadIdx.x;
  float x = in[lane];
  const float ra = in[40 + lane] * 1e-3f, rb = in[50 + lane] * 1e-9f;
  const long long t0 = clock64();
  for (int g = 0; g < groups; ++g)
  {
    const float* vr = sv[g & 3];
    const float2* tr = sr[g & 3];
    float4* ar = reinterpret_cast<float4*>(&sa[g & 3][lane * 36]);
    float v[32];
    float2 r[32];
    if (MODE == 1 || MODE == 2 || MODE == 5)
    {
#pragma unroll
      for (int u = 0; u < LA; ++u) { v[u] = vr[u * 32 + lane]; if (MODE != 1) r[u] = tr[u]; }
    }
    if (MODE == 3 || MODE == 4 || MODE == 6)
    {
#pragma unroll
      for (int u = 0; u < LA; u += 4)
      {
        const float4 t = *reinterpret_cast<const float4*>(&vr[lane * 36 + u]);
        v[u] = t.x; v[u + 1] = t.y; v[u + 2] = t.z; v[u + 3] = t.w;
        if (MODE != 6)
        {
          const float4 a = *reinterpret_cast<const float4*>(&tr[u]), b = *reinterpret_cast<const float4*>(&tr[u + 2]);
          r[u] = make_float2(a.x, a.y); r[u + 1] = make_float2(a.z, a.w); r[u + 2] = make_float2(b.x, b.y); r[u + 3] = make_float2(b.z, b.w);
        }
      }
    }
#pragma unroll
    for (int u4 = 0; u4 < 32; u4 += 4)
    {
      if (u4 + LA < 32)
      {
        if (MODE == 1 || MODE == 2 || MODE == 5)
        {
#pragma unroll
          for (int e = 0; e < 4; ++e) { v[u4 + e + LA] = vr[(u4 + e + LA) * 32 + lane]; if (MODE != 1) r[u4 + e + LA] = tr[u4 + e + LA]; }
        }
        if (MODE == 3 || MODE == 4 || MODE == 6)
        {
          const int u = u4 + LA;
          const float4 t = *reinterpret_cast<const float4*>(&vr[lane * 36 + u]);
          v[u] = t.x; v[u + 1] = t.y; v[u + 2] = t.z; v[u + 3] = t.w;
          if (MODE != 6)
          {
            const float4 a = *reinterpret_cast<const float4*>(&tr[u]), b = *reinterpret_cast<const float4*>(&tr[u + 2]);
            r[u] = make_float2(a.x, a.y); r[u + 1] = make_float2(a.z, a.w); r[u + 2] = make_float2(b.x, b.y); r[u + 3] = make_float2(b.z, b.w);
          }
        }
      }
      float4 a4;
      float* ap = reinterpret_cast<float*>(&a4);
#pragma unroll
      for (int e = 0; e < 4; ++e)
      {
        const float vv = (MODE == 0) ? ra : v[u4 + e];
        const float2 rr = (MODE == 0 || MODE == 1 || MODE == 6) ? make_float2(ra, rb) : r[u4 + e];
        const float d = __fsub_rn(vv, x);
        x = __fadd_rn(x, __fmaf_rn(d, rr.x, __fmul_rn(d, rr.y)));
        ap[e] = x;
      }
      if (MODE == 4 || MODE == 5) ar[u4 >> 2] = a4;
    }
  }
  const long long t1 = clock64();
  out[lane] = x;
  if (lane == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name)
{
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1e-3f * (i % 97) + 0.5f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int groups = 20000;
  k<MODE><<<1, 32>>>(out, in, cyc, groups);
  k<MODE><<<1, 32>>>(out, in, cyc, groups);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-62s %6.2f cycles per step\n", name, (double)c / (groups * 32.0));
}

int main()
{
  run<0>("chain only");
  run<1>("+ v: 1 LDS / step");
  run<2>("+ v: 1 LDS, r: 1 LDS.64 / step (current)");
  run<5>("+ v: 1 LDS, r: 1 LDS.64, means: STS.128 / 4 steps (current)");
  run<6>("+ v: LDS.128 / 4 steps, r in registers");
  run<3>("+ v: LDS.128 / 4 steps, r: LDS.128 / 2 steps");
  run<4>("+ v: LDS.128 / 4, r: LDS.128 / 2, means: STS.128 / 4 steps");
  return 0;
}
