// micro-benchmark: dependent FP32 chain latency of ONE warp on an otherwise idle SM (cycles per dependent op)
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, const float* in, long long* cyc, int iters)
{
  __shared__ float sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  if (threadIdx.x >= 32) { // other warps: MODE 3 spins on shared memory, else exit
    if (MODE == 3) { volatile float* p = sm; float acc = 0; for (int i = 0; i < iters * 8; ++i) acc += p[(i + threadIdx.x) & 1023]; if (acc == 123.f) out[1] = acc; }
    return;
  }
  float a = in[threadIdx.x], b = in[32 + threadIdx.x], c = in[64 + threadIdx.x], x = in[96 + threadIdx.x];
  float y = 0.f;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i)
  {
#pragma unroll
    for (int u = 0; u < 16; ++u)
    {
      if (MODE == 0) { x = __fmaf_rn(x, a, b); }                                  // 1 dependent FFMA
      if (MODE == 1) { x = __fadd_rn(x, a); }                                      // 1 dependent FADD
      if (MODE == 2 || MODE == 3) {                                                // the chain step: FADD, FMUL, FFMA, FADD + 2 LDS
        const float v = sm[(u * 32 + threadIdx.x) & 1023];
        const float2 r = *reinterpret_cast<const float2*>(&sm[(u * 2 + i) & 1022]);
        const float d = __fsub_rn(v, x);
        x = __fadd_rn(x, __fmaf_rn(d, r.x, __fmul_rn(d, r.y)));
      }
      if (MODE == 4) {                                                             // chain step without loads
        const float d = __fsub_rn(c, x);
        x = __fadd_rn(x, __fmaf_rn(d, a, __fmul_rn(d, b)));
      }
      if (MODE == 5) {                                                             // chain + independent side work (variance)
        const float d = __fsub_rn(c, x);
        const float xn = __fadd_rn(x, __fmaf_rn(d, a, __fmul_rn(d, b)));
        y = __fadd_rn(y, __fmul_rn(d, __fsub_rn(c, xn)));
        x = xn;
      }
    }
  }
  const long long t1 = clock64();
  out[threadIdx.x] = x + y;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, int ops_per_u)
{
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1e-3f * (i % 97) + 0.5f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int iters = 20000;
  k<MODE><<<1, threads>>>(out, in, cyc, iters);
  k<MODE><<<1, threads>>>(out, in, cyc, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %6.2f cycles per step (%d dependent ops) -> %.2f per op\n", name, (double)c / (iters * 16.0), ops_per_u,
         (double)c / (iters * 16.0) / ops_per_u);
}

int main()
{
  run<0>("FFMA chain, 1 warp", 32, 1);
  run<1>("FADD chain, 1 warp", 32, 1);
  run<4>("FADD-FMUL-FFMA-FADD, 1 warp", 32, 4);
  run<5>("same + variance side work, 1 warp", 32, 4);
  run<2>("same + 2 LDS per step, 1 warp", 32, 4);
  run<3>("same + 2 LDS, 3 more warps spinning on LDS", 128, 4);
  run<3>("same + 2 LDS, 7 more warps spinning on LDS", 256, 4);
  return 0;
}
