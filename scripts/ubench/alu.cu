// Issue-rate microbenchmark for the CUDA-core instructions the softmax / VQ inner loops are made of (sm_100a).
// Each kernel runs ITER x 64 independent-chain instructions per thread; reports warp-instructions / clk / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;
#define CHAINS 8

template <int OP>
__global__ void k(float* out, float a, float b, long long* clk) {
  float x[CHAINS * 2];
#pragma unroll
  for (int i = 0; i < CHAINS * 2; ++i) x[i] = a + threadIdx.x * 1e-3f + i;
  uint32_t h[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) h[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) {
        if (OP == 0) x[i] = fmaf(x[i], a, b);                               // FFMA
        if (OP == 1) {                                                       // FFMA2
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(*reinterpret_cast<uint64_t*>(&x[2 * i])) : "l"(*reinterpret_cast<uint64_t*>(&x[(2 * i + 2) % (2 * CHAINS)])), "l"(*reinterpret_cast<uint64_t*>(&x[(2 * i + 4) % (2 * CHAINS)])));
        }
        if (OP == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(*reinterpret_cast<uint64_t*>(&x[2 * i])) : "l"(*reinterpret_cast<uint64_t*>(&x[(2 * i + 2) % (2 * CHAINS)])));
        if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(*reinterpret_cast<uint64_t*>(&x[2 * i])) : "l"(*reinterpret_cast<uint64_t*>(&x[(2 * i + 2) % (2 * CHAINS)])));
        if (OP == 4) x[i] = x[i] * a;                                         // FMUL
        if (OP == 5) x[i] = x[i] + b;                                         // FADD
        if (OP == 6) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); // MUFU.EX2
        if (OP == 7) asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));  // F2FP (feeds nothing)
        if (OP == 8) { asm volatile("{.reg .b16 l, u; mov.b32 {l, u}, %1; cvt.f32.f16 %0, l;}" : "=f"(x[i]) : "r"(h[i])); }   // HADD2.F32 unpack
        if (OP == 9) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(x[(i + 1) % CHAINS]), "f"(x[(i + 2) % CHAINS]));   // FMNMX3
        if (OP == 10) x[i] = fmaxf(x[i], x[(i + 1) % CHAINS]);                // FMNMX
        if (OP == 11) { bool p = x[i] < x[(i + 3) % CHAINS]; x[i] = p ? x[(i + 1) % CHAINS] : x[i]; h[i] = p ? it : h[i]; }  // FSETP + 2 SEL
        if (OP == 12) x[i] = fmaf(x[i], x[(i + 1) % CHAINS], x[(i + 2) % CHAINS]);  // FFMA 3 distinct regs
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CHAINS * 2; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter_instr, int lanes_work) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4 * 4); cudaMalloc(&clk, 8);
  for (int warps : {4, 8, 16, 32}) {
    const int threads = warps * 32;
    k<OP><<<148, threads>>>(out, 1.0001f, 1e-7f, clk);
    k<OP><<<148, threads>>>(out, 1.0001f, 1e-7f, clk);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    const double winstr = (double)ITER * 8 * CHAINS * per_iter_instr * warps;
    printf("%-28s warps/SM %2d  %.3f warp-instr/clk/SM  (%.1f lane-ops/clk/SM)\n", name, warps, winstr / c, winstr / c * 32 * lanes_work);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  cudaFree(out); cudaFree(clk);
}

int main() {
  run<0>("FFMA (reg,imm-like a,b)", 1, 1);
  run<12>("FFMA 3 regs", 1, 1);
  run<1>("FFMA2 (f32x2)", 1, 2);
  run<2>("FMUL2", 1, 2);
  run<3>("FADD2", 1, 2);
  run<4>("FMUL", 1, 1);
  run<5>("FADD", 1, 1);
  run<6>("MUFU.EX2", 1, 1);
  run<7>("F2FP pack f16x2", 1, 2);
  run<8>("cvt f16->f32", 1, 1);
  run<9>("FMNMX3", 1, 2);
  run<10>("FMNMX", 1, 1);
  run<11>("FSETP+2SEL", 3, 1);
  return 0;
}
