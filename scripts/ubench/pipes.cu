// Micro-benchmark: issue / pipe throughput of the instructions the attention softmax is made of, per SM sub-partition,
// with 1, 2 and 4 resident warps per sub-partition.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

#define ITER 256
template <int MODE>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[16];
  uint64_t p[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i * 0.01f + threadIdx.x * 1e-4f;
#pragma unroll
  for (int i = 0; i < 8; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
    if (MODE == 0) {  // MUFU.EX2 x16 independent
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    } else if (MODE == 1) {  // FFMA2 x16 (8 regs x 2)
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i]));
    } else if (MODE == 2) {  // FFMA x16
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
    } else if (MODE == 3) {  // FADD2 x16
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(p[i]));
    } else if (MODE == 4) {  // cvt.rn.bf16x2.f32 x16
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        uint32_t w;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(a[i]), "f"(a[i + 1]));
        a[i] = __uint_as_float(w);
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(a[i + 1]), "f"(a[i]));
        a[i + 1] = __uint_as_float(w);
      }
    } else if (MODE == 5) {  // max3 x16 (fmaxf chain pairs -> FMNMX3)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaxf(fmaxf(a[i], a[(i + 1) & 15]), a[(i + 5) & 15]);
    } else if (MODE == 6) {  // mix: 8 MUFU + 16 FFMA2 interleaved (1 : 2)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i]));
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[(i + 4) & 7]));
      }
    } else if (MODE == 7) {  // mix: 8 MUFU + 8 FFMA2 + 8 cvt + 8 FADD2 (softmax-like 1 : 1 : 1 : 1)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t w;
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i]));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(a[8 + i]), "f"(a[i]));
        a[8 + i] = __uint_as_float(w);
        asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(p[(i + 3) & 7]));
      }
    } else if (MODE == 8) {  // mix: 16 MUFU + 16 FFMA (unpacked) interleaved
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[8 + i]));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  float lo, hi;
#pragma unroll
  for (int i = 0; i < 8; ++i) { asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
  for (int warps_per_smsp : {1, 2, 4}) {
    const int threads = 128 * warps_per_smsp;
    k<MODE><<<148, threads>>>(out, cyc, 0.5f);
    k<MODE><<<148, threads>>>(out, cyc, 0.5f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double per_smsp = double(ITER) * instr_per_iter * warps_per_smsp;  // warp-instructions per sub-partition
    printf("%-34s warps/SMSP=%d  cycles=%8.0f  clk per warp-instr per SMSP = %.3f\n", name, warps_per_smsp, c, c / per_smsp);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("MUFU.EX2", 16); run<1>("FFMA2", 16); run<2>("FFMA", 16); run<3>("FADD2", 16); run<4>("F2FP bf16x2", 16);
  run<5>("FMNMX3", 16); run<6>("8 MUFU + 16 FFMA2", 24); run<7>("8 MUFU+8 FFMA2+8 F2FP+8 FADD2", 32);
  run<8>("8 MUFU + 8 FFMA", 16);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
