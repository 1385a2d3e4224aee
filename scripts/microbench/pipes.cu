// per-SM pipe throughput microbenchmarks: each kernel runs `wps` warps per SM sub-partition on every SM
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 32
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm volatile("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}" : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float b0, float b1) {
  asm volatile("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\t"
      "add.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}" : "+f"(d0), "+f"(d1) : "f"(b0), "f"(b1));
}
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[N];
  for (int i = 0; i < N; ++i) a[i] = seed + threadIdx.x * 0.001f + i;
  float l0 = 0, l1 = 0, l2 = 0, l3 = 0; uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // MUFU.EX2 only
#pragma unroll
      for (int i = 0; i < N; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    } else if (MODE == 1) {  // F2FP pack only
#pragma unroll
      for (int i = 0; i < N; i += 2) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[i+1])); acc ^= r; }
    } else if (MODE == 2) {  // FFMA2 only
#pragma unroll
      for (int i = 0; i < N; i += 2) ffma2(a[i], a[i+1], a[i], a[i+1], 1.0001f, 0.5f);
    } else if (MODE == 3) {  // FADD2 only
#pragma unroll
      for (int i = 0; i < N; i += 4) { fadd2(l0, l1, a[i], a[i+1]); fadd2(l2, l3, a[i+2], a[i+3]); }
    } else if (MODE == 4) {  // softmax mix: per 4 elements 2 ffma2, 4 ex2, 2 fadd2, 2 cvt (non-volatile: compiler schedules)
      float x[N];
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
            "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}" : "=f"(x[i]), "=f"(x[i+1]) : "f"(a[i]), "f"(a[i+1]), "f"(0.25f), "f"(-seed));
      }
#pragma unroll
      for (int i = 0; i < N; ++i) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[i]) : "f"(x[i]));
#pragma unroll
      for (int i = 0; i < N; i += 4) {
        asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}" : "+f"(l0), "+f"(l1) : "f"(x[i]), "f"(x[i+1]));
        asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}" : "+f"(l2), "+f"(l3) : "f"(x[i+2]), "f"(x[i+3]));
        uint32_t r0, r1;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r0) : "f"(x[i+1]), "f"(x[i]));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r1) : "f"(x[i+3]), "f"(x[i+2]));
        acc ^= r0 ^ r1;
      }
#pragma unroll
      for (int i = 0; i < N; ++i) a[i] += 1e-7f * l0;   // keeps iterations dependent (cheap FFMA)
    } else if (MODE == 5) {  // mix without cvt
      float x[N];
#pragma unroll
      for (int i = 0; i < N; i += 2) ffma2(x[i], x[i+1], a[i], a[i+1], 0.25f, -seed);
#pragma unroll
      for (int i = 0; i < N; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
      for (int i = 0; i < N; i += 4) { fadd2(l0, l1, x[i], x[i+1]); fadd2(l2, l3, x[i+2], x[i+3]); }
    } else if (MODE == 7) {  // softmax mix with a forced software pipeline: the consumers of element pair p carry a
      // (numerically void) dependency on the MUFU result DIST elements later, so ptxas cannot schedule them right
      // behind their own MUFU.EX2 (its default: 2 MUFU, then FADD2 + F2FP on those results -> the warp stalls for the
      // full SFU latency every pair)
      constexpr int DIST = 8;
      float x[N];
#pragma unroll
      for (int i = 0; i < N; i += 2) ffma2(x[i], x[i+1], a[i], a[i+1], 0.25f, -seed);
#pragma unroll
      for (int i = 0; i < N; ++i) asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
      for (int i = 0; i < N; i += 4) {
        float y0 = x[i], y1 = x[i+1], y2 = x[i+2], y3 = x[i+3];
        if (i + DIST < N) {
          asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tmov.b64 rc, {%4, %4};\n\t"
              "fma.rn.f32x2 ra, rb, rc, ra;\n\tmov.b64 {%0, %1}, ra;\n\t}" : "+f"(y0), "+f"(y1) : "f"(x[i+DIST]), "f"(x[i+DIST+1]), "f"(0.0f));
          asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tmov.b64 rc, {%4, %4};\n\t"
              "fma.rn.f32x2 ra, rb, rc, ra;\n\tmov.b64 {%0, %1}, ra;\n\t}" : "+f"(y2), "+f"(y3) : "f"(x[i+DIST+2]), "f"(x[i+DIST+3]), "f"(0.0f));
        }
        asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}" : "+f"(l0), "+f"(l1) : "f"(y0), "f"(y1));
        asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}" : "+f"(l2), "+f"(l3) : "f"(y2), "f"(y3));
        uint32_t r0, r1;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r0) : "f"(y1), "f"(y0));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r1) : "f"(y3), "f"(y2));
        acc ^= r0 ^ r1;
      }
#pragma unroll
      for (int i = 0; i < N; ++i) a[i] += 1e-7f * l0;
    } else if (MODE == 6) {  // FMNMX3
#pragma unroll
      for (int i = 0; i < N; i += 2) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(l0) : "f"(a[i]), "f"(a[i+1]));
    }
  }
  float s = l0 + l1 + l2 + l3 + __uint_as_float(acc);
  for (int i = 0; i < N; ++i) s += a[i];
  if (s == 12345.678f) out[threadIdx.x] = s;
}
template <int MODE> void run(const char* nm, double ops_per_iter_per_thread, int wps) {
  float* d; cudaMalloc(&d, 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 4000, threads = 128 * wps;
  k<MODE><<<148, threads>>>(d, 100, 1.f);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(d, iters, 1.f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * threads * ops_per_iter_per_thread * iters;
  // report: thread-level ops per clk per SM assuming 1.90 GHz
  printf("%-28s wps=%d: %8.3f ms  %7.2f lane-ops/clk/SM (@1.9GHz)  -> %.2f clk per warp-instr per SMSP\n", nm, wps, ms,
         ops / (ms * 1e-3) / 148 / 1.9e9, 32.0 / (ops / (ms * 1e-3) / 148 / 1.9e9 / 4));
  cudaFree(d);
}
int main() {
  for (int wps : {1, 2, 4}) {
    run<0>("MUFU.EX2", N, wps);
    run<1>("F2FP.BF16.PACK (per instr)", N / 2, wps);
    run<2>("FFMA2 (per instr)", N / 2, wps);
    run<3>("FADD2 (per instr)", N / 2, wps);
    run<6>("FMNMX3 (per instr)", N / 2, wps);
    run<4>("softmax mix (per element)", N, wps);
    run<7>("mix, forced sw pipeline", N, wps);
  }
  return 0;
}
