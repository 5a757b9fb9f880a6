// Pipe-throughput micro-benchmark for the softmax instruction mix (sm_100a): cycles per warp-instruction of MUFU.EX2, F2FP
// (fp32 pair -> fp16x2), FMNMX3, FFMA2, FADD2 and of tcgen05.ld / tcgen05.st, with 1 / 2 / 4 warps per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/pipes tools/ubench/pipes.cu && gpurun_out/pipes
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
constexpr int ITERS = 256, UNROLL = 16;
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[UNROLL];
  uint32_t h[UNROLL];
  for (int i = 0; i < UNROLL; ++i) { a[i] = seed + 0.001f * (threadIdx.x + i); h[i] = 0; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i + 1) % UNROLL])); }
      if (OP == 2) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) % UNROLL]), "f"(a[(i + 2) % UNROLL]));
      if (OP == 3) { unsigned long long p, q; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a[i]), "f"(a[(i + 1) % UNROLL]));
                     asm volatile("fma.rn.f32x2 %0, %1, %1, %1;" : "=l"(q) : "l"(p)); asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[(i + 1) % UNROLL]) : "l"(q)); }
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 5) { asm volatile("shf.r.clamp.b32 %0, %1, 13, %2;" : "=r"(h[i]) : "r"(__float_as_uint(a[i])), "r"(__float_as_uint(a[(i + 1) % UNROLL]))); }
      if (OP == 6) { asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(h[i]) : "r"(__float_as_uint(a[i])), "r"(__float_as_uint(a[(i + 1) % UNROLL]))); }
    }
  }
  const long long t1 = clock64();
  float s = 0.f; uint32_t hs = 0;
  for (int i = 0; i < UNROLL; ++i) { s += a[i]; hs ^= h[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(hs);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// TMEM: each of 4 warps (one per lane quarter) issues ITERS loads / stores of 32 lanes x 32 columns
template <int ST>
__global__ void ktm(float* out, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
                   asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32];
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    const uint32_t ta = base + (uint32_t)((it & 7) * 32);
    if (ST) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                   ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                     "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    } else {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                     "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(ta) : "memory");
    }
  }
  if (ST) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 32; ++i) s ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(slot) : "memory");
}
template <int OP> void run(const char* name, float* out, long long* cyc) {
  for (int warps : {4, 8, 16}) {      // 1, 2, 4 warps per sub-partition, one CTA on one SM
    k<OP><<<1, warps * 32>>>(out, cyc, 0.5f);
    k<OP><<<1, warps * 32>>>(out, cyc, 0.5f);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-10s warps/SMSP=%d: %.2f cycles per warp-instruction per SMSP\n", name, warps / 4, (double)c / (ITERS * UNROLL) / (warps / 4));
  }
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  run<0>("MUFU.EX2", out, cyc); run<1>("F2FP.f16x2", out, cyc); run<2>("FMNMX3", out, cyc); run<3>("FFMA2", out, cyc);
  run<4>("FFMA", out, cyc); run<5>("SHF", out, cyc); run<6>("PRMT", out, cyc);
  for (int ctas : {1, 2}) {
    for (int st : {0, 1}) {
      for (int rep = 0; rep < 2; ++rep) { if (st) ktm<1><<<ctas, 128>>>(out, cyc); else ktm<0><<<ctas, 128>>>(out, cyc); }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%s 32x32b.x32, 4 warps x %d CTA(s): %.1f cycles per instruction per warp -> %.1f B/clk per CTA\n", st ? "tcgen05.st" : "tcgen05.ld", ctas,
             (double)c / ITERS, 4.0 * 4096.0 * ITERS / (double)c);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
