// Microbenchmark: issue rate of the XU (MUFU) ops the SiLU epilogue can be built from. nvcc -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) { unsigned u = __float_as_uint(v[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
      if (OP == 4) { unsigned u = __float_as_uint(v[i]); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
      if (OP == 5) v[i] = fmaf(v[i], 1.0001f, 0.5f);
      // the 16-bit pack of the epilogues: cvt.rn.bf16x2.f32 (SASS F2FP.BF16.F32.PACK_AB) and cvt.rn.f16x2.f32
      if (OP == 6) { unsigned u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(v[i]), "f"(v[(i + 1) & 7])); v[i] = __uint_as_float(u | 0x3f000000u); }
      if (OP == 7) { unsigned u; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(v[i]), "f"(v[(i + 1) & 7])); v[i] = __uint_as_float(u | 0x3f000000u); }
      // the same bf16 round-to-nearest-even pack on the integer pipe: u + 0x7fff + lsb, then a byte permute
      if (OP == 8) {
        unsigned a = __float_as_uint(v[i]), b = __float_as_uint(v[(i + 1) & 7]);
        a += 0x7fffu + ((a >> 16) & 1u); b += 0x7fffu + ((b >> 16) & 1u);
        unsigned u = __byte_perm(a, b, 0x7632);
        v[i] = __uint_as_float(u | 0x3f000000u);
      }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMallocManaged(&cyc, 8);
  const char* names[9] = {"tanh.approx.f32", "ex2.approx.f32", "rcp.approx.f32", "tanh.approx.f16x2", "tanh.approx.bf16x2", "fma.f32",
                          "cvt.rn.bf16x2.f32", "cvt.rn.f16x2.f32", "bf16x2 pack (int pipe)"};
  const int iters = 2000;
  for (int warps = 4; warps <= 16; warps *= 2)
    for (int op = 0; op < 9; ++op) {
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) k<0><<<148, warps * 32>>>(out, iters, cyc);
        if (op == 1) k<1><<<148, warps * 32>>>(out, iters, cyc);
        if (op == 2) k<2><<<148, warps * 32>>>(out, iters, cyc);
        if (op == 3) k<3><<<148, warps * 32>>>(out, iters, cyc);
        if (op == 4) k<4><<<148, warps * 32>>>(out, iters, cyc);
        if (op == 5) k<5><<<148, warps * 32>>>(out, iters, cyc);
        if (op == 6) k<6><<<148, warps * 32>>>(out, iters, cyc);
        if (op == 7) k<7><<<148, warps * 32>>>(out, iters, cyc);
        if (op == 8) k<8><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
      }
      double per = (double)*cyc / (iters * 8.0 * warps / 4.0);   // cycles per warp-instruction per SM sub-partition
      printf("warps/SM=%2d %-20s %6.2f clk per warp-instr per SMSP\n", warps, names[op], per);
    }
  return 0;
}
