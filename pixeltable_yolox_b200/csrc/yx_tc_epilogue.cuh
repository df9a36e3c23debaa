// Epilogue helpers shared by the tcgen05 kernels (implicit-GEMM conv and fused Focus stem):
// 256-bit global accesses, single-MUFU SiLU, and the 16-column bias + activation (+ residual,
// + 2x upsample copy) store.  network_blocks.py:14-24, 95-99; yolo_pafpn.py:97-99.
#pragma once
#include "yx_epilogue.cuh"

namespace yx {

// n / d for 0 <= n < 2^31 with mul = floor(2^32 / d) + 1 (d >= 2); the estimate is never low and at most 1 high
__device__ __forceinline__ int fast_div(int n, unsigned mul, int d) {
  if (d == 1) return n;
  int q = (int)__umulhi((unsigned)n, mul);
  if ((long long)q * d > n) --q;
  return q;
}
inline unsigned fast_div_mul(int d) { return d <= 1 ? 0u : (unsigned)((1ull << 32) / (unsigned)d) + 1u; }

__device__ __forceinline__ void ld_global_256(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// SiLU with one MUFU: x*sigmoid(x) = h + h*tanh(h), h = x/2 (tanh.approx.f32: ~2^-11 relative,
// below the bf16/fp16 output rounding)
__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// SiLU for fp16 outputs (11-bit significand): x * sigmoid(x) with ex2 + rcp. The single-MUFU tanh form loses the small
// results of negative inputs to cancellation in 1 + tanh(h) (1.3 % at x = -4), visible in fp16, invisible in bf16.
// (Replacing the rcp by a magic-constant seed + two Newton steps on the FMA pipe, one MUFU instead of two, measured
// SLOWER: yolox_l fp16 12.7 vs 12.0 ms per 64 images -- the epilogues are issue-bound before they are XU-bound.)
__device__ __forceinline__ float silu_exp(float x) {
  float e;
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

// true when the kernel stages 0.5 * bias in shared memory (SiLU written as h + h * tanh(h), h = x / 2)
__device__ __forceinline__ bool epi_half_bias(const EpiParams& e) { return e.act == YX_ACT_SILU && e.dtype == YX_BF16; }

// 16 accumulator columns of one pixel: bias (smem) + act (+ residual) -> 16-bit, one 32-byte store.
// `bias` holds 0.5 * bias when epi_half_bias(e).
template <bool FP16>
__device__ __forceinline__ uint32_t pack16_t(float a, float b) {
  if (FP16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <bool FP16>
__device__ __forceinline__ void unpack16_t(uint32_t w, float& a, float& b) {
  if (FP16) {
    float2 f = __half22float2(*reinterpret_cast<__half2*>(&w));
    a = f.x; b = f.y;
  } else {
    a = __uint_as_float(w << 16);            // bf16 -> fp32 is a shift
    b = __uint_as_float(w & 0xffff0000u);
  }
}

// FP16 is a compile-time parameter: a runtime flag costs a branch per packed pair in the hot loop
#ifdef YX_EPI_NOBIAS   // experiment builds only: how much of an epilogue is the shared-memory read of the bias?
#define YX_BIAS4(ptr) make_float4(0.f, 0.f, 0.f, 0.f)
#else
#define YX_BIAS4(ptr) (*reinterpret_cast<const float4*>(ptr))
#endif
template <bool FP16, bool SILU_ONLY = false>
__device__ __forceinline__ void epi_tc_chunk(const EpiParams& e, const uint32_t (&raw)[16], const float* bias,
                                             const uint32_t* res, uint16_t* dst, int b, int ho, int wo, int c0) {
  constexpr bool fp16 = FP16;
  float v[16];
  // SILU_ONLY: the caller guarantees act == SiLU, the generic activation code (a runtime switch per element) is not
  // even compiled: bf16 takes the single-MUFU tanh form, fp16 the ex2 + rcp form
  const bool silu_fast = (SILU_ONLY && !fp16) || (e.act == YX_ACT_SILU && !fp16);
  if (SILU_ONLY && fp16) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 bb = YX_BIAS4(bias + j);
      v[j + 0] = silu_exp(__uint_as_float(raw[j + 0]) + bb.x);
      v[j + 1] = silu_exp(__uint_as_float(raw[j + 1]) + bb.y);
      v[j + 2] = silu_exp(__uint_as_float(raw[j + 2]) + bb.z);
      v[j + 3] = silu_exp(__uint_as_float(raw[j + 3]) + bb.w);
    }
  } else if (silu_fast) {
    // bf16 output (8-bit significand): the 2^-11 absolute error of tanh.approx is invisible.
    // 3 instructions per element: h = fma(acc, 0.5, bias/2); t = tanh(h); y = fma(h, t, h)
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 bb = YX_BIAS4(bias + j);
      const float hb[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float h = fmaf(__uint_as_float(raw[j + q]), 0.5f, hb[q]);
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
        v[j + q] = fmaf(h, t, h);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 bb = YX_BIAS4(bias + j);
      v[j + 0] = __uint_as_float(raw[j + 0]) + bb.x;
      v[j + 1] = __uint_as_float(raw[j + 1]) + bb.y;
      v[j + 2] = __uint_as_float(raw[j + 2]) + bb.z;
      v[j + 3] = __uint_as_float(raw[j + 3]) + bb.w;
    }
  }
  if (silu_fast || SILU_ONLY) {
  } else if (e.act != YX_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = act_f<false>(v[j], e.act);
  }
  if (res) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a, c;
      unpack16_t<FP16>(res[j], a, c);
      v[2 * j] += a; v[2 * j + 1] += c;
    }
  }
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = pack16_t<FP16>(v[2 * j], v[2 * j + 1]);
  st_global_256(dst, w);
  if (e.ups) {
    const int uw = 2 * e.out_w;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long up = ((long long)b * 2 * e.out_h + 2 * ho + (q >> 1)) * uw + 2 * wo + (q & 1);
      st_global_256((uint16_t*)e.ups + up * e.ups_ld + c0, w);
    }
  }
}


}  // namespace yx
