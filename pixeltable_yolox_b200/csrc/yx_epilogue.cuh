// Conv epilogue shared by the tcgen05 implicit-GEMM kernel and the SIMT verification kernel:
// bias (folded BN shift, yolox/utils/model_utils.py:63-73) + activation (network_blocks.py:14-24)
// + Bottleneck shortcut (network_blocks.py:95-99) + fused 2x nearest upsample store
// (yolo_pafpn.py:97-99) or the YOLOX head decode (yolo_head.py:185-187, 213-251).
#pragma once
#include "yx_common.cuh"

namespace yx {

struct EpiParams {
  int out_h, out_w, out_c;
  int act;
  int dtype;     // yx_dtype of out/res/ups
  int epilogue;  // yx_epilogue
  const float* bias;
  void* out; long long out_ld;
  const void* res; long long res_ld;
  void* ups; long long ups_ld;
  void* out2; long long out2_ld; int out2_begin;   // channels >= out2_begin go to out2 (null: disabled)
  float* head_out;
  int head_anchors, head_anchor_off, head_nc, head_decode;
  float head_stride;
  float* head_cand; unsigned long long* head_keys; int* head_counts;   // fused score filter (null: off)
  float head_conf; int head_xyxy;
  int shuf_c;     // depth-to-space store (yx_conv_desc.shuffle2_c), tcgen05 path only
};

#ifdef __CUDACC__
template <bool PRECISE>
__device__ __forceinline__ float act_f(float x, int act) {
  if (act == YX_ACT_SILU) {
    if (PRECISE) return x / (1.0f + expf(-x));
    return __fdividef(x, 1.0f + __expf(-x));
  }
  if (act == YX_ACT_RELU) return fmaxf(x, 0.0f);
  if (act == YX_ACT_LRELU) return x > 0.0f ? x : 0.1f * x;
  return x;
}

// v[0..16) = raw accumulators of channels [c0, c0+16) of output pixel (b, ho, wo).
template <bool PRECISE>
__device__ __forceinline__ void epi_store16(const EpiParams& e, int b, int ho, int wo, int c0,
                                            float (&v)[16]) {
  if (e.epilogue == YX_EPI_HEAD) {
    // head_decode bit0: box decode ((v+grid)*s, exp(v)*s); bit1: sigmoid on obj/cls
    const int nch = 5 + e.head_nc;
    const long long a = (long long)b * e.head_anchors + e.head_anchor_off + (long long)ho * e.out_w + wo;
    float* dst = e.head_out + a * nch;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = c0 + j;
      if (c < nch) {
        float x = v[j] + __ldg(e.bias + c);
        if (c < 2) {
          if (e.head_decode & 1) x = (x + (c == 0 ? (float)wo : (float)ho)) * e.head_stride;
        } else if (c < 4) {
          if (e.head_decode & 1) x = expf(x) * e.head_stride;
        } else {
          if (e.head_decode & 2) x = 1.0f / (1.0f + expf(-x));
        }
        dst[c] = x;
      }
    }
    return;
  }

  const long long pix = ((long long)b * e.out_h + ho) * e.out_w + wo;
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(e.bias + c0 + j));
    v[j + 0] = act_f<PRECISE>(v[j + 0] + bb.x, e.act);
    v[j + 1] = act_f<PRECISE>(v[j + 1] + bb.y, e.act);
    v[j + 2] = act_f<PRECISE>(v[j + 2] + bb.z, e.act);
    v[j + 3] = act_f<PRECISE>(v[j + 3] + bb.w, e.act);
  }
  if (e.dtype == YX_FP32) {
    if (e.res) {
      const float4* r = reinterpret_cast<const float4*>((const float*)e.res + pix * e.res_ld + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 rr = r[j];
        v[4 * j + 0] += rr.x; v[4 * j + 1] += rr.y; v[4 * j + 2] += rr.z; v[4 * j + 3] += rr.w;
      }
    }
    float4* o = (e.out2 && c0 >= e.out2_begin)
                    ? reinterpret_cast<float4*>((float*)e.out2 + pix * e.out2_ld + (c0 - e.out2_begin))
                    : reinterpret_cast<float4*>((float*)e.out + pix * e.out_ld + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    if (e.ups) {
      const int uw = 2 * e.out_w;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long long up = ((long long)b * 2 * e.out_h + 2 * ho + (q >> 1)) * uw + 2 * wo + (q & 1);
        float4* u = reinterpret_cast<float4*>((float*)e.ups + up * e.ups_ld + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
    return;
  }

  const bool fp16 = (e.dtype == YX_FP16);
  if (e.res) {
    const uint4* r = reinterpret_cast<const uint4*>((const uint16_t*)e.res + pix * e.res_ld + c0);
    const uint4 r0 = r[0], r1 = r[1];
    const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a, c;
      unpack16(rw[j], fp16, a, c);
      v[2 * j] += a; v[2 * j + 1] += c;
    }
  }
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = pack16(v[2 * j], v[2 * j + 1], fp16);
  const uint4 o0 = make_uint4(w[0], w[1], w[2], w[3]);
  const uint4 o1 = make_uint4(w[4], w[5], w[6], w[7]);
  uint4* o = (e.out2 && c0 >= e.out2_begin)
                 ? reinterpret_cast<uint4*>((uint16_t*)e.out2 + pix * e.out2_ld + (c0 - e.out2_begin))
                 : reinterpret_cast<uint4*>((uint16_t*)e.out + pix * e.out_ld + c0);
  o[0] = o0; o[1] = o1;
  if (e.ups) {
    const int uw = 2 * e.out_w;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long up = ((long long)b * 2 * e.out_h + 2 * ho + (q >> 1)) * uw + 2 * wo + (q & 1);
      uint4* u = reinterpret_cast<uint4*>((uint16_t*)e.ups + up * e.ups_ld + c0);
      u[0] = o0; u[1] = o1;
    }
  }
}
#endif

int fill_epi_params(const yx_conv_desc* d, EpiParams* e);  // validates the epilogue part of d

}  // namespace yx
