// Bandwidth-bound helper kernels of the YOLOX forward pass (NHWC, 128-bit accesses):
//   focus_s2d   Focus space-to-depth          yolox/models/network_blocks.py:193-208
//   spp_maxpool SPP 5/9/13 max pools          yolox/models/network_blocks.py:120-142
//   dwconv3x3   depthwise conv + BN + act     yolox/models/network_blocks.py:55-74
//   pack        BN folding + weight repack    yolox/utils/model_utils.py:33-75
//   decode      YoloxHead.decode_outputs      yolox/models/yolo_head.py:233-251
//   iou         bboxes_iou                    yolox/utils/boxes.py:78-101
#include <string.h>
#include <math.h>

#include "yx_epilogue.cuh"

namespace yx {

// ------------------------------------------------------------------------------------------
// Focus: out[b, y, x, (2*dx + dy)*3 + c] = img[b, c, 2y + dy, 2x + dx]
// (concat order TL, BL, TR, BR = (dy,dx) (0,0), (1,0), (0,1), (1,1), network_blocks.py:195-207)
// One thread per output pixel; image rows are read as coalesced 2-element runs per thread.
// ------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void focus_s2d_kernel(const TI* __restrict__ img, TO* __restrict__ out, long long out_ld,
                                 int batch, int h, int w) {
  const int oh = h / 2, ow = w / 2;
  const long long total = (long long)batch * oh * ow;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % ow);
  const long long t = idx / ow;
  const int y = (int)(t % oh);
  const int b = (int)(t / oh);
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const TI* base = img + (((long long)b * 3 + c) * h + 2 * y) * w + 2 * x;
    const float tl = (float)base[0], tr = (float)base[1];
    const float bl = (float)base[w], br = (float)base[w + 1];
    v[0 + c] = tl;  // patch 0: top-left
    v[3 + c] = bl;  // patch 1: bottom-left
    v[6 + c] = tr;  // patch 2: top-right
    v[9 + c] = br;  // patch 3: bottom-right
  }
  TO* o = out + idx * out_ld;
#pragma unroll
  for (int j = 0; j < 16; ++j) o[j] = Cvt<TO>::from_f(v[j]);
}

template <typename TI>
static int focus_dispatch_out(const TI* img, void* out, long long out_ld, int out_dtype, int batch,
                              int h, int w, cudaStream_t s) {
  const long long total = (long long)batch * (h / 2) * (w / 2);
  const unsigned grid = (unsigned)ceil_div64(total, 256);
  switch (out_dtype) {
    case YX_BF16: focus_s2d_kernel<TI, __nv_bfloat16><<<grid, 256, 0, s>>>(img, (__nv_bfloat16*)out, out_ld, batch, h, w); break;
    case YX_FP16: focus_s2d_kernel<TI, __half><<<grid, 256, 0, s>>>(img, (__half*)out, out_ld, batch, h, w); break;
    case YX_FP32: focus_s2d_kernel<TI, float><<<grid, 256, 0, s>>>(img, (float*)out, out_ld, batch, h, w); break;
    default: YX_REQUIRE(false, YX_ERR_INVALID_ARG, "focus: bad out dtype %d", out_dtype);
  }
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

int focus_launch(const void* img, int img_dtype, void* out, long long out_ld, int out_dtype, int batch,
                 int h, int w, cudaStream_t s) {
  YX_REQUIRE(img && out, YX_ERR_INVALID_ARG, "focus: null pointer");
  YX_REQUIRE(h % 2 == 0 && w % 2 == 0 && h > 0 && w > 0 && batch > 0, YX_ERR_INVALID_ARG, "focus: H,W must be even and positive");
  YX_REQUIRE(out_ld >= 16, YX_ERR_INVALID_ARG, "focus: out_ld must be >= 16 (12 real + 4 zero channels)");
  if (img_dtype == YX_FP32) return focus_dispatch_out<float>((const float*)img, out, out_ld, out_dtype, batch, h, w, s);
  if (img_dtype == YX_U8) return focus_dispatch_out<uint8_t>((const uint8_t*)img, out, out_ld, out_dtype, batch, h, w, s);
  YX_REQUIRE(false, YX_ERR_INVALID_ARG, "focus: image dtype must be fp32 or uint8");
}

// ------------------------------------------------------------------------------------------
// SPP: m5, m9 = m5∘m5, m13 = m5∘m5∘m5 (exact: max is associative and the -inf padding of
// MaxPool2d commutes with the cascade). One block per (image, 8-channel group): the h x w x 8
// slab lives in shared memory as fp32; each 5x5 max is done separably (rows, then columns).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void spp_kernel(T* __restrict__ buf, long long ld, int h, int w, int c) {
  extern __shared__ float sm[];  // 3 slabs of h*w*8 floats
  const int hw = h * w;
  float* cur = sm;
  float* tmp = sm + (size_t)hw * 8;
  float* nxt = tmp + (size_t)hw * 8;
  const int groups = c / 8;
  const int b = blockIdx.x / groups;
  const int g = blockIdx.x - b * groups;
  T* base = buf + (long long)b * hw * ld + g * 8;
  for (int i = threadIdx.x; i < hw * 8; i += blockDim.x) {
    const int pix = i >> 3, ch = i & 7;
    cur[i] = Cvt<T>::to_f(base[(long long)pix * ld + ch]);
  }
  __syncthreads();
  for (int level = 1; level <= 3; ++level) {
    for (int i = threadIdx.x; i < hw * 8; i += blockDim.x) {  // horizontal 5-max
      const int pix = i >> 3, ch = i & 7;
      const int y = pix / w, x = pix - y * w;
      float m = -INFINITY;
#pragma unroll
      for (int d = -2; d <= 2; ++d) {
        const int xx = x + d;
        if (xx >= 0 && xx < w) m = fmaxf(m, cur[((y * w + xx) << 3) + ch]);
      }
      tmp[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < hw * 8; i += blockDim.x) {  // vertical 5-max
      const int pix = i >> 3, ch = i & 7;
      const int y = pix / w, x = pix - y * w;
      float m = -INFINITY;
#pragma unroll
      for (int d = -2; d <= 2; ++d) {
        const int yy = y + d;
        if (yy >= 0 && yy < h) m = fmaxf(m, tmp[((yy * w + x) << 3) + ch]);
      }
      nxt[i] = m;
      base[(long long)pix * ld + (long long)level * c + ch] = Cvt<T>::from_f(m);
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
}

// 16-bit fast path: one thread per pixel moves 8 channels as one 128-bit word; max is exact in
// bf16/fp16, so the cascade runs on packed pairs (HMNMX2) straight from shared memory.
template <typename T2>
__device__ __forceinline__ uint4 max4(const uint4 a, const uint4 b) {
  uint4 r;
  const T2* pa = reinterpret_cast<const T2*>(&a);
  const T2* pb = reinterpret_cast<const T2*>(&b);
  T2* pr = reinterpret_cast<T2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

// One CTA = one image x GW groups of 8 channels (GW*16 contiguous bytes per pixel: full 32-byte sectors for GW >= 2,
// whole 128-byte lines for GW = 8). Item i = pixel * GW + group; the 5x5 max is separable and 9 / 13 are cascades of 5.
template <typename T2, int GW>
__global__ void __launch_bounds__(512) spp16_kernel(uint16_t* __restrict__ buf, long long ld, int h, int w, int c) {
  extern __shared__ uint4 sp[];  // 2 slabs of h*w*GW 128-bit words
  const int hw = h * w, items = hw * GW;
  uint4* cur = sp;
  uint4* tmp = sp + items;
  const int chunks = c / (8 * GW);
  const int b = blockIdx.x / chunks;
  const int g = blockIdx.x - b * chunks;
  uint16_t* base = buf + (long long)b * hw * ld + g * (8 * GW);
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int px = i / GW, gi = i - px * GW;
    cur[i] = *reinterpret_cast<const uint4*>(base + (long long)px * ld + gi * 8);
  }
  __syncthreads();
  for (int level = 1; level <= 3; ++level) {
    for (int i = threadIdx.x; i < items; i += blockDim.x) {   // horizontal 5-max (window clipped = -inf padding)
      const int px = i / GW;
      const int y = px / w, x = px - y * w;
      uint4 m = cur[i];
#pragma unroll
      for (int d = -2; d <= 2; ++d) {
        const int xx = x + d;
        if (d != 0 && xx >= 0 && xx < w) m = max4<T2>(m, cur[i + d * GW]);
      }
      tmp[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < items; i += blockDim.x) {   // vertical 5-max
      const int px = i / GW, gi = i - px * GW;
      const int y = px / w;
      uint4 m = tmp[i];
#pragma unroll
      for (int d = -2; d <= 2; ++d) {
        const int yy = y + d;
        if (d != 0 && yy >= 0 && yy < h) m = max4<T2>(m, tmp[i + d * w * GW]);
      }
      cur[i] = m;   // each thread only overwrites its own item; others read tmp
      *reinterpret_cast<uint4*>(base + (long long)px * ld + (long long)level * c + gi * 8) = m;
    }
    __syncthreads();
  }
}

int spp_launch(void* buf, long long ld, int batch, int h, int w, int c, int dtype, cudaStream_t s) {
  YX_REQUIRE(buf, YX_ERR_INVALID_ARG, "spp: null buffer");
  YX_REQUIRE(c % 8 == 0 && ld >= 4 * (long long)c, YX_ERR_INVALID_ARG, "spp: c %% 8 != 0 or ld < 4c");
  const unsigned grid = (unsigned)(batch * (c / 8));
  if (dtype != YX_FP32 && (size_t)2 * h * w * 16 <= 200 * 1024 && ld % 8 == 0 && ((uintptr_t)buf & 15) == 0) {
    // widest channel group per CTA (contiguous bytes per pixel) whose two slabs fit ~100 KB, so two CTAs share an SM
    int gw = 1;
    while (gw < 8 && c % (16 * gw) == 0 && (size_t)2 * h * w * 16 * (2 * gw) <= 100 * 1024) gw *= 2;
    const size_t sm16 = (size_t)2 * h * w * 16 * gw;
    const unsigned grid16 = (unsigned)(batch * (c / (8 * gw)));
#define YX_SPP16(T2, GW)                                                                                         \
  do {                                                                                                           \
    if (sm16 > 48 * 1024) YX_CUDA(cudaFuncSetAttribute(spp16_kernel<T2, GW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm16)); \
    spp16_kernel<T2, GW><<<grid16, 512, sm16, s>>>((uint16_t*)buf, ld, h, w, c);                                  \
  } while (0)
#define YX_SPP16_T(T2) do { if (gw == 8) YX_SPP16(T2, 8); else if (gw == 4) YX_SPP16(T2, 4); else if (gw == 2) YX_SPP16(T2, 2); else YX_SPP16(T2, 1); } while (0)
    if (dtype == YX_BF16) YX_SPP16_T(__nv_bfloat162); else YX_SPP16_T(__half2);
#undef YX_SPP16_T
#undef YX_SPP16
    YX_CUDA(cudaGetLastError());
    return YX_OK;
  }
  const size_t smem = (size_t)3 * h * w * 8 * sizeof(float);
  YX_REQUIRE(smem <= 200 * 1024, YX_ERR_UNSUPPORTED, "spp: feature map %dx%d too large for the shared-memory slab", h, w);
#define YX_SPP(T)                                                                                   \
  do {                                                                                              \
    if (smem > 48 * 1024) YX_CUDA(cudaFuncSetAttribute(spp_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    spp_kernel<T><<<grid, 256, smem, s>>>((T*)buf, ld, h, w, c);                                    \
  } while (0)
  switch (dtype) {
    case YX_BF16: YX_SPP(__nv_bfloat16); break;
    case YX_FP16: YX_SPP(__half); break;
    case YX_FP32: YX_SPP(float); break;
    default: YX_REQUIRE(false, YX_ERR_INVALID_ARG, "spp: bad dtype %d", dtype);
  }
#undef YX_SPP
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

// ------------------------------------------------------------------------------------------
// Depthwise 3x3 + folded BN + activation. Thread = one output pixel x 8 channels (128-bit
// NHWC loads for 16-bit types); weights [9][c] tap-major so a tap is one 128-bit load too.
// ------------------------------------------------------------------------------------------
template <typename T, bool PRECISE>
__global__ void dwconv3x3_kernel(const T* __restrict__ in, long long in_ld, const T* __restrict__ wt,
                                 const float* __restrict__ bias, T* __restrict__ out, long long out_ld,
                                 int batch, int in_h, int in_w, int c, int out_h, int out_w, int stride,
                                 int act) {
  const int groups = c / 8;
  const long long total = (long long)batch * out_h * out_w * groups;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  long long t = idx / groups;
  const int wo = (int)(t % out_w); t /= out_w;
  const int ho = (int)(t % out_h);
  const int b = (int)(t / out_h);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  for (int fr = 0; fr < 3; ++fr) {
    const int hi = ho * stride + fr - 1;
    if (hi < 0 || hi >= in_h) continue;
    for (int fs = 0; fs < 3; ++fs) {
      const int wi = wo * stride + fs - 1;
      if (wi < 0 || wi >= in_w) continue;
      const T* ip = in + (((long long)b * in_h + hi) * in_w + wi) * in_ld + g * 8;
      const T* wp = wt + (long long)(fr * 3 + fs) * c + g * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(Cvt<T>::to_f(ip[j]), Cvt<T>::to_f(wp[j]), acc[j]);
    }
  }
  T* op = out + (((long long)b * out_h + ho) * out_w + wo) * out_ld + g * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) op[j] = Cvt<T>::from_f(act_f<PRECISE>(acc[j] + bias[g * 8 + j], act));
}

int dwconv_launch(const void* in, long long in_ld, const void* w, const float* bias, void* out,
                  long long out_ld, int batch, int in_h, int in_w, int c, int stride, int act, int dtype,
                  cudaStream_t s) {
  YX_REQUIRE(in && w && bias && out, YX_ERR_INVALID_ARG, "dwconv: null pointer");
  YX_REQUIRE(c % 8 == 0 && (stride == 1 || stride == 2), YX_ERR_INVALID_ARG, "dwconv: c %% 8 != 0 or bad stride");
  const int out_h = (in_h + 2 - 3) / stride + 1, out_w = (in_w + 2 - 3) / stride + 1;
  const long long total = (long long)batch * out_h * out_w * (c / 8);
  const unsigned grid = (unsigned)ceil_div64(total, 256);
  switch (dtype) {
    case YX_BF16: dwconv3x3_kernel<__nv_bfloat16, false><<<grid, 256, 0, s>>>((const __nv_bfloat16*)in, in_ld, (const __nv_bfloat16*)w, bias, (__nv_bfloat16*)out, out_ld, batch, in_h, in_w, c, out_h, out_w, stride, act); break;
    case YX_FP16: dwconv3x3_kernel<__half, false><<<grid, 256, 0, s>>>((const __half*)in, in_ld, (const __half*)w, bias, (__half*)out, out_ld, batch, in_h, in_w, c, out_h, out_w, stride, act); break;
    case YX_FP32: dwconv3x3_kernel<float, true><<<grid, 256, 0, s>>>((const float*)in, in_ld, (const float*)w, bias, (float*)out, out_ld, batch, in_h, in_w, c, out_h, out_w, stride, act); break;
    default: YX_REQUIRE(false, YX_ERR_INVALID_ARG, "dwconv: bad dtype %d", dtype);
  }
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

// ------------------------------------------------------------------------------------------
// BN folding + repack:  W' = W * gamma / sqrt(var + eps),  b' = beta - gamma*mean/sqrt(var+eps)
// (+ conv bias scaled the same way), written K-major [o][kh*kw][i] in the target dtype.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_kernel(const float* __restrict__ src, const float* __restrict__ gamma,
                            const float* __restrict__ beta, const float* __restrict__ mean,
                            const float* __restrict__ var, const float* __restrict__ conv_bias, float eps,
                            int o, int i, int khw, T* __restrict__ dst_w, int dst_o_off, int dst_i_off,
                            int dst_i_total, float* __restrict__ dst_b, int depthwise) {
  const long long total = (long long)o * i * khw;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int tap = (int)(idx % khw);
  long long t = idx / khw;
  const int ci = (int)(t % i);
  const int co = (int)(t / i);
  float scale = 1.0f;
  if (gamma) scale = gamma[co] / sqrtf(var[co] + eps);
  const float v = src[idx] * scale;
  if (depthwise) {
    dst_w[(long long)tap * dst_i_total + dst_i_off + co] = Cvt<T>::from_f(v);
  } else {
    dst_w[((long long)(dst_o_off + co) * khw + tap) * dst_i_total + dst_i_off + ci] = Cvt<T>::from_f(v);
  }
  if (ci == 0 && tap == 0 && dst_b) {
    float bsum = conv_bias ? conv_bias[co] * scale : 0.0f;
    if (gamma) bsum += beta[co] - mean[co] * scale;
    dst_b[(depthwise ? dst_i_off : dst_o_off) + co] = bsum;
  }
}

int pack_launch(const float* src, const float* gamma, const float* beta, const float* mean, const float* var,
                const float* conv_bias, float eps, int o, int i, int kh, int kw, void* dst_w, int dst_dtype,
                int dst_o_off, int dst_i_off, int dst_i_total, float* dst_b, int depthwise, cudaStream_t s) {
  YX_REQUIRE(src && dst_w, YX_ERR_INVALID_ARG, "pack: null pointer");
  YX_REQUIRE(!gamma || (beta && mean && var), YX_ERR_INVALID_ARG, "pack: incomplete BN parameter set");
  YX_REQUIRE(!depthwise || i == 1, YX_ERR_INVALID_ARG, "pack: depthwise weights must be [c][1][k][k]");
  const long long total = (long long)o * i * kh * kw;
  const unsigned grid = (unsigned)ceil_div64(total, 256);
#define YX_PACK(T) pack_kernel<T><<<grid, 256, 0, s>>>(src, gamma, beta, mean, var, conv_bias, eps, o, i, kh * kw, (T*)dst_w, dst_o_off, dst_i_off, dst_i_total, dst_b, depthwise)
  switch (dst_dtype) {
    case YX_BF16: YX_PACK(__nv_bfloat16); break;
    case YX_FP16: YX_PACK(__half); break;
    case YX_FP32: YX_PACK(float); break;
    default: YX_REQUIRE(false, YX_ERR_INVALID_ARG, "pack: bad dtype %d", dst_dtype);
  }
#undef YX_PACK
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

// ------------------------------------------------------------------------------------------
// decode_outputs on an undecoded [B, A, 5+nc] tensor, in place (only the 4 box columns change).
// ------------------------------------------------------------------------------------------
struct DecodeLevels {
  int n;
  int h[8], w[8], stride[8], off[8];
};

__global__ void decode_kernel(float* __restrict__ pred, int batch, int anchors, int nch, DecodeLevels lv) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)batch * anchors) return;
  const int a = (int)(idx % anchors);
  int l = 0;
  while (l + 1 < lv.n && a >= lv.off[l + 1]) ++l;
  const int r = a - lv.off[l];
  const int gy = r / lv.w[l], gx = r - gy * lv.w[l];
  const float s = (float)lv.stride[l];
  float* p = pred + idx * nch;
  p[0] = (p[0] + (float)gx) * s;
  p[1] = (p[1] + (float)gy) * s;
  p[2] = expf(p[2]) * s;
  p[3] = expf(p[3]) * s;
}

int decode_launch(float* pred, int batch, int anchors, int nc, const int* hw, const int* strides, int n_levels,
                  cudaStream_t s) {
  YX_REQUIRE(pred && hw && strides, YX_ERR_INVALID_ARG, "decode: null pointer");
  YX_REQUIRE(n_levels >= 1 && n_levels <= 8, YX_ERR_INVALID_ARG, "decode: 1..8 levels supported");
  DecodeLevels lv;
  lv.n = n_levels;
  int off = 0;
  for (int l = 0; l < n_levels; ++l) {
    lv.h[l] = hw[2 * l]; lv.w[l] = hw[2 * l + 1]; lv.stride[l] = strides[l]; lv.off[l] = off;
    off += hw[2 * l] * hw[2 * l + 1];
  }
  YX_REQUIRE(off == anchors, YX_ERR_INVALID_ARG, "decode: sum(h*w)=%d != anchors=%d", off, anchors);
  const long long total = (long long)batch * anchors;
  decode_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(pred, batch, anchors, 5 + nc, lv);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

// ------------------------------------------------------------------------------------------
// bboxes_iou (boxes.py:78-101): en = prod(tl < br); area_i = prod(br - tl) * en;
// iou = area_i / (area_a + area_b - area_i). Round-to-nearest intrinsics pin the fp32
// operation order of the reference (no FMA contraction).
// ------------------------------------------------------------------------------------------
__global__ void iou_kernel(const float* __restrict__ a, int n, const float* __restrict__ b, int m, int xyxy,
                           float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * m) return;
  const int i = (int)(idx / m), j = (int)(idx - (long long)i * m);
  const float4 A = reinterpret_cast<const float4*>(a)[i];
  const float4 B = reinterpret_cast<const float4*>(b)[j];
  float ax1, ay1, ax2, ay2, bx1, by1, bx2, by2, area_a, area_b;
  if (xyxy) {
    ax1 = A.x; ay1 = A.y; ax2 = A.z; ay2 = A.w; bx1 = B.x; by1 = B.y; bx2 = B.z; by2 = B.w;
    area_a = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
    area_b = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
  } else {
    ax1 = __fsub_rn(A.x, __fdiv_rn(A.z, 2.0f)); ay1 = __fsub_rn(A.y, __fdiv_rn(A.w, 2.0f));
    ax2 = __fadd_rn(A.x, __fdiv_rn(A.z, 2.0f)); ay2 = __fadd_rn(A.y, __fdiv_rn(A.w, 2.0f));
    bx1 = __fsub_rn(B.x, __fdiv_rn(B.z, 2.0f)); by1 = __fsub_rn(B.y, __fdiv_rn(B.w, 2.0f));
    bx2 = __fadd_rn(B.x, __fdiv_rn(B.z, 2.0f)); by2 = __fadd_rn(B.y, __fdiv_rn(B.w, 2.0f));
    area_a = __fmul_rn(A.z, A.w);
    area_b = __fmul_rn(B.z, B.w);
  }
  const float tlx = fmaxf(ax1, bx1), tly = fmaxf(ay1, by1);
  const float brx = fminf(ax2, bx2), bry = fminf(ay2, by2);
  const float en = (tlx < brx && tly < bry) ? 1.0f : 0.0f;
  const float area_i = __fmul_rn(__fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly)), en);
  out[idx] = __fdiv_rn(area_i, __fsub_rn(__fadd_rn(area_a, area_b), area_i));
}

int iou_launch(const float* a, int n, const float* b, int m, int xyxy, float* out, cudaStream_t s) {
  YX_REQUIRE(a && b && out, YX_ERR_INVALID_ARG, "iou: null pointer");
  if (n == 0 || m == 0) return YX_OK;
  const long long total = (long long)n * m;
  iou_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(a, n, b, m, xyxy, out);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

}  // namespace yx
