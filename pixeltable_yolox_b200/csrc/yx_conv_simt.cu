// CUDA-core implementation of the conv contract of yx_conv_desc (same semantics as the
// tcgen05 kernel): the fp32 verification mode (SURVEY 8c': an fp32-faithful path is needed for
// the 1e-3 head-output gate) and the on-device cross-check of the tensor-core kernel.
// BaseConv.forward: yolox/models/network_blocks.py:48-49; BN folding: utils/model_utils.py:33-75.
//
// Block = 64 output pixels x 64 output channels, 256 threads, thread = 1 pixel x 16 channels.
// K is walked tap-major then channel (the packed weight order) in chunks of 16 with fp32 FFMA
// accumulation in ascending-k order, so results are deterministic and independent of the grid.
#include <string.h>
#include <math.h>

#include "yx_epilogue.cuh"

namespace yx {

int validate_conv_geometry(const yx_conv_desc* d);

struct ConvSimtParams {
  int batch, in_h, in_w, in_c, out_h, out_w, out_c, ksize, stride, pad;
  long long M;
  const void* in; long long in_ld;
  const void* w;
  EpiParams epi;
};

template <typename T>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvSimtParams p) {
  __shared__ float As[16][64 + 1];
  __shared__ float Ws[16][64];
  const int tid = threadIdx.x;
  const int px = tid & 63;
  const int cg = tid >> 6;
  const long long m0 = (long long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const int out_hw = p.out_h * p.out_w;
  const long long K = (long long)p.ksize * p.ksize * p.in_c;

  // loader role: pixel/out-channel (tid/4), 4 consecutive k at (tid%4)*4
  const int lrow = tid >> 2;
  const int lk = (tid & 3) * 4;
  const long long lm = m0 + lrow;
  int lb = 0, lho = 0, lwo = 0;
  const bool lvalid = lm < p.M;
  if (lvalid) {
    lb = (int)(lm / out_hw);
    const int r = (int)(lm - (long long)lb * out_hw);
    lho = r / p.out_w;
    lwo = r - lho * p.out_w;
  }
  const int loc = n0 + lrow;
  const T* in = reinterpret_cast<const T*>(p.in);
  const T* w = reinterpret_cast<const T*>(p.w);

  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.0f;

  for (int tap = 0; tap < p.ksize * p.ksize; ++tap) {
    const int fr = tap / p.ksize, fs = tap - fr * p.ksize;
    const int hi = lho * p.stride + fr - p.pad;
    const int wi = lwo * p.stride + fs - p.pad;
    const bool inb = lvalid && hi >= 0 && hi < p.in_h && wi >= 0 && wi < p.in_w;
    const T* arow = in + (((long long)lb * p.in_h + hi) * p.in_w + wi) * p.in_ld;
    const T* wrow = w + (long long)loc * K + (long long)tap * p.in_c;
    for (int c0 = 0; c0 < p.in_c; c0 += 16) {
      float a4[4] = {0.f, 0.f, 0.f, 0.f}, w4[4] = {0.f, 0.f, 0.f, 0.f};
      if (inb) {
#pragma unroll
        for (int j = 0; j < 4; ++j) a4[j] = Cvt<T>::to_f(arow[c0 + lk + j]);
      }
      if (loc < p.out_c) {
#pragma unroll
        for (int j = 0; j < 4; ++j) w4[j] = Cvt<T>::to_f(wrow[c0 + lk + j]);
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        As[lk + j][lrow] = a4[j];
        Ws[lk + j][lrow] = w4[j];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float a = As[k][px];
        const float4* wv = reinterpret_cast<const float4*>(&Ws[k][cg * 16]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 ww = wv[j];
          acc[4 * j + 0] = fmaf(a, ww.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(a, ww.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(a, ww.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(a, ww.w, acc[4 * j + 3]);
        }
      }
    }
  }

  const long long m = m0 + px;
  const int c0 = n0 + cg * 16;
  if (m < p.M && c0 < p.out_c) {
    const int b = (int)(m / out_hw);
    const int r = (int)(m - (long long)b * out_hw);
    const int ho = r / p.out_w;
    const int wo = r - ho * p.out_w;
    epi_store16<true>(p.epi, b, ho, wo, c0, acc);
  }
}

int conv_simt_launch(const yx_conv_desc* d, cudaStream_t stream) {
  int rc = validate_conv_geometry(d);
  if (rc) return rc;
  ConvSimtParams p;
  memset(&p, 0, sizeof(p));
  rc = fill_epi_params(d, &p.epi);
  if (rc) return rc;
  p.batch = d->batch; p.in_h = d->in_h; p.in_w = d->in_w; p.in_c = d->in_c;
  p.out_h = d->out_h; p.out_w = d->out_w; p.out_c = d->out_c;
  p.ksize = d->ksize; p.stride = d->stride; p.pad = (d->ksize - 1) / 2;
  p.M = (long long)d->batch * d->out_h * d->out_w;
  p.in = d->in; p.in_ld = d->in_ld; p.w = d->w;
  dim3 grid((unsigned)ceil_div64(p.M, 64), (unsigned)ceil_div64(d->out_c, 64));
  switch (d->dtype) {
    case YX_BF16: conv_simt_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(p); break;
    case YX_FP16: conv_simt_kernel<__half><<<grid, 256, 0, stream>>>(p); break;
    case YX_FP32: conv_simt_kernel<float><<<grid, 256, 0, stream>>>(p); break;
    default: YX_REQUIRE(false, YX_ERR_INVALID_ARG, "conv_simt: bad dtype %d", d->dtype);
  }
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

}  // namespace yx
