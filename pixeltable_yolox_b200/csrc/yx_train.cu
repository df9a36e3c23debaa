// Training-side kernels next to SimOTA (yx_simota.cu) and the fused head losses (yx_losses.cu):
//
//  1. head_train_decode (+ backward): the training branch of YoloxHead.forward for one level,
//     yolox/models/yolo_head.py:161-201 and get_output_and_grid (:213-231): cat[reg, obj, cls] -> [B, HW, 5+nc] rows,
//     xy = (xy + grid) * stride, wh = exp(wh) * stride, raw logits for obj / cls, plus the raw regression rows
//     (`origin_preds`, :190-200) when the L1 term is on. The reference runs ~12 elementwise / permute / cat kernels per
//     level (and as many in backward); here one transposing pass each way: NCHW conv outputs are read along W
//     (coalesced), staged in shared memory and written as whole [5+nc] rows (coalesced), fp32 out.
//  2. sgd_ema: the reference's optimizer step and EMA update as ONE multi-tensor launch over every parameter and buffer:
//     torch.optim.SGD(momentum, nesterov=True, per-group weight decay) as configured in yolox/config.py:307-333, and
//     ModelEMA.update (yolox/utils/ema.py:46-58): ema = d * ema + (1 - d) * value for every floating-point state tensor.
#include <string.h>

#include "yx_common.cuh"

namespace yx {

// ------------------------------------------------------------------------------------------
// 1. training-branch head rows
// ------------------------------------------------------------------------------------------
static constexpr int kHtPix = 32;       // pixels (rows of the output) per CTA
static constexpr int kHtThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kHtThreads)
head_train_decode_kernel(const T* __restrict__ reg, const T* __restrict__ obj, const T* __restrict__ cls, int nc, int hw, int w,
                         float stride, int anchors, int anchor_off, float* __restrict__ out, float* __restrict__ origin) {
  extern __shared__ float ht_tile[];                 // [5+nc][kHtPix + 1]
  const int nch = 5 + nc;
  const int b = blockIdx.y, pix0 = blockIdx.x * kHtPix;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pix = pix0 + lane;
  const bool ok = pix < hw;
  const float gx = (float)(pix % w), gy = (float)(pix / w);
  for (int ch = warp; ch < nch; ch += kHtThreads / 32) {
    const T* src = ch < 4 ? reg + ((long long)b * 4 + ch) * hw : (ch == 4 ? obj + (long long)b * hw : cls + ((long long)b * nc + (ch - 5)) * hw);
    float v = ok ? Cvt<T>::to_f(src[pix]) : 0.0f;
    if (ch < 4 && origin && ok) origin[((long long)b * anchors + anchor_off + pix) * 4 + ch] = v;
    if (ch == 0) v = (v + gx) * stride;
    else if (ch == 1) v = (v + gy) * stride;
    else if (ch < 4) v = expf(v) * stride;
    ht_tile[ch * (kHtPix + 1) + lane] = v;
  }
  __syncthreads();
  const int rows = min(kHtPix, hw - pix0);
  float* dst = out + ((long long)b * anchors + anchor_off + pix0) * nch;
  for (int i = threadIdx.x; i < rows * nch; i += kHtThreads) {
    const int r = i / nch, c = i - r * nch;
    dst[i] = ht_tile[c * (kHtPix + 1) + r];
  }
}

template <typename T>
__global__ void __launch_bounds__(kHtThreads)
head_train_decode_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ out, const float* __restrict__ gorigin,
                             int nc, int hw, float stride, int anchors, int anchor_off, T* __restrict__ greg,
                             T* __restrict__ gobj, T* __restrict__ gcls) {
  extern __shared__ float ht_tile[];
  const int nch = 5 + nc;
  const int b = blockIdx.y, pix0 = blockIdx.x * kHtPix;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows = min(kHtPix, hw - pix0);
  const long long row0 = (long long)b * anchors + anchor_off + pix0;
  for (int i = threadIdx.x; i < rows * nch; i += kHtThreads) {
    const int r = i / nch, c = i - r * nch;
    float g = gout[row0 * nch + i];
    if (c < 2) g *= stride;                               // d (v + grid) * s / dv
    else if (c < 4) g *= out[row0 * nch + i];             // d exp(v) * s / dv = the decoded value itself
    if (c < 4 && gorigin) g += gorigin[(row0 + r) * 4 + c];
    ht_tile[c * (kHtPix + 1) + r] = g;
  }
  __syncthreads();
  const int pix = pix0 + lane;
  if (pix >= hw) return;
  for (int ch = warp; ch < nch; ch += kHtThreads / 32) {
    T* dst = ch < 4 ? greg + ((long long)b * 4 + ch) * hw : (ch == 4 ? gobj + (long long)b * hw : gcls + ((long long)b * nc + (ch - 5)) * hw);
    dst[pix] = Cvt<T>::from_f(ht_tile[ch * (kHtPix + 1) + lane]);
  }
}

static int head_train_check(int dtype, int batch, int nc, int h, int w, int anchors, int anchor_off) {
  YX_REQUIRE(dtype == YX_FP32 || dtype == YX_BF16 || dtype == YX_FP16, YX_ERR_INVALID_ARG, "head_train_decode: dtype %d", dtype);
  YX_REQUIRE(batch > 0 && nc > 0 && h > 0 && w > 0, YX_ERR_INVALID_ARG, "head_train_decode: bad sizes");
  YX_REQUIRE(anchor_off >= 0 && anchor_off + h * w <= anchors, YX_ERR_INVALID_ARG, "head_train_decode: level [%d, %d) outside %d anchors",
             anchor_off, anchor_off + h * w, anchors);
  YX_REQUIRE((size_t)(5 + nc) * (kHtPix + 1) * 4 <= 48 * 1024, YX_ERR_UNSUPPORTED, "head_train_decode: %d classes exceed the row staging", nc);
  YX_REQUIRE(batch <= 65535, YX_ERR_UNSUPPORTED, "head_train_decode: batch %d", batch);
  return YX_OK;
}

int head_train_decode_launch(const void* reg, const void* obj, const void* cls, int dtype, int batch, int nc, int h, int w,
                             float stride, int anchors, int anchor_off, float* out, float* origin, cudaStream_t s) {
  YX_REQUIRE(reg && obj && cls && out, YX_ERR_INVALID_ARG, "head_train_decode: null pointer");
  int rc = head_train_check(dtype, batch, nc, h, w, anchors, anchor_off);
  if (rc) return rc;
  const int hw = h * w;
  const dim3 grid((unsigned)((hw + kHtPix - 1) / kHtPix), (unsigned)batch);
  const size_t smem = (size_t)(5 + nc) * (kHtPix + 1) * 4;
#define YX_GO(T) head_train_decode_kernel<T><<<grid, kHtThreads, smem, s>>>((const T*)reg, (const T*)obj, (const T*)cls, nc, hw, w, stride, anchors, anchor_off, out, origin)
  if (dtype == YX_FP32) YX_GO(float); else if (dtype == YX_BF16) YX_GO(__nv_bfloat16); else YX_GO(__half);
#undef YX_GO
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

int head_train_decode_bwd_launch(const float* gout, const float* out, const float* gorigin, int dtype, int batch, int nc, int h,
                                 int w, float stride, int anchors, int anchor_off, void* greg, void* gobj, void* gcls,
                                 cudaStream_t s) {
  YX_REQUIRE(gout && out && greg && gobj && gcls, YX_ERR_INVALID_ARG, "head_train_decode_bwd: null pointer");
  int rc = head_train_check(dtype, batch, nc, h, w, anchors, anchor_off);
  if (rc) return rc;
  const int hw = h * w;
  const dim3 grid((unsigned)((hw + kHtPix - 1) / kHtPix), (unsigned)batch);
  const size_t smem = (size_t)(5 + nc) * (kHtPix + 1) * 4;
#define YX_GO(T) head_train_decode_bwd_kernel<T><<<grid, kHtThreads, smem, s>>>(gout, out, gorigin, nc, hw, stride, anchors, anchor_off, (T*)greg, (T*)gobj, (T*)gcls)
  if (dtype == YX_FP32) YX_GO(float); else if (dtype == YX_BF16) YX_GO(__nv_bfloat16); else YX_GO(__half);
#undef YX_GO
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

// ------------------------------------------------------------------------------------------
// 2. SGD (momentum, nesterov, weight decay) + EMA, every tensor in one launch
// ------------------------------------------------------------------------------------------
// table row (int64 x 6): param ptr | grad ptr (0: no optimizer step, EMA only -- BN running statistics) | momentum buffer
// ptr | ema ptr (0: none) | numel | weight-decay bits (float in the low 32 bits). chunks: (tensor index, first element).
static constexpr int kSgdThreads = 256;

__global__ void __launch_bounds__(kSgdThreads)
sgd_ema_kernel(const long long* __restrict__ table, const int* __restrict__ chunks, int chunk_elems, float lr, float momentum,
               int nesterov, int first_step, float ema_decay, float ema_rest, const float* __restrict__ hyper) {
  // hyper (device, may be null): {lr, ema_decay, 1 - ema_decay} read at run time, so that a CUDA graph that captured this
  // launch follows the learning-rate schedule and the EMA ramp without being re-captured
  if (hyper) { lr = hyper[0]; ema_decay = hyper[1]; ema_rest = hyper[2]; }
  const int t = chunks[2 * blockIdx.x], e0 = chunks[2 * blockIdx.x + 1];
  const long long* row = table + (long long)t * 6;
  float* p = reinterpret_cast<float*>(row[0]);
  const float* g = reinterpret_cast<const float*>(row[1]);
  float* buf = reinterpret_cast<float*>(row[2]);
  float* ema = reinterpret_cast<float*>(row[3]);
  const long long n = row[4];
  const float wd = __int_as_float((int)(row[5] & 0xffffffffll));
  const long long e1 = min(n, (long long)e0 + chunk_elems);
  for (long long i = e0 + threadIdx.x; i < e1; i += kSgdThreads) {
    float v = p[i];
    if (g) {
      // torch.optim.SGD._single_tensor_sgd: d = g + wd * p; buf = first ? d : momentum * buf + d;
      //                                     d = nesterov ? d + momentum * buf : buf; p -= lr * d
      float d = g[i];
      if (wd != 0.0f) d = fmaf(wd, v, d);
      if (momentum != 0.0f) {
        const float bv = first_step ? d : __fadd_rn(__fmul_rn(momentum, buf[i]), d);   // buf.mul_(momentum).add_(d): two roundings
        buf[i] = bv;
        d = nesterov ? fmaf(momentum, bv, d) : bv;
      }
      v = fmaf(-lr, d, v);
      p[i] = v;
    }
    // v *= d; v += (1 - d) * msd[k]  (ema.py:56-58): three roundings, (1 - d) evaluated in double on the host
    if (ema) ema[i] = __fadd_rn(__fmul_rn(ema[i], ema_decay), __fmul_rn(ema_rest, v));
  }
}

int sgd_ema_launch(const long long* table, const int* chunks, int n_chunks, int chunk_elems, float lr, float momentum,
                   int nesterov, int first_step, float ema_decay, float ema_rest, const float* hyper, cudaStream_t s) {
  YX_REQUIRE(table && chunks, YX_ERR_INVALID_ARG, "sgd_ema: null table");
  YX_REQUIRE(n_chunks >= 0 && chunk_elems > 0, YX_ERR_INVALID_ARG, "sgd_ema: bad chunking");
  if (n_chunks == 0) return YX_OK;
  sgd_ema_kernel<<<n_chunks, kSgdThreads, 0, s>>>(table, chunks, chunk_elems, lr, momentum, nesterov, first_step, ema_decay, ema_rest, hyper);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

}  // namespace yx
