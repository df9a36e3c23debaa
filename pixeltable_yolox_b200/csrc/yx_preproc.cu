// Either side of the detection path, byte / integer work, bit-exact:
//
//  1. letterbox_u8: the reference's test-time preprocessing (`preproc`, yolox/data/data_augment.py:140-156, called by
//     ValTransform :234-241 and YoloxProcessor.__call__, yolox/models/processor.py:30-37) on the device, straight from the
//     decoded HWC uint8 image: r = min(H/h, W/w); cv2.resize(img, (int(w*r), int(h*r)), INTER_LINEAR) into the top-left corner
//     of a 114-grey canvas; HWC -> CHW; uint8 or float32 0..255 out. Only the raw image bytes cross PCIe; the resize is
//     OpenCV's 8-bit fixed-point bilinear restated operation by operation (third-party: opencv-python >= 4.10, installed
//     4.13; oracle/preproc_oracle.py pins it against cv2 itself):
//        fx = (float)((dx + 0.5) * (double)w / nw - 0.5); sx = floor(fx); fx -= sx; clamp (sx < 0 -> 0, fx = 0;
//        sx >= w - 1 -> w - 1, fx = 0); a1 = rint(fx * 2048), a0 = rint((1 - fx) * 2048)        (round half to even)
//        same for y WITHOUT the clamp (row indices are clipped to [0, h) instead, both taps may hit the same row);
//        row_k[dx] = S[sy_k][sx] * a0 + S[sy_k][sx + 1] * a1                                   (int32)
//        out = (((b0 * (row_0 >> 4)) >> 16) + ((b1 * (row_1 >> 4)) >> 16) + 2) >> 2
//     and the exact 2:1 case, which OpenCV routes to INTER_AREA: (a + b + c + d + 2) >> 2.
//  2. coco_rows: the device half of CocoEvaluator.convert_to_coco_format (yolox/evaluators/coco_evaluator.py:205-251):
//     per kept detection box / scale (fp32 divide), xyxy -> xywh, score = obj * class_conf, category = class_ids[cls],
//     compacted over the batch in image order so that ONE device->host copy replaces the per-image .cpu() and the
//     per-row .item() calls.
#include <string.h>

#include "yx_common.cuh"

namespace yx {

struct LetterboxImage {          // one per image, device array
  const unsigned char* src;      // HWC uint8, `channels` interleaved
  int h, w;                      // source size
  long long pitch;               // bytes per source row
};

__device__ __forceinline__ void lb_coeff(int d, int src, int dst, bool clamp, int& s, int& a0, int& a1) {
  const double scale = (double)src / (double)dst;
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  s = (int)floorf(f);
  f -= (float)s;
  if (clamp) {
    if (s < 0) { f = 0.0f; s = 0; }
    if (s >= src - 1) { f = 0.0f; s = src - 1; }
  }
  a1 = __float2int_rn(__fmul_rn(f, 2048.0f));
  a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
}

template <typename TO>
__global__ void __launch_bounds__(256)
letterbox_kernel(const LetterboxImage* __restrict__ imgs, int channels, int H, int W, TO* __restrict__ out) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const LetterboxImage im = imgs[b];
  // r = min(H / h, W / w) in double, like Python's float division (data_augment.py:146-149)
  const double rh = (double)H / (double)im.h, rw = (double)W / (double)im.w;
  const double r = rh < rw ? rh : rw;
  const int nh = (int)((double)im.h * r), nw = (int)((double)im.w * r);
  TO* o = out + ((long long)b * channels * H + y) * W + x;
  const long long plane = (long long)H * W;
  if (y >= nh || x >= nw) {
    for (int c = 0; c < channels; ++c) o[c * plane] = (TO)114;
    return;
  }
  if (im.w == 2 * nw && im.h == 2 * nh) {                // exact 2:1 -> OpenCV's INTER_AREA fast path
    const unsigned char* p0 = im.src + (long long)(2 * y) * im.pitch + (long long)(2 * x) * channels;
    const unsigned char* p1 = p0 + im.pitch;
    for (int c = 0; c < channels; ++c)
      o[c * plane] = (TO)((p0[c] + p0[channels + c] + p1[c] + p1[channels + c] + 2) >> 2);
    return;
  }
  int sx, a0, a1, sy, b0, b1;
  lb_coeff(x, im.w, nw, true, sx, a0, a1);
  lb_coeff(y, im.h, nh, false, sy, b0, b1);
  const int sx1 = min(sx + 1, im.w - 1);
  const int y0 = min(max(sy, 0), im.h - 1), y1 = min(max(sy + 1, 0), im.h - 1);
  const unsigned char* r0 = im.src + (long long)y0 * im.pitch;
  const unsigned char* r1 = im.src + (long long)y1 * im.pitch;
  for (int c = 0; c < channels; ++c) {
    const int s0 = (int)r0[sx * channels + c] * a0 + (int)r0[sx1 * channels + c] * a1;
    const int s1 = (int)r1[sx * channels + c] * a0 + (int)r1[sx1 * channels + c] * a1;
    int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
    o[c * plane] = (TO)v;
  }
}

int letterbox_launch(const void* images_dev, int batch, int channels, int H, int W, void* out, int out_dtype, cudaStream_t s) {
  YX_REQUIRE(images_dev && out, YX_ERR_INVALID_ARG, "letterbox: null pointer");
  YX_REQUIRE(batch > 0 && batch <= 65535 && H > 0 && W > 0 && channels >= 1 && channels <= 4, YX_ERR_INVALID_ARG, "letterbox: bad sizes");
  YX_REQUIRE(out_dtype == YX_U8 || out_dtype == YX_FP32, YX_ERR_INVALID_ARG, "letterbox: output dtype must be uint8 or float32");
  const dim3 grid((unsigned)((W + 31) / 32), (unsigned)((H + 7) / 8), (unsigned)batch);
  const LetterboxImage* imgs = reinterpret_cast<const LetterboxImage*>(images_dev);
  if (out_dtype == YX_U8) letterbox_kernel<unsigned char><<<grid, 256, 0, s>>>(imgs, channels, H, W, (unsigned char*)out);
  else letterbox_kernel<float><<<grid, 256, 0, s>>>(imgs, channels, H, W, (float*)out);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

// ------------------------------------------------------------------------------------------
// evaluator result rows
// ------------------------------------------------------------------------------------------
// one CTA per image: its first min(det_count, max_det) rows -> flat position prefix[b] + k
__global__ void __launch_bounds__(256)
coco_rows_kernel(const float* __restrict__ dets, const int* __restrict__ det_count, int batch, int max_det,
                 const float* __restrict__ scale, const long long* __restrict__ image_ids, const int* __restrict__ class_ids,
                 int n_class_ids, float* __restrict__ bbox, float* __restrict__ score, int* __restrict__ category,
                 long long* __restrict__ image_id, int* __restrict__ total) {
  __shared__ int s_base;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int base = 0;
    for (int i = 0; i < b; ++i) base += min(det_count[i], max_det);
    s_base = base;
    if (b == batch - 1) *total = base + min(det_count[b], max_det);
  }
  __syncthreads();
  const int n = min(det_count[b], max_det);
  const float sc = scale[b];
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const float* d = dets + ((long long)b * max_det + k) * 7;
    const float x1 = __fdiv_rn(d[0], sc), y1 = __fdiv_rn(d[1], sc), x2 = __fdiv_rn(d[2], sc), y2 = __fdiv_rn(d[3], sc);
    const long long o = s_base + k;
    bbox[o * 4 + 0] = x1; bbox[o * 4 + 1] = y1;
    bbox[o * 4 + 2] = __fsub_rn(x2, x1); bbox[o * 4 + 3] = __fsub_rn(y2, y1);       // xyxy2xywh (boxes.py:123-126)
    score[o] = __fmul_rn(d[4], d[5]);
    const int c = (int)d[6];
    category[o] = (class_ids && c >= 0 && c < n_class_ids) ? class_ids[c] : c;
    image_id[o] = image_ids[b];
  }
}

int coco_rows_launch(const float* dets, const int* det_count, int batch, int max_det, const float* scale,
                     const long long* image_ids, const int* class_ids, int n_class_ids, float* bbox, float* score,
                     int* category, long long* image_id, int* total, cudaStream_t s) {
  YX_REQUIRE(dets && det_count && scale && image_ids && bbox && score && category && image_id && total, YX_ERR_INVALID_ARG,
             "coco_rows: null pointer");
  YX_REQUIRE(batch > 0 && max_det > 0, YX_ERR_INVALID_ARG, "coco_rows: bad sizes");
  coco_rows_kernel<<<batch, 256, 0, s>>>(dets, det_count, batch, max_det, scale, image_ids, class_ids, n_class_ids, bbox, score,
                                         category, image_id, total);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

}  // namespace yx
