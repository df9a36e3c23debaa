// YOLOX head losses, forward and gradient in one pass over the prediction tensor.
//   reference: YoloxHead.get_losses (yolox/models/yolo_head.py:354-418): target construction from the
//   SimOTA assignment (:354-366), IOUloss (yolox/models/losses.py:7-51, loss_type "iou" / "giou"),
//   nn.BCEWithLogitsLoss(reduction="none") on the objectness of EVERY anchor and on the classes of the
//   foreground anchors (class target = one_hot(matched class) * matched IoU), nn.L1Loss on the raw
//   regression outputs against get_l1_target (:412-418) when use_l1.
// The reference gathers the foreground rows with boolean masks (one device->host sync per step for the
// nonzero count), builds a [num_fg, nc] one-hot target tensor and runs ~25 elementwise kernels plus their
// autograd mirrors. Here one CTA stages 128 consecutive anchor rows ([5+nc] floats each) in shared memory with
// one bulk copy, every thread owns one anchor, and the SAME kernel writes d(loss)/d(pred) for that anchor
// into the staged row, which then leaves with coalesced 128-bit stores: the tensor is read once and the
// gradient written once (HBM bound: 2 * B*A*(5+nc)*4 bytes).
//   sums[0..3] (fp64) = sum of the iou / obj / cls / l1 loss terms (un-normalised: the caller divides by
//   max(num_fg, 1) and applies reg_weight = 5, yolo_head.py:382-404);
//   grad[b, a, :] = d(5*iou_sum + obj_sum + cls_sum)/d pred[b, a, :], grad_origin[b, a, :4] = d(l1_sum)/d origin.
#include <math.h>
#include <string.h>

#include "yx_common.cuh"

namespace yx {

static constexpr int kLossAnchors = 128;

__device__ __forceinline__ float bce_with_logits(float x, float t, float& dx) {
  // torch: (1 - t) * x + max(-x, 0) + log(exp(-max(-x,0)) + exp(-x - max(-x,0)))  ==  max(x,0) - x*t + log1p(exp(-|x|))
  const float e = expf(-fabsf(x));
  const float sig = x >= 0.0f ? 1.0f / (1.0f + e) : e / (1.0f + e);
  dx = sig - t;
  return fmaxf(x, 0.0f) - x * t + log1pf(e);
}

// d max(a, b) / d a as torch.maximum's backward defines it (ties split evenly)
__device__ __forceinline__ float dmax_a(float a, float b) { return a > b ? 1.0f : (a == b ? 0.5f : 0.0f); }
__device__ __forceinline__ float dmin_a(float a, float b) { return a < b ? 1.0f : (a == b ? 0.5f : 0.0f); }

__global__ void __launch_bounds__(kLossAnchors)
head_loss_kernel(const float* __restrict__ pred, const float* __restrict__ labels, int max_gt,
                 const unsigned char* __restrict__ fg_mask, const int* __restrict__ matched_gt,
                 const float* __restrict__ matched_iou, const int* __restrict__ matched_cls,
                 const float* __restrict__ origin, const float* __restrict__ x_shift, const float* __restrict__ y_shift,
                 const float* __restrict__ stride, int batch, int anchors, int nc, int giou, float reg_weight,
                 double* __restrict__ sums, float* __restrict__ grad, float* __restrict__ grad_origin) {
  extern __shared__ __align__(128) float lrows[];         // [kLossAnchors][5+nc]
  __shared__ uint64_t lbar;
  __shared__ double red[4][kLossAnchors / 32];
  const int nch = 5 + nc;
  const int chunks = (anchors + kLossAnchors - 1) / kLossAnchors;
  const int b = blockIdx.x / chunks;
  const int a0 = (blockIdx.x - b * chunks) * kLossAnchors;
  const int na = min(kLossAnchors, anchors - a0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row0 = (long long)b * anchors + a0;
  const float* g = pred + row0 * nch;
  const int total = na * nch;
  const bool aligned = (reinterpret_cast<uintptr_t>(g) & 15) == 0 && (total & 3) == 0;
  if (tid == 0) { mbar_init(&lbar, 1); fence_barrier_init(); }
  __syncthreads();
  if (aligned) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&lbar, (uint32_t)total * 4u);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(lrows)),
                   "l"(g), "r"((uint32_t)total * 4u), "r"(smem_u32(&lbar))
                   : "memory");
    }
    mbar_wait(&lbar, 0);
  } else {
    for (int i = tid; i < total; i += kLossAnchors) lrows[i] = g[i];
  }
  __syncthreads();

  double s_iou = 0.0, s_obj = 0.0, s_cls = 0.0, s_l1 = 0.0;
  if (tid < na) {
    float* row = lrows + tid * nch;       // row pitch 5+nc floats: odd for nc = 80, the strided accesses are conflict free
    const long long ra = row0 + tid;
    const bool fg = fg_mask[ra] != 0;
    // objectness: every anchor, target = fg (yolo_head.py:361, 385-387)
    float dobj;
    s_obj = bce_with_logits(row[4], fg ? 1.0f : 0.0f, dobj);
    float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
    if (fg) {
      const float* gt = labels + ((long long)b * max_gt + matched_gt[ra]) * 5;
      const float tcx = gt[1], tcy = gt[2], tw = gt[3], th = gt[4];
      const float pcx = row[0], pcy = row[1], pw = row[2], ph = row[3];
      // IOUloss.forward (losses.py:14-48)
      const float plx = pcx - pw / 2, phx = pcx + pw / 2, ply = pcy - ph / 2, phy = pcy + ph / 2;
      const float tlx_ = tcx - tw / 2, thx = tcx + tw / 2, tly_ = tcy - th / 2, thy = tcy + th / 2;
      const float tlx = fmaxf(plx, tlx_), tly = fmaxf(ply, tly_), brx = fminf(phx, thx), bry = fminf(phy, thy);
      const float area_p = pw * ph, area_g = tw * th;
      const float en = (tlx < brx && tly < bry) ? 1.0f : 0.0f;
      const float iw = brx - tlx, ih = bry - tly;
      const float area_i = iw * ih * en;
      const float area_u = area_p + area_g - area_i;
      const float U = area_u + 1e-16f;
      const float iou = area_i / U;
      // partial derivatives of area_i and area_p w.r.t. (cx, cy, w, h)
      const float a_tlx = dmax_a(plx, tlx_), a_brx = dmin_a(phx, thx), a_tly = dmax_a(ply, tly_), a_bry = dmin_a(phy, thy);
      // d iw / d cx = a_brx - a_tlx ; d iw / d w = (a_brx + a_tlx) / 2   (plx = cx - w/2, phx = cx + w/2)
      const float diw_cx = a_brx - a_tlx, diw_w = 0.5f * (a_brx + a_tlx);
      const float dih_cy = a_bry - a_tly, dih_h = 0.5f * (a_bry + a_tly);
      const float dI[4] = {en * ih * diw_cx, en * iw * dih_cy, en * ih * diw_w, en * iw * dih_h};
      const float dP[4] = {0.0f, 0.0f, ph, pw};
      float diou[4], dU[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dU[k] = dP[k] - dI[k];
        diou[k] = (dI[k] * U - area_i * dU[k]) / (U * U);
      }
      float loss, dl[4];
      if (!giou) {
        loss = 1.0f - iou * iou;
#pragma unroll
        for (int k = 0; k < 4; ++k) dl[k] = -2.0f * iou * diou[k];
      } else {
        // giou = iou - (area_c - area_u) / clamp(area_c, 1e-16); loss = 1 - clamp(giou, -1, 1)   (losses.py:37-45)
        const float clx = fminf(plx, tlx_), cly = fminf(ply, tly_), chx = fmaxf(phx, thx), chy = fmaxf(phy, thy);
        const float cw = chx - clx, chh = chy - cly;
        const float area_c = cw * chh;
        const float C = fmaxf(area_c, 1e-16f);
        const float gi = iou - (area_c - area_u) / C;
        loss = 1.0f - fminf(fmaxf(gi, -1.0f), 1.0f);
        const float c_lx = dmin_a(plx, tlx_), c_hx = dmax_a(phx, thx), c_ly = dmin_a(ply, tly_), c_hy = dmax_a(phy, thy);
        const float dcw_cx = c_hx - c_lx, dcw_w = 0.5f * (c_hx + c_lx), dch_cy = c_hy - c_ly, dch_h = 0.5f * (c_hy + c_ly);
        const float dC[4] = {chh * dcw_cx, cw * dch_cy, chh * dcw_w, cw * dch_h};
        const bool cl = area_c >= 1e-16f;          // clamp(min) passes the gradient only above the bound
        const bool inside = gi >= -1.0f && gi <= 1.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // d[(c - u) / C] = ((dc - du) * C - (c - u) * dCk) / C^2, dCk = dc when not clamped else 0
          const float dCk = cl ? dC[k] : 0.0f;
          const float dterm = ((dC[k] - dU[k]) * C - (area_c - area_u) * dCk) / (C * C);
          dl[k] = inside ? -(diou[k] - dterm) : 0.0f;
        }
      }
      s_iou = loss;
      d0 = reg_weight * dl[0]; d1 = reg_weight * dl[1]; d2 = reg_weight * dl[2]; d3 = reg_weight * dl[3];
      // classes: target = one_hot(matched class) * matched IoU (yolo_head.py:354-357, 388-392)
      const int mc = matched_cls[ra];
      const float miou = matched_iou[ra];
      float acc = 0.0f;
      for (int c = 0; c < nc; ++c) {
        float dx;
        acc += bce_with_logits(row[5 + c], c == mc ? miou : 0.0f, dx);
        row[5 + c] = dx;
      }
      s_cls = acc;
      if (origin) {
        // L1 on the raw regression outputs (yolo_head.py:393-397, 412-418)
        const float st = stride[a0 + tid];
        const float t[4] = {tcx / st - x_shift[a0 + tid], tcy / st - y_shift[a0 + tid], logf(tw / st + 1e-8f), logf(th / st + 1e-8f)};
        const float* o = origin + ra * 4;
        float go[4], l1 = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float df = o[k] - t[k];
          l1 += fabsf(df);
          go[k] = df > 0.0f ? 1.0f : (df < 0.0f ? -1.0f : 0.0f);
        }
        s_l1 = l1;
        *reinterpret_cast<float4*>(grad_origin + ra * 4) = make_float4(go[0], go[1], go[2], go[3]);
      }
    } else {
      for (int c = 0; c < nc; ++c) row[5 + c] = 0.0f;
      if (origin) *reinterpret_cast<float4*>(grad_origin + ra * 4) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    row[0] = d0; row[1] = d1; row[2] = d2; row[3] = d3; row[4] = dobj;
  }
  // block reduction of the four sums (fp64: the order of the additions must not matter at 1e-6)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_iou += __shfl_xor_sync(0xffffffffu, s_iou, o);
    s_obj += __shfl_xor_sync(0xffffffffu, s_obj, o);
    s_cls += __shfl_xor_sync(0xffffffffu, s_cls, o);
    s_l1 += __shfl_xor_sync(0xffffffffu, s_l1, o);
  }
  if (lane == 0) { red[0][warp] = s_iou; red[1][warp] = s_obj; red[2][warp] = s_cls; red[3][warp] = s_l1; }
  __syncthreads();
  if (tid < 4) {
    double t = 0.0;
    for (int w = 0; w < kLossAnchors / 32; ++w) t += red[tid][w];
    if (t != 0.0) atomicAdd(&sums[tid], t);
  }
  // the staged rows now hold the gradient: coalesced copy out
  float* go = grad + row0 * nch;
  if (aligned && (reinterpret_cast<uintptr_t>(go) & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(lrows);
    float4* g4 = reinterpret_cast<float4*>(go);
    for (int i = tid; i < total / 4; i += kLossAnchors) g4[i] = s4[i];
  } else {
    for (int i = tid; i < total; i += kLossAnchors) go[i] = lrows[i];
  }
}

int head_loss_launch(const float* pred, const float* labels, int max_gt, const unsigned char* fg_mask, const int* matched_gt,
                     const float* matched_iou, const int* matched_cls, const float* origin, const float* x_shift,
                     const float* y_shift, const float* stride, int batch, int anchors, int nc, int giou, float reg_weight,
                     double* sums, float* grad, float* grad_origin, cudaStream_t s) {
  YX_REQUIRE(pred && labels && fg_mask && matched_gt && matched_iou && matched_cls && sums && grad, YX_ERR_INVALID_ARG,
             "head_losses: null pointer");
  YX_REQUIRE(batch > 0 && anchors > 0 && nc > 0 && max_gt > 0, YX_ERR_INVALID_ARG, "head_losses: bad sizes");
  YX_REQUIRE(!origin || (x_shift && y_shift && stride && grad_origin), YX_ERR_INVALID_ARG,
             "head_losses: the L1 term needs x_shift, y_shift, stride and grad_origin");
  YX_REQUIRE(!origin || (((uintptr_t)origin | (uintptr_t)grad_origin) & 15) == 0, YX_ERR_INVALID_ARG,
             "head_losses: origin / grad_origin must be 16-byte aligned");
  const size_t smem = (size_t)kLossAnchors * (5 + nc) * sizeof(float);
  YX_REQUIRE(smem <= 200 * 1024, YX_ERR_UNSUPPORTED, "head_losses: %d classes exceed the shared-memory row staging", nc);
  static size_t configured_dev[kMaxDevices] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (configured == 0) configured = 48 * 1024;
  if (smem > configured) {
    YX_CUDA(cudaFuncSetAttribute(head_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  YX_CUDA(cudaMemsetAsync(sums, 0, 4 * sizeof(double), s));
  const int chunks = (anchors + kLossAnchors - 1) / kLossAnchors;
  head_loss_kernel<<<batch * chunks, kLossAnchors, smem, s>>>(pred, labels, max_gt, fg_mask, matched_gt, matched_iou,
                                                              matched_cls, origin, x_shift, y_shift, stride, batch, anchors,
                                                              nc, giou, reg_weight, sums, grad, grad_origin);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

}  // namespace yx
