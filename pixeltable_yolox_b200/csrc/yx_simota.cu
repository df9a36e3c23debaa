// SimOTA label assignment, batched over images, one CTA per image, no host synchronisation.
//   reference: YoloxHead.get_assignments / get_geometry_constraint / simota_matching
//   (yolox/models/yolo_head.py:420-574) and bboxes_iou (yolox/utils/boxes.py:78-101).
// The reference loops over images in Python with >= G+3 device->host syncs per image and
// materialises a [G, A', 80] BCE tensor; here the class cost of a (gt, anchor) pair is
//   S_a - (-log(1-p_a,c)) + (-log p_a,c),   S_a = sum_c -log(1-p_a,c),  p = sqrt(sig(cls)*sig(obj))
// so only a [G, A'] cost/IoU pair of matrices is ever written (A' <= 27*G in-centre anchors).
//
// simota_matching (yolo_head.py:542-574) is reproduced bit-exactly on a given cost/IoU matrix:
//   dynamic_k = max(1, int(sum of the top-10 IoUs))  -- summed in the order torch's CPU
//               inner-reduction uses for a row of 10 floats (elements 8,9 first, then 0..7);
//   per GT the dynamic_k smallest costs (ties: lower index first);
//   anchors chosen by several GTs go to argmin_g cost[g][a] over ALL g (first index on ties).
#include <string.h>
#include <math.h>

#include "yx_common.cuh"

namespace yx {

static constexpr int kSimThreads = 512;
static constexpr int kSimWarps = kSimThreads / 32;

// Select, in (value, index) lexicographic order, the next element of row[0..n) after `prev`.
// DESC: order is (value descending, index ascending); else (value ascending, index ascending).
template <bool DESC>
__device__ __forceinline__ void warp_next(const float* __restrict__ row, int n, float prev_v, int prev_i,
                                          float& out_v, int& out_i) {
  const int lane = threadIdx.x & 31;
  float bv = DESC ? -INFINITY : INFINITY;
  int bi = 0x7fffffff;
  for (int i = lane; i < n; i += 32) {
    const float v = row[i];
    const bool elig = DESC ? ((v < prev_v) || (v == prev_v && i > prev_i))
                           : ((v > prev_v) || (v == prev_v && i > prev_i));
    const bool better = DESC ? ((v > bv) || (v == bv && i < bi)) : ((v < bv) || (v == bv && i < bi));
    if (elig && (bi == 0x7fffffff || better)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    bool take;
    if (oi == 0x7fffffff) take = false;
    else if (bi == 0x7fffffff) take = true;
    else take = DESC ? ((ov > bv) || (ov == bv && oi < bi)) : ((ov < bv) || (ov == bv && oi < bi));
    if (take) { bv = ov; bi = oi; }
  }
  out_v = bv; out_i = bi;
}

// cnt: shared int[n] scratch. match_gt/match_iou: [n] outputs. Whole CTA participates.
__device__ void simota_matching_cta(const float* __restrict__ cost, const float* __restrict__ ious, int G, int n,
                                    long long ld, int* __restrict__ match_gt, float* __restrict__ match_iou,
                                    int* __restrict__ num_fg, int* cnt) {
  __shared__ int s_fg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  for (int i = tid; i < n; i += blockDim.x) { cnt[i] = 0; match_gt[i] = -1; }
  if (tid == 0) s_fg = 0;
  __syncthreads();
  const int k_top = n < 10 ? n : 10;
  for (int g = warp; g < G; g += nwarps) {
    const float* irow = ious + (long long)g * ld;
    const float* crow = cost + (long long)g * ld;
    // ---- top-k IoUs -> dynamic_k
    float tv[10];
    float pv = INFINITY; int pi = -1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      tv[r] = 0.0f;
      if (r < k_top) {
        float v; int i;
        warp_next<true>(irow, n, pv, pi, v, i);
        tv[r] = v; pv = v; pi = i;
      }
    }
    float sum = 0.0f;
    if (k_top >= 8) {
      // torch CPU inner sum of a contiguous row of k floats (one 8-lane vector + scalar tail):
      // tail first, then the eight vector lanes in order
#pragma unroll
      for (int r = 8; r < 10; ++r) if (r < k_top) sum = __fadd_rn(sum, tv[r]);
#pragma unroll
      for (int r = 0; r < 8; ++r) sum = __fadd_rn(sum, tv[r]);
    } else {
#pragma unroll
      for (int r = 0; r < 8; ++r) if (r < k_top) sum = __fadd_rn(sum, tv[r]);
    }
    int dk = (int)sum;
    if (dk < 1) dk = 1;
    if (dk > n) dk = n;
    // ---- dynamic_k smallest costs
    pv = -INFINITY; pi = -1;
    for (int r = 0; r < dk; ++r) {
      float v; int i;
      warp_next<false>(crow, n, pv, pi, v, i);
      if (i == 0x7fffffff) break;
      if (lane == 0) { atomicAdd(&cnt[i], 1); match_gt[i] = g; }
      pv = v; pi = i;
    }
  }
  __syncthreads();
  // ---- conflicts: argmin over all GTs
  int local_fg = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    const int c = cnt[i];
    int mg = -1;
    if (c > 1) {
      float best = cost[i]; mg = 0;
      for (int g = 1; g < G; ++g) {
        const float v = cost[(long long)g * ld + i];
        if (v < best) { best = v; mg = g; }
      }
      match_gt[i] = mg;
    } else if (c == 1) {
      mg = match_gt[i];
    }
    if (mg >= 0) { match_iou[i] = ious[(long long)mg * ld + i]; ++local_fg; }
    else match_iou[i] = 0.0f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local_fg += __shfl_xor_sync(0xffffffffu, local_fg, o);
  if (lane == 0 && local_fg) atomicAdd(&s_fg, local_fg);
  __syncthreads();
  if (tid == 0) *num_fg = s_fg;
}

__global__ void __launch_bounds__(kSimThreads) simota_matching_kernel(const float* cost, const float* ious, int G,
                                                                      int n, long long ld, int* match_gt,
                                                                      float* match_iou, int* num_fg) {
  extern __shared__ int dyn_cnt[];
  simota_matching_cta(cost, ious, G, n, ld, match_gt, match_iou, num_fg, dyn_cnt);
}

// ------------------------------------------------------------------------------------------
// full assignment
// ------------------------------------------------------------------------------------------
struct SimotaArgs {
  const float* pred; const float* labels;
  const float* xs; const float* ys; const float* st;
  int batch, anchors, nc, max_gt;
  unsigned char* fg_mask; int* matched_gt; float* matched_iou; int* matched_cls; int* num_fg; int* num_gt;
  // workspace (per image)
  int* cand;        // [anchors]
  float* S;         // [ncap]
  float* cost;      // [max_gt][ncap]
  float* iou;       // [max_gt][ncap]
  int* m_gt;        // [ncap]
  float* m_iou;     // [ncap]
  int ncap;
};

__device__ __forceinline__ bool in_centre(float gx, float gy, float xc, float yc, float s) {
  // yolo_head.py:520-538
  const float d = __fmul_rn(s, 1.5f);
  const float c_l = __fsub_rn(xc, __fsub_rn(gx, d));
  const float c_r = __fsub_rn(__fadd_rn(gx, d), xc);
  const float c_t = __fsub_rn(yc, __fsub_rn(gy, d));
  const float c_b = __fsub_rn(__fadd_rn(gy, d), yc);
  return fminf(fminf(c_l, c_t), fminf(c_r, c_b)) > 0.0f;
}

__global__ void __launch_bounds__(kSimThreads) simota_assign_kernel(const SimotaArgs a) {
  extern __shared__ int dyn_cnt[];        // [ncap] matching counters
  __shared__ float s_gt[128 * 5];          // cls, cx, cy, w, h
  __shared__ int s_G, s_n;
  __shared__ int warp_sums[kSimWarps];
  __shared__ int s_base;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int A = a.anchors, nch = 5 + a.nc;
  const float* pred = a.pred + (long long)b * A * nch;
  const float* lab = a.labels + (long long)b * a.max_gt * 5;

  // ---- number of GTs: rows with sum > 0 (yolo_head.py:269); the first num_gt rows are used
  if (tid == 0) { s_G = 0; s_base = 0; }
  __syncthreads();
  int valid = 0;
  if (tid < a.max_gt) {
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 5; ++j) s = __fadd_rn(s, lab[tid * 5 + j]);
    valid = s > 0.0f;
  }
  {
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (lane == 0 && bal) atomicAdd(&s_G, __popc(bal));
  }
  __syncthreads();
  const int G = s_G;
  for (int i = tid; i < G * 5; i += kSimThreads) s_gt[i] = lab[i];

  unsigned char* fg = a.fg_mask + (long long)b * A;
  int* mgt = a.matched_gt + (long long)b * A;
  float* miou = a.matched_iou + (long long)b * A;
  int* mcls = a.matched_cls + (long long)b * A;
  for (int i = tid; i < A; i += kSimThreads) { fg[i] = 0; mgt[i] = -1; miou[i] = 0.0f; mcls[i] = -1; }
  if (tid == 0) { a.num_gt[b] = G; a.num_fg[b] = 0; }
  __syncthreads();
  if (G == 0) return;

  // ---- geometry constraint + ordered compaction of the in-centre anchors (anchor order)
  int* cand = a.cand + (long long)b * A;
  for (int a0 = 0; a0 < A; a0 += kSimThreads) {
    const int an = a0 + tid;
    bool any = false;
    if (an < A) {
      const float s = a.st[an];
      const float xc = __fmul_rn(__fadd_rn(a.xs[an], 0.5f), s);
      const float yc = __fmul_rn(__fadd_rn(a.ys[an], 0.5f), s);
      for (int g = 0; g < G && !any; ++g) any = in_centre(s_gt[g * 5 + 1], s_gt[g * 5 + 2], xc, yc, s);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, any);
    const int wprefix = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_sums[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
      int v = lane < kSimWarps ? warp_sums[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      if (lane < kSimWarps) warp_sums[lane] = v;
    }
    __syncthreads();
    if (any) {
      const int dst = s_base + (warp ? warp_sums[warp - 1] : 0) + wprefix;
      if (dst < a.ncap) cand[dst] = an;
    }
    __syncthreads();
    if (tid == 0) s_base += warp_sums[kSimWarps - 1];
    __syncthreads();
  }
  if (tid == 0) s_n = s_base < a.ncap ? s_base : a.ncap;
  __syncthreads();
  const int n = s_n;
  if (n == 0) return;

  float* S = a.S + (long long)b * a.ncap;
  float* cost = a.cost + (long long)b * a.max_gt * a.ncap;
  float* iou = a.iou + (long long)b * a.max_gt * a.ncap;
  int* m_gt = a.m_gt + (long long)b * a.ncap;
  float* m_iou = a.m_iou + (long long)b * a.ncap;

  // ---- S_i = sum_c -max(log(1 - p_ic), -100)   (binary_cross_entropy clamps log at -100)
  for (int i = warp; i < n; i += kSimWarps) {
    const float* row = pred + (long long)cand[i] * nch;
    const float so = 1.0f / (1.0f + expf(-row[4]));
    float acc = 0.0f;
    for (int c = lane; c < a.nc; c += 32) {
      const float sc = 1.0f / (1.0f + expf(-row[5 + c]));
      const float p = sqrtf(sc * so);
      acc += -fmaxf(logf(1.0f - p), -100.0f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) S[i] = acc;
  }
  __syncthreads();

  // ---- pairwise IoU + cost
  for (int idx = tid; idx < G * n; idx += kSimThreads) {
    const int g = idx / n, i = idx - g * n;
    const int an = cand[i];
    const float* row = pred + (long long)an * nch;
    const float gx = s_gt[g * 5 + 1], gy = s_gt[g * 5 + 2], gw = s_gt[g * 5 + 3], gh = s_gt[g * 5 + 4];
    const float px = row[0], py = row[1], pw = row[2], ph = row[3];
    // bboxes_iou(gt, pred, xyxy=False)  (boxes.py:88-101)
    const float tlx = fmaxf(__fsub_rn(gx, __fdiv_rn(gw, 2.0f)), __fsub_rn(px, __fdiv_rn(pw, 2.0f)));
    const float tly = fmaxf(__fsub_rn(gy, __fdiv_rn(gh, 2.0f)), __fsub_rn(py, __fdiv_rn(ph, 2.0f)));
    const float brx = fminf(__fadd_rn(gx, __fdiv_rn(gw, 2.0f)), __fadd_rn(px, __fdiv_rn(pw, 2.0f)));
    const float bry = fminf(__fadd_rn(gy, __fdiv_rn(gh, 2.0f)), __fadd_rn(py, __fdiv_rn(ph, 2.0f)));
    const float en = (tlx < brx && tly < bry) ? 1.0f : 0.0f;
    const float area_i = __fmul_rn(__fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly)), en);
    const float v_iou = __fdiv_rn(area_i, __fsub_rn(__fadd_rn(__fmul_rn(gw, gh), __fmul_rn(pw, ph)), area_i));
    // class cost
    const int gc = (int)s_gt[g * 5 + 0];
    const float so = 1.0f / (1.0f + expf(-row[4]));
    const float sc = 1.0f / (1.0f + expf(-row[5 + gc]));
    const float p = sqrtf(sc * so);
    const float cls_cost = S[i] + fmaxf(logf(1.0f - p), -100.0f) - fmaxf(logf(p), -100.0f);
    const float s = a.st[an];
    const float xc = __fmul_rn(__fadd_rn(a.xs[an], 0.5f), s);
    const float yc = __fmul_rn(__fadd_rn(a.ys[an], 0.5f), s);
    const float pen = in_centre(gx, gy, xc, yc, s) ? 0.0f : 1.0e6f;
    const float iou_loss = -logf(v_iou + 1e-8f);
    cost[(long long)g * a.ncap + i] = __fadd_rn(__fadd_rn(cls_cost, __fmul_rn(3.0f, iou_loss)), pen);
    iou[(long long)g * a.ncap + i] = v_iou;
  }
  __syncthreads();

  // ---- matching, then scatter to the dense per-anchor outputs
  simota_matching_cta(cost, iou, G, n, a.ncap, m_gt, m_iou, a.num_fg + b, dyn_cnt);
  __syncthreads();
  for (int i = tid; i < n; i += kSimThreads) {
    const int g = m_gt[i];
    if (g >= 0) {
      const int an = cand[i];
      fg[an] = 1; mgt[an] = g; miou[an] = m_iou[i]; mcls[an] = (int)s_gt[g * 5 + 0];
    }
  }
}

static inline size_t a256(size_t v) { return (v + 255) & ~size_t(255); }
static int simota_ncap(int anchors, int max_gt) {
  long long c = 27LL * max_gt;  // <= 9 in-centre anchors per level per GT (radius 1.5 strides)
  if (c > anchors) c = anchors;
  return (int)((c + 3) & ~3LL);
}

long long simota_ws_bytes(int batch, int anchors, int max_gt) {
  if (batch <= 0 || anchors <= 0 || max_gt <= 0) return 256;
  const size_t ncap = (size_t)simota_ncap(anchors, max_gt);
  size_t t = 0;
  t += a256((size_t)batch * anchors * 4);
  t += a256((size_t)batch * ncap * 4);
  t += 2 * a256((size_t)batch * max_gt * ncap * 4);
  t += 2 * a256((size_t)batch * ncap * 4);
  return (long long)t;
}

int simota_assign_launch(const float* pred, const float* labels, const float* xs, const float* ys, const float* st,
                         int batch, int anchors, int nc, int max_gt, unsigned char* fg_mask, int* matched_gt,
                         float* matched_iou, int* matched_cls, int* num_fg, int* num_gt, void* ws, long long ws_bytes,
                         cudaStream_t s) {
  YX_REQUIRE(pred && labels && xs && ys && st && fg_mask && matched_gt && matched_iou && matched_cls && num_fg && num_gt && ws,
             YX_ERR_INVALID_ARG, "simota: null pointer");
  YX_REQUIRE(batch > 0 && anchors > 0 && nc > 0, YX_ERR_INVALID_ARG, "simota: bad sizes");
  YX_REQUIRE(max_gt > 0 && max_gt <= 128, YX_ERR_UNSUPPORTED, "simota: max_gt=%d (1..128 supported)", max_gt);
  YX_REQUIRE(simota_ws_bytes(batch, anchors, max_gt) <= ws_bytes, YX_ERR_CAPACITY, "simota: workspace too small");
  SimotaArgs a;
  memset(&a, 0, sizeof(a));
  a.pred = pred; a.labels = labels; a.xs = xs; a.ys = ys; a.st = st;
  a.batch = batch; a.anchors = anchors; a.nc = nc; a.max_gt = max_gt;
  a.fg_mask = fg_mask; a.matched_gt = matched_gt; a.matched_iou = matched_iou; a.matched_cls = matched_cls;
  a.num_fg = num_fg; a.num_gt = num_gt;
  a.ncap = simota_ncap(anchors, max_gt);
  uint8_t* p = reinterpret_cast<uint8_t*>(ws);
  size_t off = 0;
  a.cand = reinterpret_cast<int*>(p + off); off += a256((size_t)batch * anchors * 4);
  a.S = reinterpret_cast<float*>(p + off); off += a256((size_t)batch * a.ncap * 4);
  a.cost = reinterpret_cast<float*>(p + off); off += a256((size_t)batch * max_gt * a.ncap * 4);
  a.iou = reinterpret_cast<float*>(p + off); off += a256((size_t)batch * max_gt * a.ncap * 4);
  a.m_gt = reinterpret_cast<int*>(p + off); off += a256((size_t)batch * a.ncap * 4);
  a.m_iou = reinterpret_cast<float*>(p + off); off += a256((size_t)batch * a.ncap * 4);
  const size_t smem = (size_t)a.ncap * 4;
  simota_assign_kernel<<<batch, kSimThreads, smem, s>>>(a);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

int simota_matching_launch(const float* cost, const float* ious, int G, int n, long long ld, int* match_gt,
                           float* match_iou, int* num_fg, cudaStream_t s) {
  YX_REQUIRE(cost && ious && match_gt && match_iou && num_fg, YX_ERR_INVALID_ARG, "simota_matching: null pointer");
  YX_REQUIRE(G > 0 && n > 0 && ld >= n, YX_ERR_INVALID_ARG, "simota_matching: bad sizes");
  YX_REQUIRE(n <= 40000, YX_ERR_UNSUPPORTED, "simota_matching: n=%d exceeds the shared-memory counter capacity", n);
  const size_t smem = (size_t)n * 4;
  if (smem > 48 * 1024)
    YX_CUDA(cudaFuncSetAttribute(simota_matching_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  simota_matching_kernel<<<1, kSimThreads, smem, s>>>(cost, ious, G, n, ld, match_gt, match_iou, num_fg);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

}  // namespace yx
