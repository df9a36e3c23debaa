// SimOTA label assignment, batched over images, no host synchronisation.
//   reference: YoloxHead.get_assignments / get_geometry_constraint / simota_matching
//   (yolox/models/yolo_head.py:420-574) and bboxes_iou (yolox/utils/boxes.py:78-101).
// The reference loops over images in Python with >= G+3 device->host syncs per image and
// materialises a [G, A', 80] BCE tensor; here the class cost of a (gt, anchor) pair is
//   S_a - (-log(1-p_a,c)) + (-log p_a,c),   S_a = sum_c -log(1-p_a,c),  p = sqrt(sig(cls)*sig(obj))
// so only a [G, A'] cost/IoU pair of matrices is ever written (A' <= 9*levels*G in-centre anchors).
//
// One thread-block CLUSTER per image (1..8 CTAs, chosen so that the batch covers the chip: the 8 images per rank
// of the training step run on 64 SMs instead of 8). The CTAs of a cluster split the anchors (geometry), the
// candidate tiles (cost / IoU), the GT rows (top-k matching) and the candidates again (conflicts + scatter); they meet at
// cluster barriers (release / acquire at cluster scope orders the global-memory workspace between them).
//
// simota_matching (yolo_head.py:542-574) is reproduced bit-exactly on a given cost/IoU matrix:
//   dynamic_k = max(1, int(sum of the top-10 IoUs))  -- summed in the order torch's CPU
//               inner-reduction uses for a row of 10 floats (elements 8,9 first, then 0..7);
//   per GT the dynamic_k smallest costs (ties: lower index first);
//   anchors chosen by several GTs go to argmin_g cost[g][a] over ALL g (first index on ties).
// dynamic_k <= 10 always (ten IoUs <= 1), so both selections are "ten best of a row in (value, index) order":
// one pass over the row with a sorted ten-entry list per lane, then a ten-round merge across the warp.
#include <string.h>
#include <math.h>

#include "yx_common.cuh"

namespace yx {

static constexpr int kSimThreads = 512;
static constexpr int kSimWarps = kSimThreads / 32;
static constexpr int kSimTile = 32;          // candidates per cost tile (one 128-byte row segment per GT)
static constexpr int kSimMaxCluster = 8;
static constexpr int kNone = 0x7fffffff;

// (value, index) strict order. DESC: value descending, index ascending; else value ascending, index ascending.
// NaN never compares "before" anything and is never selected (as in the sequential selection this replaces).
template <bool DESC>
__device__ __forceinline__ bool before(float v, int i, float w, int j) {
  return DESC ? ((v > w) || (v == w && i < j)) : ((v < w) || (v == w && i < j));
}

// The ten first elements of row[0..n) in `before` order, left in lane 0..31 registers as the merged list tv/ti
// (identical in every lane). Entries past the end of the row have index kNone.
template <bool DESC>
__device__ __forceinline__ void warp_top10(const float* __restrict__ row, int n, float (&tv)[10], int (&ti)[10]) {
  const int lane = threadIdx.x & 31;
  const float worst = DESC ? -INFINITY : INFINITY;
  float lv[10]; int li[10];
#pragma unroll
  for (int r = 0; r < 10; ++r) { lv[r] = worst; li[r] = kNone; }
  // eight loads in flight per lane: the rows were written by other SMs (L2 round trips), and the insertion branch below
  // would otherwise serialise one round trip per element
  for (int i0 = lane; i0 < n; i0 += 32 * 8) {
    float v8[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v8[u] = (i0 + 32 * u < n) ? row[i0 + 32 * u] : worst;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + 32 * u;
      const float v = v8[u];
      if (i < n && before<DESC>(v, i, lv[9], li[9])) {
        lv[9] = v; li[9] = i;
#pragma unroll
        for (int r = 9; r > 0; --r) {
          if (before<DESC>(lv[r], li[r], lv[r - 1], li[r - 1])) {
            const float fv = lv[r]; lv[r] = lv[r - 1]; lv[r - 1] = fv;
            const int fi = li[r]; li[r] = li[r - 1]; li[r - 1] = fi;
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    float bv = lv[0]; int bi = li[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const bool take = oi != kNone && (bi == kNone || before<DESC>(ov, oi, bv, bi));
      if (take) { bv = ov; bi = oi; }
    }
    tv[r] = bv; ti[r] = bi;
    if (bi != kNone && li[0] == bi) {          // the winning lane pops its head
#pragma unroll
      for (int q = 0; q < 9; ++q) { lv[q] = lv[q + 1]; li[q] = li[q + 1]; }
      lv[9] = worst; li[9] = kNone;
    }
  }
}

// One GT row of simota_matching (one warp): dynamic_k from the IoU row, then the dynamic_k cheapest candidates.
__device__ __forceinline__ void match_row(const float* __restrict__ crow, const float* __restrict__ irow, int n, int g,
                                          int* __restrict__ cnt, int* __restrict__ match_gt) {
  const int lane = threadIdx.x & 31;
  const int k_top = n < 10 ? n : 10;
  float tv[10]; int ti[10];
  warp_top10<true>(irow, n, tv, ti);
#pragma unroll
  for (int r = 0; r < 10; ++r) if (r >= k_top || ti[r] == kNone) tv[r] = 0.0f;
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
  warp_top10<false>(crow, n, tv, ti);
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      if (r < dk && ti[r] != kNone) { atomicAdd(&cnt[ti[r]], 1); match_gt[ti[r]] = g; }
    }
  }
}

// anchors matched by several GTs keep argmin_g cost (first index on ties); returns the matched GT or -1
__device__ __forceinline__ int resolve(const float* __restrict__ cost, long long ld, int G, int i, int c, int single) {
  if (c > 1) {
    float best = cost[i]; int mg = 0;
    for (int g = 1; g < G; ++g) {
      const float v = cost[(long long)g * ld + i];
      if (v < best) { best = v; mg = g; }
    }
    return mg;
  }
  return c == 1 ? single : -1;
}

// stand-alone matching on a given [G, n] cost / IoU pair: one CTA; cnt: shared int[n]
__global__ void __launch_bounds__(kSimThreads) simota_matching_kernel(const float* cost, const float* ious, int G,
                                                                      int n, long long ld, int* match_gt,
                                                                      float* match_iou, int* num_fg) {
  extern __shared__ int cnt[];
  __shared__ int s_fg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n; i += kSimThreads) { cnt[i] = 0; match_gt[i] = -1; }
  if (tid == 0) s_fg = 0;
  __syncthreads();
  for (int g = warp; g < G; g += kSimWarps) match_row(cost + (long long)g * ld, ious + (long long)g * ld, n, g, cnt, match_gt);
  __syncthreads();
  int local_fg = 0;
  for (int i = tid; i < n; i += kSimThreads) {
    const int mg = resolve(cost, ld, G, i, cnt[i], match_gt[i]);
    match_gt[i] = mg;
    if (mg >= 0) { match_iou[i] = ious[(long long)mg * ld + i]; ++local_fg; }
    else match_iou[i] = 0.0f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local_fg += __shfl_xor_sync(0xffffffffu, local_fg, o);
  if (lane == 0 && local_fg) atomicAdd(&s_fg, local_fg);
  __syncthreads();
  if (tid == 0) *num_fg = s_fg;
}

// ------------------------------------------------------------------------------------------
// full assignment
// ------------------------------------------------------------------------------------------
struct SimotaArgs {
  const float* pred; const float* labels;
  const float* xs; const float* ys; const float* st;
  int batch, anchors, nc, max_gt;
  unsigned char* fg_mask; int* matched_gt; float* matched_iou; int* matched_cls; int* num_fg; int* num_gt; int* status;
  // workspace (per image)
  int* cand;        // [ncap]  anchor index of every in-centre candidate, anchor order
  int* counts;      // [kSimMaxCluster] in-centre anchors found by every CTA of the cluster
  float* cost;      // [max_gt][ncap]
  float* iou;       // [max_gt][ncap]
  int* cnt;         // [ncap]  how many GTs selected the candidate
  int* m_gt;        // [ncap]  one of them
  int ncap;
  int chunk;        // anchors per CTA of the cluster (multiple of 32)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r;
}
// every thread of every CTA of the cluster; orders global-memory accesses before / after it at cluster scope
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ bool in_centre(float gx, float gy, float xc, float yc, float s) {
  // yolo_head.py:520-538
  const float d = __fmul_rn(s, 1.5f);
  const float c_l = __fsub_rn(xc, __fsub_rn(gx, d));
  const float c_r = __fsub_rn(__fadd_rn(gx, d), xc);
  const float c_t = __fsub_rn(yc, __fsub_rn(gy, d));
  const float c_b = __fsub_rn(__fadd_rn(gy, d), yc);
  return fminf(fminf(c_l, c_t), fminf(c_r, c_b)) > 0.0f;
}

// dynamic shared memory: s_gt [max_gt*5] float | ballots [chunk/32] u32 | tile_c [max_gt][32] float | tile_i [max_gt][32]
//                        | s_gc [max_gt*5] | s_tab [16 warps][nc][2]
__global__ void __launch_bounds__(kSimThreads) simota_assign_kernel(const SimotaArgs a) {
  extern __shared__ __align__(16) unsigned char sim_smem[];
  float* s_gt = reinterpret_cast<float*>(sim_smem);                 // cls, cx, cy, w, h
  uint32_t* s_bal = reinterpret_cast<uint32_t*>(s_gt + a.max_gt * 5);
  float* tile_c = reinterpret_cast<float*>(s_bal + a.chunk / 32);
  float* tile_i = tile_c + a.max_gt * kSimTile;
  float* s_gc = tile_i + a.max_gt * kSimTile;            // [max_gt][5] GT corners + area
  float* s_tab = s_gc + a.max_gt * 5;                    // [kSimWarps][nc][2] per-class log terms of a warp's candidate
  __shared__ int s_G, s_cnt, s_bad;
  __shared__ int warp_sums[kSimWarps];
  const int CL = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
  const int b = blockIdx.x / CL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int A = a.anchors, nch = 5 + a.nc;
  const float* pred = a.pred + (long long)b * A * nch;
  const float* lab = a.labels + (long long)b * a.max_gt * 5;

  // ---- number of GTs: rows with sum > 0 (yolo_head.py:269); the first num_gt rows are used
  if (tid == 0) { s_G = 0; s_cnt = 0; s_bad = 0; }
  __syncthreads();
  for (int r0 = 0; r0 < a.max_gt; r0 += kSimThreads) {
    int valid = 0;
    if (r0 + tid < a.max_gt) {
      float s = 0.0f;
#pragma unroll
      for (int j = 0; j < 5; ++j) s = __fadd_rn(s, lab[(r0 + tid) * 5 + j]);
      valid = s > 0.0f;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (lane == 0 && bal) atomicAdd(&s_G, __popc(bal));
  }
  __syncthreads();
  const int G = s_G;
  for (int i = tid; i < G * 5; i += kSimThreads) s_gt[i] = lab[i];

  // ---- dense outputs of this CTA's anchor range default to background
  const int a_lo = rank * a.chunk, a_hi = min(A, a_lo + a.chunk);
  unsigned char* fg = a.fg_mask + (long long)b * A;
  int* mgt = a.matched_gt + (long long)b * A;
  float* miou = a.matched_iou + (long long)b * A;
  int* mcls = a.matched_cls + (long long)b * A;
  for (int i = a_lo + tid; i < a_hi; i += kSimThreads) { fg[i] = 0; mgt[i] = -1; miou[i] = 0.0f; mcls[i] = -1; }
  if (rank == 0 && tid == 0) { a.num_gt[b] = G; a.num_fg[b] = 0; if (a.status) a.status[b] = 0; }
  __syncthreads();
  if (G == 0) return;                                    // uniform over the cluster

  // ---- geometry constraint: one ballot word per 32 anchors of the range, count per CTA
  int* counts = a.counts + b * kSimMaxCluster;
  int mine = 0;
  for (int a0 = a_lo; a0 < a_lo + a.chunk; a0 += kSimThreads) {
    const int an = a0 + tid;
    bool any = false;
    if (an < a_hi) {
      const float s = a.st[an];
      const float xc = __fmul_rn(__fadd_rn(a.xs[an], 0.5f), s);
      const float yc = __fmul_rn(__fadd_rn(a.ys[an], 0.5f), s);
      for (int g = 0; g < G && !any; ++g) any = in_centre(s_gt[g * 5 + 1], s_gt[g * 5 + 2], xc, yc, s);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, any);
    if (lane == 0 && an - lane < a_lo + a.chunk) { s_bal[(an - a_lo) >> 5] = bal; mine += __popc(bal); }
  }
  if (lane == 0 && mine) atomicAdd(&s_cnt, mine);
  __syncthreads();
  if (tid == 0) counts[rank] = s_cnt;
  cluster_sync_all();

  // ---- ordered compaction (anchor order): CTA base from the cluster's counts, word prefix inside the CTA
  int base = 0, total = 0;
  for (int r = 0; r < CL; ++r) { const int c = counts[r]; if (r < rank) base += c; total += c; }
  const int n = total < a.ncap ? total : a.ncap;
  if (total > a.ncap && rank == 0 && tid == 0 && a.status) a.status[b] = YX_SIMOTA_CAPACITY;
  if (n == 0) return;                                    // uniform over the cluster
  int* cand = a.cand + (long long)b * a.ncap;
  const int words = a.chunk / 32;
  for (int w0 = 0; w0 < words; w0 += kSimThreads) {
    const int w = w0 + tid;
    uint32_t bal = w < words ? s_bal[w] : 0u;
    int v = __popc(bal);
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int q = 0; q < warp; ++q) wbase += warp_sums[q];
    int dst = base + wbase + incl - v;
    while (bal) {
      const int bit = __ffs(bal) - 1;
      bal &= bal - 1;
      if (dst < a.ncap) cand[dst] = a_lo + w * 32 + bit;
      ++dst;
    }
    int all = 0;
    for (int q = 0; q < kSimWarps; ++q) all += warp_sums[q];
    base += all;
    __syncthreads();
  }
  cluster_sync_all();

  // ---- cost / IoU tiles of 32 candidates: a warp per candidate (class sum S over the lanes, then a lane per GT),
  //      staged in shared memory so that every GT row is written as one 128-byte segment
  float* cost = a.cost + (long long)b * a.max_gt * a.ncap;
  float* iou = a.iou + (long long)b * a.max_gt * a.ncap;
  int* cnt = a.cnt + (long long)b * a.ncap;
  int* m_gt = a.m_gt + (long long)b * a.ncap;
  const int tiles = (n + kSimTile - 1) / kSimTile;
  // per-GT corners, computed once (x / 2 == x * 0.5 exactly): lo = c - wh/2, hi = c + wh/2, area = w * h
  for (int g = tid; g < G; g += kSimThreads) {
    const float gx = s_gt[g * 5 + 1], gy = s_gt[g * 5 + 2], gw = s_gt[g * 5 + 3], gh = s_gt[g * 5 + 4];
    float* q = s_gc + g * 5;
    q[0] = __fsub_rn(gx, __fmul_rn(gw, 0.5f)); q[1] = __fsub_rn(gy, __fmul_rn(gh, 0.5f));
    q[2] = __fadd_rn(gx, __fmul_rn(gw, 0.5f)); q[3] = __fadd_rn(gy, __fmul_rn(gh, 0.5f));
    q[4] = __fmul_rn(gw, gh);
  }
  __syncthreads();
  // many GTs: the two class-dependent log terms of a candidate are tabulated per class once (80 pairs of logs) instead of
  // being recomputed per (GT, candidate) pair -- the same expressions on the same inputs, so the costs are bit-identical
  const bool tabulate = G > 48;
  float* tab = s_tab + warp * 2 * a.nc;
  for (int t = rank; t < tiles; t += CL) {
    const int i0 = t * kSimTile;
    // the warp's two candidates (j = warp, warp + 16): anchor indices, then both rows' loads, are issued before any math
    // so that the two dependent L2 / HBM round trips per candidate overlap
    constexpr int kPer = kSimTile / kSimWarps;
    int an2[kPer]; float obj2[kPer]; float4 box2[kPer]; float cls2[kPer][4];
#pragma unroll
    for (int u = 0; u < kPer; ++u) an2[u] = (i0 + warp + u * kSimWarps < n) ? cand[i0 + warp + u * kSimWarps] : -1;
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      if (an2[u] < 0) continue;
      const float* row = pred + (long long)an2[u] * nch;
      obj2[u] = row[4];
      box2[u] = make_float4(row[0], row[1], row[2], row[3]);
#pragma unroll
      for (int q = 0; q < 4; ++q) cls2[u][q] = (lane + 32 * q < a.nc) ? row[5 + lane + 32 * q] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const int j = warp + u * kSimWarps;
      const int an = an2[u];
      if (an < 0) continue;                               // warp-uniform
      const float* row = pred + (long long)an * nch;
      // S_i = sum_c -max(log(1 - p_ic), -100)   (binary_cross_entropy clamps log at -100)
      const float so = 1.0f / (1.0f + expf(-obj2[u]));
      float acc = 0.0f;
      for (int c = lane, q = 0; c < a.nc; c += 32, ++q) {
        const float lg = q < 4 ? cls2[u][q < 4 ? q : 0] : row[5 + c];
        const float sc = 1.0f / (1.0f + expf(-lg));
        const float p = sqrtf(sc * so);
        const float l1 = fmaxf(logf(1.0f - p), -100.0f);
        acc += -l1;
        if (tabulate) { tab[2 * c] = l1; tab[2 * c + 1] = fmaxf(logf(p), -100.0f); }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      const float S = acc;
      const float px = box2[u].x, py = box2[u].y, pw = box2[u].z, ph = box2[u].w;
      const float plx = __fsub_rn(px, __fmul_rn(pw, 0.5f)), ply = __fsub_rn(py, __fmul_rn(ph, 0.5f));
      const float phx = __fadd_rn(px, __fmul_rn(pw, 0.5f)), phy = __fadd_rn(py, __fmul_rn(ph, 0.5f));
      const float parea = __fmul_rn(pw, ph);
      const float s = a.st[an];
      const float xc = __fmul_rn(__fadd_rn(a.xs[an], 0.5f), s);
      const float yc = __fmul_rn(__fadd_rn(a.ys[an], 0.5f), s);
      __syncwarp();
      for (int g = lane; g < G; g += 32) {
        const float* q = s_gc + g * 5;
        // bboxes_iou(gt, pred, xyxy=False)  (boxes.py:88-101)
        const float tlx = fmaxf(q[0], plx), tly = fmaxf(q[1], ply);
        const float brx = fminf(q[2], phx), bry = fminf(q[3], phy);
        const float en = (tlx < brx && tly < bry) ? 1.0f : 0.0f;
        const float area_i = __fmul_rn(__fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly)), en);
        const float v_iou = __fdiv_rn(area_i, __fsub_rn(__fadd_rn(q[4], parea), area_i));
        // class cost; a class id outside [0, nc) (F.one_hot raises in the reference) is flagged and clamped
        int gc = (int)s_gt[g * 5 + 0];
        if (gc < 0 || gc >= a.nc) { s_bad = 1; gc = gc < 0 ? 0 : a.nc - 1; }
        float l1, lp;
        if (tabulate) { l1 = tab[2 * gc]; lp = tab[2 * gc + 1]; }
        else {
          const float sc = 1.0f / (1.0f + expf(-row[5 + gc]));
          const float p = sqrtf(sc * so);
          l1 = fmaxf(logf(1.0f - p), -100.0f); lp = fmaxf(logf(p), -100.0f);
        }
        const float cls_cost = S + l1 - lp;
        const float pen = in_centre(s_gt[g * 5 + 1], s_gt[g * 5 + 2], xc, yc, s) ? 0.0f : 1.0e6f;
        const float iou_loss = -logf(v_iou + 1e-8f);
        tile_c[g * kSimTile + j] = __fadd_rn(__fadd_rn(cls_cost, __fmul_rn(3.0f, iou_loss)), pen);
        tile_i[g * kSimTile + j] = v_iou;
      }
      __syncwarp();
    }
    __syncthreads();
    const int valid = min(kSimTile, n - i0);
    for (int idx = tid; idx < G * kSimTile; idx += kSimThreads) {
      const int g = idx / kSimTile, j = idx % kSimTile;
      if (j < valid) {
        cost[(long long)g * a.ncap + i0 + j] = tile_c[idx];
        iou[(long long)g * a.ncap + i0 + j] = tile_i[idx];
      }
    }
    if (tid < valid) { cnt[i0 + tid] = 0; m_gt[i0 + tid] = -1; }
    __syncthreads();
  }
  if (tid == 0 && s_bad && a.status) atomicMax(a.status + b, YX_SIMOTA_BAD_CLASS);
  cluster_sync_all();

  // ---- matching: a warp per GT row over the whole cluster
  for (int g = rank * kSimWarps + warp; g < G; g += CL * kSimWarps)
    match_row(cost + (long long)g * a.ncap, iou + (long long)g * a.ncap, n, g, cnt, m_gt);
  cluster_sync_all();

  // ---- conflicts, then scatter to the dense per-anchor outputs
  int local_fg = 0;
  for (int i = rank * kSimThreads + tid; i < n; i += CL * kSimThreads) {
    const int g = resolve(cost, a.ncap, G, i, cnt[i], m_gt[i]);
    if (g >= 0) {
      const int an = cand[i];
      int gc = (int)s_gt[g * 5 + 0];
      fg[an] = 1; mgt[an] = g; miou[an] = iou[(long long)g * a.ncap + i]; mcls[an] = gc;
      ++local_fg;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local_fg += __shfl_xor_sync(0xffffffffu, local_fg, o);
  if (lane == 0 && local_fg) atomicAdd(a.num_fg + b, local_fg);
}

static inline size_t a256(size_t v) { return (v + 255) & ~size_t(255); }
static int simota_ncap(int anchors, int max_gt, int levels) {
  // <= 9 in-centre anchors per level per GT (radius 1.5 strides on a unit-spaced grid)
  long long c = 9LL * (levels < 1 ? 1 : levels) * max_gt;
  if (c > anchors) c = anchors;
  return (int)((c + 31) & ~31LL);
}

long long simota_ws_bytes(int batch, int anchors, int max_gt, int levels) {
  if (batch <= 0 || anchors <= 0 || max_gt <= 0) return 256;
  const size_t ncap = (size_t)simota_ncap(anchors, max_gt, levels);
  size_t t = 0;
  t += a256((size_t)batch * ncap * 4);                   // cand
  t += a256((size_t)batch * kSimMaxCluster * 4);         // counts
  t += 2 * a256((size_t)batch * max_gt * ncap * 4);      // cost, iou
  t += 2 * a256((size_t)batch * ncap * 4);               // cnt, m_gt
  return (long long)t;
}

int simota_assign_launch(const float* pred, const float* labels, const float* xs, const float* ys, const float* st,
                         int batch, int anchors, int nc, int max_gt, int levels, unsigned char* fg_mask, int* matched_gt,
                         float* matched_iou, int* matched_cls, int* num_fg, int* num_gt, int* status, void* ws,
                         long long ws_bytes, cudaStream_t s) {
  YX_REQUIRE(pred && labels && xs && ys && st && fg_mask && matched_gt && matched_iou && matched_cls && num_fg && num_gt && ws,
             YX_ERR_INVALID_ARG, "simota: null pointer");
  YX_REQUIRE(batch > 0 && anchors > 0 && nc > 0 && levels > 0, YX_ERR_INVALID_ARG, "simota: bad sizes");
  YX_REQUIRE(max_gt > 0 && max_gt <= 512, YX_ERR_UNSUPPORTED, "simota: max_gt=%d (1..512 label rows per image supported)", max_gt);
  YX_REQUIRE(simota_ws_bytes(batch, anchors, max_gt, levels) <= ws_bytes, YX_ERR_CAPACITY, "simota: workspace too small");
  SimotaArgs a;
  memset(&a, 0, sizeof(a));
  a.pred = pred; a.labels = labels; a.xs = xs; a.ys = ys; a.st = st;
  a.batch = batch; a.anchors = anchors; a.nc = nc; a.max_gt = max_gt;
  a.fg_mask = fg_mask; a.matched_gt = matched_gt; a.matched_iou = matched_iou; a.matched_cls = matched_cls;
  a.num_fg = num_fg; a.num_gt = num_gt; a.status = status;
  a.ncap = simota_ncap(anchors, max_gt, levels);
  // cluster size: the largest of 8/4/2/1 that keeps the grid within two waves of the chip
  int cl = kSimMaxCluster;
  while (cl > 1 && (long long)batch * cl > 2LL * num_sms()) cl >>= 1;
  a.chunk = (int)(((ceil_div64(anchors, cl) + 31) / 32) * 32);
  uint8_t* p = reinterpret_cast<uint8_t*>(ws);
  size_t off = 0;
  a.cand = reinterpret_cast<int*>(p + off); off += a256((size_t)batch * a.ncap * 4);
  a.counts = reinterpret_cast<int*>(p + off); off += a256((size_t)batch * kSimMaxCluster * 4);
  a.cost = reinterpret_cast<float*>(p + off); off += a256((size_t)batch * max_gt * a.ncap * 4);
  a.iou = reinterpret_cast<float*>(p + off); off += a256((size_t)batch * max_gt * a.ncap * 4);
  a.cnt = reinterpret_cast<int*>(p + off); off += a256((size_t)batch * a.ncap * 4);
  a.m_gt = reinterpret_cast<int*>(p + off); off += a256((size_t)batch * a.ncap * 4);
  const size_t smem = 2 * (size_t)max_gt * 5 * 4 + (size_t)(a.chunk / 32) * 4 + 2 * (size_t)max_gt * kSimTile * 4 +
                      (size_t)kSimWarps * nc * 8;
  YX_REQUIRE(smem <= 200 * 1024, YX_ERR_UNSUPPORTED, "simota: %zu bytes of shared memory needed (anchors=%d, max_gt=%d)", smem, anchors, max_gt);
  static size_t configured_dev[kMaxDevices] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (smem > 48 * 1024 && smem > configured) {
    YX_CUDA(cudaFuncSetAttribute(simota_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(batch * cl));
  cfg.blockDim = dim3(kSimThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  YX_CUDA(cudaLaunchKernelEx(&cfg, simota_assign_kernel, a));
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
