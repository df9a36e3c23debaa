/* C restatement (test infrastructure) of the integer/ordering-sensitive parts of the reference:
 *  - greedy NMS as torchvision.ops.nms computes it on CPU (third-party dependency of
 *    yolox/utils/boxes.py:56-67; torchvision pinned 0.17.2 in poetry.lock:2084-2085, 0.26.0
 *    installed; csrc/ops/cpu/nms_kernel.cpp: stable descending sort, fp32 areas/IoU, strict '>'
 *    against the double threshold);
 *  - YoloxHead.simota_matching (yolox/models/yolo_head.py:542-574).
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC oracle.c -o _build/liboracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float key; int idx; } kv_t;

static int cmp_desc(const void* a, const void* b) {
  const kv_t* x = (const kv_t*)a; const kv_t* y = (const kv_t*)b;
  if (x->key > y->key) return -1;
  if (x->key < y->key) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);   /* stable: lower index first */
}
static int cmp_asc(const void* a, const void* b) {
  const kv_t* x = (const kv_t*)a; const kv_t* y = (const kv_t*)b;
  if (x->key < y->key) return -1;
  if (x->key > y->key) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

/* keep[] receives indices in descending-score order; returns their number. */
int oracle_nms(const float* boxes, const float* scores, int n, double thr, int64_t* keep) {
  if (n <= 0) return 0;
  kv_t* order = (kv_t*)malloc(sizeof(kv_t) * n);
  float* areas = (float*)malloc(sizeof(float) * n);
  unsigned char* sup = (unsigned char*)calloc(n, 1);
  for (int i = 0; i < n; ++i) {
    order[i].key = scores[i]; order[i].idx = i;
    areas[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
  }
  qsort(order, n, sizeof(kv_t), cmp_desc);
  int num = 0;
  for (int _i = 0; _i < n; ++_i) {
    const int i = order[_i].idx;
    if (sup[i]) continue;
    keep[num++] = i;
    const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
    const float iarea = areas[i];
    for (int _j = _i + 1; _j < n; ++_j) {
      const int j = order[_j].idx;
      if (sup[j]) continue;
      const float xx1 = fmaxf(ix1, boxes[4 * j]), yy1 = fmaxf(iy1, boxes[4 * j + 1]);
      const float xx2 = fminf(ix2, boxes[4 * j + 2]), yy2 = fminf(iy2, boxes[4 * j + 3]);
      const float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
      const float inter = w * h;
      const float ovr = inter / (iarea + areas[j] - inter);
      if ((double)ovr > thr) sup[j] = 1;
    }
  }
  free(order); free(areas); free(sup);
  return num;
}

/* simota_matching on cost/ious [G, n] (row-major, contiguous).
 * match_gt[n] = matched GT index or -1, match_iou[n]; returns num_fg.
 * topk ties are resolved towards the lower index (stable sort); the sum of the top-10 IoUs
 * follows torch's CPU inner-reduction order for a contiguous row (8-lane vector + scalar tail). */
int oracle_simota_matching(const float* cost, const float* ious, int G, int n, int32_t* match_gt, float* match_iou) {
  int* cnt = (int*)calloc(n, sizeof(int));
  kv_t* tmp = (kv_t*)malloc(sizeof(kv_t) * n);
  for (int i = 0; i < n; ++i) { match_gt[i] = -1; match_iou[i] = 0.0f; }
  const int k = n < 10 ? n : 10;
  for (int g = 0; g < G; ++g) {
    for (int i = 0; i < n; ++i) { tmp[i].key = ious[(size_t)g * n + i]; tmp[i].idx = i; }
    qsort(tmp, n, sizeof(kv_t), cmp_desc);
    float s = 0.0f;
    if (k >= 8) {
      for (int r = 8; r < k; ++r) s = s + tmp[r].key;
      for (int r = 0; r < 8; ++r) s = s + tmp[r].key;
    } else {
      for (int r = 0; r < k; ++r) s = s + tmp[r].key;
    }
    int dk = (int)s;
    if (dk < 1) dk = 1;
    if (dk > n) dk = n;
    for (int i = 0; i < n; ++i) { tmp[i].key = cost[(size_t)g * n + i]; tmp[i].idx = i; }
    qsort(tmp, n, sizeof(kv_t), cmp_asc);
    for (int r = 0; r < dk; ++r) { cnt[tmp[r].idx] += 1; match_gt[tmp[r].idx] = g; }
  }
  int num_fg = 0;
  for (int i = 0; i < n; ++i) {
    if (cnt[i] > 1) {
      int best = 0; float bv = cost[i];
      for (int g = 1; g < G; ++g) if (cost[(size_t)g * n + i] < bv) { bv = cost[(size_t)g * n + i]; best = g; }
      match_gt[i] = best;
    }
    if (cnt[i] >= 1) { match_iou[i] = ious[(size_t)match_gt[i] * n + i]; ++num_fg; }
  }
  free(cnt); free(tmp);
  return num_fg;
}
