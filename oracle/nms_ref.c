/* CPU oracle for greedy NMS - TEST INFRASTRUCTURE ONLY (see oracle/head_oracle.py header).
 *
 * Plain-C restatement of the algorithm the reference reaches through
 *   CenterNet2/centernet/modeling/layers/ml_nms.py:30 -> d2!/layers/nms.py:21 ->
 *   torchvision.ops.boxes.batched_nms -> torchvision::nms (CPU kernel)
 * torchvision is a third-party dependency that is not under /root/reference
 * (reference pin 0.8.2+cu101, log:19; the container has 0.26.0).  Its published CPU
 * algorithm, restated: areas are precomputed in fp32; boxes are visited in stable
 * score-descending order; a later box j is suppressed by a kept box i when
 *   inter / (area_i + area_j - inter) > iou_threshold
 * with the fp32 ratio promoted to double for the comparison.  batched_nms (0.8.2) adds
 * idx * (max_coordinate + 1) to every coordinate first.
 * Pinned by tests/test_oracle_golden.py against torchvision itself and against keep
 * vectors recorded from the reference's own batched_nms call (tests/golden/ops.npz).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (no FMA contraction: every operation rounds
 * separately, like the baseline-x86-64 build of torchvision).
 */
#include <stdint.h>
#include <stdlib.h>

typedef struct { float score; int64_t idx; } entry_t;

static int cmp_desc_stable(const void* a, const void* b) {
  const entry_t* x = (const entry_t*)a;
  const entry_t* y = (const entry_t*)b;
  if (x->score > y->score) return -1;
  if (x->score < y->score) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

/* boxes [n][4] xyxy, scores [n], idxs [n] or NULL; keep [n] out. Returns number kept. */
int64_t fod_oracle_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t n,
                               double iou_threshold, int64_t* keep) {
  if (n <= 0) return 0;
  float* b = (float*)malloc(sizeof(float) * 4 * (size_t)n);
  float* area = (float*)malloc(sizeof(float) * (size_t)n);
  entry_t* order = (entry_t*)malloc(sizeof(entry_t) * (size_t)n);
  unsigned char* suppressed = (unsigned char*)calloc((size_t)n, 1);
  float maxc = boxes[0];
  for (int64_t i = 0; i < 4 * n; ++i) if (boxes[i] > maxc) maxc = boxes[i];
  const float step = maxc + 1.0f;
  for (int64_t i = 0; i < n; ++i) {
    const float off = idxs ? (float)idxs[i] * step : 0.0f;
    for (int k = 0; k < 4; ++k) b[4 * i + k] = idxs ? boxes[4 * i + k] + off : boxes[4 * i + k];
    area[i] = (b[4 * i + 2] - b[4 * i + 0]) * (b[4 * i + 3] - b[4 * i + 1]);
    order[i].score = scores[i];
    order[i].idx = i;
  }
  qsort(order, (size_t)n, sizeof(entry_t), cmp_desc_stable);
  int64_t nk = 0;
  for (int64_t _i = 0; _i < n; ++_i) {
    const int64_t i = order[_i].idx;
    if (suppressed[i]) continue;
    keep[nk++] = i;
    const float ix1 = b[4 * i], iy1 = b[4 * i + 1], ix2 = b[4 * i + 2], iy2 = b[4 * i + 3], iarea = area[i];
    for (int64_t _j = _i + 1; _j < n; ++_j) {
      const int64_t j = order[_j].idx;
      if (suppressed[j]) continue;
      const float xx1 = ix1 > b[4 * j] ? ix1 : b[4 * j];
      const float yy1 = iy1 > b[4 * j + 1] ? iy1 : b[4 * j + 1];
      const float xx2 = ix2 < b[4 * j + 2] ? ix2 : b[4 * j + 2];
      const float yy2 = iy2 < b[4 * j + 3] ? iy2 : b[4 * j + 3];
      float w = xx2 - xx1, h = yy2 - yy1;
      if (w < 0.0f) w = 0.0f;
      if (h < 0.0f) h = 0.0f;
      const float inter = w * h;
      const float ovr = inter / (iarea + area[j] - inter);
      if ((double)ovr > iou_threshold) suppressed[j] = 1;
    }
  }
  free(b); free(area); free(order); free(suppressed);
  return nk;
}
