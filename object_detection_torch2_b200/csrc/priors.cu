// P1: SSD300 default boxes, replaces SSD._get_default_bboxes (reference src/model/ssd.py:108-133).
//
// The reference evaluates (i + .5) / m, s_k * sqrt(a) etc. in Python float64 and rounds once to fp32 when
// the row is wrapped in torch.Tensor([...]) (ssd.py:130).  The 34 distinct (w, h) shapes are computed on
// the host in double and rounded to fp32 there (C's pow/sqrt and Python's ** agree after the fp32 cast);
// the kernel computes the centres in fp64 and rounds once, so the table is bit-identical.
#include <math.h>

#include "common.cuh"

namespace ssdh {

constexpr int kLevels = 6;
constexpr int kShapes = 34;   // 4 + 6 + 6 + 6 + 4 + 4

struct PriorTable {
  int cells[kLevels];        // feature map side m
  int anchors[kLevels];      // anchors per cell
  int row_off[kLevels + 1];  // first row of each level
  int shape_off[kLevels];    // first entry of each level in w/h
  float w[kShapes], h[kShapes];
};

static PriorTable build_table() {
  PriorTable t;
  const int cells[kLevels] = {38, 19, 10, 5, 3, 1};
  const int anchors[kLevels] = {4, 6, 6, 6, 4, 4};
  const double s_min = 0.2, s_max = 0.9;
  auto scale = [&](int k) { return s_min + (s_max - s_min) * (k - 1) / (kLevels - 1); };   // ssd.py:114-115
  int row = 0, sh = 0;
  for (int l = 0; l < kLevels; ++l) {
    t.cells[l] = cells[l];
    t.anchors[l] = anchors[l];
    t.row_off[l] = row;
    t.shape_off[l] = sh;
    const int k = l + 1;
    const double ratios6[5] = {1.0, 2.0, 0.5, 3.0, 1.0 / 3.0};                               // ssd.py:121
    const int n_ratio = anchors[l] - 1;
    for (int r = 0; r < n_ratio; ++r) {
      t.w[sh] = static_cast<float>(scale(k) * pow(ratios6[r], 0.5));                         // ssd.py:128
      t.h[sh] = static_cast<float>(scale(k) * pow(1.0 / ratios6[r], 0.5));                   // ssd.py:129
      ++sh;
    }
    const double extra = pow(scale(k) * scale(k + 1), 0.5);                                  // ssd.py:126
    t.w[sh] = t.h[sh] = static_cast<float>(extra);
    ++sh;
    row += cells[l] * cells[l] * anchors[l];
  }
  t.row_off[kLevels] = row;
  return t;
}

__global__ void __launch_bounds__(256) default_boxes_kernel(PriorTable t, float4* __restrict__ out) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= t.row_off[kLevels]) return;
  int l = 0;
#pragma unroll
  for (int q = 1; q < kLevels; ++q) l += (row >= t.row_off[q]);
  const int m = t.cells[l], A = t.anchors[l];
  const int local = row - t.row_off[l];
  const int a = local % A, cell = local / A;
  const int i = cell / m, j = cell % m;            // i is the OUTER loop index and drives cx (ssd.py:122-130)
  const double cx = (static_cast<double>(i) + 0.5) / static_cast<double>(m);
  const double cy = (static_cast<double>(j) + 0.5) / static_cast<double>(m);
  out[row] = make_float4(static_cast<float>(cx), static_cast<float>(cy), t.w[t.shape_off[l] + a], t.h[t.shape_off[l] + a]);
}

}  // namespace ssdh

extern "C" int ssdh_default_boxes(float* out, ssdh_stream_t stream) {
  using namespace ssdh;
  if (!out) { set_error("ssdh_default_boxes: out is NULL"); return SSDH_E_ARG; }
  if (!aligned16(out)) { set_error("ssdh_default_boxes: out must be 16-byte aligned"); return SSDH_E_ALIGN; }
  static const PriorTable table = build_table();
  const int rows = table.row_off[kLevels];
  default_boxes_kernel<<<(rows + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(table, reinterpret_cast<float4*>(out));
  return cuda_status("ssdh_default_boxes");
}
