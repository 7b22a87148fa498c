// L1-L6 stand-alone stages of the MultiBox loss: matching, offset encoding, smooth-L1, softmax
// cross-entropy, 3:1 split and the (k+1)-th value.  These mirror the reference's private helpers
// (src/model/ssd.py:231-328) one to one and are the per-stage parity probes; the training path itself
// runs the fused kernel in loss.cu.
#include "common.cuh"

namespace ssdh {

// ------------------------------------------------------------------------------------------------
// L1  src/model/ssd.py:231-250
// ------------------------------------------------------------------------------------------------
struct GtGeom {
  Corners c;
  int valid;   // g_w * g_h > 0 (ssd.py:250); otherwise the "IoU" is the gt area itself
};

__device__ __forceinline__ float match_value(const GtGeom& g, const Corners& d) {
  if (!g.valid) return g.c.area;
  const float inter = intersection(g.c, d);
  return __fdiv_rn(inter, union_area(g.c, d, inter));
}

__global__ void __launch_bounds__(256)
match_kernel(const float* __restrict__ gt, int gt_row_stride, int G, const float4* __restrict__ priors, int P,
             ThrBand band, uint64_t* __restrict__ bits, uint8_t* __restrict__ mask, int32_t* __restrict__ best_gt,
             float* __restrict__ best_iou) {
  __shared__ GtGeom s_gt[kMaxGT];
  const int n = blockIdx.y;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float* row = gt + (static_cast<size_t>(n) * G + g) * gt_row_stride;
    GtGeom q;
    q.c = make_corners(row[0], row[1], row[2], row[3]);
    q.valid = q.c.area > 0.0f;
    s_gt[g] = q;
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float4 d4 = priors[p];
  const Corners d = make_corners(d4.x, d4.y, d4.z, d4.w);
  uint64_t m = 0;
  float bv = 0.0f;
  int bg = -1;
  const bool want_best = best_gt != nullptr || best_iou != nullptr;
  for (int g = 0; g < G; ++g) {
    const GtGeom q = s_gt[g];
    bool hit;
    if (want_best) {
      const float v = match_value(q, d);
      hit = v > band.thr;
      if (bg < 0 || v > bv) { bv = v; bg = g; }
    } else if (!q.valid) {
      hit = q.c.area > band.thr;
    } else {
      const float inter = intersection(q.c, d);
      hit = quotient_gt(inter, union_area(q.c, d, inter), band);
    }
    m |= static_cast<uint64_t>(hit) << g;
  }
  const size_t o = static_cast<size_t>(n) * P + p;
  if (bits) bits[o] = m;
  if (mask)
    for (int g = 0; g < G; ++g) mask[o * G + g] = (m >> g) & 1u;
  if (best_gt) best_gt[o] = bg;
  if (best_iou) best_iou[o] = bv;
}

// arg-max over priors for one (image, gt): lowest prior index among equal IoUs.
__global__ void __launch_bounds__(256)
best_prior_kernel(const float* __restrict__ gt, int gt_row_stride, int G, const float4* __restrict__ priors, int P,
                  int32_t* __restrict__ best_prior, float* __restrict__ best_prior_iou) {
  const int g = blockIdx.x, n = blockIdx.y;
  const float* row = gt + (static_cast<size_t>(n) * G + g) * gt_row_stride;
  GtGeom q;
  q.c = make_corners(row[0], row[1], row[2], row[3]);
  q.valid = q.c.area > 0.0f;
  float bv = -INFINITY;
  int bp = 0x7fffffff;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const float4 d4 = priors[p];
    const float v = match_value(q, make_corners(d4.x, d4.y, d4.z, d4.w));
    if (v > bv) { bv = v; bp = p; }
  }
  // block arg-max, ties -> lowest index
  __shared__ float s_v[256];
  __shared__ int s_p[256];
  s_v[threadIdx.x] = bv;
  s_p[threadIdx.x] = bp;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const float ov = s_v[threadIdx.x + s];
      const int op = s_p[threadIdx.x + s];
      if (ov > s_v[threadIdx.x] || (ov == s_v[threadIdx.x] && op < s_p[threadIdx.x])) {
        s_v[threadIdx.x] = ov;
        s_p[threadIdx.x] = op;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const size_t o = static_cast<size_t>(n) * G + g;
    if (best_prior) best_prior[o] = s_p[0];
    if (best_prior_iou) best_prior_iou[o] = s_v[0];
  }
}

// ------------------------------------------------------------------------------------------------
// L2  src/model/ssd.py:252-272   out[n, p, g, :] = g-hat
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
encode_kernel(const float* __restrict__ gt, int gt_row_stride, int G, const float4* __restrict__ priors, int P,
              float4* __restrict__ out, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // over N*P*G
  if (i >= total) return;
  const int g = static_cast<int>(i % G);
  const size_t np = i / G;
  const int p = static_cast<int>(np % P);
  const size_t n = np / P;
  const float* row = gt + (n * G + g) * gt_row_stride;
  const float4 d = priors[p];
  const float gw = row[2], gh = row[3];
  float4 e;
  e.x = __fdiv_rn(__fsub_rn(row[0], d.x), d.z);
  e.y = __fdiv_rn(__fsub_rn(row[1], d.y), d.w);
  e.z = gw > 0.0f ? logf(__fdiv_rn(gw, d.z)) : gw;     // ssd.py:269
  e.w = gh > 0.0f ? logf(__fdiv_rn(gh, d.w)) : gh;     // ssd.py:270
  out[i] = e;
}

// ------------------------------------------------------------------------------------------------
// L3  src/model/ssd.py:274-283
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float smooth_l1(float x) {
  const float ax = fabsf(x);
  return ax < 1.0f ? __fmul_rn(__fmul_rn(0.5f, x), x) : __fsub_rn(ax, 0.5f);
}

__global__ void __launch_bounds__(256) smooth_l1_kernel(const float* __restrict__ x, float* __restrict__ out, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = smooth_l1(x[i]);
}

// ------------------------------------------------------------------------------------------------
// L4  src/model/ssd.py:285-298   out[n, p, g] = -sum_c gt[n, g, c] * log_softmax(pr[n, p, :])[c]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
softmax_ce_kernel(const float* __restrict__ pr, int pr_row_stride, const float* __restrict__ gt, int gt_row_stride,
                  int P, int G, int C, float* __restrict__ out) {
  extern __shared__ float s_w[];   // [G, C] class weights of this image
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < G * C; i += blockDim.x)
    s_w[i] = gt[(static_cast<size_t>(n) * G + i / C) * gt_row_stride + i % C];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float* x = pr + (static_cast<size_t>(n) * P + p) * pr_row_stride;
  float mx = x[0];
  for (int c = 1; c < C; ++c) mx = fmaxf(mx, x[c]);
  float s = 0.0f;
  for (int c = 0; c < C; ++c) s += expf(x[c] - mx);
  const float ls = logf(s);
  for (int g = 0; g < G; ++g) {
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) acc += s_w[g * C + c] * ((x[c] - mx) - ls);
    out[(static_cast<size_t>(n) * P + p) * G + g] = -acc;
  }
}

// ------------------------------------------------------------------------------------------------
// L5  src/model/ssd.py:300-311
// ------------------------------------------------------------------------------------------------
__global__ void split_pos_neg_kernel(const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
                                     int64_t* __restrict__ pos_out, int64_t* __restrict__ neg_out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = pos[i], q = neg[i];
  const bool crowded = p * 3 > q;
  // floor division like torch's `//` on int64 (neg is never negative on this path, kept general)
  int64_t third = q / 3;
  if ((q % 3 != 0) && (q < 0)) --third;
  pos_out[i] = crowded ? third : p;
  neg_out[i] = crowded ? q : p * 3;
}

// ------------------------------------------------------------------------------------------------
// L6  src/model/ssd.py:313-328  (k+1)-th largest by 4-pass 8-bit radix select; one block per row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
kplus1_kernel(const float* __restrict__ values, int len, const int64_t* __restrict__ k_in, float* __restrict__ out) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_remaining;
  const float* v = values + static_cast<size_t>(blockIdx.x) * len;
  long long k = k_in[blockIdx.x];
  if (k < 0) k = 0;
  if (k > len - 1) k = len - 1;
  if (threadIdx.x == 0) { s_prefix = 0; s_remaining = static_cast<unsigned int>(k); }
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const unsigned int prefix = s_prefix;
    const unsigned int himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      const unsigned int key = float_key(v[i]);
      if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int rem = s_remaining;    // number of elements strictly above the answer still to skip
      int b = 255;
      for (; b > 0; --b) {
        if (rem < hist[b]) break;
        rem -= hist[b];
      }
      s_remaining = rem;
      s_prefix = prefix | (static_cast<unsigned int>(b) << shift);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = key_float(s_prefix);
}

__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ x, size_t n, const float* __restrict__ scale) {
  const float s = *scale;
  if (s == 1.0f) return;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= s;
}

// L2 prefetch: one thread per 8 KB chunk, a handful of CTAs -- leaves the SMs to whatever else is running.
constexpr size_t kPrefetchChunk = 8192;
__global__ void __launch_bounds__(128) prefetch_l2_kernel(const unsigned char* __restrict__ ptr, size_t bytes) {
  const size_t chunk = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t off = chunk * kPrefetchChunk;
  if (off >= bytes) return;
  const unsigned int n = static_cast<unsigned int>(bytes - off < kPrefetchChunk ? ((bytes - off) & ~static_cast<size_t>(15)) : kPrefetchChunk);
  if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr + off), "r"(n) : "memory");
}

static int check_gt(const char* fn, int N, int G) {
  if (N <= 0 || G < 0) { set_error("%s: bad N=%d G=%d", fn, N, G); return SSDH_E_ARG; }
  if (G > kMaxGT) { set_error("%s: G=%d exceeds the limit of %d ground-truth rows per image", fn, G, kMaxGT); return SSDH_E_LIMIT; }
  return 0;
}

}  // namespace ssdh

using namespace ssdh;

extern "C" int ssdh_match(const float* gt, int gt_row_stride, int N, int G, const float* priors, int P, float thr,
                          uint64_t* match_bits, uint8_t* match_mask, int32_t* best_gt, float* best_iou,
                          int32_t* best_prior, float* best_prior_iou, ssdh_stream_t stream) {
  if (int e = check_gt("ssdh_match", N, G)) return e;
  if (!priors || P <= 0 || (G > 0 && !gt)) { set_error("ssdh_match: NULL input or P <= 0"); return SSDH_E_ARG; }
  if (!aligned16(priors)) { set_error("ssdh_match: priors must be 16-byte aligned"); return SSDH_E_ALIGN; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (match_bits || match_mask || best_gt || best_iou) {
    dim3 grid((P + 255) / 256, N);
    match_kernel<<<grid, 256, 0, st>>>(gt, gt_row_stride, G, reinterpret_cast<const float4*>(priors), P, make_band(thr),
                                       match_bits, match_mask, best_gt, best_iou);
    if (int e = cuda_status("ssdh_match")) return e;
  }
  if ((best_prior || best_prior_iou) && G > 0) {
    best_prior_kernel<<<dim3(G, N), 256, 0, st>>>(gt, gt_row_stride, G, reinterpret_cast<const float4*>(priors), P,
                                                  best_prior, best_prior_iou);
    if (int e = cuda_status("ssdh_match(best_prior)")) return e;
  }
  return 0;
}

extern "C" int ssdh_encode(const float* gt, int gt_row_stride, int N, int G, const float* priors, int P, float* out,
                           ssdh_stream_t stream) {
  if (N <= 0 || G <= 0 || P <= 0 || !gt || !priors || !out) { set_error("ssdh_encode: bad argument"); return SSDH_E_ARG; }
  if (!aligned16(priors) || !aligned16(out)) { set_error("ssdh_encode: priors/out must be 16-byte aligned"); return SSDH_E_ALIGN; }
  const size_t total = static_cast<size_t>(N) * P * G;
  encode_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      gt, gt_row_stride, G, reinterpret_cast<const float4*>(priors), P, reinterpret_cast<float4*>(out), total);
  return cuda_status("ssdh_encode");
}

extern "C" int ssdh_smooth_l1(const float* x, float* out, size_t n, ssdh_stream_t stream) {
  if (!x || !out) { set_error("ssdh_smooth_l1: NULL"); return SSDH_E_ARG; }
  if (n == 0) return 0;
  const unsigned blocks = static_cast<unsigned>(n / 256 + 1 < 148 * 16 ? n / 256 + 1 : 148 * 16);
  smooth_l1_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, n);
  return cuda_status("ssdh_smooth_l1");
}

extern "C" int ssdh_softmax_cross_entropy(const float* pr, int pr_row_stride, const float* gt, int gt_row_stride,
                                          int N, int P, int G, int C, float* out, ssdh_stream_t stream) {
  if (N <= 0 || P <= 0 || G <= 0 || C <= 0 || !pr || !gt || !out) { set_error("ssdh_softmax_cross_entropy: bad argument"); return SSDH_E_ARG; }
  if (C > kMaxClasses || G > 1024) { set_error("ssdh_softmax_cross_entropy: C or G above limit"); return SSDH_E_LIMIT; }
  const size_t smem = static_cast<size_t>(G) * C * sizeof(float);
  if (smem > 48 * 1024) { set_error("ssdh_softmax_cross_entropy: G*C too large"); return SSDH_E_LIMIT; }
  softmax_ce_kernel<<<dim3((P + 127) / 128, N), 128, smem, static_cast<cudaStream_t>(stream)>>>(pr, pr_row_stride, gt, gt_row_stride, P, G, C, out);
  return cuda_status("ssdh_softmax_cross_entropy");
}

extern "C" int ssdh_split_pos_neg(const int64_t* pos, const int64_t* neg, int64_t* pos_out, int64_t* neg_out, int n,
                                  ssdh_stream_t stream) {
  if (!pos || !neg || !pos_out || !neg_out || n < 0) { set_error("ssdh_split_pos_neg: bad argument"); return SSDH_E_ARG; }
  if (n == 0) return 0;
  split_pos_neg_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(pos, neg, pos_out, neg_out, n);
  return cuda_status("ssdh_split_pos_neg");
}

extern "C" int ssdh_kplus1_value(const float* values, int rows, int len, const int64_t* k, float* out, ssdh_stream_t stream) {
  if (!values || !k || !out || rows <= 0 || len <= 0) { set_error("ssdh_kplus1_value: bad argument"); return SSDH_E_ARG; }
  kplus1_kernel<<<rows, 512, 0, static_cast<cudaStream_t>(stream)>>>(values, len, k, out);
  return cuda_status("ssdh_kplus1_value");
}

extern "C" int ssdh_prefetch_l2(const void* ptr, size_t bytes, ssdh_stream_t stream) {
  if (!ptr) { set_error("ssdh_prefetch_l2: NULL"); return SSDH_E_ARG; }
  if (!aligned16(ptr)) { set_error("ssdh_prefetch_l2: ptr must be 16-byte aligned"); return SSDH_E_ALIGN; }
  if (bytes < 16) return 0;
  const size_t chunks = (bytes + kPrefetchChunk - 1) / kPrefetchChunk;
  prefetch_l2_kernel<<<static_cast<unsigned>((chunks + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const unsigned char*>(ptr), bytes);
  return cuda_status("ssdh_prefetch_l2");
}

extern "C" int ssdh_scale_inplace(float* x, size_t n, const float* scale, ssdh_stream_t stream) {
  if (!x || !scale) { set_error("ssdh_scale_inplace: NULL"); return SSDH_E_ARG; }
  if (n == 0) return 0;
  const unsigned blocks = static_cast<unsigned>(n / 1024 + 1 < 148 * 8 ? n / 1024 + 1 : 148 * 8);
  scale_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, scale);
  return cuda_status("ssdh_scale_inplace");
}
