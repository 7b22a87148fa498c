// "Next" row 8f-4: the true PASCAL VOC average precision on the device, opt-in beside the reference's
// recall-style AP (reference src/evaluate.py:45-67 sorts the TP column independently of the scores, which collapses to
// TP / #gt; see evaluate.average_precision_from_tallies).  Input: one (score, TP flag, class) triple per detection of the
// WHOLE dataset -- what ssdh_eval_accumulate's tp_flags give, all ranks' lists concatenated -- and the per-class ground-truth
// counts of the tallies.
//
//   1. key = class << 32 | ~order_key(score): ONE stable radix sort (cub::DeviceRadixSort, a library sort -- this is the
//      metric's bookkeeping, not the hot path) ranks every class by descending score, ties in input order;
//   2. one CTA per class walks its segment BACKWARDS in 1024-wide chunks: a reverse block scan of the TP flags gives the
//      cumulative TP count of every rank (hence precision = tp / rank and recall = tp / #gt, in fp64 as the host
//      restatement), a reverse block max-scan gives the monotone precision envelope, and
//         area AP   = sum over TP ranks of envelope / #gt          (recall only moves at TP ranks)
//         11-pt AP  = mean over t = 0, 0.1 .. 1.0 of the envelope at the first rank whose recall reaches t.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace ssdh {

constexpr int kApThreads = 1024;

__global__ void __launch_bounds__(256) ap_key_kernel(const float* __restrict__ scores, const int32_t* __restrict__ cls, int D,
                                                    unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  keys[i] = (static_cast<unsigned long long>(static_cast<uint32_t>(cls[i])) << 32) | static_cast<uint32_t>(~float_key(scores[i]));
  vals[i] = static_cast<uint32_t>(i);
}

// first index whose key is >= (c << 32): one thread per class boundary (NC + 1 of them)
__global__ void ap_bounds_kernel(const unsigned long long* __restrict__ keys, int D, int NC, int32_t* __restrict__ bounds) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > NC) return;
  const unsigned long long want = static_cast<unsigned long long>(c) << 32;
  int lo = 0, hi = D;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] < want) lo = mid + 1; else hi = mid;
  }
  bounds[c] = lo;
}

struct ApShared {
  int wsum[32];
  double wmax[32];
  double acc[32];
  double pt_env[11];
  int pt_have[11];
};

__global__ void __launch_bounds__(kApThreads) ap_class_kernel(const uint32_t* __restrict__ order, const uint8_t* __restrict__ tp,
                                                            const int32_t* __restrict__ bounds, const long long* __restrict__ tallies,
                                                            int use_07, float* __restrict__ ap_out) {
  __shared__ ApShared sh;
  const int c = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s0 = bounds[c], s1 = bounds[c + 1], n = s1 - s0;
  const long long n_gt = tallies[c * 3 + 2];
  if (n_gt <= 0) {                                   // no ground truth of this class: undefined, as in the host restatement
    if (tid == 0) ap_out[c] = __int_as_float(0x7fc00000);
    return;
  }
  if (n == 0) {
    if (tid == 0) ap_out[c] = 0.0f;
    return;
  }
  if (tid < 11) { sh.pt_env[tid] = 0.0; sh.pt_have[tid] = 0; }
  // total TP of the class
  int total = 0;
  for (int i = tid; i < n; i += kApThreads) total += tp[order[s0 + i]] == 1 ? 1 : 0;
  total = warp_sum(total);
  if (lane == 0) sh.wsum[warp] = total;
  __syncthreads();
  total = 0;
  for (int w = 0; w < 32; ++w) total += sh.wsum[w];
  __syncthreads();

  const double inv_gt = 1.0 / static_cast<double>(n_gt);
  double area = 0.0;                                 // this thread's share of sum over TP ranks of the envelope
  int tp_after = 0;                                  // TPs at ranks beyond the current chunk
  double env_after = 0.0;                            // envelope carried in from the ranks beyond the current chunk
  for (int hi = n; hi > 0; hi -= kApThreads) {
    const int i = hi - 1 - tid;                      // thread 0 takes the LAST rank of the chunk: scans run towards lower ranks
    const bool have = i >= 0;
    const int flag = have && tp[order[s0 + i]] == 1 ? 1 : 0;
    // inclusive scan over threads 0..tid = TPs at ranks >= i inside the chunk
    int incl = warp_incl_scan(flag, lane);
    if (lane == 31) sh.wsum[warp] = incl;
    __syncthreads();
    int before = 0, chunk_tp = 0;
    for (int w = 0; w < 32; ++w) { const int t = sh.wsum[w]; before += w < warp ? t : 0; chunk_tp += t; }
    incl += before;
    const int tps = total - tp_after - (incl - flag);       // cumulative TP count at rank i (inclusive)
    const double prec = have ? static_cast<double>(tps) / static_cast<double>(i + 1) : 0.0;
    // inclusive max-scan over threads 0..tid = max precision at ranks >= i inside the chunk
    double env = prec;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, env, o);
      if (lane >= o) env = fmax(env, t);
    }
    if (lane == 31) sh.wmax[warp] = env;
    __syncthreads();
    double carry = env_after, chunk_max = env_after;
    for (int w = 0; w < 32; ++w) { const double t = sh.wmax[w]; if (w < warp) carry = fmax(carry, t); chunk_max = fmax(chunk_max, t); }
    env = fmax(env, carry);
    if (flag) {
      area += env;
      if (use_07) {
        // first rank whose recall reaches t = k * 0.1: recall moves from (tps - 1) / #gt to tps / #gt at this TP rank
        const double rec = static_cast<double>(tps) * inv_gt, prev = static_cast<double>(tps - 1) * inv_gt;
        for (int k = 1; k < 11; ++k) {
          const double t = static_cast<double>(k) * 0.1;
          if (rec >= t && !(prev >= t)) { sh.pt_env[k] = env; sh.pt_have[k] = 1; }
        }
      }
    }
    if (have && i == 0 && use_07) { sh.pt_env[0] = env; sh.pt_have[0] = 1; }      // recall >= 0 holds from the first rank on
    tp_after += chunk_tp;
    env_after = chunk_max;
    __syncthreads();
  }
  area = warp_sum(area);
  if (lane == 0) sh.acc[warp] = area;
  __syncthreads();
  if (tid == 0) {
    double a = 0.0;
    if (use_07) {
      for (int k = 0; k < 11; ++k) a += sh.pt_have[k] ? sh.pt_env[k] : 0.0;
      a /= 11.0;
    } else {
      for (int w = 0; w < 32; ++w) a += sh.acc[w];
      a *= inv_gt;
    }
    ap_out[c] = static_cast<float>(a);
  }
}

static size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

static size_t vocap_layout(int D, int NC, size_t* o_keys, size_t* o_vals, size_t* o_bounds, size_t* o_temp, size_t* temp_bytes) {
  size_t off = 0;
  const size_t d = static_cast<size_t>(D > 0 ? D : 1);
  o_keys[0] = off; off += align256(d * 8);
  o_keys[1] = off; off += align256(d * 8);
  o_vals[0] = off; off += align256(d * 4);
  o_vals[1] = off; off += align256(d * 4);
  *o_bounds = off; off += align256(static_cast<size_t>(NC + 1) * 4);
  size_t tb = 0;
  cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, tb, k, v, static_cast<int>(d), 0, 40);
  *temp_bytes = tb;
  *o_temp = off; off += align256(tb);
  return off;
}

}  // namespace ssdh

using namespace ssdh;

extern "C" size_t ssdh_voc_ap_workspace_bytes(int D, int NC) {
  if (D < 0 || NC <= 0) return 0;
  size_t ok[2], ov[2], ob, ot, tb;
  return vocap_layout(D, NC, ok, ov, &ob, &ot, &tb);
}

extern "C" int ssdh_voc_ap(const float* scores, const uint8_t* tp, const int32_t* cls, int D, const int64_t* tallies, int NC,
                           int use_07_metric, float* ap_out, void* ws, size_t ws_bytes, ssdh_stream_t stream) {
  if (!ap_out || !tallies || NC <= 0 || NC > 255 || D < 0 || (D > 0 && (!scores || !tp || !cls))) { set_error("ssdh_voc_ap: bad argument"); return SSDH_E_ARG; }
  size_t ok[2], ov[2], ob, ot, tb;
  const size_t need = vocap_layout(D, NC, ok, ov, &ob, &ot, &tb);
  if (!ws || ws_bytes < need) { set_error("ssdh_voc_ap: workspace too small (%zu < %zu)", ws_bytes, need); return SSDH_E_WORKSPACE; }
  if (!aligned16(ws)) { set_error("ssdh_voc_ap: ws must be 16-byte aligned"); return SSDH_E_ALIGN; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* b = static_cast<unsigned char*>(ws);
  cub::DoubleBuffer<unsigned long long> keys(reinterpret_cast<unsigned long long*>(b + ok[0]), reinterpret_cast<unsigned long long*>(b + ok[1]));
  cub::DoubleBuffer<uint32_t> vals(reinterpret_cast<uint32_t*>(b + ov[0]), reinterpret_cast<uint32_t*>(b + ov[1]));
  int32_t* bounds = reinterpret_cast<int32_t*>(b + ob);
  if (D > 0) {
    ap_key_kernel<<<(D + 255) / 256, 256, 0, st>>>(scores, cls, D, keys.Current(), vals.Current());
    if (int e = cuda_status("ssdh_voc_ap(keys)")) return e;
    const cudaError_t ce = cub::DeviceRadixSort::SortPairs(b + ot, tb, keys, vals, D, 0, 40, st);
    if (ce != cudaSuccess) { set_error("ssdh_voc_ap: sort: %s", cudaGetErrorString(ce)); (void)cudaGetLastError(); return static_cast<int>(ce); }
  }
  ap_bounds_kernel<<<1, 256, 0, st>>>(keys.Current(), D, NC, bounds);
  if (int e = cuda_status("ssdh_voc_ap(bounds)")) return e;
  ap_class_kernel<<<NC, kApThreads, 0, st>>>(vals.Current(), tp, bounds, reinterpret_cast<const long long*>(tallies), use_07_metric ? 1 : 0, ap_out);
  return cuda_status("ssdh_voc_ap");
}
