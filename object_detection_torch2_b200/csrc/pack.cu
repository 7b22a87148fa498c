// head-output producer.  Replaces the tail of SSD.forward (reference src/model/ssd.py:96-104): every
// detector output x_k (N, A_k * (4 + C), H_k, W_k) goes through permute(0, 2, 3, 1).reshape(N, -1, 4 + C) and the six
// results are concatenated along dim 1 into (N, 8732, 4 + C) -- seven full copies of the data in the reference.
//
// Per (level, image) that is a plain 2-D transpose: the input is a [CH = A * (4 + C)] x [HW] matrix (HW contiguous), the
// output block a [HW] x [CH] matrix (row (i * W + j) * A + a, column c  <->  channel a * (4 + C) + c at cell (i, j)),
// stored at row offset off_k of the image's slab.  One launch moves all levels of all images through shared-memory tiles
// of 32 cells x all channels: coalesced reads along HW, one contiguous block written per tile, every byte read once and
// written once.  The same kernel run
// backwards (kUnpack) scatters d loss / d outputs into the detectors' NCHW gradients for autograd.
#include "common.cuh"

namespace ssdh {

constexpr int kPackMaxLevels = 8;
constexpr int kTile = 32;

struct PackParams {
  float* level[kPackMaxLevels];      // NCHW tensors, (N, ch, hw) each
  int ch[kPackMaxLevels], hw[kPackMaxLevels];
  int row_off[kPackMaxLevels];       // first slab row of the level (in rows of `width` floats)
  int tile_start[kPackMaxLevels + 1];   // prefix sum of tiles (32 cells each) per image
  int n_levels, N, width, P;
  float* slab;                       // (N, P, width)
};

// One CTA = 32 consecutive cells (hw) of one level of one image, ALL channels: the input side is ch rows of 32 floats
// (one 128-byte request per warp and channel), the output side one contiguous block of 32 * ch floats.
template <bool kUnpack>
__global__ void __launch_bounds__(256) pack_head_kernel(const PackParams p) {
  extern __shared__ float tile[];                                   // [ch][33]
  const int n = blockIdx.y;
  int t = blockIdx.x, l = 0;
#pragma unroll
  for (int q = 1; q < kPackMaxLevels; ++q) l += (q < p.n_levels && t >= p.tile_start[q]) ? 1 : 0;
  t -= p.tile_start[l];
  const int ch = p.ch[l], hw = p.hw[l];
  const int hw0 = t * kTile, nh = min(kTile, hw - hw0);
  float* lev = p.level[l] + static_cast<size_t>(n) * ch * hw + hw0;                                  // [ch][hw], at cell hw0
  float* out = p.slab + (static_cast<size_t>(n) * p.P + p.row_off[l]) * p.width + static_cast<size_t>(hw0) * ch;   // [nh][ch]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // The kernel used to be bound by instruction issue, not by memory (ncu: 77 % issue-active, ~23 instructions per element).
  // NCHW side: when the level's rows are whole float4 groups (H*W a multiple of 4: 66 % + 7 % of an SSD300 head) a lane moves
  // four consecutive cells of one channel per request (eight lanes per 128-byte channel row, four channels per warp request);
  // the shared-memory side of those four words is conflict-free in the [ch][33] layout (bank = channel + cell mod 32).
  // Slab side: the tile's block is contiguous, [nh][ch]; a warp walks one row of it with running pointers.
  const bool vec = (hw & 3) == 0 && (reinterpret_cast<uintptr_t>(lev) & 15u) == 0;
  const int q = lane & 7, cs = lane >> 3;
  if (!kUnpack) {
    if (vec) {
      const int nq = nh >> 2;
      const float* src = lev + static_cast<size_t>(warp * 4 + cs) * hw + 4 * q;
      float* dst = tile + (warp * 4 + cs) * (kTile + 1) + 4 * q;
#pragma unroll 2
      for (int c = warp * 4 + cs; c < ch; c += 32, src += static_cast<size_t>(32) * hw, dst += 32 * (kTile + 1)) {
        if (q < nq) {
          const float4 v = *reinterpret_cast<const float4*>(src);
          dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
      }
    } else {
      for (int c = warp; c < ch; c += 8)
        if (lane < nh) tile[c * (kTile + 1) + lane] = lev[static_cast<size_t>(c) * hw + lane];
    }
    __syncthreads();
    for (int h = warp; h < nh; h += 8) {
      float* orow = out + static_cast<size_t>(h) * ch;
      const float* trow = tile + h;
#pragma unroll 4
      for (int c = lane; c < ch; c += 32) orow[c] = trow[c * (kTile + 1)];
    }
  } else {
    for (int h = warp; h < nh; h += 8) {
      const float* orow = out + static_cast<size_t>(h) * ch;
      float* trow = tile + h;
#pragma unroll 4
      for (int c = lane; c < ch; c += 32) trow[c * (kTile + 1)] = orow[c];
    }
    __syncthreads();
    if (vec) {
      const int nq = nh >> 2;
      float* dstg = lev + static_cast<size_t>(warp * 4 + cs) * hw + 4 * q;
      const float* srct = tile + (warp * 4 + cs) * (kTile + 1) + 4 * q;
#pragma unroll 2
      for (int c = warp * 4 + cs; c < ch; c += 32, dstg += static_cast<size_t>(32) * hw, srct += 32 * (kTile + 1)) {
        if (q < nq) *reinterpret_cast<float4*>(dstg) = make_float4(srct[0], srct[1], srct[2], srct[3]);
      }
    } else {
      for (int c = warp; c < ch; c += 8)
        if (lane < nh) lev[static_cast<size_t>(c) * hw + lane] = tile[c * (kTile + 1) + lane];
    }
  }
}

static int run_pack(float* const* levels, const int* ch, const int* hw, int n_levels, int N, int width, float* slab, int P,
                    bool unpack, ssdh_stream_t stream, const char* fn) {
  if (!levels || !ch || !hw || !slab || n_levels <= 0 || N <= 0 || width <= 0 || P <= 0) { set_error("%s: NULL pointer or non-positive dimension", fn); return SSDH_E_ARG; }
  if (n_levels > kPackMaxLevels) { set_error("%s: at most %d levels (got %d)", fn, kPackMaxLevels, n_levels); return SSDH_E_LIMIT; }
  if (N > 65535) { set_error("%s: N <= 65535", fn); return SSDH_E_LIMIT; }
  PackParams p = {};
  long long rows = 0;
  int tiles = 0, ch_max = 0;
  for (int l = 0; l < n_levels; ++l) {
    if (!levels[l] || ch[l] <= 0 || hw[l] <= 0 || ch[l] % width != 0) {
      set_error("%s: level %d: NULL pointer, empty shape or channel count %d not a multiple of the row width %d", fn, l, ch[l], width);
      return SSDH_E_ARG;
    }
    p.level[l] = levels[l]; p.ch[l] = ch[l]; p.hw[l] = hw[l];
    p.row_off[l] = static_cast<int>(rows);
    rows += static_cast<long long>(hw[l]) * (ch[l] / width);
    p.tile_start[l] = tiles;
    tiles += (hw[l] + kTile - 1) / kTile;
    ch_max = ch[l] > ch_max ? ch[l] : ch_max;
  }
  p.tile_start[n_levels] = tiles;
  if (rows != P) { set_error("%s: the levels hold %lld rows, the slab %d", fn, rows, P); return SSDH_E_ARG; }
  const size_t smem = static_cast<size_t>(ch_max) * (kTile + 1) * sizeof(float);
  if (smem > 200 * 1024) { set_error("%s: at most %d channels per level (got %d)", fn, 200 * 1024 / ((kTile + 1) * 4), ch_max); return SSDH_E_LIMIT; }
  p.n_levels = n_levels; p.N = N; p.width = width; p.P = P; p.slab = slab;
  const dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(N));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (unpack) {
    if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(pack_head_kernel<true>), 200 * 1024, fn)) return e;
    pack_head_kernel<true><<<grid, 256, smem, st>>>(p);
  } else {
    if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(pack_head_kernel<false>), 200 * 1024, fn)) return e;
    pack_head_kernel<false><<<grid, 256, smem, st>>>(p);
  }
  return cuda_status(fn);
}

// Channels-last producers.  A detector output in torch.channels_last format is (N, H*W, ch) in memory, so every (level,
// image) block already IS the block of slab rows it belongs to: the pass is a plain copy of N * n_levels contiguous ranges,
// one CTA per 4 096 floats of a range, in the widest requests the two addresses allow (16 / 8 / 4 bytes -- the slab offset of
// a level is a multiple of 100 bytes, not of 16).
constexpr int kCopyFloats = 4096;

template <bool kUnpack>
__global__ void __launch_bounds__(256) pack_head_nhwc_kernel(const PackParams p) {
  const int n = blockIdx.y;
  int t = blockIdx.x, l = 0;
#pragma unroll
  for (int q = 1; q < kPackMaxLevels; ++q) l += (q < p.n_levels && t >= p.tile_start[q]) ? 1 : 0;
  t -= p.tile_start[l];
  const size_t block = static_cast<size_t>(p.ch[l]) * p.hw[l];                      // floats of one (level, image) block
  const size_t off = static_cast<size_t>(t) * kCopyFloats;
  const int count = static_cast<int>(block - off < kCopyFloats ? block - off : kCopyFloats);
  float* lev = p.level[l] + static_cast<size_t>(n) * block + off;
  float* slab = p.slab + (static_cast<size_t>(n) * p.P + p.row_off[l]) * p.width + off;
  const float* src = kUnpack ? slab : lev;
  float* dst = kUnpack ? lev : slab;
  const uintptr_t both = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst);
  if ((both & 15u) == 0 && (count & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 4
    for (int i = threadIdx.x; i < count / 4; i += 256) d4[i] = s4[i];
  } else if ((both & 7u) == 0 && (count & 1) == 0) {
    const float2* s2 = reinterpret_cast<const float2*>(src);
    float2* d2 = reinterpret_cast<float2*>(dst);
#pragma unroll 4
    for (int i = threadIdx.x; i < count / 2; i += 256) d2[i] = s2[i];
  } else {
#pragma unroll 4
    for (int i = threadIdx.x; i < count; i += 256) dst[i] = src[i];
  }
}

static int run_pack_nhwc(float* const* levels, const int* ch, const int* hw, int n_levels, int N, int width, float* slab, int P,
                         bool unpack, ssdh_stream_t stream, const char* fn) {
  if (!levels || !ch || !hw || !slab || n_levels <= 0 || N <= 0 || width <= 0 || P <= 0) { set_error("%s: NULL pointer or non-positive dimension", fn); return SSDH_E_ARG; }
  if (n_levels > kPackMaxLevels) { set_error("%s: at most %d levels (got %d)", fn, kPackMaxLevels, n_levels); return SSDH_E_LIMIT; }
  if (N > 65535) { set_error("%s: N <= 65535", fn); return SSDH_E_LIMIT; }
  PackParams p = {};
  long long rows = 0;
  int tiles = 0;
  for (int l = 0; l < n_levels; ++l) {
    if (!levels[l] || ch[l] <= 0 || hw[l] <= 0 || ch[l] % width != 0) {
      set_error("%s: level %d: NULL pointer, empty shape or channel count %d not a multiple of the row width %d", fn, l, ch[l], width);
      return SSDH_E_ARG;
    }
    p.level[l] = levels[l]; p.ch[l] = ch[l]; p.hw[l] = hw[l];
    p.row_off[l] = static_cast<int>(rows);
    rows += static_cast<long long>(hw[l]) * (ch[l] / width);
    p.tile_start[l] = tiles;
    tiles += static_cast<int>((static_cast<long long>(hw[l]) * ch[l] + kCopyFloats - 1) / kCopyFloats);
  }
  p.tile_start[n_levels] = tiles;
  if (rows != P) { set_error("%s: the levels hold %lld rows, the slab %d", fn, rows, P); return SSDH_E_ARG; }
  p.n_levels = n_levels; p.N = N; p.width = width; p.P = P; p.slab = slab;
  const dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(N));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (unpack) pack_head_nhwc_kernel<true><<<grid, 256, 0, st>>>(p);
  else pack_head_nhwc_kernel<false><<<grid, 256, 0, st>>>(p);
  return cuda_status(fn);
}

}  // namespace ssdh

using namespace ssdh;

extern "C" int ssdh_pack_head_nhwc(const float* const* levels, const int* ch, const int* hw, int n_levels, int N, int width,
                                   float* outputs, int P, ssdh_stream_t stream) {
  return run_pack_nhwc(const_cast<float* const*>(reinterpret_cast<const float* const*>(levels)), ch, hw, n_levels, N, width, outputs, P, false, stream, "ssdh_pack_head_nhwc");
}

extern "C" int ssdh_unpack_head_nhwc(const float* grad_outputs, float* const* level_grads, const int* ch, const int* hw, int n_levels,
                                     int N, int width, int P, ssdh_stream_t stream) {
  return run_pack_nhwc(level_grads, ch, hw, n_levels, N, width, const_cast<float*>(grad_outputs), P, true, stream, "ssdh_unpack_head_nhwc");
}

extern "C" int ssdh_pack_head(const float* const* levels, const int* ch, const int* hw, int n_levels, int N, int width,
                              float* outputs, int P, ssdh_stream_t stream) {
  return run_pack(const_cast<float* const*>(reinterpret_cast<const float* const*>(levels)), ch, hw, n_levels, N, width, outputs, P, false, stream, "ssdh_pack_head");
}

extern "C" int ssdh_unpack_head(const float* grad_outputs, float* const* level_grads, const int* ch, const int* hw, int n_levels, int N,
                                int width, int P, ssdh_stream_t stream) {
  return run_pack(level_grads, ch, hw, n_levels, N, width, const_cast<float*>(grad_outputs), P, true, stream, "ssdh_unpack_head");
}

// ------------------------------------------------------------------------------------------------------------------
// ground-truth ingest.  The reference ships dense one-hot rows [cx, cy, w, h, onehot(C)] zero-padded to the
// batch maximum by pad_sequence (src/utils.py:8-16): 100 bytes per row over PCIe for 5 numbers of information.  The host
// sends compact rows [cx, cy, w, h, label] (+ the per-image row counts) and this kernel expands them on the device into
// exactly the tensor collate_fn would have produced, which every other entry point takes unchanged.
namespace ssdh {

__global__ void __launch_bounds__(256) expand_targets_kernel(const float* __restrict__ compact, const int* __restrict__ lengths,
                                                             int N, int G, int C, float* __restrict__ dense) {
  const int width = 4 + C;
  const long long total = static_cast<long long>(N) * G * width;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % width);
    const long long rg = i / width;                       // row index n * G + g
    const int g = static_cast<int>(rg % G), n = static_cast<int>(rg / G);
    const bool real = lengths == nullptr || g < lengths[n];
    const float* src = compact + rg * 5;
    float v = 0.0f;
    if (real) {
      if (c < 4) v = src[c];
      else v = (static_cast<int>(src[4]) == c - 4) ? 1.0f : 0.0f;
    }
    dense[i] = v;
  }
}

}  // namespace ssdh

extern "C" int ssdh_expand_targets(const float* compact, const int* lengths, int N, int G, int C, float* targets, ssdh_stream_t stream) {
  if (N < 0 || G < 0 || C <= 0 || (static_cast<long long>(N) * G > 0 && (!compact || !targets))) {
    set_error("ssdh_expand_targets: NULL pointer or bad dimension");
    return SSDH_E_ARG;
  }
  const long long total = static_cast<long long>(N) * G * (4 + C);
  if (total == 0) return 0;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);      // 148 SMs x 8
  ssdh::expand_targets_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(compact, lengths, N, G, C, targets);
  return ssdh::cuda_status("ssdh_expand_targets");
}
