// Fused MultiBox loss: C ABI entry points and the production / instrumented instantiations of multibox_loss_kernel
// (loss_kernel.cuh).  Replaces SSD.loss and everything it calls (reference src/model/ssd.py:181-328) plus the autograd
// backward of that graph (src/train.py:121).
#include "loss_kernel.cuh"

namespace ssdh {
static unsigned long long* g_loss_trace = nullptr;
}

using namespace ssdh;

extern "C" size_t ssdh_multibox_loss_workspace_bytes(int N, int P, int C, int G) {
  (void)P; (void)C; (void)G;
  return 16 + static_cast<size_t>(N > 0 ? N : 0) * sizeof(ImageSlot);
}

static int multibox_loss_impl(const float* outputs, const float* targets, const float* priors, int N, int P, int C, int G,
                              float a, float thr, int n_global, float* loss, float* grad, ssdh_image_stats* stats,
                              void* ws, size_t ws_bytes, ssdh_stream_t stream, const ssdh_loss_options& opt) {
  const float* next_outputs = opt.next_outputs;
  const float* next_targets = opt.next_targets;
  if (!outputs || !priors || !loss || N <= 0 || P <= 0 || C <= 0 || G < 0 || n_global <= 0 || (G > 0 && !targets)) {
    set_error("ssdh_multibox_loss: NULL pointer or non-positive dimension");
    return SSDH_E_ARG;
  }
  if (C > kMaxClasses || G > kMaxGT || P > 65535) {
    set_error("ssdh_multibox_loss: limits are C <= %d, G <= %d, P <= 65535 (got C=%d G=%d P=%d)", kMaxClasses, kMaxGT, C, G, P);
    return SSDH_E_LIMIT;
  }
  if (!ws || ws_bytes < ssdh_multibox_loss_workspace_bytes(N, P, C, G)) { set_error("ssdh_multibox_loss: workspace too small"); return SSDH_E_WORKSPACE; }
  if (!aligned16(priors) || !aligned16(ws)) { set_error("ssdh_multibox_loss: priors and ws must be 16-byte aligned"); return SSDH_E_ALIGN; }
  const int row = 4 + C;
  LossShape shape;
  if (!pick_shape(P, C, G, &shape)) {
    set_error("ssdh_multibox_loss: one image's [P=%d, %d] slab does not fit a cluster's shared memory (8 x ~220 KB, %d rows per CTA)",
              P, row, kSlots * 384);
    return SSDH_E_LIMIT;
  }

  LossParams p;
  p.outputs = outputs; p.targets = targets; p.priors = reinterpret_cast<const float4*>(priors);
  p.next_outputs = (next_outputs && aligned16(next_outputs)) ? next_outputs : nullptr;
  p.next_targets = next_targets;
  p.N = N; p.P = P; p.C = C; p.G = G;
  p.a = a; p.band = make_band(thr);
  p.inv_n_global = 1.0f / static_cast<float>(n_global);
  p.loss = loss; p.grad = grad; p.stats = stats;
  p.ticket = reinterpret_cast<unsigned int*>(ws);
  p.slots = reinterpret_cast<ImageSlot*>(static_cast<unsigned char*>(ws) + 16);
  p.rows_per_cta = shape.rows_per_cta;
  p.n_chunks = chunks_for(P);
  // TMA bulk copies need 16-byte aligned addresses and sizes for every (image, CTA, slot) chunk.
  const bool sizes_ok = (static_cast<long long>(P) * row) % 4 == 0;     // full blocks are 32 rows; only the tail block can be odd
  p.bulk = sizes_ok && aligned16(outputs) && (grad == nullptr || aligned16(grad));
  p.early_inputs = opt.inputs_stable ? 1 : 0;
  p.trace = g_loss_trace;
  p.ce_override = opt.ce_override;
  memset(&p.xchg, 0, sizeof(p.xchg));
  if (opt.exchange) {
    const ssdh_scalar_exchange& x = *opt.exchange;
    if (x.world <= 0 || x.world > SSDH_MAX_RANKS || x.world > 32 || x.rank < 0 || x.rank >= x.world || x.ring == 0 || !x.counters) {
      set_error("ssdh_multibox_loss_ex: malformed exchange (world %d, rank %d)", x.world, x.rank);
      return SSDH_E_ARG;
    }
    for (int r = 0; r < x.world; ++r)
      if (!x.inbox[r]) { set_error("ssdh_multibox_loss_ex: exchange inbox[%d] is NULL", r); return SSDH_E_ARG; }
    p.xchg = x;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int mode = (opt.force_best_prior ? kModeForce : 0) | (opt.exact_math ? kModeExact : 0);
  if (opt.ce_override && !opt.exact_math) { set_error("ssdh_multibox_loss_ex: ce_override is a hook of the exact_math test mode"); return SSDH_E_ARG; }
  if (mode != 0) return launch_loss_extension(mode, p, shape, st);      // opt-in modes: loss_ext.cu
  // the instrumented build of the kernel (debug hook, see the bottom of this file) is a separate instantiation: the
  // production kernel carries no trace code at all
  if (p.trace != nullptr) return launch_loss_mode<kModeTrace>(p, shape, st);
  return launch_loss_mode<0>(p, shape, st);
}

static ssdh_loss_options default_options() {
  ssdh_loss_options o = {};
  o.struct_bytes = static_cast<uint32_t>(sizeof(ssdh_loss_options));
  return o;
}

extern "C" int ssdh_multibox_loss(const float* outputs, const float* targets, const float* priors, int N, int P, int C, int G,
                                  float a, float thr, int n_global, float* loss, float* grad, ssdh_image_stats* stats,
                                  void* ws, size_t ws_bytes, ssdh_stream_t stream) {
  return multibox_loss_impl(outputs, targets, priors, N, P, C, G, a, thr, n_global, loss, grad, stats, ws, ws_bytes, stream, default_options());
}

extern "C" int ssdh_multibox_loss_pipelined(const float* outputs, const float* targets, const float* priors, int N, int P, int C, int G,
                                            float a, float thr, int n_global, float* loss, float* grad, ssdh_image_stats* stats,
                                            void* ws, size_t ws_bytes, ssdh_stream_t stream, const float* next_outputs,
                                            const float* next_targets) {
  ssdh_loss_options o = default_options();
  o.inputs_stable = 1;
  o.next_outputs = next_outputs;
  o.next_targets = next_targets;
  return multibox_loss_impl(outputs, targets, priors, N, P, C, G, a, thr, n_global, loss, grad, stats, ws, ws_bytes, stream, o);
}

extern "C" int ssdh_multibox_loss_ex(const float* outputs, const float* targets, const float* priors, int N, int P, int C, int G,
                                     float a, float thr, int n_global, float* loss, float* grad, ssdh_image_stats* stats,
                                     void* ws, size_t ws_bytes, ssdh_stream_t stream, const ssdh_loss_options* options) {
  if (!options) return multibox_loss_impl(outputs, targets, priors, N, P, C, G, a, thr, n_global, loss, grad, stats, ws, ws_bytes, stream, default_options());
  if (options->struct_bytes != sizeof(ssdh_loss_options)) {
    set_error("ssdh_multibox_loss_ex: options->struct_bytes is %u, this build expects %zu", options->struct_bytes, sizeof(ssdh_loss_options));
    return SSDH_E_ARG;
  }
  return multibox_loss_impl(outputs, targets, priors, N, P, C, G, a, thr, n_global, loss, grad, stats, ws, ws_bytes, stream, *options);
}

// Debug hook (not part of the public header): device buffer of [N * 8][16] u64 phase stamps, or NULL to disable.
extern "C" __attribute__((visibility("default"))) void ssdh_debug_set_loss_trace(unsigned long long* buf) { g_loss_trace = buf; }

extern "C" int ssdh_device_info(int* sm_count, int* max_smem_optin, int* loss_cluster_size, int* loss_max_active_clusters) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("ssdh_device_info: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { set_error("ssdh_device_info: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (max_smem_optin) *max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  LossShape shape;
  pick_shape(SSDH_NUM_PRIORS, 21, 20, &shape);
  if (loss_cluster_size) *loss_cluster_size = shape.cluster;
  if (loss_max_active_clusters) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(shape.cluster * 64);
    cfg.blockDim = dim3(shape.threads);
    cfg.dynamicSmemBytes = shape.smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = shape.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int nc = 0;
    if (shape.cluster == 4) {
      ensure_dyn_smem(reinterpret_cast<const void*>(multibox_loss_kernel<21, 768, 4, 0>), static_cast<int>(kMaxDynSmem), "ssdh_device_info");
      e = cudaOccupancyMaxActiveClusters(&nc, multibox_loss_kernel<21, 768, 4, 0>, &cfg);
    } else {
      ensure_dyn_smem(reinterpret_cast<const void*>(multibox_loss_kernel<21, 384, 8, 0>), static_cast<int>(kMaxDynSmem), "ssdh_device_info");
      e = cudaOccupancyMaxActiveClusters(&nc, multibox_loss_kernel<21, 384, 8, 0>, &cfg);
    }
    if (e != cudaSuccess) { set_error("ssdh_device_info: occupancy: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return static_cast<int>(e); }
    *loss_max_active_clusters = nc;
  }
  return 0;
}
