// Step-scalar exchange over NVLink: the collective that follows the fused MultiBox loss, folded into its epilogue.
//
// Sharded training needs exactly one exchange per step: the scalar  sum_local loss_i / N_global  of every rank
// (reference src/model/ssd.py:227 is the batch mean; the gradient needs no reduced value).  A 4-byte all-reduce next to a
// 17 us kernel is pure launch latency, so instead of launching one the loss kernel's LAST CTA stores its scalar straight
// into every peer's inbox (peer-mapped device memory, one 64-bit store per peer over NVLink: value + step number in one
// word, so a reader can never see one without the other), and a tiny reduce kernel -- run once per N steps, off the step's
// critical path -- waits for the words of the steps it is asked for and adds them in rank order (bit-identical on every rank).
//
// The inbox is the one allocation this library makes (cudaMalloc + cudaIpcGetMemHandle need an allocation of their own);
// it is an explicit create / open / close / destroy resource, everything else stays caller-owned.
#include "common.cuh"

namespace ssdh {

__global__ void __launch_bounds__(64) exchange_reduce_kernel(ssdh_scalar_exchange x, int count, float* __restrict__ out, int* __restrict__ status) {
  const uint32_t base = x.counters[1];
  const unsigned long long* inbox = x.inbox[x.rank];
  for (int k = threadIdx.x; k < count; k += blockDim.x) {
    const uint32_t seq = base + static_cast<uint32_t>(k) + 1u;
    float acc = 0.0f;
    bool bad = false;
    for (int r = 0; r < x.world; ++r) {
      const unsigned long long* slot = inbox + static_cast<size_t>(r) * x.ring + ((seq - 1u) % x.ring);
      unsigned long long w = 0ull;
      long long spins = 0;
      for (;;) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(slot) : "memory");
        const uint32_t have = static_cast<uint32_t>(w >> 32);
        if (have == seq) break;
        if (static_cast<int32_t>(have - seq) > 0 || ++spins > (1ll << 26)) { bad = true; break; }     // overrun (ring too small) or a peer that never arrives
        __nanosleep(64);
      }
      acc += __uint_as_float(static_cast<uint32_t>(w));
    }
    out[k] = bad ? __int_as_float(0x7fc00000) : acc;
    if (bad && status) atomicExch(status, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) x.counters[1] = base + static_cast<uint32_t>(count);
}

}  // namespace ssdh

using namespace ssdh;

extern "C" size_t ssdh_scalar_exchange_bytes(int world) {
  if (world <= 0 || world > SSDH_MAX_RANKS) return 0;
  return static_cast<size_t>(world) * SSDH_XCHG_RING * sizeof(unsigned long long) + 64;      // inbox words + the two local counters
}

extern "C" int ssdh_scalar_exchange_create(int world, void** inbox, ssdh_ipc_handle* handle) {
  if (!inbox || !handle || world <= 0 || world > SSDH_MAX_RANKS) { set_error("ssdh_scalar_exchange_create: bad argument"); return SSDH_E_ARG; }
  void* p = nullptr;
  const size_t bytes = ssdh_scalar_exchange_bytes(world);
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("ssdh_scalar_exchange_create: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    if (p) cudaFree(p);
    return static_cast<int>(e);
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(ssdh_ipc_handle), "handle size");
  memcpy(handle, &h, sizeof(h));
  *inbox = p;
  return 0;
}

extern "C" int ssdh_scalar_exchange_open(const ssdh_ipc_handle* handle, void** peer_inbox) {
  if (!handle || !peer_inbox) { set_error("ssdh_scalar_exchange_open: bad argument"); return SSDH_E_ARG; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  const cudaError_t e = cudaIpcOpenMemHandle(peer_inbox, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("ssdh_scalar_exchange_open: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return static_cast<int>(e); }
  return 0;
}

extern "C" int ssdh_scalar_exchange_close(void* peer_inbox) {
  if (!peer_inbox) return 0;
  const cudaError_t e = cudaIpcCloseMemHandle(peer_inbox);
  if (e != cudaSuccess) { set_error("ssdh_scalar_exchange_close: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return static_cast<int>(e); }
  return 0;
}

extern "C" int ssdh_scalar_exchange_destroy(void* inbox) {
  if (!inbox) return 0;
  const cudaError_t e = cudaFree(inbox);
  if (e != cudaSuccess) { set_error("ssdh_scalar_exchange_destroy: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return static_cast<int>(e); }
  return 0;
}

extern "C" int ssdh_scalar_exchange_reduce(const ssdh_scalar_exchange* x, int count, float* out, int* status, ssdh_stream_t stream) {
  if (!x || !out || count <= 0 || x->world <= 0 || x->world > SSDH_MAX_RANKS || x->rank < 0 || x->rank >= x->world || !x->counters || x->ring == 0) {
    set_error("ssdh_scalar_exchange_reduce: bad argument");
    return SSDH_E_ARG;
  }
  if (static_cast<uint32_t>(count) > x->ring / 2) { set_error("ssdh_scalar_exchange_reduce: count must be <= ring / 2 (%u)", x->ring / 2); return SSDH_E_LIMIT; }
  for (int r = 0; r < x->world; ++r)
    if (!x->inbox[r]) { set_error("ssdh_scalar_exchange_reduce: inbox[%d] is NULL", r); return SSDH_E_ARG; }
  exchange_reduce_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(*x, count, out, status);
  return cuda_status("ssdh_scalar_exchange_reduce");
}
