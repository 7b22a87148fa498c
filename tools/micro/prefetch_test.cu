// Does cp.async.bulk.prefetch.L2 make a following bulk load an L2 hit?  Times 128 CTAs x 220.8 KB loads cold vs prefetched.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(768, 1) k_load(const char* src, int op_bytes, int ops, long long cta_stride, unsigned long long* cyc) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ unsigned long long bar;
  const int tid = threadIdx.x;
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  unsigned long long t0 = clock64();
  if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(op_bytes * ops) : "memory");
  __syncthreads();
  const char* base = src + (long long)blockIdx.x * cta_stride;
  if ((tid & 31) == 0)
    for (int o = tid >> 5; o < ops; o += 24)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + (size_t)o * op_bytes)),
                   "l"(base + (size_t)o * op_bytes), "r"(op_bytes), "r"(s32(&bar)) : "memory");
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(s32(&bar)), "r"(0) : "memory");
  unsigned long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_prefetch(const unsigned char* ptr, size_t bytes, int chunk) {
  size_t off = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * chunk;
  if (off < bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr + off), "r"(chunk) : "memory");
}
__global__ void k_touch(const int4* ptr, size_t n16, int* sink) {   // real loads of every 16 B
  int acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) { int4 v = ptr[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678) *sink = acc;
}
int main() {
  const int ctas = 128; const size_t total = 220800;
  char *a; unsigned long long* cyc; int* sink;
  const size_t span = (size_t)ctas * total;
  cudaMalloc(&a, span); cudaMalloc(&cyc, ctas * 8); cudaMalloc(&sink, 4);
  cudaMemset(a, 1, span);
  char* flush; cudaMalloc(&flush, 512 << 20);
  cudaFuncSetAttribute(k_load, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total);
  unsigned long long h[ctas];
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemset(flush, rep, 512 << 20);
      cudaDeviceSynchronize();
      if (mode == 1) k_prefetch<<<(unsigned)((span / 8192 + 127) / 128), 128>>>((const unsigned char*)a, span, 8192);
      if (mode == 2) k_touch<<<148 * 4, 256>>>((const int4*)a, span / 16, sink);
      if (mode == 3) k_prefetch<<<(unsigned)((span / 3200 + 127) / 128), 128>>>((const unsigned char*)a, span, 3200);
      cudaDeviceSynchronize();
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      if (mode) { cudaEventRecord(e0); cudaEventRecord(e1); cudaEventSynchronize(e1); }   // let async prefetches drain a little
      for (volatile int spin = 0; spin < 20000000 && mode; ++spin) {}
      k_load<<<ctas, 768, total>>>(a, 3200, 69, (long long)total, cyc);
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      unsigned long long mx = 0, mn = ~0ull; for (int i = 0; i < ctas; ++i) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
      const char* names[] = {"cold", "after cp.async.bulk.prefetch.L2 (8 KB ops)", "after a kernel that loaded every byte", "after prefetch (3200 B ops)"};
      if (rep == 2) printf("%-45s load cycles min %llu max %llu (%s)\n", names[mode], mn, mx, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
