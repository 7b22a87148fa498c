// Micro-benchmark: per-SM cp.async.bulk load / store bandwidth vs operation size and issuing threads.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(768, 1) k_load(const char* src, int op_bytes, int ops, int issuers, long long cta_stride, unsigned long long* cyc) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ unsigned long long bar;
  const int tid = threadIdx.x;
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  unsigned long long t0 = clock64();
  if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(op_bytes * ops) : "memory");
  __syncthreads();
  const char* base = src + (long long)blockIdx.x * cta_stride;
  // issuer i handles ops i, i+issuers, ...; issuers are lane 0 of consecutive warps
  if ((tid & 31) == 0 && (tid >> 5) < issuers) {
    for (int o = tid >> 5; o < ops; o += issuers)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + (size_t)o * op_bytes)),
                   "l"(base + (size_t)o * op_bytes * 4), "r"(op_bytes), "r"(s32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(s32(&bar)), "r"(0) : "memory");
  unsigned long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void __launch_bounds__(768, 1) k_store(char* dst, int op_bytes, int ops, int issuers, long long cta_stride, unsigned long long* cyc) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  for (int i = tid; i < op_bytes * ops / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  unsigned long long t0 = clock64();
  char* base = dst + (long long)blockIdx.x * cta_stride;
  if ((tid & 31) == 0 && (tid >> 5) < issuers) {
    for (int o = tid >> 5; o < ops; o += issuers)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)o * op_bytes * 4), "r"(s32(smem + (size_t)o * op_bytes)), "r"(op_bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  unsigned long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void __launch_bounds__(768, 1) k_ldgsts(const char* src, int bytes_total, long long cta_stride, int mode, unsigned long long* cyc) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  unsigned long long t0 = clock64();
  const char* base = src + (long long)blockIdx.x * cta_stride;
  const int chunks = bytes_total / 16;
  if (mode == 0) {
    for (int i = tid; i < chunks; i += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(smem + (size_t)i * 16)), "l"(base + (size_t)i * 16) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    // plain 16-byte loads, 8 in flight per thread, then stores to shared
    for (int i0 = tid; i0 < chunks; i0 += blockDim.x * 8) {
      int4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { int i = i0 + u * blockDim.x; if (i < chunks) v[u] = __ldcs((const int4*)(base) + i); }
#pragma unroll
      for (int u = 0; u < 8; ++u) { int i = i0 + u * blockDim.x; if (i < chunks) ((int4*)smem)[i] = v[u]; }
    }
  }
  __syncthreads();
  unsigned long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  const int ctas = 128;
  const size_t total = 220800;    // bytes per CTA
  char *a, *b; unsigned long long* cyc;
  const size_t span = (size_t)ctas * total * 4 + (1 << 20);
  cudaMalloc(&a, span); cudaMalloc(&b, span); cudaMalloc(&cyc, ctas * 8);
  cudaMemset(a, 1, span);
  char* flush; cudaMalloc(&flush, 512 << 20);
  cudaFuncSetAttribute(k_load, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total);
  cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total);
  for (int frac = 1; frac <= 4; ++frac)
  for (int mode = 0; mode < 1; ++mode) {
    unsigned long long hh[128]; float best = 1e9; unsigned long long mx = 0;
    const int tb = (int)total / 4 * frac;
    for (int rep = 0; rep < 4; ++rep) {
      cudaMemset(flush, rep, 512 << 20);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      k_ldgsts<<<ctas, 768, total>>>(a, tb, (long long)total * 4, mode, cyc);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      cudaMemcpy(hh, cyc, sizeof(hh), cudaMemcpyDeviceToHost);
      mx = 0; for (int i = 0; i < ctas; ++i) if (hh[i] > mx) mx = hh[i];
    }
    unsigned long long mn = ~0ull; for (int i = 0; i < ctas; ++i) if (hh[i] < mn) mn = hh[i];
    printf("%s %d B/CTA: kernel %.2f us, CTA cycles min %llu max %llu -> %.2f TB/s %s\n", mode ? "LDG.128 x8 + STS" : "LDGSTS 16B      ", tb, best * 1e3, mn, mx,
           (double)ctas * tb / (mx / 1.93e9) / 1e12, cudaGetErrorString(cudaGetLastError()));
  }
  int sizes[] = {1600, 3200, 6400, 12800, 36800, 73600, 220800};
  int issuers_l[] = {1, 24};
  unsigned long long h[ctas];
  for (int st = 0; st < 2; ++st)
    for (int is = 0; is < 2; ++is)
      for (int si = 0; si < 7; ++si) {
        int op = sizes[si], ops = total / op, issuers = issuers_l[is];
        float best = 1e9; unsigned long long med = 0;
        for (int rep = 0; rep < 4; ++rep) {
          cudaMemset(flush, rep, 512 << 20);
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          if (st == 0) k_load<<<ctas, 768, total>>>(a, op, ops, issuers, (long long)total * 4, cyc);
          else k_store<<<ctas, 768, total>>>(b, op, ops, issuers, (long long)total * 4, cyc);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
          cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
          unsigned long long mx = 0; for (int i = 0; i < ctas; ++i) if (h[i] > mx) mx = h[i];
          med = mx;
        }
        cudaError_t e = cudaGetLastError();
        printf("%s op=%6d B x %3d ops, %2d issuers: kernel %.2f us, max CTA cycles %llu -> %.2f TB/s (by cycles @1.93GHz) %s\n", st ? "STORE" : "LOAD ", op, ops, issuers,
               best * 1e3, med, (double)ctas * op * ops / (med / 1.93e9) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
