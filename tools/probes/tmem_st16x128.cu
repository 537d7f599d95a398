// Probe: does tcgen05.st.16x128b accept a lane base of 16 inside the warp's 32-lane quarter, and what is its register layout?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/probe tools/probes/tmem_st16x128.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(uint32_t* out) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&base_s)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = base_s;
  // zero 16 columns of this warp's quarter
  {
    const uint32_t z = 0xdeadbeefu;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tb + ((uint32_t)(warp * 32) << 16)), "r"(z) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  for (int half = 0; half < 2; ++half) {
    // value encodes (half, thread, reg)
    uint32_t r[4];
    for (int i = 0; i < 4; ++i) r[i] = (uint32_t)(half << 16 | lane << 8 | i);
    const uint32_t ta = tb + ((uint32_t)(warp * 32 + half * 16) << 16);
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};" ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  uint32_t v[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(tb + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 16 + i] = v[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(32u) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 128 * 16 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  static uint32_t h[128 * 16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int w = 0; w < 4; ++w)
    for (int l = 0; l < 32; ++l)
      for (int c = 0; c < 8; ++c) {
        const uint32_t v = h[(w * 32 + l) * 16 + c];
        // expected: lane l = half*16 + row, row = t/4 + 8*(i&1), col c = (t%4) + 4*(i>>1)
        const int half = l / 16, row = l % 16, t = (row % 8) * 4 + (c % 4), i = (row / 8) | ((c / 4) << 1);
        const uint32_t exp = (uint32_t)(half << 16 | t << 8 | i);
        if (v != exp) { if (bad < 12) printf("warp %d lane %d col %d: got %08x expected %08x\n", w, l, c, v, exp); ++bad; }
      }
  printf("mismatches: %d of %d\n", bad, 4 * 32 * 8);
  for (int l = 0; l < 32; l += 5) { printf("lane %2d:", l); for (int c = 0; c < 10; ++c) printf(" %08x", h[l * 16 + c]); printf("\n"); }
  return 0;
}
