// Probe: which (lane, column) lands in which thread / register for tcgen05.ld shapes .16x128b and .16x256b.
// TMEM lane L, column c is filled with L*100 + c through tcgen05.st.32x32b; then one warp reads it back.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(float* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  {  // every warp fills its own 32 lanes, columns 0..7
    uint32_t r[8];
    for (int c = 0; c < 8; ++c) r[c] = __float_as_uint((float)((32 * warp + lane) * 100 + c));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(base + ((uint32_t)(32 * warp) << 16)),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 1) {       // quadrant 1: lanes 32..63
    uint32_t a[2], b[2], c[4], d[4];
    const uint32_t q = base + ((uint32_t)32 << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(a[0]), "=r"(a[1]) : "r"(q) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(b[0]), "=r"(b[1]) : "r"(q + ((uint32_t)16 << 16) + 4) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]) : "r"(q) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(q + ((uint32_t)16 << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float* o = out + lane * 12;
    o[0] = __uint_as_float(a[0]); o[1] = __uint_as_float(a[1]); o[2] = __uint_as_float(b[0]); o[3] = __uint_as_float(b[1]);
    for (int i = 0; i < 4; ++i) { o[4 + i] = __uint_as_float(c[i]); o[8 + i] = __uint_as_float(d[i]); }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base) : "memory");
}
int main() {
  float* d; cudaMalloc(&d, 32 * 12 * 4);
  probe<<<1, 128>>>(d);
  float h[32 * 12];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("status %s\n", cudaGetErrorString(e));
  printf("value = lane*100 + col.  16x128b @lane32,col0 | 16x128b @lane48,col4 | 16x256b @lane32 | 16x256b @lane48\n");
  for (int t = 0; t < 32; ++t) {
    printf("t%2d:", t);
    for (int i = 0; i < 12; ++i) printf(" %6.0f%s", h[t * 12 + i], (i == 1 || i == 3 || i == 7) ? " |" : "");
    printf("\n");
  }
  return 0;
}
