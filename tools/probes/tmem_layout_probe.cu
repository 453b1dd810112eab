// Probe of the tcgen05.ld / tcgen05.st fragment layouts used by the attention kernel (16x256b loads of S / O, 16x128b stores
// of packed P).  Writes a known pattern with the 32x32b shape (thread = lane, consecutive columns), reads it back with the
// other shapes and prints which (lane, column) every register of every thread received.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tmem_probe tools/probes/tmem_layout_probe.cu && gpurun_out/tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(uint32_t* out_ld256, uint32_t* out_st128) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  // pattern: value = (lane << 8) | column, columns [0, 32)
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) v[c] = ((uint32_t)(warp * 32 + lane) << 8) | (uint32_t)c;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
      ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
        "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]), "r"(lane_base)
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // (1) 16x256b.x2 load: lanes [lane_off, lane_off+16) x 16 columns -> 8 registers per thread
  for (int half = 0; half < 2; ++half) {
    uint32_t r[8];
    const uint32_t addr = tmem + ((uint32_t)(warp * 32 + half * 16) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out_ld256[((warp * 2 + half) * 32 + lane) * 8 + i] = r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // (2) 16x128b.x2 store into columns [64, 72): register i of thread t carries (t << 8) | i; read back with 32x32b
  for (int half = 0; half < 2; ++half) {
    const uint32_t addr = tmem + ((uint32_t)(warp * 32 + half * 16) << 16) + 64;
    uint32_t s0 = (uint32_t)(lane << 8) | 0u, s1 = (uint32_t)(lane << 8) | 1u, s2 = (uint32_t)(lane << 8) | 2u, s3 = (uint32_t)(lane << 8) | 3u;
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%4], {%0,%1,%2,%3};" ::"r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(addr) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(lane_base + 64) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out_st128[(warp * 32 + lane) * 8 + i] = r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

int main() {
  uint32_t *d1, *d2;
  cudaMalloc(&d1, 4 * 2 * 32 * 8 * 4);
  cudaMalloc(&d2, 128 * 8 * 4);
  probe<<<1, 128>>>(d1, d2);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t h1[4 * 2 * 32 * 8], h2[128 * 8];
  cudaMemcpy(h1, d1, sizeof(h1), cudaMemcpyDeviceToHost);
  cudaMemcpy(h2, d2, sizeof(h2), cudaMemcpyDeviceToHost);
  printf("== tcgen05.ld.16x256b.x2 (warp 0, lanes 0-15): thread t register i -> (lane, column)\n");
  for (int t = 0; t < 32; ++t) {
    printf("t%02d:", t);
    for (int i = 0; i < 8; ++i) printf(" r%d=(%u,%u)", i, h1[(0 * 32 + t) * 8 + i] >> 8, h1[(0 * 32 + t) * 8 + i] & 255u);
    printf("\n");
  }
  printf("== same, warp 1 second half (lanes 48-63), threads 0,1,4,31\n");
  for (int t : {0, 1, 4, 31}) {
    printf("t%02d:", t);
    for (int i = 0; i < 8; ++i) printf(" r%d=(%u,%u)", i, h1[((1 * 2 + 1) * 32 + t) * 8 + i] >> 8, h1[((1 * 2 + 1) * 32 + t) * 8 + i] & 255u);
    printf("\n");
  }
  printf("== tcgen05.st.16x128b.x2 into columns 64..71: TMEM (lane, column 64+j) <- (thread, register)\n");
  for (int l = 0; l < 32; ++l) {
    printf("lane%02d:", l);
    for (int j = 0; j < 8; ++j) printf(" c%d=(t%u,r%u)", j, h2[l * 8 + j] >> 8, h2[l * 8 + j] & 255u);
    printf("\n");
  }
  return 0;
}
