// Hardware probe (not product code): what does tcgen05.alloc.cta_group::2 return when BOTH CTAs of a pair issue it,
// and when only the leader does?  Decides who allocates in csrc/tc_pair.cuh.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_pair_probe tmem_pair_probe.cu && ./tmem_pair_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int COLS>
__global__ void __cluster_dims__(2, 1, 1) probe(uint32_t* out, int who) {
  __shared__ uint32_t slot[4];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) slot[0] = 0xdeadbeefu, slot[1] = 0xdeadbeefu;
  __syncthreads();
  const bool mine = who == 2 || (int)rank == who;      // who: 0 leader only, 1 peer only, 2 both
  if (threadIdx.x < 32 && mine) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot[0])), "r"(COLS) : "memory");
    // a second allocation shows whether the first one consumed columns on THIS SM
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot[1])), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    out[blockIdx.x * 2 + 0] = slot[0];
    out[blockIdx.x * 2 + 1] = slot[1];
  }
  __syncthreads();
  if (threadIdx.x < 32 && mine) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(slot[1]), "r"(32) : "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(slot[0]), "r"(COLS) : "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main(int argc, char** argv) {
  // one case per process (a case that blocks in tcgen05.alloc is killed by the caller's `timeout`)
  const int who = argc > 1 ? atoi(argv[1]) : 2;          // 0 leader only, 1 peer only, 2 both CTAs
  const int cols = argc > 2 ? atoi(argv[2]) : 128;
  setvbuf(stdout, nullptr, _IONBF, 0);
  uint32_t* d;
  cudaMalloc(&d, 64);
  cudaMemset(d, 0, 64);
  const char* names[3] = {"leader only", "peer only", "both CTAs"};
  printf("%s, %d+32 columns: ", names[who], cols);
  if (cols == 256) probe<256><<<2, 64>>>(d, who); else probe<128><<<2, 64>>>(d, who);
  cudaError_t e = cudaDeviceSynchronize();
  uint32_t h[4];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("rc=%d  CTA0 first=%08x second=%08x | CTA1 first=%08x second=%08x\n", (int)e, h[0], h[1], h[2], h[3]);
  return 0;
}
