// Microbenchmark: cost of passing a ticket between the warps of one CTA (sm_100a).
//   mode 0: mbarrier arrive (release) -> try_wait (acquire), one mbarrier per warp, round robin
//   mode 1: every warp polls one shared word
//   mode 2: only the next warp polls (others nanosleep-free spin on their own flag) -- per-warp flag words
//   mode 3: mbarrier with a small payload of work (~W dependent ALU ops) inside the ticket
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long *b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_relaxed(unsigned long long *b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.relaxed.cta.shared::cta.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(unsigned long long *b, unsigned par) { unsigned ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(b)), "r"(par) : "memory"); return ok; }
__device__ __forceinline__ bool mbar_test(unsigned long long *b, unsigned par) { unsigned ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(b)), "r"(par) : "memory"); return ok; }

__global__ void k(int mode, int iters, int work, long long *out) {
  __shared__ unsigned long long mb[32];
  __shared__ volatile int turn;
  __shared__ volatile int flag[32 * 32];   // one 128-byte line per warp
  __shared__ int data[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x < 32) { mbar_init(&mb[threadIdx.x], 1); flag[threadIdx.x * 32] = 0; }
  if (threadIdx.x == 0) { turn = 0; data[0] = 1; }
  __syncthreads();
  const long long t0 = clock64();
  unsigned ph = 0;
  int acc = 0;
  for (int it = 0; it < iters; it++) {
    const int p = it * nw + warp;                       // my ticket number
    if (mode == 0 || mode == 3 || mode == 4 || mode == 5) {
      if (p > 0) {
        if (mode == 5) { while (!mbar_test(&mb[warp], ph)) { } }
        else { while (!mbar_try(&mb[warp], ph)) { } }
        ph ^= 1;
      }
    } else if (mode == 1) {
      while (turn != p) { }
    } else {
      while (flag[warp * 32] != it + (warp == 0 ? 0 : 1)) { }
    }
    // ---- ticket body: `work` dependent shared-memory read-modify-writes
    int v = data[0];
    for (int w = 0; w < work; w++) v = v * 3 + 1;
    acc += v;
    if (lane == 0) data[0] = v;
    __syncwarp();
    const int nxt = (warp + 1) % nw;
    if (lane == 0) {
      if (mode == 0 || mode == 3 || mode == 5) mbar_arrive(&mb[nxt]);
      else if (mode == 4) mbar_arrive_relaxed(&mb[nxt]);
      else if (mode == 1) turn = p + 1;
      else flag[nxt * 32] = it + 1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) { out[0] = clock64() - t0; out[1] = acc; }
}
int main() {
  long long *d; cudaMalloc(&d, 16);
  const int iters = 20000;
  for (int nw : {2, 4, 8, 16}) for (int mode : {0, 4, 5, 1, 2}) for (int work : {0, 50}) {
    k<<<1, 32 * nw>>>(mode, 100, work, d); cudaDeviceSynchronize();
    k<<<1, 32 * nw>>>(mode, iters, work, d);
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("nw %2d mode %d work %3d: %7.1f cycles per ticket %s\n", nw, mode, work, (double)h[0] / (iters * (double)nw), e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
