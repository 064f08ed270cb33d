// pdl_probe.cu -- what does griddepcontrol.wait guarantee on this driver/GPU?
// Kernel A (several waves of CTAs, each spins, then writes its slot with the iteration number) is
// followed by kernel B, launched with the programmatic-stream-serialization attribute, which waits
// (griddepcontrol.wait) and then checks that EVERY slot carries the current iteration number.
//   mode 0: A triggers (launch_dependents) at its start, no attribute on A
//   mode 1: A triggers at its start, A also launched with the attribute (chain  B(i-1) -> A(i) -> B(i))
//   mode 2: as 1, but trigger after wait in both kernels
//   mode 3: as 1, A never triggers explicitly (implicit trigger at exit)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a pdl_probe.cu -o pdl_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__global__ void __launch_bounds__(288) kA(int *slots, int it, int mode, long long spin, float *sink) {
  extern __shared__ unsigned char smem[];
  if (mode == 0 || mode == 1) { pdl_trigger(); pdl_wait(); }
  else if (mode == 2) { pdl_wait(); pdl_trigger(); }
  else { pdl_wait(); }
  const long long t0 = clock64();
  float acc = 0.f;
  while (clock64() - t0 < spin) acc += 1e-9f;
  if (threadIdx.x == 0) {
    smem[0] = (unsigned char)it;
    slots[blockIdx.x] = it;
    if (acc == 12345.f) *sink = acc;
  }
}

__global__ void __launch_bounds__(1024) kB(const int *slots, int nslots, int it, int mode, int *bad) {
  if (mode == 2) { pdl_wait(); pdl_trigger(); } else { pdl_trigger(); pdl_wait(); }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += gridDim.x * blockDim.x)
    if (slots[i] != it) atomicAdd(bad, 1);
}

// mode 5: the chain of one Arnoldi column:  A (586 x 288, 110 KB, writes slots) -> R (few CTAs of 1024 threads:
// sums the slots into h) -> G (ONE CTA of 1024 threads with dynamic shared memory: h -> h2) ->
// U (586 x 288, 110 KB: every CTA checks h2).  All four launched with the attribute, wait first, then trigger.
__global__ void __launch_bounds__(1024) kR(const int *slots, int nslots, int *h, int it) {
  pdl_wait(); pdl_trigger();
  __shared__ int sp[1024];
  int s = 0;
  for (int i = threadIdx.x; i < nslots; i += blockDim.x) s += (slots[i] == it) ? 1 : 0;
  sp[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int q = 0; q < 1024; ++q) t += sp[q];
    h[blockIdx.x] = t * 1000 + it % 1000;   // nslots*1000 + it%1000 when every slot was current
  }
}
__global__ void __launch_bounds__(1024) kG(const int *h, int nh, int *h2) {
  extern __shared__ unsigned char smem[];
  pdl_wait(); pdl_trigger();
  if ((int)threadIdx.x < nh) h2[threadIdx.x] = h[threadIdx.x];
  if (threadIdx.x == 0) smem[0] = 1;
}
__global__ void __launch_bounds__(288) kU(const int *h2, int nh, int expect, long long spin, int *bad, float *sink) {
  extern __shared__ unsigned char smem[];
  pdl_wait(); pdl_trigger();
  if ((int)threadIdx.x < nh && h2[threadIdx.x] != expect) atomicAdd(bad, 1);
  const long long t0 = clock64();
  float acc = 0.f;
  while (clock64() - t0 < spin) acc += 1e-9f;
  if (threadIdx.x == 0) { smem[0] = 1; if (acc == 12345.f) *sink = acc; }
}

template <class... Args>
static void launch(void (*k)(Args...), dim3 g, dim3 b, size_t smem, cudaStream_t s, bool attr, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = attr ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k, args...);
}

int main(int argc, char **argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  const int nblocks = 586;
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  int *slots, *bad; float *sink;
  cudaMalloc(&slots, nblocks * sizeof(int)); cudaMalloc(&bad, sizeof(int)); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  for (int mode = 0; mode < 4; ++mode) {
    for (long long spin : {2000LL, 20000LL}) {
      cudaMemset(slots, 0xff, nblocks * sizeof(int)); cudaMemset(bad, 0, sizeof(int));
      cudaDeviceSynchronize();
      for (int it = 0; it < iters; ++it) {
        launch(kA, dim3(nblocks), dim3(288), (size_t)110 * 1024, s, mode != 0 && it > 0, slots, it, mode, spin, sink);
        launch(kB, dim3(14), dim3(1024), (size_t)0, s, true, (const int *)slots, nblocks, it, mode, bad);
      }
      cudaStreamSynchronize(s);
      int h = -1; cudaMemcpy(&h, bad, sizeof(int), cudaMemcpyDeviceToHost);
      printf("mode %d spin %lld: stale slots seen by B over %d iterations: %d   (%s)\n", mode, spin, iters, h,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  // mode 4: K0 (no attribute) -> A (attribute, triggers at its start) -> cudaMemcpyAsync D2D of the slots ->
  //         B WITHOUT the attribute checks the slots and the copy.  Is a copy engine operation / a plain
  //         kernel behind a programmatic PRIMARY ordered after its completion, or after its trigger?
  {
    int *copy; cudaMalloc(&copy, nblocks * sizeof(int));
    for (long long spin : {2000LL, 20000LL}) {
      cudaMemset(slots, 0xff, nblocks * sizeof(int)); cudaMemset(copy, 0xff, nblocks * sizeof(int));
      cudaMemset(bad, 0, sizeof(int));
      int *bad2; cudaMalloc(&bad2, sizeof(int)); cudaMemset(bad2, 0, sizeof(int));
      cudaDeviceSynchronize();
      for (int it = 0; it < iters; ++it) {
        launch(kB, dim3(1), dim3(32), (size_t)0, s, false, (const int *)slots, 0, it, 0, bad);   // K0: plain kernel
        launch(kA, dim3(nblocks), dim3(288), (size_t)110 * 1024, s, true, slots, it, 0, spin, sink);
        cudaMemcpyAsync(copy, slots, nblocks * sizeof(int), cudaMemcpyDeviceToDevice, s);
        launch(kB, dim3(14), dim3(1024), (size_t)0, s, false, (const int *)slots, nblocks, it, 0, bad);
        launch(kB, dim3(14), dim3(1024), (size_t)0, s, false, (const int *)copy, nblocks, it, 0, bad2);
      }
      cudaStreamSynchronize(s);
      int h = -1, h2 = -1;
      cudaMemcpy(&h, bad, sizeof(int), cudaMemcpyDeviceToHost); cudaMemcpy(&h2, bad2, sizeof(int), cudaMemcpyDeviceToHost);
      printf("mode 4 spin %lld: stale slots seen behind a D2D copy: %d, stale entries IN the copy: %d   (%s)\n", spin, h, h2,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  {
    const int nh = 14;
    int *h, *h2; cudaMalloc(&h, nh * sizeof(int)); cudaMalloc(&h2, nh * sizeof(int));
    cudaFuncSetAttribute(kU, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    cudaFuncSetAttribute(kG, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    for (long long spin : {2000LL, 20000LL}) {
      cudaMemset(slots, 0xff, nblocks * sizeof(int)); cudaMemset(bad, 0, sizeof(int));
      cudaMemset(h, 0, nh * sizeof(int)); cudaMemset(h2, 0, nh * sizeof(int));
      cudaDeviceSynchronize();
      for (int it = 0; it < iters; ++it) {
        launch(kA, dim3(nblocks), dim3(288), (size_t)110 * 1024, s, it > 0, slots, it, 2, spin, sink);
        launch(kR, dim3(nh), dim3(1024), (size_t)0, s, true, (const int *)slots, nblocks, h, it);
        launch(kG, dim3(1), dim3(1024), (size_t)50 * 1024, s, true, (const int *)h, nh, h2);
        launch(kU, dim3(nblocks), dim3(288), (size_t)110 * 1024, s, true, (const int *)h2, nh, nblocks * 1000 + it % 1000, spin, bad, sink);
      }
      cudaStreamSynchronize(s);
      int hb = -1; cudaMemcpy(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost);
      printf("mode 5 spin %lld: wrong values seen at the end of the chain A -> R -> G -> U over %d iterations: %d   (%s)\n", spin,
             iters, hb, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}