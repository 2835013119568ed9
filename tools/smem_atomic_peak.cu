// Micro-benchmark: shared-memory atomicAdd throughput of this GPU (SURVEY.md section 8d asks for
// the box's own peak next to the HBM copy peak: MEASURED_PEAKS.json has no such figure).
// The counting kernels (k_range, k_hash) issue one shared-memory atomic per wedge, so this is
// their secondary ceiling.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/smem_atomic_peak tools/smem_atomic_peak.cu
//   tools/_build/smem_atomic_peak            -> one JSON object on stdout
//
// Patterns (u32 counters, 48 K of them per block like a k_range window, 1024 threads per block,
// one block per SM x 1 and x 2 resident):
//   stride   lane i of every warp hits bank i            (conflict-free upper bound)
//   random   a different pseudo-random counter per lane   (what wedge counting looks like)
//   same     all 32 lanes of a warp hit one counter       (worst case: serialised)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int COUNTERS = 48 * 1024;
constexpr int THREADS = 1024;
constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(THREADS) k_atomics(unsigned int* out) {
  extern __shared__ unsigned int cnt[];
  for (int i = threadIdx.x; i < COUNTERS; i += THREADS) cnt[i] = 0;
  __syncthreads();
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  #pragma unroll 8
  for (int it = 0; it < ITERS; ++it) {
    uint32_t idx;
    if (MODE == 0) idx = (warp * 32u + (uint32_t)it * 1024u + lane) % COUNTERS;          // bank = lane
    else if (MODE == 1) { x = x * 1664525u + 1013904223u; idx = (x >> 8) % COUNTERS; }   // random
    else idx = (warp * 97u + (uint32_t)it) % COUNTERS;                                   // one address per warp
    atomicAdd(cnt + idx, 1u);
  }
  __syncthreads();
  unsigned int s = 0;
  for (int i = threadIdx.x; i < COUNTERS; i += THREADS) s += cnt[i];
  if (s == 0xffffffffu) out[0] = s;      // keep the work alive
  if (threadIdx.x == 0) atomicAdd(out + 1, s);
}

template <int MODE>
double run(int blocks, unsigned int* d_out) {
  cudaFuncSetAttribute(k_atomics<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, COUNTERS * 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 3; ++w) k_atomics<MODE><<<blocks, THREADS, COUNTERS * 4>>>(d_out);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 10; ++r) {
    cudaEventRecord(a);
    k_atomics<MODE><<<blocks, THREADS, COUNTERS * 4>>>(d_out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return (double)blocks * THREADS * ITERS / (best * 1e-3) / 1e9;     // G atomics / s
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
  unsigned int* d_out;
  cudaMalloc(&d_out, 16);
  cudaMemset(d_out, 0, 16);
  const int sms = prop.multiProcessorCount;
  const double s1 = run<0>(sms, d_out), r1 = run<1>(sms, d_out), w1 = run<2>(sms, d_out);
  if (cudaGetLastError() != cudaSuccess) { fprintf(stderr, "kernel failed\n"); return 1; }
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %d, \"counters_per_block\": %d, \"threads_per_block\": %d, "
         "\"unit\": \"G shared-memory atomicAdd.u32 per second, whole GPU\", "
         "\"stride_conflict_free\": %.1f, \"random\": %.1f, \"same_address_per_warp\": %.1f, "
         "\"per_sm_per_clock_random\": %.2f}\n",
         prop.name, sms, prop.clockRate / 1000, COUNTERS, THREADS, s1, r1, w1,
         r1 * 1e9 / sms / (prop.clockRate * 1e3));
  return 0;
}
