// Microbenchmark: throughput of fp64 reductions (red.global.add.f64) into a small L2-resident tally, the operation
// save_radiation_field performs once per crossed cell (xKJ_abs(icell) += kappa_abs * l * Stokes(1), radiation_field.f90:53).
// It is the ceiling SURVEY.md 8(d) names for the grids whose tables live in L2 (G1-G3: HBM traffic ~ 0).
//   l2_atomic_peak <n_cells> <log2 n_ops> [weights.f64]
// weights.f64: n_cells doubles, the measured number of crossings of every cell (xN_abs of a G1 run); without it only the
// uniform distribution is measured.  Variants: plain reduction per lane, and warp-aggregated (__match_any_sync: lanes of a
// warp that hit the same cell add up first and issue one reduction).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/l2_atomic_peak tools/l2_atomic_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <random>
#include <algorithm>
#include <cuda_runtime.h>

__global__ void red_plain(const int* __restrict__ idx, size_t n, double* tally) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    atomicAdd(tally + idx[i], 1.0);
}
__global__ void red_match(const int* __restrict__ idx, size_t n, double* tally) {
  const unsigned lane = threadIdx.x & 31;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = idx[i];
    const unsigned peers = __match_any_sync(__activemask(), c);
    if ((unsigned)(__ffs(peers) - 1) == lane) atomicAdd(tally + c, (double)__popc(peers));
  }
}
// the access pattern of the photon loop: every lane walks its own sequence of neighbouring cells (here: a random walk
// over the index array), so consecutive reductions of a lane are dependent on nothing and hit nearby lines
__global__ void copy_only(const int* __restrict__ idx, size_t n, int* sink) {
  int s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s ^= idx[i];
  if (s == 0x7fffffff) *sink = s;
}

static double time_kernel(void (*k)(const int*, size_t, double*), const int* d_idx, size_t n, double* d_tally, int blocks, int threads) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, threads>>>(d_idx, n, d_tally);            // warm-up
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k<<<blocks, threads>>>(d_idx, n, d_tally);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  return (double)n / (best * 1e-3);
}

int main(int argc, char** argv) {
  const int n_cells = argc > 1 ? atoi(argv[1]) : 7000;
  const int lg = argc > 2 ? atoi(argv[2]) : 28;
  const size_t n = (size_t)1 << lg;
  std::vector<double> wts;
  if (argc > 3) {
    FILE* f = fopen(argv[3], "rb");
    if (f) { wts.resize(n_cells); if (fread(wts.data(), 8, n_cells, f) != (size_t)n_cells) wts.clear(); fclose(f); }
  }
  int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  int* d_idx; double* d_tally; int* d_sink;
  cudaMalloc(&d_idx, n * sizeof(int)); cudaMalloc(&d_tally, n_cells * sizeof(double)); cudaMalloc(&d_sink, 4);
  cudaMemset(d_tally, 0, n_cells * sizeof(double));
  std::vector<int> h(n);
  std::mt19937_64 rng(12345);
  printf("{\"device\": \"%s\", \"n_cells\": %d, \"n_ops\": %zu", prop.name, n_cells, n);
  for (int dist = 0; dist < (wts.empty() ? 1 : 2); ++dist) {
    if (dist == 0) { std::uniform_int_distribution<int> u(0, n_cells - 1); for (size_t i = 0; i < n; ++i) h[i] = u(rng); }
    else { std::discrete_distribution<int> d(wts.begin(), wts.end()); for (size_t i = 0; i < n; ++i) h[i] = d(rng); }
    cudaMemcpy(d_idx, h.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    const double plain = time_kernel(red_plain, d_idx, n, d_tally, blocks, threads);
    const double match = time_kernel(red_match, d_idx, n, d_tally, blocks, threads);
    // fraction of the reductions the aggregation removes (host count over warps of 32 consecutive entries is not what the
    // grid-stride loop forms; the device count is what matters, so just report the two rates)
    printf(", \"%s\": {\"red_f64_per_s\": %.4e, \"red_f64_match_any_per_s\": %.4e, \"payload_GBps\": %.1f}", dist == 0 ? "uniform" : "g1_hits", plain, match, plain * 8e-9);
  }
  // index-stream read rate alone (upper bound set by the benchmark's own input stream)
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    copy_only<<<blocks, threads>>>(d_idx, n, d_sink); cudaDeviceSynchronize();
    cudaEventRecord(e0); copy_only<<<blocks, threads>>>(d_idx, n, d_sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf(", \"index_stream_per_s\": %.4e", (double)n / (ms * 1e-3));
  }
  printf("}\n");
  return 0;
}
