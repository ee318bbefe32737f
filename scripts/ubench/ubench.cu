// ubench.cu -- micro-benchmarks behind design decisions of the Jacobian pass (development tool, not product):
//   lds  : cycles per warp-wide LDS.128 for address patterns with different amounts of broadcast
//   dfma : FP64 FMA throughput per SM as a function of independent chains per thread and warps per scheduler
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench ubench.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

// pattern p: address (in 16 B chunks) of lane l
__device__ __forceinline__ int pat_addr(int p, int l) {
  switch (p) {
    case 0: return l;                        // 32 distinct chunks, conflict-free: 4 wavefronts expected
    case 1: return l & 7;                    // 8 distinct chunks, each read by one lane of every quarter-warp
    case 2: return (l >> 2);                 // 8 distinct chunks, each read by 4 neighbouring lanes (2 distinct per quarter)
    case 3: return 0;                        // one chunk, full broadcast
    case 4: return (l & 7) * 8;              // 8 distinct chunks in the SAME bank group: 8-way conflict per quarter
    case 5: return (l & 1) + 2 * (l >> 3);   // per quarter: 2 distinct chunks; 8 distinct per warp
    case 6: return (l & 3) + 4 * (l >> 4);   // 4 distinct per quarter, 8 per warp, halves differ
    default: return l & 15;                  // 16 distinct: quarters 0,2 share, 1,3 share
  }
}
__global__ void lds_kernel(int pat, int iters, double* out, long long* cyc) {
  __shared__ double2 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_double2(i, -i);
  __syncthreads();
  int const l = threadIdx.x & 31;
  int a = pat_addr(pat, l);
  double acc = 0.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double2 v = sm[(a + 32 * k) & 1023];
      acc += v.x + v.y;
    }
    a = (a + (int)(acc * 0.0)) & 1023;  // keep the address a run-time value
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int ILP>
__global__ void dfma_kernel(int iters, double* out, long long* cyc) {
  double a[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) a[k] = threadIdx.x * 1e-3 + k;
  double const b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int k = 0; k < ILP; ++k) a[k] = fma(a[k], b, c);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* out; long long* cyc;
  CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024 * 4));
  CK(cudaMalloc(&cyc, sizeof(long long) * 4096));
  long long h[4096];
  printf("{\"sms\": %d,\n \"lds\": [", sms);
  int const iters = 2000;
  for (int warps = 1; warps <= 16; warps *= 4)
    for (int p = 0; p < 8; ++p) {
      lds_kernel<<<sms, 32 * warps>>>(p, iters, out, cyc);
      CK(cudaDeviceSynchronize());
      lds_kernel<<<sms, 32 * warps>>>(p, iters, out, cyc);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
      double m = 0; for (int i = 0; i < sms; ++i) m += h[i];
      m /= sms;
      // cycles of the SM's shared-memory pipe per warp-wide LDS.128 (all warps of the block run concurrently)
      printf("%s{\"warps\": %d, \"pattern\": %d, \"cycles_per_warp_lds128\": %.2f}", (warps == 1 && p == 0) ? "" : ", ", warps, p, m / (iters * 8.0 * warps));
    }
  printf("],\n \"dfma\": [");
  bool first = true;
  auto run = [&](auto kern, int ilp, int warps) {
    int const it = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<sms, 32 * warps>>>(it, out, cyc); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); kern<<<sms, 32 * warps>>>(it, out, cyc); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    CK(cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double m = 0; for (int i = 0; i < sms; ++i) m += h[i]; m /= sms;
    double const fmas = (double)it * 8 * ilp * 32 * warps;           // per SM
    printf("%s{\"ilp\": %d, \"warps_per_sm\": %d, \"dfma_per_clk_per_sm\": %.2f, \"tflops\": %.2f}", first ? "" : ", ", ilp, warps, fmas / m,
           2.0 * fmas * sms / (ms * 1e-3) / 1e12);
    first = false;
  };
  for (int warps : {4, 8, 12, 16, 32}) {
    run(dfma_kernel<1>, 1, warps); run(dfma_kernel<2>, 2, warps); run(dfma_kernel<4>, 4, warps); run(dfma_kernel<8>, 8, warps); run(dfma_kernel<16>, 16, warps);
  }
  printf("]}\n");
  return 0;
}
