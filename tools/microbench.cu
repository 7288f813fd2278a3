// microbench.cu -- B200 pipe-rate probes that size the /fulmov/ kernels:
// DFMA issue rate, uniform (broadcast) 128-bit L1 and shared loads, shuffles,
// fp64 global reductions.  Prints cycles per warp-instruction per SM.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_dfma(double* out, int iters, long long* cyc) {
  double a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// every lane of a warp reads the same 16 B (mode 0), or lanes read 32 consecutive 16 B (mode 1)
__global__ void k_ldg128(const double2* __restrict__ src, int nelem, double* out, int iters, int mode, long long* cyc) {
  double sx = 0, sy = 0;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int idx = (blockIdx.x * 37 + w * 11) % (nelem - 1024);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int off = idx + u * 3 + (mode ? lane : 0);
      double2 v = __ldg(src + off);
      sx += v.x; sy += v.y;
    }
    idx = (idx + 24) % (nelem - 1024);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sx + sy;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_lds128(double* out, int iters, int mode, long long* cyc) {
  __shared__ double2 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_double2(i, -i);
  __syncthreads();
  double sx = 0, sy = 0;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int idx = w * 13;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int off = (idx + u * 3 + (mode ? lane : 0)) & 2047;
      double2 v = sm[off];
      sx += v.x; sy += v.y;
    }
    idx = (idx + 24) & 1023;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sx + sy;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_shfl(double* out, int iters, long long* cyc) {
  double a[4];
  for (int i = 0; i < 4; i++) a[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] += __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 3));
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a[0] + a[1] + a[2] + a[3];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mode 0: every lane its own sector (stride 32 B); mode 1: 4 lanes share a sector (AoS moment node);
// mode 2: all lanes one address
__global__ void k_red(double* dst, long long ndst, int iters, int mode, long long* cyc) {
  const int lane = threadIdx.x & 31;
  long long base = ((long long)(blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4099 % (ndst - 4096);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    long long a;
    if (mode == 0) a = base + lane * 4;
    else if (mode == 1) a = base + lane;
    else a = base;
    atomicAdd(dst + a, 1.0);
    base = (base + 524287) % (ndst - 4096);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double max_cyc(long long* d, int n) {
  long long* h = (long long*)malloc(n * sizeof(long long));
  CK(cudaMemcpy(h, d, n * sizeof(long long), cudaMemcpyDeviceToHost));
  long long m = 0; for (int i = 0; i < n; i++) if (h[i] > m) m = h[i];
  free(h); return (double)m;
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  const int sms = pr.multiProcessorCount;
  printf("device %s, %d SMs\n", pr.name, sms);
  double* out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024 * 8));
  long long* cyc; CK(cudaMalloc(&cyc, sizeof(long long) * 65536));
  const int nelem = 1 << 20; double2* src; CK(cudaMalloc(&src, sizeof(double2) * nelem)); CK(cudaMemset(src, 0, sizeof(double2) * nelem));
  const long long ndst = 1LL << 24; double* dst; CK(cudaMalloc(&dst, sizeof(double) * ndst)); CK(cudaMemset(dst, 0, sizeof(double) * ndst));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM (one block per SM)
    const int threads = wps * 32 > 1024 ? 1024 : wps * 32, blocks = sms * (wps * 32 / threads);
    const int iters = 2000;
    k_dfma<<<blocks, threads>>>(out, iters, cyc); CK(cudaDeviceSynchronize());
    double c = max_cyc(cyc, blocks);
    printf("dfma        warps/SM %2d: %.3f cyc per warp-DFMA per SM\n", wps, c / (double)(iters * 8 * wps));
    k_shfl<<<blocks, threads>>>(out, iters, cyc); CK(cudaDeviceSynchronize());
    c = max_cyc(cyc, blocks);
    printf("shfl64      warps/SM %2d: %.3f cyc per 64-bit shuffle (2 SHFL) per SM\n", wps, c / (double)(iters * 4 * wps));
    for (int mode = 0; mode < 2; mode++) {
      k_ldg128<<<blocks, threads>>>(src, nelem, out, iters, mode, cyc); CK(cudaDeviceSynchronize());
      c = max_cyc(cyc, blocks);
      printf("ldg128 %-9s warps/SM %2d: %.3f cyc per warp-LDG.128 per SM\n", mode ? "coalesced" : "uniform", wps, c / (double)(iters * 8 * wps));
      k_lds128<<<blocks, threads>>>(out, iters, mode, cyc); CK(cudaDeviceSynchronize());
      c = max_cyc(cyc, blocks);
      printf("lds128 %-9s warps/SM %2d: %.3f cyc per warp-LDS.128 per SM\n", mode ? "coalesced" : "uniform", wps, c / (double)(iters * 8 * wps));
    }
  }
  for (int mode = 0; mode < 3; mode++) {
    const int threads = 256, blocks = sms * 4, iters = 2000;
    CK(cudaEventRecord(e0));
    k_red<<<blocks, threads>>>(dst, ndst, iters, mode, cyc);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double c = max_cyc(cyc, blocks);
    const double lane_ops = (double)blocks * threads * iters;
    printf("red.f64 mode %d (%s): %.3f cyc per lane-op per SM (issue), %.2f G lane-ops/s chip-wide (wall)\n", mode,
           mode == 0 ? "1 lane/sector" : mode == 1 ? "4 lanes/sector" : "32 lanes/address", c * sms / lane_ops, lane_ops / (ms * 1e-3) / 1e9);
  }
  return 0;
}
