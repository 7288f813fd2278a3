// Probe: 2-D tensor TMA of an fp64 [6][cap] tensor, box {32,6}, variants.  nvcc -arch=sm_100a tma_probe.cu -o tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int HINT>
__global__ void k(const __grid_constant__ CUtensorMap tm, int e, double* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1536) : "memory");
    if (HINT)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                   ::"r"(smem_u32(sm)), "l"(&tm), "r"(e), "r"(0), "r"(smem_u32(&bar)), "l"(0x12F0000000000000ull) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(sm)), "l"(&tm), "r"(e), "r"(0), "r"(smem_u32(&bar)) : "memory");
  }
  unsigned ok = 0;
  for (int spin = 0; spin < (1 << 22) && !ok; spin++)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  const double* s = reinterpret_cast<const double*>(sm);
  if (threadIdx.x < 32) for (int r = 0; r < 6; r++) out[r * 32 + threadIdx.x] = ok ? s[r * 32 + threadIdx.x] : -1.0;
}
int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fn;
  const long long cap = 2112;
  double* d; cudaMalloc(&d, cap * 6 * 8);
  double* h = new double[cap * 6];
  for (long long i = 0; i < cap * 6; i++) h[i] = (double)i;
  cudaMemcpy(d, h, cap * 6 * 8, cudaMemcpyHostToDevice);
  double* out; cudaMalloc(&out, 192 * 8);
  for (int dtype = 0; dtype < 2; dtype++) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)cap, 6}, gstr[1] = {(cuuint64_t)cap * 8};
    cuuint32_t box[2] = {32, 6}, estr[2] = {1, 1};
    CUresult r = enc(&tm, dtype ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("dtype %d encode rc %d\n", dtype, (int)r);
    for (int hint = 0; hint < 2; hint++)
      for (int e : {0, 2, 36, 100}) {
        cudaMemset(out, 0, 192 * 8);
        if (hint) k<1><<<1, 64, 2048>>>(tm, e, out); else k<0><<<1, 64, 2048>>>(tm, e, out);
        cudaError_t er = cudaDeviceSynchronize();
        double ho[192]; cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
        printf("dtype %d hint %d e %d: %s  out[0]=%g (want %d) out[32+1]=%g (want %lld) out[191]=%g (want %lld)\n", dtype, hint, e, cudaGetErrorString(er),
               ho[0], e, ho[33], cap + e + 1, ho[191], 5 * cap + e + 31);
        if (er != cudaSuccess) { cudaGetLastError(); return 1; }
      }
  }
  return 0;
}
