// Debug probe: which async-copy flavours execute on this box?  arg: 0 = 1-D bulk, 1 = 2-D tensor, 2 = 3-D tensor
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait0(uint32_t b) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(b), "r"(0) : "memory");
}
__global__ void k_bulk(const uint16_t* src, uint16_t* out, int n) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t dst = smem_u32(smem), b = smem_u32(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n * 2) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(n * 2), "r"(b) : "memory");
  }
  wait0(b);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}
__global__ void k_t2(const __grid_constant__ CUtensorMap tmap, uint16_t* out, int n, int c0, int c1) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t dst = smem_u32(smem), b = smem_u32(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n * 2) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(&tmap), "r"(b), "r"(c0), "r"(c1) : "memory");
  }
  wait0(b);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}
__global__ void k_t3(const __grid_constant__ CUtensorMap tmap, uint16_t* out, int n, int c0, int c1, int c2) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t dst = smem_u32(smem), b = smem_u32(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n * 2) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(&tmap), "r"(b), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  wait0(b);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}
int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  const int c0 = argc > 6 ? atoi(argv[6]) : 16, c1 = argc > 7 ? atoi(argv[7]) : 32;
  const int bw = argc > 2 ? atoi(argv[2]) : 64, bh = argc > 3 ? atoi(argv[3]) : 8;
  const int W = argc > 4 ? atoi(argv[4]) : 256, H = argc > 5 ? atoi(argv[5]) : 128, CT = 4;
  std::vector<uint16_t> h((size_t)W * H * CT);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(i * 7 + 3);
  uint16_t *d, *out;
  CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  const int n = bw * bh;
  CK(cudaMalloc(&out, n * 2)); CK(cudaMemset(out, 0xff, n * 2));
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fnp);
  CUtensorMap tmap;
  cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)CT}; cuuint64_t gstr[2] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}; cuuint32_t es[3] = {1, 1, 1};
  const int rank = which == 1 ? 2 : 3;
  if (which) {
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, rank, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rank %d box %dx%d -> %d\n", rank, bw, bh, (int)cr);
  }
  if (which == 0) k_bulk<<<1, 128, n * 2>>>(d + 1024, out, n);
  if (which == 1) k_t2<<<1, 128, n * 2>>>(tmap, out, n, c0, c1);
  if (which == 2) k_t3<<<1, 128, n * 2>>>(tmap, out, n, c0, c1, 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("which %d: %s\n", which, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint16_t> o(n);
  CK(cudaMemcpy(o.data(), out, n * 2, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int r = 0; r < bh; ++r) for (int c = 0; c < bw; ++c) {
    uint16_t want = which == 0 ? h[1024 + r * bw + c] : ((c0 + c < W && c0 + c >= 0 && c1 + r < H) ? h[((size_t)(which == 2 ? 1 : 0) * H + c1 + r) * W + c0 + c] : 0);
    if (o[r * bw + c] != want) ++bad;
  }
  printf("which %d mismatches %d\n", which, bad);
  return 0;
}
