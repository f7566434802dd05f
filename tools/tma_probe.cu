// Standalone probe for the TMA window load used by roi_tma.cu (debug aid, not product code).
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>  // 0: grid_constant param, 1: descriptor pointer in global memory
__global__ void probe(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, uint16_t* out, int rows, int wpu,
                      int left, int top, int ct, uint32_t* info, int fences) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  const uint32_t dst = smem_u32(smem);
  const uint32_t b = smem_u32(&bar);
  if (threadIdx.x == 0) {
    info[0] = dst; info[1] = b;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    if (fences & 1) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (fences & 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const CUtensorMap* mp = MODE == 0 ? &tmap : gmap;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(rows * wpu * 2) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(mp), "r"(b), "r"(left), "r"(top), "r"(ct) : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(b), "r"(0) : "memory");
  }
  const uint16_t* s16 = reinterpret_cast<const uint16_t*>(smem);
  for (int i = threadIdx.x; i < rows * wpu; i += blockDim.x) out[i] = s16[i];
}

int main(int argc, char** argv) {
  const int only_mode = argc > 1 ? atoi(argv[1]) : -1;
  const int arg_left = argc > 2 ? atoi(argv[2]) : 180;
  const int fences = argc > 3 ? atoi(argv[3]) : 1;
  const int W = 232, H = 200, CT = 4, rows = 50, wpu = 56;
  std::vector<uint16_t> h((size_t)W * H * CT);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(i * 7 + 3);
  uint16_t *d, *out; uint32_t* info; CUtensorMap* gmap;
  CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&out, rows * wpu * 2)); CK(cudaMalloc(&info, 16)); CK(cudaMalloc(&gmap, sizeof(CUtensorMap)));
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fnp);
  CUtensorMap tmap;
  cuuint64_t gdim[3] = {W, H, CT}; cuuint64_t gstr[2] = {W * 2, (cuuint64_t)H * W * 2};
  cuuint32_t box[3] = {wpu, rows, 1}; cuuint32_t es[3] = {1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode -> %d (alignof tmap %zu)\n", (int)cr, alignof(CUtensorMap));
  CK(cudaMemcpy(gmap, &tmap, sizeof(tmap), cudaMemcpyHostToDevice));
  const int left = arg_left, top = 150, ct = 2;   // right edge padded: 180 + 56 > 232
  for (int mode = 0; mode < 2; ++mode) {
    if (only_mode >= 0 && mode != only_mode) continue;
    CK(cudaMemset(out, 0xff, rows * wpu * 2));
    if (mode == 0) probe<0><<<1, 128, rows * wpu * 2 + 128>>>(tmap, gmap, out, rows, wpu, left, top, ct, info, fences);
    else probe<1><<<1, 128, rows * wpu * 2 + 128>>>(tmap, gmap, out, rows, wpu, left, top, ct, info, fences);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d: %s\n", mode, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint16_t> o(rows * wpu); uint32_t hi[4];
    CK(cudaMemcpy(o.data(), out, o.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hi, info, 16, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < rows; ++r) for (int c = 0; c < wpu; ++c) {
      uint16_t want = (left + c < W) ? h[((size_t)ct * H + top + r) * W + left + c] : 0;
      if (o[r * wpu + c] != want) ++bad;
    }
    printf("mode %d: smem dst 0x%x bar 0x%x mismatches %d\n", mode, hi[0], hi[1], bad);
  }
  return 0;
}
