// Edge-detection front end and circle scoring of the reference's circle finder
// (src/magnify/utils.py:100-344, used by find.py:476-491 for beads and find.py:339-360 for the
// per-ROI button refinement).
//
// Everything in this file is integer or IEEE-exact, so it reproduces the reference's NumPy /
// OpenCV results bit for bit:
//   to_uint8            utils.py:20-27   (x - min) * 255 / (max - min) in float64, truncated
//   GaussianBlur 5x5    utils.py:114     OpenCV's bit-exact fixed-point path for 8-bit images with
//                                        sigma 0: weights [1 4 6 4 1]^2 / 256, round half up,
//                                        BORDER_REFLECT_101
//   Scharr              utils.py:117-118 [3 10 3] x [-1 0 1] on the blurred image, exact integers
//   gradient quantiles  utils.py:125-126 exact order statistics of dx^2 + dy^2 by a 3-level radix
//                                        select (sqrt and the float32 sum are monotone in it)
//   Canny               utils.py:127-133 cv::Canny(dx, dy, low, high, L2gradient=true): squared
//                                        magnitudes, the TG22 non-maximum suppression, 8-connected
//                                        hysteresis on bit planes (iterated to a fixed point; order independent)
// The random circle sampling that follows in the reference (utils.py:288-344) is not
// reproducible by construction (unseeded numba RNG under prange); see circles_sample.cu.
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace {

using mgb::kThreads;

// ---------------------------------------------------------------------------------------------
// to_uint8
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) > v) {
    const unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}
__device__ __forceinline__ void atomic_max_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) < v) {
    const unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}

// images are (B, n) / (B, H, W); blockIdx.y (or .z for the 2-D kernels) selects the image
template <typename T>
__global__ void __launch_bounds__(kThreads) minmax_kernel(const T* __restrict__ x, int64_t n,
                                                          double* __restrict__ mm) {
  x += (int64_t)blockIdx.y * n;
  mm += 2 * blockIdx.y;
  double lo = INFINITY, hi = -INFINITY;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = (double)x[i];
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ double slo[kThreads / 32], shi[kThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    slo[threadIdx.x >> 5] = lo;
    shi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) {
      lo = fmin(lo, slo[w]);
      hi = fmax(hi, shi[w]);
    }
    atomic_min_f64(mm, lo);
    atomic_max_f64(mm + 1, hi);
  }
}

__global__ void minmax_init_kernel(double* mm, int64_t B) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b < B) {
    mm[2 * b] = INFINITY;
    mm[2 * b + 1] = -INFINITY;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) to_uint8_kernel(const T* __restrict__ x, int64_t n,
                                                            const double* __restrict__ mm,
                                                            uint8_t* __restrict__ out) {
  x += (int64_t)blockIdx.y * n;
  out += (int64_t)blockIdx.y * n;
  mm += 2 * blockIdx.y;
  const double lo = mm[0], range = mm[1] - mm[0];   // == max(x - min): the subtraction is monotone
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = (double)x[i] - lo;
    if (range > 0) v = __ddiv_rn(__dmul_rn(255.0, v), range);   // (255 * arr) / max, left to right
    out[i] = (uint8_t)v;                                        // astype(uint8): truncation
  }
}

// ---------------------------------------------------------------------------------------------
// blur + Scharr
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while ((unsigned)p >= (unsigned)len) p = p < 0 ? -p : 2 * len - 2 - p;
  return p;
}

__global__ void __launch_bounds__(kThreads) blur5_u8_kernel(const uint8_t* __restrict__ img, int H, int W,
                                                            uint8_t* __restrict__ out) {
  img += (int64_t)blockIdx.z * H * W;
  out += (int64_t)blockIdx.z * H * W;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y0 = blockIdx.y * 32 + (threadIdx.x >> 5) * 4;
  if (x >= W) return;
  int xs[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) xs[k] = reflect101(x + k - 2, W);
  // horizontal [1 4 6 4 1] sums of the 8 rows this thread's 4 outputs need, then the vertical pass
  int h[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int yy = reflect101(y0 + r - 2, H);
    const uint8_t* row = img + (int64_t)yy * W;
    h[r] = row[xs[0]] + 4 * row[xs[1]] + 6 * row[xs[2]] + 4 * row[xs[3]] + row[xs[4]];
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int y = y0 + r;
    if (y < H) out[(int64_t)y * W + x] = (uint8_t)((h[r] + 4 * h[r + 1] + 6 * h[r + 2] + 4 * h[r + 3] + h[r + 4] + 128) >> 8);
  }
}

__global__ void __launch_bounds__(kThreads) scharr_kernel(const uint8_t* __restrict__ b, int H, int W,
                                                          int16_t* __restrict__ dx, int16_t* __restrict__ dy) {
  b += (int64_t)blockIdx.z * H * W;
  dx += (int64_t)blockIdx.z * H * W;
  dy += (int64_t)blockIdx.z * H * W;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y0 = blockIdx.y * 32 + (threadIdx.x >> 5) * 4;
  if (x >= W) return;
  const int xl = reflect101(x - 1, W), xr = reflect101(x + 1, W);
  int l[6], c[6], r[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const uint8_t* row = b + (int64_t)reflect101(y0 + k - 1, H) * W;
    l[k] = row[xl];
    c[k] = row[x];
    r[k] = row[xr];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int y = y0 + k;
    if (y >= H) break;
    const int gx = 3 * (r[k] - l[k]) + 10 * (r[k + 1] - l[k + 1]) + 3 * (r[k + 2] - l[k + 2]);
    const int gy = 3 * (l[k + 2] - l[k]) + 10 * (c[k + 2] - c[k]) + 3 * (r[k + 2] - r[k]);
    dx[(int64_t)y * W + x] = (int16_t)gx;
    dy[(int64_t)y * W + x] = (int16_t)gy;
  }
}

// ---------------------------------------------------------------------------------------------
// radix select on m = dx^2 + dy^2 (31 bits: 11 + 10 + 10)
// ---------------------------------------------------------------------------------------------
constexpr int kMaxTargets = 4;
struct SelectLevel {
  int shift;          // bits below this level's digit
  int bins;           // 1 << digit bits
  int prefix_shift;   // m >> prefix_shift must equal a target prefix (32 = no prefix at level 0)
  int n_targets;      // histograms per image at this level
};

__global__ void __launch_bounds__(kThreads) grad_hist_kernel(const int16_t* __restrict__ dx,
                                                             const int16_t* __restrict__ dy, int64_t n,
                                                             SelectLevel lv,
                                                             const uint32_t* __restrict__ prefixes,
                                                             uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t sh[];
  const int total = lv.bins * lv.n_targets;
  dx += (int64_t)blockIdx.y * n;
  dy += (int64_t)blockIdx.y * n;
  hist += (int64_t)blockIdx.y * total;
  uint32_t prefix[kMaxTargets];
#pragma unroll
  for (int t = 0; t < kMaxTargets; ++t)
    prefix[t] = lv.prefix_shift >= 32 ? 0u : prefixes[blockIdx.y * kMaxTargets + t];
  for (int i = threadIdx.x; i < total; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // warp-uniform trip count so that the match below can use the full mask
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x - lane); b < n; b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = b + lane;
    const bool valid = i < n;
    const int gx = valid ? dx[i] : 0, gy = valid ? dy[i] : 0;
    const uint32_t m = (uint32_t)(gx * gx + gy * gy);
    const uint32_t digit = (m >> lv.shift) & (uint32_t)(lv.bins - 1);
    int slot = -1;
    if (!valid) {
      slot = -1;
    } else if (lv.prefix_shift >= 32) {
      slot = (int)digit;
    } else {
      const uint32_t p = m >> lv.prefix_shift;
      for (int t = 0; t < lv.n_targets; ++t)       // prefixes are distinct: at most one matches
        if (p == prefix[t]) slot = t * lv.bins + (int)digit;
    }
    // Gradient magnitudes are heavily concentrated (flat background): aggregate equal slots
    // within the warp before touching shared memory.
    const unsigned peers = __match_any_sync(0xffffffffu, slot);
    if (slot >= 0 && lane == __ffs(peers) - 1) atomicAdd(&sh[slot], (uint32_t)__popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// ---------------------------------------------------------------------------------------------
// Canny
// ---------------------------------------------------------------------------------------------
constexpr int kCannyShift = 15;
constexpr int kTG22 = 13573;   // (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5)

__device__ __forceinline__ int mag_at(const int16_t* __restrict__ dx, const int16_t* __restrict__ dy, int H, int W,
                                      int y, int x) {
  if ((unsigned)y >= (unsigned)H || (unsigned)x >= (unsigned)W) return 0;   // zero magnitude outside
  const int gx = dx[(int64_t)y * W + x], gy = dy[(int64_t)y * W + x];
  return gx * gx + gy * gy;
}

// Non-maximum suppression + double threshold.  The result is stored as two bit planes, one 32-bit
// word per image row and 32-pixel column block (bit i = pixel x = 32 * block + i): `strong` (above
// the high threshold: edge) and `cand` (a local maximum above the low threshold: edge if connected
// to a strong pixel).  planes = [strong (B, H, words)] [cand (B, H, words)].
__global__ void __launch_bounds__(kThreads) canny_nms_kernel(const int16_t* __restrict__ dx,
                                                             const int16_t* __restrict__ dy, int H, int W, int words,
                                                             const int32_t* __restrict__ thresholds,
                                                             uint32_t* __restrict__ strong, uint32_t* __restrict__ cand) {
  dx += (int64_t)blockIdx.z * H * W;
  dy += (int64_t)blockIdx.z * H * W;
  const int low = thresholds[2 * blockIdx.z], high = thresholds[2 * blockIdx.z + 1];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (y >= H) return;                       // whole warp
  int out = 1;                              // 0 = candidate, 1 = not an edge, 2 = strong edge
  if (x < W) {
    const int64_t at = (int64_t)y * W + x;
    const int xs = dx[at], ys = dy[at];
    const int m = xs * xs + ys * ys;
    if (m > low) {
      const int ax = abs(xs), ay = abs(ys) << kCannyShift;
      const int tg22x = ax * kTG22;
      bool keep;
      if (ay < tg22x) {
        keep = m > mag_at(dx, dy, H, W, y, x - 1) && m >= mag_at(dx, dy, H, W, y, x + 1);
      } else {
        const int tg67x = tg22x + (ax << (kCannyShift + 1));
        if (ay > tg67x) {
          keep = m > mag_at(dx, dy, H, W, y - 1, x) && m >= mag_at(dx, dy, H, W, y + 1, x);
        } else {
          const int s = (xs ^ ys) < 0 ? -1 : 1;
          keep = m > mag_at(dx, dy, H, W, y - 1, x - s) && m > mag_at(dx, dy, H, W, y + 1, x + s);
        }
      }
      if (keep) out = m > high ? 2 : 0;
    }
  }
  const uint32_t sbits = __ballot_sync(0xffffffffu, out == 2), cbits = __ballot_sync(0xffffffffu, out == 0);
  if ((threadIdx.x & 31) == 0) {
    const int64_t at = ((int64_t)blockIdx.z * H + y) * words + blockIdx.x;
    strong[at] = sbits;
    cand[at] = cbits;
  }
}

// One sweep of hysteresis, one WARP per 32x32 tile, bit-parallel on the planes above: lane r holds
// row r of the tile as 34-bit column masks (bit 0 / bit 33 = the halo pixels left / right of the
// tile, taken from the neighbouring words), lanes 0 and 31 also hold the halo rows above / below.
// A flood step is
//     h = strong | strong << 1 | strong >> 1;  reach = h | row above's h | row below's h;
//     newly = cand & reach;  strong |= newly;  cand &= ~newly
// -- two shuffles and a few logic ops for the whole tile -- repeated until no lane recruits.  Halo
// bits are read once per sweep and not updated inside it.  Each (row, block) word belongs to
// exactly one tile, so the updated strong words are stored without atomics.  A tile only has work
// when it or one of its eight neighbours recruited something in the previous sweep (`active_in`,
// one byte per tile; NULL on the first sweep = every tile); tiles that recruit set `active_out` and
// bump `changed`.  The host repeats sweeps until a sweep changes nothing.
__global__ void __launch_bounds__(kThreads) canny_hysteresis_kernel(uint32_t* __restrict__ strong_plane,
                                                                    uint32_t* __restrict__ cand_plane, int H, int words,
                                                                    int tiles_y, const uint8_t* __restrict__ active_in,
                                                                    uint8_t* __restrict__ active_out,
                                                                    int* __restrict__ changed) {
  const int lane = threadIdx.x & 31;
  const int tiles_x = words;
  const int64_t tile = blockIdx.x * (int64_t)(kThreads / 32) + (threadIdx.x >> 5);
  if (tile >= (int64_t)tiles_x * tiles_y) return;   // whole warp
  const int ty = (int)(tile / tiles_x), tx = (int)(tile - (int64_t)ty * tiles_x);
  const int64_t tile_base = (int64_t)blockIdx.z * tiles_x * tiles_y;
  if (active_in) {
    bool mine = false;
    if (lane < 9) {
      const int ny = ty + lane / 3 - 1, nx = tx + lane % 3 - 1;
      mine = ny >= 0 && ny < tiles_y && nx >= 0 && nx < tiles_x && active_in[tile_base + (int64_t)ny * tiles_x + nx];
    }
    if (!__any_sync(0xffffffffu, mine)) return;
  }
  const uint32_t* sp = strong_plane + (int64_t)blockIdx.z * H * words;
  uint32_t* cp = cand_plane + (int64_t)blockIdx.z * H * words;
  auto strong_row = [&](int y) -> uint64_t {          // 34-bit strong mask of image row y around block tx
    if ((unsigned)y >= (unsigned)H) return 0;
    const uint32_t* row = sp + (int64_t)y * words;
    uint64_t m = (uint64_t)row[tx] << 1;
    if (tx > 0) m |= row[tx - 1] >> 31;
    if (tx + 1 < words) m |= (uint64_t)(row[tx + 1] & 1u) << 33;
    return m;
  };
  const int y = ty * 32 + lane;
  uint64_t strong = strong_row(y);
  uint64_t cand = (y < H) ? (uint64_t)cp[(int64_t)y * words + tx] << 1 : 0;
  uint64_t halo = 0;                                   // lane 0: the row above the tile, lane 31: the row below
  if (lane == 0) halo = strong_row(ty * 32 - 1);
  if (lane == 31) halo = strong_row(ty * 32 + 32);
  const uint64_t before = strong;
  const uint64_t interior = 0x1fffffffeull;            // bits 1..32
  for (;;) {
    const uint64_t h = strong | (strong << 1) | (strong >> 1);
    uint64_t up = __shfl_up_sync(0xffffffffu, h, 1), down = __shfl_down_sync(0xffffffffu, h, 1);
    if (lane == 0) up = halo | (halo << 1) | (halo >> 1);
    if (lane == 31) down = halo | (halo << 1) | (halo >> 1);
    const uint64_t newly = cand & (h | up | down) & interior;
    strong |= newly;
    cand &= ~newly;
    if (!__any_sync(0xffffffffu, newly != 0)) break;
  }
  const uint64_t grew = (strong & ~before) & interior;
  if (__any_sync(0xffffffffu, grew != 0)) {
    if (grew) {
      strong_plane[((int64_t)blockIdx.z * H + y) * words + tx] = (uint32_t)(strong >> 1);
      cp[(int64_t)y * words + tx] = (uint32_t)(cand >> 1);
    }
    if (lane == 0) {
      active_out[tile_base + tile] = 1;
      atomicAdd(changed, 1);
    }
  }
}

// edges (B, H, W) uint8 0/1 from the strong plane (utils.py:139 `edges[edges != 0] = 1`)
__global__ void __launch_bounds__(kThreads) canny_edges_kernel(const uint32_t* __restrict__ strong, int H, int W,
                                                               int words, uint8_t* __restrict__ edges) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const uint32_t word = strong[((int64_t)blockIdx.z * H + y) * words + blockIdx.x];
  edges[((int64_t)blockIdx.z * H + y) * W + x] = (word >> (threadIdx.x & 31)) & 1u;
}

int grid_for(int64_t n) {
  const int64_t blocks = mgb::ceil_div(n, kThreads);
  const int64_t cap = (int64_t)mgb_sm_count() * 16;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

template <typename T>
int to_uint8_launch(const void* src, int64_t B, int64_t n, uint8_t* dst, double* mm, cudaStream_t s) {
  minmax_init_kernel<<<(unsigned)mgb::ceil_div(B, kThreads), kThreads, 0, s>>>(mm, B);
  MGB_CUDA_LAUNCH_CHECK();
  int gx = grid_for(n);
  if (B > 1) gx = (int)std::max<int64_t>(1, std::min<int64_t>(gx, mgb::ceil_div((int64_t)mgb_sm_count() * 16, B)));
  const dim3 grid((unsigned)gx, (unsigned)B);
  minmax_kernel<T><<<grid, kThreads, 0, s>>>(static_cast<const T*>(src), n, mm);
  MGB_CUDA_LAUNCH_CHECK();
  to_uint8_kernel<T><<<grid, kThreads, 0, s>>>(static_cast<const T*>(src), n, mm, dst);
  MGB_CUDA_LAUNCH_CHECK();
  return MGB_OK;
}

}  // namespace

extern "C" {

int mgb_to_uint8(const void* src, int dtype, int64_t B, int64_t n, uint8_t* dst, double* minmax, void* stream) {
  if (B < 0 || n < 0 || B > 65535 || (B * n > 0 && (!src || !dst)) || !minmax) return MGB_EINVAL;
  if (B == 0 || n == 0) return MGB_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case MGB_U8: return to_uint8_launch<uint8_t>(src, B, n, dst, minmax, s);
    case MGB_U16: return to_uint8_launch<uint16_t>(src, B, n, dst, minmax, s);
    case MGB_F32: return to_uint8_launch<float>(src, B, n, dst, minmax, s);
    case MGB_F64: return to_uint8_launch<double>(src, B, n, dst, minmax, s);
    default: return MGB_EUNSUPPORTED;
  }
}

int mgb_edge_gradients_u8(const uint8_t* image, int64_t B, int64_t H, int64_t W, uint8_t* blurred, int16_t* dx,
                          int16_t* dy, void* stream) {
  if (!image || !blurred || !dx || !dy || B <= 0 || B > 65535 || H <= 0 || W <= 0 || H > (1 << 30) || W > (1 << 30))
    return MGB_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 grid((unsigned)mgb::ceil_div(W, 32), (unsigned)mgb::ceil_div(H, 32), (unsigned)B);
  blur5_u8_kernel<<<grid, kThreads, 0, s>>>(image, (int)H, (int)W, blurred);
  MGB_CUDA_LAUNCH_CHECK();
  scharr_kernel<<<grid, kThreads, 0, s>>>(blurred, (int)H, (int)W, dx, dy);
  MGB_CUDA_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_gradient_order_stats(const int16_t* dx, const int16_t* dy, int64_t B, int64_t n, const int64_t* host_ranks,
                             int n_ranks, int64_t* host_values, uint32_t* scratch, void* stream) {
  if (!dx || !dy || !host_ranks || !host_values || !scratch || B <= 0 || B > 65535 || n <= 0 || n_ranks < 1 ||
      n_ranks > kMaxTargets)
    return MGB_EINVAL;
  for (int t = 0; t < n_ranks; ++t)
    if (host_ranks[t] < 0 || host_ranks[t] >= n) return MGB_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int shifts[3] = {20, 10, 0}, bins[3] = {2048, 1024, 1024};
  // scratch layout: [B * 4] prefixes, then [B * 4 * 2048] histogram words
  uint32_t* d_prefix = scratch;
  uint32_t* d_hist = scratch + (size_t)B * kMaxTargets;
  std::vector<uint32_t> prefix((size_t)B * kMaxTargets, 0u), slot_prefix((size_t)B * kMaxTargets, 0xffffffffu);
  std::vector<int> slot_of((size_t)B * kMaxTargets, 0);
  std::vector<int64_t> rank((size_t)B * kMaxTargets);
  for (int64_t b = 0; b < B; ++b)
    for (int t = 0; t < n_ranks; ++t) rank[b * kMaxTargets + t] = host_ranks[t];
  std::vector<uint32_t> host_hist;
  int gx = grid_for(n);
  if (B > 1) gx = (int)std::max<int64_t>(1, std::min<int64_t>(gx, mgb::ceil_div((int64_t)mgb_sm_count() * 16, B)));
  for (int level = 0; level < 3; ++level) {
    SelectLevel lv;
    lv.shift = shifts[level];
    lv.bins = bins[level];
    lv.prefix_shift = level == 0 ? 32 : shifts[level - 1];
    // ranks of one image that fell into the same bin share a histogram slot
    int max_slots = 1;
    for (int64_t b = 0; b < B; ++b) {
      int used = 0;
      for (int t = 0; t < n_ranks; ++t) {
        int found = -1;
        for (int u = 0; u < used; ++u)
          if (level == 0 || slot_prefix[b * kMaxTargets + u] == prefix[b * kMaxTargets + t]) found = u;
        if (found < 0) {
          found = used++;
          slot_prefix[b * kMaxTargets + found] = prefix[b * kMaxTargets + t];
        }
        slot_of[b * kMaxTargets + t] = found;
      }
      for (int u = used; u < kMaxTargets; ++u) slot_prefix[b * kMaxTargets + u] = 0xffffffffu;
      max_slots = std::max(max_slots, used);
    }
    lv.n_targets = level == 0 ? 1 : max_slots;
    const size_t words = (size_t)B * lv.bins * lv.n_targets;
    MGB_CUDA_TRY(cudaMemcpyAsync(d_prefix, slot_prefix.data(), (size_t)B * kMaxTargets * sizeof(uint32_t),
                                 cudaMemcpyHostToDevice, s));
    MGB_CUDA_TRY(cudaMemsetAsync(d_hist, 0, words * sizeof(uint32_t), s));
    grad_hist_kernel<<<dim3((unsigned)gx, (unsigned)B), kThreads, (size_t)lv.bins * lv.n_targets * sizeof(uint32_t), s>>>(
        dx, dy, n, lv, d_prefix, d_hist);
    MGB_CUDA_LAUNCH_CHECK();
    host_hist.resize(words);
    MGB_CUDA_TRY(cudaMemcpyAsync(host_hist.data(), d_hist, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    MGB_CUDA_TRY(cudaStreamSynchronize(s));
    for (int64_t b = 0; b < B; ++b) {
      for (int t = 0; t < n_ranks; ++t) {
        const uint32_t* h = host_hist.data() + ((size_t)b * lv.n_targets + slot_of[b * kMaxTargets + t]) * lv.bins;
        int64_t acc = 0;
        int d = 0;
        for (; d < lv.bins; ++d) {
          if (acc + h[d] > rank[b * kMaxTargets + t]) break;
          acc += h[d];
        }
        if (d == lv.bins) return MGB_EINVAL;   // rank beyond the counted elements: cannot happen
        rank[b * kMaxTargets + t] -= acc;
        uint32_t& p = prefix[b * kMaxTargets + t];
        p = level == 0 ? (uint32_t)d : ((p << 10) | (uint32_t)d);
      }
    }
  }
  for (int64_t b = 0; b < B; ++b)
    for (int t = 0; t < n_ranks; ++t) host_values[b * n_ranks + t] = (int64_t)prefix[b * kMaxTargets + t];
  return MGB_OK;
}

int mgb_canny(const int16_t* dx, const int16_t* dy, int64_t B, int64_t H, int64_t W, const int32_t* thresholds,
              uint8_t* edges, int* changed, int* host_sweeps, void* stream) {
  if (!dx || !dy || !thresholds || !edges || !changed || B <= 0 || B > 65535 || H <= 0 || W <= 0 || H > (1 << 30) ||
      W > (1 << 30))
    return MGB_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int words = (int)mgb::ceil_div(W, 32);
  const dim3 px_grid((unsigned)words, (unsigned)mgb::ceil_div(H, 8), (unsigned)B);
  const int tiles_y = (int)mgb::ceil_div(H, 32);
  const size_t plane = (size_t)B * H * words;                 // words per bit plane
  const size_t n_tiles = (size_t)B * tiles_y * words;
  // scratch: strong plane, candidate plane, two byte-per-tile activity maps (ping-ponged between sweeps)
  uint32_t* planes = nullptr;
  MGB_CUDA_TRY(mgb::scratch_alloc((void**)&planes, 2 * plane * sizeof(uint32_t) + 2 * n_tiles, s));
  uint32_t *strong = planes, *cand = planes + plane;
  uint8_t* active = reinterpret_cast<uint8_t*>(planes + 2 * plane);
  cudaError_t e = cudaSuccess;
  canny_nms_kernel<<<px_grid, kThreads, 0, s>>>(dx, dy, (int)H, (int)W, words, thresholds, strong, cand);
  mgb_count_launch_();
  e = cudaGetLastError();
  int sweeps = 0;
  while (e == cudaSuccess) {
    uint8_t* out = active + (size_t)(sweeps & 1) * n_tiles;
    const uint8_t* in = sweeps == 0 ? nullptr : active + (size_t)((sweeps + 1) & 1) * n_tiles;
    if ((e = cudaMemsetAsync(changed, 0, sizeof(int), s)) != cudaSuccess) break;
    if ((e = cudaMemsetAsync(out, 0, n_tiles, s)) != cudaSuccess) break;
    canny_hysteresis_kernel<<<dim3((unsigned)mgb::ceil_div((int64_t)words * tiles_y, kThreads / 32), 1, (unsigned)B),
                              kThreads, 0, s>>>(strong, cand, (int)H, words, tiles_y, in, out, changed);
    mgb_count_launch_();
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    ++sweeps;
    int host_changed = 0;
    if ((e = cudaMemcpyAsync(&host_changed, changed, sizeof(int), cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) break;
    if (host_changed == 0) break;
  }
  if (e == cudaSuccess) {
    canny_edges_kernel<<<px_grid, kThreads, 0, s>>>(strong, (int)H, (int)W, words, edges);
    mgb_count_launch_();
    e = cudaGetLastError();
  }
  cudaFreeAsync(planes, s);
  if (e != cudaSuccess) return (int)e;
  if (host_sweeps) *host_sweeps = sweeps;
  return MGB_OK;
}

}  // extern "C"
