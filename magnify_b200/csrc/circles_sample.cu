// Candidate circles and their scores: steps 3-5 of the reference's circle finder
// (src/magnify/utils.py:141-189, 221-249, 288-377), for a batch of B images.
//
//   grid lists     utils.py:347-377  edge pixels grouped by grid cell (cell-major, row-major inside
//                                    a cell, exactly the order of `grid_coords`)
//   sampling       utils.py:288-344  three edge pixels -> circumcircle.  The arithmetic reproduces
//                                    numba's typing of those lines bit for bit (float64 everywhere
//                                    except the radius, which is sqrt(c0*c0 + c1*c1) in float32;
//                                    pinned against the reference's own function on 3-pixel images
//                                    in tests/test_circles_host.py).  Which pixels are drawn is NOT
//                                    reproducible: the reference uses numba's unseeded per-thread
//                                    RNG under prange.  Here the draws come from a counter-based
//                                    generator (seed, image, iteration) or from a caller-supplied
//                                    table, and p0 is drawn from the cell-major list (a uniform
//                                    draw over all edge pixels either way).
//   filter + round utils.py:155-165  radius window, round-half-even to int32, off-image test
//   dedupe                           identical rounded circles are kept once (an open-addressing
//                                    table); the reference scores every duplicate again
//   scoring        utils.py:169-189, 221-249  gradient alignment along the Bresenham perimeter,
//                                    summed in float64 in perimeter order, stored as float32,
//                                    divided by the perimeter length in float32
#include "common.cuh"

#include <algorithm>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace {

using mgb::kThreads;

// ---------------------------------------------------------------------------------------------
// grid lists
// ---------------------------------------------------------------------------------------------
struct GridDims {
  int H, W, g, rows, cols;   // rows = ceil(H / g), cols = ceil(W / g)
};

// One thread per (image, cell): count (fill == false) or write (fill == true) its edge pixels in
// row-major order.  coords holds (row << 16 | col); H, W <= 65535.
template <bool kFill>
__global__ void __launch_bounds__(kThreads) cell_lists_kernel(const uint8_t* __restrict__ edges, GridDims d,
                                                              int64_t n_cells_total, int64_t* __restrict__ counts,
                                                              const int64_t* __restrict__ starts,
                                                              uint32_t* __restrict__ coords) {
  const int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (id >= n_cells_total) return;
  const int64_t per_image = (int64_t)d.rows * d.cols;
  const int64_t b = id / per_image;
  const int cell = (int)(id - b * per_image);
  const int cr = cell / d.cols, cc = cell - cr * d.cols;
  const uint8_t* img = edges + b * (int64_t)d.H * d.W;
  const int r1 = min(d.H, (cr + 1) * d.g), c1 = min(d.W, (cc + 1) * d.g);
  int64_t n = kFill ? starts[id] : 0;
  for (int r = cr * d.g; r < r1; ++r)
    for (int c = cc * d.g; c < c1; ++c)
      if (img[(int64_t)r * d.W + c]) {
        if (kFill) coords[n] = ((uint32_t)r << 16) | (uint32_t)c;
        ++n;
      }
  if (!kFill) counts[id] = n;
}

// ---------------------------------------------------------------------------------------------
// sampling
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct SampleParams {
  GridDims d;
  int64_t B, num_iter;
  float min_radius, max_radius;
  uint64_t seed;
  uint64_t table_mask;   // capacity - 1 (power of two)
};

constexpr uint64_t kEmpty = ~0ull;
// 64-bit circle key: image (16 bits) | row + bias (18) | col + bias (18) | radius (12); ascending
// keys = ascending (image, row, col, radius).  Rounded centres lie within max_radius of an image of
// at most 65535 pixels a side, so 18 bits with a bias of 2^16 always suffice.
constexpr int kCoordBias = 1 << 16;
constexpr int kMaxKeyRadius = 4095;

__device__ __forceinline__ uint64_t pack_circle(int64_t b, int row, int col, int r) {
  return ((uint64_t)b << 48) | ((uint64_t)(row + kCoordBias) << 30) | ((uint64_t)(col + kCoordBias) << 12) | (uint64_t)r;
}

// The circumcircle of p0, p0 + q1, p0 + q2 exactly as numba evaluates utils.py:317-342.
__device__ __forceinline__ void circumcircle(int p0r, int p0c, int q1r, int q1c, int q2r, int q2c, float* out) {
  const double eps = (double)1e-20f;
  const double mid1r = 0.5 * q1r, mid1c = 0.5 * q1c, mid2r = 0.5 * q2r, mid2c = 0.5 * q2c;
  const double m1 = __ddiv_rn((double)(-q1c), __dadd_rn((double)q1r, eps));
  const double m2 = __ddiv_rn((double)(-q2c), __dadd_rn((double)q2r, eps));
  const double b1 = __dsub_rn(mid1r, __dmul_rn(m1, mid1c));
  const double b2 = __dsub_rn(mid2r, __dmul_rn(m2, mid2c));
  const float c1 = (float)__ddiv_rn(__dsub_rn(b1, b2), __dadd_rn(__dsub_rn(m2, m1), eps));
  const float c0 = (float)__dadd_rn(__dmul_rn(m1, (double)c1), b1);
  const float rad = __fsqrt_rn(__fadd_rn(__fmul_rn(c0, c0), __fmul_rn(c1, c1)));
  out[0] = (float)__dadd_rn((double)c0, (double)p0r);
  out[1] = (float)__dadd_rn((double)c1, (double)p0c);
  out[2] = rad;
}

__global__ void __launch_bounds__(kThreads) sample_circles_kernel(
    SampleParams p, const uint32_t* __restrict__ coords, const int64_t* __restrict__ starts,
    const int64_t* __restrict__ counts, const int64_t* __restrict__ image_starts, const uint32_t* __restrict__ randoms,
    float* __restrict__ raw, uint64_t* __restrict__ table, uint64_t* __restrict__ unique, unsigned long long* n_unique) {
  const int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (id >= p.B * p.num_iter) return;
  const int64_t b = id / p.num_iter;
  const int64_t base = image_starts[b], n_edges = image_starts[b + 1] - base;
  float c[3] = {NAN, NAN, NAN};
  if (n_edges > 0) {
    uint32_t u0, u1, u2;
    if (randoms) {
      u0 = randoms[3 * id];
      u1 = randoms[3 * id + 1];
      u2 = randoms[3 * id + 2];
    } else {
      const uint64_t a = splitmix64(p.seed ^ (uint64_t)id * 0xD1342543DE82EF95ull);
      const uint64_t e = splitmix64(a);
      u0 = (uint32_t)(a >> 32);
      u1 = (uint32_t)(e >> 32);
      u2 = (uint32_t)e;
    }
    // index = floor(u * n / 2^32): uniform up to 2^-32 * n
    const uint32_t w0 = coords[base + (int64_t)(((uint64_t)u0 * (uint64_t)n_edges) >> 32)];
    const int p0r = (int)(w0 >> 16), p0c = (int)(w0 & 0xffff);
    const int64_t cell = b * (int64_t)p.d.rows * p.d.cols + (int64_t)(p0r / p.d.g) * p.d.cols + p0c / p.d.g;
    const int64_t st = starts[cell], cnt = counts[cell];
    const uint32_t w1 = coords[st + (int64_t)(((uint64_t)u1 * (uint64_t)cnt) >> 32)];
    const uint32_t w2 = coords[st + (int64_t)(((uint64_t)u2 * (uint64_t)cnt) >> 32)];
    circumcircle(p0r, p0c, (int)(w1 >> 16) - p0r, (int)(w1 & 0xffff) - p0c, (int)(w2 >> 16) - p0r,
                 (int)(w2 & 0xffff) - p0c, c);
  }
  if (raw) {
    raw[3 * id] = c[0];
    raw[3 * id + 1] = c[1];
    raw[3 * id + 2] = c[2];
  }
  if (!table) return;
  // utils.py:157-165: radius window on the float radius, round half to even, off-image test
  if (!(c[2] >= p.min_radius && c[2] <= p.max_radius)) return;
  const int row = (int)rintf(c[0]), col = (int)rintf(c[1]), r = (int)rintf(c[2]);
  if (!(row + r >= 0 && col + r >= 0 && row - r < p.d.H && col - r < p.d.W)) return;
  if (row <= -kCoordBias || row >= 3 * kCoordBias || col <= -kCoordBias || col >= 3 * kCoordBias || r > kMaxKeyRadius)
    return;   // cannot happen for max_radius <= 4095 (checked by the launcher)
  const uint64_t key = pack_circle(b, row, col, r);
  uint64_t slot = splitmix64(key) & p.table_mask;
  for (;;) {
    const uint64_t prev = atomicCAS((unsigned long long*)&table[slot], (unsigned long long)kEmpty, (unsigned long long)key);
    if (prev == kEmpty) {
      unique[atomicAdd(n_unique, 1ull)] = key;
      return;
    }
    if (prev == key) return;
    slot = (slot + 1) & p.table_mask;
  }
}

// ---------------------------------------------------------------------------------------------
// scoring
// ---------------------------------------------------------------------------------------------
// The score only reads the angle at edge pixels (utils.py:241-243), so only those are computed
// when an edge map is given (the float64 atan2 is the expensive part of this kernel).
__global__ void __launch_bounds__(kThreads) grad_angle_kernel(const int16_t* __restrict__ dx,
                                                              const int16_t* __restrict__ dy,
                                                              const uint8_t* __restrict__ edges, int64_t n,
                                                              float* __restrict__ angle) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    angle[i] = (!edges || edges[i]) ? (float)atan2((double)dy[i], (double)dx[i]) : 0.0f;   // np.arctan2(dy, dx)
}

// circles (N, 4) int32: image, row, col, radius.  perim_offsets[r - rmin] .. [r - rmin + 1] delimit
// radius r's perimeter points (drow, dcol) and expected angles.
__global__ void __launch_bounds__(kThreads) score_circles_kernel(
    const int32_t* __restrict__ circles, int64_t N, int H, int W, const uint8_t* __restrict__ edges,
    const float* __restrict__ angle, int rmin, int rmax, const int32_t* __restrict__ perim_offsets,
    const int16_t* __restrict__ perim_points, const double* __restrict__ perim_expected, float* __restrict__ scores) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int b = circles[4 * i], row = circles[4 * i + 1], col = circles[4 * i + 2], r = circles[4 * i + 3];
  if (r < rmin || r > rmax) {
    scores[i] = NAN;
    return;
  }
  const int p0 = perim_offsets[r - rmin], p1 = perim_offsets[r - rmin + 1];
  const uint8_t* e = edges + (int64_t)b * H * W;
  const float* a = angle + (int64_t)b * H * W;
  const double pi = 3.141592653589793, half_pi = 1.5707963267948966;
  double sum = 0.0;
  for (int j = p0; j < p1; ++j) {
    const int y = row + perim_points[2 * j], x = col + perim_points[2 * j + 1];
    if ((unsigned)y >= (unsigned)H || (unsigned)x >= (unsigned)W) continue;   // zero padding: not an edge
    const int64_t at = (int64_t)y * W + x;
    if (e[at]) {
      double diff = fabs(__dsub_rn((double)a[at], perim_expected[j]));
      if (diff > pi) diff = __dsub_rn(diff, pi);
      // 4 * |diff - pi/2| / pi - 1, evaluated left to right without contraction
      sum = __dadd_rn(sum, __dsub_rn(__ddiv_rn(__dmul_rn(4.0, fabs(__dsub_rn(diff, half_pi))), pi), 1.0));
    }
  }
  scores[i] = __fdiv_rn((float)sum, (float)(p1 - p0));
}

__global__ void __launch_bounds__(kThreads) unpack_circles_kernel(const uint64_t* __restrict__ unique, int64_t N,
                                                                  int32_t* __restrict__ circles) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint64_t k = unique[i];
  circles[4 * i] = (int32_t)(k >> 48);
  circles[4 * i + 1] = (int32_t)((k >> 30) & 0x3ffff) - kCoordBias;
  circles[4 * i + 2] = (int32_t)((k >> 12) & 0x3ffff) - kCoordBias;
  circles[4 * i + 3] = (int32_t)(k & 0xfff);
}

// key = (image << 32) | (0xffffffff - ordered(score)): ascending keys = image ascending, score
// descending; NaN scores sort last within their image.
__global__ void __launch_bounds__(kThreads) order_keys_kernel(const int32_t* __restrict__ circles,
                                                              const float* __restrict__ scores, int64_t N,
                                                              uint64_t* __restrict__ keys, int32_t* __restrict__ index) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float sc = scores[i];
  uint32_t u = __float_as_uint(sc);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);          // total order of finite floats
  if (sc != sc) u = 0;                                     // NaN: lowest score
  keys[i] = ((uint64_t)(uint32_t)circles[4 * i] << 32) | (uint64_t)(0xffffffffu - u);
  index[i] = (int32_t)i;
}

}  // namespace

extern "C" {

int mgb_order_circles(const int32_t* circles, const float* scores, int64_t N, int32_t* order, void* stream) {
  if (N < 0 || N > INT32_MAX) return MGB_EINVAL;
  if (N == 0) return MGB_OK;
  if (!circles || !scores || !order) return MGB_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint64_t* keys = nullptr;
  int32_t* index = nullptr;
  void* temp = nullptr;
  size_t temp_bytes = 0;
  // keys in, keys out (2 N x 8 bytes) + the identity permutation (N x 4 bytes)
  cudaError_t e = mgb::scratch_alloc((void**)&keys, (size_t)N * 2 * sizeof(uint64_t), s);
  if (e == cudaSuccess) e = mgb::scratch_alloc((void**)&index, (size_t)N * sizeof(int32_t), s);
  if (e == cudaSuccess) {
    order_keys_kernel<<<(unsigned)mgb::ceil_div(N, kThreads), kThreads, 0, s>>>(circles, scores, N, keys, index);
    mgb_count_launch_();
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys, keys + N, index, order, N, 0, 64, s);
  if (e == cudaSuccess) e = mgb::scratch_alloc(&temp, temp_bytes, s);
  if (e == cudaSuccess) {
    // LSD radix sort: stable, so equal (image, score) keep the (row, col, radius) order of the input
    e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys + N, index, order, N, 0, 64, s);
    mgb_count_launch_();
  }
  if (temp) cudaFreeAsync(temp, s);
  if (index) cudaFreeAsync(index, s);
  if (keys) cudaFreeAsync(keys, s);
  return e == cudaSuccess ? MGB_OK : (int)e;
}

int mgb_edge_cell_lists(const uint8_t* edges, int64_t B, int64_t H, int64_t W, int grid_length, int64_t* counts,
                        int64_t* starts, uint32_t* coords, int64_t coords_capacity, int64_t* host_total, void* stream) {
  if (!edges || !counts || !starts || !host_total || B <= 0 || B > 65535 || H <= 0 || W <= 0 || H > 65535 ||
      W > 65535 || grid_length <= 0)
    return MGB_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GridDims d{(int)H, (int)W, grid_length, (int)mgb::ceil_div(H, grid_length), (int)mgb::ceil_div(W, grid_length)};
  const int64_t cells = B * (int64_t)d.rows * d.cols;
  const unsigned blocks = (unsigned)mgb::ceil_div(cells, kThreads);
  if (coords) {
    // second call: counts / starts / *host_total come from the size query on the same edges
    if (coords_capacity < *host_total || *host_total < 0) return MGB_EINVAL;
    if (*host_total > 0) {
      cell_lists_kernel<true><<<blocks, kThreads, 0, s>>>(edges, d, cells, nullptr, starts, coords);
      MGB_CUDA_LAUNCH_CHECK();
    }
    return MGB_OK;
  }
  cell_lists_kernel<false><<<blocks, kThreads, 0, s>>>(edges, d, cells, counts, nullptr, nullptr);
  MGB_CUDA_LAUNCH_CHECK();
  // starts[0 .. cells] = exclusive prefix sum of counts (starts[cells] = total): counts has one
  // extra slot that is zeroed here so that cells + 1 entries can be scanned.
  MGB_CUDA_TRY(cudaMemsetAsync(counts + cells, 0, sizeof(int64_t), s));
  size_t temp_bytes = 0;
  MGB_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, counts, starts, cells + 1, s));
  void* temp = nullptr;
  MGB_CUDA_TRY(mgb::scratch_alloc(&temp, temp_bytes, s));
  cudaError_t e = cub::DeviceScan::ExclusiveSum(temp, temp_bytes, counts, starts, cells + 1, s);
  mgb_count_launch_();
  cudaFreeAsync(temp, s);
  if (e != cudaSuccess) return (int)e;
  int64_t total = 0;
  MGB_CUDA_TRY(cudaMemcpyAsync(&total, starts + cells, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  MGB_CUDA_TRY(cudaStreamSynchronize(s));
  *host_total = total;
  return MGB_OK;
}

int mgb_sample_circles(const uint32_t* coords, const int64_t* starts, const int64_t* counts, int64_t B, int64_t H,
                       int64_t W, int grid_length, int64_t num_iter, float min_radius, float max_radius,
                       uint64_t seed, const uint32_t* randoms, float* raw, uint64_t* table, int64_t table_capacity,
                       int32_t* circles, int64_t* host_n_unique, unsigned long long* counter, void* stream) {
  if (!coords || !starts || !counts || B <= 0 || B > 65535 || H <= 0 || W <= 0 || H > 65535 || W > 65535 ||
      grid_length <= 0 || num_iter < 0)
    return MGB_EINVAL;
  if (table && (max_radius > (float)kMaxKeyRadius || !circles || !host_n_unique || !counter || table_capacity < 2 ||
                (table_capacity & (table_capacity - 1)) != 0 || table_capacity < 2 * B * num_iter))
    return MGB_EINVAL;
  if (host_n_unique) *host_n_unique = 0;
  if (num_iter == 0) return MGB_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SampleParams p;
  p.d = GridDims{(int)H, (int)W, grid_length, (int)mgb::ceil_div(H, grid_length), (int)mgb::ceil_div(W, grid_length)};
  p.B = B;
  p.num_iter = num_iter;
  p.min_radius = min_radius;
  p.max_radius = max_radius;
  p.seed = seed;
  p.table_mask = table ? (uint64_t)table_capacity - 1 : 0;
  // image b's edges are coords[starts[b * cells_per_image] .. starts[(b + 1) * cells_per_image])
  const int64_t per_image = (int64_t)p.d.rows * p.d.cols;
  int64_t* image_starts = nullptr;
  MGB_CUDA_TRY(mgb::scratch_alloc((void**)&image_starts, (size_t)(B + 1) * sizeof(int64_t), s));
  cudaError_t e = cudaMemcpy2DAsync(image_starts, sizeof(int64_t), starts, (size_t)per_image * sizeof(int64_t),
                                    sizeof(int64_t), (size_t)(B + 1), cudaMemcpyDeviceToDevice, s);
  uint64_t* unique = nullptr;
  if (e == cudaSuccess && table) {
    e = cudaMemsetAsync(table, 0xff, (size_t)table_capacity * sizeof(uint64_t), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s);
    if (e == cudaSuccess) e = mgb::scratch_alloc((void**)&unique, (size_t)(B * num_iter) * sizeof(uint64_t), s);
  }
  if (e != cudaSuccess) {
    cudaFreeAsync(image_starts, s);
    if (unique) cudaFreeAsync(unique, s);
    return (int)e;
  }
  sample_circles_kernel<<<(unsigned)mgb::ceil_div(B * num_iter, kThreads), kThreads, 0, s>>>(
      p, coords, starts, counts, image_starts, randoms, raw, table, unique, counter);
  mgb_count_launch_();
  e = cudaGetLastError();
  int64_t n = 0;
  if (e == cudaSuccess && table) {
    unsigned long long host_n = 0;
    e = cudaMemcpyAsync(&host_n, counter, sizeof(host_n), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    n = (int64_t)host_n;
    if (e == cudaSuccess && n > 0) {
      // Order the unique circles by (image, row, col, radius): the scoring kernel's threads then
      // walk neighbouring perimeters (its edge / angle reads hit in L1/L2 instead of fetching a
      // sector per byte), and the output no longer depends on the order of the atomics above.
      uint64_t* sorted = nullptr;
      void* temp = nullptr;
      size_t temp_bytes = 0;
      e = cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, unique, sorted, n, 0, 64, s);
      if (e == cudaSuccess) e = mgb::scratch_alloc((void**)&sorted, (size_t)n * sizeof(uint64_t), s);
      if (e == cudaSuccess) e = mgb::scratch_alloc(&temp, temp_bytes, s);
      if (e == cudaSuccess) {
        e = cub::DeviceRadixSort::SortKeys(temp, temp_bytes, unique, sorted, n, 0, 64, s);
        mgb_count_launch_();
      }
      if (e == cudaSuccess) {
        unpack_circles_kernel<<<(unsigned)mgb::ceil_div(n, kThreads), kThreads, 0, s>>>(sorted, n, circles);
        mgb_count_launch_();
        e = cudaGetLastError();
      }
      if (temp) cudaFreeAsync(temp, s);
      if (sorted) cudaFreeAsync(sorted, s);
    }
  }
  cudaFreeAsync(image_starts, s);
  if (unique) cudaFreeAsync(unique, s);
  if (e != cudaSuccess) return (int)e;
  if (host_n_unique) *host_n_unique = n;
  return MGB_OK;
}

int mgb_gradient_angles(const int16_t* dx, const int16_t* dy, const uint8_t* edges, int64_t n, float* angle,
                        void* stream) {
  if (!dx || !dy || !angle || n < 0) return MGB_EINVAL;
  if (n == 0) return MGB_OK;
  const int64_t blocks = std::min<int64_t>(mgb::ceil_div(n, kThreads), (int64_t)mgb_sm_count() * 16);
  grad_angle_kernel<<<(unsigned)blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(dx, dy, edges, n, angle);
  MGB_CUDA_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_score_circles(const int32_t* circles, int64_t N, int64_t H, int64_t W, const uint8_t* edges,
                      const float* angle, int rmin, int rmax, const int32_t* perim_offsets,
                      const int16_t* perim_points, const double* perim_expected, float* scores, void* stream) {
  if (N < 0 || H <= 0 || W <= 0 || rmin < 1 || rmax < rmin) return MGB_EINVAL;
  if (N == 0) return MGB_OK;
  if (!circles || !edges || !angle || !perim_offsets || !perim_points || !perim_expected || !scores) return MGB_EINVAL;
  score_circles_kernel<<<(unsigned)mgb::ceil_div(N, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      circles, N, (int)H, (int)W, edges, angle, rmin, rmax, perim_offsets, perim_points, perim_expected, scores);
  MGB_CUDA_LAUNCH_CHECK();
  return MGB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Neighbour suppression on the device (utils.py:252-285).  The reference walks the circles best
// first; each accepted circle claims the raster ring of radius min_dist around its centre and a
// circle is rejected when its own ring touches a claimed pixel.  Whether two circles conflict
// depends only on the offset of their centres (ring and shifted ring share a pixel), so the greedy
// pass is a fixed-priority independent-set problem: a circle is REJECTED as soon as a conflicting
// circle of higher rank is KEPT, and KEPT once every conflicting circle of higher rank is
// REJECTED.  Rounds of that rule reach the sequential result (each decision is final when made).
// ---------------------------------------------------------------------------------------------
namespace {

constexpr uint8_t kUndecided = 0, kKept = 1, kRejected = 2;

struct NmsGrid {
  int cell, cells_y, cells_x, origin;   // cell edge; cells per image; coordinate offset (centres >= -origin)
  int reach;                            // conflicts need |drow|, |dcol| <= reach
};

__device__ __forceinline__ int64_t nms_cell_of(const NmsGrid& g, int b, int row, int col) {
  return ((int64_t)b * g.cells_y + (row + g.origin) / g.cell) * g.cells_x + (col + g.origin) / g.cell;
}

__global__ void __launch_bounds__(kThreads) nms_count_kernel(const int32_t* __restrict__ circles, int64_t N, NmsGrid g,
                                                             int32_t* __restrict__ counts) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < N) atomicAdd(&counts[nms_cell_of(g, circles[4 * i], circles[4 * i + 1], circles[4 * i + 2])], 1);
}

__global__ void __launch_bounds__(kThreads) nms_fill_kernel(const int32_t* __restrict__ circles, int64_t N, NmsGrid g,
                                                            const int32_t* __restrict__ starts, int32_t* __restrict__ cursor,
                                                            int32_t* __restrict__ members) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int64_t c = nms_cell_of(g, circles[4 * i], circles[4 * i + 1], circles[4 * i + 2]);
  members[starts[c] + atomicAdd(&cursor[c], 1)] = (int32_t)i;
}

// conflict bitmap: (2 reach + 1)^2 bytes, index (drow + reach) * (2 reach + 1) + dcol + reach
__global__ void __launch_bounds__(kThreads) nms_round_kernel(const int32_t* __restrict__ circles, int64_t N, NmsGrid g,
                                                             const int32_t* __restrict__ starts,
                                                             const int32_t* __restrict__ members,
                                                             const uint8_t* __restrict__ conflict,
                                                             volatile uint8_t* __restrict__ state, int* __restrict__ undecided) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N || state[i] != kUndecided) return;
  const int b = circles[4 * i], row = circles[4 * i + 1], col = circles[4 * i + 2];
  const int cy = (row + g.origin) / g.cell, cx = (col + g.origin) / g.cell;
  const int span = 2 * g.reach + 1;
  bool pending = false;
  for (int ny = max(cy - 1, 0); ny <= min(cy + 1, g.cells_y - 1); ++ny)
    for (int nx = max(cx - 1, 0); nx <= min(cx + 1, g.cells_x - 1); ++nx) {
      const int64_t c = ((int64_t)b * g.cells_y + ny) * g.cells_x + nx;
      for (int k = starts[c]; k < starts[c + 1]; ++k) {
        const int j = members[k];
        if (j >= i) continue;   // only circles of higher rank matter (the list is best first per image)
        const int dr = circles[4 * j + 1] - row, dc = circles[4 * j + 2] - col;
        if (abs(dr) > g.reach || abs(dc) > g.reach || !conflict[(dr + g.reach) * span + dc + g.reach]) continue;
        const uint8_t s = state[j];
        if (s == kKept) {
          state[i] = kRejected;
          return;
        }
        if (s == kUndecided) pending = true;
      }
    }
  if (pending) {
    atomicAdd(undecided, 1);
  } else {
    state[i] = kKept;
  }
}

}  // namespace

extern "C" int mgb_filter_neighbors_device(const int32_t* circles, int64_t N, int64_t B, int64_t H, int64_t W,
                                           int max_radius, int min_dist, const uint8_t* conflict, uint8_t* state,
                                           int* host_rounds, void* stream) {
  if (N < 0 || N > INT32_MAX || B <= 0 || H <= 0 || W <= 0 || min_dist < 1 || max_radius < 0) return MGB_EINVAL;
  if (host_rounds) *host_rounds = 0;
  if (N == 0) return MGB_OK;
  if (!circles || !conflict || !state) return MGB_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  NmsGrid g;
  g.reach = 2 * min_dist;
  g.cell = g.reach + 1;
  g.origin = max_radius + 1;   // centres lie within max_radius of the image (utils.py:160-165)
  g.cells_y = (int)mgb::ceil_div(H + 2 * (int64_t)g.origin, g.cell);
  g.cells_x = (int)mgb::ceil_div(W + 2 * (int64_t)g.origin, g.cell);
  const int64_t cells = B * (int64_t)g.cells_y * g.cells_x;
  if (cells + 1 > INT32_MAX) return MGB_EUNSUPPORTED;
  int32_t *counts = nullptr, *starts = nullptr, *members = nullptr;
  int* undecided = nullptr;
  void* temp = nullptr;
  size_t temp_bytes = 0;
  const unsigned blocks = (unsigned)mgb::ceil_div(N, kThreads);
  cudaError_t e = mgb::scratch_alloc((void**)&counts, (size_t)(cells + 1) * sizeof(int32_t), s);
  if (e == cudaSuccess) e = mgb::scratch_alloc((void**)&starts, (size_t)(cells + 1) * sizeof(int32_t), s);
  if (e == cudaSuccess) e = mgb::scratch_alloc((void**)&members, (size_t)N * sizeof(int32_t), s);
  if (e == cudaSuccess) e = mgb::scratch_alloc((void**)&undecided, sizeof(int), s);
  if (e == cudaSuccess) e = cudaMemsetAsync(counts, 0, (size_t)(cells + 1) * sizeof(int32_t), s);
  if (e == cudaSuccess) e = cudaMemsetAsync(state, 0, (size_t)N, s);
  if (e == cudaSuccess) {
    nms_count_kernel<<<blocks, kThreads, 0, s>>>(circles, N, g, counts);
    mgb_count_launch_();
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, counts, starts, cells + 1, s);
  if (e == cudaSuccess) e = mgb::scratch_alloc(&temp, temp_bytes, s);
  if (e == cudaSuccess) {
    e = cub::DeviceScan::ExclusiveSum(temp, temp_bytes, counts, starts, cells + 1, s);
    mgb_count_launch_();
  }
  if (e == cudaSuccess) e = cudaMemsetAsync(counts, 0, (size_t)(cells + 1) * sizeof(int32_t), s);   // reused as cursors
  if (e == cudaSuccess) {
    nms_fill_kernel<<<blocks, kThreads, 0, s>>>(circles, N, g, starts, counts, members);
    mgb_count_launch_();
    e = cudaGetLastError();
  }
  int rounds = 0;
  bool gave_up = false;
  while (e == cudaSuccess) {
    e = cudaMemsetAsync(undecided, 0, sizeof(int), s);
    if (e != cudaSuccess) break;
    nms_round_kernel<<<blocks, kThreads, 0, s>>>(circles, N, g, starts, members, conflict, state, undecided);
    mgb_count_launch_();
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    ++rounds;
    int left = 0;
    if ((e = cudaMemcpyAsync(&left, undecided, sizeof(int), cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) break;
    if (left == 0) break;
    if (rounds >= 256) {   // a dependency chain this long is cheaper to walk sequentially on the host
      gave_up = true;
      break;
    }
  }
  if (temp) cudaFreeAsync(temp, s);
  if (undecided) cudaFreeAsync(undecided, s);
  if (members) cudaFreeAsync(members, s);
  if (starts) cudaFreeAsync(starts, s);
  if (counts) cudaFreeAsync(counts, s);
  if (host_rounds) *host_rounds = rounds;
  if (e != cudaSuccess) return (int)e;
  return gave_up ? MGB_EUNSUPPORTED : MGB_OK;
}
