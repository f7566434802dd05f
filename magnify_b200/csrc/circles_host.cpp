// Host-side pieces of the circle finder: the Bresenham perimeter table (utils.py:433-465) that the
// scoring kernel walks, and the sequential non-maximum suppression of utils.py:252-285.  Both are
// tiny and inherently ordered, so they stay on the CPU.
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "magnify_b200.h"

extern "C" {

// Perimeter of the reference's circle raster, in the reference's order: the four axis points, then
// for every step of the first octant its eight mirror images, then the four diagonal points when
// the walk ends on the diagonal.  Steps move right while inside the circle, otherwise in (and, when
// diagonal moves are allowed, right as well).  Points are (drow, dcol) pairs.
int mgb_circle_perimeter(int r, int four_connected, int32_t* host_points, int capacity, int* host_n) {
  if (r < 1 || !host_points || !host_n || capacity < 20 * r) return MGB_EINVAL;
  int n = 0;
  auto emit = [&](long long a, long long b) {
    host_points[2 * n] = (int32_t)a;
    host_points[2 * n + 1] = (int32_t)b;
    ++n;
  };
  emit(0, -r); emit(-r, 0); emit(0, r); emit(r, 0);
  long long u = 1, v = -(long long)r;
  const long long r2 = (long long)r * r;
  while (u < -v) {
    emit(u, v); emit(v, u); emit(-u, v); emit(-v, u); emit(u, -v); emit(v, -u); emit(-u, -v); emit(-v, -u);
    if (u * u + v * v - r2 <= 0) {
      ++u;
    } else {
      ++v;
      if (!four_connected) ++u;
    }
  }
  if (v == -u) {
    emit(u, v); emit(-u, -v); emit(-u, v); emit(u, -v);
  }
  *host_n = n;
  return MGB_OK;
}

// utils.py:252-285: circles arrive best first; each accepted circle claims the perimeter ring of
// radius min_dist (4-connected raster) around its centre, and a circle is rejected as soon as one
// pixel of its own ring is already claimed.  The claim raster has the reference's geometry
// ((max row + 2 pad) x (max col + 2 pad), pad = 2 min_dist + 1); indices that fall below zero wrap
// around like NumPy's negative indexing does in the reference.
int mgb_filter_neighbors(const int32_t* host_circles, int64_t n, int min_dist, uint8_t* host_valid) {
  if (n < 0 || (n > 0 && (!host_circles || !host_valid)) || min_dist < 1) return MGB_EINVAL;
  if (n == 0) return MGB_OK;
  std::vector<int32_t> ring((size_t)40 * min_dist);
  int ring_n = 0;
  int rc = mgb_circle_perimeter(min_dist, 1, ring.data(), 20 * min_dist, &ring_n);
  if (rc != MGB_OK) return rc;
  const long long pad = 2LL * min_dist + 1;
  long long max_row = host_circles[0], max_col = host_circles[1];
  for (int64_t i = 1; i < n; ++i) {
    if (host_circles[3 * i] > max_row) max_row = host_circles[3 * i];
    if (host_circles[3 * i + 1] > max_col) max_col = host_circles[3 * i + 1];
  }
  const long long rows = max_row + 2 * pad, cols = max_col + 2 * pad;
  if (rows <= 0 || cols <= 0) return MGB_EINVAL;
  std::vector<uint64_t> claimed((size_t)((rows * cols + 63) / 64), 0);
  auto cell = [&](long long row, long long col) -> long long {
    if (row < 0) row += rows;
    if (col < 0) col += cols;
    if (row < 0 || col < 0 || row >= rows || col >= cols) return -1;   // out of the raster even after the wrap
    return row * cols + col;
  };
  for (int64_t i = 0; i < n; ++i) {
    const long long row = host_circles[3 * i] + pad, col = host_circles[3 * i + 1] + pad;
    bool ok = true;
    for (int j = 0; j < ring_n && ok; ++j) {
      const long long at = cell(ring[2 * j] + row, ring[2 * j + 1] + col);
      if (at >= 0 && ((claimed[(size_t)(at >> 6)] >> (at & 63)) & 1)) ok = false;
    }
    host_valid[i] = ok ? 1 : 0;
    if (!ok) continue;
    for (int j = 0; j < ring_n; ++j) {
      const long long at = cell(ring[2 * j] + row, ring[2 * j + 1] + col);
      if (at >= 0) claimed[(size_t)(at >> 6)] |= (1ull << (at & 63));
    }
  }
  return MGB_OK;
}

}  // extern "C"
