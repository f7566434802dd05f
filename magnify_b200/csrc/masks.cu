// F5-F8: per-ROI foreground / background masks, hand-written for sm_100a.
// Reference: src/magnify/utils.py:30-52 (circle, annulus), 380-465 (circle_labels,
// filled_circle_points, circle_points); call sites find.py:380-400 (chip), 561-586 (beads).
#include <algorithm>

#include "common.cuh"

namespace mgb {

// ---- F8: chip masks ------------------------------------------------------------------------
// cv.circle(thickness=-1) with integer centre and radius r is {dx^2 + dy^2 <= r^2} clipped to
// the canvas (pinned against cv2 4.13.0 by tests/test_oracle_geometry.py and
// tests/golden/masks_cv.npz); annulus = outer & ~inner.
__global__ void __launch_bounds__(kThreads)
chip_masks_kernel(const int32_t* __restrict__ rel, const int32_t* __restrict__ r_fg, int r_inner,
                  int r_outer, int L, uint8_t* __restrict__ fg, uint8_t* __restrict__ bg,
                  int32_t* __restrict__ counts) {
  __shared__ uint32_t red[2][kThreads / 32];
  const int64_t m = blockIdx.x;
  const int cy = rel[2 * m], cx = rel[2 * m + 1];
  const long long rf = r_fg[m];
  const long long rf2 = rf < 0 ? -1 : rf * rf;
  const long long ri2 = r_inner < 0 ? -1 : (long long)r_inner * r_inner;
  const long long ro2 = r_outer < 0 ? -1 : (long long)r_outer * r_outer;
  uint8_t* f = fg + m * (int64_t)L * L;
  uint8_t* b = bg + m * (int64_t)L * L;
  uint32_t nf = 0, nb = 0;
  const int total = L * L;
  for (int i = threadIdx.x; i < total; i += kThreads) {
    const int row = i / L, col = i - row * L;
    const long long dy = row - cy, dx = col - cx;
    const long long d2 = dx * dx + dy * dy;
    const bool isf = d2 <= rf2;
    const bool isb = (d2 <= ro2) && !(d2 <= ri2);
    f[i] = isf;
    b[i] = isb;
    nf += isf;
    nb += isb;
  }
  if (counts) {
    nf = __reduce_add_sync(0xffffffffu, nf);
    nb = __reduce_add_sync(0xffffffffu, nb);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = nf; red[1][threadIdx.x >> 5] = nb; }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t a = 0, c = 0;
      for (int i = 0; i < kThreads / 32; ++i) { a += red[0][i]; c += red[1][i]; }
      counts[2 * m] = (int32_t)a;
      counts[2 * m + 1] = (int32_t)c;
    }
  }
}

// ---- F6: bead label raster -----------------------------------------------------------------
// The first disc to touch a pixel claims it with a CAS on -1; every later touch stores -2.
// Once a pixel left -1 no CAS can succeed again, so the result is order independent and equals
// the serial loop of utils.py:384-393.
__global__ void __launch_bounds__(128)
bead_labels_kernel(const int32_t* __restrict__ beads, int64_t H, int64_t W,
                   const int32_t* __restrict__ hw, int rmax, int32_t* __restrict__ labels) {
  const int64_t i = blockIdx.x;
  const int cy = beads[3 * i], cx = beads[3 * i + 1], r = beads[3 * i + 2];
  if (r < 1 || r > rmax) return;
  const int32_t* hwr = hw + (int64_t)r * (rmax + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int d = -r + warp; d <= r; d += 4) {
    const long long yy = (long long)cy + d;
    if (yy < 0 || yy >= H) continue;
    const int half = hwr[d < 0 ? -d : d];
    for (int e = -half + lane; e <= half; e += 32) {
      const long long xx = (long long)cx + e;
      if (xx < 0 || xx >= W) continue;
      int32_t* cell = labels + yy * W + xx;
      const int old = atomicCAS(cell, -1, (int)i);
      if (old != -1) atomicExch(cell, -2);
    }
  }
}

// ---- F7: bead masks ------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
bead_masks_kernel(const int32_t* __restrict__ labels, int64_t W, const int32_t* __restrict__ boxes,
                  int L, uint8_t* __restrict__ fg, uint8_t* __restrict__ bg,
                  int32_t* __restrict__ counts) {
  __shared__ uint32_t red[2][kThreads / 32];
  const int64_t m = blockIdx.x;
  const int top = boxes[2 * m], left = boxes[2 * m + 1];
  const int32_t* src = labels + (int64_t)top * W + left;
  uint8_t* f = fg + m * (int64_t)L * L;
  uint8_t* b = bg + m * (int64_t)L * L;
  uint32_t nf = 0, nb = 0;
  const int total = L * L;
  for (int i = threadIdx.x; i < total; i += kThreads) {
    const int row = i / L, col = i - row * L;
    const int lab = __ldg(src + (int64_t)row * W + col);
    const bool isf = lab == (int)m;   // find.py:582
    const bool isb = lab == -1;       // find.py:584
    f[i] = isf;
    b[i] = isb;
    nf += isf;
    nb += isb;
  }
  if (counts) {
    nf = __reduce_add_sync(0xffffffffu, nf);
    nb = __reduce_add_sync(0xffffffffu, nb);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = nf; red[1][threadIdx.x >> 5] = nb; }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t a = 0, c = 0;
      for (int i = 0; i < kThreads / 32; ++i) { a += red[0][i]; c += red[1][i]; }
      counts[2 * m] = (int32_t)a;
      counts[2 * m + 1] = (int32_t)c;
    }
  }
}

}  // namespace mgb

using namespace mgb;

extern "C" {

int mgb_chip_masks(const int32_t* rel, const int32_t* r_fg, int r_inner, int r_outer, int64_t M,
                   int L, uint8_t* fg, uint8_t* bg, int32_t* counts, void* stream) {
  if (M < 0 || L <= 0 || L > 4096) return MGB_EINVAL;
  if (M == 0) return MGB_OK;
  if (M > INT32_MAX) return MGB_EUNSUPPORTED;
  if (!rel || !r_fg || !fg || !bg) return MGB_EINVAL;
  chip_masks_kernel<<<(unsigned)M, kThreads, 0, (cudaStream_t)stream>>>(rel, r_fg, r_inner, r_outer,
                                                                       L, fg, bg, counts);
  MGB_CUDA_LAUNCH_CHECK();
  return MGB_OK;
}

// Host restatement of the reference's perimeter walk (utils.py:433-465): start at (0, -r), emit
// the 8 mirror images of each octant point, step right while inside the circle, otherwise step
// in (diagonally).  filled_circle_points (utils.py:398-430) fills every row between its
// outermost perimeter pixels, so the disc is the span |dcol| <= hw[|drow|].
int mgb_disc_halfwidths(int r, int32_t* host_hw) {
  if (r < 1 || !host_hw) return MGB_EINVAL;
  for (int i = 0; i <= r; ++i) host_hw[i] = 0;
  auto mark = [&](long long row, long long col) {
    const long long ar = row < 0 ? -row : row, ac = col < 0 ? -col : col;
    if (ac > host_hw[ar]) host_hw[ar] = (int32_t)ac;
  };
  mark(0, -r); mark(-r, 0); mark(0, r); mark(r, 0);
  long long a = 1, b = -(long long)r;
  const long long r2 = (long long)r * r;
  while (a < -b) {
    mark(a, b); mark(b, a);   // the other six mirror images have the same (|row|, |col|)
    if (a * a + b * b - r2 <= 0) {
      ++a;
    } else {
      ++b;
      ++a;
    }
  }
  if (b == -a) mark(a, b);
  return MGB_OK;
}

int mgb_bead_labels(const int32_t* beads, int64_t M, int64_t H, int64_t W, const int32_t* hw,
                    int rmax, int32_t* labels, void* stream) {
  if (M < 0 || H <= 0 || W <= 0 || rmax < 0) return MGB_EINVAL;
  if (!labels) return MGB_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  MGB_CUDA_TRY(cudaMemsetAsync(labels, 0xff, (size_t)H * W * sizeof(int32_t), st));  // -1
  if (M == 0) return MGB_OK;
  if (M > INT32_MAX) return MGB_EUNSUPPORTED;
  if (!beads || !hw) return MGB_EINVAL;
  bead_labels_kernel<<<(unsigned)M, 128, 0, st>>>(beads, H, W, hw, rmax, labels);
  MGB_CUDA_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_bead_masks(const int32_t* labels, int64_t H, int64_t W, const int32_t* boxes, int64_t M,
                   int L, uint8_t* fg, uint8_t* bg, int32_t* counts, void* stream) {
  if (M < 0 || L <= 0 || L > 4096 || H < L || W < L) return MGB_EINVAL;
  if (M == 0) return MGB_OK;
  if (M > INT32_MAX) return MGB_EUNSUPPORTED;
  if (!labels || !boxes || !fg || !bg) return MGB_EINVAL;
  bead_masks_kernel<<<(unsigned)M, kThreads, 0, (cudaStream_t)stream>>>(labels, W, boxes, L, fg, bg,
                                                                       counts);
  MGB_CUDA_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_copy2d_async(void* dst, int64_t dst_pitch_bytes, const void* src, int64_t src_pitch_bytes,
                     int64_t width_bytes, int64_t rows, int kind, void* stream) {
  if (rows < 0 || width_bytes < 0 || dst_pitch_bytes < width_bytes || src_pitch_bytes < width_bytes) return MGB_EINVAL;
  if (rows == 0 || width_bytes == 0) return MGB_OK;
  if (!dst || !src || kind < 1 || kind > 3) return MGB_EINVAL;
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : (kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
  MGB_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)dst_pitch_bytes, src, (size_t)src_pitch_bytes, (size_t)width_bytes,
                                 (size_t)rows, k, (cudaStream_t)stream));
  return MGB_OK;
}

int mgb_abi_version(void) { return MGB_ABI_VERSION; }

int mgb_l2_persist(void* stream, const void* base, int64_t bytes, float hit_ratio) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaStreamAttrValue attr = {};
  if (!base || bytes <= 0) {           // reset: no window, persisting lines handed back
    attr.accessPolicyWindow.num_bytes = 0;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    MGB_CUDA_TRY(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
    MGB_CUDA_TRY(cudaCtxResetPersistingL2Cache());
    return MGB_OK;
  }
  int dev = 0, max_persist = 0, max_window = 0;
  MGB_CUDA_TRY(cudaGetDevice(&dev));
  MGB_CUDA_TRY(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
  MGB_CUDA_TRY(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
  if (max_persist <= 0 || max_window <= 0) return MGB_EUNSUPPORTED;
  MGB_CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist));
  attr.accessPolicyWindow.base_ptr = const_cast<void*>(base);
  attr.accessPolicyWindow.num_bytes = (size_t)(bytes < max_window ? bytes : max_window);
  attr.accessPolicyWindow.hitRatio = hit_ratio;
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  MGB_CUDA_TRY(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
  return MGB_OK;
}

static unsigned long long g_launches = 0;
void mgb_count_launch_(void) { __atomic_add_fetch(&g_launches, 1ULL, __ATOMIC_RELAXED); }
int64_t mgb_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

const char* mgb_error_string(int code) {
  switch (code) {
    case MGB_OK: return "ok";
    case MGB_EINVAL: return "invalid argument";
    case MGB_EALIGN: return "pointer or pitch not aligned for the vectorised path";
    case MGB_EUNSUPPORTED: return "shape, dtype or file encoding not supported by this build";
    case MGB_EIO: return "file could not be opened or read";
    case MGB_EFORMAT: return "not a TIFF/BigTIFF file or inconsistent directory";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown magnify_b200 error";
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// filter_nonround (reference filter.py:40-62): perimeter of every marker's foreground mask as
// cv.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) + cv.arcLength(closed) measure it.
// One warp per mask, lane 0 runs Suzuki & Abe's border following (Algorithm 1, 8-connected) on a
// zero-framed copy of the mask in shared memory: raster scan, outer / hole border starts, the
// parent rule (only outer borders whose parent is the frame are "external"), clockwise search for
// the first neighbour, counter-clockwise tracing, the NBD / -NBD marks.  The length of a closed
// border is the sum of its chain steps (1 or sqrt 2); CHAIN_APPROX_SIMPLE only drops collinear
// points, so arcLength of the compressed polygon is the same number up to float32 rounding of its
// segments (<= 6e-8 relative).
// ---------------------------------------------------------------------------------------------
namespace mgb {

// LT = label type: int16 in shared memory (L <= 160: at most L*L + 1 < 32768 borders), int32 in a
// global workspace for larger masks.  f: P*P labels, parent: one per border, is_hole: one byte per border.
template <typename LT>
__device__ __forceinline__ double trace_mask(LT* f, LT* parent, uint8_t* is_hole, const uint8_t* __restrict__ mask,
                                             int L) {
  const int P = L + 2;                       // framed side
  for (int i = threadIdx.x; i < P * P; i += 32) {
    const int y = i / P - 1, x = i % P - 1;
    f[i] = (y >= 0 && y < L && x >= 0 && x < L && mask[y * L + x]) ? 1 : 0;
  }
  __syncwarp();
  if (threadIdx.x != 0) return 0.0;
  const int di[8] = {0, -1, -1, -1, 0, 1, 1, 1};      // counter-clockwise from east (rows grow downwards)
  const int dj[8] = {1, 1, 0, -1, -1, -1, 0, 1};
  int nbd = 1;
  parent[1] = 0;
  is_hole[1] = 1;                                     // the frame counts as a hole border
  double total = 0.0;
  for (int i = 1; i <= L; ++i) {
    int lnbd = 1;
    for (int j = 1; j <= L; ++j) {
      const int v = f[i * P + j];
      if (v == 0) continue;
      int from = -1;                                  // direction of (i2, j2) seen from (i, j)
      bool hole = false;
      if (v == 1 && f[i * P + j - 1] == 0) {
        from = 4;                                     // outer border, start looking from the west pixel
      } else if (v >= 1 && f[i * P + j + 1] == 0) {
        from = 0;                                     // hole border, start from the east pixel
        hole = true;
        if (v > 1) lnbd = v;
      }
      if (from >= 0) {
        ++nbd;
        is_hole[nbd] = hole;
        parent[nbd] = (is_hole[lnbd] != hole) ? lnbd : parent[lnbd];
        const bool external = !hole && parent[nbd] == 1;
        // (3.1) clockwise from (i2, j2): first non-zero neighbour
        int first = -1;
        for (int k = 0; k < 8; ++k) {
          const int d = (from - k + 8) & 7;
          if (f[(i + di[d]) * P + j + dj[d]] != 0) {
            first = d;
            break;
          }
        }
        if (first < 0) {
          f[i * P + j] = (LT)(-nbd);             // isolated pixel: a one-point contour of length 0
        } else {
          const int i1 = i + di[first], j1 = j + dj[first];
          int i2 = i1, j2 = j1, i3 = i, j3 = j;
          double length = 0.0;
          for (;;) {
            // (3.3) counter-clockwise from the element after (i2, j2): first non-zero neighbour of (i3, j3)
            int d0 = 0;
            for (int d = 0; d < 8; ++d)
              if (i3 + di[d] == i2 && j3 + dj[d] == j2) d0 = d;
            bool east_zero = false;
            int d4 = d0;
            for (int k = 1; k <= 8; ++k) {
              const int d = (d0 + k) & 7;
              if (f[(i3 + di[d]) * P + j3 + dj[d]] != 0) {
                d4 = d;
                break;
              }
              if (d == 0) east_zero = true;
            }
            const int i4 = i3 + di[d4], j4 = j3 + dj[d4];
            // (3.4)
            if (east_zero) f[i3 * P + j3] = (LT)(-nbd);
            else if (f[i3 * P + j3] == 1) f[i3 * P + j3] = (LT)nbd;
            length += (d4 & 1) ? 1.4142135623730951 : 1.0;
            // (3.5)
            if (i4 == i && j4 == j && i3 == i1 && j3 == j1) break;
            i2 = i3; j2 = j3; i3 = i4; j3 = j4;
          }
          if (external) total += length;
        }
      }
      // (4)
      const int now = f[i * P + j];
      if (now != 1) lnbd = now < 0 ? -now : now;
    }
  }
  return total;
}

__global__ void __launch_bounds__(32) mask_perimeter_kernel(const uint8_t* __restrict__ masks, int L,
                                                            double* __restrict__ perimeter) {
  extern __shared__ int16_t sm[];
  const int P = L + 2;
  int16_t* parent = sm + P * P;
  uint8_t* is_hole = reinterpret_cast<uint8_t*>(parent + (L * L + 4));   // every pixel starts at most one border
  const double total = trace_mask<int16_t>(sm, parent, is_hole, masks + (int64_t)blockIdx.x * L * L, L);
  if (threadIdx.x == 0) perimeter[blockIdx.x] = total;
}

// The same walk with int32 labels in global memory (one workspace slice per mask of the batch).
__global__ void __launch_bounds__(32) mask_perimeter_big_kernel(const uint8_t* __restrict__ masks, int L,
                                                                int32_t* __restrict__ work, int64_t slice_words,
                                                                double* __restrict__ perimeter) {
  const int P = L + 2, borders = L * L + 4;
  int32_t* f = work + (int64_t)blockIdx.x * slice_words;
  int32_t* parent = f + P * P;
  uint8_t* is_hole = reinterpret_cast<uint8_t*>(parent + borders);
  const double total = trace_mask<int32_t>(f, parent, is_hole, masks + (int64_t)blockIdx.x * L * L, L);
  if (threadIdx.x == 0) perimeter[blockIdx.x] = total;
}


}  // namespace mgb

extern "C" int mgb_mask_perimeters(const uint8_t* masks, int64_t M, int L, double* perimeter, void* stream) {
  if (M < 0 || L <= 0 || L > 4096) return MGB_EINVAL;
  if (M == 0) return MGB_OK;
  if (!masks || !perimeter) return MGB_EINVAL;
  if (M > INT32_MAX) return MGB_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = L + 2, borders = L * L + 4;
  if (L <= 160) {   // int16 labels and the tables fit shared memory
    const size_t bytes = (size_t)P * P * 2 + (size_t)borders * 2 + (size_t)borders + 16;
    if (bytes > 48 * 1024)
      MGB_CUDA_TRY(cudaFuncSetAttribute(mgb::mask_perimeter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    mgb::mask_perimeter_kernel<<<(unsigned)M, 32, bytes, st>>>(masks, L, perimeter);
    MGB_CUDA_LAUNCH_CHECK();
    return MGB_OK;
  }
  // larger masks (max_bead_diameter > 80, chamber_diameter > 133): int32 labels in a stream-ordered
  // workspace of at most 256 MB, the masks in batches
  const int64_t slice_words = (int64_t)P * P + borders + (borders + 3) / 4 + 4;
  const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(M, (256ll << 20) / (slice_words * 4)));
  void* work = nullptr;
  MGB_CUDA_TRY(mgb::scratch_alloc(&work, (size_t)(batch * slice_words * 4), st));
  for (int64_t first = 0; first < M; first += batch) {
    const int64_t n = std::min(batch, M - first);
    mgb::mask_perimeter_big_kernel<<<(unsigned)n, 32, 0, st>>>(masks + first * (int64_t)L * L, L, (int32_t*)work,
                                                               slice_words, perimeter + first);
    mgb_count_launch_();
  }
  const cudaError_t e = cudaGetLastError();
  cudaFreeAsync(work, st);
  return e == cudaSuccess ? MGB_OK : (int)e;
}
