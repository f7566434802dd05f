// Shared pieces of the staged ROI-gather kernels (roi_tma.cu, roi_lists.cu): the launch
// parameters, mbarrier / TMA wrappers, and the copy loops that re-align a window staged in
// shared memory and store it with aligned vector stores.
#pragma once
#include <cstdlib>
#include <type_traits>

#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace mgb {

constexpr int kTmaMaxWarps = 8;

struct TmaGatherParams {
  uint16_t* roi;            // may be null (summaries only)
  const int32_t* boxes;     // (M,T,2) in elements
  const int32_t* order;     // (M) processing order of the markers, or null
  const int32_t* mask_t;    // (T) or null when !stats
  const uint8_t* fg;        // (M,Tm,rows,rows) or null
  const uint8_t* bg;
  double* stats;            // (M,C,T,8) or null
  double* peer_stats[8];    // peer-mapped copies of the gathered (ranks,M,C,T,8) buffer, this rank's block
  int n_peers;              // 0: write only `stats`
  int64_t C, T, Tm;
  int rows;                 // L
  int wu;                   // row length in 16-bit units (L * unit)
  int wpu;                  // TMA box width: roundup8(wu + 7) units
  int unit;                 // 16-bit units per element
  int n_stages;
  int loader;               // 0 = TMA tensor copy, 1 = cp.async chunks (sector granularity)
  const uint16_t* image;    // 16-bit units, for the cp.async loader
  int64_t H, Wu;
  int stage_bytes;          // rows * wpu * 2 rounded up to 128
  uint32_t vpr;             // 16-byte vectors per roi row (wu / 8) when the vector path applies, else 0
  uint32_t magic_vpr;       // ceil(2^32 / vpr)
  uint32_t half;            // wu / 2 words per row when the word path applies, else 0
  uint32_t magic_half;
  uint32_t magic_wu;        // ceil(2^32 / wu)
  // masked-value lists (roi_lists.cu)
  int cap_f, cap_b;         // list capacities (entries), multiples of 32
  int want_median;          // 0: sums / means only
  int store;                // 0: summaries only (roi == null)
  int split_parts;          // CTA layout: CTAs sharing each marker from split_from on (1: none)
  int64_t split_from;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 8 consecutive 16-bit units starting S units into the 16 units of (a, b).
template <int S>
__device__ __forceinline__ uint4 shift_units(const uint4& a, const uint4& b) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  constexpr int o = S >> 1;
  if constexpr ((S & 1) == 0) {
    return make_uint4(w[o], w[o + 1], w[o + 2], w[o + 3]);
  } else {
    return make_uint4(__funnelshift_r(w[o], w[o + 1], 16), __funnelshift_r(w[o + 1], w[o + 2], 16),
                      __funnelshift_r(w[o + 2], w[o + 3], 16), __funnelshift_r(w[o + 3], w[o + 4], 16));
  }
}

// Summary write-out: one record of kStatsRec doubles per (marker, channel, time) =
// n_fg, n_bg, sum_fg, sum_bg, mean_fg, mean_bg, median_fg, median_bg (medians NaN when the
// kernel does not compute them).  With peers, the 64-byte record goes straight into this rank's
// block of the gathered buffer of EVERY rank (NVLink peer stores, one lane per peer): the
// per-marker summaries are all-gathered by the kernel that computes them, no separate collective.
constexpr int kStatsRec = 8;
__device__ __forceinline__ double nan_f64() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ void write_stats(const TmaGatherParams& p, int lane, int64_t n, double cf, double cb,
                                            double sf, double sb, double mf, double mb) {
  // called by the whole warp (all lanes hold the reduced values): lane 0 writes the local record,
  // or lane j writes the record to peer j (one 64-byte record = four 16-byte stores per peer)
  double* o = nullptr;
  if (p.n_peers == 0) {
    if (lane == 0) o = p.stats + n * kStatsRec;
  } else if (lane < p.n_peers) {
    o = p.peer_stats[lane] + n * kStatsRec;
  }
  if (o) {
    double2* o2 = reinterpret_cast<double2*>(o);
    o2[0] = make_double2(cf, cb);
    o2[1] = make_double2(sf, sb);
    o2[2] = make_double2(sf / cf, sb / cb);   // 0/0 = NaN like nanmean
    o2[3] = make_double2(mf, mb);
  }
}

// Vector path for one window: rows of wu = 8*vpr units, output 16-byte aligned.
// VPL > 0: the loop over this lane's vectors is fully unrolled and the lane's mask bytes and
// shared-memory offsets (which do not depend on the window) live in registers for the whole CTA.
template <bool STATS, bool STORE, int S, int VPL>
__device__ __forceinline__ void consume_vec(const TmaGatherParams& p, const uint8_t* buf, uint16_t* dst,
                                            const uint8_t* fgm, const uint8_t* bgm, int lane,
                                            const uint2* fm, const uint2* bm, const uint32_t* svo,
                                            uint32_t* sf_out, uint32_t* sb_out) {
  const uint4* s4 = reinterpret_cast<const uint4*>(buf);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  const uint32_t nvec = p.rows * p.vpr;
  uint32_t sf = 0, sb = 0;
  if constexpr (VPL > 0) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const uint32_t v = lane + 32 * k;
      if (v < nvec) {
        const uint4 a = s4[svo[k]];
        const uint4 b = (S != 0) ? s4[svo[k] + 1] : a;
        const uint4 d = shift_units<S>(a, b);
        if constexpr (STORE) stg_stream(d4 + v, d);
        if constexpr (STATS) {
          sf = __dp2a_lo(d.x, fm[k].x, sf); sf = __dp2a_hi(d.y, fm[k].x, sf);
          sf = __dp2a_lo(d.z, fm[k].y, sf); sf = __dp2a_hi(d.w, fm[k].y, sf);
          sb = __dp2a_lo(d.x, bm[k].x, sb); sb = __dp2a_hi(d.y, bm[k].x, sb);
          sb = __dp2a_lo(d.z, bm[k].y, sb); sb = __dp2a_hi(d.w, bm[k].y, sb);
        }
      }
    }
  } else {
    const uint2* f2 = reinterpret_cast<const uint2*>(fgm);
    const uint2* b2 = reinterpret_cast<const uint2*>(bgm);
    const uint32_t pitch = p.wpu >> 3;
#pragma unroll 4
    for (uint32_t v = lane; v < nvec; v += 32) {
      const uint32_t row = __umulhi(v, p.magic_vpr);
      const uint32_t sv = row * pitch + (v - row * p.vpr);
      const uint4 a = s4[sv];
      const uint4 b = (S != 0) ? s4[sv + 1] : a;
      const uint4 d = shift_units<S>(a, b);
      if constexpr (STORE) stg_stream(d4 + v, d);
      if constexpr (STATS) {
        const uint2 f = f2[v];
        const uint2 g = b2[v];
        sf = __dp2a_lo(d.x, f.x, sf); sf = __dp2a_hi(d.y, f.x, sf);
        sf = __dp2a_lo(d.z, f.y, sf); sf = __dp2a_hi(d.w, f.y, sf);
        sb = __dp2a_lo(d.x, g.x, sb); sb = __dp2a_hi(d.y, g.x, sb);
        sb = __dp2a_lo(d.z, g.y, sb); sb = __dp2a_hi(d.w, g.y, sb);
      }
    }
  }
  *sf_out = sf;
  *sb_out = sb;
}

// Word / unit path for rows that are not a whole number of 16-byte vectors (e.g. L = 50).
template <bool STATS, bool STORE>
__device__ __forceinline__ void consume_generic(const TmaGatherParams& p, const uint8_t* buf, uint16_t* dst,
                                                const uint8_t* fgm, const uint8_t* bgm, int lane, int shift,
                                                uint32_t* sf_out, uint32_t* sb_out) {
  const uint16_t* s16 = reinterpret_cast<const uint16_t*>(buf);
  uint32_t sf = 0, sb = 0;
  if (p.half) {
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
    const uint16_t* f16 = reinterpret_cast<const uint16_t*>(fgm);
    const uint16_t* b16 = reinterpret_cast<const uint16_t*>(bgm);
    const uint32_t words = p.rows * p.half;
#pragma unroll 4
    for (uint32_t w = lane; w < words; w += 32) {
      const uint32_t row = __umulhi(w, p.magic_half);
      const uint32_t u = row * p.wpu + shift + 2 * (w - row * p.half);
      const uint32_t d = (uint32_t)s16[u] | ((uint32_t)s16[u + 1] << 16);
      if constexpr (STORE) d32[w] = d;
      if constexpr (STATS) {
        sf = __dp2a_lo(d, (uint32_t)f16[w], sf);
        sb = __dp2a_lo(d, (uint32_t)b16[w], sb);
      }
    }
  } else {
    const int total = p.rows * p.wu;
    for (int e = lane; e < total; e += 32) {
      const int row = e / p.wu;
      const uint32_t d = s16[row * p.wpu + shift + (e - row * p.wu)];
      if constexpr (STORE) dst[e] = (uint16_t)d;
      if constexpr (STATS) {
        sf += fgm[e] ? d : 0u;
        sb += bgm[e] ? d : 0u;
      }
    }
  }
  *sf_out = sf;
  *sb_out = sb;
}

__device__ __forceinline__ void cp_async16_cg(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}


template <int PAR>
__device__ __forceinline__ uint32_t load_pair(const uint32_t* s32, uint32_t unit_off) {
  const uint32_t w = unit_off >> 1;
  if constexpr (PAR == 0) {
    return s32[w];
  } else {
    return __funnelshift_r(s32[w], s32[w + 1], 16);
  }
}

}  // namespace mgb
