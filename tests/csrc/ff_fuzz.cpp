// Host fuzz of the flat-field fast path against the reference's exact operation order.
// Build: g++ -O2 -o ff_fuzz ff_fuzz.cpp   (driven by tests/test_ff_fastpath_host.py)
// Exit code 0 and a line "ok <n_checked> <n_slow>" when every pixel that the fast path accepts
// equals the exact value.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../magnify_b200/csrc/ff_core.cuh"

using namespace mgb;

static long long n_checked = 0, n_slow = 0, n_bad = 0;

static void check(uint16_t x, double f, double d, double M, double M2) {
  double g, b;
  ff_make_coeffs(f, d, M, M2, &g, &b);
  bool slow = false;
  int o = ff_fast_px(kFFHiBase | x, g, b, &slow);
  {  // accepted coefficients keep s inside [2^20, 2^21): the kernel relies on it
    double s = fma(ff_from_hi(kFFHiBase | x), g, b);
    if (!(s >= 1048576.0 && s < 2097152.0)) { ++n_bad; return; }
  }
  ++n_checked;
  if (slow) { ++n_slow; return; }
  uint16_t e = ff_exact_u16(x, f, d, M, M2);
  if ((uint16_t)o != e) {
    if (n_bad < 10) std::fprintf(stderr, "MISMATCH x=%u f=%.17g d=%.17g M=%.17g M2=%.17g fast=%d exact=%u\n", x, f, d, M, M2, o, e);
    ++n_bad;
  }
}

int main(int argc, char** argv) {
  long long n = argc > 1 ? std::atoll(argv[1]) : 20000000;
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> uf(0.3, 2.5), ud(-50.0, 400.0), u01(0.0, 1.0);
  for (long long it = 0; it < n / 64; ++it) {
    // One "dataset": maxima consistent with some maximal pixel.
    double fmax_pix = uf(rng), dmax_pix = ud(rng);
    uint16_t xtop = (uint16_t)(rng() % 65536);
    double M = (double)xtop - ud(rng); if (M < 1.0) M = 1.0 + u01(rng) * 60000.0;
    double M2 = M / fmax_pix * (1.0 + u01(rng));
    (void)dmax_pix;
    for (int j = 0; j < 64; ++j) {
      double f = uf(rng), d = ud(rng);
      uint16_t x = (uint16_t)(rng() % 65536);
      int kind = (int)(rng() % 8);
      if (kind == 0) { f = 1.0; }
      if (kind == 1) { d = (double)(rng() % 300); }
      if (kind == 2) { f = 1.0; d = (double)(rng() % 300); M2 = M; }
      if (kind == 3) { f = 0.5; d = 0.0; M2 = 2 * M; }
      if (kind == 4) { x = (uint16_t)(d > 0 ? (unsigned)d + (rng() % 3) : 0); }
      check(x, f, d, M, M2);
    }
  }
  // Degenerate coefficients must always be flagged slow.
  {
    double g, b; bool slow;
    ff_make_coeffs(0.0, 1.0, 10.0, 5.0, &g, &b); slow = false; ff_fast_px(kFFHiBase | 7, g, b, &slow); if (!slow) ++n_bad;
    ff_make_coeffs(1e-9, 1.0, 10.0, 5.0, &g, &b); slow = false; ff_fast_px(kFFHiBase | 7, g, b, &slow); if (!slow) ++n_bad;
    ff_make_coeffs(1.0, 1e9, 10.0, 5.0, &g, &b); slow = false; ff_fast_px(kFFHiBase | 7, g, b, &slow); if (!slow) ++n_bad;
    ff_make_coeffs(1.0, 0.0, 0.0, 0.0, &g, &b); slow = false; ff_fast_px(kFFHiBase | 7, g, b, &slow); if (!slow) ++n_bad;
  }
  std::printf("%s %lld %lld\n", n_bad ? "FAIL" : "ok", n_checked, n_slow);
  return n_bad ? 1 : 0;
}
