// Host build of noname_b200/csrc/nnpm_butterfly.cuh (test infrastructure): every radix, both signs, against the DFT
// definition evaluated in long double.  Prints one line per case: "R sign kind max_abs_err".
#include <cmath>
#include <cstdio>
#include <cstdlib>

#define __device__
#define __host__
#define __forceinline__ inline
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
using std::fma;

#include "../../noname_b200/csrc/nnpm_butterfly.cuh"

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static double rnd() {
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (double)(rng_state >> 11) / 9007199254740992.0 - 0.5;
}

template <int R, int SIGN>
static void one() {
  double2 x[R], a[R];
  long double Xr[R], Xi[R];
  const long double tau = 6.283185307179586476925286766559L;
  for (int n = 0; n < R; n++) x[n] = make_double2(rnd(), rnd());
  for (int k = 0; k < R; k++) {
    Xr[k] = Xi[k] = 0;
    for (int n = 0; n < R; n++) {
      const long double ph = SIGN * tau * (long double)((n * k) % R) / R;
      const long double c = cosl(ph), s = sinl(ph);
      Xr[k] += x[n].x * c - x[n].y * s;
      Xi[k] += x[n].x * s + x[n].y * c;
    }
  }
  constexpr int bits = zf_log2(R);
  // DIF: natural in, bit-reversed out
  for (int n = 0; n < R; n++) a[n] = x[n];
  ZDif<R, 0, SIGN>::run(a);
  double e = 0;
  for (int k = 0; k < R; k++) {
    const double2 v = a[zf_brev(k, bits)];
    e = fmax(e, fmax(fabs((double)(v.x - Xr[k])), fabs((double)(v.y - Xi[k]))));
  }
  printf("%d %d dif %.3e\n", R, SIGN, e);
  // DIT: bit-reversed in, natural out
  for (int n = 0; n < R; n++) a[zf_brev(n, bits)] = x[n];
  ZDit<R, 0, SIGN>::run(a);
  e = 0;
  for (int k = 0; k < R; k++) e = fmax(e, fmax(fabs((double)(a[k].x - Xr[k])), fabs((double)(a[k].y - Xi[k]))));
  printf("%d %d dit %.3e\n", R, SIGN, e);
  // forward DIF then inverse DIT on the bit-reversed data: R * x, no reordering in between (how the fused z pass uses them)
  for (int n = 0; n < R; n++) a[n] = x[n];
  ZDif<R, 0, SIGN>::run(a);
  ZDit<R, 0, -SIGN>::run(a);
  e = 0;
  for (int n = 0; n < R; n++) e = fmax(e, fmax(fabs(a[n].x - R * x[n].x), fabs(a[n].y - R * x[n].y)));
  printf("%d %d roundtrip %.3e\n", R, SIGN, e);
  // a sub-transform at a register offset (BASE = R): the second of two transforms held by one thread (k_rowfft pass B)
  double2 b[2 * R];
  for (int n = 0; n < 2 * R; n++) b[n] = make_double2(rnd(), rnd());
  for (int n = 0; n < R; n++) a[n] = b[R + n];
  ZDif<R, R, SIGN>::run(b);
  ZDif<R, 0, SIGN>::run(a);
  e = 0;
  for (int n = 0; n < R; n++) e = fmax(e, fmax(fabs(a[n].x - b[R + n].x), fabs(a[n].y - b[R + n].y)));
  printf("%d %d base %.3e\n", R, SIGN, e);
}

int main() {
  // twiddle table: cos / sin of 2 pi m / 32 for any integer m
  double e = 0;
  for (int m = -70; m <= 70; m++) {
    e = fmax(e, fabs((double)(zf_cos32(m) - cosl(6.283185307179586476925286766559L * m / 32))));
    e = fmax(e, fabs((double)(zf_sin32(m) - sinl(6.283185307179586476925286766559L * m / 32))));
  }
  printf("0 0 twiddles %.3e\n", e);
  one<2, -1>(); one<2, +1>(); one<4, -1>(); one<4, +1>(); one<8, -1>(); one<8, +1>();
  one<16, -1>(); one<16, +1>(); one<32, -1>(); one<32, +1>();
  return 0;
}
