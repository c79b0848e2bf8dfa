/* CPU ORACLE (test infrastructure, not product code) -- scalar C restatement of the loops
 * of yipihey/NoName src/ParticleMesh.jl, one function per reference function, loop order and
 * floating-point evaluation order as written there.  The FFTs between them are done by the
 * Python driver (oracle/pm_ref_c.py) with scipy's pocketfft, standing in for Julia Base FFTW.
 *
 * Used by: tests/ (cross-check of oracle/pm_ref.py), bench.py cpu_baseline and --impl reference.
 * PARITY UNPINNED: the reference has no tests or golden vectors (see oracle/pm_ref.py header).
 *
 * Layout is the Julia column-major memory as is: rho[i + N1*(j + N2*k)],
 * acc[c + 3*(i + N1*(j + N2*k))], x[d + rank*n], rt (complex, re/im interleaved)
 * [i + (N1/2+1)*(j + N2*k)].  Indices are 0-based (reference index minus one).
 *
 * Every function is serial, as the reference is.  Mesh loops take a [k0,k1) plane range and
 * particle loops take a pointer + count, so the Python driver can (for the all-cores arm of
 * bench.py --impl reference only) run disjoint chunks on a thread pool; libgomp is not in
 * this image.  The threaded 3-D CIC deposit bins the particle indices by the z-slab that owns
 * their cell (pmref_slab_count / pmref_slab_fill, stable), then every thread scatters its own
 * list into its own planes and one private ghost plane (pmref_cicdensity_slab): no atomics, no
 * per-thread copy of rho, and the summation order of every cell off a slab boundary is the
 * serial one.  2-D uses compare-and-swap adds (pmref_cicdensity_atomic); the threaded NGP
 * deposit keeps one private rho per thread.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef int64_t i64;

int pmref_version(void) { return 1; }

/* ParticleMesh.jl:409-415 */
void pmref_make_periodic(double *x, i64 n) {
  for (i64 i = 0; i < n; i++) {
    if (x[i] < 0.) x[i] += 1;
    if (x[i] > 1.) x[i] -= 1;
  }
}

/* index/offset triple of :386-388 shifted to 0-based; index N (x == 1.0) maps to 0 */
static inline void cell(double xv, i64 n, i64 *i0, i64 *ip, double *d) {
  double t = xv * (double)n;
  i64 ii = (i64)floor(t) + 1;
  *d = t - (double)ii + 1;
  i64 z = (ii - 1) % n;
  if (z < 0) z += n;
  *i0 = z;
  *ip = (z == n - 1) ? 0 : z + 1;
}

/* ParticleMesh.jl:355-407 (rank 2 and 3); rho must be zeroed by the caller (:290-292) */
void pmref_cicdensity(double *rho, const double *x, double m, i64 npart, int rank, i64 n1, i64 n2,
                      i64 n3) {
  if (rank == 2) {
    for (i64 n = 0; n < npart; n++) {
      i64 i, ip, j, jp;
      double dx, dy;
      cell(x[0 + 2 * n], n1, &i, &ip, &dx);
      cell(x[1 + 2 * n], n2, &j, &jp, &dy);
      rho[i + n1 * j] += m * (1. - dx) * (1. - dy);
      rho[i + n1 * jp] += m * (1. - dx) * dy;
      rho[ip + n1 * j] += m * dx * (1. - dy);
      rho[ip + n1 * jp] += m * dx * dy;
    }
    return;
  }
  for (i64 n = 0; n < npart; n++) {
    i64 i, ip, j, jp, k, kp;
    double dx, dy, dz;
    cell(x[0 + 3 * n], n1, &i, &ip, &dx);
    cell(x[1 + 3 * n], n2, &j, &jp, &dy);
    cell(x[2 + 3 * n], n3, &k, &kp, &dz);
    rho[i + n1 * (j + n2 * k)] += m * (1 - dx) * (1 - dy) * (1 - dz);
    rho[i + n1 * (jp + n2 * k)] += m * (1 - dx) * dy * (1 - dz);
    rho[ip + n1 * (j + n2 * k)] += m * dx * (1 - dy) * (1 - dz);
    rho[ip + n1 * (jp + n2 * k)] += m * dx * dy * (1 - dz);
    rho[i + n1 * (j + n2 * kp)] += m * (1 - dx) * (1 - dy) * dz;
    rho[i + n1 * (jp + n2 * kp)] += m * (1 - dx) * dy * dz;
    rho[ip + n1 * (j + n2 * kp)] += m * dx * (1 - dy) * dz;
    rho[ip + n1 * (jp + n2 * kp)] += m * dx * dy * dz;
  }
}

/* threads > 1 only: the same scatter with atomic adds into a rho shared by all threads */
static inline void atomic_add_f64(double *p, double v) {
  uint64_t *q = (uint64_t *)p;
  uint64_t old = __atomic_load_n(q, __ATOMIC_RELAXED), neu;
  do {
    double d;
    __builtin_memcpy(&d, &old, 8);
    d += v;
    __builtin_memcpy(&neu, &d, 8);
  } while (!__atomic_compare_exchange_n(q, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}
void pmref_cicdensity_atomic(double *rho, const double *x, double m, i64 npart, int rank, i64 n1, i64 n2,
                             i64 n3) {
  if (rank == 2) {
    for (i64 n = 0; n < npart; n++) {
      i64 i, ip, j, jp;
      double dx, dy;
      cell(x[0 + 2 * n], n1, &i, &ip, &dx);
      cell(x[1 + 2 * n], n2, &j, &jp, &dy);
      atomic_add_f64(&rho[i + n1 * j], m * (1. - dx) * (1. - dy));
      atomic_add_f64(&rho[i + n1 * jp], m * (1. - dx) * dy);
      atomic_add_f64(&rho[ip + n1 * j], m * dx * (1. - dy));
      atomic_add_f64(&rho[ip + n1 * jp], m * dx * dy);
    }
    return;
  }
  for (i64 n = 0; n < npart; n++) {
    i64 i, ip, j, jp, k, kp;
    double dx, dy, dz;
    cell(x[0 + 3 * n], n1, &i, &ip, &dx);
    cell(x[1 + 3 * n], n2, &j, &jp, &dy);
    cell(x[2 + 3 * n], n3, &k, &kp, &dz);
    atomic_add_f64(&rho[i + n1 * (j + n2 * k)], m * (1 - dx) * (1 - dy) * (1 - dz));
    atomic_add_f64(&rho[i + n1 * (jp + n2 * k)], m * (1 - dx) * dy * (1 - dz));
    atomic_add_f64(&rho[ip + n1 * (j + n2 * k)], m * dx * (1 - dy) * (1 - dz));
    atomic_add_f64(&rho[ip + n1 * (jp + n2 * k)], m * dx * dy * (1 - dz));
    atomic_add_f64(&rho[i + n1 * (j + n2 * kp)], m * (1 - dx) * (1 - dy) * dz);
    atomic_add_f64(&rho[i + n1 * (jp + n2 * kp)], m * (1 - dx) * dy * dz);
    atomic_add_f64(&rho[ip + n1 * (j + n2 * kp)], m * dx * (1 - dy) * dz);
    atomic_add_f64(&rho[ip + n1 * (jp + n2 * kp)], m * dx * dy * dz);
  }
}

/* threads > 1, rank 3: slab-binned scatter.  Slab t of T owns the planes [n3 t / T, n3 (t+1) / T). */
static inline i64 plane_of(double x3, i64 n3) {
  i64 z = (i64)floor(x3 * (double)n3);
  if (z >= n3 || z < 0) {                   /* x == 1.0 and strays: the periodic image, as cell() */
    z %= n3;
    if (z < 0) z += n3;
  }
  return z;
}
/* owner[k] = slab of plane k, filled by the caller's first call (n3 entries) */
void pmref_slab_table(i64 n3, i64 T, int32_t *owner) {
  for (i64 t = 0; t < T; t++)
    for (i64 k = n3 * t / T; k < n3 * (t + 1) / T; k++) owner[k] = (int32_t)t;
}
void pmref_slab_count(const double *x, i64 lo, i64 hi, i64 n3, i64 T, const int32_t *owner, i64 *counts) {
  for (i64 t = 0; t < T; t++) counts[t] = 0;
  for (i64 n = lo; n < hi; n++) counts[owner[plane_of(x[2 + 3 * n], n3)]]++;
}
/* offsets[t] = where this chunk's first particle of slab t goes in idx; advanced as the chunk is written */
void pmref_slab_fill(const double *x, i64 lo, i64 hi, i64 n3, i64 T, const int32_t *owner, i64 *offsets, i64 *idx) {
  (void)T;
  for (i64 n = lo; n < hi; n++) idx[offsets[owner[plane_of(x[2 + 3 * n], n3)]]++] = n;
}
/* the particles idx[0 .. nidx) all have k in [k0, k1): adds to the planes k < k1 go to rho, adds to plane k1
 * (kp of the slab's top plane; plane 0 for the last slab) go to ghost[n1 * n2] */
void pmref_cicdensity_slab(double *rho, double *ghost, const double *x, const i64 *idx, i64 nidx, double m, i64 n1, i64 n2,
                           i64 n3, i64 k1) {
  for (i64 q = 0; q < nidx; q++) {
    const i64 n = idx[q];
    i64 i, ip, j, jp, k, kp;
    double dx, dy, dz;
    cell(x[0 + 3 * n], n1, &i, &ip, &dx);
    cell(x[1 + 3 * n], n2, &j, &jp, &dy);
    cell(x[2 + 3 * n], n3, &k, &kp, &dz);
    double *lo = rho + n1 * n2 * k;
    double *up = (k + 1 == k1) ? ghost : rho + n1 * n2 * kp;
    lo[i + n1 * j] += m * (1 - dx) * (1 - dy) * (1 - dz);
    lo[i + n1 * jp] += m * (1 - dx) * dy * (1 - dz);
    lo[ip + n1 * j] += m * dx * (1 - dy) * (1 - dz);
    lo[ip + n1 * jp] += m * dx * dy * (1 - dz);
    up[i + n1 * j] += m * (1 - dx) * (1 - dy) * dz;
    up[i + n1 * jp] += m * (1 - dx) * dy * dz;
    up[ip + n1 * j] += m * dx * (1 - dy) * dz;
    up[ip + n1 * jp] += m * dx * dy * dz;
  }
}

/* ParticleMesh.jl:326-353 */
void pmref_ngpdensity(double *rho, const double *x, double m, i64 npart, int rank, i64 n1, i64 n2,
                      i64 n3) {
  for (i64 n = 0; n < npart; n++) {
    i64 i, j, k = 0;
    if (rank == 2) {
      i = (i64)floor(x[0 + 2 * n] * (double)n1);
      j = (i64)floor(x[1 + 2 * n] * (double)n2);
    } else {
      i = (i64)floor(x[0 + 3 * n] * (double)n1 + 1.) - 1;
      j = (i64)floor(x[1 + 3 * n] * (double)n2 + 1.) - 1;
      k = (i64)floor(x[2 + 3 * n] * (double)n3 + 1.) - 1;
      k = ((k % n3) + n3) % n3;
    }
    i = ((i % n1) + n1) % n1;
    j = ((j % n2) + n2) % n2;
    rho[i + n1 * (j + n2 * k)] += m;
  }
}

/* mean and rho .- mean(rho), :768-769 / :858 */
double pmref_sum(const double *rho, i64 ncell) {
  double s = 0;
  for (i64 i = 0; i < ncell; i++) s += rho[i];
  return s;
}
void pmref_sub(double *out, const double *rho, double mean, i64 ncell) {
  for (i64 i = 0; i < ncell; i++) out[i] = rho[i] - mean;
}
void pmref_zero(double *dst, i64 n) {                     /* the zero fill of deposit, :290-292 */
  for (i64 i = 0; i < n; i++) dst[i] = 0.;
}
void pmref_add_into(double *dst, const double *src, i64 n) {
  for (i64 i = 0; i < n; i++) dst[i] += src[i];
}

/* the k-space loop of potential_3d :482-501 / potential_2d :451-465: three sin per element,
 * divide by Ginv when non-zero, zero the DC bin.  rt is complex interleaved. */
void pmref_green(double *rt, i64 n1, i64 n2, i64 n3, i64 k0, i64 k1) {
  const i64 nc = n1 / 2 + 1;
  const double twopi = 2 * M_PI;
  for (i64 k = k0; k < k1; k++) {
    double kz = twopi * (double)k / (double)n3;
    for (i64 j = 0; j < n2; j++) {
      double ky = twopi * (double)j / (double)n2;
      for (i64 i = 0; i < nc; i++) {
        double kx = twopi * (double)i / (double)n1;
        double sx = sin(kx / 2), sy = sin(ky / 2), sz = sin(kz / 2);
        double ginv = (n3 > 1) ? -4. * (sx * sx + sy * sy + sz * sz) : -4. * (sx * sx + sy * sy);
        if (!(ginv == 0.0)) {
          double *p = rt + 2 * (i + nc * (j + n2 * k));
          p[0] /= ginv;
          p[1] /= ginv;
        }
      }
    }
  }
  if (k0 == 0) {
    rt[0] = 0.;
    rt[1] = 0.;
  }
}

/* phi[:] = irfft(...) .* 1/Nglx^2 : the scaling pass of :466 / :502 */
void pmref_scale(double *phi, i64 ncell, double n1) {
  for (i64 i = 0; i < ncell; i++) phi[i] = phi[i] * 1 / (n1 * n1);
}

/* compute_acceleration :631-637 = deriv1_3d :549-573 (or deriv1_2d :528-546), scale pass,
 * negate pass -- three passes as the reference does them. */
void pmref_compute_acceleration(double *a, const double *phi, i64 n1, i64 n2, i64 n3, i64 k0,
                                i64 k1) {
  for (i64 k = k0; k < k1; k++) {
    i64 kp1 = (k == n3 - 1) ? 0 : k + 1, km1 = (k == 0) ? n3 - 1 : k - 1;
    for (i64 j = 0; j < n2; j++) {
      i64 jp1 = (j == n2 - 1) ? 0 : j + 1, jm1 = (j == 0) ? n2 - 1 : j - 1;
      for (i64 i = 0; i < n1; i++) {
        i64 ip1 = (i == n1 - 1) ? 0 : i + 1, im1 = (i == 0) ? n1 - 1 : i - 1;
        double *o = a + 3 * (i + n1 * (j + n2 * k));
        o[0] = (phi[ip1 + n1 * (j + n2 * k)] - phi[im1 + n1 * (j + n2 * k)]) / 2;
        o[1] = (phi[i + n1 * (jp1 + n2 * k)] - phi[i + n1 * (jm1 + n2 * k)]) / 2;
        if (n3 > 1) o[2] = (phi[i + n1 * (j + n2 * kp1)] - phi[i + n1 * (j + n2 * km1)]) / 2;
      }
    }
  }
  const double fac = (double)n1;
  const i64 lo = 3 * n1 * n2 * k0, hi = 3 * n1 * n2 * k1;
  for (i64 i = lo; i < hi; i++) a[i] *= fac;
  for (i64 i = lo; i < hi; i++) a[i] = -a[i];
}

/* cic_interp_3d :694-721 (intended k index) / cic_interp_2d :676-692 */
void pmref_cic_interp(double *pa, const double *a, const double *x, i64 npart, int rank, i64 n1,
                      i64 n2, i64 n3) {
  if (rank == 2) {
    for (i64 n = 0; n < npart; n++) {
      i64 i, ip, j, jp;
      double dx, dy;
      cell(x[0 + 2 * n], n1, &i, &ip, &dx);
      cell(x[1 + 2 * n], n2, &j, &jp, &dy);
      for (int m = 0; m < 2; m++)
        pa[m + 2 * n] = (a[m + 3 * (i + n1 * j)] * (1. - dx) * (1. - dy) +
                         a[m + 3 * (ip + n1 * j)] * dx * (1. - dy) +
                         a[m + 3 * (i + n1 * jp)] * (1. - dx) * dy + a[m + 3 * (ip + n1 * jp)] * dx * dy);
    }
    return;
  }
  for (i64 n = 0; n < npart; n++) {
    i64 i, ip, j, jp, k, kp;
    double dx, dy, dz;
    cell(x[0 + 3 * n], n1, &i, &ip, &dx);
    cell(x[1 + 3 * n], n2, &j, &jp, &dy);
    cell(x[2 + 3 * n], n3, &k, &kp, &dz);
#define A(m, ii, jj, kk) a[(m) + 3 * ((ii) + n1 * ((jj) + n2 * (kk)))]
    for (int m = 0; m < 3; m++)
      pa[m + 3 * n] =
          (A(m, i, j, k) * (1. - dx) * (1. - dy) * (1. - dz) + A(m, ip, j, k) * dx * (1. - dy) * (1. - dz) +
           A(m, i, jp, k) * (1. - dx) * dy * (1. - dz) + A(m, ip, jp, k) * dx * dy * (1. - dz) +
           A(m, i, j, kp) * (1. - dx) * (1. - dy) * dz + A(m, ip, j, kp) * dx * (1. - dy) * dz +
           A(m, i, jp, kp) * (1. - dx) * dy * dz + A(m, ip, jp, kp) * dx * dy * dz);
#undef A
  }
}

/* NGP gather :642-652 (intended k index) */
void pmref_ngp_interp(double *pa, const double *a, const double *x, i64 npart, int rank, i64 n1,
                      i64 n2, i64 n3) {
  for (i64 n = 0; n < npart; n++) {
    i64 i = (i64)floor(x[0 + rank * n] * (double)n1);
    i64 j = (i64)floor(x[1 + rank * n] * (double)n2);
    i64 k = (rank == 3) ? (i64)floor(x[2 + 3 * n] * (double)n3) : 0;
    i = ((i % n1) + n1) % n1;
    j = ((j % n2) + n2) % n2;
    k = ((k % n3) + n3) % n3;
    for (int m = 0; m < rank; m++) pa[m + rank * n] = a[m + 3 * (i + n1 * (j + n2 * k))];
  }
}

/* drift :723-728 */
void pmref_drift(double *x, const double *v, double dt, i64 n) {
  for (i64 i = 0; i < n; i++) x[i] = x[i] + dt * v[i];
}

/* kick :731-736 */
void pmref_kick(double *v, const double *a, double dt, i64 n) {
  for (i64 i = 0; i < n; i++) v[i] = v[i] + dt * a[i];
}

/* comoving kick :800-807, including the VelocityMidStep temporary the reference allocates */
void pmref_kick_comoving(double *v, const double *pa, double dt, double a, double dadt, i64 n) {
  double *vmid = (double *)malloc(sizeof(double) * (size_t)n);
  for (i64 i = 0; i < n; i++) vmid[i] = v[i] + pa[i] * dt / 2;
  for (i64 i = 0; i < n; i++) v[i] += (-vmid[i] * (dadt / a) + pa[i] / a) * dt;
  free(vmid);
}

/* maximum(abs(v)+1e-30), maximum(v) : :782, :875, :878 */
void pmref_maxima(const double *v, i64 n, double *max_abs_plus, double *max_signed) {
  double ma = -INFINITY, ms = -INFINITY;
  for (i64 i = 0; i < n; i++) {
    double t = fabs(v[i]) + 1e-30;
    if (t > ma) ma = t;
    if (v[i] > ms) ms = v[i];
  }
  *max_abs_plus = ma;
  *max_signed = ms;
}
