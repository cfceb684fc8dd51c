#include <mutex>
// Device-resident divide & conquer for the real symmetric tridiagonal eigenproblem (see stedc.h).
// Tree: leaves of <= 32 rows solved by one warp each (implicit QL), then log2(n/32) levels of Cuppen merges; all
// merges of one level run batched in the same launches, data-dependent sizes (deflation counts) stay on the device.
#include "stedc.h"

#include <algorithm>
#include <cfloat>

namespace tevo {
namespace {

constexpr double EPS = 2.220446049250313e-16;

// ---------------------------------------------------------------------------------------------------
__global__ void tnorm_kernel(int n, const double* __restrict__ d, const double* __restrict__ e, double* scale) {
  __shared__ double red[32];
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    m = fmax(m, fabs(d[i]));
    if (i < n - 1) m = fmax(m, fabs(e[i]));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) {
      scale[0] = (v > 0.0 && isfinite(v)) ? 1.0 / v : 1.0;  // multiply by this
      scale[1] = (v > 0.0 && isfinite(v)) ? v : 1.0;        // multiply back by this
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// leaf solve: one warp per leaf; each lane keeps a private copy of (d,e) and its own row of the eigenvector block
__global__ void __launch_bounds__(32) leaf_kernel(int n, const double* __restrict__ d, const double* __restrict__ e,
                                                  const double* __restrict__ scale, double* __restrict__ D,
                                                  double* __restrict__ Z, long ld) {
  constexpr int L = Stedc::LEAF;
  const int lo = blockIdx.x * L;
  const int sz = min(L, n - lo);
  const int lane = threadIdx.x;
  const double sc = scale[0];
  double sd[L], se[L], zr[L];
  for (int i = 0; i < L; ++i) {
    double dv = 0.0, ev = 0.0;
    if (i < sz) {
      dv = d[lo + i] * sc;
      if (i == 0 && lo > 0) dv -= fabs(e[lo - 1] * sc);
      if (i == sz - 1 && lo + sz < n) dv -= fabs(e[lo + sz - 1] * sc);
      if (i < sz - 1) ev = e[lo + i] * sc;
    }
    sd[i] = dv;
    se[i] = ev;
    zr[i] = (i == lane) ? 1.0 : 0.0;
  }
  for (int l = 0; l < sz; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < sz - 1; ++m) {
        double dd = fabs(sd[m]) + fabs(sd[m + 1]);
        if (fabs(se[m]) <= EPS * dd) break;
      }
      if (m != l) {
        if (iter++ == 80) break;
        double g = (sd[l + 1] - sd[l]) / (2.0 * se[l]);
        double r = hypot(g, 1.0);
        g = sd[m] - sd[l] + se[l] / (g + copysign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        // the rotation chain is the critical path of the kernel (one warp per leaf, every lane the same scalars): the
        // entries travelling down the chain stay in registers (one load + one store of d, e, z per step instead of
        // two), and c, s, r come from one reciprocal square root + a Newton step instead of hypot and two divisions
        // (the matrix is scaled to unit max-norm, so f^2 + g^2 cannot overflow; tiny sums take the hypot path)
        double zhi = zr[m], dnext = sd[m];
        for (i = m - 1; i >= l; --i) {
          const double ei = se[i], di = sd[i];
          const double f = s * ei, b = c * ei;
          const double r2 = fma(f, f, g * g);
          if (r2 > 1e-280) {
            double ri = rsqrt(r2);
            ri = ri * fma(-0.5 * r2 * ri, ri, 1.5);
            r = r2 * ri;
            s = f * ri;
            c = g * ri;
          } else {
            r = hypot(f, g);
            if (r == 0.0) {
              se[i + 1] = 0.0;
              sd[i + 1] = dnext - p;
              se[m] = 0.0;
              break;
            }
            s = f / r;
            c = g / r;
          }
          se[i + 1] = r;
          g = dnext - p;
          r = (di - g) * s + 2.0 * c * b;
          p = s * r;
          sd[i + 1] = g + p;
          g = c * r - b;
          dnext = di;
          const double zlo = zr[i];
          zr[i + 1] = s * zlo + c * zhi;
          zhi = c * zlo - s * zhi;
        }
        zr[i + 1] = zhi;  // i + 1 = l after a full chain, the break position otherwise
        if (r == 0.0 && i >= l) continue;
        sd[l] -= p;
        se[l] = g;
        se[m] = 0.0;
      }
    } while (m != l);
  }
  // ascending order by rank counting (all lanes hold identical sd[])
  for (int j = 0; j < sz; ++j) {
    int rank = 0;
    for (int k = 0; k < sz; ++k) rank += (sd[k] < sd[j] || (sd[k] == sd[j] && k < j)) ? 1 : 0;
    if (lane < sz) Z[(long)(lo + rank) * ld + lo + lane] = zr[j];
    if (lane == 0) D[lo + rank] = sd[j];
  }
}

// inclusive scan of one int per thread over the CTA (blockDim.x a multiple of 32, <= 1024); sm: 32 ints
template <bool MAX>
__device__ inline int block_scan_incl(int v, int* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = MAX ? max(v, u) : v + u;
  }
  if (lane == 31) sm[w] = v;
  __syncthreads();
  if (w == 0) {
    int t = (lane < (int)(blockDim.x >> 5)) ? sm[lane] : (MAX ? -1 : 0);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t = MAX ? max(t, u) : t + u;
    }
    sm[lane] = t;
  }
  __syncthreads();
  if (w > 0) v = MAX ? max(v, sm[w - 1]) : v + sm[w - 1];
  __syncthreads();
  return v;
}

// ---------------------------------------------------------------------------------------------------
// merge step 1: z vector, sort by eigenvalue, deflation, gather order.  Deflation of tiny z components, the search for
// close pairs and the compaction run on all threads; only when a close pair exists does thread 0 walk the sorted list
// (the walk is sequential because a deflated pair changes the partner of the next test) - round 2 measured the
// always-sequential version at ~0.25 us per entry, 0.7 ms per n = 2048 factorisation.
__global__ void __launch_bounds__(1024)
merge_prepare_kernel(const MergeDesc* __restrict__ descs, const double* __restrict__ e, const double* __restrict__ scale,
                     const double* __restrict__ D, const double* __restrict__ Z, long ld, double* __restrict__ dd,
                     double* __restrict__ zz, double* __restrict__ ddefl, int* __restrict__ gsrc, int* __restrict__ rot_a,
                     int* __restrict__ rot_b, double* __restrict__ rot_c, double* __restrict__ rot_s, int* __restrict__ Kout,
                     int* __restrict__ nrot_out, double* __restrict__ rho_out, int Nmax) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* uD = reinterpret_cast<double*>(smraw);
  double* uz = uD + Nmax;
  double* sD = uz + Nmax;
  double* sz = sD + Nmax;
  int* sidx = reinterpret_cast<int*>(sz + Nmax);
  int* sflag = sidx + Nmax;
  __shared__ double red[64];
  const MergeDesc md = descs[blockIdx.x];
  const int lo = md.lo, mid = md.mid, hi = md.hi, n1 = mid - lo, N = hi - lo;
  const double es = e[mid - 1] * scale[0];
  const double rho = 2.0 * fabs(es);
  const double sgn = es >= 0.0 ? 1.0 : -1.0;
  const double rs2 = 0.70710678118654752440;
  double dmax = 0.0, zmax = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double dv = D[lo + i];
    double zv = (i < n1 ? Z[(long)(lo + i) * ld + (mid - 1)] : sgn * Z[(long)(lo + i) * ld + mid]) * rs2;
    uD[i] = dv;
    uz[i] = zv;
    dmax = fmax(dmax, fabs(dv));
    zmax = fmax(zmax, fabs(zv));
  }
  for (int o = 16; o > 0; o >>= 1) {
    dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = dmax;
    red[32 + (threadIdx.x >> 5)] = zmax;
  }
  __syncthreads();
  // stable merge of the two ascending lists
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double v = uD[i];
    int rank;
    if (i < n1) {
      int a = n1, b = N;  // first index in [n1,N) with uD >= v
      while (a < b) {
        int c = (a + b) >> 1;
        if (uD[c] < v) a = c + 1; else b = c;
      }
      rank = i + (a - n1);
    } else {
      int a = 0, b = n1;  // first index in [0,n1) with uD > v
      while (a < b) {
        int c = (a + b) >> 1;
        if (uD[c] <= v) a = c + 1; else b = c;
      }
      rank = (i - n1) + a;
    }
    sD[rank] = v;
    sz[rank] = uz[i];
    sidx[rank] = i;
    sflag[rank] = 0;
  }
  __syncthreads();
  __shared__ int sscan[32];
  __shared__ int s_any, s_nrot;
  double dm = 0.0, zm = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
    dm = fmax(dm, red[w]);
    zm = fmax(zm, red[32 + w]);
  }
  const double tol = 8.0 * EPS * fmax(dm, zm);
  const bool all_deflated = rho * zm <= tol;
  // contiguous items per thread: [j0, j1)
  const int ipt = (N + (int)blockDim.x - 1) / (int)blockDim.x;
  const int j0 = min(N, (int)threadIdx.x * ipt), j1 = min(N, j0 + ipt);
  if (threadIdx.x == 0) {
    s_any = 0;
    s_nrot = 0;
  }
  // (1) tiny components; last surviving index of the own items
  int last = -1;
  for (int j = j0; j < j1; ++j) {
    const int f = (all_deflated || rho * fabs(sz[j]) <= tol) ? 1 : 0;
    sflag[j] = f;
    if (!f) last = j;
  }
  // (2) close pairs among neighbouring survivors: |t c s| <= tol with c = z_j / tau, s = z_p / tau, tested without the
  // division as |t z_j z_p| <= tol (z_j^2 + z_p^2) (a hair generous: the walk below repeats the exact test)
  {
    const int incl = block_scan_incl<true>(last, sscan);                 // last survivor up to and including own items
    int prev = __shfl_up_sync(0xffffffffu, incl, 1);                     // ... before the own items
    if ((threadIdx.x & 31) == 0) prev = (threadIdx.x == 0) ? -1 : sscan[(threadIdx.x >> 5) - 1];
    bool any = false;
    for (int j = j0; j < j1; ++j) {
      if (sflag[j]) continue;
      if (prev >= 0) {
        const double zj = sz[j], zp = sz[prev], t = sD[j] - sD[prev];
        if (fabs(t * zj * zp) <= 1.000001 * tol * (zj * zj + zp * zp)) any = true;
      }
      prev = j;
    }
    if (any) s_any = 1;
  }
  __syncthreads();
  if (s_any && threadIdx.x == 0) {
    int nrot = 0, prev = -1;
    double zp = 0.0, dp = 0.0;
    for (int j = 0; j < N; ++j) {
      if (sflag[j]) continue;
      const double zj = sz[j], dj = sD[j];
      bool merged = false;
      if (prev >= 0) {
        const double t = dj - dp;
        if (fabs(t * zj * zp) <= 1.000001 * tol * (zj * zj + zp * zp)) {
          double s2 = zp, c = zj;
          const double tau = hypot(c, s2);
          c /= tau;
          s2 = -s2 / tau;
          if (fabs(t * c * s2) <= tol) {
            sz[j] = tau;
            sz[prev] = 0.0;
            rot_a[lo + nrot] = lo + sidx[prev];
            rot_b[lo + nrot] = lo + sidx[j];
            rot_c[lo + nrot] = c;
            rot_s[lo + nrot] = s2;
            ++nrot;
            const double tt = dp * c * c + dj * s2 * s2;
            sD[j] = dp * s2 * s2 + dj * c * c;
            sD[prev] = tt;
            sflag[prev] = 1;
            zp = tau;
            dp = sD[j];
            merged = true;
          }
        }
      }
      if (!merged) {
        zp = zj;
        dp = dj;
      }
      prev = j;
    }
    s_nrot = nrot;
  }
  __syncthreads();
  // (3) compaction: survivors first (ascending), deflated entries behind them, both in list order
  int cnt = 0;
  for (int j = j0; j < j1; ++j) cnt += sflag[j] ? 0 : 1;
  const int incl = block_scan_incl<false>(cnt, sscan);
  __shared__ int s_K;
  if (threadIdx.x == blockDim.x - 1) s_K = incl;
  __syncthreads();
  const int K = s_K;
  int t0 = incl - cnt;       // survivors before the own items
  int t1 = K + (j0 - t0);    // deflated entries before the own items
  for (int j = j0; j < j1; ++j) {
    if (!sflag[j]) {
      dd[lo + t0] = sD[j];
      zz[lo + t0] = sz[j];
      gsrc[lo + t0] = lo + sidx[j];
      ++t0;
    } else {
      ddefl[lo + t1] = sD[j];
      gsrc[lo + t1] = lo + sidx[j];
      ++t1;
    }
  }
  if (threadIdx.x == 0) {
    Kout[blockIdx.x] = K;
    nrot_out[blockIdx.x] = s_nrot;
    rho_out[blockIdx.x] = rho;
  }
}

// merge step 2: apply the deflation Givens rotations to the eigenvector columns, one thread per row
__global__ void rot_apply_kernel(const MergeDesc* __restrict__ descs, const int* __restrict__ nrot_in,
                                 const int* __restrict__ rot_a, const int* __restrict__ rot_b,
                                 const double* __restrict__ rot_c, const double* __restrict__ rot_s, double* __restrict__ Z,
                                 long ld) {
  const MergeDesc md = descs[blockIdx.y];
  const int nrot = nrot_in[blockIdx.y];
  if (nrot == 0) return;
  const int r = md.lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= md.hi) return;
  for (int q = 0; q < nrot; ++q) {
    const long a = rot_a[md.lo + q], b = rot_b[md.lo + q];
    const double c = rot_c[md.lo + q], s = rot_s[md.lo + q];
    const double x = Z[a * ld + r], y = Z[b * ld + r];
    Z[a * ld + r] = c * x + s * y;
    Z[b * ld + r] = -s * x + c * y;
  }
}

// merge step 3: Zt(:, lo+t) = Z(:, gsrc[lo+t]) on the rows of the node
__global__ void gather_kernel(const MergeDesc* __restrict__ descs, const int* __restrict__ gsrc,
                              const double* __restrict__ Z, double* __restrict__ Zt, long ld) {
  const MergeDesc md = descs[blockIdx.z];
  const int N = md.hi - md.lo;
  const int t = blockIdx.y;
  if (t >= N) return;
  const long src = gsrc[md.lo + t];
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x)
    Zt[(long)(md.lo + t) * ld + md.lo + r] = Z[src * ld + md.lo + r];
}

// merge step 4: secular equation, one warp per root
__device__ inline double secular_f(int K, const double* sdd, const double* sz2, double rho, double dorg, double mu,
                                   int lane) {
  double s = 0.0;
  for (int i = lane; i < K; i += 32) s += sz2[i] / ((sdd[i] - dorg) - mu);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return 1.0 + rho * s;
}

// f and df/dmu in one pass (Newton steps of the root finder)
__device__ inline void secular_fd(int K, const double* sdd, const double* sz2, double rho, double dorg, double mu,
                                  int lane, double& f, double& fp) {
  double s = 0.0, sp = 0.0;
  for (int i = lane; i < K; i += 32) {
    const double t = 1.0 / ((sdd[i] - dorg) - mu);
    const double zt = sz2[i] * t;
    s += zt;
    sp = fma(zt, t, sp);
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sp += __shfl_xor_sync(0xffffffffu, sp, o);
  }
  f = 1.0 + rho * s;
  fp = rho * sp;
}

__global__ void __launch_bounds__(256)
secular_kernel(const MergeDesc* __restrict__ descs, const int* __restrict__ Kin, const double* __restrict__ rho_in,
               const double* __restrict__ dd, const double* __restrict__ zz, int* __restrict__ org,
               double* __restrict__ mu, int Nmax) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* sdd = reinterpret_cast<double*>(smraw);
  double* sz2 = sdd + Nmax;
  const MergeDesc md = descs[blockIdx.y];
  const int K = Kin[blockIdx.y];
  if ((int)(blockIdx.x * 8) >= K) return;
  const double rho = rho_in[blockIdx.y];
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    sdd[i] = dd[md.lo + i];
    double z = zz[md.lo + i];
    sz2[i] = z * z;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= K) return;
  int o;
  double hi, sg;
  if (j < K - 1) {
    const double delta = sdd[j + 1] - sdd[j];
    const double fm = secular_f(K, sdd, sz2, rho, sdd[j], 0.5 * delta, lane);
    if (fm > 0.0) {
      o = j;
      sg = 1.0;
    } else if (fm < 0.0) {
      o = j + 1;
      sg = -1.0;
    } else {
      if (lane == 0) {
        org[md.lo + j] = j;
        mu[md.lo + j] = 0.5 * delta;
      }
      return;
    }
    hi = 0.5 * delta;
  } else {
    o = j;
    sg = 1.0;
    double s = 0.0;
    for (int i = lane; i < K; i += 32) s += sz2[i];
    for (int q = 16; q > 0; q >>= 1) s += __shfl_xor_sync(0xffffffffu, s, q);
    hi = rho * s;
  }
  const double dorg = sdd[o];
  // g(x) = sg * f(org, sg*x) is increasing in x on (0, hi], g(0+) = -inf
  double x = hi, xlo = 0.0;
  bool done = false;
  if (j == K - 1) {
    if (sg * secular_f(K, sdd, sz2, rho, dorg, sg * x, lane) < 0.0) done = true;  // root at the upper end (rounding)
  }
  double res = hi;
  if (!done) {
    for (int it = 0; it < 90; ++it) {
      const double xt = x * 0.0625;
      if (xt == 0.0) {
        xlo = 0.0;
        break;
      }
      if (sg * secular_f(K, sdd, sz2, rho, dorg, sg * xt, lane) <= 0.0) {
        xlo = xt;
        break;
      }
      x = xt;
    }
    double xhi = x;
    // g(xlo) <= 0 < g(xhi) (xlo = 0: the pole at the origin).  Newton steps on g, safeguarded as in "rtsafe": a step
    // that leaves the bracket, or that does not at least halve the previous one, is replaced by a bisection
    // (geometric while the bracket spans more than a factor 4).  Plain bisection to the last bit (round 1) took ~60
    // evaluations of the K-term sum per root; this takes ~8.
    double xc = (xlo > 0.0) ? sqrt(xlo) * sqrt(xhi) : 0.5 * xhi;
    double gv, gp;
    secular_fd(K, sdd, sz2, rho, dorg, sg * xc, lane, gv, gp);
    gv *= sg;  // dg/dx = f'(mu) > 0
    double dxold = xhi - xlo, dx = dxold;
    res = xc;
    for (int it = 0; it < 120; ++it) {
      if (gv == 0.0) {
        res = xc;
        break;
      }
      if (gv < 0.0)
        xlo = xc;
      else
        xhi = xc;
      const double step = gv / gp;
      const double cand = xc - step;
      const bool newton_ok = (cand > xlo) && (cand < xhi) && (fabs(2.0 * gv) <= fabs(dxold * gp));
      double xn;
      dxold = dx;
      if (newton_ok) {
        dx = step;
        xn = cand;
      } else {
        xn = (xlo > 0.0 && xhi > 4.0 * xlo) ? sqrt(xlo) * sqrt(xhi) : 0.5 * (xlo + xhi);
        dx = xn - xc;
      }
      if (!(xn > xlo) || !(xn < xhi)) {  // no representable point left inside the bracket
        res = 0.5 * (xlo + xhi);
        break;
      }
      res = xn;
      if (fabs(xn - xc) <= 2.0 * EPS * xn) break;  // converged to rounding
      xc = xn;
      secular_fd(K, sdd, sz2, rho, dorg, sg * xc, lane, gv, gp);
      gv *= sg;
    }
  }
  if (lane == 0) {
    org[md.lo + j] = o;
    mu[md.lo + j] = sg * res;
  }
}

// merge step 5: Gu-Eisenstat z-hat, one warp per component
__global__ void __launch_bounds__(256)
zhat_kernel(const MergeDesc* __restrict__ descs, const int* __restrict__ Kin, const double* __restrict__ rho_in,
            const double* __restrict__ dd, const double* __restrict__ zz, const int* __restrict__ org,
            const double* __restrict__ mu, double* __restrict__ zhat, int Nmax) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* sdd = reinterpret_cast<double*>(smraw);
  double* sdo = sdd + Nmax;  // dd[org[j]]
  double* smu = sdo + Nmax;
  const MergeDesc md = descs[blockIdx.y];
  const int K = Kin[blockIdx.y];
  if ((int)(blockIdx.x * 8) >= K) return;
  const double rho = rho_in[blockIdx.y];
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    sdd[i] = dd[md.lo + i];
    sdo[i] = dd[md.lo + org[md.lo + i]];
    smu[i] = mu[md.lo + i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= K) return;
  const double di = sdd[i];
  double p = 1.0;
  for (int j = lane; j < K; j += 32) {
    const double ndelta = smu[j] - (di - sdo[j]);  // lam_j - d_i
    if (j == i)
      p *= ndelta;
    else
      p *= ndelta / (sdd[j] - di);
  }
  for (int o = 16; o > 0; o >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, o);
  if (lane == 0) zhat[md.lo + i] = copysign(sqrt(fabs(p) / rho), zz[md.lo + i]);
}

// merge step 6: eigenvectors of the rank-one modified diagonal matrix, one CTA per column
__global__ void __launch_bounds__(128)
secvec_kernel(const MergeDesc* __restrict__ descs, const int* __restrict__ Kin, const double* __restrict__ dd,
              const int* __restrict__ org, const double* __restrict__ mu, const double* __restrict__ zhat,
              double* __restrict__ Um, long ld) {
  __shared__ double red[32];
  const MergeDesc md = descs[blockIdx.y];
  const int K = Kin[blockIdx.y];
  const int j = blockIdx.x;
  if (j >= K) return;
  const double dorg = dd[md.lo + org[md.lo + j]], m = mu[md.lo + j];
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const double u = zhat[md.lo + i] / ((dd[md.lo + i] - dorg) - m);
    v[0] += u * u;
  }
  block_sum<1>(v, red);
  const double inv = 1.0 / sqrt(v[0]);
  double* col = Um + (long)(md.lo + j) * ld + md.lo;
  for (int i = threadIdx.x; i < K; i += blockDim.x)
    col[i] = inv * zhat[md.lo + i] / ((dd[md.lo + i] - dorg) - m);
}

// merge step 8: final ascending order of roots + deflated values (rank by counting), eigenvalue write-back
__global__ void __launch_bounds__(1024)
finalize_kernel(const MergeDesc* __restrict__ descs, const int* __restrict__ Kin, const double* __restrict__ dd,
                const int* __restrict__ org, const double* __restrict__ mu, const double* __restrict__ ddefl,
                double* __restrict__ D, int* __restrict__ fpos, int Nmax) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* v = reinterpret_cast<double*>(smraw);
  const MergeDesc md = descs[blockIdx.x];
  const int K = Kin[blockIdx.x];
  const int N = md.hi - md.lo;
  for (int t = threadIdx.x; t < N; t += blockDim.x)
    v[t] = (t < K) ? dd[md.lo + org[md.lo + t]] + mu[md.lo + t] : ddefl[md.lo + t];
  __syncthreads();
  // the ranking is O(N^2): the CTAs of a merge (gridDim.y) each rank a slice of the values
  for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < N; t += gridDim.y * blockDim.x) {
    const double x = v[t];
    int rank = 0;
    for (int k = 0; k < N; ++k) rank += (v[k] < x || (v[k] == x && k < t)) ? 1 : 0;
    D[md.lo + rank] = x;
    fpos[md.lo + t] = md.lo + rank;
  }
}

// merge step 9: scatter updated / deflated columns to their final positions
__global__ void scatter_kernel(const MergeDesc* __restrict__ descs, const int* __restrict__ Kin,
                               const int* __restrict__ fpos, const double* __restrict__ Cw,
                               const double* __restrict__ Zt, double* __restrict__ Z, long ld) {
  const MergeDesc md = descs[blockIdx.z];
  const int N = md.hi - md.lo;
  const int t = blockIdx.y;
  if (t >= N) return;
  const int K = Kin[blockIdx.z];
  const double* src = (t < K ? Cw : Zt) + (long)(md.lo + t) * ld + md.lo;
  double* dst = Z + (long)fpos[md.lo + t] * ld + md.lo;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) dst[r] = src[r];
}

__global__ void unscale_kernel(int n, const double* __restrict__ D, const double* __restrict__ scale,
                               double* __restrict__ evals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) evals[i] = D[i] * scale[1];
}

template <typename T>
void dev_alloc(T*& p, size_t n) {
  TEVO_CUDA_CHECK(cudaMalloc(&p, n * sizeof(T)));
}
template <typename T>
void dev_free(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

}  // namespace

void Stedc::destroy() {
  dev_free(Zt_); dev_free(Um_); dev_free(Cw_); dev_free(zz_); dev_free(dd_); dev_free(ddefl_); dev_free(mu_);
  dev_free(zhat_); dev_free(lam_); dev_free(rot_c_); dev_free(rot_s_); dev_free(scale_); dev_free(rho_);
  dev_free(org_); dev_free(gsrc_); dev_free(rot_a_); dev_free(rot_b_); dev_free(K_); dev_free(nrot_);
  for (auto& kv : plans_) dev_free(kv.second.dev);
  plans_.clear();
  cap_ = 0;
}

void Stedc::ensure(int n) {
  if (n <= cap_) return;
  const int old = cap_;
  (void)old;
  dev_free(Zt_); dev_free(Um_); dev_free(Cw_); dev_free(zz_); dev_free(dd_); dev_free(ddefl_); dev_free(mu_);
  dev_free(zhat_); dev_free(lam_); dev_free(rot_c_); dev_free(rot_s_); dev_free(scale_); dev_free(rho_);
  dev_free(org_); dev_free(gsrc_); dev_free(rot_a_); dev_free(rot_b_); dev_free(K_); dev_free(nrot_);
  const size_t nn = (size_t)n * n;
  dev_alloc(Zt_, nn); dev_alloc(Um_, nn); dev_alloc(Cw_, nn);
  dev_alloc(zz_, n); dev_alloc(dd_, n); dev_alloc(ddefl_, n); dev_alloc(mu_, n); dev_alloc(zhat_, n); dev_alloc(lam_, n);
  dev_alloc(rot_c_, n); dev_alloc(rot_s_, n); dev_alloc(scale_, 2); dev_alloc(rho_, n);
  dev_alloc(org_, n); dev_alloc(gsrc_, n); dev_alloc(rot_a_, n); dev_alloc(rot_b_, n); dev_alloc(K_, n); dev_alloc(nrot_, n);
  cap_ = n;
}

Stedc::Plan& Stedc::plan_for(int n) {
  auto it = plans_.find(n);
  if (it != plans_.end()) return it->second;
  Plan p;
  std::vector<int> bounds;
  for (int b = 0; b < n; b += LEAF) bounds.push_back(b);
  bounds.push_back(n);
  std::vector<MergeDesc> all;
  while (bounds.size() > 2) {
    std::vector<MergeDesc> lvl;
    std::vector<int> nb;
    size_t k = 0;
    for (; k + 2 < bounds.size(); k += 2) {
      lvl.push_back(MergeDesc{bounds[k], bounds[k + 1], bounds[k + 2]});
      nb.push_back(bounds[k]);
    }
    for (; k < bounds.size(); ++k) nb.push_back(bounds[k]);
    // nb now holds the starts of the merged nodes plus (if a node was left over) its start, and the end n
    if (nb.back() != n) nb.push_back(n);
    int mx = 0;
    for (auto& m : lvl) mx = std::max(mx, m.hi - m.lo);
    p.level_off.push_back((int)all.size());
    p.level_maxN.push_back(mx);
    p.levels.push_back(lvl);
    all.insert(all.end(), lvl.begin(), lvl.end());
    bounds = nb;
  }
  if (!all.empty()) {
    dev_alloc(p.dev, all.size());
    TEVO_CUDA_CHECK(cudaMemcpy(p.dev, all.data(), all.size() * sizeof(MergeDesc), cudaMemcpyHostToDevice));
  }
  auto res = plans_.emplace(n, std::move(p));
  return res.first->second;
}

void Stedc::run(cudaStream_t st, int n, const double* d, const double* e, double* evals, double* Z) {
  if (n <= 0) return;
  ensure(n);
  Plan& plan = plan_for(n);
  const long ld = n;
  static PerDeviceOnce once;
  once([&]() {
    TEVO_CUDA_CHECK(cudaFuncSetAttribute(merge_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    TEVO_CUDA_CHECK(cudaFuncSetAttribute(secular_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    TEVO_CUDA_CHECK(cudaFuncSetAttribute(zhat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    TEVO_CUDA_CHECK(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  });
  if (n > 4096) throw TevoError(1, "stedc: n > 4096 not supported");
  double* D = lam_;  // working eigenvalues (scaled)
  TEVO_CUDA_CHECK(cudaMemsetAsync(Z, 0, sizeof(double) * (size_t)n * n, st));
  tnorm_kernel<<<1, 1024, 0, st>>>(n, d, e, scale_);
  TEVO_LAUNCHED();
  const int nleaf = (n + LEAF - 1) / LEAF;
  leaf_kernel<<<nleaf, 32, 0, st>>>(n, d, e, scale_, D, Z, ld);
  TEVO_LAUNCHED();
  for (size_t lv = 0; lv < plan.levels.size(); ++lv) {
    const int nm = (int)plan.levels[lv].size();
    if (nm == 0) continue;
    const MergeDesc* descs = plan.dev + plan.level_off[lv];
    const int Nmax = plan.level_maxN[lv];
    merge_prepare_kernel<<<nm, 1024, (size_t)Nmax * 40, st>>>(descs, e, scale_, D, Z, ld, dd_, zz_, ddefl_, gsrc_, rot_a_,
                                                             rot_b_, rot_c_, rot_s_, K_, nrot_, rho_, Nmax);
    TEVO_LAUNCHED();
    rot_apply_kernel<<<dim3((Nmax + 255) / 256, nm), 256, 0, st>>>(descs, nrot_, rot_a_, rot_b_, rot_c_, rot_s_, Z, ld);
    TEVO_LAUNCHED();
    gather_kernel<<<dim3((Nmax + 255) / 256, Nmax, nm), 256, 0, st>>>(descs, gsrc_, Z, Zt_, ld);
    TEVO_LAUNCHED();
    secular_kernel<<<dim3((Nmax + 7) / 8, nm), 256, (size_t)Nmax * 16, st>>>(descs, K_, rho_, dd_, zz_, org_, mu_, Nmax);
    TEVO_LAUNCHED();
    zhat_kernel<<<dim3((Nmax + 7) / 8, nm), 256, (size_t)Nmax * 24, st>>>(descs, K_, rho_, dd_, zz_, org_, mu_, zhat_, Nmax);
    TEVO_LAUNCHED();
    secvec_kernel<<<dim3(Nmax, nm), 128, 0, st>>>(descs, K_, dd_, org_, mu_, zhat_, Um_, ld);
    TEVO_LAUNCHED();
    dgemm_merge(st, descs, K_, nm, Nmax, Zt_, Um_, Cw_, ld);
    {
      const int fthreads = Nmax >= 1024 ? 256 : 128;
      const int fslices = std::max(1, std::min((Nmax + fthreads - 1) / fthreads, 160 / nm));
      finalize_kernel<<<dim3(nm, fslices), fthreads, (size_t)Nmax * 8, st>>>(descs, K_, dd_, org_, mu_, ddefl_, D, gsrc_,
                                                                            Nmax);
    }
    TEVO_LAUNCHED();
    scatter_kernel<<<dim3((Nmax + 255) / 256, Nmax, nm), 256, 0, st>>>(descs, K_, gsrc_, Cw_, Zt_, Z, ld);
    TEVO_LAUNCHED();
  }
  unscale_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, D, scale_, evals);
  TEVO_LAUNCHED();
}

}  // namespace tevo
