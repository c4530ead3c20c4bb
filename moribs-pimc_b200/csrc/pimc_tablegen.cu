// pimc_tablegen.cu -- rotational density-matrix TABLE GENERATORS on the device.
//
// Replaces the reference's three Fortran pre-processing programs:
//   nmv_prop/asymrho.f     asymmetric top (361 single-theta CPU jobs in the reference's workflow, nmv_prop/README)
//   symtop_prop/symrho.f   symmetric top
//   linear_prop/linden.f   linear rotor
//
// Not a port.  asymrho.f evaluates, for every grid point (phi,chi) of a theta plane, the quadruple sum
//   rho = sum_J (2J+1)/8pi^2 sum_n e^{-tau E_n} sum_{m,k} c^n_m c^n_k d^J_{mk}(theta) cos(m phi + k chi)      (:500-545)
// i.e. ~1e7 cosines per grid point.  Here the sum is factorised:
//   1. tg_eigen_kernel      rigid-rotor blocks are tridiagonal in the even-k / odd-k bases (rotmat :792-813):
//                           implicit QL per block, then S^J_w[m][k] = sum_n w_n c^n_m c^n_k, w = e^{-tau E}{1,E,E^2}
//                           (with the reference's "(2J+1)/8pi^2 e^{-tau E} < 1e-16 -> skip the state" rule, :520,597)
//   2. tg_asym_coeff_kernel A_w[m][k](theta) = sum_J (2J+1)/8pi^2 S^J_w[m][k] d^J_{mk}(theta); Wigner d by the upward
//                           three-term recurrence in J (FP64, stable) instead of Zare's alternating sum in real*16
//   3. tg_phi_kernel        T[k][phi] = sum_m A[m][k] {cos,sin}(m phi) -- the grid is integer degrees, so every
//                           trigonometric value is an exact table entry cos(n deg), n = (m*iphi) mod 360
//   4. tg_chi_gemm_kernel   raw[phi][chi] = sum_k T1[k][phi] cos(k chi) - T2[k][phi] sin(k chi): one FP64 GEMM over all
//                           (theta, weight, phi) rows, 128x128 CTA tiles, 8x8 register tiles
//   5. tg_asym_combine_kernel  even/odd parity combination and guards (:552-659) + the reference's symmetry fill
//                           (:665-709) as a gather through an index map replayed on the host
// About 45 GFLOP for the whole 181x361x361x3 table instead of ~2e14 cosine terms.
#include "../../include/pimcgpu.h"

#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

namespace pimc { int set_error(const char *msg); }

namespace {

int tfail(const char *fmt, ...)
{
   char buf[512];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof buf, fmt, ap);
   va_end(ap);
   return pimc::set_error(buf);
}
#define TCK(call)                                                                            \
   do {                                                                                      \
      cudaError_t e_ = (call);                                                               \
      if (e_ != cudaSuccess) return tfail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
   } while (0)

constexpr double TG_PI = 3.14159265358979323846;
constexpr double TG_BOLTZ = 0.6950356;       // asymrho.f:41, symrho.f:38
constexpr int NANG = 361;                    // phi, chi = 0..360 degrees
constexpr int NPLANE = NANG * NANG;

struct DevMem {                              // frees on scope exit
   std::vector<void *> p;
   ~DevMem() { for (void *q : p) cudaFree(q); }
   template <class T> cudaError_t get(T **ptr, size_t n, bool zero = false)
   {
      cudaError_t e = cudaMalloc((void **)ptr, std::max<size_t>(n, 1) * sizeof(T));
      if (e != cudaSuccess) return e;
      p.push_back(*ptr);
      return zero ? cudaMemset(*ptr, 0, std::max<size_t>(n, 1) * sizeof(T)) : cudaSuccess;
   }
};

int need_device(const char *who)
{
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return tfail("%s: no CUDA device (there is no CPU fallback)", who);
   return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Wigner d^j_{mk}(theta), convention of wigd (asymrho.f:1006-1039 = Zare Eq. 3.57 with (m,k) = (m',m)).
// Start value at j0 = max(|m|,|k|) is a single term of the sum: sqrt(C(2j0,b)) cos^a(theta/2) (+-sin(theta/2))^b, a+b = 2 j0;
// then   d^{j+1} = (2j+1)(j+1)/sqrt(((j+1)^2-m^2)((j+1)^2-k^2)) [ (cos theta - mk/(j(j+1))) d^j - sqrt((j^2-m^2)(j^2-k^2))/(j(2j+1)) d^{j-1} ].
// Going up in j the wanted solution is the growing one.  Values below 2^-500 carry a separate exponent so that a start value
// which underflows FP64 (large j0, theta near 0 or pi) still grows back into range.
struct WigD {
   double prev, cur, x;
   int e2, j, m, k;
   __device__ void start(int m_, int k_, double ch, double sh, double x_)
   {
      m = m_; k = k_; x = x_;
      const int am = abs(m), ak = abs(k);
      int a, b; bool neg;
      if (am >= ak) { j = am; if (m >= 0) { a = j + k; b = j - k; neg = (b & 1); } else { a = j - k; b = j + k; neg = false; } }
      else          { j = ak; if (k >= 0) { a = j + m; b = j - m; neg = false; }   else { a = j - m; b = j + m; neg = (b & 1); } }
      double v = 1.0;
      e2 = 0;
      const double tiny = ldexp(1.0, -500), huge = ldexp(1.0, 500);
      for (int i = 1; i <= b; i++) {
         v *= sqrt((double)(a + i) / (double)i) * sh;
         if (v < tiny && v > 0.0) { v *= huge; e2 -= 500; }
      }
      for (int i = 0; i < a; i++) {
         v *= ch;
         if (v < tiny && v > 0.0) { v *= huge; e2 -= 500; }
      }
      cur = neg ? -v : v;
      prev = 0.0;
   }
   __device__ double value() const { return e2 == 0 ? cur : (e2 < -1500 ? 0.0 : ldexp(cur, e2)); }
   __device__ void next()
   {
      double nx;
      if (j == 0) nx = x * cur;
      else {
         const double dj = j, j1 = j + 1, dm = m, dk = k;
         const double c1 = (2.0 * dj + 1.0) * j1 / sqrt((j1 * j1 - dm * dm) * (j1 * j1 - dk * dk));
         const double c0 = sqrt((dj * dj - dm * dm) * (dj * dj - dk * dk)) / (dj * (2.0 * dj + 1.0));
         nx = c1 * ((x - dm * dk / (dj * j1)) * cur - c0 * prev);
      }
      prev = cur; cur = nx; j++;
      if (e2 < 0 && fabs(cur) > ldexp(1.0, 400)) { cur *= ldexp(1.0, -500); prev *= ldexp(1.0, -500); e2 += 500; }
   }
};

// parity entry point: d[j][m+maxj][k+maxj] for all j <= maxj (zero outside |m|,|k| <= j)
__global__ void tg_wigner_kernel(int maxj, double ch, double sh, double x, double *d)
{
   const int w = 2 * maxj + 1;
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= w * w) return;
   const int m = t / w - maxj, k = t % w - maxj;
   WigD r;
   r.start(m, k, ch, sh, x);
   for (;;) {
      d[((size_t)r.j * w + (m + maxj)) * w + (k + maxj)] = r.value();
      if (r.j == maxj) break;
      r.next();
   }
}

// ------------------------------------------------------------------------------------------------------------
// Block bookkeeping for the asymmetric top: block b = (j, parity), dimension n, first k = kst (step 2)  (asymrho.f:127-137)
struct AsymBlock { int j, parity, n, kst; long off_e, off_z, off_s; };

__device__ __forceinline__ double tg_cplus(int j, int k)  { return (k >= j || k < -j) ? 0.0 : sqrt((double)j * (j + 1.0) - (double)k * (k + 1.0)); }

// One CTA per block: tridiagonal implicit QL (thread 0 generates the plane rotations of one sweep, every thread applies them
// to the eigenvector rows it owns), then the weighted projectors S_w = Z diag(w) Z^T.
__global__ void tg_eigen_kernel(const AsymBlock *blocks, double A, double B, double C, double tau, double *eng, double *Z, double *S, int *status)
{
   extern __shared__ double sh[];
   const AsymBlock bl = blocks[blockIdx.x];
   const int n = bl.n, j = bl.j;
   double *d = sh, *e = sh + n, *rc = sh + 2 * n, *rs = sh + 3 * n;
   __shared__ int s_m, s_lo, s_done, s_fail;
   double *z = Z + bl.off_z;                     // z[col*n + row]
   for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int k = bl.kst + 2 * i;
      d[i] = 0.5 * (A + C) * (double)(j * (j + 1)) + (B - 0.5 * (A + C)) * (double)(k * k);      // rotmat, k == kp
      e[i] = (i + 1 < n) ? 0.25 * (A - C) * tg_cplus(j, k) * tg_cplus(j, k + 1) : 0.0;            // rotmat, k' = k + 2
      for (int c = 0; c < n; c++) z[(size_t)c * n + i] = (c == i) ? 1.0 : 0.0;
   }
   if (threadIdx.x == 0) s_fail = 0;
   __syncthreads();
   for (int l = 0; l < n; l++) {
      for (int iter = 0;; iter++) {
         if (threadIdx.x == 0) {
            int m = l;
            for (; m < n - 1; m++) {
               const double dd = fabs(d[m]) + fabs(d[m + 1]);
               if (fabs(e[m]) <= DBL_EPSILON * dd) break;
            }
            s_m = m; s_done = (m == l); s_lo = l;
            if (m != l) {
               if (iter >= 60) { s_fail = 1; s_done = 1; }
               else {
                  double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                  double r = hypot(g, 1.0);
                  g = d[m] - d[l] + e[l] / (g + copysign(r, g));
                  double s = 1.0, c = 1.0, p = 0.0;
                  int i = m - 1;
                  for (; i >= l; i--) {
                     const double f = s * e[i], b = c * e[i];
                     r = hypot(f, g);
                     e[i + 1] = r;
                     if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
                     s = f / r; c = g / r;
                     g = d[i + 1] - p;
                     r = (d[i] - g) * s + 2.0 * c * b;
                     p = s * r;
                     d[i + 1] = g + p;
                     g = c * r - b;
                     rc[i] = c; rs[i] = s;
                  }
                  if (r == 0.0 && i >= l) s_lo = i + 1;          // rotations i+1 .. m-1 were generated
                  else { d[l] -= p; e[l] = g; e[m] = 0.0; }
               }
            }
         }
         __syncthreads();
         const int done = s_done, m = s_m, lo = s_lo;
         if (!done)
            for (int row = threadIdx.x; row < n; row += blockDim.x) {
               double hi = z[(size_t)m * n + row];
               for (int i = m - 1; i >= lo; i--) {
                  const double zi = z[(size_t)i * n + row];
                  z[(size_t)(i + 1) * n + row] = rs[i] * zi + rc[i] * hi;
                  hi = rc[i] * zi - rs[i] * hi;
               }
               z[(size_t)lo * n + row] = hi;
            }
         __syncthreads();
         if (done) break;
      }
      if (s_fail) break;
   }
   __syncthreads();
   if (s_fail) { if (threadIdx.x == 0) atomicExch(status, 1); return; }
   // weights (asymrho.f:510-520): states with (2J+1)/8pi^2 e^{-tau E} < 1e-16 do not contribute; rc/rs are reused
   const double pre = (double)(2 * j + 1) / (8.0 * TG_PI * TG_PI);
   double *w2 = e;
   for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double en = d[i];
      eng[bl.off_e + i] = en;
      double expo = exp(-tau * en);
      if (expo * pre < 1e-16) expo = 0.0;
      rc[i] = expo; rs[i] = expo * en; w2[i] = expo * en * en;
   }
   __syncthreads();
   double *s = S + bl.off_s;
   for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
      const int im = t / n, ik = t % n;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      for (int st = 0; st < n; st++) {
         const double cc = z[(size_t)st * n + im] * z[(size_t)st * n + ik];
         a0 += rc[st] * cc; a1 += rs[st] * cc; a2 += w2[st] * cc;
      }
      s[t] = a0; s[(size_t)n * n + t] = a1; s[2 * (size_t)n * n + t] = a2;
   }
}

// A[theta][parity][w][im][ik] (class-compact, np x np): class index i <-> m = -M + 2 i, M = largest |m| <= maxj of that parity.
__global__ void tg_asym_coeff_kernel(int maxj, int np, int ith0, const long *off_s /* [maxj+1][2] */, const double *S, double *Acoef)
{
   const int parity = blockIdx.y, ith = ith0 + blockIdx.z;
   const int M = ((maxj & 1) == parity) ? maxj : maxj - 1;
   if (M < 0) return;
   const int nc = M + 1;
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= nc * nc) return;
   const int im = t / nc, ik = t % nc;
   const int m = -M + 2 * im, k = -M + 2 * ik;
   const double ch = cospi((double)ith / 360.0), sh = sinpi((double)ith / 360.0), x = cospi((double)ith / 180.0);
   WigD r;
   r.start(m, k, ch, sh, x);
   double a0 = 0.0, a1 = 0.0, a2 = 0.0;
   for (;;) {
      const int j = r.j;
      const int kst = ((j & 1) == parity) ? -j : -j + 1;          // even class: -j (j even) / -j+1 ; odd class: -j+1 (j even) / -j
      const int n = (j - kst) / 2 + 1;
      const long o = off_s[2 * j + parity];
      const size_t idx = (size_t)((m - kst) / 2) * n + (k - kst) / 2;
      const double pre = (double)(2 * j + 1) / (8.0 * TG_PI * TG_PI), dv = r.value();
      a0 += pre * (S[o + idx] * dv);
      a1 += pre * (S[o + (size_t)n * n + idx] * dv);
      a2 += pre * (S[o + 2 * (size_t)n * n + idx] * dv);
      if (j == maxj) break;
      r.next();
   }
   double *A = Acoef + ((size_t)(blockIdx.z * 2 + parity) * 3) * np * np;
   A[(size_t)im * np + ik] = a0;
   A[(size_t)np * np + (size_t)im * np + ik] = a1;
   A[2 * (size_t)np * np + (size_t)im * np + ik] = a2;
}

// cos(n degrees), n = 0..359, exact to rounding
__device__ __forceinline__ void tg_load_ctab(double *ctab)
{
   for (int i = threadIdx.x; i < 360; i += blockDim.x) ctab[i] = cospi((double)i / 180.0);
}
__device__ __forceinline__ int tg_mod360(int v) { v %= 360; return v < 0 ? v + 360 : v; }

// Tt[parity][kk][row], row = (theta*3 + w)*361 + iphi, kk = ik (cos part) or np + ik (sin part); rows padded to Rpad
__global__ void tg_phi_kernel(int maxj, int np, long Rpad, const double *Acoef, double *Tt)
{
   __shared__ double ctab[360];
   tg_load_ctab(ctab);
   const int ik = blockIdx.x, tw = blockIdx.y, parity = blockIdx.z;      // tw = theta*3 + w
   const int M = ((maxj & 1) == parity) ? maxj : maxj - 1;
   __syncthreads();
   if (M < 0 || ik > M) return;
   const int nc = M + 1;
   const int theta = tw / 3, w = tw % 3;
   const double *A = Acoef + (((size_t)(theta * 2 + parity) * 3) + w) * np * np;
   for (int iphi = threadIdx.x; iphi < NANG; iphi += blockDim.x) {
      double tc = 0.0, ts = 0.0;
      for (int im = 0; im < nc; im++) {
         const int m = -M + 2 * im;
         const int a = tg_mod360(m * iphi);
         const double av = A[(size_t)im * np + ik];
         tc += av * ctab[a];
         ts += av * ctab[tg_mod360(a + 270)];
      }
      const long row = (long)tw * NANG + iphi;
      Tt[((size_t)parity * 2 * np + ik) * Rpad + row] = tc;
      Tt[((size_t)parity * 2 * np + np + ik) * Rpad + row] = ts;
   }
}

// Bm[parity][kk][l], l = 0..NPADL-1: cos(k l deg) for kk = ik, -sin(k l deg) for kk = np + ik; zero beyond the class / l > 360
constexpr int NPADL = 384;
__global__ void tg_chi_basis_kernel(int maxj, int np, double *Bm)
{
   __shared__ double ctab[360];
   tg_load_ctab(ctab);
   __syncthreads();
   const int kk = blockIdx.x, parity = blockIdx.y;
   const int M = ((maxj & 1) == parity) ? maxj : maxj - 1;
   const int ik = kk % np, part = kk / np;
   for (int l = threadIdx.x; l < NPADL; l += blockDim.x) {
      double v = 0.0;
      if (M >= 0 && ik <= M && l < NANG) {
         const int a = tg_mod360((-M + 2 * ik) * l);
         v = part == 0 ? ctab[a] : -ctab[tg_mod360(a + 270)];
      }
      Bm[((size_t)parity * 2 * np + kk) * NPADL + l] = v;
   }
}

// raw[parity][row][l] = sum_kk Tt[parity][kk][row] * Bm[parity][kk][l]      (FP64, K = 2 np, both operands K-major)
// CTA tile 128 x 128, 256 threads, 8 x 8 per thread (rows ty*4+{0..3} and 64+ty*4+{0..3}, columns 32g+2tx+{0,1}: conflict-free shared loads),
// K in slabs of TG_BK with the next slab prefetched into registers while the current one is multiplied.
constexpr int TG_BM = 128, TG_BN = 128, TG_BK = 8;
__global__ void __launch_bounds__(256, 1) tg_chi_gemm_kernel(int K, long Rpad, const double *Tt, const double *Bm, double *raw)
{
   __shared__ __align__(16) double As[2][TG_BK][TG_BM];
   __shared__ __align__(16) double Bs[2][TG_BK][TG_BN];
   const int parity = blockIdx.z;
   const double *At = Tt + (size_t)parity * K * Rpad + (size_t)blockIdx.x * TG_BM;
   const double *Bt = Bm + (size_t)parity * K * NPADL + (size_t)blockIdx.y * TG_BN;
   const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
   // loader: 8 x 128 doubles per operand per slab = 1024 doubles, 256 threads x 4 (one 32-byte row piece each)
   const int lk = tid >> 5, lc = (tid & 31) * 4;
   double acc[8][8];
#pragma unroll
   for (int i = 0; i < 8; i++)
#pragma unroll
      for (int jx = 0; jx < 8; jx++) acc[i][jx] = 0.0;
   double4 pa = *reinterpret_cast<const double4 *>(At + (size_t)lk * Rpad + lc);
   double4 pb = *reinterpret_cast<const double4 *>(Bt + (size_t)lk * NPADL + lc);
   *reinterpret_cast<double4 *>(&As[0][lk][lc]) = pa;
   *reinterpret_cast<double4 *>(&Bs[0][lk][lc]) = pb;
   __syncthreads();
   const int nslab = K / TG_BK;
   for (int sl = 0; sl < nslab; sl++) {
      const int cur = sl & 1;
      if (sl + 1 < nslab) {
         pa = *reinterpret_cast<const double4 *>(At + (size_t)((sl + 1) * TG_BK + lk) * Rpad + lc);
         pb = *reinterpret_cast<const double4 *>(Bt + (size_t)((sl + 1) * TG_BK + lk) * NPADL + lc);
      }
#pragma unroll
      for (int kk = 0; kk < TG_BK; kk++) {
         double a[8], b[8];
         const double4 a0 = *reinterpret_cast<const double4 *>(&As[cur][kk][ty * 4]);
         const double4 a1 = *reinterpret_cast<const double4 *>(&As[cur][kk][64 + ty * 4]);
         a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
         for (int g = 0; g < 4; g++) {      // columns 32 g + 2 tx + {0,1}: a quarter-warp reads 128 contiguous bytes
            const double2 bb = *reinterpret_cast<const double2 *>(&Bs[cur][kk][32 * g + tx * 2]);
            b[2 * g] = bb.x; b[2 * g + 1] = bb.y;
         }
#pragma unroll
         for (int i = 0; i < 8; i++)
#pragma unroll
            for (int jx = 0; jx < 8; jx++) acc[i][jx] = fma(a[i], b[jx], acc[i][jx]);
      }
      if (sl + 1 < nslab) {
         *reinterpret_cast<double4 *>(&As[cur ^ 1][lk][lc]) = pa;
         *reinterpret_cast<double4 *>(&Bs[cur ^ 1][lk][lc]) = pb;
      }
      __syncthreads();
   }
   double *out = raw + ((size_t)parity * Rpad + (size_t)blockIdx.x * TG_BM) * NPADL + (size_t)blockIdx.y * TG_BN;
#pragma unroll
   for (int i = 0; i < 8; i++) {
      const int r = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
#pragma unroll
      for (int g = 0; g < 4; g++)
         *reinterpret_cast<double2 *>(out + (size_t)r * NPADL + 32 * g + tx * 2) = make_double2(acc[i][2 * g], acc[i][2 * g + 1]);
   }
}

// parity combination and guards of asymrho.f:552-659, read through the symmetry-fill source map (:665-709)
__global__ void tg_asym_combine_kernel(int iodevn, long Rpad, const int *srcmap, const double *raw, double *rho, double *eng, double *esq)
{
   const int p = blockIdx.x * blockDim.x + threadIdx.x, theta = blockIdx.y;
   if (p >= NPLANE) return;
   const int src = srcmap[p], si = src / NANG, sl = src % NANG;
   double v[2][3];
#pragma unroll
   for (int par = 0; par < 2; par++)
#pragma unroll
      for (int w = 0; w < 3; w++) v[par][w] = raw[((size_t)par * Rpad + ((size_t)theta * 3 + w) * NANG + si) * NPADL + sl];
   double rhoevn = v[0][0], rotevn, esqevn, rhoodd = v[1][0], rotodd, esqodd;
   if (fabs(rhoevn) > 1.0e-16) { rotevn = v[0][1] / rhoevn; esqevn = v[0][2] / rhoevn; } else { rotevn = 0.0; esqevn = 0.0; }
   if (fabs(rhoodd) > 1.0e-16) { rotodd = v[1][1] / rhoodd; esqodd = v[1][2] / rhoodd; } else { rotodd = 0.0; esqodd = 0.0; }
   double r, e, q;
   if (iodevn == 0) { r = rhoevn; e = rotevn; q = esqevn; }
   else if (iodevn == 1) { r = rhoodd; e = rotodd; q = esqodd; }
   else {
      r = rhoevn + rhoodd;
      e = (rotodd * rhoodd + rotevn * rhoevn) / (rhoevn + rhoodd);
      q = (esqodd * rhoodd + esqevn * rhoevn) / (rhoevn + rhoodd);
      if (fabs(rhoodd + rhoevn) < 1.0e-16) { e = 0.0; q = 0.0; }
   }
   const size_t o = (size_t)theta * NPLANE + p;
   rho[o] = r; eng[o] = e; esq[o] = q;
}

// ------------------------------------------------------------------------------------------------------------
// symmetric top (symrho.f:121-168): A_w[theta][k] = sum_j (2j+1)/8pi^2 kgen d^j_{kk} e^{-tau e_jk} {1, e_jk, e_jk^2}
__global__ void tg_sym_coeff_kernel(int maxj, int kmod, int ith0, double Bz, double Bxy, double tau, double *Ak /* [theta][3][maxj+1] */)
{
   const int k = blockIdx.x * blockDim.x + threadIdx.x, theta = blockIdx.y, ith = ith0 + theta;
   if (k > maxj) return;
   double a0 = 0.0, a1 = 0.0, a2 = 0.0;
   if (k % kmod == 0) {
      const double ch = cospi((double)ith / 360.0), sh = sinpi((double)ith / 360.0), x = cospi((double)ith / 180.0);
      WigD r;
      r.start(k, k, ch, sh, x);
      const int kgen = (k == 0) ? 1 : 2;
      for (;;) {
         const int j = r.j;
         const double pre = (double)(2 * j + 1) / (8.0 * TG_PI * TG_PI);
         const double ejk = Bxy * j * (j + 1) + (Bz - Bxy) * k * k;
         const double t = pre * kgen * r.value() * exp(-tau * ejk);
         a0 += t; a1 += t * ejk; a2 += t * ejk * ejk;
         if (j == maxj) break;
         r.next();
      }
   }
   double *A = Ak + (size_t)theta * 3 * (maxj + 1);
   A[k] = a0; A[(maxj + 1) + k] = a1; A[2 * (maxj + 1) + k] = a2;
}
// one CTA per theta: f_w(icp) = sum_k A_w[k] cos(k icp deg), icp = 0..359; plane[iph][ich] = f((iph+ich) mod 360)
__global__ void tg_sym_plane_kernel(int maxj, const double *Ak, double *rho, double *eng, double *esq)
{
   __shared__ double ctab[360], f[3][360];
   tg_load_ctab(ctab);
   __syncthreads();
   const int theta = blockIdx.x;
   const double *A = Ak + (size_t)theta * 3 * (maxj + 1);
   for (int icp = threadIdx.x; icp < 360; icp += blockDim.x) {
      double r = 0.0, e = 0.0, q = 0.0;
      for (int k = 0; k <= maxj; k++) {
         const double c = ctab[tg_mod360(k * icp)];
         r += A[k] * c; e += A[(maxj + 1) + k] * c; q += A[2 * (maxj + 1) + k] * c;
      }
      f[0][icp] = r; f[1][icp] = e / r; f[2][icp] = q / r;          // no guard in the reference (symrho.f:143-144)
   }
   __syncthreads();
   for (int p = threadIdx.x; p < NPLANE; p += blockDim.x) {
      const int icp = (p / NANG + p % NANG) % 360;
      const size_t o = (size_t)theta * NPLANE + p;
      rho[o] = f[0][icp]; eng[o] = f[1][icp]; esq[o] = f[2][icp];
   }
}

// ------------------------------------------------------------------------------------------------------------
// linear rotor (linden.f:88-122, 180-193): one thread per grid point; the operation order and roundings of the Fortran
// (no FMA contraction; e^{-tau B l(l+1)} tabulated once by the host's libm exactly as exarho evaluates it) so the output
// file is byte-identical to linden.out.
__global__ void tg_linden_kernel(int npt, double cstep, int lmax, int iodevn, double bconst, const double *expo, double d1, double d2,
                                 double fourpi, const double *cost_in, double *out)
{
   const int ic = blockIdx.x * blockDim.x + threadIdx.x;
   if (ic >= npt) return;
   const double cost = cost_in ? cost_in[ic] : __dsub_rn(__dmul_rn((double)ic, cstep), 1.0);
   double pm = 1.0, pl = cost, rho = 0.0, erot = 0.0, erotsq = 0.0;       // P(0), P(1)
   for (int l = 0; l <= lmax; l++) {
      double p;
      if (l == 0) p = 1.0;
      else if (l == 1) p = cost;
      else {   // P(L+1) = ((2.0*L+1)*X*P(L) - L*P(L-1))/(L+1) with L = l-1
         const int L = l - 1;
         p = __ddiv_rn(__dsub_rn(__dmul_rn(__dmul_rn((double)(2 * L + 1), cost), pl), __dmul_rn((double)L, pm)), (double)(L + 1));
         pm = pl; pl = p;
      }
      if ((l % 2 == iodevn) || iodevn == -1) {
         double tmp = __dmul_rn((double)(2 * l + 1), p);
         tmp = __dmul_rn(tmp, expo[l]);
         rho = __dadd_rn(rho, tmp);
         const double t1 = __dmul_rn(__dmul_rn(__dmul_rn(tmp, (double)l), (double)(l + 1)), bconst);
         erot = __dadd_rn(erot, t1);
         erotsq = __dadd_rn(erotsq, __dmul_rn(__dmul_rn(__dmul_rn(t1, (double)l), (double)(l + 1)), bconst));
      }
   }
   erot = __ddiv_rn(erot, d1);
   erotsq = __ddiv_rn(erotsq, d2);
   out[4 * ic + 0] = cost;
   out[4 * ic + 1] = __ddiv_rn(rho, fourpi);
   out[4 * ic + 2] = __ddiv_rn(erot, fourpi);
   out[4 * ic + 3] = __ddiv_rn(erotsq, fourpi);
}

// host: replay of the four sequential symmetry passes (asymrho.f:665-709) on an index plane
void asym_source_map(std::vector<int> &src)
{
   src.resize(NPLANE);
   for (int p = 0; p < NPLANE; p++) src[p] = p;
   auto at = [&](int iphi, int ichi) -> int & { return src[iphi * NANG + ichi]; };
   for (int iphi = 90; iphi <= 180; iphi++)
      for (int ichi = 180 - iphi; ichi <= iphi; ichi++) at(iphi, ichi) = at(180 - ichi, 180 - iphi);
   for (int iphi = 180; iphi <= 270; iphi++)
      for (int ichi = iphi - 180; ichi <= 360 - iphi; ichi++) at(iphi, ichi) = at(180 + ichi, iphi - 180);
   for (int iphi = 180; iphi <= 360; iphi++)
      for (int ichi = 360 - iphi; ichi <= iphi; ichi++) at(iphi, ichi) = at(360 - ichi, 360 - iphi);
   for (int iphi = 0; iphi <= 360; iphi++)
      for (int ichi = iphi; ichi <= 360; ichi++) at(iphi, ichi) = at(ichi, iphi);
}

double g_last_ms[4];      // device time of the last generator call: eigen+coeff, phi stage, chi GEMM, combine

}  // namespace

extern "C" {

int pimcgpu_gen_wigner_d(int maxj, double theta, double *d)
{
   if (need_device("pimcgpu_gen_wigner_d")) return 1;
   if (maxj < 0 || maxj > 876) return tfail("pimcgpu_gen_wigner_d: maxj must be in [0,876]");
   DevMem M;
   const int w = 2 * maxj + 1;
   const size_t n = (size_t)(maxj + 1) * w * w;
   double *dd;
   TCK(M.get(&dd, n, true));
   tg_wigner_kernel<<<(w * w + 127) / 128, 128>>>(maxj, cos(0.5 * theta), sin(0.5 * theta), cos(theta), dd);
   TCK(cudaGetLastError());
   TCK(cudaMemcpy(d, dd, n * sizeof(double), cudaMemcpyDeviceToHost));
   return 0;
}

int pimcgpu_gen_asymrho(double temprt, int nslice, int iodevn, int ith0, int ith1, double Arot, double Brot, double Crot, int maxj,
                        double *rho, double *eng, double *esq, double *info)
{
   if (need_device("pimcgpu_gen_asymrho")) return 1;
   if (maxj > 876) return tfail("maxj is larger than the limit of 876");                    // asymrho.f:26
   if (maxj < 0) return tfail("pimcgpu_gen_asymrho: maxj < 0");
   if (iodevn > 1 || iodevn < -1) return tfail("iodevn can only be -1 0 1");                // :88-91
   if (ith0 < 0 || ith1 > 180 || ith1 < ith0) return tfail("pimcgpu_gen_asymrho: theta range must lie in 0..180 (weird ithe)");
   if (!(temprt > 0.0) || nslice < 1) return tfail("pimcgpu_gen_asymrho: temperature and slice count must be positive");
   const double beta = 1.0 / (TG_BOLTZ * temprt), tau = beta / (double)nslice;
   // block table
   std::vector<AsymBlock> blocks;
   std::vector<long> off_s(2 * (size_t)(maxj + 1), -1);
   long oe = 0, oz = 0, os = 0;
   for (int j = 0; j <= maxj; j++)
      for (int parity = 0; parity < 2; parity++) {
         const int kst = ((j & 1) == parity) ? -j : -j + 1;
         const int n = (j - kst) / 2 + 1;
         if (kst > j) continue;                                  // (j = 0, odd) is empty
         AsymBlock b{j, parity, n, kst, oe, oz, os};
         off_s[2 * j + parity] = os;
         blocks.push_back(b);
         oe += n; oz += (long)n * n; os += 3L * n * n;
      }
   DevMem M;
   AsymBlock *d_blocks; long *d_offs; double *d_eng, *d_Z, *d_S; int *d_status, *d_src;
   TCK(M.get(&d_blocks, blocks.size()));
   TCK(M.get(&d_offs, off_s.size()));
   TCK(M.get(&d_eng, oe));
   TCK(M.get(&d_Z, oz));
   TCK(M.get(&d_S, os));
   TCK(M.get(&d_status, 1, true));
   TCK(cudaMemcpy(d_blocks, blocks.data(), blocks.size() * sizeof(AsymBlock), cudaMemcpyHostToDevice));
   TCK(cudaMemcpy(d_offs, off_s.data(), off_s.size() * sizeof(long), cudaMemcpyHostToDevice));
   cudaEvent_t ev[5];
   for (auto &e : ev) cudaEventCreate(&e);
   cudaEventRecord(ev[0]);
   const int nmax = maxj + 1;
   tg_eigen_kernel<<<(unsigned)blocks.size(), 128, 4 * nmax * sizeof(double)>>>(d_blocks, Arot, Brot, Crot, tau, d_eng, d_Z, d_S, d_status);
   TCK(cudaGetLastError());
   int status = 0;
   std::vector<double> h_eng(oe);
   TCK(cudaMemcpy(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost));
   if (status) return tfail("  FAIL (QL iteration did not converge)");                     // asymrho.f:928
   TCK(cudaMemcpy(h_eng.data(), d_eng, oe * sizeof(double), cudaMemcpyDeviceToHost));
   // emax checks and partition sums (asymrho.f:278-425; log output of the reference, returned in info)
   double emax = -DBL_MAX;
   for (double e : h_eng) emax = std::max(emax, e);
   if (exp(-beta * emax) > 1e-8 || exp(-tau * emax) > 1e-8) return tfail("too large contribution from emax");   // :293,296
   if (info) {
      for (int pass = 0; pass < 2; pass++) {
         const double b = pass ? tau : beta;
         double z[2] = {0, 0}, es[2] = {0, 0}, eq[2] = {0, 0};
         for (const AsymBlock &bl : blocks)
            for (int i = 0; i < bl.n; i++) {
               const double en = h_eng[bl.off_e + i], wgt = (2 * bl.j + 1) * exp(-b * en);
               z[bl.parity] += wgt; es[bl.parity] += wgt * en; eq[bl.parity] += wgt * en * en;
            }
         const double kt2 = TG_BOLTZ * TG_BOLTZ * temprt * temprt;
         const double zc = z[0] + z[1], ec = (es[0] + es[1]) / zc, qc = (eq[0] + eq[1]) / zc;
         if (pass == 0) {
            for (int par = 0; par < 2; par++) {
               const double ea = es[par] / z[par], qa = eq[par] / z[par];
               info[3 * par] = z[par]; info[3 * par + 1] = ea; info[3 * par + 2] = (qa - ea * ea) / kt2;
            }
            info[6] = zc; info[7] = ec; info[8] = (qc - ec * ec) / kt2;
         } else {
            info[9] = z[0]; info[10] = es[0] / z[0]; info[11] = z[1]; info[12] = es[1] / z[1]; info[13] = zc; info[14] = ec;
         }
      }
      info[15] = emax;
   }
   std::vector<int> src;
   asym_source_map(src);
   TCK(M.get(&d_src, NPLANE));
   TCK(cudaMemcpy(d_src, src.data(), NPLANE * sizeof(int), cudaMemcpyHostToDevice));

   const int np = ((maxj + 1 + 3) / 4) * 4;          // class size padded so that K = 2 np is a multiple of TG_BK
   const int K = 2 * np;
   const int ntheta_all = ith1 - ith0 + 1;
   const int batch = std::min(ntheta_all, 32);
   const long Rpad = (((long)batch * 3 * NANG + TG_BM - 1) / TG_BM) * TG_BM;
   double *d_A, *d_Tt, *d_Bm, *d_raw, *d_rho, *d_e, *d_q;
   TCK(M.get(&d_A, (size_t)batch * 2 * 3 * np * np, true));
   TCK(M.get(&d_Tt, (size_t)2 * K * Rpad, true));
   TCK(M.get(&d_Bm, (size_t)2 * K * NPADL));
   TCK(M.get(&d_raw, (size_t)2 * Rpad * NPADL));
   TCK(M.get(&d_rho, (size_t)batch * NPLANE));
   TCK(M.get(&d_e, (size_t)batch * NPLANE));
   TCK(M.get(&d_q, (size_t)batch * NPLANE));
   tg_chi_basis_kernel<<<dim3(K, 2), 128>>>(maxj, np, d_Bm);
   TCK(cudaGetLastError());
   cudaEventRecord(ev[1]);
   float acc_ms[4] = {0, 0, 0, 0}, ms = 0;
   cudaEventSynchronize(ev[1]);
   cudaEventElapsedTime(&ms, ev[0], ev[1]);
   acc_ms[0] = ms;
   for (int t0 = 0; t0 < ntheta_all; t0 += batch) {
      const int nt = std::min(batch, ntheta_all - t0);
      const int ncmax = maxj + 1;
      cudaEventRecord(ev[0]);
      tg_asym_coeff_kernel<<<dim3((ncmax * ncmax + 127) / 128, 2, nt), 128>>>(maxj, np, ith0 + t0, d_offs, d_S, d_A);
      TCK(cudaGetLastError());
      cudaEventRecord(ev[1]);
      tg_phi_kernel<<<dim3(ncmax, nt * 3, 2), 384>>>(maxj, np, Rpad, d_A, d_Tt);
      TCK(cudaGetLastError());
      cudaEventRecord(ev[2]);
      const long rows = (((long)nt * 3 * NANG + TG_BM - 1) / TG_BM);
      tg_chi_gemm_kernel<<<dim3((unsigned)rows, NPADL / TG_BN, 2), 256>>>(K, Rpad, d_Tt, d_Bm, d_raw);
      TCK(cudaGetLastError());
      cudaEventRecord(ev[3]);
      tg_asym_combine_kernel<<<dim3((NPLANE + 255) / 256, nt), 256>>>(iodevn, Rpad, d_src, d_raw, d_rho, d_e, d_q);
      TCK(cudaGetLastError());
      cudaEventRecord(ev[4]);
      TCK(cudaMemcpy(rho + (size_t)t0 * NPLANE, d_rho, (size_t)nt * NPLANE * sizeof(double), cudaMemcpyDeviceToHost));
      TCK(cudaMemcpy(eng + (size_t)t0 * NPLANE, d_e, (size_t)nt * NPLANE * sizeof(double), cudaMemcpyDeviceToHost));
      TCK(cudaMemcpy(esq + (size_t)t0 * NPLANE, d_q, (size_t)nt * NPLANE * sizeof(double), cudaMemcpyDeviceToHost));
      for (int s = 0; s < 4; s++) { cudaEventElapsedTime(&ms, ev[s], ev[s + 1]); acc_ms[s] += ms; }
   }
   for (int s = 0; s < 4; s++) g_last_ms[s] = acc_ms[s];
   for (auto &e : ev) cudaEventDestroy(e);
   return 0;
}

int pimcgpu_gen_timing(double *ms4)
{
   for (int s = 0; s < 4; s++) ms4[s] = g_last_ms[s];
   return 0;
}

int pimcgpu_gen_symrho(double temprt, int nslice, int kmod, int ith0, int ith1, double Bz, double Bxy, int maxj, double *rho, double *eng,
                       double *esq, double *info)
{
   if (need_device("pimcgpu_gen_symrho")) return 1;
   if (maxj > 876 || maxj < 0) return tfail("maxj is larger than the limit of 876");       // symrho.f:25
   if (kmod < 1) return tfail("pimcgpu_gen_symrho: kmod must be >= 1");
   if (ith0 < 0 || ith1 > 180 || ith1 < ith0) return tfail("pimcgpu_gen_symrho: theta range must lie in 0..180 (weird ith)");
   if (!(temprt > 0.0) || nslice < 1) return tfail("pimcgpu_gen_symrho: temperature and slice count must be positive");
   const double beta = 1.0 / (TG_BOLTZ * temprt), tau = beta / (double)nslice;
   // partition sums and the truncation check (symrho.f:75-118)
   double ztau = 0, zbeta = 0, Ebeta = 0, Esqrt = 0;
   for (int j = 0; j <= maxj; j++)
      for (int k = 0; k <= j; k++)
         if (k % kmod == 0) {
            const int kgen = (k == 0) ? 1 : 2;
            const double ejk = Bxy * j * (j + 1) + (Bz - Bxy) * k * k;
            ztau += kgen * exp(-tau * ejk) * (2 * j + 1);
            zbeta += kgen * exp(-beta * ejk) * (2 * j + 1);
            Ebeta += kgen * exp(-beta * ejk) * (2 * j + 1) * ejk;
            Esqrt += kgen * exp(-beta * ejk) * (2 * j + 1) * ejk * ejk;
         }
   Ebeta = Ebeta / zbeta / TG_BOLTZ;
   Esqrt = Esqrt / zbeta / (TG_BOLTZ * TG_BOLTZ);
   if (info) { info[0] = ztau; info[1] = zbeta; info[2] = Ebeta; info[3] = Esqrt; info[4] = (Esqrt - Ebeta * Ebeta) / (temprt * temprt); }
   const double emax = (Bz > Bxy) ? Bxy * maxj * (maxj + 1) + (Bz - Bxy) * maxj * maxj : Bxy * maxj * (maxj + 1);
   const double pmax = (2 * maxj + 1) * exp(-tau * emax) / ztau;
   if (pmax > 1e-16) return tfail("pmax too large %g increase maxj", pmax);               // symrho.f:115-118
   DevMem M;
   const int nt = ith1 - ith0 + 1;
   double *d_Ak, *d_rho, *d_e, *d_q;
   TCK(M.get(&d_Ak, (size_t)nt * 3 * (maxj + 1)));
   TCK(M.get(&d_rho, (size_t)nt * NPLANE));
   TCK(M.get(&d_e, (size_t)nt * NPLANE));
   TCK(M.get(&d_q, (size_t)nt * NPLANE));
   tg_sym_coeff_kernel<<<dim3((maxj + 1 + 63) / 64, nt), 64>>>(maxj, kmod, ith0, Bz, Bxy, tau, d_Ak);
   TCK(cudaGetLastError());
   tg_sym_plane_kernel<<<nt, 512>>>(maxj, d_Ak, d_rho, d_e, d_q);
   TCK(cudaGetLastError());
   TCK(cudaMemcpy(rho, d_rho, (size_t)nt * NPLANE * sizeof(double), cudaMemcpyDeviceToHost));
   TCK(cudaMemcpy(eng, d_e, (size_t)nt * NPLANE * sizeof(double), cudaMemcpyDeviceToHost));
   TCK(cudaMemcpy(esq, d_q, (size_t)nt * NPLANE * sizeof(double), cudaMemcpyDeviceToHost));
   return 0;
}

int pimcgpu_gen_linden(double temprt, int nslice, double bconst, int npt, int iodevn, double *out, double *info)
{
   if (need_device("pimcgpu_gen_linden")) return 1;
   if (npt < 2) return tfail("pimcgpu_gen_linden: npt must be >= 2");
   if (iodevn > 1 || iodevn < -1) return tfail("pimcgpu_gen_linden: iodevn can only be -1 0 1");
   if (!(temprt > 0.0) || nslice < 1 || !(bconst > 0.0)) return tfail("pimcgpu_gen_linden: T, nslice and B must be positive");
   const int maxl = 500;                                           // linden.f:4
   const double taunit = 1.4387752224e+00, eps = 1e-16, boltz = 0.69503476e0;   // :5, :92
   const double tau = taunit / (temprt * nslice);
   int lmax = maxl;
   for (int l = 0; l <= maxl; l++)
      if (exp(-tau * bconst * l * (l + 1)) < eps) { lmax = l; break; }          // :27-36
   if (lmax < 1) return tfail("pimcgpu_gen_linden: lmax < 1");
   const double cstep = (double)2.0f / (double)(npt - 1);
   const double d1 = nslice * boltz, d2 = pow(nslice * boltz, 2.0), fourpi = (double)4.0f * TG_PI;
   std::vector<double> ex(lmax + 1), exb(lmax + 1);
   const double betat = tau * nslice;
   for (int l = 0; l <= lmax; l++) {
      ex[l] = exp(-tau * bconst * (double)(l * (l + 1)));
      exb[l] = exp(-betat * bconst * (double)(l * (l + 1)));
   }
   DevMem M;
   double *d_ex, *d_out, *d_one, *d_o1;
   TCK(M.get(&d_ex, lmax + 1));
   TCK(M.get(&d_out, (size_t)npt * 4));
   TCK(M.get(&d_one, 1));
   TCK(M.get(&d_o1, 4));
   TCK(cudaMemcpy(d_ex, ex.data(), (lmax + 1) * sizeof(double), cudaMemcpyHostToDevice));
   tg_linden_kernel<<<(npt + 127) / 128, 128>>>(npt, cstep, lmax, iodevn, bconst, d_ex, d1, d2, fourpi, nullptr, d_out);
   TCK(cudaGetLastError());
   TCK(cudaMemcpy(out, d_out, (size_t)npt * 4 * sizeof(double), cudaMemcpyDeviceToHost));
   if (info) {   // "Erot at Beta", "Cv at Beta" (linden.f:72-82): the same sum at cost = 1 with tau -> beta
      const double one = 1.0;
      double o[4];
      TCK(cudaMemcpy(d_one, &one, sizeof one, cudaMemcpyHostToDevice));
      TCK(cudaMemcpy(d_ex, exb.data(), (lmax + 1) * sizeof(double), cudaMemcpyHostToDevice));
      tg_linden_kernel<<<1, 32>>>(1, cstep, lmax, iodevn, bconst, d_ex, d1, d2, fourpi, d_one, d_o1);
      TCK(cudaGetLastError());
      TCK(cudaMemcpy(o, d_o1, sizeof o, cudaMemcpyDeviceToHost));
      const double erot = o[2] * nslice / o[1], erotsq = o[3] * nslice * nslice / o[1];
      info[0] = tau; info[1] = lmax; info[2] = erot; info[3] = (erotsq - erot * erot) / pow(temprt, 2.0);
   }
   return 0;
}

/* Fortran edit descriptors of the generators' writers: E15.8 (asymrho.f:721-723) and 1P,E15.8 (linden.f:68); buf[16] */
void pimcgpu_format_e15_8(double v, int scale1p, char *buf)
{
   char t[48];
   if (scale1p) { snprintf(t, sizeof t, "%15.8E", v); memcpy(buf, t, 15); buf[15] = 0; return; }
   snprintf(t, sizeof t, "%.7E", v);                 // [-]d.dddddddE[+-]xx
   const char *s = t;
   const bool neg = (*s == '-');
   if (neg) ++s;
   int ex = atoi(strchr(s, 'E') + 1);
   if (v != 0.0) ex += 1;
   char body[48];
   int n = 0;
   if (neg) body[n++] = '-';
   body[n++] = '0'; body[n++] = '.'; body[n++] = s[0];
   for (int i = 2; i <= 8; i++) body[n++] = s[i];
   const int ax = ex < 0 ? -ex : ex;
   if (ax < 100) { body[n++] = 'E'; body[n++] = ex < 0 ? '-' : '+'; body[n++] = '0' + ax / 10; body[n++] = '0' + ax % 10; }
   else { body[n++] = ex < 0 ? '-' : '+'; body[n++] = '0' + ax / 100; body[n++] = '0' + (ax / 10) % 10; body[n++] = '0' + ax % 10; }
   const int pad = 15 - n;
   for (int i = 0; i < pad; i++) buf[i] = ' ';
   memcpy(buf + (pad > 0 ? pad : 0), body, n);
   buf[15] = 0;
}

/* one E15.8 value per line (the rho.denXXX_rho/_eng/_esq and <type>_T<T>t<Q>.rho/.eng/.esq format read by init_rot3D,
 * mc_poten.cc:462-499); append != 0 continues an existing file (compile.x concatenates the planes) */
int pimcgpu_write_e15_8(const char *path, const double *v, long n, int append)
{
   FILE *f = fopen(path, append ? "ab" : "wb");
   if (!f) return tfail("pimcgpu_write_e15_8: cannot open %s", path);
   const long chunk = 1 << 20;
   std::vector<char> buf((size_t)std::min(n, chunk) * 16 + 16);
   for (long o = 0; o < n; o += chunk) {
      const long m = std::min(chunk, n - o);
#pragma omp parallel for schedule(static)
      for (long i = 0; i < m; i++) {
         char t[16];
         pimcgpu_format_e15_8(v[o + i], 0, t);
         memcpy(&buf[(size_t)i * 16], t, 15);
         buf[(size_t)i * 16 + 15] = '\n';
      }
      if (fwrite(buf.data(), 16, (size_t)m, f) != (size_t)m) { fclose(f); return tfail("pimcgpu_write_e15_8: short write to %s", path); }
   }
   fclose(f);
   return 0;
}

/* linden.out / <type>_T<T>t<Q>.rot: '(1p,7(1x,E15.8))' rows of cost, rho, erot, erotsq (linden.f:68; read by init_rotdens) */
int pimcgpu_write_rot(const char *path, const double *out4, int npt)
{
   FILE *f = fopen(path, "wb");
   if (!f) return tfail("pimcgpu_write_rot: cannot open %s", path);
   for (int i = 0; i < npt; i++) {
      char line[80];
      int n = 0;
      for (int c = 0; c < 4; c++) {
         line[n++] = ' ';
         pimcgpu_format_e15_8(out4[4 * i + c], 1, line + n);
         n += 15;
      }
      line[n++] = '\n';
      fwrite(line, 1, n, f);
   }
   fclose(f);
   return 0;
}

}  // extern "C"
