// C ABI of the B200 PIMC hot path (include/pimcgpu.h): context set-up, table preparation,
// state transfer, kernel launches.  Host-side numerics here (spline second derivatives, the
// short/long-range extrapolation constants, MRG32k3a stream jumps) are this library's own
// implementation of the published algorithms the reference uses at table-load time
// (mc_utils.cc:112-201, mc_poten.cc:416-437, rngstream.cc:303-321).
#include "../../include/pimcgpu.h"
#include "pimc_device.cuh"
#include "pimc_moves.cuh"
#include "pimc_estim.cuh"

#include <cmath>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

using namespace pimc;

namespace {

std::string g_err;
int fail(const char *fmt, ...)
{
   char buf[512];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof buf, fmt, ap);
   va_end(ap);
   g_err = buf;
   return 1;
}
}  // namespace
namespace pimc { int set_error(const char *msg) { g_err = msg; return 1; } }   // for the other translation units (pimc_tablegen.cu)
namespace {
#define CK(call)                                                                            \
   do {                                                                                     \
      cudaError_t e_ = (call);                                                              \
      if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
   } while (0)

struct Context {
   bool live = false;
   pimcgpu_system sys;
   Params p;
   EstBuffers e;
   cudaStream_t stream = nullptr;
   std::vector<void *> allocs;
   int threads = 0;
   size_t smem = 0;
   long step = 0;
   long nacc = 0;
   int *d_err = nullptr;
   int *d_ops = nullptr;            // [c][4] explicit symmetry operations
   bool seeded = false;
   int kind = 0;                    // rotor kind the move kernel is specialised on
   std::vector<int> h_pindex;       // [c][N]
   double *stage = nullptr;         // pinned host staging: [c][pos | ang | cosn] in the device layout
   size_t stage_chain = 0;          // doubles per chain in `stage`
   double *d_raw = nullptr;         // device scratch [3][N*P]: one chain's beads in the reference layout (import/export transposes)
   double *d_raw_all = nullptr;     // the same for every chain (batched pimcgpu_upload_states / _download_states), allocated on first use
   // split-phase transfers (pimcgpu_upload_states_begin/_commit, pimcgpu_download_states_begin/_end): a copy stream, separate
   // staging buffers in each direction, pinned staging of the small per-chain arrays of an upload in flight
   cudaStream_t copy_stream = nullptr;
   cudaEvent_t ev_up = nullptr, ev_down = nullptr, ev_done = nullptr, ev_commit = nullptr, ev_acc = nullptr;
   int up_slot = 0;                 // which half of the pinned small-array staging the upload in flight uses (two uploads alternate)
   double *d_raw_up = nullptr;
   double *d_small_up = nullptr, *d_small_down = nullptr;      // device staging of the rotor rows (split-phase upload / download)
   int *d_perm_up = nullptr;                                   // ... and of the permutation tables
   double *stage_up = nullptr;      // pinned: [c][ang | cosn] of the upload in flight
   int *stage_perm = nullptr;       // pinned: gp, gr, cst, cat, ncy of the upload in flight
   int up_first = 0, up_count = 0, down_first = 0, down_count = 0;
   double *down_angles = nullptr, *down_cosine = nullptr;
} G;

template <class T> int dalloc(T **ptr, size_t n)
{
   void *q = nullptr;
   cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
   if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
   cudaMemset(q, 0, std::max<size_t>(n, 1) * sizeof(T));
   G.allocs.push_back(q);
   *ptr = (T *)q;
   return 0;
}
template <class T> int dupload(const T **dst, const T *src, size_t n)
{
   T *q = nullptr;
   if (dalloc(&q, n)) return 1;
   cudaError_t e = cudaMemcpy(q, src, n * sizeof(T), cudaMemcpyHostToDevice);
   if (e != cudaSuccess) return fail("cudaMemcpy H2D failed: %s", cudaGetErrorString(e));
   *dst = q;
   return 0;
}

// natural-cubic-spline second derivatives with clamped end slopes taken from the first/last
// interval (Numerical-Recipes tridiagonal sweep; same choice as init_spline, mc_utils.cc:188-201)
void spline_setup(const std::vector<double> &x, const std::vector<double> &y, std::vector<double> &y2)
{
   int n = (int)x.size();
   y2.assign(n, 0.0);
   std::vector<double> u(n, 0.0);
   double yp1 = (y[1] - y[0]) / (x[1] - x[0]);
   double ypn = (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]);
   y2[0] = -0.5;
   u[0] = (3. / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - yp1);
   for (int i = 1; i < n - 1; i++) {
      double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
      double pp = sig * y2[i - 1] + 2.;
      y2[i] = (sig - 1.) / pp;
      double ui = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
      u[i] = (6. * ui / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / pp;
   }
   double qn = .5;
   double un = (3. / (x[n - 1] - x[n - 2])) * (ypn - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]));
   y2[n - 1] = (un - qn * u[n - 2]) / (qn * y2[n - 2] + 1.);
   for (int k = n - 2; k >= 0; k--) y2[k] = y2[k] * y2[k + 1] + u[k];
}
// bucket table for the interval search: lut[b] = max{k : x[k] <= x0 + b/scale} (lower bound of the bucket)
void build_lut(const std::vector<double> &x, std::vector<int> &lut, double &scale, int factor = 4)
{
   int n = (int)x.size();
   int nl = factor * n;
   lut.assign(nl, 0);
   scale = (double)nl / (x[n - 1] - x[0]);
   int k = 0;
   for (int b = 0; b < nl; b++) {
      double xb = x[0] + (double)b / scale;
      while (k < n - 2 && x[k + 1] <= xb) k++;
      lut[b] = k;
   }
}

// ---- MRG32k3a stream jumps in exact integer arithmetic -------------------------------------------
typedef unsigned long long u64;
const u64 M1 = 4294967087ull, M2 = 4294944443ull;
// A^(2^127) of the two components (L'Ecuyer, Simard, Chen, Kelton 2002; RngStreams package constants)
const u64 A1P127[3][3] = {{2427906178ull, 3580155704ull, 949770784ull}, {226153695ull, 1230515664ull, 3580155704ull}, {1988835001ull, 986791581ull, 1230515664ull}};
const u64 A2P127[3][3] = {{1464411153ull, 277697599ull, 1610723613ull}, {32183930ull, 1464411153ull, 1022607788ull}, {2824425944ull, 32183930ull, 2093834863ull}};
inline u64 mulmod(u64 a, u64 b, u64 m) { return (u64)(((unsigned __int128)a * b) % m); }
void matvec(const u64 A[3][3], const u64 s[3], u64 v[3], u64 m)
{
   u64 x[3];
   for (int i = 0; i < 3; i++) x[i] = (mulmod(A[i][0], s[0], m) + mulmod(A[i][1], s[1], m) + mulmod(A[i][2], s[2], m)) % m;
   for (int i = 0; i < 3; i++) v[i] = x[i];
}
void matmat(const u64 A[3][3], const u64 B[3][3], u64 C[3][3], u64 m)
{
   u64 W[3][3];
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) W[i][j] = (mulmod(A[i][0], B[0][j], m) + mulmod(A[i][1], B[1][j], m) + mulmod(A[i][2], B[2][j], m)) % m;
   memcpy(C, W, sizeof W);
}
// state of the s-th stream declared after SetPackageSeed(seed): (A^(2^127))^s applied to the seed
void stream_state(const u64 seed[6], u64 s, u64 st[6])
{
   u64 B1[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, B2[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, W1[3][3], W2[3][3];
   memcpy(W1, A1P127, sizeof W1);
   memcpy(W2, A2P127, sizeof W2);
   while (s > 0) {
      if (s & 1) { matmat(W1, B1, B1, M1); matmat(W2, B2, B2, M2); }
      matmat(W1, W1, W1, M1); matmat(W2, W2, W2, M2);
      s >>= 1;
   }
   matvec(B1, seed, st, M1);
   matvec(B2, seed + 3, st + 3, M2);
}

// packed per-interval records: 1/h and y'' h^2/6 folded in (the device evaluates a*ylo + b*yhi + (a^3-a)clo + (b^3-b)chi)
std::vector<SplineRec> make_records(const std::vector<double> &x, const std::vector<double> &y, const std::vector<double> &y2)
{
   int n = (int)x.size();
   std::vector<SplineRec> r(n - 1);
   for (int k = 0; k < n - 1; k++) {
      double h = x[k + 1] - x[k];
      r[k].xlo = x[k]; r[k].xhi = x[k + 1]; r[k].inv_h = 1.0 / h;
      r[k].ylo = y[k]; r[k].yhi = y[k + 1];
      r[k].clo = y2[k] * (h * h) / 6.; r[k].chi = y2[k + 1] * (h * h) / 6.; r[k].pad = 0.0;
   }
   return r;
}

int pow2floor(int v) { int r = 1; while (2 * r <= v) r *= 2; return r; }
int pow2ceil(int v) { int r = 1; while (r < v) r *= 2; return r; }

size_t smem_bytes(const Params &p, int threads)
{
   size_t d = 40;
   auto pad = [](int n) { return (size_t)((n + 1) & ~1); };
   auto padi = [](int n) { return (size_t)(((n + 1) / 2 + 1) & ~1); };
   if (p.n1d && p.poly1d) d += 4 * (size_t)(p.n1d - 1);
   else if (p.n1d) d += pad((p.n1d - 1) * (int)(sizeof(SplineRec) / sizeof(double))) + (p.uniform1d ? 0 : padi(p.nlut1d));
   if (p.rs2d) d += 2 * (size_t)(p.rs2d + p.cs2d);
   if (p.nrot && p.rot_in_smem) d += pad((p.nrot - 1) * (int)(sizeof(SplineRec) / sizeof(double))) + padi(p.nlutrot);
   if (!p.segbuf_global) d += (size_t)(threads / p.team) * p.team_buf_n;
   d += 2 * (size_t)(threads / 32);
   if (p.rot_fused) d += 3 * (size_t)((p.Q + p.cpc - 1) / p.cpc);
   size_t bytes = d * sizeof(double);
   if (p.rot_fused) bytes += (size_t)((p.Q + p.cpc - 1) / p.cpc) * sizeof(RotSlot) + (p.rot_run_cta ? (size_t)p.Q * sizeof(int) : 0);
   else if (p.rot_group > 1) bytes += (size_t)(threads / p.rot_group) * sizeof(RotSlot);
   if (p.worm_on) bytes += 16 + worm_scratch_bytes(p.N);
   return bytes;
}

void est_shapes(dim3 &g_rcf, dim3 &b_rcf)
{
   b_rcf = dim3(128);
   g_rcf = dim3((G.p.Q + 127) / 128, G.p.nchains);
}

// the step kernel variant: rotor kind in bits 0-1, worm in bit 2, free-running sweeps of a one-CTA top in bit 3
const void *steps_kernel(int kind, int worm, int run_cta = 0)
{
   if (run_cta && kind == 2 && !worm) return (const void *)pimc_steps_kernel<10>;      // bit 3: free-running sweeps of a one-CTA top
   switch (kind + 4 * (worm ? 1 : 0)) {
      case 0: return (const void *)pimc_steps_kernel<0>;
      case 1: return (const void *)pimc_steps_kernel<1>;
      case 2: return (const void *)pimc_steps_kernel<2>;
      case 4: return (const void *)pimc_steps_kernel<4>;
      case 5: return (const void *)pimc_steps_kernel<5>;
      default: return (const void *)pimc_steps_kernel<6>;
   }
}

int launch_estimators(int with_dens, int accumulate)
{
   est_energy_kernel<<<G.p.nchains * EST_BLOCKS, EST_THREADS, 64 * sizeof(double), G.stream>>>(G.p, G.e, with_dens);
   if (G.p.Q > 0 && G.p.imtype >= 0) {
      dim3 g, b;
      est_shapes(g, b);
      est_rcf_kernel<<<g, b, 0, G.stream>>>(G.p, G.e);
   }
   if (G.p.bstype >= 0) {
      const long n = (long)G.p.nchains * G.p.P * 3;
      est_com_kernel<<<(unsigned)((n + 255) / 256), 256, 0, G.stream>>>(G.p, G.e);
      est_area_kernel<<<G.p.nchains * EST_BLOCKS, EST_THREADS, 0, G.stream>>>(G.p, G.e);
   }
   est_chain_totals_kernel<<<(G.p.nchains * (5 + NAREA) + 127) / 128, 128, 0, G.stream>>>(G.p, G.e);
   est_finalize_kernel<<<8, 128, 0, G.stream>>>(G.p, G.e, accumulate);
   CK(cudaGetLastError());
   return 0;
}

} // namespace

extern "C" {

const char *pimcgpu_last_error(void) { return g_err.c_str(); }

void pimcgpu_finalize(void)
{
   if (!G.live) return;
   cudaStreamSynchronize(G.stream);
   for (void *q : G.allocs) cudaFree(q);
   G.allocs.clear();
   G.d_raw = nullptr; G.d_raw_all = nullptr;          // lazily allocated scratch belongs to the context that is going away
   G.d_raw_up = nullptr; G.d_small_up = nullptr; G.d_small_down = nullptr; G.d_perm_up = nullptr;
   if (G.copy_stream) { cudaStreamSynchronize(G.copy_stream); cudaStreamDestroy(G.copy_stream); G.copy_stream = nullptr; }
   if (G.ev_acc) { cudaEventDestroy(G.ev_acc); G.ev_acc = nullptr; }
   if (G.ev_up) { cudaEventDestroy(G.ev_up); cudaEventDestroy(G.ev_down); cudaEventDestroy(G.ev_done); cudaEventDestroy(G.ev_commit); G.ev_up = G.ev_down = G.ev_done = G.ev_commit = nullptr; }
   if (G.stage_up) { cudaFreeHost(G.stage_up); G.stage_up = nullptr; }
   if (G.stage_perm) { cudaFreeHost(G.stage_perm); G.stage_perm = nullptr; }
   G.up_count = G.down_count = 0;
   if (G.stream) cudaStreamDestroy(G.stream);
   G.stream = nullptr;
   if (G.stage) cudaFreeHost(G.stage);
   G.stage = nullptr;
   G.live = false;
   G.seeded = false;
   G.step = 0;
}

int pimcgpu_init(const pimcgpu_system *sys, const pimcgpu_tables *tab)
{
   if (G.live) pimcgpu_finalize();
   if (!sys || !tab) return fail("pimcgpu_init: null argument");
   if (sys->ntypes < 1 || sys->ntypes > MAXT) return fail("pimcgpu_init: ntypes must be 1 or 2 (one atom type, one molecule type)");
   if (sys->rotden_type != 0 && sys->rotden_type != 1) return fail("pimcgpu_init: RotDenType must be 0 (tables) or 1 (rattle-and-shake propagator)");
   if (sys->nchains < 1) return fail("pimcgpu_init: nchains must be >= 1");
   int ndev = 0;
   cudaError_t ce = cudaGetDeviceCount(&ndev);
   if (ce != cudaSuccess || ndev == 0) return fail("pimcgpu_init: no CUDA device available (%s)", cudaGetErrorString(ce));
   CK(cudaSetDevice(sys->device));
   G.sys = *sys;
   Params &p = G.p;
   memset(&p, 0, sizeof p);
   p.ntypes = sys->ntypes; p.P = sys->P; p.Q = sys->Q;
   p.N = 0; p.imtype = -1; p.bstype = -1;
   int nmolt = 0, natomt = 0;
   for (int t = 0; t < sys->ntypes; t++) {
      const pimcgpu_type &T = sys->type[t];
      if (T.numb < 1) return fail("pimcgpu_init: type %d has no particles", t);
      p.first[t] = p.N; p.N += T.numb;
      p.numb[t] = T.numb; p.molecule[t] = T.molecule; p.stat[t] = T.stat; p.levels[t] = T.levels;
      p.mcstep[t] = T.mcstep; p.rtstep[t] = T.rtstep; p.mass[t] = T.mass;
      // lambda = hbar^2/2m in K A^2: 100 hbar^2/(amu k_B) with the CODATA-86 mantissas of mc_const.h:12-14
      p.lambda[t] = 0.5 * (100.0 * (1.05457266 * 1.05457266) / (1.6605402 * 1.380658)) / T.mass;
      if (T.stat == 1) p.bstype = t;
      if (T.molecule) { p.imtype = t; nmolt++; } else natomt++;
      if (T.levels < 1 || T.levels > MAXLEV) return fail("pimcgpu_init: bisection levels must be in [1,%d]", MAXLEV);
      if ((1 << T.levels) >= sys->P) return fail("pimcgpu_init: segment size 2^%d is not smaller than the number of slices %d", T.levels, sys->P);
      if (T.molecule == 1 && T.numb > 1) return fail("pimcgpu_init: no more than one linear dopant molecule");
   }
   if (nmolt > 1 || natomt > 1) return fail("pimcgpu_init: no more than one atom type and one molecule type");
   if (sys->ntypes == 2 && sys->type[0].molecule != 0) return fail("pimcgpu_init: molecules must follow atoms");
   p.first[sys->ntypes] = p.N;
   if (sys->ntypes == 1) p.first[2] = p.N;
   p.Npad = (p.N + 3) & ~3;
   p.NM = p.imtype >= 0 ? p.numb[p.imtype] : 0;
   p.NMpad = std::max(1, p.NM);
   p.temperature = sys->temperature;
   p.beta = 1.0 / sys->temperature;
   p.tau = p.beta / (double)p.P;
   p.R = 1; p.rottau = 0.0;
   if (p.Q > 0) {
      if (p.imtype < 0) return fail("pimcgpu_init: ROTATION without a molecule type");
      if (p.P % p.Q) return fail("pimcgpu_init: NumbTimes is not proportional to NumbRotTimes");
      p.R = p.P / p.Q; p.rottau = p.beta / (double)p.Q;
   }
   p.ispher = sys->ispher; p.minimage = sys->minimage;
   p.rotden_type = sys->rotden_type; p.rnratio = std::max(1, sys->rnratio);
   p.xrot = sys->x_rot; p.yrot = sys->y_rot; p.zrot = sys->z_rot;
   for (int d = 0; d < 3; d++) p.refl[d] = sys->reflect[d] ? 1 : 0;
   p.rotsym = sys->rotsym ? 1 : 0; p.nfold = std::max(1, sys->nfold_rot);
   if ((p.refl[0] | p.refl[1] | p.refl[2]) && !(p.imtype >= 0 && p.molecule[p.imtype] == 2 && p.Q > 0))
      return fail("pimcgpu_init: REFLECTX/Y/Z need a NONLINEAR rotor with ROTATION");
   if (p.rotsym && !(p.imtype >= 0 && p.Q > 0)) return fail("pimcgpu_init: ROTSYM needs a rotor with ROTATION");
   if (sys->worm) {
      if (sys->worm_type < 0 || sys->worm_type >= sys->ntypes) return fail("pimcgpu_init: Can't find a particle type for the worm algorithm");
      if (sys->worm_m < 1 || sys->worm_m >= sys->P) return fail("pimcgpu_init: Worm algorithm: m should be smaller then M");
      if (sys->worm_m > WORM_MAXM) return fail("pimcgpu_init: Worm.m above %d is not supported on the device", WORM_MAXM);
      p.worm_on = 1; p.worm_type = sys->worm_type; p.worm_m = sys->worm_m;
      const double density = (double)p.N / (sys->box[0] * sys->box[1] * sys->box[2]);
      p.worm_norm = sys->worm_c * density;                     // Worm.c * numb * NumbTimes * Worm.m after MCWormInit's rescaling
      p.worm_twave2 = 4.0 * p.lambda[p.worm_type] * p.tau;     // twave2, mc_setup.cc:394
      p.worm_cutoff2 = 100.0 * 100.0 * ((double)p.worm_m * p.worm_twave2);
      {  // skip-ahead matrices of the worm's gaussian batches: A^(2k), k = 0 .. 3 (WORM_MAXM + 1), A = one step of MRG32k3a
         const int nj = 3 * (WORM_MAXM + 1) + 1;
         u64 A1[3][3] = {{0, 1, 0}, {0, 0, 1}, {M1 - 810728ull, 1403580ull, 0}}, A2[3][3] = {{0, 1, 0}, {0, 0, 1}, {M2 - 1370589ull, 0, 527612ull}};
         u64 S1[3][3], S2[3][3], W1[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, W2[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
         matmat(A1, A1, S1, M1); matmat(A2, A2, S2, M2);
         std::vector<uint32_t> jt((size_t)nj * 18);
         for (int k = 0; k < nj; k++) {
            for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) { jt[(size_t)k * 18 + r * 3 + cc] = (uint32_t)W1[r][cc]; jt[(size_t)k * 18 + 9 + r * 3 + cc] = (uint32_t)W2[r][cc]; }
            matmat(S1, W1, W1, M1); matmat(S2, W2, W2, M2);
         }
         if (dupload(&p.worm_jump, jt.data(), jt.size())) return 1;
      }
   }
   if (p.rotden_type == 1) {
      if (p.Q <= 0) return fail("pimcgpu_init: ROTDENSI 1 without ROTATION");
      const bool lin = p.molecule[p.imtype] == 1;
      if (!(p.xrot > 0.0) || (!lin && (!(p.yrot > 0.0) || !(p.zrot > 0.0)))) return fail("pimcgpu_init: ROTDENSI 1 needs positive rotational constants");
      if (p.Q % p.rnratio) return fail("pimcgpu_init: NumbRotTimes is not a multiple of RNratio");
      if (!lin && p.rnratio > 1 && !tab->rho3d) return fail("pimcgpu_init: RNratio > 1 needs the density-matrix tables of the coarse slices");
   }
   for (int d = 0; d < 3; d++) p.box[d] = sys->box[d];
   if (p.ispher && p.Q > 0) return fail("pimcgpu_init: ISPHER = 1 is not compatible with ROTATION");
   // PotRotEnergy aborts on "MIN IMAGE for orient pot" (mc_piqmc.cc:2012): the rotational moves of a linear rotor have no
   // minimum-image form, so the combination is refused here instead of sampling two different Hamiltonians
   if (p.minimage && p.Q > 0 && p.imtype >= 0 && p.molecule[p.imtype] == 1 && p.N > 1)
      return fail("pimcgpu_init: MINIMAGE with ROTATION of a linear rotor is not implemented (MIN IMAGE for orient pot, mc_piqmc.cc:2012)");
   // interaction branch per (type0, type1), the if-chain of mc_piqmc.cc:1847-1958
   for (int t0 = 0; t0 < sys->ntypes; t0++)
      for (int t1 = 0; t1 < sys->ntypes; t1++) {
         int m0 = p.molecule[t0], m1 = p.molecule[t1], mode;
         if (m0 == 1 || m1 == 1) mode = (m0 == 1) ? M_LIN_0MOL : M_LIN_1MOL;
         else if ((m0 == 2 || m1 == 2) && m0 != m1) mode = p.ispher ? M_SPHER : (m0 == 2 ? M_TOP_0MOL : M_TOP_1MOL);
         else if (m0 == 2 && m1 == 2 && p.numb[p.imtype] > 1) mode = M_TOPTOP;
         else mode = M_SPOT1D;
         p.mode[t0][t1] = mode;
      }
   bool need1d = false, need2d = false, need3d = false, needsph = false;
   for (int t0 = 0; t0 < sys->ntypes; t0++)
      for (int t1 = 0; t1 < sys->ntypes; t1++) {
         if (t0 == t1 && p.numb[t0] < 2) continue;
         int m = p.mode[t0][t1];
         need1d |= m == M_SPOT1D; need2d |= (m == M_LIN_0MOL || m == M_LIN_1MOL);
         need3d |= (m == M_TOP_0MOL || m == M_TOP_1MOL); needsph |= m == M_SPHER;
      }
   // ---- tables ----
   if (need1d) {
      if (!tab->grid1d || !tab->pot1d || tab->n1d < 2) return fail("pimcgpu_init: the 1-D pair potential table is required");
      int n = tab->n1d;
      std::vector<double> g(tab->grid1d, tab->grid1d + n), v(tab->pot1d, tab->pot1d + n), y2;
      spline_setup(g, v, y2);
      // short range U0 exp(-alpha r) and long range -C6/r^6 fitted to the end intervals (mc_poten.cc:422-437)
      p.alpha = log(v[0] / v[1]) / (g[1] - g[0]);
      p.unode = v[0] * exp(p.alpha * g[0]);
      p.c6 = (v[n - 1] - v[n - 2]) / (1.0 / pow(g[n - 2], 6.0) - 1.0 / pow(g[n - 1], 6.0));
      // uniform grid: the interval index is the quotient (r - x0)/h, checked against the record; otherwise a bucket table
      // fine enough that a bucket rarely holds a grid point
      const double havg = (g[n - 1] - g[0]) / (double)(n - 1);
      double dev = 0.0;
      for (int i = 0; i < n; i++) dev = std::max(dev, fabs(g[i] - (g[0] + i * havg)));
      p.uniform1d = (dev <= 0.02 * havg) ? 1 : 0;             // also grids printed with a few digits (helium.pot)
      p.x0_1d = g[0]; p.xn_1d = g[n - 1]; p.invh_1d = 1.0 / havg;
      std::vector<int> lut;
      build_lut(g, lut, p.lut1d_scale, p.uniform1d ? 4 : 16);
      p.n1d = n; p.nlut1d = (int)lut.size();
      std::vector<SplineRec> rec = make_records(g, v, y2);
      // grid uniform to rounding: per-interval cubic in t = (r - x_k)/h for the batched pair sums of the move kernel
      // (systems with a non-linear top evaluate their few atom-atom terms one by one from the packed records)
      const bool top_system = p.imtype >= 0 && p.molecule[p.imtype] == 2;
      p.poly1d = (dev <= 1e-12 * havg * n && !top_system && !getenv("PIMC_NO_POLY1D")) ? 1 : 0;
      if (p.poly1d) {
         std::vector<double2> pa(n - 1), pb(n - 1);
         for (int k = 0; k < n - 1; k++) {
            const long double h = (long double)g[k + 1] - (long double)g[k];
            const long double c0 = (long double)y2[k] * h * h / 6.0L, c1 = (long double)y2[k + 1] * h * h / 6.0L;
            pa[k] = make_double2(v[k], (double)((long double)v[k + 1] - (long double)v[k] - 2.0L * c0 - c1));
            pb[k] = make_double2((double)(3.0L * c0), (double)(c1 - c0));
         }
         if (dupload(&p.pa1d, pa.data(), pa.size()) || dupload(&p.pb1d, pb.data(), pb.size())) return 1;
      }
      if (dupload(&p.g1d, g.data(), n) || dupload(&p.v1d, v.data(), n) || dupload(&p.y2_1d, y2.data(), n) || dupload(&p.lut1d, lut.data(), lut.size()) ||
          dupload(&p.rec1d, rec.data(), rec.size())) return 1;
   }
   if (need2d) {
      if (!tab->pot2d || !tab->rgrid2d || !tab->cgrid2d) return fail("pimcgpu_init: the 2-D atom-rotor potential table is required");
      p.rs2d = tab->rsize2d; p.cs2d = tab->csize2d; p.dr2d = tab->dr2d; p.dc2d = tab->dc2d;
      p.inv_dr2d = 1.0 / p.dr2d; p.inv_dc2d = 1.0 / p.dc2d;
      std::vector<double> ir(p.rs2d, 0.0), ic(p.cs2d, 0.0);
      for (int i = 0; i + 1 < p.rs2d; i++) ir[i] = 1.0 / (tab->rgrid2d[i + 1] - tab->rgrid2d[i]);
      for (int i = 0; i + 1 < p.cs2d; i++) ic[i] = 1.0 / (tab->cgrid2d[i + 1] - tab->cgrid2d[i]);
      if (dupload(&p.rg2d, tab->rgrid2d, p.rs2d) || dupload(&p.cg2d, tab->cgrid2d, p.cs2d) || dupload(&p.v2d, tab->pot2d, (size_t)p.rs2d * p.cs2d) ||
          dupload(&p.irg2d, ir.data(), ir.size()) || dupload(&p.icg2d, ic.data(), ic.size())) return 1;
      // row-pair copy: {V[ir][ic], V[ir+1][ic]} so the four corners of a bilinear cell are 32 contiguous bytes
      {
         const int rs = p.rs2d, cs = p.cs2d;
         std::vector<double2> cell((size_t)(rs - 1) * cs);
         for (int i = 0; i < rs - 1; i++)
            for (int j = 0; j < cs; j++) cell[(size_t)i * cs + j] = make_double2(tab->pot2d[(size_t)i * cs + j], tab->pot2d[(size_t)(i + 1) * cs + j]);
         std::vector<double2> rgi(rs), cgi(cs);
         for (int i = 0; i < rs; i++) rgi[i] = make_double2(tab->rgrid2d[i], ir[i]);
         for (int i = 0; i < cs; i++) cgi[i] = make_double2(tab->cgrid2d[i], ic[i]);
         // whole-cell copy for the cached rotational sums: the four corners of cell (ir, ic) as ONE aligned 32-byte record
         // {V[ir][ic], V[ir+1][ic], V[ir][ic+1], V[ir+1][ic+1]} -- a single 256-bit gather per evaluation
         p.cell4_on = getenv("PIMC_CELL4") ? atoi(getenv("PIMC_CELL4")) : 1;          // measured on C5: 15.6 us per rotational sweep against 16.8 us with the row-pair table
         p.cell_hint = getenv("PIMC_CELL_HINT") ? atoi(getenv("PIMC_CELL_HINT")) : 1;      // gathers bypass L1 allocation: 12.8 us against 15.2 us per sweep
         if (p.cell4_on) {
            std::vector<double> c4((size_t)(rs - 1) * (cs - 1) * 4);
            for (int i = 0; i < rs - 1; i++)
               for (int j = 0; j < cs - 1; j++) {
                  double *q = &c4[((size_t)i * (cs - 1) + j) * 4];
                  q[0] = tab->pot2d[(size_t)i * cs + j]; q[1] = tab->pot2d[(size_t)(i + 1) * cs + j];
                  q[2] = tab->pot2d[(size_t)i * cs + j + 1]; q[3] = tab->pot2d[(size_t)(i + 1) * cs + j + 1];
               }
            if (dupload(&p.cell4, c4.data(), c4.size())) return 1;
         }
         if (dupload(&p.cell2d, cell.data(), cell.size()) || dupload(&p.rgi2d, rgi.data(), rgi.size()) || dupload(&p.cgi2d, cgi.data(), cgi.size())) return 1;
      }
   }
   if (need3d) {
      if (!tab->vtable) return fail("pimcgpu_init: the 3-D atom-top potential table is required");
      p.rg3 = tab->rgrd; p.thg3 = tab->thgrd; p.chg3 = tab->chgrd; p.rvmin = tab->rvmin; p.rvmax = tab->rvmax;
      p.rvstep = (p.rvmax - p.rvmin) / (double)(p.rg3 - 1);
      if (dupload(&p.v3d, tab->vtable, (size_t)p.rg3 * p.thg3 * p.chg3)) return 1;
   }
   if (needsph) {
      if (!tab->vspher) return fail("pimcgpu_init: ISPHER=1 needs the 501-entry spherical table");
      if (dupload(&p.vspher, tab->vspher, 501)) return 1;
   }
   if (p.Q > 0 && p.molecule[p.imtype] == 1 && p.rotden_type == 0) {
      if (!tab->rotgrid || tab->nrot < 2) return fail("pimcgpu_init: the linear-rotor density table (.rot) is required");
      int n = tab->nrot;
      std::vector<double> g(tab->rotgrid, tab->rotgrid + n), y2;
      const double *cols[3] = {tab->rotdens, tab->rotderv, tab->rotesqr};
      const double **dst[3] = {&p.rdens, &p.rderv, &p.resqr}, **dst2[3] = {&p.rdens2, &p.rderv2, &p.resqr2};
      for (int k = 0; k < 3; k++) {
         std::vector<double> y(cols[k], cols[k] + n);
         spline_setup(g, y, y2);
         if (dupload(dst[k], y.data(), n) || dupload(dst2[k], y2.data(), n)) return 1;
         if (k == 0) {
            std::vector<SplineRec> rec = make_records(g, y, y2);
            if (dupload(&p.recrot, rec.data(), rec.size())) return 1;
         }
      }
      std::vector<int> lut;
      build_lut(g, lut, p.lutrot_scale);
      p.nrot = n; p.nlutrot = (int)lut.size();
      if (dupload(&p.rgrid, g.data(), n) || dupload(&p.lutrot, lut.data(), lut.size())) return 1;
   }
   if (p.Q > 0 && p.molecule[p.imtype] == 2 && (p.rotden_type == 0 || tab->rho3d)) {
      if (!tab->rho3d || !tab->erot3d || !tab->esq3d) return fail("pimcgpu_init: the rho/eng/esq density-matrix tables are required");
      if (dupload(&p.rho3, tab->rho3d, PIMCGPU_SIZE_ROTDEN) || dupload(&p.erot3, tab->erot3d, PIMCGPU_SIZE_ROTDEN) || dupload(&p.esq3, tab->esq3d, PIMCGPU_SIZE_ROTDEN)) return 1;
   }
   // ---- state ----
   p.nchains = sys->nchains;
   p.S = p.P + p.Q + 8;
   const size_t C = p.nchains;
   if (dalloc(&p.pos, C * p.P * 3 * p.Npad) || dalloc(&p.ang, C * std::max(1, p.Q) * 3 * p.NMpad) || dalloc(&p.cosn, C * std::max(1, p.Q) * 3 * p.NMpad) ||
       dalloc(&p.pindex, C * p.N) || dalloc(&p.cyc_start, C * (p.N + 1)) || dalloc(&p.cyc_atoms, C * p.N) || dalloc(&p.ncyc, C * MAXT) ||
       dalloc(&p.vold, C * std::max(1, p.Q) * p.NMpad) || dalloc(&p.vepoch, C * std::max(1, p.Q) * p.NMpad) || dalloc(&p.pos_epoch, C) ||
       dalloc(&p.wstate, C * 8) || dalloc(&p.rindex, C * p.N) || dalloc(&p.qwc, C * 16) ||
       dalloc(&p.rng, C * p.S * 6) || dalloc(&p.counters, C * MAXT * 3 * 2) || dalloc(&p.scratch, C * 64) || dalloc(&G.d_err, 1)) return 1;
   G.h_pindex.assign(C * p.N, 0);
   {
      std::vector<double> q0(C * 16, 0.0);
      for (size_t c = 0; c < C; c++) q0[c * 16 + 14] = 1.0;      // countQW starts at 1 (ResetQWCounts, mc_qworm.cc:669-678)
      CK(cudaMemcpy(p.qwc, q0.data(), q0.size() * sizeof(double), cudaMemcpyHostToDevice));
   }
   G.stage_chain = (size_t)p.P * 3 * p.Npad + 2 * (size_t)std::max(1, p.Q) * 3 * p.NMpad;
   CK(cudaHostAlloc((void **)&G.stage, C * G.stage_chain * sizeof(double), cudaHostAllocDefault));
   memset(G.stage, 0, C * G.stage_chain * sizeof(double));
   if (dalloc(&G.d_raw, (size_t)3 * p.N * p.P)) return 1;
   // ---- execution geometry ----
   int seg_max = 1, seg_min = 1 << 30;
   for (int t = 0; t < p.ntypes; t++) { seg_max = std::max(seg_max, 1 << p.levels[t]); seg_min = std::min(seg_min, 1 << p.levels[t]); }
   p.seg_max = seg_max;
   p.team_buf_n = team_buf_doubles(seg_max);
   // one rotor, even Q: the potential sums of all Q proposals are one parallel stage and the decisions are pipelined
   // (rot_sweep_pipe).  A chain that lives in one CTA keeps the two-phase sweep when a top's four density look-ups
   // would have to share fewer than four threads per slice (measured on C3: 676 vs 847 M bead-updates/s).
   p.rot_fused = (p.Q > 0 && p.NM == 1 && p.Q % 2 == 0) ? 1 : 0;
   if (getenv("PIMC_NO_FUSED_ROT")) p.rot_fused = 0;
   const bool top = p.imtype >= 0 && p.molecule[p.imtype] == 2;
   int team = 1, threads = 0, cpc = 0, rot_units = 0, units = 0;
   const int nseg_widest = p.P / seg_min;
   for (int attempt = 0; attempt < 2; attempt++) {
      rot_units = p.rot_fused ? p.Q : p.Q / 2;                // rot slices that are independent within one stage
      team = sys->team;
      if (team <= 0) team = pow2ceil(std::min(32, std::max(1, p.R * (p.N - 1) / 2)));
      team = std::min(128, pow2floor(std::max(1, team)));
      units = std::max(rot_units, nseg_widest);               // widest stage: rot slices / segments of one atom
      const long rot_work = (long)rot_units * p.R * std::max(1, p.N - 1);                 // pair terms of one rotational stage
      const long bis_work = (long)nseg_widest * (seg_min - 1) * 2 * std::max(1, p.N - 1); // pair terms of one atom's bisection
      long want = std::max<long>((long)units * team, std::max(rot_work / 4, bis_work / 8));
      if (top && p.rot_fused) want = std::max<long>(want, 4L * rot_units);               // four density look-ups per slice
      threads = sys->threads_per_cta; cpc = sys->ctas_per_chain;
      if (cpc <= 0) {
         // enough CTAs for a few pair terms per thread in the widest stage, but never more CTAs than the 148 SMs can hold
         // at once.  Up to 8 CTAs per chain form a cluster; 16 CTAs per chain run as a cooperative grid with a software
         // chain barrier, because only seven 16-CTA clusters fit on a B200 at once (cudaOccupancyMaxActiveClusters)
         long useful = (want + 511) / 512;
         long fit = std::max(1, 148 / std::max(1, p.nchains));
         cpc = (int)std::max<long>(1, std::min<long>(16, std::min(useful, fit)));
         while (cpc & (cpc - 1)) cpc &= cpc - 1;                 // power of two
      }
      if (threads <= 0) {
         long per = (want + cpc - 1) / cpc;
         threads = (int)std::min<long>(512, std::max<long>(64, ((per + 31) / 32) * 32));
      }
      const int per_cta = (p.Q + cpc - 1) / std::max(1, cpc);
      // ... unless the sweeps run free (rot_run_cta): there a slice's sums overlap its neighbours' decisions and two threads per
      // slice are enough (C3: 940 vs 783 M bead-updates/s)
      const bool run_cta = top && cpc == 1 && p.imtype == p.ntypes - 1 && !p.worm_on && p.Q >= 2 && threads >= p.Q && !getenv("PIMC_NO_ROT_RUN");
      if (p.rot_fused && cpc == 1 && top && threads / std::max(1, per_cta) < 4 && !run_cta && !getenv("PIMC_FORCE_FUSED_ROT")) { p.rot_fused = 0; continue; }
      break;
   }
   if (threads % 32 || threads > PIMC_MAX_THREADS || threads < 32) return fail("pimcgpu_init: threads_per_cta must be a multiple of 32 in [32,%d]", PIMC_MAX_THREADS);
   if (cpc > 32 || (cpc & (cpc - 1))) return fail("pimcgpu_init: ctas_per_chain must be a power of two <= 32");
   if (sys->team <= 0) {
      // more threads than one team per segment can use: widen the teams (several warps per segment)
      while (team < 128 && (long)cpc * threads / team >= 2L * nseg_widest && std::max(1, p.N - 1) >= 2 * team && 2 * team <= threads) team *= 2;
   }
   if (sys->team <= 0) {
      // the gaussians and the bead updates of a segment use every lane of its team (not only the partner sums): widen up
      // to half a warp while the chain has more threads than segments (C1: 2 -> 16 lanes, +16 % measured)
      while (team < 16 && (long)cpc * threads / (2 * team) >= nseg_widest) team *= 2;
   }
   if (team > threads) team = pow2floor(threads);
   if (team > 32 && threads / team > 15) return fail("pimcgpu_init: a team wider than a warp needs at most 15 teams per CTA");
   p.team = team; p.cpc = cpc;
   p.swbar = (cpc > 8 || getenv("PIMC_SWBAR")) && cpc > 1 ? 1 : 0;
   // rot group: the threads of a CTA are split evenly over the slices it owns in one stage
   {
      int count = std::max(1, p.rot_fused ? p.Q : (p.Q + 1) / 2);
      int per_cta = (count + cpc - 1) / cpc;
      int rg = pow2floor(std::max(1, threads / std::max(1, std::min(per_cta, threads))));
      long work = (long)p.R * std::max(1, p.N - 1);           // partner terms of one rot step
      while (rg > 1 && rg > 2 * work) rg >>= 1;               // no point in more threads than terms
      p.rot_group = std::min(rg, threads);
   }
   G.threads = threads;
   {
      const int nteams = cpc * threads / team;
      p.nseg_max = ((nseg_widest + nteams - 1) / nteams) * nteams;       // whole rounds: every team of a round has its own buffer
   }
   p.rot_in_smem = 1;
   p.segbuf_global = 0;
   if (smem_bytes(p, threads) > 200 * 1024) p.rot_in_smem = 0;
   if (smem_bytes(p, threads) > 200 * 1024) { p.rot_in_smem = 1; p.segbuf_global = 1; }
   if (smem_bytes(p, threads) > 200 * 1024) p.rot_in_smem = 0;
   if (p.segbuf_global && dalloc(&p.segbuf, C * p.nseg_max * p.team_buf_n)) return 1;
   if (dalloc(&p.barrier, C * 32)) return 1;
   p.bis_piped = (!p.segbuf_global && !getenv("PIMC_NO_BIS_PIPE")) ? 1 : 0;
   p.mol_piped = getenv("PIMC_NO_MOL_PIPE") ? 0 : 1;
   // free-running rotational sweeps (rot_run): one linear rotor listed as the last species, several CTAs per chain, every CTA a
   // contiguous block of slices with one rot group per slice
   p.rot_run = 0;
   if (p.rot_fused && cpc > 1 && p.imtype == p.ntypes - 1 && p.molecule[p.imtype] == 1 && !p.worm_on && p.Q % cpc == 0 && p.Q / cpc <= threads / p.rot_group &&
       p.Q / cpc >= 2 && !getenv("PIMC_NO_ROT_RUN")) {
      if (dalloc(&p.rot_ll, C * p.Q * 8)) return 1;
      p.rot_run = 1;
   }
   p.rot_run_cta = (p.rot_fused && cpc == 1 && p.imtype == p.ntypes - 1 && p.molecule[p.imtype] == 2 && !p.worm_on && p.Q >= 2 && p.Q % 2 == 0 &&
                    p.Q <= threads / p.rot_group && (p.rot_group >= 32 || (p.Q * p.rot_group) % 32 == 0) &&      // whole warps: the groups of a warp re-converge with a full __syncwarp
                    !getenv("PIMC_NO_ROT_RUN")) ? 1 : 0;
   // geometry cache of the rotor-atom terms (rot_potential_cached): one linear rotor among atoms, pipelined sweep, no worm
   p.geo_on = 0;
   p.geo_hint = getenv("PIMC_GEO_HINT") ? atoi(getenv("PIMC_GEO_HINT")) : 0;
   if (p.rot_fused && p.imtype >= 0 && p.molecule[p.imtype] == 1 && !p.worm_on && !p.minimage && p.N > 1 && p.rs2d > 0 && p.rs2d < 4096 && !getenv("PIMC_NO_GEO")) {
      p.geo_items = p.R * (p.N - 1);
      p.geo_n = (p.geo_items + 31) / 32 * 32;
      if (dalloc(&p.geo, C * p.Q * 4 * p.geo_n)) return 1;
      p.geo_on = 1;
   }
   G.smem = smem_bytes(p, threads);
   if (G.smem > 227 * 1024) return fail("pimcgpu_init: %zu bytes of shared memory per CTA exceed the 227 KB limit", G.smem);
   G.kind = p.imtype >= 0 && p.Q > 0 ? p.molecule[p.imtype] : (p.imtype >= 0 ? p.molecule[p.imtype] : 0);
   const void *kfun = steps_kernel(G.kind, p.worm_on, p.rot_run_cta);
   CK(cudaFuncSetAttribute(kfun, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
   if (p.swbar) {
      int per_sm = 0;
      cudaDeviceProp prop;
      CK(cudaGetDeviceProperties(&prop, sys->device));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfun, threads, G.smem));
      if ((long)per_sm * prop.multiProcessorCount < (long)p.nchains * cpc)
         return fail("pimcgpu_init: %d chains x %d CTAs cannot be co-resident on %d SMs (needed by the software chain barrier)", p.nchains, cpc, prop.multiProcessorCount);
   }
   // ---- estimator buffers / accumulator layout ----
   EstBuffers &e = G.e;
   memset(&e, 0, sizeof e);
   std::vector<int> pairs;
   for (int a0 = 0; a0 < p.N - 1; a0++)
      for (int a1 = a0 + 1; a1 < p.N; a1++) { pairs.push_back(a0); pairs.push_back(a1); }
   e.npairs = (int)pairs.size() / 2;
   if (pairs.empty()) { pairs.push_back(0); pairs.push_back(0); }          // a lone particle has no pairs
   if (dupload(&e.pairs, pairs.data(), pairs.size())) return 1;
   e.has_gr3d = (need3d || needsph) ? 1 : 0;
   e.off_gr1d = 32; e.off_gr2d = e.off_gr1d + BINSR; e.off_rcf = e.off_gr2d + (long)BINSR * BINST;
   e.off_relbins = e.off_rcf + std::max(1, p.Q); e.off_area = e.off_relbins + BINST + 2 * BINSC;
   e.off_ploops = e.off_area + NAREA_ACC; e.off_rcfcnt = e.off_ploops + std::max(1, p.bstype >= 0 ? p.numb[p.bstype] : 1);
   e.off_gr3d = e.off_rcfcnt + std::max(1, p.Q);
   G.nacc = e.off_gr3d + (e.has_gr3d ? (long)BINSR * BINST * BINSC : 0);
   if (dalloc(&e.com, C * p.P * 3) || dalloc(&e.area_partials, C * EST_BLOCKS * NAREA) || dalloc(&e.chain_area, C * NAREA) || dalloc(&G.d_ops, C * 4)) return 1;
   if (dalloc(&e.acc, G.nacc) || dalloc(&e.partials, C * EST_BLOCKS * NPART) || dalloc(&e.chain_e, C * 8) || dalloc(&e.chain_rcf, C * 2 * std::max(1, p.Q))) return 1;
   CK(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
   CK(cudaDeviceSynchronize());
   G.live = true;
   G.step = 0;
   G.seeded = false;
   return 0;
}

}  // extern "C"
namespace {
// The small arrays of a split-phase upload go from their device staging area into the state (angles, axes, permutation
// tables, closed worm, stale potential cache): device work only, so a commit never queues behind a configuration that is
// still travelling on a copy engine.
__global__ void commit_small_kernel(Params p, int first, int count, const double *sd, const int *sp, int has_rot)
{
   const size_t nang = (size_t)max(1, p.Q) * 3 * p.NMpad, per = 4 * (size_t)p.N + 1 + MAXT, nve = (size_t)max(1, p.Q) * p.NMpad;
   const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
   if (has_rot)
      for (size_t i = t0; i < (size_t)count * nang; i += stride) {
         const size_t cc = i / nang, k = i - cc * nang;
         p.ang[(first + cc) * nang + k] = sd[cc * 2 * nang + k];
         p.cosn[(first + cc) * nang + k] = sd[cc * 2 * nang + nang + k];
      }
   const size_t N = (size_t)p.N;
   for (size_t i = t0; i < (size_t)count * per; i += stride) {
      const size_t cc = i / per, k = i - cc * per, c = first + cc;
      const int v = sp[i];
      if (k < N) p.pindex[c * N + k] = v;
      else if (k < 2 * N) p.rindex[c * N + (k - N)] = v;
      else if (k < 3 * N + 1) p.cyc_start[c * (N + 1) + (k - 2 * N)] = v;
      else if (k < 4 * N + 1) p.cyc_atoms[c * N + (k - 3 * N - 1)] = v;
      else p.ncyc[c * MAXT + (k - 4 * N - 1)] = v;
   }
   for (size_t i = t0; i < (size_t)count * 8; i += stride) p.wstate[(size_t)first * 8 + i] = 0;
   for (size_t i = t0; i < (size_t)count * nve; i += stride) p.vepoch[(size_t)first * nve + i] = -1;
}
// the rotor rows of a split-phase download into their device staging area ([chain][angles | axes])
__global__ void snapshot_small_kernel(Params p, int first, int count, double *sd)
{
   const size_t nang = (size_t)max(1, p.Q) * 3 * p.NMpad;
   for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)count * nang; i += (size_t)gridDim.x * blockDim.x) {
      const size_t cc = i / nang, k = i - cc * nang;
      sd[cc * 2 * nang + k] = p.ang[(first + cc) * nang + k];
      sd[cc * 2 * nang + nang + k] = p.cosn[(first + cc) * nang + k];
   }
}

// Beads between the reference layout raw[d][atom*P + it] (mc_setup.cc:139-148) and the device layout pos[it][d][Npad]:
// a tiled transpose through shared memory, both sides coalesced.  to_device != 0 imports, else exports.
__global__ void state_transpose_kernel(double *pos, double *raw, int N, int P, int Npad, int to_device)
{
   __shared__ double tile[32][33];
   const int d = blockIdx.z % 3, it0 = blockIdx.x * 32, a0 = blockIdx.y * 32;
   const size_t n = (size_t)N * P;
   pos += (size_t)(blockIdx.z / 3) * P * 3 * Npad;          // blockIdx.z = 3 * chain + dim: consecutive chains of a batch
   raw += (size_t)(blockIdx.z / 3) * 3 * n;
   if (to_device) {
      for (int r = threadIdx.y; r < 32; r += blockDim.y) {          // r: atom in tile, x: slice
         const int a = a0 + r, it = it0 + threadIdx.x;
         if (a < N && it < P) tile[r][threadIdx.x] = raw[d * n + (size_t)a * P + it];
      }
      __syncthreads();
      for (int r = threadIdx.y; r < 32; r += blockDim.y) {          // r: slice in tile, x: atom
         const int it = it0 + r, a = a0 + threadIdx.x;
         if (a < N && it < P) pos[((size_t)it * 3 + d) * Npad + a] = tile[threadIdx.x][r];
      }
   } else {
      for (int r = threadIdx.y; r < 32; r += blockDim.y) {
         const int it = it0 + r, a = a0 + threadIdx.x;
         if (a < N && it < P) tile[threadIdx.x][r] = pos[((size_t)it * 3 + d) * Npad + a];
      }
      __syncthreads();
      for (int r = threadIdx.y; r < 32; r += blockDim.y) {
         const int a = a0 + r, it = it0 + threadIdx.x;
         if (a < N && it < P) raw[d * n + (size_t)a * P + it] = tile[r][threadIdx.x];
      }
   }
}
}  // namespace
extern "C" {

int pimcgpu_upload_state(int chain, const double *coords, const double *angles, const int *pindex)
{
   if (!G.live) return fail("pimcgpu_upload_state: not initialised");
   const Params &p = G.p;
   if (chain < -1 || chain >= p.nchains) return fail("pimcgpu_upload_state: chain %d out of range", chain);
   const size_t n = (size_t)p.N * p.P;
   const size_t npos = (size_t)p.P * 3 * p.Npad, nang = (size_t)std::max(1, p.Q) * 3 * p.NMpad;
   const int cfirst = chain < 0 ? 0 : chain;
   double *hpos = G.stage + (size_t)cfirst * G.stage_chain, *hang = hpos + npos, *hcos = hang + nang;
   CK(cudaStreamSynchronize(G.stream));          // the staging area may still be in flight
   // beads: the caller's [dim][atom*P + it] array goes to the device as it is and is transposed there (state_transpose_kernel)
   CK(cudaMemcpyAsync(G.d_raw, coords, 3 * n * sizeof(double), cudaMemcpyHostToDevice, G.stream));
   if (p.imtype >= 0)
      for (int q = 0; q < p.Q; q++)
         for (int m = 0; m < p.NM; m++) {
            size_t src = (size_t)(p.first[p.imtype] + m) * p.P + q;
            double phi = angles[0 * n + src], cost = angles[1 * n + src], chi = angles[2 * n + src];
            double sint = sqrt(1.0 - cost * cost);         // MCCosine from (phi, cos theta), mc_main.cc:192-199
            size_t b = (size_t)q * 3 * p.NMpad + m;
            hang[b] = phi; hang[b + p.NMpad] = cost; hang[b + 2 * p.NMpad] = chi;
            hcos[b] = sint * cos(phi); hcos[b + p.NMpad] = sint * sin(phi); hcos[b + 2 * p.NMpad] = cost;
         }
   // permutation: pindex[] of the boson type in type-local numbering (PIndex, mc_setup.h:97) -> global next world line
   std::vector<int> gp(p.N), cstart, catoms;
   for (int a = 0; a < p.N; a++) gp[a] = a;
   if (pindex && p.bstype >= 0)
      for (int a = 0; a < p.numb[p.bstype]; a++) {
         if (pindex[a] < 0 || pindex[a] >= p.numb[p.bstype]) return fail("pimcgpu_upload_state: bad permutation entry");
         gp[p.first[p.bstype] + a] = p.first[p.bstype] + pindex[a];
      }
   std::vector<int> gr(p.N);
   for (int a = 0; a < p.N; a++) gr[gp[a]] = a;
   std::vector<int> ncyc(MAXT, 0), seen(p.N, 0);
   for (int t = 0; t < p.ntypes; t++)
      for (int a = p.first[t]; a < p.first[t] + p.numb[t]; a++) {
         if (seen[a]) continue;
         cstart.push_back((int)catoms.size());
         int b = a, guard = 0;
         do { catoms.push_back(b); seen[b] = 1; b = gp[b]; } while (b != a && ++guard <= p.N);
         if (b != a) return fail("pimcgpu_upload_state: permutation is not a bijection");
         ncyc[t]++;
      }
   while ((int)cstart.size() < p.N + 1) cstart.push_back((int)catoms.size());
   int c0 = chain < 0 ? 0 : chain, c1 = chain < 0 ? p.nchains : chain + 1;
   for (int c = c0; c < c1; c++) {
      state_transpose_kernel<<<dim3((p.P + 31) / 32, (p.N + 31) / 32, 3), dim3(32, 8), 0, G.stream>>>(p.pos + (size_t)c * npos, G.d_raw, p.N, p.P, p.Npad, 1);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(p.ang + (size_t)c * nang, hang, nang * sizeof(double), cudaMemcpyHostToDevice, G.stream));
      CK(cudaMemcpyAsync(p.cosn + (size_t)c * nang, hcos, nang * sizeof(double), cudaMemcpyHostToDevice, G.stream));
      CK(cudaMemcpyAsync(p.pindex + (size_t)c * p.N, gp.data(), p.N * sizeof(int), cudaMemcpyHostToDevice, G.stream));
      CK(cudaMemcpyAsync(p.rindex + (size_t)c * p.N, gr.data(), p.N * sizeof(int), cudaMemcpyHostToDevice, G.stream));
      CK(cudaMemsetAsync(p.wstate + (size_t)c * 8, 0, 8 * sizeof(int), G.stream));              // uploaded paths are closed (Z sector)
      CK(cudaMemcpyAsync(p.cyc_start + (size_t)c * (p.N + 1), cstart.data(), (p.N + 1) * sizeof(int), cudaMemcpyHostToDevice, G.stream));
      CK(cudaMemcpyAsync(p.cyc_atoms + (size_t)c * p.N, catoms.data(), p.N * sizeof(int), cudaMemcpyHostToDevice, G.stream));
      CK(cudaMemcpyAsync(p.ncyc + (size_t)c * MAXT, ncyc.data(), MAXT * sizeof(int), cudaMemcpyHostToDevice, G.stream));
      std::copy(gp.begin(), gp.end(), G.h_pindex.begin() + (size_t)c * p.N);
      CK(cudaMemsetAsync(p.vepoch + (size_t)c * std::max(1, p.Q) * p.NMpad, 0xff, (size_t)std::max(1, p.Q) * p.NMpad * sizeof(int), G.stream));
   }
   CK(cudaStreamSynchronize(G.stream));
   return 0;
}

int pimcgpu_download_state(int chain, double *coords, double *angles, double *cosine, int *pindex)
{
   if (!G.live) return fail("pimcgpu_download_state: not initialised");
   const Params &p = G.p;
   if (chain < 0 || chain >= p.nchains) return fail("pimcgpu_download_state: chain %d out of range", chain);
   const size_t n = (size_t)p.N * p.P;
   const size_t npos = (size_t)p.P * 3 * p.Npad, nang = (size_t)std::max(1, p.Q) * 3 * p.NMpad;
   double *hpos = G.stage + (size_t)chain * G.stage_chain, *hang = hpos + npos, *hcos = hang + nang;
   if (coords) {   // transposed on the device into the reference layout, then one copy into the caller's array
      state_transpose_kernel<<<dim3((p.P + 31) / 32, (p.N + 31) / 32, 3), dim3(32, 8), 0, G.stream>>>(p.pos + (size_t)chain * npos, G.d_raw, p.N, p.P, p.Npad, 0);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(coords, G.d_raw, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
   }
   CK(cudaMemcpyAsync(hang, p.ang + (size_t)chain * nang, nang * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
   CK(cudaMemcpyAsync(hcos, p.cosn + (size_t)chain * nang, nang * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
   CK(cudaStreamSynchronize(G.stream));
   // rotor rows: only the first Q entries of a molecule's row carry angles (README.md:228); the rest keep
   // the reference's initial values phi=0, cos(theta)=1, chi=0 (MCConfigInit, mc_setup.cc:471-487)
   #pragma omp parallel for collapse(2) schedule(static)
   for (int d = 0; d < 3; d++)
      for (size_t i = 0; i < n; i++) {
         if (angles) angles[d * n + i] = (d == 1) ? 1.0 : 0.0;
         if (cosine) cosine[d * n + i] = (d == 2) ? 1.0 : 0.0;
      }
   if (p.imtype >= 0)
      for (int q = 0; q < p.Q; q++)
         for (int m = 0; m < p.NM; m++) {
            size_t dst = (size_t)(p.first[p.imtype] + m) * p.P + q, b = (size_t)q * 3 * p.NMpad + m;
            for (int d = 0; d < 3; d++) {
               if (angles) angles[d * n + dst] = hang[b + d * p.NMpad];
               if (cosine) cosine[d * n + dst] = hcos[b + d * p.NMpad];
            }
         }
   if (pindex && p.bstype >= 0) {
      CK(cudaMemcpy(G.h_pindex.data() + (size_t)chain * p.N, p.pindex + (size_t)chain * p.N, p.N * sizeof(int), cudaMemcpyDeviceToHost));   // the worm's swaps change it
      for (int a = 0; a < p.numb[p.bstype]; a++) pindex[a] = G.h_pindex[(size_t)chain * p.N + p.first[p.bstype] + a] - p.first[p.bstype];
   }
   return 0;
}

// ---- batched state transfer: chains first .. first+count-1 in one go (arrays [count][3][N*P], permutations [count][nb]) ----
static int permutation_tables(const Params &p, const int *pindex, int *gp, int *gr, int *cstart, int *catoms, int *ncyc, const char *who)
{
   for (int a = 0; a < p.N; a++) gp[a] = a;
   if (pindex && p.bstype >= 0)
      for (int a = 0; a < p.numb[p.bstype]; a++) {
         if (pindex[a] < 0 || pindex[a] >= p.numb[p.bstype]) return fail("%s: bad permutation entry", who);
         gp[p.first[p.bstype] + a] = p.first[p.bstype] + pindex[a];
      }
   for (int a = 0; a < p.N; a++) gr[gp[a]] = a;
   std::vector<char> seen(p.N, 0);
   int ns = 0, na = 0;
   for (int t = 0; t < MAXT; t++) ncyc[t] = 0;
   for (int t = 0; t < p.ntypes; t++)
      for (int a = p.first[t]; a < p.first[t] + p.numb[t]; a++) {
         if (seen[a]) continue;
         cstart[ns++] = na;
         int b = a, guard = 0;
         do { catoms[na++] = b; seen[b] = 1; b = gp[b]; } while (b != a && ++guard <= p.N);
         if (b != a) return fail("%s: permutation is not a bijection", who);
         ncyc[t]++;
      }
   while (ns < p.N + 1) cstart[ns++] = na;
   return 0;
}

int pimcgpu_upload_states(int first, int count, const double *coords, const double *angles, const int *pindex)
{
   if (!G.live) return fail("pimcgpu_upload_states: not initialised");
   const Params &p = G.p;
   if (first < 0 || count < 1 || first + count > p.nchains) return fail("pimcgpu_upload_states: chains %d..%d out of range", first, first + count - 1);
   const size_t n = (size_t)p.N * p.P;
   const size_t npos = (size_t)p.P * 3 * p.Npad, nang = (size_t)std::max(1, p.Q) * 3 * p.NMpad;
   if (!G.d_raw_all && dalloc(&G.d_raw_all, (size_t)p.nchains * 3 * n)) return 1;
   CK(cudaStreamSynchronize(G.stream));          // the staging area may still be in flight
   CK(cudaMemcpyAsync(G.d_raw_all, coords, (size_t)count * 3 * n * sizeof(double), cudaMemcpyHostToDevice, G.stream));
   state_transpose_kernel<<<dim3((p.P + 31) / 32, (p.N + 31) / 32, 3 * count), dim3(32, 8), 0, G.stream>>>(p.pos + (size_t)first * npos, G.d_raw_all, p.N, p.P, p.Npad, 1);
   CK(cudaGetLastError());
   if (p.imtype >= 0) {
      // a few thousand sincos at most: an OpenMP team only pays off for many chains (and eight ranks share the host)
      #pragma omp parallel for schedule(static) if ((long)count * p.Q * p.NM > 65536) num_threads(4)
      for (int cc = 0; cc < count; cc++) {
         double *hang = G.stage + (size_t)(first + cc) * G.stage_chain + npos, *hcos = hang + nang;
         const double *ang = angles + (size_t)cc * 3 * n;
         for (int q = 0; q < p.Q; q++)
            for (int m = 0; m < p.NM; m++) {
               const size_t src = (size_t)(p.first[p.imtype] + m) * p.P + q;
               const double phi = ang[0 * n + src], cost = ang[1 * n + src], chi = ang[2 * n + src];
               const double sint = sqrt(1.0 - cost * cost);         // MCCosine from (phi, cos theta), mc_main.cc:192-199
               const size_t b = (size_t)q * 3 * p.NMpad + m;
               hang[b] = phi; hang[b + p.NMpad] = cost; hang[b + 2 * p.NMpad] = chi;
               hcos[b] = sint * cos(phi); hcos[b + p.NMpad] = sint * sin(phi); hcos[b + 2 * p.NMpad] = cost;
            }
      }
      const double *s0 = G.stage + (size_t)first * G.stage_chain + npos;
      CK(cudaMemcpy2DAsync(p.ang + (size_t)first * nang, nang * sizeof(double), s0, G.stage_chain * sizeof(double), nang * sizeof(double), count, cudaMemcpyHostToDevice, G.stream));
      CK(cudaMemcpy2DAsync(p.cosn + (size_t)first * nang, nang * sizeof(double), s0 + nang, G.stage_chain * sizeof(double), nang * sizeof(double), count, cudaMemcpyHostToDevice, G.stream));
   }
   const int nb = p.bstype >= 0 ? p.numb[p.bstype] : 0;
   std::vector<int> gp((size_t)count * p.N), gr((size_t)count * p.N), cst((size_t)count * (p.N + 1)), cat((size_t)count * p.N), ncy((size_t)count * MAXT);
   for (int cc = 0; cc < count; cc++)
      if (permutation_tables(p, pindex ? pindex + (size_t)cc * nb : nullptr, &gp[(size_t)cc * p.N], &gr[(size_t)cc * p.N], &cst[(size_t)cc * (p.N + 1)],
                             &cat[(size_t)cc * p.N], &ncy[(size_t)cc * MAXT], "pimcgpu_upload_states")) return 1;
   CK(cudaMemcpyAsync(p.pindex + (size_t)first * p.N, gp.data(), gp.size() * sizeof(int), cudaMemcpyHostToDevice, G.stream));
   CK(cudaMemcpyAsync(p.rindex + (size_t)first * p.N, gr.data(), gr.size() * sizeof(int), cudaMemcpyHostToDevice, G.stream));
   CK(cudaMemcpyAsync(p.cyc_start + (size_t)first * (p.N + 1), cst.data(), cst.size() * sizeof(int), cudaMemcpyHostToDevice, G.stream));
   CK(cudaMemcpyAsync(p.cyc_atoms + (size_t)first * p.N, cat.data(), cat.size() * sizeof(int), cudaMemcpyHostToDevice, G.stream));
   CK(cudaMemcpyAsync(p.ncyc + (size_t)first * MAXT, ncy.data(), ncy.size() * sizeof(int), cudaMemcpyHostToDevice, G.stream));
   CK(cudaMemsetAsync(p.wstate + (size_t)first * 8, 0, (size_t)count * 8 * sizeof(int), G.stream));                          // uploaded paths are closed
   CK(cudaMemsetAsync(p.vepoch + (size_t)first * std::max(1, p.Q) * p.NMpad, 0xff, (size_t)count * std::max(1, p.Q) * p.NMpad * sizeof(int), G.stream));
   std::copy(gp.begin(), gp.end(), G.h_pindex.begin() + (size_t)first * p.N);
   CK(cudaStreamSynchronize(G.stream));
   return 0;
}

int pimcgpu_download_states(int first, int count, double *coords, double *angles, double *cosine)
{
   if (!G.live) return fail("pimcgpu_download_states: not initialised");
   const Params &p = G.p;
   if (first < 0 || count < 1 || first + count > p.nchains) return fail("pimcgpu_download_states: chains %d..%d out of range", first, first + count - 1);
   const size_t n = (size_t)p.N * p.P;
   const size_t npos = (size_t)p.P * 3 * p.Npad, nang = (size_t)std::max(1, p.Q) * 3 * p.NMpad;
   if (!G.d_raw_all && dalloc(&G.d_raw_all, (size_t)p.nchains * 3 * n)) return 1;
   if (coords) {
      state_transpose_kernel<<<dim3((p.P + 31) / 32, (p.N + 31) / 32, 3 * count), dim3(32, 8), 0, G.stream>>>(p.pos + (size_t)first * npos, G.d_raw_all, p.N, p.P, p.Npad, 0);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(coords, G.d_raw_all, (size_t)count * 3 * n * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
   }
   double *s0 = G.stage + (size_t)first * G.stage_chain + npos;
   if (angles || cosine) {
      CK(cudaMemcpy2DAsync(s0, G.stage_chain * sizeof(double), p.ang + (size_t)first * nang, nang * sizeof(double), nang * sizeof(double), count, cudaMemcpyDeviceToHost, G.stream));
      CK(cudaMemcpy2DAsync(s0 + nang, G.stage_chain * sizeof(double), p.cosn + (size_t)first * nang, nang * sizeof(double), nang * sizeof(double), count, cudaMemcpyDeviceToHost, G.stream));
   }
   CK(cudaStreamSynchronize(G.stream));
   if (angles || cosine) {
      // rotor rows: only the first Q entries carry angles; the rest keep phi = 0, cos(theta) = 1, chi = 0 (MCConfigInit, mc_setup.cc:471-487)
      #pragma omp parallel for schedule(static) num_threads(std::min(count, 8))
      for (int cc = 0; cc < count; cc++) {
         const double *hang = G.stage + (size_t)(first + cc) * G.stage_chain + npos, *hcos = hang + nang;
         double *ang = angles ? angles + (size_t)cc * 3 * n : nullptr, *cs = cosine ? cosine + (size_t)cc * 3 * n : nullptr;
         for (int d = 0; d < 3; d++)
            for (size_t i = 0; i < n; i++) {
               if (ang) ang[d * n + i] = (d == 1) ? 1.0 : 0.0;
               if (cs) cs[d * n + i] = (d == 2) ? 1.0 : 0.0;
            }
         if (p.imtype >= 0)
            for (int q = 0; q < p.Q; q++)
               for (int m = 0; m < p.NM; m++) {
                  const size_t dst = (size_t)(p.first[p.imtype] + m) * p.P + q, b = (size_t)q * 3 * p.NMpad + m;
                  for (int d = 0; d < 3; d++) {
                     if (ang) ang[d * n + dst] = hang[b + d * p.NMpad];
                     if (cs) cs[d * n + dst] = hcos[b + d * p.NMpad];
                  }
               }
      }
   }
   return 0;
}

// Like pimcgpu_download_states, but only the rows the moves ever change are written to `angles` / `cosine`: entries
// [rotor atom * P + q], q < Q.  Everything else in the caller's arrays is left as it is -- the reference allocates
// MCAngles / MCCosine once (mc_setup.cc:135-163) and never touches the other rows after MCConfigInit (:471-487), so a
// driver that keeps its arrays across blocks gets the reference's contents with 6 Q NM doubles per chain instead of 6 N P.
int pimcgpu_download_states_rows(int first, int count, double *coords, double *angles, double *cosine)
{
   if (!G.live) return fail("pimcgpu_download_states_rows: not initialised");
   const Params &p = G.p;
   if (first < 0 || count < 1 || first + count > p.nchains) return fail("pimcgpu_download_states_rows: chains %d..%d out of range", first, first + count - 1);
   const size_t n = (size_t)p.N * p.P;
   const size_t npos = (size_t)p.P * 3 * p.Npad, nang = (size_t)std::max(1, p.Q) * 3 * p.NMpad;
   if (!G.d_raw_all && dalloc(&G.d_raw_all, (size_t)p.nchains * 3 * n)) return 1;
   if (coords) {
      state_transpose_kernel<<<dim3((p.P + 31) / 32, (p.N + 31) / 32, 3 * count), dim3(32, 8), 0, G.stream>>>(p.pos + (size_t)first * npos, G.d_raw_all, p.N, p.P, p.Npad, 0);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(coords, G.d_raw_all, (size_t)count * 3 * n * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
   }
   double *s0 = G.stage + (size_t)first * G.stage_chain + npos;
   const bool rows = (angles || cosine) && p.imtype >= 0 && p.Q > 0;
   if (rows) {
      CK(cudaMemcpy2DAsync(s0, G.stage_chain * sizeof(double), p.ang + (size_t)first * nang, nang * sizeof(double), nang * sizeof(double), count, cudaMemcpyDeviceToHost, G.stream));
      CK(cudaMemcpy2DAsync(s0 + nang, G.stage_chain * sizeof(double), p.cosn + (size_t)first * nang, nang * sizeof(double), nang * sizeof(double), count, cudaMemcpyDeviceToHost, G.stream));
   }
   CK(cudaStreamSynchronize(G.stream));
   if (rows)
      for (int cc = 0; cc < count; cc++) {
         const double *hang = G.stage + (size_t)(first + cc) * G.stage_chain + npos, *hcos = hang + nang;
         double *ang = angles ? angles + (size_t)cc * 3 * n : nullptr, *cs = cosine ? cosine + (size_t)cc * 3 * n : nullptr;
         for (int q = 0; q < p.Q; q++)
            for (int m = 0; m < p.NM; m++) {
               const size_t dst = (size_t)(p.first[p.imtype] + m) * p.P + q, b = (size_t)q * 3 * p.NMpad + m;
               for (int d = 0; d < 3; d++) {
                  if (ang) ang[d * n + dst] = hang[b + d * p.NMpad];
                  if (cs) cs[d * n + dst] = hcos[b + d * p.NMpad];
               }
            }
      }
   return 0;
}

// ---- split-phase transfers: the copies of one step overlap the move kernel of the neighbouring steps ------------------------
// upload:   _begin copies the beads, then the prepared small arrays (angles, axes, permutation tables), into device staging
//           buffers on the copy stream (may run while the move kernel of the previous step is still working on the state);
//           _commit, on the library's stream, waits for those copies and moves everything into the state with two kernels --
//           no copy-engine work on the library's stream, so a commit never queues behind a configuration still travelling.
//           The caller's bead array must stay untouched until the copy stream has read it (pimcgpu_upload_states_commit
//           followed by any synchronising call, or the next _begin).
// download: _begin snapshots the beads in the reference layout and the rotor rows on the library's stream (device kernels)
//           and lets the copy stream carry them to the host; the library's stream is free for the next upload / pass at
//           once.  _end waits for the copies and scatters the rotor rows (as pimcgpu_download_states_rows).
static int split_phase_setup(void)
{
   const Params &p = G.p;
   const size_t n = (size_t)p.N * p.P, nang = (size_t)std::max(1, p.Q) * 3 * p.NMpad;
   if (G.copy_stream) return 0;
   CK(cudaStreamCreateWithFlags(&G.copy_stream, cudaStreamNonBlocking));
   CK(cudaEventCreateWithFlags(&G.ev_up, cudaEventDisableTiming));
   CK(cudaEventCreateWithFlags(&G.ev_down, cudaEventDisableTiming));
   CK(cudaEventCreateWithFlags(&G.ev_done, cudaEventDisableTiming));
   CK(cudaEventCreateWithFlags(&G.ev_commit, cudaEventDisableTiming));
   CK(cudaEventRecord(G.ev_commit, G.stream));
   if (dalloc(&G.d_raw_up, (size_t)p.nchains * 3 * n)) return 1;
   if (dalloc(&G.d_small_up, (size_t)p.nchains * 2 * nang) || dalloc(&G.d_small_down, (size_t)p.nchains * 2 * nang)) return 1;
   if (dalloc(&G.d_perm_up, (size_t)p.nchains * (4 * (size_t)p.N + 1 + MAXT))) return 1;
   if (!G.d_raw_all && dalloc(&G.d_raw_all, (size_t)p.nchains * 3 * n)) return 1;
   CK(cudaHostAlloc((void **)&G.stage_up, 2 * (size_t)p.nchains * 2 * nang * sizeof(double), cudaHostAllocDefault));
   CK(cudaHostAlloc((void **)&G.stage_perm, 2 * (size_t)p.nchains * (4 * (size_t)p.N + 1 + MAXT) * sizeof(int), cudaHostAllocDefault));
   return 0;
}

int pimcgpu_upload_states_begin(int first, int count, const double *coords, const double *angles, const int *pindex)
{
   if (!G.live) return fail("pimcgpu_upload_states_begin: not initialised");
   const Params &p = G.p;
   if (first < 0 || count < 1 || first + count > p.nchains) return fail("pimcgpu_upload_states_begin: chains %d..%d out of range", first, first + count - 1);
   if (G.up_count) return fail("pimcgpu_upload_states_begin: an upload is already in flight (call pimcgpu_upload_states_commit)");
   if (split_phase_setup()) return 1;
   const size_t n = (size_t)p.N * p.P, nang = (size_t)std::max(1, p.Q) * 3 * p.NMpad;
   CK(cudaEventSynchronize(G.ev_up));                                 // the copies of the upload before last have left their pinned staging half
   CK(cudaStreamWaitEvent(G.copy_stream, G.ev_commit, 0));          // the previous commit has read the device staging buffers
   CK(cudaMemcpyAsync(G.d_raw_up, coords, (size_t)count * 3 * n * sizeof(double), cudaMemcpyHostToDevice, G.copy_stream));
   G.up_slot ^= 1;                                                    // the previous upload's small copies read the other half
   double *stage_up = G.stage_up + (size_t)G.up_slot * p.nchains * 2 * nang;
   int *stage_perm = G.stage_perm + (size_t)G.up_slot * p.nchains * (4 * (size_t)p.N + 1 + MAXT);
   if (p.imtype >= 0)
      for (int cc = 0; cc < count; cc++) {
         double *hang = stage_up + (size_t)cc * 2 * nang, *hcos = hang + nang;
         const double *ang = angles + (size_t)cc * 3 * n;
         for (int q = 0; q < p.Q; q++)
            for (int m = 0; m < p.NM; m++) {
               const size_t src = (size_t)(p.first[p.imtype] + m) * p.P + q;
               const double phi = ang[0 * n + src], cost = ang[1 * n + src], chi = ang[2 * n + src];
               const double sint = sqrt(1.0 - cost * cost);         // MCCosine from (phi, cos theta), mc_main.cc:192-199
               const size_t b = (size_t)q * 3 * p.NMpad + m;
               hang[b] = phi; hang[b + p.NMpad] = cost; hang[b + 2 * p.NMpad] = chi;
               hcos[b] = sint * cos(phi); hcos[b + p.NMpad] = sint * sin(phi); hcos[b + 2 * p.NMpad] = cost;
            }
      }
   const int nb = p.bstype >= 0 ? p.numb[p.bstype] : 0;
   const size_t per = 4 * (size_t)p.N + 1 + MAXT;
   for (int cc = 0; cc < count; cc++) {
      int *base = stage_perm + (size_t)cc * per;
      if (permutation_tables(p, pindex ? pindex + (size_t)cc * nb : nullptr, base, base + p.N, base + 2 * p.N, base + 3 * p.N + 1, base + 4 * p.N + 1, "pimcgpu_upload_states_begin")) return 1;
   }
   // the small arrays follow the beads on the copy stream, into device staging: the commit is device work only
   if (p.imtype >= 0) CK(cudaMemcpyAsync(G.d_small_up, stage_up, (size_t)count * 2 * nang * sizeof(double), cudaMemcpyHostToDevice, G.copy_stream));
   CK(cudaMemcpyAsync(G.d_perm_up, stage_perm, (size_t)count * per * sizeof(int), cudaMemcpyHostToDevice, G.copy_stream));
   CK(cudaEventRecord(G.ev_up, G.copy_stream));
   G.up_first = first; G.up_count = count;
   return 0;
}

int pimcgpu_upload_states_commit(void)
{
   if (!G.live) return fail("pimcgpu_upload_states_commit: not initialised");
   if (!G.up_count) return fail("pimcgpu_upload_states_commit: no upload in flight");
   const Params &p = G.p;
   const int first = G.up_first, count = G.up_count;
   const size_t npos = (size_t)p.P * 3 * p.Npad, nang = (size_t)std::max(1, p.Q) * 3 * p.NMpad, per = 4 * (size_t)p.N + 1 + MAXT;
   int *stage_perm = G.stage_perm + (size_t)G.up_slot * p.nchains * per;
   CK(cudaStreamWaitEvent(G.stream, G.ev_up, 0));
   state_transpose_kernel<<<dim3((p.P + 31) / 32, (p.N + 31) / 32, 3 * count), dim3(32, 8), 0, G.stream>>>(p.pos + (size_t)first * npos, G.d_raw_up, p.N, p.P, p.Npad, 1);
   CK(cudaGetLastError());
   commit_small_kernel<<<64, 256, 0, G.stream>>>(p, first, count, G.d_small_up, G.d_perm_up, p.imtype >= 0 ? 1 : 0);
   CK(cudaGetLastError());
   CK(cudaEventRecord(G.ev_commit, G.stream));
   for (int cc = 0; cc < count; cc++) std::copy(stage_perm + (size_t)cc * per, stage_perm + (size_t)cc * per + p.N, G.h_pindex.begin() + (size_t)(first + cc) * p.N);
   G.up_count = 0;
   return 0;
}

int pimcgpu_download_states_begin(int first, int count, double *coords, double *angles, double *cosine)
{
   if (!G.live) return fail("pimcgpu_download_states_begin: not initialised");
   const Params &p = G.p;
   if (first < 0 || count < 1 || first + count > p.nchains) return fail("pimcgpu_download_states_begin: chains %d..%d out of range", first, first + count - 1);
   if (G.down_count) return fail("pimcgpu_download_states_begin: a download is already in flight (call pimcgpu_download_states_end)");
   if (!coords) return fail("pimcgpu_download_states_begin: coords is required");
   if (split_phase_setup()) return 1;
   const size_t n = (size_t)p.N * p.P, npos = (size_t)p.P * 3 * p.Npad, nang = (size_t)std::max(1, p.Q) * 3 * p.NMpad;
   // the copy stream may still be reading the previous snapshot: order the new one behind it
   CK(cudaEventRecord(G.ev_done, G.copy_stream));
   CK(cudaStreamWaitEvent(G.stream, G.ev_done, 0));
   state_transpose_kernel<<<dim3((p.P + 31) / 32, (p.N + 31) / 32, 3 * count), dim3(32, 8), 0, G.stream>>>(p.pos + (size_t)first * npos, G.d_raw_all, p.N, p.P, p.Npad, 0);
   CK(cudaGetLastError());
   double *s0 = G.stage + (size_t)first * G.stage_chain + npos;
   const bool rows = (angles || cosine) && p.imtype >= 0 && p.Q > 0;
   if (rows) {
      snapshot_small_kernel<<<64, 256, 0, G.stream>>>(p, first, count, G.d_small_down);
      CK(cudaGetLastError());
   }
   CK(cudaEventRecord(G.ev_down, G.stream));
   CK(cudaStreamWaitEvent(G.copy_stream, G.ev_down, 0));
   // the rotor rows first (small), then the beads: both on the copy stream, nothing of it on the library's stream
   if (rows) CK(cudaMemcpy2DAsync(s0, G.stage_chain * sizeof(double), G.d_small_down, 2 * nang * sizeof(double), 2 * nang * sizeof(double), count, cudaMemcpyDeviceToHost, G.copy_stream));
   CK(cudaMemcpyAsync(coords, G.d_raw_all, (size_t)count * 3 * n * sizeof(double), cudaMemcpyDeviceToHost, G.copy_stream));
   G.down_first = first; G.down_count = count; G.down_angles = angles; G.down_cosine = cosine;
   return 0;
}

int pimcgpu_download_states_end(void)
{
   if (!G.live) return fail("pimcgpu_download_states_end: not initialised");
   if (!G.down_count) return fail("pimcgpu_download_states_end: no download in flight");
   const Params &p = G.p;
   const size_t n = (size_t)p.N * p.P, npos = (size_t)p.P * 3 * p.Npad, nang = (size_t)std::max(1, p.Q) * 3 * p.NMpad;
   CK(cudaStreamSynchronize(G.copy_stream));        // the rotor rows are in the pinned staging area, the beads in the caller's array
   double *angles = G.down_angles, *cosine = G.down_cosine;
   if ((angles || cosine) && p.imtype >= 0 && p.Q > 0)
      for (int cc = 0; cc < G.down_count; cc++) {
         const double *hang = G.stage + (size_t)(G.down_first + cc) * G.stage_chain + npos, *hcos = hang + nang;
         double *ang = angles ? angles + (size_t)cc * 3 * n : nullptr, *cs = cosine ? cosine + (size_t)cc * 3 * n : nullptr;
         for (int q = 0; q < p.Q; q++)
            for (int m = 0; m < p.NM; m++) {
               const size_t dst = (size_t)(p.first[p.imtype] + m) * p.P + q, b = (size_t)q * 3 * p.NMpad + m;
               for (int d = 0; d < 3; d++) {
                  if (ang) ang[d * n + dst] = hang[b + d * p.NMpad];
                  if (cs) cs[d * n + dst] = hcos[b + d * p.NMpad];
               }
            }
      }
   G.down_count = 0;
   return 0;
}

int pimcgpu_seed(const unsigned long seed6[6])
{
   if (!G.live) return fail("pimcgpu_seed: not initialised");
   const Params &p = G.p;
   u64 seed[6];
   for (int i = 0; i < 6; i++) seed[i] = seed6[i];
   // CheckSeed, rngstream.cc:200-236
   for (int i = 0; i < 3; i++) if (seed[i] >= M1) return fail("pimcgpu_seed: seed[%d] >= 4294967087", i);
   for (int i = 3; i < 6; i++) if (seed[i] >= M2) return fail("pimcgpu_seed: seed[%d] >= 4294944443", i);
   if (!(seed[0] | seed[1] | seed[2]) || !(seed[3] | seed[4] | seed[5])) return fail("pimcgpu_seed: a seed triple is all zero");
   std::vector<uint32_t> h((size_t)p.nchains * p.S * 6);
   u64 st[6];
   stream_state(seed, (u64)G.sys.chain_offset * (u64)p.S, st);
   for (size_t s = 0; s < (size_t)p.nchains * p.S; s++) {
      for (int k = 0; k < 6; k++) h[s * 6 + k] = (uint32_t)st[k];
      matvec(A1P127, st, st, M1);
      matvec(A2P127, st + 3, st + 3, M2);
   }
   CK(cudaMemcpy(p.rng, h.data(), h.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
   G.seeded = true;
   return 0;
}

static void launch_config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr)
{
   cfg = cudaLaunchConfig_t{};
   cfg.gridDim = dim3(G.p.nchains * G.p.cpc);
   cfg.blockDim = dim3(G.threads);
   cfg.dynamicSmemBytes = G.smem;
   cfg.stream = G.stream;
   if (G.p.swbar) {
      attr[0].id = cudaLaunchAttributeCooperative;           // all CTAs co-resident: the software chain barrier spins
      attr[0].val.cooperative = 1;
   } else {
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = G.p.cpc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
   }
   cfg.attrs = attr; cfg.numAttrs = 1;
}

int pimcgpu_steps(long nsteps)
{
   if (!G.live) return fail("pimcgpu_steps: not initialised");
   if (!G.seeded) return fail("pimcgpu_steps: call pimcgpu_seed first");
   if (nsteps <= 0) return 0;
   cudaLaunchConfig_t cfg;
   cudaLaunchAttribute attr[1];
   launch_config(cfg, attr);
   CK(cudaMemsetAsync(G.p.barrier, 0, (size_t)G.p.nchains * 32 * sizeof(unsigned), G.stream));
   if (G.p.rot_run) CK(cudaMemsetAsync(G.p.rot_ll, 0, (size_t)G.p.nchains * G.p.Q * 8 * sizeof(unsigned long long), G.stream));
   void *args[4] = {(void *)&G.p, (void *)&G.step, (void *)&nsteps, (void *)&G.d_err};
   CK(cudaLaunchKernelExC(&cfg, steps_kernel(G.kind, G.p.worm_on, G.p.rot_run_cta), args));
   G.step += nsteps;
   return 0;
}

int pimcgpu_sync(void)
{
   if (!G.live) return fail("pimcgpu_sync: not initialised");
   CK(cudaStreamSynchronize(G.stream));
   int err = 0;
   CK(cudaMemcpy(&err, G.d_err, sizeof err, cudaMemcpyDeviceToHost));
   if (err & 1) return fail("rotational density table index out of range (the reference prints 'large matrix test error' and exits)");
   if (err & 2) return fail("Rotational Moves: Negative rot density");
   return 0;
}

long pimcgpu_step_counter(void) { return G.step; }

int pimcgpu_geometry(int *out8)
{
   if (!G.live) return fail("pimcgpu_geometry: not initialised");
   out8[0] = G.p.cpc; out8[1] = G.threads; out8[2] = G.p.team; out8[3] = G.p.rot_group; out8[4] = (int)G.smem; out8[5] = G.kind + (G.p.rot_run_cta ? 8 : 0);      // the move kernel's variant (template argument without the worm bit)
   cudaLaunchConfig_t cfg;
   cudaLaunchAttribute attr[1];
   launch_config(cfg, attr);
   int nclusters = -1;
   const void *kfun = steps_kernel(G.kind, G.p.worm_on, G.p.rot_run_cta);
   if (G.p.swbar) {
      int per_sm = 0, dev = 0;
      cudaDeviceProp prop;
      cudaGetDevice(&dev);
      if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfun, G.threads, G.smem) == cudaSuccess)
         nclusters = per_sm * prop.multiProcessorCount / G.p.cpc;          // chains that can be co-resident
   } else if (cudaOccupancyMaxActiveClusters(&nclusters, kfun, &cfg) != cudaSuccess) { nclusters = -1; cudaGetLastError(); }
   out8[6] = nclusters; out8[7] = G.p.nchains;
   return 0;
}
void *pimcgpu_stream(void) { return (void *)G.stream; }

int pimcgpu_measure(void)
{
   if (!G.live) return fail("pimcgpu_measure: not initialised");
   if (launch_estimators(1, 1)) return 1;
   if (G.p.refl[0] | G.p.refl[1] | G.p.refl[2] | G.p.rotsym) return pimcgpu_symmetry_moves();
   return 0;
}

int pimcgpu_symmetry_moves(void)
{
   if (!G.live) return fail("pimcgpu_symmetry_moves: not initialised");
   if (!G.seeded) return fail("pimcgpu_symmetry_moves: call pimcgpu_seed first");
   if (!(G.p.refl[0] | G.p.refl[1] | G.p.refl[2] | G.p.rotsym)) return 0;
   symmetry_kernel<<<G.p.nchains, 256, 0, G.stream>>>(G.p, nullptr);
   CK(cudaGetLastError());
   return 0;
}

int pimcgpu_symmetry_ops(const int *ops)
{
   if (!G.live) return fail("pimcgpu_symmetry_ops: not initialised");
   if (!ops) return fail("pimcgpu_symmetry_ops: null argument");
   if (!(G.p.imtype >= 0 && G.p.Q > 0)) return fail("pimcgpu_symmetry_ops: no rotor with ROTATION");
   for (int c = 0; c < G.p.nchains; c++) {
      if ((ops[4 * c] | ops[4 * c + 1] | ops[4 * c + 2]) && G.p.molecule[G.p.imtype] != 2) return fail("pimcgpu_symmetry_ops: reflections need a NONLINEAR rotor");
      if (ops[4 * c + 3] >= G.p.NM) return fail("pimcgpu_symmetry_ops: rotor index out of range");
   }
   CK(cudaMemcpyAsync(G.d_ops, ops, (size_t)G.p.nchains * 4 * sizeof(int), cudaMemcpyHostToDevice, G.stream));
   symmetry_kernel<<<G.p.nchains, 256, 0, G.stream>>>(G.p, G.d_ops);
   CK(cudaGetLastError());
   CK(cudaStreamSynchronize(G.stream));          // `ops` is pageable host memory
   return 0;
}

int pimcgpu_worm_moves(void)
{
   if (!G.live) return fail("pimcgpu_worm_moves: not initialised");
   if (!G.p.worm_on) return fail("pimcgpu_worm_moves: the system has no WORM");
   if (!G.seeded) return fail("pimcgpu_worm_moves: call pimcgpu_seed first");
   const size_t smem = 64 * sizeof(double) + worm_scratch_bytes(G.p.N);
   if (G.kind == 2) worm_move_kernel<6><<<G.p.nchains, 256, smem, G.stream>>>(G.p);
   else if (G.kind == 1) worm_move_kernel<5><<<G.p.nchains, 256, smem, G.stream>>>(G.p);
   else worm_move_kernel<4><<<G.p.nchains, 256, smem, G.stream>>>(G.p);
   CK(cudaGetLastError());
   return 0;
}
int pimcgpu_worm_state(int chain, int *st5)
{
   if (!G.live) return fail("pimcgpu_worm_state: not initialised");
   if (chain < 0 || chain >= G.p.nchains) return fail("pimcgpu_worm_state: chain out of range");
   CK(cudaStreamSynchronize(G.stream));
   CK(cudaMemcpy(st5, G.p.wstate + (size_t)chain * 8, 5 * sizeof(int), cudaMemcpyDeviceToHost));
   return 0;
}
int pimcgpu_worm_set(int chain, const int *st5)
{
   if (!G.live) return fail("pimcgpu_worm_set: not initialised");
   if (!G.p.worm_on) return fail("pimcgpu_worm_set: the system has no WORM");
   if (chain < 0 || chain >= G.p.nchains) return fail("pimcgpu_worm_set: chain out of range");
   const int nb = G.p.numb[G.p.worm_type];
   if (st5[1] < 0 || st5[1] >= G.p.P || st5[2] < 0 || st5[2] >= G.p.P || st5[3] < 0 || st5[3] >= nb || st5[4] < 0 || st5[4] >= nb)
      return fail("pimcgpu_worm_set: worm end points out of range");
   CK(cudaStreamSynchronize(G.stream));
   CK(cudaMemcpy(G.p.wstate + (size_t)chain * 8, st5, 5 * sizeof(int), cudaMemcpyHostToDevice));
   CK(cudaMemset(G.p.vepoch + (size_t)chain * std::max(1, G.p.Q) * G.p.NMpad, 0xff, (size_t)std::max(1, G.p.Q) * G.p.NMpad * sizeof(int)));
   return 0;
}
int pimcgpu_worm_counters(double *total7, double *accep7, double *countqw)
{
   if (!G.live) return fail("pimcgpu_worm_counters: not initialised");
   std::vector<double> h((size_t)G.p.nchains * 16);
   CK(cudaStreamSynchronize(G.stream));
   CK(cudaMemcpy(h.data(), G.p.qwc, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
   for (int i = 0; i < 7; i++) { total7[i] = 0.0; accep7[i] = 0.0; }
   *countqw = 0.0;
   for (int c = 0; c < G.p.nchains; c++) {
      for (int i = 0; i < 7; i++) { total7[i] += h[(size_t)c * 16 + i]; accep7[i] += h[(size_t)c * 16 + 7 + i]; }
      *countqw += h[(size_t)c * 16 + 14];
   }
   return 0;
}

// ---- checkpoint (N4) ----
namespace {
struct CkHeader { char magic[8]; int version, N, P, Q, nchains, S, Npad, NMpad, ntypes, worm_on; long step; long chain_offset; };
struct CkItem { void *dev; size_t bytes; };
std::vector<CkItem> checkpoint_items()
{
   const Params &p = G.p;
   const size_t C = p.nchains, Qx = std::max(1, p.Q);
   return {
      {p.pos, C * p.P * 3 * p.Npad * sizeof(double)}, {p.ang, C * Qx * 3 * p.NMpad * sizeof(double)}, {p.cosn, C * Qx * 3 * p.NMpad * sizeof(double)},
      {p.pindex, C * p.N * sizeof(int)}, {p.rindex, C * p.N * sizeof(int)}, {p.cyc_start, C * (p.N + 1) * sizeof(int)}, {p.cyc_atoms, C * p.N * sizeof(int)},
      {p.ncyc, C * MAXT * sizeof(int)}, {p.wstate, C * 8 * sizeof(int)}, {p.rng, C * p.S * 6 * sizeof(uint32_t)},
      {p.vold, C * Qx * p.NMpad * sizeof(double)}, {p.vepoch, C * Qx * p.NMpad * sizeof(int)}, {p.pos_epoch, C * sizeof(int)},
   };
}
}
long pimcgpu_checkpoint_bytes(void)
{
   if (!G.live) return -1;
   size_t n = sizeof(CkHeader);
   for (const CkItem &it : checkpoint_items()) n += it.bytes;
   return (long)n;
}
int pimcgpu_checkpoint_save(void *buf, long nbytes)
{
   if (!G.live) return fail("pimcgpu_checkpoint_save: not initialised");
   if (!buf || nbytes < pimcgpu_checkpoint_bytes()) return fail("pimcgpu_checkpoint_save: buffer too small");
   CK(cudaStreamSynchronize(G.stream));
   CkHeader h;
   memset(&h, 0, sizeof h);
   memcpy(h.magic, "PIMCB200", 8);
   h.version = 1; h.N = G.p.N; h.P = G.p.P; h.Q = G.p.Q; h.nchains = G.p.nchains; h.S = G.p.S; h.Npad = G.p.Npad; h.NMpad = G.p.NMpad;
   h.ntypes = G.p.ntypes; h.worm_on = G.p.worm_on; h.step = G.step; h.chain_offset = G.sys.chain_offset;
   char *q = (char *)buf;
   memcpy(q, &h, sizeof h); q += sizeof h;
   for (const CkItem &it : checkpoint_items()) { CK(cudaMemcpy(q, it.dev, it.bytes, cudaMemcpyDeviceToHost)); q += it.bytes; }
   return 0;
}
int pimcgpu_checkpoint_load(const void *buf, long nbytes)
{
   if (!G.live) return fail("pimcgpu_checkpoint_load: not initialised");
   if (!buf || nbytes < (long)sizeof(CkHeader)) return fail("pimcgpu_checkpoint_load: truncated checkpoint");
   CkHeader h;
   memcpy(&h, buf, sizeof h);
   if (memcmp(h.magic, "PIMCB200", 8) || h.version != 1) return fail("pimcgpu_checkpoint_load: not a pimcgpu checkpoint (version 1)");
   if (h.N != G.p.N || h.P != G.p.P || h.Q != G.p.Q || h.nchains != G.p.nchains || h.S != G.p.S || h.Npad != G.p.Npad || h.NMpad != G.p.NMpad ||
       h.ntypes != G.p.ntypes || h.worm_on != G.p.worm_on || h.chain_offset != G.sys.chain_offset)
      return fail("pimcgpu_checkpoint_load: the checkpoint belongs to a different system (N %d P %d Q %d chains %d offset %ld)", h.N, h.P, h.Q, h.nchains, h.chain_offset);
   if (nbytes < pimcgpu_checkpoint_bytes()) return fail("pimcgpu_checkpoint_load: truncated checkpoint");
   CK(cudaStreamSynchronize(G.stream));
   const char *q = (const char *)buf + sizeof h;
   for (const CkItem &it : checkpoint_items()) { CK(cudaMemcpy(it.dev, q, it.bytes, cudaMemcpyHostToDevice)); q += it.bytes; }
   CK(cudaMemcpy(G.h_pindex.data(), G.p.pindex, (size_t)G.p.nchains * G.p.N * sizeof(int), cudaMemcpyDeviceToHost));
   G.step = h.step;
   G.seeded = true;
   return 0;
}

int pimcgpu_chain_areas(int chain, double *out28)
{
   if (!G.live) return fail("pimcgpu_chain_areas: not initialised");
   if (chain < 0 || chain >= G.p.nchains) return fail("pimcgpu_chain_areas: chain out of range");
   if (G.p.bstype < 0) return fail("pimcgpu_chain_areas: no BOSE type in the system");
   if (launch_estimators(0, 0)) return 1;
   CK(cudaStreamSynchronize(G.stream));
   CK(cudaMemcpy(out28, G.e.chain_area + (size_t)chain * NAREA, NAREA * sizeof(double), cudaMemcpyDeviceToHost));
   return 0;
}

long pimcgpu_accum_offset(const char *name)
{
   if (!G.live || !name) return -1;
   const std::string n(name);
   if (n == "scalars") return 0;
   if (n == "gr1d") return G.e.off_gr1d;
   if (n == "gr2d") return G.e.off_gr2d;
   if (n == "gr3d") return G.e.has_gr3d ? G.e.off_gr3d : -1;
   if (n == "rcf") return G.e.off_rcf;
   if (n == "relbins") return G.e.off_relbins;
   if (n == "area") return G.e.off_area;
   if (n == "ploops") return G.e.off_ploops;
   if (n == "rcfcnt") return G.e.off_rcfcnt;
   return -1;
}

int pimcgpu_chain_energies(int chain, double *out5)
{
   if (!G.live) return fail("pimcgpu_chain_energies: not initialised");
   if (chain < 0 || chain >= G.p.nchains) return fail("pimcgpu_chain_energies: chain out of range");
   if (launch_estimators(0, 0)) return 1;
   CK(cudaStreamSynchronize(G.stream));
   CK(cudaMemcpy(out5, G.e.chain_e + (size_t)chain * 8, 5 * sizeof(double), cudaMemcpyDeviceToHost));
   return 0;
}

int pimcgpu_chain_rcf(int chain, double *rcf0)
{
   if (!G.live) return fail("pimcgpu_chain_rcf: not initialised");
   if (chain < 0 || chain >= G.p.nchains || G.p.Q <= 0) return fail("pimcgpu_chain_rcf: bad chain or no rotation");
   if (launch_estimators(0, 0)) return 1;
   CK(cudaStreamSynchronize(G.stream));
   CK(cudaMemcpy(rcf0, G.e.chain_rcf + (size_t)chain * 2 * G.p.Q, G.p.Q * sizeof(double), cudaMemcpyDeviceToHost));
   return 0;
}

int pimcgpu_accum_layout(long *n_total, long *off_scalars, long *off_gr1d, long *off_gr2d, long *off_gr3d, long *off_rcf, long *off_relbins)
{
   if (!G.live) return fail("pimcgpu_accum_layout: not initialised");
   if (n_total) *n_total = G.nacc;
   if (off_scalars) *off_scalars = 0;
   if (off_gr1d) *off_gr1d = G.e.off_gr1d;
   if (off_gr2d) *off_gr2d = G.e.off_gr2d;
   if (off_gr3d) *off_gr3d = G.e.has_gr3d ? G.e.off_gr3d : -1;
   if (off_rcf) *off_rcf = G.e.off_rcf;
   if (off_relbins) *off_relbins = G.e.off_relbins;
   return 0;
}
void *pimcgpu_accum_device_ptr(void)
{
   if (!G.live) return nullptr;
   fold_counters_kernel<<<1, 32, 0, G.stream>>>(G.p, G.e.acc);      // ordered on the library's stream: no host round trip here
   return G.e.acc;
}
int pimcgpu_accum_download(double *host, long n)
{
   if (!G.live) return fail("pimcgpu_accum_download: not initialised");
   if (n > G.nacc) n = G.nacc;
   CK(cudaMemcpyAsync(host, G.e.acc, n * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
   CK(cudaStreamSynchronize(G.stream));
   return 0;
}
// split form: the copy is queued behind what the stream holds now (the estimators, an all-reduce the caller ordered before
// it), so a large transfer started afterwards -- pimcgpu_download_states_begin -- does not delay the small result
int pimcgpu_accum_download_begin(double *host, long n)
{
   if (!G.live) return fail("pimcgpu_accum_download_begin: not initialised");
   if (n > G.nacc) n = G.nacc;
   if (!G.ev_acc) CK(cudaEventCreateWithFlags(&G.ev_acc, cudaEventDisableTiming));
   CK(cudaMemcpyAsync(host, G.e.acc, n * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
   CK(cudaEventRecord(G.ev_acc, G.stream));
   return 0;
}
int pimcgpu_accum_download_end(void)
{
   if (!G.live) return fail("pimcgpu_accum_download_end: not initialised");
   if (!G.ev_acc) return fail("pimcgpu_accum_download_end: no download in flight");
   CK(cudaEventSynchronize(G.ev_acc));
   return 0;
}
int pimcgpu_accum_reset(void)
{
   if (!G.live) return fail("pimcgpu_accum_reset: not initialised");
   CK(cudaMemsetAsync(G.e.acc, 0, G.nacc * sizeof(double), G.stream));
   CK(cudaMemsetAsync(G.p.counters, 0, (size_t)G.p.nchains * MAXT * 3 * 2 * sizeof(double), G.stream));
   if (G.p.worm_on) {                     // ResetQWCounts, mc_main.cc:531
      std::vector<double> q0((size_t)G.p.nchains * 16, 0.0);
      for (int c = 0; c < G.p.nchains; c++) q0[(size_t)c * 16 + 14] = 1.0;
      CK(cudaStreamSynchronize(G.stream));
      CK(cudaMemcpy(G.p.qwc, q0.data(), q0.size() * sizeof(double), cudaMemcpyHostToDevice));
   }
   return 0;
}
int pimcgpu_block_scalars(pimcgpu_scalars *out)
{
   if (!G.live) return fail("pimcgpu_block_scalars: not initialised");
   double h[32];
   CK(cudaStreamSynchronize(G.stream));
   CK(cudaMemcpy(h, G.e.acc, sizeof h, cudaMemcpyDeviceToHost));
   out->count = h[0]; out->kin = h[1]; out->pot = h[2]; out->rot = h[3]; out->rotsq = h[4];
   out->cv = h[5]; out->cv_trans = h[6]; out->cv_rot = h[7];
   for (int t = 0; t < MAXT; t++)
      for (int m = 0; m < 3; m++) { out->mctotal[t][m] = h[8 + t * 3 + m]; out->mcaccep[t][m] = h[14 + t * 3 + m]; }
   return 0;
}
int pimcgpu_counters(double *mctotal, double *mcaccep)
{
   if (!G.live) return fail("pimcgpu_counters: not initialised");
   const Params &p = G.p;
   std::vector<double> h((size_t)p.nchains * MAXT * 3 * 2);
   CK(cudaStreamSynchronize(G.stream));
   CK(cudaMemcpy(h.data(), p.counters, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
   for (int i = 0; i < MAXT * 3; i++) { mctotal[i] = 0.0; mcaccep[i] = 0.0; }
   for (int c = 0; c < p.nchains; c++)
      for (int i = 0; i < MAXT * 3; i++) { mctotal[i] += h[((size_t)c * MAXT * 3 + i) * 2]; mcaccep[i] += h[((size_t)c * MAXT * 3 + i) * 2 + 1]; }
   return 0;
}

} // extern "C"

// ---- parity entry points ---------------------------------------------------------------------------
namespace {
struct DevBuf {
   std::vector<void *> ptrs;
   ~DevBuf() { for (void *q : ptrs) cudaFree(q); }
   template <class T> T *in(const T *h, size_t n) { T *d = nullptr; cudaMalloc(&d, std::max<size_t>(1, n) * sizeof(T)); cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice); ptrs.push_back(d); return d; }
   template <class T> T *out(size_t n) { T *d = nullptr; cudaMalloc(&d, std::max<size_t>(1, n) * sizeof(T)); ptrs.push_back(d); return d; }
};
template <class T> int back(T *h, const T *d, size_t n)
{
   if (!h) return 0;
   cudaError_t e = cudaMemcpy(h, d, n * sizeof(T), cudaMemcpyDeviceToHost);
   if (e != cudaSuccess) return fail("device evaluation failed: %s", cudaGetErrorString(e));
   return 0;
}
}

extern "C" {

int pimcgpu_eval_spot1d(int n, const double *r, double *v, int *klo)
{
   if (!G.live || !G.p.n1d) return fail("pimcgpu_eval_spot1d: no 1-D table loaded");
   DevBuf b; const double *dr = b.in(r, n); double *dv = b.out<double>(n); int *dk = b.out<int>(n);
   eval_spot1d_kernel<<<(n + 127) / 128, 128>>>(G.p, n, dr, dv, dk);
   return back(v, dv, n) || back(klo, dk, n);
}
int pimcgpu_eval_lpot2d(int n, const double *r, const double *c, double *v, int *ir, int *ic)
{
   if (!G.live || !G.p.rs2d) return fail("pimcgpu_eval_lpot2d: no 2-D table loaded");
   DevBuf b; const double *dr = b.in(r, n), *dc = b.in(c, n); double *dv = b.out<double>(n); int *di = b.out<int>(n), *dj = b.out<int>(n);
   eval_lpot2d_kernel<<<(n + 127) / 128, 128>>>(G.p, n, dr, dc, dv, di, dj);
   return back(v, dv, n) || back(ir, di, n) || back(ic, dj, n);
}
int pimcgpu_eval_srotdens(int n, const double *g, int which, double *v)
{
   if (!G.live || !G.p.nrot) return fail("pimcgpu_eval_srotdens: no linear-rotor density table loaded");
   DevBuf b; const double *dg = b.in(g, n); double *dv = b.out<double>(n);
   eval_srot_kernel<<<(n + 127) / 128, 128>>>(G.p, n, dg, which, dv);
   return back(v, dv, n);
}
int pimcgpu_eval_rotden(int n, const double *e1, const double *e2, double *rho, double *erot, double *esq, int *index)
{
   if (!G.live || !G.p.rho3) return fail("pimcgpu_eval_rotden: no top density-matrix tables loaded");
   DevBuf b; const double *d1 = b.in(e1, 3 * (size_t)n), *d2 = b.in(e2, 3 * (size_t)n);
   double *dr = b.out<double>(n), *de = b.out<double>(n), *dq = b.out<double>(n); int *di = b.out<int>(n);
   eval_rotden_kernel<<<(n + 127) / 128, 128>>>(G.p, n, d1, d2, dr, de, dq, di);
   return back(rho, dr, n) || back(erot, de, n) || back(esq, dq, n) || back(index, di, n);
}
int pimcgpu_eval_vcord(int n, const double *eul, const double *rcom, const double *rpt, double *v, double *rtc, int *index)
{
   if (!G.live || !G.p.v3d) return fail("pimcgpu_eval_vcord: no 3-D potential table loaded");
   DevBuf b; const double *de = b.in(eul, 3 * (size_t)n), *dc = b.in(rcom, 3 * (size_t)n), *dp = b.in(rpt, 3 * (size_t)n);
   double *dv = b.out<double>(n), *dt = b.out<double>(3 * (size_t)n); int *di = b.out<int>(n);
   eval_vcord_kernel<<<(n + 127) / 128, 128>>>(G.p, n, de, dc, dp, dv, dt, di);
   return back(v, dv, n) || back(rtc, dt, 3 * (size_t)n) || back(index, di, n);
}
int pimcgpu_eval_caleng(int n, const double *c1, const double *c2, const double *e1, const double *e2, double *e)
{
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("pimcgpu_eval_caleng: no CUDA device");
   DevBuf b; const double *a1 = b.in(c1, 3 * (size_t)n), *a2 = b.in(c2, 3 * (size_t)n), *b1 = b.in(e1, 3 * (size_t)n), *b2 = b.in(e2, 3 * (size_t)n);
   double *de = b.out<double>(n);
   eval_caleng_kernel<<<(n + 127) / 128, 128>>>(n, a1, a2, b1, b2, de);
   return back(e, de, n);
}
int pimcgpu_eval_rotpro(int n, const double *deg, double *rho, double *erot, double *esq, int *index)
{
   if (!G.live || !G.p.rho3) return fail("pimcgpu_eval_rotpro: no top density-matrix tables loaded");
   DevBuf b; const double *dd = b.in(deg, 3 * (size_t)n);
   double *dr = b.out<double>(n), *de = b.out<double>(n), *dq = b.out<double>(n); int *di = b.out<int>(n);
   eval_rotpro_kernel<<<(n + 127) / 128, 128>>>(G.p, n, dd, dr, de, dq, di);
   return back(rho, dr, n) || back(erot, de, n) || back(esq, dq, n) || back(index, di, n);
}
int pimcgpu_eval_vcalc(int n, const double *rtc, double *v, int *index)
{
   if (!G.live || !G.p.v3d) return fail("pimcgpu_eval_vcalc: no 3-D potential table loaded");
   DevBuf b; const double *dt = b.in(rtc, 3 * (size_t)n); double *dv = b.out<double>(n); int *di = b.out<int>(n);
   eval_vcalc_kernel<<<(n + 127) / 128, 128>>>(G.p, n, dt, dv, di);
   return back(v, dv, n) || back(index, di, n);
}
int pimcgpu_eval_deleul(int n, const double *e1, const double *e2, double *rel)
{
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("pimcgpu_eval_deleul: no CUDA device");
   DevBuf b; const double *d1 = b.in(e1, 3 * (size_t)n), *d2 = b.in(e2, 3 * (size_t)n); double *dr = b.out<double>(3 * (size_t)n);
   eval_deleul_kernel<<<(n + 127) / 128, 128>>>(n, d1, d2, dr);
   return back(rel, dr, 3 * (size_t)n);
}
int pimcgpu_eval_vcord_grid(int n, const double *eul, const double *rcom, const double *rpt, double *grid)
{
   if (!G.live || !G.p.v3d) return fail("pimcgpu_eval_vcord_grid: no 3-D potential table loaded");
   DevBuf b; const double *de = b.in(eul, 3 * (size_t)n), *dc = b.in(rcom, 3 * (size_t)n), *dp = b.in(rpt, 3 * (size_t)n);
   double *dg = b.out<double>(3 * (size_t)n);
   eval_vcord_grid_kernel<<<(n + 127) / 128, 128>>>(G.p, n, de, dc, dp, dg);
   return back(grid, dg, 3 * (size_t)n);
}
int pimcgpu_eval_vspher(int n, const double *r, double *v, double *rclamp)
{
   if (!G.live || !G.p.vspher) return fail("pimcgpu_eval_vspher: no spherical table loaded (ISPHER = 0)");
   DevBuf b; const double *dr = b.in(r, n); double *dv = b.out<double>(n), *dc = b.out<double>(n);
   eval_vspher_kernel<<<(n + 127) / 128, 128>>>(G.p, n, dr, dv, dc);
   return back(v, dv, n) || back(rclamp, dc, n);
}
int pimcgpu_eval_libm(int which, int n, const double *x, double *y)
{
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("pimcgpu_eval_libm: no CUDA device");
   if (which < 0 || which > 7) return fail("pimcgpu_eval_libm: unknown function %d", which);
   DevBuf b; const double *dx = b.in(x, n); double *dy = b.out<double>(n);
   eval_libm_kernel<<<(n + 127) / 128, 128>>>(which, n, dx, dy);
   return back(y, dy, n);
}
int pimcgpu_pot_energy_slice(int chain, double *v)
{
   if (!G.live) return fail("pimcgpu_pot_energy_slice: not initialised");
   if (chain < 0 || chain >= G.p.nchains) return fail("pimcgpu_pot_energy_slice: chain out of range");
   size_t n = (size_t)G.p.N * G.p.P;
   DevBuf b; double *dv = b.out<double>(n);
   CK(cudaStreamSynchronize(G.stream));
   pot_energy_slice_kernel<<<(unsigned)((n * 32 + 255) / 256), 256>>>(G.p, chain, dv);
   return back(v, dv, n);
}
// ---- host-side table preparation, exposed for the CPU-only tests (no device needed) ----
int pimcgpu_host_spline(int n, const double *x, const double *y, double *y2, double *alpha_unode_c6)
{
   if (n < 2) return fail("pimcgpu_host_spline: need at least two points");
   std::vector<double> g(x, x + n), v(y, y + n), d2;
   spline_setup(g, v, d2);
   memcpy(y2, d2.data(), n * sizeof(double));
   if (alpha_unode_c6) {
      double alpha = log(v[0] / v[1]) / (g[1] - g[0]);
      alpha_unode_c6[0] = alpha;
      alpha_unode_c6[1] = v[0] * exp(alpha * g[0]);
      alpha_unode_c6[2] = (v[n - 1] - v[n - 2]) / (1.0 / pow(g[n - 2], 6.0) - 1.0 / pow(g[n - 1], 6.0));
   }
   return 0;
}
int pimcgpu_host_stream_state(const unsigned long seed6[6], long stream, unsigned long state6[6])
{
   u64 seed[6], st[6];
   for (int i = 0; i < 6; i++) seed[i] = seed6[i];
   stream_state(seed, (u64)stream, st);
   for (int i = 0; i < 6; i++) state6[i] = st[i];
   return 0;
}
int pimcgpu_host_lut(int n, const double *x, int *lut /* [4n] */, double *scale)
{
   std::vector<double> g(x, x + n);
   std::vector<int> l;
   build_lut(g, l, *scale);
   memcpy(lut, l.data(), l.size() * sizeof(int));
   return (int)l.size();
}

#ifdef PIMC_TIMELINE
int pimcgpu_timeline(long long *out, int n)
{
   int nm = 0;
   cudaDeviceSynchronize();
   cudaMemcpyFromSymbol(&nm, g_nmarks, sizeof nm);
   if (nm > n) nm = n;
   cudaMemcpyFromSymbol(out, g_marks, nm * sizeof(long long));
   int zero = 0;
   cudaMemcpyToSymbol(g_nmarks, &zero, sizeof zero);
   return nm;
}
#endif

int pimcgpu_fp64_peak(double *tflops)
{
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("pimcgpu_fp64_peak: no CUDA device");
   cudaDeviceProp prop;
   int dev = 0;
   CK(cudaGetDevice(&dev));
   CK(cudaGetDeviceProperties(&prop, dev));
   double *d = nullptr;
   CK(cudaMalloc(&d, 8));
   const int iters = 1 << 16, blocks = prop.multiProcessorCount * 8, threads = 256;
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   double best = 0.0;
   for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      fp64_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
      best = std::max(best, fl / (ms * 1e-3) / 1e12);
   }
   cudaEventDestroy(e0); cudaEventDestroy(e1);
   cudaFree(d);
   *tflops = best;
   return 0;
}
int pimcgpu_rng_draws(long stream, int n, double *out)
{
   if (!G.live || !G.seeded) return fail("pimcgpu_rng_draws: seed the library first");
   // not kept: uses the package seed recorded at pimcgpu_seed time through the device copy of local streams
   const Params &p = G.p;
   long local = stream - G.sys.chain_offset * (long)p.S;
   if (local < 0 || local >= (long)p.nchains * p.S) return fail("pimcgpu_rng_draws: stream %ld is not owned by this context", stream);
   uint32_t st[6];
   CK(cudaStreamSynchronize(G.stream));
   CK(cudaMemcpy(st, p.rng + (size_t)local * 6, sizeof st, cudaMemcpyDeviceToHost));
   DevBuf b; double *d = b.out<double>(n);
   rng_draws_kernel<<<1, 1>>>(st[0], st[1], st[2], st[3], st[4], st[5], n, d);
   return back(out, d, n);
}

} // extern "C"
