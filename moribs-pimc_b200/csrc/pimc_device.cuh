// Device-side data model, MRG32k3a streams and leaf potentials of the PIMC hot path (sm_100a).
//
// Layout in HBM (per chain c, all FP64):
//   pos  [c][it][d][Npad]      beads of one imaginary-time slice are contiguous (SoA by component,
//                              atoms fastest) so a warp sweeping partners j reads coalesced sectors
//   ang  [c][q][3][NMpad]      (phi, cos(theta), chi) of every rotor at rot slice q
//   cosn [c][q][3][NMpad]      unit axis vector n (MCCosine)
//   rng  [c][S][6] uint32      MRG32k3a states: S = P (one per translational slice) + Q (one per
//                              rotational slice) + 8 (per-chain miscellaneous)
// Tables: the 1-D spline (grid, V, V'') and the linear-rotor density spline are staged in shared
// memory by every CTA; the 2-D/3-D potential tables and the 181x361x361 rho/E/E^2 tables stay in
// global memory (L2-resident working set) and are gathered with the reference's index arithmetic.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace pimc {

constexpr int MAXT = 2;
constexpr double RZERO = 1.0e-10;                 // mc_confg.h:60
constexpr double PI = 3.14159265358979323846;     // rotden.f:6 (== M_PI in double)
constexpr double WNO2K = 0.6950356;               // rotden.f:6, mc_const.h:15

// interaction branch of PotEnergy for an (atom0 type, atom1 type) pair, mc_piqmc.cc:1847-1958
enum Mode : int { M_SPOT1D = 0, M_LIN_0MOL = 1, M_LIN_1MOL = 2, M_TOP_0MOL = 3, M_TOP_1MOL = 4, M_SPHER = 5, M_TOPTOP = 6 };

// one spline interval [xlo, xhi]: everything splint (mc_utils.cc:156-186) needs, with 1/h and y''h^2/6 folded in
struct alignas(16) SplineRec { double xlo, xhi, inv_h, ylo, yhi, clo, chi, pad; };      // 64 bytes: four 128-bit shared-memory loads

struct Params {
   int ntypes, N, P, Q, R, Npad, NM, NMpad;
   int numb[MAXT], molecule[MAXT], stat[MAXT], levels[MAXT], first[MAXT + 1];
   double lambda[MAXT], mcstep[MAXT], rtstep[MAXT], mass[MAXT];
   double tau, rottau, beta, temperature;
   int imtype, bstype, ispher, minimage;
   int rotden_type, rnratio;     // ROTDENSI: 0 tabulated densities, 1 rattle-and-shake propagator (rsrot/rsline); RS/Noya ratio
   double xrot, yrot, zrot;      // rotational constants of the ROTDENSI line, cm^-1
   int refl[3], rotsym, nfold;   // REFLECTX/Y/Z, ROTSYM (IREFLX.., IROTSYM, NFOLD_ROT; mc_setup.h:24-30)
   // worm (WORM line, MCWormInit mc_qworm.cc:48-82): type, m-tilde, C * density, 4 lambda tau, neighbour cutoff^2
   int worm_on, worm_type, worm_m;
   double worm_norm, worm_twave2, worm_cutoff2;
   const uint32_t *worm_jump;   // [k][18]: the MRG32k3a transition matrices A1^(2k), A2^(2k) (skip-ahead of the worm's gaussian batches)
   int *wstate;                  // [c][8]: Worm.exists, ira, masha, atom_i, atom_m (atoms numbered inside the worm type)
   int *rindex;                  // [c][N] previous world line (RIndex), global atom index like pindex
   double *qwc;                  // [c][16]: QWTotal[7], QWAccep[7], countQW
   double box[3];
   int mode[MAXT][MAXT];
   // tables
   int n1d, nlut1d;  const double *g1d, *v1d, *y2_1d; const int *lut1d; double alpha, unode, c6, lut1d_scale;
   int uniform1d; double x0_1d, xn_1d, invh_1d;      // 1-D grid end points; uniform grid: interval index = (r - x0)/h directly
   const SplineRec *rec1d;       // packed per-interval records of the 1-D potential spline
   int poly1d;                   // the 1-D grid is uniform to rounding: per-interval cubic in t = (r - x_k)/h, {c0, c1} and {c2, c3}
   const double2 *pa1d, *pb1d;   // in two arrays of 16-byte entries (two conflict-poor 128-bit shared-memory loads per evaluation)
   int rs2d, cs2d;   const double *rg2d, *cg2d, *v2d; double dr2d, dc2d;
   const double *irg2d, *icg2d;  // 1/(grid[i+1]-grid[i]) of the 2-D potential axes
   double inv_dr2d, inv_dc2d;
   const double2 *cell2d;        // [(rs-1)][cs] row pairs {V[ir][ic], V[ir+1][ic]}: a bilinear cell is two adjacent 16-byte
                                 // entries (32 contiguous bytes, 2x the table instead of 4x so the hot region stays in L2)
   const double2 *rgi2d, *cgi2d; // {grid[i], 1/(grid[i+1]-grid[i])} of the two axes
   int geo_hint;                 // L2 policy of the geometry records: 0 evict_first, 1 evict_last, 2 default
   int cell4_on, cell_hint;      // whole-cell table in use (else the row-pair table); the random gathers do not allocate in L1 (they never hit
                                 // there and displace the lines the other loads reuse)
   const double *cell4;          // [(rs-1)][(cs-1)][4] whole cells {V[ir][ic], V[ir+1][ic], V[ir][ic+1], V[ir+1][ic+1]}, 32-byte aligned
   int rg3, thg3, chg3; const double *v3d; double rvmin, rvmax, rvstep;
   int nrot, nlutrot; const double *rgrid, *rdens, *rderv, *resqr, *rdens2, *rderv2, *resqr2; const int *lutrot; double lutrot_scale;
   const SplineRec *recrot;      // packed records of the linear-rotor density spline (rho column)
   const double *rho3, *erot3, *esq3;
   const double *vspher;
   // state
   int nchains, S;
   double *pos, *ang, *cosn;
   int *pindex;                  // [c][N] next world line (global atom index; identity for non-bosons)
   int *cyc_start, *cyc_atoms;   // [c][N+1], [c][N]: permutation cycles (groups moved rigidly together)
   int *ncyc;                    // [c][MAXT]
   uint32_t *rng;
   double *counters;             // [c][MAXT][3][2] (total, accepted)
   double *scratch;              // [c][64] cross-CTA partial sums
   double *vold;                 // [c][Q][NMpad] cached sum of the rotor's potential over the R slices of rot slice q
   int *vepoch;                  // [c][Q][NMpad] value of pos_epoch the cache entry was computed at (-1: invalid)
   int *pos_epoch;               // [c] bumped by every translational sweep
   // execution geometry
   int cpc, team;
   int rot_group;                // threads cooperating on one rotational slice (power of two, may span warps)
   int seg_max;
   double *segbuf;               // [c][nseg_max][team_buf_n] per-segment scratch in global memory, used when the per-team
   int segbuf_global, nseg_max;  // shared-memory buffers would not fit (many narrow teams, long segments)
   int team_buf_n;               // doubles of scratch per bisection team (team_buf_doubles)
   int rot_in_smem;              // the linear-rotor density spline is staged in shared memory
   int swbar;                    // chain barrier in global memory (cooperative grid) instead of the cluster barrier
   unsigned *barrier;            // [c][32] arrival counters of the software chain barrier (zeroed before every launch)
   int rot_fused;                // one-rotor system, even Q: potential sums of all Q proposals in one stage, decisions pipelined (rot_sweep_pipe)
   // geometry cache of the rotor-atom terms of a linear rotor (rot_potential_cached): the positions only change in the
   // translational sweeps, so between two of them every (slice, partner) term keeps its unit vector (p_j - p_g)/r, its
   // radial cell row and its radial weight; a rotational proposal only changes cos(theta) = n.u
   int rot_run;                  // linear rotor, several CTAs per chain: free-running rotational sweeps (rot_run), contiguous slice blocks per CTA
   int rot_run_cta;              // a top whose chain lives in one CTA: free-running sweeps through shared memory (rot_run_cta)
   unsigned long long *rot_ll;   // [c][Q][8] rot_run: axis hand-over records (rot_ll_publish), zeroed at every launch
   int mol_piped;                // whole-path sweep with one cross-CTA hand-over per atom (molecular_sweep_piped)
   int bis_piped;                // two-warp teams: bisection sweep software-pipelined over the atoms (bisection_sweep_piped)
   int geo_on, geo_items, geo_n; // enabled; real items per rot slice = R x (N - 1); padded stride
   double *geo;                  // [c][q][geo_n][4]: one aligned 32-byte record {ux, uy, uz, w} per item; w = the radial weight dr with
                                 // the radial cell index ir (< 4096) in the 12 lowest mantissa bits (|error| <= 2^-40 dr)
};

__host__ __device__ inline size_t pos_index(const Params &p, int c, int it, int d, int a)
{
   return (((size_t)c * p.P + it) * 3 + d) * p.Npad + a;
}
__host__ __device__ inline size_t ang_index(const Params &p, int c, int q, int d, int m)
{
   return (((size_t)c * p.Q + q) * 3 + d) * p.NMpad + m;
}

// ---------------------------------------------------------------------------------------------
// MRG32k3a (L'Ecuyer), integer form of RngStream::U01 (rngstream.cc:242-265): bit-identical output
// ---------------------------------------------------------------------------------------------
constexpr uint32_t MRG_M1 = 4294967087u, MRG_M2 = 4294944443u;

struct Mrg { uint32_t s[6]; };

__device__ __forceinline__ double mrg_u01(Mrg &g)
{
   // component 1: p1 = (1403580*s11 - 810728*s10) mod m1, with 2^32 = 209 (mod m1)
   uint64_t x = 1403580ull * g.s[1] + 810728ull * (uint64_t)(MRG_M1 - g.s[0]);
   x = (x >> 32) * 209ull + (x & 0xffffffffull);
   x = (x >> 32) * 209ull + (x & 0xffffffffull);
   if (x >= MRG_M1) x -= MRG_M1;
   uint32_t p1 = (uint32_t)x;
   g.s[0] = g.s[1]; g.s[1] = g.s[2]; g.s[2] = p1;
   // component 2: p2 = (527612*s22 - 1370589*s20) mod m2, with 2^32 = 22853 (mod m2)
   uint64_t y = 527612ull * g.s[5] + 1370589ull * (uint64_t)(MRG_M2 - g.s[3]);
   y = (y >> 32) * 22853ull + (y & 0xffffffffull);
   y = (y >> 32) * 22853ull + (y & 0xffffffffull);
   if (y >= MRG_M2) y -= MRG_M2;
   uint32_t p2 = (uint32_t)y;
   g.s[3] = g.s[4]; g.s[4] = g.s[5]; g.s[5] = p2;
   const double norm = 1.0 / (4294967087.0 + 1.0);
   return (p1 > p2) ? (double)(p1 - p2) * norm : ((double)p1 - (double)p2 + 4294967087.0) * norm;
}
// skip-ahead: state <- (J1 s1 mod m1, J2 s2 mod m2) with the 3x3 matrices in J[0..8], J[9..17] (row major)
__device__ __forceinline__ void mrg_jump(Mrg &g, const uint32_t *J)
{
   uint32_t o[6];
   #pragma unroll
   for (int r = 0; r < 3; r++) {
      uint64_t a = 0, b = 0;
      #pragma unroll
      for (int c = 0; c < 3; c++) {
         a = (a + ((uint64_t)J[r * 3 + c] * g.s[c]) % MRG_M1) % MRG_M1;
         b = (b + ((uint64_t)J[9 + r * 3 + c] * g.s[3 + c]) % MRG_M2) % MRG_M2;
      }
      o[r] = (uint32_t)a; o[3 + r] = (uint32_t)b;
   }
   #pragma unroll
   for (int i = 0; i < 6; i++) g.s[i] = o[i];
}
__device__ __forceinline__ void mrg_load(Mrg &g, const uint32_t *st)
{
   #pragma unroll
   for (int i = 0; i < 6; i++) g.s[i] = st[i];
}
__device__ __forceinline__ void mrg_store(const Mrg &g, uint32_t *st)
{
   #pragma unroll
   for (int i = 0; i < 6; i++) st[i] = g.s[i];
}
// gauss(alpha) of mc_randg.cc:138-150: sqrt(-ln r1) cos(2 pi r2) / sqrt(alpha)
__device__ __forceinline__ double gauss_u(double alpha, double r1, double r2)
{
   return sqrt(-log(r1)) * cos(2.0 * PI * r2) / sqrt(alpha);
}

// ---------------------------------------------------------------------------------------------
// spline look-ups
// ---------------------------------------------------------------------------------------------
// interval index of the reference's bisection search in splint (mc_utils.cc:170-176):
// klo = max{k : xa[k] <= x}, clamped to [0, n-2]; bucket LUT guess corrected against the grid.
__device__ __forceinline__ int spline_klo(const double *xa, int n, const int *lut, int nlut, double scale, double x)
{
   int b = (int)((x - xa[0]) * scale);
   b = b < 0 ? 0 : (b >= nlut ? nlut - 1 : b);
   int k = lut[b];
   while (k < n - 2 && xa[k + 1] <= x) k++;
   while (k > 0 && xa[k] > x) k--;
   return k;
}
__device__ __forceinline__ double splint_eval(const double *xa, const double *ya, const double *y2a, int klo, double x)
{
   int khi = klo + 1;
   double h = xa[khi] - xa[klo];
   double a = (xa[khi] - x) / h;
   double b = (x - xa[klo]) / h;
   return a * ya[klo] + b * ya[khi] + ((a * a * a - a) * y2a[klo] + (b * b * b - b) * y2a[khi]) * (h * h) / 6.;
}

// shared-memory (or global) views of the small tables, set up once per CTA
struct SmallTables {
   const double *g1d, *v1d, *y2_1d; const int *lut1d;
   const double *rgrid, *rdens, *rdens2; const int *lutrot;
   const SplineRec *rec1d, *recrot;
   const double2 *rgi2d, *cgi2d;   // axis tables of the 2-D potential (shared memory in the move kernel)
   const double2 *pa1d, *pb1d;     // per-interval cubic of the 1-D potential (shared memory in the move kernel)
   // the same tables as 32-bit shared-window addresses, valid in the move kernel only (stage_tables): explicit ld.shared
   // instead of generic loads (which take the global-memory scoreboard and an address-window check each)
   uint32_t s_pa1d, s_pb1d, s_rgi2d, s_cgi2d;
};
__device__ __forceinline__ double2 lds_d2(uint32_t base, int i)
{
   double2 v;
   asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(base + 16u * (uint32_t)i));
   return v;
}

// interval search + cubic evaluation on the packed records: same interval as the reference's bisection search
// (klo = max{k : x_k <= x}); a, b use the stored 1/h, the curvature terms the stored y'' h^2/6
__device__ __forceinline__ double spline_rec_eval(const SplineRec *rec, int n, const int *lut, int nlut, double scale, double x0, double x, int *klo_out)
{
   int b = (int)((x - x0) * scale);
   b = b < 0 ? 0 : (b >= nlut ? nlut - 1 : b);
   int k = lut[b];
   while (k < n - 2 && rec[k].xhi <= x) k++;
   while (k > 0 && rec[k].xlo > x) k--;
   if (klo_out) *klo_out = k;
   const SplineRec r = rec[k];
   double a = (r.xhi - x) * r.inv_h;
   double bb = (x - r.xlo) * r.inv_h;
   return a * r.ylo + bb * r.yhi + ((a * a * a - a) * r.clo + (bb * bb * bb - bb) * r.chi);
}

// SPot1D, mc_poten.cc:624-639
__device__ __forceinline__ double spot1d(const Params &p, const SmallTables &t, double r, int *klo_out = nullptr)
{
   int n = p.n1d;
   if (klo_out) *klo_out = -1;
   const double x0 = t.rec1d[0].xlo, xn = t.rec1d[n - 2].xhi;
   if (r >= xn) return -p.c6 / pow(r, 6.0);
   if (r <= x0) return p.unode * exp(-p.alpha * r);
   return spline_rec_eval(t.rec1d, n, t.lut1d, p.nlut1d, p.lut1d_scale, x0, r, klo_out);
}

// SPot1D inside the move kernel.  The two extrapolation branches (beyond either end of the grid, rare) are kept out of
// line; the interval is guessed from the uniform-grid quotient or the bucket table and checked against the record, so
// it is always the reference's klo = max{k : x_k <= r}.
__device__ __noinline__ double spot1d_tail(const Params &p, double r)
{
   if (r >= p.xn_1d) return -p.c6 / pow(r, 6.0);
   return p.unode * exp(-p.alpha * r);
}
__device__ __forceinline__ double spot1d_move(const Params &p, const SmallTables &t, double r)
{
   const int n = p.n1d;
   if (__builtin_expect(!(r < p.xn_1d && r > p.x0_1d), 0)) return spot1d_tail(p, r);
   int k;
   if (p.uniform1d) k = min((int)((r - p.x0_1d) * p.invh_1d), n - 2);
   else {
      int b = (int)((r - p.x0_1d) * p.lut1d_scale);
      k = t.lut1d[min(b, p.nlut1d - 1)];
   }
   const SplineRec *rec = t.rec1d;
   SplineRec rc = rec[k];
   if (__builtin_expect(!(rc.xlo <= r && r < rc.xhi), 0)) {
      while (k < n - 2 && rec[k].xhi <= r) k++;
      while (k > 0 && rec[k].xlo > r) k--;
      rc = rec[k];
   }
   const double a = (rc.xhi - r) * rc.inv_h, bb = (r - rc.xlo) * rc.inv_h;
   return a * rc.ylo + bb * rc.yhi + ((a * a * a - a) * rc.clo + (bb * bb * bb - bb) * rc.chi);
}

// Branch-free form for batches of independent evaluations on a (quasi-)uniform grid: value from the guessed interval,
// `bad` set when the guess is not the reference's interval or r lies beyond the grid; the caller then redoes the batch
// with spot1d_move.  No data-dependent branch, so the evaluations of a batch interleave in the pipeline.
__device__ __forceinline__ double spot1d_try(const Params &p, const SmallTables &t, double r, bool &bad)
{
   int k = (int)((r - p.x0_1d) * p.invh_1d);
   k = max(0, min(k, p.n1d - 2));
   const SplineRec rc = t.rec1d[k];
   bad |= !(rc.xlo <= r && r < rc.xhi && r > p.x0_1d);
   const double a = (rc.xhi - r) * rc.inv_h, bb = (r - rc.xlo) * rc.inv_h;
   return a * rc.ylo + bb * rc.yhi + ((a * a * a - a) * rc.clo + (bb * bb * bb - bb) * rc.chi);
}

// The same spline piece as a cubic in t = (r - x_k)/h on a grid that is uniform to rounding (parah2.pot, isoH2H208.pot):
// c0 + t(c1 + t(c2 + t c3)) with c0 = y_k, c1 = y_k+1 - y_k - 2C_k - C_k+1, c2 = 3C_k, c3 = C_k+1 - C_k, C = y'' h^2/6
// (splint's a y_k + b y_k+1 + (a^3 - a)C_k + (b^3 - b)C_k+1 with a = 1 - t, b = t, mc_utils.cc:178-185).  The interval
// comes from the quotient; a point within rounding of a knot may land in the neighbouring piece, where the C^2 spline
// has the same value to rounding.  `bad` is set beyond either end of the grid (the caller redoes those with spot1d_move).
__device__ __forceinline__ double spot1d_poly(const Params &p, const SmallTables &t, double r, bool &bad)
{
   const double xq = (r - p.x0_1d) * p.invh_1d;
   int k = (int)xq;
   k = max(0, min(k, p.n1d - 2));
   bad |= !(r > p.x0_1d && r < p.xn_1d);
   const double tt = xq - (double)k;
   const double2 A = lds_d2(t.s_pa1d, k), B = lds_d2(t.s_pb1d, k);
   return fma(fma(fma(B.y, tt, B.x), tt, A.y), tt, A.x);
}

// floor(x/delta) by true division: the rare exact path of lpot2d's index selection, kept out of line so ptxas does
// not if-convert the FP64 division into the common path
__device__ __noinline__ double exact_floor_div(double x, double delta) { return floor(x / delta); }

// the four corner values of a bilinear cell from the row-pair table: two adjacent 128-bit gathers (usually one
// sector) instead of eight scalar ones
__device__ __forceinline__ void load_cell(const double2 *cell, double &y1, double &y2, double &y3, double &y4)
{
   const double2 a = __ldg(cell), b = __ldg(cell + 1);
   y1 = a.x; y2 = a.y; y4 = b.x; y3 = b.y;
}
// one bilinear cell from the whole-cell table: a single 256-bit gather (LDG.E.256 on sm_100a), kept in L2 with priority --
// the hot part of the table (tens of MB) competes with the streamed geometry records for one L2 partition
__device__ __forceinline__ void load_cell4(const double *cell, double &y1, double &y2, double &y3, double &y4)
{
   asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(y1), "=d"(y2), "=d"(y4), "=d"(y3) : "l"(cell));
}
__device__ __forceinline__ void load_cell4_keep(const double *cell, double &y1, double &y2, double &y3, double &y4)
{
   asm("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(y1), "=d"(y2), "=d"(y4), "=d"(y3) : "l"(cell));
}
__device__ __forceinline__ void load_geo4_keep(const double *rec, double &ux, double &uy, double &uz, double &w)
{
   asm volatile("ld.global.L1::no_allocate.L2::evict_last.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(ux), "=d"(uy), "=d"(uz), "=d"(w) : "l"(rec) : "memory");
}
__device__ __forceinline__ void load_geo4_plain(const double *rec, double &ux, double &uy, double &uz, double &w)
{
   asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(ux), "=d"(uy), "=d"(uz), "=d"(w) : "l"(rec) : "memory");
}
__device__ __forceinline__ void load_cell4_stream(const double *cell, double &y1, double &y2, double &y3, double &y4)
{
   asm("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(y1), "=d"(y2), "=d"(y4), "=d"(y3) : "l"(cell));
}
// one geometry record: streamed once per rotational sweep, must not displace the table in L2
__device__ __forceinline__ void load_geo4(const double *rec, double &ux, double &uy, double &uz, double &w)
{
   asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(ux), "=d"(uy), "=d"(uz), "=d"(w) : "l"(rec) : "memory");
}
// index selection of LPot2D, mc_poten.cc:696-704: floor((x - xmin)/delta) exactly as the reference -- the product with
// the stored reciprocal decides unless it lands within 1e-7 of an integer, where the true quotient is taken
// (|x*inv - x/delta| < 1e-11 here) -- then the clamp to [0, size-2]
__device__ __forceinline__ int lpot_index(double x, double inv_delta, double delta, int size)
{
   // nearest integer of x*inv_delta by the 2^52+2^51 trick (no FP64<->int conversion instructions), then floor
   const double MAGIC = 6755399441055744.0;
   const double xq = x * inv_delta;
   const double tq = xq + MAGIC;
   int i = __double2loint(tq);
   const double d = xq - (tq - MAGIC);
   if (__builtin_expect(fabs(d) < 1e-7, 0)) i = (int)exact_floor_div(x, delta);
   else if (d < 0.0) i -= 1;
   return max(0, min(i, size - 2));
}
// r and 1/r from r^2 with one reciprocal-square-root (<= 1 ulp each); used inside the move kernel where the IEEE
// sqrt + divide of the reference would cost ~40 more instructions per pair (the batched parity entry points keep
// the exact forms)
__device__ __forceinline__ void fast_r_invr(double r2, double &r, double &invr)
{
   // MUFU.RSQ64H seed (about 20 bits) + two Newton steps y <- y(1.5 - 0.5 r2 y^2); r2 is a squared distance between
   // distinct beads (positive, normal), so the special-case handling of the library rsqrt() is not needed
   double y;
   asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));
   const double h = 0.5 * r2;
   y = y * fma(-h * y, y, 1.5);
   y = y * fma(-h * y, y, 1.5);
   invr = y;
   r = r2 * y;
}
// LPot2D, mc_poten.cc:688-729
__device__ __forceinline__ double lpot2d(const Params &p, const SmallTables &t, double r, double cost, int *pir = nullptr, int *pic = nullptr)
{
   const double rmin = t.rgi2d[0].x, cmin = t.cgi2d[0].x;
   const int ir = lpot_index(r - rmin, p.inv_dr2d, p.dr2d, p.rs2d);
   const int ic = lpot_index(cost - cmin, p.inv_dc2d, p.dc2d, p.cs2d);
   if (pir) *pir = ir;
   if (pic) *pic = ic;
   double y1, y2, y3, y4;
   load_cell(p.cell2d + (size_t)ir * p.cs2d + ic, y1, y2, y3, y4);
   const double2 gr = t.rgi2d[ir], gc = t.cgi2d[ic];
   double dr = (r - gr.x) * gr.y;
   double dc = (cost - gc.x) * gc.y;
   return (1.0 - dr) * (1.0 - dc) * y1 + dr * (1.0 - dc) * y2 + dr * dc * y3 + (1.0 - dr) * dc * y4;
}
// NB independent LPot2D evaluations: index arithmetic first, the NB cell gathers next, the bilinear forms last
template <int NB>
__device__ __forceinline__ void lpot2d_xn(const Params &p, const SmallTables &t, const double *r, const double *cost, double *out)
{
   const double rmin = t.rgi2d[0].x, cmin = t.cgi2d[0].x;
   int ir[NB], ic[NB];
   #pragma unroll
   for (int u = 0; u < NB; u++) {
      ir[u] = lpot_index(r[u] - rmin, p.inv_dr2d, p.dr2d, p.rs2d);
      ic[u] = lpot_index(cost[u] - cmin, p.inv_dc2d, p.dc2d, p.cs2d);
   }
   double y1[NB], y2[NB], y3[NB], y4[NB];
   #pragma unroll
   for (int u = 0; u < NB; u++) load_cell(p.cell2d + (size_t)ir[u] * p.cs2d + ic[u], y1[u], y2[u], y3[u], y4[u]);
   #pragma unroll
   for (int u = 0; u < NB; u++) {
      const double2 gr = lds_d2(t.s_rgi2d, ir[u]), gc = lds_d2(t.s_cgi2d, ic[u]);      // move kernel only: axis tables staged in shared memory
      double dr = (r[u] - gr.x) * gr.y;
      double dc = (cost[u] - gc.x) * gc.y;
      out[u] = (1.0 - dr) * (1.0 - dc) * y1[u] + dr * (1.0 - dc) * y2[u] + dr * dc * y3[u] + (1.0 - dr) * dc * y4[u];
   }
}

// SRotDens / SRotDensDeriv / SRotDensEsqrt, mc_poten.cc:548-622 (which = 0,1,2)
__device__ __forceinline__ double srot_eval(const Params &p, const double *g, const double *y, const double *y2,
                                            const int *lut, double gamma, int which)
{
   int size = p.nrot;
   if (gamma > g[size - 1]) return which == 0 ? y[size - 1] : 0.0;
   if (gamma < g[0]) {
      double rl = g[0], rr = g[1];
      double salpha = (y[1] - y[0]) / (rr - rl);
      double sbeta = (y[0] * rr - y[1] * rl) / (rr - rl);
      return salpha * gamma + sbeta;
   }
   int k = spline_klo(g, size, lut, p.nlutrot, p.lutrot_scale, gamma);
   return splint_eval(g, y, y2, k, gamma);
}
__device__ __forceinline__ double srotdens(const Params &p, const SmallTables &t, double gamma)
{
   const int n = p.nrot;
   const SplineRec *rec = t.recrot;
   if (gamma > rec[n - 2].xhi) return rec[n - 2].yhi;
   if (gamma < rec[0].xlo) {
      double rl = rec[0].xlo, rr = rec[0].xhi, y0 = rec[0].ylo, y1 = rec[0].yhi;
      double salpha = (y1 - y0) / (rr - rl);
      double sbeta = (y0 * rr - y1 * rl) / (rr - rl);
      return salpha * gamma + sbeta;
   }
   return spline_rec_eval(rec, n, t.lutrot, p.nlutrot, p.lutrot_scale, rec[0].xlo, gamma, nullptr);
}

// ---------------------------------------------------------------------------------------------
// rigid-body leaves
// ---------------------------------------------------------------------------------------------
struct Mat3 { double m[3][3]; };   // m[i][j] = rotmat(i+1, j+1)

// matpre, rotden.f:136-163; eul = (phi, theta, chi)
__device__ __forceinline__ void matpre(double phi, double theta, double chi, Mat3 &r)
{
   double sp, cp, st, ct, sk, ck;
   sincos(phi, &sp, &cp);
   sincos(theta, &st, &ct);
   sincos(chi, &sk, &ck);
   r.m[0][0] = cp * ct * ck - sp * sk;
   r.m[0][1] = -cp * ct * sk - sp * ck;
   r.m[0][2] = cp * st;
   r.m[1][0] = sp * ct * ck + cp * sk;
   r.m[1][1] = -sp * ct * sk + cp * ck;
   r.m[1][2] = sp * st;
   r.m[2][0] = -st * ck;
   r.m[2][1] = st * sk;
   r.m[2][2] = ct;
}
__device__ __forceinline__ double within(double v) { return v > 1.0 ? 1.0 : (v < -1.0 ? -1.0 : v); }

// (phi, theta, chi) of a rotation matrix: the extraction shared by deleul (rotden.f:62-119) and rflmfx/y/z (vcord.f:293-343)
__device__ __forceinline__ void euler_from_matrix(const double (&m)[3][3], double &phi2, double &theta2, double &chi2)
{
   const double small = 1.0e-08;
   double cost = within(m[2][2]);
   theta2 = acos(cost);
   double sint = sin(theta2);
   if (fabs(1.0 - cost) < small) {
      phi2 = 0.0;
      double cchi = within(m[0][0]), schi = within(m[1][0]);
      chi2 = (schi > 0.0) ? acos(cchi) : 2.0 * PI - acos(cchi);
   } else if (fabs(1.0 + cost) < small) {
      phi2 = 0.0;
      double cchi = within(m[1][1]), schi = within(m[0][1]);
      chi2 = (schi > 0.0) ? acos(cchi) : 2.0 * PI - acos(cchi);
   } else {
      double cphi = within(m[0][2] / sint), sphi = within(m[1][2] / sint);
      double cchi = within(-m[2][0] / sint), schi = within(m[2][1] / sint);
      phi2 = (sphi > 0.0) ? acos(cphi) : 2.0 * PI - acos(cphi);
      chi2 = (schi > 0.0) ? acos(cchi) : 2.0 * PI - acos(cchi);
   }
}
// Euler angles of R = R1^T R2 (deleul, rotden.f:32-134), radians
__device__ __forceinline__ void deleul(const Mat3 &r1, const Mat3 &r2, double &phi2, double &theta2, double &chi2)
{
   double m[3][3];
   #pragma unroll
   for (int i = 0; i < 3; i++)
      #pragma unroll
      for (int j = 0; j < 3; j++) {
         double s = 0.0;
         #pragma unroll
         for (int k = 0; k < 3; k++) s = s + r1.m[k][i] * r2.m[k][j];
         m[i][j] = s;
      }
   euler_from_matrix(m, phi2, theta2, chi2);
}
// rsrot_ (rotden.f:218-286, the live branch iodevn = -1): rattle-and-shake propagator exponent numerator `rho`
// and energy estimator `erot` (K) for a top with rotational constants x, y, z (cm^-1); tauC = MCRotTau (1/K)
__device__ __forceinline__ double rsrot(const Params &p, const Mat3 &r1, const Mat3 &r2, double *erot)
{
   const double tau = p.rottau / WNO2K;
   const double b0 = 1.0 / p.xrot, b1 = 1.0 / p.yrot, b2 = 1.0 / p.zrot;
   double dg[3];
   #pragma unroll
   for (int i = 0; i < 3; i++) {
      double s = 0.0;
      #pragma unroll
      for (int j = 0; j < 3; j++) s = s + r1.m[j][i] * r2.m[j][i];
      dg[i] = s;
   }
   const double sumaxs = (b0 - b1 - b2) * (1.0 - dg[0]) + (b1 - b2 - b0) * (1.0 - dg[1]) + (b2 - b0 - b1) * (1.0 - dg[2]);
   if (erot) {
      double e = sumaxs / (4.0 * tau * tau);
      e = e + 1.5 / tau + 0.25 * (p.xrot + p.yrot + p.zrot);
      *erot = e / WNO2K;
   }
   return sumaxs;
}
// rsline_ (rotden.f:356-375): -(1 - gamma)/(2 B tau) for a linear rotor with constant B = X_Rot
__device__ __forceinline__ double rsline(const Params &p, double dprd, double *erot)
{
   const double tau = p.rottau / WNO2K;
   const double r = (1.0 - dprd) / (2.0 * p.xrot * tau);
   if (erot) *erot = ((1.0 - r) / tau) / WNO2K;
   return -r;
}

// rotpro (rotpro_sub.f:1-64): a pure function of the three relative angles in DEGREES -- index selection by int()
// truncation (:7-9), forward differences along chi, phi, theta, zero difference on the last grid line.
// erot/esq are still in the tables' cm^-1 units here (rotden_ converts).
__device__ __forceinline__ double rotpro(const Params &p, double phi, double theta, double chi, double *erot, double *esq,
                                         int *index, int *istop)
{
   int ichi = (int)chi, iphi = (int)phi, itheta = (int)theta;
   if (ichi > 360 || ichi < 0) { ichi = 0; if (istop) *istop = 1; }
   if (iphi > 360 || iphi < 0) { iphi = 0; if (istop) *istop = 1; }
   if (itheta > 180 || itheta < 0) { itheta = 0; if (istop) *istop = 1; }
   int ind = (itheta * 361 + iphi) * 361 + ichi;
   if (index) *index = ind;
   int kc = (ichi != 360) ? ind + 1 : ind;
   int kp = (iphi != 360) ? ind + 361 : ind;
   int kt = (itheta != 180) ? ind + 361 * 361 : ind;
   double fc = chi - (double)ichi, fp = phi - (double)iphi, ft = theta - (double)itheta;
   double rho0 = __ldg(p.rho3 + ind);
   double rho = rho0 + (__ldg(p.rho3 + kc) - rho0) * fc + (__ldg(p.rho3 + kp) - rho0) * fp + (__ldg(p.rho3 + kt) - rho0) * ft;
   if (erot) {
      double e0 = __ldg(p.erot3 + ind);
      *erot = e0 + (__ldg(p.erot3 + kc) - e0) * fc + (__ldg(p.erot3 + kp) - e0) * fp + (__ldg(p.erot3 + kt) - e0) * ft;
      double q0 = __ldg(p.esq3 + ind);
      *esq = q0 + (__ldg(p.esq3 + kc) - q0) * fc + (__ldg(p.esq3 + kp) - q0) * fp + (__ldg(p.esq3 + kt) - q0) * ft;
   }
   return rho;
}

// rotden_ (rotden.f:1-31): deleul, radians -> degrees, rotpro, cm^-1 -> K.  Returns rho; erot/esq only when the
// pointers are non-null (moves need rho alone).  *istop mirrors the Fortran's out-of-range flag.
__device__ __forceinline__ double rotden(const Params &p, const Mat3 &r1, const Mat3 &r2, double *rel, double *erot,
                                         double *esq, int *index, int *istop)
{
   double phi, theta, chi;
   deleul(r1, r2, phi, theta, chi);
   if (rel) { rel[0] = phi; rel[1] = theta; rel[2] = chi; }
   phi = phi * 180.0 / PI;
   theta = theta * 180.0 / PI;
   chi = chi * 180.0 / PI;
   const double rho = rotpro(p, phi, theta, chi, erot, esq, index, istop);
   if (erot) {
      *erot = *erot / WNO2K;
      *esq = *esq / (WNO2K * WNO2K);
   }
   return rho;
}

// vcalc (vcalc.f:1-65): a pure function of r (bohr), theta and chi (degrees) -- r clamped to the table's range, indices by
// int() truncation clamped to the grid (:16-25), forward differences, zero gradient on the last grid line (:33-58)
__device__ __forceinline__ double vcalc(const Params &p, double r, double theta, double chi, int *index)
{
   int maxrpt = p.rg3 - 1, mxthpt = p.thg3 - 1, mxchpt = p.chg3 - 1;
   if (r < p.rvmin) r = p.rvmin;
   if (r > p.rvmax) r = p.rvmax;
   int ir = (int)((r - p.rvmin) / p.rvstep);
   // "index sits on the last grid line" is decided on the doubles (theta, chi >= 0 here), not by comparing the
   // clamped integers: ptxas 12.9 fuses min/max/compare into a predicated VIMNMX whose predicate came out
   // inverted on sm_100a (angular gradient terms silently dropped) -- see DESIGN.md "Toolchain notes".
   const bool th_last = !(theta < (double)mxthpt), ch_last = !(chi < (double)mxchpt);
   int ith = th_last ? mxthpt : (int)theta;
   int ich = ch_last ? mxchpt : (int)chi;
   if (ith < 0) ith = 0;
   if (ich < 0) ich = 0;
   int ind = (ir * p.thg3 + ith) * p.chg3 + ich;
   if (index) *index = ind;
   double v0 = __ldg(p.v3d + ind);
   double gradr = 0.0, delr = 0.0;
   if (ir != maxrpt) { gradr = (__ldg(p.v3d + ind + p.thg3 * p.chg3) - v0) / p.rvstep; delr = r - (p.rvmin + ir * p.rvstep); }
   const double gradth = __ldg(p.v3d + (th_last ? ind : ind + p.chg3)) - v0;
   const double delth = th_last ? 0.0 : theta - (double)ith;
   const double gradch = __ldg(p.v3d + (ch_last ? ind : ind + 1)) - v0;
   const double delch = ch_last ? 0.0 : chi - (double)ich;
   return v0 + gradr * delr + gradth * delth + gradch * delch;
}

// vcord_ (vcord.f:1-98) for a rotor with rotation matrix `rm` at `rcom` and a point particle at `rpt`.
// rtc (optional) receives radret, theret, chiret; grid (optional) the arguments handed to vcalc (bohr, degrees).
__device__ __forceinline__ double vcord(const Params &p, const Mat3 &rm, const double *rcom, const double *rpt,
                                        double *rtc, int *index, double *grid = nullptr)
{
   const double small = 1.0e-08, bo2ang = 0.529177249;
   double R[3] = {rpt[0] - rcom[0], rpt[1] - rcom[1], rpt[2] - rcom[2]};
   // hatx, haty, hatz = columns of the rotation matrix (rottrn of the unit vectors)
   double dx = 0.0, dy = 0.0, dz = 0.0, nz = 0.0, rr = 0.0;
   #pragma unroll
   for (int i = 0; i < 3; i++) {
      dx = dx + R[i] * rm.m[i][0];
      dy = dy + R[i] * rm.m[i][1];
      dz = dz + rm.m[i][2] * R[i];
      nz = nz + rm.m[i][2] * rm.m[i][2];
      rr = rr + R[i] * R[i];
   }
   double radwff = sqrt(rr);
   double ca = dz / (sqrt(nz) * radwff);
   ca = ca > 1.0 ? 1.0 : (ca < -1.0 ? -1.0 : ca);
   double thewff = acos(ca);
   double chiwff;
   if (fabs(dx) < small) chiwff = PI / 2.0;
   else chiwff = atan(fabs(dy / dx));
   double chiret;
   if (dx >= 0.0 && dy >= 0.0) chiret = chiwff;
   else if (dx < 0.0 && dy >= 0.0) chiret = PI - chiwff;
   else if (dx < 0.0 && dy < 0.0) chiret = PI + chiwff;
   else chiret = 2 * PI - chiwff;
   if (rtc) { rtc[0] = radwff; rtc[1] = thewff; rtc[2] = chiret; }
   if (p.chg3 == 181) { chiwff = chiret; if (chiret > PI) chiwff = 2 * PI - chiret; }
   else if (p.chg3 == 361) chiwff = chiret;
   double r = radwff / bo2ang;
   double theta = thewff * 180.0 / PI;
   double chi = chiwff * 180.0 / PI;
   if (grid) { grid[0] = r; grid[1] = theta; grid[2] = chi; }
   return vcalc(p, r, theta, chi, index);
}

// vspher_, vspher.f:519-543 (r in Angstrom).  ang2bo is a REAL*4 literal widened to double in the Fortran PARAMETER
// statement (no D exponent), like the table entries.  *rclamp (optional) receives what the Fortran leaves in its `r`
// argument: the clamped distance in bohr, which GetPotEnergy_Densities then bins (mc_estim.cc:631-637).
__device__ __forceinline__ double vspher(const Params &p, double r, double *rclamp = nullptr)
{
   const double r0 = 3.0, rmax = 26.0, rstep = 0.046, ang2bo = (double)0.5291772f;
   r = r / ang2bo;
   if (r < r0) r = r0;
   if (r > rmax) r = rmax;
   if (rclamp) *rclamp = r;
   int ir = (int)((r - r0) / rstep);
   double v0 = __ldg(p.vspher + ir);
   if (ir == 500) return v0;
   double gradr = (__ldg(p.vspher + ir + 1) - v0) / rstep;
   return v0 + gradr * (r - (r0 + ir * rstep));
}

// TIP4P site frame: rottrn of the four body-frame sites (caleng_tip4p_gg.f:37-94)
struct Tip4pSites { double o[3], m[3], h1[3], h2[3]; };
__device__ __forceinline__ void tip4p_sites(const Mat3 &r, const double *com, Tip4pSites &s)
{
   #pragma unroll
   for (int i = 0; i < 3; i++) {
      // rsf(i) = rcom(i) + sum_j rotmat(i,j)*rwf(j), body sites O(0,0,.06562) M(0,0,-.08438) H(+-.7557,0,-.5223)
      // (the zero components of the body-frame sites add exact zeros and are dropped)
      s.o[i]  = com[i] + r.m[i][2] * 0.06562;
      s.m[i]  = com[i] + r.m[i][2] * -0.08438;
      s.h1[i] = (com[i] + r.m[i][0] * 0.7557) + r.m[i][2] * -0.5223;
      s.h2[i] = (com[i] + r.m[i][0] * -0.7557) + r.m[i][2] * -0.5223;
   }
}
__device__ __forceinline__ double dist2(const double *a, const double *b)
{
   double s = 0.0;
   #pragma unroll
   for (int i = 0; i < 3; i++) s = s + (a[i] - b[i]) * (a[i] - b[i]);
   return s;
}
// caleng_, caleng_tip4p_gg.f:95-183
__device__ __forceinline__ double caleng(const Tip4pSites &a, const Tip4pSites &b)
{
   const double qm = -1.04, qh = 0.520, br2ang = 0.52917721092, hr2k = 3.1577465e5, kcal2k = 503.218978939;
   double roo = dist2(a.o, b.o);
   double rmm = sqrt(dist2(a.m, b.m));
   double roo4 = roo * roo, roo6 = roo4 * roo, roo12 = roo6 * roo6;
   double v_o2lj = 6.0e5 / roo12 - 610.0 / roo6;
   double rhm1 = sqrt(dist2(a.m, b.h1)), rhm2 = sqrt(dist2(a.m, b.h2));
   double rhm3 = sqrt(dist2(b.m, a.h1)), rhm4 = sqrt(dist2(b.m, a.h2));
   double rhh1 = sqrt(dist2(a.h1, b.h1)), rhh2 = sqrt(dist2(a.h1, b.h2));
   double rhh3 = sqrt(dist2(a.h2, b.h1)), rhh4 = sqrt(dist2(a.h2, b.h2));
   double v_mh = qm * qh * (1.0 / rhm1 + 1.0 / rhm2 + 1.0 / rhm3 + 1.0 / rhm4);
   double v_hh = qh * qh * (1.0 / rhh1 + 1.0 / rhh2 + 1.0 / rhh3 + 1.0 / rhh4);
   double v_mm = qm * qm * (1.0 / rmm);
   return v_o2lj * kcal2k + (v_mh + v_mm + v_hh) * hr2k * br2ang;
}

// ---------------------------------------------------------------------------------------------
// one pair term of PotEnergy (mc_piqmc.cc:1822-1959): atom0 with bead position pos0 against atom1
// at slice `it` of chain c.  rm0/n0 (optional) override atom0's stored orientation: PotRotE3D's
// Eulang argument (mc_piqmc.cc:2046-2151) and PotRotEnergy's cosine argument (:1967-2044).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int type_of(const Params &p, int atom) { return atom < p.first[1] ? 0 : 1; }

__device__ __forceinline__ void load_rotmat(const Params &p, int c, int q, int m, Mat3 &r)
{
   double phi = p.ang[ang_index(p, c, q, 0, m)];
   double cth = p.ang[ang_index(p, c, q, 1, m)];
   double chi = p.ang[ang_index(p, c, q, 2, m)];
   matpre(phi, acos(cth), chi, r);
}

// KIND prunes the branches at compile time: -1 all, 0 atoms only, 1 linear-rotor system, 2 top system.
template <int KIND = -1>
__device__ __forceinline__ double pair_energy(const Params &p, const SmallTables &t, int c, int atom0, const double *pos0,
                                              int atom1, int it, const Mat3 *rm0, const double *n0)
{
   int type0 = type_of(p, atom0), type1 = type_of(p, atom1);
   int mode = p.mode[type0][type1];
   if (KIND == 0) mode = M_SPOT1D;
   if (KIND == 1 && mode != M_LIN_0MOL && mode != M_LIN_1MOL) mode = M_SPOT1D;
   if (KIND == 2 && (mode == M_LIN_0MOL || mode == M_LIN_1MOL)) mode = M_SPOT1D;
   double p1[3], dr[3], dr2 = 0.0;
   #pragma unroll
   for (int d = 0; d < 3; d++) {
      p1[d] = p.pos[pos_index(p, c, it, d, atom1)];
      dr[d] = pos0[d] - p1[d];
      if (p.minimage) dr[d] -= p.box[d] * rint(dr[d] / p.box[d]);
      dr2 += dr[d] * dr[d];
   }
   int q = it / p.R;
   if ((KIND == -1 || KIND == 1) && (mode == M_LIN_0MOL || mode == M_LIN_1MOL)) {
      double r = sqrt(dr2);
      double n[3];
      int sgn;
      if (mode == M_LIN_0MOL) {
         sgn = -1;
         if (n0) { n[0] = n0[0]; n[1] = n0[1]; n[2] = n0[2]; }
         else { int m = atom0 - p.first[p.imtype]; for (int d = 0; d < 3; d++) n[d] = p.cosn[ang_index(p, c, q, d, m)]; }
      } else {
         sgn = 1;
         int m = atom1 - p.first[p.imtype];
         for (int d = 0; d < 3; d++) n[d] = p.cosn[ang_index(p, c, q, d, m)];
      }
      double cost = 0.0;
      #pragma unroll
      for (int d = 0; d < 3; d++) cost += n[d] * dr[d];
      cost /= r;
      cost *= sgn;
      return lpot2d(p, t, r, cost);
   }
   if (KIND == -1 || KIND == 2) {
      if (mode == M_TOP_0MOL) {
         Mat3 rl;
         const Mat3 *rm = rm0;
         if (!rm) { load_rotmat(p, c, q, atom0 - p.first[p.imtype], rl); rm = &rl; }
         return vcord(p, *rm, pos0, p1, nullptr, nullptr);
      }
      if (mode == M_TOP_1MOL) {
         Mat3 rl;
         load_rotmat(p, c, q, atom1 - p.first[p.imtype], rl);
         return vcord(p, rl, p1, pos0, nullptr, nullptr);
      }
      if (mode == M_SPHER) return vspher(p, sqrt(dr2));
      if (mode == M_TOPTOP) {
         Mat3 ra, rb;
         const Mat3 *rm = rm0;
         if (!rm) { load_rotmat(p, c, q, atom0 - p.first[p.imtype], ra); rm = &ra; }
         load_rotmat(p, c, q, atom1 - p.first[p.imtype], rb);
         Tip4pSites sa, sb;
         tip4p_sites(*rm, pos0, sa);
         tip4p_sites(rb, p1, sb);
         return caleng(sa, sb);
      }
   }
   return spot1d(p, t, sqrt(dr2));
}

} // namespace pimc
