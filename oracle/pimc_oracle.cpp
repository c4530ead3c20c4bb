// ORACLE (test infrastructure, not product code).
//
// CPU restatement of MoRiBS-PIMC's sampling hot path: per-bead potential
// sums, tabulated/analytic pair potentials, rotational density look-ups,
// translational and rotational Metropolis moves and the per-slice estimator
// sums.  Every function cites the reference file:line it follows.  The
// Fortran leaves are reached through the gfortran-ABI symbols of
// oracle/fortran_shim.cpp.  Pinned against the reference's own C++ objects
// (oracle/_ref/libpimcref.so) by tests/test_oracle_vs_ref.py; the Fortran
// leaves themselves are "parity unpinned" (see fortran_shim.cpp).
//
// Build: g++ -O2 -ffp-contract=off (oracle/Makefile).
#include "pimc_oracle.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <array>
#include <deque>

extern "C" {
void rotden_(double *, double *, double *, double *, double *, double *, double *, double *, double *, int *);
void vcord_(double *, double *, double *, double *, int *, int *, int *, double *, double *, double *,
            double *, double *, double *, double *, double *, double *, double *, int *);
void caleng_(double *, double *, double *, double *, double *);
void vspher_(double *, double *);
void rsrot_(double *, double *, double *, double *, double *, double *, int *, double *, double *, double *);
void rsline_(double *, double *, double *, double *, double *);
void rflmfx_(double *, double *, double *, double *, double *);
void rflmfy_(double *, double *, double *, double *, double *);
void rflmfz_(double *, double *, double *, double *, double *);
void oracle_set_vspher_table(const double *);
extern int oracle_last_rotden_index, oracle_last_vcord_index;
void oracle_rotpro(double, double, double, const double *, const double *, const double *, double *, double *, double *, int *, int *);
double oracle_vcalc(double, double, double, double, double, double, int, int, int, const double *, int *);
void oracle_deleul(const double *, const double *, double *);
}

namespace {

// mc_const.h:12-15, mc_confg.h:60
const double HBAR = 1.05457266, AMU = 1.6605402, K_B = 1.380658, WNO2K = 0.6950356;
const double RZERO = 1.0e-10;
const int PHI = 0, CTH = 1, CHI = 2;
const int MCMOLEC = 0, MCMULTI = 1, MCROTAT = 2;
// mc_estim.cc:25-30
const int MC_BINSR = 300, MC_BINST = 50, MC_BINSC = 100;
const double MAX_RADIUS = 15.0, MIN_RADIUS = 0.0;

// ---- cubic spline, mc_utils.cc:112-154 ------------------------------------
void spline(const double *x, const double *y, int n, double yp1, double ypn, double *y2)
{
   std::vector<double> u(n);
   double p, qn, sig, un;
   if (yp1 > 0.99e30) y2[0] = u[0] = 0.0;
   else {
      y2[0] = -0.5;
      u[0] = (3. / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - yp1);
   }
   for (int i = 1; i < (n - 1); i++) {
      sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
      p = sig * y2[i - 1] + 2.;
      y2[i] = (sig - 1.) / p;
      u[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
      u[i] = (6. * u[i] / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
   }
   if (ypn > 0.99e30) qn = un = 0.;
   else {
      qn = .5;
      un = (3. / (x[n - 1] - x[n - 2])) * (ypn - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]));
   }
   y2[n - 1] = (un - qn * u[n - 2]) / (qn * y2[n - 2] + 1.);
   for (int k = n - 2; k >= 0; k--) y2[k] = y2[k] * y2[k + 1] + u[k];
}
// mc_utils.cc:188-201
void init_spline(const double *grid, const double *data, double *sdata, int n)
{
   double drl = grid[1] - grid[0];
   double dpl = (data[1] - data[0]) / drl;
   double drr = grid[n - 1] - grid[n - 2];
   double dpr = (data[n - 1] - data[n - 2]) / drr;
   spline(grid, data, n, dpl, dpr, sdata);
}
// mc_utils.cc:156-186
double splint(const double *xa, const double *ya, const double *y2a, int n, double x, int *klo_out)
{
   int klo = 0, khi = n - 1;
   while (khi - klo > 1) {
      int k = (khi + klo) >> 1;
      if (xa[k] > x) khi = k; else klo = k;
   }
   double h = xa[khi] - xa[klo];
   if (klo_out) *klo_out = klo;
   double a = (xa[khi] - x) / h;
   double b = (x - xa[klo]) / h;
   return a * ya[klo] + b * ya[khi] + ((a * a * a - a) * y2a[klo] + (b * b * b - b) * y2a[khi]) * (h * h) / 6.;
}

// ---- MRG32k3a in the reference's double arithmetic, rngstream.cc:21-126,242-265
const double m1 = 4294967087.0, m2 = 4294944443.0, norm = 1.0 / (m1 + 1.0);
const double a12 = 1403580.0, a13n = 810728.0, a21 = 527612.0, a23n = 1370589.0;
const double two17 = 131072.0, two53 = 9007199254740992.0;
const double A1p127[3][3] = {{2427906178.0, 3580155704.0, 949770784.0},
                             {226153695.0, 1230515664.0, 3580155704.0},
                             {1988835001.0, 986791581.0, 1230515664.0}};
const double A2p127[3][3] = {{1464411153.0, 277697599.0, 1610723613.0},
                             {32183930.0, 1464411153.0, 1022607788.0},
                             {2824425944.0, 32183930.0, 2093834863.0}};
double MultModM(double a, double s, double c, double m)
{
   double v = a * s + c;
   long a1;
   if (v >= two53 || v <= -two53) {
      a1 = (long)(a / two17); a -= a1 * two17;
      v = a1 * s;
      a1 = (long)(v / m); v -= a1 * m;
      v = v * two17 + a * s + c;
   }
   a1 = (long)(v / m);
   if ((v -= a1 * m) < 0.0) return v += m; else return v;
}
void MatVecModM(const double A[3][3], const double s[3], double v[3], double m)
{
   double x[3];
   for (int i = 0; i < 3; ++i) {
      x[i] = MultModM(A[i][0], s[0], 0.0, m);
      x[i] = MultModM(A[i][1], s[1], x[i], m);
      x[i] = MultModM(A[i][2], s[2], x[i], m);
   }
   for (int i = 0; i < 3; ++i) v[i] = x[i];
}
void MatMatModM(const double A[3][3], const double B[3][3], double C[3][3], double m)
{
   double V[3], W[3][3];
   for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) V[j] = B[j][i];
      MatVecModM(A, V, V, m);
      for (int j = 0; j < 3; ++j) W[j][i] = V[j];
   }
   for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i][j] = W[i][j];
}
// state of the s-th stream = (A^(2^127))^s * seed, by binary powering of the jump matrix
void mrg_stream_state(const unsigned long *seed6, long s, double *st)
{
   double B1[3][3], B2[3][3], W1[3][3], W2[3][3];
   for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
      W1[i][j] = A1p127[i][j]; W2[i][j] = A2p127[i][j];
      B1[i][j] = B2[i][j] = (i == j) ? 1.0 : 0.0;
   }
   long n = s;
   while (n > 0) {
      if (n % 2) { MatMatModM(W1, B1, B1, m1); MatMatModM(W2, B2, B2, m2); }
      MatMatModM(W1, W1, W1, m1); MatMatModM(W2, W2, W2, m2);
      n /= 2;
   }
   double s1[3] = {(double)seed6[0], (double)seed6[1], (double)seed6[2]};
   double s2[3] = {(double)seed6[3], (double)seed6[4], (double)seed6[5]};
   MatVecModM(B1, s1, st, m1);
   MatVecModM(B2, s2, st + 3, m2);
}
// rngstream.cc:242-265 (anti == false)
double mrg_u01(double *Cg)
{
   long k;
   double p1, p2;
   p1 = a12 * Cg[1] - a13n * Cg[0];
   k = (long)(p1 / m1); p1 -= k * m1; if (p1 < 0.0) p1 += m1;
   Cg[0] = Cg[1]; Cg[1] = Cg[2]; Cg[2] = p1;
   p2 = a21 * Cg[5] - a23n * Cg[3];
   k = (long)(p2 / m2); p2 -= k * m2; if (p2 < 0.0) p2 += m2;
   Cg[3] = Cg[4]; Cg[4] = Cg[5]; Cg[5] = p2;
   return (p1 > p2) ? (p1 - p2) * norm : (p1 - p2 + m1) * norm;
}
// mc_randg.cc:138-150 with explicit uniforms
inline double gauss_u(double alpha, double r1, double r2)
{
   double x1 = sqrt(-log(r1)) * cos(2.0 * M_PI * r2);
   return x1 / sqrt(alpha);
}

} // namespace

struct orc {
   orc_system_t sys;
   int N, P, Q, R;
   int offset_atom[3];                // first global atom of each type
   double lambda[2], beta, tau, rottau;
   int imtype, bstype;
   std::vector<int> mctype;           // MCType
   // tables
   int n1d = 0; std::vector<double> g1d, v1d, y2_1d; double alpha = 0, unode = 0, c6 = 0;
   int rs2d = 0, cs2d = 0; double dr2d = 0, dc2d = 0; std::vector<double> rg2d, cg2d, v2d;
   int rg3 = 0, thg3 = 0, chg3 = 0; double rvmin = 0, rvmax = 0, rvstep = 0; const double *v3d = nullptr;
   int nrot = 0; std::vector<double> rgrid, rdens, rderv, resqr, rdens2, rderv2, resqr2;
   const double *rho3 = nullptr, *erot3 = nullptr, *esq3 = nullptr;
   // state, reference layout [dim][atom*P+it]
   std::vector<double> coords[3], angles[3], cosine[3], newc[3];
   std::vector<int> pindex, rindex;
   double mctotal[2][3], mcaccep[2][3];
   // histograms
   std::vector<double> gr1d, gr2d, gr3d[2], relthe, relphi, relchi;
   double delta_radius, delta_theta, delta_chi;
   // schedule streams: P translational, Q rotational, 8 misc
   std::vector<std::array<double, 6>> streams;
   double ErotSQ = 0, Erot_termSQ = 0;
   // worm (mc_qworm.h:32-47 TPathWorm, mc_qworm.cc:15-46): atoms are numbered inside the worm type
   struct { int on = 0, type = 0, exists = 0, ira = 0, masha = 0, atom_i = 0, atom_m = 0, m = 1;
            double c = 0, qw_norm = 0, cutoff2 = 0, twave2 = 0; } worm;
   double qwtotal[7] = {0, 0, 0, 0, 0, 0, 0}, qwaccep[7] = {0, 0, 0, 0, 0, 0, 0}, countqw = 1.0;
   std::deque<double> wq[15];          // explicit uniforms per SPRNG stream (mc_randg.cc:90-174), test mode
   int wmode = 0;                      // 0: queues, 1: the chain's worm stream of the device schedule
   std::vector<double> dr2_list, ptable; std::vector<int> atm_list;

   int type_offset(int t) const { return offset_atom[t] * P; }
};

// WorldLine(atom, pt), mc_qworm.cc:553-575: false when bead pt of world line `atom` lies in the ira-masha gap
static bool WorldLine(const orc_t *o, int atom, int pt)
{
   const auto &W = o->worm;
   bool wline = true;
   if ((atom == W.atom_m) || (atom == W.atom_i)) {
      if ((W.atom_i != W.atom_m) || (W.ira > W.masha)) {
         if (((atom == W.atom_m) && (pt < W.masha)) || ((atom == W.atom_i) && (pt > W.ira))) wline = false;
      } else {
         if ((pt > W.ira) && (pt < W.masha)) wline = false;
      }
   }
   return wline;
}
// the mask every PotEnergy variant applies to its partner loop (mc_piqmc.cc:1226-1227,1816-1817,2000-2001,2082-2083)
static inline bool partner_on_line(const orc_t *o, int atom1, int it)
{
   if (!(o->worm.on && o->worm.exists)) return true;
   int type1 = o->mctype[atom1];
   if (o->worm.type != type1) return true;
   return WorldLine(o, atom1 - o->offset_atom[type1], it);
}

// ----------------------------------------------------------------------------
// leaf potentials
// ----------------------------------------------------------------------------
static double SPot1D(orc_t *o, double r, int *klo)      // mc_poten.cc:624-639
{
   int n = o->n1d;
   if (klo) *klo = -1;
   if (r >= o->g1d[n - 1]) return -o->c6 / pow(r, 6.0);
   if (r <= o->g1d[0]) return o->unode * exp(-o->alpha * r);
   return splint(o->g1d.data(), o->v1d.data(), o->y2_1d.data(), n, r, klo);
}
static double LPot2D(orc_t *o, double r, double cost, int *pir, int *pic)   // mc_poten.cc:688-729
{
   double rmin = o->rg2d[0], cmin = o->cg2d[0];
   int rsize = o->rs2d, csize = o->cs2d;
   int ir = (int)floor((r - rmin) / o->dr2d);
   int ic = (int)floor((cost - cmin) / o->dc2d);
   if (ir < 0) ir = 0; else if (ir >= (rsize - 1)) ir = rsize - 2;
   if (ic < 0) ic = 0; else if (ic >= (csize - 1)) ic = csize - 2;
   if (pir) *pir = ir;
   if (pic) *pic = ic;
   const double *pot = o->v2d.data();
   double y1 = pot[ir * csize + ic], y2 = pot[(ir + 1) * csize + ic];
   double y3 = pot[(ir + 1) * csize + ic + 1], y4 = pot[ir * csize + ic + 1];
   double r1 = o->rg2d[ir], r2 = o->rg2d[ir + 1], c1 = o->cg2d[ic], c2 = o->cg2d[ic + 1];
   double dr = (r - r1) / (r2 - r1), dc = (cost - c1) / (c2 - c1);
   return (1.0 - dr) * (1.0 - dc) * y1 + dr * (1.0 - dc) * y2 + dr * dc * y3 + (1.0 - dr) * dc * y4;
}
// mc_poten.cc:548-622; which = 0 SRotDens, 1 SRotDensDeriv, 2 SRotDensEsqrt
static double SRot(orc_t *o, double gamma, int which)
{
   int size = o->nrot;
   const double *g = o->rgrid.data();
   const double *y = which == 0 ? o->rdens.data() : which == 1 ? o->rderv.data() : o->resqr.data();
   const double *y2 = which == 0 ? o->rdens2.data() : which == 1 ? o->rderv2.data() : o->resqr2.data();
   if (gamma > g[size - 1]) return which == 0 ? y[size - 1] : 0.0;
   if (gamma < g[0]) {
      double rl = g[0], rr = g[1];
      double salpha = (y[1] - y[0]) / (rr - rl);
      double sbeta = (y[0] * rr - y[1] * rl) / (rr - rl);
      return salpha * gamma + sbeta;
   }
   return splint(g, y, y2, size, gamma, nullptr);
}

static double vcord_call(orc_t *o, const double *eul, const double *rcom, const double *rpt, double *rtc)
{
   double E[3] = {eul[0], eul[1], eul[2]}, C[3] = {rcom[0], rcom[1], rcom[2]}, Pt[3] = {rpt[0], rpt[1], rpt[2]};
   double v, rad, the, chi, hx[3], hy[3], hz[3];
   int iv = 0;
   // argument order &Rvmax,&Rvmin as at mc_piqmc.cc:1316
   vcord_(E, C, Pt, const_cast<double *>(o->v3d), &o->rg3, &o->thg3, &o->chg3, &o->rvmax, &o->rvmin, &o->rvstep,
          &v, &rad, &the, &chi, hx, hy, hz, &iv);
   if (rtc) { rtc[0] = rad; rtc[1] = the; rtc[2] = chi; }
   return v;
}

// One pair term of PotEnergy: atom0 (bead pos0, orientation row tm0) with atom1
// at slice `it`.  eul0/cos0 override atom0's stored orientation when non-NULL
// (PotRotE3D / PotRotEnergy).  mc_piqmc.cc:1822-1959.
static double pair_energy(orc_t *o, int atom0, const double *pos0, int atom1, int it, const double *eul0,
                          const double *cos0, int hist, double *hist_out)
{
   int P = o->P;
   int type0 = o->mctype[atom0], type1 = o->mctype[atom1];
   int offset0 = P * atom0, offset1 = P * atom1;
   int t1 = offset1 + it;
   const orc_type_t &T0 = o->sys.type[type0], &T1 = o->sys.type[type1];
   double dr[3], dr2 = 0.0;
   for (int id = 0; id < 3; id++) {
      dr[id] = pos0[id] - o->coords[id][t1];
      if (o->sys.minimage) dr[id] -= o->sys.box[id] * rint(dr[id] / o->sys.box[id]);
      dr2 += dr[id] * dr[id];
   }
   double r = sqrt(dr2);
   (void)hist; (void)hist_out;
   if (T0.molecule == 1 || T1.molecule == 1) {
      int sgn = 1;
      int tm = offset1 + it / o->R;
      const double *cs = nullptr;
      if (T0.molecule == 1) { sgn = -1; tm = offset0 + it / o->R; cs = cos0; }
      double cost = 0.0;
      for (int id = 0; id < 3; id++) cost += (cs ? cs[id] : o->cosine[id][tm]) * dr[id];
      cost /= r;
      cost *= sgn;
      return LPot2D(o, r, cost, nullptr, nullptr);
   } else if ((T0.molecule == 2 || T1.molecule == 2) && o->sys.ispher == 0 && T0.molecule != T1.molecule) {
      double RCOM[3], Rpt[3], Eul[3];
      int tm;
      if (T0.molecule == 2) {
         tm = offset0 + it / o->R;
         for (int id = 0; id < 3; id++) { RCOM[id] = pos0[id]; Rpt[id] = o->coords[id][t1]; }
         if (eul0) { Eul[0] = eul0[0]; Eul[1] = eul0[1]; Eul[2] = eul0[2]; }
         else { Eul[PHI] = o->angles[PHI][tm]; Eul[CTH] = acos(o->angles[CTH][tm]); Eul[CHI] = o->angles[CHI][tm]; }
      } else {
         tm = offset1 + it / o->R;
         for (int id = 0; id < 3; id++) { Rpt[id] = pos0[id]; RCOM[id] = o->coords[id][t1]; }
         Eul[PHI] = o->angles[PHI][tm]; Eul[CTH] = acos(o->angles[CTH][tm]); Eul[CHI] = o->angles[CHI][tm];
      }
      return vcord_call(o, Eul, RCOM, Rpt, nullptr);
   } else if ((T0.molecule == 2 || T1.molecule == 2) && o->sys.ispher == 1 && T0.molecule != T1.molecule) {
      double radret = r, v;
      vspher_(&radret, &v);
      return v;
   } else if (T0.molecule == 2 && T1.molecule == 2 && o->sys.type[o->imtype].numb > 1) {
      double c1[3], c2[3], e1[3], e2[3], E;
      for (int id = 0; id < 3; id++) { c1[id] = pos0[id]; c2[id] = o->coords[id][t1]; }
      int tm0 = offset0 + it / o->R, tm1 = offset1 + it / o->R;
      if (eul0) { e1[0] = eul0[0]; e1[1] = eul0[1]; e1[2] = eul0[2]; }
      else { e1[PHI] = o->angles[PHI][tm0]; e1[CTH] = acos(o->angles[CTH][tm0]); e1[CHI] = o->angles[CHI][tm0]; }
      e2[PHI] = o->angles[PHI][tm1]; e2[CTH] = acos(o->angles[CTH][tm1]); e2[CHI] = o->angles[CHI][tm1];
      caleng_(c1, c2, &E, e1, e2);
      return E;
   }
   return SPot1D(o, r, nullptr);
}

// PotEnergy(atom0,pos,it), mc_piqmc.cc:1796-1965
static double PotEnergy_it(orc_t *o, int atom0, const double *pos0, int it)
{
   double spot = 0.0;
   for (int atom1 = 0; atom1 < o->N; atom1++)
      if (atom1 != atom0 && partner_on_line(o, atom1, it)) spot += pair_energy(o, atom0, pos0, atom1, it, nullptr, nullptr, 0, nullptr);
   return spot;
}
// PotEnergy(atom0,pos), mc_piqmc.cc:1201-1383: per partner, sum over slices, then add
static double PotEnergy_path(orc_t *o, int atom0, const double *shift)
{
   double spot = 0.0;
   int P = o->P;
   for (int atom1 = 0; atom1 < o->N; atom1++)
      if (atom1 != atom0) {
         double spot_pair = 0.0;
         for (int it = 0; it < P; it++) {
            if (!partner_on_line(o, atom1, it)) continue;
            double pos0[3];
            for (int id = 0; id < 3; id++) {
               pos0[id] = o->coords[id][P * atom0 + it];
               if (shift) pos0[id] += shift[id];
            }
            spot_pair += pair_energy(o, atom0, pos0, atom1, it, nullptr, nullptr, 0, nullptr);
         }
         spot += spot_pair;
      }
   return spot;
}
// PotRotEnergy, mc_piqmc.cc:1967-2044 (linear rotor; `cosine` row passed explicitly)
static double PotRotEnergy(orc_t *o, int atom0, const double *cos3, int it)
{
   double spot = 0.0;
   int P = o->P;
   double pos0[3];
   for (int id = 0; id < 3; id++) pos0[id] = o->coords[id][P * atom0 + it];
   for (int atom1 = 0; atom1 < o->N; atom1++)
      if (atom1 != atom0 && partner_on_line(o, atom1, it)) spot += pair_energy(o, atom0, pos0, atom1, it, nullptr, cos3, 0, nullptr);
   return spot;
}
// PotRotE3D, mc_piqmc.cc:2046-2151
static double PotRotE3D(orc_t *o, int atom0, const double *eul, int it)
{
   double spot = 0.0;
   int P = o->P;
   double pos0[3];
   for (int id = 0; id < 3; id++) pos0[id] = o->coords[id][P * atom0 + it];
   for (int atom1 = 0; atom1 < o->N; atom1++)
      if (atom1 != atom0 && partner_on_line(o, atom1, it)) spot += pair_energy(o, atom0, pos0, atom1, it, eul, nullptr, 0, nullptr);
   return spot;
}

static void rotden_call(orc_t *o, const double *e1, const double *e2, double *rel, double *rho, double *erot,
                        double *esq, int *istop)
{
   double a[3] = {e1[0], e1[1], e1[2]}, b[3] = {e2[0], e2[1], e2[2]};
   *istop = 0;
   rotden_(a, b, rel, rho, erot, esq, const_cast<double *>(o->rho3), const_cast<double *>(o->erot3),
           const_cast<double *>(o->esq3), istop);
}

// ----------------------------------------------------------------------------
// moves with explicit uniforms
// ----------------------------------------------------------------------------
// MCBisectionMove / MCBisectionMoveExchange for ONE atom, mc_piqmc.cc:194-419.
// u_gauss: pairs (r1,r2) per midpoint per dimension in sampling order; u_acc:
// consumed sequentially, only when deltav >= 0 (rnd3 short-circuit, :270-271).
static int bisection_move(orc_t *o, int type, int atom, int time0, const double *ug, const double *ua, int exch, int *consumed)
{
   int P = o->P;
   const orc_type_t &T = o->sys.type[type];
   double mclambda = o->lambda[type];
   int mclevels = T.levels, seg_size = 1 << mclevels;
   int offset0 = o->type_offset(type) + P * atom, offset1 = offset0;
   int time1 = time0 + seg_size, timep = time1 % P;
   if (exch && timep != time1) offset1 = o->type_offset(type) + P * o->pindex[atom];
   for (int id = 0; id < 3; id++) {
      o->newc[id][offset0 + time0] = o->coords[id][offset0 + time0];
      o->newc[id][offset1 + timep] = o->coords[id][offset1 + timep];
   }
   double bnorm = 1.0 / (mclambda * o->tau);
   bool Accepted = false;
   double pot0 = 0.0, pot1 = 0.0;
   int ig = 0, ia = 0;
   for (int level = 0; level < mclevels; level++) {
      int lss = (int)pow(2.0, (mclevels - level));
      double bkin_norm = bnorm / (double)lss;
      double bpot_norm = o->tau * (double)(lss / 2);
      pot1 = pot0; pot0 = 0.0;
      int t0, t1, t2 = 0;
      do {
         t0 = t2; t2 = t0 + lss; t1 = (t0 + t2) / 2;
         int pt0 = (time0 + t0) % P, pt1 = (time0 + t1) % P, pt2 = (time0 + t2) % P;
         int off0 = offset0, off1 = offset0, off2 = offset0;
         if (exch) {
            if (pt0 != (time0 + t0)) off0 = offset1;
            if (pt1 != (time0 + t1)) off1 = offset1;
            if (pt2 != (time0 + t2)) off2 = offset1;
         }
         for (int id = 0; id < 3; id++) {
            o->newc[id][off1 + pt1] = 0.5 * (o->newc[id][off0 + pt0] + o->newc[id][off2 + pt2]);
            o->newc[id][off1 + pt1] += gauss_u(bkin_norm, ug[ig], ug[ig + 1]);
            ig += 2;
         }
         // the reference evaluates with gatom0 = off0/P even when pt1 wrapped (:373-379)
         int gatom0 = off0 / P;
         double pn[3], po[3];
         for (int id = 0; id < 3; id++) { pn[id] = o->newc[id][P * gatom0 + pt1]; po[id] = o->coords[id][P * gatom0 + pt1]; }
         pot0 += PotEnergy_it(o, gatom0, pn, pt1) - PotEnergy_it(o, gatom0, po, pt1);
         if (t0 != 0) {
            for (int id = 0; id < 3; id++) { pn[id] = o->newc[id][P * gatom0 + pt0]; po[id] = o->coords[id][P * gatom0 + pt0]; }
            pot0 += PotEnergy_it(o, gatom0, pn, pt0) - PotEnergy_it(o, gatom0, po, pt0);
         }
      } while (t2 < seg_size);
      double deltav = (pot0 - 2.0 * pot1);
      deltav *= bpot_norm;
      Accepted = false;
      if (deltav < 0.0) Accepted = true;
      else if (exp(-deltav) > ua[ia++]) Accepted = true;
      if (!Accepted) break;
   }
   if (consumed) { consumed[0] = ig; consumed[1] = ia; }
   o->mctotal[type][MCMULTI] += 1.0;
   if (Accepted) {
      o->mcaccep[type][MCMULTI] += 1.0;
      for (int id = 0; id < 3; id++)
         for (int it = time0; it <= time1; it++) {
            int pit = it % P;
            int offset = offset0;
            if (exch && pit != it) offset = offset1;
            o->coords[id][offset + pit] = o->newc[id][offset + pit];
         }
   }
   return Accepted ? 1 : 0;
}

// MCMolecularMove for ONE atom, mc_piqmc.cc:54-102
static int molecular_move(orc_t *o, int type, int atom, const double *u3, double uacc)
{
   int P = o->P;
   int offset = o->type_offset(type) + P * atom;
   int gatom = offset / P;
   double disp[3];
   for (int id = 0; id < 3; id++) disp[id] = o->sys.type[type].mcstep * (u3[id] - 0.5);
   double deltav = 0.0;
   deltav += (PotEnergy_path(o, gatom, disp) - PotEnergy_path(o, gatom, nullptr));
   bool Accepted = false;
   if (deltav < 0.0) Accepted = true;
   else if (exp(-deltav * o->tau) > uacc) Accepted = true;
   o->mctotal[type][MCMOLEC] += 1.0;
   if (Accepted) {
      o->mcaccep[type][MCMOLEC] += 1.0;
      for (int id = 0; id < 3; id++)
         for (int it = 0; it < P; it++) {
            // the reference forms newcoords = MCCoords; newcoords += disp (:73-74)
            double v = o->coords[id][offset + it];
            v += disp[id];
            o->coords[id][offset + it] = v;
         }
   }
   return Accepted ? 1 : 0;
}

// MCRot3Dstep, mc_piqmc.cc:938-1199 (RotDenType 0 and 1)
static int rot3d_step(orc_t *o, int it1, int atom0, int type, double rand1, double rand2, double rand3, double rand4)
{
   int P = o->P, Q = o->Q;
   int offset = o->type_offset(type) + P * atom0;
   int gatom = offset / P;
   double step = o->sys.type[type].rtstep;
   int it0 = it1 - 1, it2 = it1 + 1;
   if (it0 < 0) it0 += Q;
   if (it2 >= Q) it2 -= Q;
   int t0 = offset + it0, t1 = offset + it1, t2 = offset + it2;
   double cost = o->angles[CTH][t1], phi = o->angles[PHI][t1], chi = o->angles[CHI][t1];
   cost += (step * (rand1 - 0.5));
   phi += 2.0 * M_PI * (step * (rand2 - 0.5));
   chi += 2.0 * M_PI * (step * (rand3 - 0.5));
   if (phi < 0.0) phi = 2.0 * M_PI + phi;
   if (chi < 0.0) chi = 2.0 * M_PI + chi;
   phi = fmod(phi, 2.0 * M_PI);
   chi = fmod(chi, 2.0 * M_PI);
   if (cost > 1.0) cost = 2.0 - cost;
   if (cost < -1.0) cost = -2.0 - cost;

   double rho, erot, esq, Eul1[3], Eul2[3], Eulrel[3];
   int istop = 0;
   auto dens = [&](const double *a, const double *b) -> double {
      if (o->sys.rotden_type == 0) {
         rotden_call(o, a, b, Eulrel, &rho, &erot, &esq, &istop);
         if (istop == 1) { fprintf(stderr, "large matrix test error\n"); exit(0); }
      } else {
         double A[3] = {a[0], a[1], a[2]}, B[3] = {b[0], b[1], b[2]};
         rsrot_(A, B, &o->sys.x_rot, &o->sys.y_rot, &o->sys.z_rot, &o->rottau, &o->sys.rot_odevn, &o->sys.rot_eoff, &rho, &erot);
      }
      return rho;
   };
   auto eul_old = [&](int t, double *e) { e[0] = o->angles[PHI][t]; e[1] = acos(o->angles[CTH][t]); e[2] = o->angles[CHI][t]; };
   double Enew[3] = {phi, acos(cost), chi};

   eul_old(t0, Eul1); eul_old(t1, Eul2);
   double r = dens(Eul1, Eul2);
   double dens_old = r, rhoold = r;
   eul_old(t1, Eul1); eul_old(t2, Eul2);
   r = dens(Eul1, Eul2);
   dens_old = dens_old * r; rhoold = rhoold + r;
   if (fabs(dens_old) < RZERO) dens_old = 0.0;
   if (dens_old < 0.0) dens_old = fabs(dens_old);

   double pot_old = 0.0;
   int itr0 = it1 * o->R, itr1 = itr0 + o->R;
   for (int it = itr0; it < itr1; it++) pot_old += PotRotE3D(o, gatom, Eul1, it);

   eul_old(t0, Eul1);
   r = dens(Eul1, Enew);
   double dens_new = r, rhonew = r;
   eul_old(t2, Eul2);
   r = dens(Enew, Eul2);
   dens_new = dens_new * r; rhonew = rhonew + r;
   if (fabs(dens_new) < RZERO) dens_new = 0.0;
   if (dens_new < 0.0) dens_new = fabs(dens_new);

   double pot_new = 0.0;
   for (int it = itr0; it < itr1; it++) pot_new += PotRotE3D(o, gatom, Enew, it);

   double rd;
   bool Accepted = false;
   if (o->sys.rotden_type == 0) {
      if (dens_old > RZERO) rd = dens_new / dens_old; else rd = 1.0;
      rd *= exp(-o->tau * (pot_new - pot_old));
      if (rd > 1.0) Accepted = true; else if (rd > rand4) Accepted = true;
   } else {
      rd = (rhonew - rhoold) / (4.0 * (o->rottau / WNO2K));
      rd -= o->tau * (pot_new - pot_old);
      if (rd > 0.0) Accepted = true; else if (rd > log(rand4)) Accepted = true;
   }
   o->mctotal[type][MCROTAT] += 1.0;
   if (Accepted) {
      o->mcaccep[type][MCROTAT] += 1.0;
      o->angles[CTH][t1] = cost; o->angles[PHI][t1] = phi; o->angles[CHI][t1] = chi;
      double sint = sqrt(1.0 - cost * cost);
      o->cosine[0][t1] = sint * cos(phi);
      o->cosine[1][t1] = sint * sin(phi);
      o->cosine[2][t1] = cost;
   }
   return Accepted ? 1 : 0;
}

// MCRotLinStep, mc_piqmc.cc:781-936
static int rotlin_step(orc_t *o, int it1, int type, double rand1, double rand2, double rand3)
{
   int P = o->P, Q = o->Q;
   int offset = o->type_offset(type);
   int gatom = offset / P;
   double step = o->sys.type[type].rtstep;
   int it0 = it1 - 1, it2 = it1 + 1;
   if (it0 < 0) it0 += Q;
   if (it2 >= Q) it2 -= Q;
   int t0 = offset + it0, t1 = offset + it1, t2 = offset + it2;
   double cost = o->angles[CTH][t1], phi = o->angles[PHI][t1];
   cost += (step * (rand1 - 0.5));
   phi += (step * (rand2 - 0.5));
   if (cost > 1.0) cost = 2.0 - cost;
   if (cost < -1.0) cost = -2.0 - cost;
   double sint = sqrt(1.0 - cost * cost);
   double nn[3] = {sint * cos(phi), sint * sin(phi), cost};

   double p0 = 0.0, p1 = 0.0;
   for (int id = 0; id < 3; id++) { p0 += o->cosine[id][t0] * o->cosine[id][t1]; p1 += o->cosine[id][t1] * o->cosine[id][t2]; }
   double dens_old, rho1, rho2, erot;
   if (o->sys.rotden_type == 0) dens_old = SRot(o, p0, 0) * SRot(o, p1, 0);
   else {
      rsline_(&o->sys.x_rot, &p0, &o->rottau, &rho1, &erot);
      rsline_(&o->sys.x_rot, &p1, &o->rottau, &rho2, &erot);
      dens_old = rho1 + rho2;
   }
   if (fabs(dens_old) < RZERO) dens_old = 0.0;
   if (dens_old < 0.0 && o->sys.rotden_type == 0) { printf("Rotational Moves: Negative rot density\n"); exit(1); }

   double pot_old = 0.0;
   int itr0 = it1 * o->R, itr1 = itr0 + o->R;
   double cold[3] = {o->cosine[0][t1], o->cosine[1][t1], o->cosine[2][t1]};
   for (int it = itr0; it < itr1; it++) pot_old += PotRotEnergy(o, gatom, cold, it);

   p0 = 0.0; p1 = 0.0;
   for (int id = 0; id < 3; id++) { p0 += o->cosine[id][t0] * nn[id]; p1 += nn[id] * o->cosine[id][t2]; }
   double dens_new;
   if (o->sys.rotden_type == 0) dens_new = SRot(o, p0, 0) * SRot(o, p1, 0);
   else {
      rsline_(&o->sys.x_rot, &p0, &o->rottau, &rho1, &erot);
      rsline_(&o->sys.x_rot, &p1, &o->rottau, &rho2, &erot);
      dens_new = rho1 + rho2;
   }
   if (fabs(dens_new) < RZERO) dens_new = 0.0;
   if (dens_new < 0.0 && o->sys.rotden_type == 0) { printf("Rotational Moves: Negative rot density\n"); exit(1); }

   double pot_new = 0.0;
   for (int it = itr0; it < itr1; it++) pot_new += PotRotEnergy(o, gatom, nn, it);

   double rd;
   bool Accepted = false;
   if (o->sys.rotden_type == 0) {
      if (dens_old > RZERO) rd = dens_new / dens_old; else rd = 1.0;
      rd *= exp(-o->tau * (pot_new - pot_old));
      if (rd > 1.0) Accepted = true; else if (rd > rand3) Accepted = true;
   } else {
      rd = dens_new - dens_old - o->tau * (pot_new - pot_old);
      if (rd > 0.0) Accepted = true; else if (rd > log(rand3)) Accepted = true;
   }
   o->mctotal[type][MCROTAT] += 1.0;
   if (Accepted) {
      o->mcaccep[type][MCROTAT] += 1.0;
      o->angles[CTH][t1] = cost; o->angles[PHI][t1] = phi;
      for (int id = 0; id < 3; id++) o->cosine[id][t1] = nn[id];
   }
   return Accepted ? 1 : 0;
}

// ----------------------------------------------------------------------------
// estimators
// ----------------------------------------------------------------------------
static void bin_1D(orc_t *o, double r)                         // mc_estim.cc:1275-1286
{
   int bin_r = (int)floor((r - MIN_RADIUS) / o->delta_radius);
   if (bin_r < MC_BINSR && bin_r >= 0) o->gr1d[bin_r] += 1.0;
}
static void bin_2D(orc_t *o, double r, double cost)            // mc_estim.cc:1233-1250
{
   int bin_r = (int)floor((r - MIN_RADIUS) / o->delta_radius);
   if (bin_r < MC_BINSR && bin_r >= 0) {
      double theta = acos(cost);
      int bin_t = (int)floor(theta / o->delta_theta);
      if (bin_t < MC_BINST && bin_t >= 0) o->gr2d[bin_r * MC_BINST + bin_t] += 1.0;
   }
}
static void bin_3D(orc_t *o, double r, double theta, double chi, int dtype)   // mc_estim.cc:1252-1273
{
   int bin_r = (int)floor((r - MIN_RADIUS) / o->delta_radius);
   if (bin_r < MC_BINSR && bin_r >= 0) {
      int bin_t = (int)floor(theta / o->delta_theta);
      if (bin_t < MC_BINST && bin_t >= 0) {
         int bin_c = (int)floor(chi / o->delta_chi);
         if (bin_c < MC_BINSC && bin_c >= 0) o->gr3d[dtype][((size_t)bin_r * MC_BINST + bin_t) * MC_BINSC + bin_c] += 1.0;
      }
   }
}

// GetPotEnergy_Densities / GetPotEnergy, mc_estim.cc:500-874
static double GetPot(orc_t *o, int dens)
{
   int P = o->P, N = o->N, R = o->R;
   double spot = 0.0;
   for (int atom0 = 0; atom0 < N - 1; atom0++)
      for (int atom1 = atom0 + 1; atom1 < N; atom1++) {
         int type0 = o->mctype[atom0], type1 = o->mctype[atom1];
         const orc_type_t &T0 = o->sys.type[type0], &T1 = o->sys.type[type1];
         int offset0 = P * atom0, offset1 = P * atom1;
         double spot_pair = 0.0;
         for (int it = 0; it < P; it++) {
            int t0 = offset0 + it, t1 = offset1 + it;
            double dr[3], dr2 = 0.0;
            for (int id = 0; id < 3; id++) {
               dr[id] = o->coords[id][t0] - o->coords[id][t1];
               if (o->sys.minimage) dr[id] -= o->sys.box[id] * rint(dr[id] / o->sys.box[id]);
               dr2 += dr[id] * dr[id];
            }
            double r = sqrt(dr2);
            if (T0.molecule == 1 || T1.molecule == 1) {
               int sgn = 1, tm = offset1 + it / R;
               if (T0.molecule == 1) { sgn = -1; tm = offset0 + it / R; }
               double cost = 0.0;
               for (int id = 0; id < 3; id++) cost += o->cosine[id][tm] * dr[id];
               cost /= r; cost *= sgn;
               if (dens) bin_2D(o, r, cost);
               spot_pair += LPot2D(o, r, cost, nullptr, nullptr);
            } else if ((T0.molecule == 2 || T1.molecule == 2) && T0.molecule != T1.molecule) {
               int tm, typed;
               double RCOM[3], Rpt[3], Eul[3], v, rtc[3];
               if (T0.molecule == 2) {
                  typed = type1; tm = offset0 + it / R;
                  for (int id = 0; id < 3; id++) { RCOM[id] = o->coords[id][t0]; Rpt[id] = o->coords[id][t1]; }
               } else {
                  typed = type0; tm = offset1 + it / R;
                  for (int id = 0; id < 3; id++) { Rpt[id] = o->coords[id][t0]; RCOM[id] = o->coords[id][t1]; }
               }
               Eul[PHI] = o->angles[PHI][tm]; Eul[CTH] = acos(o->angles[CTH][tm]); Eul[CHI] = o->angles[CHI][tm];
               if (o->sys.ispher == 0) v = vcord_call(o, Eul, RCOM, Rpt, rtc);
               else { rtc[0] = r; vspher_(&rtc[0], &v); rtc[1] = 0.0; rtc[2] = 0.0; }
               if (dens) bin_3D(o, rtc[0], rtc[1], rtc[2], typed);
               spot_pair += v;
            } else if (T0.molecule == 2 && T1.molecule == 2 && o->sys.type[o->imtype].numb > 1) {
               double c1[3], c2[3], e1[3], e2[3], E;
               for (int id = 0; id < 3; id++) { c1[id] = o->coords[id][t0]; c2[id] = o->coords[id][t1]; }
               int tm0 = offset0 + it / R, tm1 = offset1 + it / R;
               e1[PHI] = o->angles[PHI][tm0]; e1[CTH] = acos(o->angles[CTH][tm0]); e1[CHI] = o->angles[CHI][tm0];
               e2[PHI] = o->angles[PHI][tm1]; e2[CTH] = acos(o->angles[CTH][tm1]); e2[CHI] = o->angles[CHI][tm1];
               caleng_(c1, c2, &E, e1, e2);
               spot_pair += E;
            } else if (type0 == type1 && T0.molecule == 0) {
               if (dens) bin_1D(o, r);
               spot_pair += SPot1D(o, r, nullptr);
            }
         }
         spot += spot_pair;
      }
   return spot / (double)P;
}

// GetKinEnergy, mc_estim.cc:876-937
static double GetKin(orc_t *o)
{
   int P = o->P, N = o->N;
   int numb = 0;
   double r2avr = 0.0;
   for (int atom = 0; atom < N; atom++) {
      numb++;
      int type = o->mctype[atom];
      int offset0 = P * atom;
      int gatom = o->offset_atom[type];
      double sum = 0.0;
      for (int it = 0; it < P; it++) {
         int t0 = offset0 + it;
         int offset1 = offset0;
         if (o->sys.type[type].stat == 1 && (it + 1) == P) offset1 = P * (gatom + o->pindex[atom - gatom]);
         int t1 = offset1 + (it + 1) % P;
         for (int dim = 0; dim < 3; dim++) {
            double dr = o->coords[dim][t0] - o->coords[dim][t1];
            if (o->sys.minimage) dr -= o->sys.box[dim] * rint(dr / o->sys.box[dim]);
            sum += dr * dr;
         }
      }
      r2avr += sum / (4.0 * o->beta * o->lambda[type]);
   }
   return (double)P * o->sys.temperature * (0.5 * (double)(3 * numb) - r2avr);
}

// GetRotEnergy, mc_estim.cc:939-987
static double GetRotEnergy(orc_t *o)
{
   int type = o->imtype, Q = o->Q;
   int offset = o->type_offset(type);
   double srot = 0.0;
   o->ErotSQ = 0.0; o->Erot_termSQ = 0.0;
   for (int it0 = 0; it0 < Q; it0++) {
      int t0 = offset + it0, t1 = offset + (it0 + 1) % Q;
      double p0 = 0.0;
      for (int id = 0; id < 3; id++) p0 += o->cosine[id][t0] * o->cosine[id][t1];
      if (o->sys.rotden_type == 0) {
         double rdens = SRot(o, p0, 0);
         if (fabs(rdens) > RZERO) srot += SRot(o, p0, 1) / rdens;
         o->Erot_termSQ += (SRot(o, p0, 1) / rdens) * (SRot(o, p0, 1) / rdens);
         o->ErotSQ += SRot(o, p0, 2) / rdens;
      } else {
         double rho, erot;
         rsline_(&o->sys.x_rot, &p0, &o->rottau, &rho, &erot);
         srot += rho;
      }
   }
   if (o->sys.rotden_type == 1) {
      srot = srot / (double)Q;
      srot = srot / o->rottau + 1.0 / o->rottau;
   }
   return srot;
}

// GetRotE3D, mc_estim.cc:989-1096 (including the accumulating `offset +=` at :1002)
static double GetRotE3D(orc_t *o)
{
   int type = o->imtype, P = o->P, Q = o->Q;
   int offset = o->type_offset(type);
   double ERot3D = 0.0;
   o->ErotSQ = 0.0; o->Erot_termSQ = 0.0;
   for (int atom = 0; atom < o->sys.type[type].numb; atom++) {
      offset += P * atom;
      double srot = 0.0, sesq = 0.0, se_termsq = 0.0;
      int RNskip = (o->sys.rotden_type == 0) ? 1 : o->sys.rnratio;
      for (int it0 = 0; it0 < Q; it0 = it0 + RNskip) {
         int t0 = offset + it0, t1 = offset + (it0 + RNskip) % Q;
         double rho, erot, esq, Eul1[3], Eul2[3], Eulrel[3];
         int istop = 0;
         Eul1[0] = o->angles[PHI][t0]; Eul1[1] = acos(o->angles[CTH][t0]); Eul1[2] = o->angles[CHI][t0];
         Eul2[0] = o->angles[PHI][t1]; Eul2[1] = acos(o->angles[CTH][t1]); Eul2[2] = o->angles[CHI][t1];
         rotden_call(o, Eul1, Eul2, Eulrel, &rho, &erot, &esq, &istop);
         double phirel = Eulrel[0], therel = Eulrel[1], chirel = Eulrel[2];
         int bin_t = (int)floor(therel / o->delta_theta);
         if (bin_t < MC_BINST && bin_t >= 0) o->relthe[bin_t] += (double)RNskip;
         int bin_p = (int)floor(phirel / o->delta_chi);
         if (bin_p < MC_BINSC && bin_p >= 0) o->relphi[bin_p] += (double)RNskip;
         int bin_c = (int)floor(chirel / o->delta_chi);
         if (bin_c < MC_BINSC && bin_c >= 0) o->relchi[bin_c] += (double)RNskip;
         if (o->sys.rotden_type == 1 && o->sys.rnratio == 1) {
            rsrot_(Eul1, Eul2, &o->sys.x_rot, &o->sys.y_rot, &o->sys.z_rot, &o->rottau, &o->sys.rot_odevn, &o->sys.rot_eoff, &rho, &erot);
            srot += rho;
         } else srot += erot;
         sesq += esq;
         se_termsq += erot * erot;
      }
      double nq = (double)(Q / RNskip);
      srot = srot / nq;
      sesq = sesq / (nq * nq);
      se_termsq = se_termsq / (nq * nq);
      ERot3D += srot; o->ErotSQ += sesq; o->Erot_termSQ += se_termsq;
      if (o->sys.rotden_type == 1 && o->sys.rnratio == 1) {
         double tc = o->rottau / WNO2K;
         ERot3D = ERot3D / (4.0 * tc * tc);
         ERot3D += 0.25 * (o->sys.x_rot + o->sys.y_rot + o->sys.z_rot) + 1.5 / tc;
         ERot3D = ERot3D / WNO2K;
      }
   }
   return ERot3D;
}

// GetRCF row 0, mc_estim.cc:1099-1139
static void GetRCF(orc_t *o, double *rcf0)
{
   int Q = o->Q, offset = o->type_offset(o->imtype);
   for (int it0 = 0; it0 < Q; it0++) {
      int t0 = offset + it0;
      for (int itc = 0; itc < Q; itc++) {
         int tc = offset + (it0 + itc) % Q;
         double p0 = 0.0;
         for (int id = 0; id < 3; id++) p0 += o->cosine[id][t0] * o->cosine[id][tc];
         rcf0[itc] += p0;
      }
   }
}


// ----------------------------------------------------------------------------
// worm moves (N1), mc_qworm.cc:93-667.  Uniforms come from per-stream queues in
// test mode (same stream numbers as mc_randg.cc:90-174) or, in schedule mode,
// all from the chain's worm stream in program order (what the device does).
// ----------------------------------------------------------------------------
namespace {
inline double draw(orc_t *o, int s);
double wr(orc_t *o, int stream)
{
   if (o->wmode == 1) return draw(o, o->P + o->Q + 1);
   if (o->wq[stream].empty()) { fprintf(stderr, "oracle: worm RNG queue %d empty\n", stream); exit(2); }
   double v = o->wq[stream].front(); o->wq[stream].pop_front(); return v;
}
inline double w_gauss(orc_t *o, double alpha)        // mc_randg.cc:138-150
{
   double r1 = wr(o, 8), r2 = wr(o, 9);
   double x1 = sqrt(-log(r1)) * cos(2.0 * M_PI * r2);
   return (x1 / sqrt(alpha));
}
inline int w_nrnd(orc_t *o, int k, int n) { return (int)floor(n * wr(o, 9 + k)); }   // nrnd1..3 -> streams 10..12

// get_potential, mc_qworm.cc:400-422: sum of PotEnergy over the open interval (it0, it1); the moving atom's beads are
// read from `pos` (MCCoords or the swap path)
typedef std::vector<double> *coord3;
double get_potential(orc_t *o, int it0, int it1, int atom0, int atom1, std::vector<double> *coords)
{
   int P = o->P, aoff = o->offset_atom[o->worm.type];
   int pit0 = it0 % P;
   double pot = 0.0;
   int atom = atom0;
   for (int it = (it0 + 1); it < it1; it++) {
      int pit = it % P;
      if ((pit != it) && (pit0 == it0)) atom = atom1;
      double pos0[3];
      for (int id = 0; id < 3; id++) pos0[id] = coords[id][P * (aoff + atom) + pit];
      pot += PotEnergy_it(o, aoff + atom, pos0, pit);
   }
   return pot;
}
// sample_middle, mc_qworm.cc:240-287
void sample_middle(orc_t *o, int it0, int it2, int atom0, int atom2, std::vector<double> *coords)
{
   if ((it2 - it0) < 2) return;
   int P = o->P;
   int it1 = (int)rint(0.5 * (double)(it0 + it2));
   int pt0 = it0 % P, pt1 = it1 % P, pt2 = it2 % P;
   int atom1 = atom0;
   if ((pt1 != it1) && (pt0 == it0)) atom1 = atom2;
   int offset = o->type_offset(o->worm.type);
   pt0 += (offset + atom0 * P); pt1 += (offset + atom1 * P); pt2 += (offset + atom2 * P);
   double s0 = (double)(it1 - it0), s2 = (double)(it2 - it1);
   double gkin = (s0 + s2) / (o->worm.twave2 * s0 * s2);
   for (int id = 0; id < 3; id++) {
      coords[id][pt1] = (s2 * coords[id][pt0] + s0 * coords[id][pt2]) / (s0 + s2);
      coords[id][pt1] += w_gauss(o, gkin);
   }
   sample_middle(o, it0, it1, atom0, atom1, coords);
   sample_middle(o, it1, it2, atom1, atom2, coords);
}
// qw_open_prob, mc_qworm.cc:127-153
double qw_open_prob(orc_t *o, int segm)
{
   auto &W = o->worm;
   int P = o->P, offset = o->type_offset(W.type);
   double kin = 0.0;
   int pt0 = offset + W.atom_i * P + W.ira, pt1 = offset + W.atom_m * P + W.masha;
   for (int id = 0; id < 3; id++) {
      double dr = o->coords[id][pt0] - o->coords[id][pt1];
      if (o->sys.minimage) dr -= (o->sys.box[id] * rint(dr / o->sys.box[id]));
      kin += (dr * dr);
   }
   kin /= (W.twave2 * (double)segm);
   double pot = get_potential(o, W.ira, W.ira + segm, W.atom_i, W.atom_m, o->coords);
   pot *= o->tau;
   return (W.qw_norm * pow((double)segm, 0.5 * 3.0) * exp(kin + pot));
}
void qworm_open(orc_t *o)            // mc_qworm.cc:155-182
{
   auto &W = o->worm;
   int P = o->P;
   o->qwtotal[0] += 1.0;
   W.atom_i = w_nrnd(o, 1, o->sys.type[W.type].numb);
   W.ira = w_nrnd(o, 2, P);
   int segm = w_nrnd(o, 3, W.m) + 1;
   W.masha = (W.ira + segm) % P;
   W.atom_m = W.atom_i;
   if (W.masha != (W.ira + segm)) W.atom_m = o->pindex[W.atom_i];
   double prob = qw_open_prob(o, segm);
   bool Accepted = false;
   if (prob >= 1.0) Accepted = true;
   else if (prob > wr(o, 1)) Accepted = true;
   if (Accepted) { W.exists = 1; o->qwaccep[0] += 1.0; }
}
void qworm_close(orc_t *o)           // mc_qworm.cc:184-238
{
   auto &W = o->worm;
   int P = o->P;
   o->qwtotal[1] += 1.0;
   int segm = W.masha - W.ira;
   if (segm < 0) segm += P;
   if (segm > W.m) return;
   int it0 = W.ira, it2 = W.ira + segm;
   sample_middle(o, it0, it2, W.atom_i, W.atom_m, o->coords);
   double prob = 1.0 / qw_open_prob(o, segm);
   bool Accepted = false;
   if (prob >= 1.0) Accepted = true;
   else if (prob > wr(o, 1)) Accepted = true;
   if (Accepted) { W.exists = 0; o->qwaccep[1] += 1.0; }
}
void qworm_advance(orc_t *o)         // mc_qworm.cc:299-357
{
   auto &W = o->worm;
   int P = o->P;
   o->qwtotal[4] += 1.0;
   int segm = W.masha - W.ira;
   if (segm < 0) segm += P;
   int advance = w_nrnd(o, 3, W.m) + 1;
   if (segm - advance <= 0) return;
   int type = W.type, offset = o->type_offset(type);
   int it0 = W.ira, it2 = W.ira + advance;
   int ira_new = it2 % P, atom_i_new = W.atom_i;
   if (ira_new != it2) atom_i_new = W.atom_m;
   double gvar = 1.0 / ((double)advance * W.twave2);
   int pt0 = offset + W.atom_i * P + it0 % P, pt2 = offset + atom_i_new * P + it2 % P;
   for (int id = 0; id < 3; id++) o->coords[id][pt2] = o->coords[id][pt0] + w_gauss(o, gvar);
   sample_middle(o, it0, it2, W.atom_i, atom_i_new, o->coords);
   double pot = get_potential(o, it0, it2 + 1, W.atom_i, atom_i_new, o->coords);
   bool Accepted = false;
   if (pot < 0.0) Accepted = true;
   else if (exp(-pot * o->tau) > wr(o, 2)) Accepted = true;
   if (Accepted) { o->qwaccep[4] += 1.0; W.ira = ira_new; W.atom_i = atom_i_new; }
}
void qworm_recede(orc_t *o)          // mc_qworm.cc:359-398
{
   auto &W = o->worm;
   int P = o->P;
   o->qwtotal[5] += 1.0;
   int segm = W.ira - W.masha;
   if (segm < 0) segm += P;
   int recede = w_nrnd(o, 3, W.m) + 1;
   if ((segm - recede) < 1) return;
   int it0 = (W.ira - recede), it1 = W.ira;
   int atom0 = W.atom_i, atom1 = W.atom_i;
   if (it0 < 0) { it0 += P; it1 += P; atom0 = o->rindex[atom1]; }
   double pot = get_potential(o, it0, it1 + 1, atom0, atom1, o->coords);
   bool Accepted = false;
   if (pot > 0.0) Accepted = true;
   else if (exp(pot * o->tau) > wr(o, 2)) Accepted = true;
   if (Accepted) { W.ira = it0 % P; W.atom_i = atom0; o->qwaccep[5] += 1.0; }
}
// get_ptable, mc_qworm.cc:577-643 (entries 1..count)
int get_ptable(orc_t *o, int atomw, int pt0, int pt1, int segm, int t1)
{
   auto &W = o->worm;
   int P = o->P, type = W.type, offset = o->type_offset(type);
   int itw = offset + atomw * P + pt0;
   int count = 0;
   for (int atom1 = 0; atom1 < o->sys.type[type].numb; atom1++)
      if (WorldLine(o, atom1, pt1)) {
         int atom0 = atom1;
         if (t1 != pt1) atom0 = o->rindex[atom1];
         if (atom0 != W.atom_i) {
            int it1 = offset + atom1 * P + pt1;
            double dr2 = 0.0;
            for (int id = 0; id < 3; id++) {
               double dx = o->coords[id][itw] - o->coords[id][it1];
               if (o->sys.minimage) dx -= (o->sys.box[id] * rint(dx / o->sys.box[id]));
               dr2 += (dx * dx);
            }
            if (dr2 < W.cutoff2) { count++; o->dr2_list[count] = dr2; o->atm_list[count] = atom1; }
         }
      }
   for (int j = 2; j <= count; j++) {          // mmsort, mc_utils.cc:206-231
      double dtmp = o->dr2_list[j]; int itmp = o->atm_list[j];
      int i = j - 1;
      while ((i > 0) && (o->dr2_list[i] > dtmp)) { o->dr2_list[i + 1] = o->dr2_list[i]; o->atm_list[i + 1] = o->atm_list[i]; i--; }
      o->dr2_list[i + 1] = dtmp; o->atm_list[i + 1] = itmp;
   }
   if (count > 100) count = 100;               // MAXNEIGHBORS
   double norm = 1.0 / ((double)segm * W.twave2);
   for (int ic = 1; ic <= count; ic++) o->ptable[ic] = exp(-norm * o->dr2_list[ic]);
   return count;
}
int atom2swap(orc_t *o, int count, double &pnorm)    // mc_qworm.cc:645-667
{
   pnorm = 0.0;
   for (int ic = 1; ic <= count; ic++) pnorm += o->ptable[ic];
   double prand = pnorm * wr(o, 3);
   double sum = 0.0;
   int ic = 1;
   while ((ic <= count) && (sum < prand)) { sum += o->ptable[ic]; ic++; }
   ic--;
   return (o->atm_list[ic]);
}
void qworm_swap(orc_t *o)            // mc_qworm.cc:424-551
{
   auto &W = o->worm;
   int P = o->P;
   o->qwtotal[6] += 1.0;
   int segm = W.m;
   int it0 = W.ira, it1 = it0 + segm;
   int pit0 = it0, pit1 = it1 % P;
   int atomw = W.atom_i;
   int count = get_ptable(o, atomw, pit0, pit1, segm, it1);
   if (count <= 0) return;
   double pnorm_old, pnorm_new;
   int atom1 = atom2swap(o, count, pnorm_old);
   if (atom1 < 0) return;
   int atom0 = atom1;
   if (pit1 != it1) atom0 = o->rindex[atom1];
   int type = W.type;
   int offset0 = o->type_offset(type) + atom0 * P, offset1 = o->type_offset(type) + atom1 * P, offsetw = o->type_offset(type) + atomw * P;
   for (int id = 0; id < 3; id++) {
      o->newc[id][offset0 + pit0] = o->coords[id][offsetw + pit0];
      o->newc[id][offset1 + pit1] = o->coords[id][offset1 + pit1];
   }
   sample_middle(o, it0, it1, atom0, atom1, o->newc);
   int gatom0 = offset0 / P, gatom1 = offset1 / P;
   double pot = 0.0;
   int gatom = gatom0;
   for (int it = (it0 + 1); it < it1; it++) {
      int pit = it % P;
      if (pit != it) gatom = gatom1;
      double pn[3], po[3];
      for (int id = 0; id < 3; id++) { pn[id] = o->newc[id][P * gatom + pit]; po[id] = o->coords[id][P * gatom + pit]; }
      pot += PotEnergy_it(o, gatom, pn, pit);
      pot -= PotEnergy_it(o, gatom, po, pit);
   }
   double prob = exp(-pot * o->tau);
   count = get_ptable(o, atom0, pit0, pit1, segm, it1);
   pnorm_new = 0.0;
   for (int ic = 1; ic <= count; ic++) pnorm_new += o->ptable[ic];
   prob *= (pnorm_old / pnorm_new);
   bool Accepted = false;
   if (prob >= 1.0) Accepted = true;
   else if (prob > wr(o, 4)) Accepted = true;
   if (Accepted) {
      o->qwaccep[6] += 1.0;
      for (int id = 0; id < 3; id++) {
         int offset = offset0;
         for (int it = (it0 + 1); it < it1; it++) {
            int pit = it % P;
            if (pit != it) offset = offset1;
            o->coords[id][offset + pit] = o->newc[id][offset + pit];
         }
      }
      for (int id = 0; id < 3; id++)
         for (int it = 0; it <= it0; it++) {
            o->newc[id][offset0 + it] = o->coords[id][offset0 + it];
            o->coords[id][offset0 + it] = o->coords[id][offsetw + it];
            o->coords[id][offsetw + it] = o->newc[id][offset0 + it];
         }
      int ratomw = o->rindex[atomw], ratom0 = o->rindex[atom0];
      o->pindex[ratomw] = atom0; o->rindex[atom0] = ratomw;
      o->pindex[ratom0] = atomw; o->rindex[atomw] = ratom0;
      if (W.ira > W.masha) {
         if (atom0 == W.atom_m) W.atom_m = W.atom_i;
         else if (W.atom_i == W.atom_m) W.atom_m = atom0;
      }
   }
}
// MCWormMove, mc_qworm.cc:93-125
void worm_move(orc_t *o)
{
   auto &W = o->worm;
   for (int atom = 0; atom < o->sys.type[W.type].numb; atom++) {
      o->countqw += 1.0;
      if (W.exists) qworm_close(o); else qworm_open(o);
      if (W.exists) {
         o->countqw += 1.0;
         double r = wr(o, 5);
         if (r > 0.5) qworm_advance(o); else qworm_recede(o);
      }
      if (o->bstype >= 0 && W.type == o->bstype) {
         o->countqw += 1.0;
         if (W.exists) qworm_swap(o);
      }
   }
}
} // namespace

// ----------------------------------------------------------------------------
// area / exchange estimators (a18) and symmetry operations (a19)
// ----------------------------------------------------------------------------
// body axes of a top from its Euler angles: vcord_ with ivcord = 1 (vcord.f:36-44)
static void body_axes(orc_t *o, const double *eul, const double *rcom, double *hx, double *hy, double *hz)
{
   double E[3] = {eul[0], eul[1], eul[2]}, C[3] = {rcom[0], rcom[1], rcom[2]}, Pt[3] = {0, 0, 0};
   double v, rad, the, chi;
   int iv = 1;
   vcord_(E, C, Pt, const_cast<double *>(o->v3d), &o->rg3, &o->thg3, &o->chg3, &o->rvmax, &o->rvmin, &o->rvstep,
          &v, &rad, &the, &chi, hx, hy, hz, &iv);
}

// GetExchangeLength, mc_estim.cc:1997-2019: ploops[cycle length - 1] += 1 for every permutation cycle
static void GetExchangeLength(orc_t *o, double *ploops)
{
   int nb = o->sys.type[o->bstype].numb;
   std::vector<int> flag(nb, 0);
   for (int atom = 0; atom < nb; atom++)
      if (flag[atom] == 0) {
         int clen = 0, patom = o->pindex[atom];
         while (patom != atom) { flag[patom] = 1; patom = o->pindex[patom]; clen++; }
         ploops[clen] += 1.0;
      }
}

// GetAreaEstimators, mc_estim.cc:2087-2250 (reference point: the dopant's centre of mass, option (ii)).
// out[0..3] = area_perp, area_parl, inert_perp, inert_parl (the inertias before the division by NumbTimes)
static void GetAreaEstimators(orc_t *o, double *out4)
{
   int P = o->P;
   int moff = o->type_offset(o->imtype), boff = o->type_offset(o->bstype);
   double area_perp = 0.0, area_parl = 0.0, inert_perp = 0.0, inert_parl = 0.0;
   for (int atom = 0; atom < o->sys.type[o->bstype].numb; atom++)
      for (int it0 = 0; it0 < P; it0++) {
         int it1 = (it0 + 1) % P;
         int pt0 = boff + P * atom, pt1 = pt0;
         if (it1 != (it0 + 1)) pt1 = boff + P * o->pindex[atom];
         pt0 += it0; pt1 += it1;
         double dr0[3], dr1[3], n_parl[3], n_perp[3], area[3], rn0[3], rn1[3];
         for (int d = 0; d < 3; d++) {
            dr0[d] = o->coords[d][pt0] - o->coords[d][moff + it0];
            dr1[d] = o->coords[d][pt1] - o->coords[d][moff + it1];
         }
         int it_rot = it0 / o->R;
         for (int d = 0; d < 3; d++) n_parl[d] = o->cosine[d][moff + it_rot];
         const double zero = 10e-4;
         double tg = 0.0, st = 1.0;
         if (fabs(n_parl[0]) > zero) { tg = n_parl[1] / n_parl[0]; st = sqrt(1.0 + tg * tg); }
         n_perp[0] = tg / st; n_perp[1] = -1.0 / st; n_perp[2] = 0.0;
         area[0] = 0.5 * (dr0[1] * dr1[2] - dr0[2] * dr1[1]);
         area[1] = 0.5 * (dr0[2] * dr1[0] - dr0[0] * dr1[2]);
         area[2] = 0.5 * (dr0[0] * dr1[1] - dr0[1] * dr1[0]);
         for (int d = 0; d < 3; d++) { area_perp += (n_perp[d] * area[d]); area_parl += (n_parl[d] * area[d]); }
         for (int k = 0; k < 2; k++) {
            const double *n = k == 0 ? n_perp : n_parl;
            rn0[0] = n[1] * dr0[2] - n[2] * dr0[1]; rn0[1] = n[2] * dr0[0] - n[0] * dr0[2]; rn0[2] = n[0] * dr0[1] - n[1] * dr0[0];
            rn1[0] = n[1] * dr1[2] - n[2] * dr1[1]; rn1[1] = n[2] * dr1[0] - n[0] * dr1[2]; rn1[2] = n[0] * dr1[1] - n[1] * dr1[0];
            for (int d = 0; d < 3; d++) { if (k == 0) inert_perp += (rn0[d] * rn1[d]); else inert_parl += (rn0[d] * rn1[d]); }
         }
      }
   out4[0] = area_perp; out4[1] = area_parl; out4[2] = inert_perp; out4[3] = inert_parl;
}

// GetAreaEstim3D(iframe), mc_estim.cc:2252-2594: iframe 0 = space-fixed frame about the total centre of mass,
// 1 = dopant-fixed frame about the dopant.  area_proj[3], inert[9] (before the division by NumbTimes)
static void GetAreaEstim3D(orc_t *o, int iframe, double *area_proj, double *inert)
{
   int P = o->P, N = o->N;
   std::vector<double> com[3];
   for (int d = 0; d < 3; d++) {
      com[d].assign(P, 0.0);
      for (int it = 0; it < P; it++) {
         if (iframe == 0) {
            double tmass = 0.0, c = 0.0;
            for (int atom = 0; atom < N; atom++) {
               double mass = o->sys.type[o->mctype[atom]].mass;
               c += (mass * o->coords[d][atom * P + it]);
               tmass += mass;
            }
            com[d][it] = c / tmass;
         } else com[d][it] = o->coords[d][o->type_offset(o->imtype) + it];
      }
   }
   int boff = o->type_offset(o->bstype);
   double bmass = o->sys.type[o->bstype].mass;
   double ap[3] = {0, 0, 0}, ic[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
   for (int it0 = 0; it0 < P; it0++) {
      double hat[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
      if (iframe == 1) {
         int it_rot = it0 / o->R + o->type_offset(o->imtype);
         double E[3] = {o->angles[PHI][it_rot], acos(o->angles[CTH][it_rot]), o->angles[CHI][it_rot]}, C0[3] = {0, 0, 0};
         body_axes(o, E, C0, hat[0], hat[1], hat[2]);
      }
      for (int atom = 0; atom < o->sys.type[o->bstype].numb; atom++) {
         int it1 = (it0 + 1) % P;
         int pt0 = boff + P * atom, pt1 = pt0;
         if (it1 != (it0 + 1)) pt1 = boff + P * o->pindex[atom];
         pt0 += it0; pt1 += it1;
         double dr0[3], dr1[3], area[3], rn0[3], rn1[3];
         for (int d = 0; d < 3; d++) { dr0[d] = o->coords[d][pt0] - com[d][it0]; dr1[d] = o->coords[d][pt1] - com[d][it1]; }
         area[0] = 0.5 * (dr0[1] * dr1[2] - dr0[2] * dr1[1]);
         area[1] = 0.5 * (dr0[2] * dr1[0] - dr0[0] * dr1[2]);
         area[2] = 0.5 * (dr0[0] * dr1[1] - dr0[1] * dr1[0]);
         for (int id = 0; id < 3; id++) { ap[0] += area[id] * hat[0][id]; ap[1] += area[id] * hat[1][id]; ap[2] += area[id] * hat[2][id]; }
         for (int id = 0; id < 3; id++) {
            const double *h = hat[id];
            rn0[0] = dr0[1] * h[2] - dr0[2] * h[1]; rn0[1] = dr0[2] * h[0] - dr0[0] * h[2]; rn0[2] = dr0[0] * h[1] - dr0[1] * h[0];
            rn1[0] = dr1[1] * h[2] - dr1[2] * h[1]; rn1[1] = dr1[2] * h[0] - dr1[0] * h[2]; rn1[2] = dr1[0] * h[1] - dr1[1] * h[0];
            double sum = 0.0;
            for (int d = 0; d < 3; d++) sum += rn0[d] * rn1[d] * bmass;
            ic[id * 3 + id] += sum;
            double dr0_id = 0.0;
            for (int d = 0; d < 3; d++) dr0_id += h[d] * dr0[d];
            for (int jd = 0; jd < 3; jd++)
               if (jd != id) {
                  double dr1_jd = 0.0;
                  for (int d = 0; d < 3; d++) dr1_jd += hat[jd][d] * dr1[d];
                  ic[id * 3 + jd] += -bmass * dr0_id * dr1_jd;
               }
         }
      }
   }
   for (int i = 0; i < 3; i++) area_proj[i] = ap[i];
   for (int i = 0; i < 9; i++) inert[i] = ic[i];
}

// Reflect_MF_XZ / _YZ / _XY, mc_piqmc.cc:1385-1708: plane 0 = XZ (REFLECTY, rflmfy), 1 = YZ (REFLECTX, rflmfx),
// 2 = XY (REFLECTZ, rflmfz).  Euler angles of every rotor slice are re-extracted from the flipped body axes, the
// y coordinate of every bead is negated; MCCosine is NOT refreshed (as in the reference).
static void Reflect_MF(orc_t *o, int plane)
{
   int P = o->P, type = o->imtype;
   for (int molec = 0; molec < o->sys.type[type].numb; molec++) {
      int offset = o->type_offset(type) + molec * P;
      for (int it_rot = 0; it_rot < P / o->R; it_rot++) {
         int pMF = it_rot + offset;
         double RCOM[3] = {o->coords[0][pMF], o->coords[1][pMF], o->coords[2][pMF]};
         double E[3] = {o->angles[PHI][pMF], acos(o->angles[CTH][pMF]), o->angles[CHI][pMF]};
         double hx[3], hy[3], hz[3];
         body_axes(o, E, RCOM, hx, hy, hz);
         if (plane == 0) rflmfy_(RCOM, hx, hy, hz, E);
         else if (plane == 1) rflmfx_(RCOM, hx, hy, hz, E);
         else rflmfz_(RCOM, hx, hy, hz, E);
         o->angles[PHI][pMF] = E[0];
         o->angles[CTH][pMF] = cos(E[1]);
         o->angles[CHI][pMF] = E[2];
      }
   }
   for (size_t i = 0; i < o->coords[1].size(); i++) o->coords[1][i] *= -1.0;
}

// RotSymConfig, mc_piqmc.cc:1710-1794: one rotor, picked by `rand`, is turned by its symmetry operation
// (top: chi += 2 pi / nfold; linear: n -> -n)
static void RotSymConfig(orc_t *o, double rand, int nfold)
{
   int P = o->P, type = o->imtype, numb = o->sys.type[type].numb;
   for (int molec = 0; molec < numb; molec++)
      if (rand > (double)molec / (double)numb && rand <= (double)(molec + 1) / (double)numb) {
         int offset = o->type_offset(type) + molec * P;
         for (int it_rot = 0; it_rot < P / o->R; it_rot++) {
            int pMF = it_rot + offset;
            if (o->sys.type[type].molecule == 2) {
               double chi = o->angles[CHI][pMF] + 2.0 * M_PI / (double)nfold;
               chi = fmod(chi, 2.0 * M_PI);
               if (chi < 0.0) chi = 2.0 * M_PI + chi;
               o->angles[CHI][pMF] = chi;
            } else if (o->sys.type[type].molecule == 1) {
               double phi = o->angles[PHI][pMF] + M_PI;
               o->angles[CTH][pMF] *= -1.0;
               phi = fmod(phi, 2.0 * M_PI);
               if (phi < 0.0) phi = 2.0 * M_PI + phi;
               o->angles[PHI][pMF] = phi;
               double cost = o->angles[CTH][pMF];
               double sint = sqrt(1.0 - cost * cost);
               o->cosine[0][pMF] = sint * cos(phi);
               o->cosine[1][pMF] = sint * sin(phi);
               o->cosine[2][pMF] = cost;
            }
         }
      }
}

// ----------------------------------------------------------------------------
// device-schedule replay (DESIGN.md "Schedule"); the per-move mathematics is
// the reference's, the ORDER of moves and the stream addressing are the CUDA
// path's.
// ----------------------------------------------------------------------------
namespace {
inline double draw(orc_t *o, int s) { return mrg_u01(o->streams[s].data()); }

// rigid shift of a whole permutation cycle (identity permutation: one atom)
void sched_molecular(orc_t *o, int type)
{
   int P = o->P, numb = o->sys.type[type].numb, base = o->offset_atom[type];
   int MS = P + o->Q;
   std::vector<int> flag(numb, 0);
   for (int atom = 0; atom < numb; atom++) {
      if (flag[atom]) continue;
      std::vector<int> cyc;
      int a = atom;
      do { cyc.push_back(a); flag[a] = 1; a = (o->sys.type[type].stat == 1) ? o->pindex[a] : a; } while (a != atom);
      double u[4];
      for (int k = 0; k < 4; k++) u[k] = draw(o, MS);
      double disp[3];
      for (int id = 0; id < 3; id++) disp[id] = o->sys.type[type].mcstep * (u[id] - 0.5);
      // dV of the rigid shift: members against non-members only
      double deltav = 0.0;
      for (int a0 : cyc) {
         int g0 = base + a0;
         for (int atom1 = 0; atom1 < o->N; atom1++) {
            bool member = false;
            for (int a1 : cyc) if (base + a1 == atom1) member = true;
            if (member) continue;
            for (int it = 0; it < P; it++) {
               if (!partner_on_line(o, atom1, it)) continue;         // world-line mask of PotEnergy(atom,pos), mc_piqmc.cc:1226-1227
               double pn[3], po[3];
               for (int id = 0; id < 3; id++) { po[id] = o->coords[id][P * g0 + it]; pn[id] = po[id] + disp[id]; }
               deltav += pair_energy(o, g0, pn, atom1, it, nullptr, nullptr, 0, nullptr) -
                         pair_energy(o, g0, po, atom1, it, nullptr, nullptr, 0, nullptr);
            }
         }
      }
      bool acc = (deltav < 0.0) || (exp(-deltav * o->tau) > u[3]);
      o->mctotal[type][MCMOLEC] += 1.0;
      if (acc) {
         o->mcaccep[type][MCMOLEC] += 1.0;
         for (int a0 : cyc)
            for (int id = 0; id < 3; id++)
               for (int it = 0; it < P; it++) o->coords[id][P * (base + a0) + it] += disp[id];
      }
   }
}

// one segment [s0, s0+seg] of world line `atom` (continuing on pindex[atom] past beta for BOSE)
void sched_bisect(orc_t *o, int type, int atom, int s0)
{
   int P = o->P;
   const orc_type_t &T = o->sys.type[type];
   int L = T.levels, seg = 1 << L, base = o->offset_atom[type];
   int gA = base + atom;
   int gB = (T.stat == 1) ? base + o->pindex[atom] : gA;
   std::vector<double> nx((seg + 1) * 3);
   auto gat = [&](int t) { return (s0 + t >= P) ? gB : gA; };
   auto sl = [&](int t) { return (s0 + t) % P; };
   for (int id = 0; id < 3; id++) {
      nx[0 * 3 + id] = o->coords[id][P * gat(0) + sl(0)];
      nx[seg * 3 + id] = o->coords[id][P * gat(seg) + sl(seg)];
   }
   double bnorm = 1.0 / (o->lambda[type] * o->tau);
   // unit normals of all interior slices up front, six uniforms per slice from the slice's own stream
   std::vector<double> xi((seg + 1) * 3, 0.0);
   for (int t = 1; t < seg; t++)
      for (int id = 0; id < 3; id++) {
         double r1 = draw(o, sl(t)), r2 = draw(o, sl(t));
         xi[t * 3 + id] = sqrt(-log(r1)) * cos(2.0 * M_PI * r2);
      }
   double S = 0.0;
   bool acc = true;
   for (int level = 0; level < L; level++) {
      int lss = seg >> level, half = lss / 2;
      double bkin = bnorm / (double)lss;
      double D = 0.0;
      for (int t1 = half; t1 < seg; t1 += lss) {
         int p = sl(t1), g = gat(t1);
         double po[3];
         for (int id = 0; id < 3; id++) {
            nx[t1 * 3 + id] = 0.5 * (nx[(t1 - half) * 3 + id] + nx[(t1 + half) * 3 + id]) + xi[t1 * 3 + id] / sqrt(bkin);
            po[id] = o->coords[id][P * g + p];
         }
         D += PotEnergy_it(o, g, &nx[t1 * 3], p) - PotEnergy_it(o, g, po, p);
      }
      double deltav = (D - S) * (o->tau * (double)half);
      S += D;
      acc = false;
      if (deltav < 0.0) acc = true;
      else if (exp(-deltav) > draw(o, sl(0))) acc = true;
      if (!acc) break;
   }
   o->mctotal[type][MCMULTI] += 1.0;
   if (acc) {
      o->mcaccep[type][MCMULTI] += 1.0;
      for (int t = 1; t < seg; t++)
         for (int id = 0; id < 3; id++) o->coords[id][P * gat(t) + sl(t)] = nx[t * 3 + id];
   }
}

void sched_rot_slice(orc_t *o, int type, int q)
{
   int RS = o->P + q;
   for (int a = 0; a < o->sys.type[type].numb; a++) {
      if (o->sys.type[type].molecule == 2) {
         double r1 = draw(o, RS), r2 = draw(o, RS), r3 = draw(o, RS), r4 = draw(o, RS);
         rot3d_step(o, q, a, type, r1, r2, r3, r4);
      } else {
         double r1 = draw(o, RS), r2 = draw(o, RS), r3 = draw(o, RS);
         rotlin_step(o, q, type, r1, r2, r3);
      }
   }
}
} // namespace

// ----------------------------------------------------------------------------
// C interface
// ----------------------------------------------------------------------------
extern "C" {

orc_t *orc_create(const orc_system_t *sys)
{
   orc_t *o = new orc();
   o->sys = *sys;
   o->P = sys->P; o->Q = sys->Q;
   o->N = 0;
   o->imtype = -1; o->bstype = -1;
   for (int t = 0; t < sys->ntypes; t++) {
      o->offset_atom[t] = o->N;
      o->N += sys->type[t].numb;
      for (int a = 0; a < sys->type[t].numb; a++) o->mctype.push_back(t);
      // mc_setup.cc:206-215
      double lam = 100.0 * (HBAR * HBAR) / (AMU * K_B);
      o->lambda[t] = 0.5 * lam / sys->type[t].mass;
      if (sys->type[t].stat == 1) o->bstype = t;
      if (sys->type[t].molecule) o->imtype = t;
   }
   o->offset_atom[sys->ntypes] = o->N;
   // mc_setup.cc:366-383
   o->beta = 1.0 / sys->temperature;
   o->tau = o->beta / (double)o->P;
   o->R = 1;
   o->rottau = 0.0;
   if (o->Q > 0) { o->rottau = o->beta / (double)o->Q; o->R = o->P / o->Q; }
   size_t n = (size_t)o->N * o->P;
   for (int d = 0; d < 3; d++) { o->coords[d].assign(n, 0.0); o->angles[d].assign(n, 0.0); o->cosine[d].assign(n, 0.0); o->newc[d].assign(n, 0.0); }
   o->pindex.resize(o->N); o->rindex.resize(o->N);
   for (int a = 0; a < o->N; a++) { o->pindex[a] = a; o->rindex[a] = a; }
   memset(o->mctotal, 0, sizeof(o->mctotal)); memset(o->mcaccep, 0, sizeof(o->mcaccep));
   // mc_estim.cc:254-261
   o->delta_radius = (MAX_RADIUS - MIN_RADIUS) / (double)MC_BINSR;
   o->delta_theta = M_PI / (double)(MC_BINST - 1);
   o->delta_chi = 2.0 * M_PI / (double)(MC_BINSC - 1);
   orc_reset_hist(o);
   return o;
}
void orc_destroy(orc_t *o) { delete o; }

void orc_set_pot1d(orc_t *o, int n, const double *grid, const double *v)   // mc_poten.cc:379-438
{
   o->n1d = n; o->g1d.assign(grid, grid + n); o->v1d.assign(v, v + n); o->y2_1d.assign(n, 0.0);
   init_spline(o->g1d.data(), o->v1d.data(), o->y2_1d.data(), n);
   double fr = o->v1d[0] / o->v1d[1];
   double dr = o->g1d[1] - o->g1d[0];
   o->alpha = log(fr) / dr;
   o->unode = o->v1d[0] * exp(o->alpha * o->g1d[0]);
   double r0 = pow(o->g1d[n - 2], 6.0), r1 = pow(o->g1d[n - 1], 6.0);
   fr = o->v1d[n - 1] - o->v1d[n - 2];
   o->c6 = fr / (1.0 / r0 - 1.0 / r1);
}
void orc_get_pot1d_setup(orc_t *o, double *y2, double *auc)
{
   memcpy(y2, o->y2_1d.data(), sizeof(double) * o->n1d);
   auc[0] = o->alpha; auc[1] = o->unode; auc[2] = o->c6;
}
void orc_set_pot2d(orc_t *o, int rs, int cs, double dr, double dc, const double *rg, const double *cg, const double *v)
{
   o->rs2d = rs; o->cs2d = cs; o->dr2d = dr; o->dc2d = dc;
   o->rg2d.assign(rg, rg + rs); o->cg2d.assign(cg, cg + cs); o->v2d.assign(v, v + (size_t)rs * cs);
}
void orc_set_pot3d(orc_t *o, int rg, int thg, int chg, double rvmin, double rvmax, const double *v)
{
   o->rg3 = rg; o->thg3 = thg; o->chg3 = chg; o->rvmin = rvmin; o->rvmax = rvmax;
   o->rvstep = (rvmax - rvmin) / (double)(rg - 1);       // mc_poten.cc:288
   o->v3d = v;
}
void orc_set_rotlin(orc_t *o, int n, const double *grid, const double *dens, const double *derv, const double *esqr)
{
   o->nrot = n;
   o->rgrid.assign(grid, grid + n); o->rdens.assign(dens, dens + n); o->rderv.assign(derv, derv + n); o->resqr.assign(esqr, esqr + n);
   o->rdens2.assign(n, 0.0); o->rderv2.assign(n, 0.0); o->resqr2.assign(n, 0.0);
   init_spline(o->rgrid.data(), o->rdens.data(), o->rdens2.data(), n);     // mc_poten.cc:543-545
   init_spline(o->rgrid.data(), o->rderv.data(), o->rderv2.data(), n);
   init_spline(o->rgrid.data(), o->resqr.data(), o->resqr2.data(), n);
}
void orc_set_rot3d(orc_t *o, const double *rho, const double *erot, const double *esq) { o->rho3 = rho; o->erot3 = erot; o->esq3 = esq; }
void orc_set_vspher(orc_t *, const double *t501) { oracle_set_vspher_table(t501); }

void orc_set_state(orc_t *o, const double *coords, const double *angles, const int *pindex)
{
   size_t n = (size_t)o->N * o->P;
   for (int d = 0; d < 3; d++)
      for (size_t i = 0; i < n; i++) { o->coords[d][i] = coords[d * n + i]; o->angles[d][i] = angles[d * n + i]; }
   for (size_t i = 0; i < n; i++) {                     // mc_main.cc:192-199
      double phi = o->angles[PHI][i], cost = o->angles[CTH][i];
      double sint = sqrt(1.0 - cost * cost);
      o->cosine[0][i] = sint * cos(phi); o->cosine[1][i] = sint * sin(phi); o->cosine[2][i] = cost;
   }
   if (pindex && o->bstype >= 0)
      for (int a = 0; a < o->sys.type[o->bstype].numb; a++) { o->pindex[a] = pindex[a]; o->rindex[pindex[a]] = a; }
}
void orc_get_state(orc_t *o, double *coords, double *angles, double *cosine)
{
   size_t n = (size_t)o->N * o->P;
   for (int d = 0; d < 3; d++)
      for (size_t i = 0; i < n; i++) {
         coords[d * n + i] = o->coords[d][i]; angles[d * n + i] = o->angles[d][i];
         if (cosine) cosine[d * n + i] = o->cosine[d][i];
      }
}

double orc_spot1d(orc_t *o, double r, int *klo) { return SPot1D(o, r, klo); }
double orc_lpot2d(orc_t *o, double r, double c, int *ir, int *ic) { return LPot2D(o, r, c, ir, ic); }
double orc_srotdens(orc_t *o, double g, int which) { return SRot(o, g, which); }
void orc_rotden(orc_t *o, const double *e1, const double *e2, double *rel, double *rho, double *erot, double *esq, int *index, int *istop)
{
   rotden_call(o, e1, e2, rel, rho, erot, esq, istop);
   if (index) *index = oracle_last_rotden_index;
}
double orc_vcord(orc_t *o, const double *eul, const double *rcom, const double *rpt, double *rtc, int *index)
{
   double v = vcord_call(o, eul, rcom, rpt, rtc);
   if (index) *index = oracle_last_vcord_index;
   return v;
}
// pure table look-ups on this handle's tables (identical doubles in -> indices and values out)
void orc_rotpro(orc_t *o, const double *deg3 /* phi, theta, chi */, double *rho, double *erot, double *esq, int *index, int *jstop)
{
   oracle_rotpro(deg3[0], deg3[1], deg3[2], o->rho3, o->erot3, o->esq3, rho, erot, esq, index, jstop);
}
double orc_vcalc(orc_t *o, const double *rtc /* r bohr, theta deg, chi deg */, int *index)
{
   return oracle_vcalc(rtc[0], rtc[1], rtc[2], o->rvmin, o->rvmax, o->rvstep, o->rg3, o->thg3, o->chg3, o->v3d, index);
}
void orc_deleul(const double *e1, const double *e2, double *rel) { oracle_deleul(e1, e2, rel); }
double orc_vspher(double r, double *rclamp)
{
   double rr = r, v;
   vspher_(&rr, &v);
   if (rclamp) *rclamp = rr;
   return v;
}
double orc_caleng(const double *c1, const double *c2, const double *e1, const double *e2)
{
   double a[3] = {c1[0], c1[1], c1[2]}, b[3] = {c2[0], c2[1], c2[2]}, x[3] = {e1[0], e1[1], e1[2]}, y[3] = {e2[0], e2[1], e2[2]}, E;
   caleng_(a, b, &E, x, y);
   return E;
}

double orc_pot_energy_it(orc_t *o, int atom, const double *pos3, int it)
{
   double p[3];
   for (int id = 0; id < 3; id++) p[id] = pos3 ? pos3[id] : o->coords[id][o->P * atom + it];
   return PotEnergy_it(o, atom, p, it);
}
double orc_pot_energy_path(orc_t *o, int atom, const double *shift3) { return PotEnergy_path(o, atom, shift3); }
double orc_pot_rot_energy(orc_t *o, int atom, const double *c3, int it) { return PotRotEnergy(o, atom, c3, it); }
double orc_pot_rot_e3d(orc_t *o, int atom, const double *e3, int it) { return PotRotE3D(o, atom, e3, it); }

int orc_bisection_move(orc_t *o, int type, int atom, int time, const double *ug, const double *ua, int exch, int *consumed) { return bisection_move(o, type, atom, time, ug, ua, exch, consumed); }
int orc_molecular_move(orc_t *o, int type, int atom, const double *u3, double ua) { return molecular_move(o, type, atom, u3, ua); }
int orc_rot3d_step(orc_t *o, int it1, int atom0, int type, double r1, double r2, double r3, double r4) { return rot3d_step(o, it1, atom0, type, r1, r2, r3, r4); }
int orc_rotlin_step(orc_t *o, int it1, int type, double r1, double r2, double r3) { return rotlin_step(o, it1, type, r1, r2, r3); }

double orc_get_kin(orc_t *o) { return GetKin(o); }
double orc_get_pot(orc_t *o, int dens) { return GetPot(o, dens); }
double orc_get_rot_energy(orc_t *o, double *esq, double *eterm) { double s = GetRotEnergy(o); *esq = o->ErotSQ; *eterm = o->Erot_termSQ; return s; }
double orc_get_rot_e3d(orc_t *o, double *esq, double *eterm) { double s = GetRotE3D(o); *esq = o->ErotSQ; *eterm = o->Erot_termSQ; return s; }
void orc_get_rcf(orc_t *o, double *rcf0) { for (int i = 0; i < o->Q; i++) rcf0[i] = 0.0; GetRCF(o, rcf0); }
void orc_reset_hist(orc_t *o)
{
   o->gr1d.assign(MC_BINSR, 0.0); o->gr2d.assign(MC_BINSR * MC_BINST, 0.0);
   for (int d = 0; d < 2; d++) o->gr3d[d].assign((size_t)MC_BINSR * MC_BINST * MC_BINSC, 0.0);
   o->relthe.assign(MC_BINST, 0.0); o->relphi.assign(MC_BINSC, 0.0); o->relchi.assign(MC_BINSC, 0.0);
}
void orc_get_hist(orc_t *o, double *g1, double *g2, double *g3a, double *g3m, double *rt, double *rp, double *rc)
{
   if (g1) memcpy(g1, o->gr1d.data(), sizeof(double) * o->gr1d.size());
   if (g2) memcpy(g2, o->gr2d.data(), sizeof(double) * o->gr2d.size());
   if (g3a) memcpy(g3a, o->gr3d[0].data(), sizeof(double) * o->gr3d[0].size());
   if (g3m) memcpy(g3m, o->gr3d[1].data(), sizeof(double) * o->gr3d[1].size());
   if (rt) memcpy(rt, o->relthe.data(), sizeof(double) * MC_BINST);
   if (rp) memcpy(rp, o->relphi.data(), sizeof(double) * MC_BINSC);
   if (rc) memcpy(rc, o->relchi.data(), sizeof(double) * MC_BINSC);
}

void orc_exchange_length(orc_t *o, double *ploops) { GetExchangeLength(o, ploops); }
void orc_area_estimators(orc_t *o, double *out4) { GetAreaEstimators(o, out4); }
void orc_area_estim3d(orc_t *o, int iframe, double *area_proj3, double *inert9) { GetAreaEstim3D(o, iframe, area_proj3, inert9); }
void orc_reflect(orc_t *o, int plane) { Reflect_MF(o, plane); }
void orc_rotsym(orc_t *o, double u, int nfold) { RotSymConfig(o, u, nfold); }
// the tail of MCGetAverage (mc_main.cc:647-692) with the uniforms taken from the chain's miscellaneous stream
// in the device's order: REFLECTY, REFLECTX, REFLECTZ, ROTSYM (+ the rotor pick)
void orc_sched_symmetry(orc_t *o, int refl_x, int refl_y, int refl_z, int rotsym, int nfold)
{
   int MS = o->P + o->Q;
   if (refl_y && draw(o, MS) < 0.5) Reflect_MF(o, 0);
   if (refl_x && draw(o, MS) < 0.5) Reflect_MF(o, 1);
   if (refl_z && draw(o, MS) < 0.5) Reflect_MF(o, 2);
   if (rotsym && draw(o, MS) < 0.5) RotSymConfig(o, draw(o, MS), nfold);
}

// ---- worm (N1) ----
void orc_worm_init(orc_t *o, int type, double c_input, int m)      // MCWormInit, mc_qworm.cc:48-82
{
   auto &W = o->worm;
   W.on = 1; W.type = type; W.exists = 0; W.m = m;
   int numb = o->sys.type[type].numb;
   double density = (double)o->N / (o->sys.box[0] * o->sys.box[1] * o->sys.box[2]);
   W.c = c_input * (density / (numb * o->P * m));
   W.qw_norm = W.c * numb * o->P * m;
   W.twave2 = 4.0 * o->lambda[type] * o->tau;                      // mc_setup.cc:394
   W.cutoff2 = 100.0 * 100.0 * ((double)m * 4.0 * o->lambda[type] * o->tau);
   o->dr2_list.assign(numb + 2, 0.0); o->atm_list.assign(numb + 2, 0); o->ptable.assign(numb + 2, 0.0);
   o->countqw = 1.0;
   for (int i = 0; i < 7; i++) { o->qwtotal[i] = 0; o->qwaccep[i] = 0; }
}
void orc_worm_set(orc_t *o, const int *st5) { auto &W = o->worm; W.exists = st5[0]; W.ira = st5[1]; W.masha = st5[2]; W.atom_i = st5[3]; W.atom_m = st5[4]; }
void orc_worm_get(orc_t *o, int *st5) { auto &W = o->worm; st5[0] = W.exists; st5[1] = W.ira; st5[2] = W.masha; st5[3] = W.atom_i; st5[4] = W.atom_m; }
void orc_worm_push(orc_t *o, int stream, const double *u, int n) { for (int i = 0; i < n; i++) o->wq[stream].push_back(u[i]); }
void orc_worm_clear(orc_t *o) { for (auto &q : o->wq) q.clear(); }
int  orc_worm_pending(orc_t *o, int stream) { return (int)o->wq[stream].size(); }
/* which: 0 open, 1 close, 4 advance, 5 recede, 6 swap (the QW_* codes), 7 the whole MCWormMove */
void orc_worm_op(orc_t *o, int which, int sched_stream)
{
   o->wmode = sched_stream ? 1 : 0;
   switch (which) {
      case 0: qworm_open(o); break;
      case 1: qworm_close(o); break;
      case 4: qworm_advance(o); break;
      case 5: qworm_recede(o); break;
      case 6: qworm_swap(o); break;
      default: worm_move(o);
   }
}
void orc_worm_counters(orc_t *o, double *total7, double *accep7, double *countqw)
{
   for (int i = 0; i < 7; i++) { total7[i] = o->qwtotal[i]; accep7[i] = o->qwaccep[i]; }
   *countqw = o->countqw;
}
void orc_get_perm(orc_t *o, int *pindex, int *rindex, int n) { for (int i = 0; i < n; i++) { pindex[i] = o->pindex[i]; rindex[i] = o->rindex[i]; } }
int  orc_world_line(orc_t *o, int atom, int pt) { return WorldLine(o, atom, pt) ? 1 : 0; }

void orc_mrg_stream_state(const unsigned long *seed6, long stream, double *st) { mrg_stream_state(seed6, stream, st); }
void orc_mrg_draws(const unsigned long *seed6, long first, int nstream, int ndraw, double *out)
{
   for (int s = 0; s < nstream; s++) {
      double st[6];
      mrg_stream_state(seed6, first + s, st);
      for (int k = 0; k < ndraw; k++) out[(size_t)s * ndraw + k] = mrg_u01(st);
   }
}

void orc_sched_seed(orc_t *o, const unsigned long *seed6, long chain_global)
{
   int S = o->P + o->Q + 8;
   o->streams.resize(S);
   double st[6];
   mrg_stream_state(seed6, chain_global * S, st);
   double s1[3] = {st[0], st[1], st[2]}, s2[3] = {st[3], st[4], st[5]};
   for (int s = 0; s < S; s++) {
      for (int k = 0; k < 3; k++) { o->streams[s][k] = s1[k]; o->streams[s][3 + k] = s2[k]; }
      MatVecModM(A1p127, s1, s1, m1);                  // rngstream.cc:319-320
      MatVecModM(A2p127, s2, s2, m2);
   }
}

void orc_sched_run(orc_t *o, long t0, long nsteps)
{
   int P = o->P, Q = o->Q;
   for (long t = t0; t < t0 + nsteps; t++) {
      int time = (int)(t % P);
      for (int type = 0; type < o->sys.ntypes; type++) {
         const orc_type_t &T = o->sys.type[type];
         bool closed = true;
         if (o->worm.on && type == o->worm.type) {        // mc_main.cc:355-379: worm moves, then the path moves in the Z sector only
            o->wmode = 1;
            worm_move(o);
            closed = !o->worm.exists;
         }
         if (time == 0 && closed) sched_molecular(o, type);
         int seg = 1 << T.levels, nseg = P / seg;
         if (time % nseg == 0 && closed) {
            int off = (time / nseg) % P;
            for (int atom = 0; atom < T.numb; atom++)
               for (int k = 0; k < nseg; k++) sched_bisect(o, type, atom, (off + k * seg) % P);
         }
         if (type == o->imtype && Q > 0) {
            int qe = (Q % 2 == 1 && Q > 1) ? Q - 1 : Q;      // odd Q: the last slice gets its own phase
            for (int q = 0; q < qe; q += 2) sched_rot_slice(o, type, q);
            for (int q = 1; q < qe; q += 2) sched_rot_slice(o, type, q);
            if (qe != Q) sched_rot_slice(o, type, Q - 1);
         }
      }
   }
}
void orc_sched_counters(orc_t *o, double *tot, double *acc)
{
   for (int t = 0; t < 2; t++) for (int m = 0; m < 3; m++) { tot[t * 3 + m] = o->mctotal[t][m]; acc[t * 3 + m] = o->mcaccep[t][m]; }
}

} // extern "C"
