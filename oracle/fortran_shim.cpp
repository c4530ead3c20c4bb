// ORACLE (test infrastructure, not product code).
//
// C++ restatement of the ten Fortran-77 entry points of MoRiBS-PIMC's hot path
// that cannot be compiled in this image (no gfortran).  Each routine follows
// the gfortran ABI (lower-case name + '_', all arguments by pointer, 3x3
// arrays column-major, no hidden string lengths) so that the reference's own
// C++ translation units link against it unchanged (oracle/_ref), and so that
// oracle/pimc_oracle.cpp (the travelling CPU port) calls the same leaf math.
//
// Evaluation order follows the Fortran source left to right; build with
// -O2 -ffp-contract=off.  PARITY STATUS of these ten routines: "parity
// unpinned" -- the reference ships no numeric golden vectors for them and the
// Fortran itself cannot be run here; they are a careful line-by-line
// restatement cited below.
//
//   rotden_/deleul/matpre/rottrn/within   rotden.f:1-216
//   rsrot_ (live branch only)             rotden.f:218-286
//   rsline_                               rotden.f:356-375
//   rotpro                                rotpro_sub.f:1-64
//   vcord_ + dotprd/dnorm/dotang/crsprd   vcord.f:1-98,143-188
//   rflmfy_/rflmfx_/rflmfz_               vcord.f:257-546
//   vcalc                                 vcalc.f:1-65
//   caleng_                               caleng_tip4p_gg.f:2-186
//   vspher_                               vspher.f:12-544 (table supplied at run time)
//   initconf_                             initconf.f:1-27
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

const double PI = 3.14159265358979323846;

// column-major 3x3 accessor, 1-based like the Fortran
inline double &M(double *m, int i, int j) { return m[(j - 1) * 3 + (i - 1)]; }
inline double Mc(const double *m, int i, int j) { return m[(j - 1) * 3 + (i - 1)]; }

// rotden.f:136-163
void matpre(const double *eul, double *rotmat)
{
   double phi = eul[0], theta = eul[1], chi = eul[2];
   double cp = cos(phi), sp = sin(phi);
   double ct = cos(theta), st = sin(theta);
   double ck = cos(chi), sk = sin(chi);
   M(rotmat, 1, 1) = cp * ct * ck - sp * sk;
   M(rotmat, 1, 2) = -cp * ct * sk - sp * ck;
   M(rotmat, 1, 3) = cp * st;
   M(rotmat, 2, 1) = sp * ct * ck + cp * sk;
   M(rotmat, 2, 2) = -sp * ct * sk + cp * ck;
   M(rotmat, 2, 3) = sp * st;
   M(rotmat, 3, 1) = -st * ck;
   M(rotmat, 3, 2) = st * sk;
   M(rotmat, 3, 3) = ct;
}

// rotden.f:165-177   rsf = rcom + rotmat * rwf
void rottrn(const double *rotmat, const double *rwf, double *rsf, const double *rcom)
{
   for (int i = 1; i <= 3; i++) {
      double s = rcom[i - 1];
      for (int j = 1; j <= 3; j++) s = s + Mc(rotmat, i, j) * rwf[j - 1];
      rsf[i - 1] = s;
   }
}

// rotden.f:208-216
inline void within(double &v)
{
   if (v > 1.0) v = 1.0;
   if (v < -1.0) v = -1.0;
}

// Euler-angle extraction shared by deleul (rotden.f:62-119) and the three
// rflmf* routines (vcord.f:293-343 and copies).
void euler_from_matrix(const double *rotma2, double *out)
{
   const double small = 1.0e-08, zero = 0.0;
   double cost = Mc(rotma2, 3, 3);
   within(cost);
   double theta2 = acos(cost);
   double sint = sin(theta2);
   double phi2, chi2;
   if (fabs(1.0 - cost) < small) {
      phi2 = 0.0;
      double cchi = Mc(rotma2, 1, 1), schi = Mc(rotma2, 2, 1);
      within(cchi); within(schi);
      if (schi > zero) chi2 = acos(cchi); else chi2 = 2.0 * PI - acos(cchi);
   } else if (fabs(1.0 + cost) < small) {
      phi2 = 0.0;
      double cchi = Mc(rotma2, 2, 2), schi = Mc(rotma2, 1, 2);
      within(cchi); within(schi);
      if (schi > zero) chi2 = acos(cchi); else chi2 = 2.0 * PI - acos(cchi);
   } else {
      double cphi = Mc(rotma2, 1, 3) / sint;
      double sphi = Mc(rotma2, 2, 3) / sint;
      double cchi = -Mc(rotma2, 3, 1) / sint;
      double schi = Mc(rotma2, 3, 2) / sint;
      within(cphi); within(sphi); within(cchi); within(schi);
      if (sphi > zero) phi2 = acos(cphi); else phi2 = 2.0 * PI - acos(cphi);
      if (schi > zero) chi2 = acos(cchi); else chi2 = 2.0 * PI - acos(cchi);
   }
   out[0] = phi2; out[1] = theta2; out[2] = chi2;
}

// rotden.f:32-134
void deleul(const double *Eulan1, const double *Eulan2, double *Eulrel, int *istop)
{
   double rotmat[9], rotma1[9], rotma2[9];
   *istop = 0;
   matpre(Eulan2, rotmat);
   matpre(Eulan1, rotma1);
   for (int i = 1; i <= 3; i++)
      for (int j = 1; j <= 3; j++) {
         double s = 0.0;
         for (int k = 1; k <= 3; k++) s = s + Mc(rotma1, k, i) * Mc(rotmat, k, j);
         M(rotma2, i, j) = s;
      }
   euler_from_matrix(rotma2, Eulrel);
}

// rotpro_sub.f:1-64.  The Fortran leaves delch2/delch3 (etc.) uninitialised
// when an index sits on the last grid line; here they are zero (documented fence).
void rotpro(double chi, double phi, double theta, double *rho, double *erot, double *esq,
            const double *rhoprp, const double *erotpr, const double *erotsq, int *jstop,
            int *idx_out)
{
   int ichi = (int)chi, iphi = (int)phi, itheta = (int)theta;
   if (ichi > 360 || ichi < 0) { printf("ichi out or range %d %g\n", ichi, chi); ichi = 0; *jstop = 1; }
   if (iphi > 360 || iphi < 0) { printf("iphi out or range %d %g\n", iphi, phi); iphi = 0; *jstop = 1; }
   if (itheta > 180 || itheta < 0) { printf("itheta out or range %d %g\n", itheta, theta); itheta = 0; *jstop = 1; }

   int ind = (itheta * 361 + iphi) * 361 + ichi;
   if (idx_out) *idx_out = ind;
   double rho0 = rhoprp[ind], erot0 = erotpr[ind], esq0 = erotsq[ind];
   double delchi = 0, delphi = 0, delthe = 0;
   double delch2 = 0, delph2 = 0, delth2 = 0, delch3 = 0, delph3 = 0, delth3 = 0;
   if (ichi != 360) {
      int k = (itheta * 361 + iphi) * 361 + ichi + 1;
      delchi = rhoprp[k] - rho0; delch2 = erotpr[k] - erot0; delch3 = erotsq[k] - esq0;
   }
   if (iphi != 360) {
      int k = (itheta * 361 + iphi + 1) * 361 + ichi;
      delphi = rhoprp[k] - rho0; delph2 = erotpr[k] - erot0; delph3 = erotsq[k] - esq0;
   }
   if (itheta != 180) {
      int k = ((itheta + 1) * 361 + iphi) * 361 + ichi;
      delthe = rhoprp[k] - rho0; delth2 = erotpr[k] - erot0; delth3 = erotsq[k] - esq0;
   }
   double fc = chi - (double)ichi, fp = phi - (double)iphi, ft = theta - (double)itheta;
   *rho = rho0 + delchi * fc + delphi * fp + delthe * ft;
   *erot = erot0 + delch2 * fc + delph2 * fp + delth2 * ft;
   *esq = esq0 + delch3 * fc + delph3 * fp + delth3 * ft;
}

inline double dotprd(const double *a, const double *b)
{
   double d = 0.0;
   for (int i = 0; i < 3; i++) d = d + a[i] * b[i];
   return d;
}
inline double dnorm(const double *a) { return sqrt(dotprd(a, a)); }
inline double dotang(const double *a, const double *b)
{
   double d = dotprd(a, b) / (dnorm(a) * dnorm(b));
   if (d > 1.0) d = 1.0;
   if (d < -1.0) d = -1.0;
   return acos(d);
}
inline void crsprd(const double *a, const double *b, double *c)
{
   c[0] = a[1] * b[2] - a[2] * b[1];
   c[1] = a[2] * b[0] - a[0] * b[2];
   c[2] = a[0] * b[1] - a[1] * b[0];
}

// vcalc.f:1-65 (r in bohr, theta and chi in degrees)
double vcalc(double r, double theta, double chi, double r0, double rmax, double rstep,
             int nrgrd, int nthgrd, int nchgrd, const double *vtable, int *idx_out)
{
   int maxrpt = nrgrd - 1, mxthpt = nthgrd - 1, mxchpt = nchgrd - 1;
   if (r < r0) r = r0;
   if (r > rmax) r = rmax;
   int ir = (int)((r - r0) / rstep);
   int ith = (int)theta;
   int ich = (int)chi;
   if (ich > mxchpt) ich = mxchpt;
   if (ich < 0) ich = 0;
   if (ith > mxthpt) ith = mxthpt;
   if (ith < 0) ith = 0;
   int ind = (ir * nthgrd + ith) * nchgrd + ich;
   if (idx_out) *idx_out = ind;
   double v0 = vtable[ind];
   double gradr, delr, gradth, delth, gradch, delch;
   if (ir == maxrpt) { gradr = 0.0; delr = 0.0; }
   else {
      gradr = (vtable[((ir + 1) * nthgrd + ith) * nchgrd + ich] - v0) / rstep;
      delr = r - (r0 + ir * rstep);
   }
   if (ith == mxthpt) { gradth = 0.0; delth = 0.0; }
   else {
      gradth = vtable[(ir * nthgrd + (ith + 1)) * nchgrd + ich] - v0;
      delth = theta - (double)ith;
   }
   if (ich == mxchpt) { gradch = 0.0; delch = 0.0; }
   else {
      gradch = vtable[(ir * nthgrd + ith) * nchgrd + ich + 1] - v0;
      delch = chi - (double)ich;
   }
   return v0 + gradr * delr + gradth * delth + gradch * delch;
}

#include "../moribs-pimc_b200/data/vspher_table.h"   // DATA vtable of vspher.f:15-517, extracted by oracle/extract_vspher.py (input data)
const double *g_vspher_table = PIMC_VSPHER_TABLE;   // 501 entries; oracle_set_vspher_table may point it elsewhere

void reflect_finish(double *hatx, double *haty, double *hatz, double *eulang)
{
   double rotma2[9];
   for (int i = 1; i <= 3; i++) {
      M(rotma2, i, 1) = hatx[i - 1];
      M(rotma2, i, 2) = haty[i - 1];
      M(rotma2, i, 3) = hatz[i - 1];
   }
   euler_from_matrix(rotma2, eulang);
}

} // namespace

extern "C" {

// last flattened table indices touched by rotden_/vcord_ (test hook: bit-exact index parity)
int oracle_last_rotden_index = -1;
int oracle_last_vcord_index = -1;

void oracle_set_vspher_table(const double *t501) { g_vspher_table = t501 ? t501 : PIMC_VSPHER_TABLE; }

// pure-function views of the table look-ups (test hooks: the same doubles go to the device's selectors)
// rotpro_sub.f:1-64: angles in degrees; erot/esq in the tables' units (cm^-1)
void oracle_rotpro(double phi, double theta, double chi, const double *rhoprp, const double *erotpr, const double *erotsq,
                   double *rho, double *erot, double *esq, int *index, int *jstop)
{
   *jstop = 0;
   rotpro(chi, phi, theta, rho, erot, esq, rhoprp, erotpr, erotsq, jstop, index);
}
// vcalc.f:1-65: r in bohr, theta and chi in degrees
double oracle_vcalc(double r, double theta, double chi, double r0, double rmax, double rstep, int nrgrd, int nthgrd, int nchgrd,
                    const double *vtable, int *index)
{
   return vcalc(r, theta, chi, r0, rmax, rstep, nrgrd, nthgrd, nchgrd, vtable, index);
}
// deleul, rotden.f:32-134: relative Euler angles (radians) of two orientations
void oracle_deleul(const double *e1, const double *e2, double *rel)
{
   int istop = 0;
   deleul(e1, e2, rel, &istop);
}

// rotden.f:1-31
void rotden_(double *Eulan1, double *Eulan2, double *Eulrel, double *rho, double *erot, double *esq,
             double *rhoprp, double *erotpr, double *erotsq, int *istop)
{
   const double wno2k = 0.6950356;
   deleul(Eulan1, Eulan2, Eulrel, istop);
   double phi = Eulrel[0] * 180.0 / PI;
   double theta = Eulrel[1] * 180.0 / PI;
   double chi = Eulrel[2] * 180.0 / PI;
   int jstop = 0;
   rotpro(chi, phi, theta, rho, erot, esq, rhoprp, erotpr, erotsq, &jstop, &oracle_last_rotden_index);
   if (jstop == 1) *istop = 1;
   *erot = *erot / wno2k;
   *esq = *esq / (wno2k * wno2k);
}

// vcord.f:1-98
void vcord_(double *Eulang, double *RCOM, double *RpH2, double *vtable, int *nrgrd, int *nthgrd,
            int *nchgrd, double *rvmax, double *rvmin, double *rvstep, double *vpot, double *radret,
            double *theret, double *chiret, double *hatx, double *haty, double *hatz, int *ivcord)
{
   const double small = 1.0e-08, bo2ang = 0.529177249;
   const double unx[3] = {1, 0, 0}, uny[3] = {0, 1, 0}, unz[3] = {0, 0, 1}, origin[3] = {0, 0, 0};
   double rotmat[9], RH2COM[3];
   matpre(Eulang, rotmat);
   rottrn(rotmat, unx, hatx, origin);
   rottrn(rotmat, uny, haty, origin);
   rottrn(rotmat, unz, hatz, origin);
   for (int i = 0; i < 3; i++) RH2COM[i] = RpH2[i] - RCOM[i];
   if (*ivcord == 1) return;

   double thewff = dotang(hatz, RH2COM);
   double chiwff;
   if (fabs(dotprd(RH2COM, hatx)) < small) {
      chiwff = PI / 2.0;
   } else {
      double tanchi = dotprd(RH2COM, haty) / dotprd(RH2COM, hatx);
      chiwff = atan(fabs(tanchi));
   }
   double radwff = dnorm(RH2COM);
   *radret = radwff;
   *theret = thewff;
   double Rdotx = dotprd(RH2COM, hatx);
   double Rdoty = dotprd(RH2COM, haty);
   if (Rdotx >= 0.0 && Rdoty >= 0.0) *chiret = chiwff;
   else if (Rdotx < 0.0 && Rdoty >= 0.0) *chiret = PI - chiwff;
   else if (Rdotx < 0.0 && Rdoty < 0.0) *chiret = PI + chiwff;
   else *chiret = 2 * PI - chiwff;
   if (*nchgrd == 91) {
   } else if (*nchgrd == 181) {
      chiwff = *chiret;
      if (*chiret > PI) chiwff = 2 * PI - *chiret;
   } else if (*nchgrd == 361) {
      chiwff = *chiret;
   }
   radwff = radwff / bo2ang;
   thewff = thewff * 180.0 / PI;
   chiwff = chiwff * 180.0 / PI;
   *vpot = vcalc(radwff, thewff, chiwff, *rvmin, *rvmax, *rvstep, *nrgrd, *nthgrd, *nchgrd, vtable,
                 &oracle_last_vcord_index);
}

// caleng_tip4p_gg.f:2-186
void caleng_(double *com_1, double *com_2, double *E_2H2O, double *Eulang_1, double *Eulang_2)
{
   const double qm = -1.04, qh = 0.520, br2ang = 0.52917721092, hr2k = 3.1577465e5,
                kcal2k = 503.218978939;
   const double ROwf[3] = {0.0, 0.0, 0.06562}, RH1wf[3] = {0.7557, 0.0, -0.5223},
                RH2wf[3] = {-0.7557, 0.0, -0.5223}, RMwf[3] = {0.0, 0.0, -0.08438};
   double rot1[9], rot2[9];
   double RO1[3], RM1[3], RH11[3], RH21[3], RO2[3], RM2[3], RH12[3], RH22[3];
   matpre(Eulang_1, rot1);
   rottrn(rot1, ROwf, RO1, com_1);
   rottrn(rot1, RMwf, RM1, com_1);
   rottrn(rot1, RH1wf, RH11, com_1);
   rottrn(rot1, RH2wf, RH21, com_1);
   matpre(Eulang_2, rot2);
   rottrn(rot2, ROwf, RO2, com_2);
   rottrn(rot2, RMwf, RM2, com_2);
   rottrn(rot2, RH1wf, RH12, com_2);
   rottrn(rot2, RH2wf, RH22, com_2);

   double roo = 0, rmm = 0;
   for (int i = 0; i < 3; i++) {
      roo = roo + (RO1[i] - RO2[i]) * (RO1[i] - RO2[i]);
      rmm = rmm + (RM1[i] - RM2[i]) * (RM1[i] - RM2[i]);
   }
   rmm = sqrt(rmm);
   double roo4 = roo * roo, roo6 = roo4 * roo, roo12 = roo6 * roo6;
   const double A_param = 6.0e5, B_param = 610.0;
   double v_o2lj = A_param / roo12 - B_param / roo6;

   double rhm1 = 0, rhm2 = 0, rhm3 = 0, rhm4 = 0, rhh1 = 0, rhh2 = 0, rhh3 = 0, rhh4 = 0;
   for (int i = 0; i < 3; i++) {
      rhm1 = rhm1 + (RM1[i] - RH12[i]) * (RM1[i] - RH12[i]);
      rhm2 = rhm2 + (RM1[i] - RH22[i]) * (RM1[i] - RH22[i]);
      rhm3 = rhm3 + (RM2[i] - RH11[i]) * (RM2[i] - RH11[i]);
      rhm4 = rhm4 + (RM2[i] - RH21[i]) * (RM2[i] - RH21[i]);
      rhh1 = rhh1 + (RH11[i] - RH12[i]) * (RH11[i] - RH12[i]);
      rhh2 = rhh2 + (RH11[i] - RH22[i]) * (RH11[i] - RH22[i]);
      rhh3 = rhh3 + (RH21[i] - RH12[i]) * (RH21[i] - RH12[i]);
      rhh4 = rhh4 + (RH21[i] - RH22[i]) * (RH21[i] - RH22[i]);
   }
   rhm1 = sqrt(rhm1); rhm2 = sqrt(rhm2); rhm3 = sqrt(rhm3); rhm4 = sqrt(rhm4);
   rhh1 = sqrt(rhh1); rhh2 = sqrt(rhh2); rhh3 = sqrt(rhh3); rhh4 = sqrt(rhh4);
   // the O-H and O-O Coulomb terms carry qo = 0 and are dropped by the Fortran (:178-181)
   double v_mhcolm = qm * qh * (1.0 / rhm1 + 1.0 / rhm2 + 1.0 / rhm3 + 1.0 / rhm4);
   double v_hhcolm = qh * qh * (1.0 / rhh1 + 1.0 / rhh2 + 1.0 / rhh3 + 1.0 / rhh4);
   double v_mmcolm = qm * qm * (1.0 / rmm);
   *E_2H2O = v_o2lj * kcal2k + (v_mhcolm + v_mmcolm + v_hhcolm) * hr2k * br2ang;
}

// vspher.f:12-544.  The 501-entry radial table is DATA in the Fortran source; it
// is handed in through oracle_set_vspher_table (oracle/_ref extracts it at build time).
void vspher_(double *r, double *vpot)
{
   const double r0 = 3.0, rmax = 26.0, rstep = 0.046, ang2bo = (double)0.5291772f;   // REAL*4 literal widened (vspher.f:519 has no D exponent)
   const int maxrpt = 500;
   if (!g_vspher_table) { printf("vspher_: table not set\n"); exit(1); }
   *r = *r / ang2bo;
   if (*r < r0) *r = r0;
   if (*r > rmax) *r = rmax;
   int ir = (int)((*r - r0) / rstep);
   double v0 = g_vspher_table[ir];
   double gradr, delr;
   if (ir == maxrpt) { gradr = 0.0; delr = 0.0; }
   else { gradr = (g_vspher_table[ir + 1] - v0) / rstep; delr = *r - (r0 + ir * rstep); }
   *vpot = v0 + gradr * delr;
}

// rotden.f:218-286 (everything after the early return at :286 is dead code)
void rsrot_(double *Eulan1, double *Eulan2, double *xrot, double *yrot, double *zrot, double *tauC,
            int *iodevn, double *eoff, double *rho, double *erot)
{
   const double wno2k = 0.6950356;
   (void)eoff;
   double rotma1[9], rotma2[9], digrel[3], blist[3];
   double tau = *tauC / wno2k;
   blist[0] = 1.0 / *xrot; blist[1] = 1.0 / *yrot; blist[2] = 1.0 / *zrot;
   matpre(Eulan1, rotma1);
   matpre(Eulan2, rotma2);
   for (int i = 1; i <= 3; i++) {
      double s = 0.0;
      for (int j = 1; j <= 3; j++) s = s + Mc(rotma1, j, i) * Mc(rotma2, j, i);
      digrel[i - 1] = s;
   }
   double sumaxs = (blist[0] - blist[1] - blist[2]) * (1.0 - digrel[0]) +
                   (blist[1] - blist[2] - blist[0]) * (1.0 - digrel[1]) +
                   (blist[2] - blist[0] - blist[1]) * (1.0 - digrel[2]);
   if (*iodevn == -1) {
      *rho = sumaxs;
      double e = sumaxs / (4.0 * tau * tau);
      e = e + 1.5 / tau + 0.25 * (*xrot + *yrot + *zrot);
      *erot = e / wno2k;
   }
}

// rotden.f:356-375
void rsline_(double *Brot, double *dprd, double *tauC, double *rho, double *erot)
{
   const double wno2k = 0.6950356;
   double tau = *tauC / wno2k;
   double r = (1.0 - *dprd) / (2.0 * *Brot * tau);
   double e = (1.0 - r) / tau;
   *rho = -r;
   *erot = e / wno2k;
}

// vcord.f:257-352: negate y of x-hat and z-hat, y-hat = z x x
void rflmfy_(double *rcom, double *hatx, double *haty, double *hatz, double *eulang)
{
   (void)rcom;
   hatx[1] = -hatx[1];
   hatz[1] = -hatz[1];
   crsprd(hatz, hatx, haty);
   reflect_finish(hatx, haty, hatz, eulang);
}
// vcord.f:354-449: negate y of y-hat and z-hat, x-hat = y x z
void rflmfx_(double *rcom, double *hatx, double *haty, double *hatz, double *eulang)
{
   (void)rcom;
   haty[1] = -haty[1];
   hatz[1] = -hatz[1];
   crsprd(haty, hatz, hatx);
   reflect_finish(hatx, haty, hatz, eulang);
}
// vcord.f:451-546: negate y of x-hat and y-hat, z-hat = x x y
void rflmfz_(double *rcom, double *hatx, double *haty, double *hatz, double *eulang)
{
   (void)rcom;
   hatx[1] = -hatx[1];
   haty[1] = -haty[1];
   crsprd(hatx, haty, hatz);
   reflect_finish(hatx, haty, hatz, eulang);
}

// initconf.f:1-27: list-directed read of xyz.init in the CWD
void initconf_(double *coords, double *angles, int *ntotal, int *nboson, int *indexp, int *indexr)
{
   FILE *f = fopen("xyz.init", "r");
   if (!f) { printf("initconf_: cannot open xyz.init\n"); exit(1); }
   int idump;
   if (fscanf(f, "%d", &idump) != 1) { printf("initconf_: bad header\n"); exit(1); }
   for (int i = 0; i < *nboson; i++)
      if (fscanf(f, "%d", &indexp[i]) != 1) { printf("initconf_: bad permutation\n"); exit(1); }
   for (int i = 0; i < *nboson; i++) indexr[indexp[i]] = i;
   int c;
   while ((c = fgetc(f)) != '\n' && c != EOF) {}          // rest of line 1
   while ((c = fgetc(f)) != '\n' && c != EOF) {}          // comment line 2
   char label[64];
   for (int i = 0; i < *ntotal; i++) {
      if (fscanf(f, "%63s", label) != 1) { printf("initconf_: short file\n"); exit(1); }
      for (int j = 0; j < 3; j++)
         if (fscanf(f, "%lf %lf", &coords[i * 3 + j], &angles[i * 3 + j]) != 2) {
            printf("initconf_: bad bead %d\n", i); exit(1);
         }
   }
   fclose(f);
}

} // extern "C"
