// ORACLE (test infrastructure, not product code): CPU restatement of the reference's rotational
// density-matrix TABLE GENERATORS, in the evaluation order of the Fortran:
//
//   nmv_prop/asymrho.f      asymmetric top: rigid-rotor blocks (rotmat :792-813), TRED2/TQL (:839-1004),
//                           Wigner d in real*16 from Zare Eq. 3.57 (wigd :1006-1039, calfac :1041-1051),
//                           grid-point sums (:488-659), symmetry fill (:665-709), partition sums (:299-425)
//   symtop_prop/symrho.f    symmetric top: :57-168
//   linear_prop/linden.f    linear rotor: exarho :88-122, lgnd :180-193, driver loop :22-69
//
// real*16 is IEEE binary128 (__float128 + libquadmath here), everything else double, -ffp-contract=off.
// PINNED on the reference's own golden outputs (tests/test_tablegen_oracle.py): nmv_prop/rho.den010_{rho,eng,esq}
// and the "AT BETA"/"AT TAU" lines of nmv_prop/log (argument list of nmv_prop/a-run), symtop_prop/rho.den0{00,10}_*
// (symtop_prop/a-run), examples/*/N2O_T0.5t128.rot and CO2_T100t4.rot (linden.f output).
// Only tests/ (and oracle-side scripts) may load this library.
#include <quadmath.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef __float128 quad;

namespace {

const double PI = 3.14159265358979323846e+00;
const double BOLTZ = 0.6950356e0;       // asymrho.f:41, symrho.f:38
const int MAXFAC = 1754;                // asymrho.f:44

// calfac, asymrho.f:1041-1051 (fact(i)=fact(i-1)*dfloat(i) in real*16)
std::vector<quad> g_fact;
void calfac() {
  if (!g_fact.empty()) return;
  g_fact.resize(MAXFAC + 1);
  g_fact[0] = 1.0Q;
  for (int i = 1; i <= MAXFAC; ++i) g_fact[i] = g_fact[i - 1] * (quad)(double)i;
}

// integer power by repeated squaring (what a Fortran x**n with integer n compiles to)
quad qpowi(quad a, int n) {
  quad pow = 1.0Q, x = a;
  unsigned u = (unsigned)(n < 0 ? -n : n);
  if (u) for (;;) {
    if (u & 1u) pow *= x;
    u >>= 1;
    if (u) x *= x; else break;
  }
  return n < 0 ? 1.0Q / pow : pow;
}

// wigd, asymrho.f:1006-1039
double wigd(int j, int m, int k, double theta) {
  calfac();
  const quad* fact = g_fact.data();
  quad pre1 = sqrtq(fact[j + k]) * sqrtq(fact[j - k]) * sqrtq(fact[j + m]) * sqrtq(fact[j - m]);
  int nulow = (0 > k - m) ? 0 : k - m;
  int nuup = (j + k < j - m) ? j + k : j - m;
  quad thehlf = 0.5Q * (quad)theta;
  quad acc = 0.0Q;
  for (int nu = nulow; nu <= nuup; ++nu) {
    quad denorm = fact[j - m - nu] * fact[j + k - nu] * fact[nu + m - k] * fact[nu] * (quad)((nu & 1) ? -1 : 1);
    quad pre2 = pre1 / denorm;
    quad cosfac = cosq(thehlf);
    quad sinfac = -sinq(thehlf);
    cosfac = qpowi(cosfac, 2 * j + k - m - 2 * nu);
    sinfac = qpowi(sinfac, m - k + 2 * nu);
    acc = acc + pre2 * cosfac * sinfac;
  }
  return (double)acc;
}

// cplus / cminus / rotmat, asymrho.f:752-813
double cplus(int j, int k) {
  if (k >= j || k < -j) return 0.0;
  double dj = j, dk = k;
  return std::sqrt(dj * (dj + 1.0) - dk * (dk + 1.0));
}
double cminus(int j, int k) {
  if (k <= -j || k > j) return 0.0;
  double dj = j, dk = k;
  return std::sqrt(dj * (dj + 1.0) - dk * (dk - 1.0));
}
double rotmat(int j, int k, int kp, double A, double B, double C) {
  int dk = std::abs(k - kp);
  if (dk != 0 && dk != 2) return 0.0;
  if (k == kp) return 0.5 * (A + C) * (double)(j * (j + 1)) + (B - 0.5 * (A + C)) * (double)(k * k);
  if (k == kp + 2) return 0.25 * (A - C) * cplus(j, kp) * cplus(j, kp + 1);
  return 0.25 * (A - C) * cminus(j, kp) * cminus(j, kp - 1);
}

// column-major (MD,MD) accessors, 1-based like the Fortran
struct Mat {
  int md; std::vector<double> a;
  explicit Mat(int n) : md(n), a((size_t)n * n, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)(j - 1) * md + (i - 1)]; }
};

// TRED2, asymrho.f:931-1004 (Householder reduction, accumulating the transformation in Z)
void tred2(int N, Mat& Z, std::vector<double>& D, std::vector<double>& E) {
  const double BETA = 1e-20;
  for (int I = N; I >= 2; --I) {
    int IM1 = I - 1, L = I - 2;
    double F = Z(I, IM1), G = 0.0, H;
    if (L > 0) for (int K = 1; K <= L; ++K) G = G + Z(I, K) * Z(I, K);
    H = G + F * F;
    if (G - BETA <= 0.0) {
      E[I] = F;
      H = 0.0;
    } else {
      L = L + 1;
      if (F >= 0.0) { E[I] = -std::sqrt(H); G = E[I]; }
      else          { E[I] = std::sqrt(H);  G = E[I]; }
      H = H - F * G;
      Z(I, IM1) = F - G;
      F = 0.0;
      for (int J = 1; J <= L; ++J) {
        Z(J, I) = Z(I, J) / H;
        G = 0.0;
        for (int K = 1; K <= J; ++K) G = G + Z(J, K) * Z(I, K);
        for (int K = J + 1; K <= L; ++K) G = G + Z(K, J) * Z(I, K);
        E[J] = G / H;
        F = F + G * Z(J, I);
      }
      double HH = F / (H + H);
      for (int J = 1; J <= L; ++J) {
        F = Z(I, J);
        E[J] = E[J] - HH * F;
        G = E[J];
        for (int K = 1; K <= J; ++K) Z(J, K) = Z(J, K) - F * E[K] - G * Z(I, K);
      }
    }
    D[I] = H;
  }
  D[1] = 0.0;
  E[1] = 0.0;
  for (int I = 1; I <= N; ++I) {
    int L = I - 1;
    if (D[I] != 0.0 && L > 0) {
      for (int J = 1; J <= L; ++J) {
        double G = 0.0;
        for (int K = 1; K <= L; ++K) G = G + Z(I, K) * Z(K, J);
        for (int K = 1; K <= L; ++K) Z(K, J) = Z(K, J) - G * Z(K, I);
      }
    }
    D[I] = Z(I, I);
    Z(I, I) = 1.0;
    if (L > 0) for (int J = 1; J <= L; ++J) { Z(I, J) = 0.0; Z(J, I) = 0.0; }
  }
}

// TQL, asymrho.f:839-929 (QL with implicit shifts, then selection sort ascending). Returns false on 'FAIL'.
bool tql(int N, Mat& Z, std::vector<double>& D, std::vector<double>& E) {
  const double EPS = 1e-12;
  const int NITER = 50;
  tred2(N, Z, D, E);
  for (int I = 2; I <= N; ++I) E[I - 1] = E[I];
  double F = 0.0, B = 0.0;
  E[N] = 0.0;
  for (int L = 1; L <= N; ++L) {
    int J = 0;
    double H = EPS * (std::fabs(D[L]) + std::fabs(E[L]));
    int LP1 = L + 1;
    if (B - H < 0.0) B = H;
    int M;
    for (M = L; M <= N; ++M) if (std::fabs(E[M]) - B <= 0.0) break;
    if (M != L) {
      for (;;) {
        if (J == NITER) return false;
        J = J + 1;
        double P = (D[LP1] - D[L]) / (2 * E[L]);
        double R = std::sqrt(P * P + 1);
        if (P < 0.0) H = D[L] - E[L] / (P - R);
        else         H = D[L] - E[L] / (P + R);
        for (int I = L; I <= N; ++I) D[I] = D[I] - H;
        F = F + H;
        P = D[M];
        double C = 1.0, S = 0.0;
        int MM1 = M - 1;
        if (MM1 - L >= 0) {
          for (int LMIP = L; LMIP <= MM1; ++LMIP) {
            int I = L + MM1 - LMIP, IP1 = I + 1;
            double G = C * E[I];
            H = C * P;
            if (std::fabs(P) - std::fabs(E[I]) >= 0.0) {
              C = E[I] / P;
              R = std::sqrt(C * C + 1.0);
              E[IP1] = S * P * R;
              S = C / R;
              C = 1.0 / R;
            } else {
              C = P / E[I];
              R = std::sqrt(C * C + 1);
              E[IP1] = S * E[I] * R;
              S = 1 / R;
              C = C / R;
            }
            P = C * D[I] - S * G;
            D[IP1] = H + S * (C * G + S * D[I]);
            for (int K = 1; K <= N; ++K) {
              H = Z(K, IP1);
              Z(K, IP1) = S * Z(K, I) + C * H;
              Z(K, I) = C * Z(K, I) - S * H;
            }
          }
        }
        E[L] = S * P;
        D[L] = C * P;
        if (std::fabs(E[L]) - B <= 0.0) break;
      }
    }
    D[L] = D[L] + F;
  }
  for (int I = 1; I <= N; ++I) {
    int K = I;
    double P = D[I];
    for (int J = I + 1; J <= N; ++J) if (D[J] - P < 0.0) { K = J; P = D[J]; }
    if (K != I) {
      D[K] = D[I];
      D[I] = P;
      for (int J = 1; J <= N; ++J) { P = Z(J, I); Z(J, I) = Z(J, K); Z(J, K) = P; }
    }
  }
  return true;
}

struct Asym {
  double temprt, beta, tau, A, B, C;
  int nslice, iodevn, maxj, jmax;
  std::vector<double> engevn, engodd, eigevn, eigodd;
  int nstev = 0, nstod = 0;
  // dlist(j,m,k) for ONE theta (the reference keeps all thetas of the run; a run is one theta in practice)
  int dl_ithe = -1;
  std::vector<double> dlist;
  double& dl(int j, int m, int k) { int w = 2 * maxj + 1; return dlist[((size_t)j * w + (m + maxj)) * w + (k + maxj)]; }
  double info[16];
};

}  // namespace

extern "C" {

double tg_wigd(int j, int m, int k, double theta) { return wigd(j, m, k, theta); }

// asymrho.f:93-425: blocks, eigen-decomposition, emax checks, partition sums.
// info[0..5]  = AT BETA: Z_even, E_even (cm-1), Cv_even, Z_odd, E_odd, Cv_odd ; info[6..8] = classical Z, E, Cv
// info[9..14] = AT TAU : Z_even, E_even, Z_odd, E_odd, Z_cl, E_cl ; info[15] = emax
// returns NULL (and a message in err) where the Fortran would STOP.
void* tg_asym_setup(double temprt, int nslice, int iodevn, double Arot, double Brot, double Crot, int maxj, double* info,
                    char* err, int errlen) {
  auto fail = [&](const char* m) -> void* { if (err) snprintf(err, errlen, "%s", m); return nullptr; };
  if (maxj > 876) return fail("maxj is larger than the limit of 876");
  if (iodevn > 1 || iodevn < -1) return fail("iodevn can only be -1 0 1");
  Asym* S = new Asym;
  S->temprt = temprt; S->nslice = nslice; S->iodevn = iodevn; S->A = Arot; S->B = Brot; S->C = Crot; S->maxj = maxj;
  S->beta = 1.0 / (BOLTZ * temprt);
  S->tau = S->beta / (double)nslice;
  const int maxd = 2 * maxj + 1;
  Mat H(maxd);
  std::vector<double> eigval(maxd + 2), work(maxd + 2);
  for (int j = 0; j <= maxj; ++j) {
    int kevnst, koddst, ndimev, ndimod;
    if (j % 2 == 0) { kevnst = -j; koddst = -j + 1; ndimev = j + 1; ndimod = j; }
    else            { kevnst = -j + 1; koddst = -j; ndimev = j; ndimod = j + 1; }
    for (int parity = 0; parity < 2; ++parity) {
      if (parity == 1 && j == 0) break;
      int kst = parity ? koddst : kevnst, ndim = parity ? ndimod : ndimev;
      int irow = 0;
      for (int k = kst; k <= j; k += 2) {
        ++irow;
        int jcol = 0;
        for (int kp = kst; kp <= k; kp += 2) {
          ++jcol;
          double e = rotmat(j, k, kp, Arot, Brot, Crot);
          H(irow, jcol) = e;
          H(jcol, irow) = e;
        }
      }
      if (irow != ndim) { delete S; return fail("wrong dimension"); }
      if (!tql(ndim, H, eigval, work)) { delete S; return fail("  FAIL"); }
      std::vector<double>& eng = parity ? S->engodd : S->engevn;
      std::vector<double>& eig = parity ? S->eigodd : S->eigevn;
      for (int ist = 1; ist <= ndim; ++ist) {
        eng.push_back(eigval[ist]);
        for (int ibs = 1; ibs <= ndim; ++ibs) eig.push_back(H(ibs, ist));
      }
    }
  }
  S->jmax = maxj;
  S->nstev = (int)S->engevn.size();
  S->nstod = (int)S->engodd.size();
  double emax = -1e300;      // bubble_sort(esort); emax = esort(nsttot)
  for (double e : S->engevn) if (e > emax) emax = e;
  for (double e : S->engodd) if (e > emax) emax = e;
  if (std::exp(-S->beta * emax) > 1e-8 || std::exp(-S->tau * emax) > 1e-8) { delete S; return fail("too large contribution from emax"); }
  // partition sums, asymrho.f:299-425
  for (int pass = 0; pass < 2; ++pass) {
    double b = pass ? S->tau : S->beta;
    double zparev = 0, zparod = 0, eavrev = 0, eavrod = 0, esqevn = 0, esqodd = 0;
    int istevn = 0, istodd = 0;
    for (int j = 0; j <= S->jmax; ++j) {
      int ndimev = (j % 2 == 0) ? j + 1 : j, ndimod = (j % 2 == 0) ? j : j + 1;
      int ndegen = 2 * j + 1;
      double sumevn = 0, sengev = 0, seevsq = 0;
      for (int i = 1; i <= ndimev; ++i) {
        double energy = S->engevn[istevn + i - 1];
        sumevn = sumevn + std::exp(-b * energy);
        sengev = sengev + energy * std::exp(-b * energy);
        seevsq = seevsq + energy * energy * std::exp(-b * energy);
      }
      zparev = zparev + ndegen * sumevn; eavrev = eavrev + ndegen * sengev; esqevn = esqevn + ndegen * seevsq;
      istevn += ndimev;
      double sumodd = 0, sengod = 0, seodsq = 0;
      for (int i = 1; i <= ndimod; ++i) {
        double energy = S->engodd[istodd + i - 1];
        sumodd = sumodd + std::exp(-b * energy);
        sengod = sengod + energy * std::exp(-b * energy);
        seodsq = seodsq + energy * energy * std::exp(-b * energy);
      }
      zparod = zparod + ndegen * sumodd; eavrod = eavrod + ndegen * sengod; esqodd = esqodd + ndegen * seodsq;
      istodd += ndimod;
    }
    double eavrcl = eavrev + eavrod, esqcla = esqevn + esqodd;
    eavrev = eavrev / zparev; eavrod = eavrod / zparod; esqevn = esqevn / zparev; esqodd = esqodd / zparod;
    double kt2 = BOLTZ * BOLTZ * temprt * temprt;
    double zparcl = zparev + zparod;
    eavrcl = eavrcl / zparcl; esqcla = esqcla / zparcl;
    if (pass == 0) {
      double v[9] = {zparev, eavrev, (esqevn - eavrev * eavrev) / kt2, zparod, eavrod, (esqodd - eavrod * eavrod) / kt2,
                     zparcl, eavrcl, (esqcla - eavrcl * eavrcl) / kt2};
      memcpy(S->info, v, sizeof v);
    } else {
      double v[6] = {zparev, eavrev, zparod, eavrod, zparcl, eavrcl};
      memcpy(S->info + 9, v, sizeof v);
    }
  }
  S->info[15] = emax;
  if (info) memcpy(info, S->info, sizeof S->info);
  return S;
}

void tg_asym_free(void* h) { delete (Asym*)h; }

int tg_asym_nstates(void* h, int parity) { Asym* S = (Asym*)h; return parity ? S->nstod : S->nstev; }
void tg_asym_energies(void* h, int parity, double* out) {
  Asym* S = (Asym*)h;
  const std::vector<double>& e = parity ? S->engodd : S->engevn;
  memcpy(out, e.data(), e.size() * sizeof(double));
}

// asymrho.f:102-112 for one theta
void tg_asym_dlist(void* h, int ithe) {
  Asym* S = (Asym*)h;
  if (S->dl_ithe == ithe) return;
  int w = 2 * S->maxj + 1;
  S->dlist.assign((size_t)(S->maxj + 1) * w * w, 0.0);
  double th = (double)ithe * PI / 180.0;
  for (int j = 0; j <= S->maxj; ++j)
    for (int m = -j; m <= j; ++m)
      for (int k = -j; k <= j; ++k) S->dl(j, m, k) = wigd(j, m, k, th);
  S->dl_ithe = ithe;
}

// one grid point, asymrho.f:486-659: out = {rho, erot, esq}
void tg_asym_point(void* h, int ithe, int iphi, int ichi, double* out) {
  Asym* S = (Asym*)h;
  const double eps = 1e-16;
  tg_asym_dlist(h, ithe);
  double phi = (double)iphi * PI / 180.0, chi = (double)ichi * PI / 180.0;
  double rhoevn = 0, rhoodd = 0, rotevn = 0, rotodd = 0, esqevn = 0, esqodd = 0;
  for (int parity = 0; parity < 2; ++parity) {
    if (parity == 0 && S->iodevn == 1) continue;
    if (parity == 1 && S->iodevn == 0) continue;
    const std::vector<double>& eng = parity ? S->engodd : S->engevn;
    const std::vector<double>& eig = parity ? S->eigodd : S->eigevn;
    double rhop = 0, rotp = 0, esqp = 0;
    int istate = 0, ivec = 0;
    for (int j = 0; j <= S->jmax; ++j) {
      int kst, ndim;
      if (j % 2 == 0) { kst = parity ? -j + 1 : -j; ndim = parity ? j : j + 1; }
      else            { kst = parity ? -j : -j + 1; ndim = parity ? j + 1 : j; }
      double pre = (double)(2 * j + 1) / (8.0 * PI * PI);
      double rho1 = 0, erot1 = 0, esq1 = 0;
      for (int ist = 1; ist <= ndim; ++ist) {
        double energy = eng[istate++];
        double expo = std::exp(-S->tau * energy);
        if (!(expo * pre < eps)) {
          int im = 0;
          double rho2 = 0;
          for (int m = kst; m <= j; m += 2) {
            ++im;
            double coef1 = eig[im + ivec - 1];
            int ik = 0;
            double rho3 = 0;
            for (int k = kst; k <= j; k += 2) {
              ++ik;
              double coef2 = eig[ik + ivec - 1];
              rho3 = rho3 + coef2 * S->dl(j, m, k) * std::cos((double)m * phi + (double)k * chi);
            }
            rho2 = rho2 + coef1 * rho3;
          }
          rho1 = rho1 + rho2 * expo;
          erot1 = erot1 + expo * rho2 * energy;
          esq1 = esq1 + expo * rho2 * energy * energy;
        }
        ivec += ndim;
      }
      rhop = rhop + pre * rho1; rotp = rotp + pre * erot1; esqp = esqp + pre * esq1;
    }
    if (std::fabs(rhop) > 1.0e-16) { rotp = rotp / rhop; esqp = esqp / rhop; }
    else { rotp = 0; esqp = 0; }
    if (parity) { rhoodd = rhop; rotodd = rotp; esqodd = esqp; }
    else        { rhoevn = rhop; rotevn = rotp; esqevn = esqp; }
  }
  if (S->iodevn == 0) { out[0] = rhoevn; out[1] = rotevn; out[2] = esqevn; return; }
  if (S->iodevn == 1) { out[0] = rhoodd; out[1] = rotodd; out[2] = esqodd; return; }
  double rhocla = rhoodd + rhoevn;
  out[0] = rhoevn + rhoodd;
  out[1] = (rotodd * rhoodd + rotevn * rhoevn) / (rhoevn + rhoodd);
  out[2] = (esqodd * rhoodd + esqevn * rhoevn) / (rhoevn + rhoodd);
  if (std::fabs(rhocla) < 1.0e-16) { out[1] = 0; out[2] = 0; }
}

// the chi range the reference computes directly for a given phi, asymrho.f:475-483
int tg_asym_maxchi(int iphi) {
  if (iphi >= 0 && iphi <= 90) return iphi;
  if (iphi > 90 && iphi <= 180) return 180 - iphi;
  if (iphi > 180 && iphi <= 270) return iphi - 180;
  return 360 - iphi;
}

// the four sequential symmetry passes, asymrho.f:665-709, on one [361][361] plane (phi outer, chi inner), in place
void tg_asym_symfill(double* p) {
  auto at = [&](int iphi, int ichi) -> double& { return p[iphi * 361 + ichi]; };
  for (int iphi = 90; iphi <= 180; ++iphi)
    for (int ichi = 180 - iphi; ichi <= iphi; ++ichi) at(iphi, ichi) = at(180 - ichi, 180 - iphi);
  for (int iphi = 180; iphi <= 270; ++iphi)
    for (int ichi = iphi - 180; ichi <= 360 - iphi; ++ichi) at(iphi, ichi) = at(180 + ichi, iphi - 180);
  for (int iphi = 180; iphi <= 360; ++iphi)
    for (int ichi = 360 - iphi; ichi <= iphi; ++ichi) at(iphi, ichi) = at(360 - ichi, 360 - iphi);
  for (int iphi = 0; iphi <= 360; ++iphi)
    for (int ichi = iphi; ichi <= 360; ++ichi) at(iphi, ichi) = at(ichi, iphi);
}

// whole theta plane (direct region + symmetry fill); planes are [361][361]; stride thins the work for tests
// (points with iphi%stride || ichi%stride are left 0 and the fill is skipped when stride>1)
void tg_asym_plane(void* h, int ithe, int stride, double* rho, double* eng, double* esq) {
  memset(rho, 0, 361 * 361 * sizeof(double));
  memset(eng, 0, 361 * 361 * sizeof(double));
  memset(esq, 0, 361 * 361 * sizeof(double));
  for (int iphi = 0; iphi <= 360; iphi += stride)
    for (int ichi = 0; ichi <= tg_asym_maxchi(iphi); ichi += stride) {
      double o[3];
      tg_asym_point(h, ithe, iphi, ichi, o);
      rho[iphi * 361 + ichi] = o[0]; eng[iphi * 361 + ichi] = o[1]; esq[iphi * 361 + ichi] = o[2];
    }
  if (stride == 1) { tg_asym_symfill(rho); tg_asym_symfill(eng); tg_asym_symfill(esq); }
}

// symrho.f:31-168 for one theta: planes [361][361]; info = {ztau, zbeta, Ebeta (K), Esqrt (K^2), Cv}
// returns 0, or 1 where the Fortran stops with 'pmax too large'
int tg_symrho_plane(double temprt, int nslice, int kmod, int ith, double Bz, double Bxy, int maxj, double* rhopro,
                    double* erotpr, double* erotsq, double* info) {
  const double eps = 1e-16;
  double beta = 1.0 / (BOLTZ * temprt);
  double tau = beta / (double)nslice;
  std::vector<double> dlist((size_t)(maxj + 1) * (2 * maxj + 1), 0.0);
  auto dl = [&](int j, int k) -> double& { return dlist[(size_t)j * (2 * maxj + 1) + k + maxj]; };
  double th = (double)ith * PI / 180.0;
  for (int j = 0; j <= maxj; ++j)
    for (int k = -j; k <= j; ++k) dl(j, k) = wigd(j, k, k, th);
  double ztau = 0, zbeta = 0, Ebeta = 0, Esqrt = 0;
  for (int j = 0; j <= maxj; ++j)
    for (int k = 0; k <= j; ++k)
      if (k % kmod == 0) {
        int kgen = 2 - (k == 0 ? 1 : 0);
        double ejk = Bxy * j * (j + 1) + (Bz - Bxy) * k * k;
        ztau = ztau + kgen * std::exp(-tau * ejk) * (2 * j + 1);
        zbeta = zbeta + kgen * std::exp(-beta * ejk) * (2 * j + 1);
        Ebeta = Ebeta + kgen * std::exp(-beta * ejk) * (2 * j + 1) * ejk;
        Esqrt = Esqrt + kgen * std::exp(-beta * ejk) * (2 * j + 1) * ejk * ejk;
      }
  Ebeta = Ebeta / zbeta;
  Esqrt = Esqrt / zbeta;
  Ebeta = Ebeta / BOLTZ;
  Esqrt = Esqrt / (BOLTZ * BOLTZ);
  double Cv = (Esqrt - Ebeta * Ebeta) / (temprt * temprt);
  if (info) { info[0] = ztau; info[1] = zbeta; info[2] = Ebeta; info[3] = Esqrt; info[4] = Cv; }
  double emax = (Bz > Bxy) ? Bxy * maxj * (maxj + 1) + (Bz - Bxy) * maxj * maxj : Bxy * maxj * (maxj + 1);
  double pmax = (2 * maxj + 1) * std::exp(-tau * emax) / ztau;
  if (pmax > eps) return 1;
  memset(rhopro, 0, 361 * 361 * sizeof(double));
  memset(erotpr, 0, 361 * 361 * sizeof(double));
  memset(erotsq, 0, 361 * 361 * sizeof(double));
  for (int icp = 0; icp <= 360; ++icp) {
    double cph = (double)icp * PI / 180.0;
    double rho = 0, erot = 0, esq = 0;
    for (int j = 0; j <= maxj; ++j) {
      double pre = (double)(2 * j + 1) / (8.0 * PI * PI);
      for (int k = 0; k <= j; ++k)
        if (k % kmod == 0) {
          int kgen = 2 - (k == 0 ? 1 : 0);
          double ejk = Bxy * j * (j + 1) + (Bz - Bxy) * k * k;
          double t = pre * kgen * dl(j, k) * std::cos(k * cph) * std::exp(-tau * ejk);
          rho = rho + t;
          erot = erot + t * ejk;
          esq = esq + t * ejk * ejk;
        }
    }
    erot = erot / rho;
    esq = esq / rho;
    for (int iph = 0; iph <= 360; ++iph)
      for (int ich = 0; ich <= iph; ++ich)
        if ((iph + ich) % 360 == icp) {
          rhopro[iph * 361 + ich] = rho; erotpr[iph * 361 + ich] = erot; erotsq[iph * 361 + ich] = esq;
          rhopro[ich * 361 + iph] = rho; erotpr[ich * 361 + iph] = erot; erotsq[ich * 361 + iph] = esq;
        }
  }
  return 0;
}

// linden.f: lgnd :180-193, exarho :88-122
static void lgnd(int lmax, double x, double* p) {
  p[0] = 1.0;
  p[1] = x;
  for (int l = 1; l <= lmax - 1; ++l) p[l + 1] = ((double)(2.0f * l + 1) * x * p[l] - l * p[l - 1]) / (l + 1);
}
static void exarho(double cost, int lmax, double* pl, double* rho_, double* erot_, double tau, double bconst, int iodevn,
                   int nslice, double* erotsq_) {
  const double boltz = 0.69503476e0;          // linden.f:92 (differs from asymrho's 0.6950356)
  lgnd(lmax, cost, pl);
  double rho = 0, erot = 0, erotsq = 0;
  for (int l = 0; l <= lmax; ++l)
    if ((l % 2 == iodevn) || iodevn == -1) {
      double tmp = (double)(2 * l + 1) * pl[l];
      tmp = tmp * std::exp(-tau * bconst * (double)(l * (l + 1)));
      rho = rho + tmp;
      erot = erot + tmp * l * (l + 1) * bconst;
      erotsq = erotsq + tmp * l * (l + 1) * bconst * l * (l + 1) * bconst;
    }
  erot = erot / (nslice * boltz);
  erotsq = erotsq / std::pow(nslice * boltz, 2.0);
  rho = rho / ((double)4.0f * PI);
  erot = erot / ((double)4.0f * PI);
  erotsq = erotsq / ((double)4.0f * PI);
  *rho_ = rho; *erot_ = erot; *erotsq_ = erotsq;
}

// linden.f:22-82: out[npt][4] = cost, rho, erot, erotsq ; info = {tau, lmax, Erot at beta, Cv at beta}
void tg_linden(double temprt, int nslice, double bconst, int npt, int iodevn, double* out, double* info) {
  const int maxl = 500;
  const double taunit = 1.4387752224e+00, eps = 1e-16;
  std::vector<double> pl(maxl + 2);
  double tau = taunit / (temprt * nslice);
  int lmax = maxl;
  for (int l = 0; l <= maxl; ++l)
    if (std::exp(-tau * bconst * l * (l + 1)) < eps) { lmax = l; break; }
  double cstep = (double)2.0f / (double)(npt - 1);
  for (int ic = 1; ic <= npt; ++ic) {
    double cost = (ic - 1) * cstep - 1.0;
    double rho, erot1, erotsq;
    exarho(cost, lmax, pl.data(), &rho, &erot1, tau, bconst, iodevn, nslice, &erotsq);
    out[(ic - 1) * 4 + 0] = cost; out[(ic - 1) * 4 + 1] = rho; out[(ic - 1) * 4 + 2] = erot1; out[(ic - 1) * 4 + 3] = erotsq;
  }
  double cost = 1.0, beta = tau * nslice, rho, erot, erotsq;
  exarho(cost, lmax, pl.data(), &rho, &erot, beta, bconst, iodevn, nslice, &erotsq);
  erot = erot * nslice / rho;
  erotsq = erotsq * nslice * nslice / rho;
  double Cv = (erotsq - erot * erot) / std::pow(temprt, 2.0);
  if (info) { info[0] = tau; info[1] = lmax; info[2] = erot; info[3] = Cv; }
}

// Fortran edit descriptors used by the generators' writers: E15.8 ("0.dddddddde+xx", asymrho.f:721-723)
// and 1P,E15.8 ("d.dddddddde+xx", linden.f:68). buf needs 32 bytes.
void tg_fmt_e15_8(double v, int scale1p, char* buf) {
  char t[40];
  if (scale1p) { snprintf(buf, 32, "%15.8E", v); return; }
  snprintf(t, sizeof t, "%.7E", v);                 // [-]d.dddddddE[+-]xx
  const char* s = t;
  bool neg = (*s == '-');
  if (neg) ++s;
  char dig[9] = {s[0], s[2], s[3], s[4], s[5], s[6], s[7], s[8], 0};
  int ex = atoi(strchr(s, 'E') + 1);
  if (v != 0.0) ex += 1;
  char body[40];
  if (ex > -100 && ex < 100) snprintf(body, sizeof body, "%s0.%sE%c%02d", neg ? "-" : "", dig, ex < 0 ? '-' : '+', std::abs(ex));
  else snprintf(body, sizeof body, "%s0.%s%c%03d", neg ? "-" : "", dig, ex < 0 ? '-' : '+', std::abs(ex));
  snprintf(buf, 32, "%15s", body);
}

}  // extern "C"
