// ORACLE (test infrastructure, not product code).
//
// Harness that drives the UNMODIFIED reference C++ translation units (compiled
// from /root/reference by oracle/Makefile into oracle/_ref/libpimcref.so).
// It only *calls* reference functions and pokes the reference's global state
// (mc_setup.h:176-191); no reference source is copied.  Used
//   * by tests (here, where /root/reference exists) to pin oracle/pimc_oracle.cpp,
//   * by bench.py --impl reference / cpu_baseline to time the reference's own
//     CPU hot path (PIMCPass loop, mc_main.cc:349-381) on the host cores.
//
// The serial SPRNG wrappers of mc_randg.cc are compiled under renamed symbols
// (sprng_rnd1 ...) and re-exported here through a switch: queue mode feeds
// test-supplied uniforms, otherwise the vendored SPRNG streams are used.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <string>
#include <unistd.h>
#include <omp.h>

#include "mc_confg.h"
#include "mc_setup.h"
#include "mc_input.h"
#include "mc_utils.h"
#include "mc_poten.h"
#include "mc_piqmc.h"
#include "mc_qworm.h"
#include "mc_estim.h"
#include "mc_const.h"
#include "rngstream.h"
#include "omprng.h"

// ---- renamed originals from mc_randg.cc (see oracle/Makefile -D flags) ----
double sprng_rnd(void);  double sprng_rnd1(void); double sprng_rnd2(void); double sprng_rnd3(void);
double sprng_rnd4(void); double sprng_rnd5(void); double sprng_rnd6(void); double sprng_rnd7(void);
double sprng_gauss(double);
int sprng_nrnd1(int); int sprng_nrnd2(int); int sprng_nrnd3(int); int sprng_nrnd4(int);
void RandomInit(int, int);

// reference functions without a header declaration
void PIMCPass(int, int);
void MCGetAverage(void);
void MCResetBlockAverage(void);
void InitTotalAverage(void);
void init_pot1D(int);
void init_pot2D(int);
void init_pot3D(int);
extern double avergCount, totalCount;
extern double _bpot, _bkin, _brot, _brotsq, _bCv, _bCv_trans, _bCv_rot;
extern double **_rcf;
extern double **_gr1D;
extern double ***_gr2D;
extern double **_gr3D;
extern double *_relthe_sum, *_relphi_sum, *_relchi_sum;
extern double *_ploops;
extern double _areas[2], _area2[2], _inert[2];
void qworm_open(void); void qworm_close(void); void qworm_advance(void); void qworm_recede(void); void qworm_swap(void);

static int g_queue_mode = 0;
static std::deque<double> g_q[15];

static double popq(int s)
{
   if (g_q[s].empty()) { printf("ref_harness: RNG queue %d empty\n", s); exit(2); }
   double v = g_q[s].front(); g_q[s].pop_front(); return v;
}

double rnd(void)  { return g_queue_mode ? popq(0) : sprng_rnd(); }
double rnd1(void) { return g_queue_mode ? popq(1) : sprng_rnd1(); }
double rnd2(void) { return g_queue_mode ? popq(2) : sprng_rnd2(); }
double rnd3(void) { return g_queue_mode ? popq(3) : sprng_rnd3(); }
double rnd4(void) { return g_queue_mode ? popq(4) : sprng_rnd4(); }
double rnd5(void) { return g_queue_mode ? popq(5) : sprng_rnd5(); }
double rnd6(void) { return g_queue_mode ? popq(6) : sprng_rnd6(); }
double rnd7(void) { return g_queue_mode ? popq(7) : sprng_rnd7(); }
double gauss(double alpha)
{
   if (!g_queue_mode) return sprng_gauss(alpha);
   // mc_randg.cc:138-150 with the two uniforms taken from queues 8 and 9
   double r1 = popq(8), r2 = popq(9);
   double x1 = sqrt(-log(r1)) * cos(2.0 * M_PI * r2);
   return (x1 / sqrt(alpha));
}
int nrnd1(int n) { return g_queue_mode ? (int)floor(n * popq(10)) : sprng_nrnd1(n); }
int nrnd2(int n) { return g_queue_mode ? (int)floor(n * popq(11)) : sprng_nrnd2(n); }
int nrnd3(int n) { return g_queue_mode ? (int)floor(n * popq(12)) : sprng_nrnd3(n); }
int nrnd4(int n) { return g_queue_mode ? (int)floor(n * popq(13)) : sprng_nrnd4(n); }

extern "C" {

void ref_rng_queue_mode(int on) { g_queue_mode = on; }
void ref_rng_push(int stream, const double *u, int n) { for (int i = 0; i < n; i++) g_q[stream].push_back(u[i]); }
void ref_rng_clear(void) { for (int s = 0; s < 15; s++) g_q[s].clear(); }
int  ref_rng_pending(int stream) { return (int)g_q[stream].size(); }

// Set-up sequence of mc_main.cc:103-127,151-155,233-295 driven from `workdir`
// (must hold qmc.input and the small table files).  Big tables (3-D potential,
// top density matrices) may be injected instead of parsed: pass vtable!=NULL /
// rho!=NULL.
int ref_init(const char *workdir, double *vtab, int rg, int thg, int chg, double rvmin, double rvmax,
             double *rho, double *erot, double *esq, int nthreads)
{
   if (chdir(workdir) != 0) { printf("ref_init: cannot chdir to %s\n", workdir); return 1; }
   if (nthreads > 0) omp_set_num_threads(nthreads);
   MPIsize = 1; MPIrank = MPI_MASTER;
   int restart = 0;
   IOReadParams(FINPUT, restart);
   MCInitParams();
   MCSetUnits();
   MCMemAlloc();
   MemAllocMCCounts();
   MemAllocQWCounts();
   MCInit();
   if (WORM) MCWormInit();
   SEED = 985456376;
   RandomInit(MPIrank, MPIsize);
   MCConfigInit();

   // InitPotentials (mc_poten.cc:93-119) with optional injection of the 3-D table
   for (int atype = 0; atype < NumbTypes; atype++) {
      if (MCAtom[atype].molecule == 1) init_pot2D(atype);
      else if (MCAtom[atype].molecule == 2) {
         if (NumbTypes > 1) {
            if (vtab) {
               vtable = vtab; Rgrd = rg; THgrd = thg; CHgrd = chg; Rvmin = rvmin; Rvmax = rvmax;
               Rvstep = (Rvmax - Rvmin) / (double)(Rgrd - 1);
            } else if (!ISPHER) init_pot3D(atype);      // ISPHER = 1 never reads vtable (vspher_ carries its own DATA table)
         }
      } else init_pot1D(atype);
   }
   if (ROTATION) {
      if (MCAtom[IMTYPE].molecule == 2 && RotDenType == 0 && rho) {
         rhoprp = rho; erotpr = erot; erotsq = esq;   // MCMemAlloc's untouched buffers are leaked
      } else InitRotDensity();
   }
   InitMCEstims();
   InitTotalAverage();
   ResetMCCounts();
   ResetQWCounts();
   fixedseed();                 // omprng.cc:14-18 instead of the wall-clock randomseed()
   MCResetBlockAverage();
   return 0;
}

int ref_numb_atoms(void) { return NumbAtoms; }
int ref_numb_times(void) { return NumbTimes; }
int ref_numb_rot_times(void) { return NumbRotTimes; }
int ref_numb_types(void) { return NumbTypes; }
double ref_tau(void) { return MCTau; }
double ref_rot_tau(void) { return MCRotTau; }
double ref_lambda(int type) { return MCAtom[type].lambda; }

// state in the reference layout [dim][atom*P + it]
void ref_set_state(const double *coords, const double *angles, const int *pindex)
{
   int n = NumbAtoms * NumbTimes;
   for (int d = 0; d < 3; d++)
      for (int i = 0; i < n; i++) {
         MCCoords[d][i] = coords[d * n + i];
         MCAngles[d][i] = angles[d * n + i];
      }
   for (int i = 0; i < n; i++) {                // mc_main.cc:192-199
      double phi = MCAngles[PHI][i], cost = MCAngles[CTH][i];
      double sint = sqrt(1.0 - cost * cost);
      MCCosine[AXIS_X][i] = sint * cos(phi);
      MCCosine[AXIS_Y][i] = sint * sin(phi);
      MCCosine[AXIS_Z][i] = cost;
   }
   if (pindex && BOSONS)
      for (int a = 0; a < MCAtom[BSTYPE].numb; a++) { PIndex[a] = pindex[a]; RIndex[pindex[a]] = a; }
}
void ref_get_state(double *coords, double *angles, double *cosine)
{
   int n = NumbAtoms * NumbTimes;
   for (int d = 0; d < 3; d++)
      for (int i = 0; i < n; i++) {
         coords[d * n + i] = MCCoords[d][i];
         angles[d * n + i] = MCAngles[d][i];
         cosine[d * n + i] = MCCosine[d][i];
      }
}
void ref_get_initial_lattice(double *coords)
{
   int n = NumbAtoms * NumbTimes;
   for (int d = 0; d < 3; d++) for (int i = 0; i < n; i++) coords[d * n + i] = MCCoords[d][i];
}

// leaf functions
double ref_SPot1D(double r, int type) { return SPot1D(r, type); }
double ref_LPot2D(double r, double c, int type) { return LPot2D(r, c, type); }
double ref_SRotDens(double g, int type) { return SRotDens(g, type); }
double ref_SRotDensDeriv(double g, int type) { return SRotDensDeriv(g, type); }
double ref_SRotDensEsqrt(double g, int type) { return SRotDensEsqrt(g, type); }

// per-bead potential sums (mc_piqmc.cc:1201-1383,1796-2151) on the current MCCoords
double ref_PotEnergy_it(int atom, int it) { return PotEnergy(atom, MCCoords, it); }
double ref_PotEnergy_path(int atom) { return PotEnergy(atom, MCCoords); }
double ref_PotRotEnergy(int atom, int it) { return PotRotEnergy(atom, MCCosine, it); }
double ref_PotRotE3D(int atom, const double *eul, int it)
{
   double e[3] = {eul[0], eul[1], eul[2]};
   return PotRotE3D(atom, e, it);
}

// moves
void ref_MCMolecularMove(int type) { MCMolecularMove(type); }
void ref_MCBisectionMove(int type, int time) { MCBisectionMove(type, time); }
void ref_MCMolecularMoveExchange(int type) { MCMolecularMoveExchange(type); }
void ref_MCBisectionMoveExchange(int type, int time) { MCBisectionMoveExchange(type, time); }
int ref_MCRot3Dstep(int it1, int atom0, int type, double r1, double r2, double r3, double r4)
{
   int offset = MCAtom[type].offset + NumbTimes * atom0;
   int gatom = offset / NumbTimes;
   double tot = 0, acp = 0;
   MCRot3Dstep(it1, offset, gatom, type, MCAtom[type].rtstep, r1, r2, r3, r4, IROTSYM, NFOLD_ROT, tot, acp);
   return (int)acp;
}
int ref_MCRotLinStep(int it1, int type, double r1, double r2, double r3)
{
   int offset = MCAtom[type].offset;
   int gatom = offset / NumbTimes;
   double tot = 0, acp = 0;
   MCRotLinStep(it1, offset, gatom, type, MCAtom[type].rtstep, r1, r2, r3, tot, acp);
   return (int)acp;
}
void ref_counters(double *tot, double *acc)
{
   for (int t = 0; t < NumbTypes; t++)
      for (int m = 0; m < MCMAXMOVES; m++) { tot[t * MCMAXMOVES + m] = MCTotal[t][m]; acc[t * MCMAXMOVES + m] = MCAccep[t][m]; }
}

// estimators (mc_estim.cc)
double ref_GetKinEnergy(void) { return GetKinEnergy(); }
double ref_GetPotEnergy(void) { return GetPotEnergy(); }
double ref_GetPotEnergy_Densities(void) { return GetPotEnergy_Densities(); }
double ref_GetRotEnergy(double *erotsq, double *eterm)
{
   double s = GetRotEnergy(); *erotsq = ErotSQ; *eterm = Erot_termSQ; return s;
}
double ref_GetRotE3D(double *erotsq, double *eterm)
{
   double s = GetRotE3D(); *erotsq = ErotSQ; *eterm = Erot_termSQ; return s;
}
void ref_GetRCF(double *rcf0)
{
   for (int i = 0; i < NumbRotTimes; i++) _rcf[0][i] = 0.0;
   GetRCF();
   for (int i = 0; i < NumbRotTimes; i++) rcf0[i] = _rcf[0][i];
}
void ref_reset_block(void) { MCResetBlockAverage(); }
void ref_get_gr1D(double *out) { for (int i = 0; i < 300; i++) out[i] = _gr1D[0][i]; }
void ref_get_gr2D(double *out) { for (int i = 0; i < 300; i++) for (int j = 0; j < 50; j++) out[i * 50 + j] = _gr2D[0][i][j]; }
void ref_get_gr3D(int dtype, double *out) { memcpy(out, _gr3D[dtype], sizeof(double) * 300 * 50 * 100); }
void ref_get_relbins(double *the, double *phi, double *chi)
{
   for (int i = 0; i < 50; i++) the[i] = _relthe_sum[i];
   for (int i = 0; i < 100; i++) { phi[i] = _relphi_sum[i]; chi[i] = _relchi_sum[i]; }
}
void ref_zero_relbins(void)
{
   for (int i = 0; i < 50; i++) _relthe_sum[i] = 0;
   for (int i = 0; i < 100; i++) { _relphi_sum[i] = 0; _relchi_sum[i] = 0; }
}
// The reference's own density writers (mc_estim.cc:1327-1820,1930-1995) on test-supplied histograms: block and accumulated
// arrays are both set to the given counts; files appear under `prefix` exactly as MCSaveBlockAverages / main write them.
extern "C++" {
extern double **_gr1D_sum;
extern double ***_gr2D_sum;
extern double **_gr3D_sum;
}
void ref_save_densities(const char *prefix, double acount, const double *gr1d, const double *gr2d, const double *gr3d, const double *rel)
{
   for (int id = 0; id < NumbTypes; id++) {
      if (MCAtom[id].molecule) continue;
      for (int i = 0; i < 300; i++) { _gr1D[id][i] = gr1d[i]; _gr1D_sum[id][i] = gr1d[i]; }
   }
   if (IMPURITY && MCAtom[IMTYPE].molecule == 1)
      for (int i = 0; i < 300; i++) for (int j = 0; j < 50; j++) { _gr2D[0][i][j] = gr2d[i * 50 + j]; _gr2D_sum[0][i][j] = gr2d[i * 50 + j]; }
   if (IMPURITY && MCAtom[IMTYPE].molecule == 2) {
      for (int id = 0; id < NumbTypes; id++)
         for (long i = 0; i < 300L * 50 * 100; i++) { double v = (MCAtom[id].molecule == 0) ? gr3d[i] : 0.0; _gr3D[id][i] = v; _gr3D_sum[id][i] = v; }
      for (int i = 0; i < 50; i++) _relthe_sum[i] = rel[i];
      for (int i = 0; i < 100; i++) { _relphi_sum[i] = rel[50 + i]; _relchi_sum[i] = rel[150 + i]; }
   }
   if (IMPURITY && MCAtom[IMTYPE].molecule == 1) {          // mc_main.cc:718-724, 453-454
      SaveDensities1D(prefix, acount);
      SaveDensities2D(prefix, acount, MC_BLOCK);
      SaveDensities2D(prefix, acount, MC_TOTAL);
   }
   if (IMPURITY && MCAtom[IMTYPE].molecule == 2) {          // mc_main.cc:726-737, 456-461
      SaveDensities1D(prefix, acount);
      SaveRho1D(prefix, acount, MC_BLOCK);
      SaveRhoThetaChi(prefix, acount, MC_BLOCK);
      SaveDensities3D(prefix, acount, MC_TOTAL);
      SaveRho1D(prefix, acount, MC_TOTAL);
   }
}
// The reference's own per-block writers (SaveEnergy, SaveSumEnergy, SaveRCF, SaveGraSum, SaveExchangeLength,
// SaveAreaEstimators, SaveAreaEstim3D) on test-supplied accumulator values.
extern "C++" {
void SaveEnergy(const char [], double, long int);
void SaveSumEnergy(double, double);
extern double **_rcf_sum;
extern double _kin_total, _pot_total, _rot_total, _rotsq_total, _Cv_total, _Cv_trans_total, _Cv_rot_total;
extern double _areas3DMFF[6], _inert3DMFF[9], _areas3DSFF[6], _inert3DSFF[9];
extern std::fstream _feng;
}
void ref_save_block(const char *prefix, long block, double acount, const double *scal7, const double *rcf0, const double *rcf19,
                    const double *gr1d, const double *ploops, const int *pindex, const double *area40)
{
   avergCount = acount;
   _bkin = scal7[0]; _bpot = scal7[1]; _brot = scal7[2]; _brotsq = scal7[3]; _bCv = scal7[4]; _bCv_trans = scal7[5]; _bCv_rot = scal7[6];
   _kin_total = scal7[0]; _pot_total = scal7[1]; _rot_total = scal7[2]; _rotsq_total = scal7[3]; _Cv_total = scal7[4]; _Cv_trans_total = scal7[5]; _Cv_rot_total = scal7[6];
   SaveEnergy(prefix, acount, block);
   std::string fs = std::string(prefix) + "_sum.eng";
   if (_feng.is_open()) _feng.close();
   _feng.clear();
   _feng.open(fs.c_str(), std::ios::out); io_setout(_feng);
   SaveSumEnergy(acount, 3.0);
   _feng.close();
   if (ROTATION) {
      for (int it = 0; it < NumbRotTimes; it++) {
         _rcf[0][it] = rcf0[it]; _rcf_sum[0][it] = rcf0[it];
         for (int ip = 1; ip < NUMB_RCF; ip++) { _rcf[ip][it] = rcf19[it]; _rcf_sum[ip][it] = rcf19[it]; }
      }
      SaveRCF(prefix, acount, MC_BLOCK);
      SaveRCF(prefix, acount, MC_TOTAL);
   }
   for (int id = 0; id < NumbTypes; id++)
      if (!MCAtom[id].molecule) for (int i = 0; i < 300; i++) _gr1D_sum[id][i] = gr1d[i];
   SaveGraSum(prefix, acount);
   if (BOSONS) {
      for (int a = 0; a < MCAtom[BSTYPE].numb; a++) { _ploops[a] = ploops[a]; PIndex[a] = pindex[a]; }
      PrintXYZprl = 0;
      SaveExchangeLength(prefix, acount, block);
      for (int k = 0; k < 2; k++) { _areas[k] = area40[k]; _area2[k] = area40[2 + k]; _inert[k] = area40[4 + k]; }
      for (int k = 0; k < 6; k++) { _areas3DSFF[k] = area40[6 + k]; _areas3DMFF[k] = area40[21 + k]; }
      for (int k = 0; k < 9; k++) { _inert3DSFF[k] = area40[12 + k]; _inert3DMFF[k] = area40[27 + k]; }
      if (MCAtom[IMTYPE].molecule == 1) SaveAreaEstimators(prefix, acount, block);
      SaveAreaEstim3D(prefix, acount, block, 0);
      if (MCAtom[IMTYPE].molecule == 2 && ISPHER == 0) SaveAreaEstim3D(prefix, acount, block, 1);
   }
}
// IOxyz / IOxyzAng of the current reference state (set with ref_set_state)
void ref_write_xyz(const char *path_xyz, const char *prefix_ang)
{
   IOxyz(IOWrite, path_xyz);
   IOxyzAng(IOWrite, prefix_ang);
}
void ref_GetAreaEstimators(double *areas, double *area2, double *inert)
{
   GetAreaEstimators();
   for (int i = 0; i < 2; i++) { areas[i] = _areas[i]; area2[i] = _area2[i]; inert[i] = _inert[i]; }
}
void ref_GetAreaEstim3D(int iframe, double *areas6, double *inert9)
{
   GetAreaEstim3D(iframe);
   for (int i = 0; i < 6; i++) areas6[i] = iframe ? _areas3DMFF[i] : _areas3DSFF[i];
   for (int i = 0; i < 9; i++) inert9[i] = iframe ? _inert3DMFF[i] : _inert3DSFF[i];
}
// instantaneous values: the block accumulators are zeroed first (mc_estim.cc:2232-2249, 2565-2590)
void ref_area_estimators(double *out4)
{
   for (int i = 0; i < 2; i++) { _areas[i] = 0; _area2[i] = 0; _inert[i] = 0; }
   GetAreaEstimators();
   out4[0] = _areas[0]; out4[1] = _areas[1]; out4[2] = _inert[0] * (double)NumbTimes; out4[3] = _inert[1] * (double)NumbTimes;
}
void ref_area_estim3d(int iframe, double *areas6, double *inert9)
{
   for (int i = 0; i < 6; i++) { _areas3DMFF[i] = 0; _areas3DSFF[i] = 0; }
   for (int i = 0; i < 9; i++) { _inert3DMFF[i] = 0; _inert3DSFF[i] = 0; }
   GetAreaEstim3D(iframe);
   for (int i = 0; i < 6; i++) areas6[i] = iframe ? _areas3DMFF[i] : _areas3DSFF[i];
   for (int i = 0; i < 9; i++) inert9[i] = (iframe ? _inert3DMFF[i] : _inert3DSFF[i]) * (double)NumbTimes;
}
void ref_exchange_length(double *ploops)
{
   int nb = MCAtom[BSTYPE].numb;
   for (int i = 0; i < nb; i++) _ploops[i] = 0;
   GetExchangeLength();
   for (int i = 0; i < nb; i++) ploops[i] = _ploops[i];
}
void ref_reflect(int plane)
{
   PrintYrfl = 0; PrintXrfl = 0; PrintZrfl = 0;
   if (plane == 0) Reflect_MF_XZ(); else if (plane == 1) Reflect_MF_YZ(); else Reflect_MF_XY();
}
void ref_rotsym(void) { RotSymConfig(); }       // rotor pick from rnd1 (queue 1)
void ref_set_nfold(int n) { NFOLD_ROT = n; }
// worm moves (mc_qworm.cc; the move functions have external linkage but no header declaration)
void ref_worm_set(const int *st5) { Worm.exists = st5[0]; Worm.ira = st5[1]; Worm.masha = st5[2]; Worm.atom_i = st5[3]; Worm.atom_m = st5[4]; }
void ref_worm_get(int *st5) { st5[0] = Worm.exists; st5[1] = Worm.ira; st5[2] = Worm.masha; st5[3] = Worm.atom_i; st5[4] = Worm.atom_m; }
void ref_worm_op(int which)
{
   switch (which) {
      case 0: qworm_open(); break;
      case 1: qworm_close(); break;
      case 4: qworm_advance(); break;
      case 5: qworm_recede(); break;
      case 6: qworm_swap(); break;
      default: MCWormMove();
   }
}
void ref_worm_counters(double *t7, double *a7, double *cq)
{
   for (int i = 0; i < QWMAXMOVES; i++) { t7[i] = QWTotal[0][i]; a7[i] = QWAccep[0][i]; }
   *cq = countQW;
}
void ref_get_perm(int *pi, int *ri, int n) { for (int i = 0; i < n; i++) { pi[i] = PIndex[i]; ri[i] = RIndex[i]; } }
int ref_world_line(int atom, int pt) { return WorldLine(atom, pt) ? 1 : 0; }
int ref_worm_enabled(void) { return WORM ? 1 : 0; }
void ref_MCGetAverage(double *out7)
{
   MCGetAverage();
   out7[0] = _bkin; out7[1] = _bpot; out7[2] = _brot; out7[3] = _brotsq;
   out7[4] = _bCv; out7[5] = _bCv_trans; out7[6] = _bCv_rot;
}

// MRG32k3a: first `ndraw` uniforms of the first `nstream` RngStream objects
// constructed after SetPackageSeed(seed) (rngstream.cc:303-321)
void ref_rngstream_draws(const unsigned long *seed6, int nstream, int ndraw, double *out)
{
   RngStream::SetPackageSeed(seed6);
   for (int s = 0; s < nstream; s++) {
      RngStream g;
      for (int k = 0; k < ndraw; k++) out[(size_t)s * ndraw + k] = g.RandU01();
   }
}

// The hot loop of mc_main.cc:348-381 (non-worm branch) for `nsteps` values of
// `time` starting at time0; returns wall seconds (omp_get_wtime).
double ref_run_steps(int time0, int nsteps)
{
   double t0 = omp_get_wtime();
   for (int s = 0; s < nsteps; s++) {
      int time = (time0 + s) % NumbTimes;
      for (int type = 0; type < NumbTypes; type++) PIMCPass(type, time);
   }
   return omp_get_wtime() - t0;
}

// The hot loop of mc_main.cc:348-381 including the worm branch (:355-379): MCWormMove, then -- in the Z sector only --
// one bisection move at a random slice and the whole-path move at time 0.
double ref_run_steps_worm(int time0, int nsteps)
{
   double t0 = omp_get_wtime();
   for (int s = 0; s < nsteps; s++) {
      int time = (time0 + s) % NumbTimes;
      for (int type = 0; type < NumbTypes; type++)
         if (WORM && (type == Worm.type)) {
            MCWormMove();
            if (!Worm.exists) {
               int rt = nrnd2(NumbTimes);
               if ((type == BSTYPE) || (type == FERMTYPE)) MCBisectionMoveExchange(type, rt);
               else MCBisectionMove(type, rt);
               if (time == 0) {
                  if ((type == BSTYPE) || (type == FERMTYPE)) MCMolecularMoveExchange(type);
                  else MCMolecularMove(type);
               }
            }
         } else PIMCPass(type, time);
   }
   return omp_get_wtime() - t0;
}
int ref_worm_exists(void) { return Worm.exists; }
// block accumulators of the exchange / area estimators as MCGetAverage left them
void ref_get_exchange_acc(double *ploops, double *sff_area6, double *sff_inert9)
{
   for (int i = 0; i < MCAtom[BSTYPE].numb; i++) ploops[i] = _ploops[i];
   for (int i = 0; i < 6; i++) sff_area6[i] = _areas3DSFF[i];
   for (int i = 0; i < 9; i++) sff_inert9[i] = _inert3DSFF[i];
}
// every block accumulator the observables of north_star need, as MCGetAverage left them: _rcf[0][0..Q-1],
// linear-dopant area sums {_areas[2], _area2[2], _inert[2]} and the 3-D tensors {areas6, inert9} in both frames
void ref_get_block_acc(double *rcf0, double *lin6, double *sff15, double *mff15)
{
   if (rcf0) for (int i = 0; i < NumbRotTimes; i++) rcf0[i] = _rcf[0][i];
   if (lin6) for (int i = 0; i < 2; i++) { lin6[i] = _areas[i]; lin6[2 + i] = _area2[i]; lin6[4 + i] = _inert[i]; }
   if (sff15) { for (int i = 0; i < 6; i++) sff15[i] = _areas3DSFF[i]; for (int i = 0; i < 9; i++) sff15[6 + i] = _inert3DSFF[i]; }
   if (mff15) { for (int i = 0; i < 6; i++) mff15[i] = _areas3DMFF[i]; for (int i = 0; i < 9; i++) mff15[6 + i] = _inert3DMFF[i]; }
}
void ref_reset_area_acc(void)
{
   for (int i = 0; i < 2; i++) { _areas[i] = 0.0; _area2[i] = 0.0; _inert[i] = 0.0; }
   for (int i = 0; i < 6; i++) { _areas3DSFF[i] = 0.0; _areas3DMFF[i] = 0.0; }
   for (int i = 0; i < 9; i++) { _inert3DSFF[i] = 0.0; _inert3DMFF[i] = 0.0; }
}
void ref_reset_exchange_acc(void)
{
   for (int i = 0; i < MCAtom[BSTYPE].numb; i++) _ploops[i] = 0.0;
   for (int i = 0; i < 6; i++) { _areas3DSFF[i] = 0.0; _areas3DMFF[i] = 0.0; }
   for (int i = 0; i < 9; i++) { _inert3DSFF[i] = 0.0; _inert3DMFF[i] = 0.0; }
}

} // extern "C"
