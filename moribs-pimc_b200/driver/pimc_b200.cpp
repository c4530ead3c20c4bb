// pimc_b200 -- C++ host driver of the B200 PIMC hot path: a drop-in for the reference's `pimc` binary on the
// sampling path.  It reads the same `qmc.input` from the current directory (keywords and semantics of
// mc_input.cc:18-57,115-343), the same table files (1-D/2-D/3-D potentials, <type>_T<T>t<Q>.rot or
// .rho/.eng/.esq; formats of mc_poten.cc:254-546 and README.md:78-149), `xyz.init` (initconf.f) when
// READMCCOORDS is set, runs the block loop of mc_main.cc:340-484 with the moves and estimator sums on the GPU
// through include/pimcgpu.h, and writes <prefix>.eng, <prefix>_sum.eng, <prefix>NNN.rcf, <prefix>_sum.rcf,
// <prefix>_sum.gra, <prefix>.prl / .sup / .sffs3d / .mffs3d (exchange lengths and superfluid area estimators),
// <prefix>.xyz and the yw001.stat/.conf/.tabl checkpoints in the reference's formats
// (mc_main.cc:704-836, mc_estim.cc:1141-1191,1288-1326,2021-2085,2596-2729, mc_input.cc:517-794).
//
// Independent Markov chains: `--chains C` chains per GPU (default 1); with `--ranks N --rank r` (or the
// RANK/WORLD_SIZE environment of a launcher) every rank drives one GPU, the block accumulators are all-reduced
// with NCCL over NVLink before rank 0 writes the block files.
//
// ROTDENSI 1 (rattle-and-shake propagator), REFLECTX/Y/Z, ROTSYM and WORM (exchange sampling with the worm algorithm,
// mc_qworm.cc) are honoured on the device; yw001.worm is written in the reference's layout for chain 0.  yw001.rand (SPRNG state) has no
// counterpart: the MRG32k3a package seed and step counter are written to yw001.mrg instead.
#include "../../include/pimcgpu.h"
#include "../data/vspher_table.h"
#include "pimc_writers.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <chrono>
#include <thread>
#include <algorithm>
#include <array>
#include <iterator>
#include <sys/stat.h>

#ifdef PIMC_WITH_NCCL
#include <nccl.h>
#include <cuda_runtime.h>
#endif

using namespace std;

static const int IO_WIDTH = 14, IO_WIDTH_BLOCK = 4, IO_PRECISION = 6;
static const char BLANK[] = "   ";

[[noreturn]] static void die(const string &proc, const string &msg)
{
   // nrerror, mc_utils.cc:99-108: message on stdout, exit status 1
   cout << endl << "run-time error..." << endl << proc << ":  " << msg << endl << "...now exiting to system..." << endl << endl;
   exit(1);
}
static void ck(int rc, const char *what)
{
   if (rc) die(what, pimcgpu_last_error());
}

struct Species { string name, fpot; int numb, molecule, stat, levels; double mcstep, rtstep, mass; };

struct Deck {
   vector<Species> types;
   string outdir = "./", prefix = "pimc";
   int P = 0, Q = 0, ispher = 0, minimage = 0, rotden_type = 0, read_coords = 0, worm = 0;
   int rot_odevn = 0, rnratio = 1, refl[3] = {0, 0, 0}, rotsym = 0, nfold = 1, worm_m = 0, restart = 0;
   string worm_type;
   double worm_c = 0;
   double rot_eoff = 0, xrot = 0, yrot = 0, zrot = 0;
   double temperature = 0, density = 0.02;
   long passes = 1, blocks = 1, eq_blocks = 0;
   int skip_ratio = 100000, skip_total = 10000, skip_averg = 1;
   int N() const { int n = 0; for (auto &t : types) n += t.numb; return n; }
};

static double species_mass(const string &s)
{
   // MCInitParams, mc_setup.cc:244-319 with the masses of mc_const.h:50-56
   const double H1 = 1.0078, H2 = 2.015650642, HE4 = 4.0026032497, C12 = 12.0, N14 = 14.003, O16 = 15.994915, S32 = 31.972;
   if (s == "He4") return HE4;
   if (s == "H2") return H2;
   if (s == "OCS") return O16 + C12 + S32;
   if (s == "N2O") return 2.0 * N14 + O16;
   if (s == "CO2") return C12 + 2.0 * O16;
   if (s == "CO") return C12 + O16;
   if (s == "HCN") return H1 + C12 + N14;
   if (s == "HCCCN") return H1 + 3.0 * C12 + N14;
   if (s == "H2O") return 2.0 * H1 + O16;
   if (s == "SO2") return 2.0 * O16 + S32;
   if (s == "HCOOCH3") return 4.0 * H1 + 2.0 * O16 + 2.0 * C12;
   die("MCInitParams", "Unknown atom/molecule type");
}

static Deck read_deck(const char *path)
{
   ifstream inf(path);
   if (!inf.good()) die("IOReadParams", string("Can't open input file  [") + path + "]");
   Deck d;
   string key, rot_type;
   double rot_step = 0;
   bool impurity = false;
   while (inf >> key) {
      if (key == "OUTPUTDIR") inf >> d.outdir;
      else if (key == "FILENAMEPREFIX") inf >> d.prefix;
      else if (key == "TEMPERATURE") inf >> d.temperature;
      else if (key == "DENSITY") inf >> d.density;
      else if (key == "ATOM" || key == "MOLECULE" || key == "NONLINEAR") {
         Species s;
         string sstat, smod;
         inf >> s.name >> s.numb >> sstat >> s.mcstep >> s.levels >> s.fpot >> smod;
         if (s.numb < 0) { d.ispher = 1; s.numb = -s.numb; }
         if (sstat == "BOSE") s.stat = 1; else if (sstat == "BOLTZMANN") s.stat = 0; else die("IOReadParams", "Unknown statistics");
         if (smod != "PRIMITIVE" && smod != "EFFECTIVE") die("IOReadParams", "Unknown model of interaction");
         if (smod == "EFFECTIVE") s.fpot += ".eff"; else s.fpot += ".pot";
         s.molecule = key == "ATOM" ? 0 : (key == "MOLECULE" ? 1 : 2);
         if (s.molecule == 0 && impurity) die("IOReadParams", "Molecules should follow atoms in input file");
         if (s.molecule) impurity = true;
         s.rtstep = 0;
         if (s.numb > 0) { s.mass = species_mass(s.name); d.types.push_back(s); }
      }
      else if (key == "NUMBEROFSLICES") inf >> d.P;
      else if (key == "NUMBEROFPASSES") inf >> d.passes;
      else if (key == "NUMBEROFBLOCKS") inf >> d.blocks >> d.eq_blocks;
      else if (key == "ROTATION") inf >> rot_type >> rot_step >> d.Q;
      else if (key == "ROTDENSI") inf >> d.rotden_type >> d.rot_odevn >> d.rot_eoff >> d.xrot >> d.yrot >> d.zrot >> d.rnratio;   // mc_input.cc:274-284
      else if (key == "REFLECTX") inf >> d.refl[0];                                                                               // mc_input.cc:296-330
      else if (key == "REFLECTY") inf >> d.refl[1];
      else if (key == "REFLECTZ") inf >> d.refl[2];
      else if (key == "ROTSYM") { d.rotsym = 1; inf >> d.nfold; }
      else if (key == "WORM") { d.worm = 1; inf >> d.worm_type >> d.worm_c >> d.worm_m; }       // mc_input.cc:286-293
      else if (key == "MINIMAGE") d.minimage = 1;
      else if (key == "RESTART") d.restart = 1;                                               // mc_input.cc:117-120
      else if (key == "READMCCOORDS") d.read_coords = 1;
      else if (key == "MCSKIP_RATIO") inf >> d.skip_ratio;
      else if (key == "MCSKIP_TOTAL") inf >> d.skip_total;
      else if (key == "MCSKIP_AVERG") inf >> d.skip_averg;
      string rest;
      getline(inf, rest);                 // the rest of the line is a comment
   }
   if (d.Q) {
      bool found = false;
      for (auto &t : d.types) if (t.name == rot_type) { if (!t.molecule) die("IOReadParams", "Rotational degrees of freedom for molecules only"); t.rtstep = rot_step; found = true; }
      if (!found) die("IOReadParams", "Can't find a particle type to sample rotational degrees of freedom");
   }
   if (d.types.empty() || d.types.size() > 2) die("IOReadParams", "No more then one atom/molecule type: densities and potential energy");
   return d;
}

// read_datafile, mc_poten.cc:757-865: whitespace-separated columns, lines starting with '#' skipped
static vector<vector<double>> read_columns(const string &path, int ncol)
{
   ifstream f(path);
   if (!f.good()) die("read_datafile", "Can't open input file  [" + path + "]");
   vector<vector<double>> c(ncol);
   string line;
   while (getline(f, line)) {
      istringstream is(line);
      string tok;
      if (!(is >> tok) || tok == "#") continue;
      c[0].push_back(strtod(tok.c_str(), nullptr));
      for (int k = 1; k < ncol; k++) { is >> tok; c[k].push_back(strtod(tok.c_str(), nullptr)); }
   }
   return c;
}
// one number per token, fast path for the 23.6 M-line density-matrix files (init_rot3D, mc_poten.cc:462-499)
static vector<double> read_numbers(const string &path, size_t n)
{
   FILE *f = fopen(path.c_str(), "rb");
   if (!f) die("init_rot3D", "Can't open input file  [" + path + "]");
   fseek(f, 0, SEEK_END);
   long sz = ftell(f);
   fseek(f, 0, SEEK_SET);
   string buf(sz, '\0');
   if (fread(&buf[0], 1, sz, f) != (size_t)sz) die("init_rot3D", "short read of " + path);
   fclose(f);
   vector<double> v;
   v.reserve(n);
   const char *p = buf.c_str(), *end = p + sz;
   while (p < end && v.size() < n) {
      char *q;
      double x = strtod(p, &q);
      if (q == p) { p++; continue; }
      v.push_back(x);
      p = q;
   }
   if (v.size() != n) die("init_rot3D", "wrong number of entries in " + path);
   return v;
}

// values as a run would read them back from the generated files (E15.8 / 1P,E15.8 text)
static string fortran_1p(double v) { char b[16]; pimcgpu_format_e15_8(v, 1, b); return b; }
static void round_e15_8(vector<double> &v)
{
#pragma omp parallel for schedule(static)
   for (long i = 0; i < (long)v.size(); i++) { char b[32]; snprintf(b, sizeof b, "%.7E", v[i]); v[i] = strtod(b, nullptr); }
}

static string cxx_double(double x) { ostringstream o; o << x; return o.str(); }     // how init_rot3D spells the temperature

struct Writer {
   static void setout(ostream &o) { o << setprecision(IO_PRECISION) << setiosflags(ios::scientific); }
};

int main(int argc, char **argv)
{
   int chains = 1, rank = 0, ranks = 1;
   unsigned long seed[6] = {12345, 12345, 12345, 12345, 12345, 12345};          // fixedseed(), omprng.cc:14-18
   if (getenv("RANK")) rank = atoi(getenv("RANK"));
   if (getenv("WORLD_SIZE")) ranks = atoi(getenv("WORLD_SIZE"));
   for (int i = 1; i < argc; i++) {
      string a = argv[i];
      if (a == "--chains" && i + 1 < argc) chains = atoi(argv[++i]);
      else if (a == "--rank" && i + 1 < argc) rank = atoi(argv[++i]);
      else if (a == "--ranks" && i + 1 < argc) ranks = atoi(argv[++i]);
      else if (a == "--seed" && i + 1 < argc) { unsigned long s = strtoul(argv[++i], nullptr, 10); for (int k = 0; k < 6; k++) seed[k] = s + k; }
      else die("main", "usage: pimc_b200 [--chains C] [--ranks N --rank r] [--seed s]   (reads ./qmc.input)");
   }
   Deck d = read_deck("qmc.input");
   const int N = d.N(), P = d.P, Q = d.Q;
   const size_t n = (size_t)N * P;

   pimcgpu_system sys;
   memset(&sys, 0, sizeof sys);
   sys.ntypes = (int)d.types.size();
   int imtype = -1, bstype = -1, first[3] = {0, 0, 0};
   for (int t = 0; t < sys.ntypes; t++) {
      const Species &s = d.types[t];
      sys.type[t].numb = s.numb; sys.type[t].molecule = s.molecule; sys.type[t].stat = s.stat; sys.type[t].levels = s.levels;
      sys.type[t].mass = s.mass; sys.type[t].mcstep = s.mcstep; sys.type[t].rtstep = s.rtstep;
      first[t + 1] = first[t] + s.numb;
      if (s.molecule) imtype = t;
      if (s.stat == 1) bstype = t;
   }
   sys.P = P; sys.Q = Q; sys.temperature = d.temperature; sys.ispher = d.ispher; sys.minimage = d.minimage;
   for (int k = 0; k < 3; k++) sys.box[k] = pow((double)N / d.density, 1.0 / 3.0);        // MCInit, mc_setup.cc:339,359-360
   sys.rotden_type = d.rotden_type; sys.rot_odevn = d.rot_odevn; sys.rot_eoff = d.rot_eoff;
   sys.x_rot = d.xrot; sys.y_rot = d.yrot; sys.z_rot = d.zrot; sys.rnratio = d.rnratio;
   for (int k = 0; k < 3; k++) sys.reflect[k] = d.refl[k];
   sys.rotsym = d.rotsym; sys.nfold_rot = d.nfold;
   if (d.worm) {
      sys.worm = 1; sys.worm_type = -1; sys.worm_c = d.worm_c; sys.worm_m = d.worm_m;
      for (int t = 0; t < sys.ntypes; t++) if (d.types[t].name == d.worm_type) sys.worm_type = t;
      if (sys.worm_type < 0) die("MCWormInit", "Can't find a particle type for the worm algorithm");
      if (d.worm_m >= P) die("IOReadParams", "Worm algorithm: m should be smaller then M");
   }
   sys.nchains = chains; sys.chain_offset = (long)rank * chains; sys.device = rank;
#ifdef PIMC_WITH_NCCL
   { int nd = 1; cudaGetDeviceCount(&nd); sys.device = rank % std::max(1, nd); }
#else
   sys.device = 0;
#endif

   // mc_main.cc:129-144: never run over the files of an earlier simulation (permutation.tab always; the checkpoint files unless
   // this is a RESTART).  _io_error prints "<proc>: <message> <file>" through nrerror.
   auto file_exists = [](const char *f) { struct stat sb; return stat(f, &sb) == 0; };
   if (rank == 0) {
      if (file_exists("permutation.tab")) die("QMC ->", "File already exists: permutation.tab");
      if (!d.restart)
         for (const char *f : {"yw001.stat", "yw001.conf", "yw001.rand"})
            if (file_exists(f)) die("QMC ->", string("File already exists: ") + f);
   }
   // ---- tables: InitPotentials / InitRotDensity, mc_poten.cc:93-164 ----
   pimcgpu_tables tab;
   memset(&tab, 0, sizeof tab);
   vector<vector<double>> t1d, trot;
   vector<double> rg2, cg2, v2, v3, rho, erot, esq;
   for (int t = 0; t < sys.ntypes; t++) {
      const Species &s = d.types[t];
      if (s.molecule == 0) {
         t1d = read_columns(s.fpot, 2);
         tab.n1d = (int)t1d[0].size(); tab.grid1d = t1d[0].data(); tab.pot1d = t1d[1].data();
      } else if (s.molecule == 1) {
         ifstream f(s.fpot);
         if (!f.good()) die("init_pot2D", "Can't open input file  [" + s.fpot + "]");
         f >> tab.rsize2d >> tab.csize2d >> tab.dr2d >> tab.dc2d;
         rg2.resize(tab.rsize2d); cg2.resize(tab.csize2d); v2.resize((size_t)tab.rsize2d * tab.csize2d);
         for (auto &x : rg2) f >> x;
         for (auto &x : cg2) f >> x;
         for (auto &x : v2) f >> x;
         tab.rgrid2d = rg2.data(); tab.cgrid2d = cg2.data(); tab.pot2d = v2.data();
      } else if (sys.ntypes > 1) {
         ifstream f(s.fpot);
         if (!f.good()) die("init_pot3D", "Can't open input file  [" + s.fpot + "]");
         f >> tab.rgrd >> tab.thgrd >> tab.chgrd >> tab.rvmin >> tab.rvmax;
         v3.resize((size_t)tab.rgrd * tab.thgrd * tab.chgrd);
         for (auto &x : v3) f >> x;
         tab.vtable = v3.data();
         cout << "Rgrd=" << tab.rgrd << " THgrd=" << tab.thgrd << " CHgrd=" << tab.chgrd << " Rvmin=" << tab.rvmin << " Rvmax=" << tab.rvmax << endl;
      }
   }
   if (d.ispher) tab.vspher = PIMC_VSPHER_TABLE;      // negative species count: spherical H2O-pH2 treatment, DATA table of vspher.f:15-517
   if (Q > 0 && imtype >= 0) {
      const Species &s = d.types[imtype];
      string base = s.name + "_T" + cxx_double(d.temperature) + "t" + to_string(Q);          // mc_poten.cc:443-458,518-524
      if (d.rotden_type == 1) {
         // InitRotDensity loads nothing for the rattle-and-shake propagator (mc_poten.cc:148-164)
      } else if (s.molecule == 1) {
         if (!ifstream(base + ".rot").good() && d.xrot > 0.0) {
            // table absent: what linden.x T Q B npt iodevn writes (linear_prop/README uses 1500 points), generated on the
            // device.  rho falls by 1/e over 2 B tau in cos(gamma) (linden.f:129): at least ten grid points per decay length
            const double btau = d.xrot * 1.4387752224 / (d.temperature * Q);
            const int npt = (int)std::min(20001.0, std::max(1500.0, ceil(10.0 / btau) + 1.0));
            cout << "generating " << base << ".rot on the device (linden: B=" << d.xrot << " cm-1, " << npt << " points)" << endl;
            vector<double> t4((size_t)npt * 4);
            ck(pimcgpu_gen_linden(d.temperature, Q, d.xrot, npt, d.rot_odevn, t4.data(), nullptr), "pimcgpu_gen_linden");
            if (rank == 0) {      // temporary name + rename: a rank that starts later sees either no file or the whole file
               ck(pimcgpu_write_rot((base + ".rot.tmp").c_str(), t4.data(), npt), "pimcgpu_write_rot");
               if (rename((base + ".rot.tmp").c_str(), (base + ".rot").c_str()) != 0) die("init_rotdens", "cannot rename " + base + ".rot.tmp");
            }
            trot.assign(4, vector<double>(npt));
            for (int i = 0; i < npt; i++) for (int k = 0; k < 4; k++) trot[k][i] = strtod(fortran_1p(t4[4 * i + k]).c_str(), nullptr);
         } else
            trot = read_columns(base + ".rot", 4);
         tab.nrot = (int)trot[0].size(); tab.rotgrid = trot[0].data(); tab.rotdens = trot[1].data(); tab.rotderv = trot[2].data(); tab.rotesqr = trot[3].data();
      } else {
         cout << base << ".rho " << base << ".eng " << base << ".esq" << endl;
         if (!ifstream(base + ".rho").good() && d.xrot > 0.0 && d.yrot > 0.0 && d.zrot > 0.0) {
            // tables absent: what 181 asymrho.x jobs + compile.x produce (nmv_prop/README), generated on the device from the
            // ROTDENSI constants.  rotmat (asymrho.f:792-813) quantises along z: Arot = X_Rot, Brot = Z_Rot, Crot = Y_Rot.
            // maxj: first j whose lowest level falls under the generator's own cut (2j+1)/8pi^2 e^{-tau E} < 1e-16 (:520)
            const double tau = 1.0 / (0.6950356 * d.temperature) / Q, cmin = std::min(d.xrot, std::min(d.yrot, d.zrot));
            int maxj = 4;
            while (maxj < 876 && (2 * maxj + 1) / (8.0 * M_PI * M_PI) * exp(-tau * cmin * maxj * (maxj + 1.0)) >= 1e-16) maxj++;
            cout << "generating the tables on the device (asymrho: A=" << d.xrot << " B=" << d.zrot << " C=" << d.yrot << " cm-1, maxj=" << maxj << ")" << endl;
            rho.resize(PIMCGPU_SIZE_ROTDEN); erot.resize(PIMCGPU_SIZE_ROTDEN); esq.resize(PIMCGPU_SIZE_ROTDEN);
            ck(pimcgpu_gen_asymrho(d.temperature, Q, d.rot_odevn, 0, 180, d.xrot, d.zrot, d.yrot, maxj, rho.data(), erot.data(), esq.data(), nullptr), "pimcgpu_gen_asymrho");
            // the run uses the values a later run would read back from the files: 8 significant digits (E15.8)
            round_e15_8(rho); round_e15_8(erot); round_e15_8(esq);
            if (rank == 0) {      // .rho last and by rename: its presence is what the other ranks (and later runs) test
               ck(pimcgpu_write_e15_8((base + ".eng").c_str(), erot.data(), PIMCGPU_SIZE_ROTDEN, 0), "pimcgpu_write_e15_8");
               ck(pimcgpu_write_e15_8((base + ".esq").c_str(), esq.data(), PIMCGPU_SIZE_ROTDEN, 0), "pimcgpu_write_e15_8");
               ck(pimcgpu_write_e15_8((base + ".rho.tmp").c_str(), rho.data(), PIMCGPU_SIZE_ROTDEN, 0), "pimcgpu_write_e15_8");
               if (rename((base + ".rho.tmp").c_str(), (base + ".rho").c_str()) != 0) die("init_rot3D", "cannot rename " + base + ".rho.tmp");
            }
         } else {
         rho = read_numbers(base + ".rho", PIMCGPU_SIZE_ROTDEN);
         erot = read_numbers(base + ".eng", PIMCGPU_SIZE_ROTDEN);
         esq = read_numbers(base + ".esq", PIMCGPU_SIZE_ROTDEN);
         }
         tab.rho3d = rho.data(); tab.erot3d = erot.data(); tab.esq3d = esq.data();
      }
   }
   ck(pimcgpu_init(&sys, &tab), "pimcgpu_init");

   // ---- initial configuration ----
   vector<double> coords(3 * n, 0.0), angles(3 * n, 0.0), cosine(3 * n, 0.0);
   vector<int> pindex(N), rindex(N);
   for (int a = 0; a < N; a++) pindex[a] = rindex[a] = a;
   for (size_t i = 0; i < n; i++) angles[n + i] = 1.0;                                       // MCConfigInit, mc_setup.cc:471-487
   if (d.read_coords) {
      // initconf.f:1-27 + mc_main.cc:184-201
      ifstream f("xyz.init");
      if (!f.good()) die("initconf", "Can't open input file  [xyz.init]");
      long ntot; f >> ntot;
      int nb = bstype >= 0 ? d.types[bstype].numb : 0;
      if (nb > N) die("initconf", "more bosons than particles");
      vector<char> seen_p(max(1, nb), 0);
      for (int i = 0; i < nb; i++) {
         f >> pindex[i];
         if (!f || pindex[i] < 0 || pindex[i] >= nb || seen_p[pindex[i]]) die("initconf", "the permutation on the first line of xyz.init is not a bijection of 0.." + to_string(nb - 1));
         seen_p[pindex[i]] = 1;
         rindex[pindex[i]] = i;
      }
      string line; getline(f, line); getline(f, line);
      for (size_t i = 0; i < n; i++) {
         string label; f >> label;
         for (int k = 0; k < 3; k++) f >> coords[k * n + i] >> angles[k * n + i];
      }
      if (!f) die("initconf", "short xyz.init");
   } else {
      // classical start: molecules in a row through the origin, atoms on the nearest sites of a simple cubic lattice
      // (the reference's initLattice_config draws a different lattice; only the equilibrated ensemble matters)
      vector<array<double, 4>> sites;
      const double a0 = 3.8;
      for (int i = -8; i <= 8; i++) for (int j = -8; j <= 8; j++) for (int k = -8; k <= 8; k++) {
         double x = i * a0, y = j * a0, z = k * a0, r = sqrt(x * x + y * y + z * z);
         if (r > 3.3) sites.push_back({r, x, y, z});
      }
      sort(sites.begin(), sites.end());
      int isite = 0, atom = 0;
      for (int t = 0; t < sys.ntypes; t++)
         for (int k = 0; k < d.types[t].numb; k++, atom++) {
            double c[3] = {0, 0, 0};
            if (d.types[t].molecule == 0) { c[0] = sites[isite][1]; c[1] = sites[isite][2]; c[2] = sites[isite][3]; isite++; }
            else c[0] = 2.9 * (k - 0.5 * (d.types[t].numb - 1));
            for (int it = 0; it < P; it++) for (int dd = 0; dd < 3; dd++) coords[dd * n + (size_t)atom * P + it] = c[dd];
         }
   }
   ck(pimcgpu_upload_state(-1, coords.data(), angles.data(), bstype >= 0 ? pindex.data() : nullptr), "pimcgpu_upload_state");
   ck(pimcgpu_seed(seed), "pimcgpu_seed");
   // RESTART (mc_main.cc:223-231): block counter from yw001.stat; the full sampler state -- every chain's beads, angles,
   // permutation tables, worm, MRG32k3a streams -- from this rank's side file yw001.b200[.rank] (row N4)
   long start_block = 0;
   const string ckname = ranks > 1 ? "yw001.b200." + to_string(rank) : string("yw001.b200");
   if (d.restart) {
      ifstream fs("yw001.stat");
      string key2;
      if (!fs.good()) die("StatusIO", "Can't open input file  [yw001.stat]");
      while (fs >> key2) { if (key2 == "STARTBLOCK") fs >> start_block; string rest; getline(fs, rest); }
      ifstream fc(ckname, ios::binary);
      if (!fc.good()) die("ConfigIO", "Can't open input file  [" + ckname + "]");
      vector<char> blob((istreambuf_iterator<char>(fc)), istreambuf_iterator<char>());
      ck(pimcgpu_checkpoint_load(blob.data(), (long)blob.size()), "pimcgpu_checkpoint_load");
      if (rank == 0) cout << "RESTART at block " << start_block << ", step " << pimcgpu_step_counter() << endl;
   }

#ifdef PIMC_WITH_NCCL
   ncclComm_t comm = nullptr;
   if (ranks > 1) {
      ncclUniqueId id;
      // Rendezvous through the output directory.  A stale id file of an earlier run must never be taken for this run's:
      // every other rank first drops a hello file, rank 0 waits for all of them, and only then writes a FRESH id file
      // (temporary name + rename); a rank accepts the id file only if it is newer than its own hello.  MASTER_PORT (set by
      // torchrun-style launchers) keeps concurrent runs in one directory apart.  Rank 0 removes the files afterwards.
      const char *nonce = getenv("PIMC_NCCL_NONCE") ? getenv("PIMC_NCCL_NONCE") : (getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0");
      const string idf = d.outdir + ".pimc_nccl_id." + nonce;
      auto mtime_ns = [](const string &f, long long &t) { struct stat sb; if (stat(f.c_str(), &sb) != 0) return false; t = (long long)sb.st_mtim.tv_sec * 1000000000LL + sb.st_mtim.tv_nsec; return true; };
      if (rank == 0) {
         remove(idf.c_str());
         for (int r = 1; r < ranks; r++) {
            long long t;
            const string hf = idf + ".hello." + to_string(r);
            for (int tries = 0; !mtime_ns(hf, t); tries++) { if (tries > 12000) die("nccl", "rank " + to_string(r) + " never arrived"); this_thread::sleep_for(chrono::milliseconds(50)); }
         }
         if (ncclGetUniqueId(&id) != ncclSuccess) die("nccl", "ncclGetUniqueId failed");
         FILE *f = fopen((idf + ".tmp").c_str(), "wb");
         if (!f || fwrite(&id, sizeof id, 1, f) != 1 || fclose(f) != 0) die("nccl", "cannot write " + idf + ".tmp");
         if (rename((idf + ".tmp").c_str(), idf.c_str()) != 0) die("nccl", "cannot rename " + idf + ".tmp");
      } else {
         const string hf = idf + ".hello." + to_string(rank);
         { FILE *f = fopen(hf.c_str(), "wb"); if (!f) die("nccl", "cannot write " + hf); fputc('h', f); fclose(f); }
         long long t_hello = 0, t_id = 0;
         mtime_ns(hf, t_hello);
         for (int tries = 0;; tries++) {
            if (mtime_ns(idf, t_id) && t_id >= t_hello) {
               FILE *f = fopen(idf.c_str(), "rb");
               if (f) { const bool ok = fread(&id, sizeof id, 1, f) == 1; fclose(f); if (ok) break; }
            }
            if (tries > 12000) die("nccl", "no id file from rank 0");
            this_thread::sleep_for(chrono::milliseconds(50));
         }
      }
      if (ncclCommInitRank(&comm, ranks, id, rank) != ncclSuccess) die("nccl", "ncclCommInitRank failed");
      if (rank == 0) { remove(idf.c_str()); for (int r = 1; r < ranks; r++) remove((idf + ".hello." + to_string(r)).c_str()); }     // the collective init has consumed them
   }
#else
   if (ranks > 1) die("main", "built without NCCL: multi-rank runs need -DPIMC_WITH_NCCL");
#endif

   long n_acc = 0, off_gr1d = 0, off_gr2d = 0, off_gr3d = 0, off_rcf = 0, off_rel = 0;
   ck(pimcgpu_accum_layout(&n_acc, nullptr, &off_gr1d, &off_gr2d, &off_gr3d, &off_rcf, &off_rel), "pimcgpu_accum_layout");
   vector<double> acc(n_acc), gr1d_sum(PIMCGPU_BINSR, 0.0), rcf_sum(max(1, Q), 0.0), rcf_rows_sum(max(1, Q), 0.0);
   // _gr2D_sum, _gr3D_sum, _relthe_sum/_relphi_sum/_relchi_sum (mc_estim.cc:40-49): accumulated over the blocks by the host
   vector<double> gr2d_sum((size_t)PIMCGPU_BINSR * PIMCGPU_BINST, 0.0), rel_sum(PIMCGPU_BINST + 2 * PIMCGPU_BINSC, 0.0);
   vector<double> gr3d_sum(off_gr3d >= 0 ? (size_t)PIMCGPU_BINSR * PIMCGPU_BINST * PIMCGPU_BINSC : 0, 0.0);
   DensityWriters dw;
   dw.P = P; dw.Q = Q; dw.ntypes = sys.ntypes; dw.imtype = imtype;
   for (int t = 0; t < sys.ntypes; t++) { dw.numb[t] = d.types[t].numb; dw.molecule[t] = d.types[t].molecule; if (!d.types[t].molecule) dw.atype = t; }
   dw.volume = sys.box[0] * sys.box[1] * sys.box[2];
   const string fname = d.outdir + d.prefix;
   const double beta = 1.0 / d.temperature, rottau = Q ? beta / Q : 0.0;
   double kin_tot = 0, pot_tot = 0, rot_tot = 0, rotsq_tot = 0, cv_tot = 0, cvt_tot = 0, cvr_tot = 0, total_count = 0, sums = 0;
   FILE *fsum = nullptr;
   if (rank == 0) fsum = fopen((fname + "_sum.eng").c_str(), "w");
   auto t_start = chrono::steady_clock::now();
   double bead_updates = 0;

   for (long block = start_block + 1; block <= start_block + d.blocks; block++) {
      ck(pimcgpu_accum_reset(), "pimcgpu_accum_reset");
      const long steps_block = d.passes * P;
      long done = 0;
      bool print_xyz = true;                                                                 // MCResetBlockAverage, mc_main.cc:543
      long sum_row_at = -1;                                                                  // step of this block's last _sum.eng row
      // MCSaveAcceptRatio (mc_main.cc:440-441, 879-943): one line every MCSKIP_RATIO steps of the block with the acceptance ratios
      // accumulated since the block began; worm species get the reference's in-line "open/close [..] .. swap [..] .." columns
      auto accept_ratio_line = [&](long step) {
         pimcgpu_scalars sc;
         pimcgpu_accum_device_ptr();
         ck(pimcgpu_sync(), "pimcgpu_sync");
         ck(pimcgpu_block_scalars(&sc), "pimcgpu_block_scalars");
         if (rank != 0) return;
         const long pass = (step + P - 1) / P;                                                 // passCount of the step (1-based)
         cout << "BLOCK:" << setw(8) << block << BLANK << "PASS:" << setw(8) << pass << BLANK << "STEP:" << setw(8) << step << BLANK;
         for (int t = 0; t < sys.ntypes; t++) {
            if (d.worm && t == sys.worm_type) {
               double qt[7], qa[7], cq = 0;
               ck(pimcgpu_worm_counters(qt, qa, &cq), "pimcgpu_worm_counters");
               cout << setw(8) << "open/close" << " [ " << qt[0] / cq << "-" << qt[1] / cq << " ] " << BLANK << setw(8) << qa[0] / qt[0] << BLANK << setw(8) << qa[1] / qt[1] << BLANK;
               cout << setw(8) << "advance/recede" << " [ " << qt[4] / cq << "-" << qt[5] / cq << " ] " << BLANK << setw(8) << qa[4] / qt[4] << BLANK << setw(8) << qa[5] / qt[5] << BLANK;
               cout << setw(8) << "swap" << " [ " << qt[6] / cq << " ] " << BLANK << setw(8) << qa[6] / qt[6] << BLANK;
            } else
               cout << setw(8) << d.types[t].name << BLANK << setw(8) << sc.mcaccep[t][0] / sc.mctotal[t][0] << BLANK << setw(8) << sc.mcaccep[t][1] / sc.mctotal[t][1] << BLANK;
         }
         if (Q) cout << BLANK << "Rot: " << setw(8) << sc.mcaccep[imtype][2] / sc.mctotal[imtype][2] << BLANK;
         cout << endl;
      };
      while (done < steps_block) {
         long chunk = min<long>(d.skip_averg - done % d.skip_averg, steps_block - done);
         if (block <= d.eq_blocks) chunk = min<long>(steps_block - done, 4L * P);            // no estimators while equilibrating
         if (d.skip_ratio > 0) chunk = min<long>(chunk, d.skip_ratio - done % d.skip_ratio);  // stop where the reference prints its line
         ck(pimcgpu_steps(chunk), "pimcgpu_steps");
         done += chunk;
         if (d.skip_ratio > 0 && done % d.skip_ratio == 0) accept_ratio_line(done);
         if (block > d.eq_blocks && done % d.skip_averg == 0) {
            ck(pimcgpu_measure(), "pimcgpu_measure");
            if (bstype >= 0 && rank == 0 && !getenv("PIMC_NO_PERMUTATION_TAB")) {
               // GetPermutation (mc_estim.cc:2731-2753, called by MCGetAverage when there are bosons): one line "PIndex[0] PIndex[1] ..."
               // per measurement appended to permutation.tab; chain 0's permutation (closed paths only, like the measurement itself)
               int st[5] = {0, 0, 0, 0, 0};
               if (d.worm) ck(pimcgpu_worm_state(0, st), "pimcgpu_worm_state");
               if (!st[0]) {
                  ck(pimcgpu_download_state(0, nullptr, nullptr, nullptr, pindex.data()), "pimcgpu_download_state");
                  static FILE *fperm = nullptr;
                  if (!fperm) fperm = fopen("permutation.tab", "a");
                  if (!fperm) die("GetPermutation", "Can't open input file permutation.tab");
                  ostringstream row;
                  for (int i = 0; i < d.types[bstype].numb; i++) row << pindex[i] << " ";
                  fprintf(fperm, "%s\n", row.str().c_str());
                  if (done >= steps_block) fflush(fperm);
               }
            }
            if (print_xyz && rank == 0) {
               // instantaneous configuration of the block's first measured (closed) path: IOxyzAng into <prefix>NNN.xyz
               // (PrintXYZprl, mc_main.cc:395-427,543); chain 0
               int st[5] = {0, 0, 0, 0, 0};
               if (d.worm) ck(pimcgpu_worm_state(0, st), "pimcgpu_worm_state");
               if (!st[0]) {
                  ck(pimcgpu_download_state(0, coords.data(), angles.data(), nullptr, bstype >= 0 ? pindex.data() : nullptr), "pimcgpu_download_state");
                  vector<string> names; vector<int> numbs;
                  for (auto &t : d.types) { names.push_back(t.name); numbs.push_back(t.numb); }
                  ostringstream bc; bc << setw(3) << setfill('0') << block;
                  XyzWriters::xyz_ang(fname + bc.str(), sys.ntypes, names.data(), numbs.data(), P, coords.data(), angles.data(),
                                      bstype >= 0 ? d.types[bstype].numb : 0, pindex.data());
                  print_xyz = false;
               }
            }
            if (done % d.skip_total == 0) {
               // SaveSumEnergy every MCSKIP_TOTAL steps of a measuring block (mc_main.cc:431-437): totals of the finished blocks
               // plus what this block has accumulated so far, summed over the ranks (eight scalars through a scratch buffer; the
               // accumulator buffer itself is only reduced at the end of the block)
               pimcgpu_scalars part;
               bool have = false;
#ifdef PIMC_WITH_NCCL
               if (comm) {
                  static double *d_tmp8 = nullptr;
                  if (!d_tmp8 && cudaMalloc((void **)&d_tmp8, 8 * sizeof(double)) != cudaSuccess) die("cuda", "cudaMalloc failed");
                  double h8[8];
                  cudaStream_t st = (cudaStream_t)pimcgpu_stream();
                  const double *dacc8 = (const double *)pimcgpu_accum_device_ptr();
                  cudaMemcpyAsync(d_tmp8, dacc8, sizeof h8, cudaMemcpyDeviceToDevice, st);
                  if (ncclAllReduce(d_tmp8, d_tmp8, 8, ncclDouble, ncclSum, comm, st) != ncclSuccess) die("nccl", "ncclAllReduce failed");
                  cudaMemcpyAsync(h8, d_tmp8, sizeof h8, cudaMemcpyDeviceToHost, st);
                  cudaStreamSynchronize(st);
                  part.count = h8[0]; part.kin = h8[1]; part.pot = h8[2]; part.rot = h8[3]; part.rotsq = h8[4]; part.cv = h8[5]; part.cv_trans = h8[6]; part.cv_rot = h8[7];
                  have = true;
               }
#endif
               if (!have) ck(pimcgpu_block_scalars(&part), "pimcgpu_block_scalars");
               if (rank == 0 && part.count > 0) {
                  sums += 1.0;
                  BlockWriters::sum_energy(fsum, sums, total_count + part.count, kin_tot + part.kin, pot_tot + part.pot, rot_tot + part.rot, rotsq_tot + part.rotsq,
                                           cv_tot + part.cv, cvt_tot + part.cv_trans, cvr_tot + part.cv_rot, N, P, d.temperature);
                  sum_row_at = done;
               }
            }
         }
      }
      ck(pimcgpu_sync(), "pimcgpu_sync");
      double *dacc = (double *)pimcgpu_accum_device_ptr();                                   // move counters folded in
#ifdef PIMC_WITH_NCCL
      if (comm) {
         cudaStream_t st = (cudaStream_t)pimcgpu_stream();
         if (ncclAllReduce(dacc, dacc, n_acc, ncclDouble, ncclSum, comm, st) != ncclSuccess) die("nccl", "ncclAllReduce failed");
         cudaStreamSynchronize(st);
      }
#else
      (void)dacc;
#endif
      ck(pimcgpu_accum_download(acc.data(), n_acc), "pimcgpu_accum_download");
      pimcgpu_scalars sc;
      ck(pimcgpu_block_scalars(&sc), "pimcgpu_block_scalars");
      for (int t = 0; t < sys.ntypes; t++)
         bead_updates += sc.mctotal[t][0] * P + sc.mctotal[t][1] * ((1 << d.types[t].levels) - 1) + sc.mctotal[t][2];
      {  // side file with the full state of this rank's chains, written before the reference-format files
         vector<char> blob(pimcgpu_checkpoint_bytes());
         ck(pimcgpu_checkpoint_save(blob.data(), (long)blob.size()), "pimcgpu_checkpoint_save");
         ofstream f(ckname + ".tmp", ios::binary);
         f.write(blob.data(), blob.size());
         f.close();
         rename((ckname + ".tmp").c_str(), ckname.c_str());
      }
      if (rank != 0) continue;
      // end-of-block summary of the same ratios (not in the reference, whose lines come every MCSKIP_RATIO steps above)
      cout << "BLOCK-END:" << setw(8) << block << BLANK << "PASS:" << setw(8) << d.passes << BLANK << "STEP:" << setw(8) << steps_block << BLANK;
      for (int t = 0; t < sys.ntypes; t++)
         cout << setw(8) << d.types[t].name << BLANK << setw(8) << sc.mcaccep[t][0] / sc.mctotal[t][0] << BLANK << setw(8) << sc.mcaccep[t][1] / sc.mctotal[t][1] << BLANK;
      if (Q) cout << BLANK << "Rot: " << setw(8) << sc.mcaccep[imtype][2] / sc.mctotal[imtype][2] << BLANK;
      cout << endl;
      if (d.worm) {
         // worm part of MCSaveAcceptRatio, mc_main.cc:886-921 (this rank's chains)
         double qt[7], qa[7], cq = 0;
         ck(pimcgpu_worm_counters(qt, qa, &cq), "pimcgpu_worm_counters");
         cout << "WORM: open/close " << qa[0] / qt[0] << BLANK << qa[1] / qt[1] << BLANK << qt[0] / cq << "-" << qt[1] / cq << BLANK
              << "advance/recede " << qa[4] / qt[4] << BLANK << qa[5] / qt[5] << BLANK << qt[4] / cq << "-" << qt[5] / cq << BLANK
              << "swap " << qa[6] / qt[6] << BLANK << qt[6] / cq << endl;
      }
      if (block > d.eq_blocks && sc.count > 0) {
         const double ac = sc.count;
         BlockWriters::energy(fname, block, ac, sc.kin, sc.pot, sc.rot, sc.rotsq, sc.cv, sc.cv_trans, sc.cv_rot);                 // SaveEnergy
         // SaveSumEnergy, mc_main.cc:797-836: rows every MCSKIP_TOTAL steps come from the step loop above; a block whose length is
         // not a multiple of MCSKIP_TOTAL still gets its closing row here (the reference would write none)
         kin_tot += sc.kin; pot_tot += sc.pot; rot_tot += sc.rot; rotsq_tot += sc.rotsq; cv_tot += sc.cv; cvt_tot += sc.cv_trans; cvr_tot += sc.cv_rot;
         total_count += ac;
         const double tc = total_count;
         if (sum_row_at != steps_block) {
            sums += 1.0;
            BlockWriters::sum_energy(fsum, sums, tc, kin_tot, pot_tot, rot_tot, rotsq_tot, cv_tot, cvt_tot, cvr_tot, N, P, d.temperature);
         }
         if (Q) {   // SaveRCF, block and accumulated (mc_main.cc:739-740, 463-464)
            const long orc = pimcgpu_accum_offset("rcfcnt");
            for (int it = 0; it < Q; it++) { rcf_sum[it] += acc[off_rcf + it]; rcf_rows_sum[it] += (double)Q * ac; }   // _rcf_sum[1..9] += 1 per time origin
            ostringstream bc; bc << setw(3) << setfill('0') << block;
            BlockWriters::rcf(fname + bc.str(), Q, rottau, ac, &acc[off_rcf], &acc[orc]);
            BlockWriters::rcf(fname + "_sum", Q, rottau, tc, rcf_sum.data(), rcf_rows_sum.data());
         }
         // SaveGraSum, mc_estim.cc:1288-1326
         int na = 0, natypes = 0;
         for (auto &t : d.types) if (!t.molecule) { na = t.numb; natypes = 1; }
         for (int ir = 0; ir < PIMCGPU_BINSR; ir++) gr1d_sum[ir] += acc[off_gr1d + ir];
         BlockWriters::gra_sum(fname, tc, P, natypes, na, gr1d_sum.data());
      }
      if (block > d.eq_blocks && sc.count > 0 && imtype >= 0) {
         // density part of MCSaveBlockAverages (mc_main.cc:715-737) and the accumulated densities (:451-461)
         const double ac = sc.count, tc = total_count;
         ostringstream bc; bc << setw(3) << setfill('0') << block;
         const string bname = fname + bc.str();
         const double *g1 = &acc[off_gr1d], *g2 = &acc[off_gr2d], *g3 = off_gr3d >= 0 ? &acc[off_gr3d] : nullptr;
         for (size_t i = 0; i < gr2d_sum.size(); i++) gr2d_sum[i] += g2[i];
         for (size_t i = 0; i < gr3d_sum.size(); i++) gr3d_sum[i] += g3[i];
         for (size_t i = 0; i < rel_sum.size(); i++) rel_sum[i] += acc[off_rel + i];
         // the reference rewrites the 180 MB <prefix>_sum.g3d after every block (mc_main.cc:458); PIMC_G3D_LAST_ONLY=1 defers it
         const bool g3d_now = block == start_block + d.blocks || !getenv("PIMC_G3D_LAST_ONLY");
         dw.block_and_total(fname, bname, ac, tc, g1, g2, g3, &acc[off_rel], gr2d_sum.data(), gr3d_sum.empty() ? nullptr : gr3d_sum.data(), rel_sum.data(), g3d_now);
      }
      if (block > d.eq_blocks && sc.count > 0 && bstype >= 0) {
         const double ac = sc.count, bmass = d.types[bstype].mass;
         const double lambda = 0.5 * (100.0 * (1.05457266 * 1.05457266) / (1.6605402 * 1.380658)) / bmass;    // mc_setup.cc:206-215
         const long oa = pimcgpu_accum_offset("area"), op = pimcgpu_accum_offset("ploops");
         const int nb = d.types[bstype].numb;
         ck(pimcgpu_download_state(0, nullptr, nullptr, nullptr, pindex.data()), "pimcgpu_download_state");
         BlockWriters::exchange_length(fname, block, ac, nb, &acc[op], pindex.data());      // the permutation row is chain 0's
         if (imtype >= 0 && d.types[imtype].molecule == 1) BlockWriters::area_estimators(fname, block, ac, &acc[oa], beta, lambda, bmass);
         for (int iframe = 0; iframe < 2; iframe++) {
            if (iframe == 1 && !(imtype >= 0 && d.types[imtype].molecule == 2 && d.ispher == 0)) break;
            const double *A = &acc[oa + 6 + 15 * iframe];
            BlockWriters::area_estim3d(fname, block, ac, A, A + 6, iframe, beta, lambda, bmass);
         }
      }
      // checkpoint, mc_main.cc:471-483: yw001.stat / .conf / .tabl in the reference's byte layout, chain 0
      ck(pimcgpu_download_state(0, coords.data(), angles.data(), cosine.data(), bstype >= 0 ? pindex.data() : nullptr), "pimcgpu_download_state");
      auto backup = [&](const char *f) {      // IOFileBackUp, mc_input.cc:796-815: cp <file> <file>.old
         ifstream in(f, ios::binary);
         if (in.good()) { ofstream out(string(f) + ".old", ios::binary); out << in.rdbuf(); }
      };
      backup("yw001.stat"); backup("yw001.conf"); backup("yw001.tabl");
      if (d.worm) backup("yw001.worm");
      { ofstream f("yw001.stat"); f << "STARTBLOCK " << block << endl; }
      {
         ofstream f("yw001.conf", ios::binary);
         streamsize size = sizeof(double) * n;
         f.write((char *)&size, sizeof(streamsize));
         f.write((char *)coords.data(), size);         // MCCoords[0]: the x row only, as the reference does (mc_input.cc:587-590)
         f.write((char *)cosine.data(), size);         // MCCosine[0]
      }
      {
         ofstream f("yw001.tabl", ios::binary);
         for (int a = 0; a < N; a++) rindex[pindex[a]] = a;
         streamsize size = sizeof(int) * N;
         f.write((char *)&size, sizeof(streamsize));
         f.write((char *)pindex.data(), size);
         f.write((char *)rindex.data(), size);
      }
      if (d.worm) {
         // QWormsIO, mc_input.cc:653-688: the TPathWorm record of mc_qworm.h:32-47 (chain 0)
         struct { char stype[80]; int type, exists, ira, masha, atom_i, atom_m; double c; int m; } rec;
         memset(&rec, 0, sizeof rec);
         int st[5];
         ck(pimcgpu_worm_state(0, st), "pimcgpu_worm_state");
         strncpy(rec.stype, d.worm_type.c_str(), sizeof rec.stype - 1);
         rec.type = sys.worm_type; rec.exists = st[0]; rec.ira = st[1]; rec.masha = st[2]; rec.atom_i = st[3]; rec.atom_m = st[4];
         rec.c = d.worm_c * d.density / ((double)d.types[sys.worm_type].numb * P * d.worm_m); rec.m = d.worm_m;
         ofstream f("yw001.worm", ios::binary);
         f.write((char *)&rec, sizeof rec);
      }
      { ofstream f("yw001.mrg"); f << "SEED"; for (int k = 0; k < 6; k++) f << " " << seed[k]; f << "\nSTEP " << pimcgpu_step_counter() << "\nCHAINS " << chains << " RANKS " << ranks << endl; }
      {  // IOxyz, mc_input.cc:690-794 (the checkpoint-time dump of chain 0, mc_main.cc:476-477)
         vector<string> names; vector<int> numbs;
         for (auto &t : d.types) { names.push_back(t.name); numbs.push_back(t.numb); }
         XyzWriters::xyz(fname + ".xyz", sys.ntypes, names.data(), numbs.data(), P, coords.data(), cosine.data());
      }
   }
   double secs = chrono::duration<double>(chrono::steady_clock::now() - t_start).count();
   if (rank == 0)
      cout << "pimc_b200: " << d.blocks << " blocks, " << chains * ranks << " chains, " << secs << " s, " << bead_updates / secs << " bead-updates/s" << endl;
#ifdef PIMC_WITH_NCCL
   if (comm) ncclCommDestroy(comm);
#endif
   pimcgpu_finalize();
   return 0;
}
