// pimc_tables -- command-line front end of the device rho-table generators (include/pimcgpu.h, csrc/pimc_tablegen.cu).
//
// One binary for the reference's three pre-processing programs, each with that program's own argument list, log lines
// and output files, so the reference's workflow (README.md:55, nmv_prop/README, symtop_prop/README, linear_prop/README)
// carries over unchanged:
//
//   pimc_tables asymrho T P iodevn ith0 ithend Arot Brot Crot maxj [--table NAME]     (nmv_prop/asymrho.f, a-run:4)
//   pimc_tables symrho  T P kmod   ith0 ithend Bz Bxy maxj         [--table NAME]     (symtop_prop/symrho.f, a-run:4)
//   pimc_tables linden  T P B npt iodevn                           [--out FILE]       (linear_prop/linden.f)
//
// asymrho/symrho write rho.denXXX (regular table, '(3(I5),3(1x,E15.8))', asymrho.f:716), rho.denXXX_rho, _eng, _esq (one
// E15.8 value per line) for every theta of the range.  With --table NAME and the full range 0..180 they also write
// NAME.rho, NAME.eng, NAME.esq -- what nmv_prop/compile.x concatenates from 181 single-theta jobs and what init_rot3D
// reads (mc_poten.cc:443-499); NAME follows the reference's rule <type>_T<T>t<Q>.  linden writes linden.out (or FILE).
#include "../../include/pimcgpu.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using std::string;

static void die(const char *what)
{
   printf("%s: %s\n", what, pimcgpu_last_error());
   exit(1);
}
static int usage()
{
   printf("usage: pimc_tables asymrho T P iodevn ith0 ithend Arot Brot Crot maxj [--table NAME]\n"
          "       pimc_tables symrho  T P kmod ith0 ithend Bz Bxy maxj [--table NAME]\n"
          "       pimc_tables linden  T P B npt iodevn [--out FILE]\n");
   return 1;
}
static string plane_name(int ith)
{
   char b[32];
   snprintf(b, sizeof b, "rho.den%03d", ith);        // asymrho.f:437-446
   return b;
}
// regular output table of one theta plane (file 2 of the Fortran)
static void write_regular(const string &path, int ith, const double *r, const double *e, const double *q, const string &header)
{
   FILE *f = fopen(path.c_str(), "wb");
   if (!f) { printf("cannot open %s\n", path.c_str()); exit(1); }
   if (ith == 0) fputs(header.c_str(), f);
   char a[16], b[16], c[16];
   for (int iphi = 0; iphi <= 360; iphi++)
      for (int ichi = 0; ichi <= 360; ichi++) {
         const int i = iphi * 361 + ichi;
         pimcgpu_format_e15_8(r[i], 0, a); pimcgpu_format_e15_8(e[i], 0, b); pimcgpu_format_e15_8(q[i], 0, c);
         fprintf(f, "%5d%5d%5d %s %s %s\n", ith, iphi, ichi, a, b, c);
      }
   fclose(f);
}
static void write_planes(int ith0, int ith1, const std::vector<double> &rho, const std::vector<double> &eng, const std::vector<double> &esq,
                         const string &header, const string &table)
{
   const long np = 361L * 361L;
   for (int ith = ith0; ith <= ith1; ith++) {
      const long o = (long)(ith - ith0) * np;
      const string base = plane_name(ith);
      printf("%s\n", base.c_str());
      write_regular(base, ith, &rho[o], &eng[o], &esq[o], header);
      if (pimcgpu_write_e15_8((base + "_rho").c_str(), &rho[o], np, 0)) die("write");
      if (pimcgpu_write_e15_8((base + "_eng").c_str(), &eng[o], np, 0)) die("write");
      if (pimcgpu_write_e15_8((base + "_esq").c_str(), &esq[o], np, 0)) die("write");
   }
   if (!table.empty()) {
      if (ith0 != 0 || ith1 != 180) { printf("--table needs the full theta range 0 180\n"); exit(1); }
      if (pimcgpu_write_e15_8((table + ".rho").c_str(), rho.data(), 181 * np, 0)) die("write");
      if (pimcgpu_write_e15_8((table + ".eng").c_str(), eng.data(), 181 * np, 0)) die("write");
      if (pimcgpu_write_e15_8((table + ".esq").c_str(), esq.data(), 181 * np, 0)) die("write");
      printf("%s.rho %s.eng %s.esq\n", table.c_str(), table.c_str(), table.c_str());
   }
}

int main(int argc, char **argv)
{
   if (argc < 2) return usage();
   const string prog = argv[1];
   std::vector<string> pos;
   string table, out = "linden.out";
   for (int i = 2; i < argc; i++) {
      const string a = argv[i];
      if (a == "--table" && i + 1 < argc) table = argv[++i];
      else if (a == "--out" && i + 1 < argc) out = argv[++i];
      else pos.push_back(a);
   }
   auto D = [&](int i) { return atof(pos[i].c_str()); };
   auto I = [&](int i) { return atoi(pos[i].c_str()); };
   const double boltz = 0.6950356;
   const auto t0 = std::chrono::steady_clock::now();
   if (prog == "asymrho") {
      if (pos.size() != 9) return usage();
      const double T = D(0), A = D(5), B = D(6), C = D(7);
      const int P = I(1), iodevn = I(2), ith0 = I(3), ith1 = I(4), maxj = I(8);
      if (ith1 < ith0 || ith0 < 0 || ith1 > 180) { printf("weird ithe\n"); return 1; }
      const size_t n = (size_t)(ith1 - ith0 + 1) * 361 * 361;
      std::vector<double> rho(n), eng(n), esq(n);
      double info[16];
      printf("tau=%10.5f\n", 1.0 / (boltz * T) / P);                                        // asymrho.f:96
      if (pimcgpu_gen_asymrho(T, P, iodevn, ith0, ith1, A, B, C, maxj, rho.data(), eng.data(), esq.data(), info)) die("asymrho");
      printf(" jmax=%12d\n emax=%20.12f\n", maxj, info[15]);
      const char *lab[3] = {"EVEN K:   ", "ODD  K:   ", "CLASSICAL:"};
      printf("\nAT BETA\n");                                                                // :354-368
      for (int k = 0; k < 3; k++)
         printf("%s Z=%12.6f E=%12.6f CM-1 E=%12.6f K Cv=%12.6f Kb\n", lab[k], info[3 * k], info[3 * k + 1], info[3 * k + 1] / boltz, info[3 * k + 2]);
      printf("\nAT TAU\n");                                                                 // :413-425
      for (int k = 0; k < 3; k++)
         printf("%s Z=%12.6f E=%12.6f CM-1 E=%12.6f K\n", lab[k], info[9 + 2 * k], info[10 + 2 * k], info[10 + 2 * k] / boltz);
      char hdr[160];
      snprintf(hdr, sizeof hdr, "# T=%10.5f NSLICE=%5d IODEVN=%5d\n# the  phi  chi       rho            engrot\n", T, P, iodevn);   // :464-466
      write_planes(ith0, ith1, rho, eng, esq, hdr, table);
   } else if (prog == "symrho") {
      if (pos.size() != 8) return usage();
      const double T = D(0), Bz = D(5), Bxy = D(6);
      const int P = I(1), kmod = I(2), ith0 = I(3), ith1 = I(4), maxj = I(7);
      if (ith1 < ith0 || ith0 < 0 || ith1 > 180) { printf("weird ith\n"); return 1; }
      const size_t n = (size_t)(ith1 - ith0 + 1) * 361 * 361;
      std::vector<double> rho(n), eng(n), esq(n);
      double info[5];
      printf("tau=%10.5f\n", 1.0 / (boltz * T) / P);                                        // symrho.f:52
      if (pimcgpu_gen_symrho(T, P, kmod, ith0, ith1, Bz, Bxy, maxj, rho.data(), eng.data(), esq.data(), info)) die("symrho");
      printf(" ztau= %.15g\n zbeta= %.15g\n Ebeta= %.15g K\n Esqrt= %.15g K^2\n Cv= %.15g Kb\n", info[0], info[1], info[2], info[3], info[4]);   // :100-104
      char hdr[160];
      snprintf(hdr, sizeof hdr, "# T=%10.5f NSLICE=%5d KMOD=%5d\n# the  phi  chi       rho            engrot\n", T, P, kmod);         // :196-198
      write_planes(ith0, ith1, rho, eng, esq, hdr, table);
   } else if (prog == "linden") {
      if (pos.size() != 5) return usage();
      const double T = D(0), B = D(2);
      const int P = I(1), npt = I(3), iodevn = I(4);
      if (npt < 2) return usage();
      std::vector<double> tab((size_t)npt * 4);
      double info[4];
      if (pimcgpu_gen_linden(T, P, B, npt, iodevn, tab.data(), info)) die("linden");
      printf("tau=%10.5f\nlmax=%4d\n", info[0], (int)info[1]);                              // linden.f:24,37
      if (pimcgpu_write_rot(out.c_str(), tab.data(), npt)) die("write");
      char a[16], b[16];
      pimcgpu_format_e15_8(info[2], 0, a); pimcgpu_format_e15_8(info[3], 0, b);
      printf(" beta=  %.15g\nErot at Beta:%s\nCv at Beta:%s\n", info[0] * P, a, b);           // :75,81-82
   } else
      return usage();
   const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
   printf("pimc_tables: %s done in %.3f s\n", prog.c_str(), sec);
   return 0;
}
