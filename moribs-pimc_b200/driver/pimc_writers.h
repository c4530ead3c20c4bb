// pimc_writers.h -- host-side writers of the reference's density output files, shared by the driver and by the CPU test
// that compares them byte for byte with the reference's own Save* functions (tests/test_writers.py).
#ifndef PIMC_WRITERS_H
#define PIMC_WRITERS_H
#include "../../include/pimcgpu.h"

#include <cmath>
#include <cstdio>
#include <string>

using std::string;

// The density writers of mc_estim.cc:1327-1820,1930-1995 (SaveDensities1D, SaveRho1D, SaveDensities2D, SaveDensities3D,
// SaveRhoThetaChi) on the histograms the device accumulates.  "The density type corresponds to the atom type"
// (mc_estim.cc:228): column id is species id of the deck; only the atom species is ever binned (:577,637,677).
struct DensityWriters {
   int P = 0, Q = 0, ntypes = 0, imtype = -1, atype = -1;
   int numb[PIMCGPU_MAX_TYPES] = {0, 0}, molecule[PIMCGPU_MAX_TYPES] = {0, 0};
   double volume = 1.0;
   static constexpr int BR = PIMCGPU_BINSR, BT = PIMCGPU_BINST, BC = PIMCGPU_BINSC;
   const double dr = 15.0 / BR, dth = M_PI / (BT - 1), dch = 2.0 * M_PI / (BC - 1);     // densities_init, mc_estim.cc:248-254
   static void num(FILE *f, double v) { fprintf(f, "%14.6e   ", v); }                    // setw(IO_WIDTH) << v << BLANK
   int natomtypes() const { return atype >= 0 ? 1 : 0; }
   int ndens3d() const { return ntypes; }                                              // NUMB_ATOMTYPES + NUMB_MOLCTYPES (:246)
   double h3(const double *g, int id, int ir, int it, int ic) const { return (g && id == atype) ? g[((size_t)ir * BT + it) * BC + ic] : 0.0; }

   // SaveDensities1D, mc_estim.cc:1327-1443: <name>.gra and, around a linear rotor, <name>.gri / .grt out of the 2-D histogram
   void densities1d(const string &name, double acount, const double *gr1d, const double *gr2d) const
   {
      FILE *f = fopen((name + ".gra").c_str(), "w");
      if (!f) return;
      const double norm0 = 2.0 * M_PI * dr * acount * (double)P / volume;
      for (int ir = 0; ir < BR; ir++) {
         const double r = ir * dr + 0.5 * dr, r2 = r * r;
         num(f, r);
         for (int id = 0; id < natomtypes(); id++) num(f, gr1d[ir] / (r2 * (norm0 * (double)(numb[atype] * (numb[atype] - 1)))));
         fputc('\n', f);
      }
      fclose(f);
      if (!(imtype >= 0 && molecule[imtype] == 1)) return;
      f = fopen((name + ".gri").c_str(), "w");
      const double norm1 = dr * (double)P * acount * (4.0 * M_PI);
      for (int ir = 0; ir < BR; ir++) {
         const double r = ir * dr + 0.5 * dr, r2 = r * r;
         num(f, r);
         for (int id = 0; id < natomtypes(); id++) {
            double densr = 0.0;
            for (int it = 0; it < BT; it++) densr += gr2d[ir * BT + it];
            num(f, densr / (norm1 * r2));
         }
         fputc('\n', f);
      }
      fclose(f);
      f = fopen((name + ".grt").c_str(), "w");
      const double norm2 = (double)P * acount * dth;
      for (int it = 0; it < BT; it++) {
         num(f, (it * dth + 0.5 * dth) * 180.0 / M_PI);
         for (int id = 0; id < natomtypes(); id++) {
            double denst = 0.0;
            for (int ir = 0; ir < BR; ir++) denst += gr2d[ir * BT + it];
            num(f, denst / (norm2 * (double)numb[atype]));
         }
         fputc('\n', f);
      }
      fclose(f);
   }
   // SaveDensities2D, mc_estim.cc:1666-1725: theta, r, density triples, theta outer
   void densities2d(const string &name, double acount, const double *gr2d) const
   {
      FILE *f = fopen((name + ".g2d").c_str(), "w");
      if (!f) return;
      const double norm3 = dr * dth * (double)P * acount;
      for (int it = 0; it < BT; it++)
         for (int ir = 0; ir < BR; ir++) {
            for (int id = 0; id < natomtypes(); id++) { num(f, (it * dth + 0.5 * dth) * 180.0 / M_PI); num(f, ir * dr + 0.5 * dr); num(f, gr2d[ir * BT + it] / norm3); }
            fputc('\n', f);
         }
      fclose(f);
   }
   // SaveRho1D, mc_estim.cc:1445-1663: .gri/.grt/.grc out of the 3-D histogram (top), total mode adds the relative Euler angles
   void rho1d(const string &name, double acount, const double *g3, const double *rel, bool total) const
   {
      const string base = name + (total ? "_sum" : "");
      FILE *f = fopen((base + ".gri").c_str(), "w");
      if (!f) return;
      const double norm1 = dr * (double)P * acount;                       // no 4 pi here, unlike SaveDensities1D (:1455-1456)
      for (int ir = 0; ir < BR; ir++) {
         num(f, ir * dr + 0.5 * dr);
         for (int id = 0; id < ndens3d(); id++) {
            double densr = 0.0;
            for (int it = 0; it < BT; it++) for (int ic = 0; ic < BC; ic++) densr += h3(g3, id, ir, it, ic);
            num(f, densr / norm1);
         }
         fputc('\n', f);
      }
      fclose(f);
      f = fopen((base + ".grt").c_str(), "w");
      const double norm2 = (double)P * acount * dth * (180.0 / M_PI);
      for (int it = 0; it < BT; it++) {
         num(f, (it * dth + 0.5 * dth) * 180.0 / M_PI);
         for (int id = 0; id < ndens3d(); id++) {
            double denst = 0.0;
            for (int ir = 0; ir < BR; ir++) for (int ic = 0; ic < BC; ic++) denst += h3(g3, id, ir, it, ic);
            num(f, denst / (norm2 * (double)numb[id]));
         }
         fputc('\n', f);
      }
      fclose(f);
      f = fopen((base + ".grc").c_str(), "w");
      const double norm4 = (double)P * acount * dch * (180.0 / M_PI);
      for (int ic = 0; ic < BC; ic++) {
         num(f, (ic * dch + 0.5 * dch) * 180.0 / M_PI);
         for (int id = 0; id < ndens3d(); id++) {
            double densc = 0.0;
            for (int ir = 0; ir < BR; ir++) for (int it = 0; it < BT; it++) densc += h3(g3, id, ir, it, ic);
            num(f, densc / (norm4 * (double)numb[id]));
         }
         fputc('\n', f);
      }
      fclose(f);
      if (!total) return;
      const double norm5 = (double)Q * acount * dch * (180.0 / M_PI), norm6 = (double)Q * acount * dth * (180.0 / M_PI);
      const char *ext[3] = {".eulphi", ".eulchi", ".eulthe"};
      for (int k = 0; k < 3; k++) {
         f = fopen((base + ext[k]).c_str(), "w");
         const int nb = k == 2 ? BT : BC;
         const double *h = k == 0 ? rel + BT : k == 1 ? rel + BT + BC : rel;            // relphi, relchi, relthe
         for (int i = 0; i < nb; i++) {
            const double d = k == 2 ? dth : dch;
            fprintf(f, "%14.6e   %14.6e\n", (i * d + 0.5 * d) * 180.0 / M_PI, h[i] / (k == 2 ? norm6 : norm5));
         }
         fclose(f);
      }
   }
   // SaveDensities3D, mc_estim.cc:1728-1820 (accumulated averages only: mc_main.cc:458, the block call is commented out, :732)
   void densities3d(const string &name, double acount, const double *g3) const
   {
      FILE *f = fopen((name + "_sum.g3d").c_str(), "w");
      if (!f) return;
      const double norm5 = dr * dth * dch * (double)P * acount;
      for (int ir = 0; ir < BR; ir++) {
         const double r = ir * dr + 0.5 * dr;
         for (int it = 0; it < BT; it++) {
            const double theta = it * dth + 0.5 * dth;
            for (int ic = 0; ic < BC; ic++) {
               const double chi = ic * dch + 0.5 * dch;
               for (int id = 0; id < ndens3d(); id++) { num(f, r); num(f, theta * 180.0 / M_PI); num(f, chi * 180.0 / M_PI); num(f, h3(g3, id, ir, it, ic) / (norm5 * r * r * sin(theta))); }
               fputc('\n', f);
            }
         }
      }
      fclose(f);
   }
   // SaveRhoThetaChi, mc_estim.cc:1930-1995: (theta, chi) density, a blank line after every theta row
   void rho_theta_chi(const string &name, double acount, const double *g3) const
   {
      FILE *f = fopen((name + ".gtc").c_str(), "w");
      if (!f) return;
      const double norm6 = dth * dch * (double)P * acount * (180.0 * 180.0 / (M_PI * M_PI));
      for (int it = 0; it < BT; it++) {
         for (int ic = 0; ic < BC; ic++) {
            num(f, (it * dth + 0.5 * dth) * 180.0 / M_PI); num(f, (ic * dch + 0.5 * dch) * 180.0 / M_PI);
            for (int id = 0; id < ndens3d(); id++) {
               double denstc = 0.0;
               for (int ir = 0; ir < BR; ir++) denstc = denstc + h3(g3, id, ir, it, ic);
               num(f, denstc / (norm6 * (double)numb[id]));
            }
            fputc('\n', f);
         }
         fputc('\n', f);
      }
      fclose(f);
   }
   // the density part of MCSaveBlockAverages (mc_main.cc:715-737) followed by the accumulated files main() rewrites after
   // every block (:451-461).  bname = prefix + block number, fname = prefix; *_sum = histograms summed over the blocks.
   void block_and_total(const string &fname, const string &bname, double ac, double tc, const double *g1, const double *g2, const double *g3,
                        const double *rel, const double *g2_sum, const double *g3_sum, const double *rel_sum, bool write_g3d) const
   {
      if (imtype < 0) return;
      if (molecule[imtype] == 1) {
         densities1d(bname, ac, g1, g2);
         densities2d(bname, ac, g2);
         densities2d(fname + "_sum", tc, g2_sum);
      } else {
         densities1d(bname, ac, g1, g2);
         rho1d(bname, ac, g3, rel, false);
         rho_theta_chi(bname, ac, g3);
         if (write_g3d) densities3d(fname, tc, g3_sum);
         rho1d(fname, tc, g3_sum, rel_sum, true);
      }
   }
};

// The per-block scalar writers of mc_main.cc:764-836 and mc_estim.cc:1141-1191,1288-1326,2021-2085,2596-2729.
struct BlockWriters {
   static void num(FILE *f, double v) { fprintf(f, "%14.6e   ", v); }
   // SaveEnergy, mc_main.cc:764-795: appends one row to <prefix>.eng
   static void energy(const string &fname, long block, double ac, double kin, double pot, double rot, double rotsq, double cv, double cvt, double cvr)
   {
      FILE *f = fopen((fname + ".eng").c_str(), "a");
      if (!f) return;
      fprintf(f, "%4ld   ", block);
      num(f, kin / ac); num(f, pot / ac); num(f, (kin + pot) / ac); num(f, rot / ac); num(f, rotsq / ac); num(f, (kin + pot + rot) / ac);
      num(f, cv / ac); num(f, cvt / ac); num(f, cvr / ac);
      fputc('\n', f);
      fclose(f);
   }
   // SaveSumEnergy, mc_main.cc:797-836: one row of the accumulated averages on the open <prefix>_sum.eng
   static void sum_energy(FILE *f, double numb, double acount, double kin, double pot, double rot, double rotsq, double cv, double cvt, double cvr,
                          int natoms, int P, double T)
   {
      const double beta = 1.0 / T;
      fprintf(f, "%4.6e   ", numb);
      num(f, kin / acount); num(f, pot / acount); num(f, (kin + pot) / acount); num(f, rot / acount); num(f, rotsq / acount); num(f, (kin + pot + rot) / acount);
      double Cv = 0.5 * (double)(3 * natoms * P * T) - (kin + pot + rot) / acount;
      Cv = Cv * Cv + cv / acount;
      Cv = -Cv * beta / T;
      num(f, Cv);
      double Cvt = 0.5 * (double)(3 * natoms * P * T) - kin / acount;
      Cvt = Cvt * Cvt + cvt / acount;
      Cvt = -Cvt * beta / T;
      num(f, Cvt);
      double Cvr = -rot / acount;
      Cvr = Cvr * Cvr + cvr / acount;
      Cvr = -Cvr * beta / T;
      num(f, Cvr);
      fputc('\n', f);
      fflush(f);
   }
   // SaveRCF, mc_estim.cc:1141-1191: <n(0).n(t)>, two blank lines, a comment line, then the nine Legendre rows (all equal: rows1to9)
   static void rcf(const string &name, int Q, double rottau, double acount, const double *row0, const double *rows1to9)
   {
      FILE *f = fopen((name + ".rcf").c_str(), "w");
      if (!f) return;
      const double norm = acount * (double)Q;
      for (int it = 0; it <= Q; it++) { num(f, (double)it * rottau); num(f, row0[it % Q] / norm); fputc('\n', f); }
      fputs("\n\n#\n", f);
      for (int it = 0; it <= Q; it++) {
         num(f, (double)it * rottau);
         for (int ip = 1; ip < 10; ip++) num(f, rows1to9[it % Q] / norm);
         fputc('\n', f);
      }
      fclose(f);
   }
   // SaveGraSum, mc_estim.cc:1288-1326
   static void gra_sum(const string &fname, double tc, int P, int natomtypes, int numb, const double *gr1d_sum)
   {
      FILE *f = fopen((fname + "_sum.gra").c_str(), "w");
      if (!f) return;
      const double dr = 15.0 / PIMCGPU_BINSR, norma = dr * tc * (double)P;
      for (int ir = 0; ir < PIMCGPU_BINSR; ir++) {
         num(f, ir * dr + 0.5 * dr);
         for (int id = 0; id < natomtypes; id++) num(f, gr1d_sum[ir] / (norma * (numb * (numb - 1)) / 2.0));
         fputc('\n', f);
      }
      fclose(f);
   }
   // SaveExchangeLength, mc_estim.cc:2021-2085 (GSLOOP_MAX = 7, mc_confg.h:67; PrintXYZprl = 0)
   static void exchange_length(const string &fname, long block, double ac, int nb, const double *ploops, const int *pindex)
   {
      FILE *f = fopen((fname + ".prl").c_str(), "a");
      if (!f) return;
      fprintf(f, "%4ld   ", block);
      double excited = 0.0, ground = 0.0;
      for (int cl = 0; cl < nb; cl++) {
         const double norm = (double)(cl + 1) / (ac * (double)nb);
         if (cl <= 7) excited += (ploops[cl] * norm); else ground += (ploops[cl] * norm);
      }
      num(f, ground); num(f, excited); num(f, ground + excited);
      for (int cl = 0; cl < nb; cl++) num(f, ploops[cl] * ((double)(cl + 1) / (ac * (double)nb)));
      fputc('\n', f);
      for (int a = 0; a < nb; a++) fprintf(f, "%14d   ", pindex[a]);
      fputs("0\n", f);
      fclose(f);
   }
   // SaveAreaEstimators, mc_estim.cc:2596-2640: A = _areas[2] _area2[2] _inert[2]
   static void area_estimators(const string &fname, long block, double ac, const double *A, double beta, double lambda, double mass)
   {
      FILE *f = fopen((fname + ".sup").c_str(), "a");
      if (!f) return;
      const double norm = 2.0 / (beta * lambda);
      fprintf(f, "%4ld   ", block);
      num(f, A[2] * norm / A[4]); num(f, A[3] * norm / A[5]);
      num(f, A[4] * 1.0 * mass / ac); num(f, A[5] * 1.0 * mass / ac);
      num(f, A[0] * sqrt(norm / A[4]) / ac); num(f, A[1] * sqrt(norm / A[5]) / ac);
      fputc('\n', f);
      fclose(f);
   }
   // SaveAreaEstim3D, mc_estim.cc:2670-2729: 9 inertia components, then 6 of 4m^2/(hbar^2 beta) <A_i A_j>
   static void area_estim3d(const string &fname, long block, double ac, const double *areas6, const double *inert9, int iframe, double beta, double lambda, double bmass)
   {
      FILE *f = fopen((fname + (iframe ? ".mffs3d" : ".sffs3d")).c_str(), "a");
      if (!f) return;
      const double norm = 2.0 * bmass / (beta * lambda);
      fprintf(f, "%4ld   ", block);
      for (int k = 0; k < 9; k++) num(f, inert9[k] / ac);
      for (int k = 0; k < 6; k++) num(f, areas6[k] * norm / ac);
      fputc('\n', f);
      fclose(f);
   }
};

// IOxyz (mc_input.cc:690-794) and IOxyzAng (mc_estim.cc:1822-1928): beads of every particle, x/y/z interleaved with the unit
// axis (IOxyz) or with phi, cos(theta), chi (IOxyzAng, which also lists the permutation of the bosons on its first line).
// arrays are the reference layout [dim][atom*P + it]
struct XyzWriters {
   static void num(FILE *f, double v) { fprintf(f, "%14.6e   ", v); }
   static void xyz(const string &path, int ntypes, const string *names, const int *numb, int P, const double *coords, const double *cosine)
   {
      FILE *f = fopen(path.c_str(), "w");
      if (!f) return;
      size_t n = 0;
      for (int t = 0; t < ntypes; t++) n += (size_t)numb[t] * P;
      fprintf(f, "%zu\n#   xyz format:  [atom type]  x y z (Angstrom) \n", n);
      size_t atom = 0;
      for (int t = 0; t < ntypes; t++)
         for (int k = 0; k < numb[t]; k++, atom++)
            for (int it = 0; it < P; it++) {
               const string lab = names[t] + std::to_string(k + 1);
               fprintf(f, "%5s   ", lab.c_str());
               for (int d = 0; d < 3; d++) { num(f, coords[d * n + atom * P + it]); num(f, cosine[d * n + atom * P + it]); }
               fputc('\n', f);
            }
      fclose(f);
   }
   static void xyz_ang(const string &name, int ntypes, const string *names, const int *numb, int P, const double *coords, const double *angles,
                       int nbosons, const int *pindex)
   {
      FILE *f = fopen((name + ".xyz").c_str(), "w");
      if (!f) return;
      size_t n = 0;
      for (int t = 0; t < ntypes; t++) n += (size_t)numb[t] * P;
      fprintf(f, "%zu ", n);
      for (int a = 0; a < nbosons; a++) fprintf(f, "  %d   ", pindex[a]);
      fprintf(f, "\n#   xyz format:  [atom type]  x y z (Angstrom) \n");
      size_t atom = 0;
      for (int t = 0; t < ntypes; t++)
         for (int k = 0; k < numb[t]; k++, atom++)
            for (int it = 0; it < P; it++) {
               fprintf(f, "%s%d", names[t].c_str(), k + 1);
               for (int d = 0; d < 3; d++) { num(f, coords[d * n + atom * P + it]); num(f, angles[d * n + atom * P + it]); }
               fputc('\n', f);
            }
      fclose(f);
   }
};

#endif
