// Test shim (not product code): exposes the driver's density writers (moribs-pimc_b200/driver/pimc_writers.h) to ctypes so
// tests/test_writers.py can compare their files with the reference's own Save* functions.  Built on the fly with g++.
#include "../moribs-pimc_b200/driver/pimc_writers.h"

extern "C" void shim_save_densities(const char *prefix, int P, int Q, int ntypes, const int *numb, const int *molecule, double volume, double acount,
                                    const double *gr1d, const double *gr2d, const double *gr3d, const double *rel)
{
   DensityWriters dw;
   dw.P = P; dw.Q = Q; dw.ntypes = ntypes;
   for (int t = 0; t < ntypes; t++) { dw.numb[t] = numb[t]; dw.molecule[t] = molecule[t]; if (molecule[t]) dw.imtype = t; else dw.atype = t; }
   dw.volume = volume;
   dw.block_and_total(prefix, prefix, acount, acount, gr1d, gr2d, gr3d, rel, gr2d, gr3d, rel, true);
}

extern "C" void shim_save_block(const char *prefix, long block, int natoms, int P, int Q, double T, int natomtypes, int na, int nb, int linear, int mff,
                                double lambda, double bmass, double acount, const double *scal7, const double *rcf0, const double *rcf19,
                                const double *gr1d, const double *ploops, const int *pindex, const double *area40)
{
   const string fname = prefix;
   const double beta = 1.0 / T;
   BlockWriters::energy(fname, block, acount, scal7[0], scal7[1], scal7[2], scal7[3], scal7[4], scal7[5], scal7[6]);
   FILE *fs = fopen((fname + "_sum.eng").c_str(), "w");
   BlockWriters::sum_energy(fs, 3.0, acount, scal7[0], scal7[1], scal7[2], scal7[3], scal7[4], scal7[5], scal7[6], natoms, P, T);
   fclose(fs);
   if (Q) {
      BlockWriters::rcf(fname, Q, beta / Q, acount, rcf0, rcf19);
      BlockWriters::rcf(fname + "_sum", Q, beta / Q, acount, rcf0, rcf19);
   }
   BlockWriters::gra_sum(fname, acount, P, natomtypes, na, gr1d);
   if (nb > 0) {
      BlockWriters::exchange_length(fname, block, acount, nb, ploops, pindex);
      if (linear) BlockWriters::area_estimators(fname, block, acount, area40, beta, lambda, bmass);
      BlockWriters::area_estim3d(fname, block, acount, area40 + 6, area40 + 12, 0, beta, lambda, bmass);
      if (mff) BlockWriters::area_estim3d(fname, block, acount, area40 + 21, area40 + 27, 1, beta, lambda, bmass);
   }
}

extern "C" void shim_write_xyz(const char *path_xyz, const char *prefix_ang, int ntypes, const char *const *names, const int *numb, int P,
                               const double *coords, const double *angles, const double *cosine, int nbosons, const int *pindex)
{
   std::string nm[PIMCGPU_MAX_TYPES];
   for (int t = 0; t < ntypes; t++) nm[t] = names[t];
   XyzWriters::xyz(path_xyz, ntypes, nm, numb, P, coords, cosine);
   XyzWriters::xyz_ang(prefix_ang, ntypes, nm, numb, P, coords, angles, nbosons, pindex);
}
