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
