// Per-slice estimator sums on the device: GetKinEnergy, GetPotEnergy_Densities, GetRotEnergy /
// GetRotE3D, GetRCF and the Cv algebra of MCGetAverage (mc_estim.cc:500-1139, mc_main.cc:551-616).
//
// Each estimator kernel uses EB blocks per chain; a block reduces its share in a fixed order and
// writes one partial per quantity, est_finalize adds the partials in block order and the chains in
// chain order (deterministic).  Histogram counts are integer-valued, so they are accumulated with
// FP64 atomics straight into the block accumulator buffer (order-independent).
#pragma once
#include "pimc_device.cuh"
#include "pimc_worm.cuh"

namespace pimc {

constexpr int EST_BLOCKS = 64;       // blocks per chain in the estimator kernels
constexpr int EST_THREADS = 256;
constexpr int NPART = 8;             // partial slots per block: r2avr, pot, srot, sesq, setermsq
constexpr int NAREA = 28;            // area sums per chain: lin (area_perp, area_parl, inert_perp, inert_parl), SFF (area[3], inert[9]), MFF (same)
constexpr int NAREA_ACC = 40;        // accumulator slots: _areas[2] _area2[2] _inert[2] _areas3DSFF[6] _inert3DSFF[9] _areas3DMFF[6] _inert3DMFF[9]
constexpr int BINSR = 300, BINST = 50, BINSC = 100;          // mc_estim.cc:25-27
constexpr double MAX_RADIUS = 15.0, MIN_RADIUS = 0.0;        // mc_estim.cc:29-30

struct EstBuffers {
   double *partials;      // [c][EST_BLOCKS][NPART]
   double *chain_e;       // [c][8]: skin, spot, srot, ErotSQ, Erot_termSQ
   double *chain_rcf;     // [c][2][Q]: sum_it0 n(it0).n(it0+itc), then the number of it0 with n.n < PLONE (rows 1..9 of _rcf)
   double *acc;           // accumulator buffer
   double *com;           // [c][P][3] total centre of mass per slice (space-fixed-frame area estimator)
   double *area_partials; // [c][EST_BLOCKS][NAREA]
   double *chain_area;    // [c][NAREA]
   long off_gr1d, off_gr2d, off_gr3d, off_rcf, off_relbins, off_area, off_ploops, off_rcfcnt;
   int has_gr3d;
   const int *pairs;      // [npairs][2]
   int npairs;
};

__device__ __forceinline__ double block_sum(double v, double *red)
{
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
   __syncthreads();
   if ((threadIdx.x & 31) == 0) red[warp] = v;
   __syncthreads();
   double s = 0.0;
   for (int w = 0; w < nwarp; w++) s += red[w];
   return s;
}

__device__ __forceinline__ void bin_r(const EstBuffers &e, double r, int *bin)
{
   const double delta_radius = (MAX_RADIUS - MIN_RADIUS) / (double)BINSR;
   *bin = (int)floor((r - MIN_RADIUS) / delta_radius);
}

// one thread-block slice of all three energy estimators for chain c = blockIdx.x / EST_BLOCKS
__global__ void __launch_bounds__(EST_THREADS)
est_energy_kernel(const __grid_constant__ Params p, const __grid_constant__ EstBuffers e, int with_dens)
{
   extern __shared__ double smem[];
   double *red = smem;
   const int c = blockIdx.x / EST_BLOCKS, b = blockIdx.x % EST_BLOCKS;
   const int P = p.P, N = p.N, Q = p.Q;
   const int gt = b * blockDim.x + threadIdx.x, nt = EST_BLOCKS * blockDim.x;
   if (p.worm_on && p.wstate[(size_t)c * 8]) return;      // G sector: no estimators (mc_main.cc:389-391, mc_estim.cc:506-507)
   // the pair-distance histogram gets P N(N-1)/2 counts per chain and measurement on a few hundred bins: count in shared
   // memory (integers, so the result does not depend on the order) and flush once per block
   __shared__ unsigned int h1[BINSR];
   for (int i = threadIdx.x; i < BINSR; i += blockDim.x) h1[i] = 0u;
   __syncthreads();
   SmallTables t;
   t.g1d = p.g1d; t.v1d = p.v1d; t.y2_1d = p.y2_1d; t.lut1d = p.lut1d;
   t.rgrid = p.rgrid; t.rdens = p.rdens; t.rdens2 = p.rdens2; t.lutrot = p.lutrot;
   t.rec1d = p.rec1d; t.recrot = p.recrot; t.rgi2d = p.rgi2d; t.cgi2d = p.cgi2d;
   const double delta_theta = PI / (double)(BINST - 1), delta_chi = 2.0 * PI / (double)(BINSC - 1);

   // ---- GetKinEnergy, mc_estim.cc:876-937: sum_atoms sum_it |r_it - r_it+1|^2 / (4 beta lambda)
   double r2 = 0.0;
   for (long i = gt; i < (long)N * P; i += nt) {
      int a = (int)(i % N), it = (int)(i / N);
      int type = type_of(p, a);
      int a1 = a, it1 = it + 1;
      if (it1 == P) { it1 = 0; if (p.stat[type] == 1) a1 = p.pindex[(size_t)c * N + a]; }
      double s = 0.0;
      #pragma unroll
      for (int d = 0; d < 3; d++) {
         double dr = p.pos[pos_index(p, c, it, d, a)] - p.pos[pos_index(p, c, it1, d, a1)];
         if (p.minimage) dr -= p.box[d] * rint(dr / p.box[d]);
         s += dr * dr;
      }
      r2 += s / (4.0 * p.beta * p.lambda[type]);
   }
   r2 = block_sum(r2, red);

   // ---- GetPotEnergy_Densities, mc_estim.cc:500-687
   double pot = 0.0;
   // (pair, slice) items; 32-bit index arithmetic whenever the count allows (a 64-bit divide per item costs more than the spline)
   const long nitems = (long)e.npairs * P;
   const bool small = nitems < (1L << 31);
   for (long i = gt; i < nitems; i += nt) {
      int pi, it;
      if (small) { const unsigned iu = (unsigned)i, np = (unsigned)e.npairs; it = (int)(iu / np); pi = (int)(iu - (unsigned)it * np); }
      else { pi = (int)(i % e.npairs); it = (int)(i / e.npairs); }
      int a0 = e.pairs[2 * pi], a1 = e.pairs[2 * pi + 1];
      int mode = p.mode[type_of(p, a0)][type_of(p, a1)];
      double p0[3], p1[3], dr[3], dr2 = 0.0;
      #pragma unroll
      for (int d = 0; d < 3; d++) {
         p0[d] = p.pos[pos_index(p, c, it, d, a0)];
         p1[d] = p.pos[pos_index(p, c, it, d, a1)];
         dr[d] = p0[d] - p1[d];
         if (p.minimage) dr[d] -= p.box[d] * rint(dr[d] / p.box[d]);
         dr2 += dr[d] * dr[d];
      }
      int q = it / p.R;
      if (mode == M_LIN_0MOL || mode == M_LIN_1MOL) {
         double r = sqrt(dr2);
         int m = ((mode == M_LIN_0MOL) ? a0 : a1) - p.first[p.imtype];
         double cost = 0.0;
         #pragma unroll
         for (int d = 0; d < 3; d++) cost += p.cosn[ang_index(p, c, q, d, m)] * dr[d];
         cost /= r;
         cost *= (mode == M_LIN_0MOL) ? -1 : 1;
         if (with_dens) {          // bin_2Ddensity, mc_estim.cc:1233-1250
            int br; bin_r(e, r, &br);
            if (br < BINSR && br >= 0) {
               int bt = (int)floor(acos(cost) / delta_theta);
               if (bt < BINST && bt >= 0) atomicAdd(e.acc + e.off_gr2d + br * BINST + bt, 1.0);
            }
         }
         pot += lpot2d(p, t, r, cost);
      } else if (mode == M_TOP_0MOL || mode == M_TOP_1MOL || mode == M_SPHER) {
         double rtc[3], v;
         if (p.ispher == 0) {
            Mat3 rm;
            if (mode == M_TOP_0MOL) { load_rotmat(p, c, q, a0 - p.first[p.imtype], rm); v = vcord(p, rm, p0, p1, rtc, nullptr); }
            else { load_rotmat(p, c, q, a1 - p.first[p.imtype], rm); v = vcord(p, rm, p1, p0, rtc, nullptr); }
         } else {
            // vspher_ overwrites its r argument with the clamped value in bohr, which is then binned (mc_estim.cc:631-637)
            double r = sqrt(dr2);
            v = vspher(p, r, &rtc[0]);
            rtc[1] = 0.0; rtc[2] = 0.0;
         }
         if (with_dens && e.has_gr3d) {   // bin_3Ddensity, mc_estim.cc:1252-1273
            int br; bin_r(e, rtc[0], &br);
            if (br < BINSR && br >= 0) {
               int bt = (int)floor(rtc[1] / delta_theta);
               if (bt < BINST && bt >= 0) {
                  int bc = (int)floor(rtc[2] / delta_chi);
                  if (bc < BINSC && bc >= 0) atomicAdd(e.acc + e.off_gr3d + ((size_t)br * BINST + bt) * BINSC + bc, 1.0);
               }
            }
         }
         pot += v;
      } else if (mode == M_TOPTOP) {
         Mat3 ra, rb;
         load_rotmat(p, c, q, a0 - p.first[p.imtype], ra);
         load_rotmat(p, c, q, a1 - p.first[p.imtype], rb);
         Tip4pSites sa, sb;
         tip4p_sites(ra, p0, sa);
         tip4p_sites(rb, p1, sb);
         pot += caleng(sa, sb);
      } else {
         double r = sqrt(dr2);
         if (with_dens) { int br; bin_r(e, r, &br); if (br < BINSR && br >= 0) atomicAdd(&h1[br], 1u); }
         // grid uniform to rounding: the per-interval cubic of the move kernel (same spline piece, 1e-15), else the packed records
         bool bad = !p.poly1d;
         double v = 0.0;
         if (p.poly1d) {
            const double xq = (r - p.x0_1d) * p.invh_1d;
            int k = max(0, min((int)xq, p.n1d - 2));
            bad = !(r > p.x0_1d && r < p.xn_1d);
            const double tt = xq - (double)k;
            const double2 A = __ldg(p.pa1d + k), B = __ldg(p.pb1d + k);
            v = fma(fma(fma(B.y, tt, B.x), tt, A.y), tt, A.x);
         }
         pot += bad ? spot1d(p, t, r) : v;
      }
   }
   pot = block_sum(pot, red);
   __syncthreads();
   for (int i = threadIdx.x; i < BINSR; i += blockDim.x)
      if (h1[i]) atomicAdd(e.acc + e.off_gr1d + i, (double)h1[i]);

   // ---- GetRotE3D (mc_estim.cc:989-1096) / GetRotEnergy (:939-987), RotDenType 0
   double srot = 0.0, sesq = 0.0, sterm = 0.0;
   if (Q > 0 && p.imtype >= 0 && p.rotden_type == 1) {
      // rattle-and-shake estimators (GetRotE3D with RotDenType 1, mc_estim.cc:1003-1092; GetRotEnergy :970-984).  The
      // reference folds the analytic offset into the running sum once per molecule, so the molecules are taken in
      // sequence by block 0 of the chain.
      if (b == 0) {
         const int nm = p.numb[p.imtype];
         if (p.molecule[p.imtype] == 2) {
            const int skip = p.rnratio;
            const double nq = (double)(Q / skip), tc = p.rottau / WNO2K;
            double E = 0.0, ESQ = 0.0, ETERM = 0.0;
            for (int m = 0; m < nm; m++) {
               double s = 0.0, q2 = 0.0, t2 = 0.0;
               for (int q0 = threadIdx.x * skip; q0 < Q; q0 += blockDim.x * skip) {
                  const int q1 = (q0 + skip) % Q;
                  Mat3 r0, r1;
                  load_rotmat(p, c, q0, m, r0);
                  load_rotmat(p, c, q1, m, r1);
                  double rel[3], erot = 0.0, esq = 0.0, rho = 0.0;
                  if (p.rho3) rho = rotden(p, r0, r1, rel, &erot, &esq, nullptr, nullptr);      // Noya tables for the coarse slices
                  else deleul(r0, r1, rel[0], rel[1], rel[2]);
                  if (with_dens) {
                     int bt = (int)floor(rel[1] / delta_theta);
                     if (bt < BINST && bt >= 0) atomicAdd(e.acc + e.off_relbins + bt, (double)skip);
                     int bp = (int)floor(rel[0] / delta_chi);
                     if (bp < BINSC && bp >= 0) atomicAdd(e.acc + e.off_relbins + BINST + bp, (double)skip);
                     int bc = (int)floor(rel[2] / delta_chi);
                     if (bc < BINSC && bc >= 0) atomicAdd(e.acc + e.off_relbins + BINST + BINSC + bc, (double)skip);
                  }
                  if (skip == 1) { rho = rsrot(p, r0, r1, &erot); s += rho; }
                  else s += erot;
                  q2 += esq;
                  t2 += erot * erot;
               }
               s = block_sum(s, red); q2 = block_sum(q2, red); t2 = block_sum(t2, red);
               E += s / nq; ESQ += q2 / (nq * nq); ETERM += t2 / (nq * nq);
               if (skip == 1) {
                  E = E / (4.0 * tc * tc);
                  E += 0.25 * (p.xrot + p.yrot + p.zrot) + 1.5 / tc;
                  E = E / WNO2K;
               }
            }
            if (threadIdx.x == 0) { srot = E; sesq = ESQ; sterm = ETERM; }
         } else {
            double s = 0.0;
            for (int q0 = threadIdx.x; q0 < Q; q0 += blockDim.x) {
               const int q1 = (q0 + 1) % Q;
               double p0 = 0.0;
               #pragma unroll
               for (int d = 0; d < 3; d++) p0 += p.cosn[ang_index(p, c, q0, d, 0)] * p.cosn[ang_index(p, c, q1, d, 0)];
               s += rsline(p, p0, nullptr);
            }
            s = block_sum(s, red);
            if (threadIdx.x == 0) { s = s / (double)Q; srot = s / p.rottau + 1.0 / p.rottau; }
         }
      }
   } else if (Q > 0 && p.imtype >= 0) {
      const int nm = p.numb[p.imtype];
      const bool top = p.molecule[p.imtype] == 2;
      for (int i = gt; i < nm * Q; i += nt) {
         int m = i / Q, q0 = i % Q, q1 = (q0 + 1) % Q;
         if (top) {
            Mat3 r0, r1;
            load_rotmat(p, c, q0, m, r0);
            load_rotmat(p, c, q1, m, r1);
            double rel[3], erot, esq;
            rotden(p, r0, r1, rel, &erot, &esq, nullptr, nullptr);
            if (with_dens) {
               int bt = (int)floor(rel[1] / delta_theta);
               if (bt < BINST && bt >= 0) atomicAdd(e.acc + e.off_relbins + bt, 1.0);
               int bp = (int)floor(rel[0] / delta_chi);
               if (bp < BINSC && bp >= 0) atomicAdd(e.acc + e.off_relbins + BINST + bp, 1.0);
               int bc = (int)floor(rel[2] / delta_chi);
               if (bc < BINSC && bc >= 0) atomicAdd(e.acc + e.off_relbins + BINST + BINSC + bc, 1.0);
            }
            srot += erot / (double)Q;
            sesq += esq / ((double)Q * (double)Q);
            sterm += erot * erot / ((double)Q * (double)Q);
         } else if (m == 0) {
            double p0 = 0.0;
            #pragma unroll
            for (int d = 0; d < 3; d++) p0 += p.cosn[ang_index(p, c, q0, d, 0)] * p.cosn[ang_index(p, c, q1, d, 0)];
            double rdens = srot_eval(p, p.rgrid, p.rdens, p.rdens2, p.lutrot, p0, 0);
            double rderv = srot_eval(p, p.rgrid, p.rderv, p.rderv2, p.lutrot, p0, 1);
            double resqr = srot_eval(p, p.rgrid, p.resqr, p.resqr2, p.lutrot, p0, 2);
            if (fabs(rdens) > RZERO) srot += rderv / rdens;
            sterm += (rderv / rdens) * (rderv / rdens);
            sesq += resqr / rdens;
         }
      }
   }
   srot = block_sum(srot, red);
   sesq = block_sum(sesq, red);
   sterm = block_sum(sterm, red);
   if (threadIdx.x == 0) {
      double *o = e.partials + ((size_t)c * EST_BLOCKS + b) * NPART;
      o[0] = r2; o[1] = pot; o[2] = srot; o[3] = sesq; o[4] = sterm;
   }
}

// total centre of mass of every slice (GetAreaEstim3D iframe 0, mc_estim.cc:2283-2302): atoms in sequence, one thread
// per (chain, slice, dimension)
__global__ void est_com_kernel(const __grid_constant__ Params p, const __grid_constant__ EstBuffers e)
{
   const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= (long)p.nchains * p.P * 3) return;
   const int d = (int)(i % 3), it = (int)((i / 3) % p.P), c = (int)(i / (3L * p.P));
   double tmass = 0.0, s = 0.0;
   for (int a = 0; a < p.N; a++) {
      const double mass = p.mass[type_of(p, a)];
      s += (mass * p.pos[pos_index(p, c, it, d, a)]);
      tmass += mass;
   }
   e.com[i] = s / tmass;
}

// dr0 x n, dr1 x n products of the classical-inertia terms
__device__ __forceinline__ void cross3(const double *a, const double *b, double *o)
{
   o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
// area vector + inertia tensor of one link (dr0 -> dr1) on the axes hat[3][3]: GetAreaEstim3D, mc_estim.cc:2418-2538
__device__ __forceinline__ void area3d_link(const double *dr0, const double *dr1, const double (&hat)[3][3], double bmass, double *out12)
{
   double area[3];
   cross3(dr0, dr1, area);
   #pragma unroll
   for (int k = 0; k < 3; k++) area[k] *= 0.5;
   #pragma unroll
   for (int id = 0; id < 3; id++) {
      #pragma unroll
      for (int k = 0; k < 3; k++) out12[k] += area[id] * hat[k][id];
   }
   #pragma unroll
   for (int id = 0; id < 3; id++) {
      double rn0[3], rn1[3];
      cross3(dr0, hat[id], rn0);
      cross3(dr1, hat[id], rn1);
      double sum = 0.0;
      #pragma unroll
      for (int d = 0; d < 3; d++) sum += rn0[d] * rn1[d] * bmass;
      out12[3 + id * 3 + id] += sum;
      double dr0_id = 0.0;
      #pragma unroll
      for (int d = 0; d < 3; d++) dr0_id += hat[id][d] * dr0[d];
      #pragma unroll
      for (int jd = 0; jd < 3; jd++)
         if (jd != id) {
            double dr1_jd = 0.0;
            #pragma unroll
            for (int d = 0; d < 3; d++) dr1_jd += hat[jd][d] * dr1[d];
            out12[3 + id * 3 + jd] += -bmass * dr0_id * dr1_jd;
         }
   }
}

// GetAreaEstimators (mc_estim.cc:2087-2250) and GetAreaEstim3D for both frames (:2252-2594): one item per
// (slice, boson); links close onto world line PIndex[atom] at it0 = P-1
__global__ void __launch_bounds__(EST_THREADS)
est_area_kernel(const __grid_constant__ Params p, const __grid_constant__ EstBuffers e)
{
   __shared__ double red[32];
   const int c = blockIdx.x / EST_BLOCKS, b = blockIdx.x % EST_BLOCKS;
   if (p.worm_on && p.wstate[(size_t)c * 8]) return;
   const int P = p.P, N = p.N, bt = p.bstype;
   const int nb = p.numb[bt], b0 = p.first[bt];
   const int gt = b * blockDim.x + threadIdx.x, nt = EST_BLOCKS * blockDim.x;
   const bool lin = p.imtype >= 0 && p.molecule[p.imtype] == 1, mff = p.imtype >= 0 && p.molecule[p.imtype] == 2 && p.ispher == 0;
   const int gm = p.imtype >= 0 ? p.first[p.imtype] : 0;        // the (first) dopant molecule
   const double bmass = p.mass[bt];
   double acc[NAREA];
   #pragma unroll
   for (int k = 0; k < NAREA; k++) acc[k] = 0.0;
   const double ident[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
   for (long i = gt; i < (long)P * nb; i += nt) {
      const int a = b0 + (int)(i % nb), it0 = (int)(i / nb);
      int it1 = it0 + 1, a1 = a;
      if (it1 == P) { it1 = 0; a1 = p.pindex[(size_t)c * N + a]; }
      double r0[3], r1[3], dr0[3], dr1[3];
      #pragma unroll
      for (int d = 0; d < 3; d++) { r0[d] = p.pos[pos_index(p, c, it0, d, a)]; r1[d] = p.pos[pos_index(p, c, it1, d, a1)]; }
      // space-fixed frame about the total centre of mass
      #pragma unroll
      for (int d = 0; d < 3; d++) { dr0[d] = r0[d] - e.com[((size_t)c * P + it0) * 3 + d]; dr1[d] = r1[d] - e.com[((size_t)c * P + it1) * 3 + d]; }
      area3d_link(dr0, dr1, ident, bmass, acc + 4);
      if (lin || mff) {
         #pragma unroll
         for (int d = 0; d < 3; d++) { dr0[d] = r0[d] - p.pos[pos_index(p, c, it0, d, gm)]; dr1[d] = r1[d] - p.pos[pos_index(p, c, it1, d, gm)]; }
      }
      const int q = it0 / p.R;
      if (mff) {
         Mat3 rm;
         load_rotmat(p, c, q, 0, rm);
         double hat[3][3];                         // hatx, haty, hatz = columns of the rotation matrix (vcord_ ivcord = 1)
         #pragma unroll
         for (int k = 0; k < 3; k++)
            #pragma unroll
            for (int d = 0; d < 3; d++) hat[k][d] = rm.m[d][k];
         area3d_link(dr0, dr1, hat, bmass, acc + 16);
      }
      if (lin) {
         double n_parl[3], n_perp[3], area[3], rn0[3], rn1[3];
         #pragma unroll
         for (int d = 0; d < 3; d++) n_parl[d] = p.cosn[ang_index(p, c, q, d, 0)];
         const double zero = 10e-4;
         double tg = 0.0, st = 1.0;
         if (fabs(n_parl[0]) > zero) { tg = n_parl[1] / n_parl[0]; st = sqrt(1.0 + tg * tg); }
         n_perp[0] = tg / st; n_perp[1] = -1.0 / st; n_perp[2] = 0.0;
         cross3(dr0, dr1, area);
         #pragma unroll
         for (int d = 0; d < 3; d++) { acc[0] += (n_perp[d] * (0.5 * area[d])); acc[1] += (n_parl[d] * (0.5 * area[d])); }
         cross3(n_perp, dr0, rn0); cross3(n_perp, dr1, rn1);
         #pragma unroll
         for (int d = 0; d < 3; d++) acc[2] += (rn0[d] * rn1[d]);
         cross3(n_parl, dr0, rn0); cross3(n_parl, dr1, rn1);
         #pragma unroll
         for (int d = 0; d < 3; d++) acc[3] += (rn0[d] * rn1[d]);
      }
   }
   double *o = e.area_partials + ((size_t)c * EST_BLOCKS + b) * NAREA;
   for (int k = 0; k < NAREA; k++) {
      double v = block_sum(acc[k], red);
      if (threadIdx.x == 0) o[k] = v;
   }
}

// GetRCF row 0 (mc_estim.cc:1099-1139): rcf[itc] = sum_it0 n(it0).n(it0+itc) for the first rotor
__global__ void est_rcf_kernel(const __grid_constant__ Params p, const __grid_constant__ EstBuffers e)
{
   const int c = blockIdx.y, Q = p.Q;
   const int itc = blockIdx.x * blockDim.x + threadIdx.x;
   if (itc >= Q) return;
   double s = 0.0, below = 0.0;
   for (int it0 = 0; it0 < Q; it0++) {
      int tc = (it0 + itc) % Q;
      double p0 = 0.0;
      #pragma unroll
      for (int d = 0; d < 3; d++) p0 += p.cosn[ang_index(p, c, it0, d, 0)] * p.cosn[ang_index(p, c, tc, d, 0)];
      s += p0;
      // rows 1..9 of the BLOCK array: `if (p0<PLONE)` governs `_rcf[in][itc] += pleg` with pleg = 1 since the Legendre call is
      // commented out (mc_estim.cc:1127-1137, PLONE = 0.9999999 mc_confg.h:64); _rcf_sum[in] gets 1 unconditionally
      if (p0 < 0.9999999) below += 1.0;
   }
   e.chain_rcf[(size_t)c * 2 * Q + itc] = s;
   e.chain_rcf[(size_t)c * 2 * Q + Q + itc] = below;
}

// per-chain totals of the block partials, one thread per (chain, quantity), blocks summed in their fixed order
__global__ void est_chain_totals_kernel(const __grid_constant__ Params p, const __grid_constant__ EstBuffers e)
{
   const int per = 5 + NAREA;
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= p.nchains * per) return;
   const int c = t / per, k = t % per;
   if (p.worm_on && p.wstate[(size_t)c * 8]) return;       // G sector: the estimator kernels wrote nothing for this chain
   double s = 0.0;
   if (k < 5) {
      for (int b = 0; b < EST_BLOCKS; b++) s += e.partials[((size_t)c * EST_BLOCKS + b) * NPART + k];
      e.chain_e[(size_t)c * 8 + k] = s;            // raw sums; est_finalize_kernel turns them into the estimators
   } else if (p.bstype >= 0) {
      for (int b = 0; b < EST_BLOCKS; b++) s += e.area_partials[((size_t)c * EST_BLOCKS + b) * NAREA + (k - 5)];
      e.chain_area[(size_t)c * NAREA + (k - 5)] = s;
   }
}

// Cv algebra (mc_main.cc:589-616), accumulation in chain order
__global__ void est_finalize_kernel(const __grid_constant__ Params p, const __grid_constant__ EstBuffers e, int accumulate)
{
   const int tid = blockIdx.x * blockDim.x + threadIdx.x;
   const int Q = p.Q;
   if (tid == 0) {
      for (int c = 0; c < p.nchains; c++) {
         if (p.worm_on && p.wstate[(size_t)c * 8]) continue;      // a chain in the G sector contributes no sample
         double *ce = e.chain_e + (size_t)c * 8;               // raw block sums from est_chain_totals_kernel
         const double r2 = ce[0], pot = ce[1], srot = ce[2], sesq = ce[3], sterm = ce[4];
         double skin = (double)p.P * p.temperature * (0.5 * (double)(3 * p.N) - r2);
         double spot = pot / (double)p.P;
         ce[0] = skin; ce[1] = spot; ce[2] = srot; ce[3] = sesq; ce[4] = sterm;
         if (accumulate) {
            double nd = (double)(3 * p.N);
            double kterm = nd * 0.5 / p.tau - skin;
            double sCv = -0.5 * nd / (p.beta * p.tau) - (kterm - spot - srot) * (kterm - spot - srot) + (2.0 / p.beta) * (0.5 * nd / p.tau - skin) + sterm - sesq;
            double sCv_trans = -0.5 * nd / (p.beta * p.tau) - kterm * kterm + (2.0 / p.beta) * (0.5 * nd / p.tau - skin);
            double sCv_rot = -srot * srot + sterm - sesq;
            double *a = e.acc;
            a[0] += 1.0; a[1] += skin; a[2] += spot; a[3] += srot; a[4] += sesq; a[5] += sCv; a[6] += sCv_trans; a[7] += sCv_rot;
         }
      }
   }
   if (tid == 32 && p.bstype >= 0) {
      const int nb = p.numb[p.bstype], b0 = p.first[p.bstype];
      const bool lin = p.imtype >= 0 && p.molecule[p.imtype] == 1, mff = p.imtype >= 0 && p.molecule[p.imtype] == 2 && p.ispher == 0;
      for (int c = 0; c < p.nchains; c++) {
         if (p.worm_on && p.wstate[(size_t)c * 8]) continue;
         const double *ca = e.chain_area + (size_t)c * NAREA;   // summed by est_chain_totals_kernel
         if (!accumulate) continue;
         double *a = e.acc + e.off_area;
         if (lin) {            // mc_estim.cc:2242-2249
            a[0] += ca[0]; a[1] += ca[1]; a[2] += ca[0] * ca[0]; a[3] += ca[1] * ca[1];
            a[4] += ca[2] / (double)p.P; a[5] += ca[3] / (double)p.P;
         }
         for (int f = 0; f < 2; f++) {     // space-fixed, then dopant-fixed frame (mc_estim.cc:2565-2590)
            if (f == 1 && !mff) break;
            const double *ap = ca + 4 + 12 * f, *ic = ap + 3;
            double *aa = a + 6 + 15 * f, *ai = aa + 6;
            int ind = 0;
            for (int id = 0; id < 3; id++)
               for (int jd = 0; jd <= id; jd++) aa[ind++] += ap[id] * ap[jd];
            for (int k = 0; k < 9; k++) ai[k] += ic[k] / (double)p.P;
         }
         // GetExchangeLength, mc_estim.cc:1997-2019: one count per permutation cycle, binned by length - 1
         const int *pi = p.pindex + (size_t)c * p.N;
         for (int at = 0; at < nb; at++) {
            int len = 0, pa = pi[b0 + at] - b0;
            bool first = true;                    // `at` is the smallest member of its cycle
            while (pa != at) { if (pa < at) first = false; pa = pi[b0 + pa] - b0; len++; }
            if (first) e.acc[e.off_ploops + len] += 1.0;
         }
      }
   }
   if (accumulate && Q > 0)
      for (int itc = tid; itc < Q; itc += gridDim.x * blockDim.x) {
         double s = 0.0, b = 0.0;
         for (int c = 0; c < p.nchains; c++)
            if (!(p.worm_on && p.wstate[(size_t)c * 8])) { s += e.chain_rcf[(size_t)c * 2 * Q + itc]; b += e.chain_rcf[(size_t)c * 2 * Q + Q + itc]; }
         e.acc[e.off_rcf + itc] += s;
         e.acc[e.off_rcfcnt + itc] += b;
      }
}

// MCTotal/MCAccep of every chain folded into the accumulator scalars (slots 8..19)
__global__ void fold_counters_kernel(const __grid_constant__ Params p, double *acc)
{
   int i = threadIdx.x;       // (type, move, which) -> 12 slots
   if (i >= MAXT * 3 * 2) return;
   int which = i % 2, move = (i / 2) % 3, type = i / 6;
   double s = 0.0;
   for (int c = 0; c < p.nchains; c++) s += p.counters[(((size_t)c * MAXT + type) * 3 + move) * 2 + which];
   acc[8 + which * 6 + type * 3 + move] = s;
}

// Symmetry operations at the end of MCGetAverage (mc_main.cc:647-692): Reflect_MF_XZ / _YZ / _XY (mc_piqmc.cc:1385-1708
// with rflmfy/x/z, vcord.f:257-546) and RotSymConfig (mc_piqmc.cc:1710-1794).  One block per chain.  ops == nullptr:
// each enabled operation is applied with probability 1/2, uniforms from the chain's miscellaneous stream in the order
// REFLECTY, REFLECTX, REFLECTZ, ROTSYM (+ one more for the rotor pick); otherwise ops[c][4] = {XZ, YZ, XY flags, rotor
// index or -1}.  As in the reference, the reflections leave MCCosine untouched and flip the y coordinate of every bead.
__global__ void symmetry_kernel(const __grid_constant__ Params p, const int *ops)
{
   __shared__ int op[4];
   const int c = blockIdx.x, Q = p.Q, nm = p.NM;
   if (p.worm_on && p.wstate[(size_t)c * 8]) return;      // applied inside MCGetAverage: Z sector only
   if (threadIdx.x == 0) {
      op[0] = op[1] = op[2] = 0; op[3] = -1;
      if (ops) { for (int k = 0; k < 4; k++) op[k] = ops[c * 4 + k]; }
      else {
         uint32_t *sp = p.rng + ((size_t)c * p.S + p.P + p.Q) * 6;
         Mrg rs;
         mrg_load(rs, sp);
         if (p.refl[1] && mrg_u01(rs) < 0.5) op[0] = 1;
         if (p.refl[0] && mrg_u01(rs) < 0.5) op[1] = 1;
         if (p.refl[2] && mrg_u01(rs) < 0.5) op[2] = 1;
         if (p.rotsym && mrg_u01(rs) < 0.5) {
            const double u = mrg_u01(rs);
            for (int m = 0; m < nm; m++)
               if (u > (double)m / (double)nm && u <= (double)(m + 1) / (double)nm) op[3] = m;
         }
         mrg_store(rs, sp);
      }
      if (op[0] | op[1] | op[2] | (op[3] >= 0)) p.pos_epoch[c] += 1;        // cached rotor potentials are stale
   }
   __syncthreads();
   if (!(op[0] | op[1] | op[2]) && op[3] < 0) return;
   const bool top = p.imtype >= 0 && p.molecule[p.imtype] == 2;
   for (int i = threadIdx.x; i < nm * Q; i += blockDim.x) {
      const int m = i / Q, q = i % Q;
      double phi = p.ang[ang_index(p, c, q, 0, m)], cth = p.ang[ang_index(p, c, q, 1, m)], chi = p.ang[ang_index(p, c, q, 2, m)];
      if (top)
         for (int k = 0; k < 3; k++) {
            if (!op[k]) continue;
            Mat3 r;
            matpre(phi, acos(cth), chi, r);
            double hx[3], hy[3], hz[3];
            #pragma unroll
            for (int d = 0; d < 3; d++) { hx[d] = r.m[d][0]; hy[d] = r.m[d][1]; hz[d] = r.m[d][2]; }
            if (k == 0) { hx[1] = -hx[1]; hz[1] = -hz[1]; cross3(hz, hx, hy); }          // rflmfy: y-hat = z x x
            else if (k == 1) { hy[1] = -hy[1]; hz[1] = -hz[1]; cross3(hy, hz, hx); }     // rflmfx: x-hat = y x z
            else { hx[1] = -hx[1]; hy[1] = -hy[1]; cross3(hx, hy, hz); }                 // rflmfz: z-hat = x x y
            double mm[3][3], theta;
            #pragma unroll
            for (int d = 0; d < 3; d++) { mm[d][0] = hx[d]; mm[d][1] = hy[d]; mm[d][2] = hz[d]; }
            euler_from_matrix(mm, phi, theta, chi);
            cth = cos(theta);
         }
      if (op[3] == m) {
         if (top) {
            chi = chi + 2.0 * PI / (double)p.nfold;
            chi = fmod(chi, 2.0 * PI);
            if (chi < 0.0) chi = 2.0 * PI + chi;
         } else {
            phi = phi + PI;
            cth *= -1.0;
            phi = fmod(phi, 2.0 * PI);
            if (phi < 0.0) phi = 2.0 * PI + phi;
            const double sint = sqrt(1.0 - cth * cth);
            p.cosn[ang_index(p, c, q, 0, m)] = sint * cos(phi);
            p.cosn[ang_index(p, c, q, 1, m)] = sint * sin(phi);
            p.cosn[ang_index(p, c, q, 2, m)] = cth;
         }
      }
      p.ang[ang_index(p, c, q, 0, m)] = phi; p.ang[ang_index(p, c, q, 1, m)] = cth; p.ang[ang_index(p, c, q, 2, m)] = chi;
   }
   if ((op[0] + op[1] + op[2]) & 1)
      for (long i = threadIdx.x; i < (long)p.P * p.Npad; i += blockDim.x) {
         const int it = (int)(i / p.Npad), a = (int)(i % p.Npad);
         p.pos[pos_index(p, c, it, 1, a)] *= -1.0;
      }
}

// one MCWormMove per chain outside the step kernel (parity entry point; one CTA per chain)
template <int KIND>
__global__ void __launch_bounds__(256) worm_move_kernel(const __grid_constant__ Params p)
{
   extern __shared__ double wsm[];
   SmallTables t;
   t.g1d = p.g1d; t.v1d = p.v1d; t.y2_1d = p.y2_1d; t.lut1d = p.lut1d; t.rgrid = p.rgrid; t.rdens = p.rdens; t.rdens2 = p.rdens2;
   t.lutrot = p.lutrot; t.rec1d = p.rec1d; t.recrot = p.recrot; t.rgi2d = p.rgi2d; t.cgi2d = p.cgi2d;
   worm_sweep_cta<KIND>(p, t, blockIdx.x, wsm, reinterpret_cast<unsigned char *>(wsm + 64));
}

// ---- parity kernels -----------------------------------------------------------------------------
__global__ void eval_spot1d_kernel(const __grid_constant__ Params p, int n, const double *r, double *v, int *klo)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   SmallTables t; t.g1d = p.g1d; t.v1d = p.v1d; t.y2_1d = p.y2_1d; t.lut1d = p.lut1d; t.rec1d = p.rec1d; t.recrot = p.recrot; t.rgi2d = p.rgi2d; t.cgi2d = p.cgi2d;
   int k; v[i] = spot1d(p, t, r[i], &k); klo[i] = k;
}
__global__ void eval_lpot2d_kernel(const __grid_constant__ Params p, int n, const double *r, const double *c, double *v, int *ir, int *ic)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   SmallTables t; t.rgi2d = p.rgi2d; t.cgi2d = p.cgi2d;
   int a, b; v[i] = lpot2d(p, t, r[i], c[i], &a, &b); ir[i] = a; ic[i] = b;
}
__global__ void eval_srot_kernel(const __grid_constant__ Params p, int n, const double *g, int which, double *v)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   const double *y = which == 0 ? p.rdens : which == 1 ? p.rderv : p.resqr;
   const double *y2 = which == 0 ? p.rdens2 : which == 1 ? p.rderv2 : p.resqr2;
   v[i] = srot_eval(p, p.rgrid, y, y2, p.lutrot, g[i], which);
}
__global__ void eval_rotden_kernel(const __grid_constant__ Params p, int n, const double *e1, const double *e2, double *rho, double *erot, double *esq, int *idx)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   Mat3 a, b;
   matpre(e1[3 * i], e1[3 * i + 1], e1[3 * i + 2], a);
   matpre(e2[3 * i], e2[3 * i + 1], e2[3 * i + 2], b);
   double er, es; int k, istop = 0;
   rho[i] = rotden(p, a, b, nullptr, &er, &es, &k, &istop);
   erot[i] = er; esq[i] = es; idx[i] = istop ? -1 - k : k;
}
__global__ void eval_vcord_kernel(const __grid_constant__ Params p, int n, const double *eul, const double *rcom, const double *rpt, double *v, double *rtc, int *idx)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   Mat3 a;
   matpre(eul[3 * i], eul[3 * i + 1], eul[3 * i + 2], a);
   int k; v[i] = vcord(p, a, rcom + 3 * i, rpt + 3 * i, rtc + 3 * i, &k); idx[i] = k;
}
__global__ void eval_rotpro_kernel(const __grid_constant__ Params p, int n, const double *deg, double *rho, double *erot, double *esq, int *idx)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   double er, es; int k, istop = 0;
   rho[i] = rotpro(p, deg[3 * i], deg[3 * i + 1], deg[3 * i + 2], &er, &es, &k, &istop);
   erot[i] = er; esq[i] = es; idx[i] = istop ? -1 - k : k;
}
__global__ void eval_vcalc_kernel(const __grid_constant__ Params p, int n, const double *rtc, double *v, int *idx)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   int k; v[i] = vcalc(p, rtc[3 * i], rtc[3 * i + 1], rtc[3 * i + 2], &k); idx[i] = k;
}
__global__ void eval_deleul_kernel(int n, const double *e1, const double *e2, double *rel)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   Mat3 a, b;
   matpre(e1[3 * i], e1[3 * i + 1], e1[3 * i + 2], a);
   matpre(e2[3 * i], e2[3 * i + 1], e2[3 * i + 2], b);
   deleul(a, b, rel[3 * i], rel[3 * i + 1], rel[3 * i + 2]);
}
__global__ void eval_vcord_grid_kernel(const __grid_constant__ Params p, int n, const double *eul, const double *rcom, const double *rpt, double *grid)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   Mat3 a;
   matpre(eul[3 * i], eul[3 * i + 1], eul[3 * i + 2], a);
   vcord(p, a, rcom + 3 * i, rpt + 3 * i, nullptr, nullptr, grid + 3 * i);
}
__global__ void eval_vspher_kernel(const __grid_constant__ Params p, int n, const double *r, double *v, double *rc)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   double c; v[i] = vspher(p, r[i], &c); rc[i] = c;
}
__global__ void eval_libm_kernel(int which, int n, const double *x, double *y)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   const double a = x[i];
   double r = 0.0;
   switch (which) {
      case 0: r = sin(a); break;
      case 1: r = cos(a); break;
      case 2: r = acos(a); break;
      case 3: r = atan(a); break;
      case 4: r = exp(a); break;
      case 5: r = log(a); break;
      case 6: r = sqrt(a); break;
      case 7: r = fmod(a, 2.0 * PI); break;
   }
   y[i] = r;
}
__global__ void eval_caleng_kernel(int n, const double *c1, const double *c2, const double *e1, const double *e2, double *e)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   Mat3 a, b;
   matpre(e1[3 * i], e1[3 * i + 1], e1[3 * i + 2], a);
   matpre(e2[3 * i], e2[3 * i + 1], e2[3 * i + 2], b);
   Tip4pSites sa, sb;
   tip4p_sites(a, c1 + 3 * i, sa);
   tip4p_sites(b, c2 + 3 * i, sb);
   e[i] = caleng(sa, sb);
}
// PotEnergy(atom, MCCoords, it) for all (atom, it) of chain c: one warp per bead, lanes over partners
__global__ void pot_energy_slice_kernel(const __grid_constant__ Params p, int c, double *v)
{
   SmallTables t; t.g1d = p.g1d; t.v1d = p.v1d; t.y2_1d = p.y2_1d; t.lut1d = p.lut1d;
   t.rgrid = p.rgrid; t.rdens = p.rdens; t.rdens2 = p.rdens2; t.lutrot = p.lutrot; t.rec1d = p.rec1d; t.recrot = p.recrot; t.rgi2d = p.rgi2d; t.cgi2d = p.cgi2d;
   int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   if (w >= p.N * p.P) return;
   int atom = w / p.P, it = w % p.P;
   double p0[3];
   for (int d = 0; d < 3; d++) p0[d] = p.pos[pos_index(p, c, it, d, atom)];
   double s = 0.0;
   for (int j = lane; j < p.N; j += 32)
      if (j != atom && partner_on_line(p, c, j, it)) s += pair_energy(p, t, c, atom, p0, j, it, nullptr, nullptr);
   for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
   if (lane == 0) v[(size_t)atom * p.P + it] = s;
}
__global__ void rng_draws_kernel(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, uint32_t s4, uint32_t s5, int n, double *out)
{
   Mrg g; g.s[0] = s0; g.s[1] = s1; g.s[2] = s2; g.s[3] = s3; g.s[4] = s4; g.s[5] = s5;
   for (int i = 0; i < n; i++) out[i] = mrg_u01(g);
}

// FP64 FMA throughput microbenchmark (the denominator of the FP64 roofline)
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b)
{
   double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
   for (int i = 0; i < iters; i++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
   }
   double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
   if (s == 12345.678) out[0] = s;
}

} // namespace pimc
