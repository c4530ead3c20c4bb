// Persistent move kernel: translational multilevel bisection, whole-path (molecular) moves and
// rotational checkerboard sweeps for many independent Markov chains (sm_100a, FP64).
//
// Execution model: one thread-block CLUSTER per chain (cluster size 1..16).  The kernel stays
// resident for `nsteps` iterations of the reference's `time` loop (mc_main.cc:349-381) and walks
// the stages of one step in order; dependent stages are separated by a chain-wide barrier
// (__syncthreads for a 1-CTA chain, barrier.cluster otherwise).  Inside a stage the independent
// units (bisection segments, rotational slices of one parity) are dealt to TEAMS of `team` lanes
// (power of two, <= 32); the lanes of a team split the partner loop of the pair-action sum and
// combine with a fixed-order warp-shuffle butterfly, so results are bit-reproducible.
//
// Schedule (DESIGN.md "Schedule"; the CPU replay is oracle/pimc_oracle.cpp:orc_sched_run):
//   step t, time = t mod P, for each type:
//     time == 0                  -> whole-path move of every permutation cycle     (MCMolecularMove*)
//     time mod (P/seg) == 0      -> bisection sweep with offset time/(P/seg): all P/seg disjoint
//                                   segments of every atom, atoms in sequence       (MCBisectionMove*)
//     rotor type                 -> even slices, then odd slices                    (MCRotations3D / MCRotationsMove)
#pragma once
#include "pimc_device.cuh"
#include <cooperative_groups.h>

namespace pimc {
namespace cg = cooperative_groups;

struct Ctx {
   int c, crank;
   int tid, gthread, nthreads_chain;
   int T, lane_t, team_lane0, team_id, nteams_chain;
   SmallTables t;
   double *team_buf;   // shared: (seg_max+1)*3 doubles per team
   double *red;        // shared: 40 doubles
};

__device__ __forceinline__ void chain_sync(const Params &p)
{
   if (p.cpc == 1) __syncthreads();
   else cg::this_cluster().sync();
}
__device__ __forceinline__ double team_sum(double v, int T)
{
   for (int o = T >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}
__device__ __forceinline__ double team_prod(double v, int T)
{
   for (int o = T >> 1; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}
__device__ __forceinline__ int team_or(int v, int T)
{
   for (int o = T >> 1; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}
__device__ __forceinline__ uint32_t *stream_ptr(const Params &p, int c, int s)
{
   return p.rng + ((size_t)c * p.S + s) * 6;
}
__device__ __forceinline__ double *counter_ptr(const Params &p, int c, int type, int move)
{
   return p.counters + (((size_t)c * MAXT + type) * 3 + move) * 2;
}

// deterministic chain-wide sum: warp butterfly, per-CTA shared slots, per-chain global slots
__device__ double chain_reduce(const Params &p, Ctx &x, double v)
{
   v = team_sum(v, 32);
   int warp = x.tid >> 5, nwarp = blockDim.x >> 5;
   if ((x.tid & 31) == 0) x.red[warp] = v;
   __syncthreads();
   double s = 0.0;
   for (int w = 0; w < nwarp; w++) s += x.red[w];
   __syncthreads();
   if (p.cpc == 1) return s;
   if (x.tid == 0) p.scratch[(size_t)x.c * 64 + x.crank] = s;
   cg::this_cluster().sync();
   double tot = 0.0;
   for (int r = 0; r < p.cpc; r++) tot += __ldcg(p.scratch + (size_t)x.c * 64 + r);
   return tot;
}

// ---------------------------------------------------------------------------------------------
// whole-path move of every permutation cycle of `type` (MCMolecularMove / MCMolecularMoveExchange,
// mc_piqmc.cc:54-192).  dV of the rigid shift: cycle members against non-members, all P slices.
// ---------------------------------------------------------------------------------------------
__device__ void molecular_sweep(const Params &p, Ctx &x, int type)
{
   const int c = x.c, P = p.P, N = p.N;
   const int *cyc_start = p.cyc_start + (size_t)c * (N + 1);
   const int *cyc_atoms = p.cyc_atoms + (size_t)c * N;
   int g0 = 0;
   for (int t = 0; t < type; t++) g0 += p.ncyc[c * MAXT + t];
   int g1 = g0 + p.ncyc[c * MAXT + type];
   uint32_t *ms = stream_ptr(p, c, P + p.Q);
   for (int g = g0; g < g1; g++) {
      Mrg rs;
      mrg_load(rs, ms);
      double u0 = mrg_u01(rs), u1 = mrg_u01(rs), u2 = mrg_u01(rs), u3 = mrg_u01(rs);
      double disp[3] = {p.mcstep[type] * (u0 - 0.5), p.mcstep[type] * (u1 - 0.5), p.mcstep[type] * (u2 - 0.5)};
      int b = cyc_start[g], len = cyc_start[g + 1] - b;
      long nitems = (long)len * P * N;
      double part = 0.0;
      for (long i = x.gthread; i < nitems; i += x.nthreads_chain) {
         int j = (int)(i % N);
         long r = i / N;
         int it = (int)(r % P);
         int a0 = cyc_atoms[b + (int)(r / P)];
         bool member = false;
         for (int k = 0; k < len; k++) member |= (cyc_atoms[b + k] == j);
         if (member) continue;
         double po[3], pn[3];
         #pragma unroll
         for (int d = 0; d < 3; d++) { po[d] = p.pos[pos_index(p, c, it, d, a0)]; pn[d] = po[d] + disp[d]; }
         part += pair_energy(p, x.t, c, a0, pn, j, it, nullptr, nullptr) - pair_energy(p, x.t, c, a0, po, j, it, nullptr, nullptr);
      }
      double deltav = chain_reduce(p, x, part);
      bool acc = (deltav < 0.0) || (exp(-deltav * p.tau) > u3);
      if (acc) {
         long nw = (long)len * P * 3;
         for (long i = x.gthread; i < nw; i += x.nthreads_chain) {
            int d = (int)(i % 3);
            long r = i / 3;
            int it = (int)(r % P);
            int a0 = cyc_atoms[b + (int)(r / P)];
            p.pos[pos_index(p, c, it, d, a0)] += disp[d];
         }
      }
      if (x.gthread == 0) {
         mrg_store(rs, ms);
         double *cn = counter_ptr(p, c, type, 0);
         cn[0] += 1.0;
         if (acc) cn[1] += 1.0;
      }
      chain_sync(p);
   }
}

// ---------------------------------------------------------------------------------------------
// bisection sweep (MCBisectionMove / MCBisectionMoveExchange, mc_piqmc.cc:194-419): every atom of
// `type` in sequence; for one atom all P/seg segments [s0, s0+seg], s0 = off + k*seg, in parallel,
// one team per segment.  Level sums are kept incrementally: pot0(l) = S(l-1) + D(l) where D(l) is
// the dV of the midpoints sampled at level l, so delta = (D - S)*tau*seg_l/2 equals the
// reference's (pot0 - 2 pot1)*tau*(seg_l/2) without re-evaluating earlier levels (:263 "inefficient").
// ---------------------------------------------------------------------------------------------
__device__ void bisection_sweep(const Params &p, Ctx &x, int type, int off)
{
   const int c = x.c, P = p.P, N = p.N, T = x.T;
   const int L = p.levels[type], seg = 1 << L, nseg = P / seg;
   const int base = p.first[type];
   const double bnorm = 1.0 / (p.lambda[type] * p.tau);
   double *nx = x.team_buf;
   const int nrounds = (nseg + x.nteams_chain - 1) / x.nteams_chain;
   for (int a = 0; a < p.numb[type]; a++) {
      const int gA = base + a;
      const int gB = (p.stat[type] == 1) ? p.pindex[(size_t)c * N + gA] : gA;
      for (int rd = 0; rd < nrounds; rd++) {
         const int k = rd * x.nteams_chain + x.team_id;
         const bool active = k < nseg;
         const int s0 = active ? (off + k * seg) % P : 0;
         if (active)
            for (int i = x.lane_t; i < 6; i += T) {
               int e = i / 3, d = i - 3 * e, t = e * seg;
               int g = (s0 + t >= P) ? gB : gA;
               nx[t * 3 + d] = p.pos[pos_index(p, c, (s0 + t) % P, d, g)];
            }
         __syncwarp();
         double S = 0.0;
         bool alive = active;
         for (int level = 0; level < L; level++) {
            const int lss = seg >> level, half = lss >> 1, nmid = 1 << level;
            if (alive) {
               const double bkin = bnorm / (double)lss;
               for (int m = x.lane_t; m < nmid; m += T) {
                  int t1 = half + m * lss;
                  int sl = (s0 + t1) % P;
                  uint32_t *sp = stream_ptr(p, c, sl);
                  Mrg rs;
                  mrg_load(rs, sp);
                  #pragma unroll
                  for (int d = 0; d < 3; d++) {
                     double r1 = mrg_u01(rs), r2 = mrg_u01(rs);
                     nx[t1 * 3 + d] = 0.5 * (nx[(t1 - half) * 3 + d] + nx[(t1 + half) * 3 + d]) + gauss_u(bkin, r1, r2);
                  }
                  mrg_store(rs, sp);
               }
            }
            __syncwarp();
            double D = 0.0;
            if (alive) {
               const int nitems = nmid * N;
               for (int i = x.lane_t; i < nitems; i += T) {
                  int m = i / N, j = i - m * N;
                  int t1 = half + m * lss;
                  int sl = (s0 + t1) % P;
                  int g = (s0 + t1 >= P) ? gB : gA;
                  if (j == g) continue;
                  double po[3], pn[3];
                  #pragma unroll
                  for (int d = 0; d < 3; d++) { po[d] = p.pos[pos_index(p, c, sl, d, g)]; pn[d] = nx[t1 * 3 + d]; }
                  D += pair_energy(p, x.t, c, g, pn, j, sl, nullptr, nullptr) - pair_energy(p, x.t, c, g, po, j, sl, nullptr, nullptr);
               }
            }
            D = team_sum(D, T);
            const double deltav = (D - S) * (p.tau * (double)half);
            S += D;
            int acc = 1;
            if (alive && x.lane_t == 0) {
               if (!(deltav < 0.0)) {
                  uint32_t *sp = stream_ptr(p, c, s0);
                  Mrg rs;
                  mrg_load(rs, sp);
                  double u = mrg_u01(rs);
                  mrg_store(rs, sp);
                  acc = (exp(-deltav) > u) ? 1 : 0;
               }
            }
            acc = __shfl_sync(0xffffffffu, acc, x.team_lane0);
            if (!acc) alive = false;
         }
         if (active && x.lane_t == 0) {
            double *cn = counter_ptr(p, c, type, 1);
            atomicAdd(cn, 1.0);
            if (alive) atomicAdd(cn + 1, 1.0);
         }
         if (alive)
            for (int i = x.lane_t; i < (seg - 1) * 3; i += T) {
               int t = 1 + i / 3, d = i % 3;
               int g = (s0 + t >= P) ? gB : gA;
               p.pos[pos_index(p, c, (s0 + t) % P, d, g)] = nx[t * 3 + d];
            }
         __syncwarp();
      }
      chain_sync(p);
   }
}

// ---------------------------------------------------------------------------------------------
// one rotational Metropolis step at rot slice q for rotor m of `type`
// (MCRot3Dstep mc_piqmc.cc:938-1199, MCRotLinStep :781-936; RotDenType 0)
// ---------------------------------------------------------------------------------------------
__device__ void rot_step(const Params &p, Ctx &x, int type, int q, int m, bool active, int *err)
{
   const int c = x.c, N = p.N, Q = p.Q, R = p.R, T = x.T;
   const int g = p.first[type] + m;
   const bool top = p.molecule[type] == 2;
   double r1 = 0, r2 = 0, r3 = 0, r4 = 0;
   if (active && x.lane_t == 0) {
      uint32_t *sp = stream_ptr(p, c, p.P + q);
      Mrg rs;
      mrg_load(rs, sp);
      r1 = mrg_u01(rs); r2 = mrg_u01(rs); r3 = mrg_u01(rs);
      if (top) r4 = mrg_u01(rs);
      mrg_store(rs, sp);
   }
   r1 = __shfl_sync(0xffffffffu, r1, x.team_lane0);
   r2 = __shfl_sync(0xffffffffu, r2, x.team_lane0);
   r3 = __shfl_sync(0xffffffffu, r3, x.team_lane0);
   r4 = __shfl_sync(0xffffffffu, r4, x.team_lane0);
   if (!top) r4 = r3;            // the linear step uses its third uniform for the accept test

   int q0 = q - 1, q2 = q + 1;
   if (q0 < 0) q0 += Q;
   if (q2 >= Q) q2 -= Q;
   const double step = p.rtstep[type];
   double dens_old = 1.0, dens_new = 1.0, dV = 0.0;
   double cost = 0, phi = 0, chi = 0, nn[3] = {0, 0, 1};
   int bad = 0;
   if (active) {
      cost = p.ang[ang_index(p, c, q, 1, m)];
      phi = p.ang[ang_index(p, c, q, 0, m)];
      chi = p.ang[ang_index(p, c, q, 2, m)];
      const double cost_old = cost, phi_old = phi, chi_old = chi;
      cost += step * (r1 - 0.5);
      if (top) {
         phi += 2.0 * PI * (step * (r2 - 0.5));
         chi += 2.0 * PI * (step * (r3 - 0.5));
         if (phi < 0.0) phi = 2.0 * PI + phi;
         if (chi < 0.0) chi = 2.0 * PI + chi;
         phi = fmod(phi, 2.0 * PI);
         chi = fmod(chi, 2.0 * PI);
      } else {
         phi += step * (r2 - 0.5);
      }
      if (cost > 1.0) cost = 2.0 - cost;
      if (cost < -1.0) cost = -2.0 - cost;
      const int it0 = q * R;
      const int nitems = R * N;
      if (top) {
         Mat3 R0, R1, R2, Rn;
         load_rotmat(p, c, q0, m, R0);
         matpre(phi_old, acos(cost_old), chi_old, R1);
         load_rotmat(p, c, q2, m, R2);
         matpre(phi, acos(cost), chi, Rn);
         double po = 1.0, pn = 1.0;
         for (int i = x.lane_t; i < 4; i += T) {
            int istop = 0;
            if (i == 0) po *= rotden(p, R0, R1, nullptr, nullptr, nullptr, nullptr, &istop);
            else if (i == 1) po *= rotden(p, R1, R2, nullptr, nullptr, nullptr, nullptr, &istop);
            else if (i == 2) pn *= rotden(p, R0, Rn, nullptr, nullptr, nullptr, nullptr, &istop);
            else pn *= rotden(p, Rn, R2, nullptr, nullptr, nullptr, nullptr, &istop);
            bad |= istop;
         }
         dens_old = po; dens_new = pn;
         for (int i = x.lane_t; i < nitems; i += T) {
            int r = i / N, j = i - r * N;
            if (j == g) continue;
            int it = it0 + r;
            double pg[3];
            #pragma unroll
            for (int d = 0; d < 3; d++) pg[d] = p.pos[pos_index(p, c, it, d, g)];
            dV += pair_energy(p, x.t, c, g, pg, j, it, &Rn, nullptr) - pair_energy(p, x.t, c, g, pg, j, it, &R1, nullptr);
         }
      } else {
         const double sint = sqrt(1.0 - cost * cost);
         double sp_, cp_;
         sincos(phi, &sp_, &cp_);
         nn[0] = sint * cp_; nn[1] = sint * sp_; nn[2] = cost;
         double n0[3], n1[3], n2[3];
         #pragma unroll
         for (int d = 0; d < 3; d++) {
            n0[d] = p.cosn[ang_index(p, c, q0, d, m)];
            n1[d] = p.cosn[ang_index(p, c, q, d, m)];
            n2[d] = p.cosn[ang_index(p, c, q2, d, m)];
         }
         double po = 1.0, pn = 1.0;
         for (int i = x.lane_t; i < 4; i += T) {
            const double *a = (i == 0 || i == 2) ? n0 : (i == 1 ? n1 : nn);
            const double *b = (i == 0) ? n1 : (i == 2 ? nn : n2);
            double dot = 0.0;
            #pragma unroll
            for (int d = 0; d < 3; d++) dot += a[d] * b[d];
            double rho = srotdens(p, x.t, dot);
            if (i < 2) po *= rho; else pn *= rho;
         }
         dens_old = po; dens_new = pn;
         for (int i = x.lane_t; i < nitems; i += T) {
            int r = i / N, j = i - r * N;
            if (j == g) continue;
            int it = it0 + r;
            double pg[3];
            #pragma unroll
            for (int d = 0; d < 3; d++) pg[d] = p.pos[pos_index(p, c, it, d, g)];
            dV += pair_energy(p, x.t, c, g, pg, j, it, nullptr, nn) - pair_energy(p, x.t, c, g, pg, j, it, nullptr, n1);
         }
      }
   }
   dens_old = team_prod(dens_old, T);
   dens_new = team_prod(dens_new, T);
   dV = team_sum(dV, T);
   bad = team_or(bad, T);
   if (active && x.lane_t == 0) {
      if (fabs(dens_old) < RZERO) dens_old = 0.0;
      if (fabs(dens_new) < RZERO) dens_new = 0.0;
      if (top) { dens_old = fabs(dens_old); dens_new = fabs(dens_new); }
      else if (dens_old < 0.0 || dens_new < 0.0) bad = 2;     // "Negative rot density" is fatal in the reference
      double rd = (dens_old > RZERO) ? dens_new / dens_old : 1.0;
      rd *= exp(-p.tau * dV);
      bool acc = (rd > 1.0) || (rd > r4);
      if (bad) { acc = false; atomicOr(err, bad); }
      double *cn = counter_ptr(p, c, type, 2);
      atomicAdd(cn, 1.0);
      if (acc) {
         atomicAdd(cn + 1, 1.0);
         p.ang[ang_index(p, c, q, 1, m)] = cost;
         p.ang[ang_index(p, c, q, 0, m)] = phi;
         if (top) {
            p.ang[ang_index(p, c, q, 2, m)] = chi;
            const double sint = sqrt(1.0 - cost * cost);
            double sp_, cp_;
            sincos(phi, &sp_, &cp_);
            nn[0] = sint * cp_; nn[1] = sint * sp_; nn[2] = cost;
         }
         #pragma unroll
         for (int d = 0; d < 3; d++) p.cosn[ang_index(p, c, q, d, m)] = nn[d];
      }
   }
   __syncwarp();
}

// even slices, then odd slices (MCRotations3D mc_piqmc.cc:737-771, MCRotationsMove :490-512);
// an odd slice count gets a third phase for the last slice.
__device__ void rot_sweep(const Params &p, Ctx &x, int type, int *err)
{
   const int Q = p.Q;
   const int qe = (Q % 2 == 1 && Q > 1) ? Q - 1 : Q;
   for (int phase = 0; phase < 3; phase++) {
      int count = (phase == 0) ? (qe + 1) / 2 : (phase == 1 ? qe / 2 : (qe != Q ? 1 : 0));
      if (count == 0) continue;
      int nrounds = (count + x.nteams_chain - 1) / x.nteams_chain;
      for (int rd = 0; rd < nrounds; rd++) {
         int idx = rd * x.nteams_chain + x.team_id;
         bool active = idx < count;
         int q = (phase == 0) ? 2 * idx : (phase == 1 ? 2 * idx + 1 : Q - 1);
         if (!active) q = 0;
         for (int m = 0; m < p.numb[type]; m++) rot_step(p, x, type, q, m, active, err);
      }
      chain_sync(p);
   }
}

// ---------------------------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_tables(const Params &p, SmallTables &t, double *&cursor)
{
   // copies the small spline tables into shared memory; `cursor` walks the dynamic smem block
   auto put = [&](const double *src, int n) -> const double * {
      double *dst = cursor;
      for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
      cursor += (n + 1) & ~1;
      return dst;
   };
   auto puti = [&](const int *src, int n) -> const int * {
      int *dst = reinterpret_cast<int *>(cursor);
      for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
      cursor += ((n + 1) / 2 + 1) & ~1;
      return dst;
   };
   t.g1d = t.v1d = t.y2_1d = nullptr; t.lut1d = nullptr;
   t.rgrid = t.rdens = t.rdens2 = nullptr; t.lutrot = nullptr;
   if (p.n1d) { t.g1d = put(p.g1d, p.n1d); t.v1d = put(p.v1d, p.n1d); t.y2_1d = put(p.y2_1d, p.n1d); t.lut1d = puti(p.lut1d, p.nlut1d); }
   if (p.nrot) { t.rgrid = put(p.rgrid, p.nrot); t.rdens = put(p.rdens, p.nrot); t.rdens2 = put(p.rdens2, p.nrot); t.lutrot = puti(p.lutrot, p.nlutrot); }
   __syncthreads();
}

__global__ void __launch_bounds__(512, 1)
pimc_steps_kernel(const __grid_constant__ Params p, long t0, long nsteps, int *err)
{
   extern __shared__ double smem[];
   Ctx x;
   x.tid = threadIdx.x;
   x.c = blockIdx.x / p.cpc;
   x.crank = blockIdx.x % p.cpc;
   x.gthread = x.crank * blockDim.x + x.tid;
   x.nthreads_chain = p.cpc * blockDim.x;
   x.T = p.team;
   x.lane_t = x.tid & (x.T - 1);
   x.team_lane0 = (x.tid & 31) & ~(x.T - 1);
   x.team_id = x.gthread / x.T;
   x.nteams_chain = x.nthreads_chain / x.T;
   double *cursor = smem;
   x.red = cursor; cursor += 40;
   stage_tables(p, x.t, cursor);
   x.team_buf = cursor + (size_t)(x.tid / x.T) * ((p.seg_max + 1) * 3);

   for (long s = 0; s < nsteps; s++) {
      const long t = t0 + s;
      const int time = (int)(t % p.P);
      for (int type = 0; type < p.ntypes; type++) {
         if (time == 0) molecular_sweep(p, x, type);
         const int seg = 1 << p.levels[type], nseg = p.P / seg;
         if (time % nseg == 0) bisection_sweep(p, x, type, (time / nseg) % p.P);
         if (type == p.imtype && p.Q > 0) rot_sweep(p, x, type, err);
      }
   }
}

} // namespace pimc
