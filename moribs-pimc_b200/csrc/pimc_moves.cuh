// Persistent move kernel: translational multilevel bisection, whole-path (molecular) moves and
// rotational checkerboard sweeps for many independent Markov chains (sm_100a, FP64).
//
// Execution model: `cpc` CTAs per chain (1 CTA per SM).  Up to 8 CTAs form one thread-block CLUSTER and use the
// hardware cluster barrier; 16 CTAs per chain (8 chains x 16 = 128 of the 148 SMs -- only seven 16-CTA clusters fit
// on a B200 at once, measured) run as a cooperative grid with a per-chain barrier in global memory.  The kernel
// stays resident for `nsteps` iterations of the reference's `time` loop (mc_main.cc:349-381) and walks the stages of
// one step in order; dependent stages are separated by the chain barrier.  Inside a stage the independent units are
// dealt to groups of threads:
//   * bisection segments -> TEAMS of `team` threads (power of two, 1..128; more than 32 = several warps of one CTA):
//     the lanes split the (midpoint, partner) terms of the pair-action sum; warp butterfly + fixed-order combine;
//   * rotational slices -> ROT GROUPS of `rot_group` threads (power of two, may span several warps of one CTA).
// Every reduction has a fixed order, so a trajectory is bit-reproducible for a given geometry.
// The kernel is specialised at compile time on the rotor kind (KIND & 3: 0 atoms only, 1 linear rotor,
// 2 non-linear top) so each variant carries only its own interaction branches, and on the worm (KIND & 4):
// the worm moves and the world-line masks of the pair loops exist only in the variants that need them.
//
// Schedule (DESIGN.md "Schedule"; the CPU replay is oracle/pimc_oracle.cpp:orc_sched_run):
//   step t, time = t mod P, for each type:
//     time == 0                  -> whole-path move of every permutation cycle     (MCMolecularMove*)
//     time mod (P/seg) == 0      -> bisection sweep with offset time/(P/seg): all P/seg disjoint
//                                   segments of every atom, atoms in sequence       (MCBisectionMove*)
//     rotor type                 -> even slices, then odd slices                    (MCRotations3D / MCRotationsMove)
// A system with ONE rotor evaluates the potential part of all Q proposals in a single parallel stage (the pair action
// is diagonal in imaginary time and the proposal of a slice depends on that slice alone); the even and the odd
// accept/reject decisions, which couple neighbouring slices only through the density matrix, are pipelined across
// time steps behind split arrive/wait counters (rot_sweep_pipe).
#pragma once
#include "pimc_device.cuh"
#include "pimc_worm.cuh"
#include <cooperative_groups.h>

namespace pimc {
namespace cg = cooperative_groups;

#ifdef PIMC_TIMELINE
#define MARK(x, id) do { if ((x).gthread == 0 && (x).c == 0 && g_nmarks < 4000) { g_marks[g_nmarks++] = ((long long)(id) << 48) | (clock64() & 0xffffffffffffLL); } } while (0)
#else
#define MARK(x, id) do { } while (0)
#endif

constexpr int MAXLEV = 8;         // bisection levels supported by the per-team scratch
constexpr int TEAM_EXTRA = 4 + MAXLEV + 3 * MAXLEV;   // doubles per team behind the segment buffers: per-warp partial sums, candidate accept uniforms and their stream states

// per rot group scratch in shared memory
struct RotSlot {
   double u4;                 // uniform of the accept test
   double cost, phi, chi;     // proposal
   double a[9], b[9];         // KIND 2: rotation matrices of the new (a) and current (b) orientation; KIND 1: a[0..2] = n_new, b[0..2] = n_cur
   double rho[4];             // the four density-matrix factors
   double vnew, vold;         // potential sums of the proposed / current orientation (pipelined sweep)
   double cur[3];             // pipelined sweep: current (cost, phi, chi) of the owned slice, kept across time steps
   double vcache;             // pipelined sweep: cached potential sum of the current orientation ...
   unsigned nbw[12];          // rot_run: the two neighbours' axes as received (2 x 3 doubles in 32-bit halves)
   int vep, epoch;            // ... the position epoch it belongs to, and the epoch seen by the current sweep
   int need_old, bad;
   int gep;                   // position epoch the geometry cache of this slice was filled at (-1: never)
};

struct Ctx {
   int c, crank;
   int tid, gthread, nthreads_chain;
   int T, lane_t, team_lane0, team_id, nteams_chain, team_cta;
   int W, tw, wl, wlanes;     // warps per team, this thread's warp inside the team, lane inside that warp, lanes per warp part
   int G, gl, grp, ngrp;      // rot group size, index inside the group, group index in the CTA, groups per CTA
   unsigned gmask;            // lanes of this thread's rot group inside its warp (all lanes when the group fills or spans warps)
   SmallTables t;
   double *team_buf;   // shared: this thread's team scratch (segment positions, unit normals, stream cache, partial sums)
   volatile int *rot_done;   // shared: rot_run_cta, sweeps completed by every rot slice of the chain in this launch
   double *red;        // shared: 40 doubles
   RotSlot *slot;      // shared: this thread's rot group slot (thread-private storage when the group is one thread)
   double *part;       // shared: per-warp partial sums (new, old) of this thread's rot group, used when it spans warps
   uint32_t *rrng;     // shared: MRG32k3a states of the rot slices this CTA owns in the fused rotational sweep
   unsigned bar_target;
   int red_par;
   int rot_iter;       // pipelined rotational sweeps done in this launch
   unsigned mol_target[2];   // arrivals expected on the two whole-path hand-over counters (molecular_sweep_piped)
};

// doubles of scratch per bisection team (host and device use the same formula)
__host__ __device__ inline int team_buf_doubles(int seg_max) { return (seg_max + 1) * 12 + seg_max * 3 + TEAM_EXTRA; }   // two (positions, normals) sets: pipelined sweep

__device__ __forceinline__ void chain_sync(const Params &p, Ctx &x)
{
   if (p.cpc == 1) { __syncthreads(); return; }
   if (!p.swbar) { cg::this_cluster().sync(); return; }
   // cooperative grid: arrive/spin on the chain's counter (the pattern of a grid-wide barrier)
   __syncthreads();
   if (x.tid == 0) {
      x.bar_target += (unsigned)p.cpc;
      unsigned *b = p.barrier + (size_t)x.c * 32;
      __threadfence();
      atomicAdd(b, 1u);
      unsigned v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(b) : "memory"); } while ((int)(v - x.bar_target) < 0);
      __threadfence();
   }
   __syncthreads();
}
__device__ __forceinline__ double team_sum(double v, int T)
{
   for (int o = T >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
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

// deterministic chain-wide sum: warp butterfly, per-CTA shared slots, per-chain global slots (two alternating sets)
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
   x.red_par ^= 1;
   double *slots = p.scratch + (size_t)x.c * 64 + 32 * x.red_par;
   if (x.tid == 0) slots[x.crank] = s;
   chain_sync(p, x);
   double tot = 0.0;
   for (int r = 0; r < p.cpc; r++) tot += __ldcg(slots + r);
   return tot;
}

// V(bead at pn) - V(bead at po) of atom g against partner j at slice it; for atom-atom pairs of a spline system the
// partner position is loaded once
template <int KIND>
__device__ __forceinline__ double pair_diff(const Params &p, const SmallTables &t, int c, int g, const double *pn, const double *po, int j, int it)
{
   if ((KIND & 3) != 2 && p.mode[type_of(p, g)][type_of(p, j)] == M_SPOT1D && !p.minimage) {
      double d2n = 0.0, d2o = 0.0;
      #pragma unroll
      for (int d = 0; d < 3; d++) {
         double pj = p.pos[pos_index(p, c, it, d, j)];
         d2n += (pn[d] - pj) * (pn[d] - pj);
         d2o += (po[d] - pj) * (po[d] - pj);
      }
      return spot1d_move(p, t, sqrt(d2n)) - spot1d_move(p, t, sqrt(d2o));
   }
   if ((KIND & 3) == 1 && p.mode[type_of(p, g)][type_of(p, j)] == M_LIN_1MOL && !p.minimage) {
      // atom bead (new and old position) against the linear rotor: shared partner loads, both bilinear forms in flight
      const int q = it / p.R, m = j - p.first[p.imtype];
      double rr[2], cs[2], e[2], d2n = 0.0, d2o = 0.0, dn = 0.0, dold = 0.0;
      #pragma unroll
      for (int d = 0; d < 3; d++) {
         const double pj = p.pos[pos_index(p, c, it, d, j)], n = p.cosn[ang_index(p, c, q, d, m)];
         d2n += (pn[d] - pj) * (pn[d] - pj); dn += n * (pn[d] - pj);
         d2o += (po[d] - pj) * (po[d] - pj); dold += n * (po[d] - pj);
      }
      double inv;
      fast_r_invr(d2n, rr[0], inv); cs[0] = dn * inv;
      fast_r_invr(d2o, rr[1], inv); cs[1] = dold * inv;
      lpot2d_xn<2>(p, t, rr, cs, e);
      return e[0] - e[1];
   }
   return pair_energy<(KIND & 3)>(p, t, c, g, pn, j, it, nullptr, nullptr) - pair_energy<(KIND & 3)>(p, t, c, g, po, j, it, nullptr, nullptr);
}

// sum over the partners j = lane, lane+stride, ... of V(g at pn) - V(g at po) at slice it.  When the moved bead is an atom of
// a spline system (`fast_atoms`) only its atom partners are summed here, four per lane at a time with their position loads
// issued together; the caller deals the few remaining (rotor) partners to separate lanes.
template <int KIND>
__device__ __forceinline__ bool fast_atoms(const Params &p, int tg)
{
   return ((KIND & 3) != 2) && p.molecule[tg] == 0 && !p.minimage && p.n1d > 0;      // moved bead is an atom of a spline system
}
template <int KIND>
__device__ __forceinline__ double partner_sum_diff(const Params &p, const SmallTables &t, int c, int g, const double *pn, const double *po,
                                                   int it, int lane, int stride, int jx = -1)
{
   double D = 0.0;
   const int N = p.N;
   const int tg = type_of(p, g);
   const bool fast = fast_atoms<KIND>(p, tg);
   if (fast) {
      const int ja = p.first[tg], jb = ja + p.numb[tg];          // atom partners [ja, jb)
      const double *px = p.pos + pos_index(p, c, it, 0, 0), *py = px + p.Npad, *pz = py + p.Npad;
      for (int j0 = ja + lane; j0 < jb; j0 += 4 * stride) {
         double rn[4], ro[4];
         bool ok[4];
         #pragma unroll
         for (int u = 0; u < 4; u++) {
            const int j = j0 + u * stride;
            ok[u] = j < jb && j != g && j != jx;                    // jx: a partner whose term the caller adds later (pipelined sweep)
            const int jj = ok[u] ? j : (g == ja ? ja + 1 : ja);     // masked slots: any atom other than g (a real distance)
            const double qx = px[jj], qy = py[jj], qz = pz[jj];
            double inv_;
            fast_r_invr((pn[0] - qx) * (pn[0] - qx) + (pn[1] - qy) * (pn[1] - qy) + (pn[2] - qz) * (pn[2] - qz), rn[u], inv_);
            fast_r_invr((po[0] - qx) * (po[0] - qx) + (po[1] - qy) * (po[1] - qy) + (po[2] - qz) * (po[2] - qz), ro[u], inv_);
         }
         double e[4];
         bool bad = !p.uniform1d;
         if (p.poly1d) {
            #pragma unroll
            for (int u = 0; u < 4; u++) e[u] = spot1d_poly(p, t, rn[u], bad) - spot1d_poly(p, t, ro[u], bad);
         } else if (p.uniform1d) {
            #pragma unroll
            for (int u = 0; u < 4; u++) e[u] = spot1d_try(p, t, rn[u], bad) - spot1d_try(p, t, ro[u], bad);
         }
         if (__builtin_expect(bad, 0)) {
            #pragma unroll
            for (int u = 0; u < 4; u++) e[u] = spot1d_move(p, t, rn[u]) - spot1d_move(p, t, ro[u]);
         }
         #pragma unroll
         for (int u = 0; u < 4; u++) D += ok[u] ? e[u] : 0.0;
      }
      return D;
   }
   for (int j = lane; j < N; j += stride) {
      if (j == g || !partner_on_line<KIND>(p, c, j, it)) continue;         // world-line mask of an open worm (a22)
      D += pair_diff<KIND>(p, t, c, g, pn, po, j, it);
   }
   return D;
}

// NU full sub-batches of 32 atom partners starting at atom j0 (every lane has a partner in every sub-batch): the body of the
// fast path above without the range test.  g: the moved atom, jx: a partner left out (or -1).
template <int NU>
__device__ __forceinline__ double spline_batch(const Params &p, const SmallTables &t, const double *px, const double *py, const double *pz,
                                               int j0, int g, int jx, int jalt, const double *pn, const double *po)
{
   double rn[NU], ro[NU];
   bool ok[NU];
   #pragma unroll
   for (int u = 0; u < NU; u++) {
      const int j = j0 + 32 * u;
      ok[u] = j != g && j != jx;
      const int jj = ok[u] ? j : jalt;
      const double qx = px[jj], qy = py[jj], qz = pz[jj];
      double inv_;
      fast_r_invr((pn[0] - qx) * (pn[0] - qx) + (pn[1] - qy) * (pn[1] - qy) + (pn[2] - qz) * (pn[2] - qz), rn[u], inv_);
      fast_r_invr((po[0] - qx) * (po[0] - qx) + (po[1] - qy) * (po[1] - qy) + (po[2] - qz) * (po[2] - qz), ro[u], inv_);
   }
   double e[NU];
   bool bad = false;
   #pragma unroll
   for (int u = 0; u < NU; u++) e[u] = spot1d_poly(p, t, rn[u], bad) - spot1d_poly(p, t, ro[u], bad);
   if (__builtin_expect(bad, 0)) {
      #pragma unroll
      for (int u = 0; u < NU; u++) e[u] = spot1d_move(p, t, rn[u]) - spot1d_move(p, t, ro[u]);
   }
   double D = 0.0;
   #pragma unroll
   for (int u = 0; u < NU; u++) D += ok[u] ? e[u] : 0.0;
   return D;
}
// the first 32 * nfull atom partners of the species of atom g at slice `it` (lane = partner within a sub-batch); the
// remaining numb - 32 nfull partners of ALL the beads a warp works on go through one extra pass of the caller (tail_pair),
// so no sub-batch runs with most of its lanes masked (100 partners: 3 full sub-batches + 4 tail partners instead of 4 sub-batches)
template <int KIND>
__device__ __forceinline__ double partner_sum_main(const Params &p, const SmallTables &t, int c, int g, const double *pn, const double *po,
                                                   int it, int lane, int nfull, int jx)
{
   const int tg = type_of(p, g), ja = p.first[tg];
   const double *px = p.pos + pos_index(p, c, it, 0, 0), *py = px + p.Npad, *pz = py + p.Npad;
   const int jalt = (g == ja || jx == ja) ? ((g == ja + 1 || jx == ja + 1) ? ja + 2 : ja + 1) : ja;       // any atom other than g and jx
   double D = 0.0;
   int u0 = 0;
   for (; nfull - u0 >= 4; u0 += 4) D += spline_batch<4>(p, t, px, py, pz, ja + lane + 32 * u0, g, jx, jalt, pn, po);
   switch (nfull - u0) {
      case 3: D += spline_batch<3>(p, t, px, py, pz, ja + lane + 32 * u0, g, jx, jalt, pn, po); break;
      case 2: D += spline_batch<2>(p, t, px, py, pz, ja + lane + 32 * u0, g, jx, jalt, pn, po); break;
      case 1: D += spline_batch<1>(p, t, px, py, pz, ja + lane + 32 * u0, g, jx, jalt, pn, po); break;
      default: break;
   }
   return D;
}
// one (bead, partner) term of the tail pass: exact spline path
__device__ __forceinline__ double tail_pair(const Params &p, const SmallTables &t, int c, int it, int j, const double *pn, const double *po)
{
   double d2n = 0.0, d2o = 0.0;
   #pragma unroll
   for (int d = 0; d < 3; d++) {
      const double pj = p.pos[pos_index(p, c, it, d, j)];
      d2n += (pn[d] - pj) * (pn[d] - pj);
      d2o += (po[d] - pj) * (po[d] - pj);
   }
   double rn, ro, inv_;
   fast_r_invr(d2n, rn, inv_);
   fast_r_invr(d2o, ro, inv_);
   bool bad = false;
   double e = spot1d_poly(p, t, rn, bad) - spot1d_poly(p, t, ro, bad);
   if (__builtin_expect(bad, 0)) e = spot1d_move(p, t, rn) - spot1d_move(p, t, ro);
   return e;
}

__device__ __forceinline__ void bump_pos_epoch(const Params &p, Ctx &x)
{
   if (x.gthread == 0) p.pos_epoch[x.c] += 1;      // read by the rot leaders after the next chain barrier
}

// ---------------------------------------------------------------------------------------------
// whole-path move of every permutation cycle of `type` (MCMolecularMove / MCMolecularMoveExchange,
// mc_piqmc.cc:54-192).  dV of the rigid shift: cycle members against non-members, all P slices.
// A cycle of one world line (the common case) gives every warp of the chain whole slices and its
// lanes the partners; longer cycles use the flat (member, slice, partner) loop.
// ---------------------------------------------------------------------------------------------
template <int KIND>
__device__ void molecular_sweep(const Params &p, Ctx &x, int type)
{
   const int c = x.c, P = p.P, N = p.N;
   const int *cyc_start = p.cyc_start + (size_t)c * (N + 1);
   const int *cyc_atoms = p.cyc_atoms + (size_t)c * N;
   int g0 = 0;
   for (int t = 0; t < type; t++) g0 += p.ncyc[c * MAXT + t];
   int g1 = g0 + p.ncyc[c * MAXT + type];
   uint32_t *ms = stream_ptr(p, c, P + p.Q);
   bump_pos_epoch(p, x);
   const int gwarp = x.gthread >> 5, nwarps = x.nthreads_chain >> 5, lane = x.tid & 31;
   for (int g = g0; g < g1; g++) {
      Mrg rs;
      mrg_load(rs, ms);
      double u0 = mrg_u01(rs), u1 = mrg_u01(rs), u2 = mrg_u01(rs), u3 = mrg_u01(rs);
      double disp[3] = {p.mcstep[type] * (u0 - 0.5), p.mcstep[type] * (u1 - 0.5), p.mcstep[type] * (u2 - 0.5)};
      int b = cyc_start[g], len = cyc_start[g + 1] - b;
      double part = 0.0;
      if (len == 1) {
         const int a0 = cyc_atoms[b];
         const bool fast = fast_atoms<KIND>(p, type);
         const int nother = N - p.numb[type], base = p.first[type];
         for (int it = gwarp; it < P; it += nwarps) {
            double po[3], pn[3];
            #pragma unroll
            for (int d = 0; d < 3; d++) { po[d] = p.pos[pos_index(p, c, it, d, a0)]; pn[d] = po[d] + disp[d]; }
            part += partner_sum_diff<KIND>(p, x.t, c, a0, pn, po, it, lane, 32);
            if (fast)
               for (int jo = lane; jo < nother; jo += 32) {
                  const int j = (jo < base) ? jo : jo + p.numb[type];
                  part += pair_diff<KIND>(p, x.t, c, a0, pn, po, j, it);
               }
         }
      } else {
         long nitems = (long)len * P * N;
         for (long i = x.gthread; i < nitems; i += x.nthreads_chain) {
            int j = (int)(i % N);
            long r = i / N;
            int it = (int)(r % P);
            int a0 = cyc_atoms[b + (int)(r / P)];
            bool member = false;
            for (int k = 0; k < len; k++) member |= (cyc_atoms[b + k] == j);
            if (member || !partner_on_line<KIND>(p, c, j, it)) continue;
            double po[3], pn[3];
            #pragma unroll
            for (int d = 0; d < 3; d++) { po[d] = p.pos[pos_index(p, c, it, d, a0)]; pn[d] = po[d] + disp[d]; }
            part += pair_diff<KIND>(p, x.t, c, a0, pn, po, j, it);
         }
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
      chain_sync(p, x);
   }
}

// ---------------------------------------------------------------------------------------------
// The whole-path sweep of an atomic species without exchange cycles, one cross-CTA hand-over per atom instead of two
// chain barriers.  Every warp of the chain owns a fixed set of slices for the whole sweep: the pair action is diagonal
// in imaginary time, so a warp only ever reads and shifts beads of its own slices -- positions need no synchronisation
// at all.  What couples the chain is the scalar dV of an atom: every CTA publishes its partial sum (release) and every
// CTA waits for all partials (acquire), adds them in CTA order and takes the same decision from the same uniform.
// The sum of atom a+1 without its partner a is formed BEFORE the wait for atom a; the one missing pair term per slice
// follows the decision.  Draws (4 per atom from the chain's miscellaneous stream) and acceptance as in molecular_sweep.
// ---------------------------------------------------------------------------------------------
template <int KIND>
__device__ double molecular_partial(const Params &p, Ctx &x, int type, int a0, const double *disp, int jx)
{
   const int c = x.c, P = p.P, N = p.N, base = p.first[type], na = p.numb[type], nother = N - na;
   const int nwc = blockDim.x >> 5, gwarp = x.crank * nwc + (x.tid >> 5), nwarps = p.cpc * nwc, lane = x.tid & 31;
   double part = 0.0;
   const int nfull = na / 32, ntail = na - 32 * nfull;
   const int nmine = (P - gwarp + nwarps - 1) / nwarps;               // slices of this warp
   const bool split = p.poly1d && nfull >= 1 && nmine * ntail <= 32;
   if (split && lane < nmine * ntail) {                                // tail partners of all my slices in one pass
      const int k = lane / ntail, j = base + 32 * nfull + (lane - k * ntail), it = gwarp + k * nwarps;
      if (j != a0 && j != jx) {
         double po[3], pn[3];
         #pragma unroll
         for (int d = 0; d < 3; d++) { po[d] = p.pos[pos_index(p, c, it, d, a0)]; pn[d] = po[d] + disp[d]; }
         part += tail_pair(p, x.t, c, it, j, pn, po);
      }
   }
   for (int it = gwarp; it < P; it += nwarps) {
      double po[3], pn[3];
      #pragma unroll
      for (int d = 0; d < 3; d++) { po[d] = p.pos[pos_index(p, c, it, d, a0)]; pn[d] = po[d] + disp[d]; }
      if (split) part += partner_sum_main<KIND>(p, x.t, c, a0, pn, po, it, lane, nfull, jx);
      else part += partner_sum_diff<KIND>(p, x.t, c, a0, pn, po, it, lane, 32, jx);
      for (int jo = lane; jo < nother; jo += 32) {
         const int j = (jo < base) ? jo : jo + na;
         part += pair_diff<KIND>(p, x.t, c, a0, pn, po, j, it);
      }
   }
   return part;
}

template <int KIND>
__device__ void molecular_sweep_piped(const Params &p, Ctx &x, int type)
{
   const int c = x.c, P = p.P, base = p.first[type], na = p.numb[type];
   const int nwc = blockDim.x >> 5, warp = x.tid >> 5, gwarp = x.crank * nwc + warp, nwarps = p.cpc * nwc, lane = x.tid & 31;
   uint32_t *ms = stream_ptr(p, c, P + p.Q);
   double *slots = p.scratch + (size_t)c * 64;                      // [parity of the atom][CTA]
   unsigned *cnt = p.barrier + (size_t)c * 32 + 24;                 // [parity] arrivals, monotonic within a launch
   __shared__ int s_acc;
   bump_pos_epoch(p, x);
   Mrg rs;
   mrg_load(rs, ms);
   const double step = p.mcstep[type];
   double disp[3], u3, dnext[3], u3next = 0.0;
   { const double u0 = mrg_u01(rs), u1 = mrg_u01(rs), u2 = mrg_u01(rs); u3 = mrg_u01(rs); disp[0] = step * (u0 - 0.5); disp[1] = step * (u1 - 0.5); disp[2] = step * (u2 - 0.5); }
   double part = molecular_partial<KIND>(p, x, type, base, disp, -1);
   for (int a = 0; a < na; a++) {
      const int par = a & 1;
      // CTA total of atom a in warp order, published for the other CTAs
      part = team_sum(part, 32);
      if (lane == 0) x.red[warp] = part;
      __syncthreads();
      if (x.tid == 0) {
         double sct = 0.0;
         for (int w = 0; w < nwc; w++) sct += x.red[w];
         slots[par * 32 + x.crank] = sct;
         asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cnt + par) : "memory");
      }
      // meanwhile: the sum of atom a+1 without its partner a
      double pre = 0.0;
      if (a + 1 < na) {
         const double u0 = mrg_u01(rs), u1 = mrg_u01(rs), u2 = mrg_u01(rs); u3next = mrg_u01(rs);
         dnext[0] = step * (u0 - 0.5); dnext[1] = step * (u1 - 0.5); dnext[2] = step * (u2 - 0.5);
         pre = molecular_partial<KIND>(p, x, type, base + a + 1, dnext, base + a);
      }
      // every partial of atom a, the decision
      if (x.tid == 0) {
         x.mol_target[par] += (unsigned)p.cpc;
         unsigned v;
         do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt + par) : "memory"); } while ((int)(v - x.mol_target[par]) < 0);
         double dv = 0.0;
         for (int r = 0; r < p.cpc; r++) dv += __ldcg(slots + par * 32 + r);
         s_acc = ((dv < 0.0) || (exp(-dv * p.tau) > u3)) ? 1 : 0;
      }
      __syncthreads();
      const bool acc = s_acc != 0;
      if (acc && lane < 3)
         for (int it = gwarp; it < P; it += nwarps) p.pos[pos_index(p, c, it, lane, base + a)] += disp[lane];
      if (x.gthread == 0) {
         double *cn = counter_ptr(p, c, type, 0);
         cn[0] += 1.0;
         if (acc) cn[1] += 1.0;
      }
      __syncwarp();
      // the missing partner term of atom a+1, slice by slice (a lane per owned slice)
      part = pre;
      if (a + 1 < na) {
         int k = 0;
         for (int it = gwarp; it < P; it += nwarps, k++) {
            if ((k & 31) != lane) continue;
            double d2n = 0.0, d2o = 0.0;
            #pragma unroll
            for (int d = 0; d < 3; d++) {
               const double pj = p.pos[pos_index(p, c, it, d, base + a)], po = p.pos[pos_index(p, c, it, d, base + a + 1)], pn = po + dnext[d];
               d2n += (pn - pj) * (pn - pj);
               d2o += (po - pj) * (po - pj);
            }
            part += spot1d_move(p, x.t, sqrt(d2n)) - spot1d_move(p, x.t, sqrt(d2o));
         }
         disp[0] = dnext[0]; disp[1] = dnext[1]; disp[2] = dnext[2]; u3 = u3next;
      }
   }
   if (x.gthread == 0) mrg_store(rs, ms);
   chain_sync(p, x);
}

// ---------------------------------------------------------------------------------------------
// bisection sweep (MCBisectionMove / MCBisectionMoveExchange, mc_piqmc.cc:194-419): every atom of
// `type` in sequence; for one atom all P/seg segments [s0, s0+seg], s0 = off + k*seg, in parallel,
// one team per segment.  Level sums are kept incrementally: pot0(l) = S(l-1) + D(l) where D(l) is
// the dV of the midpoints sampled at level l, so delta = (D - S)*tau*seg_l/2 equals the
// reference's (pot0 - 2 pot1)*tau*(seg_l/2) without re-evaluating earlier levels (:263 "inefficient").
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void team_sync(const Ctx &x)
{
   if (x.T <= 32) __syncwarp();
   else asm volatile("bar.sync %0, %1;" ::"r"(x.team_cta + 1), "r"(x.T) : "memory");     // one named barrier per team (<= 15 teams per CTA)
}
// fixed-order sum over the team; all lanes return the total
__device__ __forceinline__ double team_reduce(const Ctx &x, double v, double *red)
{
   v = team_sum(v, x.wlanes);
   if (x.W == 1) return v;
   if (x.wl == 0) red[x.tw] = v;
   team_sync(x);
   double s = 0.0;
   for (int w = 0; w < x.W; w++) s += red[w];
   return s;
}

template <int KIND>
__device__ void bisection_sweep(const Params &p, Ctx &x, int type, int off)
{
   const int c = x.c, P = p.P, N = p.N, T = x.T, W = x.W;
   const int L = p.levels[type], seg = 1 << L, nseg = P / seg;
   const int base = p.first[type];
   const double bnorm = 1.0 / (p.lambda[type] * p.tau);
   const int nrounds = (nseg + x.nteams_chain - 1) / x.nteams_chain;
   bump_pos_epoch(p, x);
   // The pair action is diagonal in imaginary time and the segment end points are fixed during a sweep, so
   // segment k of every atom only ever reads slices of segment k: a team keeps its segment index and walks
   // through the atoms in sequence with no chain-wide barrier; the sweep ends with one barrier.
   for (int rd = 0; rd < nrounds; rd++) {
      const int k = rd * x.nteams_chain + x.team_id;
      const bool active = k < nseg;
      const int s0 = active ? (off + k * seg) % P : 0;
      double *nx = p.segbuf_global ? p.segbuf + ((size_t)c * p.nseg_max + k) * p.team_buf_n : x.team_buf;      // nseg_max is rounded up to whole rounds
      double *xi = nx + (p.seg_max + 1) * 3;
      uint32_t *rc = reinterpret_cast<uint32_t *>(nx + (p.seg_max + 1) * 6);       // streams of slices s0 .. s0+seg-1, cached for the sweep
      double *tred = nx + (p.seg_max + 1) * 6 + p.seg_max * 3;
      double *ucand = tred + 4;                                                  // next L uniforms of the segment's accept stream ...
      uint32_t *cst = reinterpret_cast<uint32_t *>(tred + 4 + MAXLEV);           // ... and the stream state after each of them
      if (active)
         for (int i = x.lane_t; i < seg * 6; i += T) rc[i] = stream_ptr(p, c, (s0 + i / 6) % P)[i % 6];
      team_sync(x);
      for (int a = 0; a < p.numb[type]; a++) {
         const int gA = base + a;
         const int gB = (p.stat[type] == 1) ? p.pindex[(size_t)c * N + gA] : gA;
         if (active)
            for (int i = x.lane_t; i < 6; i += T) {
               int e = i / 3, d = i - 3 * e, t = e * seg;
               int g = (s0 + t >= P) ? gB : gA;
               nx[t * 3 + d] = p.pos[pos_index(p, c, (s0 + t) % P, d, g)];
            }
         // candidate uniforms of the level tests (stream of slice s0): drawn ahead by the lanes that idle below, consumed
         // only where a level needs one (delta >= 0), the stream is advanced by the number actually used
         if (active)
            for (int k = T - 1 - x.lane_t; k < L; k += T) {
               Mrg rs;
               mrg_load(rs, rc);
               double u = 0.0;
               for (int j = 0; j <= k; j++) u = mrg_u01(rs);
               ucand[k] = u;
               mrg_store(rs, cst + k * 6);
            }
         int used = 0;
         // unit normals for all interior slices of the segment, drawn up front from the slices' own streams
         // (6 uniforms per slice: gauss() of mc_randg.cc:138-150 per dimension)
         MARK(x, 20);
         for (int i0 = 0; i0 < (seg - 1) * 3; i0 += T) {
            const int i = i0 + x.lane_t;
            const bool valid = active && i < (seg - 1) * 3;
            const int t = valid ? 1 + i / 3 : 1, d = i - 3 * (t - 1);
            Mrg rs;
            if (valid) {
               mrg_load(rs, rc + t * 6);
               double r1 = 0, r2 = 0;
               for (int kk = 0; kk <= d; kk++) { r1 = mrg_u01(rs); r2 = mrg_u01(rs); }
               for (int kk = d + 1; kk < 3; kk++) { mrg_u01(rs); mrg_u01(rs); }
               xi[t * 3 + d] = sqrt(-log(r1)) * cos(2.0 * PI * r2);
            }
            team_sync(x);                               // every lane of a slice has read the state before it advances
            if (valid && d == 2) mrg_store(rs, rc + t * 6);
            team_sync(x);
         }
         MARK(x, 21);
         double S = 0.0;
         bool alive = active;
         for (int level = 0; level < L; level++) {
            const int lss = seg >> level, half = lss >> 1, nmid = 1 << level;
            if (alive) {
               const double sq = sqrt(bnorm / (double)lss);
               for (int i = x.lane_t; i < nmid * 3; i += T) {
                  int m = i / 3, d = i - 3 * m;
                  int t1 = half + m * lss;
                  nx[t1 * 3 + d] = 0.5 * (nx[(t1 - half) * 3 + d] + nx[(t1 + half) * 3 + d]) + xi[t1 * 3 + d] / sq;
               }
            }
            team_sync(x);
            double D = 0.0;
            if (alive && fast_atoms<KIND>(p, type)) {
               // the partners that are not atoms of this type (the rotor): one lane per (midpoint, partner)
               // (dealt from the top lane down, spread over the team: the low lanes carry the most atom partners)
               const int nother = N - p.numb[type];
               const int nt = nmid * nother, sdeal = max(1, T / max(1, nt)), rl = T - 1 - x.lane_t;
               for (int i = (rl % sdeal == 0) ? rl / sdeal : nt; i < nt; i += (T + sdeal - 1) / sdeal) {
                  const int m = i / nother, jo = i - m * nother;
                  const int j = (jo < base) ? jo : jo + p.numb[type];
                  const int t1 = half + m * lss;
                  const int sl = (s0 + t1) % P;
                  const int g = (s0 + t1 >= P) ? gB : gA;
                  double po[3], pn[3];
                  #pragma unroll
                  for (int d = 0; d < 3; d++) { po[d] = p.pos[pos_index(p, c, sl, d, g)]; pn[d] = nx[t1 * 3 + d]; }
                  D += pair_diff<KIND>(p, x.t, c, g, pn, po, j, sl);
               }
            }
            // general pair terms (tops, rotors, minimum image, worm masks) of a one-warp team with few partners: the
            // (midpoint, partner) pairs of the level are evaluated flat over the team -- a system of two particles (C1) has ONE
            // partner per midpoint, and sixteen midpoints in turn on one lane was most of its sweep -- into the spare half of
            // the team's buffer; every lane then adds the terms it used to evaluate itself, in the order it used to (midpoints
            // in turn, its partners in stride order), so the sum is the same bit for bit
            double *esc = nx + (p.seg_max + 1) * 6 + p.seg_max * 3 + TEAM_EXTRA;
            const bool flat = !fast_atoms<KIND>(p, type) && W == 1 && N < 2 * T && nmid * N <= (p.seg_max + 1) * 6;
            if (flat) {
               if (alive)
                  for (int i = x.lane_t; i < nmid * N; i += T) {
                     const int m = i / N, j = i - m * N;
                     const int t1 = half + m * lss;
                     const int sl = (s0 + t1) % P;
                     const int g = (s0 + t1 >= P) ? gB : gA;
                     double e = 0.0;
                     if (j != g && partner_on_line<KIND>(p, c, j, sl)) {
                        double po[3], pn[3];
                        #pragma unroll
                        for (int d = 0; d < 3; d++) { po[d] = p.pos[pos_index(p, c, sl, d, g)]; pn[d] = nx[t1 * 3 + d]; }
                        e = pair_diff<KIND>(p, x.t, c, g, pn, po, j, sl);
                     }
                     esc[i] = e;
                  }
               team_sync(x);
               if (alive)
                  for (int m = 0; m < nmid; m++) {
                     const int t1 = half + m * lss;
                     const int sl = (s0 + t1) % P;
                     const int g = (s0 + t1 >= P) ? gB : gA;
                     double sm = 0.0;
                     for (int j = x.lane_t; j < N; j += T) {
                        if (j == g || !partner_on_line<KIND>(p, c, j, sl)) continue;
                        sm += esc[m * N + j];
                     }
                     D += sm;
                  }
            } else if (alive) {
               // the warps of the team take the midpoints in turn; with fewer midpoints than warps several warps share one
               const int wpm = (nmid >= W) ? 1 : W / nmid, mstep = W / wpm;
               const int li = (x.tw % wpm) * x.wlanes + x.wl, stride = wpm * x.wlanes;
               for (int m = x.tw / wpm; m < nmid; m += mstep) {
                  const int t1 = half + m * lss;
                  const int sl = (s0 + t1) % P;
                  const int g = (s0 + t1 >= P) ? gB : gA;
                  double po[3], pn[3];
                  #pragma unroll
                  for (int d = 0; d < 3; d++) { po[d] = p.pos[pos_index(p, c, sl, d, g)]; pn[d] = nx[t1 * 3 + d]; }
                  D += partner_sum_diff<KIND>(p, x.t, c, g, pn, po, sl, li, stride);
               }
            }
            MARK(x, 22);
            D = team_reduce(x, D, tred);
            const double deltav = (D - S) * (p.tau * (double)half);
            S += D;
            // every lane holds the same D (symmetric butterfly, fixed-order combine), so the test needs no broadcast
            int acc = 1;
            if (alive && !(deltav < 0.0)) { acc = (exp(-deltav) > ucand[used]) ? 1 : 0; used++; }
            if (!acc) alive = false;
            MARK(x, 23);
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
         if (active && used > 0)
            for (int i = x.lane_t; i < 6; i += T) rc[i] = cst[(used - 1) * 6 + i];
         team_sync(x);
         MARK(x, 24);
      }
      if (active)
         for (int i = x.lane_t; i < seg * 6; i += T) stream_ptr(p, c, (s0 + i / 6) % P)[i % 6] = rc[i];
      team_sync(x);
   }
   chain_sync(p, x);
   MARK(x, 25);
}

// ---------------------------------------------------------------------------------------------
// The same sweep software-pipelined over the atoms, for two-warp teams of an atomic species of a spline system.
// Atom a only feels the outcome of atom a-1 through ONE partner term per bead (the pair action is a sum over partners),
// so warp (a & 1) of the team proposes atom a -- normals, the midpoints of every level, the partner sums of every level
// without the partner a-1 -- while the other warp is still deciding atom a-1; once that atom is final it adds the
// missing term, runs the level tests and writes back.  Draws per stream, order of the level tests, early rejection and
// the acceptance rule are those of bisection_sweep (the sums of levels below a rejection are computed and dropped); the
// per-atom critical path shrinks from "everything" to "one partner term + the tests".
// Hand-over between the two warps: two counters in shared memory (normals drawn up to atom .., final up to atom ..).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_wait_ge(volatile int *flag, int v)
{
   while (*flag < v) { }
   __syncwarp();
   __threadfence_block();
}
__device__ __forceinline__ void warp_post(volatile int *flag, int v)
{
   __syncwarp();
   __threadfence_block();
   if ((threadIdx.x & 31) == 0) *flag = v;
}

template <int KIND>
__device__ void bisection_sweep_piped(const Params &p, Ctx &x, int type, int off)
{
   const int c = x.c, P = p.P, N = p.N;
   const int L = p.levels[type], seg = 1 << L, nseg = P / seg;
   const int base = p.first[type], na = p.numb[type];
   const double bnorm = 1.0 / (p.lambda[type] * p.tau);
   const int nrounds = (nseg + x.nteams_chain - 1) / x.nteams_chain;
   const int w = x.tw, lane = x.wl;                       // warp of the team (0, 1), lane of the warp
   const int nother = N - na;
   bump_pos_epoch(p, x);
   for (int rd = 0; rd < nrounds; rd++) {
      const int k = rd * x.nteams_chain + x.team_id;
      const bool active = k < nseg;
      const int s0 = active ? (off + k * seg) % P : 0;
      double *buf = x.team_buf;
      double *nx = buf + (size_t)w * (p.seg_max + 1) * 6, *xi = nx + (p.seg_max + 1) * 3;      // this warp's positions / normals
      uint32_t *rc = reinterpret_cast<uint32_t *>(buf + (p.seg_max + 1) * 12);                  // streams of slices s0 .. s0+seg-1
      double *dsum = buf + (p.seg_max + 1) * 12 + p.seg_max * 3 + w * (MAXLEV + 2);             // this warp's level sums
      volatile int *flags = reinterpret_cast<volatile int *>(buf + (p.seg_max + 1) * 12 + p.seg_max * 3 + 2 * (MAXLEV + 2));   // [0] normals drawn, [1] atoms final
      if (active)
         for (int i = x.lane_t; i < seg * 6; i += x.T) rc[i] = stream_ptr(p, c, (s0 + i / 6) % P)[i % 6];
      if (x.lane_t == 0) { flags[0] = 0; flags[1] = 0; }
      team_sync(x);
      if (active)
      for (int a = w; a < na; a += 2) {
         const int gA = base + a;
         const int gB = (p.stat[type] == 1) ? p.pindex[(size_t)c * N + gA] : gA;
         // the partner still in flight: atom a-1 (its beads past beta sit on its successor world line)
         const int hA = a > 0 ? gA - 1 : -1;
         const int hB = a > 0 ? ((p.stat[type] == 1) ? p.pindex[(size_t)c * N + hA] : hA) : -1;
         if (lane < 6) {
            const int e = lane / 3, d = lane - 3 * e, t = e * seg;
            nx[t * 3 + d] = p.pos[pos_index(p, c, (s0 + t) % P, d, (s0 + t >= P) ? gB : gA)];
         }
         // unit normals of the interior slices from the slices' own streams, after those of atom a-1
         if (a > 0) warp_wait_ge(flags, a);
         for (int i0 = 0; i0 < (seg - 1) * 3; i0 += 32) {
            const int i = i0 + lane;
            const bool valid = i < (seg - 1) * 3;
            const int t = valid ? 1 + i / 3 : 1, d = i - 3 * (t - 1);
            Mrg rs;
            if (valid) {
               mrg_load(rs, rc + t * 6);
               double r1 = 0, r2 = 0;
               for (int kk = 0; kk <= d; kk++) { r1 = mrg_u01(rs); r2 = mrg_u01(rs); }
               for (int kk = d + 1; kk < 3; kk++) { mrg_u01(rs); mrg_u01(rs); }
               xi[t * 3 + d] = sqrt(-log(r1)) * cos(2.0 * PI * r2);
            }
            __syncwarp();                               // every lane of a slice has read the state before it advances
            if (valid && d == 2) mrg_store(rs, rc + t * 6);
            __syncwarp();
         }
         warp_post(flags, a + 1);
         // midpoints of every level, then the partner sums of every level without the partner in flight
         for (int level = 0; level < L; level++) {
            const int lss = seg >> level, half = lss >> 1, nmid = 1 << level;
            const double sq = sqrt(bnorm / (double)lss);
            for (int i = lane; i < nmid * 3; i += 32) {
               const int m = i / 3, d = i - 3 * m, t1 = half + m * lss;
               nx[t1 * 3 + d] = 0.5 * (nx[(t1 - half) * 3 + d] + nx[(t1 + half) * 3 + d]) + xi[t1 * 3 + d] / sq;
            }
            __syncwarp();
         }
         // partners of the other species (the rotor): every (midpoint of any level, partner) term in ONE pass, a lane per term;
         // midpoint index gm = 2^level - 1 + m enumerates the levels in order
         double eo = 0.0;
         int eo_level = -1;
         if ((seg - 1) * nother <= 32) {
            if (lane < (seg - 1) * nother) {
               const int gm = lane / nother, jo = lane - gm * nother;
               const int level = 31 - __clz(gm + 1), m = gm + 1 - (1 << level);
               const int lss = seg >> level, half = lss >> 1;
               const int j = (jo < base) ? jo : jo + na;
               const int t1 = half + m * lss, sl = (s0 + t1) % P, g = (s0 + t1 >= P) ? gB : gA;
               double po[3], pn[3];
               #pragma unroll
               for (int d = 0; d < 3; d++) { po[d] = p.pos[pos_index(p, c, sl, d, g)]; pn[d] = nx[t1 * 3 + d]; }
               eo = pair_diff<KIND>(p, x.t, c, g, pn, po, j, sl);
               eo_level = level;
            }
         }
         // atom partners beyond the last full sub-batch of 32: every (midpoint, tail partner) term in one pass
         const int nfull = na / 32, ntail = na - 32 * nfull;
         const bool split = p.poly1d && nfull >= 1 && (seg - 1) * ntail <= 32;
         double et = 0.0;
         int et_level = -1;
         if (split && lane < (seg - 1) * ntail) {
            const int gm = lane / ntail, j = base + 32 * nfull + (lane - gm * ntail);
            const int level = 31 - __clz(gm + 1), m = gm + 1 - (1 << level);
            const int lss = seg >> level, half = lss >> 1;
            const int t1 = half + m * lss, sl = (s0 + t1) % P;
            const bool wrap = s0 + t1 >= P;
            const int g = wrap ? gB : gA;
            if (j != g && j != (wrap ? hB : hA)) {
               double po[3], pn[3];
               #pragma unroll
               for (int d = 0; d < 3; d++) { po[d] = p.pos[pos_index(p, c, sl, d, g)]; pn[d] = nx[t1 * 3 + d]; }
               et = tail_pair(p, x.t, c, sl, j, pn, po);
            }
            et_level = level;
         }
         for (int level = 0; level < L; level++) {
            const int lss = seg >> level, half = lss >> 1, nmid = 1 << level;
            double D = ((eo_level == level) ? eo : 0.0) + ((et_level == level) ? et : 0.0);
            if ((seg - 1) * nother > 32)
               for (int i = lane; i < nmid * nother; i += 32) {       // many terms: level by level
                  const int m = i / nother, jo = i - m * nother;
                  const int j = (jo < base) ? jo : jo + na;
                  const int t1 = half + m * lss, sl = (s0 + t1) % P, g = (s0 + t1 >= P) ? gB : gA;
                  double po[3], pn[3];
                  #pragma unroll
                  for (int d = 0; d < 3; d++) { po[d] = p.pos[pos_index(p, c, sl, d, g)]; pn[d] = nx[t1 * 3 + d]; }
                  D += pair_diff<KIND>(p, x.t, c, g, pn, po, j, sl);
               }
            for (int m = 0; m < nmid; m++) {
               const int t1 = half + m * lss, sl = (s0 + t1) % P;
               const bool wrap = s0 + t1 >= P;
               const int g = wrap ? gB : gA;
               double po[3], pn[3];
               #pragma unroll
               for (int d = 0; d < 3; d++) { po[d] = p.pos[pos_index(p, c, sl, d, g)]; pn[d] = nx[t1 * 3 + d]; }
               if (split) D += partner_sum_main<KIND>(p, x.t, c, g, pn, po, sl, lane, nfull, wrap ? hB : hA);
               else D += partner_sum_diff<KIND>(p, x.t, c, g, pn, po, sl, lane, 32, wrap ? hB : hA);
            }
            D = team_sum(D, 32);
            if (lane == 0) dsum[level] = D;
         }
         __syncwarp();
         // atom a-1 is final: its term, the level tests, the write-back
         if (a > 0) warp_wait_ge(flags + 1, a);
         double S = 0.0;
         bool alive = true;
         int used = 0;
         Mrg as;
         if (lane == 0) mrg_load(as, rc);
         // the term of atom a-1 for every midpoint of every level in one pass (a lane per midpoint), summed per level
         if (a > 0) {
            for (int gm = lane; gm < seg - 1; gm += 32) {
               const int level = 31 - __clz(gm + 1), m = gm + 1 - (1 << level);
               const int lss = seg >> level, half = lss >> 1;
               const int t1 = half + m * lss, sl = (s0 + t1) % P;
               const bool wrap = s0 + t1 >= P;
               const int g = wrap ? gB : gA, j = wrap ? hB : hA;
               double d2n = 0.0, d2o = 0.0;
               #pragma unroll
               for (int d = 0; d < 3; d++) {
                  const double pj = __ldcg(p.pos + pos_index(p, c, sl, d, j)), pg = p.pos[pos_index(p, c, sl, d, g)], pn = nx[t1 * 3 + d];
                  d2n += (pn - pj) * (pn - pj);
                  d2o += (pg - pj) * (pg - pj);
               }
               xi[(gm + 1) * 3] = spot1d_move(p, x.t, sqrt(d2n)) - spot1d_move(p, x.t, sqrt(d2o));     // the normals are spent: reuse their slots
            }
            __syncwarp();
         }
         for (int level = 0; level < L && alive; level++) {
            const int half = seg >> (level + 1), nmid = 1 << level;
            double e = 0.0;
            if (a > 0)
               for (int m = 0; m < nmid; m++) e += xi[(nmid + m) * 3];          // gm + 1 = 2^level + m, fixed order
            const double D = dsum[level] + e;
            const double deltav = (D - S) * (p.tau * (double)half);
            S += D;
            if (!(deltav < 0.0)) {
               double u = 0.0;
               if (lane == 0) { u = mrg_u01(as); }
               u = __shfl_sync(0xffffffffu, u, 0);
               used++;
               if (!(exp(-deltav) > u)) alive = false;
            }
         }
         if (lane == 0) {
            if (used > 0) mrg_store(as, rc);
            double *cn = counter_ptr(p, c, type, 1);
            atomicAdd(cn, 1.0);
            if (alive) atomicAdd(cn + 1, 1.0);
         }
         if (alive)
            for (int i = lane; i < (seg - 1) * 3; i += 32) {
               const int t = 1 + i / 3, d = i % 3;
               p.pos[pos_index(p, c, (s0 + t) % P, d, (s0 + t >= P) ? gB : gA)] = nx[t * 3 + d];
            }
         warp_post(flags + 1, a + 1);
      }
      team_sync(x);
      if (active)
         for (int i = x.lane_t; i < seg * 6; i += x.T) stream_ptr(p, c, (s0 + i / 6) % P)[i % 6] = rc[i];
      team_sync(x);
   }
   chain_sync(p, x);
}

// ---------------------------------------------------------------------------------------------
// rotational Metropolis step at rot slice q for rotor m (MCRot3Dstep mc_piqmc.cc:938-1199,
// MCRotLinStep :781-936), executed by one rot group:
//   leader: uniforms, proposal, orientation matrices          -> shared slot          | group sync
//   all   : the four density factors (first four threads) and the partner/slice sum of the
//           potential for the proposed orientation; the sum for the current orientation is taken
//           from the per-slice cache unless a translational sweep or a partner rotor invalidated it
//           (mathematically the reference's pot_old, :1069-1075)                      | group sync
//   leader: acceptance, state and cache update
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void group_sync(const Ctx &x)
{
   if (x.G <= 32) __syncwarp(x.gmask);          // groups that share a warp may be in different places (rot_run)
   else if (x.ngrp <= 15) asm volatile("bar.sync %0, %1;" ::"r"(x.grp + 1), "r"(x.G) : "memory");   // one named barrier per rot group
   else __syncthreads();
}

// sum over the R translational slices of rot slice q and all partners of the rotor's potential
// for the orientation in `o` (KIND 2: rotation matrix; KIND 1: unit vector); strided over the group
template <int KIND>
__device__ __forceinline__ double rot_potential(const Params &p, Ctx &x, int g, int q, const double *o)
{
   const int c = x.c, N = p.N, R = p.R, it0 = q * R;
   double v = 0.0;
   if ((KIND & 3) == 2) {
      Mat3 ro;
      #pragma unroll
      for (int i = 0; i < 9; i++) ro.m[i / 3][i % 3] = o[i];
      // rotor partners: the partner's orientation is the same for all R slices, one partner per lane
      const int m0 = p.first[p.imtype];
      for (int j = m0 + x.gl; j < m0 + p.numb[p.imtype]; j += x.G) {
         if (j == g) continue;
         Mat3 rb;
         load_rotmat(p, c, q, j - m0, rb);
         for (int r = 0; r < R; r++) {
            double pg[3], pj[3];
            #pragma unroll
            for (int d = 0; d < 3; d++) { pg[d] = p.pos[pos_index(p, c, it0 + r, d, g)]; pj[d] = p.pos[pos_index(p, c, it0 + r, d, j)]; }
            Tip4pSites sa, sb;
            tip4p_sites(ro, pg, sa);
            tip4p_sites(rb, pj, sb);
            v += caleng(sa, sb);
         }
      }
      // atom partners: the (atom, slice) terms dealt flat to the lanes of the group, so a lone atom's R terms (C1) are
      // evaluated side by side instead of one after the other
      if (p.ntypes > 1) {
         const int at = 1 - p.imtype, a0 = p.first[at], nitems = p.numb[at] * R;
         for (int i = x.gl; i < nitems; i += x.G) {
            const int jj = i / R, r = i - jj * R, j = a0 + jj;
            if (!partner_on_line<KIND>(p, c, j, it0 + r)) continue;
            double pg[3], pj[3];
            #pragma unroll
            for (int d = 0; d < 3; d++) { pg[d] = p.pos[pos_index(p, c, it0 + r, d, g)]; pj[d] = p.pos[pos_index(p, c, it0 + r, d, j)]; }
            v += p.ispher ? vspher(p, sqrt(dist2(pg, pj))) : vcord(p, ro, pg, pj, nullptr, nullptr);
         }
      }
   } else {
      // items (slice r, partner j) without integer division: the warps of the group stride over the R slices,
      // their lanes over the partners (coalesced over j); a group narrower than a warp walks the slices in turn
      const int lanes = (x.G < 32) ? x.G : 32, lane = x.gl & (lanes - 1);
      const int gw = x.gl / lanes, ngw = x.G / lanes;
      for (int r = gw; r < R; r += ngw) {
         const int it = it0 + r;
         const double gx = p.pos[pos_index(p, c, it, 0, g)], gy = p.pos[pos_index(p, c, it, 1, g)], gz = p.pos[pos_index(p, c, it, 2, g)];
         const double *px = p.pos + pos_index(p, c, it, 0, 0), *py = px + p.Npad, *pz = py + p.Npad;
         // four partners per lane in flight: position loads, then all table gathers, then the bilinear forms, so
         // the L2 latencies of independent evaluations overlap (the loop body alone exposes ~2 k cycles per partner)
         for (int j0 = lane; j0 < N; j0 += 4 * lanes) {
            double rr[4], cs[4];
            bool ok[4];
            #pragma unroll
            for (int u = 0; u < 4; u++) {
               const int j = j0 + u * lanes;
               ok[u] = j < N && j != g && partner_on_line<KIND>(p, c, min(j, N - 1), it);
               const int jj = ok[u] ? j : (g == 0 ? 1 : 0);        // a valid, distinct partner for masked slots
               double dx = gx - px[jj], dy = gy - py[jj], dz = gz - pz[jj];
               double dr2 = dx * dx + dy * dy + dz * dz;
               double dot = o[0] * dx + o[1] * dy + o[2] * dz;
               double invr;
               fast_r_invr(dr2, rr[u], invr);
               cs[u] = -dot * invr;
            }
            if (!(ok[0] | ok[1] | ok[2] | ok[3])) continue;       // nothing to evaluate (e.g. a lone rotor: no table loaded)
            double e[4];
            lpot2d_xn<4>(p, x.t, rr, cs, e);
            #pragma unroll
            for (int u = 0; u < 4; u++) v += ok[u] ? e[u] : 0.0;
         }
      }
   }
   return v;
}

// ---------------------------------------------------------------------------------------------
// geometry cache of a linear rotor's partner terms (one rotor, partners are atoms, no worm, no minimum image).
// Item k = r * (N - 1) + jj of rot slice q is the pair (translational slice q R + r, jj-th atom other than the rotor).
// Thread gl of the slice's rot group owns the items k = gl, gl + G, ... in BOTH the fill and the evaluation, so a
// thread only ever reads what it wrote itself: no barrier between the two.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void geo_fill(const Params &p, Ctx &x, int g, int q)
{
   const int c = x.c, NA = p.N - 1, it0 = q * p.R;
   double *gb = p.geo + ((size_t)c * p.Q + q) * 4 * p.geo_n;
   const double rmin = x.t.rgi2d[0].x;
   for (int k = x.gl; k < p.geo_items; k += x.G) {
      const int r = k / NA, jj = k - r * NA, j = jj < g ? jj : jj + 1, it = it0 + r;
      const double dx = p.pos[pos_index(p, c, it, 0, j)] - p.pos[pos_index(p, c, it, 0, g)];
      const double dy = p.pos[pos_index(p, c, it, 1, j)] - p.pos[pos_index(p, c, it, 1, g)];
      const double dz = p.pos[pos_index(p, c, it, 2, j)] - p.pos[pos_index(p, c, it, 2, g)];
      double rr, invr;
      fast_r_invr(dx * dx + dy * dy + dz * dz, rr, invr);
      const int ir = lpot_index(rr - rmin, p.inv_dr2d, p.dr2d, p.rs2d);          // LPot2D radial cell, mc_poten.cc:696-701
      const double2 gr = x.t.rgi2d[ir];
      const double dr = (rr - gr.x) * gr.y;
      const double w = __longlong_as_double((__double_as_longlong(dr) & ~0xfffLL) | (long long)ir);      // ir rides in the low mantissa bits
      *reinterpret_cast<double4 *>(gb + 4 * (size_t)k) = make_double4(dx * invr, dy * invr, dz * invr, w);
   }
}

// NB cached items against one orientation n: cos(theta) = n.u, angular cell, the NB cell gathers, the bilinear forms
#ifndef PIMC_GEO_BATCH
#define PIMC_GEO_BATCH 7
#endif
template <int NB>
__device__ __forceinline__ double geo_evaln(const Params &p, const SmallTables &t, double cmin, double n0, double n1, double n2,
                                            const double *ux, const double *uy, const double *uz, const double *dr, const int *ib, const bool *ok)
{
   double cs[NB];
   int ic[NB];
   #pragma unroll
   for (int u = 0; u < NB; u++) {
      cs[u] = n0 * ux[u] + n1 * uy[u] + n2 * uz[u];
      ic[u] = lpot_index(cs[u] - cmin, p.inv_dc2d, p.dc2d, p.cs2d);            // LPot2D angular cell, mc_poten.cc:702-704
   }
   double y1[NB], y2[NB], y3[NB], y4[NB];
   if (p.cell4_on) {
      if (p.cell_hint == 2) {
         #pragma unroll
         for (int u = 0; u < NB; u++) load_cell4_stream(p.cell4 + 4 * (size_t)(ib[u] + ic[u]), y1[u], y2[u], y3[u], y4[u]);
      } else if (p.cell_hint) {
         #pragma unroll
         for (int u = 0; u < NB; u++) load_cell4_keep(p.cell4 + 4 * (size_t)(ib[u] + ic[u]), y1[u], y2[u], y3[u], y4[u]);
      } else {
         #pragma unroll
         for (int u = 0; u < NB; u++) load_cell4(p.cell4 + 4 * (size_t)(ib[u] + ic[u]), y1[u], y2[u], y3[u], y4[u]);
      }
   } else {
      #pragma unroll
      for (int u = 0; u < NB; u++) load_cell(p.cell2d + (size_t)(ib[u] + ic[u]), y1[u], y2[u], y3[u], y4[u]);
   }
   double v = 0.0;
   #pragma unroll
   for (int u = 0; u < NB; u++) {
      const double2 gc = lds_d2(t.s_cgi2d, ic[u]);
      const double dc = (cs[u] - gc.x) * gc.y;
      const double lo = y1[u] + dr[u] * (y2[u] - y1[u]);         // V at cos(theta)_ic, interpolated in r
      const double hi = y4[u] + dr[u] * (y3[u] - y4[u]);         // V at cos(theta)_ic+1
      const double e = lo + dc * (hi - lo);
      v += ok[u] ? e : 0.0;
   }
   return v;
}

// sum over the cached items of rot slice q of LPot2D(r, cos(theta) = n.u) for ONE orientation (NO = 1) or for the
// proposed and the current orientation (NO = 2: the geometry loads are shared).  PIMC_GEO_BATCH items per lane in flight:
// the sums of a slice are two dependent memory round trips per batch, so fewer, larger batches shorten the critical
// path of the slice (12.5 items per lane on C5: two batches of 7).
template <int NO>
__device__ __forceinline__ void rot_potential_cached(const Params &p, Ctx &x, int q, const double *o0, const double *o1, double *vout)
{
   constexpr int NB = (NO == 1) ? PIMC_GEO_BATCH : 4;
   const int c = x.c, G = x.G, n = p.geo_items, gn = p.geo_n;
   const double *gb = p.geo + ((size_t)c * p.Q + q) * 4 * gn;
   const int rowlen = p.cell4_on ? p.cs2d - 1 : p.cs2d;
   const double cmin = x.t.cgi2d[0].x;
   const double a0 = o0[0], a1 = o0[1], a2 = o0[2];
   double b0 = 0.0, b1 = 0.0, b2 = 0.0;
   if (NO == 2) { b0 = o1[0]; b1 = o1[1]; b2 = o1[2]; }
   double v0 = 0.0, v1 = 0.0;
   for (int k0 = x.gl; k0 < n; k0 += NB * G) {
      double ux[NB], uy[NB], uz[NB], dr[NB];
      int ib[NB];
      bool ok[NB];
      #pragma unroll
      for (int u = 0; u < NB; u++) {
         const int k = k0 + u * G;
         ok[u] = k < n;
         const int kk = ok[u] ? k : k0;                 // masked slots repeat the first item of the batch (a valid look-up)
         if (p.geo_hint == 1) load_geo4_keep(gb + 4 * (size_t)kk, ux[u], uy[u], uz[u], dr[u]);
         else if (p.geo_hint == 2) load_geo4_plain(gb + 4 * (size_t)kk, ux[u], uy[u], uz[u], dr[u]);
         else load_geo4(gb + 4 * (size_t)kk, ux[u], uy[u], uz[u], dr[u]);
      }
      #pragma unroll
      for (int u = 0; u < NB; u++) {
         const long long bits = __double_as_longlong(dr[u]);
         ib[u] = (int)(bits & 0xfffLL) * rowlen;        // row offset of the radial cell in the table
         dr[u] = __longlong_as_double(bits & ~0xfffLL);
      }
#ifdef PIMC_TIMELINE
      asm volatile("" ::"d"(dr[NB - 1]), "d"(ux[0]));
      MARK(x, 30);
#endif
      v0 += geo_evaln<NB>(p, x.t, cmin, a0, a1, a2, ux, uy, uz, dr, ib, ok);
#ifdef PIMC_TIMELINE
      asm volatile("" ::"d"(v0));
      MARK(x, 31);
#endif
      if (NO == 2) v1 += geo_evaln<NB>(p, x.t, cmin, b0, b1, b2, ux, uy, uz, dr, ib, ok);
   }
   vout[0] = v0;
   if (NO == 2) vout[1] = v1;
}

// proposal of a rotational step (mc_piqmc.cc:950-995 top, :796-814 linear): angles in (cost, phi, chi), the
// orientation (KIND 2: rotation matrix, 9 doubles; KIND 1: unit vector) in `o`
template <int KIND>
__device__ __forceinline__ void rot_propose(const Params &p, int type, double r1, double r2, double r3, double &cost, double &phi, double &chi, double *o)
{
   const double step = p.rtstep[type];
   cost += step * (r1 - 0.5);
   if ((KIND & 3) == 2) {
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
   if ((KIND & 3) == 2) {
      Mat3 Rn;
      matpre(phi, acos(cost), chi, Rn);
      #pragma unroll
      for (int i = 0; i < 9; i++) o[i] = Rn.m[i / 3][i % 3];
   } else {
      const double sint = sqrt(1.0 - cost * cost);
      double sp_, cp_;
      sincos(phi, &sp_, &cp_);
      o[0] = sint * cp_; o[1] = sint * sp_; o[2] = cost;
   }
}

// acceptance of a rotational step from the four density factors rho = {(q0->cur), (cur->q2), (q0->new), (new->q2)} and
// the potential sums (mc_piqmc.cc:895-921, 1145-1178); *bad collects the reference's fatal conditions
template <int KIND>
__device__ __forceinline__ bool rot_accept(const Params &p, const double *rho, double vnew, double vold, double u4, int *bad)
{
   if (p.rotden_type == 0) {
      double dens_old = rho[0] * rho[1], dens_new = rho[2] * rho[3];
      if (fabs(dens_old) < RZERO) dens_old = 0.0;
      if (fabs(dens_new) < RZERO) dens_new = 0.0;
      if ((KIND & 3) == 2) { dens_old = fabs(dens_old); dens_new = fabs(dens_new); }
      else if (dens_old < 0.0 || dens_new < 0.0) *bad |= 2;     // "Negative rot density" is fatal in the reference
      double rd = (dens_old > RZERO) ? dens_new / dens_old : 1.0;
      rd *= exp(-p.tau * (vnew - vold));
      return (rd > 1.0) || (rd > u4);
   }
   // rattle-and-shake propagator: the acceptance works on exponents (mc_piqmc.cc:903-921, 1152-1178)
   double rd;
   if ((KIND & 3) == 2) rd = ((rho[2] + rho[3]) - (rho[0] + rho[1])) / (4.0 * (p.rottau / WNO2K));
   else {
      double dens_old = rho[0] + rho[1], dens_new = rho[2] + rho[3];
      if (fabs(dens_old) < RZERO) dens_old = 0.0;
      if (fabs(dens_new) < RZERO) dens_new = 0.0;
      rd = dens_new - dens_old;
   }
   rd -= p.tau * (vnew - vold);
   return (rd > 0.0) || (rd > log(u4));
}

// density factor i of a step at slice q: i = 0 rho(q0 -> cur), 1 rho(cur -> q2), 2 rho(q0 -> new), 3 rho(new -> q2);
// cur / nw are the orientations of the slice (matrix or unit vector)
// nb0 / nb2 (optional): the neighbours' current rotation matrices when the caller holds them (all slices of a one-CTA chain
// live in the CTA's slots) -- otherwise they are rebuilt from the stored Euler angles (three sincos and an acos each)
template <int KIND>
__device__ __forceinline__ double rot_density(const Params &p, const SmallTables &t, int c, int q0, int q2, int m, int i, const double *cur, const double *nw, int *bad,
                                              const double *nb0 = nullptr, const double *nb2 = nullptr)
{
   const double *mid = (i < 2) ? cur : nw;
   if ((KIND & 3) == 2) {
      Mat3 A, B;
      if ((i == 0 || i == 2) && nb0) {
         #pragma unroll
         for (int k = 0; k < 9; k++) { A.m[k / 3][k % 3] = nb0[k]; B.m[k / 3][k % 3] = mid[k]; }
      } else if (!(i == 0 || i == 2) && nb2) {
         #pragma unroll
         for (int k = 0; k < 9; k++) { A.m[k / 3][k % 3] = mid[k]; B.m[k / 3][k % 3] = nb2[k]; }
      } else
      if (i == 0 || i == 2) {
         load_rotmat(p, c, q0, m, A);
         #pragma unroll
         for (int k = 0; k < 9; k++) B.m[k / 3][k % 3] = mid[k];
      } else {
         #pragma unroll
         for (int k = 0; k < 9; k++) A.m[k / 3][k % 3] = mid[k];
         load_rotmat(p, c, q2, m, B);
      }
      if (p.rotden_type == 1) return rsrot(p, A, B, nullptr);
      int istop = 0;
      const double r = rotden(p, A, B, nullptr, nullptr, nullptr, nullptr, &istop);
      if (istop) *bad |= 1;
      return r;
   }
   const int qq = (i == 0 || i == 2) ? q0 : q2;
   double dot = 0.0;
   #pragma unroll
   for (int d = 0; d < 3; d++) {
      double nb = p.cosn[ang_index(p, c, qq, d, m)];
      dot += (i == 0 || i == 2) ? nb * mid[d] : mid[d] * nb;
   }
   return (p.rotden_type == 1) ? rsline(p, dot, nullptr) : srotdens(p, t, dot);
}

// state update of an accepted step (mc_piqmc.cc:923-935, 1180-1198)
template <int KIND>
__device__ __forceinline__ void rot_commit(const Params &p, int c, int q, int m, double cost, double phi, double chi)
{
   p.ang[ang_index(p, c, q, 1, m)] = cost;
   p.ang[ang_index(p, c, q, 0, m)] = phi;
   if ((KIND & 3) == 2) p.ang[ang_index(p, c, q, 2, m)] = chi;
   const double sint = sqrt(1.0 - cost * cost);
   double sp_, cp_;
   sincos(phi, &sp_, &cp_);
   p.cosn[ang_index(p, c, q, 0, m)] = sint * cp_;
   p.cosn[ang_index(p, c, q, 1, m)] = sint * sp_;
   p.cosn[ang_index(p, c, q, 2, m)] = cost;
}

template <int KIND>
__device__ void rot_step(const Params &p, Ctx &x, int type, int q, int m, bool active, int *err)
{
   const int c = x.c, Q = p.Q, G = x.G;
   const int g = p.first[type] + m;
   RotSlot own;
   RotSlot *sl = (G == 1) ? &own : x.slot;
   int q0 = q - 1, q2 = q + 1;
   if (q0 < 0) q0 += Q;
   if (q2 >= Q) q2 -= Q;
   double *vcache = p.vold + ((size_t)c * Q + q) * p.NMpad + m;
   int *vep = p.vepoch + ((size_t)c * Q + q) * p.NMpad + m;

   MARK(x, 1);
   if (active && x.gl == 0) {
      uint32_t *sp = stream_ptr(p, c, p.P + q);
      Mrg rs;
      mrg_load(rs, sp);
      double r1 = mrg_u01(rs), r2 = mrg_u01(rs), r3 = mrg_u01(rs), r4 = r3;
      if ((KIND & 3) == 2) r4 = mrg_u01(rs);
      mrg_store(rs, sp);
      double cost = p.ang[ang_index(p, c, q, 1, m)], phi = p.ang[ang_index(p, c, q, 0, m)], chi = p.ang[ang_index(p, c, q, 2, m)];
      if ((KIND & 3) == 2) {
         Mat3 R1;
         matpre(phi, acos(cost), chi, R1);
         #pragma unroll
         for (int i = 0; i < 9; i++) sl->b[i] = R1.m[i / 3][i % 3];
      } else {
         #pragma unroll
         for (int d = 0; d < 3; d++) sl->b[d] = p.cosn[ang_index(p, c, q, d, m)];
      }
      rot_propose<KIND>(p, type, r1, r2, r3, cost, phi, chi, sl->a);
      sl->u4 = r4; sl->cost = cost; sl->phi = phi; sl->chi = chi;
      sl->need_old = (*vep != p.pos_epoch[c]) ? 1 : 0;
      sl->bad = 0;
   }
   MARK(x, 2);
   group_sync(x);
   MARK(x, 3);

   double vnew = 0.0, vold = 0.0;
   if (active) {
      for (int i = x.gl; i < 4; i += G) {
         int bad = 0;
         sl->rho[i] = rot_density<KIND>(p, x.t, c, q0, q2, m, i, sl->b, sl->a, &bad);
         if (bad) { if (G == 1) sl->bad |= bad; else atomicOr(&sl->bad, bad); }
      }
      MARK(x, 4);
      vnew = rot_potential<KIND>(p, x, g, q, sl->a);
      if (sl->need_old) vold = rot_potential<KIND>(p, x, g, q, sl->b);
   }
   MARK(x, 5);
   // group reduction in a fixed order
   const int gw = (G < 32) ? G : 32;
   vnew = team_sum(vnew, gw);
   vold = team_sum(vold, gw);
   if (G > 32 && (x.tid & 31) == 0) { x.part[2 * (x.gl >> 5)] = vnew; x.part[2 * (x.gl >> 5) + 1] = vold; }
   group_sync(x);
   MARK(x, 6);

   if (active && x.gl == 0) {
      if (G > 32) {
         vnew = 0.0; vold = 0.0;
         for (int w = 0; w < (G >> 5); w++) { vnew += x.part[2 * w]; vold += x.part[2 * w + 1]; }
      }
      if (!sl->need_old) vold = *vcache;
      int bad = sl->bad;
      bool acc = rot_accept<KIND>(p, sl->rho, vnew, vold, sl->u4, &bad);
      if (bad) { acc = false; atomicOr(err, bad); }
      double *cn = counter_ptr(p, c, type, 2);
      atomicAdd(cn, 1.0);
      *vcache = acc ? vnew : vold;
      *vep = p.pos_epoch[c];
      if (acc) {
         atomicAdd(cn + 1, 1.0);
         rot_commit<KIND>(p, c, q, m, sl->cost, sl->phi, sl->chi);
         if ((KIND & 3) == 2)
            for (int mm = 0; mm < p.NM; mm++)            // partner rotors at this slice see a new neighbour
               if (mm != m) p.vepoch[((size_t)c * Q + q) * p.NMpad + mm] = -1;
      }
   }
   MARK(x, 7);
   group_sync(x);
   MARK(x, 8);
}

// even slices, then odd slices (MCRotations3D mc_piqmc.cc:737-771, MCRotationsMove :490-512);
// an odd slice count gets a third phase for the last slice.  CTA `crank` of the chain owns the
// slices idx = s*cpc + crank of a phase; its rot groups take them ngrp at a time.
template <int KIND>
__device__ void rot_sweep(const Params &p, Ctx &x, int type, int *err)
{
   const int Q = p.Q;
   const int qe = (Q % 2 == 1 && Q > 1) ? Q - 1 : Q;
   for (int phase = 0; phase < 3; phase++) {
      int count = (phase == 0) ? (qe + 1) / 2 : (phase == 1 ? qe / 2 : (qe != Q ? 1 : 0));
      if (count == 0) continue;
      int S = (count + p.cpc - 1) / p.cpc;
      int nrounds = (S + x.ngrp - 1) / x.ngrp;
      for (int rd = 0; rd < nrounds; rd++) {
         int s = rd * x.ngrp + x.grp;
         int idx = s * p.cpc + x.crank;
         bool active = s < S && idx < count;
         int q = (phase == 0) ? 2 * idx : (phase == 1 ? 2 * idx + 1 : Q - 1);
         if (!active) q = 0;
         for (int m = 0; m < p.numb[type]; m++) rot_step<KIND>(p, x, type, q, m, active, err);
      }
      MARK(x, 9);
      chain_sync(p, x);
      MARK(x, 10);
   }
}

// ---------------------------------------------------------------------------------------------
// rotational sweep of a ONE-rotor system with an even number of rot slices, software-pipelined over the
// time steps.  The potential part of a proposal depends on its own slice only (the pair action is diagonal
// in imaginary time), the accept/reject decision couples a slice to its two neighbours through the density
// matrix.  Per sweep every rot group first evaluates the potential sums of ALL slices its CTA owns, then
//   CTAs of even slices: wait until the odd decisions of the PREVIOUS sweep are published, decide, publish;
//   CTAs of odd slices : wait until the even decisions of THIS sweep are published, decide, publish
// (per-chain arrival counters in global memory, split arrive / wait).  An even-slice CTA therefore starts the
// sums of the next time step while the odd slices are still being decided: the chain-wide barriers of the
// two-phase sweep disappear from the critical path.  Draws per slice and step are those of rot_step, the order
// of decisions (all even, then all odd) is unchanged, so the trajectory is that of the two-phase sweep.
// A chain that lives in one CTA (cpc = 1) runs sums | even decisions | odd decisions with CTA barriers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rot_arrive(const Params &p, Ctx &x, int parity)
{
   __syncthreads();
   if (x.tid == 0) {
      __threadfence();
      atomicAdd(p.barrier + (size_t)x.c * 32 + 8 + 8 * parity, 1u);
   }
}
__device__ __forceinline__ void rot_wait(const Params &p, Ctx &x, int parity, unsigned target)
{
   if (x.tid == 0) {
      unsigned *b = p.barrier + (size_t)x.c * 32 + 8 + 8 * parity, v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(b) : "memory"); } while ((int)(v - target) < 0);
      __threadfence();
   }
   __syncthreads();
}

// decisions of the owned slices of parity `par` (par < 0: all owned slices)
template <int KIND>
__device__ __forceinline__ void rot_decide_owned(const Params &p, Ctx &x, int type, int par, int *err)
{
   const int c = x.c, Q = p.Q, G = x.G, m = 0;
   const int nown = (Q + p.cpc - 1) / p.cpc;
   const int nrounds = (nown + x.ngrp - 1) / x.ngrp;
   for (int rd = 0; rd < nrounds; rd++) {
      const int ls = rd * x.ngrp + x.grp;
      const int q = ls * p.cpc + x.crank;
      const bool active = ls < nown && q < Q && (par < 0 || (q & 1) == par);
      RotSlot *sl = x.slot + (ls < nown ? ls : 0);
      int q0 = q - 1, q2 = q + 1;
      if (q0 < 0) q0 += Q;
      if (q2 >= Q) q2 -= Q;
      // a chain that lives in one CTA holds every slice's current matrix in its slots: the neighbours' are read, not rebuilt
      const double *nb0 = (p.cpc == 1 && (KIND & 3) == 2) ? x.slot[q0].b : nullptr, *nb2 = (p.cpc == 1 && (KIND & 3) == 2) ? x.slot[q2].b : nullptr;
      if (active)
         for (int i = x.gl; i < 4; i += G) {
            int bad = 0;
            sl->rho[i] = rot_density<KIND>(p, x.t, c, q0, q2, m, i, sl->b, sl->a, &bad, nb0, nb2);
            if (bad) { if (G == 1) sl->bad |= bad; else atomicOr(&sl->bad, bad); }
         }
      group_sync(x);
      if (active && x.gl == 0) {
         int bad = sl->bad;
         const double vnew = sl->vnew, vold = sl->vold;
         bool acc = rot_accept<KIND>(p, sl->rho, vnew, vold, sl->u4, &bad);
         if (bad) { acc = false; atomicOr(err, bad); }
         double *cn = counter_ptr(p, c, type, 2);
         atomicAdd(cn, 1.0);
         sl->vcache = acc ? vnew : vold;
         sl->vep = sl->epoch;
         if (acc) {
            atomicAdd(cn + 1, 1.0);
            rot_commit<KIND>(p, c, q, m, sl->cost, sl->phi, sl->chi);
            sl->cur[0] = sl->cost; sl->cur[1] = sl->phi; sl->cur[2] = sl->chi;
            #pragma unroll
            for (int i = 0; i < ((KIND & 3) == 2 ? 9 : 3); i++) sl->b[i] = sl->a[i];
         }
      }
      group_sync(x);
   }
}

template <int KIND>
__device__ void rot_sweep_pipe(const Params &p, Ctx &x, int type, int *err)
{
   const int c = x.c, Q = p.Q, G = x.G, m = 0;
   const int g = p.first[type];
   const int nown = (Q + p.cpc - 1) / p.cpc;
   const int nrounds = (nown + x.ngrp - 1) / x.ngrp;
   MARK(x, 1);
   for (int rd = 0; rd < nrounds; rd++) {
      const int ls = rd * x.ngrp + x.grp;
      const int q = ls * p.cpc + x.crank;
      const bool active = ls < nown && q < Q;
      RotSlot *sl = x.slot + (ls < nown ? ls : 0);
      if (active && x.gl == 0) {
         Mrg rs;
         mrg_load(rs, x.rrng + ls * 6);
         double r1 = mrg_u01(rs), r2 = mrg_u01(rs), r3 = mrg_u01(rs), r4 = r3;
         if ((KIND & 3) == 2) r4 = mrg_u01(rs);
         mrg_store(rs, x.rrng + ls * 6);
         const int epoch = p.pos_epoch[c];
         double cost = sl->cur[0], phi = sl->cur[1], chi = sl->cur[2];     // current state and orientation live in the slot
         rot_propose<KIND>(p, type, r1, r2, r3, cost, phi, chi, sl->a);
         sl->u4 = r4; sl->cost = cost; sl->phi = phi; sl->chi = chi;
         sl->epoch = epoch;
         sl->need_old = (sl->vep != epoch) ? 1 : 0;
         sl->bad = 0;
      }
      MARK(x, 2);
      group_sync(x);
      MARK(x, 3);
      double vnew = 0.0, vold = 0.0;
      if (active) {
         if ((KIND & 7) == 1 && p.geo_on) {
            // linear rotor among atoms: distance part of every term from the cache, refilled after a translational sweep
            const bool refill = sl->gep != sl->epoch;
            if (refill) geo_fill(p, x, g, q);
            double vv[2];
            if (sl->need_old) { rot_potential_cached<2>(p, x, q, sl->a, sl->b, vv); vnew = vv[0]; vold = vv[1]; }
            else { rot_potential_cached<1>(p, x, q, sl->a, nullptr, vv); vnew = vv[0]; }
         } else {
            vnew = rot_potential<KIND>(p, x, g, q, sl->a);
            if (sl->need_old) vold = rot_potential<KIND>(p, x, g, q, sl->b);
         }
      }
      MARK(x, 4);
      const int gw = (G < 32) ? G : 32;
      vnew = team_sum(vnew, gw);
      vold = team_sum(vold, gw);
      if (G > 32 && (x.tid & 31) == 0) { x.part[2 * (x.gl >> 5)] = vnew; x.part[2 * (x.gl >> 5) + 1] = vold; }
      group_sync(x);
      if (active && x.gl == 0) {
         if (G > 32) {
            vnew = 0.0; vold = 0.0;
            for (int w = 0; w < (G >> 5); w++) { vnew += x.part[2 * w]; vold += x.part[2 * w + 1]; }
         }
         if (!sl->need_old) vold = sl->vcache;
         sl->vnew = vnew; sl->vold = vold;
         sl->gep = sl->epoch;
      }
      group_sync(x);
   }
   MARK(x, 5);
   if (p.cpc == 1) {
      __syncthreads();
      rot_decide_owned<KIND>(p, x, type, 0, err);
      __syncthreads();
      rot_decide_owned<KIND>(p, x, type, 1, err);
      __syncthreads();
   } else {
      const int par = x.crank & 1, half = p.cpc >> 1;
      if (par == 0) rot_wait(p, x, 1, (unsigned)(x.rot_iter * half));              // odd decisions of the previous sweep
      else rot_wait(p, x, 0, (unsigned)((x.rot_iter + 1) * half));                 // even decisions of this sweep
      MARK(x, 6);
      rot_decide_owned<KIND>(p, x, type, -1, err);
      MARK(x, 9);
      rot_arrive(p, x, par);
   }
   x.rot_iter++;
   MARK(x, 10);
}

// rot slice owned by local slot ls of this CTA: a contiguous block for the free-running sweeps, round-robin otherwise
__device__ __forceinline__ int owned_slice(const Params &p, const Ctx &x, int ls)
{
   return p.rot_run ? x.crank * (p.Q / p.cpc) + ls : ls * p.cpc + x.crank;
}

// ---------------------------------------------------------------------------------------------
// `nrun` consecutive rotational sweeps of a linear rotor with no translational sweep in between, free-running: every
// rot group owns ONE slice for the whole run and loops  propose -> potential sums (gathers) -> wait for the two
// neighbouring slices -> four density factors, accept, commit -> post  on its own.  An even slice's n-th decision needs
// the (n-1)-th decisions of its odd neighbours, an odd slice's n-th decision the n-th of its even neighbours: the order
// "all even, then all odd" of every sweep is kept, but no barrier couples slices that do not depend on each other, so the
// gathers of the slices that may run ahead fill the time the others spend in their (latency-bound) decisions.
// Hand-over through global memory: the deciding thread commits (phi, cos theta, n) and then stores the slice's decision
// count with release semantics; a waiting group polls its two neighbours with acquire loads and reads their axes from L2.
// Draws, proposals and acceptance are those of rot_step / rot_sweep_pipe.
// ---------------------------------------------------------------------------------------------
// publishes the axis of a slice after `done` sweeps of this launch: six 64-bit words {done + 1 : 32 bits of the axis}.
// Every word certifies its own payload (a 64-bit store is single-copy atomic), so neither side needs a fence; a reader
// cannot see words of two versions because the publisher's next decision waits for that reader's.
__device__ __forceinline__ void rot_ll_publish(unsigned long long *dst, const double *axis, int done)
{
   const unsigned long long tag = (unsigned long long)(unsigned)(done + 1) << 32;
   #pragma unroll
   for (int d = 0; d < 3; d++) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(axis[d]);
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(dst + 2 * d), "l"(tag | (bits & 0xffffffffULL)) : "memory");
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(dst + 2 * d + 1), "l"(tag | (bits >> 32)) : "memory");
   }
}

template <int KIND>
__device__ void rot_run(const Params &p, Ctx &x, int type, int nrun, int *err)
{
   const int c = x.c, Q = p.Q, G = x.G, m = 0;
   const int g = p.first[type];
   const int nown = Q / p.cpc, ls = x.grp;
   const bool active = ls < nown;
   const int q = x.crank * nown + (active ? ls : 0);
   RotSlot *sl = x.slot + (active ? ls : 0);
   int q0 = q - 1, q2 = q + 1;
   if (q0 < 0) q0 += Q;
   if (q2 >= Q) q2 -= Q;
   unsigned long long *ll = p.rot_ll + (size_t)c * Q * 8;
   // the geometry cache follows the positions: stale after a translational sweep (the slot's gep is set by the leader's
   // first decision; the chain barrier between sweeps orders it with this read)
   if (active && p.geo_on && sl->gep != p.pos_epoch[c]) geo_fill(p, x, g, q);
   if (active)
   for (int it = 0; it < nrun; it++) {
      const int n = x.rot_iter + it;                     // sweeps this slice has completed in this launch
      MARK(x, 1);
      if (x.gl == 0) {
         Mrg rs;
         mrg_load(rs, x.rrng + ls * 6);
         const double r1 = mrg_u01(rs), r2 = mrg_u01(rs), r3 = mrg_u01(rs);
         mrg_store(rs, x.rrng + ls * 6);
         const int epoch = p.pos_epoch[c];
         double cost = sl->cur[0], phi = sl->cur[1], chi = sl->cur[2];
         rot_propose<KIND>(p, type, r1, r2, r3, cost, phi, chi, sl->a);
         sl->u4 = r3; sl->cost = cost; sl->phi = phi; sl->chi = chi;
         sl->epoch = epoch;
         sl->need_old = (sl->vep != epoch) ? 1 : 0;
         sl->bad = 0;
      }
      MARK(x, 2);
      group_sync(x);
      double vnew = 0.0, vold = 0.0;
      if (p.geo_on) {
         double vv[2];
         if (sl->need_old) { rot_potential_cached<2>(p, x, q, sl->a, sl->b, vv); vnew = vv[0]; vold = vv[1]; }
         else { rot_potential_cached<1>(p, x, q, sl->a, nullptr, vv); vnew = vv[0]; }
      } else {
         vnew = rot_potential<KIND>(p, x, g, q, sl->a);
         if (sl->need_old) vold = rot_potential<KIND>(p, x, g, q, sl->b);
      }
      const int gw = (G < 32) ? G : 32;
      for (int o = gw >> 1; o > 0; o >>= 1) { vnew += __shfl_xor_sync(x.gmask, vnew, o); vold += __shfl_xor_sync(x.gmask, vold, o); }
      if (G > 32 && (x.tid & 31) == 0) { x.part[2 * (x.gl >> 5)] = vnew; x.part[2 * (x.gl >> 5) + 1] = vold; }
      MARK(x, 4);
      // the two neighbours' axes: odd slices need this sweep's even decisions, even slices the previous sweep's odd ones.
      // Twelve lanes poll one 64-bit word each (32 bits of payload under the publisher's sweep count): the payload is valid
      // the moment its own tag is, one L2 round trip, no fence on either side.
      {
         const unsigned need = (unsigned)((q & 1) ? n + 1 : n) + 1u;
         for (int w = x.gl; w < 12; w += G) {
            const unsigned long long *src = ll + (size_t)(w < 6 ? q0 : q2) * 8 + (w < 6 ? w : w - 6);
            unsigned long long v;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory"); } while ((unsigned)(v >> 32) < need);
            sl->nbw[w] = (unsigned)v;
         }
      }
      group_sync(x);
      MARK(x, 6);
      // density factors rho(q0 -> cur), rho(cur -> q2), rho(q0 -> new), rho(new -> q2)
      if (x.gl < 4) {
         const int i = x.gl;
         const double *mid = (i < 2) ? sl->b : sl->a;
         const double *nbv = reinterpret_cast<const double *>(sl->nbw) + ((i == 0 || i == 2) ? 0 : 3);
         double dot = 0.0;
         #pragma unroll
         for (int d = 0; d < 3; d++) {
            const double nb = nbv[d];
            dot += (i == 0 || i == 2) ? nb * mid[d] : mid[d] * nb;
         }
         sl->rho[i] = (p.rotden_type == 1) ? rsline(p, dot, nullptr) : srotdens(p, x.t, dot);
      }
      group_sync(x);
      MARK(x, 7);
      if (x.gl == 0) {
         if (G > 32) {
            vnew = 0.0; vold = 0.0;
            for (int w = 0; w < (G >> 5); w++) { vnew += x.part[2 * w]; vold += x.part[2 * w + 1]; }
         }
         if (!sl->need_old) vold = sl->vcache;
         int bad = 0;
         bool acc = rot_accept<KIND>(p, sl->rho, vnew, vold, sl->u4, &bad);
         if (bad) { acc = false; atomicOr(err, bad); }
         double *cn = counter_ptr(p, c, type, 2);
         sl->vcache = acc ? vnew : vold;
         sl->vep = sl->epoch;
         sl->gep = sl->epoch;
         // hand the slice's axis to its neighbours first (rot_ll_publish): for a linear rotor the committed axis is the
         // proposal's (same expressions, same inputs: sl->a), no second sincos / sqrt
         rot_ll_publish(ll + (size_t)q * 8, acc ? sl->a : sl->b, n + 1);
         if (acc) {
            p.ang[ang_index(p, c, q, 1, m)] = sl->cost;
            p.ang[ang_index(p, c, q, 0, m)] = sl->phi;
            #pragma unroll
            for (int d = 0; d < 3; d++) { const double nd = sl->a[d]; p.cosn[ang_index(p, c, q, d, m)] = nd; sl->b[d] = nd; }
            sl->cur[0] = sl->cost; sl->cur[1] = sl->phi; sl->cur[2] = sl->chi;
         }
         atomicAdd(cn, 1.0);
         if (acc) atomicAdd(cn + 1, 1.0);
      }
      MARK(x, 9);
   }
   x.rot_iter += nrun;
   chain_sync(p, x);
}

// ---------------------------------------------------------------------------------------------
// The same free-running stretch for a top whose chain lives in ONE CTA (C1-C3): every slice has its rot group, its slot in
// shared memory holds the current rotation matrix, and the hand-over is a per-slice sweep count next to the slots
// (volatile shared memory + __threadfence_block).  rot_sweep_pipe runs such a chain in three CTA-wide stages per sweep
// (sums | even decisions | odd decisions, a barrier after each: every stage as slow as its slowest slice and nothing
// overlapping); here a slice's proposal and sums run while its neighbours decide.  Draws, sums, the order of the
// decisions seen by every slice, acceptance and commit are those of rot_sweep_pipe.
// ---------------------------------------------------------------------------------------------
template <int KIND>
__device__ void rot_run_cta(const Params &p, Ctx &x, int type, int nrun, int *err)
{
   const int c = x.c, Q = p.Q, G = x.G, m = 0;
   const int g = p.first[type];
   // the even slices go to the first half of the rot groups, the odd ones to the second: the groups that share a warp are
   // in the same phase (a polling group and a working one in one warp take turns at the warp's issue slots)
   const int ls = x.grp;
   const bool active = ls < Q;
   const int q = active ? ((ls < Q / 2) ? 2 * ls : 2 * (ls - Q / 2) + 1) : 0;
   RotSlot *sl = x.slot + q;
   int q0 = q - 1, q2 = q + 1;
   if (q0 < 0) q0 += Q;
   if (q2 >= Q) q2 -= Q;
   volatile int *done = x.rot_done;
   if (active)
   for (int it = 0; it < nrun; it++) {
      const int n = x.rot_iter + it;
      if (x.gl == 0) {
         Mrg rs;
         mrg_load(rs, x.rrng + q * 6);
         double r1 = mrg_u01(rs), r2 = mrg_u01(rs), r3 = mrg_u01(rs), r4 = r3;
         if ((KIND & 3) == 2) r4 = mrg_u01(rs);
         mrg_store(rs, x.rrng + q * 6);
         const int epoch = p.pos_epoch[c];
         double cost = sl->cur[0], phi = sl->cur[1], chi = sl->cur[2];
         rot_propose<KIND>(p, type, r1, r2, r3, cost, phi, chi, sl->a);
         sl->u4 = r4; sl->cost = cost; sl->phi = phi; sl->chi = chi;
         sl->epoch = epoch;
         sl->need_old = (sl->vep != epoch) ? 1 : 0;
         sl->bad = 0;
      }
      group_sync(x);
      double vnew = rot_potential<KIND>(p, x, g, q, sl->a), vold = 0.0;
      if (sl->need_old) vold = rot_potential<KIND>(p, x, g, q, sl->b);
      const int gw = (G < 32) ? G : 32;
      for (int o = gw >> 1; o > 0; o >>= 1) { vnew += __shfl_xor_sync(x.gmask, vnew, o); vold += __shfl_xor_sync(x.gmask, vold, o); }
      if (G > 32 && (x.tid & 31) == 0) { x.part[2 * (x.gl >> 5)] = vnew; x.part[2 * (x.gl >> 5) + 1] = vold; }
      // odd slices need this sweep's even decisions, even slices the previous sweep's odd ones
      if (x.gl < 2) {
         const int target = (q & 1) ? n + 1 : n;
         const volatile int *f = done + (x.gl == 0 ? q0 : q2);
         while (*f < target) { }
      }
      // groups that share a warp go on together: lanes that left the poll at different times would otherwise run the rest
      // of the sweep as separate paths of the warp, one after the other
      if (G < 32) __syncwarp(); else group_sync(x);
      __threadfence_block();
      for (int i = x.gl; i < 4; i += G) {
         int bad = 0;
         sl->rho[i] = rot_density<KIND>(p, x.t, c, q0, q2, m, i, sl->b, sl->a, &bad, x.slot[q0].b, x.slot[q2].b);
         if (bad) { if (G == 1) sl->bad |= bad; else atomicOr(&sl->bad, bad); }
      }
      group_sync(x);
      if (x.gl == 0) {
         if (G > 32) {
            vnew = 0.0; vold = 0.0;
            for (int w = 0; w < (G >> 5); w++) { vnew += x.part[2 * w]; vold += x.part[2 * w + 1]; }
         }
         if (!sl->need_old) vold = sl->vcache;
         int bad = sl->bad;
         bool acc = rot_accept<KIND>(p, sl->rho, vnew, vold, sl->u4, &bad);
         if (bad) { acc = false; atomicOr(err, bad); }
         sl->vcache = acc ? vnew : vold;
         sl->vep = sl->epoch;
         if (acc) {
            #pragma unroll
            for (int i = 0; i < ((KIND & 3) == 2 ? 9 : 3); i++) sl->b[i] = sl->a[i];
            sl->cur[0] = sl->cost; sl->cur[1] = sl->phi; sl->cur[2] = sl->chi;
         }
         __threadfence_block();
         done[q] = n + 1;                                 // the neighbours may go on; the global copy of the angles follows
         double *cn = counter_ptr(p, c, type, 2);
         atomicAdd(cn, 1.0);
         if (acc) {
            atomicAdd(cn + 1, 1.0);
            rot_commit<KIND>(p, c, q, m, sl->cost, sl->phi, sl->chi);
         }
      }
   }
   x.rot_iter += nrun;
   __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_tables(const Params &p, SmallTables &t, double *&cursor)
{
   // copies the packed 1-D spline records and their bucket table into shared memory; `cursor` walks the dynamic
   // smem block.  The linear-rotor density spline is touched four times per rot step and stays in global/L1.
   t.g1d = p.g1d; t.v1d = p.v1d; t.y2_1d = p.y2_1d; t.lut1d = p.lut1d; t.rec1d = p.rec1d;
   t.rgrid = p.rgrid; t.rdens = p.rdens; t.rdens2 = p.rdens2; t.lutrot = p.lutrot; t.recrot = p.recrot;
   t.rgi2d = p.rgi2d; t.cgi2d = p.cgi2d;
   t.pa1d = p.pa1d; t.pb1d = p.pb1d;
   t.s_pa1d = t.s_pb1d = t.s_rgi2d = t.s_cgi2d = 0;
   if (p.rs2d) {
      double2 *d2 = reinterpret_cast<double2 *>(cursor);
      for (int i = threadIdx.x; i < p.rs2d; i += blockDim.x) d2[i] = p.rgi2d[i];
      for (int i = threadIdx.x; i < p.cs2d; i += blockDim.x) d2[p.rs2d + i] = p.cgi2d[i];
      t.rgi2d = d2; t.cgi2d = d2 + p.rs2d;
      t.s_rgi2d = (uint32_t)__cvta_generic_to_shared(d2); t.s_cgi2d = t.s_rgi2d + 16u * (uint32_t)p.rs2d;
      cursor += 2 * (size_t)(p.rs2d + p.cs2d);
   }
   if (p.nrot && p.rot_in_smem) {
      const int nd = (p.nrot - 1) * (int)(sizeof(SplineRec) / sizeof(double));
      const double *src = reinterpret_cast<const double *>(p.recrot);
      for (int i = threadIdx.x; i < nd; i += blockDim.x) cursor[i] = src[i];
      t.recrot = reinterpret_cast<const SplineRec *>(cursor);
      cursor += (nd + 1) & ~1;
      int *li = reinterpret_cast<int *>(cursor);
      for (int i = threadIdx.x; i < p.nlutrot; i += blockDim.x) li[i] = p.lutrot[i];
      t.lutrot = li;
      cursor += ((p.nlutrot + 1) / 2 + 1) & ~1;
   }
   if (p.n1d && p.poly1d) {
      // uniform grid: only the per-interval cubics go to shared memory; the packed records (ends of the grid, the
      // one-off evaluations outside the batched sums) are read from global memory
      double2 *d2 = reinterpret_cast<double2 *>(cursor);
      const int nk = p.n1d - 1;
      for (int i = threadIdx.x; i < nk; i += blockDim.x) { d2[i] = p.pa1d[i]; d2[nk + i] = p.pb1d[i]; }
      t.pa1d = d2; t.pb1d = d2 + nk;
      t.s_pa1d = (uint32_t)__cvta_generic_to_shared(d2); t.s_pb1d = t.s_pa1d + 16u * (uint32_t)nk;
      cursor += 4 * (size_t)nk;
   } else if (p.n1d) {
      const int nd = (p.n1d - 1) * (int)(sizeof(SplineRec) / sizeof(double));
      const double *src = reinterpret_cast<const double *>(p.rec1d);
      for (int i = threadIdx.x; i < nd; i += blockDim.x) cursor[i] = src[i];
      t.rec1d = reinterpret_cast<const SplineRec *>(cursor);
      cursor += (nd + 1) & ~1;
      if (!p.uniform1d) {
         int *li = reinterpret_cast<int *>(cursor);
         for (int i = threadIdx.x; i < p.nlut1d; i += blockDim.x) li[i] = p.lut1d[i];
         t.lut1d = li;
         cursor += ((p.nlut1d + 1) / 2 + 1) & ~1;
      }
   }
   __syncthreads();
}

#ifndef PIMC_MAX_THREADS
#define PIMC_MAX_THREADS 512
#endif
template <int KIND>
__global__ void __launch_bounds__(PIMC_MAX_THREADS, 1)
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
   x.team_cta = x.tid / x.T;
   x.nteams_chain = x.nthreads_chain / x.T;
   x.W = x.T > 32 ? x.T >> 5 : 1;
   x.tw = x.T > 32 ? x.lane_t >> 5 : 0;
   x.wl = x.T > 32 ? (x.tid & 31) : x.lane_t;
   x.wlanes = x.T > 32 ? 32 : x.T;
   x.G = p.rot_group;
   x.gl = x.tid & (x.G - 1);
   x.grp = x.tid / x.G;
   x.ngrp = blockDim.x / x.G;
   x.gmask = x.G >= 32 ? 0xffffffffu : (((1u << x.G) - 1u) << ((x.tid & 31) & ~(x.G - 1)));
   x.bar_target = 0;
   x.red_par = 0;
   x.mol_target[0] = x.mol_target[1] = 0;
   double *cursor = smem;
   x.red = cursor; cursor += 40;
   stage_tables(p, x.t, cursor);
   x.team_buf = cursor + (size_t)x.team_cta * p.team_buf_n;
   if (!p.segbuf_global) cursor += (size_t)(blockDim.x / x.T) * p.team_buf_n;
   x.part = cursor + 2 * (size_t)((x.tid >> 5) - (x.gl >> 5));      // first warp of this thread's rot group
   cursor += 2 * (size_t)(blockDim.x >> 5);
   x.rrng = reinterpret_cast<uint32_t *>(cursor);
   const bool piped = ((KIND & 3) != 0) && p.rot_fused && p.Q > 0;
   const int nown = (p.Q + p.cpc - 1) / p.cpc;
   x.rot_iter = 0;
   if (piped) {
      cursor += (size_t)nown * 3;
      for (int i = x.tid; i < nown * 6; i += blockDim.x) {
         const int q = owned_slice(p, x, i / 6);
         if (q < p.Q) x.rrng[i] = stream_ptr(p, x.c, p.P + q)[i % 6];
      }
      x.slot = reinterpret_cast<RotSlot *>(cursor);                  // one slot per owned slice
      for (int ls = x.tid; ls < nown; ls += blockDim.x) {
         const int q = owned_slice(p, x, ls);
         if (q >= p.Q) continue;
         RotSlot *sl = x.slot + ls;
         const double phi = p.ang[ang_index(p, x.c, q, 0, 0)], cost = p.ang[ang_index(p, x.c, q, 1, 0)], chi = p.ang[ang_index(p, x.c, q, 2, 0)];
         sl->cur[0] = cost; sl->cur[1] = phi; sl->cur[2] = chi;
         if ((KIND & 3) == 2) {
            Mat3 R1;
            matpre(phi, acos(cost), chi, R1);
            for (int i = 0; i < 9; i++) sl->b[i] = R1.m[i / 3][i % 3];
         } else {
            for (int d = 0; d < 3; d++) sl->b[d] = p.cosn[ang_index(p, x.c, q, d, 0)];
         }
         sl->vcache = p.vold[((size_t)x.c * p.Q + q) * p.NMpad];
         sl->vep = p.vepoch[((size_t)x.c * p.Q + q) * p.NMpad];
         sl->gep = -1;
         if (((KIND & 7) == 1) && p.rot_run) rot_ll_publish(p.rot_ll + ((size_t)x.c * p.Q + q) * 8, sl->b, 0);
      }
      if (KIND & 8) {
         x.rot_done = reinterpret_cast<volatile int *>(x.slot + nown);
         for (int i = x.tid; i < p.Q; i += blockDim.x) x.rot_done[i] = 0;
      }
      __syncthreads();
   } else x.slot = reinterpret_cast<RotSlot *>(cursor) + x.grp;      // one slot per rot group

   // worm scratch (WormShared + neighbour lists) behind the rot slots
   unsigned char *worm_scr = nullptr;
   if ((KIND & 4) && p.worm_on) {
      size_t off = (size_t)(piped ? nown : (x.G > 1 ? (int)(blockDim.x / x.G) : 0)) * sizeof(RotSlot);
      worm_scr = reinterpret_cast<unsigned char *>(cursor) + ((off + 15) & ~(size_t)15);
   }
   // time = t mod P and, per type, time mod (P/seg) and time / (P/seg), advanced incrementally (no divisions in the loop)
   int time = (int)(t0 % p.P), tmod[MAXT], toff[MAXT], tnseg[MAXT];
   for (int type = 0; type < p.ntypes; type++) {
      tnseg[type] = p.P / (1 << p.levels[type]);
      tmod[type] = time % tnseg[type];
      toff[type] = time / tnseg[type];
   }
   const bool use_run = ((KIND & 7) == 1) && piped && p.rot_run;
   const bool use_run_cta = (KIND & 8) != 0;             // the variant is only launched where it applies (pimcgpu_init)
   for (long s = 0; s < nsteps; s++) {
      if (use_run || use_run_cta) {
         // translational sweeps of this step, then every rotational sweep up to the next translational one in a single
         // free-running stretch (rot_run): the rotor is the last species, so its sweep closes the step
         for (int type = 0; type < p.ntypes; type++) {
            if (time == 0) {
               if (p.mol_piped && p.numb[type] > 1 && fast_atoms<KIND>(p, type) && p.ncyc[x.c * MAXT + type] == p.numb[type]) molecular_sweep_piped<KIND>(p, x, type);
               else molecular_sweep<KIND>(p, x, type);
            }
            if (tmod[type] == 0) {
               if (p.bis_piped && x.W == 2 && p.numb[type] > 1 && fast_atoms<KIND>(p, type)) bisection_sweep_piped<KIND>(p, x, type, toff[type]);
               else bisection_sweep<KIND>(p, x, type, toff[type]);
            }
         }
         int nrun = 1;
         for (;;) {
            if (s + nrun >= nsteps) break;
            bool tr = false;
            for (int type = 0; type < p.ntypes; type++) tr |= ((time + nrun) % tnseg[type]) == 0;     // includes time + nrun == P
            if (tr) break;
            nrun++;
         }
         if (use_run) rot_run<KIND>(p, x, p.imtype, nrun, err);
         else if (KIND & 8) rot_run_cta<KIND>(p, x, p.imtype, nrun, err);
         for (int k = 0; k < nrun; k++) {
            if (++time == p.P) time = 0;
            for (int type = 0; type < p.ntypes; type++) {
               if (time == 0) { tmod[type] = 0; toff[type] = 0; }
               else if (++tmod[type] == tnseg[type]) { tmod[type] = 0; toff[type]++; }
            }
         }
         s += nrun - 1;
         continue;
      }
      // does the NEXT step start with a translational sweep?  (decided before this step's rotational sweep runs ahead)
      bool next_trans = false;
      for (int type = 0; type < p.ntypes; type++) next_trans |= (tmod[type] + 1 == tnseg[type]) || (time + 1 == p.P) || p.worm_on;
      for (int type = 0; type < p.ntypes; type++) {
         bool closed = true;
         if ((KIND & 4) && p.worm_on && type == p.worm_type) {
            // mc_main.cc:355-379: MCWormMove, then the path moves of this type in the Z sector only
            // one CTA per chain holds the current rotation matrix of every slice in its rot slots
            if (x.crank == 0) worm_sweep_cta<KIND>(p, x.t, x.c, x.red, worm_scr, (piped && p.cpc == 1 && (KIND & 3) == 2) ? x.slot[0].b : nullptr, sizeof(RotSlot));
            chain_sync(p, x);
            closed = p.wstate[(size_t)x.c * 8] == 0;
         }
         if (time == 0 && closed) {
            if (!(KIND & 4) && p.mol_piped && p.numb[type] > 1 && fast_atoms<KIND>(p, type) && p.ncyc[x.c * MAXT + type] == p.numb[type]) molecular_sweep_piped<KIND>(p, x, type);
            else molecular_sweep<KIND>(p, x, type);
         }
         if (tmod[type] == 0 && closed) {
            if (!(KIND & 4) && p.bis_piped && x.W == 2 && p.numb[type] > 1 && fast_atoms<KIND>(p, type)) bisection_sweep_piped<KIND>(p, x, type, toff[type]);
            else bisection_sweep<KIND>(p, x, type, toff[type]);
         }
         if ((KIND & 3) != 0 && type == p.imtype && p.Q > 0) {
            if (KIND & 8) { }                             // free-running variant: every sweep goes through rot_run_cta above
            else if (piped) {
               rot_sweep_pipe<KIND>(p, x, type, err);
               // a translational sweep (or the end of the launch) needs every decision of this sweep: full barrier
               if ((s == nsteps - 1 || type != p.ntypes - 1 || next_trans) && p.cpc > 1) chain_sync(p, x);
            } else rot_sweep<KIND>(p, x, type, err);
         }
      }
      if (++time == p.P) time = 0;
      for (int type = 0; type < p.ntypes; type++) {
         if (time == 0) { tmod[type] = 0; toff[type] = 0; }
         else if (++tmod[type] == tnseg[type]) { tmod[type] = 0; toff[type]++; }
      }
   }
   if (piped) {
      __syncthreads();
      for (int i = x.tid; i < nown * 6; i += blockDim.x) {
         const int q = owned_slice(p, x, i / 6);
         if (q < p.Q) stream_ptr(p, x.c, p.P + q)[i % 6] = x.rrng[i];
      }
      for (int ls = x.tid; ls < nown; ls += blockDim.x) {
         const int q = owned_slice(p, x, ls);
         if (q >= p.Q) continue;
         p.vold[((size_t)x.c * p.Q + q) * p.NMpad] = x.slot[ls].vcache;
         p.vepoch[((size_t)x.c * p.Q + q) * p.NMpad] = x.slot[ls].vep;
      }
   }
}

} // namespace pimc
