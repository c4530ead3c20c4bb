// Worm moves on the device: open, close, advance, recede and swap of
// mc_qworm.cc:93-667 for one chain, executed by CTA 0 of the chain.  The moves of one MCWormMove call are inherently
// sequential (each starts from the worm the previous one left), so thread 0 carries the control flow -- uniforms,
// Levy-bridge sampling of the <= m new beads, the permutation table of the swap -- while every potential sum over
// (slices of the segment) x (partners) is dealt to all threads of the CTA and reduced in a fixed order.
//
// All uniforms come from ONE MRG32k3a stream of the chain (index P + Q + 1) in program order; the CPU replay is
// oracle/pimc_oracle.cpp:worm_move in schedule mode.  Worm atoms are numbered inside the worm's type, as in the
// reference (Worm.atom_i, Worm.atom_m, PIndex, RIndex); the device arrays pindex/rindex hold global atom indices.
#pragma once
#include "pimc_device.cuh"

// clock64 timeline of chain 0 / thread 0 (profiles/timeline.py; builds with -DPIMC_TIMELINE only)
#ifdef PIMC_TIMELINE
namespace pimc {
__device__ long long g_marks[4096];
__device__ int g_nmarks;
}
#define WMARK(c, id) do { if (threadIdx.x == 0 && (c) == 0 && g_nmarks < 4000) { g_marks[g_nmarks++] = ((long long)(id) << 48) | (clock64() & 0xffffffffffffLL); } } while (0)
#else
#define WMARK(c, id) do { } while (0)
#endif

namespace pimc {

constexpr int QW_OPEN = 0, QW_CLOSE = 1, QW_ADVANCE = 4, QW_RECEDE = 5, QW_SWAP = 6;     // mc_qworm.h:58-64
constexpr int WORM_MAXM = 64;          // largest Worm.m the per-CTA scratch holds
constexpr int WORM_MAXNEIGHBORS = 100; // mc_qworm.cc:12

// WorldLine(atom, pt), mc_qworm.cc:553-575; st = {exists, ira, masha, atom_i, atom_m}
__device__ __forceinline__ bool worm_world_line(const int *st, int atom, int pt)
{
   const int ira = st[1], masha = st[2], atom_i = st[3], atom_m = st[4];
   bool wline = true;
   if ((atom == atom_m) || (atom == atom_i)) {
      if ((atom_i != atom_m) || (ira > masha)) {
         if (((atom == atom_m) && (pt < masha)) || ((atom == atom_i) && (pt > ira))) wline = false;
      } else {
         if ((pt > ira) && (pt < masha)) wline = false;
      }
   }
   return wline;
}
// the mask of the PotEnergy partner loops (mc_piqmc.cc:1226-1227,1816-1817,2000-2001,2082-2083): false when partner j
// must be skipped at slice it
template <int KIND = 4>
__device__ __forceinline__ bool partner_on_line(const Params &p, int c, int j, int it)
{
   if (!(KIND & 4)) return true;          // kernel variant without a worm: no mask code at all
   if (!p.worm_on) return true;
   const int *st = p.wstate + (size_t)c * 8;
   if (!st[0] || type_of(p, j) != p.worm_type) return true;
   return worm_world_line(st, j - p.first[p.worm_type], it);
}

struct WormShared {
   int st[5];                 // exists, ira, masha, atom_i, atom_m
   int it0, it1, atom0, atom1, use_path, diff, go, changed, perm_changed;
   double path[(WORM_MAXM + 2) * 3];     // swap: the proposed path, point k = it - it0
   double result;
   int count;                            // entries of the permutation table (worm_get_ptable)
   int wb;                               // 1: the bridge of a close/advance move sits in path[] and goes to the state before the sums
   int ng;                               // gaussians of the coming Levy bridge, drawn as one batch (worm_gauss_batch)
   uint32_t gstate[2][6];                // the worm stream before the batch / after it
   // the recursion of sample_middle as a node list in its own pre-order (worm_bridge_plan): points relative to the left end
   short nd_it0[WORM_MAXM + 1], nd_it1[WORM_MAXM + 1], nd_it2[WORM_MAXM + 1], nd_depth[WORM_MAXM + 1];
   int nnodes, ndepth, gi0;
   int pl_it0, pl_it2;                   // the bridge to lay out (worm_bridge_fill): end points, pl_it2 - pl_it0 < 2: none
   double gs[3 * (WORM_MAXM + 1)];       // sqrt(-log u1) cos(2 pi u2), still to be divided by sqrt(alpha)
};

// sum over the open interval (it0, it1) of PotEnergy(atom(it), pos, it mod P) -- get_potential, mc_qworm.cc:400-422 -- with
// the moving atom's beads from the state (use_path = 0) or from the proposed path; diff = 1 subtracts the same sum for the
// beads of the state (qworm_swap, mc_qworm.cc:479-489).  All threads of the CTA; returns the total to every thread.
template <int KIND>
__device__ double worm_pot_sum(const Params &p, const SmallTables &t, int c, const WormShared &w, double *red, int &flip,
                               const double *slot_b, size_t slot_stride)
{
   const int P = p.P, N = p.N, base = p.first[p.worm_type];
   const int it0 = w.it0, it1 = w.it1, nint = it1 - it0 - 1;
   const int pit0 = it0 % P;
   const int nv = w.diff ? 2 : 1;                         // swap: the proposed path (v = 0) minus the state (v = 1), one thread each
   double s = 0.0;
   for (int i = threadIdx.x; i < nv * nint * N; i += blockDim.x) {
      // the slice runs fastest: the lanes of a warp share the partner, hence the branch of the pair term
      const int vj = i / nint, k = i - vj * nint;
      const int v = vj / N, j = vj - v * N;
      const int it = it0 + 1 + k, pit = it % P;
      int atom = w.atom0;
      if (w.diff) { if (pit != it) atom = w.atom1; }                       // qworm_swap switches on the wrap alone
      else if ((pit != it) && (pit0 == it0)) atom = w.atom1;
      const int g = base + atom;
      if (j == g) continue;
      // the mask uses the worm as it stands in shared memory (the open move evaluates with exists = 0)
      if (w.st[0] && type_of(p, j) == p.worm_type && !worm_world_line(w.st, j - base, pit)) continue;
      double px[3];
      #pragma unroll
      for (int d = 0; d < 3; d++) px[d] = (w.use_path && v == 0) ? w.path[(it - it0) * 3 + d] : p.pos[pos_index(p, c, pit, d, g)];
      double e;
      if ((KIND & 3) == 2 && slot_b && p.mode[type_of(p, g)][type_of(p, j)] == M_TOP_1MOL) {
         // the top's rotation matrix of this slice from the CTA's rot slot (what load_rotmat would rebuild with an acos and
         // three sincos from the stored angles: the slot holds matpre of exactly those)
         const double *rb = reinterpret_cast<const double *>(reinterpret_cast<const char *>(slot_b) + (size_t)(pit / p.R) * slot_stride);
         Mat3 rm;
         #pragma unroll
         for (int q = 0; q < 9; q++) rm.m[q / 3][q % 3] = rb[q];
         double p1[3];
         #pragma unroll
         for (int d = 0; d < 3; d++) p1[d] = p.pos[pos_index(p, c, pit, d, j)];
         e = vcord(p, rm, p1, px, nullptr, nullptr);
      } else e = pair_energy<(KIND & 3)>(p, t, c, g, px, j, pit, nullptr, nullptr);
      s += v ? -e : e;
   }
   // fixed-order block reduction; the two halves of red[] alternate between calls, so one barrier per call is enough (a
   // half is rewritten two calls later, behind the barrier of the call in between)
   for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
   const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
   double *r = red + 16 * (flip & 1);
   flip ^= 1;
   if ((threadIdx.x & 31) == 0) r[warp] = s;
   __syncthreads();
   double tot = 0.0;
   for (int k = 0; k < nwarp; k++) tot += r[k];
   return tot;
}

__device__ __forceinline__ int w_nrnd(Mrg &g, int n) { return (int)floor(n * mrg_u01(g)); }

// The gaussians of one Levy bridge as a batch (all threads of the CTA).  Gaussian k uses draws 2k and 2k+1 of the worm's
// stream: thread k jumps a copy of the generator ahead by 2k steps with the tabulated transition matrices (exact integer
// arithmetic), draws its two uniforms and evaluates the transcendental part of gauss (mc_randg.cc:138-150); thread 0 then
// continues with the state behind the last draw and consumes gs[] in program order -- same draws, same operations, same bits
// as the sequential form, without its ~600 cycles of dependent latency per number on the thread that carries the control flow.
// the parallel part: w.ng and w.gstate[0] published by a barrier before, gs[] and gstate[1] by one after
__device__ __forceinline__ void worm_gauss_fill(const Params &p, WormShared &w)
{
   const int ng = w.ng;
   for (int k = threadIdx.x; k < ng; k += blockDim.x) {
      Mrg h;
      mrg_load(h, w.gstate[0]);
      mrg_jump(h, p.worm_jump + (size_t)k * 18);
      const double u1 = mrg_u01(h), u2 = mrg_u01(h);
      w.gs[k] = sqrt(-log(u1)) * cos(2.0 * PI * u2);
      if (k == ng - 1) mrg_store(h, w.gstate[1]);
   }
}
__device__ __forceinline__ void worm_gauss_batch(const Params &p, WormShared &w, Mrg &g)
{
   if (threadIdx.x == 0) mrg_store(g, w.gstate[0]);
   __syncthreads();                                   // w.ng and the generator published by thread 0
   const int ng = w.ng;
   if (ng <= 0) return;
   for (int k = threadIdx.x; k < ng; k += blockDim.x) {
      Mrg h;
      mrg_load(h, w.gstate[0]);
      mrg_jump(h, p.worm_jump + (size_t)k * 18);
      const double u1 = mrg_u01(h), u2 = mrg_u01(h);
      w.gs[k] = sqrt(-log(u1)) * cos(2.0 * PI * u2);
      if (k == ng - 1) mrg_store(h, w.gstate[1]);
   }
   __syncthreads();
   if (threadIdx.x == 0) mrg_load(g, w.gstate[1]);
}

// sample_middle, mc_qworm.cc:240-287: the recursion (midpoint it1 = rint((it0 + it2)/2), left half first) samples every
// interior point of the bridge exactly once, so a segment of length L holds L - 1 nodes and, in the recursion's own
// pre-order, the left child of node k is node k + 1 and the right child node k + (it1 - it0).  Node k consumes the
// gaussians gi0 + 3k .. gi0 + 3k + 2 exactly as the recursive form does.  Thread 0 only names the end points
// (worm_bridge_request); every interior point then finds its own node by walking down from the root (all threads), and
// the first warp fills the bridge level by level -- nodes of one depth only depend on shallower ones.
// The bridge lives in path[] (point index = it - pl_it0), both end points already in place.
__device__ __forceinline__ void worm_bridge_request(WormShared &w, int it0r, int it2r, int gi0)
{
   w.pl_it0 = it0r; w.pl_it2 = it2r; w.gi0 = gi0;
}
// all threads, after a barrier that published the request, the end points and gs[]
__device__ __forceinline__ void worm_bridge_fill(const Params &p, WormShared &w)
{
   const int it0r = w.pl_it0, L = w.pl_it2 - w.pl_it0, gi0 = w.gi0;
   if (L < 2) return;
   const int nn = L - 1;
   for (int j = 1 + threadIdx.x; j < L; j += blockDim.x) {
      int a = it0r, b = w.pl_it2, k = 0, dep = 0;
      const int target = it0r + j;
      for (;;) {
         const int mid = (int)rint(0.5 * (double)(a + b));
         if (mid == target) {
            w.nd_it0[k] = (short)(a - it0r); w.nd_it1[k] = (short)j; w.nd_it2[k] = (short)(b - it0r); w.nd_depth[k] = (short)dep;
            const double s0 = (double)(mid - a), s2 = (double)(b - mid);
            const double gkin = (s0 + s2) / (p.worm_twave2 * s0 * s2);
            const double sq = sqrt(gkin);
            #pragma unroll
            for (int d = 0; d < 3; d++) w.gs[gi0 + 3 * k + d] = w.gs[gi0 + 3 * k + d] / sq;      // the node's displacement
            break;
         }
         if (target < mid) { b = mid; k += 1; }
         else { k += mid - a; a = mid; }
         dep++;
      }
   }
   __syncthreads();
   if (threadIdx.x < 32) {
      int nd = 0;
      for (int k = threadIdx.x; k < nn; k += 32) nd = max(nd, (int)w.nd_depth[k] + 1);
      for (int o = 16; o > 0; o >>= 1) nd = max(nd, __shfl_xor_sync(0xffffffffu, nd, o));
      for (int dep = 0; dep < nd; dep++) {
         for (int k = threadIdx.x; k < nn; k += 32) {      // one lane per node, its three coordinates side by side
            if (w.nd_depth[k] != dep) continue;
            const int i0 = w.nd_it0[k], i1 = w.nd_it1[k], i2 = w.nd_it2[k];
            const double s0 = (double)(i1 - i0), s2 = (double)(i2 - i1);
            // a division by 2^n is the multiplication by 2^-n, bit for bit (no subnormals among coordinates): the levels of a
            // bridge form a dependent chain, and a double-precision division is most of a level
            const int len = i2 - i0;
            const bool pow2 = (len & (len - 1)) == 0;
            const double inv = __hiloint2double((1023 - (31 - __clz(len))) << 20, 0);
            #pragma unroll
            for (int d = 0; d < 3; d++) {
               const double x0 = w.path[i0 * 3 + d], x2 = w.path[i2 * 3 + d];
               const double num = s2 * x0 + s0 * x2;
               double x1 = pow2 ? num * inv : num / (s0 + s2);
               x1 += w.gs[gi0 + 3 * k + d];
               w.path[i1 * 3 + d] = x1;
            }
         }
         __syncwarp();
      }
   }
   __syncthreads();
}

// get_ptable, mc_qworm.cc:577-643: neighbours of world line atomw at slice pt0 among the beads at slice pt1, sorted by
// distance (mmsort, mc_utils.cc:206-231), weights exp(-dr^2 / (segm * 4 lambda tau)); entries 1..count.  All threads of the
// CTA: one thread per candidate world line for the distances, thread 0 for the (order-preserving) compaction and the
// insertion sort, one thread per entry for the weights.  `flag` is an int scratch of numb entries.
// worm_ptable_dist: the distances (no barrier; the LAST threads of the CTA take the candidates, so the phase can share a
// stretch with work that is laid out from thread 0 upwards); worm_ptable_finish: the rest, behind a barrier.
__device__ __forceinline__ void worm_ptable_dist(const Params &p, int c, const WormShared &w, int atomw, int pt0, int pt1, int t1, double *ptable, int *flag)
{
   const int base = p.first[p.worm_type], numb = p.numb[p.worm_type];
   const int *rindex = p.rindex + (size_t)c * p.N;
   const int *st = w.st;
   for (int atom1 = blockDim.x - 1 - threadIdx.x; atom1 < numb; atom1 += blockDim.x) {
      int ok = 0;
      double dr2 = 0.0;
      if (worm_world_line(st, atom1, pt1)) {
         int atom0 = atom1;
         if (t1 != pt1) atom0 = rindex[base + atom1] - base;
         if (atom0 != st[3]) {
            #pragma unroll
            for (int d = 0; d < 3; d++) {
               double dx = p.pos[pos_index(p, c, pt0, d, base + atomw)] - p.pos[pos_index(p, c, pt1, d, base + atom1)];
               if (p.minimage) dx -= (p.box[d] * rint(dx / p.box[d]));
               dr2 += (dx * dx);
            }
            if (dr2 < p.worm_cutoff2) ok = 1;
         }
      }
      flag[atom1] = ok;
      ptable[1 + atom1] = dr2;                          // candidate scratch until the weights are written
   }
}
__device__ int worm_ptable_finish(const Params &p, WormShared &w, int segm, double *dr2_list, int *atm_list, double *ptable, const int *flag)
{
   const int numb = p.numb[p.worm_type];
   __syncthreads();
   if (threadIdx.x == 0) {
      int count = 0;
      for (int atom1 = 0; atom1 < numb; atom1++)
         if (flag[atom1]) { count++; dr2_list[count] = ptable[1 + atom1]; atm_list[count] = atom1; }
      for (int j = 2; j <= count; j++) {
         const double dtmp = dr2_list[j];
         const int itmp = atm_list[j];
         int i = j - 1;
         while ((i > 0) && (dr2_list[i] > dtmp)) { dr2_list[i + 1] = dr2_list[i]; atm_list[i + 1] = atm_list[i]; i--; }
         dr2_list[i + 1] = dtmp; atm_list[i + 1] = itmp;
      }
      if (count > WORM_MAXNEIGHBORS) count = WORM_MAXNEIGHBORS;
      w.count = count;
   }
   __syncthreads();
   const int count = w.count;
   const double norm = 1.0 / ((double)segm * p.worm_twave2);
   for (int ic = 1 + threadIdx.x; ic <= count; ic += blockDim.x) ptable[ic] = exp(-norm * dr2_list[ic]);
   __syncthreads();
   return count;
}

__device__ int worm_get_ptable(const Params &p, int c, WormShared &w, int atomw, int pt0, int pt1, int segm, int t1,
                               double *dr2_list, int *atm_list, double *ptable, int *flag)
{
   worm_ptable_dist(p, c, w, atomw, pt0, pt1, t1, ptable, flag);
   return worm_ptable_finish(p, w, segm, dr2_list, atm_list, ptable, flag);
}

// permutation cycles of every type from pindex (what pimcgpu_upload_state prepares on the host); thread 0 only
__device__ void worm_rebuild_cycles(const Params &p, int c, int *seen /* [N] scratch */)
{
   const int N = p.N;
   const int *pindex = p.pindex + (size_t)c * N;
   int *cstart = p.cyc_start + (size_t)c * (N + 1), *catoms = p.cyc_atoms + (size_t)c * N, *ncyc = p.ncyc + (size_t)c * MAXT;
   for (int a = 0; a < N; a++) seen[a] = 0;
   int ns = 0, na = 0;
   for (int t = 0; t < p.ntypes; t++) {
      ncyc[t] = 0;
      for (int a = p.first[t]; a < p.first[t] + p.numb[t]; a++) {
         if (seen[a]) continue;
         cstart[ns++] = na;
         int b = a;
         do { catoms[na++] = b; seen[b] = 1; b = pindex[b]; } while (b != a && na <= N);
         ncyc[t]++;
      }
   }
   while (ns < N + 1) cstart[ns++] = na;
}

// beads 1 .. it1-it0-1 of a bridge built in path[] go to the state (close, advance): before the wrap they belong to atom0,
// after it to atom1 -- the assignment sample_middle makes.  All threads; the potential sums that follow read the state.
__device__ __forceinline__ void worm_write_back(const Params &p, int c, const WormShared &w)
{
   if (!w.wb) return;
   const int P = p.P, base = p.first[p.worm_type], it0 = w.it0, n = w.it1 - w.it0 - 1;
   for (int i = threadIdx.x; i < n * 3; i += blockDim.x) {
      const int k = 1 + i / 3, d = i % 3, it = it0 + k, pit = it % P;
      p.pos[pos_index(p, c, pit, d, base + (pit != it ? w.atom1 : w.atom0))] = w.path[k * 3 + d];
   }
   __syncthreads();
}

// MCWormMove, mc_qworm.cc:93-125, for chain c by the calling CTA.  `scratch` holds WormShared and the neighbour lists.
template <int KIND>
__device__ void worm_sweep_cta(const Params &p, const SmallTables &t, int c, double *red, unsigned char *scratch,
                               const double *slot_b = nullptr, size_t slot_stride = 0)
{
   int flip = 0;
   WormShared &w = *reinterpret_cast<WormShared *>(scratch);
   double *dr2_list = reinterpret_cast<double *>(scratch + ((sizeof(WormShared) + 15) & ~15));
   double *ptable = dr2_list + (p.N + 2);
   int *atm_list = reinterpret_cast<int *>(ptable + (p.N + 2));
   int *seen = atm_list + (p.N + 2);
   const int P = p.P, base = p.first[p.worm_type], numb = p.numb[p.worm_type], tid = threadIdx.x;
   int *gst = p.wstate + (size_t)c * 8;
   int *pindex = p.pindex + (size_t)c * p.N, *rindex = p.rindex + (size_t)c * p.N;
   double *qw = p.qwc + (size_t)c * 16;
   uint32_t *sp = p.rng + ((size_t)c * p.S + p.P + p.Q + 1) * 6;
   Mrg g;
   if (tid == 0) {
      mrg_load(g, sp);
      for (int i = 0; i < 5; i++) w.st[i] = gst[i];
      w.changed = 0; w.perm_changed = 0;
   }
   __syncthreads();
   const bool bose_worm = p.bstype >= 0 && p.worm_type == p.bstype;
   for (int atom = 0; atom < numb; atom++) {
      // ---------------- open / close ----------------
      WMARK(c, 40);
      int segm = 0, gi = 0;
      if (tid == 0) {
         w.ng = 0;
         if (w.st[0]) {
            segm = w.st[2] - w.st[1];
            if (segm < 0) segm += P;
            if (segm <= p.worm_m && segm >= 2) w.ng = 3 * (segm - 1);
         }
      }
      worm_gauss_batch(p, w, g);
      WMARK(c, 41);
      if (tid == 0) {
         qw[14] += 1.0;
         w.go = 0;
         w.wb = 0;
         w.pl_it0 = 0; w.pl_it2 = 0;
         if (w.st[0]) {                                     // qworm_close, mc_qworm.cc:184-238
            qw[QW_CLOSE] += 1.0;
            if (segm <= p.worm_m) {
               // the bridge is built in shared memory (no dependent global round trips) and written to the state by all threads
               gi = 0;
               #pragma unroll
               for (int d = 0; d < 3; d++) {
                  w.path[d] = p.pos[pos_index(p, c, w.st[1], d, base + w.st[3])];
                  w.path[segm * 3 + d] = p.pos[pos_index(p, c, w.st[2], d, base + w.st[4])];
               }
               worm_bridge_request(w, w.st[1], w.st[1] + segm, 0);
               w.wb = 1;
               w.go = 1;
            }
         } else {                                           // qworm_open, mc_qworm.cc:155-182
            qw[QW_OPEN] += 1.0;
            w.st[3] = w_nrnd(g, numb);
            w.st[1] = w_nrnd(g, P);
            segm = w_nrnd(g, p.worm_m) + 1;
            w.st[2] = (w.st[1] + segm) % P;
            w.st[4] = w.st[3];
            if (w.st[2] != (w.st[1] + segm)) w.st[4] = pindex[base + w.st[3]] - base;
            w.go = 1;
         }
         w.it0 = w.st[1]; w.it1 = w.st[1] + segm; w.atom0 = w.st[3]; w.atom1 = w.st[4]; w.use_path = 0; w.diff = 0;
      }
      __syncthreads();
      WMARK(c, 42);
      if (w.go) {
         worm_bridge_fill(p, w);
         WMARK(c, 43);
         worm_write_back(p, c, w);
         WMARK(c, 44);
         // qw_open_prob, mc_qworm.cc:127-153; the end-to-end distance does not depend on the sum (the beads between moved,
         // the two ends did not): its loads go out first
         double kin = 0.0;
         if (tid == 0) {
            #pragma unroll
            for (int d = 0; d < 3; d++) {
               double dr = p.pos[pos_index(p, c, w.st[1], d, base + w.st[3])] - p.pos[pos_index(p, c, w.st[2], d, base + w.st[4])];
               if (p.minimage) dr -= (p.box[d] * rint(dr / p.box[d]));
               kin += (dr * dr);
            }
            kin /= (p.worm_twave2 * (double)segm);
         }
         const double pot = worm_pot_sum<KIND>(p, t, c, w, red, flip, slot_b, slot_stride);
         WMARK(c, 45);
         if (tid == 0) {
            const double popen = p.worm_norm * pow((double)segm, 0.5 * 3.0) * exp(kin + pot * p.tau);
            const double prob = w.st[0] ? 1.0 / popen : popen;
            bool acc = false;
            if (prob >= 1.0) acc = true;
            else if (prob > mrg_u01(g)) acc = true;
            if (acc) {
               if (w.st[0]) { w.st[0] = 0; qw[7 + QW_CLOSE] += 1.0; }
               else { w.st[0] = 1; qw[7 + QW_OPEN] += 1.0; }
               w.changed = 1;
            }
         }
      }
      __syncthreads();
      WMARK(c, 46);
      // ---------------- advance / recede ----------------
      if (w.st[0]) {                                        // block-uniform: w.st is shared and only thread 0 writes it between barriers
         double r = 0.0;
         int steps = 0, sg = 0;
         if (tid == 0) {
            // the draws that precede the bridge's gaussians in the stream: advance-or-recede, then the length of the move
            r = mrg_u01(g);
            steps = w_nrnd(g, p.worm_m) + 1;
            w.ng = 0;
            if (r > 0.5) {
               sg = w.st[2] - w.st[1];
               if (sg < 0) sg += P;
               if (sg - steps > 0) w.ng = 3 * steps;         // the new head, then the steps - 1 beads between
            }
         }
         worm_gauss_batch(p, w, g);
         WMARK(c, 47);
         if (tid == 0) {
            qw[14] += 1.0;
            w.go = 0;
            w.wb = 0;
            w.pl_it0 = 0; w.pl_it2 = 0;
            if (r > 0.5) {                                  // qworm_advance, mc_qworm.cc:299-357
               qw[QW_ADVANCE] += 1.0;
               const int advance = steps;
               if (sg - advance > 0) {
                  const int it0 = w.st[1], it2 = w.st[1] + advance;
                  const int ira_new = it2 % P;
                  int atom_i_new = w.st[3];
                  if (ira_new != it2) atom_i_new = w.st[4];
                  const double gvar = 1.0 / ((double)advance * p.worm_twave2);
                  gi = 0;
                  #pragma unroll
                  for (int d = 0; d < 3; d++) {
                     w.path[d] = p.pos[pos_index(p, c, it0 % P, d, base + w.st[3])];
                     w.path[advance * 3 + d] = w.path[d] + (w.gs[gi++] / sqrt(gvar));          // the new head
                  }
                  worm_bridge_request(w, it0, it2, gi);
                  w.it0 = it0; w.it1 = it2 + 1; w.atom0 = w.st[3]; w.atom1 = atom_i_new; w.use_path = 0; w.diff = 0;
                  w.wb = 1;
                  w.go = 1;
               }
            } else {                                        // qworm_recede, mc_qworm.cc:359-398
               qw[QW_RECEDE] += 1.0;
               sg = w.st[1] - w.st[2];
               if (sg < 0) sg += P;
               const int recede = steps;
               if ((sg - recede) >= 1) {
                  int it0 = w.st[1] - recede, it1 = w.st[1];
                  int atom0 = w.st[3];
                  const int atom1 = w.st[3];
                  if (it0 < 0) { it0 += P; it1 += P; atom0 = rindex[base + atom1] - base; }
                  w.it0 = it0; w.it1 = it1 + 1; w.atom0 = atom0; w.atom1 = atom1; w.use_path = 0; w.diff = 0;
                  w.go = 2;
               }
            }
         }
         __syncthreads();
         WMARK(c, 48);
         if (w.go) {
            worm_bridge_fill(p, w);
            WMARK(c, 49);
            worm_write_back(p, c, w);
            WMARK(c, 50);
            const double pot = worm_pot_sum<KIND>(p, t, c, w, red, flip, slot_b, slot_stride);
            WMARK(c, 51);
            if (tid == 0) {
               bool acc = false;
               if (w.go == 1) {
                  if (pot < 0.0) acc = true;
                  else if (exp(-pot * p.tau) > mrg_u01(g)) acc = true;
                  if (acc) { qw[7 + QW_ADVANCE] += 1.0; w.st[1] = (w.it1 - 1) % P; w.st[3] = w.atom1; w.changed = 1; }
               } else {
                  if (pot > 0.0) acc = true;
                  else if (exp(pot * p.tau) > mrg_u01(g)) acc = true;
                  if (acc) { w.st[1] = w.it0 % P; w.st[3] = w.atom0; qw[7 + QW_RECEDE] += 1.0; w.changed = 1; }
               }
            }
         }
         __syncthreads();
         WMARK(c, 52);
      }
      // ---------------- swap (qworm_swap, mc_qworm.cc:424-551) ----------------
      if (bose_worm) {
         if (tid == 0) qw[14] += 1.0;
         if (w.st[0]) {
            double pnorm_old = 0.0, u_swap = 0.0;
            int sw_atom0 = -1, sw_atom1 = -1;
            int count;
            // A non-empty table draws ONE uniform (atom2swap), and a chosen partner then the gaussians of a bridge of fixed
            // length: both are drawn ahead, next to the table's distance loads, and handed back if the move stops earlier.
            Mrg g_before;
            if (tid == 0) {
               g_before = g;
               u_swap = mrg_u01(g);
               w.ng = (p.worm_m >= 2) ? 3 * (p.worm_m - 1) : 0;
               mrg_store(g, w.gstate[0]);
            }
            __syncthreads();
            worm_gauss_fill(p, w);
            {
               const int sg = p.worm_m, it0 = w.st[1], it1 = it0 + sg;
               count = worm_get_ptable(p, c, w, w.st[3], it0, it1 % P, sg, it1, dr2_list, atm_list, ptable, seen);
            }
            WMARK(c, 53);
            if (tid == 0) {
               qw[QW_SWAP] += 1.0;
               w.go = 0;
               const int sg = p.worm_m, it0 = w.st[1], it1 = it0 + sg, pit0 = it0, pit1 = it1 % P, atomw = w.st[3];
               if (count == 0) g = g_before;              // nothing was drawn
               if (count > 0) {
                  // atom2swap, mc_qworm.cc:645-667
                  for (int ic = 1; ic <= count; ic++) pnorm_old += ptable[ic];
                  const double prand = pnorm_old * u_swap;
                  double sum = 0.0;
                  int ic = 1;
                  while ((ic <= count) && (sum < prand)) { sum += ptable[ic]; ic++; }
                  ic--;
                  const int atom1 = atm_list[ic];
                  if (atom1 >= 0) {
                     int atom0 = atom1;
                     if (pit1 != it1) atom0 = rindex[base + atom1] - base;
                     #pragma unroll
                     for (int d = 0; d < 3; d++) {
                        w.path[0 * 3 + d] = p.pos[pos_index(p, c, pit0, d, base + atomw)];
                        w.path[sg * 3 + d] = p.pos[pos_index(p, c, pit1, d, base + atom1)];
                     }
                     sw_atom0 = atom0; sw_atom1 = atom1;
                     if (sg >= 2) mrg_load(g, w.gstate[1]);      // the bridge's gaussians are consumed
                  }
               }
            }
            WMARK(c, 55);
            if (tid == 0 && sw_atom1 >= 0) {
               const int sg = p.worm_m, it0 = w.st[1], it1 = it0 + sg;
               worm_bridge_request(w, it0, it1, 0);
               w.it0 = it0; w.it1 = it1; w.atom0 = sw_atom0; w.atom1 = sw_atom1; w.use_path = 1; w.diff = 1;
               w.go = 1;
            }
            __syncthreads();
            WMARK(c, 56);
            if (w.go) {
               worm_bridge_fill(p, w);
               WMARK(c, 57);
               worm_ptable_dist(p, c, w, w.atom0, w.it0, w.it1 % P, w.it1, ptable, seen);      // the reverse move's table: its loads fly during the sum
               const double pot = worm_pot_sum<KIND>(p, t, c, w, red, flip, slot_b, slot_stride);
               WMARK(c, 58);
               const int count2 = worm_ptable_finish(p, w, p.worm_m, dr2_list, atm_list, ptable, seen);
               if (tid == 0) {
                  double prob = exp(-pot * p.tau);
                  const int count = count2;
                  double pnorm_new = 0.0;
                  for (int ic = 1; ic <= count; ic++) pnorm_new += ptable[ic];
                  prob *= (pnorm_old / pnorm_new);
                  bool acc = false;
                  if (prob >= 1.0) acc = true;
                  else if (prob > mrg_u01(g)) acc = true;
                  w.go = acc ? 1 : 0;
               }
               __syncthreads();
               WMARK(c, 59);
               if (w.go) {
                  const int it0 = w.it0, it1 = w.it1, atom0 = w.atom0, atom1 = w.atom1, atomw = w.st[3];
                  // the proposed beads replace the state's between it0 and it1 ...
                  for (int i = tid; i < (it1 - it0 - 1) * 3; i += blockDim.x) {
                     const int it = it0 + 1 + i / 3, d = i % 3, pit = it % P;
                     p.pos[pos_index(p, c, pit, d, base + (pit != it ? atom1 : atom0))] = w.path[(it - it0) * 3 + d];
                  }
                  // ... and the world lines of atom0 and atomw trade their beads 0..it0
                  for (int i = tid; i < (it0 + 1) * 3; i += blockDim.x) {
                     const int it = i / 3, d = i % 3;
                     const size_t a = pos_index(p, c, it, d, base + atom0), b = pos_index(p, c, it, d, base + atomw);
                     const double va = p.pos[a], vb = p.pos[b];
                     p.pos[a] = vb; p.pos[b] = va;
                  }
                  __syncthreads();          // every thread has read the move's description before thread 0 updates the worm
                  if (tid == 0) {
                     qw[7 + QW_SWAP] += 1.0;
                     const int ratomw = rindex[base + atomw] - base, ratom0 = rindex[base + atom0] - base;
                     pindex[base + ratomw] = base + atom0; rindex[base + atom0] = base + ratomw;
                     pindex[base + ratom0] = base + atomw; rindex[base + atomw] = base + ratom0;
                     if (w.st[1] > w.st[2]) {
                        if (atom0 == w.st[4]) w.st[4] = w.st[3];
                        else if (w.st[3] == w.st[4]) w.st[4] = atom0;
                     }
                     w.changed = 1; w.perm_changed = 1;
                  }
               }
            }
            __syncthreads();
            WMARK(c, 61);
         }
      }
   }
   if (tid == 0) {
      mrg_store(g, sp);
      for (int i = 0; i < 5; i++) gst[i] = w.st[i];
      if (w.changed) p.pos_epoch[c] += 1;                   // cached rotor potentials are stale (beads moved or masks changed)
      if (w.perm_changed) worm_rebuild_cycles(p, c, seen);
   }
   __syncthreads();
}

__host__ __device__ inline size_t worm_scratch_bytes(int N)
{
   return ((sizeof(WormShared) + 15) & ~(size_t)15) + 2 * (size_t)(N + 2) * sizeof(double) + 2 * (size_t)(N + 2) * sizeof(int) + 16;
}

} // namespace pimc
