/* pimcgpu.h -- C ABI of the B200 (sm_100a) PIMC sampling hot path.
 *
 * Drop-in boundary for MoRiBS-PIMC's hot path.  The reference has no plugin
 * ABI: its driver (mc_main.cc) calls free functions that share global arrays
 * (mc_setup.h:176-191).  Each entry point below names the reference
 * interface it replaces; INTEGRATION.md shows the patch a maintainer applies
 * to mc_main.cc to route PIMCPass/MCGetAverage through this library.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on
 * success and nonzero on failure with a message in pimcgpu_last_error(); the
 * library never calls exit().  Host state arrays use the reference layout
 * [dim][atom*P + it] (mc_setup.cc:139-148, mc_utils.cc:10-30); the device
 * keeps its own slice-major SoA mirror.  Not re-entrant; one context per
 * process (= per GPU rank), like the reference's globals.
 */
#ifndef PIMCGPU_H
#define PIMCGPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define PIMCGPU_MAX_TYPES 2      /* one atom type + one molecule type (mc_input.cc:395-396) */
#define PIMCGPU_BINSR 300        /* MC_BINSR, mc_estim.cc:25 */
#define PIMCGPU_BINST 50         /* MC_BINST, mc_estim.cc:26 */
#define PIMCGPU_BINSC 100        /* MC_BINSC, mc_estim.cc:27 */
#define PIMCGPU_SIZE_ROTDEN (181 * 361 * 361)   /* SizeRotDen, mc_setup.h:116 */

/* TParticle (mc_setup.h:66-91): one species line of qmc.input (mc_input.cc:147-214) */
typedef struct {
   int    numb;       /* number of atoms/molecules                                  */
   int    molecule;   /* 0 ATOM, 1 MOLECULE (linear rotor), 2 NONLINEAR             */
   int    stat;       /* 0 BOLTZMANN, 1 BOSE                                        */
   int    levels;     /* bisection levels; segment = 2^levels (mlsegm)              */
   double mass;       /* amu (MCInitParams, mc_setup.cc:244-319)                    */
   double mcstep;     /* whole-path move step, Angstrom                             */
   double rtstep;     /* rotational step (ROTATION line, mc_input.cc:261-273)       */
} pimcgpu_type;

/* the scalars of mc_setup.h:13-63,103-107 plus how to spread chains over the device */
typedef struct {
   int          ntypes;
   pimcgpu_type type[PIMCGPU_MAX_TYPES];
   int          P;             /* NumbTimes                                          */
   int          Q;             /* NumbRotTimes, 0 without ROTATION                   */
   double       temperature;   /* Kelvin                                             */
   int          ispher;        /* ISPHER                                             */
   int          minimage;      /* MINIMAGE                                           */
   double       box[3];        /* BoxSize                                            */
   int          rotden_type;   /* RotDenType; only 0 (tabulated) runs on the device  */
   int          nchains;       /* independent Markov chains on this device           */
   long         chain_offset;  /* global index of local chain 0 (rank * nchains)     */
   int          device;        /* CUDA device ordinal                                */
   int          ctas_per_chain;/* thread-block cluster size per chain, 0 = auto      */
   int          threads_per_cta;/* 0 = auto                                          */
   int          team;          /* lanes cooperating on one segment / rot slice, 0 = auto */
   /* ROTDENSI line (mc_input.cc:274-284): with rotden_type = 1 the rattle-and-shake propagator of
    * rsrot_/rsline_ (rotden.f:218-286,356-375) replaces the tables in moves and estimators     */
   int          rot_odevn;     /* RotOdEvn (only the -1 branch of rsrot_ is live)               */
   double       rot_eoff;      /* RotEoff, cm^-1 (unused by the live branch)                    */
   double       x_rot, y_rot, z_rot;  /* X_Rot Y_Rot Z_Rot, cm^-1                                */
   int          rnratio;       /* RNratio: RS slices per Noya slice in GetRotE3D (0 -> 1)       */
   /* symmetry operations applied with probability 1/2 at the end of every measurement
    * (MCGetAverage, mc_main.cc:647-692): REFLECTX/Y/Z and ROTSYM (mc_input.cc:296-330)         */
   int          reflect[3];    /* IREFLX, IREFLY, IREFLZ                                        */
   int          rotsym;        /* IROTSYM                                                       */
   int          nfold_rot;     /* NFOLD_ROT                                                     */
   /* WORM line (mc_input.cc:286-293) + MCWormInit (mc_qworm.cc:48-82): worm moves on the device */
   int          worm;          /* 1: sample exchange with the worm algorithm                    */
   int          worm_type;     /* Worm.type: index of the species the worm lives in             */
   double       worm_c;        /* Worm.c as in qmc.input (normalised with the density at init)  */
   int          worm_m;        /* Worm.m (m-tilde), < P                                         */
} pimcgpu_system;

/* host pointers to the tables the reference loads in InitPotentials / InitRotDensity
 * (mc_poten.cc:93-164); all are copied to the device by pimcgpu_init.  Unused ones NULL. */
typedef struct {
   /* 1-D pair potential of the atom type: init_pot1D, mc_poten.cc:379-438 (K, Angstrom) */
   int           n1d;
   const double *grid1d, *pot1d;
   /* 2-D atom--linear-rotor potential: init_pot2D, mc_poten.cc:305-377 */
   int           rsize2d, csize2d;
   double        dr2d, dc2d;
   const double *rgrid2d, *cgrid2d, *pot2d;           /* pot2d[rsize][csize]        */
   /* 3-D atom--top potential: init_pot3D, mc_poten.cc:254-303 (r in bohr, 1-degree grids) */
   int           rgrd, thgrd, chgrd;
   double        rvmin, rvmax;
   const double *vtable;                              /* [rgrd][thgrd][chgrd]       */
   /* linear-rotor density matrix columns of <type>_T<T>t<Q>.rot: init_rotdens, mc_poten.cc:508-546 */
   int           nrot;
   const double *rotgrid, *rotdens, *rotderv, *rotesqr;
   /* top density matrix / energy / energy^2 tables: init_rot3D, mc_poten.cc:440-506 */
   const double *rho3d, *erot3d, *esq3d;              /* PIMCGPU_SIZE_ROTDEN each   */
   /* 501-entry spherical H2O-pH2 table of vspher.f:15-517 (ISPHER=1), else NULL     */
   const double *vspher;
} pimcgpu_tables;

/* Block accumulators: the file-statics of mc_estim.cc:39-103 and mc_main.cc:45-64,
 * SUMS and COUNTS only (ratios are formed after the cross-GPU reduction).          */
typedef struct {
   double count;                 /* avergCount summed over chains                    */
   double kin, pot, rot, rotsq;  /* _bkin _bpot _brot _brotsq                        */
   double cv, cv_trans, cv_rot;  /* _bCv _bCv_trans _bCv_rot                         */
   double mctotal[PIMCGPU_MAX_TYPES][3];   /* MCTotal[type][MCMOLEC,MCMULTI,MCROTAT]  */
   double mcaccep[PIMCGPU_MAX_TYPES][3];   /* MCAccep                                 */
} pimcgpu_scalars;

/* ---- life cycle: MCMemAlloc/MCInit/InitPotentials/InitRotDensity (mc_main.cc:110-238) ---- */
int  pimcgpu_init(const pimcgpu_system *sys, const pimcgpu_tables *tab);
void pimcgpu_finalize(void);                       /* MCMemFree/Done* (mc_main.cc:495-505)  */
const char *pimcgpu_last_error(void);

/* ---- state: the arrays of mc_setup.h:176-191; chain = -1 addresses every chain ---- */
int pimcgpu_upload_state(int chain, const double *coords, const double *angles, const int *pindex);
int pimcgpu_download_state(int chain, double *coords, double *angles, double *cosine, int *pindex);

/* the same for chains first .. first+count-1 in one call: arrays [count][3][N*P], pindex [count][numb of the BOSE species] or
 * NULL (identity); one host<->device copy and one transposing kernel for the whole batch instead of per-chain round trips   */
int pimcgpu_upload_states(int first, int count, const double *coords, const double *angles, const int *pindex);
int pimcgpu_download_states(int first, int count, double *coords, double *angles, double *cosine);
/* as above, but `angles` / `cosine` only receive the rows the moves change -- [rotor atom * P + q], q < NumbRotTimes; the
 * other entries of the caller's MCAngles / MCCosine (allocated once, mc_setup.cc:135-163; never written after MCConfigInit,
 * :471-487) are left untouched: 6 Q doubles per rotor and chain cross the bus instead of 6 N P                                  */
int pimcgpu_download_states_rows(int first, int count, double *coords, double *angles, double *cosine);
/* split-phase forms of the two calls above: the copies ride a second stream and overlap the move kernel of the neighbouring
 * steps (a driver that keeps two sets of chains on the host alternates them on the device; mc_main.cc has no counterpart --
 * its state never leaves the host).  upload: _begin starts the bead copy into a staging buffer (the state may still be in use by
 * a running pimcgpu_steps) and prepares angles / permutations, _commit installs everything on the library's stream; the host
 * arrays must not change between the two.  download: _begin snapshots the beads on the library's stream and hands them to the
 * copy stream, _end waits and writes the rotor rows of angles / cosine like pimcgpu_download_states_rows.                      */
int pimcgpu_upload_states_begin(int first, int count, const double *coords, const double *angles, const int *pindex);
int pimcgpu_upload_states_commit(void);
int pimcgpu_download_states_begin(int first, int count, double *coords, double *angles, double *cosine);
int pimcgpu_download_states_end(void);

/* ---- MRG32k3a package seed: RngStream::SetPackageSeed (rngstream.cc:346-353) ---- */
int pimcgpu_seed(const unsigned long seed6[6]);

/* ---- moves: nsteps iterations of the `time` loop body mc_main.cc:349-381 (PIMCPass :510-522)
 *      for every chain, asynchronous on the library's stream                              ---- */
int pimcgpu_steps(long nsteps);
int pimcgpu_sync(void);
long pimcgpu_step_counter(void);                   /* passTotal of mc_main.cc:346              */
/* launch geometry chosen by pimcgpu_init: out8 = {ctas_per_chain, threads_per_cta, team, rot_group, smem bytes,
 * move-kernel variant (rotor kind 0/1/2; +8: free-running rotational sweeps of a top whose chain lives in one CTA),
 * clusters the device can keep resident at once, nchains}                                                     */
int pimcgpu_geometry(int *out8);

/* ---- estimators: the device part of MCGetAverage (mc_main.cc:551-646): GetKinEnergy,
 *      GetPotEnergy_Densities, GetRotEnergy/GetRotE3D, GetRCF, Cv terms -> accumulators   ---- */
int pimcgpu_measure(void);
/* accumulator buffer (doubles): layout from pimcgpu_accum_layout, lives on the device so the
 * host plumbing can all-reduce it in place over NCCL before the Save* writers run            */
int    pimcgpu_accum_layout(long *n_total, long *off_scalars, long *off_gr1d, long *off_gr2d, long *off_gr3d,
                            long *off_rcf, long *off_relbins);
void  *pimcgpu_accum_device_ptr(void);
int    pimcgpu_accum_download(double *host, long n);
/* the same in two halves: _begin queues the copy on the library's stream (host should be pinned), _end waits for it.
 * Queued before pimcgpu_download_states_begin, the block's sums (what SaveEnergy etc. need, mc_main.cc:700-760) reach
 * the host ahead of the much larger configuration.                                             */
int    pimcgpu_accum_download_begin(double *host, long n);
int    pimcgpu_accum_download_end(void);
int    pimcgpu_accum_reset(void);                  /* MCResetBlockAverage, mc_main.cc:524-549  */
int    pimcgpu_block_scalars(pimcgpu_scalars *out);/* reads the (possibly all-reduced) buffer  */
int    pimcgpu_counters(double *mctotal, double *mcaccep);   /* [types][3], MCTotal/MCAccep    */
void  *pimcgpu_stream(void);                       /* cudaStream_t the library launches on     */

/* offset (in doubles) of a named region of the accumulator buffer, -1 if unknown:
 *   "scalars" "gr1d" "gr2d" "gr3d" "rcf" "relbins"  as in pimcgpu_accum_layout
 *   "area"   40 doubles: _areas[PERP,PARL] _area2[2] _inert[2] (GetAreaEstimators, mc_estim.cc:2087-2250), then
 *            _areas3DSFF[6] _inert3DSFF[9] _areas3DMFF[6] _inert3DMFF[9] (GetAreaEstim3D, :2252-2594)
 *   "ploops" one double per boson: _ploops (GetExchangeLength, :1997-2019)
 *   "rcfcnt" Q doubles: rows 1..9 of the block array _rcf, i.e. the number of time origins with n(0).n(t) < PLONE
 *            (GetRCF, :1127-1137; the Legendre call is commented out, so every row holds this count)      */
long   pimcgpu_accum_offset(const char *name);

/* ---- symmetry operations of MCGetAverage (mc_main.cc:647-692): Reflect_MF_XZ/YZ/XY (mc_piqmc.cc:1385-1708) and
 *      RotSymConfig (:1710-1794) on the device.  pimcgpu_measure applies them after the estimators when the system
 *      enables them; pimcgpu_symmetry_moves does the same on its own.  pimcgpu_symmetry_ops applies explicit
 *      operations: ops[chain][4] = {XZ, YZ, XY reflection flags, rotor index for the symmetry rotation or -1}  ---- */
int pimcgpu_symmetry_moves(void);
int pimcgpu_symmetry_ops(const int *ops);

/* ---- worm moves (mc_qworm.cc:93-667).  With sys.worm set, pimcgpu_steps runs MCWormMove for the worm's species in
 *      every step and its path moves only while the worm is closed (mc_main.cc:355-379); pimcgpu_measure skips chains
 *      whose worm is open.  Worm atoms are numbered inside their species, as Worm.atom_i / atom_m.            ---- */
int pimcgpu_worm_moves(void);                               /* one MCWormMove per chain (parity entry point)        */
int pimcgpu_worm_state(int chain, int *st5);                /* Worm.exists, ira, masha, atom_i, atom_m              */
int pimcgpu_worm_set(int chain, const int *st5);
int pimcgpu_worm_counters(double *total7, double *accep7, double *countqw);   /* QWTotal, QWAccep, countQW over all chains */

/* ---- checkpoint for an EXACT restart.  The reference's yw001.conf holds only the x row of MCCoords and
 *      MCCosine (mc_input.cc:560-609), so its restart is lossy; this blob holds every chain's beads, angles, permutation
 *      tables, worm, MRG32k3a streams, rotor-potential cache and the step counter.  A context initialised with the same
 *      system and tables that loads it continues bit-identically.  The driver writes it to the side file yw001.b200 and
 *      leaves yw001.stat/.conf/.tabl/.worm byte-compatible.                                                        ---- */
long pimcgpu_checkpoint_bytes(void);
int  pimcgpu_checkpoint_save(void *buf, long nbytes);
int  pimcgpu_checkpoint_load(const void *buf, long nbytes);

/* instantaneous area estimators of one chain, out[28]: area_perp, area_parl, inert_perp, inert_parl of
 * GetAreaEstimators (linear dopant; inertia sums before the division by NumbTimes), then area_proj[3] and
 * inert3D[9] of GetAreaEstim3D in the space-fixed frame, then the same in the dopant-fixed frame             */
int pimcgpu_chain_areas(int chain, double *out28);

/* instantaneous estimator values of one chain (parity with GetKinEnergy, GetPotEnergy,
 * GetRotEnergy|GetRotE3D, ErotSQ, Erot_termSQ; mc_estim.cc:689-1096): out[5]                 */
int pimcgpu_chain_energies(int chain, double *out5);
int pimcgpu_chain_rcf(int chain, double *rcf0 /* [Q] */);    /* GetRCF row 0, mc_estim.cc:1099-1139 */

/* ---- batched leaf evaluations on the device (parity entry points; pure functions) ---- */
int pimcgpu_eval_spot1d(int n, const double *r, double *v, int *klo);                 /* SPot1D  mc_poten.cc:624-639 */
int pimcgpu_eval_lpot2d(int n, const double *r, const double *cost, double *v, int *ir, int *ic); /* LPot2D :688-729 */
int pimcgpu_eval_srotdens(int n, const double *gamma, int which, double *v);          /* SRotDens* :548-622 */
int pimcgpu_eval_rotden(int n, const double *eul1, const double *eul2, double *rho, double *erot, double *esq,
                        int *index);                                                  /* rotden_ rotden.f:1-31 */
int pimcgpu_eval_vcord(int n, const double *eul, const double *rcom, const double *rpt, double *v, double *rtc,
                       int *index);                                                   /* vcord_ vcord.f:1-98 */
int pimcgpu_eval_caleng(int n, const double *com1, const double *com2, const double *eul1, const double *eul2,
                        double *e);                                                   /* caleng_ caleng_tip4p_gg.f:2-186 */
/* pure table selectors: identical doubles in, flat index and interpolated value out (bit-exact index contract) */
int pimcgpu_eval_rotpro(int n, const double *deg /* [n][3] phi, theta, chi in degrees */, double *rho, double *erot,
                        double *esq, int *index);                                     /* rotpro rotpro_sub.f:1-64; erot, esq in cm^-1; index < 0: out of range (-1-index) */
int pimcgpu_eval_vcalc(int n, const double *rtc /* [n][3] r bohr, theta deg, chi deg */, double *v, int *index); /* vcalc vcalc.f:1-65 */
int pimcgpu_eval_deleul(int n, const double *eul1, const double *eul2, double *rel /* [n][3] rad */);      /* deleul rotden.f:32-134 */
int pimcgpu_eval_vcord_grid(int n, const double *eul, const double *rcom, const double *rpt, double *grid /* [n][3] r bohr, theta deg, chi deg handed to vcalc */); /* vcord.f:86-95 */
int pimcgpu_eval_vspher(int n, const double *r, double *v, double *rclamp);           /* vspher_ vspher.f:12-544; rclamp = the overwritten r argument (bohr) */
/* device libm against the host's: which = 0 sin, 1 cos, 2 acos, 3 atan, 4 exp, 5 log, 6 sqrt, 7 fmod(x, 2 pi) */
int pimcgpu_eval_libm(int which, int n, const double *x, double *y);
/* PotEnergy(atom, MCCoords, it) for every atom and slice of one chain: v[N][P] (mc_piqmc.cc:1796-1965) */
int pimcgpu_pot_energy_slice(int chain, double *v);
/* first n uniforms of MRG32k3a stream `stream` (global stream index, 2^127 spacing): RngStream::RandU01 */
int pimcgpu_rng_draws(long stream, int n, double *out);

/* ---- host-side table preparation (no device needed): what pimcgpu_init does with the raw tables ----
 * spline second derivatives + short/long-range constants: init_spline/init_pot1D (mc_utils.cc:188-201,
 * mc_poten.cc:416-437); MRG32k3a state of the s-th stream (rngstream.cc:303-321); interval-search bucket table */
int pimcgpu_host_spline(int n, const double *x, const double *y, double *y2, double *alpha_unode_c6);
int pimcgpu_host_stream_state(const unsigned long seed6[6], long stream, unsigned long state6[6]);
int pimcgpu_host_lut(int n, const double *x, int *lut /* [4n] */, double *scale);   /* returns the table length */

/* ---- rotational density-matrix table generators on the device (csrc/pimc_tablegen.cu).  They need no
 *      pimcgpu_init context.  Each replaces one of the reference's Fortran pre-processing programs and takes that
 *      program's command-line arguments in the same order; outputs are host arrays in the layout the programs write.
 *      There is no CPU fallback: without a CUDA device they fail with a message.                                     ---- */
/* nmv_prop/asymrho.f (asymrho.x T P iodevn ith0 ithend A B C maxj; a-run:4): theta planes ith0..ith1 (degrees, 0..180) of the
 * asymmetric-top propagator, energy and energy-square estimators; rho/eng/esq are [ith1-ith0+1][361 phi][361 chi], i.e. the
 * line order of rho.denXXX_rho/_eng/_esq (asymrho.f:713-725) and, for ith0=0, ith1=180, of the complete table read by
 * init_rot3D (mc_poten.cc:462-499).  info[16] (may be NULL) = the log lines "AT BETA" Z,E(cm-1),Cv for even k, odd k, classical
 * (asymrho.f:354-368), "AT TAU" Z,E for even, odd, classical (:413-425), emax.  Errors mirror the Fortran STOPs
 * (maxj > 876, iodevn outside -1..1, 'too large contribution from emax', QL 'FAIL').                                         */
int pimcgpu_gen_asymrho(double temprt, int nslice, int iodevn, int ith0, int ith1, double Arot, double Brot, double Crot,
                        int maxj, double *rho, double *eng, double *esq, double *info);
/* symtop_prop/symrho.f (symrho.x T P kmod ith0 ithend Bz Bxy maxj; a-run:4); same output layout; info[5] = ztau, zbeta,
 * Ebeta (K), Esqrt (K^2), Cv (symrho.f:100-104); error 'pmax too large' as symrho.f:115-118                                  */
int pimcgpu_gen_symrho(double temprt, int nslice, int kmod, int ith0, int ith1, double Bz, double Bxy, int maxj, double *rho,
                       double *eng, double *esq, double *info);
/* linear_prop/linden.f (linden.x T P B npt iodevn): out[npt][4] = cos(gamma), rho, erot, erotsq -- the four columns of
 * linden.out / <type>_T<T>t<Q>.rot read by init_rotdens (mc_poten.cc:518-545); bit-identical to the Fortran's arithmetic.
 * info[4] = tau, lmax, "Erot at Beta", "Cv at Beta" (linden.f:72-82)                                                         */
int pimcgpu_gen_linden(double temprt, int nslice, double bconst, int npt, int iodevn, double *out, double *info);
/* parity entry point: Wigner d^j_{mk}(theta) in the convention of wigd (asymrho.f:1006-1039), d[(maxj+1)][2maxj+1][2maxj+1] */
int pimcgpu_gen_wigner_d(int maxj, double theta, double *d);
/* device milliseconds of the last pimcgpu_gen_asymrho call: eigen-solve + projectors + Fourier coefficients, phi stage,
 * chi GEMM, combine                                                                                                          */
int pimcgpu_gen_timing(double *ms4);
/* host-side writers in the Fortran edit descriptors of the generators: E15.8 (scale1p = 0; asymrho.f:721-723) or 1P,E15.8
 * (scale1p = 1; linden.f:68) into buf[16]; one E15.8 value per line (append != 0 continues a file, as compile.x concatenates
 * planes); the four-column .rot file                                                                                         */
void pimcgpu_format_e15_8(double v, int scale1p, char *buf);
int  pimcgpu_write_e15_8(const char *path, const double *v, long n, int append);
int  pimcgpu_write_rot(const char *path, const double *out4, int npt);

/* measured FP64 FMA throughput of the current device in TFLOP/s (roofline denominator; not part of the path) */
int pimcgpu_fp64_peak(double *tflops);

#ifdef __cplusplus
}
#endif
#endif
