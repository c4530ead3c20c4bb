/* ORACLE (test infrastructure, not product code).
 *
 * C interface of the CPU restatement of MoRiBS-PIMC's sampling hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library.  State arrays use the reference layout
 * [dim][atom*P + it] (mc_setup.cc:139-148).
 */
#ifndef PIMC_ORACLE_H
#define PIMC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc orc_t;

/* per-type description, mirrors TParticle (mc_setup.h:66-91) */
typedef struct {
   int    numb;       /* atoms/molecules of this type                      */
   int    molecule;   /* 0 atom, 1 linear rotor, 2 non-linear rotor        */
   int    stat;       /* 0 BOLTZMANN, 1 BOSE                               */
   int    levels;     /* bisection levels, segment = 2^levels              */
   double mass;       /* amu                                               */
   double mcstep;     /* whole-path displacement step (Angstrom)           */
   double rtstep;     /* rotational step                                   */
} orc_type_t;

typedef struct {
   int        ntypes;        /* <= 2: at most one atom type and one molecule type (mc_input.cc:395-396) */
   orc_type_t type[2];
   int        P;             /* NumbTimes                                   */
   int        Q;             /* NumbRotTimes (0: no ROTATION)               */
   double     temperature;   /* Kelvin                                      */
   int        ispher;        /* ISPHER                                      */
   int        minimage;      /* MINIMAGE                                    */
   double     box[3];        /* BoxSize                                     */
   int        rotden_type;   /* RotDenType (0 tables, 1 rattle-shake)       */
   int        rot_odevn;     /* RotOdEvn                                    */
   double     rot_eoff, x_rot, y_rot, z_rot; /* ROTDENSI line               */
   int        rnratio;
} orc_system_t;

orc_t *orc_create(const orc_system_t *sys);
void   orc_destroy(orc_t *);

/* tables; 1-D/2-D/linear-rotor arrays are copied, the big 3-D ones are borrowed */
void orc_set_pot1d(orc_t *, int n, const double *grid, const double *v);
void orc_set_pot2d(orc_t *, int rsize, int csize, double dr, double dc, const double *rgrid,
                   const double *cgrid, const double *v);
void orc_set_pot3d(orc_t *, int rgrd, int thgrd, int chgrd, double rvmin, double rvmax, const double *v);
void orc_set_rotlin(orc_t *, int n, const double *grid, const double *dens, const double *derv,
                    const double *esqr);
void orc_set_rot3d(orc_t *, const double *rho, const double *erot, const double *esq);
void orc_set_vspher(orc_t *, const double *t501);
/* pure table look-ups (test hooks): rotpro_sub.f angles in degrees, vcalc.f r in bohr + degrees, deleul, vspher (clamped r out) */
void orc_rotpro(orc_t *, const double *deg3, double *rho, double *erot, double *esq, int *index, int *jstop);
double orc_vcalc(orc_t *, const double *rtc, int *index);
void orc_deleul(const double *e1, const double *e2, double *rel);
double orc_vspher(double r, double *rclamp);
void orc_get_pot1d_setup(orc_t *, double *y2, double *alpha_unode_c6);

/* state */
void orc_set_state(orc_t *, const double *coords, const double *angles, const int *pindex);
void orc_get_state(orc_t *, double *coords, double *angles, double *cosine);

/* leaf functions (a7-a12); optional index outputs for the bit-exact index tests */
double orc_spot1d(orc_t *, double r, int *klo);
double orc_lpot2d(orc_t *, double r, double cost, int *ir, int *ic);
double orc_srotdens(orc_t *, double gamma, int which); /* 0 rho, 1 deriv, 2 esq */
void   orc_rotden(orc_t *, const double *eul1, const double *eul2, double *eulrel, double *rho,
                  double *erot, double *esq, int *index, int *istop);
double orc_vcord(orc_t *, const double *eul, const double *rcom, const double *rpt, double *rtc,
                 int *index);
double orc_caleng(const double *com1, const double *com2, const double *eul1, const double *eul2);

/* per-bead potential sums (a6) on the current state */
double orc_pot_energy_it(orc_t *, int atom, const double *pos3, int it);   /* pos3==NULL: own bead */
double orc_pot_energy_path(orc_t *, int atom, const double *shift3);       /* shift3==NULL: unshifted */
double orc_pot_rot_energy(orc_t *, int atom, const double *cosine3, int it);
double orc_pot_rot_e3d(orc_t *, int atom, const double *eul3, int it);

/* single moves with explicit uniforms (a2, a4, a5); return 1 if accepted */
int orc_bisection_move(orc_t *, int type, int atom, int time, const double *u_gauss, const double *u_acc,
                       int exchange, int *consumed2 /* out: #gauss uniforms, #accept uniforms used; may be NULL */);
int orc_molecular_move(orc_t *, int type, int atom, const double *u3, double u_acc);
int orc_rot3d_step(orc_t *, int it1, int atom0, int type, double r1, double r2, double r3, double r4);
int orc_rotlin_step(orc_t *, int it1, int type, double r1, double r2, double r3);

/* estimators (a14-a17) */
double orc_get_kin(orc_t *);
double orc_get_pot(orc_t *, int with_densities);
double orc_get_rot_energy(orc_t *, double *erotsq, double *eterm);
double orc_get_rot_e3d(orc_t *, double *erotsq, double *eterm);
void   orc_get_rcf(orc_t *, double *rcf0);
void   orc_reset_hist(orc_t *);
void   orc_get_hist(orc_t *, double *gr1d, double *gr2d, double *gr3d_atoms, double *gr3d_mols,
                    double *relthe, double *relphi, double *relchi);

/* area / exchange estimators (a18): instantaneous values on the current state */
void orc_exchange_length(orc_t *, double *ploops /* [numb bosons], += */);            /* GetExchangeLength mc_estim.cc:1997-2019 */
void orc_area_estimators(orc_t *, double *out4 /* area_perp, area_parl, inert_perp, inert_parl */); /* :2087-2250 */
void orc_area_estim3d(orc_t *, int iframe, double *area_proj3, double *inert9);       /* GetAreaEstim3D :2252-2594 */
/* symmetry operations (a19) */
void orc_reflect(orc_t *, int plane /* 0 XZ (REFLECTY), 1 YZ (REFLECTX), 2 XY (REFLECTZ) */); /* mc_piqmc.cc:1385-1708 */
void orc_rotsym(orc_t *, double u, int nfold);                                        /* RotSymConfig mc_piqmc.cc:1710-1794 */
void orc_sched_symmetry(orc_t *, int refl_x, int refl_y, int refl_z, int rotsym, int nfold);

/* worm moves (N1, mc_qworm.cc:93-667) and the world-line mask (a22); worm atoms are numbered inside their type */
void orc_worm_init(orc_t *, int type, double c_input, int m);                 /* WORM line of qmc.input + MCWormInit */
void orc_worm_set(orc_t *, const int *st5 /* exists, ira, masha, atom_i, atom_m */);
void orc_worm_get(orc_t *, int *st5);
void orc_worm_push(orc_t *, int stream, const double *u, int n);              /* explicit uniforms per SPRNG stream number */
void orc_worm_clear(orc_t *);
int  orc_worm_pending(orc_t *, int stream);
void orc_worm_op(orc_t *, int which /* 0 open 1 close 4 advance 5 recede 6 swap, else MCWormMove */, int sched_stream);
void orc_worm_counters(orc_t *, double *total7, double *accep7, double *countqw);
void orc_get_perm(orc_t *, int *pindex, int *rindex, int n);
int  orc_world_line(orc_t *, int atom, int pt);                               /* WorldLine, mc_qworm.cc:553-575 */

/* MRG32k3a (a13): state of the s-th RngStream after SetPackageSeed(seed), and draws */
void orc_mrg_stream_state(const unsigned long *seed6, long stream, double *state6);
void orc_mrg_draws(const unsigned long *seed6, long first_stream, int nstream, int ndraw, double *out);

/* device-schedule replay: the same stream addressing, stage order and accept
   logic as the CUDA path (DESIGN.md "Schedule"), on the CPU */
void orc_sched_seed(orc_t *, const unsigned long *seed6, long chain_global);
void orc_sched_run(orc_t *, long t0, long nsteps);
void orc_sched_counters(orc_t *, double *total6, double *accep6);

#ifdef __cplusplus
}
#endif
#endif
