/* cph_b200.h -- C ABI of libcph_b200.so: the sm_100a CUDA implementation of the
 * per-timestep hot path of LAMMPS `fix constant_pH`.
 *
 * Reference being replaced: MahdiTavakol/Constant_pH, fix_constant_pH.{h,cpp}
 * (cited below as h:N / cpp:N).  Each entry point names the reference interface
 * it stands in for.  Where the reference has nothing (pair
 * re-evaluation, q(lambda), per-site dU/dlambda, neighbour list, restart) the
 * comment says "north_star" and the CPU oracle under oracle/ is the spec.
 *
 * Conventions
 *   - every function returns 0 on success, a negative CPH_ERR_* otherwise; it never
 *     throws and never aborts.  cph_last_error() gives the message.
 *   - plain pointers and sizes only.  `where` selects the address space of the
 *     caller's per-atom buffers: CPH_HOST (pageable or pinned host memory; the
 *     library copies) or CPH_DEVICE (device memory on the handle's GPU; the library
 *     reads/writes it in place on its own stream).
 *   - per-atom arrays are in the CALLER's order (LAMMPS local index 0..nlocal-1);
 *     the library keeps its own cell-sorted order internally.
 *   - one handle per rank/GPU; a handle is not thread-safe, distinct handles are
 *     independent (all kernel constants travel as launch parameters; nothing lives in
 *     device-global state).  All device work is issued on one library-owned stream;
 *     functions that return data to the host synchronise that stream themselves.
 *   - types follow the default LAMMPS build (-DLAMMPS_SMALLBIG): tagint = int32,
 *     bigint = int64.
 *
 * The CPU oracle (oracle/cph_oracle.cpp, test infrastructure only) exports the
 * single-rank subset of these functions with the prefix orc_, so the parity tests
 * drive both through one table (constant_ph_b200/capi.py).
 */
#ifndef CPH_B200_H
#define CPH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cph_handle cph_handle;

/* ---- error codes ------------------------------------------------------------ */
#define CPH_OK            0
#define CPH_ERR_ARG      -1   /* bad argument (the reference's error->all on bad input, cpp:36-45) */
#define CPH_ERR_STATE    -2   /* call out of order (e.g. pair pass before atoms were set) */
#define CPH_ERR_CUDA     -3   /* CUDA runtime failure, or no CUDA device: there is no CPU fallback */
#define CPH_ERR_COMM     -4   /* NCCL / rank-group failure */
#define CPH_ERR_OVERFLOW -5   /* capacity exceeded after automatic regrow failed */
#define CPH_ERR_DOMAIN   -6   /* box / sub-box smaller than the ghost cutoff allows */

#define CPH_HOST   0
#define CPH_DEVICE 1

/* pair styles named by BASELINE.json */
#define CPH_PAIR_LJ_CUT_COUL_CUT 0
#define CPH_PAIR_LJ_CUT_COUL_DSF 1
#define CPH_PAIR_LJ_CUT_COUL_LONG 2  /* real-space part of an Ewald sum, alpha = g_ewald */

/* k-space styles (cph_set_kspace) */
#define CPH_KSPACE_NONE  0
#define CPH_KSPACE_EWALD 1

/* what drives the lambda force (cph_set_mode) */
#define CPH_DUDL_REFERENCE 0  /* HB-HA from the per-atom energy partition, cpp:264-267 + cpp:111 */
#define CPH_DUDL_CHARGE    1  /* north_star: dU/dlambda_s = sum_i dq_i (phi_i + 2 q_i C_self) with q(lambda) */

/* how lambda is advanced */
#define CPH_INTEGRATE_REFERENCE 0  /* one kinematic step inside post_force, cpp:109-117 */
#define CPH_INTEGRATE_VV        1  /* velocity-Verlet halves in initial/final_integrate (north_star) */

/* bias derivative flavour */
#define CPH_BIAS_EXACT      0  /* exact d/dlambda of cpp:132-136, erf in fp64 */
#define CPH_BIAS_AS_WRITTEN 1  /* cpp:123, cpp:137-141 verbatim, erff in fp32 */

/* force scale of cpp:162-170 */
#define CPH_FSCALE_LAMBDA     0  /* f *= lambda, as written (cpp:166-168) */
#define CPH_FSCALE_ONE_MINUS  1  /* f *= (1-lambda), consistent with cpp:114 */

/* ---- lifecycle -------------------------------------------------------------- */
int cph_version(void);
/* FixConstantPH::FixConstantPH (cpp:33) owns one of these per MPI rank. */
int cph_create(int device, cph_handle **out);
/* FixConstantPH::~FixConstantPH (cpp:60-63; leaks H_atom there). */
int cph_destroy(cph_handle *h);
/* error->all message text (cpp:38-45, 183).  h may be NULL (last error of cph_create). */
const char *cph_last_error(cph_handle *h);

/* ---- configuration (any time before cph_set_atoms) ---------------------------- */
/* force->qqrd2e, force->boltz (the undeclared `R` of cpp:111), force->ftm2v (D9). */
int cph_set_units(cph_handle *h, double qqrd2e, double boltz, double ftm2v);
/* The pair style whose eatom the reference reads at cpp:216-219, restated inside the path.
 * epsilon/sigma: (ntypes+1)^2 row-major mixed coefficients (index 0 unused);
 * cut_lj: (ntypes+1)^2 per-pair LJ cutoffs or NULL for cut_lj_global everywhere;
 * special_lj/special_coul: force->special_lj / special_coul [0..3]. */
int cph_set_pair(cph_handle *h, int style, int ntypes, const double *epsilon, const double *sigma,
                 const double *cut_lj, double cut_lj_global, double cut_coul, double alpha,
                 const double *special_lj, const double *special_coul);
/* domain->boxlo/boxhi/periodicity, domain->sublo/subhi, comm->procgrid/myloc, neighbor->skin. */
int cph_set_domain(cph_handle *h, const double *boxlo, const double *boxhi, const int *periodic,
                   const double *sublo, const double *subhi, const int *procgrid, const int *myloc,
                   double skin);
/* Constructor arguments arg[3..8] (cpp:37-49): nevery, hydrogen-group bit, water-group bit, pK, pH, T. */
int cph_set_fix(cph_handle *h, int nevery, int groupHbit, int groupWbit, double pK, double pH, double T);
/* init() constants (cpp:86-96): w s h k a b r m d, m_lambda; plus the derivative flavour. */
int cph_set_bias(cph_handle *h, double w, double s, double hbar, double k, double a, double b,
                 double r, double m, double d, double m_lambda, int bias_mode);
int cph_set_mode(cph_handle *h, int dudl_mode, int integrator_mode, int fscale_mode);
/* north_star's lambda/theta variables (absent from the reference, which integrates lambda itself and
 * confines it with the wall terms U4/U5 of cpp:135-136).  CPH_COORD_THETA: the dynamical coordinate is theta
 * with lambda = sin^2(theta); v_lambda, the mass and the restart record then refer to theta and
 * F_theta = F_lambda * sin(2 theta).  cph_set_lambda still takes lambda (theta = asin(sqrt(lambda))). */
/* Nose-Hoover thermostat on the site velocities at the fix's temperature T (velocity-Verlet form only;
 * absent from the reference): tau = period in time units, 0 = off.  The restart record then also carries
 * the thermostat state (xi, eta, K), and out8[7] of cph_get_scalars is the thermostat energy Q xi^2/2 + S k T eta,
 * which together with H_lambda is conserved for frozen atoms. */
int cph_set_thermostat(cph_handle *h, double tau);
/* Special-bond pairs whose lj AND coul weights are both zero under lj/cut/coul/dsf.  drop == 0 (default, SURVEY.md
 * Appendix A): they stay in the list and contribute -(1-factor_coul)*qqrd2e*qi*qj/r, the undamped term the damped
 * sum over periodic images must not contain -- what the pair style's own factor_coul branch computes.  drop != 0:
 * they are left out of the list and contribute nothing, as under the plain cut styles (what a LAMMPS build does
 * whose Neighbor::init does not match "lj/cut/coul/dsf" against its anchored "^coul/dsf" pattern). */
int cph_set_excluded_policy(cph_handle *h, int drop);
#define CPH_COORD_LAMBDA 0
#define CPH_COORD_THETA  1
int cph_set_coordinate(cph_handle *h, int coordinate);
/* modify_water() (h:58: declared, never defined nor called; TODO at cpp:268; the constructor insists on a
 * 3-atom water group, cpp:44-45).  When enabled (charge mode), every atom of the water group carries
 * q_base - (1/n_W) sum_s lambda_s dQ_s, dQ_s = sum_{t in s}(qB_t - qA_t), so the box keeps the total charge
 * it has at lambda = 0, and dU/dlambda_s gains -(dQ_s/n_W) sum_{a in W} dE/dq_a.  Off by default. */
int cph_set_water_buffer(cph_handle *h, int enable);
/* north_star multi-site tables.  Site s has pK[s]; titratable atom t (global tag titr_tag[t])
 * belongs to site titr_site[t] and has end-state charges qA[t], qB[t].  With nsites == 0 the
 * reference's single global lambda (one site = the whole hydrogen group, pK from cph_set_fix) is used. */
int cph_set_sites(cph_handle *h, int nsites, const double *pK, int ntitr, const int *titr_tag,
                  const int *titr_site, const double *qA, const double *qB);
/* LJ end states (north_star "end states"; absent from the reference, which only rescales forces, cpp:149-171).
 * typeB[t] is the atom type titratable atom t of cph_set_sites has in state B (0: it keeps one LJ identity);
 * its own atom->type is the state-A type.  Such an atom interacts as the lambda-weighted superposition of its two
 * types: E_LJ(i,j) = sum_ab w_i^a w_j^b E_LJ(type_i^a, type_j^b; r), w^A = 1 - lambda_site, w^B = lambda_site
 * (1 and 0 for ordinary atoms), forces likewise, and dU/dlambda_s gains sum_{i in s} sum_j sum_b w_j^b
 * (E_LJ(type_i^B, type_j^b) - E_LJ(type_i^A, type_j^b)), special-bond weights applied.  ntitr must equal the count
 * given to cph_set_sites; call after cph_set_sites / cph_set_pair_style and before cph_set_atoms.  Evaluated by a
 * separate kernel over the few pairs that touch such an atom; the main pair kernel is unchanged. */
int cph_set_lj_states(cph_handle *h, int ntitr, const int *typeB);
/* lambda / v_lambda initial values (never initialised in the reference). */
int cph_set_lambda(cph_handle *h, const double *lambda, const double *v_lambda);

/* ---- rank group: replaces MPI_Allreduce (cpp:274) and comm->reverse_comm (cpp:253) --- */
int cph_comm_unique_id(char *id128);                                   /* ncclGetUniqueId */
int cph_comm_init_nccl(cph_handle *h, int nranks, int rank, const char *id128);

/* ---- atoms: called on every re-neighbouring step (LAMMPS post_neighbor) ------- */
/* atom->x q type tag mask molecule nspecial special of the nlocal OWNED atoms (inside
 * [sublo,subhi)).  The library sorts them into cells, builds its own ghost atoms
 * (periodic images and neighbour-rank copies), maps tags to titration sites and builds
 * the Verlet list (init_list, h:40, never defined in the reference).
 * molecule may be NULL.  When given it is used only to prune the special-bond lookup during
 * the list build (a candidate is compared with special[i] only if molecule[j] == molecule[i]),
 * so it must be LAMMPS-consistent: bonded atoms share a molecule id.  Pass NULL otherwise. */
int cph_set_atoms(cph_handle *h, int where, int nlocal, const double *x, const double *q,
                  const int *type, const int *tag, const int *mask, const int *molecule,
                  const int *nspecial, const int *special, int maxspecial);

/* ---- per step ------------------------------------------------------------------ */
/* atom->x of the owned atoms after the host integrator moved them. */
int cph_set_x(cph_handle *h, int where, const double *x);
/* neighbor->decide(): has any owned atom (on any rank) moved more than skin/2 since the list was built? */
int cph_check_rebuild(cph_handle *h, int *flag);
/* comm->forward_comm(): refresh ghost x and q. */
int cph_forward(cph_handle *h);
/* pair->compute(eflag): forces, and with eflag per-atom energy (the eatom of cpp:217) and the
 * electrostatic potential phi_i. */
int cph_pair_pass(cph_handle *h, int eflag);
/* compute_Hs() (cpp:177-280): HA, HB, per-site HB_s-HA_s and dU/dlambda_s, summed over ranks. */
int cph_site_reduce(cph_handle *h);
/* The other energy sources compute_Hs adds to H_atom (cpp:221-249: bond, angle, dihedral, improper, kspace,
 * fix energies) stay with LAMMPS on the host.  The fix folds their ghost shares (cpp:253, 287-308), partitions
 * them as cpp:264-267 does and hands this rank's two sums over; they enter HA and HB (and HB-HA of the
 * reference's single site) at the next site reduce and are then cleared. */
int cph_set_extra_partition(cph_handle *h, double dHA, double dHB);
/* The same sources seen by north_star's charge derivative: dU/dlambda_s = sum_i dq_i dE/dq_i needs dE/dq_i of
 * every term that depends on the charges.  KSpace (cpp:241-244) stays with LAMMPS; its per-atom energy is
 * e_i = q_i phi_i / 2 with phi_i = dE_kspace/dq_i (E is a quadratic form of the charges), so the fix recovers
 * phi_i = 2 eatom_i / q_i on the titratable atoms it owns and hands over this rank's sums
 * dudl[s] = sum_{i in s, owned} (qB_i - qA_i) phi_i.  They are added to the rank's per-site sums before the
 * all-reduce that replaces cpp:274, at the next site reduce, and are then cleared.  nsites must equal the
 * site count of cph_set_sites. */
int cph_set_extra_dudl(cph_handle *h, int nsites, const double *dudl);
/* ... or the k-space source itself on the device: `kspace_style ewald`, the reciprocal part of
 * the Ewald sum with splitting parameter g_ewald over the wave vectors 2 pi (nx/Lx, ny/Ly, nz/Lz), |n_d| <= k?max,
 * k^2 <= max_d (2 pi k_d max / L_d)^2 -- what force->kspace holds as g_ewald and kxmax/kymax/kzmax.  The pass runs
 * behind every pair pass and adds forces, dE/dq_i and the per-atom energy e_i = q_i phi_i / 2 (self and
 * neutralising-background terms included) to the pair results, so HA/HB (cpp:241-244, 264-267) and every site's
 * dU/dlambda contain the k-space part and follow q(lambda).  Use it with CPH_PAIR_LJ_CUT_COUL_LONG (alpha =
 * g_ewald) for the real-space part.  After cph_set_domain (the wave vectors follow the box; a later
 * cph_set_domain recomputes them); fully periodic boxes only.  The sum is O(atoms x wave vectors): a tool for
 * boxes up to ~1e5 atoms, not a mesh solver.  style CPH_KSPACE_NONE switches it off. */
int cph_set_kspace(cph_handle *h, int style, double g_ewald, int kxmax, int kymax, int kzmax);
/* this rank's share (its owned atoms) of the k-space energy of the last pass with eflag; the sum over ranks is
 * the E_long LAMMPS prints.  It is part of E_coul in cph_get_scalars. */
int cph_get_kspace_energy(cph_handle *h, double *e);
/* calculate_df + calculate_dU + integrate_lambda (cpp:109-145), dt = nevery*update->dt. */
int cph_integrate_lambda(cph_handle *h, double dt);
/* north_star hooks absent from the reference. */
int cph_initial_integrate(cph_handle *h, double dt);
int cph_final_integrate(cph_handle *h, double dt);
/* q_i = (1-lambda_s) qA_i + lambda_s qB_i on the titratable atoms (north_star). */
int cph_apply_charges(cph_handle *h);
/* set_force() (cpp:149-171): scale the forces of hydrogen-group atoms. */
int cph_set_force(cph_handle *h);
/* post_force() (cpp:67-79) in one call: [set_x] forward, pair pass, and on nevery steps
 * site reduce + lambda update; then set_force / charge update as the modes require.
 * x may be NULL (positions already set); f may be NULL (forces stay on the device),
 * otherwise the owned atoms' forces are written to it (caller order). */
int cph_post_force(cph_handle *h, int64_t ntimestep, double dt, int where, const double *x, double *f);
/* setup(int) (h:35: declared, never defined in the reference).  Everything cph_post_force evaluates on an active
 * step -- forces, energy partition, site sums, f/df/U/dU, F_lambda, H_lambda, the force rescale of cpp:149-171 --
 * without the lambda step: LAMMPS calls setup() at the start of every run, and the reference never integrates
 * there. */
int cph_setup(cph_handle *h, int64_t ntimestep, int where, const double *x, double *f);
/* How cph_post_force / cph_setup hand forces to a HOST array: accumulate == 0 stores them, != 0 ADDS them to what
 * the array already holds (the fix under `pair_modify compute no`: atom->f keeps what bonded styles, KSpace and
 * other fixes put there, cpp:149-171 runs on the sum).  Pageable arrays (LAMMPS' atom->x, atom->f) are staged
 * through page-locked memory by the library in pieces that overlap the DMA; page-locked arrays go straight
 * through the copy engine. */
int cph_set_force_mode(cph_handle *h, int accumulate);
/* page-locked host memory for callers that want the direct path (cudaMallocHost / cudaFreeHost) */
int cph_alloc_host(size_t bytes, void **p);
int cph_free_host(void *p);

/* ---- results ---------------------------------------------------------------------- */
int cph_get_forces(cph_handle *h, int where, double *f);     /* nlocal*3 */
int cph_get_eatom(cph_handle *h, int where, double *eatom);  /* nlocal; pair eatom of cpp:217 */
int cph_get_phi(cph_handle *h, int where, double *phi);      /* nlocal; d E_coul / d q_i */
int cph_get_q(cph_handle *h, int where, double *q);          /* nlocal; current charges */
/* out[0]=HA out[1]=HB (cpp:276-277) out[2]=E_vdwl out[3]=E_coul out[4]=H_lambda (cpp:114)
 * out[5]=sum of site kinetic energies out[6]=max displacement^2 at the last check out[7]=thermostat energy */
int cph_get_scalars(cph_handle *h, double *out8);
/* per-site arrays, each nsites long or NULL: lambda, v_lambda, dU/dlambda (charge), HB_s-HA_s,
 * F_lambda (cpp:111), f, df (cpp:122-123), U, dU (cpp:143-144). */
int cph_get_sites(cph_handle *h, double *lambda, double *v_lambda, double *dudl, double *hdiff,
                  double *f_lambda, double *f, double *df, double *U, double *dU);
/* compute_scalar() (h:37) = H_lambda; compute_vector(i) (h:38) = [lambda_s, v_s, dudl_s, F_s] * S */
int cph_compute_scalar(cph_handle *h, double *out);
int cph_compute_vector(cph_handle *h, int i, double *out);
/* memory_usage() (cpp:314-318): bytes held on the device. */
int cph_memory_usage(cph_handle *h, double *bytes);
/* out[0]=nlocal out[1]=nghost out[2]=stored neighbours (sum) out[3]=max per atom out[4]=special pairs
 * out[5]=list builds so far out[6]=titratable atoms owned out[7]=nsites */
int cph_get_counts(cph_handle *h, int64_t *out8);
/* how the per-step ghost refresh travels: 0 = single rank (periodic self images only), 1 = ncclSend/ncclRecv,
 * 2 = stores into the neighbours' receive buffers over NVLink (CUDA IPC peer memory), 3 = the same plus the two
 * small per-step all-reduces (decision flags, site sums) as one-shot stores into peer mailboxes instead of NCCL */
int cph_get_halo_mode(cph_handle *h, int *mode);
/* bookkeeping checks (bit-exact parity): site index of every owned atom (-1 = none), caller order */
int cph_get_site_map(cph_handle *h, int *site_of_atom);
/* neighbour list as sets: numneigh[i] per owned atom (caller order), then keys
 * ((int64)tag_j << 8 | special_class << 5 | image_code), image_code = (ix+1)+3(iy+1)+9(iz+1)
 * of the periodic shift of j.  Pass keys == NULL to get the counts only. */
int cph_get_neighbors(cph_handle *h, int *numneigh, int64_t *keys, int64_t keys_capacity);

/* ---- bonded terms and atom dynamics on the device ------------------ */
/* The reference adds bond->eatom and angle->eatom to H_atom before the HA/HB partition (cpp:221-229).
 * With a topology set, cph_pair_pass / cph_post_force evaluate bond_style harmonic, E = K (r - r0)^2, and
 * angle_style harmonic, E = K (theta - theta0)^2, right behind the pair pass and ADD forces and per-atom
 * energy (1/2 per bond atom, 1/3 per angle atom, LAMMPS' ev_tally) to the pair results, so cph_get_forces,
 * cph_get_eatom and the partition of cph_site_reduce include them; cph_set_extra_partition then only
 * carries what still lives on the host (dihedral, improper, kspace).
 * Coefficient tables are indexed by type, 1-based, [0] unused (bond_coeff / angle_coeff; theta0 in radians). */
int cph_set_bonded(cph_handle *h, int nbondtypes, const double *bond_k, const double *bond_r0,
                   int nangletypes, const double *angle_k, const double *angle_theta0);
/* atom->num_bond, bond_type, bond_atom, num_angle, angle_type, angle_atom1/2/3 of the owned atoms, in the
 * order of the last cph_set_atoms (call it again after every cph_set_atoms), row-major [nlocal][maxbond] /
 * [nlocal][maxangle], host pointers.  Layout of `newton_bond off`: every bond is listed with both of its
 * atoms and every angle with all three, so each rank can evaluate its owned atoms' shares without a
 * reverse exchange.  Partners are resolved through the special-bond tables given to cph_set_atoms. */
int cph_set_topology(cph_handle *h, int nlocal, int maxbond, const int *num_bond, const int *bond_type,
                     const int *bond_atom, int maxangle, const int *num_angle, const int *angle_type,
                     const int *angle_atom1, const int *angle_atom2, const int *angle_atom3);
/* out[0] = E_bond, out[1] = E_angle of the last pass with eflag, summed over ranks. */
int cph_get_bonded_energy(cph_handle *h, double *out2);
/* `fix nve` on the device, so a box can run real dynamics with positions resident in HBM:
 * cph_md_initial_integrate: v += dt/2 ftm2v f/m, x += dt v;  cph_md_final_integrate: v += dt/2 ftm2v f/m.
 * Step: cph_md_initial_integrate, cph_post_force(x = NULL), cph_md_final_integrate.  mass is per type,
 * 1-based.  Velocities follow the atoms through list rebuilds; cph_set_atoms discards them (send them
 * again).  Atoms are remapped into the periodic box at list rebuilds along dimensions the rank spans
 * alone; an atom that drifts more than the skin out of a decomposed sub-box is the host's to migrate
 * (CPH_ERR_DOMAIN, as with host-driven positions). */
int cph_set_mass(cph_handle *h, int ntypes, const double *mass);
int cph_set_v(cph_handle *h, int where, const double *v);    /* nlocal*3, caller order */
int cph_md_initial_integrate(cph_handle *h, double dt);
int cph_md_final_integrate(cph_handle *h, double dt);
int cph_get_x(cph_handle *h, int where, double *x);          /* nlocal*3, caller order */
int cph_get_v(cph_handle *h, int where, double *v);          /* nlocal*3, caller order */

/* ---- restart (absent from the reference; LAMMPS write_restart/restart layout) ---------- */
int cph_restart_size(cph_handle *h, int *ndoubles);
int cph_pack_restart(cph_handle *h, double *buf);
int cph_unpack_restart(cph_handle *h, const double *buf, int ndoubles);

/* ---- timing -------------------------------------------------------------------------- */
int cph_sync(cph_handle *h);
/* Library stream (cudaStream_t as void*), so callers can order their own device work on it. */
int cph_stream(cph_handle *h, void **stream);
/* CUDA-event stopwatch on the library stream: whole region ... */
int cph_timer_start(cph_handle *h);
int cph_timer_stop(cph_handle *h, double *ms);
/* ... and per kernel class while enabled.  which: 0 pair evaluation, 1 inner-list prune (+ fp32 record
 * refresh), 2 site reduce, 3 lambda integrator, 4 charge / force update, 5 halo + allreduce, 6 list build
 * (all stages), 7 new positions + displacement check, 10 k-space structure factors, 11 k-space per-atom sums.  Always
 * available: 8 -> launches = kernels of this library launched so far, 9 -> launches = inner-list prunes so far. */
int cph_profile(cph_handle *h, int enable);
int cph_profile_get(cph_handle *h, int which, double *ms_total, int64_t *launches);

/* out[0] = entries of the pruned inner rows the pair evaluation walks, out[1] = the same with every row padded
 * to a multiple of 32 (= 32 x the loop trips of the evaluation kernel): the work count behind bench.py's
 * fp64 roofline. */
int cph_get_inner_counts(cph_handle *h, int64_t *out2);

/* ---- device microbenchmarks (no handle): the measured fp64 peak SURVEY.md section 7 asks for, and the
 * accuracy of the hardware-seeded 1/sqrt(x), 1/x the pair evaluation is built on ------------------------- */
/* sustained DFMA rate of `device`: warp-level DFMA instructions per second, and 2 x thread-level DFMAs as TFLOP/s */
int cph_bench_fp64_peak(int device, double *dfma_warp_instr_per_s, double *tflops);
/* worst relative error over the argument ranges of the pair kernel: out[0] 1/sqrt seed, out[1] 1/x seed,
 * out[2] refined 1/sqrt, out[3] refined 1/x */
int cph_bench_seed_error(int device, double *out4);
/* orders of the Newton steps behind the seeds this library was built with: 10 * order(1/sqrt) + order(1/x) */
int cph_refine_order(void);

#ifdef __cplusplus
}
#endif
#endif /* CPH_B200_H */
