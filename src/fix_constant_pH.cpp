/* ----------------------------------------------------------------------
   fix constant_pH -- B200 drop-in for MahdiTavakol/Constant_pH fix_constant_pH.cpp

   Syntax (reference cpp:36-49; the keyword loop at cpp:51-54 is empty there):
     fix ID group constant_pH nevery groupH groupW pK pH T [keyword value ...]
       sites FILE            per-site table (north_star multi-site; without it the reference's
                             single global lambda over groupH is used, pK from the arguments)
       dudl charge|reference dU/dlambda from q(lambda) (north_star) or HB-HA of cpp:264-267
       integrator reference|vv   cpp:109-117 in post_force, or velocity-Verlet halves
       fscale lambda|oneminus    cpp:166-168 as written, or (1-lambda)
       bias exact|aswritten      exact derivatives (D13-D16) or cpp:123, 137-141 verbatim
       mlambda VALUE         lambda mass (cpp:96 hard-codes 20)
       lambda0 VALUE         initial lambda when no site file is given (never set in the reference)
       buffer yes|no         modify_water() (h:58): keep the box charge constant through groupW
       coordinate lambda|theta   integrate lambda itself (reference) or theta with lambda = sin^2(theta)
       tlambda TAU           Nose-Hoover thermostat (period TAU) on the site velocities at T; needs integrator vv
       excluded keep|drop    lj/cut/coul/dsf only: keep fully excluded special pairs in the list and subtract their
                             undamped Coulomb term (default), or drop them as plain cut styles do
       ewald KX KY KZ        run `kspace_style ewald` on the device (reciprocal sum over |n_d| <= K?, g_ewald from
                             force->kspace); needs pair lj/cut/coul/long and `kspace_modify compute no`.  Without it
                             a host KSpace style feeds its per-atom energy in as the reference does (cpp:241-244)
       bias_w|bias_s|bias_h|bias_k|bias_a|bias_b|bias_r|bias_m|bias_d VALUE
                             override one constant of the bias potential (init() loads Donnini's table, cpp:86-94)

   The host side stays a LAMMPS Fix; every per-timestep loop of the reference (cpp:149-171,
   cpp:212-267) and the pair arithmetic north_star pulls into the path run in libcph_b200.so.
   There is no CPU fallback: without a CUDA device init() aborts through error->all.
------------------------------------------------------------------------- */

#include "fix_constant_pH.h"

#include "angle.h"
#include "atom.h"
#include "bond.h"
#include "comm.h"
#include "dihedral.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "group.h"
#include "improper.h"
#include "kspace.h"
#include "memory.h"
#include "neighbor.h"
#include "pair.h"
#include "update.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cph_b200.h"

using namespace LAMMPS_NS;
using namespace FixConst;

namespace {

// keywords whose value is one of two words
struct Choice {
  const char *key, *word0, *word1;
  int val0, val1;
};
const Choice kChoices[] = {
    {"dudl", "reference", "charge", CPH_DUDL_REFERENCE, CPH_DUDL_CHARGE},
    {"integrator", "reference", "vv", CPH_INTEGRATE_REFERENCE, CPH_INTEGRATE_VV},
    {"fscale", "lambda", "oneminus", CPH_FSCALE_LAMBDA, CPH_FSCALE_ONE_MINUS},
    {"bias", "exact", "aswritten", CPH_BIAS_EXACT, CPH_BIAS_AS_WRITTEN},
    {"buffer", "no", "yes", 0, 1},
    {"coordinate", "lambda", "theta", CPH_COORD_LAMBDA, CPH_COORD_THETA},
    {"excluded", "keep", "drop", 0, 1},
};
const int kNumChoices = sizeof(kChoices) / sizeof(kChoices[0]);
const char kBiasNames[] = "wshkabrmd";      // bias_<letter> keywords, in the order of cpp:86-94

// one rank per GPU: the launcher's local rank picks the device
int local_device()
{
  for (const char *name : {"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID", "MV2_COMM_WORLD_LOCAL_RANK"})
    if (const char *v = getenv(name)) return atoi(v);
  return 0;
}

}    // namespace

/* ----------------------------------------------------------------------
   constructor: the six positional arguments of the reference (cpp:36-49), then keywords
------------------------------------------------------------------------- */

FixConstantPH::FixConstantPH(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg)
{
  cph = nullptr;
  host_energy = nullptr;
  host_energy_cap = 0;
  force_out = nullptr;
  force_cap = 0;
  charge_buf = nullptr;
  charge_cap = 0;
  pending_restart = nullptr;
  pending_n = 0;
  resend_atoms = true;
  sites_on_device = false;
  rank_group_ready = false;
  part[0] = part[1] = 0.0;
  tab = Sites{0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  opt = Options{CPH_DUDL_REFERENCE, CPH_INTEGRATE_REFERENCE, CPH_FSCALE_LAMBDA, CPH_BIAS_EXACT, 0, CPH_COORD_LAMBDA, 0,
                0.0, 0.5, nullptr, {}, {0, 0, 0}};
  for (double &u : opt.bias_user) u = NAN;
  bias = Bias{0, 0, 0, 0, 0, 0, 0, 0, 0, 20.0};         // mass: cpp:96; the rest is loaded in init()
  lambda_cached = opt.lambda_start;

  const int npositional = 9;                            // ID group style + 6 (cpp:36)
  if (narg < npositional) utils::missing_cmd_args(FLERR, "fix constant_pH", error);

  nevery = utils::inumeric(FLERR, arg[3], false, lmp);
  if (nevery <= 0)                                      // the reference lets 0 through to a modulo
    error->all(FLERR, "Illegal fix constant pH every value {}", nevery);

  auto lookup_group = [&](const char *name, const char *what) {
    const int id = group->find(name);
    if (id < 0) error->all(FLERR, "Cannot find the {} group for fix constant_pH", what);
    return id;
  };
  in.hyd_group = lookup_group(arg[4], "hydrogens");
  in.wat_group = lookup_group(arg[5], "water");
  in.hyd_bit = group->bitmask[in.hyd_group];
  in.wat_bit = group->bitmask[in.wat_group];
  const bigint nwater = group->count(in.wat_group);
  if (nwater != 3)                                      // cpp:44-45
    error->all(FLERR, "Number of atoms in the water molecule for the fix constant_pH is {} instead of three", nwater);

  in.pK = utils::numeric(FLERR, arg[6], false, lmp);
  in.pH = utils::numeric(FLERR, arg[7], false, lmp);
  in.temperature = utils::numeric(FLERR, arg[8], false, lmp);

  int *const choice_target[kNumChoices] = {&opt.dudl, &opt.integrator, &opt.fscale, &opt.bias_form, &opt.buffer, &opt.theta,
                                              &opt.excluded_drop};
  for (int iarg = npositional; iarg < narg; iarg += 2) {
    if (iarg + 1 >= narg) utils::missing_cmd_args(FLERR, "fix constant_pH", error);
    const char *key = arg[iarg], *val = arg[iarg + 1];
    int c = 0;
    while (c < kNumChoices && strcmp(key, kChoices[c].key) != 0) c++;
    if (c < kNumChoices) {
      if (strcmp(val, kChoices[c].word0) == 0) *choice_target[c] = kChoices[c].val0;
      else if (strcmp(val, kChoices[c].word1) == 0) *choice_target[c] = kChoices[c].val1;
      else error->all(FLERR, "Illegal fix constant_pH {} value {}", key, val);
    } else if (strcmp(key, "sites") == 0) {
      opt.site_file = strdup(val);
      opt.dudl = CPH_DUDL_CHARGE;
    } else if (strcmp(key, "mlambda") == 0) {
      bias.mass = utils::numeric(FLERR, val, false, lmp);
      if (bias.mass <= 0.0) error->all(FLERR, "Illegal fix constant_pH mlambda value {}", bias.mass);
    } else if (strcmp(key, "tlambda") == 0) {
      opt.thermostat_period = utils::numeric(FLERR, val, false, lmp);
      if (opt.thermostat_period < 0.0)
        error->all(FLERR, "Illegal fix constant_pH tlambda value {}", opt.thermostat_period);
    } else if (strcmp(key, "ewald") == 0) {
      if (iarg + 3 >= narg) utils::missing_cmd_args(FLERR, "fix constant_pH", error);
      for (int d = 0; d < 3; d++) {
        opt.ewald_kmax[d] = utils::inumeric(FLERR, arg[iarg + 1 + d], false, lmp);
        if (opt.ewald_kmax[d] < 1) error->all(FLERR, "Illegal fix constant_pH ewald value {}", opt.ewald_kmax[d]);
      }
      iarg += 2;      // three values instead of one
    } else if (strcmp(key, "lambda0") == 0) {
      opt.lambda_start = lambda_cached = utils::numeric(FLERR, val, false, lmp);
    } else if (strncmp(key, "bias_", 5) == 0 && key[5] && !key[6] && strchr(kBiasNames, key[5])) {
      opt.bias_user[strchr(kBiasNames, key[5]) - kBiasNames] = utils::numeric(FLERR, val, false, lmp);
    } else {
      error->all(FLERR, "Unknown fix constant_pH keyword: {}", key);
    }
  }

  scalar_flag = 1;          // compute_scalar(): H_lambda (cpp:114)
  vector_flag = 1;          // compute_vector(): [lambda_s, v_s, dU/dlambda_s, F_s] per site
  size_vector = 4;
  global_freq = 1;
  extscalar = 1;
  extvector = 0;
  restart_global = 1;       // write_restart / restart (absent from the reference)
  comm_reverse = 1;         // one double per ghost for the host-tallied energies (cpp:253, 282-284)
  if (opt.site_file) load_site_table(opt.site_file);
}

/* ---------------------------------------------------------------------- */

FixConstantPH::~FixConstantPH()
{
  if (cph) cph_destroy(cph);
  memory->destroy(host_energy);
  void *owned[] = {opt.site_file, tab.pK, tab.lambda0, tab.qA, tab.qB, tab.tag, tab.site, tab.typeB, pending_restart, force_out,
                   charge_buf};
  for (void *p : owned) free(p);
}

/* ---------------------------------------------------------------------- */

int FixConstantPH::setmask()
{
  int hooks = POST_FORCE | POST_NEIGHBOR;     // post_force is the reference's only hook (cpp:67)
  if (opt.integrator == CPH_INTEGRATE_VV) hooks |= INITIAL_INTEGRATE | FINAL_INTEGRATE;
  return hooks;
}

/* ---------------------------------------------------------------------- */

void FixConstantPH::require(int rc, const char *what)
{
  if (rc != CPH_OK) error->all(FLERR, "fix constant_pH: {} failed: {}", what, cph_last_error(cph));
}

/* ----------------------------------------------------------------------
   site table:  line 1 "nsites ntitr"; nsites lines "pK lambda0"; ntitr lines "tag site qA qB [typeB]"
   typeB (optional): the atom type the atom has in state B -- LJ end states, cph_set_lj_states
------------------------------------------------------------------------- */

void FixConstantPH::load_site_table(const char *path)
{
  FILE *fp = fopen(path, "r");
  if (!fp) error->all(FLERR, "Cannot open fix constant_pH site file {}", path);
  if (fscanf(fp, "%d %d", &tab.nsites, &tab.natoms) != 2 || tab.nsites < 1 || tab.natoms < 0)
    error->all(FLERR, "Bad header in fix constant_pH site file {}", path);
  tab.pK = (double *) calloc(tab.nsites, sizeof(double));
  tab.lambda0 = (double *) calloc(tab.nsites, sizeof(double));
  tab.tag = (int *) calloc(tab.natoms + 1, sizeof(int));
  tab.site = (int *) calloc(tab.natoms + 1, sizeof(int));
  tab.qA = (double *) calloc(tab.natoms + 1, sizeof(double));
  tab.qB = (double *) calloc(tab.natoms + 1, sizeof(double));
  tab.typeB = (int *) calloc(tab.natoms + 1, sizeof(int));
  tab.lj_states = 0;
  for (int s = 0; s < tab.nsites; s++)
    if (fscanf(fp, "%lf %lf", &tab.pK[s], &tab.lambda0[s]) != 2)
      error->all(FLERR, "Bad site line {} in fix constant_pH site file", s);
  char line[256];
  if (!fgets(line, sizeof(line), fp)) line[0] = 0;      // rest of the last site line
  for (int t = 0; t < tab.natoms; t++) {
    do {
      if (!fgets(line, sizeof(line), fp)) error->all(FLERR, "Bad atom line {} in fix constant_pH site file", t);
    } while (line[strspn(line, " \t\r\n")] == 0);     // blank lines
    const int got = sscanf(line, "%d %d %lf %lf %d", &tab.tag[t], &tab.site[t], &tab.qA[t], &tab.qB[t], &tab.typeB[t]);
    if (got < 4) error->all(FLERR, "Bad atom line {} in fix constant_pH site file", t);
    if (got < 5) tab.typeB[t] = 0;
    if (tab.typeB[t]) tab.lj_states = 1;
  }
  fclose(fp);
}

/* ----------------------------------------------------------------------
   init: bias constants (cpp:85-96) and the whole configuration of the device side
------------------------------------------------------------------------- */

void FixConstantPH::init()
{
  // Donnini, Ullmann, J Chem Theory Comput 2016, Table S2 -- the values at cpp:86-94
  bias = Bias{200.0, 0.3, 4.0, 2.533, 0.034041, 0.005238, 16.458, 0.1507, 2.0, bias.mass};
  double *slot[9] = {&bias.w, &bias.s, &bias.h, &bias.k, &bias.a, &bias.b, &bias.r, &bias.m, &bias.d};
  for (int c = 0; c < 9; c++)
    if (!std::isnan(opt.bias_user[c])) *slot[c] = opt.bias_user[c];

  if (!atom->q_flag) error->all(FLERR, "fix constant_pH requires atom attribute q");
  if (domain->triclinic) error->all(FLERR, "fix constant_pH does not support triclinic boxes");
  if (opt.thermostat_period > 0.0 && opt.integrator != CPH_INTEGRATE_VV)
    error->all(FLERR, "fix constant_pH tlambda requires integrator vv");

  if (!cph) {
    cph_handle *fresh = nullptr;
    if (cph_create(local_device(), &fresh) != CPH_OK)
      error->all(FLERR, "fix constant_pH: no usable CUDA device: {}", cph_last_error(nullptr));
    cph = fresh;
  }

  // One rank per GPU: the library's rank group stands in for MPI_Allreduce(..., world) (cpp:274) and for the ghost
  // exchange behind comm->reverse_comm(this) (cpp:253).  Its id travels over LAMMPS' own communicator once; the
  // rank inside the group is this rank's place in the processor grid, x fastest, whatever `processors map` did.
  if (comm->nprocs > 1 && !rank_group_ready) {
    char id[128];
    memset(id, 0, sizeof(id));
    int rc = CPH_OK;
    if (comm->me == 0) rc = cph_comm_unique_id(id);
    MPI_Bcast(&rc, 1, MPI_INT, 0, world);
    if (rc != CPH_OK) error->all(FLERR, "fix constant_pH: cannot create the NCCL rank group id");
    MPI_Bcast(id, 128, MPI_BYTE, 0, world);
    const int brick = (comm->myloc[2] * comm->procgrid[1] + comm->myloc[1]) * comm->procgrid[0] + comm->myloc[0];
    require(cph_comm_init_nccl(cph, comm->nprocs, brick, id), "cph_comm_init_nccl");
    rank_group_ready = true;
  }

  require(cph_set_units(cph, force->qqrd2e, force->boltz, force->ftm2v), "cph_set_units");

  // the pair style whose eatom the reference reads (cpp:216-219); its arithmetic runs in the library
  int style = -1;
  Pair *pair = force->pair_match("lj/cut/coul/dsf", 1);
  if (pair) style = CPH_PAIR_LJ_CUT_COUL_DSF;
  else if ((pair = force->pair_match("lj/cut/coul/cut", 1))) style = CPH_PAIR_LJ_CUT_COUL_CUT;
  else if ((pair = force->pair_match("lj/cut/coul/long", 1))) style = CPH_PAIR_LJ_CUT_COUL_LONG;
  if (style < 0)
    error->all(FLERR, "fix constant_pH supports pair styles lj/cut/coul/cut, lj/cut/coul/dsf and lj/cut/coul/long");
  if (style == CPH_PAIR_LJ_CUT_COUL_LONG && !force->kspace)
    error->all(FLERR, "fix constant_pH: pair style lj/cut/coul/long requires a KSpace style");
  if (opt.ewald_kmax[0] > 0) {
    if (style != CPH_PAIR_LJ_CUT_COUL_LONG) error->all(FLERR, "fix constant_pH ewald requires pair style lj/cut/coul/long");
    if (force->kspace->compute_flag)
      error->all(FLERR, "fix constant_pH ewald runs the k-space sum itself: use kspace_modify compute no");
  }
  int dim = 0;
  double **eps = (double **) pair->extract("epsilon", dim);
  double **sig = (double **) pair->extract("sigma", dim);
  const double *cut_coul = (double *) pair->extract("cut_coul", dim);
  const double *cut_lj = (double *) pair->extract("cut_lj", dim);
  const double *alpha = (double *) pair->extract("alpha", dim);
  if (style == CPH_PAIR_LJ_CUT_COUL_LONG) alpha = &force->kspace->g_ewald;    // what the pair style's init_style takes
  if (!eps || !sig || !cut_coul) error->all(FLERR, "fix constant_pH: pair style does not expose epsilon/sigma/cut_coul");
  if (style == CPH_PAIR_LJ_CUT_COUL_DSF && !alpha) error->all(FLERR, "fix constant_pH: pair style does not expose alpha");
  const int nt1 = atom->ntypes + 1;
  std::vector<double> e_tab((size_t) nt1 * nt1, 0.0), s_tab((size_t) nt1 * nt1, 0.0);
  for (int i = 1; i < nt1; i++)
    for (int j = 1; j < nt1; j++) {
      const int lo = std::min(i, j), hi = std::max(i, j);      // Pair::init_one fills i <= j
      e_tab[(size_t) i * nt1 + j] = eps[lo][hi];
      s_tab[(size_t) i * nt1 + j] = sig[lo][hi];
    }
  require(cph_set_pair(cph, style, atom->ntypes, e_tab.data(), s_tab.data(), nullptr, cut_lj ? *cut_lj : *cut_coul,
                       *cut_coul, alpha ? *alpha : 0.0, force->special_lj, force->special_coul), "cph_set_pair");

  require(cph_set_domain(cph, domain->boxlo, domain->boxhi, domain->periodicity, domain->sublo, domain->subhi,
                         comm->procgrid, comm->myloc, neighbor->skin), "cph_set_domain");
  if (opt.ewald_kmax[0] > 0)
    require(cph_set_kspace(cph, CPH_KSPACE_EWALD, force->kspace->g_ewald, opt.ewald_kmax[0], opt.ewald_kmax[1],
                           opt.ewald_kmax[2]), "cph_set_kspace");
  require(cph_set_fix(cph, nevery, in.hyd_bit, in.wat_bit, in.pK, in.pH, in.temperature), "cph_set_fix");
  require(cph_set_bias(cph, bias.w, bias.s, bias.h, bias.k, bias.a, bias.b, bias.r, bias.m, bias.d, bias.mass,
                       opt.bias_form), "cph_set_bias");
  require(cph_set_mode(cph, opt.dudl, opt.integrator, opt.fscale), "cph_set_mode");
  require(cph_set_thermostat(cph, opt.thermostat_period), "cph_set_thermostat");
  require(cph_set_coordinate(cph, opt.theta), "cph_set_coordinate");
  require(cph_set_excluded_policy(cph, opt.excluded_drop), "cph_set_excluded_policy");
  require(cph_set_water_buffer(cph, opt.buffer ? (int) group->count(in.wat_group) : 0), "cph_set_water_buffer");
  // LAMMPS calls init() at the start of EVERY run: the site table and the dynamical state (lambda, v_lambda,
  // a_lambda, thermostat) are sent once; later runs continue from where the previous one stopped.
  if (!sites_on_device) {
    require(cph_set_sites(cph, tab.nsites, tab.pK, tab.natoms, tab.tag, tab.site, tab.qA, tab.qB), "cph_set_sites");
    if (tab.lj_states) require(cph_set_lj_states(cph, tab.natoms, tab.typeB), "cph_set_lj_states");
    if (!pending_restart)
      require(cph_set_lambda(cph, tab.nsites ? tab.lambda0 : &opt.lambda_start, nullptr), "cph_set_lambda");
    sites_on_device = true;
  }
  if (pending_restart) {
    require(cph_unpack_restart(cph, pending_restart, pending_n), "cph_unpack_restart");
    free(pending_restart);
    pending_restart = nullptr;
  }
  size_vector = 4 * std::max(tab.nsites, 1);
  resend_atoms = true;
}

/* ---------------------------------------------------------------------- */

void FixConstantPH::init_list(int /*id*/, NeighList * /*ptr*/)
{
  // The reference declares this hook (h:40) but never requests a list.  The Verlet list of
  // this fix lives on the device and is built by cph_set_atoms / cph_post_force.
}

/* ---------------------------------------------------------------------- */

void FixConstantPH::upload_atoms()
{
  const int n = atom->nlocal;
  const bool specials = atom->maxspecial > 0 && atom->nspecial && atom->special && n > 0;
  require(cph_set_atoms(cph, CPH_HOST, n, n ? atom->x[0] : nullptr, atom->q, atom->type, atom->tag, atom->mask,
                        atom->molecule_flag ? atom->molecule : nullptr, specials ? atom->nspecial[0] : nullptr,
                        specials ? atom->special[0] : nullptr, specials ? atom->maxspecial : 0), "cph_set_atoms");
  resend_atoms = false;
}

void FixConstantPH::post_neighbor()
{
  resend_atoms = true;     // LAMMPS migrated and re-sorted atoms: local indices changed
}

void FixConstantPH::setup(int /*vflag*/)
{
  // The reference declares setup (h:35) without a body and never integrates there: evaluate everything
  // (forces, partition, F_lambda, H_lambda, force rescale) at the current lambda, leave lambda where it is.
  upload_atoms();
  compute_Hs();
  run_device_step(true);
}

/* ----------------------------------------------------------------------
   velocity-Verlet halves of the lambda dynamics.  LAMMPS calls these hooks on every step; lambda lives on the
   nevery grid of post_force (cpp:69), so both act only on the steps whose post_force is active, with
   t_lambda = nevery*dt (cpp:113): kick+drift at the start of such a step, the closing kick after its forces.
------------------------------------------------------------------------- */

void FixConstantPH::initial_integrate(int /*vflag*/)
{
  if (update->ntimestep % nevery) return;
  require(cph_initial_integrate(cph, update->dt * nevery), "cph_initial_integrate");
}

void FixConstantPH::final_integrate()
{
  if (update->ntimestep % nevery) return;
  require(cph_final_integrate(cph, update->dt * nevery), "cph_final_integrate");
}

/* ----------------------------------------------------------------------
   post_force (cpp:67-79): on nevery steps the energy partition, df, dU and the lambda step;
   on every step the force rescale.  One library call runs the sequence on the device.
------------------------------------------------------------------------- */

void FixConstantPH::post_force(int /*vflag*/)
{
  if (resend_atoms) upload_atoms();
  if (update->ntimestep % nevery == 0) compute_Hs();       // host-tallied energy sources (cpp:221-253)
  run_device_step(false);
}

/* ----------------------------------------------------------------------
   one device step: positions in, pair forces out.  Under `pair_modify compute no` the GPU pair pass IS the pair
   computation: the library ADDS its forces to atom->f (cph_set_force_mode) while it copies them back, so no
   host loop over 3N doubles runs here.  setup_only: cph_setup instead of cph_post_force (no lambda step).
------------------------------------------------------------------------- */

void FixConstantPH::run_device_step(bool setup_only)
{
  const int n = atom->nlocal;
  const bool lammps_owns_pair_forces = force->pair && force->pair->compute_flag;
  const bool single_site_rescale = opt.dudl == CPH_DUDL_REFERENCE && tab.nsites == 0;
  double *x = n ? atom->x[0] : nullptr;
  double *f = (n && !lammps_owns_pair_forces) ? atom->f[0] : nullptr;

  // The reference scales the TOTAL force on the hydrogen group (cpp:162-170).  The library scales the pair part
  // it owns; whatever else LAMMPS put into atom->f (bonded terms, other fixes) gets the factor of the lambda the
  // step starts from here, before the pair part is added -- except when lambda is about to move (reference
  // integrator on an active step): then the new lambda is read back first and the scaling follows the add.
  require(cph_set_force_mode(cph, 1), "cph_set_force_mode");
  if (setup_only) require(cph_setup(cph, update->ntimestep, CPH_HOST, x, f), "cph_setup");
  else require(cph_post_force(cph, update->ntimestep, update->dt, CPH_HOST, x, f), "cph_post_force");

  if (single_site_rescale) {
    require(cph_get_sites(cph, &lambda_cached, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr),
            "cph_get_sites");
    if (lammps_owns_pair_forces) set_force();           // LAMMPS computed the pair forces itself: the factor applies to everything
    else if (host_forces_present()) scale_host_part();  // bonded / kspace / other fixes: same factor as the pair part got
  }
  // q(lambda) back into atom->q where the host consumes charges (its own pair compute, KSpace)
  if (opt.dudl == CPH_DUDL_CHARGE && (lammps_owns_pair_forces || force->kspace)) pull_charges();
}

/* ----------------------------------------------------------------------
   set_force (cpp:149-171): multiply the forces of the hydrogen group
------------------------------------------------------------------------- */

void FixConstantPH::scale_hydrogen_forces(double factor)
{
  const int n = atom->nlocal;
  const int *mask = atom->mask;
  for (int i = 0; i < n; i++) {
    if (!(mask[i] & in.hyd_bit)) continue;
    double *fi = atom->f[i];
    for (int c = 0; c < 3; c++) fi[c] *= factor;
  }
}

void FixConstantPH::set_force()
{
  scale_hydrogen_forces(opt.fscale == CPH_FSCALE_LAMBDA ? lambda_cached : 1.0 - lambda_cached);
}

bool FixConstantPH::host_forces_present() const
{
  return force->bond || force->angle || force->dihedral || force->improper || (force->kspace && force->kspace->compute_flag);
}

// atom->f = (host part) + (pair part, already scaled on the device by s): make the host part follow,
// f <- s*f_host + f_pair = s*(f - f_pair) + f_pair, for the hydrogen group only.  The pair part of those few
// atoms is fetched from the device.
void FixConstantPH::scale_host_part()
{
  const int n = atom->nlocal;
  const double s = opt.fscale == CPH_FSCALE_LAMBDA ? lambda_cached : 1.0 - lambda_cached;
  if (n > force_cap) {
    force_cap = n + n / 8 + 16;
    free(force_out);
    force_out = (double *) malloc(sizeof(double) * 3 * force_cap);
  }
  require(cph_get_forces(cph, CPH_HOST, force_out), "cph_get_forces");
  const int *mask = atom->mask;
  for (int i = 0; i < n; i++) {
    if (!(mask[i] & in.hyd_bit)) continue;
    for (int c = 0; c < 3; c++) {
      const double fp = force_out[3 * i + c];
      atom->f[i][c] = s * (atom->f[i][c] - fp) + fp;
    }
  }
}

// q_i(lambda) of the titratable atoms (and of the water buffer) from the device into atom->q
void FixConstantPH::pull_charges()
{
  const int n = atom->nlocal;
  if (n > charge_cap) {
    charge_cap = n + n / 8 + 16;
    free(charge_buf);
    charge_buf = (double *) malloc(sizeof(double) * charge_cap);
  }
  require(cph_get_q(cph, CPH_HOST, charge_buf), "cph_get_q");
  memcpy(atom->q, charge_buf, sizeof(double) * n);
}

/* ----------------------------------------------------------------------
   compute_Hs (cpp:177-280).  The pair part of the per-atom energy is produced and partitioned
   on the device.  This is the rest of that routine, for the sources LAMMPS keeps on the host
   (bond, angle, dihedral, improper, kspace: cpp:221-244): accumulate them per atom, fold the
   ghost shares back (cpp:253), split into "all atoms" and "all but the hydrogen group"
   (cpp:264-267) and hand the two sums to the library, which adds them before the all-reduce
   that replaces cpp:274.
------------------------------------------------------------------------- */

void FixConstantPH::compute_Hs()
{
  struct Source {
    const double *eatom;
    int count;
  };
  const int nlocal = atom->nlocal, nall = nlocal + atom->nghost;
  const int nbonded = force->newton_bond ? nall : nlocal;                       // cpp:206
  const bool tip4p = force->kspace && force->kspace->tip4pflag;
  const int nkspace = tip4p ? nall : nlocal;                                    // cpp:208
  std::vector<Source> sources;
  if (force->bond && force->bond->eatom) sources.push_back({force->bond->eatom, nbonded});
  if (force->angle && force->angle->eatom) sources.push_back({force->angle->eatom, nbonded});
  if (force->dihedral && force->dihedral->eatom) sources.push_back({force->dihedral->eatom, nbonded});
  if (force->improper && force->improper->eatom) sources.push_back({force->improper->eatom, nbonded});
  if (force->kspace && force->kspace->compute_flag && force->kspace->eatom)
    sources.push_back({force->kspace->eatom, nkspace});
  if (sources.empty()) return;

  if (update->eflag_atom != update->ntimestep)                                  // cpp:181-183
    error->all(FLERR, "Per-atom energy was not tallied on needed timestep");
  if (atom->nmax > host_energy_cap) {                                           // cpp:188-192
    memory->destroy(host_energy);
    host_energy_cap = atom->nmax;
    memory->create(host_energy, host_energy_cap, "constant_pH:host_energy");
  }

  std::fill(host_energy, host_energy + std::min(nall, host_energy_cap), 0.0);
  for (const Source &src : sources)
    for (int i = 0; i < src.count; i++) host_energy[i] += src.eatom[i];
  if (force->newton || tip4p) comm->reverse_comm(this);                         // cpp:253

  double everyone = 0.0, without_hydrogens = 0.0;
  const int *mask = atom->mask;
  for (int i = 0; i < nlocal; i++) {
    everyone += host_energy[i];
    if ((mask[i] & in.hyd_bit) == 0) without_hydrogens += host_energy[i];
  }
  require(cph_set_extra_partition(cph, everyone, without_hydrogens), "cph_set_extra_partition");

  // north_star's charge derivative needs the k-space share of dE/dq_i as well
  if (opt.dudl == CPH_DUDL_CHARGE && tab.nsites > 0 && force->kspace && force->kspace->compute_flag &&
      force->kspace->eatom)
    kspace_site_derivative(force->kspace->eatom);
}

/* ----------------------------------------------------------------------
   KSpace (cpp:241-244) stays with LAMMPS.  Its energy is a quadratic form of the charges, so the per-atom
   energy it tallies is e_i = q_i phi_i / 2 with phi_i = dE_kspace/dq_i (self and neutralising terms included).
   dU/dlambda_s gains sum_{i in s} (qB_i - qA_i) phi_i over the titratable atoms this rank owns; the library
   adds the sums to its own before the all-reduce.  An atom whose charge is exactly zero at this lambda gives
   no handle on phi_i and contributes nothing.
------------------------------------------------------------------------- */

void FixConstantPH::kspace_site_derivative(const double *ekspace)
{
  if (force->kspace->tip4pflag)
    error->all(FLERR, "fix constant_pH dudl charge cannot take the k-space potential from a TIP4P KSpace style");
  if (atom->map_style == Atom::MAP_NONE)
    error->all(FLERR, "fix constant_pH with a KSpace style requires an atom map, see atom_modify");
  std::vector<double> site_sum(tab.nsites, 0.0);
  const int nlocal = atom->nlocal;
  for (int t = 0; t < tab.natoms; t++) {
    const int i = atom->map(tab.tag[t]);
    if (i < 0 || i >= nlocal) continue;
    const double q = atom->q[i];
    if (q == 0.0) continue;
    site_sum[tab.site[t]] += (tab.qB[t] - tab.qA[t]) * 2.0 * ekspace[i] / q;
  }
  require(cph_set_extra_dudl(cph, tab.nsites, site_sum.data()), "cph_set_extra_dudl");
}

/* ---------------------------------------------------------------------- */

double FixConstantPH::compute_scalar()
{
  double value = 0.0;
  require(cph_compute_scalar(cph, &value), "cph_compute_scalar");
  return value;
}

double FixConstantPH::compute_vector(int i)
{
  double value = 0.0;
  require(cph_compute_vector(cph, i, &value), "cph_compute_vector");
  return value;
}

/* ----------------------------------------------------------------------
   memory usage: the host-side per-atom arrays (cpp:310-318) plus the device buffers
------------------------------------------------------------------------- */

double FixConstantPH::memory_usage()
{
  double device_bytes = 0.0;
  if (cph) cph_memory_usage(cph, &device_bytes);
  return device_bytes + sizeof(double) * ((double) host_energy_cap + 3.0 * force_cap);
}

/* ----------------------------------------------------------------------
   restart: [version, S, (lambda|theta, v, a) * S, (xi, eta, K if tlambda)] as doubles,
   LAMMPS global-restart layout
------------------------------------------------------------------------- */

void FixConstantPH::write_restart(FILE *fp)
{
  int n = 0;
  require(cph_restart_size(cph, &n), "cph_restart_size");
  std::vector<double> record(n);
  require(cph_pack_restart(cph, record.data()), "cph_pack_restart");
  if (comm->me == 0) {
    const int bytes = n * (int) sizeof(double);
    fwrite(&bytes, sizeof(int), 1, fp);
    fwrite(record.data(), sizeof(double), n, fp);
  }
}

void FixConstantPH::restart(char *buf)
{
  const double *record = (const double *) buf;
  const int nsites_in_record = (int) record[1];
  pending_n = 2 + 3 * nsites_in_record + (opt.thermostat_period > 0.0 ? 3 : 0);
  free(pending_restart);
  pending_restart = (double *) malloc(sizeof(double) * pending_n);
  memcpy(pending_restart, record, sizeof(double) * pending_n);
  if (cph) {      // already initialised: apply now
    require(cph_unpack_restart(cph, pending_restart, pending_n), "cph_unpack_restart");
    free(pending_restart);
    pending_restart = nullptr;
  }
}

/* ----------------------------------------------------------------------
   reverse communication of the host-tallied per-atom energies (the reference has these bodies at
   cpp:287-308 but does not declare them).  The pair energies live on the device, where a full
   neighbour list leaves nothing to fold.
------------------------------------------------------------------------- */

int FixConstantPH::pack_reverse_comm(int n, int first, double *buf)
{
  memcpy(buf, host_energy + first, sizeof(double) * n);
  return n;
}

void FixConstantPH::unpack_reverse_comm(int n, int *list, double *buf)
{
  for (int k = 0; k < n; k++) host_energy[list[k]] += buf[k];
}
