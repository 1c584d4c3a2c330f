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

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "cph_b200.h"

using namespace LAMMPS_NS;
using namespace FixConst;

/* ---------------------------------------------------------------------- */

FixConstantPH::FixConstantPH(LAMMPS *lmp, int narg, char **arg) :
  Fix(lmp, narg, arg), H_atom(nullptr), cph(nullptr), sitefile(nullptr), site_pK(nullptr),
  site_lambda0(nullptr), titr_qA(nullptr), titr_qB(nullptr), titr_tag(nullptr), titr_site(nullptr),
  restart_buf(nullptr), xbuf(nullptr), fbuf(nullptr)
{
  if (narg < 9) utils::missing_cmd_args(FLERR, "fix constant_pH", error);              // cpp:36 (D3)
  nevery = utils::inumeric(FLERR, arg[3], false, lmp);                                 // cpp:37
  if (nevery <= 0) error->all(FLERR, "Illegal fix constant pH every value {}", nevery); // cpp:38 (D4)
  igroupH = group->find(arg[4]);                                                       // cpp:39
  if (igroupH == -1) error->all(FLERR, "Cannot find the hydrogens group for fix constant_pH");  // cpp:40 (D5)
  groupHbit = group->bitmask[igroupH];                                                 // cpp:41
  igroupW = group->find(arg[5]);                                                       // cpp:42
  if (igroupW == -1) error->all(FLERR, "Cannot find the water group for fix constant_pH");      // cpp:43
  if (group->count(igroupW) != 3)                                                      // cpp:44-45
    error->all(FLERR, "Number of atoms in the water molecule for the fix constant_pH is {} instead of three",
               group->count(igroupW));
  groupWbit = group->bitmask[igroupW];                                                 // cpp:46
  pK = utils::numeric(FLERR, arg[6], false, lmp);                                      // cpp:47
  pH = utils::numeric(FLERR, arg[7], false, lmp);                                      // cpp:48
  T = utils::numeric(FLERR, arg[8], false, lmp);                                       // cpp:49

  dudl_mode = CPH_DUDL_REFERENCE;
  integrator_mode = CPH_INTEGRATE_REFERENCE;
  fscale_mode = CPH_FSCALE_LAMBDA;
  bias_mode = CPH_BIAS_EXACT;
  m_lambda = 20.0;                                                                     // cpp:96
  water_buffer = 0;
  coord_theta = 0;
  t_lambda_period = 0.0;
  lambda_host = 0.5;
  nsites = ntitr = 0;
  restart_n = 0;
  atoms_sent = false;
  bufmax = 0;
  nmax = 1;
  HA = HB = 0.0;

  int iarg = 9;                                                                        // cpp:51
  while (iarg < narg) {                                                                // cpp:52 (D6: advance or fail)
    if (iarg + 1 >= narg) utils::missing_cmd_args(FLERR, "fix constant_pH", error);
    const char *key = arg[iarg], *val = arg[iarg + 1];
    if (strcmp(key, "sites") == 0) {
      sitefile = strdup(val);
      dudl_mode = CPH_DUDL_CHARGE;
    } else if (strcmp(key, "dudl") == 0) {
      if (strcmp(val, "charge") == 0) dudl_mode = CPH_DUDL_CHARGE;
      else if (strcmp(val, "reference") == 0) dudl_mode = CPH_DUDL_REFERENCE;
      else error->all(FLERR, "Illegal fix constant_pH dudl value {}", val);
    } else if (strcmp(key, "integrator") == 0) {
      if (strcmp(val, "vv") == 0) integrator_mode = CPH_INTEGRATE_VV;
      else if (strcmp(val, "reference") == 0) integrator_mode = CPH_INTEGRATE_REFERENCE;
      else error->all(FLERR, "Illegal fix constant_pH integrator value {}", val);
    } else if (strcmp(key, "fscale") == 0) {
      if (strcmp(val, "oneminus") == 0) fscale_mode = CPH_FSCALE_ONE_MINUS;
      else if (strcmp(val, "lambda") == 0) fscale_mode = CPH_FSCALE_LAMBDA;
      else error->all(FLERR, "Illegal fix constant_pH fscale value {}", val);
    } else if (strcmp(key, "bias") == 0) {
      if (strcmp(val, "aswritten") == 0) bias_mode = CPH_BIAS_AS_WRITTEN;
      else if (strcmp(val, "exact") == 0) bias_mode = CPH_BIAS_EXACT;
      else error->all(FLERR, "Illegal fix constant_pH bias value {}", val);
    } else if (strcmp(key, "mlambda") == 0) {
      m_lambda = utils::numeric(FLERR, val, false, lmp);
      if (m_lambda <= 0.0) error->all(FLERR, "Illegal fix constant_pH mlambda value {}", m_lambda);
    } else if (strcmp(key, "tlambda") == 0) {
      t_lambda_period = utils::numeric(FLERR, val, false, lmp);
      if (t_lambda_period < 0.0) error->all(FLERR, "Illegal fix constant_pH tlambda value {}", t_lambda_period);
    } else if (strcmp(key, "coordinate") == 0) {
      if (strcmp(val, "theta") == 0) coord_theta = 1;
      else if (strcmp(val, "lambda") == 0) coord_theta = 0;
      else error->all(FLERR, "Illegal fix constant_pH coordinate value {}", val);
    } else if (strcmp(key, "buffer") == 0) {
      if (strcmp(val, "yes") == 0) water_buffer = 1;
      else if (strcmp(val, "no") == 0) water_buffer = 0;
      else error->all(FLERR, "Illegal fix constant_pH buffer value {}", val);
    } else if (strcmp(key, "lambda0") == 0) {
      lambda_host = utils::numeric(FLERR, val, false, lmp);
    } else {
      error->all(FLERR, "Unknown fix constant_pH keyword: {}", key);
    }
    iarg += 2;
  }

  scalar_flag = 1;          // compute_scalar(): H_lambda (cpp:114)
  vector_flag = 1;          // compute_vector(): [lambda_s, v_s, dU/dlambda_s, F_s] per site
  size_vector = 4;
  global_freq = 1;
  extscalar = 1;
  extvector = 0;
  restart_global = 1;       // write_restart / restart (absent from the reference)
  comm_reverse = 1;         // cpp:253, 282-284 (D18)
  if (sitefile) read_sites(sitefile);
}

/* ---------------------------------------------------------------------- */

FixConstantPH::~FixConstantPH()
{
  if (cph) cph_destroy(cph);
  memory->destroy(H_atom);                                                             // D7
  free(sitefile);
  free(site_pK); free(site_lambda0); free(titr_qA); free(titr_qB); free(titr_tag); free(titr_site);
  free(restart_buf); free(xbuf); free(fbuf);
}

/* ---------------------------------------------------------------------- */

int FixConstantPH::setmask()
{
  int mask = 0;
  mask |= POST_FORCE;                    // the reference's only hook (cpp:67)
  mask |= POST_NEIGHBOR;                 // atoms were migrated / re-sorted: resend them
  if (integrator_mode == CPH_INTEGRATE_VV) {
    mask |= INITIAL_INTEGRATE;
    mask |= FINAL_INTEGRATE;
  }
  return mask;
}

/* ---------------------------------------------------------------------- */

void FixConstantPH::check(int rc, const char *what)
{
  if (rc == CPH_OK) return;
  error->all(FLERR, "fix constant_pH: {} failed: {}", what, cph_last_error(cph));
}

/* ----------------------------------------------------------------------
   site table:  line 1 "nsites ntitr"; nsites lines "pK lambda0"; ntitr lines "tag site qA qB"
------------------------------------------------------------------------- */

void FixConstantPH::read_sites(const char *path)
{
  FILE *fp = fopen(path, "r");
  if (!fp) error->all(FLERR, "Cannot open fix constant_pH site file {}", path);
  if (fscanf(fp, "%d %d", &nsites, &ntitr) != 2 || nsites < 1 || ntitr < 0)
    error->all(FLERR, "Bad header in fix constant_pH site file {}", path);
  site_pK = (double *) malloc(sizeof(double) * nsites);
  site_lambda0 = (double *) malloc(sizeof(double) * nsites);
  titr_tag = (int *) malloc(sizeof(int) * (ntitr + 1));
  titr_site = (int *) malloc(sizeof(int) * (ntitr + 1));
  titr_qA = (double *) malloc(sizeof(double) * (ntitr + 1));
  titr_qB = (double *) malloc(sizeof(double) * (ntitr + 1));
  for (int s = 0; s < nsites; s++)
    if (fscanf(fp, "%lf %lf", &site_pK[s], &site_lambda0[s]) != 2)
      error->all(FLERR, "Bad site line {} in fix constant_pH site file", s);
  for (int t = 0; t < ntitr; t++)
    if (fscanf(fp, "%d %d %lf %lf", &titr_tag[t], &titr_site[t], &titr_qA[t], &titr_qB[t]) != 4)
      error->all(FLERR, "Bad atom line {} in fix constant_pH site file", t);
  fclose(fp);
}

/* ---------------------------------------------------------------------- */

void FixConstantPH::init()
{
  // default values from Donnini, Ullmann, J Chem Theory Comput 2016 - Table S2   (cpp:85-94)
  w = 200.0;
  s = 0.3;
  h = 4.0;
  k = 2.533;
  a = 0.034041;
  b = 0.005238;
  r = 16.458;
  m = 0.1507;
  d = 2.0;

  if (atom->nmax > nmax) {                                                             // cpp:100-104
    memory->destroy(H_atom);
    nmax = atom->nmax;
    memory->create(H_atom, nmax, "constant_pH:H_atom");
  }

  if (!atom->q_flag) error->all(FLERR, "fix constant_pH requires atom attribute q");
  if (domain->triclinic) error->all(FLERR, "fix constant_pH does not support triclinic boxes");

  if (!cph) {
    int dev = 0;
    const char *lr = getenv("LOCAL_RANK");
    if (!lr) lr = getenv("OMPI_COMM_WORLD_LOCAL_RANK");
    if (!lr) lr = getenv("SLURM_LOCALID");
    if (lr) dev = atoi(lr);
    cph_handle *hnd = nullptr;
    int rc = cph_create(dev, &hnd);
    if (rc != CPH_OK) error->all(FLERR, "fix constant_pH: no usable CUDA device: {}", cph_last_error(nullptr));
    cph = hnd;
  }

  check(cph_set_units(cph, force->qqrd2e, force->boltz, force->ftm2v), "cph_set_units");   // D8, D9

  // the pair style whose eatom the reference reads (cpp:216-219); its arithmetic runs in the library
  int style = -1;
  Pair *pair = force->pair_match("lj/cut/coul/dsf", 1);
  if (pair) style = CPH_PAIR_LJ_CUT_COUL_DSF;
  else if ((pair = force->pair_match("lj/cut/coul/cut", 1))) style = CPH_PAIR_LJ_CUT_COUL_CUT;
  if (style < 0) error->all(FLERR, "fix constant_pH supports pair styles lj/cut/coul/cut and lj/cut/coul/dsf");
  int dim = 0;
  double **eps = (double **) pair->extract("epsilon", dim);
  double **sig = (double **) pair->extract("sigma", dim);
  double *cut_coul = (double *) pair->extract("cut_coul", dim);
  double *cut_lj = (double *) pair->extract("cut_lj", dim);
  double *alpha = (double *) pair->extract("alpha", dim);
  if (!eps || !sig || !cut_coul) error->all(FLERR, "fix constant_pH: pair style does not expose epsilon/sigma/cut_coul");
  if (style == CPH_PAIR_LJ_CUT_COUL_DSF && !alpha) error->all(FLERR, "fix constant_pH: pair style does not expose alpha");
  const int nt = atom->ntypes;
  double *e1 = (double *) calloc((size_t)(nt + 1) * (nt + 1), sizeof(double));
  double *s1 = (double *) calloc((size_t)(nt + 1) * (nt + 1), sizeof(double));
  for (int i = 1; i <= nt; i++)
    for (int j = 1; j <= nt; j++) {
      const int lo = i < j ? i : j, hi = i < j ? j : i;      // init_one fills i <= j
      e1[i * (nt + 1) + j] = eps[lo][hi];
      s1[i * (nt + 1) + j] = sig[lo][hi];
    }
  check(cph_set_pair(cph, style, nt, e1, s1, nullptr, cut_lj ? *cut_lj : *cut_coul, *cut_coul, alpha ? *alpha : 0.0,
                     force->special_lj, force->special_coul), "cph_set_pair");
  free(e1);
  free(s1);

  check(cph_set_domain(cph, domain->boxlo, domain->boxhi, domain->periodicity, domain->sublo, domain->subhi,
                       comm->procgrid, comm->myloc, neighbor->skin), "cph_set_domain");
  check(cph_set_fix(cph, nevery, groupHbit, groupWbit, pK, pH, T), "cph_set_fix");
  check(cph_set_bias(cph, w, s, h, k, a, b, r, m, d, m_lambda, bias_mode), "cph_set_bias");   // cpp:86-96
  check(cph_set_mode(cph, dudl_mode, integrator_mode, fscale_mode), "cph_set_mode");
  if (t_lambda_period > 0.0 && integrator_mode != CPH_INTEGRATE_VV)
    error->all(FLERR, "fix constant_pH tlambda requires integrator vv");
  check(cph_set_thermostat(cph, t_lambda_period), "cph_set_thermostat");
  check(cph_set_coordinate(cph, coord_theta ? CPH_COORD_THETA : CPH_COORD_LAMBDA), "cph_set_coordinate");
  check(cph_set_water_buffer(cph, water_buffer ? (int) group->count(igroupW) : 0), "cph_set_water_buffer");
  check(cph_set_sites(cph, nsites, site_pK, ntitr, titr_tag, titr_site, titr_qA, titr_qB), "cph_set_sites");
  if (restart_buf) {
    check(cph_unpack_restart(cph, restart_buf, restart_n), "cph_unpack_restart");
    free(restart_buf);
    restart_buf = nullptr;
  } else if (nsites) {
    check(cph_set_lambda(cph, site_lambda0, nullptr), "cph_set_lambda");
  } else {
    check(cph_set_lambda(cph, &lambda_host, nullptr), "cph_set_lambda");
  }
  size_vector = 4 * (nsites ? nsites : 1);
  atoms_sent = false;
}

/* ---------------------------------------------------------------------- */

void FixConstantPH::init_list(int /*id*/, NeighList * /*ptr*/)
{
  // The reference declares this hook (h:40) but never requests a list.  The Verlet list of
  // this fix lives on the device and is built by cph_set_atoms / cph_post_force.
}

/* ---------------------------------------------------------------------- */

void FixConstantPH::send_atoms()
{
  const int nlocal = atom->nlocal;
  const double *x = nlocal ? &atom->x[0][0] : nullptr;
  const int *nspecial = (atom->maxspecial && atom->nspecial && nlocal) ? &atom->nspecial[0][0] : nullptr;
  const int *special = (atom->maxspecial && atom->special && nlocal) ? &atom->special[0][0] : nullptr;
  check(cph_set_atoms(cph, CPH_HOST, nlocal, x, atom->q, atom->type, atom->tag, atom->mask,
                      atom->molecule_flag ? atom->molecule : nullptr, nspecial, special,
                      nspecial ? atom->maxspecial : 0), "cph_set_atoms");
  if (nlocal > bufmax) {
    bufmax = nlocal + nlocal / 8 + 16;
    free(fbuf);
    fbuf = (double *) malloc(sizeof(double) * 3 * bufmax);
  }
  atoms_sent = true;
}

void FixConstantPH::post_neighbor()
{
  atoms_sent = false;      // LAMMPS migrated and re-sorted atoms: local indices changed
}

void FixConstantPH::setup(int vflag)
{
  send_atoms();
  post_force(vflag);       // as most fixes do; the reference declares setup (h:35) without a body
}

/* ---------------------------------------------------------------------- */

void FixConstantPH::initial_integrate(int /*vflag*/)
{
  check(cph_initial_integrate(cph, update->dt * nevery), "cph_initial_integrate");
}

void FixConstantPH::final_integrate()
{
  check(cph_final_integrate(cph, update->dt * nevery), "cph_final_integrate");
}

/* ---------------------------------------------------------------------- */

void FixConstantPH::post_force(int /*vflag*/)
{
  if (!atoms_sent) send_atoms();
  const int nlocal = atom->nlocal;
  const double *x = nlocal ? &atom->x[0][0] : nullptr;

  // cpp:221-253: energy sources that stay with LAMMPS on the host enter the HA/HB partition
  if (update->ntimestep % nevery == 0) compute_Hs();

  // cpp:69-78: on nevery steps compute_Hs, calculate_df, calculate_dU, integrate_lambda;
  // set_force on every step.  One library call runs the whole sequence on the device.
  check(cph_post_force(cph, update->ntimestep, update->dt, CPH_HOST, x, fbuf), "cph_post_force");

  double sc[8];
  check(cph_get_scalars(cph, sc), "cph_get_scalars");
  HA = sc[0];                                                                          // cpp:276
  HB = sc[1];                                                                          // cpp:277

  const bool pair_on_host = force->pair && force->pair->compute_flag;
  double **f = atom->f;
  if (dudl_mode == CPH_DUDL_REFERENCE) {
    // cpp:162-170 scales the TOTAL force on the hydrogen group.  The library already scaled the
    // pair part it owns; scale whatever else LAMMPS put in atom->f (bonded terms, other fixes).
    check(cph_get_sites(cph, &lambda_host, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr),
          "cph_get_sites");
    if (nsites == 0 && !pair_on_host) {
      const double scale = fscale_mode == CPH_FSCALE_LAMBDA ? lambda_host : 1.0 - lambda_host;
      int *mask = atom->mask;
      for (int i = 0; i < nlocal; i++)
        if (mask[i] & groupHbit) {
          f[i][0] *= scale;
          f[i][1] *= scale;
          f[i][2] *= scale;
        }
    }
  }
  if (!pair_on_host) {
    // `pair_modify compute no`: the GPU pair pass IS the pair computation; add its forces
    for (int i = 0; i < nlocal; i++) {
      f[i][0] += fbuf[3 * i];
      f[i][1] += fbuf[3 * i + 1];
      f[i][2] += fbuf[3 * i + 2];
    }
  } else if (dudl_mode == CPH_DUDL_REFERENCE) {
    set_force();           // LAMMPS computed the pair forces itself: reference behaviour, cpp:78
  }
}

/* ---------------------------------------------------------------------- */

void FixConstantPH::set_force()
{
  // cpp:149-171 verbatim, for the configuration in which LAMMPS (not the library) owns the forces
  double **f = atom->f;
  int *mask = atom->mask;
  int nlocal = atom->nlocal;
  const double lambda = fscale_mode == CPH_FSCALE_LAMBDA ? lambda_host : 1.0 - lambda_host;
  for (int i = 0; i < nlocal; i++) {
    if (mask[i] & groupHbit) {
      f[i][0] *= lambda;
      f[i][1] *= lambda;
      f[i][2] *= lambda;
    }
  }
}

/* ----------------------------------------------------------------------
   the reference's helpers: their arithmetic is in the library, these entry points let the
   sequence of cpp:69-73 be driven step by step (tests, debugging)
------------------------------------------------------------------------- */

void FixConstantPH::compute_Hs()
{
  // The pair part of H_atom (cpp:216-219) is produced and partitioned on the device.  What follows is
  // the rest of the reference routine (cpp:200-267, "taken from src/compute_pe_atom.cpp") for the
  // sources LAMMPS keeps on the host: bonded styles and KSpace.  Their two partition sums are handed
  // to the library, which adds them to HA/HB before the all-reduce that replaces cpp:274.
  const bool any = (force->bond && force->bond->eatom) || (force->angle && force->angle->eatom) ||
                   (force->dihedral && force->dihedral->eatom) || (force->improper && force->improper->eatom) ||
                   (force->kspace && force->kspace->compute_flag && force->kspace->eatom);
  if (!any) return;
  if (update->eflag_atom != update->ntimestep)                                         // cpp:181-183
    error->all(FLERR, "Per-atom energy was not tallied on needed timestep");

  if (atom->nmax > nmax) {                                                             // cpp:188-192
    memory->destroy(H_atom);
    nmax = atom->nmax;
    memory->create(H_atom, nmax, "constant_pH:H_atom");
  }

  int i;
  int nlocal = atom->nlocal;                                                           // cpp:200-208
  int nbond = nlocal;
  int ntotal = nlocal;
  int nkspace = nlocal;
  if (force->newton_bond) nbond += atom->nghost;
  if (force->newton) ntotal += atom->nghost;
  if (force->kspace && force->kspace->tip4pflag) nkspace += atom->nghost;

  for (i = 0; i < ntotal; i++) H_atom[i] = 0.0;                                        // cpp:212

  if (force->bond && force->bond->eatom) {                                             // cpp:221-224
    double *eatom = force->bond->eatom;
    for (i = 0; i < nbond; i++) H_atom[i] += eatom[i];
  }
  if (force->angle && force->angle->eatom) {                                           // cpp:226-229
    double *eatom = force->angle->eatom;
    for (i = 0; i < nbond; i++) H_atom[i] += eatom[i];
  }
  if (force->dihedral && force->dihedral->eatom) {                                     // cpp:231-234
    double *eatom = force->dihedral->eatom;
    for (i = 0; i < nbond; i++) H_atom[i] += eatom[i];
  }
  if (force->improper && force->improper->eatom) {                                     // cpp:236-239
    double *eatom = force->improper->eatom;
    for (i = 0; i < nbond; i++) H_atom[i] += eatom[i];
  }
  if (force->kspace && force->kspace->compute_flag && force->kspace->eatom) {          // cpp:241-244
    double *eatom = force->kspace->eatom;
    for (i = 0; i < nkspace; i++) H_atom[i] += eatom[i];
  }

  // communicate ghost energy between neighbor procs                                    cpp:251-253
  if (force->newton || (force->kspace && force->kspace->tip4pflag)) comm->reverse_comm(this);

  int *mask = atom->mask;                                                              // cpp:257-267
  double HA_local = 0.0;
  double HB_local = 0.0;
  for (i = 0; i < nlocal; i++) {
    HA_local += H_atom[i];
    if (!(mask[i] & groupHbit)) HB_local += H_atom[i];
  }
  check(cph_set_extra_partition(cph, HA_local, HB_local), "cph_set_extra_partition");
}

void FixConstantPH::calculate_df() {}                       // cpp:120-124: fused into the integrator kernel
void FixConstantPH::calculate_dU() {}                       // cpp:128-145: fused into the integrator kernel

void FixConstantPH::integrate_lambda()
{
  check(cph_integrate_lambda(cph, nevery * update->dt), "cph_integrate_lambda");       // cpp:109-117
}

void FixConstantPH::modify_water()
{
  // h:58: declared, never defined nor called in the reference (TODO at cpp:268).  With `buffer yes`
  // the library moves -(1/3) sum_s lambda_s dQ_s onto each atom of the water group whenever it
  // applies q(lambda) (cph_apply_charges inside cph_post_force), so the box charge stays constant.
  check(cph_apply_charges(cph), "cph_apply_charges");
}

/* ---------------------------------------------------------------------- */

double FixConstantPH::compute_scalar()
{
  double v = 0.0;
  check(cph_compute_scalar(cph, &v), "cph_compute_scalar");
  return v;
}

double FixConstantPH::compute_vector(int i)
{
  double v = 0.0;
  check(cph_compute_vector(cph, i, &v), "cph_compute_vector");
  return v;
}

/* ----------------------------------------------------------------------
   memory usage of local atom-based array (cpp:310-318) plus the device buffers
------------------------------------------------------------------------- */

double FixConstantPH::memory_usage()
{
  double bytes = (double) nmax * sizeof(double);
  double dev = 0.0;
  if (cph) cph_memory_usage(cph, &dev);
  return bytes + dev;
}

/* ----------------------------------------------------------------------
   restart: [version, S, (lambda|theta, v, a) * S, (xi, eta, K if tlambda)] as doubles, LAMMPS global-restart layout
------------------------------------------------------------------------- */

void FixConstantPH::write_restart(FILE *fp)
{
  int n = 0;
  check(cph_restart_size(cph, &n), "cph_restart_size");
  double *list = (double *) malloc(sizeof(double) * n);
  check(cph_pack_restart(cph, list), "cph_pack_restart");
  if (comm->me == 0) {
    int size = n * sizeof(double);
    fwrite(&size, sizeof(int), 1, fp);
    fwrite(list, sizeof(double), n, fp);
  }
  free(list);
}

void FixConstantPH::restart(char *buf)
{
  double *list = (double *) buf;
  const int S = (int) list[1];
  restart_n = 2 + 3 * S + (t_lambda_period > 0.0 ? 3 : 0);     // + thermostat state (xi, eta, K)
  free(restart_buf);
  restart_buf = (double *) malloc(sizeof(double) * restart_n);
  memcpy(restart_buf, list, sizeof(double) * restart_n);
  if (cph) {      // already initialised: apply now
    check(cph_unpack_restart(cph, restart_buf, restart_n), "cph_unpack_restart");
    free(restart_buf);
    restart_buf = nullptr;
  }
}

/* ----------------------------------------------------------------------
   cpp:287-308.  The pair energies live on the device (full neighbour list: nothing to fold); H_atom
   carries the host-side sources of compute_Hs, whose ghost shares are folded exactly as in the reference.
------------------------------------------------------------------------- */

int FixConstantPH::pack_reverse_comm(int n, int first, double *buf)
{
  int i, m, last;

  m = 0;
  last = first + n;
  for (i = first; i < last; i++) buf[m++] = H_atom[i];                                 // cpp:293
  return m;
}

/* ---------------------------------------------------------------------- */

void FixConstantPH::unpack_reverse_comm(int n, int *list, double *buf)
{
  int i, j, m;

  m = 0;
  for (i = 0; i < n; i++) {                                                            // cpp:304-307
    j = list[i];
    H_atom[j] += buf[m++];
  }
}
