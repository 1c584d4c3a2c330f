// harness.cpp -- drives the drop-in FixConstantPH the way LAMMPS' Verlet loop would
// (initial_integrate -> [neighbor decide] -> force_clear -> pair -> post_force ->
// final_integrate -> output), against the LAMMPS shim under shim/lammps/.
//
//   harness BOX.bin NSTEPS [harness options ...] [fix keywords ...]
//
// BOX.bin is written by constant_ph_b200.synth.write_harness_input(); the fix is created from
// the same argument vector a LAMMPS input line would produce:
//   fix cph all constant_pH <nevery> Hgrp Wgrp <pK> <pH> <T> [keywords]
// Harness options (in this order, each optional, before the fix keywords):
//   nevery N        the fix's nevery argument
//   bonded SCALE    a synthetic bond style whose per-atom energy is SCALE*(1 + i % 7) (compute_Hs sources, cpp:221-253)
//   ghosts K        (after `bonded`) every K-th owned atom gets one periodic-image ghost; the bond style tallies a
//                   quarter of that atom's energy on the ghost (newton_bond on) and the fix must fold it back with
//                   comm->reverse_comm(this) -> pack/unpack_reverse_comm (cpp:253, 287-308) to get the same HA/HB
//   kspace SCALE    a synthetic KSpace style: potential phi_i = SCALE*(1 + tag_i % 5) on every atom, per-atom energy
//                   e_i = q_i phi_i / 2 from the CURRENT atom->q before every force evaluation (cpp:241-244)
//   runs R          split NSTEPS into R `run` commands: init() + setup() at the start of each, as LAMMPS does
//   jiggle AMP      atoms move: x_i(t) = x0_i + AMP sin(2 pi t / T_i + phi_i + d) per dimension d, T_i = 60 + i % 80,
//                   phi_i = 0.37 i (constant_ph_b200.synth.harness_jiggle replays it for the oracle)
//   timing          print "TIMING_MS_PER_STEP <wall ms per Verlet step through the fix>" (steps after the first 5)
// Multi-rank: CPH_SHIM_RANK / CPH_SHIM_NRANKS / CPH_SHIM_DIR in the environment (one process per GPU, LOCAL_RANK picks
// the device); the box is split into NRANKS bricks along x and every process keeps the atoms of its brick.
// Output (rank 0): one line per step "step H_lambda lambda_0 ... lambda_{S-1}" (full precision); every rank:
// "FORCES_ABS_SUM <v>" over its own atoms.  Exit code 2 when the fix aborts through error->all.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fix_constant_pH.h"
#include "lammps_shim.h"

using namespace LAMMPS_NS;

template <typename T>
static std::vector<T> rd(FILE *fp, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, fp) != n) { fprintf(stderr, "short read\n"); exit(3); }
  return v;
}

template <typename T>
static std::vector<T> pick(const std::vector<T> &v, const std::vector<int> &rows, size_t width) {
  std::vector<T> out(rows.size() * width);
  for (size_t k = 0; k < rows.size(); k++)
    for (size_t c = 0; c < width; c++) out[k * width + c] = v[(size_t)rows[k] * width + c];
  return out;
}

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: harness BOX.bin NSTEPS [nevery N] [bonded S] [ghosts K] [kspace S] [runs R] [jiggle A] [timing] [fix keywords...]\n"); return 1; }
  FILE *fp = fopen(argv[1], "rb");
  if (!fp) { perror(argv[1]); return 1; }
  const int nsteps = atoi(argv[2]);
  auto hi = rd<int>(fp, 8);      // n, ntypes, maxspecial, style, nsites, ntitr, Hbit, Wbit
  const int nall = hi[0], ntypes = hi[1], maxspecial = hi[2], style = hi[3];
  auto hd = rd<double>(fp, 22);  // boxlo3 boxhi3 cut_lj cut_coul alpha skin slj4 scoul4 pH T dt pK0
  auto x_all = rd<double>(fp, 3 * (size_t)nall);
  auto q_all = rd<double>(fp, nall);
  auto type_all = rd<int>(fp, nall);
  auto tag_all = rd<int>(fp, nall);
  auto mask_all = rd<int>(fp, nall);
  auto mol_all = rd<int>(fp, nall);
  auto nspecial_all = rd<int>(fp, 3 * (size_t)nall);
  auto special_all = rd<int>(fp, (size_t)nall * maxspecial);
  auto eps = rd<double>(fp, (size_t)(ntypes + 1) * (ntypes + 1));
  auto sig = rd<double>(fp, (size_t)(ntypes + 1) * (ntypes + 1));
  fclose(fp);

  // ---- this rank's brick (along x) and its atoms ------------------------------------------------
  ShimWorld &world = ShimWorld::get();
  const int P = world.size, me = world.rank;
  const double Lx = hd[3] - hd[0];
  const double sublo_x = hd[0] + Lx * me / P, subhi_x = hd[0] + Lx * (me + 1) / P;
  std::vector<int> rows;
  for (int i = 0; i < nall; i++)
    if (P == 1 || (x_all[3 * (size_t)i] >= sublo_x && x_all[3 * (size_t)i] < subhi_x)) rows.push_back(i);
  const int n = (int)rows.size();
  auto x = pick(x_all, rows, 3);
  auto q = pick(q_all, rows, 1);
  auto type = pick(type_all, rows, 1);
  auto tag = pick(tag_all, rows, 1);
  auto mask = pick(mask_all, rows, 1);
  auto mol = pick(mol_all, rows, 1);
  auto nspecial = pick(nspecial_all, rows, 3);
  auto special = pick(special_all, rows, (size_t)maxspecial);
  const std::vector<double> x0 = x;

  LAMMPS lmp;
  Atom *atom = lmp.atom;
  atom->nlocal = n; atom->nmax = n; atom->natoms = nall; atom->ntypes = ntypes; atom->maxspecial = maxspecial;
  std::vector<double> f(3 * (size_t)n, 0.0);
  std::vector<double *> xrow(n), frow(n);
  std::vector<int *> nsrow(n), sprow(n);
  for (int i = 0; i < n; i++) {
    xrow[i] = &x[3 * (size_t)i]; frow[i] = &f[3 * (size_t)i];
    nsrow[i] = &nspecial[3 * (size_t)i]; sprow[i] = maxspecial ? &special[(size_t)i * maxspecial] : nullptr;
  }
  atom->x = xrow.data(); atom->f = frow.data(); atom->q = q.data(); atom->type = type.data();
  atom->tag = tag.data(); atom->mask = mask.data(); atom->molecule = mol.data();
  atom->nspecial = nsrow.data(); atom->special = sprow.data();
  atom->map_array.assign((size_t)nall + 2, -1);
  for (int i = 0; i < n; i++) atom->map_array[tag[i]] = i;
  lmp.group->add("Hgrp", hi[6]);
  lmp.group->add("Wgrp", hi[7]);
  lmp.group->count_override = 3;            // group->count is a global (all-rank) count upstream
  for (int k = 0; k < 3; k++) {
    lmp.domain->boxlo[k] = lmp.domain->sublo[k] = hd[k];
    lmp.domain->boxhi[k] = lmp.domain->subhi[k] = hd[3 + k];
  }
  lmp.domain->sublo[0] = sublo_x;
  lmp.domain->subhi[0] = subhi_x;
  lmp.comm->me = me; lmp.comm->nprocs = P;
  lmp.comm->procgrid[0] = P; lmp.comm->myloc[0] = me;
  lmp.neighbor->skin = hd[9];
  for (int k = 0; k < 4; k++) { lmp.force->special_lj[k] = hd[10 + k]; lmp.force->special_coul[k] = hd[14 + k]; }
  lmp.update->dt = hd[20];
  Pair pair;
  pair.style = style == 2 ? "lj/cut/coul/long" : style == 1 ? "lj/cut/coul/dsf" : "lj/cut/coul/cut";
  pair.compute_flag = 0;   // pair_modify compute no: the fix's GPU pass is the pair computation
  std::vector<double *> erow(ntypes + 1), srow(ntypes + 1);
  for (int i = 0; i <= ntypes; i++) { erow[i] = &eps[(size_t)i * (ntypes + 1)]; srow[i] = &sig[(size_t)i * (ntypes + 1)]; }
  pair.epsilon = erow.data(); pair.sigma = srow.data();
  pair.cut_lj_global = hd[6]; pair.cut_coul = hd[7]; pair.alpha = hd[8];
  lmp.force->pair = &pair;

  // fix cph all constant_pH nevery Hgrp Wgrp pK pH T [keywords]
  std::vector<std::string> a = {"cph", "all", "constant_pH", "1", "Hgrp", "Wgrp",
                                std::to_string(hd[21]), std::to_string(hd[18]), std::to_string(hd[19])};
  int kw = 3;
  auto opt = [&](const char *name) { return kw + 1 < argc && !strcmp(argv[kw], name); };
  if (opt("nevery")) { a[3] = argv[kw + 1]; kw += 2; }
  Bond bond;
  std::vector<double> bond_eatom;
  if (opt("bonded")) {
    const double sc = atof(argv[kw + 1]);
    bond_eatom.resize(n);
    for (int i = 0; i < n; i++) bond_eatom[i] = sc * (1 + rows[i] % 7);
    bond.eatom = bond_eatom.data();
    lmp.force->bond = &bond;
    kw += 2;
    if (opt("ghosts")) {
      const int every = std::max(1, atoi(argv[kw + 1]));
      for (int i = 0; i < n; i += every) lmp.comm->ghost_owner.push_back(i);
      const int ng = (int)lmp.comm->ghost_owner.size();
      lmp.comm->first_ghost = n;
      atom->nghost = ng;
      atom->nmax = n + ng;
      bond_eatom.resize(n + ng);
      for (int g = 0; g < ng; g++) {
        const int i = lmp.comm->ghost_owner[g];
        bond_eatom[n + g] = 0.25 * bond_eatom[i];
        bond_eatom[i] *= 0.75;
      }
      bond.eatom = bond_eatom.data();
      kw += 2;
    }
  }
  KSpace kspace;
  if (style == 2) {        // lj/cut/coul/long needs a KSpace style; `kspace_modify compute no`: the fix runs the sum (keyword ewald)
    kspace.g_ewald = hd[8];
    kspace.compute_flag = 0;
    lmp.force->kspace = &kspace;
  }
  std::vector<double> kspace_eatom;
  double kspace_scale = 0.0;
  if (opt("kspace")) {
    kspace_scale = atof(argv[kw + 1]);
    kspace_eatom.assign(n, 0.0);
    kspace.eatom = kspace_eatom.data();
    lmp.force->kspace = &kspace;
    kw += 2;
  }
  auto kspace_compute = [&]() {                            // KSpace::compute(eflag_atom) on the current charges
    for (int i = 0; i < (int)kspace_eatom.size(); i++) kspace_eatom[i] = 0.5 * q[i] * kspace_scale * (1 + tag[i] % 5);
  };
  int nruns = 1;
  if (opt("runs")) { nruns = std::max(1, atoi(argv[kw + 1])); kw += 2; }
  double jiggle = 0.0;
  if (opt("jiggle")) { jiggle = atof(argv[kw + 1]); kw += 2; }
  bool timing = false;
  if (kw < argc && !strcmp(argv[kw], "timing")) { timing = true; kw += 1; }
  for (int k = kw; k < argc; k++) a.push_back(argv[k]);
  std::vector<char *> av;
  for (auto &s : a) av.push_back((char *)s.c_str());

  auto move_atoms = [&](long step) {
    if (jiggle == 0.0) return;
    const double t = step * lmp.update->dt;
    for (int i = 0; i < n; i++) {
      const int g = rows[i];
      const double w = 2.0 * M_PI / (60.0 + g % 80), ph = 0.37 * g;
      for (int d = 0; d < 3; d++) x[3 * (size_t)i + d] = x0[3 * (size_t)i + d] + jiggle * sin(w * t + ph + d);
    }
  };

  try {
    FixConstantPH fix(&lmp, (int)av.size(), av.data());
    const int mask_bits = fix.setmask();
    int S = 1;
    auto report = [&](long step) {
      if (me != 0) return;
      printf("%ld %.17g", step, fix.compute_scalar());
      for (int s = 0; s < S; s++) printf(" %.17g", fix.compute_vector(4 * s));
      printf("\n");
    };
    long step = 0;
    double timed_ms = 0.0;
    long timed_steps = 0;
    for (int run = 0; run < nruns; run++) {
      // `run N`: Modify::init -> fix.init(), then Verlet::setup -> fix.setup()
      fix.init();
      S = fix.size_vector / 4;
      lmp.update->ntimestep = step;
      lmp.update->eflag_atom = step;
      std::fill(f.begin(), f.end(), 0.0);
      kspace_compute();
      fix.setup(0);
      if (run == 0) report(0);
      const long last = (long)nsteps * (run + 1) / nruns;
      while (step < last) {
        step++;
        lmp.update->ntimestep = step;
        lmp.update->eflag_atom = step;
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
          return std::chrono::duration<double, std::milli>(b - a).count();
        };
        const auto t0 = now();
        if (mask_bits & FixConst::INITIAL_INTEGRATE) fix.initial_integrate(0);
        const auto t1 = now();
        move_atoms(step);                                   // the host integrator's job
        std::fill(f.begin(), f.end(), 0.0);                 // force_clear(); pair->compute is off
        kspace_compute();
        const auto t2 = now();
        fix.post_force(0);
        if (mask_bits & FixConst::FINAL_INTEGRATE) fix.final_integrate();
        const auto t3 = now();
        if (step > 5) {    // the fix's hooks only: not the stand-in integrator, not LAMMPS' force_clear
          timed_ms += ms(t0, t1) + ms(t2, t3);
          timed_steps++;
        }
        if (!timing) report(step);
      }
    }
    if (timing) report(step);
    double fs = 0;
    for (double v : f) fs += v < 0 ? -v : v;
    printf("FORCES_ABS_SUM %.17g\n", fs);
    if (timing && timed_steps) printf("TIMING_MS_PER_STEP %.6f\n", timed_ms / timed_steps);
    // restart round trip through the LAMMPS hooks
    FILE *rf = tmpfile();
    fix.write_restart(rf);
    rewind(rf);
    int size = 0;
    if (fread(&size, sizeof(int), 1, rf) != 1) size = 0;
    std::vector<char> buf(size);
    if (size && fread(buf.data(), 1, size, rf) != (size_t)size) size = 0;
    fclose(rf);
    if (size) fix.restart(buf.data());
    if (me == 0) printf("RESTART_BYTES %d\n", size);
    printf("MEMORY_USAGE %.0f\n", fix.memory_usage());
  } catch (const LammpsAbort &e) {
    fprintf(stderr, "%s\n", e.what());
    return 2;
  }
  return 0;
}
