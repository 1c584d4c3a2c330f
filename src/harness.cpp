// harness.cpp -- drives the drop-in FixConstantPH the way LAMMPS' Verlet loop would
// (initial_integrate -> [neighbor decide] -> force_clear -> pair -> post_force ->
// final_integrate -> output), against the LAMMPS shim under shim/lammps/.
//
//   harness BOX.bin NSTEPS [fix keywords ...]
//
// BOX.bin is written by constant_ph_b200.synth.write_harness_input(); the fix is created from
// the same argument vector a LAMMPS input line would produce:
//   fix cph all constant_pH <nevery> Hgrp Wgrp <pK> <pH> <T> [keywords]
// Output: one line per step "step H_lambda lambda_0 ... lambda_{S-1}" (full precision), then
// "FORCES_ABS_SUM <v>".  Exit code 2 when the fix aborts through error->all.
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

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: harness BOX.bin NSTEPS [nevery N] [fix keywords...]\n"); return 1; }
  FILE *fp = fopen(argv[1], "rb");
  if (!fp) { perror(argv[1]); return 1; }
  const int nsteps = atoi(argv[2]);
  auto hi = rd<int>(fp, 8);      // n, ntypes, maxspecial, style, nsites, ntitr, Hbit, Wbit
  const int n = hi[0], ntypes = hi[1], maxspecial = hi[2], style = hi[3];
  auto hd = rd<double>(fp, 22);  // boxlo3 boxhi3 cut_lj cut_coul alpha skin slj4 scoul4 pH T dt pK0
  auto x = rd<double>(fp, 3 * (size_t)n);
  auto q = rd<double>(fp, n);
  auto type = rd<int>(fp, n);
  auto tag = rd<int>(fp, n);
  auto mask = rd<int>(fp, n);
  auto mol = rd<int>(fp, n);
  auto nspecial = rd<int>(fp, 3 * (size_t)n);
  auto special = rd<int>(fp, (size_t)n * maxspecial);
  auto eps = rd<double>(fp, (size_t)(ntypes + 1) * (ntypes + 1));
  auto sig = rd<double>(fp, (size_t)(ntypes + 1) * (ntypes + 1));
  fclose(fp);

  LAMMPS lmp;
  Atom *atom = lmp.atom;
  atom->nlocal = n; atom->nmax = n; atom->natoms = n; atom->ntypes = ntypes; atom->maxspecial = maxspecial;
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
  lmp.group->add("Hgrp", hi[6]);
  lmp.group->add("Wgrp", hi[7]);
  for (int k = 0; k < 3; k++) {
    lmp.domain->boxlo[k] = lmp.domain->sublo[k] = hd[k];
    lmp.domain->boxhi[k] = lmp.domain->subhi[k] = hd[3 + k];
  }
  lmp.neighbor->skin = hd[9];
  for (int k = 0; k < 4; k++) { lmp.force->special_lj[k] = hd[10 + k]; lmp.force->special_coul[k] = hd[14 + k]; }
  lmp.update->dt = hd[20];
  Pair pair;
  pair.style = style == 1 ? "lj/cut/coul/dsf" : "lj/cut/coul/cut";
  pair.compute_flag = 0;   // pair_modify compute no: the fix's GPU pass is the pair computation
  std::vector<double *> erow(ntypes + 1), srow(ntypes + 1);
  for (int i = 0; i <= ntypes; i++) { erow[i] = &eps[(size_t)i * (ntypes + 1)]; srow[i] = &sig[(size_t)i * (ntypes + 1)]; }
  pair.epsilon = erow.data(); pair.sigma = srow.data();
  pair.cut_lj_global = hd[6]; pair.cut_coul = hd[7]; pair.alpha = hd[8];
  lmp.force->pair = &pair;

  // fix cph all constant_pH nevery Hgrp Wgrp pK pH T [keywords]
  std::vector<std::string> a = {"cph", "all", "constant_pH", "1", "Hgrp", "Wgrp",
                                std::to_string(hd[21]), std::to_string(hd[18]), std::to_string(hd[19])};
  int first_kw = 3;
  if (argc > 4 && !strcmp(argv[3], "nevery")) { a[3] = argv[4]; first_kw = 5; }
  // "bonded SCALE": a synthetic bond style whose per-atom energy is SCALE*(1 + i % 7), to exercise the
  // host-side sources of compute_Hs (cpp:221-253)
  Bond bond;
  std::vector<double> bond_eatom;
  if (argc > first_kw + 1 && !strcmp(argv[first_kw], "bonded")) {
    const double sc = atof(argv[first_kw + 1]);
    bond_eatom.resize(n);
    for (int i = 0; i < n; i++) bond_eatom[i] = sc * (1 + i % 7);
    bond.eatom = bond_eatom.data();
    lmp.force->bond = &bond;
    first_kw += 2;
  }
  for (int k = first_kw; k < argc; k++) a.push_back(argv[k]);
  std::vector<char *> av;
  for (auto &s : a) av.push_back((char *)s.c_str());

  try {
    FixConstantPH fix(&lmp, (int)av.size(), av.data());
    const int mask_bits = fix.setmask();
    fix.init();
    lmp.update->ntimestep = 0;
    lmp.update->eflag_atom = 0;
    fix.setup(0);
    const int S = fix.size_vector / 4;
    auto report = [&](long step) {
      printf("%ld %.17g", step, fix.compute_scalar());
      for (int s = 0; s < S; s++) printf(" %.17g", fix.compute_vector(4 * s));
      printf("\n");
    };
    report(0);
    for (int step = 1; step <= nsteps; step++) {
      lmp.update->ntimestep = step;
      lmp.update->eflag_atom = step;
      if (mask_bits & FixConst::INITIAL_INTEGRATE) fix.initial_integrate(0);
      std::fill(f.begin(), f.end(), 0.0);                 // force_clear(); pair->compute is off
      fix.post_force(0);
      if (mask_bits & FixConst::FINAL_INTEGRATE) fix.final_integrate();
      report(step);
    }
    double fs = 0;
    for (double v : f) fs += v < 0 ? -v : v;
    printf("FORCES_ABS_SUM %.17g\n", fs);
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
    printf("RESTART_BYTES %d\n", size);
    printf("MEMORY_USAGE %.0f\n", fix.memory_usage());
  } catch (const LammpsAbort &e) {
    fprintf(stderr, "%s\n", e.what());
    return 2;
  }
  return 0;
}
