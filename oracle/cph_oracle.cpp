// cph_oracle.cpp -- CPU ORACLE for the per-timestep hot path of LAMMPS `fix constant_pH`.
//
// THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
// (constant_ph_b200/csrc/, libcph_b200.so) never links, loads or calls this file.
//
// PARITY UNPINNED for everything beyond the closed-form bias/integrator formulae: the
// reference (MahdiTavakol/Constant_pH, fix_constant_pH.{h,cpp}, 383 lines) does not
// compile, has no tests and no golden vectors, and upstream LAMMPS is
// not available here.  What this file restates, with the reference line it follows
// (cpp:N = /root/reference/fix_constant_pH.cpp line N):
//   - fix arguments and gating                     cpp:36-49, cpp:69, cpp:75-78
//   - Donnini-2016 bias constants                  cpp:86-96   (set by the caller through orc_set_bias)
//   - lambda integrator                            cpp:109-117
//   - f(lambda), df                                cpp:120-124 (df exact per SURVEY D13, or as written)
//   - U1..U5 and dU                                cpp:128-145 (exact derivatives / erf per D14-D16, or as written)
//   - force rescale of the hydrogen group          cpp:149-171
//   - HA / HB partition of per-atom energy         cpp:264-267, allreduce cpp:274 (single rank: identity)
// and, because north_star moves it inside the path, LAMMPS's pair arithmetic for
// lj/cut/coul/cut and lj/cut/coul/dsf as written down in SURVEY.md Appendix A
// (half neighbour list, Newton's third law, ev_tally's half-half per-atom energy,
// polynomial erfc), plus the charge-derivative dU/dlambda of Appendix B.
//
// Design choices that make this an independent check of the CUDA path: half list with
// i<j (the CUDA path uses a full list), minimum-image shifts stored per list entry (the
// CUDA path uses explicit ghost atoms), per-thread accumulation arrays (the CUDA path
// reduces per warp), long-double totals.
//
// Single rank only: sublo/subhi/procgrid are ignored; every periodic box length must be
// >= 2*(cutoff+skin).
//
// Build: see oracle/Makefile  (g++ -O3 -march=native -fopenmp -shared -fPIC).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr double EWALD_P = 0.3275911;
constexpr double A1 = 0.254829592, A2 = -0.284496736, A3 = 1.421413741, A4 = -1.453152027,
                 A5 = 1.061405429;
constexpr double MY_PIS = 1.77245385090551602729;  // sqrt(pi)
constexpr double MY_PI = 3.14159265358979323846;
constexpr int SBSHIFT = 30;
constexpr int NEIGHMASK = 0x1FFFFFFF;

struct Oracle {
  std::string err;
  // units
  double qqrd2e = 332.06371, boltz = 0.0019872067, ftm2v = 1.0;
  // pair
  int style = 0, ntypes = 0;
  std::vector<double> lj1, lj2, lj3, lj4, cut_ljsq;
  double cut_lj_max = 0, cut_coul = 0, alpha = 0;
  double special_lj[4] = {1, 0, 0, 0}, special_coul[4] = {1, 0, 0, 0};
  double e_shift = 0, f_shift = 0;
  bool have_pair = false;
  bool drop_excluded = false;   // orc_set_excluded_policy
  bool force_add = false;   // orc_set_force_mode: post_force ADDS to the caller's force array
  // domain
  double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, skin = 2.0;
  int periodic[3] = {1, 1, 1};
  bool have_domain = false;
  // fix
  int nevery = 1, Hbit = 0, Wbit = 0;
  double pK0 = 0, pH = 7, T = 300;
  // bias
  double bw = 200.0, bs = 0.3, bh = 4.0, bk = 2.533, ba = 0.034041, bb = 0.005238, br = 16.458,
         bm = 0.1507, bd = 2.0, m_lambda = 20.0;
  int bias_mode = 0, dudl_mode = 1, integ_mode = 0, fscale_mode = 0;
  double nh_tau = 0, nh_xi = 0, nh_eta = 0, nh_energy = 0;   // Nose-Hoover thermostat on the site velocities
  int coord_theta = 0;            // 1: the dynamical coordinate is theta, lambda = sin^2 theta
  std::vector<double> theta;
  double extra_HA = 0, extra_HB = 0;   // host-tallied sources of compute_Hs (cpp:221-249)
  std::vector<double> extra_dudl;      // their per-site dE/dlambda (host KSpace), consumed by the next site_reduce
  int water_buffer = 0;           // modify_water(): keep the box charge constant through the 3-atom water group
  std::vector<double> qbase;      // charges as supplied by the host (the lambda = 0 state of the buffer atoms)
  // sites
  int S = 1;
  bool implicit_site = true;  // nsites==0: one global lambda over the hydrogen group (reference)
  std::vector<double> pK;
  std::vector<int> titr_tag, titr_site;
  std::vector<double> qA, qB;
  std::vector<double> lam, vlam, alam;
  // LJ end states (docs/SPEC.md "LJ end states"): B-state atom type per titration entry (0 = one LJ identity),
  // per atom after set_atoms; g[i] = dE_vdwl/dlambda carried by atom i's own end states
  std::vector<int> titr_typeB, typeB_of;
  std::vector<double> lj_g;
  bool lj_states = false;
  // per-site outputs
  std::vector<double> dudl, hdiff, flam, fs, dfs, Us, dUs;
  double HA = 0, HB = 0, evdwl = 0, ecoul = 0, Hlambda = 0, ke_sites = 0, maxdisp2 = 0;
  // atoms
  int n = 0;
  std::vector<double> x, q, xbuild;
  std::vector<int> type, tag, mask;
  std::vector<int> site_of;   // site index of every atom, -1 = none
  std::vector<int> titr_of;   // index into titr arrays, -1 = none
  std::vector<std::vector<int>> spec[3];  // partner tags by class 1-2, 1-3, 1-4
  // neighbour list (half, i<j)
  std::vector<int64_t> first;
  std::vector<int> neigh;
  std::vector<uint8_t> nimg;
  int64_t nspecial_pairs = 0, nbuilds = 0;
  int maxneigh_full = 0;
  // results
  std::vector<double> f, eatom, phi;
  bool have_atoms = false, have_pass = false;
  // f2: bonded terms (bond_style harmonic, angle_style harmonic) whose eatom cpp:221-229 adds to the partition;
  // per-atom incident lists in LAMMPS' newton_bond-off layout (every bond with both atoms, every angle with all three)
  std::vector<double> bond_k, bond_r0, angle_k, angle_t0;      // indexed by type, 1-based
  int maxbond = 0, maxangle = 0;
  std::vector<int> num_bond, bond_type, bond_atom, num_angle, angle_type, angle_a1, angle_a2, angle_a3;
  std::vector<int> index_of_tag;
  bool have_topology = false;
  double e_bond = 0, e_angle = 0;
  // f2: plain velocity-Verlet of the atoms (fix nve) so that boxes can run real dynamics
  std::vector<double> mass, v;
  bool md = false;
  // f4: kspace_style ewald (reciprocal part of the Ewald sum; pair style 2 = lj/cut/coul/long is its real-space part)
  bool kspace = false;
  double g_ewald = 0;
  int kmax[3] = {0, 0, 0};
  std::vector<double> kvec, ug;   // half-space wave vectors (3 per entry) and 4 pi/V exp(-k^2/4g^2)/k^2
  double e_kspace = 0;
};

int fail(Oracle *o, int code, const char *msg) {
  if (o) o->err = msg;
  return code;
}

inline double cself(const Oracle *o) {
  // e_self = -(e_shift/2 + alpha/sqrt(pi)) * qi^2 * qqrd2e   (dsf only)
  return o->style == 1 ? -(o->e_shift / 2.0 + o->alpha / MY_PIS) * o->qqrd2e : 0.0;
}

void build_list(Oracle *o) {
  const int n = o->n;
  if (o->md)   // self-propelled atoms: remap into the periodic box when re-neighbouring, as LAMMPS does
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++)
        if (o->periodic[k]) {
          const double Lk = o->hi[k] - o->lo[k];
          o->x[3 * i + k] -= std::floor((o->x[3 * i + k] - o->lo[k]) / Lk) * Lk;
        }
  const double cutmax = std::max(o->cut_lj_max, o->cut_coul);
  const double rlist = cutmax + o->skin;
  const double rlist2 = rlist * rlist;
  double L[3];
  int nb[3];
  for (int k = 0; k < 3; k++) {
    L[k] = o->hi[k] - o->lo[k];
    nb[k] = std::max(1, (int)std::floor(L[k] / (0.5 * rlist)));
  }
  double lob[3], bw[3];
  for (int k = 0; k < 3; k++) {
    lob[k] = o->lo[k];
    if (!o->periodic[k]) {  // non-periodic: bins must cover wherever the atoms are
      double mn = 1e300, mx = -1e300;
      for (int i = 0; i < n; i++) {
        mn = std::min(mn, o->x[3 * i + k]);
        mx = std::max(mx, o->x[3 * i + k]);
      }
      if (n == 0) { mn = o->lo[k]; mx = o->hi[k]; }
      lob[k] = std::min(mn, o->lo[k]);
      double top = std::max(mx, o->hi[k]);
      L[k] = top - lob[k] + 1e-9;
      nb[k] = std::max(1, (int)std::floor(L[k] / (0.5 * rlist)));
    }
    bw[k] = L[k] / nb[k];
  }
  const double Lbox[3] = {o->hi[0] - o->lo[0], o->hi[1] - o->lo[1], o->hi[2] - o->lo[2]};
  const int64_t nbins = (int64_t)nb[0] * nb[1] * nb[2];
  std::vector<int> binof(n), binstart(nbins + 1, 0), binatoms(n);
  for (int i = 0; i < n; i++) {
    int c[3];
    for (int k = 0; k < 3; k++) {
      int b = (int)std::floor((o->x[3 * i + k] - lob[k]) / bw[k]);
      if (o->periodic[k]) { b %= nb[k]; if (b < 0) b += nb[k]; }
      else b = std::min(std::max(b, 0), nb[k] - 1);
      c[k] = b;
    }
    binof[i] = (c[2] * nb[1] + c[1]) * nb[0] + c[0];
    binstart[binof[i] + 1]++;
  }
  for (int64_t b = 0; b < nbins; b++) binstart[b + 1] += binstart[b];
  {
    std::vector<int> fill(binstart.begin(), binstart.end() - 1);
    for (int i = 0; i < n; i++) binatoms[fill[binof[i]]++] = i;
  }
  // stencil extents: enough bins to cover rlist
  int sx[3];
  for (int k = 0; k < 3; k++) {
    sx[k] = (int)std::ceil(rlist / bw[k]);
  }
  // coul/dsf keeps fully excluded pairs for the damped-term correction (Appendix A) unless told to drop them
  const bool keep_all_special = (o->style >= 1) && !o->drop_excluded;
  o->first.assign(n + 1, 0);
  std::vector<int> cnt(n, 0);
  // for each i visit distinct bins within the stencil (periodic wrap may alias bins when nb is small)
  auto visit = [&](int i, auto &&emit) {
    int c0 = binof[i] % nb[0], c1 = (binof[i] / nb[0]) % nb[1], c2 = binof[i] / (nb[0] * nb[1]);
    int cc[3] = {c0, c1, c2};
    std::vector<int> list[3];
    for (int k = 0; k < 3; k++) {
      if (o->periodic[k] && 2 * sx[k] + 1 >= nb[k]) {
        for (int b = 0; b < nb[k]; b++) list[k].push_back(b);
      } else {
        for (int d = -sx[k]; d <= sx[k]; d++) {
          int b = cc[k] + d;
          if (o->periodic[k]) { b %= nb[k]; if (b < 0) b += nb[k]; }
          else if (b < 0 || b >= nb[k]) continue;
          list[k].push_back(b);
        }
      }
    }
    const double xi = o->x[3 * i], yi = o->x[3 * i + 1], zi = o->x[3 * i + 2];
    for (int bz : list[2]) for (int by : list[1]) for (int bx : list[0]) {
      int b = (bz * nb[1] + by) * nb[0] + bx;
      for (int p = binstart[b]; p < binstart[b + 1]; p++) {
        int j = binatoms[p];
        if (j <= i) continue;
        double d[3] = {xi - o->x[3 * j], yi - o->x[3 * j + 1], zi - o->x[3 * j + 2]};
        int img[3] = {0, 0, 0};
        double xj[3];
        for (int k = 0; k < 3; k++) {
          if (o->periodic[k]) img[k] = (int)std::nearbyint(d[k] / Lbox[k]);
          xj[k] = o->x[3 * j + k] + img[k] * Lbox[k];  // the ghost image of j that LAMMPS would hold
        }
        double dx = xi - xj[0], dy = yi - xj[1], dz = zi - xj[2];
        double rsq = dx * dx + dy * dy + dz * dz;
        if (rsq >= rlist2) continue;
        // special-bond class by tag (LAMMPS find_special)
        int sb = 0;
        int tj = o->tag[j];
        for (int c = 0; c < 3 && !sb; c++)
          for (int t : o->spec[c][i]) if (t == tj) { sb = c + 1; break; }
        if (sb && !keep_all_special && o->special_lj[sb] == 0.0 && o->special_coul[sb] == 0.0) continue;
        int code = (img[0] + 1) + 3 * (img[1] + 1) + 9 * (img[2] + 1);
        emit(j, sb, code);
      }
    }
  };
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; i++) {
    int c = 0;
    visit(i, [&](int, int, int) { c++; });
    cnt[i] = c;
  }
  for (int i = 0; i < n; i++) o->first[i + 1] = o->first[i] + cnt[i];
  o->neigh.resize(o->first[n]);
  o->nimg.resize(o->first[n]);
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; i++) {
    int64_t p = o->first[i];
    visit(i, [&](int j, int sb, int code) {
      o->neigh[p] = j | (sb << SBSHIFT);
      o->nimg[p] = (uint8_t)code;
      p++;
    });
  }
  int64_t nsp = 0;
  for (int64_t p = 0; p < o->first[n]; p++) if ((o->neigh[p] >> SBSHIFT) & 3) nsp++;
  o->nspecial_pairs = nsp;
  // full-list row lengths (for the counts the CUDA path reports)
  std::vector<int> full(n, 0);
  for (int i = 0; i < n; i++) {
    full[i] += cnt[i];
    for (int64_t p = o->first[i]; p < o->first[i + 1]; p++) full[o->neigh[p] & NEIGHMASK]++;
  }
  o->maxneigh_full = n ? *std::max_element(full.begin(), full.end()) : 0;
  o->xbuild = o->x;
  o->nbuilds++;
}

// One pass over the half list: forces, per-atom energy (half to each end, ev_tally), phi.
void pair_pass(Oracle *o, int eflag) {
  const int n = o->n;
  const int nt1 = o->ntypes + 1;
  const double cut_coulsq = o->cut_coul * o->cut_coul;
  const double L[3] = {o->hi[0] - o->lo[0], o->hi[1] - o->lo[1], o->hi[2] - o->lo[2]};
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  // Threading as in LAMMPS' OPENMP package: each thread owns a contiguous chunk of i and a
  // private accumulation array, reduced afterwards.  The private array only spans the
  // window of atoms the chunk can touch (its own i range plus the j it references), so the
  // reduction cost does not grow with the thread count when atoms are in spatial order.
  std::vector<std::vector<double>> acc(nthreads);
  std::vector<int> wlo(nthreads, 0), whi(nthreads, 0), c0(nthreads, 0), c1(nthreads, 0);
  std::vector<long double> tv(nthreads, 0.0L), tc(nthreads, 0.0L);
  std::vector<std::vector<std::pair<int, double>>> glist(nthreads);
  const bool lj_states = o->lj_states && !o->typeB_of.empty();
  const double qqrd2e = o->qqrd2e, alpha = o->alpha, e_shift = o->e_shift, f_shift = o->f_shift;
  const int style = o->style;
  {
    // chunks balanced on the number of stored pairs
    const int64_t total = o->first[n];
    int i = 0;
    for (int t = 0; t < nthreads; t++) {
      c0[t] = i;
      int64_t target = total * (t + 1) / nthreads;
      while (i < n && (o->first[i + 1] <= target || t == nthreads - 1)) i++;
      c1[t] = i;
    }
    c1[nthreads - 1] = n;
  }
#pragma omp parallel num_threads(nthreads)
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    int lo = c0[tid], hi = c1[tid];
    for (int64_t p = o->first[c0[tid]]; p < o->first[c1[tid]]; p++) {
      int j = o->neigh[p] & NEIGHMASK;
      lo = std::min(lo, j);
      hi = std::max(hi, j + 1);
    }
    wlo[tid] = lo; whi[tid] = hi;
    std::vector<double> &abuf = acc[tid];
    abuf.assign((size_t)std::max(0, hi - lo) * 5, 0.0);
    double *a = abuf.data() - (size_t)lo * 5;   // indexable by global atom index
    long double ev_t = 0.0L, ec_t = 0.0L;
    std::vector<std::pair<int, double>> &gl = glist[tid];
    gl.clear();
    for (int i = c0[tid]; i < c1[tid]; i++) {
      const double xi = o->x[3 * i], yi = o->x[3 * i + 1], zi = o->x[3 * i + 2];
      const double qi = o->q[i];
      const int ti = o->type[i];
      double fx = 0, fy = 0, fz = 0, ei = 0, phii = 0;
      if (eflag && style == 1) {
        double e_self = -(e_shift / 2.0 + alpha / MY_PIS) * qi * qi * qqrd2e;
        ei += e_self;
        ec_t += e_self;
      }
      for (int64_t p = o->first[i]; p < o->first[i + 1]; p++) {
        int jraw = o->neigh[p];
        int sb = (jraw >> SBSHIFT) & 3;
        int j = jraw & NEIGHMASK;
        int code = o->nimg[p];
        int ix = code % 3 - 1, iy = (code / 3) % 3 - 1, iz = code / 9 - 1;
        double xj = o->x[3 * j] + ix * L[0], yj = o->x[3 * j + 1] + iy * L[1], zj = o->x[3 * j + 2] + iz * L[2];
        double delx = xi - xj, dely = yi - yj, delz = zi - zj;
        double rsq = delx * delx + dely * dely + delz * delz;
        const int tj = o->type[j];
        const int tt = ti * nt1 + tj;
        const double cutljsq = o->cut_ljsq[tt];
        double cutsq = std::max(cutljsq, cut_coulsq);
        if (lj_states && (o->typeB_of[i] | o->typeB_of[j])) cutsq = std::max(cutsq, o->cut_lj_max * o->cut_lj_max);
        if (rsq >= cutsq) continue;
        const double factor_lj = o->special_lj[sb], factor_coul = o->special_coul[sb];
        const double qj = o->q[j];
        double r2inv = 1.0 / rsq;
        double forcecoul = 0.0, forcelj = 0.0, ecoul = 0.0, evdwl = 0.0, kij = 0.0;
        const int tBi = lj_states ? o->typeB_of[i] : 0, tBj = lj_states ? o->typeB_of[j] : 0;
        if (tBi | tBj) {
          // an atom with LJ end states is the lambda-weighted superposition of its two types:
          // E = sum_ab w_i^a w_j^b E_LJ(t_i^a, t_j^b),  w^A = 1 - lambda, w^B = lambda  (1, 0 for ordinary atoms)
          const double li = tBi ? o->lam[o->site_of[i]] : 0.0, lj = tBj ? o->lam[o->site_of[j]] : 0.0;
          const double wi[2] = {1.0 - li, li}, wj[2] = {1.0 - lj, lj};
          const int tis[2] = {ti, tBi}, tjs[2] = {tj, tBj};
          const double r6inv = r2inv * r2inv * r2inv;
          double dEi = 0.0, dEj = 0.0;
          for (int a = 0; a <= (tBi ? 1 : 0); a++)
            for (int b = 0; b <= (tBj ? 1 : 0); b++) {
              const int t2 = tis[a] * nt1 + tjs[b];
              if (rsq >= o->cut_ljsq[t2]) continue;
              const double e_ab = r6inv * (o->lj3[t2] * r6inv - o->lj4[t2]);
              forcelj += wi[a] * wj[b] * (r6inv * (o->lj1[t2] * r6inv - o->lj2[t2]));
              evdwl += wi[a] * wj[b] * e_ab;
              if (tBi) dEi += (a ? wj[b] : -wj[b]) * e_ab;
              if (tBj) dEj += (b ? wi[a] : -wi[a]) * e_ab;
            }
          evdwl *= factor_lj;
          if (eflag) {
            if (tBi) gl.push_back({i, factor_lj * dEi});
            if (tBj) gl.push_back({j, factor_lj * dEj});
          }
        } else if (rsq < cutljsq) {
          double r6inv = r2inv * r2inv * r2inv;
          forcelj = r6inv * (o->lj1[tt] * r6inv - o->lj2[tt]);
          evdwl = factor_lj * (r6inv * (o->lj3[tt] * r6inv - o->lj4[tt]));
        }
        if (style == 0) {
          if (rsq < cut_coulsq) {
            double rinv = std::sqrt(r2inv);
            kij = qqrd2e * factor_coul * rinv;  // E_ij = qi qj kij
            forcecoul = factor_coul * qqrd2e * qi * qj * rinv;
            ecoul = factor_coul * qqrd2e * qi * qj * rinv;
          }
        } else {
          if (rsq < cut_coulsq) {
            double r = std::sqrt(rsq);
            double prefactor = qqrd2e * qi * qj / r;
            double erfcd = std::exp(-alpha * alpha * rsq);
            double t = 1.0 / (1.0 + EWALD_P * alpha * r);
            double erfcc = t * (A1 + t * (A2 + t * (A3 + t * (A4 + t * A5)))) * erfcd;
            forcecoul = prefactor * (erfcc / r + 2.0 * alpha / MY_PIS * erfcd + r * f_shift) * r;
            double kk = erfcc - r * e_shift - rsq * f_shift;
            ecoul = prefactor * kk;
            if (factor_coul < 1.0) {
              forcecoul -= (1.0 - factor_coul) * prefactor;
              ecoul -= (1.0 - factor_coul) * prefactor;
              kk -= (1.0 - factor_coul);
            }
            kij = qqrd2e / r * kk;
          }
        }
        double fpair = (forcecoul + factor_lj * forcelj) * r2inv;
        fx += delx * fpair; fy += dely * fpair; fz += delz * fpair;
        a[5 * (size_t)j + 0] -= delx * fpair;
        a[5 * (size_t)j + 1] -= dely * fpair;
        a[5 * (size_t)j + 2] -= delz * fpair;
        if (eflag) {
          double e = evdwl + ecoul;
          ei += 0.5 * e;
          a[5 * (size_t)j + 3] += 0.5 * e;
          phii += qj * kij;
          a[5 * (size_t)j + 4] += qi * kij;
          ev_t += evdwl;
          ec_t += ecoul;
        }
      }
      a[5 * (size_t)i + 0] += fx; a[5 * (size_t)i + 1] += fy; a[5 * (size_t)i + 2] += fz;
      a[5 * (size_t)i + 3] += ei; a[5 * (size_t)i + 4] += phii;
    }
    tv[tid] = ev_t;
    tc[tid] = ec_t;
  }
  o->f.assign((size_t)3 * n, 0.0);
  if (eflag) { o->eatom.assign(n, 0.0); o->phi.assign(n, 0.0); }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) {
    double s[5] = {0, 0, 0, 0, 0};
    for (int t = 0; t < nthreads; t++)
      if (i >= wlo[t] && i < whi[t])
        for (int c = 0; c < 5; c++) s[c] += acc[t][5 * (size_t)(i - wlo[t]) + c];
    o->f[3 * i] = s[0]; o->f[3 * i + 1] = s[1]; o->f[3 * i + 2] = s[2];
    if (eflag) { o->eatom[i] = s[3]; o->phi[i] = s[4] + 2.0 * o->q[i] * cself(o); }
  }
  if (eflag) {
    o->lj_g.assign(lj_states ? n : 0, 0.0);
    if (lj_states)
      for (int t = 0; t < nthreads; t++)
        for (const auto &c : glist[t]) o->lj_g[c.first] += c.second;
    long double ev = 0, ec = 0;
    for (int t = 0; t < nthreads; t++) { ev += tv[t]; ec += tc[t]; }
    o->evdwl = (double)ev;
    o->ecoul = (double)ec;
  }
  o->have_pass = true;
}

// f2 -- bond_style harmonic and angle_style harmonic as upstream LAMMPS writes them
// [UPSTREAM-LAMMPS bond_harmonic.cpp / angle_harmonic.cpp, from memory]: E_bond = K (r - r0)^2,
// E_angle = K (theta - theta0)^2, each term evaluated once, forces to all its atoms (Newton),
// energy shared equally between them (ev_tally: 1/2 per bond atom, 1/3 per angle atom).
// Runs after pair_pass and ADDS to f / eatom, so the partition of cpp:264-267 sees the bonded
// energy the way cpp:221-229 adds bond->eatom and angle->eatom to H_atom.
void bonded_pass(Oracle *o, int eflag) {
  if (!o->have_topology) return;
  const int n = o->n;
  double L[3];
  for (int k = 0; k < 3; k++) L[k] = o->hi[k] - o->lo[k];
  auto delta = [&](int a, int b, double *d) {       // x_a - x_b, closest image
    for (int k = 0; k < 3; k++) {
      d[k] = o->x[3 * a + k] - o->x[3 * b + k];
      if (o->periodic[k]) d[k] -= L[k] * std::nearbyint(d[k] / L[k]);
    }
  };
  long double eb = 0, ea = 0;
  for (int i = 0; i < n; i++) {
    for (int m = 0; m < o->num_bond[i]; m++) {
      const int tj = o->bond_atom[(size_t)i * o->maxbond + m];
      if (o->tag[i] > tj) continue;                 // stored with both atoms: evaluate once
      const int j = o->index_of_tag[tj], bt = o->bond_type[(size_t)i * o->maxbond + m];
      double d[3];
      delta(i, j, d);
      const double rsq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2], r = std::sqrt(rsq);
      const double dr = r - o->bond_r0[bt], rk = o->bond_k[bt] * dr;
      const double fbond = r > 0.0 ? -2.0 * rk / r : 0.0;
      for (int k = 0; k < 3; k++) { o->f[3 * i + k] += d[k] * fbond; o->f[3 * j + k] -= d[k] * fbond; }
      if (eflag) {
        const double e = rk * dr;
        o->eatom[i] += 0.5 * e; o->eatom[j] += 0.5 * e;
        eb += e;
      }
    }
    for (int m = 0; m < o->num_angle[i]; m++) {
      const size_t am = (size_t)i * o->maxangle + m;
      if (o->angle_a2[am] != o->tag[i]) continue;   // stored with all three atoms: evaluate at the centre
      const int i1 = o->index_of_tag[o->angle_a1[am]], i2 = i, i3 = o->index_of_tag[o->angle_a3[am]];
      const int at = o->angle_type[am];
      double d1[3], d2[3];
      delta(i1, i2, d1);
      delta(i3, i2, d2);
      const double rsq1 = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2], r1 = std::sqrt(rsq1);
      const double rsq2 = d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2], r2 = std::sqrt(rsq2);
      double c = (d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2]) / (r1 * r2);
      c = std::min(1.0, std::max(-1.0, c));
      double sn = std::sqrt(1.0 - c * c);
      if (sn < 0.001) sn = 0.001;
      sn = 1.0 / sn;
      const double dtheta = std::acos(c) - o->angle_t0[at], tk = o->angle_k[at] * dtheta;
      const double a = -2.0 * tk * sn, a11 = a * c / rsq1, a12 = -a / (r1 * r2), a22 = a * c / rsq2;
      for (int k = 0; k < 3; k++) {
        const double f1 = a11 * d1[k] + a12 * d2[k], f3 = a22 * d2[k] + a12 * d1[k];
        o->f[3 * i1 + k] += f1; o->f[3 * i2 + k] -= f1 + f3; o->f[3 * i3 + k] += f3;
      }
      if (eflag) {
        const double e = tk * dtheta;
        o->eatom[i1] += e / 3.0; o->eatom[i2] += e / 3.0; o->eatom[i3] += e / 3.0;
        ea += e;
      }
    }
  }
  if (eflag) { o->e_bond = (double)eb; o->e_angle = (double)ea; }
}

// f2 -- fix nve [UPSTREAM-LAMMPS fix_nve.cpp, from memory]: dtf = dt/2 * ftm2v;
// initial: v += dtf f/m, x += dt v;  final: v += dtf f/m
void md_kick(Oracle *o, double dt, bool drift) {
  const double dtf = 0.5 * dt * o->ftm2v;
  for (int i = 0; i < o->n; i++) {
    const double dtfm = dtf / o->mass[o->type[i]];
    for (int k = 0; k < 3; k++) {
      o->v[3 * i + k] += dtfm * o->f[3 * i + k];
      if (drift) o->x[3 * i + k] += dt * o->v[3 * i + k];
    }
  }
}

// f4 -- kspace_style ewald [UPSTREAM-LAMMPS ewald.cpp, from memory; standard Ewald summation].  The reference only
// reads force->kspace->eatom (cpp:241-244); north_star's charge derivative needs the k-space potential as well, so
// the reciprocal sum joins the path.  With S(k) = sum_j q_j exp(i k.r_j) over the half space of wave vectors
// k = 2 pi (nx/Lx, ny/Ly, nz/Lz), |n_d| <= kmax_d, k^2 <= gsqmx:
//   E = qqrd2e [ sum_k ug(k) |S(k)|^2 - g/sqrt(pi) sum q_i^2 - pi (sum q_i)^2 / (2 g^2 V) ],  ug = 4 pi/V exp(-k^2/4g^2)/k^2
//   phi_i = dE/dq_i = qqrd2e [ sum_k 2 ug (cos(k.r_i) Re S + sin(k.r_i) Im S) - 2 g q_i/sqrt(pi) - pi sum q/(g^2 V) ]
//   f_i = qqrd2e q_i sum_k 2 ug k (sin(k.r_i) Re S - cos(k.r_i) Im S),   eatom_i = q_i phi_i / 2  (E is a quadratic form)
// The real-space part erfc(g r)/r is pair style 2 (lj/cut/coul/long).
void ewald_setup(Oracle *o) {
  o->kvec.clear(); o->ug.clear();
  if (!o->kspace) return;
  double L[3], unitk[3];
  for (int d = 0; d < 3; d++) { L[d] = o->hi[d] - o->lo[d]; unitk[d] = 2.0 * MY_PI / L[d]; }
  const double V = L[0] * L[1] * L[2];
  double gsqmx = 0;
  for (int d = 0; d < 3; d++) gsqmx = std::max(gsqmx, unitk[d] * unitk[d] * o->kmax[d] * o->kmax[d]);
  gsqmx *= 1.00001;
  for (int nx = 0; nx <= o->kmax[0]; nx++)
    for (int ny = -o->kmax[1]; ny <= o->kmax[1]; ny++)
      for (int nz = -o->kmax[2]; nz <= o->kmax[2]; nz++) {
        if (!(nx > 0 || (nx == 0 && ny > 0) || (nx == 0 && ny == 0 && nz > 0))) continue;
        const double kx = unitk[0] * nx, ky = unitk[1] * ny, kz = unitk[2] * nz;
        const double sqk = kx * kx + ky * ky + kz * kz;
        if (sqk > gsqmx) continue;
        o->kvec.push_back(kx); o->kvec.push_back(ky); o->kvec.push_back(kz);
        o->ug.push_back(4.0 * MY_PI / V * std::exp(-0.25 * sqk / (o->g_ewald * o->g_ewald)) / sqk);
      }
}

void kspace_pass(Oracle *o, int eflag) {
  if (!o->kspace) return;
  const int n = o->n;
  const size_t K = o->ug.size();
  std::vector<double> Sre(K), Sim(K);
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < K; k++) {
    long double sr = 0, si = 0;
    const double kx = o->kvec[3 * k], ky = o->kvec[3 * k + 1], kz = o->kvec[3 * k + 2];
    for (int j = 0; j < n; j++) {
      const double a = kx * o->x[3 * j] + ky * o->x[3 * j + 1] + kz * o->x[3 * j + 2];
      sr += o->q[j] * std::cos(a);
      si += o->q[j] * std::sin(a);
    }
    Sre[k] = (double)sr; Sim[k] = (double)si;
  }
  long double qsum = 0;
  for (int j = 0; j < n; j++) qsum += o->q[j];
  const double L[3] = {o->hi[0] - o->lo[0], o->hi[1] - o->lo[1], o->hi[2] - o->lo[2]};
  const double V = L[0] * L[1] * L[2], g = o->g_ewald;
  std::vector<double> ek(n, 0.0);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) {
    long double pot = 0, fx = 0, fy = 0, fz = 0;
    for (size_t k = 0; k < K; k++) {
      const double kx = o->kvec[3 * k], ky = o->kvec[3 * k + 1], kz = o->kvec[3 * k + 2];
      const double a = kx * o->x[3 * i] + ky * o->x[3 * i + 1] + kz * o->x[3 * i + 2];
      const double c = std::cos(a), sn = std::sin(a);
      pot += 2.0 * o->ug[k] * (c * Sre[k] + sn * Sim[k]);
      const double w = 2.0 * o->ug[k] * (sn * Sre[k] - c * Sim[k]);
      fx += w * kx; fy += w * ky; fz += w * kz;
    }
    const double qi = o->q[i];
    const double phi = o->qqrd2e * ((double)pot - 2.0 * g * qi / MY_PIS - MY_PI * (double)qsum / (g * g * V));
    o->f[3 * i] += o->qqrd2e * qi * (double)fx;
    o->f[3 * i + 1] += o->qqrd2e * qi * (double)fy;
    o->f[3 * i + 2] += o->qqrd2e * qi * (double)fz;
    if (eflag) {
      o->phi[i] += phi;
      ek[i] = 0.5 * qi * phi;
      o->eatom[i] += ek[i];
    }
  }
  if (eflag) {
    long double e = 0;
    for (int i = 0; i < n; i++) e += ek[i];
    o->e_kspace = (double)e;
    o->ecoul += (double)e;
  }
}

// compute_Hs tail (cpp:259-277) plus the per-site sums.
void site_reduce(Oracle *o) {
  long double HA = 0, HB = 0;
  for (int i = 0; i < o->n; i++) {
    HA += o->eatom[i];                                   // cpp:265
    if (!(o->mask[i] & o->Hbit)) HB += o->eatom[i];      // cpp:266
  }
  HA += o->extra_HA;
  HB += o->extra_HB;
  o->HA = (double)HA;
  o->HB = (double)HB;
  std::vector<long double> d(o->S, 0.0L), hd(o->S, 0.0L);
  for (int i = 0; i < o->n; i++) {
    int s = o->site_of[i];
    if (s < 0) continue;
    if (o->mask[i] & o->Hbit) hd[s] -= o->eatom[i];      // HB_s - HA_s = -sum_{i in H of s} eatom_i
    int t = o->titr_of[i];
    if (t >= 0) d[s] += (long double)(o->qB[t] - o->qA[t]) * o->phi[i];  // Appendix B (phi holds dE/dq_i)
    if (!o->lj_g.empty()) d[s] += o->lj_g[i];                            // LJ end states of atom i
  }
  if (o->water_buffer && o->dudl_mode == 1) {
    // modify_water (h:58; TODO at cpp:268): the buffer atoms carry -(1/nW) sum_s lambda_s dQ_s each,
    // so dE/dlambda_s gains -(dQ_s/nW) * sum_{a in W} dE/dq_a
    long double phiw = 0;
    int nw = 0;
    for (int i = 0; i < o->n; i++) if (o->mask[i] & o->Wbit) { phiw += o->phi[i]; nw++; }
    if (nw) {
      std::vector<long double> dQ(o->S, 0.0L);
      for (size_t t = 0; t < o->qA.size(); t++) dQ[o->titr_site[t]] += (long double)(o->qB[t] - o->qA[t]);
      for (int s = 0; s < o->S; s++) d[s] -= dQ[s] / nw * phiw;
    }
  }
  if (o->implicit_site) hd[0] += (long double)o->extra_HB - (long double)o->extra_HA;
  o->extra_HA = o->extra_HB = 0;
  for (size_t s = 0; s < o->extra_dudl.size() && s < (size_t)o->S; s++) d[s] += o->extra_dudl[s];
  o->extra_dudl.clear();
  for (int s = 0; s < o->S; s++) { o->dudl[s] = (double)d[s]; o->hdiff[s] = (double)hd[s]; }
}

// calculate_df (cpp:120-124) and calculate_dU (cpp:128-145) for one lambda.
void bias_terms(const Oracle *o, double lambda, double &f, double &df, double &U, double &dU) {
  const double a = o->ba, b = o->bb, s = o->bs, k = o->bk, d = o->bd, w = o->bw, r = o->br, m = o->bm;
  double ex = std::exp(-50.0 * (lambda - 0.5));
  f = 1.0 / (1.0 + ex);                                                  // cpp:122
  double U1 = -k * std::exp(-(lambda - 1 - b) * (lambda - 1 - b) / (2 * a * a));   // cpp:132
  double U2 = -k * std::exp(-(lambda + b) * (lambda + b) / (2 * a * a));           // cpp:133
  double U3 = d * std::exp(-(lambda - 0.5) * (lambda - 0.5) / (2 * s * s));        // cpp:134
  double U4, U5, dU1, dU2, dU3, dU4, dU5;
  if (o->bias_mode == 0) {
    df = 50.0 * ex * f * f;                                              // D13
    U4 = 0.5 * w * (1 - std::erf(r * (lambda + m)));                     // cpp:135 with erf (D16)
    U5 = 0.5 * w * (1 + std::erf(r * (lambda - 1 - m)));                 // cpp:136
    dU1 = -((lambda - 1 - b) / (a * a)) * U1;                            // D14
    dU2 = -((lambda + b) / (a * a)) * U2;
    dU3 = -((lambda - 0.5) / (s * s)) * U3;                              // cpp:139
    dU4 = -0.5 * w * r * 2 * std::exp(-r * r * (lambda + m) * (lambda + m)) / std::sqrt(MY_PI);   // D15
    dU5 = 0.5 * w * r * 2 * std::exp(-r * r * (lambda - 1 - m) * (lambda - 1 - m)) / std::sqrt(MY_PI);
  } else {
    df = 50.0 * ex / (f * f);                                            // cpp:123 verbatim
    U4 = 0.5 * w * (1 - (double)erff((float)(r * (lambda + m))));        // cpp:135 verbatim (fp32 erff)
    U5 = 0.5 * w * (1 + (double)erff((float)(r * (lambda - 1 - m))));    // cpp:136
    dU1 = -((lambda - 1 - b) / (2 * a * a)) * U1;                        // cpp:137
    dU2 = -((lambda + b) / (2 * a * a)) * U2;                            // cpp:138
    dU3 = -((lambda - 0.5) / (s * s)) * U3;                              // cpp:139
    dU4 = -0.5 * w * r * 2 * std::exp(-r * r * (lambda + 0.5) * (lambda + 0.5)) / std::sqrt(MY_PI);  // cpp:140
    dU5 = 0.5 * w * r * 2 * std::exp(-r * r * (lambda - 1 - m) * (lambda - 1 - m)) / std::sqrt(MY_PI);  // cpp:141
  }
  U = U1 + U2 + U3 + U4 + U5;          // cpp:143
  dU = dU1 + dU2 + dU3 + dU4 + dU5;    // cpp:144
}

// phase 0: reference kinematic step (cpp:109-117).  phase 1: VV first half (kick+drift).
// phase 2: VV force evaluation (a <- F/m).  phase 3: VV second half kick + H_lambda.
void integrate(Oracle *o, double dt, int phase) {
  const double ln10 = std::log(10.0);
  long double hsum = 0, ke = 0, eff_ref = 0;
  const bool thermo = o->nh_tau > 0 && o->integ_mode == 1 && (phase == 1 || phase == 3);
  const double SkT = o->S * o->boltz * o->T, Q = SkT * o->nh_tau * o->nh_tau;
  double nh = 1.0;
  if (thermo && phase == 1) {          // first Nose-Hoover half step (K of the previous step)
    o->nh_xi += 0.5 * dt * (2.0 * o->ke_sites - SkT) / Q;
    o->nh_eta += 0.5 * dt * o->nh_xi;
  }
  if (thermo) nh = std::exp(-0.5 * dt * o->nh_xi);
  for (int s = 0; s < o->S; s++) {
    double &cq = o->coord_theta ? o->theta[s] : o->lam[s];
    double &v = o->vlam[s], &acc = o->alam[s];
    if (phase == 1) {
      v = v * nh + 0.5 * acc * dt;
      cq += v * dt;
      if (o->coord_theta) { double sn = std::sin(cq); o->lam[s] = sn * sn; }
      continue;
    }
    if (phase == 3) v = (v + 0.5 * acc * dt) * nh;
    double lambda = cq, chain = 1.0;
    if (o->coord_theta) { double sn = std::sin(cq); lambda = sn * sn; chain = std::sin(2.0 * cq); }
    double f, df, U, dU;
    bias_terms(o, lambda, f, df, U, dU);
    const double pK = o->implicit_site ? o->pK0 : o->pK[s];
    const double dE = (o->dudl_mode == 0) ? o->hdiff[s] : o->dudl[s];
    const double ph = o->boltz * o->T * ln10 * (pK - o->pH);
    double f_lambda = -(dE + df * ph + dU);                                 // cpp:111
    double a_lambda = f_lambda * chain / o->m_lambda * o->ftm2v;            // cpp:112 (+ D9); theta: F_theta = F_lambda sin 2theta
    o->fs[s] = f; o->dfs[s] = df; o->Us[s] = U; o->dUs[s] = dU; o->flam[s] = f_lambda;
    double kin = 0.5 * o->m_lambda * v * v / o->ftm2v;
    hsum += f * ph + U + kin;                                               // cpp:114 site terms
    eff_ref += lambda * o->hdiff[s];                                        // cpp:114 lambda*(HB-HA)
    ke += kin;
    if (phase == 0) {
      cq = 0.5 * a_lambda * dt * dt + v * dt + cq;                          // cpp:115
      v = a_lambda * dt + v;                                                // cpp:116
    }
    if (o->coord_theta) { double sn = std::sin(cq); o->lam[s] = sn * sn; }
    acc = a_lambda;
  }
  if (phase == 1) return;
  // cpp:114: (1-lambda)*HA + lambda*HB = HA + lambda*(HB-HA); with several sites the
  // partition is per site.  Charge mode: the force-field energy at the current charges.
  double eff = (o->dudl_mode == 0) ? o->HA + (double)eff_ref : o->evdwl + o->ecoul;
  o->Hlambda = eff + (double)hsum;
  o->ke_sites = (double)ke;
  if (thermo && phase == 3 && dt > 0) {   // second Nose-Hoover half step
    o->nh_eta += 0.5 * dt * o->nh_xi;
    o->nh_xi += 0.5 * dt * (2.0 * o->ke_sites - SkT) / Q;
    o->nh_energy = 0.5 * Q * o->nh_xi * o->nh_xi + SkT * o->nh_eta;
  }
}

void apply_charges(Oracle *o) {
  for (int i = 0; i < o->n; i++) {
    int t = o->titr_of[i];
    if (t < 0) continue;
    double l = o->lam[o->titr_site[t]];
    o->q[i] = o->qA[t] + l * (o->qB[t] - o->qA[t]);   // q(lambda) = (1-lambda) qA + lambda qB
  }
  if (o->water_buffer) {
    // modify_water: total charge stays what it is at lambda = 0
    long double tot = 0;
    for (size_t t = 0; t < o->qA.size(); t++) tot += (long double)o->lam[o->titr_site[t]] * (o->qB[t] - o->qA[t]);
    int nw = 0;
    for (int i = 0; i < o->n; i++) nw += (o->mask[i] & o->Wbit) ? 1 : 0;
    for (int i = 0; i < o->n; i++)
      if (o->mask[i] & o->Wbit) o->q[i] = o->qbase[i] - (double)(tot / nw);
  }
}

void set_force(Oracle *o) {
  for (int i = 0; i < o->n; i++) {                       // cpp:162
    if (!(o->mask[i] & o->Hbit)) continue;               // cpp:164
    int s = o->site_of[i];
    if (s < 0) continue;
    double sc = o->fscale_mode == 0 ? o->lam[s] : 1.0 - o->lam[s];   // cpp:166-168 / D17
    o->f[3 * i] *= sc; o->f[3 * i + 1] *= sc; o->f[3 * i + 2] *= sc;
  }
}

void size_sites(Oracle *o) {
  int S = o->S;
  o->lam.resize(S, 0.5); o->vlam.resize(S, 0.0); o->alam.resize(S, 0.0); o->theta.resize(S, 0.78539816339744830962);
  o->dudl.assign(S, 0); o->hdiff.assign(S, 0); o->flam.assign(S, 0); o->fs.assign(S, 0);
  o->dfs.assign(S, 0); o->Us.assign(S, 0); o->dUs.assign(S, 0);
}

double max_disp2(Oracle *o) {
  double m = 0;
  for (int i = 0; i < o->n; i++) {
    double dx = o->x[3 * i] - o->xbuild[3 * i], dy = o->x[3 * i + 1] - o->xbuild[3 * i + 1],
           dz = o->x[3 * i + 2] - o->xbuild[3 * i + 2];
    m = std::max(m, dx * dx + dy * dy + dz * dz);
  }
  o->maxdisp2 = m;
  return m;
}

}  // namespace

#define ORC ((Oracle *)h)
extern "C" {

int orc_version(void) { return 1; }
int orc_create(int, void **out) { *out = new Oracle(); size_sites((Oracle *)*out); return 0; }
int orc_destroy(void *h) { delete ORC; return 0; }
const char *orc_last_error(void *h) { return h ? ORC->err.c_str() : ""; }
int orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int orc_set_units(void *h, double qqrd2e, double boltz, double ftm2v) {
  ORC->qqrd2e = qqrd2e; ORC->boltz = boltz; ORC->ftm2v = ftm2v; return 0;
}

int orc_set_pair(void *h, int style, int ntypes, const double *eps, const double *sig, const double *cut_lj,
                 double cut_lj_global, double cut_coul, double alpha, const double *slj, const double *scoul) {
  Oracle *o = ORC;
  if (style < 0 || style > 2 || ntypes < 1 || cut_coul <= 0) return fail(o, -1, "bad pair arguments");
  o->e_shift = o->f_shift = 0.0;   // style 2 (lj/cut/coul/long, alpha = g_ewald): the damped kernel of dsf without its shifts
  o->style = style; o->ntypes = ntypes; o->cut_coul = cut_coul; o->alpha = alpha;
  int m = (ntypes + 1) * (ntypes + 1);
  o->lj1.assign(m, 0); o->lj2.assign(m, 0); o->lj3.assign(m, 0); o->lj4.assign(m, 0); o->cut_ljsq.assign(m, 0);
  o->cut_lj_max = 0;
  for (int t = 0; t < m; t++) {
    double e = eps[t], s = sig[t];
    o->lj1[t] = 48.0 * e * std::pow(s, 12.0);
    o->lj2[t] = 24.0 * e * std::pow(s, 6.0);
    o->lj3[t] = 4.0 * e * std::pow(s, 12.0);
    o->lj4[t] = 4.0 * e * std::pow(s, 6.0);
    double c = cut_lj ? cut_lj[t] : cut_lj_global;
    o->cut_ljsq[t] = c * c;
    o->cut_lj_max = std::max(o->cut_lj_max, c);
  }
  for (int k = 0; k < 4; k++) { o->special_lj[k] = slj[k]; o->special_coul[k] = scoul[k]; }
  if (style == 1) {  // init of coul/dsf (Appendix A)
    double cut_coulsq = cut_coul * cut_coul;
    double erfcc = std::erfc(alpha * cut_coul);
    double erfcd = std::exp(-alpha * alpha * cut_coul * cut_coul);
    o->f_shift = -(erfcc / cut_coulsq + 2.0 / MY_PIS * alpha * erfcd / cut_coul);
    o->e_shift = erfcc / cut_coul - o->f_shift * cut_coul;
  }
  o->have_pair = true;
  return 0;
}

int orc_set_domain(void *h, const double *lo, const double *hi, const int *per, const double *, const double *,
                   const int *, const int *, double skin) {
  Oracle *o = ORC;
  for (int k = 0; k < 3; k++) { o->lo[k] = lo[k]; o->hi[k] = hi[k]; o->periodic[k] = per[k]; }
  o->skin = skin; o->have_domain = true;
  ewald_setup(o);   // the wave vectors follow the box
  return 0;
}

int orc_set_fix(void *h, int nevery, int Hbit, int Wbit, double pK, double pH, double T) {
  if (nevery <= 0) return fail(ORC, -1, "Illegal fix constant_pH every value");  // cpp:38 with D4
  ORC->nevery = nevery; ORC->Hbit = Hbit; ORC->Wbit = Wbit; ORC->pK0 = pK; ORC->pH = pH; ORC->T = T;
  return 0;
}

int orc_set_bias(void *h, double w, double s, double hbar, double k, double a, double b, double r, double m,
                 double d, double m_lambda, int mode) {
  Oracle *o = ORC;
  o->bw = w; o->bs = s; o->bh = hbar; o->bk = k; o->ba = a; o->bb = b; o->br = r; o->bm = m; o->bd = d;
  o->m_lambda = m_lambda; o->bias_mode = mode;
  return 0;
}

int orc_set_thermostat(void *h, double tau) { ORC->nh_tau = tau; return 0; }
int orc_set_extra_partition(void *h, double a, double b) { ORC->extra_HA = a; ORC->extra_HB = b; return 0; }
int orc_set_extra_dudl(void *h, int nsites, const double *d) {
  if (nsites != ORC->S || !d) return -1;
  ORC->extra_dudl.assign(d, d + nsites);
  return 0;
}
int orc_set_coordinate(void *h, int c) { ORC->coord_theta = c == 1; return 0; }
int orc_set_water_buffer(void *h, int enable) { ORC->water_buffer = enable ? 1 : 0; return 0; }

int orc_set_mode(void *h, int dudl, int integ, int fscale) {
  ORC->dudl_mode = dudl; ORC->integ_mode = integ; ORC->fscale_mode = fscale; return 0;
}

int orc_set_sites(void *h, int nsites, const double *pK, int ntitr, const int *ttag, const int *tsite,
                  const double *qA, const double *qB) {
  Oracle *o = ORC;
  if (nsites < 0 || ntitr < 0) return fail(o, -1, "bad site arguments");
  o->implicit_site = (nsites == 0);
  o->S = nsites == 0 ? 1 : nsites;
  o->pK.assign(pK, pK + nsites);
  o->titr_tag.assign(ttag, ttag + ntitr); o->titr_site.assign(tsite, tsite + ntitr);
  o->qA.assign(qA, qA + ntitr); o->qB.assign(qB, qB + ntitr);
  for (int t = 0; t < ntitr; t++) if (tsite[t] < 0 || tsite[t] >= o->S) return fail(o, -1, "site index out of range");
  o->lam.clear(); o->vlam.clear(); o->alam.clear(); o->theta.clear();
  o->titr_typeB.clear(); o->lj_states = false;
  size_sites(o);
  return 0;
}

// LJ end states: typeB[t] = atom type of titration entry t in state B (0: the atom has one LJ identity); the
// atom's own type is its state-A type.  Before set_atoms.
int orc_set_lj_states(void *h, int ntitr, const int *typeB) {
  Oracle *o = ORC;
  if (ntitr != (int)o->titr_tag.size()) return fail(o, -1, "set_lj_states: one entry per titratable atom of set_sites");
  if (o->have_atoms) return fail(o, -2, "set_lj_states before set_atoms");
  o->lj_states = false;
  for (int t = 0; t < ntitr; t++) {
    if (typeB[t] < 0 || typeB[t] > o->ntypes) return fail(o, -1, "set_lj_states: type out of range");
    if (typeB[t]) o->lj_states = true;
  }
  o->titr_typeB.assign(typeB, typeB + ntitr);
  return 0;
}

int orc_set_lambda(void *h, const double *l, const double *v) {
  Oracle *o = ORC;
  for (int s = 0; s < o->S; s++) {
    if (l) { o->lam[s] = l[s]; o->theta[s] = std::asin(std::sqrt(std::min(1.0, std::max(0.0, l[s])))); }
    if (v) o->vlam[s] = v[s];
  }
  return 0;
}

int orc_set_atoms(void *h, int, int n, const double *x, const double *q, const int *type, const int *tag,
                  const int *mask, const int *, const int *nspecial, const int *special, int maxspecial) {
  Oracle *o = ORC;
  if (!o->have_pair || !o->have_domain) return fail(o, -2, "set_pair/set_domain before set_atoms");
  const double rlist = std::max(o->cut_lj_max, o->cut_coul) + o->skin;
  for (int k = 0; k < 3; k++)
    if (o->periodic[k] && o->hi[k] - o->lo[k] < 2 * rlist) return fail(o, -6, "oracle needs box >= 2*(cut+skin)");
  o->n = n;
  o->x.assign(x, x + 3 * (size_t)n); o->q.assign(q, q + n); o->qbase.assign(q, q + n);
  o->type.assign(type, type + n); o->tag.assign(tag, tag + n); o->mask.assign(mask, mask + n);
  for (int c = 0; c < 3; c++) o->spec[c].assign(n, {});
  if (nspecial && special)
    for (int i = 0; i < n; i++) {
      int e0 = 0;
      for (int c = 0; c < 3; c++) {
        int e1 = nspecial[3 * i + c];
        for (int k = e0; k < e1 && k < maxspecial; k++) o->spec[c][i].push_back(special[(size_t)i * maxspecial + k]);
        e0 = e1;
      }
    }
  // tag -> titration entry; bit-exact bookkeeping the CUDA path must reproduce
  int maxtag = 0;
  for (int i = 0; i < n; i++) maxtag = std::max(maxtag, tag[i]);
  std::vector<int> t_of_tag(maxtag + 1, -1);
  for (size_t t = 0; t < o->titr_tag.size(); t++)
    if (o->titr_tag[t] >= 0 && o->titr_tag[t] <= maxtag) t_of_tag[o->titr_tag[t]] = (int)t;
  o->site_of.assign(n, -1); o->titr_of.assign(n, -1);
  for (int i = 0; i < n; i++) {
    int t = t_of_tag[tag[i]];
    if (t >= 0) { o->titr_of[i] = t; o->site_of[i] = o->titr_site[t]; }
    else if (o->implicit_site && (mask[i] & o->Hbit)) o->site_of[i] = 0;
  }
  o->typeB_of.assign(o->lj_states ? n : 0, 0);
  if (o->lj_states)
    for (int i = 0; i < n; i++)
      if (o->titr_of[i] >= 0) o->typeB_of[i] = o->titr_typeB[o->titr_of[i]];
  o->f.assign(3 * (size_t)n, 0); o->eatom.assign(n, 0); o->phi.assign(n, 0);
  o->have_topology = false;   // per-atom lists follow the atom order: resend after every set_atoms
  o->md = false;
  build_list(o);
  o->have_atoms = true;
  if (o->dudl_mode == 1 && !o->qA.empty()) apply_charges(o);   // charges follow lambda from the first pass on
  return 0;
}

int orc_set_x(void *h, int, const double *x) {
  if (!ORC->have_atoms) return fail(ORC, -2, "set_atoms first");
  ORC->x.assign(x, x + 3 * (size_t)ORC->n); return 0;
}
int orc_check_rebuild(void *h, int *flag) {
  double m = max_disp2(ORC);
  *flag = m > 0.25 * ORC->skin * ORC->skin;
  return 0;
}
int orc_rebuild(void *h) { build_list(ORC); return 0; }
int orc_forward(void *) { return 0; }
int orc_pair_pass(void *h, int eflag) {
  if (!ORC->have_atoms) return fail(ORC, -2, "set_atoms first");
  pair_pass(ORC, eflag); bonded_pass(ORC, eflag); kspace_pass(ORC, eflag); return 0;
}
int orc_site_reduce(void *h) {
  if (!ORC->have_pass) return fail(ORC, -2, "pair pass first");
  site_reduce(ORC); return 0;
}
int orc_integrate_lambda(void *h, double dt) { integrate(ORC, dt, ORC->integ_mode == 0 ? 0 : 2); return 0; }
int orc_initial_integrate(void *h, double dt) {
  if (ORC->integ_mode != 1) return 0;
  integrate(ORC, dt, 1);
  if (ORC->dudl_mode == 1) apply_charges(ORC);
  return 0;
}
int orc_final_integrate(void *h, double dt) {
  if (ORC->integ_mode != 1) return 0;
  integrate(ORC, dt, 3); return 0;
}
int orc_apply_charges(void *h) { apply_charges(ORC); return 0; }
int orc_set_force(void *h) { set_force(ORC); return 0; }

// post_force (cpp:67-79) in one call; same sequence as cph_post_force.
// advance == false is setup() (h:35, declared without a body): everything is evaluated, lambda does not move
static int post_force_impl(Oracle *o, int64_t ntimestep, double dt, const double *x, double *f, bool advance) {
  if (!o->have_atoms) return fail(o, -2, "set_atoms first");
  if (x) o->x.assign(x, x + 3 * (size_t)o->n);
  if (max_disp2(o) > 0.25 * o->skin * o->skin) build_list(o);
  bool active = !advance || (ntimestep % o->nevery) == 0;   // cpp:69
  pair_pass(o, active ? 1 : 0);
  bonded_pass(o, active ? 1 : 0);               // cpp:221-229: bonded eatom joins the partition
  kspace_pass(o, active ? 1 : 0);               // cpp:241-244: so does the k-space eatom
  if (active) {
    site_reduce(o);                             // cpp:70
    const int phase = (o->integ_mode == 0 && advance) ? 0 : 2;
    integrate(o, dt * o->nevery, phase);        // cpp:71-73, t_lambda = nevery*dt (cpp:113)
    if (o->dudl_mode == 1 && phase == 0) apply_charges(o);
  }
  if (o->dudl_mode == 0) set_force(o);          // cpp:78, every step
  if (f) {
    if (o->force_add) for (size_t k = 0; k < 3 * (size_t)o->n; k++) f[k] += o->f[k];
    else std::memcpy(f, o->f.data(), sizeof(double) * 3 * (size_t)o->n);
  }
  return 0;
}
int orc_post_force(void *h, int64_t ntimestep, double dt, int, const double *x, double *f) {
  return post_force_impl(ORC, ntimestep, dt, x, f, true);
}
int orc_setup(void *h, int64_t ntimestep, int, const double *x, double *f) {
  return post_force_impl(ORC, ntimestep, 0.0, x, f, false);
}
// style 0: none, 1: ewald.  After orc_set_domain (the wave vectors follow the box).
int orc_set_kspace(void *h, int style, double g_ewald, int kxmax, int kymax, int kzmax) {
  Oracle *o = ORC;
  if (style == 0) { o->kspace = false; ewald_setup(o); return 0; }
  if (style != 1 || g_ewald <= 0 || kxmax < 1 || kymax < 1 || kzmax < 1) return fail(o, -1, "bad kspace arguments");
  if (!o->have_domain) return fail(o, -2, "set_domain first");
  if (!(o->periodic[0] && o->periodic[1] && o->periodic[2])) return fail(o, -1, "ewald needs a fully periodic box");
  o->kspace = true; o->g_ewald = g_ewald;
  o->kmax[0] = kxmax; o->kmax[1] = kymax; o->kmax[2] = kzmax;
  ewald_setup(o);
  return 0;
}
int orc_get_kspace_energy(void *h, double *e) { *e = ORC->e_kspace; return 0; }
int orc_set_excluded_policy(void *h, int drop) { ORC->drop_excluded = drop != 0; return 0; }
int orc_set_force_mode(void *h, int accumulate) { ORC->force_add = accumulate != 0; return 0; }

int orc_get_forces(void *h, int, double *f) { std::memcpy(f, ORC->f.data(), sizeof(double) * 3 * (size_t)ORC->n); return 0; }
int orc_get_eatom(void *h, int, double *e) { std::memcpy(e, ORC->eatom.data(), sizeof(double) * ORC->n); return 0; }
int orc_get_phi(void *h, int, double *p) { std::memcpy(p, ORC->phi.data(), sizeof(double) * ORC->n); return 0; }
int orc_get_q(void *h, int, double *q) { std::memcpy(q, ORC->q.data(), sizeof(double) * ORC->n); return 0; }
int orc_get_scalars(void *h, double *out) {
  Oracle *o = ORC;
  out[0] = o->HA; out[1] = o->HB; out[2] = o->evdwl; out[3] = o->ecoul; out[4] = o->Hlambda;
  out[5] = o->ke_sites; out[6] = o->maxdisp2; out[7] = o->nh_energy;
  return 0;
}
int orc_get_sites(void *h, double *lambda, double *v, double *dudl, double *hdiff, double *flam, double *f,
                  double *df, double *U, double *dU) {
  Oracle *o = ORC;
  size_t b = sizeof(double) * o->S;
  if (lambda) std::memcpy(lambda, o->lam.data(), b);
  if (v) std::memcpy(v, o->vlam.data(), b);
  if (dudl) std::memcpy(dudl, o->dudl.data(), b);
  if (hdiff) std::memcpy(hdiff, o->hdiff.data(), b);
  if (flam) std::memcpy(flam, o->flam.data(), b);
  if (f) std::memcpy(f, o->fs.data(), b);
  if (df) std::memcpy(df, o->dfs.data(), b);
  if (U) std::memcpy(U, o->Us.data(), b);
  if (dU) std::memcpy(dU, o->dUs.data(), b);
  return 0;
}
int orc_compute_scalar(void *h, double *out) { *out = ORC->Hlambda; return 0; }
int orc_compute_vector(void *h, int i, double *out) {
  Oracle *o = ORC;
  if (i < 0 || i >= 4 * o->S) return fail(o, -1, "compute_vector index out of range");
  int s = i / 4;
  switch (i % 4) {
    case 0: *out = o->lam[s]; break;
    case 1: *out = o->vlam[s]; break;
    case 2: *out = o->dudl_mode == 0 ? o->hdiff[s] : o->dudl[s]; break;
    default: *out = o->flam[s];
  }
  return 0;
}
int orc_get_counts(void *h, int64_t *out) {
  Oracle *o = ORC;
  int64_t nt = 0;
  for (int i = 0; i < o->n; i++) nt += o->titr_of[i] >= 0;
  out[0] = o->n; out[1] = 0; out[2] = 2 * o->first[o->n]; out[3] = o->maxneigh_full;
  out[4] = 2 * o->nspecial_pairs; out[5] = o->nbuilds; out[6] = nt; out[7] = o->implicit_site ? 0 : o->S;
  return 0;
}
int orc_get_site_map(void *h, int *site) { std::memcpy(site, ORC->site_of.data(), sizeof(int) * ORC->n); return 0; }

// Full-list view of the half list, as keys (tag_j<<8 | sb<<5 | image code of j relative to i).
int orc_get_neighbors(void *h, int *numneigh, int64_t *keys, int64_t cap) {
  Oracle *o = ORC;
  int n = o->n;
  std::vector<int64_t> cnt(n, 0);
  for (int i = 0; i < n; i++)
    for (int64_t p = o->first[i]; p < o->first[i + 1]; p++) { cnt[i]++; cnt[o->neigh[p] & NEIGHMASK]++; }
  for (int i = 0; i < n; i++) numneigh[i] = (int)cnt[i];
  if (!keys) return 0;
  std::vector<int64_t> off(n + 1, 0);
  for (int i = 0; i < n; i++) off[i + 1] = off[i] + cnt[i];
  if (off[n] > cap) return fail(o, -5, "keys capacity too small");
  std::vector<int64_t> fill(off.begin(), off.end() - 1);
  for (int i = 0; i < n; i++)
    for (int64_t p = o->first[i]; p < o->first[i + 1]; p++) {
      int j = o->neigh[p] & NEIGHMASK, sb = (o->neigh[p] >> SBSHIFT) & 3, code = o->nimg[p];
      int ix = code % 3 - 1, iy = (code / 3) % 3 - 1, iz = code / 9 - 1;
      int rcode = (1 - ix) + 3 * (1 - iy) + 9 * (1 - iz);   // image of i as seen from j
      keys[fill[i]++] = ((int64_t)o->tag[j] << 8) | (sb << 5) | code;
      keys[fill[j]++] = ((int64_t)o->tag[i] << 8) | (sb << 5) | rcode;
    }
  for (int i = 0; i < n; i++) std::sort(keys + off[i], keys + off[i + 1]);
  return 0;
}

int orc_restart_size(void *h, int *nd) { *nd = 2 + 3 * ORC->S + (ORC->nh_tau > 0 ? 3 : 0); return 0; }
int orc_pack_restart(void *h, double *buf) {
  Oracle *o = ORC;
  buf[0] = o->coord_theta ? 2.0 : 1.0; buf[1] = o->S;
  for (int s = 0; s < o->S; s++) { buf[2 + 3 * s] = o->coord_theta ? o->theta[s] : o->lam[s]; buf[3 + 3 * s] = o->vlam[s]; buf[4 + 3 * s] = o->alam[s]; }
  if (o->nh_tau > 0) { buf[2 + 3 * o->S] = o->nh_xi; buf[3 + 3 * o->S] = o->nh_eta; buf[4 + 3 * o->S] = o->ke_sites; }
  return 0;
}
int orc_unpack_restart(void *h, const double *buf, int nd) {
  Oracle *o = ORC;
  const int extra = o->nh_tau > 0 ? 3 : 0;
  if (nd < 2 || (int)buf[1] != o->S || nd != 2 + 3 * o->S + extra) return fail(o, -1, "restart does not match the site table");
  if (extra) { o->nh_xi = buf[2 + 3 * o->S]; o->nh_eta = buf[3 + 3 * o->S]; o->ke_sites = buf[4 + 3 * o->S]; }
  if ((buf[0] == 2.0) != (o->coord_theta != 0)) return fail(o, -1, "restart record was written with the other lambda coordinate");
  for (int s = 0; s < o->S; s++) {
    if (o->coord_theta) { o->theta[s] = buf[2 + 3 * s]; double sn = std::sin(o->theta[s]); o->lam[s] = sn * sn; }
    else { o->lam[s] = buf[2 + 3 * s]; o->theta[s] = std::asin(std::sqrt(std::min(1.0, std::max(0.0, o->lam[s])))); }
    o->vlam[s] = buf[3 + 3 * s]; o->alam[s] = buf[4 + 3 * s];
  }
  if (o->dudl_mode == 1 && o->have_atoms) apply_charges(o);
  return 0;
}

// closed-form helper for the KATs of SURVEY.md §4: bias terms at one lambda.
int orc_bias_terms(void *h, double lambda, double *out4) {
  bias_terms(ORC, lambda, out4[0], out4[1], out4[2], out4[3]);
  return 0;
}
// ---- f2: bonded terms and atom dynamics ----------------------------------------------------------
int orc_set_bonded(void *h, int nbondtypes, const double *k, const double *r0, int nangletypes, const double *ak,
                   const double *theta0) {
  Oracle *o = ORC;
  o->bond_k.assign(k, k + nbondtypes + 1); o->bond_r0.assign(r0, r0 + nbondtypes + 1);
  o->angle_k.assign(ak, ak + nangletypes + 1); o->angle_t0.assign(theta0, theta0 + nangletypes + 1);
  return 0;
}
int orc_set_topology(void *h, int n, int maxbond, const int *num_bond, const int *bond_type, const int *bond_atom,
                     int maxangle, const int *num_angle, const int *angle_type, const int *a1, const int *a2,
                     const int *a3) {
  Oracle *o = ORC;
  if (!o->have_atoms || n != o->n) return fail(o, -2, "set_topology after set_atoms, same atom count");
  o->maxbond = maxbond; o->maxangle = maxangle;
  o->num_bond.assign(num_bond, num_bond + n);
  o->bond_type.assign(bond_type, bond_type + (size_t)n * maxbond);
  o->bond_atom.assign(bond_atom, bond_atom + (size_t)n * maxbond);
  o->num_angle.assign(num_angle, num_angle + n);
  o->angle_type.assign(angle_type, angle_type + (size_t)n * maxangle);
  o->angle_a1.assign(a1, a1 + (size_t)n * maxangle);
  o->angle_a2.assign(a2, a2 + (size_t)n * maxangle);
  o->angle_a3.assign(a3, a3 + (size_t)n * maxangle);
  int maxtag = 0;
  for (int i = 0; i < n; i++) maxtag = std::max(maxtag, o->tag[i]);
  o->index_of_tag.assign(maxtag + 1, -1);
  for (int i = 0; i < n; i++) o->index_of_tag[o->tag[i]] = i;
  o->have_topology = true;
  return 0;
}
int orc_get_bonded_energy(void *h, double *out2) { out2[0] = ORC->e_bond; out2[1] = ORC->e_angle; return 0; }
int orc_set_mass(void *h, int ntypes, const double *mass) { ORC->mass.assign(mass, mass + ntypes + 1); return 0; }
int orc_set_v(void *h, int, const double *v) {
  Oracle *o = ORC;
  if (!o->have_atoms || o->mass.empty()) return fail(o, -2, "set_atoms and set_mass before set_v");
  o->v.assign(v, v + 3 * (size_t)o->n);
  o->md = true;
  return 0;
}
int orc_md_initial_integrate(void *h, double dt) {
  if (!ORC->md || !ORC->have_pass) return fail(ORC, -2, "set_v and a force pass first");
  md_kick(ORC, dt, true); return 0;
}
int orc_md_final_integrate(void *h, double dt) {
  if (!ORC->md || !ORC->have_pass) return fail(ORC, -2, "set_v and a force pass first");
  md_kick(ORC, dt, false); return 0;
}
int orc_get_x(void *h, int, double *x) { std::memcpy(x, ORC->x.data(), sizeof(double) * 3 * (size_t)ORC->n); return 0; }
int orc_get_v(void *h, int, double *v) {
  if (!ORC->md) return fail(ORC, -2, "set_v first");
  std::memcpy(v, ORC->v.data(), sizeof(double) * 3 * (size_t)ORC->n); return 0;
}
int orc_sync(void *) { return 0; }

}  // extern "C"
