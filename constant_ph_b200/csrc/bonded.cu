// bonded.cu -- SURVEY.md §8 row f2: the bonded terms of flexible molecules and plain atom dynamics.
//
// The reference adds bond->eatom and angle->eatom to its per-atom energy before the HA/HB partition
// (fix_constant_pH.cpp:221-229).  Here those terms are evaluated on the device, right behind the pair
// pass, and ADDED to its forces and per-atom energies, so cph_site_reduce sees them exactly where the
// reference's partition loop (cpp:264-267) would.  Styles: bond_style harmonic, E = K (r - r0)^2, and
// angle_style harmonic, E = K (theta - theta0)^2 [upstream LAMMPS formulae, SURVEY Appendix A conventions];
// per-atom energy is shared equally between the atoms of a term (ev_tally).
//
// Layout.  The host hands over LAMMPS' per-atom incident lists (newton_bond off: every bond with both of its
// atoms, every angle with all three; partner ids are tags).  After each list build a resolve kernel turns
// the tags into indices of the current internal order: a bonded partner is a 1-2 or 1-3 special neighbour,
// so it is found among the special entries at the end of the atom's Verlet row (owned atom or the ghost
// image the row holds).  The per-step kernel is one thread per owned atom: each atom evaluates the terms it
// takes part in and keeps only ITS OWN force and energy share -- no atomics, no reverse halo, same answer
// on any decomposition.  HBM traffic per step: 32 B {x,y,z,q} + 24 B force RMW + 8 B energy RMW per atom,
// plus 4 B per stored partner index; the partners' coordinates come out of L2 (they are the atom's nearest
// neighbours in the cell-sorted order).
//
// The second half is `fix nve` [upstream LAMMPS: v += dt/2 * ftm2v * f/m; x += dt * v; ...; v += dt/2 * ftm2v * f/m]
// so that benchmark boxes can run real dynamics with positions resident in HBM.  Atoms are remapped into
// the periodic box when the list is rebuilt; in a decomposed run an atom that leaves its sub-box by more
// than the skin still has to be migrated by the host (cph_set_atoms), which cph_rebuild reports.
#include <cmath>

#include "cph_internal.h"

namespace {

constexpr int TPB = 128;
inline int nblk(int n) { return (n + TPB - 1) / TPB; }

__device__ __forceinline__ int find_partner(int want, const double4 p, const int *__restrict__ row_end, int nsp,
                                            const int *__restrict__ tag, const double4 *__restrict__ xq) {
  int best = -1;
  double bestd = 1.0e300;
  for (int s = 0; s < nsp; s++) {
    const int j = row_end[-s] & CPH_NEIGHMASK;
    if (tag[j] != want) continue;
    const double4 r = xq[j];
    const double dx = p.x - r.x, dy = p.y - r.y, dz = p.z - r.z;
    const double d = dx * dx + dy * dy + dz * dz;
    if (d < bestd) { bestd = d; best = j; }
  }
  return best;
}

// partner tags -> indices of the current internal order
__global__ void resolve_kernel(int n, const int *__restrict__ perm, const int *__restrict__ tag,
                               const double4 *__restrict__ xq, const int *__restrict__ neigh,
                               const int *__restrict__ numspec, int rowcap, int maxbond,
                               const int *__restrict__ num_bond, const int *__restrict__ bond_type,
                               const int *__restrict__ bond_atom, int maxangle, const int *__restrict__ num_angle,
                               const int *__restrict__ angle_type, const int *__restrict__ a1,
                               const int *__restrict__ a2, const int *__restrict__ a3, int *bcount, int *bond_j,
                               int *bond_t, int *angle_j, int *angle_t, unsigned int *flags) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int c = perm[k], me = tag[k];
  const double4 p = xq[k];
  const int nsp = numspec[k];
  const int *row_end = neigh + (size_t)k * rowcap + (rowcap - 1);
  const int nb = min(num_bond[c], maxbond), na = min(num_angle[c], maxangle);
  bool missing = false;
  for (int m = 0; m < nb; m++) {
    const int j = find_partner(bond_atom[(size_t)c * maxbond + m], p, row_end, nsp, tag, xq);
    missing |= j < 0;
    bond_j[(size_t)k * maxbond + m] = j;
    bond_t[(size_t)k * maxbond + m] = bond_type[(size_t)c * maxbond + m];
  }
  for (int m = 0; m < na; m++) {
    const size_t e = (size_t)c * maxangle + m;
    const int t1 = a1[e], t2 = a2[e], t3 = a3[e];
    const int role = t2 == me ? 1 : (t1 == me ? 0 : 2);
    // role 1 (centre): the two ends in stored order; role 0/2 (an end): the centre, then the far end
    const int wa = role == 1 ? t1 : t2, wb = role == 0 ? t3 : (role == 1 ? t3 : t1);
    const int ja = find_partner(wa, p, row_end, nsp, tag, xq), jb = find_partner(wb, p, row_end, nsp, tag, xq);
    missing |= ja < 0 || jb < 0 || (t1 != me && t2 != me && t3 != me);
    angle_j[2 * ((size_t)k * maxangle + m)] = ja;
    angle_j[2 * ((size_t)k * maxangle + m) + 1] = jb;
    angle_t[(size_t)k * maxangle + m] = angle_type[e] | (role << 16);
  }
  bcount[k] = nb | (na << 8);
  if (missing) atomicOr(flags + 6, 1u);
}

template <int EFLAG>
__global__ void __launch_bounds__(TPB)
bonded_kernel(int n, const double4 *__restrict__ xq, const int *__restrict__ bcount, int maxbond,
              const int *__restrict__ bond_j, const int *__restrict__ bond_t, const double2 *__restrict__ bond_coef,
              int maxangle, const int *__restrict__ angle_j, const int *__restrict__ angle_t,
              const double2 *__restrict__ angle_coef, double *__restrict__ f, double *__restrict__ eatom,
              double *__restrict__ etot) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  double eb = 0.0, ea = 0.0;
  if (k < n) {
    const int cnt = bcount[k], nb = cnt & 255, na = cnt >> 8;
    if (cnt) {
      const double4 p = xq[k];
      double fx = 0.0, fy = 0.0, fz = 0.0;
      for (int m = 0; m < nb; m++) {
        const double4 r = xq[bond_j[(size_t)k * maxbond + m]];
        const double2 c = bond_coef[bond_t[(size_t)k * maxbond + m]];
        const double dx = p.x - r.x, dy = p.y - r.y, dz = p.z - r.z;
        const double rsq = dx * dx + dy * dy + dz * dz, rr = sqrt(rsq);
        const double dr = rr - c.y, rk = c.x * dr;
        const double fbond = rr > 0.0 ? -2.0 * rk / rr : 0.0;
        fx += dx * fbond; fy += dy * fbond; fz += dz * fbond;
        if (EFLAG) eb += 0.5 * (rk * dr);
      }
      for (int m = 0; m < na; m++) {
        const size_t e = (size_t)k * maxangle + m;
        const int tr = angle_t[e], role = tr >> 16;
        const double2 c = angle_coef[tr & 0xffff];
        const double4 qa = xq[angle_j[2 * e]], qb = xq[angle_j[2 * e + 1]];
        // i1 - i2 (centre) - i3 with this atom in position `role`
        const double4 p1 = role == 0 ? p : (role == 1 ? qa : qb);
        const double4 p2 = role == 1 ? p : qa;
        const double4 p3 = role == 2 ? p : qb;
        const double d1x = p1.x - p2.x, d1y = p1.y - p2.y, d1z = p1.z - p2.z;
        const double d2x = p3.x - p2.x, d2y = p3.y - p2.y, d2z = p3.z - p2.z;
        const double rsq1 = d1x * d1x + d1y * d1y + d1z * d1z, r1 = sqrt(rsq1);
        const double rsq2 = d2x * d2x + d2y * d2y + d2z * d2z, r2 = sqrt(rsq2);
        double cs = (d1x * d2x + d1y * d2y + d1z * d2z) / (r1 * r2);
        cs = fmin(1.0, fmax(-1.0, cs));
        double sn = sqrt(1.0 - cs * cs);
        if (sn < 0.001) sn = 0.001;
        sn = 1.0 / sn;
        const double dtheta = acos(cs) - c.y, tk = c.x * dtheta;
        const double a = -2.0 * tk * sn, a11 = a * cs / rsq1, a12 = -a / (r1 * r2), a22 = a * cs / rsq2;
        const double f1x = a11 * d1x + a12 * d2x, f1y = a11 * d1y + a12 * d2y, f1z = a11 * d1z + a12 * d2z;
        const double f3x = a22 * d2x + a12 * d1x, f3y = a22 * d2y + a12 * d1y, f3z = a22 * d2z + a12 * d1z;
        if (role == 0) { fx += f1x; fy += f1y; fz += f1z; }
        else if (role == 2) { fx += f3x; fy += f3y; fz += f3z; }
        else { fx -= f1x + f3x; fy -= f1y + f3y; fz -= f1z + f3z; }
        if (EFLAG) ea += (tk * dtheta) / 3.0;
      }
      f[3 * (size_t)k] += fx;
      f[3 * (size_t)k + 1] += fy;
      f[3 * (size_t)k + 2] += fz;
      if (EFLAG) eatom[k] += eb + ea;
    }
  }
  if (EFLAG) {
    for (int o = 16; o; o >>= 1) {
      eb += __shfl_xor_sync(0xffffffffu, eb, o);
      ea += __shfl_xor_sync(0xffffffffu, ea, o);
    }
    if ((threadIdx.x & 31) == 0 && (eb != 0.0 || ea != 0.0)) {
      atomicAdd(etot, eb);
      atomicAdd(etot + 1, ea);
    }
  }
}

struct MassTable {
  double inv[CPH_MAXNT1];
};

// fix nve: v += dtf f / m  (and x += dt v in the first half)
__global__ void nve_kernel(int n, double4 *__restrict__ xq, double3 *__restrict__ v, const double *__restrict__ f,
                           const int *__restrict__ type, MassTable mt, double dtf, double dt, int drift) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double dtfm = dtf * mt.inv[type[k]];
  double3 vk = v[k];
  vk.x += dtfm * f[3 * (size_t)k];
  vk.y += dtfm * f[3 * (size_t)k + 1];
  vk.z += dtfm * f[3 * (size_t)k + 2];
  v[k] = vk;
  if (drift) {
    double4 p = xq[k];
    p.x += dt * vk.x; p.y += dt * vk.y; p.z += dt * vk.z;
    xq[k] = p;
  }
}

__global__ void wrap_kernel(int n, double4 *__restrict__ xq, double3 lo, double3 len, int3 on) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double4 p = xq[k];
  if (on.x) p.x -= floor((p.x - lo.x) / len.x) * len.x;
  if (on.y) p.y -= floor((p.y - lo.y) / len.y) * len.y;
  if (on.z) p.z -= floor((p.z - lo.z) / len.z) * len.z;
  xq[k] = p;
}

__global__ void scatter_v_kernel(int n, const int *__restrict__ perm, const double *__restrict__ vc, double3 *v) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const size_t c = (size_t)perm[k];
  v[k] = make_double3(vc[3 * c], vc[3 * c + 1], vc[3 * c + 2]);
}

template <typename T>
int put(cph_handle *h, DevBuf<T> &buf, const T *src, size_t count) {
  CPH_CUDA(h, buf.reserve(count + 1));
  if (count) CPH_CUDA(h, cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  return 0;
}

}  // namespace

int cph_bonded_set_coef(cph_handle *h, int nbondtypes, const double *bk, const double *br0, int nangletypes,
                        const double *ak, const double *at0) {
  if (nbondtypes < 0 || nangletypes < 0 || nbondtypes > 0xffff || nangletypes > 0xffff)
    return cph_fail(h, CPH_ERR_ARG, "bad bond/angle type counts %d/%d", nbondtypes, nangletypes);
  if ((nbondtypes && (!bk || !br0)) || (nangletypes && (!ak || !at0)))
    return cph_fail(h, CPH_ERR_ARG, "NULL bonded coefficient table");
  std::vector<double2> b(nbondtypes + 1, make_double2(0, 0)), a(nangletypes + 1, make_double2(0, 0));
  for (int t = 1; t <= nbondtypes; t++) b[t] = make_double2(bk[t], br0[t]);
  for (int t = 1; t <= nangletypes; t++) a[t] = make_double2(ak[t], at0[t]);
  CPH_TRY(put(h, h->d_bond_coef, b.data(), b.size()));
  CPH_TRY(put(h, h->d_angle_coef, a.data(), a.size()));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  h->nbondtypes = nbondtypes;
  h->nangletypes = nangletypes;
  h->have_bonded_coef = true;
  return 0;
}

int cph_bonded_set_topology(cph_handle *h, int nlocal, int maxbond, const int *num_bond, const int *bond_type,
                            const int *bond_atom, int maxangle, const int *num_angle, const int *angle_type,
                            const int *a1, const int *a2, const int *a3) {
  if (!h->have_atoms || nlocal != h->nlocal)
    return cph_fail(h, CPH_ERR_STATE, "cph_set_topology follows cph_set_atoms with the same atom count (%d vs %d)",
                    nlocal, h->nlocal);
  if (!h->have_bonded_coef) return cph_fail(h, CPH_ERR_STATE, "cph_set_bonded first");
  if (maxbond < 0 || maxangle < 0 || maxbond > 255 || maxangle > 255)
    return cph_fail(h, CPH_ERR_ARG, "maxbond/maxangle %d/%d outside [0,255]", maxbond, maxangle);
  if (h->maxspecial == 0 && nlocal > 0 && (maxbond || maxangle))
    return cph_fail(h, CPH_ERR_STATE, "bonded terms need the special-bond tables in cph_set_atoms (partners are resolved through them)");
  const size_t n = (size_t)nlocal;
  if (n && ((maxbond && (!num_bond || !bond_type || !bond_atom)) ||
            (maxangle && (!num_angle || !angle_type || !a1 || !a2 || !a3))))
    return cph_fail(h, CPH_ERR_ARG, "NULL topology array");
  for (size_t i = 0; i < n; i++) {
    const int nb = maxbond ? num_bond[i] : 0, na = maxangle ? num_angle[i] : 0;
    if (nb < 0 || nb > maxbond || na < 0 || na > maxangle)
      return cph_fail(h, CPH_ERR_ARG, "atom %zu: %d bonds / %d angles outside [0,%d] / [0,%d]", i, nb, na, maxbond, maxangle);
    for (int m = 0; m < nb; m++)
      if (bond_type[i * maxbond + m] < 1 || bond_type[i * maxbond + m] > h->nbondtypes)
        return cph_fail(h, CPH_ERR_ARG, "atom %zu: bond type %d outside [1,%d]", i, bond_type[i * maxbond + m], h->nbondtypes);
    for (int m = 0; m < na; m++)
      if (angle_type[i * maxangle + m] < 1 || angle_type[i * maxangle + m] > h->nangletypes)
        return cph_fail(h, CPH_ERR_ARG, "atom %zu: angle type %d outside [1,%d]", i, angle_type[i * maxangle + m], h->nangletypes);
  }
  cudaSetDevice(h->device);
  std::vector<int> zeros(n, 0);
  CPH_TRY(put(h, h->d_num_bond, maxbond ? num_bond : zeros.data(), n));
  CPH_TRY(put(h, h->d_num_angle, maxangle ? num_angle : zeros.data(), n));
  CPH_TRY(put(h, h->d_bond_type, bond_type, n * maxbond));
  CPH_TRY(put(h, h->d_bond_atom, bond_atom, n * maxbond));
  CPH_TRY(put(h, h->d_angle_type, angle_type, n * maxangle));
  CPH_TRY(put(h, h->d_angle_a1, a1, n * maxangle));
  CPH_TRY(put(h, h->d_angle_a2, a2, n * maxangle));
  CPH_TRY(put(h, h->d_angle_a3, a3, n * maxangle));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));   // the host arrays may go away after this call
  h->maxbond = maxbond;
  h->maxangle = maxangle;
  h->have_topology = true;
  h->have_pass = false;
  // fully excluded special pairs are left out of the rows under coul/cut; the partners are looked up
  // there, so such a list is rebuilt with them kept
  if (h->last_dropmask != 0) return cph_rebuild(h);
  return cph_bonded_resolve(h);
}

int cph_bonded_resolve(cph_handle *h) {
  if (!h->have_topology) return 0;
  const int n = h->nlocal;
  const size_t nn = (size_t)n;
  CPH_CUDA(h, h->d_bcount.reserve(nn + 1));
  CPH_CUDA(h, h->d_bond_j.reserve(nn * h->maxbond + 1));
  CPH_CUDA(h, h->d_bond_t.reserve(nn * h->maxbond + 1));
  CPH_CUDA(h, h->d_angle_j.reserve(2 * nn * h->maxangle + 1));
  CPH_CUDA(h, h->d_angle_t.reserve(nn * h->maxangle + 1));
  CPH_CUDA(h, h->d_bonded_e.reserve(2));
  cudaStream_t st = h->stream;
  CPH_CUDA(h, cudaMemsetAsync(h->d_bonded_e.p, 0, 2 * sizeof(double), st));
  if (n == 0) return 0;
  CPH_CUDA(h, cudaMemsetAsync(h->d_flags.p + 6, 0, sizeof(unsigned int), st));
  h->nlaunch++;
  resolve_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_perm.p, h->d_tag.p, h->d_xq.p, h->d_neigh.p, h->d_numspec.p, h->rowcap,
                                          h->maxbond, h->d_num_bond.p, h->d_bond_type.p, h->d_bond_atom.p, h->maxangle,
                                          h->d_num_angle.p, h->d_angle_type.p, h->d_angle_a1.p, h->d_angle_a2.p,
                                          h->d_angle_a3.p, h->d_bcount.p, h->d_bond_j.p, h->d_bond_t.p, h->d_angle_j.p,
                                          h->d_angle_t.p, h->d_flags.p);
  CPH_CUDA(h, cudaGetLastError());
  unsigned int missing = 0;
  CPH_CUDA(h, cudaMemcpyAsync(&missing, h->d_flags.p + 6, sizeof(missing), cudaMemcpyDeviceToHost, st));
  CPH_CUDA(h, cudaStreamSynchronize(st));
  if (missing)
    return cph_fail(h, CPH_ERR_STATE, "a bond or angle partner is not among the atom's special neighbours "
                                      "(topology and special tables disagree, or a bond is longer than the list cutoff)");
  return 0;
}

int cph_launch_bonded(cph_handle *h, int eflag) {
  if (!h->have_topology || h->nlocal == 0) return 0;
  ProfScope ps(h, 4);
  const int n = h->nlocal;
  cudaStream_t st = h->stream;
  if (eflag) CPH_CUDA(h, cudaMemsetAsync(h->d_bonded_e.p, 0, 2 * sizeof(double), st));
  h->nlaunch++;
  if (eflag)
    bonded_kernel<1><<<nblk(n), TPB, 0, st>>>(n, h->d_xq.p, h->d_bcount.p, h->maxbond, h->d_bond_j.p, h->d_bond_t.p,
                                              h->d_bond_coef.p, h->maxangle, h->d_angle_j.p, h->d_angle_t.p,
                                              h->d_angle_coef.p, h->d_f.p, h->d_eatom.p, h->d_bonded_e.p);
  else
    bonded_kernel<0><<<nblk(n), TPB, 0, st>>>(n, h->d_xq.p, h->d_bcount.p, h->maxbond, h->d_bond_j.p, h->d_bond_t.p,
                                              h->d_bond_coef.p, h->maxangle, h->d_angle_j.p, h->d_angle_t.p,
                                              h->d_angle_coef.p, h->d_f.p, h->d_eatom.p, h->d_bonded_e.p);
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

int cph_bonded_energy(cph_handle *h, double *out2) {
  out2[0] = out2[1] = 0.0;
  if (!h->have_topology) return 0;
  CPH_CUDA(h, cudaMemcpyAsync(out2, h->d_bonded_e.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->nranks > 1) {   // owned shares -> box totals
    CPH_TRY(put(h, h->d_stage, out2, 2));
    CPH_TRY(cph_comm_allreduce(h, h->d_stage.p, 2));
    CPH_CUDA(h, cudaMemcpyAsync(out2, h->d_stage.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return 0;
}

// ---- fix nve ----------------------------------------------------------------------------------------
int cph_md_set_v(cph_handle *h, int where, const double *v) {
  const int n = h->nlocal;
  for (int t = 1; t <= h->pp.ntypes; t++)
    if (!(h->mass_h[t] > 0.0)) return cph_fail(h, CPH_ERR_STATE, "cph_set_mass first (type %d has no mass)", t);
  CPH_CUDA(h, h->d_v.reserve((size_t)n + 1));
  CPH_CUDA(h, h->d_v2.reserve((size_t)n + 1));
  if (n) {
    const double *vd = v;
    if (where == CPH_HOST) {
      CPH_TRY(put(h, h->d_stage, v, 3 * (size_t)n));
      vd = h->d_stage.p;
    }
    h->nlaunch++;
    scatter_v_kernel<<<nblk(n), TPB, 0, h->stream>>>(n, h->d_perm.p, vd, h->d_v.p);
    CPH_CUDA(h, cudaGetLastError());
    CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  h->md_on = true;
  return 0;
}

int cph_md_kick(cph_handle *h, double dt, int drift) {
  const int n = h->nlocal;
  if (n == 0) return 0;
  MassTable mt;
  for (int t = 0; t < CPH_MAXNT1; t++) mt.inv[t] = h->mass_h[t] > 0.0 ? 1.0 / h->mass_h[t] : 0.0;
  h->nlaunch++;
  nve_kernel<<<nblk(n), TPB, 0, h->stream>>>(n, h->d_xq.p, h->d_v.p, h->d_f.p, h->d_type.p, mt,
                                             0.5 * dt * h->fix.ftm2v, dt, drift);
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

int cph_md_wrap(cph_handle *h) {
  if (!h->md_on || h->nlocal == 0) return 0;
  // only along dimensions this rank spans on its own: across a decomposed dimension an atom that
  // leaves belongs to the neighbour rank, which is the host's migration (cph_set_atoms)
  int3 on = make_int3(h->periodic[0] && h->procgrid[0] == 1, h->periodic[1] && h->procgrid[1] == 1,
                      h->periodic[2] && h->procgrid[2] == 1);
  if (!on.x && !on.y && !on.z) return 0;
  h->nlaunch++;
  wrap_kernel<<<nblk(h->nlocal), TPB, 0, h->stream>>>(
      h->nlocal, h->d_xq.p, make_double3(h->boxlo[0], h->boxlo[1], h->boxlo[2]),
      make_double3(h->boxhi[0] - h->boxlo[0], h->boxhi[1] - h->boxlo[1], h->boxhi[2] - h->boxlo[2]), on);
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

void cph_bonded_release(cph_handle *h) {
  h->d_bond_coef.release(); h->d_angle_coef.release();
  DevBuf<int> *ib[] = {&h->d_num_bond, &h->d_bond_type, &h->d_bond_atom, &h->d_num_angle, &h->d_angle_type,
                       &h->d_angle_a1, &h->d_angle_a2, &h->d_angle_a3, &h->d_bcount, &h->d_bond_j, &h->d_bond_t,
                       &h->d_angle_j, &h->d_angle_t};
  for (auto *b : ib) b->release();
  h->d_bonded_e.release();
  h->d_v.release();
  h->d_v2.release();
}
