// ljstates.cu -- LJ end states of titratable atoms (docs/SPEC.md "LJ end states"; include/cph_b200.h
// cph_set_lj_states).  The reference has no such term (it only rescales the forces on the hydrogen group,
// fix_constant_pH.cpp:149-171); north_star asks for "end states", and a titratable proton that keeps its LJ site
// in the deprotonated state is the usual gap of a charge-only scheme.
//
// An atom with a B-state type is the lambda-weighted superposition of its two types,
//     E_LJ(i,j) = sum_ab w_i^a w_j^b E_LJ(type_i^a, type_j^b; r),   w^A = 1 - lambda_site, w^B = lambda_site
// (w = 1, 0 for ordinary atoms).  The main pair kernel (pair.cu) is untouched: it evaluates every pair with the
// A types.  What is here adds the DIFFERENCE for the few pairs that touch such an atom:
//   es_map_kernel      at every list build: B type and site of every owned + ghost atom (tag look-up)
//   es_scan_kernel     after every prune: per owned atom, the entries of its inner row that need the correction
//                      (all of them when the atom itself has end states, else the partners that have), as a CSR
//                      list -- two passes (count, fill) around one exclusive scan, entry order = row order
//   es_pair_kernel     every step, one warp per owned atom with a non-empty list: forces, per-atom energy and
//                      g_i = dE/dlambda carried by atom i's own end states (joins dU/dlambda_s in K3, sites.cu).
// Every atom's correction comes from its OWN row (full list), so no atomics, no dependence on launch order,
// and ghosts with end states act on the owned atoms around them without any extra communication (lambda is
// replicated on every rank).
#include <cub/cub.cuh>

#include "cph_internal.h"

namespace {

constexpr int TPB = 256;

__global__ void es_map_kernel(int nall, const int *__restrict__ tag, int ntitr, const int *__restrict__ tsorted,
                              const int *__restrict__ entry_of_sorted, const int *__restrict__ titr_typeB,
                              const int *__restrict__ titr_site, const int *__restrict__ type, int *es_tB,
                              int *es_site, unsigned int *tmask) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nall) return;
  int tb = 0, site = -1;
  if (k < nall) {                                   // slot nall is the far-away dummy atom
    const int t = tag[k];
    int lo = 0, hi = ntitr - 1;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1, v = tsorted[mid];
      if (v == t) {
        const int e = entry_of_sorted[mid];
        tb = titr_typeB[e];
        site = titr_site[e];
        break;
      }
      if (v < t) lo = mid + 1; else hi = mid - 1;
    }
  }
  es_tB[k] = tb;
  es_site[k] = tb ? site : -1;
  if (tb) atomicOr(tmask, 1u << type[k]);           // the scan's type filter: A-state types that carry end states
}

// FILL == false: cnt[i] = number of entries of atom i's inner row that touch an end-state atom.
// FILL == true : the entries themselves (inner-row encoding: index | class << 26 | type << 28) at off[i]...
template <bool FILL>
__global__ void __launch_bounds__(TPB)
es_scan_kernel(int nlocal, const int *__restrict__ neigh2, const int *__restrict__ numneigh2, int rowcap2,
               const int *__restrict__ es_tB, const unsigned int *__restrict__ tmask_p, int dummy, int *cnt_or_off,
               int *ent) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= nlocal) return;
  const int n2 = numneigh2[i] & 0xffffff;
  const unsigned int tmask = *tmask_p;
  const bool self = es_tB[i] != 0;
  const int *row = neigh2 + (size_t)i * rowcap2;
  int cnt = 0;
  const int off = FILL ? cnt_or_off[i] : 0;
  for (int k0 = 0; k0 < n2; k0 += 32) {
    const int k = k0 + lane;
    bool q = false;
    int e = 0;
    if (k < n2) {
      e = row[k];
      const int j = e & CPH_JMASK;
      // the type filter saves the gather of es_tB[j] for nearly every entry
      q = j != dummy && (self || (((tmask >> ((unsigned int)e >> CPH_TYPESHIFT)) & 1u) && es_tB[j] != 0));
    }
    const unsigned int m = __ballot_sync(0xffffffffu, q);
    if (FILL && q) ent[off + cnt + __popc(m & ((1u << lane) - 1))] = e;
    cnt += __popc(m);
  }
  if (!FILL && lane == 0) cnt_or_off[i] = cnt;
}

struct EsArgs {
  int nlocal, nt1;
  const double4 *xq;
  const int *type, *es_tB, *es_site, *off, *ent;
  const double *lam;
  const double4 *coef;     // {12 lj3, 6 lj4, lj3, lj4}
  const double2 *cuts;     // {cut_ljsq, cutsq}
  double flj[4];
  double *f, *evdwl, *eatom, *g;
};

template <int EFLAG>
__global__ void __launch_bounds__(TPB) es_pair_kernel(const __grid_constant__ EsArgs A) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= A.nlocal) return;
  const int o0 = A.off[i], o1 = A.off[i + 1];
  const int tBi = A.es_tB[i];
  if (o0 == o1) {
    if (EFLAG && lane == 0) A.g[i] = 0.0;
    return;
  }
  const double4 pi = A.xq[i];
  const int tAi = A.type[i];
  const double li = tBi ? A.lam[A.es_site[i]] : 0.0;
  const double wi[2] = {1.0 - li, li};
  const int tis[2] = {tAi, tBi};
  double fx = 0, fy = 0, fz = 0, ev = 0, g = 0;
  for (int k = o0 + lane; k < o1; k += 32) {
    const int e = A.ent[k];
    const int j = e & CPH_JMASK, sb = (e >> CPH_SB2SHIFT) & 3, tAj = (int)((unsigned int)e >> CPH_TYPESHIFT);
    const int tBj = A.es_tB[j];
    const double lj = tBj ? A.lam[A.es_site[j]] : 0.0;
    const double wj[2] = {1.0 - lj, lj};
    const int tjs[2] = {tAj, tBj};
    const double4 pj = A.xq[j];
    const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
    const double rsq = dx * dx + dy * dy + dz * dz;
    const double r2inv = 1.0 / rsq, r6inv = r2inv * r2inv * r2inv;
    double emix = 0, fmix = 0, dEi = 0, eAA = 0, fAA = 0;
    for (int a = 0; a <= (tBi ? 1 : 0); a++)
      for (int b = 0; b <= (tBj ? 1 : 0); b++) {
        const int t2 = tis[a] * A.nt1 + tjs[b];
        if (rsq >= A.cuts[t2].x) continue;
        const double4 c = A.coef[t2];
        const double e_ab = r6inv * (c.z * r6inv - c.w);
        const double f_ab = r6inv * (c.x * r6inv - c.y);
        const double w = wi[a] * wj[b];
        emix += w * e_ab;
        fmix += w * f_ab;
        if (tBi) dEi += (a ? wj[b] : -wj[b]) * e_ab;
        if (a == 0 && b == 0) { eAA = e_ab; fAA = f_ab; }   // what the main kernel has already added
      }
    const double fl = A.flj[sb];
    const double fp = fl * (fmix - fAA) * r2inv;
    fx += dx * fp; fy += dy * fp; fz += dz * fp;
    if (EFLAG) {
      ev += 0.5 * fl * (emix - eAA);
      g += fl * dEi;
    }
  }
  for (int o = 16; o; o >>= 1) {
    fx += __shfl_xor_sync(0xffffffffu, fx, o);
    fy += __shfl_xor_sync(0xffffffffu, fy, o);
    fz += __shfl_xor_sync(0xffffffffu, fz, o);
    if (EFLAG) {
      ev += __shfl_xor_sync(0xffffffffu, ev, o);
      g += __shfl_xor_sync(0xffffffffu, g, o);
    }
  }
  if (lane == 0) {
    A.f[3 * (size_t)i] += fx; A.f[3 * (size_t)i + 1] += fy; A.f[3 * (size_t)i + 2] += fz;
    if (EFLAG) {
      A.evdwl[i] += ev;
      A.eatom[i] += ev;
      A.g[i] = g;
    }
  }
}

inline int nblk(int n) { return (n + TPB - 1) / TPB; }

}  // namespace

// typeB per titration entry in the CALLER's order of cph_set_sites; stored in the library's site-major order
int cph_ljstates_set(cph_handle *h, int ntitr, const int *typeB) {
  if (!h->have_sites || ntitr != h->ntitr)
    return cph_fail(h, CPH_ERR_ARG, "cph_set_lj_states: %d entries, cph_set_sites was given %d", ntitr, h->ntitr);
  if (!h->have_pair) return cph_fail(h, CPH_ERR_STATE, "cph_set_pair_style before cph_set_lj_states");
  if (h->have_atoms) return cph_fail(h, CPH_ERR_STATE, "cph_set_lj_states before cph_set_atoms");
  if (ntitr > 0 && !typeB) return cph_fail(h, CPH_ERR_ARG, "typeB is NULL");
  std::vector<int> tb(ntitr, 0);
  bool any = false;
  for (int k = 0; k < ntitr; k++) {
    const int t = typeB[h->titr_order_h[k]];
    if (t < 0 || t > h->pp.ntypes) return cph_fail(h, CPH_ERR_ARG, "cph_set_lj_states: type %d out of range", t);
    tb[k] = t;
    any = any || t != 0;
  }
  h->lj_states = any;
  if (!any) return CPH_OK;
  CPH_CUDA(h, h->d_titr_typeB.reserve(ntitr + 1));
  CPH_CUDA(h, cudaMemcpyAsync(h->d_titr_typeB.p, tb.data(), ntitr * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return CPH_OK;
}

// after a list build: per-atom B type / site for owned atoms, ghosts and the dummy slot; the type filter of the scan
int cph_ljstates_map(cph_handle *h) {
  if (!h->lj_states) return 0;
  const int nall = h->nall;
  CPH_CUDA(h, h->d_es_tB.reserve((size_t)nall + 2));
  CPH_CUDA(h, h->d_es_site.reserve((size_t)nall + 2));
  CPH_CUDA(h, h->d_es_tmask.reserve(1));
  CPH_CUDA(h, cudaMemsetAsync(h->d_es_tmask.p, 0, sizeof(unsigned int), h->stream));
  es_map_kernel<<<nblk(nall + 1), TPB, 0, h->stream>>>(nall, h->d_tag.p, h->ntitr, h->d_titr_tag_sorted.p,
                                                      h->d_titr_entry_of_sorted.p, h->d_titr_typeB.p,
                                                      h->d_titr_site.p, h->d_type.p, h->d_es_tB.p, h->d_es_site.p,
                                                      h->d_es_tmask.p);
  CPH_CUDA(h, cudaGetLastError());
  h->nlaunch++;
  return 0;
}

// after a prune: the CSR list of inner-row entries that need the correction
int cph_ljstates_collect(cph_handle *h) {
  if (!h->lj_states) return 0;
  const int n = h->nlocal;
  if (n == 0) return 0;
  cudaStream_t st = h->stream;
  CPH_CUDA(h, h->d_es_cnt.reserve((size_t)n + 2));
  CPH_CUDA(h, h->d_es_off.reserve((size_t)n + 2));
  const int blocks = (int)(((size_t)n * 32 + TPB - 1) / TPB);
  CPH_CUDA(h, cudaMemsetAsync(h->d_es_cnt.p + n, 0, sizeof(int), st));
  es_scan_kernel<false><<<blocks, TPB, 0, st>>>(n, h->d_neigh2.p, h->d_numneigh2.p, h->rowcap2, h->d_es_tB.p,
                                                h->d_es_tmask.p, h->nall, h->d_es_cnt.p, nullptr);
  size_t tmp = 0;
  CPH_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, tmp, h->d_es_cnt.p, h->d_es_off.p, n + 1, st));
  CPH_CUDA(h, h->d_cubtmp.reserve(tmp));
  CPH_CUDA(h, cub::DeviceScan::ExclusiveSum(h->d_cubtmp.p, tmp, h->d_es_cnt.p, h->d_es_off.p, n + 1, st));
  int total = 0;
  CPH_CUDA(h, cudaMemcpyAsync(&total, h->d_es_off.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  CPH_CUDA(h, cudaStreamSynchronize(st));
  h->es_entries = total;
  CPH_CUDA(h, h->d_es_ent.reserve((size_t)total + 32));
  es_scan_kernel<true><<<blocks, TPB, 0, st>>>(n, h->d_neigh2.p, h->d_numneigh2.p, h->rowcap2, h->d_es_tB.p,
                                               h->d_es_tmask.p, h->nall, h->d_es_off.p, h->d_es_ent.p);
  CPH_CUDA(h, cudaGetLastError());
  h->nlaunch += 3;
  return 0;
}

// every step, behind the pair pass: adds the end-state difference to f / evdwl / eatom, writes g
int cph_launch_ljstates(cph_handle *h, int eflag) {
  if (!h->lj_states) return 0;
  const int n = h->nlocal;
  if (n == 0) return 0;
  EsArgs A;
  A.nlocal = n; A.nt1 = h->pp.ntypes + 1;
  A.xq = h->d_xq.p; A.type = h->d_type.p; A.es_tB = h->d_es_tB.p; A.es_site = h->d_es_site.p;
  A.off = h->d_es_off.p; A.ent = h->d_es_ent.p; A.lam = h->d_lam.p;
  A.coef = h->d_coef4.p; A.cuts = h->d_cut2.p;
  for (int k = 0; k < 4; k++) A.flj[k] = h->pp.special_lj[k];
  CPH_CUDA(h, h->d_es_g.reserve((size_t)n + 2));
  A.f = h->d_f.p; A.evdwl = h->d_evdwl.p; A.eatom = h->d_eatom.p; A.g = h->d_es_g.p;
  const int blocks = (int)(((size_t)n * 32 + TPB - 1) / TPB);
  if (eflag) es_pair_kernel<1><<<blocks, TPB, 0, h->stream>>>(A);
  else es_pair_kernel<0><<<blocks, TPB, 0, h->stream>>>(A);
  CPH_CUDA(h, cudaGetLastError());
  h->nlaunch++;
  return 0;
}
