// ljstates.cu -- LJ end states of titratable atoms (docs/SPEC.md "LJ end states"; include/cph_b200.h
// cph_set_lj_states).  The reference has no such term (it only rescales the forces on the hydrogen group,
// fix_constant_pH.cpp:149-171); north_star asks for "end states", and a titratable proton that keeps its LJ site
// in the deprotonated state is the usual gap of a charge-only scheme.
//
// An atom with a B-state type is the lambda-weighted superposition of its two types,
//     E_LJ(i,j) = sum_ab w_i^a w_j^b E_LJ(type_i^a, type_j^b; r),   w^A = 1 - lambda_site, w^B = lambda_site
// (w = 1, 0 for ordinary atoms).  The main pair kernel (pair.cu) is untouched: it evaluates every pair with the
// A types.  What is here adds the DIFFERENCE for the pairs that touch such an atom:
//   es_map_kernel          at every list build: B type and site of every owned + ghost atom (tag look-up), the
//                          list of owned atoms that have end states, the filter of A types that carry them
//   prune_kernel<ES=true>  (pair.cu) every prune: an atom WITHOUT end states gets the short list of the survivors
//                          of its inner row that have them (typically one or two entries)
//   es_pair_thread_kernel  every step, one THREAD per owned atom over that short list
//   es_pair_warp_kernel    every step, one WARP per owned atom WITH end states over its whole inner row; also
//                          g_i = dE/dlambda carried by the atom's own end states (joins dU/dlambda_s in K3)
// Every atom's correction comes from its OWN row (full list) in a fixed order: no atomics, no dependence on
// launch order, and ghosts with end states act on the owned atoms around them without any extra communication
// (lambda is replicated on every rank).
#include "cph_internal.h"

namespace {

constexpr int TPB = 256;

__global__ void es_map_kernel(int nall, int nlocal, const int *__restrict__ tag, int ntitr,
                              const int *__restrict__ tsorted, const int *__restrict__ entry_of_sorted,
                              const int *__restrict__ titr_typeB, const int *__restrict__ titr_site,
                              const int *__restrict__ type, int *es_tB, int *es_site, unsigned int *tmask,
                              int *own, int *nown, double *g) {
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
  if (k < nlocal) g[k] = 0.0;                       // only atoms with end states ever write theirs
  if (tb) {
    atomicOr(tmask, 1u << type[k]);                 // the prune's type filter: A-state types that carry end states
    if (k < nlocal) own[atomicAdd(nown, 1)] = k;    // any order: every atom is corrected on its own
  }
}

struct EsArgs {
  int nlocal, nt1, rowcap2, dummy, nown;
  const double4 *xq;
  const int *type, *es_tB, *es_site, *cnt, *ent, *own, *neigh2, *numneigh2;
  const double *lam;
  const double4 *coef;     // {12 lj3, 6 lj4, lj3, lj4}
  const double2 *cuts;     // {cut_ljsq, cutsq}
  double flj[4];
  double *f, *evdwl, *eatom, *g;
};

struct EsAcc {
  double fx = 0, fy = 0, fz = 0, ev = 0, g = 0;
};

// one corrected pair seen from atom i (types tis[], weights wi[]); e is the inner-row entry of its partner
template <int EFLAG>
__device__ __forceinline__ void es_one(const EsArgs &A, const double4 &pi, const int tBi, const int *tis,
                                       const double *wi, const int e, EsAcc &a) {
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
  for (int ia = 0; ia <= (tBi ? 1 : 0); ia++)
    for (int ib = 0; ib <= (tBj ? 1 : 0); ib++) {
      const int t2 = tis[ia] * A.nt1 + tjs[ib];
      if (rsq >= A.cuts[t2].x) continue;
      const double4 c = A.coef[t2];
      const double e_ab = r6inv * (c.z * r6inv - c.w);
      const double f_ab = r6inv * (c.x * r6inv - c.y);
      const double w = wi[ia] * wj[ib];
      emix += w * e_ab;
      fmix += w * f_ab;
      if (tBi) dEi += (ia ? wj[ib] : -wj[ib]) * e_ab;
      if (ia == 0 && ib == 0) { eAA = e_ab; fAA = f_ab; }   // what the main kernel has already added
    }
  const double fl = A.flj[sb];
  const double fp = fl * (fmix - fAA) * r2inv;
  a.fx += dx * fp; a.fy += dy * fp; a.fz += dz * fp;
  if (EFLAG) {
    a.ev += 0.5 * fl * (emix - eAA);
    a.g += fl * dEi;
  }
}

// atoms without end states: the short list the prune wrote (entry-major: consecutive threads, consecutive words)
template <int EFLAG>
__global__ void __launch_bounds__(TPB) es_pair_thread_kernel(const __grid_constant__ EsArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.nlocal) return;
  const int c = A.cnt[i];
  if (c == 0) return;
  const double4 pi = A.xq[i];
  const int tis[2] = {A.type[i], 0};
  const double wi[2] = {1.0, 0.0};
  EsAcc a;
  for (int k = 0; k < c; k++) es_one<EFLAG>(A, pi, 0, tis, wi, A.ent[(size_t)k * A.nlocal + i], a);
  A.f[3 * (size_t)i] += a.fx; A.f[3 * (size_t)i + 1] += a.fy; A.f[3 * (size_t)i + 2] += a.fz;
  if (EFLAG) {
    A.evdwl[i] += a.ev;
    A.eatom[i] += a.ev;
  }
}

// atoms with end states: every entry of the inner row, lanes striding, fixed shuffle tree
template <int EFLAG>
__global__ void __launch_bounds__(TPB) es_pair_warp_kernel(const __grid_constant__ EsArgs A) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= A.nown) return;
  const int i = A.own[w];
  const int tBi = A.es_tB[i];
  const double4 pi = A.xq[i];
  const double li = A.lam[A.es_site[i]];
  const double wi[2] = {1.0 - li, li};
  const int tis[2] = {A.type[i], tBi};
  const int n2 = A.numneigh2[i] & 0xffffff;
  const int *row = A.neigh2 + (size_t)i * A.rowcap2;
  EsAcc a;
  for (int k = lane; k < n2; k += 32) {
    const int e = row[k];
    if ((e & CPH_JMASK) != A.dummy) es_one<EFLAG>(A, pi, tBi, tis, wi, e, a);
  }
  for (int o = 16; o; o >>= 1) {
    a.fx += __shfl_xor_sync(0xffffffffu, a.fx, o);
    a.fy += __shfl_xor_sync(0xffffffffu, a.fy, o);
    a.fz += __shfl_xor_sync(0xffffffffu, a.fz, o);
    if (EFLAG) {
      a.ev += __shfl_xor_sync(0xffffffffu, a.ev, o);
      a.g += __shfl_xor_sync(0xffffffffu, a.g, o);
    }
  }
  if (lane == 0) {
    A.f[3 * (size_t)i] += a.fx; A.f[3 * (size_t)i + 1] += a.fy; A.f[3 * (size_t)i + 2] += a.fz;
    if (EFLAG) {
      A.evdwl[i] += a.ev;
      A.eatom[i] += a.ev;
      A.g[i] = a.g;
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

// after a list build: per-atom B type / site for owned atoms, ghosts and the dummy slot; the owned atoms that have
// end states; the type filter of the prune
int cph_ljstates_map(cph_handle *h) {
  if (!h->lj_states) return 0;
  const int nall = h->nall, n = h->nlocal;
  cudaStream_t st = h->stream;
  CPH_CUDA(h, h->d_es_tB.reserve((size_t)nall + 2));
  CPH_CUDA(h, h->d_es_site.reserve((size_t)nall + 2));
  CPH_CUDA(h, h->d_es_tmask.reserve(1));
  CPH_CUDA(h, h->d_es_word.reserve(2));               // [0] number of owned atoms with end states, [1] largest list
  CPH_CUDA(h, h->d_es_own.reserve((size_t)std::min(n, h->ntitr) + 2));
  CPH_CUDA(h, h->d_es_g.reserve((size_t)n + 2));
  CPH_CUDA(h, cudaMemsetAsync(h->d_es_tmask.p, 0, sizeof(unsigned int), st));
  CPH_CUDA(h, cudaMemsetAsync(h->d_es_word.p, 0, 2 * sizeof(int), st));
  es_map_kernel<<<nblk(nall + 1), TPB, 0, st>>>(nall, n, h->d_tag.p, h->ntitr, h->d_titr_tag_sorted.p,
                                                h->d_titr_entry_of_sorted.p, h->d_titr_typeB.p, h->d_titr_site.p,
                                                h->d_type.p, h->d_es_tB.p, h->d_es_site.p, h->d_es_tmask.p,
                                                h->d_es_own.p, h->d_es_word.p, h->d_es_g.p);
  CPH_CUDA(h, cudaGetLastError());
  CPH_CUDA(h, cudaMemcpyAsync(&h->es_nown, h->d_es_word.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  CPH_CUDA(h, cudaStreamSynchronize(st));
  h->nlaunch++;
  return 0;
}

// what prune_kernel<ES> writes: sized for the current atom count and list capacity
int cph_ljstates_lists(cph_handle *h, const int **tB, const unsigned int **tmask, int **cnt, int **ent, int **over,
                       int *cap) {
  const int n = h->nlocal;
  CPH_CUDA(h, h->d_es_cnt.reserve((size_t)n + 2));
  CPH_CUDA(h, h->d_es_ent.reserve((size_t)n * h->es_cap + 32));
  CPH_CUDA(h, cudaMemsetAsync(h->d_es_word.p + 1, 0, sizeof(int), h->stream));
  *tB = h->d_es_tB.p; *tmask = h->d_es_tmask.p; *cnt = h->d_es_cnt.p; *ent = h->d_es_ent.p;
  *over = h->d_es_word.p + 1; *cap = h->es_cap;
  return 0;
}

// every step, behind the pair pass: adds the end-state difference to f / evdwl / eatom, writes g
int cph_launch_ljstates(cph_handle *h, int eflag) {
  if (!h->lj_states) return 0;
  const int n = h->nlocal;
  if (n == 0) return 0;
  EsArgs A;
  A.nlocal = n; A.nt1 = h->pp.ntypes + 1; A.rowcap2 = h->rowcap2; A.dummy = h->nall; A.nown = h->es_nown;
  A.xq = h->d_xq.p; A.type = h->d_type.p; A.es_tB = h->d_es_tB.p; A.es_site = h->d_es_site.p;
  A.cnt = h->d_es_cnt.p; A.ent = h->d_es_ent.p; A.own = h->d_es_own.p;
  A.neigh2 = h->d_neigh2.p; A.numneigh2 = h->d_numneigh2.p; A.lam = h->d_lam.p;
  A.coef = h->d_coef4.p; A.cuts = h->d_cut2.p;
  for (int k = 0; k < 4; k++) A.flj[k] = h->pp.special_lj[k];
  A.f = h->d_f.p; A.evdwl = h->d_evdwl.p; A.eatom = h->d_eatom.p; A.g = h->d_es_g.p;
  if (eflag) es_pair_thread_kernel<1><<<nblk(n), TPB, 0, h->stream>>>(A);
  else es_pair_thread_kernel<0><<<nblk(n), TPB, 0, h->stream>>>(A);
  if (h->es_nown > 0) {
    const int blocks = (int)(((size_t)h->es_nown * 32 + TPB - 1) / TPB);
    if (eflag) es_pair_warp_kernel<1><<<blocks, TPB, 0, h->stream>>>(A);
    else es_pair_warp_kernel<0><<<blocks, TPB, 0, h->stream>>>(A);
    h->nlaunch++;
  }
  CPH_CUDA(h, cudaGetLastError());
  h->nlaunch++;
  return 0;
}
