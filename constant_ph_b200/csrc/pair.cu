// pair.cu -- K2: fused fp64 pair pass for lj/cut/coul/cut and lj/cut/coul/dsf.
//
// Stands in for the LAMMPS pair style whose per-atom energies the reference reads at
// fix_constant_pH.cpp:216-219 (`force->pair->eatom`), restated from SURVEY.md Appendix A,
// and adds what north_star asks for in the same pass: forces, the electrostatic
// potential phi_i = dE_coul/dq_i that gives every site's dU/dlambda analytically
// (Appendix B), and the van-der-Waals share of the per-atom energy.
//
// Mapping: one warp per owned atom.  The warp streams the atom's neighbour row with
// 128-byte coalesced loads, gathers {x,y,z,q} of each candidate (one 32-byte sector),
// and tests the cutoff.  Because the Verlet skin makes ~40 % of the candidates fail the
// test, accepted candidates are ballot-compacted into a per-warp shared-memory tile; the
// expensive fp64 evaluation (rsqrt, exp, polynomial erfc) then runs on full tiles of 32
// with every lane active.  Accumulators are reduced across the warp with shuffles; a full
// list means no atomics and a fixed summation order (run-to-run bit reproducible).
//
// Per-atom outputs: f (3), evdwl_i = 1/2 sum_j evdwl_ij, phi_i, eatom_i = evdwl_i +
// 1/2 q_i phi_i  (== ev_tally's half-half split plus the dsf self term).
#include "cph_internal.h"

namespace {

constexpr int WARPS = 8;
constexpr int TPB = WARPS * 32;
constexpr int QCAP = 64;  // per-warp compaction tile (ring)

constexpr double EWALD_P = 0.3275911;
constexpr double A1 = 0.254829592, A2 = -0.284496736, A3 = 1.421413741, A4 = -1.453152027, A5 = 1.061405429;
constexpr double MY_PIS = 1.77245385090551602729;

// 1/sqrt(x): hardware seed (2^-23) + one third-order step -> ~2^-66, i.e. correctly rounded
// to within 1 ulp; far inside the 1e-10 parity budget and ~4x cheaper than the IEEE path.
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double t = x * y;
  double e = fma(-t, y, 1.0);
  double p = fma(0.375, e, 0.5);
  p = p * e;
  return fma(y, p, y);
}
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  double p = fma(e, e, e);
  return fma(y, p, y);
}

template <int STYLE, int EFLAG>
struct Acc {
  double fx = 0, fy = 0, fz = 0, ev = 0, phi = 0;
};

// evaluate one in-range pair; del = xi - xj
template <int STYLE, int EFLAG>
__device__ __forceinline__ void eval_pair(const PairParams &pp, const PairCoef &c, double delx, double dely,
                                          double delz, double rsq, double qi, double qj, int sb,
                                          Acc<STYLE, EFLAG> &a) {
  if (rsq >= c.cutsq) return;
  double factor_lj = 1.0, factor_coul = 1.0;
  if (sb) {   // rare (2 of ~420 pairs in water); selects instead of indexing keep pp out of local memory
    factor_lj = sb == 1 ? pp.special_lj[1] : (sb == 2 ? pp.special_lj[2] : pp.special_lj[3]);
    factor_coul = sb == 1 ? pp.special_coul[1] : (sb == 2 ? pp.special_coul[2] : pp.special_coul[3]);
  }
  const double rinv = fast_rsqrt(rsq);
  const double r2inv = rinv * rinv;
  double fpair = 0.0;
  if (rsq < c.cut_ljsq) {
    const double r6inv = r2inv * r2inv * r2inv;
    // lj1 = 12*lj3, lj2 = 6*lj4 (Appendix A coefficients)
    double forcelj = r6inv * (12.0 * c.lj3 * r6inv - 6.0 * c.lj4);
    fpair = factor_lj * forcelj * r2inv;
    if (EFLAG) a.ev += factor_lj * (r6inv * (c.lj3 * r6inv - c.lj4));
  }
  if (rsq < pp.cut_coulsq) {
    if (STYLE == CPH_PAIR_LJ_CUT_COUL_CUT) {
      const double k = pp.qqrd2e * factor_coul * rinv;   // E_ij = qi qj k
      fpair += qi * qj * k * r2inv;
      if (EFLAG) a.phi += qj * k;
    } else {
      const double r = rsq * rinv;
      const double erfcd = exp(-pp.alpha * pp.alpha * rsq);
      const double t = fast_rcp(fma(EWALD_P * pp.alpha, r, 1.0));
      const double erfcc = t * (A1 + t * (A2 + t * (A3 + t * (A4 + t * A5)))) * erfcd;
      const double pre = pp.qqrd2e * rinv;               // prefactor / (qi qj)
      // forcecoul*r2inv = prefactor*(erfcc/r + 2a/sqrt(pi)*erfcd + r*f_shift)*r * r2inv
      double fc = fma(erfcc, rinv, fma(2.0 * pp.alpha / MY_PIS, erfcd, r * pp.f_shift));
      double kk = erfcc - r * pp.e_shift - rsq * pp.f_shift;
      if (sb) {
        fc -= (1.0 - factor_coul) * rinv;
        kk -= (1.0 - factor_coul);
      }
      fpair += qi * qj * pre * fc * rinv;
      if (EFLAG) a.phi += qj * pre * kk;
    }
  }
  a.fx = fma(delx, fpair, a.fx);
  a.fy = fma(dely, fpair, a.fy);
  a.fz = fma(delz, fpair, a.fz);
}

template <int STYLE, int EFLAG>
__global__ void __launch_bounds__(TPB)
pair_kernel(int nlocal, const double4 *__restrict__ xq, const int *__restrict__ type,
            const int *__restrict__ neigh, const int *__restrict__ numneigh, int rowcap, PairParams pp,
            const PairCoef *__restrict__ coef, double *__restrict__ f, double *__restrict__ evdwl,
            double *__restrict__ phi, double *__restrict__ eatom) {
  __shared__ PairCoef s_coef[CPH_MAXNT1 * CPH_MAXNT1];
  __shared__ int s_q[WARPS][QCAP];
  const int nt1 = pp.ntypes + 1;
  for (int k = threadIdx.x; k < nt1 * nt1; k += TPB) s_coef[k] = coef[k];
  __syncthreads();

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * WARPS + w;
  if (i >= nlocal) return;
  int *q = s_q[w];
  const double4 pi = xq[i];
  const int ti = type[i];
  const PairCoef *ci = s_coef + ti * nt1;
  const int nn = numneigh[i];
  const int *row = neigh + (size_t)i * rowcap;
  Acc<STYLE, EFLAG> a;
  int head = 0, tail = 0;   // ring indices (warp-uniform)

  auto drain = [&](int count) {
    // lanes [0,count) each evaluate one queued pair
    if (lane < count) {
      int raw = q[(head + lane) & (QCAP - 1)];
      int j = raw & CPH_NEIGHMASK, sb = (raw >> CPH_SBSHIFT) & 3;
      double4 pj = xq[j];
      int tj = type[j];
      double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
      double rsq = delx * delx + dely * dely + delz * delz;
      eval_pair<STYLE, EFLAG>(pp, ci[tj], delx, dely, delz, rsq, pi.w, pj.w, sb, a);
    }
    head += count;
  };

  for (int k0 = 0; k0 < nn; k0 += 32) {
    int k = k0 + lane;
    bool in = false;
    int raw = 0;
    if (k < nn) {
      raw = row[k];
      double4 pj = xq[raw & CPH_NEIGHMASK];
      double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
      double rsq = delx * delx + dely * dely + delz * delz;
      in = rsq < pp.cutsq_max;
    }
    unsigned int m = __ballot_sync(0xffffffffu, in);
    if (in) q[(tail + __popc(m & ((1u << lane) - 1))) & (QCAP - 1)] = raw;
    tail += __popc(m);
    __syncwarp();
    if (tail - head >= 32) {
      drain(32);
      __syncwarp();
    }
  }
  if (tail - head > 0) drain(tail - head);

  // warp reduction in a fixed order
  for (int o = 16; o; o >>= 1) {
    a.fx += __shfl_xor_sync(0xffffffffu, a.fx, o);
    a.fy += __shfl_xor_sync(0xffffffffu, a.fy, o);
    a.fz += __shfl_xor_sync(0xffffffffu, a.fz, o);
    if (EFLAG) {
      a.ev += __shfl_xor_sync(0xffffffffu, a.ev, o);
      a.phi += __shfl_xor_sync(0xffffffffu, a.phi, o);
    }
  }
  if (lane == 0) {
    f[3 * (size_t)i] = a.fx;
    f[3 * (size_t)i + 1] = a.fy;
    f[3 * (size_t)i + 2] = a.fz;
    if (EFLAG) {
      double ev = 0.5 * a.ev;
      double ph = a.phi + 2.0 * pi.w * pp.c_self;   // dE_coul/dq_i including the dsf self term
      evdwl[i] = ev;
      phi[i] = ph;
      eatom[i] = ev + 0.5 * pi.w * ph;
    }
  }
}

}  // namespace

int cph_launch_pair(cph_handle *h, int eflag) {
  ProfScope ps(h, 0);
  const int n = h->nlocal;
  if (n == 0) return 0;
  const int blocks = (n + WARPS - 1) / WARPS;
  const PairCoef *dc = h->d_coef.p;
#define LAUNCH(S, E)                                                                                             \
  pair_kernel<S, E><<<blocks, TPB, 0, h->stream>>>(n, h->d_xq.p, h->d_type.p, h->d_neigh.p, h->d_numneigh.p,    \
                                                   h->rowcap, h->pp, dc, h->d_f.p, h->d_evdwl.p, h->d_phi.p,   \
                                                   h->d_eatom.p)
  if (h->pp.style == CPH_PAIR_LJ_CUT_COUL_CUT) {
    if (eflag) LAUNCH(CPH_PAIR_LJ_CUT_COUL_CUT, 1); else LAUNCH(CPH_PAIR_LJ_CUT_COUL_CUT, 0);
  } else {
    if (eflag) LAUNCH(CPH_PAIR_LJ_CUT_COUL_DSF, 1); else LAUNCH(CPH_PAIR_LJ_CUT_COUL_DSF, 0);
  }
#undef LAUNCH
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}
