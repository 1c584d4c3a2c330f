// sites.cu -- K3 (energy partition + per-site reduction), K4 (lambda integrator),
// K5 (charge / force update) and the caller-order <-> internal-order movers.
//
// Reference lines restated here (cpp:N = fix_constant_pH.cpp):
//   partition_kernel    cpp:259-267  HA = sum eatom, HB = sum eatom over atoms NOT in the hydrogen group
//   site_sum_kernel     cpp:264-267 per site, plus north_star's dU/dlambda_s = sum dq_i * dE/dq_i
//   integrate_kernel    cpp:109-117 (integrator), cpp:120-124 (f, df), cpp:128-145 (U, dU)
//   set_force_kernel    cpp:149-171
// All sums are two-stage with a fixed block count so results are bit-reproducible run to run.
#include <cmath>

#include "cph_internal.h"

namespace {

constexpr int TPB = 256;
constexpr int MAXPART = 1024;   // stage-1 blocks

inline int nblk(int n) { return (n + TPB - 1) / TPB; }

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double *partials) {
  __shared__ double sm[NV][TPB / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < NV; c++) {
    double x = v[c];
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sm[c][w] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0;
    for (int k = 0; k < TPB / 32; k++) s += sm[threadIdx.x][k];
    partials[(size_t)blockIdx.x * NV + threadIdx.x] = s;
  }
}

// stage 2: out[c] = sum_b partials[b][c], one warp per value, fixed order
template <int NV>
__global__ void final_sum_kernel(int nb, const double *__restrict__ partials, double *out) {
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
  if (c >= NV) return;
  double s = 0;
  for (int b = lane; b < nb; b += 32) s += partials[(size_t)b * NV + c];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[c] = s;
}

// K3, one launch: compute_Hs tail (cpp:259-267: HA, HB over the owned atoms, plus E_vdwl, E_coul) in the first
// nbP blocks, the per-site sums (cpp:264-267 per site + north_star's dU/dlambda_s = sum dq_i dE/dq_i) in the
// rest: a group of lanes per site over its contiguous range of the site-major titratable-atom table, lanes striding,
// shuffle butterfly -- a fixed order whatever the site size, no atomics.  The last block to finish (ticket counter)
// adds the block partials in a fixed order and writes red[0..3].
__global__ void __launch_bounds__(TPB)
site_partition_kernel(int n, const double *__restrict__ eatom, const double *__restrict__ evdwl,
                      const int *__restrict__ mask, int Hbit, int nbP, double *partials, int S, int lps,
                      const int *__restrict__ site_start, const int *__restrict__ titr_local,
                      const double *__restrict__ titr_dq, const double *__restrict__ phi,
                      const double *__restrict__ lj_g, const double *__restrict__ extra_dudl, int implicit_site,
                      double extra_HA, double extra_HB, const double *__restrict__ bonded_e, double *red,
                      unsigned int *ticket, const __grid_constant__ MailRed mr) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if ((int)blockIdx.x < nbP) {
    double v[4] = {0, 0, 0, 0};   // HA, HB, E_vdwl, E_coul
    for (int k = blockIdx.x * TPB + threadIdx.x; k < n; k += nbP * TPB) {
      double e = eatom[k], ev = evdwl[k];
      v[0] += e;                                  // cpp:265
      if (!(mask[k] & Hbit)) v[1] += e;           // cpp:266
      v[2] += ev;
      v[3] += e - ev;
    }
    block_reduce_store<4>(v, partials);
  } else {
    // lps lanes per site (a power of two chosen from the largest site: 1 for one-atom sites, 8 for carboxyl /
    // amine groups, 32 beyond), lanes striding over the site's range, butterfly inside the group
    const int site = ((int)blockIdx.x - nbP) * (TPB / lps) + (int)threadIdx.x / lps;
    const int sub = (int)threadIdx.x & (lps - 1);
    double d = 0, hd = 0;
    if (site < S) {
      for (int t = site_start[site] + sub; t < site_start[site + 1]; t += lps) {
        const int k = titr_local[t];
        if (k >= 0) {                                         // owned by this rank (cpp:264: i < nlocal)
          d += titr_dq[t] * phi[k];                           // Appendix B
          if (lj_g) d += lj_g[k];                             // LJ end states of atom k (ljstates.cu)
          if (mask[k] & Hbit) hd -= eatom[k];                 // HB_s - HA_s
        }
      }
    }
    for (int o = lps >> 1; o; o >>= 1) {
      d += __shfl_xor_sync(0xffffffffu, d, o);
      hd += __shfl_xor_sync(0xffffffffu, hd, o);
    }
    if (site < S) {
      // host-tallied dE/dlambda_s of this rank (KSpace stays with LAMMPS: cpp:241-244, cph_set_extra_dudl)
      if (extra_dudl) d += extra_dudl[site];
      if (sub == 0) {
        red[4 + site] = d;
        if (!implicit_site) red[4 + S + site] = hd;
      }
      // one-shot all-reduce over NVLink, push side: this rank's sums of the site go straight into its slot of
      // every rank's mailbox (the group's lanes share the ranks; its own mailbox included)
      for (int p = sub; p < mr.P; p += lps) {
        mr.dst[p][4 + site] = d;
        if (!implicit_site) mr.dst[p][4 + S + site] = hd;
      }
    }
  }
  __shared__ unsigned int s_last;
  if (mr.P) __threadfence_system(); else __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ double out[4];
  if (w < 4) {
    double a = 0;
    for (int b = lane; b < nbP; b += 32) a += partials[(size_t)b * 4 + w];
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    // extra_HA/HB: this rank's share of the per-atom energies LAMMPS tallied on the host (bonded styles,
    // KSpace: cpp:221-244), already partitioned by the fix exactly as cpp:264-267 does
    if (lane == 0) {
      if (w == 0) a += extra_HA;
      if (w == 1) a += extra_HB;
      // E_coul is summed as eatom - evdwl; with bonded terms on the device (f2) eatom also holds their shares
      if (w == 3 && bonded_e) a -= bonded_e[0] + bonded_e[1];
      out[w] = a;
      red[w] = a;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (implicit_site) red[4 + S] = out[1] - out[0];   // reference single site: HB - HA of the hydrogen group (cpp:111)
    *ticket = 0u;
  }
  // every block has fenced its stores system-wide before taking its ticket: the scalars go out, then the
  // sequence number that tells the gathering kernels on all ranks that this rank's block is complete
  if ((int)threadIdx.x < mr.P) {
    double *dst = mr.dst[threadIdx.x];
    for (int c = 0; c < 4; c++) dst[c] = out[c];
    if (implicit_site) dst[4 + S] = out[1] - out[0];
    dst[4 + 2 * S] = red[4 + 2 * S];                   // modify_water slot (filled before this launch)
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(dst + mr.seq_index) = mr.seq;
  }
}

// One-shot all-reduce, gather side: wait until every rank's block of this reduction has arrived in MY mailbox.
// Called by all threads of a block; the first P threads watch one rank each.
__device__ __forceinline__ void mail_wait(const MailRed &mr, unsigned int *status) {
  if ((int)threadIdx.x < mr.P) {
    const volatile unsigned long long *sq =
        reinterpret_cast<const volatile unsigned long long *>(mr.src[threadIdx.x] + mr.seq_index);
    const long long t0 = clock64();
    while (*sq != mr.seq) {
      if (clock64() - t0 > 400000000000LL) { atomicOr(status, 1u); break; }   // minutes: a rank is gone
      __nanosleep(200);
    }
    __threadfence_system();
  }
  __syncthreads();
}
// total of entry k over the ranks, in rank order: every rank computes bit-identical sums
__device__ __forceinline__ double mail_total(const MailRed &mr, int k) {
  double t = 0.0;
  for (int p = 0; p < mr.P; p++) t += reinterpret_cast<const volatile double *>(mr.src[p])[k];
  return t;
}

// stand-alone gather (cph_site_reduce called on its own): totals of all entries into red[]
__global__ void __launch_bounds__(TPB)
red_gather_kernel(int nred, const __grid_constant__ MailRed mr, double *red, unsigned int *status) {
  mail_wait(mr, status);
  const int k = blockIdx.x * TPB + threadIdx.x;
  if (k < nred) red[k] = mail_total(mr, k);
}

struct BiasOut { double f, df, U, dU; };

__device__ __forceinline__ BiasOut bias_terms(const BiasParams &bp, double lambda) {
  const double a = bp.a, b = bp.b, s = bp.s, k = bp.k, d = bp.d, w = bp.w, r = bp.r, m = bp.m;
  const double SQRT_PI = 1.77245385090551602729;
  BiasOut o;
  double ex = exp(-50.0 * (lambda - 0.5));
  o.f = 1.0 / (1.0 + ex);                                                        // cpp:122
  double U1 = -k * exp(-(lambda - 1 - b) * (lambda - 1 - b) / (2 * a * a));      // cpp:132
  double U2 = -k * exp(-(lambda + b) * (lambda + b) / (2 * a * a));              // cpp:133
  double U3 = d * exp(-(lambda - 0.5) * (lambda - 0.5) / (2 * s * s));           // cpp:134
  double U4, U5, dU1, dU2, dU3, dU4, dU5;
  if (bp.mode == CPH_BIAS_EXACT) {
    o.df = 50.0 * ex * o.f * o.f;
    U4 = 0.5 * w * (1 - erf(r * (lambda + m)));                                  // cpp:135, erf (D16)
    U5 = 0.5 * w * (1 + erf(r * (lambda - 1 - m)));                              // cpp:136
    dU1 = -((lambda - 1 - b) / (a * a)) * U1;                                    // D14
    dU2 = -((lambda + b) / (a * a)) * U2;
    dU3 = -((lambda - 0.5) / (s * s)) * U3;                                      // cpp:139
    dU4 = -0.5 * w * r * 2 * exp(-r * r * (lambda + m) * (lambda + m)) / SQRT_PI;          // D15
    dU5 = 0.5 * w * r * 2 * exp(-r * r * (lambda - 1 - m) * (lambda - 1 - m)) / SQRT_PI;   // cpp:141
  } else {
    o.df = 50.0 * ex / (o.f * o.f);                                              // cpp:123 verbatim
    U4 = 0.5 * w * (1 - (double)erff((float)(r * (lambda + m))));                // cpp:135 verbatim
    U5 = 0.5 * w * (1 + (double)erff((float)(r * (lambda - 1 - m))));            // cpp:136
    dU1 = -((lambda - 1 - b) / (2 * a * a)) * U1;                                // cpp:137
    dU2 = -((lambda + b) / (2 * a * a)) * U2;                                    // cpp:138
    dU3 = -((lambda - 0.5) / (s * s)) * U3;                                      // cpp:139
    dU4 = -0.5 * w * r * 2 * exp(-r * r * (lambda + 0.5) * (lambda + 0.5)) / SQRT_PI;      // cpp:140
    dU5 = 0.5 * w * r * 2 * exp(-r * r * (lambda - 1 - m) * (lambda - 1 - m)) / SQRT_PI;   // cpp:141
  }
  o.U = U1 + U2 + U3 + U4 + U5;             // cpp:143
  o.dU = dU1 + dU2 + dU3 + dU4 + dU5;       // cpp:144
  return o;
}

// The same terms with the eight transcendental evaluations of a site spread over the eight lanes of its group
// (sub = lane within the group): one exp / erf call deep instead of eight.  Every lane returns all four values.
__device__ __forceinline__ BiasOut bias_terms_lanes(const BiasParams &bp, double lambda, int sub, int lane_base) {
  const double a = bp.a, b = bp.b, s = bp.s, k = bp.k, d = bp.d, w = bp.w, r = bp.r, m = bp.m;
  const double SQRT_PI = 1.77245385090551602729;
  const bool exact = bp.mode == CPH_BIAS_EXACT;
  const double m4 = exact ? m : 0.5;                                              // cpp:140 as written uses (lambda + 0.5)
  double arg;
  switch (sub) {
    case 0: arg = -50.0 * (lambda - 0.5); break;                                  // cpp:122
    case 1: arg = -(lambda - 1 - b) * (lambda - 1 - b) / (2 * a * a); break;      // cpp:132
    case 2: arg = -(lambda + b) * (lambda + b) / (2 * a * a); break;              // cpp:133
    case 3: arg = -(lambda - 0.5) * (lambda - 0.5) / (2 * s * s); break;          // cpp:134
    case 4: arg = r * (lambda + m); break;                                        // cpp:135
    case 5: arg = r * (lambda - 1 - m); break;                                    // cpp:136
    case 6: arg = -r * r * (lambda + m4) * (lambda + m4); break;                  // cpp:140
    default: arg = -r * r * (lambda - 1 - m) * (lambda - 1 - m); break;           // cpp:141
  }
  double val;
  if (sub == 4 || sub == 5) val = exact ? erf(arg) : (double)erff((float)arg);    // D16
  else val = exp(arg);
  double e[8];
#pragma unroll
  for (int q = 0; q < 8; q++) e[q] = __shfl_sync(0xffffffffu, val, lane_base + q);
  BiasOut o;
  o.f = 1.0 / (1.0 + e[0]);
  const double U1 = -k * e[1], U2 = -k * e[2], U3 = d * e[3];
  const double U4 = 0.5 * w * (1 - e[4]), U5 = 0.5 * w * (1 + e[5]);
  double dU1, dU2;
  if (exact) {
    o.df = 50.0 * e[0] * o.f * o.f;
    dU1 = -((lambda - 1 - b) / (a * a)) * U1;                                     // D14
    dU2 = -((lambda + b) / (a * a)) * U2;
  } else {
    o.df = 50.0 * e[0] / (o.f * o.f);                                             // cpp:123 verbatim
    dU1 = -((lambda - 1 - b) / (2 * a * a)) * U1;                                 // cpp:137
    dU2 = -((lambda + b) / (2 * a * a)) * U2;                                     // cpp:138
  }
  const double dU3 = -((lambda - 0.5) / (s * s)) * U3;                            // cpp:139
  const double dU4 = -0.5 * w * r * 2 * e[6] / SQRT_PI;
  const double dU5 = 0.5 * w * r * 2 * e[7] / SQRT_PI;
  o.U = U1 + U2 + U3 + U4 + U5;             // cpp:143
  o.dU = dU1 + dU2 + dU3 + dU4 + dU5;       // cpp:144
  return o;
}

struct LambdaArgs {
  int S, phase, thermo, apply, thermo_post, nw, lanes;
  double dt, SkT, Q, inv_nw;         // inv_nw = 1 / n_W when the water buffer is on, else 0
  BiasParams bp;
  FixParams fx;
  const double *pK, *dQ, *wq, *titr_qA, *titr_dq;
  const int *site_start, *titr_local, *wlocal;
  double *red, *lam, *theta, *vlam, *alam, *flam, *fs, *dfs, *Us, *dUs, *partials, *scal;
  double4 *xq;
  unsigned int *ticket;
  MailRed mr;                        // mr.P != 0: the site sums of all ranks wait in my mailbox (gather them first)
  unsigned int *status;
};

// K4 + K5, one launch.  Per site (one thread): calculate_df, calculate_dU, integrate_lambda (cpp:109-145) in
// the requested phase, then q_i = (1-lambda_s) qA_i + lambda_s qB_i on the site's own atoms (north_star).
//   phase 0: reference kinematic step (cpp:109-117)   phase 1: VV kick+drift
//   phase 2: VV force evaluation (a <- F/m)           phase 3: VV second kick + H_lambda
// The last block to finish (ticket counter) adds the block partials of H_lambda in a fixed order, takes the
// second Nose-Hoover half step and, with the water buffer on, moves -(1/n_W) sum_s lambda_s dQ_s onto the
// buffer atoms (modify_water, h:58).
__global__ void __launch_bounds__(TPB)
lambda_update_kernel(const __grid_constant__ LambdaArgs A) {
  const int S = A.S, phase = A.phase;
  const BiasParams &bp = A.bp;
  const FixParams &fx = A.fx;
  // thermo: Nose-Hoover on the site velocities (velocity-Verlet form only); scal[9] = exp(-xi dt/2)
  const double nh = A.thermo ? A.scal[9] : 1.0;
  const double dt = A.dt;
  // theta != NULL: the dynamical coordinate is theta with lambda = sin^2(theta) (north_star's lambda/theta
  // variables; absent from the reference, which integrates lambda itself and confines it with U4/U5);
  // velocity, acceleration and mass then refer to theta and F_theta = F_lambda * sin(2 theta).
  double *theta = A.theta;
  const bool gather = A.mr.P != 0;
  double wphi = 0.0;         // modify_water: -(dQ_s/n_W) sum_W dE/dq still to be applied to the gathered dU/dlambda_s
  if (gather) {
    mail_wait(A.mr, A.status);
    if (A.inv_nw != 0.0 && fx.dudl_mode == CPH_DUDL_CHARGE) wphi = A.inv_nw * mail_total(A.mr, 4 + 2 * S);
  }
  double v[3] = {0, 0, 0};   // sum of site terms of H_lambda, sum lambda*(HB_s-HA_s), kinetic
  // eight lanes per site: they share the site's eight exp / erf evaluations and its titratable atoms; lane 0 of
  // the group does the bookkeeping.  The loop bound is warp-uniform (whole groups), inactive groups idle inside.
  // (A.lanes == 1, chosen for very many sites where throughput matters more than latency: one thread per site.)
  const int gshift = A.lanes == 8 ? 3 : 0, glanes = 1 << gshift;
  const int sub = threadIdx.x & (glanes - 1), lane_base = threadIdx.x & 24;
  const int ngroups = gridDim.x * (TPB >> gshift);
  for (int s0 = blockIdx.x * (TPB >> gshift); s0 < S; s0 += ngroups) {
    const int s = min(s0 + (int)(threadIdx.x >> gshift), S - 1);
    const bool live = s0 + (int)(threadIdx.x >> gshift) < S, lead = live && sub == 0;
    double cq = theta ? theta[s] : A.lam[s];
    double vel = A.vlam[s], acc = A.alam[s];
    __syncwarp();            // every lane of the group holds the site's state before its lead lane overwrites it
    double lambda;
    if (phase == 1) {
      vel = vel * nh + 0.5 * acc * dt;
      cq += vel * dt;
      lambda = cq;
      if (theta) { const double sn = sin(cq); lambda = sn * sn; }
      if (lead) {
        if (theta) theta[s] = cq;
        A.lam[s] = lambda;
        A.vlam[s] = vel;
      }
    } else {
      double chain = 1.0;
      lambda = cq;
      if (theta) { const double sn = sin(cq); lambda = sn * sn; chain = sin(2.0 * cq); }
      if (phase == 3) vel = (vel + 0.5 * acc * dt) * nh;
      BiasOut b = glanes == 8 ? bias_terms_lanes(bp, lambda, sub, lane_base) : bias_terms(bp, lambda);
      const double pk = fx.implicit_site ? fx.pK : A.pK[s];
      double hd, dq_e;
      if (gather) {          // fused all-reduce: totals over the ranks, written back for the getters
        hd = mail_total(A.mr, 4 + S + s);
        dq_e = mail_total(A.mr, 4 + s);
        if (wphi != 0.0) dq_e -= A.dQ[s] * wphi;
        if (lead) {
          A.red[4 + S + s] = hd;
          A.red[4 + s] = dq_e;
        }
      } else {
        hd = A.red[4 + S + s];
        dq_e = A.red[4 + s];
      }
      const double dE = (fx.dudl_mode == CPH_DUDL_REFERENCE) ? hd : dq_e;
      const double ph = fx.boltz * fx.T * log(10.0) * (pk - fx.pH);
      const double f_lambda = -(dE + b.df * ph + b.dU);                 // cpp:111
      const double a_lambda = f_lambda * chain / bp.m_lambda * fx.ftm2v;   // cpp:112 (+ SURVEY D9)
      const double kin = 0.5 * bp.m_lambda * vel * vel / fx.ftm2v;
      if (lead) {
        v[0] += b.f * ph + b.U + kin;                                   // cpp:114 site terms
        v[1] += lambda * hd;                                            // cpp:114 lambda*(HB-HA)
        v[2] += kin;
        A.fs[s] = b.f; A.dfs[s] = b.df; A.Us[s] = b.U; A.dUs[s] = b.dU; A.flam[s] = f_lambda;
      }
      if (phase == 0) {
        cq = 0.5 * a_lambda * dt * dt + vel * dt + cq;                  // cpp:115
        vel = a_lambda * dt + vel;                                      // cpp:116
      }
      lambda = cq;
      if (theta) { const double sn = sin(cq); lambda = sn * sn; }
      if (lead) {
        if (theta) theta[s] = cq;
        A.lam[s] = lambda;
        A.vlam[s] = vel;
        A.alam[s] = a_lambda;
      }
    }
    if (A.apply && live)    // the site's own atoms follow its lambda at once (owned atoms only), one lane each
      for (int t = A.site_start[s] + sub; t < A.site_start[s + 1]; t += glanes) {
        const int k = A.titr_local[t];
        if (k >= 0) A.xq[k].w = A.titr_qA[t] + lambda * A.titr_dq[t];
      }
  }
  const bool sums = phase != 1;
  const bool water = A.apply && A.nw > 0;
  if (!sums && !water) return;
  if (sums) block_reduce_store<3>(v, A.partials);
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(A.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
  if (sums) {
    __shared__ double out[3];
    if (c < 3) {
      double a = 0;
      for (int b = lane; b < (int)gridDim.x; b += 32) a += A.partials[(size_t)b * 3 + c];
      for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) out[c] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (gather) {
        for (int c2 = 0; c2 < 4; c2++) A.red[c2] = mail_total(A.mr, c2);
        A.red[4 + 2 * S] = mail_total(A.mr, 4 + 2 * S);
      }
      // cpp:114: (1-lambda)HA + lambda HB = HA + lambda (HB-HA); charge mode: E_ff at the current charges
      const double eff = (fx.dudl_mode == CPH_DUDL_REFERENCE) ? A.red[0] + out[1] : A.red[2] + A.red[3];
      A.scal[4] = eff + out[0];
      A.scal[5] = out[2];
      if (A.thermo_post) {   // second Nose-Hoover half step, with the kinetic energy after the scaled final kick
        A.scal[10] += 0.5 * dt * A.scal[8];
        A.scal[8] += 0.5 * dt * (2.0 * out[2] - A.SkT) / A.Q;
        A.scal[11] = 0.5 * A.Q * A.scal[8] * A.scal[8] + A.SkT * A.scal[10];    // thermostat energy (conserved with H_lambda)
      }
    }
  }
  if (water) {   // tot = sum_s lambda_s dQ_s in a fixed order; the owned buffer atoms get q_base - tot / n_W
    __shared__ double sm[TPB];
    double a = 0;
    for (int s = threadIdx.x; s < S; s += TPB) a += A.lam[s] * A.dQ[s];
    sm[threadIdx.x] = a;
    __syncthreads();
    for (int o = TPB / 2; o; o >>= 1) {
      if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
      __syncthreads();
    }
    if ((int)threadIdx.x < A.nw && A.wlocal[threadIdx.x] >= 0)
      A.xq[A.wlocal[threadIdx.x]].w = A.wq[threadIdx.x] - sm[0] * A.inv_nw;
  }
  __syncthreads();
  if (threadIdx.x == 0) *A.ticket = 0u;
}

// Nose-Hoover half step before the first kick: xi += dt/2 (2K - S kT)/Q, eta += xi dt/2,
// scal[9] = exp(-xi dt/2).  scal[5] = K of the previous step, scal[8] = xi, scal[10] = eta.
__global__ void nh_pre_kernel(double *scal, double dt, double SkT, double Q) {
  if (threadIdx.x || blockIdx.x) return;
  double xi = scal[8] + 0.5 * dt * (2.0 * scal[5] - SkT) / Q;
  scal[8] = xi;
  scal[10] += 0.5 * dt * xi;
  scal[9] = exp(-0.5 * dt * xi);
}

// q_i = (1-lambda_s) qA_i + lambda_s qB_i  (north_star; the reference never touches atom->q)
__global__ void apply_charges_kernel(int ntitr, const int *__restrict__ titr_site, const int *__restrict__ titr_local,
                                     const double *__restrict__ qA, const double *__restrict__ dq,
                                     const double *__restrict__ lam, double4 *xq) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntitr) return;
  int k = titr_local[t];
  if (k < 0) return;
  xq[k].w = qA[t] + lam[titr_site[t]] * dq[t];
}

// modify_water (h:58, TODO at cpp:268): sum of dE/dq over the buffer atoms this rank owns
__global__ void water_phi_kernel(int nw, const int *__restrict__ wlocal, const double *__restrict__ phi, double *out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0;
  for (int k = 0; k < nw; k++)
    if (wlocal[k] >= 0) s += phi[wlocal[k]];
  *out = s;
}
// every buffer atom carries -(1/n_W) sum_s lambda_s dQ_s: dE/dlambda_s gains -(dQ_s/n_W) * sum_W dE/dq
__global__ void water_dudl_kernel(int S, const double *__restrict__ dQ, double inv_nw, double *red) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < S) red[4 + s] -= dQ[s] * inv_nw * red[4 + 2 * S];
}
// one block: tot = sum_s lambda_s dQ_s in a fixed order, then the owned buffer atoms get q_base - tot/n_W
__global__ void __launch_bounds__(TPB)
water_apply_kernel(int S, const double *__restrict__ lam, const double *__restrict__ dQ, double inv_nw, int nw,
                   const int *__restrict__ wlocal, const double *__restrict__ wq, double4 *xq) {
  __shared__ double sm[TPB];
  double v = 0;
  for (int s = threadIdx.x; s < S; s += TPB) v += lam[s] * dQ[s];
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int o = TPB / 2; o; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  const double tot = sm[0];
  if (threadIdx.x < nw && wlocal[threadIdx.x] >= 0) xq[wlocal[threadIdx.x]].w = wq[threadIdx.x] - tot * inv_nw;
}

// set_force (cpp:149-171) over the compact hydrogen-group list instead of a scan of mask[]
__global__ void set_force_kernel(int nh, const int *__restrict__ hlist, const int *__restrict__ site_of,
                                 const double *__restrict__ lam, int fscale_mode, double *f) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nh) return;
  int k = hlist[m];
  double l = lam[site_of[k]];
  double sc = fscale_mode == CPH_FSCALE_LAMBDA ? l : 1.0 - l;   // cpp:166-168 / SURVEY D17
  f[3 * (size_t)k] *= sc;
  f[3 * (size_t)k + 1] *= sc;
  f[3 * (size_t)k + 2] *= sc;
}

// new positions from the caller (caller order) + the neighbor->decide() displacement test
__global__ void set_x_kernel(int n, const double *__restrict__ xc, const int *__restrict__ perm,
                             const double *__restrict__ xbuild, double thresh2, const double *__restrict__ xinner,
                             double thresh2_inner, double3 lo, double3 hi, double4 *xq, unsigned int *flags) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  float d2f = 0.f, drift = 0.f;
  bool over = false, over_in = false;
  if (k < n) {
    size_t c = (size_t)perm[k] * 3;
    double x = xc[c], y = xc[c + 1], z = xc[c + 2];
    xq[k].x = x; xq[k].y = y; xq[k].z = z;
    // how far outside its sub-box the atom sits (the next list build sizes the ghost shell with it)
    const double e = fmax(fmax(lo.x - x, x - hi.x), fmax(fmax(lo.y - y, y - hi.y), fmax(lo.z - z, z - hi.z)));
    drift = e > 0 ? __double2float_ru(e) : 0.f;
    double dx = x - xbuild[3 * (size_t)k], dy = y - xbuild[3 * (size_t)k + 1], dz = z - xbuild[3 * (size_t)k + 2];
    double d2 = dx * dx + dy * dy + dz * dz;
    over = d2 > thresh2;
    d2f = __double2float_ru(d2);
    if (xinner) {   // the pruned inner rows stay valid while nobody moved more than inner_skin/2
      dx = x - xinner[3 * (size_t)k]; dy = y - xinner[3 * (size_t)k + 1]; dz = z - xinner[3 * (size_t)k + 2];
      over_in = dx * dx + dy * dy + dz * dz > thresh2_inner;
    }
  }
  for (int o = 16; o; o >>= 1) {
    d2f = fmaxf(d2f, __shfl_xor_sync(0xffffffffu, d2f, o));
    drift = fmaxf(drift, __shfl_xor_sync(0xffffffffu, drift, o));
  }
  unsigned int any = __ballot_sync(0xffffffffu, over);
  unsigned int any_in = __ballot_sync(0xffffffffu, over_in);
  if ((threadIdx.x & 31) == 0) {
    if (d2f > 0.f) atomicMax(flags + 0, __float_as_uint(d2f));
    if (drift > 0.f) atomicMax(flags + 2, __float_as_uint(drift));
    if (any) atomicOr(flags + 4, 1u);
    if (any_in) atomicOr(flags + 5, 1u);
  }
}

__global__ void check_kernel(int n, const double *__restrict__ xbuild, double thresh2, const double *__restrict__ xinner,
                             double thresh2_inner, double3 lo, double3 hi, const double4 *__restrict__ xq,
                             unsigned int *flags) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  float d2f = 0.f, drift = 0.f;
  bool over = false, over_in = false;
  if (k < n) {
    double4 p = xq[k];
    const double e = fmax(fmax(lo.x - p.x, p.x - hi.x), fmax(fmax(lo.y - p.y, p.y - hi.y), fmax(lo.z - p.z, p.z - hi.z)));
    drift = e > 0 ? __double2float_ru(e) : 0.f;
    double dx = p.x - xbuild[3 * (size_t)k], dy = p.y - xbuild[3 * (size_t)k + 1], dz = p.z - xbuild[3 * (size_t)k + 2];
    double d2 = dx * dx + dy * dy + dz * dz;
    over = d2 > thresh2;
    d2f = __double2float_ru(d2);
    if (xinner) {
      dx = p.x - xinner[3 * (size_t)k]; dy = p.y - xinner[3 * (size_t)k + 1]; dz = p.z - xinner[3 * (size_t)k + 2];
      over_in = dx * dx + dy * dy + dz * dz > thresh2_inner;
    }
  }
  for (int o = 16; o; o >>= 1) {
    d2f = fmaxf(d2f, __shfl_xor_sync(0xffffffffu, d2f, o));
    drift = fmaxf(drift, __shfl_xor_sync(0xffffffffu, drift, o));
  }
  unsigned int any = __ballot_sync(0xffffffffu, over);
  unsigned int any_in = __ballot_sync(0xffffffffu, over_in);
  if ((threadIdx.x & 31) == 0) {
    if (d2f > 0.f) atomicMax(flags + 0, __float_as_uint(d2f));
    if (drift > 0.f) atomicMax(flags + 2, __float_as_uint(drift));
    if (any) atomicOr(flags + 4, 1u);
    if (any_in) atomicOr(flags + 5, 1u);
  }
}

// internal order -> caller order
__global__ void gather_out_kernel(int n, int width, const int *__restrict__ inv, const double *__restrict__ src,
                                  const double4 *__restrict__ xq, double *out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  int k = inv[c];
  if (xq) {
    const double4 p = xq[k];
    if (width == 1) out[c] = p.w;                        // charges
    else { out[3 * (size_t)c] = p.x; out[3 * (size_t)c + 1] = p.y; out[3 * (size_t)c + 2] = p.z; }   // positions
    return;
  }
  for (int d = 0; d < width; d++) out[(size_t)c * width + d] = src[(size_t)k * width + d];
}

}  // namespace

// slots of reduction number `seq` on the push side (my slot everywhere) and on the gather side (everyone's slot here)
static MailRed mail_red_args(const cph_handle *h, unsigned long long seq) {
  MailRed mr;
  mr.P = h->nranks;
  mr.seq = seq;
  mr.seq_index = (int)h->mail_red_cap;
  const int par = (int)(seq & 1);
  for (int p = 0; p < h->nranks; p++) {
    mr.dst[p] = (double *)((unsigned char *)h->mail_base[p] + mail_red_off(h->nranks, h->mail_red_cap, par, h->rank));
    mr.src[p] = (const double *)((const unsigned char *)h->d_mail.p + mail_red_off(h->nranks, h->mail_red_cap, par, p));
  }
  return mr;
}

bool cph_mail_red_usable(const cph_handle *h) {
  return h->nranks > 1 && h->peer_halo && h->mail_ok && (size_t)(4 + 2 * h->S + 1) <= h->mail_red_cap;
}

// The totals of the reduction pushed last, gathered into d_red by a kernel of its own (cph_site_reduce called alone).
int cph_launch_red_gather(cph_handle *h) {
  if (!h->red_pending) return 0;
  ProfScope ps(h, 5);
  const int nred = 4 + 2 * h->S + 1;
  h->nlaunch++;
  red_gather_kernel<<<nblk(nred), TPB, 0, h->stream>>>(nred, mail_red_args(h, h->seq_red), h->d_red.p, h->d_flags.p + 83);
  h->red_pending = false;
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

// push: also store this rank's block into every rank's mailbox (the caller checked cph_mail_red_usable)
int cph_launch_partition(cph_handle *h, bool push) {
  ProfScope ps(h, 2);
  const int n = h->nlocal, S = h->S;
  cudaStream_t st = h->stream;
  MailRed mr;
  if (push) {
    mr = mail_red_args(h, ++h->seq_red);
    h->red_pending = true;
  }
  CPH_CUDA(h, h->d_part.reserve((size_t)MAXPART * 4));
  const int nbP = std::max(1, std::min(MAXPART, nblk(n)));
  const int lps = h->site_lps;
  const int nbS = (S + TPB / lps - 1) / (TPB / lps);
  site_partition_kernel<<<nbP + nbS, TPB, 0, st>>>(n, h->d_eatom.p, h->d_evdwl.p, h->d_mask.p, h->fix.Hbit, nbP,
                                                  h->d_part.p, S, lps, h->d_site_start.p, h->d_titr_local.p,
                                                  h->d_titr_dq.p, h->d_phi.p,
                                                  h->lj_states ? h->d_es_g.p : nullptr,
                                                  h->extra_dudl ? h->d_extra_dudl.p : nullptr, h->fix.implicit_site, h->extra_HA,
                                                  h->extra_HB, h->have_topology ? h->d_bonded_e.p : nullptr,
                                                  h->d_red.p, h->d_flags.p + 80, mr);
  h->extra_HA = h->extra_HB = 0.0;   // consumed
  h->extra_dudl = false;
  h->nlaunch += 1;
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

// phase as in lambda_update_kernel; apply: also move the charges of the titratable atoms (and of the water
// buffer) to the new lambda in the same launch
int cph_launch_integrate(cph_handle *h, double dt, int phase, bool apply) {
  ProfScope ps(h, 3);
  const int S = h->S;
  cudaStream_t st = h->stream;
  CPH_CUDA(h, h->d_part.reserve((size_t)MAXPART * 4));
  const int lanes = S <= 8192 ? 8 : 1;     // eight lanes per site (latency) up to a few thousand sites, one beyond (throughput)
  const int nb = std::max(1, std::min(MAXPART, (S * lanes + TPB - 1) / TPB));
  const int thermo = (h->nh_tau > 0 && h->fix.integ_mode == CPH_INTEGRATE_VV && (phase == 1 || phase == 3)) ? 1 : 0;
  const double SkT = S * h->fix.boltz * h->fix.T, Q = SkT * h->nh_tau * h->nh_tau;
  if (thermo && phase == 1) {
    h->nlaunch++;
    nh_pre_kernel<<<1, 32, 0, st>>>(h->d_scal.p, dt, SkT, Q);
  }
  LambdaArgs A;
  A.S = S; A.phase = phase; A.thermo = thermo; A.lanes = lanes;
  A.apply = (apply && h->ntitr > 0 && h->have_atoms) ? 1 : 0;
  A.thermo_post = (thermo && phase == 3 && dt > 0) ? 1 : 0;
  const bool water = h->water_n > 0 && h->fix.dudl_mode == CPH_DUDL_CHARGE;
  A.nw = (water && h->have_atoms) ? h->nw_local : 0;
  A.dt = dt; A.SkT = SkT; A.Q = Q; A.inv_nw = water ? 1.0 / h->water_n : 0.0;
  A.bp = h->bias; A.fx = h->fix;
  A.pK = h->d_pK.p; A.dQ = h->d_dQ.p; A.wq = h->d_wq.p; A.titr_qA = h->d_titr_qA.p; A.titr_dq = h->d_titr_dq.p;
  A.site_start = h->d_site_start.p; A.titr_local = h->d_titr_local.p; A.wlocal = h->d_wlocal.p;
  A.red = h->d_red.p; A.lam = h->d_lam.p; A.theta = h->coord_theta ? h->d_theta.p : nullptr; A.vlam = h->d_vlam.p;
  A.alam = h->d_alam.p; A.flam = h->d_flam.p; A.fs = h->d_fs.p; A.dfs = h->d_dfs.p; A.Us = h->d_Us.p; A.dUs = h->d_dUs.p;
  A.partials = h->d_part.p; A.scal = h->d_scal.p; A.xq = h->d_xq.p; A.ticket = h->d_flags.p + 81;
  A.status = h->d_flags.p + 83;
  if (h->red_pending && phase != 1) {          // the all-reduce of the site sums completes inside this launch
    A.mr = mail_red_args(h, h->seq_red);
    h->red_pending = false;
  }
  lambda_update_kernel<<<nb, TPB, 0, st>>>(A);
  h->nlaunch += 1;
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

int cph_launch_apply_charges(cph_handle *h) {
  if (h->ntitr == 0) return 0;
  ProfScope ps(h, 4);
  h->nlaunch++;
  apply_charges_kernel<<<nblk(h->ntitr), TPB, 0, h->stream>>>(h->ntitr, h->d_titr_site.p, h->d_titr_local.p,
                                                             h->d_titr_qA.p, h->d_titr_dq.p, h->d_lam.p, h->d_xq.p);
  if (h->water_n > 0 && h->nw_local > 0) {
    h->nlaunch++;
    water_apply_kernel<<<1, TPB, 0, h->stream>>>(h->S, h->d_lam.p, h->d_dQ.p, 1.0 / h->water_n, h->nw_local,
                                                 h->d_wlocal.p, h->d_wq.p, h->d_xq.p);
  }
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

int cph_launch_water_phi(cph_handle *h) {
  if (h->water_n <= 0) return 0;
  h->nlaunch++;
  water_phi_kernel<<<1, 32, 0, h->stream>>>(h->nw_local, h->d_wlocal.p, h->d_phi.p, h->d_red.p + 4 + 2 * (size_t)h->S);
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

int cph_launch_water_dudl(cph_handle *h) {
  if (h->water_n <= 0) return 0;
  h->nlaunch++;
  water_dudl_kernel<<<nblk(h->S), TPB, 0, h->stream>>>(h->S, h->d_dQ.p, 1.0 / h->water_n, h->d_red.p);
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

int cph_launch_set_force(cph_handle *h) {
  if (h->nh == 0) return 0;
  ProfScope ps(h, 4);
  h->nlaunch++;
  set_force_kernel<<<nblk(h->nh), TPB, 0, h->stream>>>(h->nh, h->d_hlist.p, h->d_site_of.p, h->d_lam.p,
                                                      h->fix.fscale_mode, h->d_f.p);
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

int cph_launch_set_x(cph_handle *h, const double *xc) {
  ProfScope ps(h, 7);
  const int n = h->nlocal;
  cudaStream_t st = h->stream;
  // words 0 (max displacement), 4 (re-neighbour) and 5 (prune) are this step's; 1..3 belong to the list
  // build, which clears them itself before use, so one memset covers the lot
  CPH_CUDA(h, cudaMemsetAsync(h->d_flags.p, 0, 6 * sizeof(unsigned int), st));
  const double thresh2 = 0.25 * h->skin * h->skin;
  const double thresh2_in = 0.25 * h->inner_skin * h->inner_skin;
  const double *xin = h->inner_valid ? h->d_xinner.p : nullptr;
  if (n) {
    h->nlaunch++;
    const double3 slo = make_double3(h->sublo[0], h->sublo[1], h->sublo[2]);
    const double3 shi = make_double3(h->subhi[0], h->subhi[1], h->subhi[2]);
    if (xc) set_x_kernel<<<nblk(n), TPB, 0, st>>>(n, xc, h->d_perm.p, h->d_xbuild.p, thresh2, xin, thresh2_in, slo, shi,
                                                 h->d_xq.p, h->d_flags.p);
    else check_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_xbuild.p, thresh2, xin, thresh2_in, slo, shi, h->d_xq.p,
                                               h->d_flags.p);
  }
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

int cph_launch_gather_out(cph_handle *h, int what, double *out, cudaStream_t on) {
  const int n = h->nlocal;
  if (n == 0) return 0;
  cudaStream_t st = on ? on : h->stream;
  const double *src = what == 0 ? h->d_f.p : what == 1 ? h->d_eatom.p : what == 5 ? (const double *)h->d_v.p : h->d_phi.p;
  const bool three = what == 0 || what == 4 || what == 5;
  h->nlaunch++;
  gather_out_kernel<<<nblk(n), TPB, 0, st>>>(n, three ? 3 : 1, h->d_inv.p, src,
                                                   (what == 3 || what == 4) ? h->d_xq.p : nullptr, out);
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

namespace {
__global__ void pack_xq_kernel(int n, const double *__restrict__ x, const double *__restrict__ q, double4 *xq) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) xq[k] = make_double4(x[3 * (size_t)k], x[3 * (size_t)k + 1], x[3 * (size_t)k + 2], q[k]);
}
}  // namespace

// device-resident caller arrays -> packed {x,y,z,q} (cph_set_atoms with CPH_DEVICE)
int cph_launch_pack_xq(cph_handle *h, int n, const double *x, const double *q) {
  if (n == 0) return 0;
  h->nlaunch++;
  pack_xq_kernel<<<nblk(n), TPB, 0, h->stream>>>(n, x, q, h->d_xq.p);
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}
