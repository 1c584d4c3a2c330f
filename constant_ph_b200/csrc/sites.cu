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

// compute_Hs tail (cpp:259-267) over the owned atoms + energy totals
__global__ void __launch_bounds__(TPB)
partition_kernel(int n, const double *__restrict__ eatom, const double *__restrict__ evdwl,
                 const int *__restrict__ mask, int Hbit, double *partials) {
  double v[4] = {0, 0, 0, 0};   // HA, HB, E_vdwl, E_coul
  for (int k = blockIdx.x * TPB + threadIdx.x; k < n; k += gridDim.x * TPB) {
    double e = eatom[k], ev = evdwl[k];
    v[0] += e;                                  // cpp:265
    if (!(mask[k] & Hbit)) v[1] += e;           // cpp:266
    v[2] += ev;
    v[3] += e - ev;
  }
  block_reduce_store<4>(v, partials);
}

__global__ void partition_final_kernel(int nb, const double *__restrict__ partials, double *red, int implicit_site,
                                       int S, double extra_HA, double extra_HB, const double *__restrict__ bonded_e) {
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
  __shared__ double out[4];
  if (c < 4) {
    double s = 0;
    for (int b = lane; b < nb; b += 32) s += partials[(size_t)b * 4 + c];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    // extra_HA/HB: this rank's share of the per-atom energies LAMMPS tallied on the host (bonded styles,
    // KSpace: cpp:221-244), already partitioned by the fix exactly as cpp:264-267 does
    if (lane == 0) {
      if (c == 0) s += extra_HA;
      if (c == 1) s += extra_HB;
      // E_coul is summed as eatom - evdwl; with bonded terms on the device (f2) eatom also holds their shares
      if (c == 3 && bonded_e) s -= bonded_e[0] + bonded_e[1];
      out[c] = s;
      red[c] = s;
    }
  }
  __syncthreads();
  // reference single site: HB - HA of the whole hydrogen group (cpp:111)
  if (implicit_site && threadIdx.x == 0) red[4 + S] = out[1] - out[0];
}

// warp-shuffle segmented reduction over the site-major titratable-atom table
__global__ void __launch_bounds__(TPB)
site_sum_kernel(int ntitr, const int *__restrict__ titr_site, const int *__restrict__ titr_local,
                const double *__restrict__ titr_dq, const double *__restrict__ phi, const double *__restrict__ eatom,
                const int *__restrict__ mask, int Hbit, int S, int implicit_site, double *red) {
  const int t = blockIdx.x * TPB + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int site = -1;
  double d = 0, hd = 0;
  if (t < ntitr) {
    site = titr_site[t];
    int k = titr_local[t];
    if (k >= 0) {                                         // owned by this rank (cpp:264: i < nlocal)
      d = titr_dq[t] * phi[k];                            // Appendix B
      if (mask[k] & Hbit) hd = -eatom[k];                 // HB_s - HA_s
    }
  }
  for (int o = 1; o < 32; o <<= 1) {
    double ud = __shfl_up_sync(0xffffffffu, d, o), uh = __shfl_up_sync(0xffffffffu, hd, o);
    int us = __shfl_up_sync(0xffffffffu, site, o);
    if (lane >= o && us == site) { d += ud; hd += uh; }
  }
  int next = __shfl_down_sync(0xffffffffu, site, 1);
  bool tail = (lane == 31) || (next != site);
  if (site >= 0 && tail) {
    atomicAdd(red + 4 + site, d);
    if (!implicit_site) atomicAdd(red + 4 + S + site, hd);
  }
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

// phase 0: reference kinematic step (cpp:109-117)   phase 1: VV kick+drift
// phase 2: VV force evaluation (a <- F/m)           phase 3: VV second kick + H_lambda
__global__ void __launch_bounds__(TPB)
integrate_kernel(int S, double dt, int phase, BiasParams bp, FixParams fx, const double *__restrict__ pK,
                 const double *__restrict__ red, double *lam, double *theta, double *vlam, double *alam, double *flam,
                 double *fs, double *dfs, double *Us, double *dUs, double *partials, const double *__restrict__ scal,
                 int thermo) {
  // thermo: Nose-Hoover on the site velocities (velocity-Verlet form only); scal[9] = exp(-xi dt/2)
  const double nh = thermo ? scal[9] : 1.0;
  // theta != NULL: the dynamical coordinate is theta with lambda = sin^2(theta) (north_star's lambda/theta
  // variables; absent from the reference, which integrates lambda itself and confines it with U4/U5);
  // velocity, acceleration and mass then refer to theta and F_theta = F_lambda * sin(2 theta).
  double v[3] = {0, 0, 0};   // sum of site terms of H_lambda, sum lambda*(HB_s-HA_s), kinetic
  for (int s = blockIdx.x * TPB + threadIdx.x; s < S; s += gridDim.x * TPB) {
    double cq = theta ? theta[s] : lam[s];
    double vel = vlam[s], acc = alam[s];
    if (phase == 1) {
      vel = vel * nh + 0.5 * acc * dt;
      cq += vel * dt;
      if (theta) { theta[s] = cq; const double sn = sin(cq); lam[s] = sn * sn; }
      else lam[s] = cq;
      vlam[s] = vel;
      continue;
    }
    double lambda = cq, chain = 1.0;
    if (theta) { const double sn = sin(cq); lambda = sn * sn; chain = sin(2.0 * cq); }
    if (phase == 3) vel = (vel + 0.5 * acc * dt) * nh;
    BiasOut b = bias_terms(bp, lambda);
    const double pk = fx.implicit_site ? fx.pK : pK[s];
    const double hd = red[4 + S + s];
    const double dE = (fx.dudl_mode == CPH_DUDL_REFERENCE) ? hd : red[4 + s];
    const double ph = fx.boltz * fx.T * log(10.0) * (pk - fx.pH);
    const double f_lambda = -(dE + b.df * ph + b.dU);                 // cpp:111
    const double a_lambda = f_lambda * chain / bp.m_lambda * fx.ftm2v;   // cpp:112 (+ SURVEY D9)
    const double kin = 0.5 * bp.m_lambda * vel * vel / fx.ftm2v;
    v[0] += b.f * ph + b.U + kin;                                     // cpp:114 site terms
    v[1] += lambda * hd;                                              // cpp:114 lambda*(HB-HA)
    v[2] += kin;
    fs[s] = b.f; dfs[s] = b.df; Us[s] = b.U; dUs[s] = b.dU; flam[s] = f_lambda;
    if (phase == 0) {
      cq = 0.5 * a_lambda * dt * dt + vel * dt + cq;                  // cpp:115
      vel = a_lambda * dt + vel;                                      // cpp:116
    }
    if (theta) { theta[s] = cq; const double sn = sin(cq); lam[s] = sn * sn; }
    else lam[s] = cq;
    vlam[s] = vel;
    alam[s] = a_lambda;
  }
  if (phase != 1) block_reduce_store<3>(v, partials);
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

__global__ void integrate_final_kernel(int nb, const double *__restrict__ partials, const double *__restrict__ red,
                                       int dudl_mode, double *scal, int thermo_post, double dt, double SkT, double Q) {
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
  __shared__ double out[3];
  if (c < 3) {
    double s = 0;
    for (int b = lane; b < nb; b += 32) s += partials[(size_t)b * 3 + c];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[c] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // cpp:114: (1-lambda)HA + lambda HB = HA + lambda (HB-HA); charge mode: E_ff at the current charges
    double eff = (dudl_mode == CPH_DUDL_REFERENCE) ? red[0] + out[1] : red[2] + red[3];
    scal[4] = eff + out[0];
    scal[5] = out[2];
    if (thermo_post) {   // second Nose-Hoover half step, with the kinetic energy after the scaled final kick
      scal[10] += 0.5 * dt * scal[8];
      scal[8] += 0.5 * dt * (2.0 * out[2] - SkT) / Q;
      scal[11] = 0.5 * Q * scal[8] * scal[8] + SkT * scal[10];    // thermostat energy (conserved with H_lambda)
    }
  }
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
                             double thresh2_inner, double4 *xq, unsigned int *flags) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  float d2f = 0.f;
  bool over = false, over_in = false;
  if (k < n) {
    size_t c = (size_t)perm[k] * 3;
    double x = xc[c], y = xc[c + 1], z = xc[c + 2];
    xq[k].x = x; xq[k].y = y; xq[k].z = z;
    double dx = x - xbuild[3 * (size_t)k], dy = y - xbuild[3 * (size_t)k + 1], dz = z - xbuild[3 * (size_t)k + 2];
    double d2 = dx * dx + dy * dy + dz * dz;
    over = d2 > thresh2;
    d2f = __double2float_ru(d2);
    if (xinner) {   // the pruned inner rows stay valid while nobody moved more than inner_skin/2
      dx = x - xinner[3 * (size_t)k]; dy = y - xinner[3 * (size_t)k + 1]; dz = z - xinner[3 * (size_t)k + 2];
      over_in = dx * dx + dy * dy + dz * dz > thresh2_inner;
    }
  }
  for (int o = 16; o; o >>= 1) d2f = fmaxf(d2f, __shfl_xor_sync(0xffffffffu, d2f, o));
  unsigned int any = __ballot_sync(0xffffffffu, over);
  unsigned int any_in = __ballot_sync(0xffffffffu, over_in);
  if ((threadIdx.x & 31) == 0) {
    if (d2f > 0.f) atomicMax(flags + 0, __float_as_uint(d2f));
    if (any) atomicOr(flags + 4, 1u);
    if (any_in) atomicOr(flags + 5, 1u);
  }
}

__global__ void check_kernel(int n, const double *__restrict__ xbuild, double thresh2, const double *__restrict__ xinner,
                             double thresh2_inner, const double4 *__restrict__ xq, unsigned int *flags) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  float d2f = 0.f;
  bool over = false, over_in = false;
  if (k < n) {
    double4 p = xq[k];
    double dx = p.x - xbuild[3 * (size_t)k], dy = p.y - xbuild[3 * (size_t)k + 1], dz = p.z - xbuild[3 * (size_t)k + 2];
    double d2 = dx * dx + dy * dy + dz * dz;
    over = d2 > thresh2;
    d2f = __double2float_ru(d2);
    if (xinner) {
      dx = p.x - xinner[3 * (size_t)k]; dy = p.y - xinner[3 * (size_t)k + 1]; dz = p.z - xinner[3 * (size_t)k + 2];
      over_in = dx * dx + dy * dy + dz * dz > thresh2_inner;
    }
  }
  for (int o = 16; o; o >>= 1) d2f = fmaxf(d2f, __shfl_xor_sync(0xffffffffu, d2f, o));
  unsigned int any = __ballot_sync(0xffffffffu, over);
  unsigned int any_in = __ballot_sync(0xffffffffu, over_in);
  if ((threadIdx.x & 31) == 0) {
    if (d2f > 0.f) atomicMax(flags + 0, __float_as_uint(d2f));
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

int cph_launch_partition(cph_handle *h) {
  ProfScope ps(h, 2);
  const int n = h->nlocal, S = h->S;
  cudaStream_t st = h->stream;
  CPH_CUDA(h, h->d_part.reserve((size_t)MAXPART * 4));
  CPH_CUDA(h, cudaMemsetAsync(h->d_red.p, 0, (4 + 2 * (size_t)S + 1) * sizeof(double), st));
  int nb = std::max(1, std::min(MAXPART, nblk(n)));
  partition_kernel<<<nb, TPB, 0, st>>>(n, h->d_eatom.p, h->d_evdwl.p, h->d_mask.p, h->fix.Hbit, h->d_part.p);
  partition_final_kernel<<<1, 128, 0, st>>>(nb, h->d_part.p, h->d_red.p, h->fix.implicit_site, S, h->extra_HA,
                                            h->extra_HB, h->have_topology ? h->d_bonded_e.p : nullptr);
  h->extra_HA = h->extra_HB = 0.0;   // consumed
  if (h->ntitr)
    site_sum_kernel<<<nblk(h->ntitr), TPB, 0, st>>>(h->ntitr, h->d_titr_site.p, h->d_titr_local.p, h->d_titr_dq.p,
                                                    h->d_phi.p, h->d_eatom.p, h->d_mask.p, h->fix.Hbit, S,
                                                    h->fix.implicit_site, h->d_red.p);
  h->nlaunch += h->ntitr ? 3 : 2;
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

int cph_launch_integrate(cph_handle *h, double dt, int phase) {
  ProfScope ps(h, 3);
  const int S = h->S;
  cudaStream_t st = h->stream;
  CPH_CUDA(h, h->d_part.reserve((size_t)MAXPART * 4));
  int nb = std::max(1, std::min(MAXPART, nblk(S)));
  const int thermo = (h->nh_tau > 0 && h->fix.integ_mode == CPH_INTEGRATE_VV && (phase == 1 || phase == 3)) ? 1 : 0;
  const double SkT = S * h->fix.boltz * h->fix.T, Q = SkT * h->nh_tau * h->nh_tau;
  if (thermo && phase == 1) {
    h->nlaunch++;
    nh_pre_kernel<<<1, 32, 0, st>>>(h->d_scal.p, dt, SkT, Q);
  }
  integrate_kernel<<<nb, TPB, 0, st>>>(S, dt, phase, h->bias, h->fix, h->d_pK.p, h->d_red.p, h->d_lam.p,
                                       h->coord_theta ? h->d_theta.p : nullptr, h->d_vlam.p,
                                       h->d_alam.p, h->d_flam.p, h->d_fs.p, h->d_dfs.p, h->d_Us.p, h->d_dUs.p,
                                       h->d_part.p, h->d_scal.p, thermo);
  if (phase != 1)
    integrate_final_kernel<<<1, 96, 0, st>>>(nb, h->d_part.p, h->d_red.p, h->fix.dudl_mode, h->d_scal.p,
                                             thermo && phase == 3 && dt > 0, dt, SkT, Q);
  h->nlaunch += phase != 1 ? 2 : 1;
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
    if (xc) set_x_kernel<<<nblk(n), TPB, 0, st>>>(n, xc, h->d_perm.p, h->d_xbuild.p, thresh2, xin, thresh2_in, h->d_xq.p,
                                                 h->d_flags.p);
    else check_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_xbuild.p, thresh2, xin, thresh2_in, h->d_xq.p, h->d_flags.p);
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
