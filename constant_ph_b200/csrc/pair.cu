// pair.cu -- K2: fused fp64 pair pass for lj/cut/coul/cut and lj/cut/coul/dsf.
//
// Stands in for the LAMMPS pair style whose per-atom energies the reference reads at
// fix_constant_pH.cpp:216-219 (`force->pair->eatom`), restated from SURVEY.md Appendix A,
// and adds what north_star asks for in the same pass: forces, the electrostatic
// potential phi_i = dE_coul/dq_i that gives every site's dU/dlambda analytically
// (Appendix B), and the van-der-Waals share of the per-atom energy.
//
// Mapping: one warp per owned atom, APW atoms per warp in sequence.
//   1. The atom's neighbour row streams HBM -> shared memory through a per-warp ring of
//      512-byte tiles filled by TMA bulk copies (cp.async.bulk + mbarrier), issued several
//      tiles ahead so the index stream's DRAM latency is never on the critical path.  Rows
//      are padded to whole tiles with a far-away dummy atom, so there is no tail logic.
//   2. Filter: each lane takes 4 candidates of a tile, gathers their packed fp32
//      {x,y,z,type} (16 B; half the traffic of the fp64 record) and applies a CONSERVATIVE
//      fp32 cutoff test (cutoff^2 + margin).  The Verlet skin makes ~40 % of candidates fail.
//   3. Survivors are ballot-compacted into a per-warp shared-memory queue; whenever 32 are
//      queued the warp evaluates them with every lane active: one 256-bit load of the fp64
//      {x,y,z,q}, the exact fp64 cutoff test, rsqrt / exp / polynomial erfc in fp64.
//      Special-bond pairs (2 of ~420 per water atom) take a separate slow path so the hot
//      evaluation carries no exclusion arithmetic.
//   4. Accumulators are reduced across the warp with shuffles; a full list means no
//      atomics and a fixed summation order (bit reproducible run to run).
// The fp32 test only prunes; every pair inside the cutoff is decided and evaluated in fp64,
// so results are identical to an all-fp64 filter.
//
// The kernel is issue-bound, not HBM-bound: all polynomial constants live in
// __constant__ memory so DFMA takes them as c[bank][offset] operands instead of two UMOVs
// each, exp() is a 32-entry-table + degree-6 polynomial without special cases (its argument
// is in [-alpha^2 rc^2, 0]), 1/sqrt and 1/x are the hardware seed plus one third-order step.
//
// Per-atom outputs: f (3), evdwl_i = 1/2 sum_j evdwl_ij, phi_i, eatom_i = evdwl_i +
// 1/2 q_i phi_i  (== ev_tally's half-half split plus the dsf self term).
#include "cph_internal.h"

namespace {

constexpr int WARPS = 8;
constexpr int TPB = WARPS * 32;
constexpr int APW = 4;            // atoms per warp (sequential)
constexpr int APB = WARPS * APW;  // atoms per block
constexpr int CH = 128;           // candidates per tile (4 per lane)
constexpr int NBUF = 4;           // tiles in flight per warp
constexpr int QCAP = 128;         // per-warp compaction queue (ring): <= 63 left over + 64 pushed
constexpr int MAXTILES = 64;      // per-warp tile schedule entries (rows of up to 2048 neighbours)

// constants of the hot evaluation, addressed as c[3][..] operands
struct EvalConst {
  double ewp_alpha;        // EWALD_P * alpha
  double a1, a2, a3, a4, a5;
  double neg_alpha2;       // -alpha^2
  double two_alpha_pis;    // 2 alpha / sqrt(pi)
  double qqrd2e, e_shift, f_shift, cut_coulsq, cutsq_max;
  double exp_scale;        // 32 / ln 2
  double exp_magic;        // 2^52 + 2^51
  double exp_c1;           // ln2/32
  double p2, p3, p4, p5, p6;
  double one_m_fc[4], flj[4], fcoul[4];   // special-bond factors (slow path only)
};
__constant__ EvalConst kc;
__constant__ double kexp2[32];   // 2^(j/32)

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
// exp(x) for x in [-700, 0]: x = (32k + j) ln2/32 + r, |r| <= ln2/64; exp = 2^k * 2^(j/32) * P6(r).
// Relative error < 2e-16 (P6 truncation 4e-18, table correctly rounded).
__device__ __forceinline__ double fast_exp_neg(double x, const double *s_exp2) {
  const double tm = fma(x, kc.exp_scale, kc.exp_magic);
  const int ni = __double2loint(tm);
  const double nd = tm - kc.exp_magic;
  const double r = fma(nd, -kc.exp_c1, x);   // |nd| < 2^15: representation error of c1 adds < 1e-13 relative
  double p = fma(r, kc.p6, kc.p5);
  p = fma(r, p, kc.p4);
  p = fma(r, p, kc.p3);
  p = fma(r, p, kc.p2);
  p = fma(r, p, 1.0);
  p = fma(r, p, 1.0);
  const double v = s_exp2[ni & 31] * p;
  return __hiloint2double(__double2hiint(v) + ((ni >> 5) << 20), __double2loint(v));
}

// the fp64 record {x,y,z,q} of one atom.  CPH_LD256=1: one 32-byte request per lane
// (LDG.E.ENL2.256 on sm_100a; measured to be served from L2, L1 hit rate 27 %);
// CPH_LD256=0: two 16-byte read-only loads that allocate in L1.
#ifndef CPH_LD256
#define CPH_LD256 1
#endif
__device__ __forceinline__ double4 ld256(const double4 *p) {
  double4 v;
#if CPH_LD256
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
#else
  const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
  const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
  v = make_double4(a.x, a.y, b.x, b.y);
#endif
  return v;
}

// ---- TMA bulk copy + mbarrier (per-warp ring) -------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned int bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned int bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned int dst, const void *src, unsigned int bytes, unsigned int bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int bar, unsigned int parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

struct Acc {
  double fx = 0, fy = 0, fz = 0, ev = 0, phi = 0;
};

// Hot evaluation of one pair without special-bond factors, as STRAIGHT-LINE code (selects, no
// branches) so that two independent pairs per lane interleave in the fp64 pipe.  Returns
// fpair (force / r), the vdW energy and qj*K (the pair's contribution to phi_i).
// UNI: every type pair has cut_lj == cut_coul == the global cutoff (one exact test).
template <int STYLE, int EFLAG, int UNI>
__device__ __forceinline__ void eval_pair(const double4 *s_coef, const double2 *s_cut, int tt, double rsq,
                                          double qi, double qj, const double *s_exp2, bool LJ, double &fpair_out,
                                          double &ev_out, double &phi_out) {
  bool in, lj_on = true, coul_on = true;
  if (UNI) {
    in = rsq < kc.cutsq_max;           // the exact (fp64) cutoff decision
  } else {
    const double2 cc = s_cut[tt];      // {cut_ljsq, cutsq}
    in = rsq < cc.y;
    lj_on = rsq < cc.x;
    coul_on = rsq < kc.cut_coulsq;
  }
  const double rinv = fast_rsqrt(rsq);
  const double r2inv = rinv * rinv;
  double fpair = 0.0, ev = 0.0, ph = 0.0;
  if (LJ) {
    const double4 c = s_coef[tt];      // {12*lj3, 6*lj4, lj3, lj4}
    const double r6inv = r2inv * r2inv * r2inv;
    fpair = r6inv * fma(c.x, r6inv, -c.y) * r2inv;
    if (EFLAG) ev = r6inv * fma(c.z, r6inv, -c.w);
    if (!UNI) {
      fpair = lj_on ? fpair : 0.0;
      ev = lj_on ? ev : 0.0;
    }
  }
  double fcoul;
  if (STYLE == CPH_PAIR_LJ_CUT_COUL_CUT) {
    const double k = kc.qqrd2e * rinv;   // E_ij = qi qj k
    fcoul = qi * qj * k * r2inv;
    if (EFLAG) ph = qj * k;
  } else {
    const double r = rsq * rinv;
    const double erfcd = fast_exp_neg(kc.neg_alpha2 * rsq, s_exp2);
    const double t = fast_rcp(fma(kc.ewp_alpha, r, 1.0));
    double poly = fma(t, kc.a5, kc.a4);
    poly = fma(t, poly, kc.a3);
    poly = fma(t, poly, kc.a2);
    poly = fma(t, poly, kc.a1);
    const double erfcc = t * poly * erfcd;
    const double u = (kc.qqrd2e * rinv) * qj;          // prefactor / qi, shared by force and potential
    // forcecoul*r2inv = prefactor*(erfcc/r + 2a/sqrt(pi)*erfcd + r*f_shift)*r * r2inv
    const double fc = fma(erfcc, rinv, fma(kc.two_alpha_pis, erfcd, r * kc.f_shift));
    fcoul = (qi * u) * (fc * rinv);
    if (EFLAG) ph = u * fma(-rsq, kc.f_shift, fma(-r, kc.e_shift, erfcc));
  }
  if (!UNI) {
    fcoul = coul_on ? fcoul : 0.0;
    ph = coul_on ? ph : 0.0;
  }
  fpair += fcoul;
  fpair_out = in ? fpair : 0.0;
  ev_out = in ? ev : 0.0;
  phi_out = in ? ph : 0.0;
}

// Slow path for special-bond pairs (factor_lj / factor_coul and, under
// dsf, the -(1-factor_coul)*prefactor correction).  A handful of lanes per atom.
template <int STYLE, int EFLAG>
__device__ __noinline__ void eval_special(const double4 *s_coef, const double2 *s_cut, int tt, double delx,
                                          double dely, double delz, double rsq, double qi, double qj, int sb,
                                          double *out5) {
  const double2 cc = s_cut[tt];
  out5[0] = out5[1] = out5[2] = out5[3] = out5[4] = 0.0;
  if (rsq >= cc.y) return;
  const double factor_lj = kc.flj[sb], factor_coul = kc.fcoul[sb];
  const double rinv = 1.0 / sqrt(rsq);
  const double r2inv = rinv * rinv;
  double fpair = 0.0, ev = 0.0, ph = 0.0;
  if (rsq < cc.x) {
    const double4 c = s_coef[tt];
    const double r6inv = r2inv * r2inv * r2inv;
    fpair = factor_lj * r6inv * (c.x * r6inv - c.y) * r2inv;
    ev = factor_lj * (r6inv * (c.z * r6inv - c.w));
  }
  if (rsq < kc.cut_coulsq) {
    if (STYLE == CPH_PAIR_LJ_CUT_COUL_CUT) {
      const double k = kc.qqrd2e * factor_coul * rinv;
      fpair += qi * qj * k * r2inv;
      ph = qj * k;
    } else {
      const double r = rsq * rinv;
      const double erfcd = exp(kc.neg_alpha2 * rsq);
      const double t = 1.0 / (1.0 + kc.ewp_alpha * r);
      const double erfcc = t * (kc.a1 + t * (kc.a2 + t * (kc.a3 + t * (kc.a4 + t * kc.a5)))) * erfcd;
      const double pre = kc.qqrd2e * rinv;
      double fc = erfcc * rinv + kc.two_alpha_pis * erfcd + r * kc.f_shift;
      double kk = erfcc - r * kc.e_shift - rsq * kc.f_shift;
      fc -= kc.one_m_fc[sb] * rinv;
      kk -= kc.one_m_fc[sb];
      fpair += qi * qj * pre * fc * rinv;
      ph = qj * pre * kk;
    }
  }
  out5[0] = delx * fpair; out5[1] = dely * fpair; out5[2] = delz * fpair; out5[3] = ev; out5[4] = ph;
}

struct WarpSmem {
  int tile[NBUF][CH];                 // neighbour tiles (TMA destination), 16-byte aligned
  int2 queue[QCAP];                   // {neighbour index, type pair index}
  unsigned long long bar[NBUF];
  int sched[MAXTILES];                // tile schedule: tile index (row offset / CH) of each tile of this warp
};

#ifndef CPH_PAIR_MINBLOCKS
#define CPH_PAIR_MINBLOCKS 3
#endif

template <int STYLE, int EFLAG, int UNI>
__global__ void __launch_bounds__(TPB, CPH_PAIR_MINBLOCKS)
pair_fused_kernel(int nlocal, const double4 *__restrict__ xq, const float4 *__restrict__ xt,
            const int *__restrict__ neigh, const int *__restrict__ numneigh, const int *__restrict__ numspec,
            int rowcap, int nt1, float cutf, const double4 *__restrict__ coef, const double2 *__restrict__ cuts,
            const int *__restrict__ type_has_lj, double *__restrict__ f, double *__restrict__ evdwl,
            double *__restrict__ phi, double *__restrict__ eatom, double c_self) {
  __shared__ double4 s_coef[CPH_MAXNT1 * CPH_MAXNT1];
  __shared__ double2 s_cut[CPH_MAXNT1 * CPH_MAXNT1];
  __shared__ double s_exp2[32];
  __shared__ __align__(128) WarpSmem s_w[WARPS];
  for (int k = threadIdx.x; k < nt1 * nt1; k += TPB) {
    s_coef[k] = coef[k];
    s_cut[k] = cuts[k];
  }
  if (threadIdx.x < 32) s_exp2[threadIdx.x] = kexp2[threadIdx.x];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned int ltmask = (1u << lane) - 1;
  WarpSmem &sm = s_w[w];
  const unsigned int bar0 = smem_u32(&sm.bar[0]);
  const unsigned int tile0 = smem_u32(&sm.tile[0][0]);
  if (lane == 0)
    for (int b = 0; b < NBUF; b++) mbar_init(bar0 + 8 * b, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");

  // this warp's atoms: base + n*WARPS, n = 0..APW-1 (the block's warps walk adjacent atoms together)
  const int base = blockIdx.x * APB + w;
  const int my_atom = base + lane * WARPS;
  const bool mine = lane < APW && my_atom < nlocal;
  const int nn_mine = mine ? numneigh[my_atom] : 0;
  const int nsp_mine = mine ? numspec[my_atom] : 0;
  const int nt_mine = (nn_mine + CH - 1) / CH;            // tiles of "my" atom
  // exclusive scan over the APW atoms -> tile schedule in shared memory
  int pre = nt_mine;
  for (int o = 1; o < APW; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, pre, o);
    if (lane >= o) pre += v;
  }
  const int total_tiles = min(__shfl_sync(0xffffffffu, pre, APW - 1), MAXTILES);
  if (lane < APW) {
    int t0 = pre - nt_mine;
    for (int c = 0; c < nt_mine && t0 + c < MAXTILES; c++)
      sm.sched[t0 + c] = (base + lane * WARPS) * (rowcap / CH) + c;
  }
  __syncthreads();

  int issued = 0;
  auto issue = [&]() {
    if (issued < total_tiles && lane == 0) {
      const unsigned int slot = issued & (NBUF - 1);
      mbar_expect_tx(bar0 + 8 * slot, CH * 4);
      bulk_g2s(tile0 + slot * (CH * 4), neigh + (size_t)sm.sched[issued] * CH, CH * 4, bar0 + 8 * slot);
    }
    issued++;
  };
#pragma unroll
  for (int b = 0; b < NBUF; b++) issue();

  int cslot = 0;
  for (int n = 0; n < APW; n++) {
    const int i = base + n * WARPS;
    if (i >= nlocal) break;
    const int ntile = __shfl_sync(0xffffffffu, nt_mine, n);
    const int nsp = __shfl_sync(0xffffffffu, nsp_mine, n);
    const double4 pi = xq[i];
    const float4 pti = xt[i];
    const int ti = __float_as_int(pti.w);
    const int tbase = ti * nt1;
    const bool has_lj = type_has_lj[ti] != 0;       // warp-uniform: water H (2/3 of atoms) skips all LJ work
    Acc a;
    int head = 0, tail = 0;   // queue ring indices (warp-uniform)

    // special-bond partners sit at the end of the row; a handful of lanes take the slow path
    if (nsp > 0) {
      if (lane < nsp) {
        const int raw = neigh[(size_t)i * rowcap + (rowcap - 1 - lane)];
        const int j = raw & CPH_NEIGHMASK, sb = (raw >> CPH_SBSHIFT) & 3;
        const double4 pq = ld256(xq + j);
        const int tt = tbase + __float_as_int(xt[j].w);
        const double delx = pi.x - pq.x, dely = pi.y - pq.y, delz = pi.z - pq.z;
        const double rsq = fma(delz, delz, fma(dely, dely, delx * delx));
        double o5[5];
        eval_special<STYLE, EFLAG>(s_coef, s_cut, tt, delx, dely, delz, rsq, pi.w, pq.w, sb, o5);
        a.fx = o5[0]; a.fy = o5[1]; a.fz = o5[2];
        if (EFLAG) { a.ev = o5[3]; a.phi = o5[4]; }
      }
      __syncwarp();
    }

    // evaluate `count` (<= 32) queued pairs, one per lane
    auto drain = [&](int count) {
      if (lane < count) {
        const int2 e = sm.queue[(head + lane) & (QCAP - 1)];
        const double4 pj = ld256(xq + e.x);
        const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
        const double rsq = fma(delz, delz, fma(dely, dely, delx * delx));
        double fp, ev, ph;
        // has_lj is warp-uniform: one copy of the evaluation, the LJ block behind a uniform branch
        // (two template copies per call site pushed the kernel body past the 32 KB L1.5 I-cache)
        eval_pair<STYLE, EFLAG, UNI>(s_coef, s_cut, e.y, rsq, pi.w, pj.w, s_exp2, has_lj, fp, ev, ph);
        a.fx = fma(delx, fp, a.fx); a.fy = fma(dely, fp, a.fy); a.fz = fma(delz, fp, a.fz);
        if (EFLAG) { a.ev += ev; a.phi += ph; }
      }
      head += count;
    };
    // two candidates per lane: fp32 test, ballot-compact into the queue
    auto push2 = [&](int ra, const float4 &pa, int rb, const float4 &pb) {
      const float dxa = pti.x - pa.x, dya = pti.y - pa.y, dza = pti.z - pa.z;
      const float dxb = pti.x - pb.x, dyb = pti.y - pb.y, dzb = pti.z - pb.z;
      const bool ina = fmaf(dxa, dxa, fmaf(dya, dya, dza * dza)) < cutf;
      const bool inb = fmaf(dxb, dxb, fmaf(dyb, dyb, dzb * dzb)) < cutf;
      const unsigned int ma = __ballot_sync(0xffffffffu, ina);
      const unsigned int mb = __ballot_sync(0xffffffffu, inb);
      const int ca = __popc(ma);
      if (ina) sm.queue[(tail + __popc(ma & ltmask)) & (QCAP - 1)] = make_int2(ra, tbase + __float_as_int(pa.w));
      if (inb) sm.queue[(tail + ca + __popc(mb & ltmask)) & (QCAP - 1)] = make_int2(rb, tbase + __float_as_int(pb.w));
      tail += ca + __popc(mb);
      __syncwarp();
      while (tail - head >= 32) drain(32);
    };

    for (int t = 0; t < ntile; t++) {
      const unsigned int slot = cslot & (NBUF - 1);
      mbar_wait(bar0 + 8 * slot, (cslot / NBUF) & 1);
      const int *tp = &sm.tile[slot][lane];
      const int r0 = tp[0], r1 = tp[32], r2 = tp[64], r3 = tp[96];
      __syncwarp();
      cslot++;
      issue();                      // refill the slot just consumed
      const float4 p0 = xt[r0], p1 = xt[r1], p2 = xt[r2], p3 = xt[r3];
      push2(r0, p0, r1, p1);
      push2(r2, p2, r3, p3);
    }
    if (tail - head >= 32) drain(32);
    if (tail - head > 0) drain(tail - head);
    __syncwarp();

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
        const double ev = 0.5 * a.ev;
        const double ph = a.phi + 2.0 * pi.w * c_self;   // dE_coul/dq_i including the dsf self term
        evdwl[i] = ev;
        phi[i] = ph;
        eatom[i] = ev + 0.5 * pi.w * ph;
      }
    }
  }
}


// ============================================================================================
// Two-level list ("rolling prune"): K2a prunes each Verlet row (rc + skin, rebuilt from cells
// every ~20 steps) down to an inner row (rc + inner skin) every few steps, in fp32, without any
// fp64 math; K2b evaluates the inner rows every step with all lanes busy.  The inner row is a
// conservative superset of the pairs in range while no atom has moved more than inner_skin/2
// since the prune (tracked next to neighbor->decide()); K2b still decides every pair in fp64.
// ============================================================================================

// K2a: one warp per atom; TMA tile ring over the outer row; survivors written to the inner row
// as (j | type_j << 28), padded to a multiple of 32 with the dummy atom.
__global__ void __launch_bounds__(TPB, 4)
prune_kernel(int nlocal, const float4 *__restrict__ xt, const int *__restrict__ neigh,
             const int *__restrict__ numneigh, int rowcap, float cutf_inner, int dummy, int *__restrict__ neigh2,
             int *__restrict__ numneigh2) {
  __shared__ __align__(128) WarpSmem s_w[WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned int ltmask = (1u << lane) - 1;
  WarpSmem &sm = s_w[w];
  const unsigned int bar0 = smem_u32(&sm.bar[0]);
  const unsigned int tile0 = smem_u32(&sm.tile[0][0]);
  if (lane == 0)
    for (int b = 0; b < NBUF; b++) mbar_init(bar0 + 8 * b, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const int base = blockIdx.x * APB + w;
  const int my_atom = base + lane * WARPS;
  const bool mine = lane < APW && my_atom < nlocal;
  const int nn_mine = mine ? numneigh[my_atom] : 0;
  const int nt_mine = (nn_mine + CH - 1) / CH;
  int pre = nt_mine;
  for (int o = 1; o < APW; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, pre, o);
    if (lane >= o) pre += v;
  }
  const int total_tiles = min(__shfl_sync(0xffffffffu, pre, APW - 1), MAXTILES);
  if (lane < APW) {
    int t0 = pre - nt_mine;
    for (int c = 0; c < nt_mine && t0 + c < MAXTILES; c++) sm.sched[t0 + c] = (base + lane * WARPS) * (rowcap / CH) + c;
  }
  __syncthreads();
  int issued = 0;
  auto issue = [&]() {
    if (issued < total_tiles && lane == 0) {
      const unsigned int slot = issued & (NBUF - 1);
      mbar_expect_tx(bar0 + 8 * slot, CH * 4);
      bulk_g2s(tile0 + slot * (CH * 4), neigh + (size_t)sm.sched[issued] * CH, CH * 4, bar0 + 8 * slot);
    }
    issued++;
  };
#pragma unroll
  for (int b = 0; b < NBUF; b++) issue();

  int cslot = 0;
  for (int n = 0; n < APW; n++) {
    const int i = base + n * WARPS;
    if (i >= nlocal) break;
    const int ntile = __shfl_sync(0xffffffffu, nt_mine, n);
    const float4 pti = xt[i];
    int *row2 = neigh2 + (size_t)i * rowcap;
    int cnt = 0;
    for (int t = 0; t < ntile; t++) {
      const unsigned int slot = cslot & (NBUF - 1);
      mbar_wait(bar0 + 8 * slot, (cslot / NBUF) & 1);
      const int *tp = &sm.tile[slot][lane];
      const int r0 = tp[0], r1 = tp[32], r2 = tp[64], r3 = tp[96];
      __syncwarp();
      cslot++;
      issue();
      const float4 p0 = xt[r0], p1 = xt[r1], p2 = xt[r2], p3 = xt[r3];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int raw = u == 0 ? r0 : u == 1 ? r1 : u == 2 ? r2 : r3;
        const float4 pj = u == 0 ? p0 : u == 1 ? p1 : u == 2 ? p2 : p3;
        const float dx = pti.x - pj.x, dy = pti.y - pj.y, dz = pti.z - pj.z;
        const bool in = fmaf(dx, dx, fmaf(dy, dy, dz * dz)) < cutf_inner;
        const unsigned int m = __ballot_sync(0xffffffffu, in);
        if (in) row2[cnt + __popc(m & ltmask)] = raw | (__float_as_int(pj.w) << CPH_TYPESHIFT);
        cnt += __popc(m);
      }
    }
    const int padded = (cnt + 31) & ~31;
    if (cnt + lane < padded) row2[cnt + lane] = dummy | (1 << CPH_TYPESHIFT);
    if (lane == 0) numneigh2[i] = cnt;
  }
}

#ifndef CPH_EVAL_ILP2
#define CPH_EVAL_ILP2 0
#endif
#ifndef CPH_EVAL_MINBLOCKS
#define CPH_EVAL_MINBLOCKS 3
#endif
// atoms per warp in the evaluation kernel: 8 amortises the per-CTA table load best at 1M atoms,
// 2 gives the finer work granularity that wins below ~300k atoms per rank (both measured)

// K2b: one warp per atom over the pruned inner row.  Every lane evaluates one pair per
// iteration with the next entry already loaded; no queue, no ballots: all issue slots go to the
// fp64 evaluation.
// launch shape measured on B200 (profiles/r1_scaling_and_bench.md): 64-thread CTAs capped at 72
// registers (28 resident warps/SM, no spills) beat 256-thread CTAs at 80 registers by 8 %
#ifndef CPH_EVAL_WARPS
#define CPH_EVAL_WARPS 2
#endif
#ifndef CPH_EVAL_MAXNREG
#define CPH_EVAL_MAXNREG 72
#endif
constexpr int EWARPS = CPH_EVAL_WARPS;
constexpr int ETPB = EWARPS * 32;
#ifdef CPH_EVAL_MAXNREG
#define CPH_EVAL_BOUNDS __maxnreg__(CPH_EVAL_MAXNREG)
#else
#define CPH_EVAL_BOUNDS __launch_bounds__(ETPB, CPH_EVAL_MINBLOCKS)
#endif

template <int STYLE, int EFLAG, int UNI, int EAPW>
__global__ void CPH_EVAL_BOUNDS
pair_eval_kernel(int nlocal, const double4 *__restrict__ xq, const int *__restrict__ type,
                 const int *__restrict__ neigh, const int *__restrict__ numspec, const int *__restrict__ neigh2,
                 const int *__restrict__ numneigh2, int rowcap, int dummy, int nt1, const double4 *__restrict__ coef,
                 const double2 *__restrict__ cuts, const int *__restrict__ type_has_lj, double *__restrict__ f,
                 double *__restrict__ evdwl, double *__restrict__ phi, double *__restrict__ eatom, double c_self,
                 const unsigned int *__restrict__ gate) {
  // speculative launch (cph_post_force): enqueued before the host has seen this step's list flags;
  // if a re-neighbouring or a prune turns out to be due, the whole grid retires and the host
  // launches the pass again behind the rebuilt rows
  if (gate != nullptr && (gate[4] | gate[5]) != 0u) return;
  __shared__ double4 s_coef[CPH_MAXNT1 * CPH_MAXNT1];
  __shared__ double2 s_cut[CPH_MAXNT1 * CPH_MAXNT1];
  __shared__ double s_exp2[32];
  for (int k = threadIdx.x; k < nt1 * nt1; k += ETPB) {
    s_coef[k] = coef[k];
    if (!UNI) s_cut[k] = cuts[k];     // per-pair cutoffs are only read when they differ from the global one
  }
  if (threadIdx.x < 32) s_exp2[threadIdx.x] = kexp2[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int base = blockIdx.x * (EWARPS * EAPW) + w;
  for (int n = 0; n < EAPW; n++) {
    const int i = base + n * EWARPS;
    if (i >= nlocal) break;
    const double4 pi = xq[i];
    const int ti = type[i];
    const int tbase = ti * nt1;
    const bool has_lj = type_has_lj[ti] != 0;
    const int n2 = numneigh2[i];
    const int *row2 = neigh2 + (size_t)i * rowcap;
    // software pipeline: entry k+64 and the fp64 record of entry k+32 are in flight while
    // entry k is evaluated (the gather latency was the top stall of the unpipelined loop)
    int e0 = lane < n2 ? row2[lane] : dummy;
    int e1 = lane + 32 < n2 ? row2[lane + 32] : dummy;
    double4 p0 = ld256(xq + (e0 & CPH_JMASK));
    Acc a;
    const int nsp = numspec[i];
    if (nsp > 0) {   // special-bond partners sit at the end of the OUTER row
      if (lane < nsp) {
        const int raw = neigh[(size_t)i * rowcap + (rowcap - 1 - lane)];
        const int j = raw & CPH_NEIGHMASK, sb = (raw >> CPH_SBSHIFT) & 3;
        const double4 pq = ld256(xq + j);
        const int tt = tbase + type[j];
        const double delx = pi.x - pq.x, dely = pi.y - pq.y, delz = pi.z - pq.z;
        const double rsq = fma(delz, delz, fma(dely, dely, delx * delx));
        double o5[5];
        eval_special<STYLE, EFLAG>(s_coef, cuts /* global: rare path */, tt, delx, dely, delz, rsq, pi.w, pq.w, sb, o5);
        a.fx = o5[0]; a.fy = o5[1]; a.fz = o5[2];
        if (EFLAG) { a.ev = o5[3]; a.phi = o5[4]; }
      }
      __syncwarp();
    }
#if CPH_EVAL_ILP2
    // two independent pairs per lane per iteration (entries k and k+32): the dependent fp64
    // chains of the two evaluations interleave
    for (int k0 = 0; k0 < n2; k0 += 64) {
      const double4 p1 = ld256(xq + (e1 & CPH_JMASK));
      const int ka = k0 + 64 + lane, kb = k0 + 96 + lane;
      const int e2 = ka < n2 ? row2[ka] : dummy;
      const int e3 = kb < n2 ? row2[kb] : dummy;
      const double dx0 = pi.x - p0.x, dy0 = pi.y - p0.y, dz0 = pi.z - p0.z;
      const double dx1 = pi.x - p1.x, dy1 = pi.y - p1.y, dz1 = pi.z - p1.z;
      const double rs0 = fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0));
      const double rs1 = fma(dz1, dz1, fma(dy1, dy1, dx1 * dx1));
      double f0, f1, v0, v1, h0, h1;
      eval_pair<STYLE, EFLAG, UNI>(s_coef, s_cut, tbase + ((e0 >> CPH_TYPESHIFT) & 15), rs0, pi.w, p0.w, s_exp2,
                                   has_lj, f0, v0, h0);
      eval_pair<STYLE, EFLAG, UNI>(s_coef, s_cut, tbase + ((e1 >> CPH_TYPESHIFT) & 15), rs1, pi.w, p1.w, s_exp2,
                                   has_lj, f1, v1, h1);
      if (k0 + lane >= n2) { f0 = 0.0; v0 = 0.0; h0 = 0.0; }
      if (k0 + 32 + lane >= n2) { f1 = 0.0; v1 = 0.0; h1 = 0.0; }
      a.fx = fma(dx0, f0, a.fx); a.fy = fma(dy0, f0, a.fy); a.fz = fma(dz0, f0, a.fz);
      a.fx = fma(dx1, f1, a.fx); a.fy = fma(dy1, f1, a.fy); a.fz = fma(dz1, f1, a.fz);
      if (EFLAG) { a.ev += v0 + v1; a.phi += h0 + h1; }
      e0 = e2; e1 = e3;
      p0 = ld256(xq + (e0 & CPH_JMASK));
    }
#else
    for (int k0 = 0; k0 < n2; k0 += 32) {
      const double4 p1 = ld256(xq + (e1 & CPH_JMASK));
      const int kn = k0 + 64 + lane;
      const int e2 = kn < n2 ? row2[kn] : dummy;
      if (k0 + lane < n2) {
        const double delx = pi.x - p0.x, dely = pi.y - p0.y, delz = pi.z - p0.z;
        const double rsq = fma(delz, delz, fma(dely, dely, delx * delx));
        double fp, ev, ph;
        eval_pair<STYLE, EFLAG, UNI>(s_coef, s_cut, tbase + ((e0 >> CPH_TYPESHIFT) & 15), rsq, pi.w, p0.w, s_exp2,
                                     has_lj, fp, ev, ph);
        a.fx = fma(delx, fp, a.fx); a.fy = fma(dely, fp, a.fy); a.fz = fma(delz, fp, a.fz);
        if (EFLAG) { a.ev += ev; a.phi += ph; }
      }
      e0 = e1; p0 = p1; e1 = e2;
    }
#endif
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
        const double ev = 0.5 * a.ev;
        const double ph = a.phi + 2.0 * pi.w * c_self;
        evdwl[i] = ev;
        phi[i] = ph;
        eatom[i] = ev + 0.5 * pi.w * ph;
      }
    }
  }
}

__global__ void snapshot_kernel(int n, const double4 *__restrict__ xq, double *xs) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double4 p = xq[k];
  xs[3 * (size_t)k] = p.x; xs[3 * (size_t)k + 1] = p.y; xs[3 * (size_t)k + 2] = p.z;
}

// packed fp32 {x - origin, y - origin, z - origin, type} for the prefilter
__global__ void xt_kernel(int nall, const double4 *__restrict__ xq, const int *__restrict__ type, double3 origin,
                          float4 *xt) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nall) return;   // index nall is the far-away dummy atom
  double4 p = xq[k];
  xt[k] = make_float4((float)(p.x - origin.x), (float)(p.y - origin.y), (float)(p.z - origin.z),
                      __int_as_float(type[k]));
}

}  // namespace

// Constants of the hot loop live in __constant__ memory, which is per device, not per handle:
// the handle that launches re-uploads them when another handle (other pair parameters) was the
// last user of the device.  Handles with DIFFERENT pair parameters must not run concurrently.
static const cph_handle *g_kc_owner[64] = {nullptr};

int cph_pair_upload_constants(cph_handle *h) {
  if (h->device >= 0 && h->device < 64) {
    if (g_kc_owner[h->device] && g_kc_owner[h->device] != h) cudaDeviceSynchronize();
    g_kc_owner[h->device] = h;
  }
  const PairParams &pp = h->pp;
  EvalConst c{};
  c.ewp_alpha = 0.3275911 * pp.alpha;
  c.a1 = 0.254829592; c.a2 = -0.284496736; c.a3 = 1.421413741; c.a4 = -1.453152027; c.a5 = 1.061405429;
  c.neg_alpha2 = -pp.alpha * pp.alpha;
  c.two_alpha_pis = 2.0 * pp.alpha / 1.77245385090551602729;
  c.qqrd2e = pp.qqrd2e; c.e_shift = pp.e_shift; c.f_shift = pp.f_shift;
  c.cut_coulsq = pp.cut_coulsq; c.cutsq_max = pp.cutsq_max;
  const double ln2 = 0.693147180559945309417232121458;
  c.exp_scale = 32.0 / ln2;
  c.exp_magic = 6755399441055744.0;
  c.exp_c1 = ln2 / 32.0;
  c.p2 = 1.0 / 2; c.p3 = 1.0 / 6; c.p4 = 1.0 / 24; c.p5 = 1.0 / 120; c.p6 = 1.0 / 720;
  for (int k = 0; k < 4; k++) {
    c.flj[k] = pp.special_lj[k];
    c.fcoul[k] = pp.special_coul[k];
    c.one_m_fc[k] = 1.0 - pp.special_coul[k];
  }
  double e2[32];
  for (int j = 0; j < 32; j++) e2[j] = (double)exp2l((long double)j / 32.0L);
  CPH_CUDA(h, cudaMemcpyToSymbolAsync(kc, &c, sizeof(c), 0, cudaMemcpyHostToDevice, h->stream));
  CPH_CUDA(h, cudaMemcpyToSymbolAsync(kexp2, e2, sizeof(e2), 0, cudaMemcpyHostToDevice, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// fp32 prefilter records of every atom (owned, ghost, dummy); positions relative to the grid origin
int cph_launch_xt(cph_handle *h) {
  ProfScope ps(h, 1);
  const double3 origin = make_double3(h->grid.lo[0], h->grid.lo[1], h->grid.lo[2]);
  CPH_CUDA(h, h->d_xt.reserve((size_t)h->nall + 2));
  h->nlaunch++;
  xt_kernel<<<(h->nall + 256) / 256, 256, 0, h->stream>>>(h->nall, h->d_xq.p, h->d_type.p, origin, h->d_xt.p);
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

void cph_pair_forget(cph_handle *h) {
  for (auto &o : g_kc_owner)
    if (o == h) o = nullptr;
}

// K2a launcher: prune the Verlet rows to rc + inner skin and remember where the atoms were
int cph_launch_prune(cph_handle *h) {
  const int n = h->nlocal;
  if (n == 0) { h->inner_valid = true; return 0; }
  CPH_TRY(cph_launch_xt(h));
  ProfScope ps(h, 1);
  double extent = 0;
  for (int k = 0; k < 3; k++) extent = std::max(extent, h->grid.n[k] / h->grid.inv[k]);
  const double cut = std::sqrt(h->pp.cutsq_max) + h->inner_skin;
  const float cutf = (float)(cut * cut + 32.0 * cut * extent * 5.97e-8 + 1e-5 * cut * cut);
  CPH_CUDA(h, h->d_neigh2.reserve((size_t)n * h->rowcap));
  CPH_CUDA(h, h->d_numneigh2.reserve(n + 1));
  CPH_CUDA(h, h->d_xinner.reserve(3 * (size_t)n + 3));
  if (h->rowcap / CH * APW > MAXTILES)
    return cph_fail(h, CPH_ERR_OVERFLOW, "neighbour rows of %d entries exceed the prune kernel's tile schedule", h->rowcap);
  h->nlaunch += 2;
  prune_kernel<<<(n + APB - 1) / APB, TPB, 0, h->stream>>>(n, h->d_xt.p, h->d_neigh.p, h->d_numneigh.p, h->rowcap, cutf,
                                                          h->nall, h->d_neigh2.p, h->d_numneigh2.p);
  snapshot_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(n, h->d_xq.p, h->d_xinner.p);
  CPH_CUDA(h, cudaGetLastError());
  h->inner_valid = true;
  h->nprunes++;
  return 0;
}

int cph_launch_pair(cph_handle *h, int eflag, const unsigned int *gate) {
  const int n = h->nlocal;
  if (n == 0) return 0;
  if (gate && (h->fused_pair || !h->inner_valid))
    return cph_fail(h, CPH_ERR_STATE, "a gated pair pass needs valid inner rows");
  if (h->device < 0 || h->device >= 64 || g_kc_owner[h->device] != h || h->kc_dirty) {
    CPH_TRY(cph_pair_upload_constants(h));
    h->kc_dirty = false;
  }
  if (h->fused_pair) return cph_launch_pair_fused(h, eflag);
  if (!h->inner_valid) CPH_TRY(cph_launch_prune(h));
  ProfScope ps(h, 0);
  const int nt1 = h->pp.ntypes + 1;
  const bool small = n < 300000;
  // atoms per warp: fewer on small boxes so the grid still covers the 148 SMs several times over
  static const int eapw_env = getenv("CPH_EAPW") ? atoi(getenv("CPH_EAPW")) : 0;
  // measured on B200 (profiles/r1_scaling_and_bench.md): 2 and 4 tie at 125k atoms, 4 wins by 1 % at 250k
  const int eapw = (eapw_env == 1 || eapw_env == 2 || eapw_env == 4 || eapw_env == 8) ? eapw_env
                   : n < 200000 ? 2 : small ? 4 : 8;
  const int blocks = (n + EWARPS * eapw - 1) / (EWARPS * eapw);
  h->nlaunch++;
#define LAUNCH(S, E, U)                                                                 \
  do {                                                                                  \
    if (eapw == 1) LAUNCH_A(S, E, U, 1); else if (eapw == 2) LAUNCH_A(S, E, U, 2);      \
    else if (eapw == 4) LAUNCH_A(S, E, U, 4); else LAUNCH_A(S, E, U, 8);                \
  } while (0)
#define LAUNCH_A(S, E, U, A)                                                                                        \
  pair_eval_kernel<S, E, U, A><<<blocks, ETPB, 0, h->stream>>>(n, h->d_xq.p, h->d_type.p, h->d_neigh.p, h->d_numspec.p, \
                                                           h->d_neigh2.p, h->d_numneigh2.p, h->rowcap, h->nall, nt1, \
                                                           h->d_coef4.p, h->d_cut2.p, h->d_type_has_lj.p,          \
                                                           h->d_f.p, h->d_evdwl.p, h->d_phi.p, h->d_eatom.p,       \
                                                           h->pp.c_self, gate)
#define LAUNCH_E(S, U) do { if (eflag) LAUNCH(S, 1, U); else LAUNCH(S, 0, U); } while (0)
  if (h->pp.style == CPH_PAIR_LJ_CUT_COUL_CUT) {
    if (h->uniform_cut) LAUNCH_E(CPH_PAIR_LJ_CUT_COUL_CUT, 1); else LAUNCH_E(CPH_PAIR_LJ_CUT_COUL_CUT, 0);
  } else {
    if (h->uniform_cut) LAUNCH_E(CPH_PAIR_LJ_CUT_COUL_DSF, 1); else LAUNCH_E(CPH_PAIR_LJ_CUT_COUL_DSF, 0);
  }
#undef LAUNCH_E
#undef LAUNCH
#undef LAUNCH_A
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

// single-kernel variant (filter + evaluation fused), kept for A/B runs: CPH_PAIR_FUSED=1
int cph_launch_pair_fused(cph_handle *h, int eflag) {
  const int n = h->nlocal;
  if (n == 0) return 0;
  CPH_TRY(cph_launch_xt(h));
  if (h->device < 0 || h->device >= 64 || g_kc_owner[h->device] != h || h->kc_dirty) {
    CPH_TRY(cph_pair_upload_constants(h));
    h->kc_dirty = false;
  }
  ProfScope ps(h, 0);
  // conservative fp32 cutoff: coordinates relative to the grid origin are below `extent`, so
  // |r2_fp32 - r2_fp64| <= ~8 * cut * extent * 2^-24; the margin is 4x that plus a relative term
  double extent = 0;
  for (int k = 0; k < 3; k++) extent = std::max(extent, h->grid.n[k] / h->grid.inv[k]);
  const double cut = std::sqrt(h->pp.cutsq_max);
  const float cutf = (float)(h->pp.cutsq_max + 32.0 * cut * extent * 5.97e-8 + 1e-5 * h->pp.cutsq_max);
  const int blocks = (n + APB - 1) / APB;
  if (h->rowcap / CH * APW > MAXTILES)
    return cph_fail(h, CPH_ERR_OVERFLOW, "neighbour rows of %d entries exceed the pair kernel's tile schedule", h->rowcap);
  h->nlaunch++;
  const int nt1 = h->pp.ntypes + 1;
#define LAUNCH(S, E, U)                                                                                           \
  pair_fused_kernel<S, E, U><<<blocks, TPB, 0, h->stream>>>(n, h->d_xq.p, h->d_xt.p, h->d_neigh.p, h->d_numneigh.p,    \
                                                      h->d_numspec.p, h->rowcap, nt1, cutf, h->d_coef4.p, h->d_cut2.p,           \
                                                      h->d_type_has_lj.p, h->d_f.p, h->d_evdwl.p, h->d_phi.p,    \
                                                      h->d_eatom.p, h->pp.c_self)
#define LAUNCH_E(S, U) do { if (eflag) LAUNCH(S, 1, U); else LAUNCH(S, 0, U); } while (0)
  if (h->pp.style == CPH_PAIR_LJ_CUT_COUL_CUT) {
    if (h->uniform_cut) LAUNCH_E(CPH_PAIR_LJ_CUT_COUL_CUT, 1); else LAUNCH_E(CPH_PAIR_LJ_CUT_COUL_CUT, 0);
  } else {
    if (h->uniform_cut) LAUNCH_E(CPH_PAIR_LJ_CUT_COUL_DSF, 1); else LAUNCH_E(CPH_PAIR_LJ_CUT_COUL_DSF, 0);
  }
#undef LAUNCH_E
#undef LAUNCH
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}
