// pair.cu -- K2: fp64 pair pass for lj/cut/coul/cut and lj/cut/coul/dsf over a two-level list.
//
// Stands in for the LAMMPS pair style whose per-atom energies the reference reads at
// fix_constant_pH.cpp:216-219 (`force->pair->eatom`), restated from SURVEY.md Appendix A,
// and adds what north_star asks for in the same pass: forces, the electrostatic
// potential phi_i = dE_coul/dq_i that gives every site's dU/dlambda analytically
// (Appendix B), and the van-der-Waals share of the per-atom energy.
//
// Two kernels ("rolling prune"):
//   K2a prune_kernel (fp32, every few steps): one warp per atom streams the Verlet row
//       (rc + skin) HBM -> shared memory through a per-warp ring of 512-byte tiles filled by
//       TMA bulk copies (cp.async.bulk + mbarrier), tests every candidate on the packed fp32
//       record against (rc + inner skin)^2 + margin and ballot-compacts the survivors into the
//       inner row (j | type_j << 28, padded to a multiple of 32 with a far-away dummy atom).
//   K2b pair_eval_kernel (fp64, every step): one warp per atom over the inner row, one pair per
//       lane per iteration, manually unrolled by two so the gathers of the next chunk are in
//       flight while the current one is evaluated and no register rotation is needed.
//
// K2b is bound by the fp64 pipe (a warp DFMA issues every second cycle on sm_100a), so the
// loop is written to spend as few fp64 AND as few other issue slots per pair as possible
// (round-2 instruction diet, profiles/r2_*):
//   * 1/sqrt and 1/x: hardware seed (MUFU.RSQ64H / MUFU.RCP64H, relative error 2^-20 measured by
//     cph_bench_seed_error) + one Newton step (third order for 1/r, second order for the erfc
//     argument: fastmath.cuh);
//   * exp(-alpha^2 r^2): conflict-free 16-entry table of 2^(j/16) in shared memory times a
//     degree-7 polynomial whose coefficients absorb -alpha^2, so the argument is r^2 itself;
//   * qqrd2e and q_i are applied once per atom after the warp reduction, not per pair;
//   * out-of-range lanes are neutralised by ONE select on the high word of q_j/r (the value
//     becomes a denormal that cannot change any sum), not by 64-bit selects on every output;
//   * LJ is a separate instantiation of the row loop chosen per atom (water H: 2/3 of the atoms
//     have no LJ partner at all), and the per-type-pair cutoff variant another;
//   * all tables are addressed with 32-bit shared-memory addresses (no generic-pointer
//     conversion inside the loop), the record address is one IMAD.WIDE;
//   * constants travel as a __grid_constant__ kernel parameter, so handles with different pair
//     parameters are independent (nothing lives in device-global __constant__ memory).
// Every pair inside the cutoff is still DECIDED and EVALUATED in fp64 with LAMMPS' polynomial
// erfc (same coefficients as the oracle -- it is not erfc()).  Full list => no atomics and a
// fixed summation order => bit-reproducible run to run.
//
// Per-atom outputs: f (3), evdwl_i = 1/2 sum_j evdwl_ij, phi_i, eatom_i = evdwl_i +
// 1/2 q_i phi_i  (== ev_tally's half-half split plus the dsf self term).
#include "cph_internal.h"
#include "fastmath.cuh"

namespace {

constexpr int WARPS = 8;
constexpr int TPB = WARPS * 32;
constexpr int APW = 4;            // atoms per warp (sequential) in the prune kernel
constexpr int APB = WARPS * APW;  // atoms per block
constexpr int CH = 128;           // candidates per tile (4 per lane)
constexpr int NBUF = 4;           // tiles in flight per warp
constexpr int MAXTILES = 64;      // per-warp tile schedule entries (rows of up to 2048 neighbours)

// ---- small PTX helpers ------------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds_f64(unsigned int a) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ double2 lds_v2f64(unsigned int a) {
  double2 v;
  asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
// the fp64 record {x,y,z,q} of one atom: one 32-byte request per lane (LDG.E.ENL2.256 on sm_100a)
__device__ __forceinline__ double4 ld256(const void *p) {
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
// inner-row entries are read exactly once: do not let them displace the neighbour records in L1
__device__ __forceinline__ int ld_stream(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// xq + 32 * j as one IMAD.WIDE.U32
__device__ __forceinline__ const void *rec_addr(const double4 *xq, unsigned int j) {
  unsigned long long a;
  asm("mad.wide.u32 %0, %1, 32, %2;" : "=l"(a) : "r"(j), "l"(xq));
  return (const void *)a;
}
// prune kernel: fp32 record of atom j and the store of inner-row entry `pos`, one IMAD.WIDE.U32 per address
__device__ __forceinline__ float4 ld_xt(const float4 *xt, int j) {
  unsigned long long a;
  asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(j), "l"(xt));
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a));
  return v;
}
__device__ __forceinline__ void st_entry(unsigned long long rowp, unsigned int pos, int v) {
  unsigned long long a;
  asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(a) : "r"(pos), "l"(rowp));
  asm volatile("st.global.b32 [%0], %1;" ::"l"(a), "r"(v) : "memory");
}
// ---- TMA bulk copy + mbarrier (per-warp ring of the prune kernel) ------------------------------
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

// One pair of the hot loop, straight-line.  Accumulates into `a`:
//   LJ == false: a.f += del * (q_j/r * fc/r)   (Coulomb only; the caller multiplies by q_i qqrd2e once per atom)
//   LJ == true : a.f += del * (q_i qqrd2e * that + LJ)            a.ev += evdwl_ij
//   a.phi += q_j * K(r) / qqrd2e
// UNI: every type pair has cut_lj == cut_coul == the one global cutoff (one exact fp64 decision);
// otherwise the three decisions of SURVEY Appendix A are taken from the per-type-pair table.
// SPECIAL: the chunk may hold special-bond partners (class in bits 26-27 of the entry; only the LAST chunk of a row
// does): their LJ and Coulomb terms are weighted with special_lj / special_coul, and under dsf the undamped term
// (1 - factor_coul) q_i q_j qqrd2e / r is taken out of energy and force, branch-free.
template <int STYLE, int EFLAG, bool LJ, bool UNI, bool SPECIAL = false>
__device__ __forceinline__ void eval_one(const EvalConst &c, const double4 &pi, const double4 &pj, const int e,
                                         const double qiq, const unsigned int coef_i, const unsigned int cut_i,
                                         const unsigned int exp_tab, Acc &a) {
  const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
  const double s = fma(dz, dz, fma(dy, dy, dx * dx));
  bool in_c, in_lj;
  if (UNI) {
    in_c = in_lj = s < c.cutsq;                          // the exact (fp64) cutoff decision
  } else {
    const double2 cc = lds_v2f64(cut_i + (((unsigned int)e >> CPH_TYPESHIFT) << 4));   // {cut_ljsq, cutsq}
    in_c = s < c.cut_coulsq && s < cc.y;
    in_lj = s < cc.x && s < cc.y;
  }
  const double y = fast_rsqrt(s);                        // 1/r
  double u = keep_if(in_c, pj.w * y);                    // q_j / r, neutralised when out of range
  double f_lj = 1.0, omf = 0.0;                          // factor_lj, 1 - factor_coul
  if (SPECIAL) {
    const int sb = (e >> CPH_SB2SHIFT) & 3;
    f_lj = c.flj[sb];
    omf = c.one_m_fc[sb];
    if (STYLE == CPH_PAIR_LJ_CUT_COUL_CUT) u *= c.fcoul[sb];
  }
  double fp;
  if (STYLE == CPH_PAIR_LJ_CUT_COUL_CUT) {
    if (EFLAG) a.phi += u;                               // E_ij = factor_coul qqrd2e q_i q_j / r
    fp = u * (y * y);
  } else {
    const double r = s * y;
    // erfcd = exp(-alpha^2 s) = 2^(n/16) * P7(rho),  n = round(-alpha^2 s 16/ln2),  rho = s + n ln2/(16 alpha^2).
    // 16 table entries fill one 128-byte row of shared memory, one bank pair each: the look-up is conflict-free
    // whatever the lanes ask for (a 256-entry table with a degree-4 polynomial cost 7 wavefronts per look-up
    // on the L1 data pipe, which is this kernel's busiest unit: profiles/r2b_eval_ncu.txt)
    const double tm = fma(s, c.exp_scale, c.exp_magic);
    int ni = __double2loint(tm);
    ni = in_c ? ni : 0;                                  // out-of-range lanes (the far-away dummy) stay finite
    const double nd = tm - c.exp_magic;
    const double rho = fma(nd, c.exp_c1s, s);
    double p = fma(rho, c.b7, c.b6);
    p = fma(rho, p, c.b5);
    p = fma(rho, p, c.b4);
    p = fma(rho, p, c.b3);
    p = fma(rho, p, c.b2);
    p = fma(rho, p, c.b1);
    p = fma(rho, p, 1.0);
    const double v = lds_f64(exp_tab + ((ni & 15) << 3)) * p;
    const double D = __hiloint2double(__double2hiint(v) + (int)((unsigned int)(ni & ~15) << 16), __double2loint(v));
    // LAMMPS' polynomial erfc: t = 1/(1 + p alpha r), erfcc = t (a1 + t (a2 + ...)) erfcd
    // special partners are all bonded pairs of like sign pattern (O-H, H-H): their terms add up coherently, and so
    // would the always-negative 1e-12 error of the second-order step; the one chunk that holds them takes the third
    const double t = fast_rcp<SPECIAL ? 3 : CPH_REFINE_RCP>(fma(c.pa, r, 1.0));
    double q = fma(t, c.a5, c.a4);
    q = fma(t, q, c.a3);
    q = fma(t, q, c.a2);
    q = fma(t, q, c.a1);
    const double E = (t * q) * D;
    double kk = fma(-s, c.f_shift, fma(-r, c.e_shift, E));
    // forcecoul / r^2 = q_i q_j qqrd2e / r * (erfcc/r + 2 alpha/sqrt(pi) erfcd + r f_shift) / r
    double fc = fma(E, y, fma(c.cD, D, r * c.f_shift));
    if (SPECIAL) {
      kk -= omf;
      fc = fma(-omf, y, fc);
    }
    if (EFLAG) a.phi = fma(u, kk, a.phi);
    fp = u * (fc * y);
  }
  if (LJ) {
    // one 16-byte shared-memory load per pair: {lj3, lj4}; the force needs 12 lj3 and 6 lj4, which cost
    // two fp64 multiplies -- the fp64 pipe has the slack, the L1 data pipe does not (profiles/r2a_eval_ncu.txt)
    const double2 c34 = lds_v2f64(coef_i + (((unsigned int)e >> CPH_TYPESHIFT) << 4));
    const double r2 = y * y;
    double r6 = keep_if(in_lj, r2 * r2 * r2);
    double lj_scale = r2;
    if (SPECIAL) lj_scale *= f_lj;
    const double t3 = c34.x * r6;                        // lj3 r^-6
    fp = fma(qiq, fp, r6 * fma(12.0, t3, -6.0 * c34.y) * lj_scale);
    if (EFLAG) a.ev = fma(SPECIAL ? r6 * f_lj : r6, t3 - c34.y, a.ev);
  }
  a.fx = fma(dx, fp, a.fx);
  a.fy = fma(dy, fp, a.fy);
  a.fz = fma(dz, fp, a.fz);
}

// The inner row of one atom, two 32-entry chunks per trip.  While chunk A is evaluated the record
// gather of chunk B and the index load of the chunk after it are in flight, and vice versa; the
// index of a chunk past the row end is clamped to this lane's entry of the last chunk (a valid
// index whose record is loaded but never evaluated).  n2pad is a non-zero multiple of 32.
template <int STYLE, int EFLAG, bool LJ, bool UNI>
__device__ __forceinline__ void row_loop(const EvalConst &c, const double4 *__restrict__ xq, const int *__restrict__ row,
                                         const int n2pad, const int lane, const double4 &pi, const double qiq,
                                         const unsigned int coef_i, const unsigned int cut_i, const unsigned int exp_tab,
                                         Acc &a) {
  const int last = n2pad - 32 + lane;
  const int nmain = n2pad - 32;                          // the last chunk (special partners, padding) is evaluated apart
  // indices run TWO trips ahead of their use (they stream from HBM / L2), records one chunk ahead (mostly L1 hits)
  int eA = ld_stream(row + lane);
  int eB = ld_stream(row + min(lane + 32, last));
  int eC = ld_stream(row + min(lane + 64, last));
  int eD = ld_stream(row + min(lane + 96, last));
  double4 pA = ld256(rec_addr(xq, (unsigned int)eA & CPH_JMASK));
  for (int k = 0; k < nmain; k += 64) {
    const double4 pB = ld256(rec_addr(xq, (unsigned int)eB & CPH_JMASK));
    const int eE = ld_stream(row + min(k + 128 + lane, last));
    const int eF = ld_stream(row + min(k + 160 + lane, last));
    eval_one<STYLE, EFLAG, LJ, UNI>(c, pi, pA, eA, qiq, coef_i, cut_i, exp_tab, a);
    if (k + 32 >= nmain) {                               // warp-uniform: chunk B is the last chunk of the row
      pA = pB;
      eA = eB;
      break;
    }
    pA = ld256(rec_addr(xq, (unsigned int)eC & CPH_JMASK));
    eval_one<STYLE, EFLAG, LJ, UNI>(c, pi, pB, eB, qiq, coef_i, cut_i, exp_tab, a);
    eA = eC; eB = eD; eC = eE; eD = eF;
  }
  eval_one<STYLE, EFLAG, LJ, UNI, true>(c, pi, pA, eA, qiq, coef_i, cut_i, exp_tab, a);
}

struct WarpSmem {
  int tile[NBUF][CH];                 // neighbour tiles (TMA destination), 16-byte aligned
  unsigned long long bar[NBUF];
  int sched[MAXTILES];                // tile schedule: tile index (row offset / CH) of each tile of this warp
};

// K2a: one warp per atom; TMA tile ring over the outer row; survivors written to the inner row
// as (j | type_j << 28), padded to a multiple of 32 with the dummy atom.
// ES (LJ end states declared, ljstates.cu): while an atom WITHOUT end states is pruned, the survivors that have
// them are also written to the atom's short correction list (entry-major, es.cap slots per atom).
struct PruneEs {
  const int *tB;                 // [nall+1] B-state type of every owned / ghost atom, 0 = none
  const unsigned int *tmask;     // bit t: some atom of A-state type t has end states (saves the tB gather)
  int *cnt, *ent, *over;         // per-atom count, entries [k * nlocal + i], largest count seen
  int cap;
};
template <bool ES>
__global__ void __launch_bounds__(TPB, 4)
prune_kernel(int nlocal, const float4 *__restrict__ xt, const int *__restrict__ neigh,
             const int *__restrict__ numneigh, const int *__restrict__ numspec, int rowcap, float cutf_inner,
             int dummy, const int *__restrict__ type_has_lj, int *__restrict__ neigh2, int rowcap2,
             int *__restrict__ numneigh2, const __grid_constant__ PruneEs es) {
  __shared__ __align__(128) WarpSmem s_w[WARPS];
  const int lane = threadIdx.x & 31;
  // the warp index taken from lane 0: ptxas then KNOWS it is warp-uniform (ring addresses and the tile schedule
  // live in uniform registers, the bulk-copy issue needs no per-lane waterfall)
  const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
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
    int *row2 = neigh2 + (size_t)i * rowcap2;
    const unsigned long long row2p = (unsigned long long)row2;
    int cnt = 0;
    // correction list of an atom without end states of its own (an atom WITH end states needs none: every entry
    // of its row is corrected)
    int ecnt = 0;
    unsigned int tmask = 0;
    if (ES) tmask = es.tB[i] ? 0u : *es.tmask;
    auto collect = [&](const bool in, const int entry) {
      const bool q = in && ((tmask >> ((unsigned int)entry >> CPH_TYPESHIFT)) & 1u) && es.tB[entry & CPH_JMASK] != 0;
      const unsigned int mq = __ballot_sync(0xffffffffu, q);
      if (mq) {
        const int pos = ecnt + __popc(mq & ltmask);
        if (q && pos < es.cap) es.ent[(size_t)pos * nlocal + i] = entry;
        ecnt += __popc(mq);
      }
    };
    for (int t = 0; t < ntile; t++) {
      const unsigned int slot = cslot & (NBUF - 1);
      mbar_wait(bar0 + 8 * slot, (cslot / NBUF) & 1);
      const int *tp = &sm.tile[slot][lane];
      const int r0 = tp[0], r1 = tp[32], r2 = tp[64], r3 = tp[96];
      __syncwarp();
      cslot++;
      issue();
      const float4 p0 = ld_xt(xt, r0), p1 = ld_xt(xt, r1), p2 = ld_xt(xt, r2), p3 = ld_xt(xt, r3);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int raw = u == 0 ? r0 : u == 1 ? r1 : u == 2 ? r2 : r3;
        const float4 pj = u == 0 ? p0 : u == 1 ? p1 : u == 2 ? p2 : p3;
        const float dx = pti.x - pj.x, dy = pti.y - pj.y, dz = pti.z - pj.z;
        const bool in = fmaf(dx, dx, fmaf(dy, dy, dz * dz)) < cutf_inner;
        const unsigned int m = __ballot_sync(0xffffffffu, in);
        const int entry = raw | (__float_as_int(pj.w) << CPH_TYPESHIFT);
        if (in) st_entry(row2p, (unsigned int)(cnt + __popc(m & ltmask)), entry);
        cnt += __popc(m);
        if (ES) collect(in, entry);
      }
    }
    // Special-bond partners (at the end of the outer row, class in bits 30-31) ride in the LAST chunk of the inner
    // row, class in bits 26-27, where the evaluation kernel applies their factors; they use lanes that would
    // otherwise hold padding.  If they do not fit behind the ordinary entries of the last chunk, that chunk is
    // padded out and they get one of their own.  No distance test: an excluded pair is a bonded pair.
    const int nsp = min(numspec[i], 32);
    if (nsp > 0) {
      if ((cnt & 31) + nsp > 32) {
        const int padded0 = (cnt + 31) & ~31;
        if (cnt + lane < padded0) row2[cnt + lane] = dummy | (1 << CPH_TYPESHIFT);
        cnt = padded0;
      }
      int entry = 0;
      if (lane < nsp) {
        const int raw = neigh[(size_t)i * rowcap + (rowcap - 1 - lane)];
        const int j = raw & CPH_NEIGHMASK, sb = (raw >> CPH_SBSHIFT) & 3;
        entry = j | (sb << CPH_SB2SHIFT) | (__float_as_int(xt[j].w) << CPH_TYPESHIFT);
        row2[cnt + lane] = entry;
      }
      if (ES) collect(lane < nsp, entry);
      cnt += nsp;
    }
    // an inner row holds at most its outer row (ordinary + special entries) plus padding: <= rowcap2 (cph_launch_prune)
    const int padded = max((cnt + 31) & ~31, 32);
    if (cnt + lane < padded) row2[cnt + lane] = dummy | (1 << CPH_TYPESHIFT);
    if (lane == 0) {                                     // at least one (all-padding) chunk, so every row has a last chunk
      const int ti = __float_as_int(pti.w);
      numneigh2[i] = max(cnt, 1) | (ti << 24) | (type_has_lj[ti] ? 1 << 30 : 0);
      if (ES) {
        es.cnt[i] = min(ecnt, es.cap);
        if (ecnt > es.cap) atomicMax(es.over, ecnt);
      }
    }
  }
}

// launch shape measured on B200 (profiles/r1_scaling_and_bench.md): 64-thread CTAs under a register
// cap beat 256-thread CTAs; the grid is persistent (a multiple of the SM count) and walks the atom
// blocks with a stride, so the 2 KB exp table and the coefficient table are staged once per CTA
#ifndef CPH_EVAL_WARPS
#define CPH_EVAL_WARPS 2
#endif
#ifndef CPH_EVAL_MAXNREG
#define CPH_EVAL_MAXNREG 64
#endif
constexpr int EWARPS = CPH_EVAL_WARPS;
#ifndef CPH_EVAL_CLAIM
#define CPH_EVAL_CLAIM 2
#endif
constexpr int CLAIM = CPH_EVAL_CLAIM;   // atoms per queue claim
constexpr int ETPB = EWARPS * 32;

// K2b: one warp per atom over the pruned inner row.
template <int STYLE, int EFLAG, bool UNI>
__global__ void __maxnreg__(CPH_EVAL_MAXNREG)
pair_eval_kernel(const __grid_constant__ EvalArgs A) {
  // speculative launch (cph_post_force): enqueued before the host has seen this step's list flags;
  // if a re-neighbouring or a prune turns out to be due, the whole grid retires and the host
  // launches the pass again behind the rebuilt rows
  if (A.gate != nullptr && (A.gate[4] | A.gate[5]) != 0u) return;
  __shared__ __align__(16) double2 s_coef[CPH_MAXNT1 * CPH_MAXNT1];   // {lj3, lj4}
  __shared__ __align__(16) double2 s_cut[UNI ? 1 : CPH_MAXNT1 * CPH_MAXNT1];
  __shared__ __align__(128) double s_exp2[16];
  const EvalConst &c = A.c;
  const int nt1 = A.nt1;
  for (int k = threadIdx.x; k < nt1 * nt1; k += ETPB) {
    s_coef[k] = make_double2(A.coef[k].z, A.coef[k].w);
    if (!UNI) s_cut[k] = A.cuts[k];   // per-pair cutoffs are only read when they differ from the global one
  }
  if (STYLE == CPH_PAIR_LJ_CUT_COUL_DSF)
    if (threadIdx.x < 16) s_exp2[threadIdx.x] = A.exp2[threadIdx.x];
  __syncthreads();
  const unsigned int exp_tab = smem_u32(s_exp2), coef0 = smem_u32(s_coef), cut0 = smem_u32(s_cut);
  const int lane = threadIdx.x & 31;
  // Work distribution: the atoms (cell-sorted, so neighbours in index are neighbours in space) are split into
  // one contiguous range per SM, and every warp of an SM takes the NEXT atom of its SM's range from an atomic
  // counter.  The 32 warps resident on an SM therefore work on ~32 consecutive atoms at any moment, whose
  // neighbour records overlap almost completely: the gathers hit in L1 instead of travelling to L2.  A warp
  // whose own range is exhausted steals from the other ranges (same counters), so every atom is evaluated
  // exactly once whatever the block placement, and the tail balances itself.  Results do not depend on who
  // evaluates an atom.
  unsigned int smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  const int nq = A.nqueues;
  const int per_q = (A.nlocal + nq - 1) / nq;
  // main launch: a warp starts on its SM's range and, when that is dry, helps on the next maxscan - 1 ranges (load
  // balance at the end without walking all the ranges: every hop is an atomic round trip to L2).  The sweep launch
  // that follows binds one CTA to every range instead (no hops): whatever the main launch left -- a range whose SM
  // hosted no CTA, say -- is evaluated there, so the cover is exact whatever the block placement was.
  const int maxscan = A.sweep ? 1 : A.maxscan;
  int q = __shfl_sync(0xffffffffu, (int)((A.sweep ? blockIdx.x : smid) % (unsigned int)nq), 0), scanned = 0;
  // a claim is CLAIM consecutive atoms (one atomic round trip to L2 per CLAIM atoms)
  auto next_claim = [&]() -> int {
    while (scanned < maxscan) {
      int t = 0;
      if (lane == 0) t = atomicAdd(A.qnext + q, CLAIM);
      t = __shfl_sync(0xffffffffu, t, 0);                 // warp-uniform, and the compiler can tell
      const int cand = q * per_q + t;
      if (t < per_q && cand < A.nlocal) return cand;
      scanned++;                                          // this range is done: on to the next one
      q = q + 1 == nq ? 0 : q + 1;
    }
    return -1;
  };
  for (int i0 = next_claim(); i0 >= 0; i0 = next_claim()) {
    const int iend = min(min(i0 + CLAIM, (q + 1) * per_q), A.nlocal);   // claims do not cross a range end (q: the claim's range)
    for (int i = i0; i < iend; i++) {
      // the next atom's record, row length and the head of its row: into L1 now, so that its prologue does not wait on L2
      if (i + 1 < iend && lane == 0) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.xq + i + 1));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.numneigh2 + i + 1));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.neigh2 + (size_t)(i + 1) * A.rowcap2));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.neigh2 + (size_t)(i + 1) * A.rowcap2 + 32));
      }
      // The inner rows stream from HBM exactly once; a row is asked for PF atoms before a warp of this SM gets to
      // it (the SM's warps take consecutive atoms), so its index loads find it in L2 instead of waiting ~1 us.
      if (lane == 0 && i + A.pf_atoms < A.nlocal)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A.neigh2 + (size_t)(i + A.pf_atoms) * A.rowcap2),
                     "r"(A.pf_bytes) : "memory");
      const double4 pi = A.xq[i];
      const int packed = A.numneigh2[i];                // row length | type << 24 | "has an LJ partner" << 30 (prune kernel)
      const int ti = (packed >> 24) & 15;
      const bool has_lj = (packed >> 30) & 1;           // warp-uniform: water H (2/3 of the atoms) skips all LJ work
      const int n2pad = ((packed & 0xffffff) + 31) & ~31;
      const int *row2 = A.neigh2 + (size_t)i * A.rowcap2;
      const double qiq = pi.w * c.qqrd2e;
      const unsigned int coef_i = coef0 + (unsigned int)(ti * nt1) * 16u, cut_i = cut0 + (unsigned int)(ti * nt1) * 16u;
      Acc a;
      if (n2pad) {
        if (has_lj) {
          row_loop<STYLE, EFLAG, true, UNI>(c, A.xq, row2, n2pad, lane, pi, qiq, coef_i, cut_i, exp_tab, a);
        } else {
          row_loop<STYLE, EFLAG, false, UNI>(c, A.xq, row2, n2pad, lane, pi, qiq, coef_i, cut_i, exp_tab, a);
          a.fx *= qiq; a.fy *= qiq; a.fz *= qiq;
        }
        if (EFLAG) a.phi *= c.qqrd2e;
      }
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
        A.f[3 * (size_t)i] = a.fx;
        A.f[3 * (size_t)i + 1] = a.fy;
        A.f[3 * (size_t)i + 2] = a.fz;
        if (EFLAG) {
          const double ev = 0.5 * a.ev;
          const double ph = a.phi + 2.0 * pi.w * c.c_self;   // dE_coul/dq_i including the dsf self term
          A.evdwl[i] = ev;
          A.phi[i] = ph;
          A.eatom[i] = ev + 0.5 * pi.w * ph;
        }
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

__global__ void sum_int_kernel(int n, const int *__restrict__ v, unsigned long long *out2) {
  unsigned long long s = 0, sp = 0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int len = v[k] & 0xffffff;
    s += (unsigned long long)len;
    sp += (unsigned long long)((len + 31) & ~31);
  }
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sp += __shfl_xor_sync(0xffffffffu, sp, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out2, s); atomicAdd(out2 + 1, sp); }
}

}  // namespace

// Constants of the evaluation loop; they travel with every launch as a kernel parameter.
int cph_pair_fill_constants(cph_handle *h) {
  const PairParams &pp = h->pp;
  EvalConst c{};
  const double ln2 = 0.693147180559945309417232121458;
  c.cutsq = pp.cutsq_max;
  c.cut_coulsq = pp.cut_coulsq;
  c.qqrd2e = pp.qqrd2e;
  c.c_self = pp.c_self;
  c.e_shift = pp.e_shift;
  c.f_shift = pp.f_shift;
  c.a1 = 0.254829592; c.a2 = -0.284496736; c.a3 = 1.421413741; c.a4 = -1.453152027; c.a5 = 1.061405429;
  c.pa = 0.3275911 * pp.alpha;
  c.cD = 2.0 * pp.alpha / 1.77245385090551602729;
  c.neg_alpha2 = -pp.alpha * pp.alpha;
  c.exp_magic = 6755399441055744.0;   // 2^52 + 2^51
  if (pp.style == CPH_PAIR_LJ_CUT_COUL_DSF) {
    const double A = pp.alpha * pp.alpha;
    c.exp_scale = -A * 16.0 / ln2;
    c.exp_c1s = ln2 / (16.0 * A);
    double bk = 1.0;
    double *b[7] = {&c.b1, &c.b2, &c.b3, &c.b4, &c.b5, &c.b6, &c.b7};
    for (int k = 1; k <= 7; k++) { bk *= -A / k; *b[k - 1] = bk; }      // (-alpha^2)^k / k!
  }
  for (int k = 0; k < 4; k++) {
    c.flj[k] = pp.special_lj[k];
    c.fcoul[k] = pp.special_coul[k];
    c.one_m_fc[k] = 1.0 - pp.special_coul[k];
  }
  h->eval_const = c;
  if (!h->d_exp2.p) {
    double e2[16];
    for (int j = 0; j < 16; j++) e2[j] = (double)exp2l((long double)j / 16.0L);
    CPH_CUDA(h, h->d_exp2.reserve(16));
    CPH_CUDA(h, cudaMemcpyAsync(h->d_exp2.p, e2, sizeof(e2), cudaMemcpyHostToDevice, h->stream));
    CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  h->kc_dirty = false;
  return 0;
}

// fp32 prefilter records of every atom (owned, ghost, dummy); positions relative to the grid origin
int cph_launch_xt(cph_handle *h) {
  const double3 origin = make_double3(h->grid.lo[0], h->grid.lo[1], h->grid.lo[2]);
  CPH_CUDA(h, h->d_xt.reserve((size_t)h->nall + 2));
  h->nlaunch++;
  xt_kernel<<<(h->nall + 256) / 256, 256, 0, h->stream>>>(h->nall, h->d_xq.p, h->d_type.p, origin, h->d_xt.p);
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

// K2a launcher: prune the Verlet rows to rc + inner skin and remember where the atoms were
int cph_launch_prune(cph_handle *h) {
  const int n = h->nlocal;
  if (n == 0) { h->inner_valid = true; return 0; }
  ProfScope ps(h, 1);                       // one slot for the fp32 record refresh + the prune + the snapshot
  CPH_TRY(cph_launch_xt(h));
  double extent = 0;
  for (int k = 0; k < 3; k++) extent = std::max(extent, h->grid.n[k] / h->grid.inv[k]);
  const double cut = std::sqrt(h->pp.cutsq_max) + h->inner_skin;
  const float cutf = (float)(cut * cut + 32.0 * cut * extent * 5.97e-8 + 1e-5 * cut * cut);
  // an inner row holds at most what its outer row holds (+ padding to 32)
  h->rowcap2 = ((h->maxneigh + 31) & ~31) + 64;   // ordinary + special entries, padding of the last ordinary chunk, one more chunk
  if (h->nall + 1 > CPH_JMASK)
    return cph_fail(h, CPH_ERR_OVERFLOW, "%d atoms and ghosts on this rank exceed the 26-bit indices of the inner rows", h->nall);
  CPH_CUDA(h, h->d_neigh2.reserve((size_t)n * h->rowcap2));
  CPH_CUDA(h, h->d_numneigh2.reserve(n + 1));
  CPH_CUDA(h, h->d_xinner.reserve(3 * (size_t)n + 3));
  if (h->rowcap / CH * APW > MAXTILES)
    return cph_fail(h, CPH_ERR_OVERFLOW, "neighbour rows of %d entries exceed the prune kernel's tile schedule", h->rowcap);
  h->nlaunch += 2;
  const int pblocks = (n + APB - 1) / APB;
  if (!h->lj_states) {
    prune_kernel<false><<<pblocks, TPB, 0, h->stream>>>(n, h->d_xt.p, h->d_neigh.p, h->d_numneigh.p, h->d_numspec.p,
                                                        h->rowcap, cutf, h->nall, h->d_type_has_lj.p, h->d_neigh2.p,
                                                        h->rowcap2, h->d_numneigh2.p, PruneEs{});
  } else {
    // LJ end states: the correction lists are written by the same pass; a list that outgrows its slots is the
    // only reason to run it again (the capacity then fits the largest list seen)
    for (int attempt = 0; attempt < 2; attempt++) {
      PruneEs es;
      CPH_TRY(cph_ljstates_lists(h, &es.tB, &es.tmask, &es.cnt, &es.ent, &es.over, &es.cap));
      prune_kernel<true><<<pblocks, TPB, 0, h->stream>>>(n, h->d_xt.p, h->d_neigh.p, h->d_numneigh.p, h->d_numspec.p,
                                                         h->rowcap, cutf, h->nall, h->d_type_has_lj.p, h->d_neigh2.p,
                                                         h->rowcap2, h->d_numneigh2.p, es);
      int over = 0;
      CPH_CUDA(h, cudaMemcpyAsync(&over, es.over, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CPH_CUDA(h, cudaStreamSynchronize(h->stream));
      if (over <= h->es_cap) break;
      if (attempt == 1) return cph_fail(h, CPH_ERR_OVERFLOW, "LJ end-state lists grew past %d entries twice", over);
      h->es_cap = (over + 7) & ~7;
      h->nlaunch++;
    }
  }
  snapshot_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(n, h->d_xq.p, h->d_xinner.p);
  CPH_CUDA(h, cudaGetLastError());
  h->inner_valid = true;
  h->nprunes++;
  return 0;
}

int cph_launch_pair(cph_handle *h, int eflag, const unsigned int *gate) {
  const int n = h->nlocal;
  if (n == 0) return 0;
  if (gate && !h->inner_valid) return cph_fail(h, CPH_ERR_STATE, "a gated pair pass needs valid inner rows");
  if (h->kc_dirty) CPH_TRY(cph_pair_fill_constants(h));
  if (!h->inner_valid) CPH_TRY(cph_launch_prune(h));
  ProfScope ps(h, 0);
  static const int ctas_env = getenv("CPH_EVAL_CTAS_PER_SM") ? atoi(getenv("CPH_EVAL_CTAS_PER_SM")) : 0;
  EvalArgs A;
  A.c = h->eval_const;
  A.nlocal = n; A.nt1 = h->pp.ntypes + 1; A.rowcap2 = h->rowcap2;
  A.xq = h->d_xq.p;
  A.neigh2 = h->d_neigh2.p; A.numneigh2 = h->d_numneigh2.p; A.coef = h->d_coef4.p; A.cuts = h->d_cut2.p;
  A.exp2 = h->d_exp2.p;
  A.f = h->d_f.p; A.evdwl = h->d_evdwl.p; A.phi = h->d_phi.p; A.eatom = h->d_eatom.p; A.gate = gate;
  // persistent grid: as many 64-thread CTAs as are resident at once (16 per SM at <= 64 registers), but no more
  // warps than atoms; the per-SM queue heads are cleared in front of every launch
  const int per_sm = ctas_env > 0 ? ctas_env : 16;
  const int blocks = std::max(1, std::min(h->num_sms * per_sm, (n + EWARPS - 1) / EWARPS));
  A.nqueues = h->num_sms;
  {
    // bytes of a typical inner row (from the mean Verlet row and the two radii), a little generously, whole 128-byte lines
    static const int pf_env = getenv("CPH_EVAL_PREFETCH") ? atoi(getenv("CPH_EVAL_PREFETCH")) : -1;
    const double rc = std::sqrt(h->pp.cutsq_max), ratio = (rc + h->inner_skin) / (rc + h->skin);
    const double mean_inner = (double)h->stored_neigh / std::max(1, n) * ratio * ratio * ratio;
    const int entries = std::min(h->rowcap2, (int)(mean_inner * 1.1) + 64);
    A.pf_bytes = (unsigned int)((entries * 4 + 127) / 128 * 128);
    A.pf_atoms = pf_env >= 0 ? pf_env : 2 * EWARPS * per_sm;     // two rounds of the SM's warps ahead
    if (A.pf_atoms == 0) A.pf_atoms = 1 << 30;                   // CPH_EVAL_PREFETCH=0: off
  }
  CPH_CUDA(h, h->d_qnext.reserve(h->num_sms));
  A.qnext = h->d_qnext.p;
  CPH_CUDA(h, cudaMemsetAsync(h->d_qnext.p, 0, h->num_sms * sizeof(int), h->stream));
  h->nlaunch++;
  static const int carve_env = getenv("CPH_EVAL_CARVEOUT") ? atoi(getenv("CPH_EVAL_CARVEOUT")) : -1;
  static const int scan_env = getenv("CPH_EVAL_MAXSCAN") ? atoi(getenv("CPH_EVAL_MAXSCAN")) : 0;
  A.maxscan = std::min(A.nqueues, scan_env > 0 ? scan_env : 4);
  EvalArgs B = A;                                   // the sweep: one CTA per range, bound to it
  B.sweep = 1;
  A.sweep = 0;
  h->nlaunch++;
#define LAUNCH(S, E, U)                                                                                               \
  do {                                                                                                                \
    if (carve_env >= 0)                                                                                               \
      cudaFuncSetAttribute(pair_eval_kernel<S, E, U>, cudaFuncAttributePreferredSharedMemoryCarveout, carve_env);     \
    pair_eval_kernel<S, E, U><<<blocks, ETPB, 0, h->stream>>>(A);                                                     \
    pair_eval_kernel<S, E, U><<<A.nqueues, ETPB, 0, h->stream>>>(B);                                                  \
  } while (0)
#define LAUNCH_E(S, U) do { if (eflag) LAUNCH(S, 1, U); else LAUNCH(S, 0, U); } while (0)
  if (h->pp.style == CPH_PAIR_LJ_CUT_COUL_CUT) {
    if (h->uniform_cut) LAUNCH_E(CPH_PAIR_LJ_CUT_COUL_CUT, true); else LAUNCH_E(CPH_PAIR_LJ_CUT_COUL_CUT, false);
  } else {
    if (h->uniform_cut) LAUNCH_E(CPH_PAIR_LJ_CUT_COUL_DSF, true); else LAUNCH_E(CPH_PAIR_LJ_CUT_COUL_DSF, false);
  }
#undef LAUNCH_E
#undef LAUNCH
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

// out[0] = inner-row entries (sum of numneigh2), out[1] = the same padded to 32 per row (= 32 x loop trips of K2b)
int cph_inner_counts(cph_handle *h, int64_t *out2) {
  out2[0] = out2[1] = 0;
  if (!h->inner_valid || h->nlocal == 0) return 0;
  CPH_CUDA(h, h->d_scr_stats.reserve(4));
  CPH_CUDA(h, cudaMemsetAsync(h->d_scr_stats.p + 2, 0, 2 * sizeof(unsigned long long), h->stream));
  sum_int_kernel<<<256, 256, 0, h->stream>>>(h->nlocal, h->d_numneigh2.p, h->d_scr_stats.p + 2);
  unsigned long long v[2];
  CPH_CUDA(h, cudaMemcpyAsync(v, h->d_scr_stats.p + 2, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  out2[0] = (int64_t)v[0];
  out2[1] = (int64_t)v[1];
  return 0;
}
