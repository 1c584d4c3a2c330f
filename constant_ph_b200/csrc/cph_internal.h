// cph_internal.h -- handle layout and kernel entry points of libcph_b200.so (sm_100a only).
// Not part of the ABI; the ABI is include/cph_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/cph_b200.h"

#define CPH_NEIGHMASK 0x1FFFFFFF   // LAMMPS NEIGHMASK: low 29 bits = atom index
#define CPH_SBSHIFT 30             // LAMMPS SBBITS: top 2 bits = special-bond class (outer rows)
#define CPH_TYPESHIFT 28           // inner rows: type of j in the top 4 bits, special-bond class in bits 26-27, index in the low 26
#define CPH_SB2SHIFT 26
#define CPH_JMASK 0x03FFFFFF
#define CPH_MAIL_MAXP 8             // peer mailboxes serve up to 8 ranks (one NVSwitch domain); beyond that NCCL
#define CPH_MAIL_FSLOT 16           // 32-bit words per flag slot: [0..5] flags, [8..9] sequence number
#define CPH_MAXNT1 12              // ntypes+1 <= 12: the (ntypes+1)^2 * 32 B coefficient table lives in shared memory

// device buffer that only ever grows
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t reserve(size_t n, bool keep = false, cudaStream_t s = 0) {
    if (n <= cap) return cudaSuccess;
    size_t ncap = n + n / 8 + 64;
    T *np_ = nullptr;
    cudaError_t e = cudaMalloc((void **)&np_, ncap * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep && p && cap) {
      e = cudaMemcpyAsync(np_, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) return e;
      cudaStreamSynchronize(s);
    }
    if (p) cudaFree(p);
    p = np_;
    cap = ncap;
    return cudaSuccess;
  }
  // exactly n elements (no head room): used where several buffers must share one capacity
  cudaError_t reserve_exact(size_t n, bool keep = false, cudaStream_t s = 0) {
    if (n <= cap) return cudaSuccess;
    T *np_ = nullptr;
    cudaError_t e = cudaMalloc((void **)&np_, n * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep && p && cap) {
      e = cudaMemcpyAsync(np_, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) return e;
      cudaStreamSynchronize(s);
    }
    if (p) cudaFree(p);
    p = np_;
    cap = n;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  size_t bytes() const { return cap * sizeof(T); }
};

// per type pair: lj3, lj4, cut_ljsq, cutsq (max of lj and coul cutoffs) -- 32 B, one LDS.128 x2
struct PairCoef {
  double lj3, lj4, cut_ljsq, cutsq;
};

struct PairParams {
  int style, ntypes;
  double qqrd2e, alpha, cut_coulsq, cutsq_max, e_shift, f_shift, c_self;
  double special_lj[4], special_coul[4];
};

// constants of the evaluation loop (pair.cu); passed to the kernel by value with every launch
struct EvalConst {
  double cutsq, cut_coulsq;            // global cutoff^2 (max over type pairs), Coulomb cutoff^2
  double exp_scale, exp_magic, exp_c1s;   // -alpha^2 16/ln2, 2^52+2^51, ln2/(16 alpha^2)
  double b1, b2, b3, b4, b5, b6, b7;   // (-alpha^2)^k / k!
  double pa, a1, a2, a3, a4, a5;       // EWALD_P alpha and LAMMPS' erfc polynomial
  double cD, e_shift, f_shift;         // 2 alpha/sqrt(pi), dsf shifts
  double qqrd2e, c_self;
  double neg_alpha2;                   // slow path only
  double one_m_fc[4], flj[4], fcoul[4];
};

struct EvalArgs {
  EvalConst c;
  int nlocal, nt1, rowcap2, nqueues, pf_atoms, maxscan, sweep;
  unsigned int pf_bytes;               // L2 prefetch of the inner rows: how far ahead, how many bytes
  int *qnext;                          // per-SM queue heads (pair.cu)
  const double4 *xq;
  const int *neigh2, *numneigh2;
  const double4 *coef;
  const double2 *cuts;
  const double *exp2;
  double *f, *evdwl, *phi, *eatom;
  const unsigned int *gate;
};

struct BiasParams {
  double w, s, h, k, a, b, r, m, d, m_lambda;
  int mode;
};

struct FixParams {
  int nevery, Hbit, Wbit;
  double pK, pH, T, boltz, ftm2v;
  int dudl_mode, integ_mode, fscale_mode, implicit_site;
};

struct Grid {         // cell grid over the extended sub-box
  double lo[3];       // origin (sublo - ghost cutoff)
  double inv[3];      // 1 / cell width
  int n[3];           // cells per dim
  int ncell;
};

// byte offsets inside a mailbox
inline size_t mail_flag_off(int P, int parity, int src) { return ((size_t)parity * P + src) * CPH_MAIL_FSLOT * 4; }
inline size_t mail_red_base(int P) { return ((size_t)2 * P * CPH_MAIL_FSLOT * 4 + 255) / 256 * 256; }
inline size_t mail_red_off(int P, size_t cap, int parity, int src) {
  return mail_red_base(P) + ((size_t)parity * P + src) * (cap + 2) * sizeof(double);
}
// third region: 32-int blocks for the all-gather of the list rebuild (send counts per direction), sequence number behind
#define CPH_MAIL_GSLOT 40            // ints per gather slot: [0..31] payload, [32..33] sequence number
inline size_t mail_gather_off(int P, size_t cap, int parity, int src) {
  return mail_red_off(P, cap, 2, 0) + ((size_t)parity * P + src) * CPH_MAIL_GSLOT * sizeof(int);
}
inline size_t mail_bytes(int P, size_t cap) { return mail_gather_off(P, cap, 2, 0); }

// where this rank's site-sum block goes (one slot per destination rank), and where the blocks of all ranks arrive
struct MailRed {
  int P = 0;                         // 0: mailboxes off (single rank or NCCL fallback)
  int seq_index = 0;                 // the sequence number sits at doubles[seq_index] of a slot
  unsigned long long seq = 0;
  double *dst[CPH_MAIL_MAXP];        // push side: my slot in rank p's mailbox
  const double *src[CPH_MAIL_MAXP];  // gather side: rank p's slot in MY mailbox
};

struct ProfSlot {
  double ms = 0;
  int64_t launches = 0;
};

struct NcclApi;  // comm.cu

struct cph_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;    // side stream: decision flags to the host while the halo runs
  cudaEvent_t ev_flags = nullptr;
  cudaEvent_t ev_force = nullptr;    // forces final: their copy to the host overlaps the lambda tail of the step
  cudaEvent_t ev_xstage = nullptr;   // the staged upload of x has left the page-locked staging buffer
  cudaEvent_t ev_fchunk[4]{};        // pieces of the force read-back
  bool force_add = false;            // host forces are ADDED to the caller's array (cph_set_force_mode)
  double *h_fpin = nullptr;          // page-locked staging of the force read-back
  size_t h_fpin_bytes = 0;
  DevBuf<double> d_xstage;           // device-side landing buffer of x (caller order)
  unsigned int *h_flags = nullptr;   // pinned
  std::string err;
  // configuration
  PairParams pp{};
  std::vector<PairCoef> coef_h;
  double cut_lj_max = 0, cut_coul = 0, skin = 2.0;
  bool have_pair = false, have_domain = false, have_atoms = false, have_pass = false, have_sites = false;
  double boxlo[3]{}, boxhi[3]{}, sublo[3]{}, subhi[3]{};
  int periodic[3]{1, 1, 1}, procgrid[3]{1, 1, 1}, myloc[3]{0, 0, 0};
  BiasParams bias{200.0, 0.3, 4.0, 2.533, 0.034041, 0.005238, 16.458, 0.1507, 2.0, 20.0, 0};
  FixParams fix{1, 0, 0, 0.0, 7.0, 300.0, 0.0019872067, 1.0, 1, 0, 0, 1};
  double qqrd2e = 332.06371;
  double extra_HA = 0.0, extra_HB = 0.0;   // host-tallied energy sources for the next site reduce (cpp:221-244)
  DevBuf<double> d_extra_dudl;             // their per-site dE/dlambda (host KSpace, cpp:241-244), this rank's share
  bool extra_dudl = false;                 // set for the next site reduce, then cleared
  // sites (host copies, site-major order)
  int S = 1, ntitr = 0;
  std::vector<int> titr_tag_sorted_h;  // titr tags sorted ascending (for the device binary search)
  std::vector<int> titr_entry_of_sorted_h;
  std::vector<int> titr_order_h;       // site-major entry k = caller's entry titr_order_h[k] of cph_set_sites
  // LJ end states (ljstates.cu): B-state type per titration entry; per-atom views; CSR list of the inner-row
  // entries that touch an atom with end states; g = dE/dlambda carried by each owned atom's own end states
  bool lj_states = false;
  DevBuf<int> d_titr_typeB, d_es_tB, d_es_site, d_es_cnt, d_es_ent, d_es_own, d_es_word;
  DevBuf<unsigned int> d_es_tmask;     // bit t set: some owned or ghost atom of (A-state) type t has end states
  DevBuf<double> d_es_g;
  int es_cap = 8;                      // slots per atom in the correction lists (grown to the largest list seen)
  int es_nown = 0;                     // owned atoms that have end states themselves
  // device: sites
  DevBuf<double> d_pK, d_lam, d_vlam, d_alam, d_flam, d_fs, d_dfs, d_Us, d_dUs;
  DevBuf<double> d_theta;            // dynamical coordinate when coord_theta (lambda = sin^2 theta)
  bool coord_theta = false;
  double nh_tau = 0.0;               // Nose-Hoover period of the lambda thermostat (0 = off); xi, eta in d_scal[8], [10]
  // one contiguous reduction buffer so a single allreduce covers everything (cpp:274):
  // [0]=HA [1]=HB [2]=E_vdwl [3]=E_coul [4..4+S)=dU/dlambda_s [4+S..4+2S)=HB_s-HA_s
  // [4+2S]=sum of dE/dq over the owned water-buffer atoms (modify_water)
  DevBuf<double> d_red;
  DevBuf<PairCoef> d_coef;
  // modify_water (h:58): charge buffer on the water group
  int water_n = 0;                   // atoms in the water group (0 = buffer off)
  int nw_local = 0;                  // of which this rank owns
  std::vector<int> wtag_h;           // tags of the owned buffer atoms
  std::vector<double> wq_h;          // their charges as supplied (lambda = 0 state)
  DevBuf<int> d_wtag, d_wlocal;
  DevBuf<double> d_wq, d_dQ;         // d_dQ[s] = sum_{t in s} (qB - qA)
  DevBuf<double4> d_coef4;      // {12 lj3, 6 lj4, lj3, lj4} per type pair
  DevBuf<double2> d_cut2;       // {cut_ljsq, cutsq} per type pair
  DevBuf<int> d_type_has_lj;    // type i has a non-zero LJ partner
  bool uniform_cut = true, kc_dirty = true;
  DevBuf<int> d_titr_tag_sorted, d_titr_entry_of_sorted;  // [ntitr]
  DevBuf<int> d_titr_site, d_titr_local;                   // [ntitr] site-major; local = owned index or -1
  int site_lps = 32;                                       // lanes per site in the site-sum kernel (power of two)
  DevBuf<int> d_site_start;                                // [S+1] range of every site in the site-major table
  DevBuf<double> d_titr_qA, d_titr_dq;                     // [ntitr]
  DevBuf<double> d_scal;                                   // 16 doubles of scalar results
  DevBuf<double> d_part;                                   // block partials for deterministic sums
  // device: atoms in internal (cell-sorted) order; owned [0,nlocal), ghosts [nlocal,nall)
  int nlocal = 0, nghost = 0, nall = 0, maxspecial = 0;
  size_t atom_cap = 0;      // common capacity of every per-atom buffer (owned + ghost + dummy)
  DevBuf<double4> d_xq;
  DevBuf<float4> d_xt;      // fp32 {x-origin, y-origin, z-origin, type}: prefilter record
  DevBuf<float4> d_xb;      // fp32 {x-origin, y-origin, z-origin, molecule id}: list-build record
  DevBuf<int> d_molecule;   // caller order, optional
  bool have_mol = false;
  DevBuf<int> d_type, d_tag, d_mask;
  DevBuf<int> d_perm;       // internal -> caller index   [nlocal]
  DevBuf<int> d_inv;        // caller -> internal index   [nlocal]
  DevBuf<int> d_site_of;    // [nlocal] internal order
  DevBuf<int> d_titr_of;    // [nlocal] internal order, entry in site-major titr arrays
  DevBuf<int> d_nspecial, d_special;  // caller order, as uploaded
  DevBuf<double> d_xbuild;  // [3*nlocal] positions at list build
  DevBuf<int> d_ghost_src;  // [nghost] owned internal index (self image) or -1-slot in the receive buffer
  DevBuf<int> d_ghost_code; // [nghost] direction 0..26 (self image) or 32+image code (received copy)
  DevBuf<int> d_mol;        // molecule id, internal order, owned + ghost (optional)
  // halo (K6): records (owner, direction) sorted by direction; send/receive staging
  int nrec = 0, nsend = 0, nrecv = 0;
  int rec_start[28]{}, send_count[27]{}, send_off[27]{}, recv_count[27]{}, recv_off[27]{};
  DevBuf<int> d_rec_src, d_rec_dir;
  std::vector<int> peer_rank, peer_soff, peer_scnt, peer_roff, peer_rcnt;   // one message per neighbour rank
  DevBuf<double4> d_sendx, d_recvx;   // d_recvx holds TWO halves of recv_half records (alternating per step)
  size_t recv_half = 0;
  int halo_parity = 0;
  // peer-memory halo: neighbours' receive buffers mapped through CUDA IPC; the pack kernel stores
  // {x,y,z,q} straight into them over NVLink (no ncclSend/ncclRecv on the per-step path)
  bool peer_halo = false, peer_halo_wanted = true;
  std::vector<void *> peer_base;                 // [nranks] mapped base of every neighbour's d_recvx
  std::vector<unsigned char> peer_handle_cache;  // [nranks * 64] handle each mapping was opened from
  std::vector<int> peer_table;                   // [nranks * 28] recv_off[27] + recv_half of every rank
  DevBuf<unsigned char> d_ipc_stage;
  // Peer mailboxes: a one-shot all-reduce over NVLink for the two small per-step reductions (decision flags at the
  // start of a step, site sums before the lambda update).  Every rank stores its block straight into a slot of
  // every other rank's mailbox (mapped through CUDA IPC like the halo buffers) and publishes a sequence number;
  // the consuming kernel waits for all sequence numbers and combines the blocks in rank order, so every rank gets
  // bit-identical totals without NCCL and without a host round trip.  Two parities: a writer can be at most one
  // reduction ahead of the slowest reader.
  DevBuf<unsigned char> d_mail;
  size_t mail_red_cap = 0;                       // doubles per site-sum slot (the sequence number sits behind them)
  bool mail_ok = false;                          // every rank's mailbox is mapped on every rank
  bool mail_wanted = true;                       // CPH_MAIL=0: keep the two NCCL all-reduces (A/B)
  std::vector<void *> mail_base;                 // [nranks] mapped base of every rank's mailbox (own entry: d_mail.p)
  std::vector<unsigned char> mail_handle_cache;  // [nranks * 64]
  unsigned long long seq_flags = 0, seq_red = 0, seq_gather = 0; // reductions / gathers published so far
  bool red_pending = false;                      // site sums pushed; the totals are gathered by the next consumer
  DevBuf<int4> d_sendmeta, d_recvmeta;
  DevBuf<double> d_f, d_evdwl, d_phi, d_eatom;  // [3*nlocal], [nlocal]...
  DevBuf<int> d_hlist;      // owned atoms in the hydrogen group
  int nh = 0;
  // staging
  DevBuf<double> d_stage;   // H2D / D2H staging for caller-order arrays
  DevBuf<int> d_istage;
  DevBuf<unsigned long long> d_keys, d_keys2;
  DevBuf<int> d_vals, d_vals2, d_tmpi;
  DevBuf<int> d_scr_i, d_scr_src, d_scr_code, d_scr_off;   // persistent rebuild scratch (no malloc/free per rebuild)
  DevBuf<unsigned long long> d_scr_stats;
  DevBuf<double4> d_xq2;
  DevBuf<unsigned char> d_cubtmp;
  double *h_pin = nullptr;  // pinned host scratch
  size_t h_pin_bytes = 0;
  // cells + list
  Grid grid{};
  double ghost_cut = 0;
  DevBuf<int> d_cell_start_o, d_cell_start_g;
  // row i: [0,numneigh) ordinary neighbours, padded to a 128 multiple with the dummy atom;
  // special-bond partners (entry = j | class<<30) at the END of the row, numspec of them
  DevBuf<int> d_neigh, d_numneigh, d_numspec;
  // inner rows (rolling prune): pairs within rc + inner_skin, entry = j | type_j<<28
  DevBuf<int> d_neigh2, d_numneigh2;
  DevBuf<double> d_xinner;           // positions at the last prune
  double inner_skin = 0.4;           // measured best of 0.3..0.8 at config 3 (CPH_INNER_SKIN overrides)
  bool inner_valid = false;
  int rowcap2 = 0;                   // pitch of the inner rows
  EvalConst eval_const{};
  DevBuf<double> d_exp2;             // 2^(j/16), staged into shared memory by the evaluation kernel
  DevBuf<int> d_qnext;               // per-SM queue heads of the evaluation kernel
  int num_sms = 148;
  bool speculate = true;        // enqueue the pair pass before the host has read the list flags (CPH_SPECULATE=0: off)
  int64_t nprunes = 0;
  bool drift_known = false;          // the step's all-reduced flags already carry the drift beyond the sub-box
  float drift_value = 0.f;
  int rowcap = 0;
  int64_t nbuilds = 0, stored_neigh = 0, special_pairs = 0;
  int64_t nlaunch = 0;               // kernels of this library launched so far
  int maxneigh = 0;
  DevBuf<unsigned int> d_flags;  // [0] max displacement^2 as float bits, [1] overflow, [2] drift...
  // scalars mirrored on host after site_reduce / integrate
  double scal_h[16]{};
  // f2: bonded terms of flexible molecules (bond_style harmonic, angle_style harmonic) + fix-nve atom dynamics
  int nbondtypes = 0, nangletypes = 0, maxbond = 0, maxangle = 0;
  bool have_bonded_coef = false, have_topology = false, md_on = false;
  bool drop_excluded = false;        // cph_set_excluded_policy: dsf drops fully excluded specials like the cut styles
  int last_dropmask = 0;             // special classes the last list build left out of the rows
  DevBuf<double2> d_bond_coef, d_angle_coef;          // {K, r0} / {K, theta0} by type
  DevBuf<int> d_num_bond, d_bond_type, d_bond_atom;   // caller order, as uploaded (partner ids are tags)
  DevBuf<int> d_num_angle, d_angle_type, d_angle_a1, d_angle_a2, d_angle_a3;
  DevBuf<int> d_bcount;              // internal order: bonds | angles << 8 of the atom
  DevBuf<int> d_bond_j, d_bond_t;    // internal order [nlocal * maxbond]: partner index (owned or ghost), type
  DevBuf<int> d_angle_j;             // internal order [nlocal * maxangle * 2]: indices of the other two atoms
  DevBuf<int> d_angle_t;             // type | role << 16 (role 0/2: an end, 1: the centre)
  DevBuf<double> d_bonded_e;         // [2] E_bond, E_angle of the owned shares
  double mass_h[CPH_MAXNT1]{};
  DevBuf<double3> d_v, d_v2;         // internal order velocities (owned atoms)
  // f4: kspace_style ewald (kspace.cu); pair style CPH_PAIR_LJ_CUT_COUL_LONG is its real-space part
  int kspace_style = 0;
  double g_ewald = 0.0, kspace_self2 = 0.0, kspace_bg = 0.0;
  int kmax[3]{0, 0, 0};
  int nkvec = 0;                     // half-space wave vectors (the buffer holds one more: the zero vector)
  bool coul_long = false;            // the damped kernel runs without shifts and without its pair-level self term
  DevBuf<double4> d_kvec;            // {kx, ky, kz, ug}
  DevBuf<int> d_kidx;                // nx | (ny+512)<<10 | (nz+512)<<20 (factorised kernels)
  double kspace_unitk[3]{0, 0, 0};   // 2 pi / L_d
  bool kspace_fact = true;           // factorised kernels (CPH_EWALD=direct: one sincos per atom and wave vector)
  int kspace_tile = 32;              // atoms per shared-memory tile of the factorised structure-factor kernel
  int kspace_tune[5]{256, 6, 256, 32, 6};   // launch shapes of the k-space kernels (kspace.cu, CPH_EWALD_TUNE)
  bool kspace_rows = true;           // per-atom sums by the row-walking kernel (CPH_EWALD=tables: the table kernel)
  int kspace_nrows = 0;
  DevBuf<int4> d_krows;              // (nx, ny) rows of the wave-vector list
  DevBuf<double4> d_kpart;           // [slice][atom] partial {pot, fx, fy, fz} of the row-walking kernel
  DevBuf<double2> d_sfac_part, d_sfac;   // structure factors: chunk partials, totals
  DevBuf<double> d_ekspace;          // [nlocal] per-atom k-space energy, [nlocal] their sum
  // comm
  int nranks = 1, rank = 0;
  void *nccl_comm = nullptr;
  // timing
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool profiling = false;
  ProfSlot prof[12];                 // 0..7 per kernel class, 10 / 11 k-space (8 and 9 are counters, see cph_profile_get)
  cudaEvent_t pev0 = nullptr, pev1 = nullptr;
};

// ---- kernels (launchers) ----------------------------------------------------------------
// neigh.cu
int cph_rebuild(cph_handle *h);                 // sort, ghosts, cells, list, site map
int cph_forward_ghosts(cph_handle *h);          // refresh ghost x and q (send + barrier if needed + finish)
int cph_flags_allreduce(cph_handle *h, unsigned int *dst);   // max over ranks of this step's decision flags -> dst[0..5] (mailboxes or NCCL)
int cph_halo_send(cph_handle *h);               // pack + ship the copies (peer stores over NVLink, or NCCL p2p)
int cph_halo_finish(cph_handle *h);             // ghost atoms from self images + received copies
void cph_halo_close(cph_handle *h);
void cph_mail_close(cph_handle *h);
int cph_neighbors_to_host(cph_handle *h, int *numneigh, int64_t *keys, int64_t cap);
// pair.cu
int cph_launch_pair(cph_handle *h, int eflag, const unsigned int *gate = nullptr);
int cph_launch_prune(cph_handle *h);
int cph_pair_fill_constants(cph_handle *h);
int cph_inner_counts(cph_handle *h, int64_t *out2);
int cph_launch_xt(cph_handle *h);
// ljstates.cu
int cph_ljstates_set(cph_handle *h, int ntitr, const int *typeB);
int cph_ljstates_map(cph_handle *h);            // after a list build
// buffers the prune kernel fills when end states are declared (sized for the current atom count and capacity)
int cph_ljstates_lists(cph_handle *h, const int **tB, const unsigned int **tmask, int **cnt, int **ent, int **over,
                       int *cap);
int cph_launch_ljstates(cph_handle *h, int eflag);   // after the pair pass: adds the end-state difference
// sites.cu
int cph_launch_partition(cph_handle *h, bool push = false);   // HA, HB, E_vdwl, E_coul + per-site sums (+ push to the mailboxes)
bool cph_mail_red_usable(const cph_handle *h);
int cph_launch_red_gather(cph_handle *h);
int cph_launch_integrate(cph_handle *h, double dt, int phase, bool apply = false);
int cph_launch_apply_charges(cph_handle *h);
int cph_launch_water_phi(cph_handle *h);      // red[4+2S] = sum of dE/dq over owned buffer atoms
int cph_launch_water_dudl(cph_handle *h);     // dU/dlambda_s -= dQ_s / n_W * red[4+2S] (after the allreduce)
int cph_launch_set_force(cph_handle *h);
int cph_launch_set_x(cph_handle *h, const double *x_dev_caller_order);
int cph_launch_gather_out(cph_handle *h, int what, double *out_dev, cudaStream_t on = nullptr);  // 0 f, 1 eatom, 2 phi, 3 q, 4 x, 5 v
int cph_launch_pack_xq(cph_handle *h, int n, const double *x, const double *q);
// bonded.cu
int cph_bonded_resolve(cph_handle *h);          // partner tags -> indices in the current internal order (after a list build)
int cph_launch_bonded(cph_handle *h, int eflag);   // adds bond + angle forces / per-atom energy to the pair results
int cph_md_wrap(cph_handle *h);                 // remap self-propelled atoms into the periodic box (before re-neighbouring)
void cph_bonded_release(cph_handle *h);
int cph_bonded_set_coef(cph_handle *h, int nbondtypes, const double *bk, const double *br0, int nangletypes,
                        const double *ak, const double *at0);
int cph_bonded_set_topology(cph_handle *h, int nlocal, int maxbond, const int *num_bond, const int *bond_type,
                            const int *bond_atom, int maxangle, const int *num_angle, const int *angle_type,
                            const int *a1, const int *a2, const int *a3);
int cph_bonded_energy(cph_handle *h, double *out2);
int cph_md_set_v(cph_handle *h, int where, const double *v_caller_order);
int cph_md_kick(cph_handle *h, double dt, int drift);
// kspace.cu
int cph_kspace_setup(cph_handle *h);            // wave vectors for the current box
int cph_launch_kspace(cph_handle *h, int eflag);   // after the pair pass: adds the reciprocal Ewald sum
int cph_kspace_energy(cph_handle *h, double *out);
void cph_kspace_release(cph_handle *h);
// comm.cu
int cph_comm_allreduce(cph_handle *h, double *buf, int n);            // sum
int cph_comm_allreduce_max_u32(cph_handle *h, unsigned int *buf, int n);
int cph_comm_allreduce_max_u32_dev(cph_handle *h, unsigned int *dbuf, int n);   // in place, on the stream
int cph_comm_exchange(cph_handle *h, int npeers, const int *peers, const void *const *sendbuf, const size_t *sendbytes,
                      void *const *recvbuf, const size_t *recvbytes);
int cph_comm_allgather(cph_handle *h, const void *sendbuf, void *recvbuf, size_t bytes_per_rank);
int cph_comm_exchange_counts(cph_handle *h, const int *active, const int *peer, const int *from, const int *send_count,
                             int *recv_count);
void cph_comm_destroy(cph_handle *h);

int cph_fail(cph_handle *h, int code, const char *fmt, ...);

#define CPH_CUDA(h, call)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return cph_fail(h, CPH_ERR_CUDA, "%s at %s:%d: %s", #call, __FILE__, __LINE__,        \
                      cudaGetErrorString(e__));                                             \
  } while (0)

#define CPH_TRY(call)          \
  do {                         \
    int rc__ = (call);         \
    if (rc__ != 0) return rc__; \
  } while (0)

struct ProfScope {
  cph_handle *h;
  int which;
  ProfScope(cph_handle *h_, int w) : h(h_), which(w) {
    if (h->profiling) cudaEventRecord(h->pev0, h->stream);
  }
  ~ProfScope() {
    if (h->profiling) {
      cudaEventRecord(h->pev1, h->stream);
      cudaEventSynchronize(h->pev1);
      float ms = 0;
      cudaEventElapsedTime(&ms, h->pev0, h->pev1);
      h->prof[which].ms += ms;
      h->prof[which].launches += 1;
    }
  }
};
