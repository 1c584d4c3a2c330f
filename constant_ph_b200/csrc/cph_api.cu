// cph_api.cu -- the C ABI of include/cph_b200.h: argument checking, host orchestration,
// host<->device staging.  Every entry point cites the reference interface it replaces in
// the header.  No CPU fallback exists: without a CUDA device cph_create fails.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include <omp.h>

#include "cph_internal.h"

static thread_local std::string g_create_error;

int cph_fail(cph_handle *h, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

namespace {

int need(cph_handle *h, bool cond, const char *what) {
  if (!cond) return cph_fail(h, CPH_ERR_STATE, "%s", what);
  return 0;
}

int ensure_pinned(cph_handle *h, size_t bytes) {
  if (bytes <= h->h_pin_bytes) return 0;
  if (h->h_pin) cudaFreeHost(h->h_pin);
  h->h_pin = nullptr;
  h->h_pin_bytes = 0;
  CPH_CUDA(h, cudaMallocHost((void **)&h->h_pin, bytes + bytes / 8));
  h->h_pin_bytes = bytes + bytes / 8;
  return 0;
}

template <typename T>
int upload(cph_handle *h, DevBuf<T> &buf, const T *src, size_t n, int where = CPH_HOST) {
  CPH_CUDA(h, buf.reserve(n + 1));
  if (n)
    CPH_CUDA(h, cudaMemcpyAsync(buf.p, src, n * sizeof(T),
                                where == CPH_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}

int size_sites(cph_handle *h) {
  size_t S = (size_t)h->S;
  DevBuf<double> *bufs[] = {&h->d_lam, &h->d_vlam, &h->d_alam, &h->d_flam, &h->d_fs, &h->d_dfs, &h->d_Us, &h->d_dUs,
                            &h->d_theta};
  for (auto *b : bufs) {
    CPH_CUDA(h, b->reserve(S + 1));
    CPH_CUDA(h, cudaMemsetAsync(b->p, 0, (S + 1) * sizeof(double), h->stream));
  }
  CPH_CUDA(h, h->d_red.reserve(4 + 2 * S + 8));
  CPH_CUDA(h, cudaMemsetAsync(h->d_red.p, 0, (4 + 2 * S + 8) * sizeof(double), h->stream));
  CPH_CUDA(h, h->d_scal.reserve(16));
  CPH_CUDA(h, cudaMemsetAsync(h->d_scal.p, 0, 16 * sizeof(double), h->stream));
  std::vector<double> half(S, 0.5), quarter_pi(S, 0.78539816339744830962);   // lambda = 1/2 <=> theta = pi/4
  CPH_CUDA(h, cudaMemcpyAsync(h->d_lam.p, half.data(), S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CPH_CUDA(h, cudaMemcpyAsync(h->d_theta.p, quarter_pi.data(), S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// caller-order per-atom result to the caller's buffer (host or device)
int fetch_atoms(cph_handle *h, int what, int width, int where, double *out) {
  const size_t n = (size_t)h->nlocal * width;
  if (n == 0) return 0;
  if (where == CPH_DEVICE) return cph_launch_gather_out(h, what, out);
  CPH_CUDA(h, h->d_stage.reserve(n));
  CPH_TRY(cph_launch_gather_out(h, what, h->d_stage.p));
  CPH_CUDA(h, cudaMemcpyAsync(out, h->d_stage.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// Is this host pointer page-locked (cudaMallocHost / cudaHostRegister)?  LAMMPS' atom->x and atom->f are not.
bool is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

constexpr int HCHUNKS = 4;   // host staging pipeline depth

int host_threads() {
  static const int nt = std::max(1, std::min(16, omp_get_max_threads()));
  return nt;
}

// caller's host array -> device.  Page-locked memory goes straight through the copy engine; pageable memory
// (LAMMPS' own arrays) is copied by the host cores into the library's page-locked staging in HCHUNKS pieces,
// each shipped as soon as it is complete, so the DMA of one piece overlaps the copy of the next.
int upload_host_array(cph_handle *h, double *dst_dev, const double *src, size_t n) {
  if (n == 0) return 0;
  if (is_pinned(src)) {
    CPH_CUDA(h, cudaMemcpyAsync(dst_dev, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return 0;
  }
  CPH_CUDA(h, cudaEventSynchronize(h->ev_xstage));          // the previous upload has left the staging buffer
  CPH_TRY(ensure_pinned(h, n * sizeof(double) + 64));
  double *stage = h->h_pin;
  const int nt = host_threads();
  for (int c = 0; c < HCHUNKS; c++) {
    const size_t lo = n * c / HCHUNKS, hi = n * (c + 1) / HCHUNKS;
    if (hi == lo) continue;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int t = 0; t < nt; t++) {
      const size_t a = lo + (hi - lo) * t / nt, b = lo + (hi - lo) * (t + 1) / nt;
      memcpy(stage + a, src + a, (b - a) * sizeof(double));
    }
    CPH_CUDA(h, cudaMemcpyAsync(dst_dev + lo, stage + lo, (hi - lo) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  CPH_CUDA(h, cudaEventRecord(h->ev_xstage, h->stream));
  return 0;
}

// device forces (already gathered to caller order in d_stage on `st`) -> caller's host array: stored, or
// ADDED to what the array holds (cph_set_force_mode: the fix under `pair_modify compute no`, cpp:149-171 keeps
// the other force contributions LAMMPS put there).  Straight DMA when the destination is page-locked and the
// mode is "store"; otherwise HCHUNKS pieces through page-locked staging, each added / copied by the host cores
// while the next one is in flight.  Returns with the data in place.
int deliver_forces(cph_handle *h, double *f, size_t n3, cudaStream_t st) {
  if (n3 == 0) return 0;
  if (!h->force_add && is_pinned(f)) {
    CPH_CUDA(h, cudaMemcpyAsync(f, h->d_stage.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CPH_CUDA(h, cudaStreamSynchronize(st));
    return 0;
  }
  if (n3 * sizeof(double) > h->h_fpin_bytes) {
    if (h->h_fpin) cudaFreeHost(h->h_fpin);
    h->h_fpin = nullptr;
    h->h_fpin_bytes = 0;
    const size_t want = n3 * sizeof(double) + n3 + 64;
    CPH_CUDA(h, cudaMallocHost((void **)&h->h_fpin, want));
    h->h_fpin_bytes = want;
  }
  double *stage = h->h_fpin;
  for (int c = 0; c < HCHUNKS; c++) {
    const size_t lo = n3 * c / HCHUNKS, hi = n3 * (c + 1) / HCHUNKS;
    if (hi > lo)
      CPH_CUDA(h, cudaMemcpyAsync(stage + lo, h->d_stage.p + lo, (hi - lo) * sizeof(double), cudaMemcpyDeviceToHost, st));
    CPH_CUDA(h, cudaEventRecord(h->ev_fchunk[c], st));
  }
  const int nt = host_threads();
  const bool add = h->force_add;
  for (int c = 0; c < HCHUNKS; c++) {
    const size_t lo = n3 * c / HCHUNKS, hi = n3 * (c + 1) / HCHUNKS;
    CPH_CUDA(h, cudaEventSynchronize(h->ev_fchunk[c]));
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int t = 0; t < nt; t++) {
      const size_t a = lo + (hi - lo) * t / nt, b = lo + (hi - lo) * (t + 1) / nt;
      if (add) for (size_t k = a; k < b; k++) f[k] += stage[k];
      else memcpy(f + a, stage + a, (b - a) * sizeof(double));
    }
  }
  return 0;
}

int read_flags(cph_handle *h, unsigned int *flags_h) {
  CPH_CUDA(h, cudaMemcpyAsync(flags_h, h->d_flags.p, 8 * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

}  // namespace

extern "C" {

int cph_version(void) { return 100; }

int cph_create(int device, cph_handle **out) {
  if (!out) return cph_fail(nullptr, CPH_ERR_ARG, "cph_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return cph_fail(nullptr, CPH_ERR_CUDA, "no CUDA device (%s); libcph_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return cph_fail(nullptr, CPH_ERR_ARG, "device %d out of range [0,%d)", device, ndev);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major < 10)
    return cph_fail(nullptr, CPH_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
  cph_handle *h = new cph_handle();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return cph_fail(nullptr, CPH_ERR_CUDA, "cannot create a stream on device %d", device);
  }
  {
    cudaError_t rc = cudaSuccess;
    auto keep = [&](cudaError_t e2) { if (rc == cudaSuccess && e2 != cudaSuccess) rc = e2; };
    keep(cudaEventCreate(&h->ev0)); keep(cudaEventCreate(&h->ev1));
    keep(cudaEventCreate(&h->pev0)); keep(cudaEventCreate(&h->pev1));
    keep(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
    keep(cudaEventCreateWithFlags(&h->ev_flags, cudaEventDisableTiming));
    keep(cudaEventCreateWithFlags(&h->ev_force, cudaEventDisableTiming));
    keep(cudaEventCreateWithFlags(&h->ev_xstage, cudaEventDisableTiming));
    for (auto &ev : h->ev_fchunk) keep(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    keep(cudaMallocHost((void **)&h->h_flags, 16 * sizeof(unsigned int)));
    keep(h->d_flags.reserve(96));
    if (rc == cudaSuccess) keep(cudaMemsetAsync(h->d_flags.p, 0, 96 * sizeof(unsigned int), h->stream));
    if (rc != cudaSuccess) {
      cph_fail(nullptr, CPH_ERR_CUDA, "cph_create: %s while setting up streams, events and flag buffers on device %d",
               cudaGetErrorString(rc), device);
      cph_destroy(h);
      return CPH_ERR_CUDA;
    }
  }
  if (const char *e = getenv("CPH_INNER_SKIN")) h->inner_skin = std::max(0.0, atof(e));   // tuning knobs
  if (const char *e = getenv("CPH_SPECULATE")) h->speculate = atoi(e) != 0;
  if (const char *e = getenv("CPH_HALO")) h->peer_halo_wanted = strcmp(e, "nccl") != 0;   // "nccl" forces ncclSend/Recv
  if (const char *e = getenv("CPH_MAIL")) h->mail_wanted = atoi(e) != 0;                   // 0: NCCL for the small all-reduces
  int rc = size_sites(h);
  if (rc == 0) {   // no site table yet: the reference's single site with an empty titratable-atom range
    const int zero2[2] = {0, 0};
    rc = upload(h, h->d_site_start, zero2, 2);
    if (rc == 0 && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = CPH_ERR_CUDA;
  }
  if (rc) { g_create_error = h->err; delete h; return rc; }
  *out = h;
  return CPH_OK;
}

int cph_destroy(cph_handle *h) {
  if (!h) return CPH_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cph_halo_close(h);
  cph_mail_close(h);
  cph_comm_destroy(h);
  cph_bonded_release(h);
  cph_kspace_release(h);
  DevBuf<double> *db[] = {&h->d_pK, &h->d_lam, &h->d_vlam, &h->d_alam, &h->d_flam, &h->d_fs, &h->d_dfs, &h->d_Us,
                          &h->d_dUs, &h->d_theta, &h->d_red, &h->d_titr_qA, &h->d_titr_dq, &h->d_scal, &h->d_part, &h->d_xbuild,
                          &h->d_f, &h->d_evdwl, &h->d_phi, &h->d_eatom, &h->d_stage, &h->d_wq, &h->d_dQ};
  for (auto *b : db) b->release();
  DevBuf<int> *ib[] = {&h->d_titr_tag_sorted, &h->d_titr_entry_of_sorted, &h->d_titr_site, &h->d_titr_local, &h->d_site_start, &h->d_type,
                       &h->d_tag, &h->d_mask, &h->d_perm, &h->d_inv, &h->d_site_of, &h->d_titr_of, &h->d_nspecial,
                       &h->d_special, &h->d_ghost_src, &h->d_ghost_code, &h->d_hlist, &h->d_istage, &h->d_vals,
                       &h->d_vals2, &h->d_tmpi, &h->d_cell_start_o, &h->d_cell_start_g, &h->d_neigh, &h->d_numneigh, &h->d_numspec, &h->d_neigh2, &h->d_numneigh2, &h->d_scr_i, &h->d_scr_src, &h->d_scr_code, &h->d_scr_off, &h->d_mol, &h->d_rec_src, &h->d_rec_dir,
                       &h->d_wtag, &h->d_wlocal};
  for (auto *b : ib) b->release();
  h->d_xb.release(); h->d_molecule.release(); h->d_coef.release(); h->d_coef4.release(); h->d_cut2.release(); h->d_type_has_lj.release(); h->d_xt.release(); h->d_xq.release(); h->d_xq2.release(); h->d_keys.release(); h->d_keys2.release();
  h->d_xinner.release(); h->d_exp2.release(); h->d_qnext.release(); h->d_cubtmp.release(); h->d_flags.release(); h->d_scr_stats.release(); h->d_ipc_stage.release();
  h->d_mail.release(); h->d_sendx.release(); h->d_recvx.release(); h->d_sendmeta.release(); h->d_recvmeta.release();
  if (h->h_pin) cudaFreeHost(h->h_pin);
  cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaEventDestroy(h->pev0); cudaEventDestroy(h->pev1);
  cudaEventDestroy(h->ev_flags);
  cudaEventDestroy(h->ev_force);
  if (h->ev_xstage) cudaEventDestroy(h->ev_xstage);
  for (auto ev : h->ev_fchunk) if (ev) cudaEventDestroy(ev);
  if (h->h_fpin) cudaFreeHost(h->h_fpin);
  h->d_xstage.release();
  cudaStreamDestroy(h->stream2);
  cudaFreeHost(h->h_flags);
  cudaStreamDestroy(h->stream);
  delete h;
  return CPH_OK;
}

const char *cph_last_error(cph_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

// ---- configuration ---------------------------------------------------------------------------
int cph_set_units(cph_handle *h, double qqrd2e, double boltz, double ftm2v) {
  if (!(ftm2v > 0)) return cph_fail(h, CPH_ERR_ARG, "ftm2v must be positive");
  h->qqrd2e = qqrd2e;
  h->pp.qqrd2e = qqrd2e;
  h->kc_dirty = true;
  h->fix.boltz = boltz;
  h->fix.ftm2v = ftm2v;
  return CPH_OK;
}

int cph_set_pair(cph_handle *h, int style, int ntypes, const double *epsilon, const double *sigma,
                 const double *cut_lj, double cut_lj_global, double cut_coul, double alpha,
                 const double *special_lj, const double *special_coul) {
  if (style != CPH_PAIR_LJ_CUT_COUL_CUT && style != CPH_PAIR_LJ_CUT_COUL_DSF && style != CPH_PAIR_LJ_CUT_COUL_LONG)
    return cph_fail(h, CPH_ERR_ARG, "unknown pair style %d", style);
  // lj/cut/coul/long (alpha = g_ewald) IS the damped kernel of coul/dsf with both shifts at zero; its self energy
  // belongs to the k-space part (kspace.cu).  Everything downstream sees the dsf style: same kernels, same list rules
  // (LAMMPS keeps fully excluded special pairs in the list whenever a KSpace style is defined).
  h->coul_long = style == CPH_PAIR_LJ_CUT_COUL_LONG;
  if (h->coul_long) style = CPH_PAIR_LJ_CUT_COUL_DSF;
  if (ntypes < 1 || ntypes + 1 > CPH_MAXNT1) return cph_fail(h, CPH_ERR_ARG, "ntypes %d outside [1,%d]", ntypes, CPH_MAXNT1 - 1);
  if (!epsilon || !sigma || !special_lj || !special_coul) return cph_fail(h, CPH_ERR_ARG, "NULL coefficient table");
  if (!(cut_coul > 0)) return cph_fail(h, CPH_ERR_ARG, "cut_coul must be positive");
  cudaSetDevice(h->device);
  const int nt1 = ntypes + 1;
  h->coef_h.assign((size_t)nt1 * nt1, PairCoef{0, 0, 0, 0});
  h->cut_lj_max = 0;
  const double cut_coulsq = cut_coul * cut_coul;
  for (int t = 0; t < nt1 * nt1; t++) {
    double e = epsilon[t], s = sigma[t];
    double c = cut_lj ? cut_lj[t] : cut_lj_global;
    h->coef_h[t].lj3 = 4.0 * e * std::pow(s, 12.0);
    h->coef_h[t].lj4 = 4.0 * e * std::pow(s, 6.0);
    h->coef_h[t].cut_ljsq = c * c;
    h->coef_h[t].cutsq = std::max(c * c, cut_coulsq);
    h->cut_lj_max = std::max(h->cut_lj_max, c);
  }
  PairParams &pp = h->pp;
  pp.style = style;
  pp.ntypes = ntypes;
  pp.qqrd2e = h->qqrd2e;
  pp.alpha = alpha;
  pp.cut_coulsq = cut_coulsq;
  double cm = std::max(h->cut_lj_max, cut_coul);
  pp.cutsq_max = cm * cm;
  pp.e_shift = pp.f_shift = pp.c_self = 0.0;
  if (style == CPH_PAIR_LJ_CUT_COUL_DSF && !h->coul_long) {   // init_style of coul/dsf
    const double MY_PIS = 1.77245385090551602729;
    double erfcc = std::erfc(alpha * cut_coul);
    double erfcd = std::exp(-alpha * alpha * cut_coul * cut_coul);
    pp.f_shift = -(erfcc / cut_coulsq + 2.0 / MY_PIS * alpha * erfcd / cut_coul);
    pp.e_shift = erfcc / cut_coul - pp.f_shift * cut_coul;
    pp.c_self = -(pp.e_shift / 2.0 + alpha / MY_PIS) * h->qqrd2e;
  }
  for (int k = 0; k < 4; k++) { pp.special_lj[k] = special_lj[k]; pp.special_coul[k] = special_coul[k]; }
  h->cut_coul = cut_coul;
  if (style == CPH_PAIR_LJ_CUT_COUL_DSF && alpha * alpha * cut_coulsq > 700.0)
    return cph_fail(h, CPH_ERR_ARG, "alpha*cut_coul = %g is outside the range of the damped kernel", alpha * cut_coul);
  CPH_TRY(upload(h, h->d_coef, h->coef_h.data(), h->coef_h.size()));
  {
    std::vector<double4> c4((size_t)nt1 * nt1);
    std::vector<double2> c2((size_t)nt1 * nt1);
    std::vector<int> has(nt1, 0);
    h->uniform_cut = true;
    for (int t = 0; t < nt1 * nt1; t++) {
      const PairCoef &c = h->coef_h[t];
      c4[t] = make_double4(12.0 * c.lj3, 6.0 * c.lj4, c.lj3, c.lj4);
      c2[t] = make_double2(c.cut_ljsq, c.cutsq);
      if ((c.lj3 != 0.0 || c.lj4 != 0.0) && t / nt1 >= 1 && t % nt1 >= 1) has[t / nt1] = 1;
      if (t / nt1 >= 1 && t % nt1 >= 1 && (c.cut_ljsq != pp.cutsq_max || cut_coulsq != pp.cutsq_max)) h->uniform_cut = false;
    }
    CPH_TRY(upload(h, h->d_coef4, c4.data(), c4.size()));
    CPH_TRY(upload(h, h->d_cut2, c2.data(), c2.size()));
    CPH_TRY(upload(h, h->d_type_has_lj, has.data(), has.size()));
  }
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  h->kc_dirty = true;
  h->have_pair = true;
  h->rowcap = 0;
  return CPH_OK;
}

int cph_set_domain(cph_handle *h, const double *boxlo, const double *boxhi, const int *periodic,
                   const double *sublo, const double *subhi, const int *procgrid, const int *myloc, double skin) {
  if (!boxlo || !boxhi || !periodic) return cph_fail(h, CPH_ERR_ARG, "NULL box");
  if (!(skin >= 0)) return cph_fail(h, CPH_ERR_ARG, "skin must be >= 0");
  for (int k = 0; k < 3; k++) {
    if (!(boxhi[k] > boxlo[k])) return cph_fail(h, CPH_ERR_ARG, "box dimension %d is empty", k);
    h->boxlo[k] = boxlo[k]; h->boxhi[k] = boxhi[k]; h->periodic[k] = periodic[k];
    h->sublo[k] = sublo ? sublo[k] : boxlo[k];
    h->subhi[k] = subhi ? subhi[k] : boxhi[k];
    h->procgrid[k] = procgrid ? procgrid[k] : 1;
    h->myloc[k] = myloc ? myloc[k] : 0;
    if (h->procgrid[k] < 1 || h->myloc[k] < 0 || h->myloc[k] >= h->procgrid[k])
      return cph_fail(h, CPH_ERR_ARG, "bad processor grid in dim %d", k);
  }
  int np = h->procgrid[0] * h->procgrid[1] * h->procgrid[2];
  if (np != h->nranks)
    return cph_fail(h, CPH_ERR_ARG, "procgrid has %d ranks but the rank group has %d (call cph_comm_init_nccl first)", np,
                    h->nranks);
  {
    const int me = (h->myloc[2] * h->procgrid[1] + h->myloc[1]) * h->procgrid[0] + h->myloc[0];
    if (h->nranks > 1 && me != h->rank)
      return cph_fail(h, CPH_ERR_ARG, "myloc maps to rank %d but this handle is rank %d (x-fastest brick order expected)", me,
                      h->rank);
  }
  h->skin = skin;
  h->have_domain = true;
  CPH_TRY(cph_kspace_setup(h));   // the wave vectors follow the box
  h->rowcap = 0;
  return CPH_OK;
}

int cph_set_fix(cph_handle *h, int nevery, int groupHbit, int groupWbit, double pK, double pH, double T) {
  if (nevery <= 0) return cph_fail(h, CPH_ERR_ARG, "Illegal fix constant_pH every value %d", nevery);  // cpp:38 (+D4)
  h->fix.nevery = nevery; h->fix.Hbit = groupHbit; h->fix.Wbit = groupWbit;
  h->fix.pK = pK; h->fix.pH = pH; h->fix.T = T;
  return CPH_OK;
}

int cph_set_bias(cph_handle *h, double w, double s, double hbar, double k, double a, double b, double r, double m,
                 double d, double m_lambda, int bias_mode) {
  if (bias_mode != CPH_BIAS_EXACT && bias_mode != CPH_BIAS_AS_WRITTEN) return cph_fail(h, CPH_ERR_ARG, "bad bias mode");
  if (!(m_lambda > 0) || a == 0 || s == 0) return cph_fail(h, CPH_ERR_ARG, "bad bias constants");
  h->bias = BiasParams{w, s, hbar, k, a, b, r, m, d, m_lambda, bias_mode};
  return CPH_OK;
}

int cph_set_mode(cph_handle *h, int dudl_mode, int integrator_mode, int fscale_mode) {
  if ((dudl_mode | 1) != 1 || (integrator_mode | 1) != 1 || (fscale_mode | 1) != 1) return cph_fail(h, CPH_ERR_ARG, "bad mode");
  h->fix.dudl_mode = dudl_mode; h->fix.integ_mode = integrator_mode; h->fix.fscale_mode = fscale_mode;
  return CPH_OK;
}

int cph_set_thermostat(cph_handle *h, double tau) {
  if (tau < 0) return cph_fail(h, CPH_ERR_ARG, "thermostat period must be >= 0");
  h->nh_tau = tau;
  return CPH_OK;
}

int cph_set_extra_partition(cph_handle *h, double dHA, double dHB) {
  h->extra_HA = dHA;
  h->extra_HB = dHB;
  return CPH_OK;
}

int cph_set_kspace(cph_handle *h, int style, double g_ewald, int kxmax, int kymax, int kzmax) {
  if (style == CPH_KSPACE_NONE) {
    h->kspace_style = CPH_KSPACE_NONE;
    h->nkvec = 0;
    return CPH_OK;
  }
  if (style != CPH_KSPACE_EWALD) return cph_fail(h, CPH_ERR_ARG, "unknown kspace style %d", style);
  if (!(g_ewald > 0.0) || kxmax < 1 || kymax < 1 || kzmax < 1)
    return cph_fail(h, CPH_ERR_ARG, "ewald: g_ewald %g, kmax %d %d %d", g_ewald, kxmax, kymax, kzmax);
  CPH_TRY(need(h, h->have_domain, "cph_set_domain must precede cph_set_kspace"));
  if (!(h->periodic[0] && h->periodic[1] && h->periodic[2]))
    return cph_fail(h, CPH_ERR_ARG, "ewald needs a fully periodic box");
  if ((double)(kxmax + 1) * (2 * kymax + 1) * (2 * kzmax + 1) > 4.0e6)
    return cph_fail(h, CPH_ERR_ARG, "ewald: %d x %d x %d wave vectors is mesh-solver territory", kxmax, kymax, kzmax);
  cudaSetDevice(h->device);
  h->kspace_style = style;
  h->g_ewald = g_ewald;
  h->kmax[0] = kxmax; h->kmax[1] = kymax; h->kmax[2] = kzmax;
  return cph_kspace_setup(h);
}

int cph_get_kspace_energy(cph_handle *h, double *e) {
  if (!e) return cph_fail(h, CPH_ERR_ARG, "NULL output");
  cudaSetDevice(h->device);
  return cph_kspace_energy(h, e);
}

int cph_set_extra_dudl(cph_handle *h, int nsites, const double *dudl) {
  CPH_TRY(need(h, h->have_sites, "cph_set_sites first"));
  if (nsites != h->S || !dudl)
    return cph_fail(h, CPH_ERR_ARG, "cph_set_extra_dudl: %d values for %d sites", dudl ? nsites : 0, h->S);
  cudaSetDevice(h->device);
  // a pageable source has been read completely when cudaMemcpyAsync returns: the caller may reuse its array
  CPH_TRY(upload(h, h->d_extra_dudl, dudl, (size_t)nsites));
  h->extra_dudl = true;
  return CPH_OK;
}

int cph_set_excluded_policy(cph_handle *h, int drop) {
  h->drop_excluded = drop != 0;
  h->rowcap = 0;
  return CPH_OK;
}

int cph_set_coordinate(cph_handle *h, int coordinate) {
  if (coordinate != CPH_COORD_LAMBDA && coordinate != CPH_COORD_THETA) return cph_fail(h, CPH_ERR_ARG, "bad coordinate");
  h->coord_theta = coordinate == CPH_COORD_THETA;
  return CPH_OK;
}

int cph_set_water_buffer(cph_handle *h, int enable) {
  if (enable < 0) return cph_fail(h, CPH_ERR_ARG, "water buffer: negative atom count");
  // `enable` is the number of atoms in the water group (the fix insists on 3, cpp:44-45); 1 is read as 3
  h->water_n = enable == 1 ? 3 : enable;
  return CPH_OK;
}

int cph_set_sites(cph_handle *h, int nsites, const double *pK, int ntitr, const int *titr_tag, const int *titr_site,
                  const double *qA, const double *qB) {
  if (nsites < 0 || ntitr < 0) return cph_fail(h, CPH_ERR_ARG, "negative site/atom count");
  if (nsites > 0 && !pK) return cph_fail(h, CPH_ERR_ARG, "pK is NULL");
  if (ntitr > 0 && (!titr_tag || !titr_site || !qA || !qB)) return cph_fail(h, CPH_ERR_ARG, "NULL titratable-atom table");
  cudaSetDevice(h->device);
  h->fix.implicit_site = (nsites == 0);
  h->S = nsites == 0 ? 1 : nsites;
  h->ntitr = ntitr;
  for (int t = 0; t < ntitr; t++)
    if (titr_site[t] < 0 || titr_site[t] >= h->S) return cph_fail(h, CPH_ERR_ARG, "site index %d out of range", titr_site[t]);
  // site-major order (stable in tag) so the per-site reduction is a segmented scan
  std::vector<int> order(ntitr);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    return titr_site[a] != titr_site[b] ? titr_site[a] < titr_site[b] : titr_tag[a] < titr_tag[b];
  });
  h->titr_order_h = order;
  h->lj_states = false;                      // cph_set_lj_states follows a new table
  std::vector<int> site(ntitr), tag(ntitr);
  std::vector<double> a(ntitr), dq(ntitr);
  for (int k = 0; k < ntitr; k++) {
    int t = order[k];
    site[k] = titr_site[t]; tag[k] = titr_tag[t]; a[k] = qA[t]; dq[k] = qB[t] - qA[t];
  }
  // tag-sorted view for the device binary search
  std::vector<int> byt(ntitr);
  std::iota(byt.begin(), byt.end(), 0);
  std::sort(byt.begin(), byt.end(), [&](int x, int y) { return tag[x] < tag[y]; });
  h->titr_tag_sorted_h.resize(ntitr);
  h->titr_entry_of_sorted_h.resize(ntitr);
  for (int k = 0; k < ntitr; k++) {
    h->titr_tag_sorted_h[k] = tag[byt[k]];
    h->titr_entry_of_sorted_h[k] = byt[k];
    if (k && h->titr_tag_sorted_h[k] == h->titr_tag_sorted_h[k - 1])
      return cph_fail(h, CPH_ERR_ARG, "atom tag %d appears twice in the titratable-atom table", tag[byt[k]]);
  }
  {
    std::vector<int> start(h->S + 1, 0);
    for (int k = 0; k < ntitr; k++) start[site[k] + 1]++;
    int biggest = 1;
    for (int s2 = 0; s2 < h->S; s2++) { biggest = std::max(biggest, start[s2 + 1]); start[s2 + 1] += start[s2]; }
    h->site_lps = 1;
    while (h->site_lps < 32 && h->site_lps < biggest) h->site_lps *= 2;
    CPH_TRY(upload(h, h->d_site_start, start.data(), start.size()));
  }
  CPH_TRY(upload(h, h->d_titr_site, site.data(), ntitr));
  CPH_TRY(upload(h, h->d_titr_qA, a.data(), ntitr));
  CPH_TRY(upload(h, h->d_titr_dq, dq.data(), ntitr));
  CPH_TRY(upload(h, h->d_titr_tag_sorted, h->titr_tag_sorted_h.data(), ntitr));
  CPH_TRY(upload(h, h->d_titr_entry_of_sorted, h->titr_entry_of_sorted_h.data(), ntitr));
  {
    std::vector<double> dQ(h->S, 0.0);
    for (int t = 0; t < ntitr; t++) dQ[titr_site[t]] += qB[t] - qA[t];
    CPH_TRY(upload(h, h->d_dQ, dQ.data(), dQ.size()));
  }
  CPH_TRY(upload(h, h->d_pK, pK, nsites));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  CPH_TRY(size_sites(h));
  h->have_sites = true;
  return CPH_OK;
}

int cph_set_lj_states(cph_handle *h, int ntitr, const int *typeB) {
  cudaSetDevice(h->device);
  return cph_ljstates_set(h, ntitr, typeB);
}

int cph_set_lambda(cph_handle *h, const double *lambda, const double *v_lambda) {
  cudaSetDevice(h->device);
  size_t b = (size_t)h->S * sizeof(double);
  if (lambda) CPH_CUDA(h, cudaMemcpyAsync(h->d_lam.p, lambda, b, cudaMemcpyHostToDevice, h->stream));
  std::vector<double> th;
  if (lambda) {   // theta = asin(sqrt(lambda)) on [0, pi/2]; kept in step with lambda in either mode
    th.resize(h->S);
    for (int s = 0; s < h->S; s++) th[s] = std::asin(std::sqrt(std::min(1.0, std::max(0.0, lambda[s]))));
    CPH_CUDA(h, cudaMemcpyAsync(h->d_theta.p, th.data(), b, cudaMemcpyHostToDevice, h->stream));
  }
  if (v_lambda) CPH_CUDA(h, cudaMemcpyAsync(h->d_vlam.p, v_lambda, b, cudaMemcpyHostToDevice, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->have_atoms && h->fix.dudl_mode == CPH_DUDL_CHARGE) {
    CPH_TRY(cph_launch_apply_charges(h));
    CPH_TRY(cph_forward_ghosts(h));
  }
  return CPH_OK;
}

// ---- atoms -------------------------------------------------------------------------------------
int cph_set_atoms(cph_handle *h, int where, int nlocal, const double *x, const double *q, const int *type,
                  const int *tag, const int *mask, const int *molecule, const int *nspecial, const int *special,
                  int maxspecial) {
  CPH_TRY(need(h, h->have_pair && h->have_domain, "cph_set_pair and cph_set_domain must precede cph_set_atoms"));
  if (nlocal < 0) return cph_fail(h, CPH_ERR_ARG, "nlocal < 0");
  if (nlocal > 0 && (!x || !q || !type || !tag || !mask)) return cph_fail(h, CPH_ERR_ARG, "NULL per-atom array");
  if (nlocal > CPH_NEIGHMASK / 2) return cph_fail(h, CPH_ERR_OVERFLOW, "too many atoms for 29-bit neighbour indices");
  if (maxspecial < 0 || (maxspecial > 0 && (!nspecial || !special))) return cph_fail(h, CPH_ERR_ARG, "bad special tables");
  cudaSetDevice(h->device);
  cudaStream_t st = h->stream;
  const size_t n = (size_t)nlocal;
  h->nlocal = nlocal;
  h->maxspecial = maxspecial;
  // pack {x,y,z,q} on the host side of the copy: one pinned staging buffer, one H2D
  CPH_CUDA(h, h->d_xq.reserve(n + 1));
  if (where == CPH_HOST) {
    CPH_TRY(ensure_pinned(h, n * sizeof(double4) + 64));
    double4 *p = (double4 *)h->h_pin;
    for (size_t i = 0; i < n; i++) p[i] = make_double4(x[3 * i], x[3 * i + 1], x[3 * i + 2], q[i]);
    CPH_CUDA(h, cudaMemcpyAsync(h->d_xq.p, p, n * sizeof(double4), cudaMemcpyHostToDevice, st));
  } else {
    CPH_TRY(cph_launch_pack_xq(h, nlocal, x, q));
  }
  CPH_TRY(upload(h, h->d_type, type, n, where));
  CPH_TRY(upload(h, h->d_tag, tag, n, where));
  CPH_TRY(upload(h, h->d_mask, mask, n, where));
  h->have_mol = molecule != nullptr && maxspecial > 0;
  if (h->have_mol) CPH_TRY(upload(h, h->d_molecule, molecule, n, where));
  if (maxspecial) {
    CPH_TRY(upload(h, h->d_nspecial, nspecial, 3 * n, where));
    CPH_TRY(upload(h, h->d_special, special, n * maxspecial, where));
  }
  // modify_water: remember the owned buffer atoms and the charge they were given (lambda = 0 state)
  h->wtag_h.clear();
  h->wq_h.clear();
  if (h->water_n > 0) {
    if (where != CPH_HOST) return cph_fail(h, CPH_ERR_ARG, "the water buffer needs host-resident mask/tag/q in cph_set_atoms");
    for (size_t i = 0; i < n; i++)
      if (mask[i] & h->fix.Wbit) { h->wtag_h.push_back(tag[i]); h->wq_h.push_back(q[i]); }
    if ((int)h->wtag_h.size() > h->water_n || h->wtag_h.size() > 64)
      return cph_fail(h, CPH_ERR_ARG, "the water group has %zu atoms on this rank, expected at most %d", h->wtag_h.size(), h->water_n);
  }
  h->nw_local = (int)h->wtag_h.size();
  CPH_TRY(upload(h, h->d_wtag, h->wtag_h.data(), h->wtag_h.size()));
  CPH_TRY(upload(h, h->d_wq, h->wq_h.data(), h->wq_h.size()));
  CPH_CUDA(h, h->d_wlocal.reserve(h->wtag_h.size() + 1));
  std::vector<int> ident(n);
  std::iota(ident.begin(), ident.end(), 0);
  CPH_TRY(upload(h, h->d_perm, ident.data(), n));
  CPH_CUDA(h, cudaStreamSynchronize(st));
  for (size_t i = 0; i < n; i++)
    if (type[i] < 1 || type[i] > h->pp.ntypes) {
      if (where == CPH_HOST) return cph_fail(h, CPH_ERR_ARG, "atom %zu has type %d outside [1,%d]", i, type[i], h->pp.ntypes);
      break;
    }
  h->have_atoms = false;
  h->have_topology = false;   // per-atom lists and velocities follow the atom order: the host sends them again
  h->md_on = false;
  CPH_TRY(cph_rebuild(h));
  h->have_atoms = true;
  h->have_pass = false;
  if (h->fix.dudl_mode == CPH_DUDL_CHARGE && h->ntitr) {
    // charges follow lambda from the first pass on (q(lambda), north_star)
    CPH_TRY(cph_launch_apply_charges(h));
    CPH_TRY(cph_forward_ghosts(h));
  }
  return CPH_OK;
}

// ---- per step -------------------------------------------------------------------------------------
int cph_set_x(cph_handle *h, int where, const double *x) {
  CPH_TRY(need(h, h->have_atoms, "cph_set_atoms first"));
  if (!x) return cph_fail(h, CPH_ERR_ARG, "x is NULL");
  cudaSetDevice(h->device);
  const size_t n3 = 3 * (size_t)h->nlocal;
  const double *xd = x;
  if (where == CPH_HOST) {
    CPH_CUDA(h, h->d_xstage.reserve(n3 + 1));
    CPH_TRY(upload_host_array(h, h->d_xstage.p, x, n3));
    xd = h->d_xstage.p;
  }
  return cph_launch_set_x(h, xd);
}

int cph_check_rebuild(cph_handle *h, int *flag) {
  CPH_TRY(need(h, h->have_atoms, "cph_set_atoms first"));
  cudaSetDevice(h->device);
  CPH_TRY(cph_launch_set_x(h, nullptr));
  unsigned int fl[8];
  CPH_TRY(cph_comm_allreduce_max_u32_dev(h, h->d_flags.p, 6));   // global neighbor->decide()
  CPH_TRY(read_flags(h, fl));
  unsigned int any = fl[4];
  if (fl[5]) h->inner_valid = false;      // someone moved more than inner_skin/2 since the last prune
  float md;
  memcpy(&md, &fl[0], 4);
  h->scal_h[6] = md;
  if (flag) *flag = any ? 1 : 0;
  return CPH_OK;
}

int cph_forward(cph_handle *h) {
  CPH_TRY(need(h, h->have_atoms, "cph_set_atoms first"));
  cudaSetDevice(h->device);
  return cph_forward_ghosts(h);
}

int cph_pair_pass(cph_handle *h, int eflag) {
  CPH_TRY(need(h, h->have_atoms, "cph_set_atoms first"));
  cudaSetDevice(h->device);
  CPH_TRY(cph_launch_pair(h, eflag ? 1 : 0));
  CPH_TRY(cph_launch_bonded(h, eflag ? 1 : 0));
  CPH_TRY(cph_launch_ljstates(h, eflag ? 1 : 0));
  CPH_TRY(cph_launch_kspace(h, eflag ? 1 : 0));
  h->have_pass = true;
  return CPH_OK;
}

// compute_Hs() tail: partition + per-site sums + the all-reduce that replaces MPI_Allreduce (cpp:274).
// fused: the caller launches the lambda update next, which completes the mailbox all-reduce itself.
static int site_reduce_impl(cph_handle *h, bool fused) {
  CPH_TRY(need(h, h->have_pass, "cph_pair_pass (with eflag) first"));
  cudaSetDevice(h->device);
  if (h->red_pending) CPH_TRY(cph_launch_red_gather(h));        // an earlier reduction nobody consumed
  if (cph_mail_red_usable(h)) {
    // one-shot all-reduce over NVLink: every rank stores its block into every rank's mailbox
    CPH_TRY(cph_launch_water_phi(h));                             // its slot travels with the block
    CPH_TRY(cph_launch_partition(h, true));
    if (!fused) {
      CPH_TRY(cph_launch_red_gather(h));
      if (h->fix.dudl_mode == CPH_DUDL_CHARGE) CPH_TRY(cph_launch_water_dudl(h));
    }
    return CPH_OK;
  }
  CPH_TRY(cph_launch_partition(h));
  CPH_TRY(cph_launch_water_phi(h));
  CPH_TRY(cph_comm_allreduce(h, h->d_red.p, 4 + 2 * h->S + 1));   // cpp:274
  if (h->fix.dudl_mode == CPH_DUDL_CHARGE) CPH_TRY(cph_launch_water_dudl(h));
  return CPH_OK;
}

int cph_site_reduce(cph_handle *h) { return site_reduce_impl(h, false); }

int cph_integrate_lambda(cph_handle *h, double dt) {
  cudaSetDevice(h->device);
  return cph_launch_integrate(h, dt, h->fix.integ_mode == CPH_INTEGRATE_REFERENCE ? 0 : 2);
}

int cph_initial_integrate(cph_handle *h, double dt) {
  if (h->fix.integ_mode != CPH_INTEGRATE_VV) return CPH_OK;
  cudaSetDevice(h->device);
  const bool charge = h->fix.dudl_mode == CPH_DUDL_CHARGE && h->have_atoms;
  CPH_TRY(cph_launch_integrate(h, dt, 1, charge));      // kick + drift, charges follow in the same launch
  if (charge) CPH_TRY(cph_forward_ghosts(h));
  return CPH_OK;
}

int cph_final_integrate(cph_handle *h, double dt) {
  if (h->fix.integ_mode != CPH_INTEGRATE_VV) return CPH_OK;
  cudaSetDevice(h->device);
  return cph_launch_integrate(h, dt, 3);
}

int cph_apply_charges(cph_handle *h) {
  CPH_TRY(need(h, h->have_atoms, "cph_set_atoms first"));
  cudaSetDevice(h->device);
  CPH_TRY(cph_launch_apply_charges(h));
  return cph_forward_ghosts(h);
}

int cph_set_force(cph_handle *h) {
  CPH_TRY(need(h, h->have_pass, "cph_pair_pass first"));
  cudaSetDevice(h->device);
  return cph_launch_set_force(h);
}

// post_force() (cpp:67-79).  advance == false is setup(): everything is evaluated (forces, partition, site sums,
// bias, F_lambda, H_lambda, force rescale) but lambda does not move -- the reference declares setup (h:35)
// without a body and never integrates there.
static int post_force_impl(cph_handle *h, int64_t ntimestep, double dt, int where, const double *x, double *f,
                           bool advance) {
  CPH_TRY(need(h, h->have_atoms, "cph_set_atoms first"));
  cudaSetDevice(h->device);
  // new positions + neighbor->decide()
  if (x) CPH_TRY(cph_set_x(h, where, x));
  else CPH_TRY(cph_launch_set_x(h, nullptr));
  // Halo first (speculatively: a rebuild redoes it, which costs nothing extra).  In peer mode the
  // pack kernel has stored this rank's copies into the neighbours' buffers; the all-reduce of the
  // decision flags that follows is also the barrier after which every rank may read its own buffer.
  CPH_TRY(cph_halo_send(h));
  CPH_TRY(cph_flags_allreduce(h, h->d_flags.p));                  // global decision, one host sync
  // the flags travel to the host on a side stream while the main stream builds the ghost atoms
  CPH_CUDA(h, cudaEventRecord(h->ev_flags, h->stream));
  CPH_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_flags, 0));
  CPH_CUDA(h, cudaMemcpyAsync(h->h_flags, h->d_flags.p, 8 * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream2));
  CPH_TRY(cph_halo_finish(h));
  const bool active = !advance || (ntimestep % h->fix.nevery) == 0;       // cpp:69
  // Most steps need neither a re-neighbouring nor a prune.  Enqueue the pair pass on that assumption,
  // gated on the device copy of the flags, so the GPU has work while the host waits for its copy;
  // when the guess is wrong the gated grid retires at once and the pass is launched again below.
  const bool guessed = h->speculate && h->inner_valid && !h->profiling && h->nlocal > 0;
  if (guessed) CPH_TRY(cph_launch_pair(h, active ? 1 : 0, h->d_flags.p));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream2));
  const unsigned int *fl = h->h_flags;
  unsigned int any = fl[4];
  if (fl[5]) h->inner_valid = false;      // someone moved more than inner_skin/2 since the last prune
  float md;
  memcpy(&md, &fl[0], 4);
  h->scal_h[6] = md;
  if (any) {
    memcpy(&h->drift_value, &fl[2], 4);   // all-reduced with the other decision words: no second round trip
    h->drift_known = true;
    CPH_TRY(cph_rebuild(h));
  }
  if (!guessed || any || fl[5]) CPH_TRY(cph_launch_pair(h, active ? 1 : 0));
  CPH_TRY(cph_launch_bonded(h, active ? 1 : 0));                        // cpp:221-229: bonded eatom joins the partition
  CPH_TRY(cph_launch_ljstates(h, active ? 1 : 0));                      // LJ end states, when the caller declared any
  CPH_TRY(cph_launch_kspace(h, active ? 1 : 0));                        // cpp:241-244: the k-space eatom joins the partition
  h->have_pass = true;
  // In charge mode nothing after this point touches the forces, so their way back to the host
  // (gather to caller order + D2H) runs on the side stream under the site reduce / lambda update.
  const size_t n3 = 3 * (size_t)h->nlocal;
  const bool early_f = f && where == CPH_HOST && h->fix.dudl_mode == CPH_DUDL_CHARGE && h->nlocal > 0 && !h->profiling;
  if (early_f) {
    CPH_CUDA(h, h->d_stage.reserve(n3));
    CPH_CUDA(h, cudaEventRecord(h->ev_force, h->stream));
    CPH_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_force, 0));
    CPH_TRY(cph_launch_gather_out(h, 0, h->d_stage.p, h->stream2));
  }
  if (active) {
    CPH_TRY(site_reduce_impl(h, true));                                 // cpp:70
    const int phase = (h->fix.integ_mode == CPH_INTEGRATE_REFERENCE && advance) ? 0 : 2;
    // cpp:71-73; t_lambda = nevery*dt (cpp:113); the charges follow lambda in the same launch
    CPH_TRY(cph_launch_integrate(h, dt * h->fix.nevery, phase, h->fix.dudl_mode == CPH_DUDL_CHARGE && phase == 0));
  }
  if (h->fix.dudl_mode == CPH_DUDL_REFERENCE) CPH_TRY(cph_launch_set_force(h));   // cpp:78, every step
  if (early_f) {
    CPH_TRY(deliver_forces(h, f, n3, h->stream2));
    CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  } else if (f && where == CPH_HOST && h->nlocal > 0) {
    CPH_CUDA(h, h->d_stage.reserve(n3));
    CPH_TRY(cph_launch_gather_out(h, 0, h->d_stage.p));
    CPH_TRY(deliver_forces(h, f, n3, h->stream));
  } else if (f) {
    CPH_TRY(fetch_atoms(h, 0, 3, where, f));
  }
  return CPH_OK;
}

int cph_post_force(cph_handle *h, int64_t ntimestep, double dt, int where, const double *x, double *f) {
  return post_force_impl(h, ntimestep, dt, where, x, f, true);
}

int cph_setup(cph_handle *h, int64_t ntimestep, int where, const double *x, double *f) {
  return post_force_impl(h, ntimestep, 0.0, where, x, f, false);
}

int cph_set_force_mode(cph_handle *h, int accumulate) {
  h->force_add = accumulate != 0;
  return CPH_OK;
}

int cph_alloc_host(size_t bytes, void **p) {
  if (!p) return CPH_ERR_ARG;
  *p = nullptr;
  return cudaMallocHost(p, bytes ? bytes : 1) == cudaSuccess ? CPH_OK : CPH_ERR_CUDA;
}

int cph_free_host(void *p) {
  if (p) cudaFreeHost(p);
  return CPH_OK;
}

// ---- results -----------------------------------------------------------------------------------------
int cph_get_forces(cph_handle *h, int where, double *f) {
  CPH_TRY(need(h, h->have_pass, "no pair pass yet"));
  cudaSetDevice(h->device);
  return fetch_atoms(h, 0, 3, where, f);
}
int cph_get_eatom(cph_handle *h, int where, double *e) {
  CPH_TRY(need(h, h->have_pass, "no pair pass yet"));
  cudaSetDevice(h->device);
  return fetch_atoms(h, 1, 1, where, e);
}
int cph_get_phi(cph_handle *h, int where, double *p) {
  CPH_TRY(need(h, h->have_pass, "no pair pass yet"));
  cudaSetDevice(h->device);
  return fetch_atoms(h, 2, 1, where, p);
}
int cph_get_q(cph_handle *h, int where, double *q) {
  CPH_TRY(need(h, h->have_atoms, "cph_set_atoms first"));
  cudaSetDevice(h->device);
  return fetch_atoms(h, 3, 1, where, q);
}

// ---- f2: bonded terms and atom dynamics ------------------------------------------------------------
int cph_set_bonded(cph_handle *h, int nbondtypes, const double *bond_k, const double *bond_r0, int nangletypes,
                   const double *angle_k, const double *angle_theta0) {
  cudaSetDevice(h->device);
  return cph_bonded_set_coef(h, nbondtypes, bond_k, bond_r0, nangletypes, angle_k, angle_theta0);
}
int cph_set_topology(cph_handle *h, int nlocal, int maxbond, const int *num_bond, const int *bond_type,
                     const int *bond_atom, int maxangle, const int *num_angle, const int *angle_type,
                     const int *angle_atom1, const int *angle_atom2, const int *angle_atom3) {
  cudaSetDevice(h->device);
  return cph_bonded_set_topology(h, nlocal, maxbond, num_bond, bond_type, bond_atom, maxangle, num_angle, angle_type,
                                 angle_atom1, angle_atom2, angle_atom3);
}
int cph_get_bonded_energy(cph_handle *h, double *out2) {
  CPH_TRY(need(h, h->have_pass, "no pair pass yet"));
  cudaSetDevice(h->device);
  return cph_bonded_energy(h, out2);
}
int cph_set_mass(cph_handle *h, int ntypes, const double *mass) {
  if (ntypes < 1 || ntypes + 1 > CPH_MAXNT1 || !mass) return cph_fail(h, CPH_ERR_ARG, "cph_set_mass: ntypes %d outside [1,%d]", ntypes, CPH_MAXNT1 - 1);
  for (int t = 1; t <= ntypes; t++) {
    if (!(mass[t] > 0.0)) return cph_fail(h, CPH_ERR_ARG, "mass of type %d must be positive", t);
    h->mass_h[t] = mass[t];
  }
  return CPH_OK;
}
int cph_set_v(cph_handle *h, int where, const double *v) {
  CPH_TRY(need(h, h->have_atoms, "cph_set_atoms first"));
  if (!v && h->nlocal) return cph_fail(h, CPH_ERR_ARG, "v is NULL");
  cudaSetDevice(h->device);
  return cph_md_set_v(h, where, v);
}
int cph_md_initial_integrate(cph_handle *h, double dt) {
  CPH_TRY(need(h, h->md_on && h->have_pass, "cph_set_v and a force pass first"));
  cudaSetDevice(h->device);
  return cph_md_kick(h, dt, 1);
}
int cph_md_final_integrate(cph_handle *h, double dt) {
  CPH_TRY(need(h, h->md_on && h->have_pass, "cph_set_v and a force pass first"));
  cudaSetDevice(h->device);
  return cph_md_kick(h, dt, 0);
}
int cph_get_x(cph_handle *h, int where, double *x) {
  CPH_TRY(need(h, h->have_atoms, "cph_set_atoms first"));
  cudaSetDevice(h->device);
  return fetch_atoms(h, 4, 3, where, x);
}
int cph_get_v(cph_handle *h, int where, double *v) {
  CPH_TRY(need(h, h->have_atoms && h->md_on, "cph_set_v first"));
  cudaSetDevice(h->device);
  return fetch_atoms(h, 5, 3, where, v);
}

int cph_get_scalars(cph_handle *h, double *out8) {
  cudaSetDevice(h->device);
  double red[4], sc[8];
  CPH_CUDA(h, cudaMemcpyAsync(red, h->d_red.p, sizeof(red), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaMemcpyAsync(sc, h->d_scal.p, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  out8[0] = red[0]; out8[1] = red[1]; out8[2] = red[2]; out8[3] = red[3];
  double nhe = 0;
  CPH_CUDA(h, cudaMemcpyAsync(&nhe, h->d_scal.p + 11, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  out8[4] = sc[4]; out8[5] = sc[5]; out8[6] = h->scal_h[6]; out8[7] = nhe;
  return CPH_OK;
}

int cph_get_sites(cph_handle *h, double *lambda, double *v_lambda, double *dudl, double *hdiff, double *f_lambda,
                  double *f, double *df, double *U, double *dU) {
  cudaSetDevice(h->device);
  const size_t S = (size_t)h->S, b = S * sizeof(double);
  struct { double *dst; const double *src; } m[] = {
      {lambda, h->d_lam.p}, {v_lambda, h->d_vlam.p}, {dudl, h->d_red.p + 4}, {hdiff, h->d_red.p + 4 + S},
      {f_lambda, h->d_flam.p}, {f, h->d_fs.p}, {df, h->d_dfs.p}, {U, h->d_Us.p}, {dU, h->d_dUs.p}};
  for (auto &e : m)
    if (e.dst) CPH_CUDA(h, cudaMemcpyAsync(e.dst, e.src, b, cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return CPH_OK;
}

int cph_compute_scalar(cph_handle *h, double *out) {
  double s[8];
  CPH_TRY(cph_get_scalars(h, s));
  *out = s[4];
  return CPH_OK;
}

int cph_compute_vector(cph_handle *h, int i, double *out) {
  if (i < 0 || i >= 4 * h->S) return cph_fail(h, CPH_ERR_ARG, "compute_vector index %d out of range [0,%d)", i, 4 * h->S);
  cudaSetDevice(h->device);
  const int s = i / 4;
  const size_t S = (size_t)h->S;
  const double *src;
  switch (i % 4) {
    case 0: src = h->d_lam.p + s; break;
    case 1: src = h->d_vlam.p + s; break;
    case 2: src = h->d_red.p + 4 + (h->fix.dudl_mode == CPH_DUDL_REFERENCE ? S : 0) + s; break;
    default: src = h->d_flam.p + s;
  }
  CPH_CUDA(h, cudaMemcpyAsync(out, src, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return CPH_OK;
}

int cph_memory_usage(cph_handle *h, double *bytes) {
  double b = 0;
  b += h->d_xq.bytes() + h->d_xq2.bytes() + h->d_neigh.bytes() + h->d_numneigh.bytes();
  b += h->d_type.bytes() + h->d_tag.bytes() + h->d_mask.bytes() + h->d_perm.bytes() + h->d_inv.bytes();
  b += h->d_f.bytes() + h->d_evdwl.bytes() + h->d_phi.bytes() + h->d_eatom.bytes() + h->d_xbuild.bytes();
  b += h->d_keys.bytes() + h->d_keys2.bytes() + h->d_vals.bytes() + h->d_vals2.bytes() + h->d_tmpi.bytes();
  b += h->d_stage.bytes() + h->d_cubtmp.bytes() + h->d_cell_start_o.bytes() + h->d_cell_start_g.bytes();
  b += h->d_nspecial.bytes() + h->d_special.bytes() + h->d_ghost_src.bytes() + h->d_ghost_code.bytes();
  b += h->d_neigh2.bytes() + h->d_numneigh2.bytes() + h->d_xinner.bytes() + h->d_xt.bytes() + h->d_xb.bytes();
  b += h->d_bond_j.bytes() + h->d_bond_t.bytes() + h->d_angle_j.bytes() + h->d_angle_t.bytes() + h->d_bcount.bytes();
  b += h->d_bond_type.bytes() + h->d_bond_atom.bytes() + h->d_angle_type.bytes() + h->d_angle_a1.bytes() +
       h->d_angle_a2.bytes() + h->d_angle_a3.bytes() + h->d_num_bond.bytes() + h->d_num_angle.bytes();
  b += h->d_v.bytes() + h->d_v2.bytes();
  *bytes = b;
  return CPH_OK;
}

int cph_get_counts(cph_handle *h, int64_t *out8) {
  cudaSetDevice(h->device);
  int64_t nt = 0;
  if (h->ntitr && h->have_atoms) {
    std::vector<int> loc(h->ntitr);
    CPH_CUDA(h, cudaMemcpyAsync(loc.data(), h->d_titr_local.p, h->ntitr * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CPH_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int v : loc) nt += v >= 0;
  }
  out8[0] = h->nlocal; out8[1] = h->nghost; out8[2] = h->stored_neigh; out8[3] = h->maxneigh;
  out8[4] = h->special_pairs; out8[5] = h->nbuilds; out8[6] = nt; out8[7] = h->fix.implicit_site ? 0 : h->S;
  return CPH_OK;
}

int cph_get_halo_mode(cph_handle *h, int *mode) {
  *mode = h->nranks == 1 ? 0 : (h->peer_halo ? (h->mail_ok ? 3 : 2) : 1);
  return CPH_OK;
}

int cph_get_inner_counts(cph_handle *h, int64_t *out2) {
  CPH_TRY(need(h, h->have_atoms, "cph_set_atoms first"));
  cudaSetDevice(h->device);
  return cph_inner_counts(h, out2);
}

int cph_get_site_map(cph_handle *h, int *site_of_atom) {
  CPH_TRY(need(h, h->have_atoms, "cph_set_atoms first"));
  cudaSetDevice(h->device);
  const int n = h->nlocal;
  std::vector<int> s(n), perm(n);
  CPH_CUDA(h, cudaMemcpyAsync(s.data(), h->d_site_of.p, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaMemcpyAsync(perm.data(), h->d_perm.p, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int k = 0; k < n; k++) site_of_atom[perm[k]] = s[k];
  return CPH_OK;
}

int cph_get_neighbors(cph_handle *h, int *numneigh, int64_t *keys, int64_t keys_capacity) {
  CPH_TRY(need(h, h->have_atoms, "cph_set_atoms first"));
  cudaSetDevice(h->device);
  return cph_neighbors_to_host(h, numneigh, keys, keys_capacity);
}

// ---- restart: [version, S, (lambda, v, a) * S] as doubles (LAMMPS write_restart layout) ----------------------
int cph_restart_size(cph_handle *h, int *ndoubles) { *ndoubles = 2 + 3 * h->S + (h->nh_tau > 0 ? 3 : 0); return CPH_OK; }

int cph_pack_restart(cph_handle *h, double *buf) {
  cudaSetDevice(h->device);
  const int S = h->S;
  std::vector<double> l(S), v(S), a(S);
  // version 1: the coordinate is lambda; version 2: the coordinate is theta (lambda = sin^2 theta)
  CPH_CUDA(h, cudaMemcpyAsync(l.data(), h->coord_theta ? h->d_theta.p : h->d_lam.p, S * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaMemcpyAsync(v.data(), h->d_vlam.p, S * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaMemcpyAsync(a.data(), h->d_alam.p, S * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  buf[0] = h->coord_theta ? 2.0 : 1.0; buf[1] = S;
  for (int s = 0; s < S; s++) { buf[2 + 3 * s] = l[s]; buf[3 + 3 * s] = v[s]; buf[4 + 3 * s] = a[s]; }
  if (h->nh_tau > 0) {   // thermostat state: xi, eta and the kinetic energy its next half step starts from
    double sc[12];
    CPH_CUDA(h, cudaMemcpyAsync(sc, h->d_scal.p, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
    CPH_CUDA(h, cudaStreamSynchronize(h->stream));
    buf[2 + 3 * S] = sc[8];
    buf[3 + 3 * S] = sc[10];
    buf[4 + 3 * S] = sc[5];
  }
  return CPH_OK;
}

int cph_unpack_restart(cph_handle *h, const double *buf, int nd) {
  cudaSetDevice(h->device);
  const int S = h->S;
  const int extra = h->nh_tau > 0 ? 3 : 0;
  if (!buf || nd < 2 || (int)buf[1] != S || nd != 2 + 3 * S + extra)
    return cph_fail(h, CPH_ERR_ARG, "restart record does not match the site table (%d sites)", S);
  if ((buf[0] == 2.0) != h->coord_theta)
    return cph_fail(h, CPH_ERR_ARG, "restart record was written with the other lambda coordinate");
  std::vector<double> l(S), v(S), a(S), th(S);
  for (int s = 0; s < S; s++) { l[s] = buf[2 + 3 * s]; v[s] = buf[3 + 3 * s]; a[s] = buf[4 + 3 * s]; }
  if (h->coord_theta) {
    th = l;
    for (int s = 0; s < S; s++) { const double sn = std::sin(th[s]); l[s] = sn * sn; }
  } else {
    for (int s = 0; s < S; s++) th[s] = std::asin(std::sqrt(std::min(1.0, std::max(0.0, l[s]))));
  }
  CPH_CUDA(h, cudaMemcpyAsync(h->d_theta.p, th.data(), S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CPH_CUDA(h, cudaMemcpyAsync(h->d_lam.p, l.data(), S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CPH_CUDA(h, cudaMemcpyAsync(h->d_vlam.p, v.data(), S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CPH_CUDA(h, cudaMemcpyAsync(h->d_alam.p, a.data(), S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (extra) {
    CPH_CUDA(h, cudaMemcpyAsync(h->d_scal.p + 8, &buf[2 + 3 * S], sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CPH_CUDA(h, cudaMemcpyAsync(h->d_scal.p + 10, &buf[3 + 3 * S], sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CPH_CUDA(h, cudaMemcpyAsync(h->d_scal.p + 5, &buf[4 + 3 * S], sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->have_atoms && h->fix.dudl_mode == CPH_DUDL_CHARGE) {
    CPH_TRY(cph_launch_apply_charges(h));
    CPH_TRY(cph_forward_ghosts(h));
  }
  return CPH_OK;
}

// ---- timing ----------------------------------------------------------------------------------------------------
int cph_sync(cph_handle *h) {
  cudaSetDevice(h->device);
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return CPH_OK;
}
int cph_stream(cph_handle *h, void **stream) { *stream = (void *)h->stream; return CPH_OK; }
int cph_timer_start(cph_handle *h) {
  cudaSetDevice(h->device);
  CPH_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  return CPH_OK;
}
int cph_timer_stop(cph_handle *h, double *ms) {
  cudaSetDevice(h->device);
  CPH_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  CPH_CUDA(h, cudaEventSynchronize(h->ev1));
  float t = 0;
  CPH_CUDA(h, cudaEventElapsedTime(&t, h->ev0, h->ev1));
  *ms = t;
  return CPH_OK;
}
int cph_profile(cph_handle *h, int enable) {
  h->profiling = enable != 0;
  if (enable) for (auto &p : h->prof) p = ProfSlot{};
  return CPH_OK;
}
int cph_profile_get(cph_handle *h, int which, double *ms_total, int64_t *launches) {
  if (which == 8) {          // total kernel launches of this library (always counted)
    *ms_total = 0.0;
    *launches = h->nlaunch;
    return CPH_OK;
  }
  if (which == 9) {          // inner-list prunes so far
    *ms_total = 0.0;
    *launches = h->nprunes;
    return CPH_OK;
  }
  if (which < 0 || which > 11) return cph_fail(h, CPH_ERR_ARG, "profile slot %d out of range", which);
  *ms_total = h->prof[which].ms;
  *launches = h->prof[which].launches;
  return CPH_OK;
}

}  // extern "C"
