// kspace.cu -- SURVEY.md §8 row f4: the k-space source of compute_Hs (fix_constant_pH.cpp:241-244) on the device.
//
// The reference adds force->kspace->eatom to its per-atom energy before the HA/HB partition and never looks at the
// potential; north_star's charge derivative dU/dlambda_s = sum_i dq_i dE/dq_i needs dE_kspace/dq_i as well.  This file
// is `kspace_style ewald` [upstream LAMMPS ewald.cpp conventions, standard Ewald summation]: the reciprocal sum whose
// real-space partner erfc(g r)/r is pair style CPH_PAIR_LJ_CUT_COUL_LONG (the damped kernel of pair.cu with its
// shifts at zero).  With S(k) = sum_j q_j exp(i k.r_j) over the half space of wave vectors k = 2 pi (nx/Lx, ny/Ly,
// nz/Lz), |n_d| <= kmax_d, k^2 <= gsqmx, and ug(k) = 4 pi/V exp(-k^2/4g^2)/k^2:
//   E      = qqrd2e [ sum_k ug |S|^2 - g/sqrt(pi) sum q_i^2 - pi (sum q_i)^2 / (2 g^2 V) ]
//   phi_i  = dE/dq_i = qqrd2e [ sum_k 2 ug (cos(k.r_i) Re S + sin(k.r_i) Im S) - 2 g q_i/sqrt(pi) - pi sum q/(g^2 V) ]
//   f_i    = qqrd2e q_i sum_k 2 ug k (sin(k.r_i) Re S - cos(k.r_i) Im S)
//   eatom_i = q_i phi_i / 2   (E is a quadratic form of the charges; what LAMMPS' per-atom tally adds up to)
// The pass runs right behind the pair pass and ADDS to its forces, phi and per-atom energy (like bonded.cu), so the
// partition (cpp:264-267) and the per-site sums see the k-space part exactly where cpp:241-244 puts it, and it
// follows q(lambda) every step.
//
// Kernels.  (1) structure factors: one thread per wave vector, atoms staged through shared memory in tiles, the atom
// range split into chunks over blockIdx.y so that small boxes still fill the 148 SMs; (2) ewald_sfac_sum_kernel: chunk
// partials added in a fixed order (bit-reproducible), then -- several ranks -- one all-reduce of the 2(K+1) doubles (a
// zero wave vector at the end carries sum q); (3) per-atom sums: one warp per owned atom, lanes stride over the wave
// vectors (coalesced records), shuffle butterfly, lane 0 adds the atom's terms; (4) a one-block fixed-order sum of
// the per-atom k-space energies for cph_get_kspace_energy.
// (1) and (3) exist in several forms, cross-checked against each other to 1e-15 (tools/ewald_timing.py).  *Direct*:
// sincos(k.r) for every atom and wave vector (2 N K calls of ~120 issue slots; CPH_EWALD=direct, also the fallback
// when the tables would not fit shared memory).  *Tables*: exp(i k.r) = exp(i kx x) exp(i ky y) exp(i kz z), so per
// atom only the kxmax+kymax+kzmax+3 phase factors of the three axes are tabulated in shared memory (recurrence
// E[n] = E[n-1] E[1] in the structure-factor kernel, sincos(n theta) in the per-atom kernel) and a wave vector
// (nx, ny, nz) costs two complex products of table entries, conjugated for negative indices: the structure-factor
// kernel in use, and the per-atom kernel below 4096 atoms (CPH_EWALD=tables: always).  *Row walking* (per-atom sums,
// default): one thread per atom walks the (nx, ny) rows of the list with the phases in registers -- see the kernel.
// Measured at 32k atoms x 22 236 wave vectors: 4.69 ms direct, 3.69 ms tables, 2.10 ms default = 0.50 of the measured
// DFMA peak (profiles/r2z_summary.md).  The pass is bound by the fp64 pipe, not by HBM: this is the O(N K) Ewald sum,
// meant for the boxes the reference itself targets (configs 1-2); a mesh solver (PPPM) is what 1M atoms would need
// and is not built.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "cph_internal.h"

namespace {

constexpr int KTPB = 128;    // wave vectors per block of the structure-factor kernel
constexpr int TILE = 256;    // atoms staged per round
constexpr int ATPB = 256;    // 8 atoms per block in the per-atom kernel

__global__ void __launch_bounds__(KTPB)
ewald_sfac_kernel(int n, const double4 *__restrict__ xq, int K1, const double4 *__restrict__ kv, int nchunk,
                  double2 *__restrict__ part) {
  __shared__ double4 tile[TILE];
  const int k = blockIdx.x * KTPB + threadIdx.x;
  const int c = blockIdx.y;
  const int lo = (int)((long long)n * c / nchunk), hi = (int)((long long)n * (c + 1) / nchunk);
  const double4 w = k < K1 ? kv[k] : make_double4(0.0, 0.0, 0.0, 0.0);
  double sr = 0.0, si = 0.0;
  for (int base = lo; base < hi; base += TILE) {
    const int m = min(TILE, hi - base);
    __syncthreads();
    for (int t = threadIdx.x; t < m; t += KTPB) tile[t] = xq[base + t];
    __syncthreads();
    for (int t = 0; t < m; t++) {
      const double4 p = tile[t];
      double sn, co;
      sincos(w.x * p.x + w.y * p.y + w.z * p.z, &sn, &co);
      sr = fma(p.w, co, sr);
      si = fma(p.w, sn, si);
    }
  }
  if (k < K1) part[(size_t)c * K1 + k] = make_double2(sr, si);
}

__global__ void ewald_sfac_sum_kernel(int K1, int nchunk, const double2 *__restrict__ part, double2 *__restrict__ S) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K1) return;
  double sr = 0.0, si = 0.0;
  for (int c = 0; c < nchunk; c++) {
    const double2 v = part[(size_t)c * K1 + k];
    sr += v.x;
    si += v.y;
  }
  S[k] = make_double2(sr, si);
}

// self2 = 2 g/sqrt(pi), bg = pi/(g^2 V); S[K].x = sum of all charges (the zero wave vector)
__global__ void __launch_bounds__(ATPB)
ewald_atom_kernel(int n, const double4 *__restrict__ xq, int K, const double4 *__restrict__ kv,
                  const double2 *__restrict__ S, double qqrd2e, double self2, double bg, int eflag, double *f,
                  double *phi, double *eatom, double *ek) {
  const int i = (blockIdx.x * ATPB + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  const double4 p = xq[i];
  double pot = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
  for (int k = lane; k < K; k += 32) {
    const double4 w = kv[k];
    const double2 s = S[k];
    double sn, co;
    sincos(w.x * p.x + w.y * p.y + w.z * p.z, &sn, &co);
    const double u2 = 2.0 * w.w;
    pot = fma(u2, co * s.x + sn * s.y, pot);
    const double g = u2 * (sn * s.x - co * s.y);
    fx = fma(g, w.x, fx);
    fy = fma(g, w.y, fy);
    fz = fma(g, w.z, fz);
  }
  for (int o = 16; o; o >>= 1) {
    pot += __shfl_xor_sync(0xffffffffu, pot, o);
    fx += __shfl_xor_sync(0xffffffffu, fx, o);
    fy += __shfl_xor_sync(0xffffffffu, fy, o);
    fz += __shfl_xor_sync(0xffffffffu, fz, o);
  }
  if (lane == 0) {
    const double ph = qqrd2e * (pot - self2 * p.w - bg * S[K].x);
    const double c = qqrd2e * p.w;
    f[3 * (size_t)i] += c * fx;
    f[3 * (size_t)i + 1] += c * fy;
    f[3 * (size_t)i + 2] += c * fz;
    if (eflag) {
      const double e = 0.5 * p.w * ph;
      phi[i] += ph;
      eatom[i] += e;
      ek[i] = e;
    }
  }
}

// ---- factorised variants ------------------------------------------------------------------------------------------
constexpr int FTPB = 256;    // wave vectors per block
constexpr int FTA = 32;      // atoms per tile (fewer when the tables are long)

// packed wave-vector indices: nx | (ny + 512) << 10 | (nz + 512) << 20
__device__ __forceinline__ void unpack_kidx(int w, int nx1, int ny1, int &ix, int &iy, int &iz, double &sy, double &sz) {
  const int ny = ((w >> 10) & 1023) - 512, nz = ((w >> 20) & 1023) - 512;
  ix = w & 1023;
  iy = nx1 + abs(ny);
  iz = nx1 + ny1 + abs(nz);
  sy = ny < 0 ? -1.0 : 1.0;      // exp(-i n theta) = conj exp(i n theta)
  sz = nz < 0 ? -1.0 : 1.0;
}

// tab[a][0..nx1) = q_a exp(i n ux x_a), [nx1..nx1+ny1) = exp(i n uy y_a), then z: nx1 = kxmax+1, ...
__global__ void __launch_bounds__(FTPB)
ewald_sfac_fact_kernel(int n, const double4 *__restrict__ xq, int K1, const int *__restrict__ kidx, int nchunk, int ta,
                       int nx1, int ny1, int nz1, double ux, double uy, double uz, double2 *__restrict__ part) {
  extern __shared__ double2 tab[];
  const int stride = nx1 + ny1 + nz1;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  const int lo = (int)((long long)n * c / nchunk), hi = (int)((long long)n * (c + 1) / nchunk);
  int ix = 0, iy = nx1, iz = nx1 + ny1;
  double sy = 1.0, sz = 1.0;
  if (k < K1) unpack_kidx(kidx[k], nx1, ny1, ix, iy, iz, sy, sz);
  double sr = 0.0, si = 0.0;
  for (int base = lo; base < hi; base += ta) {
    const int m = min(ta, hi - base);
    __syncthreads();
    if ((int)threadIdx.x < 3 * m) {      // thread (atom a, axis d): E[n] = E[n-1] E[1]
      const int a = threadIdx.x / 3, d = threadIdx.x - 3 * a;
      const double4 p = xq[base + a];
      const double ang = d == 0 ? ux * p.x : d == 1 ? uy * p.y : uz * p.z;
      const int cnt = d == 0 ? nx1 : d == 1 ? ny1 : nz1;
      double2 *e = tab + a * stride + (d == 0 ? 0 : d == 1 ? nx1 : nx1 + ny1);
      double s1, c1;
      sincos(ang, &s1, &c1);
      double re = d == 0 ? p.w : 1.0, im = 0.0;      // the charge rides on the x factor
      e[0] = make_double2(re, im);
      for (int q = 1; q < cnt; q++) {
        const double nre = re * c1 - im * s1, nim = re * s1 + im * c1;
        re = nre;
        im = nim;
        e[q] = make_double2(re, im);
      }
    }
    __syncthreads();
    for (int a = 0; a < m; a++) {
      const double2 *e = tab + a * stride;
      const double2 ex = e[ix], ey = e[iy], ez = e[iz];
      const double eyi = sy * ey.y, ezi = sz * ez.y;
      const double tr = ex.x * ey.x - ex.y * eyi, ti = ex.x * eyi + ex.y * ey.x;
      sr += tr * ez.x - ti * ezi;
      si += tr * ezi + ti * ez.x;
    }
  }
  if (k < K1) part[(size_t)c * K1 + k] = make_double2(sr, si);
}

__global__ void __launch_bounds__(ATPB)
ewald_atom_fact_kernel(int n, const double4 *__restrict__ xq, int K, const double4 *__restrict__ kv,
                       const int *__restrict__ kidx, const double2 *__restrict__ S, int nx1, int ny1, int nz1, double ux,
                       double uy, double uz, double qqrd2e, double self2, double bg, int eflag, double *f, double *phi,
                       double *eatom, double *ek) {
  extern __shared__ double2 tab[];      // one table per warp
  const int stride = nx1 + ny1 + nz1;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * (ATPB / 32) + wl;
  if (i >= n) return;                   // whole warps leave; only __syncwarp below
  double2 *e = tab + wl * stride;
  const double4 p = xq[i];
  for (int t = lane; t < stride; t += 32) {
    const int d = t < nx1 ? 0 : t < nx1 + ny1 ? 1 : 2;
    const int q = t - (d == 0 ? 0 : d == 1 ? nx1 : nx1 + ny1);
    const double ang = (d == 0 ? ux * p.x : d == 1 ? uy * p.y : uz * p.z) * q;
    double sn, co;
    sincos(ang, &sn, &co);
    e[t] = make_double2(co, sn);
  }
  __syncwarp();
  double pot = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
  for (int k = lane; k < K; k += 32) {
    int ix, iy, iz;
    double sy, sz;
    unpack_kidx(kidx[k], nx1, ny1, ix, iy, iz, sy, sz);
    const double4 w = kv[k];
    const double2 s = S[k];
    const double2 ex = e[ix], ey = e[iy], ez = e[iz];
    const double eyi = sy * ey.y, ezi = sz * ez.y;
    const double tr = ex.x * ey.x - ex.y * eyi, ti = ex.x * eyi + ex.y * ey.x;
    const double co = tr * ez.x - ti * ezi, sn = tr * ezi + ti * ez.x;
    const double u2 = 2.0 * w.w;
    pot = fma(u2, co * s.x + sn * s.y, pot);
    const double g = u2 * (sn * s.x - co * s.y);
    fx = fma(g, w.x, fx);
    fy = fma(g, w.y, fy);
    fz = fma(g, w.z, fz);
  }
  for (int o = 16; o; o >>= 1) {
    pot += __shfl_xor_sync(0xffffffffu, pot, o);
    fx += __shfl_xor_sync(0xffffffffu, fx, o);
    fy += __shfl_xor_sync(0xffffffffu, fy, o);
    fz += __shfl_xor_sync(0xffffffffu, fz, o);
  }
  if (lane == 0) {
    const double ph = qqrd2e * (pot - self2 * p.w - bg * S[K].x);
    const double c = qqrd2e * p.w;
    f[3 * (size_t)i] += c * fx;
    f[3 * (size_t)i + 1] += c * fy;
    f[3 * (size_t)i + 2] += c * fz;
    if (eflag) {
      const double en = 0.5 * p.w * ph;
      phi[i] += ph;
      eatom[i] += en;
      ek[i] = en;
    }
  }
}

// ---- per-atom sums, row-walking variant ------------------------------------------------------------------------------
// The wave-vector list is ordered (nx, ny) row by row with nz ascending inside a row, so a row is a contiguous run
// [zlo, zhi] of nz (row descriptor: nx | (ny+512)<<10, zlo, zhi, index of its first entry).  One THREAD per atom walks
// the rows of its slice: exp(i nx thetax) and exp(i ny thetay) by sincos once per row, exp(i nz thetaz) by the recurrence
// E <- E * exp(i thetaz) along the row, everything in registers.  The per-wave-vector data (S, ug) are the same for all
// lanes of a warp: one broadcast load instead of 32.  No tables, no shuffles; rows are split into slices over
// blockIdx.y so that small boxes fill the machine, and a second kernel adds the slices in a fixed order.
constexpr int RTPB = 256;      // default block; the launch may use 64..256 (CPH_EWALD_TUNE)

__global__ void __launch_bounds__(256)
ewald_atom_rows_kernel(int n, const double4 *__restrict__ xq, int nrows, const int4 *__restrict__ rows,
                       const double4 *__restrict__ kv, const double2 *__restrict__ S, double ux, double uy, double uz,
                       int nslice, double4 *__restrict__ part) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int sl = blockIdx.y;
  const int r0 = (int)((long long)nrows * sl / nslice), r1 = (int)((long long)nrows * (sl + 1) / nslice);
  const double4 p = xq[min(i, n - 1)];
  const double tx = ux * p.x, ty = uy * p.y, tz = uz * p.z;
  double s1, c1;
  sincos(tz, &s1, &c1);
  double pot = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
  int cur_nx = -1;
  double exr = 1.0, exi = 0.0;
  for (int r = r0; r < r1; r++) {
    const int4 row = rows[r];
    const int nx = row.x & 1023, ny = ((row.x >> 10) & 1023) - 512;
    if (nx != cur_nx) {        // warp-uniform
      sincos(tx * nx, &exi, &exr);
      cur_nx = nx;
    }
    double eyr, eyi, ezr, ezi;
    sincos(ty * ny, &eyi, &eyr);
    sincos(tz * row.y, &ezi, &ezr);
    const double er = exr * eyr - exi * eyi, ei = exr * eyi + exi * eyr;     // exp(i (nx thetax + ny thetay))
    double grow = 0.0, gz = 0.0, dnz = (double)row.y;
    const double4 *kvr = kv + row.w;
    const double2 *Sr = S + row.w;
    const int len = row.z - row.y + 1;
    for (int m = 0; m < len; m++) {
      const double2 s = Sr[m];
      const double u2 = 2.0 * kvr[m].w;
      const double co = er * ezr - ei * ezi, sn = er * ezi + ei * ezr;
      pot = fma(u2, co * s.x + sn * s.y, pot);
      const double g = u2 * (sn * s.x - co * s.y);
      grow += g;
      gz = fma(g, dnz, gz);
      dnz += 1.0;
      const double nr = ezr * c1 - ezi * s1, ni = ezr * s1 + ezi * c1;
      ezr = nr;
      ezi = ni;
    }
    fx = fma(ux * nx, grow, fx);
    fy = fma(uy * ny, grow, fy);
    fz = fma(uz, gz, fz);
  }
  if (i < n) part[(size_t)sl * n + i] = make_double4(pot, fx, fy, fz);
}

__global__ void ewald_atom_combine_kernel(int n, const double4 *__restrict__ xq, int nslice,
                                          const double4 *__restrict__ part, const double2 *__restrict__ S, int K,
                                          double qqrd2e, double self2, double bg, int eflag, double *f, double *phi,
                                          double *eatom, double *ek) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double pot = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
  for (int sl = 0; sl < nslice; sl++) {
    const double4 v = part[(size_t)sl * n + i];
    pot += v.x; fx += v.y; fy += v.z; fz += v.w;
  }
  const double q = xq[i].w;
  const double ph = qqrd2e * (pot - self2 * q - bg * S[K].x);
  const double c = qqrd2e * q;
  f[3 * (size_t)i] += c * fx;
  f[3 * (size_t)i + 1] += c * fy;
  f[3 * (size_t)i + 2] += c * fz;
  if (eflag) {
    const double en = 0.5 * q * ph;
    phi[i] += ph;
    eatom[i] += en;
    ek[i] = en;
  }
}

// one block, fixed order: out = sum_i ek[i]
__global__ void __launch_bounds__(256) ewald_energy_kernel(int n, const double *__restrict__ ek, double *out) {
  __shared__ double sm[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) a += ek[i];
  sm[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sm[0];
}

}  // namespace

// wave vectors and their weights for the current box (cph_set_kspace, and again whenever cph_set_domain changes the box)
int cph_kspace_setup(cph_handle *h) {
  h->nkvec = 0;
  if (h->kspace_style != CPH_KSPACE_EWALD) return CPH_OK;
  cudaSetDevice(h->device);
  const double PI = 3.14159265358979323846;
  double L[3], unitk[3];
  for (int d = 0; d < 3; d++) {
    L[d] = h->boxhi[d] - h->boxlo[d];
    unitk[d] = 2.0 * PI / L[d];
  }
  const double V = L[0] * L[1] * L[2], g = h->g_ewald;
  double gsqmx = 0.0;
  for (int d = 0; d < 3; d++) gsqmx = std::max(gsqmx, unitk[d] * unitk[d] * h->kmax[d] * h->kmax[d]);
  gsqmx *= 1.00001;
  std::vector<double4> kv;
  std::vector<int> kidx;
  std::vector<int4> rows;      // (nx, ny) rows of the list: nx | (ny+512)<<10, first nz, last nz, index of the first entry
  for (int nx = 0; nx <= h->kmax[0]; nx++)
    for (int ny = -h->kmax[1]; ny <= h->kmax[1]; ny++) {
      const size_t row_first = kv.size();
      int zlo = 0, zhi = -1;
      for (int nz = -h->kmax[2]; nz <= h->kmax[2]; nz++) {
        if (!(nx > 0 || (nx == 0 && ny > 0) || (nx == 0 && ny == 0 && nz > 0))) continue;   // half space
        const double kx = unitk[0] * nx, ky = unitk[1] * ny, kz = unitk[2] * nz;
        const double sqk = kx * kx + ky * ky + kz * kz;
        if (sqk > gsqmx) continue;
        kv.push_back(make_double4(kx, ky, kz, 4.0 * PI / V * std::exp(-0.25 * sqk / (g * g)) / sqk));
        kidx.push_back((nx & 1023) | (((ny + 512) & 1023) << 10) | (((nz + 512) & 1023) << 20));
        if (zhi < zlo) zlo = nz;
        zhi = nz;
      }
      // the sphere and the half space both cut a row down to ONE run of consecutive nz (checked: the run is as long
      // as the number of entries the row added)
      if (zhi >= zlo) {
        if ((size_t)(zhi - zlo + 1) != kv.size() - row_first) return cph_fail(h, CPH_ERR_STATE, "ewald: broken row of wave vectors");
        rows.push_back(make_int4((nx & 1023) | (((ny + 512) & 1023) << 10), zlo, zhi, (int)row_first));
      }
    }
  h->nkvec = (int)kv.size();
  kv.push_back(make_double4(0.0, 0.0, 0.0, 0.0));   // zero wave vector: its "structure factor" is sum q
  kidx.push_back(0 | (512 << 10) | (512 << 20));
  CPH_CUDA(h, h->d_kvec.reserve(kv.size()));
  CPH_CUDA(h, h->d_kidx.reserve(kidx.size()));
  CPH_CUDA(h, cudaMemcpyAsync(h->d_kvec.p, kv.data(), kv.size() * sizeof(double4), cudaMemcpyHostToDevice, h->stream));
  CPH_CUDA(h, cudaMemcpyAsync(h->d_kidx.p, kidx.data(), kidx.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  h->kspace_nrows = (int)rows.size();
  CPH_CUDA(h, h->d_krows.reserve(rows.size() + 1));
  CPH_CUDA(h, cudaMemcpyAsync(h->d_krows.p, rows.data(), rows.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int d = 0; d < 3; d++) h->kspace_unitk[d] = unitk[d];
  // factorised kernels: phase tables of kxmax+kymax+kzmax+3 entries per atom in shared memory (48 KB without opt-in)
  const int stride = h->kmax[0] + h->kmax[1] + h->kmax[2] + 3;
  const char *mode = getenv("CPH_EWALD");
  h->kspace_tile = std::min(FTA, (int)(48 * 1024 / ((size_t)stride * sizeof(double2))));
  h->kspace_fact = !(mode && !strcmp(mode, "direct")) && h->kspace_tile >= 4 &&
                   (size_t)(ATPB / 32) * stride * sizeof(double2) <= 48 * 1024 &&
                   h->kmax[0] <= 511 && h->kmax[1] <= 511 && h->kmax[2] <= 511;
  // per-atom sums: row-walking kernel (default) or the table kernel (CPH_EWALD=tables)
  h->kspace_rows = h->kspace_fact && !(mode && !strcmp(mode, "tables"));
  // launch shapes (CPH_EWALD_TUNE="rows_block,rows_blocks_per_sm,sfac_block,sfac_tile,sfac_blocks_per_sm" overrides)
  // defaults from the sweep on config 2 (profiles/r2z6_ewald_tune.json): rows 256 threads x 6 blocks per SM (1.12 ms
  // against 1.44 at 128 x 4), structure factors 256 threads, 32-atom tiles, 6 blocks per SM (0.98 against 1.11 ms)
  int tune[5] = {RTPB, 6, FTPB, FTA, 6};
  if (const char *t = getenv("CPH_EWALD_TUNE")) sscanf(t, "%d,%d,%d,%d,%d", &tune[0], &tune[1], &tune[2], &tune[3], &tune[4]);
  h->kspace_tune[0] = std::max(32, std::min(256, tune[0] & ~31));
  h->kspace_tune[1] = std::max(1, std::min(16, tune[1]));
  h->kspace_tune[2] = std::max(96, std::min(256, tune[2] & ~31));
  h->kspace_tune[3] = std::max(4, std::min(h->kspace_tile, std::min(tune[3], h->kspace_tune[2] / 3)));
  h->kspace_tune[4] = std::max(1, std::min(16, tune[4]));
  h->kspace_self2 = 2.0 * g / 1.77245385090551602729;
  h->kspace_bg = PI / (g * g * V);
  return CPH_OK;
}

// after the pair pass (and the other terms that add to it): the reciprocal sum at the current positions and charges
int cph_launch_kspace(cph_handle *h, int eflag) {
  if (h->kspace_style != CPH_KSPACE_EWALD) return CPH_OK;
  const int n = h->nlocal, K = h->nkvec, K1 = K + 1;
  cudaStream_t st = h->stream;
  const bool fact = h->kspace_fact;
  const int ktpb = fact ? h->kspace_tune[2] : KTPB, tile = fact ? h->kspace_tune[3] : TILE;
  const int kblocks = (K1 + ktpb - 1) / ktpb;
  // enough (wave-vector block, atom chunk) pairs for a few blocks per SM, chunks of at least one tile
  const int bps = fact ? h->kspace_tune[4] : 4;
  const int nchunk = std::max(1, std::min(std::min(64, (n + tile - 1) / tile), (bps * h->num_sms + kblocks - 1) / kblocks));
  const int nx1 = h->kmax[0] + 1, ny1 = h->kmax[1] + 1, nz1 = h->kmax[2] + 1;
  const size_t stride_bytes = (size_t)(nx1 + ny1 + nz1) * sizeof(double2);
  const double *u = h->kspace_unitk;
  {
  ProfScope ps(h, 10);      // structure factors
  CPH_CUDA(h, h->d_sfac_part.reserve((size_t)nchunk * K1));
  CPH_CUDA(h, h->d_sfac.reserve((size_t)K1));
  CPH_CUDA(h, h->d_ekspace.reserve((size_t)n + 2));
  if (fact)
    ewald_sfac_fact_kernel<<<dim3(kblocks, nchunk), ktpb, tile * stride_bytes, st>>>(
        n, h->d_xq.p, K1, h->d_kidx.p, nchunk, tile, nx1, ny1, nz1, u[0], u[1], u[2], h->d_sfac_part.p);
  else
    ewald_sfac_kernel<<<dim3(kblocks, nchunk), KTPB, 0, st>>>(n, h->d_xq.p, K1, h->d_kvec.p, nchunk, h->d_sfac_part.p);
  ewald_sfac_sum_kernel<<<(K1 + 255) / 256, 256, 0, st>>>(K1, nchunk, h->d_sfac_part.p, h->d_sfac.p);
  h->nlaunch += 2;
  CPH_CUDA(h, cudaGetLastError());
  }
  // several ranks: every rank summed over the atoms it owns
  CPH_TRY(cph_comm_allreduce(h, reinterpret_cast<double *>(h->d_sfac.p), 2 * K1));
  ProfScope ps(h, 11);      // per-atom sums
  if (n > 0) {
    const int ablocks = (n + ATPB / 32 - 1) / (ATPB / 32);
    // small boxes: the warp-per-atom table kernel has more parallelism (0.025 against 0.042 ms at 3k atoms)
    if (h->kspace_rows && n >= 4096) {
      const int rtpb = h->kspace_tune[0];
      const int rblocks = (n + rtpb - 1) / rtpb;
      const int nslice = std::max(1, std::min(std::min(64, h->kspace_nrows),
                                              (h->kspace_tune[1] * h->num_sms + rblocks - 1) / rblocks));
      CPH_CUDA(h, h->d_kpart.reserve((size_t)nslice * n));
      ewald_atom_rows_kernel<<<dim3(rblocks, nslice), rtpb, 0, st>>>(n, h->d_xq.p, h->kspace_nrows, h->d_krows.p,
                                                                    h->d_kvec.p, h->d_sfac.p, u[0], u[1], u[2], nslice,
                                                                    h->d_kpart.p);
      ewald_atom_combine_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, h->d_xq.p, nslice, h->d_kpart.p, h->d_sfac.p, K,
                                                                h->qqrd2e, h->kspace_self2, h->kspace_bg, eflag, h->d_f.p,
                                                                h->d_phi.p, h->d_eatom.p, h->d_ekspace.p);
      h->nlaunch++;
    } else if (fact)
      ewald_atom_fact_kernel<<<ablocks, ATPB, (ATPB / 32) * stride_bytes, st>>>(
          n, h->d_xq.p, K, h->d_kvec.p, h->d_kidx.p, h->d_sfac.p, nx1, ny1, nz1, u[0], u[1], u[2], h->qqrd2e,
          h->kspace_self2, h->kspace_bg, eflag, h->d_f.p, h->d_phi.p, h->d_eatom.p, h->d_ekspace.p);
    else
      ewald_atom_kernel<<<ablocks, ATPB, 0, st>>>(n, h->d_xq.p, K, h->d_kvec.p, h->d_sfac.p, h->qqrd2e, h->kspace_self2,
                                                 h->kspace_bg, eflag, h->d_f.p, h->d_phi.p, h->d_eatom.p, h->d_ekspace.p);
    h->nlaunch++;
  }
  if (eflag) {
    ewald_energy_kernel<<<1, 256, 0, st>>>(n, h->d_ekspace.p, h->d_ekspace.p + n);
    h->nlaunch++;
  }
  CPH_CUDA(h, cudaGetLastError());
  return CPH_OK;
}

int cph_kspace_energy(cph_handle *h, double *out) {
  *out = 0.0;
  if (h->kspace_style != CPH_KSPACE_EWALD || h->d_ekspace.cap < (size_t)h->nlocal + 1) return CPH_OK;
  CPH_CUDA(h, cudaMemcpyAsync(out, h->d_ekspace.p + h->nlocal, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return CPH_OK;
}

void cph_kspace_release(cph_handle *h) {
  h->d_kvec.release();
  h->d_kidx.release();
  h->d_krows.release();
  h->d_kpart.release();
  h->d_sfac_part.release();
  h->d_sfac.release();
  h->d_ekspace.release();
}
