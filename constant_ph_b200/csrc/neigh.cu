// neigh.cu -- GPU-resident Verlet neighbour list (K1) and ghost atoms (K6 local part).
//
// Stands in for LAMMPS Neighbor/NeighList + Comm::borders, which the reference only
// gestures at: `init_list` is declared (fix_constant_pH.h:40) and never defined, and no
// list is ever requested.  Semantics restated from SURVEY.md Appendix A:
// list cutoff = max pair cutoff + skin; entries carry the special-bond class in the top
// two bits; pairs whose lj and coul weights are both zero are dropped, except under
// coul/dsf where they are kept for the damped-term correction.
//
// Layout in HBM (internal order): owned atoms [0,nlocal) sorted by (cell, tag), ghosts
// [nlocal,nall) sorted the same way; positions+charge packed as double4 (32 B, one
// sector) so a neighbour gather is one aligned 32-byte request; the list is a fixed-pitch
// matrix neigh[i*rowcap + k] so a warp reads its atom's row with 128-byte coalesced loads.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "cph_internal.h"

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ int cell_of(const Grid &g, double x, double y, double z) {
  int cx = (int)floor((x - g.lo[0]) * g.inv[0]);
  int cy = (int)floor((y - g.lo[1]) * g.inv[1]);
  int cz = (int)floor((z - g.lo[2]) * g.inv[2]);
  cx = min(max(cx, 0), g.n[0] - 1);
  cy = min(max(cy, 0), g.n[1] - 1);
  cz = min(max(cz, 0), g.n[2] - 1);
  return (cz * g.n[1] + cy) * g.n[0] + cx;
}

// how far outside the sub-box owned atoms have drifted (max over dims), as float bits
__global__ void drift_kernel(int n, const double4 *__restrict__ xq, double3 lo, double3 hi,
                             unsigned int *flags) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  float d = 0.f;
  if (k < n) {
    double4 p = xq[k];
    double e = fmax(fmax(lo.x - p.x, p.x - hi.x), fmax(fmax(lo.y - p.y, p.y - hi.y), fmax(lo.z - p.z, p.z - hi.z)));
    d = e > 0 ? __double2float_ru(e) : 0.f;
  }
  for (int o = 16; o; o >>= 1) d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o));
  if ((threadIdx.x & 31) == 0 && d > 0.f) atomicMax(flags + 2, __float_as_uint(d));
}

__global__ void key_kernel(int n, const double4 *__restrict__ xq, const int *__restrict__ tag, Grid g,
                           unsigned long long *keys, int *vals) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double4 p = xq[k];
  unsigned long long c = (unsigned long long)cell_of(g, p.x, p.y, p.z);
  keys[k] = (c << 32) | (unsigned int)tag[k];
  vals[k] = k;
}

template <typename T>
__global__ void gather_kernel(int n, const int *__restrict__ idx, const T *__restrict__ src, T *dst) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[idx[k]];
}

__global__ void after_sort_kernel(int n, const unsigned long long *__restrict__ keys, const int *__restrict__ perm,
                                  const double4 *__restrict__ xq, int *inv, double *xbuild, int *cellid) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  if (inv) inv[perm[k]] = k;
  if (xbuild) {
    double4 p = xq[k];
    xbuild[3 * (size_t)k] = p.x;
    xbuild[3 * (size_t)k + 1] = p.y;
    xbuild[3 * (size_t)k + 2] = p.z;
  }
  cellid[k] = (int)(keys[k] >> 32);
}

// cell_start[c] = first sorted atom whose cell >= c; cell_start[ncell] = n
__global__ void cell_start_kernel(int n, int ncell, const int *__restrict__ cellid, int *start) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n) return;
  int cprev = (k == 0) ? -1 : cellid[k - 1];
  int ccur = (k == n) ? ncell : cellid[k];
  for (int c = cprev + 1; c <= ccur; c++) start[c] = k;
}

struct GhostDirs {
  double shift[27][3];   // translation applied to the copy
  int active[27];        // 0: no copy in this direction, 1: periodic self image, 2: copy goes to another rank
  int imgcode[27];       // (ix+1)+3(iy+1)+9(iz+1) of the periodic shift
  int peer[27];          // rank the copy goes to (active == 2)
  int from[27];          // rank whose direction-d copies arrive here
};

struct DirTable {        // record layout of the direction-sorted ghost records
  int start[28];         // first record of direction d
  int out[27];           // output offset of direction d (send buffer for remote, ghost slot for local), -1 = unused
};

// bit d of the result is set when the atom must be copied in direction d
__device__ __forceinline__ unsigned int ghost_mask(double4 p, const double3 lo, const double3 hi, double gc,
                                                   const GhostDirs &gd) {
  int up[3] = {p.x >= hi.x - gc, p.y >= hi.y - gc, p.z >= hi.z - gc};
  int dn[3] = {p.x < lo.x + gc, p.y < lo.y + gc, p.z < lo.z + gc};
  unsigned int m = 0;
  for (int d = 0; d < 27; d++) {
    if (d == 13 || !gd.active[d]) continue;
    int dx = d % 3 - 1, dy = (d / 3) % 3 - 1, dz = d / 9 - 1;
    bool ok = (dx == 0 || (dx > 0 ? up[0] : dn[0])) && (dy == 0 || (dy > 0 ? up[1] : dn[1])) &&
              (dz == 0 || (dz > 0 ? up[2] : dn[2]));
    if (ok) m |= 1u << d;
  }
  return m;
}

__global__ void ghost_count_kernel(int n, const double4 *__restrict__ xq, double3 lo, double3 hi, double gc,
                                   GhostDirs gd, int *count) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  count[k] = __popc(ghost_mask(xq[k], lo, hi, gc, gd));
}

__global__ void ghost_fill_kernel(int n, const double4 *__restrict__ xq, double3 lo, double3 hi, double gc,
                                  GhostDirs gd, const int *__restrict__ offset, int *src, unsigned long long *dirkey,
                                  int *vals) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  unsigned int m = ghost_mask(xq[k], lo, hi, gc, gd);
  int o = offset[k];
  while (m) {
    int d = __ffs(m) - 1;
    m &= m - 1;
    src[o] = k;
    dirkey[o] = (unsigned long long)d;
    vals[o] = o;
    o++;
  }
}

__global__ void dir_of_key_kernel(int n, const unsigned long long *__restrict__ keys, int *dir) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dir[k] = (int)keys[k];
}

// K6 pack: copies that go to other ranks.  Records are sorted by direction; tab.out[d] is the
// offset of direction d in the send buffer.  meta (build only): {type | imgcode<<8, tag, mask, molecule}
__global__ void halo_pack_kernel(int nrec, const int *__restrict__ rsrc, const int *__restrict__ rdir, DirTable tab,
                                 GhostDirs gd, const double4 *__restrict__ xq, const int *__restrict__ type,
                                 const int *__restrict__ tag, const int *__restrict__ mask,
                                 const int *__restrict__ mol, double4 *sendx, int4 *sendmeta) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrec) return;
  const int d = rdir[r];
  if (gd.active[d] != 2) return;
  const int o = tab.out[d] + (r - tab.start[d]);
  const int s = rsrc[r];
  double4 p = xq[s];
  p.x += gd.shift[d][0];
  p.y += gd.shift[d][1];
  p.z += gd.shift[d][2];
  sendx[o] = p;
  if (sendmeta) sendmeta[o] = make_int4(type[s] | (gd.imgcode[d] << 8), tag[s], mask[s], mol ? mol[s] : 0);
}

struct PeerTab {
  double4 *dst[27];      // dst[d][r - start[d]]: where record r of direction d lands in the neighbour's buffer
};

struct MailFlags {
  int P;                                   // 0: no mailboxes
  unsigned long long seq;
  unsigned int *dst[CPH_MAIL_MAXP];        // my flag slot in every rank's mailbox (my own included)
};

// K6 fused pack + NVLink store: every copy that belongs to another rank is shifted across the
// periodic boundary and stored straight into that rank's receive buffer (mapped peer memory).
// With mailboxes the same kernel also publishes this rank's decision flags: the last block to
// finish (ticket counter; every block has fenced its stores system-wide by then) writes the six
// flag words and then the sequence number into its slot on every rank -- the consumer that sees
// the sequence number therefore also sees the positions.
__global__ void halo_pack_peer_kernel(int nrec, const int *__restrict__ rsrc, const int *__restrict__ rdir,
                                      DirTable tab, GhostDirs gd, const double4 *__restrict__ xq, PeerTab pt,
                                      MailFlags mf, const unsigned int *__restrict__ flags, unsigned int *ticket) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < nrec) {
    const int d = rdir[r];
    if (gd.active[d] == 2) {
      double4 p = xq[rsrc[r]];
      p.x += gd.shift[d][0];
      p.y += gd.shift[d][1];
      p.z += gd.shift[d][2];
      pt.dst[d][r - tab.start[d]] = p;
    }
  }
  __threadfence_system();   // the stores must have landed before this rank joins the next collective
  if (mf.P == 0) return;
  __shared__ unsigned int s_last;
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if ((int)threadIdx.x < mf.P) {
    unsigned int *dst = mf.dst[threadIdx.x];
    for (int k = 0; k < 6; k++) dst[k] = flags[k];
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(dst + 8) = mf.seq;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// The other half of the one-shot all-reduce: wait until every rank's flags of this step have arrived
// (lane p watches rank p's slot in MY mailbox), take the maximum, leave it where the ncclAllReduce left it.
__global__ void flags_gather_kernel(const unsigned int *mail, int P, unsigned long long seq, unsigned int *out,
                                    unsigned int *status) {
  const int lane = threadIdx.x;
  unsigned int v[6] = {0, 0, 0, 0, 0, 0};
  if (lane < P) {
    const unsigned int *slot = mail + (size_t)lane * CPH_MAIL_FSLOT;
    const volatile unsigned long long *sq = reinterpret_cast<const volatile unsigned long long *>(slot + 8);
    const long long t0 = clock64();
    while (*sq != seq) {
      if (clock64() - t0 > 400000000000LL) { atomicOr(status, 1u); break; }   // minutes: a rank is gone
      __nanosleep(200);
    }
    __threadfence_system();
    for (int k = 0; k < 6; k++) v[k] = reinterpret_cast<const volatile unsigned int *>(slot)[k];
  }
  for (int k = 0; k < 6; k++) {
    unsigned int m = v[k];
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) out[k] = m;
  }
}

// unsorted ghost candidates: local self images first (direction order), then received copies
__global__ void ghost_candidates_kernel(int nrec, const int *__restrict__ rsrc, const int *__restrict__ rdir,
                                        DirTable tab, GhostDirs gd, int nloc, int nrecv,
                                        const double4 *__restrict__ xq, const int *__restrict__ tag,
                                        const double4 *__restrict__ recvx, const int4 *__restrict__ recvmeta, Grid g,
                                        int *gsrc, int *gcode, unsigned long long *keys, int *vals) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < nrec) {
    const int d = rdir[r];
    if (gd.active[d] == 1) {
      const int o = tab.out[d] + (r - tab.start[d]);
      const int s = rsrc[r];
      const double4 p = xq[s];
      const unsigned long long c = (unsigned long long)cell_of(g, p.x + gd.shift[d][0], p.y + gd.shift[d][1],
                                                               p.z + gd.shift[d][2]);
      gsrc[o] = s;
      gcode[o] = d;
      keys[o] = (c << 32) | (unsigned int)tag[s];
      vals[o] = o;
    }
  }
  if (r < nrecv) {
    const int o = nloc + r;
    const double4 p = recvx[r];
    const int4 m = recvmeta[r];
    const unsigned long long c = (unsigned long long)cell_of(g, p.x, p.y, p.z);
    gsrc[o] = -1 - r;
    gcode[o] = 32 + (m.x >> 8);
    keys[o] = (c << 32) | (unsigned int)m.y;
    vals[o] = o;
  }
}

// xq/type/tag/mask/molecule of ghost g from its owner (a local self image: owned atom + shift) or
// from the receive buffer (a copy another rank sent, already shifted); used at build (all fields)
// and every step (xq only)
__global__ void ghost_copy_kernel(int nghost, int nlocal, const int *__restrict__ src, const int *__restrict__ code,
                                  GhostDirs gd, const double4 *__restrict__ recvx, const int4 *__restrict__ recvmeta,
                                  double4 *xq, int *type, int *tag, int *mask, int *mol, int all) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  const int s = src[g];
  double4 p;
  if (s >= 0) {
    const int d = code[g];
    p = xq[s];
    p.x += gd.shift[d][0];
    p.y += gd.shift[d][1];
    p.z += gd.shift[d][2];
    if (all) {
      type[nlocal + g] = type[s];
      tag[nlocal + g] = tag[s];
      mask[nlocal + g] = mask[s];
      if (mol) mol[nlocal + g] = mol[s];
    }
  } else {
    p = recvx[-1 - s];
    if (all) {
      const int4 m = recvmeta[-1 - s];
      type[nlocal + g] = m.x & 255;
      tag[nlocal + g] = m.y;
      mask[nlocal + g] = m.z;
      if (mol) mol[nlocal + g] = m.w;
    }
  }
  xq[nlocal + g] = p;
}

__global__ void mol_owned_kernel(int n, const int *__restrict__ perm, const int *__restrict__ molecule, int *mol) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) mol[k] = molecule[perm[k]];
}

// build record of atom j / store of row entry `pos`: one IMAD.WIDE.U32 each for the address
__device__ __forceinline__ float4 ld_xb(const float4 *xb, int j) {
  unsigned long long a;
  asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(j), "l"(xb));
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a));
  return v;
}
__device__ __forceinline__ void st_row(unsigned long long rowp, unsigned int pos, int j) {
  unsigned long long a;
  asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(a) : "r"(pos), "l"(rowp));
  asm volatile("st.global.b32 [%0], %1;" ::"l"(a), "r"(j) : "memory");
}

// One warp per owned atom.  Lanes sweep, for each of the 5x5 (y,z) cell rows around the atom,
// the x-run of cells that can hold a neighbour (contiguous in the sorted order; the run is
// trimmed to the chord of the cutoff sphere at that row), test candidates on the packed fp32
// record with a +-margin band, re-test the few borderline ones in fp64 (so membership is
// exactly the fp64 rule), and ballot-compact accepted candidates into the atom's row:
// ordinary neighbours from the front, special-bond partners from the back.
__global__ void __launch_bounds__(TPB)
list_build_kernel(int nlocal, const double4 *__restrict__ xq, const float4 *__restrict__ xt,
                  const int *__restrict__ tag, const int *__restrict__ perm, const int *__restrict__ nspecial,
                  const int *__restrict__ special, int maxspecial, Grid g, const int *__restrict__ start_o,
                  const int *__restrict__ start_g, int3 gl, double rlist2, float margin, int dropmask,
                  int rowcap, int dummy, int *neigh, int *numneigh,
                  int *numspec, unsigned int *flags, unsigned long long *stats) {
  const int lane = threadIdx.x & 31;
  // the atom index is the same in every lane; taking it from lane 0 lets ptxas SEE that (warp-uniform control flow
  // below: no divergence guards around the votes, loop bounds in uniform registers)
  const int i = __shfl_sync(0xffffffffu, (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), 0);
  if (i >= nlocal) return;
  const double4 pi = xq[i];
  const float4 pti = xt[i];
  const int ci = cell_of(g, pi.x, pi.y, pi.z);
  const int cx = ci % g.n[0], cy = (ci / g.n[0]) % g.n[1], cz = ci / (g.n[0] * g.n[1]);
  // special partners of i (tags), by class
  const int ic = perm[i];
  int ns1 = 0, ns2 = 0, ns3 = 0;
  if (nspecial) {
    ns1 = nspecial[3 * ic];
    ns2 = nspecial[3 * ic + 1];
    ns3 = min(nspecial[3 * ic + 2], maxspecial);
  }
  const int *sp = special ? special + (size_t)ic * maxspecial : nullptr;
  // xt.w here is the molecule id (0 = unknown): only same-molecule candidates can be special partners
  const int moli = __float_as_int(pti.w);
  int *row = neigh + (size_t)i * rowcap;
  const unsigned long long rowp = (unsigned long long)row;
  int cnt = 0, nsp = 0;
  const float rl2 = (float)rlist2;
  const float lo2 = rl2 - margin, hi2 = rl2 + margin;
  const float wy = (float)(1.0 / g.inv[1]), wz = (float)(1.0 / g.inv[2]);
  const float ivx = (float)g.inv[0];
  // ghost cells live in the outer `gl` cell layers of the grid (everything outside the sub-box);
  // an atom whose +-2 stencil stays inside the interior never meets a ghost: skip that half of the loops
  const bool near_ghost = cx < gl.x + 2 || cx >= g.n[0] - gl.x - 2 || cy < gl.y + 2 || cy >= g.n[1] - gl.y - 2 ||
                          cz < gl.z + 2 || cz >= g.n[2] - gl.z - 2;
  // The 25 rows' candidate runs are worked out by 25 lanes AT ONCE (one pass of arithmetic and one round of
  // cell-start loads instead of 25 dependent ones), then handed out by shuffles in the fixed order
  // dz, dy, owned-before-ghost -- the order the rows have always had.
  int run_so = 0, run_eo = 0, run_sg = 0, run_eg = 0;
  if (lane < 25) {
    const int dz = lane / 5 - 2, dy = lane % 5 - 2;
    const int z = cz + dz, y = cy + dy;
    if (z >= 0 && z < g.n[2] && y >= 0 && y < g.n[1]) {
      // distance from the atom to the z-slab / y-slab of that cell row (xt is relative to the grid origin)
      const float gz = dz == 0 ? 0.f : (dz > 0 ? z * wz - pti.z : pti.z - (z + 1) * wz);
      const float gy = dy == 0 ? 0.f : (dy > 0 ? y * wy - pti.y : pti.y - (y + 1) * wy);
      const float rem = hi2 - fmaxf(gz, 0.f) * fmaxf(gz, 0.f) - fmaxf(gy, 0.f) * fmaxf(gy, 0.f);
      if (rem >= 0.f) {
        const float xr = sqrtf(rem) * 1.0001f + 1e-3f;
        const int x0 = max((int)floorf((pti.x - xr) * ivx), 0), x1 = min((int)floorf((pti.x + xr) * ivx), g.n[0] - 1);
        if (x1 >= x0) {
          const int c0 = (z * g.n[1] + y) * g.n[0];
          run_so = start_o[c0 + x0];
          run_eo = start_o[c0 + x1 + 1];
          if (near_ghost) {
            run_sg = start_g[c0 + x0];
            run_eg = start_g[c0 + x1 + 1];
          }
        }
      }
    }
  }
  const unsigned int ltmask = (1u << lane) - 1;
  auto sweep = [&](const int s, const int e, const int base) {
    for (int p0 = s; p0 < e; p0 += 32) {
      const int p = p0 + lane;
      // lanes past the end of the run re-read its last record (a valid address) and are masked out below: the
      // test runs on predicates only, no divergent region around the load
      const int j = base + min(p, e - 1);
      const float4 pj = ld_xb(xt, j);
      const float fx = pti.x - pj.x, fy = pti.y - pj.y, fz = pti.z - pj.z;
      const float r2 = fmaf(fx, fx, fmaf(fy, fy, fz * fz));
      bool ok = p < e && r2 < hi2 && j != i;
      if (ok && r2 > lo2) {   // borderline (rare): decide in fp64, exactly as the reference rule
        const double4 qj = xq[j];
        const double dx = pi.x - qj.x, dy_ = pi.y - qj.y, dz_ = pi.z - qj.z;
        ok = dx * dx + dy_ * dy_ + dz_ * dz_ < rlist2;
      }
      // candidate for a special-bond partner (same molecule): rare
      const bool spc = ok && ns3 != 0 && __float_as_int(pj.w) == moli;
      if (__any_sync(0xffffffffu, spc)) {
        // slow path (a few chunks per atom): look the tag up in special[i]; special-bond partners
        // fill the row from the back, the rest of the chunk goes on to the fast path below
        int sb = 0;
        if (spc) {
          const int tj = tag[j];
          for (int k = 0; k < ns3; k++)
            if (sp[k] == tj) { sb = k < ns1 ? 1 : (k < ns2 ? 2 : 3); break; }
          if ((dropmask >> sb) & 1) { ok = false; sb = 0; }   // both weights zero (and not dsf): not stored
        }
        const unsigned int ms = __ballot_sync(0xffffffffu, ok && sb);
        const int poss = nsp + __popc(ms & ltmask);
        if (ok && sb && poss < rowcap) row[rowcap - 1 - poss] = j | (sb << CPH_SBSHIFT);
        nsp += __popc(ms);
        if (sb) ok = false;
      }
      // ordinary neighbours fill the row from the front
      const unsigned int m = __ballot_sync(0xffffffffu, ok);
      const unsigned int pos = (unsigned int)(cnt + __popc(m & ltmask));
      if (ok && pos < (unsigned int)rowcap) st_row(rowp, pos, j);
      cnt += __popc(m);
    }
  };
  auto uni = [](int v, int r) { return __shfl_sync(0xffffffffu, __shfl_sync(0xffffffffu, v, r), 0); };
  for (int r = 0; r < 25; r++) {
    sweep(uni(run_so, r), uni(run_eo, r), 0);
    if (near_ghost) sweep(uni(run_sg, r), uni(run_eg, r), nlocal);
  }
  // pad the row to a whole 128-entry tile with the far-away dummy atom, so the pair kernel
  // streams full tiles and needs no tail logic
  const int padded = (cnt + 127) & ~127;
  const bool fits = padded + nsp <= rowcap;
  if (fits)
    for (int k = cnt + lane; k < padded; k += 32) row[k] = dummy;
  if (lane == 0) {
    numneigh[i] = fits ? cnt : 0;
    numspec[i] = fits ? nsp : 0;
    if (!fits) atomicMax(flags + 1, (unsigned int)(padded + nsp));
    if (nsp > 32) atomicMax(flags + 7, (unsigned int)nsp);   // the evaluation kernel takes one special partner per lane
    atomicAdd(stats, (unsigned long long)(cnt + nsp));
    if (nsp) atomicAdd(stats + 1, (unsigned long long)nsp);
    atomicMax(flags + 3, (unsigned int)(cnt + nsp));
  }
}

// fp32 build records: position relative to the grid origin + molecule id (0 when the caller gave none)
__global__ void xb_kernel(int nall, const double4 *__restrict__ xq, const int *__restrict__ mol, double3 origin,
                          float4 *xb) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nall) return;
  double4 p = xq[k];
  int m = (mol && k < nall) ? mol[k] : 0;
  xb[k] = make_float4((float)(p.x - origin.x), (float)(p.y - origin.y), (float)(p.z - origin.z), __int_as_float(m));
}

// modify_water: where the owned buffer atoms sit in the internal order
__global__ void water_map_kernel(int nlocal, const int *__restrict__ tag, const int *__restrict__ mask, int Wbit,
                                 int nw, const int *__restrict__ wtag, int *wlocal) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nlocal || !(mask[k] & Wbit)) return;
  for (int s = 0; s < nw; s++)
    if (wtag[s] == tag[k]) wlocal[s] = k;
}

// site bookkeeping: owned atom -> titration entry by binary search of its tag
__global__ void site_map_kernel(int nlocal, const int *__restrict__ tag, const int *__restrict__ mask, int ntitr,
                                const int *__restrict__ tsorted, const int *__restrict__ entry_of_sorted,
                                const int *__restrict__ titr_site, int implicit_site, int Hbit,
                                int *site_of, int *titr_of, int *titr_local) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nlocal) return;
  int t = tag[k], lo = 0, hi = ntitr - 1, e = -1;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1;
    int v = tsorted[mid];
    if (v == t) { e = entry_of_sorted[mid]; break; }
    if (v < t) lo = mid + 1; else hi = mid - 1;
  }
  int s = -1;
  if (e >= 0) { s = titr_site[e]; titr_local[e] = k; }
  else if (implicit_site && (mask[k] & Hbit)) s = 0;
  site_of[k] = s;
  titr_of[k] = e;
}

__global__ void hflag_kernel(int nlocal, const int *__restrict__ mask, const int *__restrict__ site_of, int Hbit,
                             int *flag) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nlocal) flag[k] = ((mask[k] & Hbit) && site_of[k] >= 0) ? 1 : 0;
}
__global__ void hlist_kernel(int nlocal, const int *__restrict__ flag, const int *__restrict__ off, int *hlist) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nlocal && flag[k]) hlist[off[k]] = k;
}

__global__ void fill_int_kernel(int n, int *a, int v) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) a[k] = v;
}

inline int nblk(int n) { return (n + TPB - 1) / TPB; }

int sort_pairs(cph_handle *h, int n, DevBuf<unsigned long long> &kin, DevBuf<unsigned long long> &kout,
               DevBuf<int> &vin, DevBuf<int> &vout, int end_bit) {
  size_t tmp = 0;
  CPH_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin.p, kout.p, vin.p, vout.p, n, 0, end_bit, h->stream));
  CPH_CUDA(h, h->d_cubtmp.reserve(tmp));
  CPH_CUDA(h, cub::DeviceRadixSort::SortPairs(h->d_cubtmp.p, tmp, kin.p, kout.p, vin.p, vout.p, n, 0, end_bit, h->stream));
  return 0;
}

int exclusive_scan(cph_handle *h, int n, int *in, int *out) {
  size_t tmp = 0;
  CPH_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, h->stream));
  CPH_CUDA(h, h->d_cubtmp.reserve(tmp));
  CPH_CUDA(h, cub::DeviceScan::ExclusiveSum(h->d_cubtmp.p, tmp, in, out, n, h->stream));
  return 0;
}

// the 27 directions as seen from the rank at grid position `myloc` (any rank's, not only this one's: the steady-state
// rebuild derives every rank's receive layout locally from one all-gather of the send counts)
void make_ghost_dirs_at(const cph_handle *h, const int *myloc, GhostDirs &gd) {
  auto rank_of = [&](const int *loc) { return (loc[2] * h->procgrid[1] + loc[1]) * h->procgrid[0] + loc[0]; };
  for (int d = 0; d < 27; d++) {
    int dv[3] = {d % 3 - 1, (d / 3) % 3 - 1, d / 9 - 1};
    int active = (d != 13), img[3] = {0, 0, 0}, to[3], fr[3];
    bool remote = false;
    for (int k = 0; k < 3; k++) {
      gd.shift[d][k] = 0.0;
      to[k] = fr[k] = myloc[k];
      if (dv[k] == 0) continue;
      const int pg = h->procgrid[k];
      int nb = myloc[k] + dv[k], nf = myloc[k] - dv[k];
      if (!h->periodic[k]) {
        // non-periodic: a copy only exists if there is a neighbour rank in that direction
        if (nb < 0 || nb >= pg) active = 0;
        // (the mirror test for `from` is done by the sender; counts of absent senders are zero)
      } else {
        // periodic wrap: the copy lands across the box
        if (dv[k] > 0 && myloc[k] == pg - 1) { gd.shift[d][k] = -(h->boxhi[k] - h->boxlo[k]); img[k] = -1; }
        if (dv[k] < 0 && myloc[k] == 0) { gd.shift[d][k] = (h->boxhi[k] - h->boxlo[k]); img[k] = 1; }
      }
      to[k] = ((nb % pg) + pg) % pg;
      fr[k] = ((nf % pg) + pg) % pg;
      if (pg > 1) remote = true;
    }
    gd.active[d] = active ? (remote ? 2 : 1) : 0;
    gd.imgcode[d] = (img[0] + 1) + 3 * (img[1] + 1) + 9 * (img[2] + 1);
    gd.peer[d] = rank_of(to);
    gd.from[d] = rank_of(fr);
    // does the rank at `fr` really send in direction d?  (non-periodic edges)
    bool from_ok = (d != 13);
    for (int k = 0; k < 3; k++)
      if (dv[k] != 0 && !h->periodic[k]) {
        int nf = myloc[k] - dv[k];
        if (nf < 0 || nf >= h->procgrid[k]) from_ok = false;
      }
    if (!from_ok || !remote) gd.from[d] = -1;
  }
}

void make_ghost_dirs(const cph_handle *h, GhostDirs &gd) { make_ghost_dirs_at(h, h->myloc, gd); }

template <typename T>
int permute_buf(cph_handle *h, int n, const int *idx, DevBuf<T> &buf, DevBuf<T> &tmp, size_t total) {
  CPH_CUDA(h, tmp.reserve(total));
  gather_kernel<T><<<nblk(n), TPB, 0, h->stream>>>(n, idx, buf.p, tmp.p);
  std::swap(buf.p, tmp.p);
  std::swap(buf.cap, tmp.cap);
  return 0;
}

}  // namespace

// exchange the packed copies with the spatial neighbours (grouped ncclSend/ncclRecv, one message
// per direction and array); with_meta: also the build-time records
static int halo_exchange(cph_handle *h, const GhostDirs &, bool with_meta) {
  // one message per neighbour RANK (all directions that point at it are contiguous in the
  // staging buffers): on a 2x2x2 grid 7 sends + 7 receives instead of 26 + 26
  const void *sb[64]; void *rb[64]; size_t sn[64], rn[64]; int peers[64];
  int np = 0;
  for (size_t k = 0; k < h->peer_rank.size() && np < 60; k++) {
    if (h->peer_scnt[k]) {
      peers[np] = h->peer_rank[k]; sb[np] = h->d_sendx.p + h->peer_soff[k]; sn[np] = (size_t)h->peer_scnt[k] * sizeof(double4);
      rb[np] = nullptr; rn[np] = 0; np++;
      if (with_meta) {
        peers[np] = h->peer_rank[k]; sb[np] = h->d_sendmeta.p + h->peer_soff[k]; sn[np] = (size_t)h->peer_scnt[k] * sizeof(int4);
        rb[np] = nullptr; rn[np] = 0; np++;
      }
    }
    if (h->peer_rcnt[k]) {
      peers[np] = h->peer_rank[k]; rb[np] = h->d_recvx.p + (size_t)h->halo_parity * h->recv_half + h->peer_roff[k]; rn[np] = (size_t)h->peer_rcnt[k] * sizeof(double4);
      sb[np] = nullptr; sn[np] = 0; np++;
      if (with_meta) {
        peers[np] = h->peer_rank[k]; rb[np] = h->d_recvmeta.p + h->peer_roff[k]; rn[np] = (size_t)h->peer_rcnt[k] * sizeof(int4);
        sb[np] = nullptr; sn[np] = 0; np++;
      }
    }
  }
  if (np) CPH_TRY(cph_comm_exchange(h, np, peers, sb, sn, rb, rn));
  return 0;
}

static DirTable send_table(const cph_handle *h) {
  DirTable t;
  for (int d = 0; d < 27; d++) { t.start[d] = h->rec_start[d]; t.out[d] = h->send_off[d]; }
  t.start[27] = h->rec_start[27];
  return t;
}

// comm->forward_comm(), first half: ship the copies that belong to other ranks.  Peer mode: one
// kernel packs AND stores them into the neighbours' receive buffers over NVLink; the caller's
// next collective (the decision-flag all-reduce) is the barrier that makes them visible.
// NCCL mode: pack + one ncclSend/ncclRecv per neighbour rank.
int cph_halo_send(cph_handle *h) {
  if (h->nranks == 1) return 0;
  ProfScope ps(h, 5);
  GhostDirs gd;
  make_ghost_dirs(h, gd);
  cudaStream_t st = h->stream;
  h->halo_parity ^= 1;     // every rank toggles in lockstep: writers never touch the half being read
  if (h->peer_halo) {
    MailFlags mf;
    mf.P = 0;
    if (h->mail_ok) {      // this step's decision flags travel with the positions
      mf.P = h->nranks;
      mf.seq = ++h->seq_flags;
      const int par = (int)(mf.seq & 1);
      for (int p = 0; p < h->nranks; p++)
        mf.dst[p] = (unsigned int *)((unsigned char *)h->mail_base[p] + mail_flag_off(h->nranks, par, h->rank));
    }
    if (h->nsend || mf.P) {
      PeerTab pt;
      for (int d = 0; d < 27; d++) {
        pt.dst[d] = nullptr;
        if (gd.active[d] != 2) continue;
        const int p = gd.peer[d];
        const int *row = h->peer_table.data() + (size_t)p * 28;
        pt.dst[d] = (double4 *)h->peer_base[p] + (size_t)h->halo_parity * (size_t)row[27] + row[d];
      }
      h->nlaunch++;
      halo_pack_peer_kernel<<<std::max(1, nblk(h->nrec)), TPB, 0, st>>>(h->nrec, h->d_rec_src.p, h->d_rec_dir.p,
                                                                       send_table(h), gd, h->d_xq.p, pt, mf, h->d_flags.p,
                                                                       h->d_flags.p + 82);
    }
  } else {
    if (h->nsend) {
      h->nlaunch++;
      halo_pack_kernel<<<nblk(h->nrec), TPB, 0, st>>>(h->nrec, h->d_rec_src.p, h->d_rec_dir.p, send_table(h), gd,
                                                     h->d_xq.p, nullptr, nullptr, nullptr, nullptr, h->d_sendx.p,
                                                     nullptr);
    }
    CPH_TRY(halo_exchange(h, gd, false));
  }
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

// second half: self images and received copies -> ghost atoms
int cph_halo_finish(cph_handle *h) {
  if (h->nghost == 0) return 0;
  ProfScope ps(h, 5);
  GhostDirs gd;
  make_ghost_dirs(h, gd);
  h->nlaunch++;
  ghost_copy_kernel<<<nblk(h->nghost), TPB, 0, h->stream>>>(h->nghost, h->nlocal, h->d_ghost_src.p, h->d_ghost_code.p, gd,
                                                           h->d_recvx.p + (size_t)h->halo_parity * h->recv_half, nullptr,
                                                           h->d_xq.p, nullptr, nullptr, nullptr, nullptr, 0);
  CPH_CUDA(h, cudaGetLastError());
  return 0;
}

// max over ranks of the six decision words of this step -> dst[0..5].  With mailboxes: the flags were published by
// cph_halo_send's pack kernel; one warp waits for all of them and reduces (no NCCL, no host).  Otherwise ncclAllReduce
// in place on d_flags (dst is then d_flags itself).  Either way it is also the barrier behind which a rank may read
// the copies its neighbours stored into its receive buffer.
int cph_flags_allreduce(cph_handle *h, unsigned int *dst) {
  if (h->nranks == 1) return 0;
  if (h->peer_halo && h->mail_ok) {
    ProfScope ps(h, 5);
    const int par = (int)(h->seq_flags & 1);
    h->nlaunch++;
    flags_gather_kernel<<<1, 32, 0, h->stream>>>(
        (const unsigned int *)((unsigned char *)h->d_mail.p + mail_flag_off(h->nranks, par, 0)), h->nranks, h->seq_flags, dst,
        h->d_flags.p + 83);
    CPH_CUDA(h, cudaGetLastError());
    return 0;
  }
  if (dst != h->d_flags.p) CPH_CUDA(h, cudaMemcpyAsync(dst, h->d_flags.p, 6 * sizeof(unsigned int), cudaMemcpyDeviceToDevice, h->stream));
  return cph_comm_allreduce_max_u32_dev(h, dst, 6);
}

int cph_forward_ghosts(cph_handle *h) {
  CPH_TRY(cph_halo_send(h));
  // stand-alone call: a (tiny) collective orders the peer stores before the reads; its result is not used
  if (h->nranks > 1 && h->peer_halo) CPH_TRY(cph_flags_allreduce(h, h->d_flags.p + 84));
  return cph_halo_finish(h);
}

void cph_halo_close(cph_handle *h) {
  for (size_t p = 0; p < h->peer_base.size(); p++)
    if (h->peer_base[p]) cudaIpcCloseMemHandle(h->peer_base[p]);
  h->peer_base.clear();
  h->peer_handle_cache.clear();
  h->peer_halo = false;
}

void cph_mail_close(cph_handle *h) {
  for (size_t p = 0; p < h->mail_base.size(); p++)
    if (h->mail_base[p] && (int)p != h->rank) cudaIpcCloseMemHandle(h->mail_base[p]);
  h->mail_base.clear();
  h->mail_handle_cache.clear();
  h->mail_ok = false;
}

namespace {
// All-gather of one 32-int block per rank through the mailboxes: post mine into my slot on every rank, wait for
// everybody's sequence number in MY mailbox, copy the P blocks out.  One block of 32 * P threads.
__global__ void mail_gather_kernel(const int *mine, int P, int me, unsigned long long seq, int *const *dst_slots,
                                   const int *my_slots, int *out, unsigned int *status) {
  const int p = threadIdx.x >> 5, k = threadIdx.x & 31;
  if (p < P) dst_slots[p][k] = mine[k];                       // my block into rank p's mailbox
  __threadfence_system();
  __syncthreads();
  if (k == 0 && p < P) *reinterpret_cast<volatile unsigned long long *>(dst_slots[p] + 32) = seq;
  if (p < P) {
    const int *slot = my_slots + (size_t)p * CPH_MAIL_GSLOT;  // rank p's block in MY mailbox
    if (k == 0) {
      const volatile unsigned long long *sq = reinterpret_cast<const volatile unsigned long long *>(slot + 32);
      const long long t0 = clock64();
      while (*sq != seq) {
        if (clock64() - t0 > 400000000000LL) { atomicOr(status, 1u); break; }
        __nanosleep(200);
      }
      __threadfence_system();
    }
    __syncwarp();
    out[p * 32 + k] = reinterpret_cast<const volatile int *>(slot)[k];
  }
}
}  // namespace

// host blocks in (32 ints) and out (32 ints per rank); one kernel, one synchronisation
static int mail_gather32(cph_handle *h, const int *mine32, int *all) {
  const int P = h->nranks;
  const unsigned long long seq = ++h->seq_gather;
  const int par = (int)(seq & 1);
  CPH_CUDA(h, h->d_ipc_stage.reserve(4096));
  int *d_mine = (int *)h->d_ipc_stage.p;                      // [32] mine, [64..64+8 ptrs] slots, [256..] out
  int **d_dst = (int **)(h->d_ipc_stage.p + 256);
  int *d_out = (int *)(h->d_ipc_stage.p + 512);
  int *dst_h[CPH_MAIL_MAXP];
  for (int p = 0; p < P; p++)
    dst_h[p] = (int *)((unsigned char *)h->mail_base[p] + mail_gather_off(P, h->mail_red_cap, par, h->rank));
  CPH_CUDA(h, cudaMemcpyAsync(d_mine, mine32, 32 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CPH_CUDA(h, cudaMemcpyAsync(d_dst, dst_h, P * sizeof(int *), cudaMemcpyHostToDevice, h->stream));
  h->nlaunch++;
  mail_gather_kernel<<<1, 32 * P, 0, h->stream>>>(d_mine, P, h->rank, seq, d_dst,
      (const int *)((unsigned char *)h->d_mail.p + mail_gather_off(P, h->mail_red_cap, par, 0)), d_out, h->d_flags.p + 83);
  CPH_CUDA(h, cudaGetLastError());
  CPH_CUDA(h, cudaMemcpyAsync(all, d_out, (size_t)P * 32 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// receive layout of the rank at grid position loc, from the matrix of every rank's send counts per direction:
// the same rule cph_rebuild applies to itself (blocks grouped by source rank, directions in order inside a block)
static void recv_layout_of(const cph_handle *h, const int *loc, const int *matrix32, int *recv_count, int *recv_off,
                           int *nrecv_out) {
  GhostDirs gd;
  make_ghost_dirs_at(h, loc, gd);
  int nrecv = 0;
  for (int d = 0; d < 27; d++) {
    recv_count[d] = gd.from[d] >= 0 ? matrix32[gd.from[d] * 32 + d] : 0;
    recv_off[d] = 0;
  }
  for (int p = 0; p < h->nranks; p++)
    for (int d = 0; d < 27; d++)
      if (gd.from[d] == p) { recv_off[d] = nrecv; nrecv += recv_count[d]; }
  *nrecv_out = nrecv;
}

// Map the neighbours' receive buffers (CUDA IPC).  Called at every list build: handles and
// layouts are all-gathered, mappings are reopened only when a neighbour's buffer moved.
static int halo_map_peers(cph_handle *h, const GhostDirs &gd) {
  const int P = h->nranks;
  h->peer_halo = false;
  h->mail_ok = false;
  if (P == 1 || !h->peer_halo_wanted) return 0;
  // this rank's mailbox, allocated once (its address is exported): room for twice the current site table
  if (!h->d_mail.p && P <= CPH_MAIL_MAXP && h->mail_wanted) {
    h->mail_red_cap = std::max((size_t)32768, (size_t)2 * (4 + 2 * (size_t)h->S + 1));
    const size_t bytes = mail_bytes(P, h->mail_red_cap);
    if (h->d_mail.reserve_exact(bytes) == cudaSuccess) cudaMemsetAsync(h->d_mail.p, 0, bytes, h->stream);
    else { cudaGetLastError(); h->mail_red_cap = 0; }
  }
  // record: IPC handle of the receive buffer + recv_off[27] + recv_half + IPC handle of the mailbox + its slot size
  const size_t rec = 64 + 28 * sizeof(int) + 64 + sizeof(unsigned long long);
  CPH_CUDA(h, h->d_ipc_stage.reserve(rec * (P + 1)));
  std::vector<unsigned char> mine(rec, 0), all(rec * P, 0);
  cudaIpcMemHandle_t hd;
  unsigned int ok = cudaIpcGetMemHandle(&hd, h->d_recvx.p) == cudaSuccess ? 1u : 0u;
  if (!ok) cudaGetLastError();
  memcpy(mine.data(), &hd, 64);
  int tab[28];
  for (int d = 0; d < 27; d++) tab[d] = h->recv_off[d];
  tab[27] = (int)h->recv_half;
  memcpy(mine.data() + 64, tab, sizeof(tab));
  unsigned int mok = 0;
  if (h->d_mail.p) {
    cudaIpcMemHandle_t mh;
    mok = cudaIpcGetMemHandle(&mh, h->d_mail.p) == cudaSuccess ? 1u : 0u;
    if (!mok) cudaGetLastError();
    memcpy(mine.data() + 64 + sizeof(tab), &mh, 64);
    const unsigned long long cap = h->mail_red_cap;
    memcpy(mine.data() + 64 + sizeof(tab) + 64, &cap, sizeof(cap));
  }
  CPH_CUDA(h, cudaMemcpyAsync(h->d_ipc_stage.p, mine.data(), rec, cudaMemcpyHostToDevice, h->stream));
  CPH_TRY(cph_comm_allgather(h, h->d_ipc_stage.p, h->d_ipc_stage.p + rec, rec));
  unsigned int status = 0;   // a waiting kernel gave up on a rank earlier: say so now (the host synchronises here anyway)
  CPH_CUDA(h, cudaMemcpyAsync(all.data(), h->d_ipc_stage.p + rec, rec * P, cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaMemcpyAsync(&status, h->d_flags.p + 83, sizeof(status), cudaMemcpyDeviceToHost, h->stream));
  CPH_CUDA(h, cudaStreamSynchronize(h->stream));
  if (status) return cph_fail(h, CPH_ERR_COMM, "a rank did not publish its block of a mailbox reduction in time");
  h->peer_base.resize(P, nullptr);
  h->peer_handle_cache.resize((size_t)P * 64, 0);
  h->peer_table.assign((size_t)P * 28, 0);
  for (int p = 0; p < P; p++) memcpy(h->peer_table.data() + (size_t)p * 28, all.data() + rec * p + 64, 28 * sizeof(int));
  for (int d = 0; d < 27 && ok; d++) {
    if (gd.active[d] != 2) continue;
    const int p = gd.peer[d];
    if (p == h->rank) continue;
    const unsigned char *hp = all.data() + rec * p;
    if (h->peer_base[p] && memcmp(hp, h->peer_handle_cache.data() + (size_t)p * 64, 64) == 0) continue;
    if (h->peer_base[p]) { cudaIpcCloseMemHandle(h->peer_base[p]); h->peer_base[p] = nullptr; }
    cudaIpcMemHandle_t ph;
    memcpy(&ph, hp, 64);
    void *ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
    h->peer_base[p] = ptr;
    memcpy(h->peer_handle_cache.data() + (size_t)p * 64, hp, 64);
  }
  // mailboxes: EVERY rank maps EVERY other rank's (the reductions are all-to-all), same slot size everywhere
  h->mail_base.resize(P, nullptr);
  h->mail_handle_cache.resize((size_t)P * 64, 0);
  for (int p = 0; p < P && mok; p++) {
    const unsigned char *hp = all.data() + rec * p + 64 + sizeof(tab);
    unsigned long long cap = 0;
    memcpy(&cap, hp + 64, sizeof(cap));
    if (cap != h->mail_red_cap) { mok = 0; break; }
    if (p == h->rank) { h->mail_base[p] = h->d_mail.p; continue; }
    if (h->mail_base[p] && memcmp(hp, h->mail_handle_cache.data() + (size_t)p * 64, 64) == 0) continue;
    if (h->mail_base[p]) { cudaIpcCloseMemHandle(h->mail_base[p]); h->mail_base[p] = nullptr; }
    cudaIpcMemHandle_t ph;
    memcpy(&ph, hp, 64);
    void *ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mok = 0; break; }
    h->mail_base[p] = ptr;
    memcpy(h->mail_handle_cache.data() + (size_t)p * 64, hp, 64);
  }
  // every rank must agree: one failure anywhere falls back to NCCL everywhere
  unsigned int bad[2] = {ok ? 0u : 1u, (ok && mok) ? 0u : 1u};
  CPH_TRY(cph_comm_allreduce_max_u32(h, bad, 2));
  h->peer_halo = bad[0] == 0;
  h->mail_ok = bad[1] == 0;
  return 0;
}

// cudaMalloc/cudaFree cost hundreds of milliseconds on this platform, so every per-atom buffer
// is sized once, to one common capacity with head room, and a steady-state rebuild allocates
// nothing.  `need` counts owned + ghost atoms + the dummy.
static int ensure_atom_capacity(cph_handle *h, size_t need) {
  cudaStream_t st = h->stream;
  if (need <= h->atom_cap) return 0;
  const size_t cap = need + need / 8 + 1024;   // head room once; all buffers get exactly this
  CPH_CUDA(h, h->d_xq.reserve_exact(cap, true, st));
  CPH_CUDA(h, h->d_type.reserve_exact(cap, true, st));
  CPH_CUDA(h, h->d_tag.reserve_exact(cap, true, st));
  CPH_CUDA(h, h->d_mask.reserve_exact(cap, true, st));
  CPH_CUDA(h, h->d_perm.reserve_exact(cap, true, st));
  CPH_CUDA(h, h->d_scr_off.reserve_exact(cap, true, st));   // holds the ghost offsets across the regrow
  CPH_CUDA(h, h->d_rec_src.reserve_exact(cap, true, st));   // halo records live from build to build
  CPH_CUDA(h, h->d_rec_dir.reserve_exact(cap, true, st));
  CPH_CUDA(h, h->d_mol.reserve_exact(cap, true, st));
  CPH_CUDA(h, h->d_xq2.reserve_exact(cap));
  CPH_CUDA(h, h->d_xt.reserve_exact(cap));
  CPH_CUDA(h, h->d_xb.reserve_exact(cap));
  DevBuf<int> *ib[] = {&h->d_scr_i, &h->d_vals, &h->d_vals2, &h->d_tmpi, &h->d_scr_src, &h->d_scr_code,
                       &h->d_ghost_src, &h->d_ghost_code};
  for (auto *b : ib) CPH_CUDA(h, b->reserve_exact(cap));
  CPH_CUDA(h, h->d_keys.reserve_exact(cap));
  CPH_CUDA(h, h->d_keys2.reserve_exact(cap));
  // the swap-based permutation needs identical capacities
  h->d_xq.cap = h->d_xq2.cap = cap;
  h->d_type.cap = h->d_tag.cap = h->d_mask.cap = h->d_perm.cap = h->d_scr_i.cap = cap;
  h->atom_cap = cap;
  return 0;
}

// CPH_TRACE=1 prints host wall-clock per rebuild phase (debug aid; adds stream syncs)
struct PhaseTrace {
  cph_handle *h;
  bool on;
  std::chrono::steady_clock::time_point t0;
  explicit PhaseTrace(cph_handle *h_) : h(h_), on(getenv("CPH_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void mark(const char *what) {
    if (!on) return;
    cudaStreamSynchronize(h->stream);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[cph rebuild %lld] %-12s %8.3f ms\n", (long long)h->nbuilds, what,
            std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

int cph_rebuild(cph_handle *h) {
  ProfScope ps(h, 6);
  PhaseTrace tr(h);
  const int n = h->nlocal;
  cudaStream_t st = h->stream;
  const double cutmax = std::max(h->cut_lj_max, h->cut_coul);
  const double rlist = cutmax + h->skin;

  // ---- grid over the extended sub-box (fixed per domain) -------------------------------
  Grid &g = h->grid;
  const double margin = rlist + h->skin;
  g.ncell = 1;
  for (int k = 0; k < 3; k++) {
    double lo = h->sublo[k] - margin, hi = h->subhi[k] + margin;
    int nc = std::max(1, (int)std::floor((hi - lo) / (0.5 * rlist)));
    g.lo[k] = lo;
    g.inv[k] = nc / (hi - lo);
    g.n[k] = nc;
    g.ncell *= nc;
  }
  CPH_CUDA(h, h->d_flags.reserve(8));
  CPH_CUDA(h, cudaMemsetAsync(h->d_flags.p, 0, 8 * sizeof(unsigned int), st));
  CPH_TRY(cph_md_wrap(h));   // device-side dynamics: remap into the periodic box, as LAMMPS does when re-neighbouring

  // ---- drift of owned atoms outside the sub-box -> ghost cutoff ---------------------------
  double3 slo = make_double3(h->sublo[0], h->sublo[1], h->sublo[2]);
  double3 shi = make_double3(h->subhi[0], h->subhi[1], h->subhi[2]);
  unsigned int flags_h[8];
  float drift;
  if (h->drift_known && !h->md_on) {
    // measured by this step's set_x / check kernel and all-reduced with the decision flags
    drift = h->drift_value;
  } else {
    if (n) drift_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_xq.p, slo, shi, h->d_flags.p);
    CPH_CUDA(h, cudaMemcpyAsync(flags_h, h->d_flags.p, sizeof(flags_h), cudaMemcpyDeviceToHost, st));
    CPH_CUDA(h, cudaStreamSynchronize(st));
    CPH_TRY(cph_comm_allreduce_max_u32(h, flags_h + 2, 1));
    memcpy(&drift, &flags_h[2], 4);
  }
  h->drift_known = false;
  if (drift > h->skin)
    return cph_fail(h, CPH_ERR_DOMAIN, "owned atoms drifted %.3f beyond the sub-box (> skin %.3f): "
                    "the host must migrate atoms and call cph_set_atoms", drift, h->skin);
  h->ghost_cut = rlist + drift;
  for (int k = 0; k < 3; k++) {
    bool needs = h->periodic[k] || h->procgrid[k] > 1;
    if (needs && h->subhi[k] - h->sublo[k] < h->ghost_cut)
      return cph_fail(h, CPH_ERR_DOMAIN, "sub-box extent %.3f in dim %d is smaller than the ghost cutoff %.3f",
                      h->subhi[k] - h->sublo[k], k, h->ghost_cut);
  }

  tr.mark("drift");
  {
    // expected owned + ghost count from the shell volume, with head room
    double fac = 1.0;
    for (int k = 0; k < 3; k++)
      if (h->periodic[k] || h->procgrid[k] > 1) fac *= 1.0 + 2.0 * (rlist + h->skin) / (h->subhi[k] - h->sublo[k]);
    CPH_TRY(ensure_atom_capacity(h, (size_t)(n * fac * 1.05) + 4096));
  }
  // ---- sort owned atoms by (cell, tag) -------------------------------------------------------
  CPH_CUDA(h, h->d_keys.reserve(n + 1));
  CPH_CUDA(h, h->d_keys2.reserve(n + 1));
  CPH_CUDA(h, h->d_vals.reserve(n + 1));
  CPH_CUDA(h, h->d_vals2.reserve(n + 1));
  CPH_CUDA(h, h->d_tmpi.reserve(n + 1));
  if (n) {
    key_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_xq.p, h->d_tag.p, g, h->d_keys.p, h->d_vals.p);
    CPH_TRY(sort_pairs(h, n, h->d_keys, h->d_keys2, h->d_vals, h->d_vals2, 64));
    // d_vals2 = old index of the atom now at position k
    CPH_TRY(permute_buf(h, n, h->d_vals2.p, h->d_xq, h->d_xq2, h->atom_cap));
    if (h->md_on) CPH_TRY(permute_buf(h, n, h->d_vals2.p, h->d_v, h->d_v2, h->d_v.cap));
    // d_scr_i is a persistent scratch buffer; after each swap it holds the previous array
    DevBuf<int> &tmp = h->d_scr_i;
    const size_t icap = h->d_type.cap;   // == tag/mask/perm/scr_i capacity (ensure_atom_capacity)
    CPH_TRY(permute_buf(h, n, h->d_vals2.p, h->d_type, tmp, icap));
    CPH_TRY(permute_buf(h, n, h->d_vals2.p, h->d_tag, tmp, icap));
    CPH_TRY(permute_buf(h, n, h->d_vals2.p, h->d_mask, tmp, icap));
    CPH_TRY(permute_buf(h, n, h->d_vals2.p, h->d_perm, tmp, icap));
    CPH_CUDA(h, h->d_inv.reserve(n));
    CPH_CUDA(h, h->d_xbuild.reserve(3 * (size_t)n));
    after_sort_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_keys2.p, h->d_perm.p, h->d_xq.p, h->d_inv.p, h->d_xbuild.p,
                                               h->d_tmpi.p);
  }
  CPH_CUDA(h, h->d_cell_start_o.reserve(g.ncell + 1));
  CPH_CUDA(h, h->d_cell_start_g.reserve(g.ncell + 1));
  cell_start_kernel<<<nblk(n + 1), TPB, 0, st>>>(n, g.ncell, h->d_tmpi.p, h->d_cell_start_o.p);

  tr.mark("sort");
  // ---- ghosts: periodic self images + copies from the spatial neighbours (K6) ---------------------
  GhostDirs gd;
  make_ghost_dirs(h, gd);
  int nrec = 0;
  if (h->have_mol && n) mol_owned_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_perm.p, h->d_molecule.p, h->d_mol.p);
  if (n) {
    ghost_count_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_xq.p, slo, shi, h->ghost_cut, gd, h->d_vals.p);
    CPH_TRY(exclusive_scan(h, n + 1, h->d_vals.p, h->d_vals2.p));
    CPH_CUDA(h, cudaMemcpyAsync(&nrec, h->d_vals2.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    CPH_CUDA(h, cudaStreamSynchronize(st));
  }
  if (n) CPH_CUDA(h, cudaMemcpyAsync(h->d_scr_off.p, h->d_vals2.p, (n + 1) * sizeof(int), cudaMemcpyDeviceToDevice, st));
  CPH_TRY(ensure_atom_capacity(h, (size_t)n + nrec + 2));
  h->nrec = nrec;
  for (int d = 0; d < 28; d++) h->rec_start[d] = 0;
  if (nrec) {
    // records (owner, direction), sorted by direction (stable: owner order kept)
    ghost_fill_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_xq.p, slo, shi, h->ghost_cut, gd, h->d_scr_off.p,
                                               h->d_scr_src.p, h->d_keys.p, h->d_vals.p);
    CPH_TRY(sort_pairs(h, nrec, h->d_keys, h->d_keys2, h->d_vals, h->d_vals2, 5));
    gather_kernel<int><<<nblk(nrec), TPB, 0, st>>>(nrec, h->d_vals2.p, h->d_scr_src.p, h->d_rec_src.p);
    dir_of_key_kernel<<<nblk(nrec), TPB, 0, st>>>(nrec, h->d_keys2.p, h->d_rec_dir.p);
    cell_start_kernel<<<nblk(nrec + 1), TPB, 0, st>>>(nrec, 27, h->d_rec_dir.p, h->d_scr_code.p);
    CPH_CUDA(h, cudaMemcpyAsync(h->rec_start, h->d_scr_code.p, 28 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CPH_CUDA(h, cudaStreamSynchronize(st));
  }
  // offsets: self images -> ghost slots [0,nloc); remote copies -> send buffer [0,nsend), grouped
  // by destination rank so that each neighbour rank gets ONE contiguous message
  int nloc = 0, nsend = 0;
  DirTable ltab;
  for (int d = 0; d < 27; d++) {
    const int c = h->rec_start[d + 1] - h->rec_start[d];
    ltab.start[d] = h->rec_start[d];
    ltab.out[d] = -1;
    h->send_count[d] = gd.active[d] == 2 ? c : 0;
    h->send_off[d] = 0;
    if (gd.active[d] == 1) { ltab.out[d] = nloc; nloc += c; }
  }
  ltab.start[27] = h->rec_start[27];
  int nrecv = 0;
  for (int d = 0; d < 27; d++) { h->recv_count[d] = 0; h->recv_off[d] = 0; }
  h->peer_rank.clear(); h->peer_soff.clear(); h->peer_scnt.clear(); h->peer_roff.clear(); h->peer_rcnt.clear();
  if (h->nranks > 1) {
    // Counts first (one int per direction), then the records.  Steady state (mailboxes mapped by an earlier build):
    // ONE all-gather of every rank's 27 send counts through the mailboxes; each rank then derives its own receive
    // counts, everybody's receive layout and the "does anyone have to grow its receive buffer" decision locally,
    // so the five small NCCL calls + host synchronisations of the first build are not repeated.
    const int P = h->nranks;
    const bool fast = h->mail_ok && h->peer_halo && (int)h->peer_table.size() == P * 28;
    std::vector<int> matrix;
    auto loc_of = [&](int r, int *loc) {
      loc[0] = r % h->procgrid[0]; loc[1] = (r / h->procgrid[0]) % h->procgrid[1]; loc[2] = r / (h->procgrid[0] * h->procgrid[1]);
    };
    if (fast) {
      int mine[32] = {0};
      for (int d = 0; d < 27; d++) mine[d] = h->send_count[d];
      matrix.resize((size_t)P * 32);
      CPH_TRY(mail_gather32(h, mine, matrix.data()));
      int off_unused[27], n_unused;
      recv_layout_of(h, h->myloc, matrix.data(), h->recv_count, off_unused, &n_unused);
    } else {
      CPH_TRY(cph_comm_exchange_counts(h, gd.active, gd.peer, gd.from, h->send_count, h->recv_count));
    }
    for (int p = 0; p < h->nranks; p++) {
      int so = nsend, sc = 0, ro = nrecv, rc = 0;
      for (int d = 0; d < 27; d++) {
        if (gd.active[d] == 2 && gd.peer[d] == p) { h->send_off[d] = nsend; nsend += h->send_count[d]; sc += h->send_count[d]; }
        if (gd.from[d] == p) { h->recv_off[d] = nrecv; nrecv += h->recv_count[d]; rc += h->recv_count[d]; }
      }
      if (sc || rc) {
        h->peer_rank.push_back(p); h->peer_soff.push_back(so); h->peer_scnt.push_back(sc);
        h->peer_roff.push_back(ro); h->peer_rcnt.push_back(rc);
      }
    }
    h->nsend = nsend;
    CPH_CUDA(h, h->d_sendx.reserve(nsend + 1));
    CPH_CUDA(h, h->d_sendmeta.reserve(nsend + 1));
    {
      // two halves (alternating per step) so that a neighbour's next write never lands in the
      // half this rank is still reading.  The buffer is exported through CUDA IPC, so it may only
      // be reallocated after EVERY rank has closed its mappings: agree on "someone must grow",
      // close, synchronise, then grow (with head room, so this is rare).
      unsigned int grow = ((size_t)nrecv + 1 > h->recv_half) ? 1u : 0u;
      unsigned int any_grow = grow;
      if (fast) {   // every rank's need, from the matrix and the buffer sizes exchanged at the last mapping
        for (int r = 0; r < P; r++) {
          int loc[3], cnt[27], off[27], nr;
          loc_of(r, loc);
          recv_layout_of(h, loc, matrix.data(), cnt, off, &nr);
          if ((size_t)nr + 1 > (size_t)h->peer_table[(size_t)r * 28 + 27]) any_grow = 1;
        }
      } else {
        CPH_TRY(cph_comm_allreduce_max_u32(h, &any_grow, 1));
      }
      if (any_grow) {
        cph_halo_close(h);
        unsigned int closed = 1;
        CPH_TRY(cph_comm_allreduce_max_u32(h, &closed, 1));   // barrier: nobody maps anybody any more
        if (grow) {
          const size_t half = (size_t)nrecv + 1;
          h->recv_half = half + half / 4 + 256;
          h->d_recvx.release();
          CPH_CUDA(h, h->d_recvx.reserve_exact(2 * h->recv_half));
        }
      }
    }
    CPH_CUDA(h, h->d_recvmeta.reserve(nrecv + 1));
    if (nsend)
      halo_pack_kernel<<<nblk(nrec), TPB, 0, st>>>(nrec, h->d_rec_src.p, h->d_rec_dir.p, send_table(h), gd, h->d_xq.p,
                                                  h->d_type.p, h->d_tag.p, h->d_mask.p,
                                                  h->have_mol ? h->d_mol.p : nullptr, h->d_sendx.p, h->d_sendmeta.p);
    CPH_TRY(halo_exchange(h, gd, true));
    if (fast && h->peer_halo) {
      // nobody grew: buffers, mappings and sizes are those of the last mapping; the receive offsets of every rank
      // follow from the matrix
      for (int r = 0; r < P; r++) {
        int loc[3], cnt[27], off[27], nr;
        loc_of(r, loc);
        recv_layout_of(h, loc, matrix.data(), cnt, off, &nr);
        for (int d = 0; d < 27; d++) h->peer_table[(size_t)r * 28 + d] = off[d];
      }
    } else {
      CPH_TRY(halo_map_peers(h, gd));
    }
  }
  h->nrecv = nrecv;
  const int nghost = nloc + nrecv;
  h->nghost = nghost;
  h->nall = n + nghost;
  CPH_TRY(ensure_atom_capacity(h, (size_t)h->nall + 2));   // + the far-away dummy atom that pads neighbour rows
  if (nghost) {
    const int nthreads = std::max(nrec, nrecv);
    ghost_candidates_kernel<<<nblk(nthreads), TPB, 0, st>>>(nrec, h->d_rec_src.p, h->d_rec_dir.p, ltab, gd, nloc, nrecv,
                                                           h->d_xq.p, h->d_tag.p,
                                                           h->d_recvx.p + (size_t)h->halo_parity * h->recv_half,
                                                           h->d_recvmeta.p, g,
                                                           h->d_scr_src.p, h->d_scr_code.p, h->d_keys.p, h->d_vals.p);
    CPH_TRY(sort_pairs(h, nghost, h->d_keys, h->d_keys2, h->d_vals, h->d_vals2, 64));
    gather_kernel<int><<<nblk(nghost), TPB, 0, st>>>(nghost, h->d_vals2.p, h->d_scr_src.p, h->d_ghost_src.p);
    gather_kernel<int><<<nblk(nghost), TPB, 0, st>>>(nghost, h->d_vals2.p, h->d_scr_code.p, h->d_ghost_code.p);
    ghost_copy_kernel<<<nblk(nghost), TPB, 0, st>>>(nghost, n, h->d_ghost_src.p, h->d_ghost_code.p, gd,
                                                    h->d_recvx.p + (size_t)h->halo_parity * h->recv_half,
                                                    h->d_recvmeta.p, h->d_xq.p, h->d_type.p, h->d_tag.p, h->d_mask.p,
                                                    h->have_mol ? h->d_mol.p : nullptr, 1);
    after_sort_kernel<<<nblk(nghost), TPB, 0, st>>>(nghost, h->d_keys2.p, nullptr, nullptr, nullptr, nullptr,
                                                    h->d_tmpi.p);
    cell_start_kernel<<<nblk(nghost + 1), TPB, 0, st>>>(nghost, g.ncell, h->d_tmpi.p, h->d_cell_start_g.p);
  } else {
    fill_int_kernel<<<nblk(g.ncell + 1), TPB, 0, st>>>(g.ncell + 1, h->d_cell_start_g.p, 0);
  }

  {
    // dummy atom at index nall: far outside any cutoff, zero charge, a valid type
    const double far = 1.0e15;
    double4 dq = make_double4(far, far, far, 0.0);
    int one = 1, zero = 0;
    CPH_CUDA(h, cudaMemcpyAsync(h->d_xq.p + h->nall, &dq, sizeof(dq), cudaMemcpyHostToDevice, st));
    CPH_CUDA(h, cudaMemcpyAsync(h->d_type.p + h->nall, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    CPH_CUDA(h, cudaMemcpyAsync(h->d_tag.p + h->nall, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
    CPH_CUDA(h, cudaMemcpyAsync(h->d_mask.p + h->nall, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
    CPH_CUDA(h, cudaStreamSynchronize(st));
  }

  tr.mark("ghosts");
  // ---- Verlet list ----------------------------------------------------------------------------------
  if (h->rowcap == 0) {
    // expected row length from the mean density, with head room; regrown on overflow
    double vol = 1.0;
    for (int k = 0; k < 3; k++) vol *= (h->subhi[k] - h->sublo[k]);
    double rho = n / std::max(vol, 1e-30);
    double expect = 4.0 / 3.0 * M_PI * rlist * rlist * rlist * rho;
    h->rowcap = ((int)(expect * 1.25) + 96 + 127) / 128 * 128;
  }
  CPH_CUDA(h, h->d_numneigh.reserve(n + 1));
  CPH_CUDA(h, h->d_numspec.reserve(n + 1));
  DevBuf<unsigned long long> &stats = h->d_scr_stats;
  CPH_CUDA(h, stats.reserve(2));
  // build-time fp32 records {x, y, z, molecule id} of all atoms (owned + ghost + dummy)
  xb_kernel<<<nblk(h->nall + 1), TPB, 0, st>>>(h->nall, h->d_xq.p, h->have_mol ? h->d_mol.p : nullptr,
                                              make_double3(g.lo[0], g.lo[1], g.lo[2]), h->d_xb.p);
  int dropmask = 0;   // special class c is not stored when both weights are zero, except under coul/dsf
  if ((h->pp.style != CPH_PAIR_LJ_CUT_COUL_DSF || h->drop_excluded) && !h->have_topology)   // bonded partners are looked up among the specials
    for (int c = 1; c <= 3; c++)
      if (h->pp.special_lj[c] == 0.0 && h->pp.special_coul[c] == 0.0) dropmask |= 1 << c;
  h->last_dropmask = dropmask;
  double extent = 0;
  for (int k = 0; k < 3; k++) extent = std::max(extent, g.n[k] / g.inv[k]);
  const float fmargin = (float)(32.0 * rlist * extent * 5.97e-8 + 1e-5 * rlist * rlist);
  // grid layers (from each face of the grid) that can hold ghost atoms: a ghost is a copy of an atom
  // some rank owns, so it sits outside this sub-box or at most `skin` inside it (drifted owners)
  int3 ghost_layers;
  {
    int gl[3];
    for (int k = 0; k < 3; k++) {
      const double w = 1.0 / g.inv[k];
      gl[k] = (int)std::floor((h->sublo[k] + h->skin - g.lo[k]) / w) + 1;
    }
    ghost_layers = make_int3(gl[0], gl[1], gl[2]);
  }
  for (int attempt = 0; attempt < 4 && n; attempt++) {
    CPH_CUDA(h, h->d_neigh.reserve((size_t)n * h->rowcap));
    CPH_CUDA(h, cudaMemsetAsync(stats.p, 0, 2 * sizeof(unsigned long long), st));
    CPH_CUDA(h, cudaMemsetAsync(h->d_flags.p + 1, 0, sizeof(unsigned int), st));
    CPH_CUDA(h, cudaMemsetAsync(h->d_flags.p + 3, 0, sizeof(unsigned int), st));
    int warps_per_block = TPB / 32;
    int blocks = (n + warps_per_block - 1) / warps_per_block;
    list_build_kernel<<<blocks, TPB, 0, st>>>(
        n, h->d_xq.p, h->d_xb.p, h->d_tag.p, h->d_perm.p, h->maxspecial ? h->d_nspecial.p : nullptr,
        h->maxspecial ? h->d_special.p : nullptr, h->maxspecial, g, h->d_cell_start_o.p, h->d_cell_start_g.p,
        ghost_layers, rlist * rlist, fmargin, dropmask, h->rowcap, h->nall, h->d_neigh.p,
        h->d_numneigh.p, h->d_numspec.p, h->d_flags.p, stats.p);
    CPH_CUDA(h, cudaGetLastError());
    unsigned long long stats_h[2];
    CPH_CUDA(h, cudaMemcpyAsync(flags_h, h->d_flags.p, sizeof(flags_h), cudaMemcpyDeviceToHost, st));
    CPH_CUDA(h, cudaMemcpyAsync(stats_h, stats.p, sizeof(stats_h), cudaMemcpyDeviceToHost, st));
    CPH_CUDA(h, cudaStreamSynchronize(st));
    if (flags_h[1] > (unsigned int)h->rowcap) {
      h->rowcap = ((int)flags_h[1] + 64 + 127) / 128 * 128;  // regrow and redo
      if (attempt == 3) return cph_fail(h, CPH_ERR_OVERFLOW, "neighbour rows overflowed after regrow");
      continue;
    }
    if (flags_h[7] > 32u)
      return cph_fail(h, CPH_ERR_OVERFLOW, "an atom has %u special-bond partners inside the list cutoff; at most 32 are supported",
                      flags_h[7]);
    h->stored_neigh = (int64_t)stats_h[0];
    h->special_pairs = (int64_t)stats_h[1];
    h->maxneigh = (int)flags_h[3];
    break;
  }

  tr.mark("list");
  // ---- site bookkeeping -----------------------------------------------------------------------------
  CPH_CUDA(h, h->d_site_of.reserve(n + 1));
  CPH_CUDA(h, h->d_titr_of.reserve(n + 1));
  CPH_CUDA(h, h->d_titr_local.reserve(h->ntitr + 1));
  if (h->ntitr) fill_int_kernel<<<nblk(h->ntitr), TPB, 0, st>>>(h->ntitr, h->d_titr_local.p, -1);
  h->nh = 0;
  if (h->nw_local > 0 && n) {
    fill_int_kernel<<<1, TPB, 0, st>>>(h->nw_local, h->d_wlocal.p, -1);
    water_map_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_tag.p, h->d_mask.p, h->fix.Wbit, h->nw_local, h->d_wtag.p,
                                              h->d_wlocal.p);
  }
  if (n) {
    site_map_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_tag.p, h->d_mask.p, h->ntitr, h->d_titr_tag_sorted.p,
                                             h->d_titr_entry_of_sorted.p, h->d_titr_site.p, h->fix.implicit_site,
                                             h->fix.Hbit, h->d_site_of.p, h->d_titr_of.p, h->d_titr_local.p);
    CPH_CUDA(h, h->d_vals.reserve(n + 1));
    CPH_CUDA(h, h->d_vals2.reserve(n + 1));
    hflag_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_mask.p, h->d_site_of.p, h->fix.Hbit, h->d_vals.p);
    CPH_TRY(exclusive_scan(h, n + 1, h->d_vals.p, h->d_vals2.p));
    int nh = 0;
    CPH_CUDA(h, cudaMemcpyAsync(&nh, h->d_vals2.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    CPH_CUDA(h, cudaStreamSynchronize(st));
    h->nh = nh;
    CPH_CUDA(h, h->d_hlist.reserve(nh + 1));
    if (nh) hlist_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_vals.p, h->d_vals2.p, h->d_hlist.p);
  }
  CPH_TRY(cph_ljstates_map(h));     // LJ end states: B type / site of every owned and ghost atom
  // result arrays
  CPH_CUDA(h, h->d_f.reserve(3 * (size_t)n + 3));
  CPH_CUDA(h, h->d_evdwl.reserve(n + 1));
  CPH_CUDA(h, h->d_phi.reserve(n + 1));
  CPH_CUDA(h, h->d_eatom.reserve(n + 1));
  CPH_CUDA(h, cudaGetLastError());
  tr.mark("sites");
  h->nlaunch += 27;          // this file's kernels per rebuild (cub sort/scan kernels not counted)
  h->inner_valid = false;    // the inner (pruned) rows are rebuilt from the new Verlet rows
  h->nbuilds++;
  CPH_TRY(cph_bonded_resolve(h));   // bond / angle partners in the new internal order
  return 0;
}

// ---- bookkeeping getters (bit-exact parity checks) -------------------------------------------------------
namespace {
__global__ void neighbor_keys_kernel(int nlocal, const int *__restrict__ neigh, const int *__restrict__ numneigh,
                                     const int *__restrict__ numspec, int rowcap, const int *__restrict__ tag, const int *__restrict__ ghost_code,
                                     GhostDirs gd, const int *__restrict__ perm, const long long *__restrict__ off,
                                     long long *keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  long long o = off[perm[i]];
  int nn = numneigh[i], ns = numspec[i];
  for (int k = 0; k < nn + ns; k++) {
    int raw = neigh[(size_t)i * rowcap + (k < nn ? k : rowcap - 1 - (k - nn))];
    int j = raw & CPH_NEIGHMASK, sb = (raw >> CPH_SBSHIFT) & 3;
    int code = 13;
    if (j >= nlocal) {
      int c = ghost_code[j - nlocal];
      code = c >= 32 ? c - 32 : gd.imgcode[c];
    }
    keys[o + k] = ((long long)tag[j] << 8) | (sb << 5) | code;
  }
}
}  // namespace

// numneigh in caller order; keys (optional) concatenated in caller order, sorted per row on the host
int cph_neighbors_to_host(cph_handle *h, int *numneigh, int64_t *keys, int64_t cap) {
  const int n = h->nlocal;
  cudaStream_t st = h->stream;
  std::vector<int> nn_int(n), ns_int(n), perm(n);
  CPH_CUDA(h, cudaMemcpyAsync(nn_int.data(), h->d_numneigh.p, n * sizeof(int), cudaMemcpyDeviceToHost, st));
  CPH_CUDA(h, cudaMemcpyAsync(ns_int.data(), h->d_numspec.p, n * sizeof(int), cudaMemcpyDeviceToHost, st));
  CPH_CUDA(h, cudaMemcpyAsync(perm.data(), h->d_perm.p, n * sizeof(int), cudaMemcpyDeviceToHost, st));
  CPH_CUDA(h, cudaStreamSynchronize(st));
  for (int k = 0; k < n; k++) numneigh[perm[k]] = nn_int[k] + ns_int[k];
  if (!keys) return 0;
  std::vector<long long> off(n + 1, 0);
  for (int c = 0; c < n; c++) off[c + 1] = off[c] + numneigh[c];
  if (off[n] > cap) return cph_fail(h, CPH_ERR_OVERFLOW, "keys capacity %lld < %lld", (long long)cap, off[n]);
  DevBuf<long long> d_off, d_keys;
  CPH_CUDA(h, d_off.reserve(n + 1));
  CPH_CUDA(h, d_keys.reserve(off[n] + 1));
  CPH_CUDA(h, cudaMemcpyAsync(d_off.p, off.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
  GhostDirs gd;
  make_ghost_dirs(h, gd);
  neighbor_keys_kernel<<<nblk(n), TPB, 0, st>>>(n, h->d_neigh.p, h->d_numneigh.p, h->d_numspec.p, h->rowcap, h->d_tag.p,
                                                h->d_ghost_code.p, gd, h->d_perm.p, d_off.p, d_keys.p);
  CPH_CUDA(h, cudaMemcpyAsync(keys, d_keys.p, off[n] * sizeof(long long), cudaMemcpyDeviceToHost, st));
  CPH_CUDA(h, cudaStreamSynchronize(st));
  for (int c = 0; c < n; c++) std::sort(keys + off[c], keys + off[c + 1]);
  d_off.release();
  d_keys.release();
  return 0;
}
