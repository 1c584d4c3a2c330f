"""Seeded synthetic boxes for the constant-pH hot path.

The reference ships no input decks, so every configuration that
BASELINE.json names is generated here, identically for the CPU oracle and the
CUDA path: SPC/E water on a jittered lattice plus 8-atom titratable solutes
(acetic acid: carboxyl site, methylammonium: amine site) that carry a
protonated (A) and a deprotonated (B) charge set, LAMMPS-style special-bond
tables, and the two mask groups the reference's constructor takes
(fix_constant_pH.cpp:39-46: hydrogen group, 3-atom water group).

Pure numpy; no device code.  `units real` throughout.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

# ---- unit system: LAMMPS `units real` -----------------
QQRD2E = 332.06371
BOLTZ = 0.0019872067
FTM2V = 1.0 / 48.88821291 / 48.88821291

STYLE_COUL_CUT = 0   # pair lj/cut/coul/cut
STYLE_COUL_DSF = 1   # pair lj/cut/coul/dsf

GROUP_ALL_BIT = 1
GROUP_H_BIT = 2      # arg[4] of the fix (fix_constant_pH.cpp:39-41)
GROUP_W_BIT = 4      # arg[5] of the fix (fix_constant_pH.cpp:42-46), exactly 3 atoms

# ---- atom types (1-based, LAMMPS convention) ----------------------------------
#            eps [kcal/mol], sigma [A]
_TYPES = {
    1: ("OW", 0.1553, 3.166),
    2: ("HW", 0.0, 0.0),
    3: ("CT", 0.080, 3.50),
    4: ("CC", 0.070, 3.56),
    5: ("OC", 0.120, 3.03),
    6: ("OH", 0.152, 3.15),
    7: ("HO", 0.046, 0.40),
    8: ("HC", 0.022, 2.35),
    9: ("NH", 0.200, 3.29),
    10: ("HN", 0.046, 0.40),
}
NTYPES = 10

# water geometry (SPC/E)
_R_OH = 1.0
_THETA = np.deg2rad(109.47)
_WATER_LOCAL = np.array([
    [0.0, 0.0, 0.0],
    [_R_OH * np.sin(_THETA / 2), _R_OH * np.cos(_THETA / 2), 0.0],
    [-_R_OH * np.sin(_THETA / 2), _R_OH * np.cos(_THETA / 2), 0.0],
])
_WATER_Q = np.array([-0.8476, 0.4238, 0.4238])
_WATER_TYPE = np.array([1, 2, 2], dtype=np.int32)
_WATER_BONDS = [(0, 1), (0, 2)]

# acetic acid, 8 atoms, C-C axis along x, centred near the origin
_ACID_LOCAL = np.array([
    [-0.76, 0.00, 0.00],    # 0 CT methyl carbon
    [0.76, 0.00, 0.00],     # 1 CC carboxyl carbon
    [1.38, 1.05, 0.00],     # 2 OC carbonyl oxygen
    [1.44, -1.12, 0.00],    # 3 OH hydroxyl oxygen
    [2.40, -1.02, 0.00],    # 4 HO titratable proton
    [-1.14, 1.02, 0.00],    # 5 HC
    [-1.14, -0.51, 0.88],   # 6 HC
    [-1.14, -0.51, -0.88],  # 7 HC
])
_ACID_TYPE = np.array([3, 4, 5, 6, 7, 8, 8, 8], dtype=np.int32)
_ACID_QA = np.array([-0.30, 0.75, -0.55, -0.61, 0.44, 0.09, 0.09, 0.09])   # protonated, sum 0
_ACID_QB = np.array([-0.37, 0.62, -0.76, -0.76, 0.00, 0.09, 0.09, 0.09])   # deprotonated, sum -1
_ACID_BONDS = [(0, 1), (1, 2), (1, 3), (3, 4), (0, 5), (0, 6), (0, 7)]
_ACID_HGROUP = [4]

# methylammonium, 8 atoms
_AMINE_LOCAL = np.array([
    [-0.745, 0.00, 0.00],   # 0 CT
    [0.745, 0.00, 0.00],    # 1 NH
    [1.105, -0.96, 0.00],   # 2 HN titratable proton
    [1.105, 0.48, 0.83],    # 3 HN
    [1.105, 0.48, -0.83],   # 4 HN
    [-1.125, 1.02, 0.00],   # 5 HC
    [-1.125, -0.51, 0.88],  # 6 HC
    [-1.125, -0.51, -0.88], # 7 HC
])
_AMINE_TYPE = np.array([3, 9, 10, 10, 10, 8, 8, 8], dtype=np.int32)
_AMINE_QA = np.array([0.13, -0.30, 0.33, 0.33, 0.33, 0.06, 0.06, 0.06])    # protonated, sum +1
_AMINE_QB = np.array([0.10, -0.96, 0.00, 0.34, 0.34, 0.06, 0.06, 0.06])    # deprotonated, sum 0
_AMINE_BONDS = [(0, 1), (1, 2), (1, 3), (1, 4), (0, 5), (0, 6), (0, 7)]
_AMINE_HGROUP = [2]

# poly(acrylic acid) repeat unit -CH2-CH(COOH)-, 9 atoms, backbone along x at one lattice spacing per
# monomer (a slightly stretched all-trans chain), the carboxyl group in the y-z plane; consecutive
# monomers are bonded CH(3) - CH2'(0), so the special-bond lists run across monomers (BASELINE configs[3])
_PAA_LOCAL = np.array([
    [-0.776, -0.30, 0.00],   # 0 CT backbone CH2
    [-0.776, -0.93, 0.89],   # 1 HC
    [-0.776, -0.93, -0.89],  # 2 HC
    [0.776, 0.30, 0.00],     # 3 CT backbone CH
    [0.776, -0.25, 0.95],    # 4 HC
    [0.776, 1.82, 0.00],     # 5 CC carboxyl carbon
    [0.776, 2.44, 1.05],     # 6 OC carbonyl oxygen
    [0.776, 2.50, -1.12],    # 7 OH hydroxyl oxygen
    [0.776, 3.46, -1.02],    # 8 HO titratable proton
])
_PAA_TYPE = np.array([3, 8, 8, 3, 8, 4, 5, 6, 7], dtype=np.int32)
_PAA_QA = np.array([-0.18, 0.09, 0.09, -0.12, 0.09, 0.75, -0.55, -0.61, 0.44])   # protonated, sum 0
_PAA_QB = np.array([-0.18, 0.09, 0.09, -0.19, 0.09, 0.62, -0.76, -0.76, 0.00])   # deprotonated, sum -1
_PAA_BONDS = [(0, 1), (0, 2), (0, 3), (3, 4), (3, 5), (5, 6), (5, 7), (7, 8)]
_PAA_LINK = (3, 0)           # CH of monomer m - CH2 of monomer m+1
_PAA_HGROUP = [8]
_PAA_Y_MID = 1.265           # middle of the monomer's y extent: it sits across two emptied lattice rows


def _paa_chain(nmono, pitch):
    """Local coordinates, types and bond list of one chain of `nmono` repeat units."""
    xs = [_PAA_LOCAL + np.array([m * pitch, 0.0, 0.0]) for m in range(nmono)]
    bonds = []
    for m in range(nmono):
        bonds += [(9 * m + a, 9 * m + b) for a, b in _PAA_BONDS]
        if m + 1 < nmono:
            bonds.append((9 * m + _PAA_LINK[0], 9 * (m + 1) + _PAA_LINK[1]))
    return np.concatenate(xs), np.tile(_PAA_TYPE, nmono), bonds


PK_CARBOXYL = 4.76
PK_AMINE = 10.6


def _special_template(natoms, bonds):
    """1-2 / 1-3 / 1-4 partner lists of a small molecule (LAMMPS `special` layout:
    per atom [1-2 ..., 1-3 ..., 1-4 ...] with cumulative counts in nspecial)."""
    adj = [set() for _ in range(natoms)]
    for a, b in bonds:
        adj[a].add(b)
        adj[b].add(a)
    out = []
    for i in range(natoms):
        d = {i: 0}
        frontier = [i]
        for depth in (1, 2, 3):
            nxt = []
            for u in frontier:
                for v in sorted(adj[u]):
                    if v not in d:
                        d[v] = depth
                        nxt.append(v)
            frontier = nxt
        l12 = sorted(v for v, k in d.items() if k == 1)
        l13 = sorted(v for v, k in d.items() if k == 2)
        l14 = sorted(v for v, k in d.items() if k == 3)
        out.append((l12, l13, l14))
    return out


@dataclass
class Box:
    """One synthetic system in LAMMPS per-atom layout (host arrays, caller order)."""
    name: str
    boxlo: np.ndarray
    boxhi: np.ndarray
    x: np.ndarray          # (n,3) f64
    q: np.ndarray          # (n,) f64 -- charges at lambda0 (QL mode) / qA
    type: np.ndarray       # (n,) i32, 1-based
    tag: np.ndarray        # (n,) i32, 1-based global ids
    mask: np.ndarray       # (n,) i32 group bits
    molecule: np.ndarray   # (n,) i32
    nspecial: np.ndarray   # (n,3) i32 cumulative 1-2,1-3,1-4 counts
    special: np.ndarray    # (n,maxspecial) i32 partner tags
    maxspecial: int
    ntypes: int
    epsilon: np.ndarray    # (ntypes+1, ntypes+1) mixed
    sigma: np.ndarray
    style: int
    cut_lj: float
    cut_coul: float
    alpha: float
    special_lj: np.ndarray   # (4,) [1, f12, f13, f14]
    special_coul: np.ndarray
    skin: float
    # titration sites
    nsites: int
    pK: np.ndarray         # (S,)
    titr_tag: np.ndarray   # (A_t,) i32 tags of atoms whose charge depends on lambda
    titr_site: np.ndarray  # (A_t,) i32 site index
    qA: np.ndarray         # (A_t,)
    qB: np.ndarray         # (A_t,)
    lambda0: np.ndarray    # (S,)
    v0: np.ndarray         # (S,)
    pH: float = 4.8
    T: float = 300.0
    dt: float = 1.0
    meta: dict = field(default_factory=dict)

    @property
    def n(self):
        return int(self.x.shape[0])

    def charges_at(self, lam):
        """q(lambda) = (1-lambda) qA + lambda qB for titratable atoms (north_star)."""
        q = self.q.copy()
        pos = self.meta["tag_to_index"][self.titr_tag]
        q[pos] = self.qA + lam[self.titr_site] * (self.qB - self.qA)
        return q


def mix_geometric(eps, sig):
    """LAMMPS default `pair_modify mix geometric`."""
    nt = len(eps) - 1
    e = np.zeros((nt + 1, nt + 1))
    s = np.zeros((nt + 1, nt + 1))
    for i in range(1, nt + 1):
        for j in range(1, nt + 1):
            e[i, j] = np.sqrt(eps[i] * eps[j])
            s[i, j] = np.sqrt(sig[i] * sig[j])
    return e, s


def _random_rotations(rng, n):
    """n uniformly random rotation matrices from unit quaternions."""
    qn = rng.normal(size=(n, 4))
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    w, x, y, z = qn[:, 0], qn[:, 1], qn[:, 2], qn[:, 3]
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - z * w)
    R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w)
    R[:, 2, 1] = 2 * (y * z + x * w)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _lattice_dims(n_atoms_target):
    """Near-cubic (nx,ny,nz) whose slot count best matches n_atoms_target/3 waters."""
    slots = n_atoms_target / 3.0
    c = int(round(slots ** (1.0 / 3.0)))
    best = None
    for nx in range(max(2, c - 2), c + 3):
        for ny in range(nx, c + 3):
            for nz in range(ny, c + 3):
                err = abs(nx * ny * nz - slots)
                if best is None or err < best[0]:
                    best = (err, (nx, ny, nz))
    return best[1]


def make_box(name="custom", n_atoms=3000, n_acid=1, n_amine=0, style=STYLE_COUL_CUT,
             seed=1, cut=10.0, skin=2.0, alpha=0.2, pH=4.8, T=300.0, dense_titr_frac=0.0,
             jitter=0.3, shuffle=False, special_14=0.5, md_safe=False, n_paa=0, chain_len=25):
    """Build a water + solute box.

    n_acid / n_amine solutes each take three lattice slots along x (the solute
    sits in the middle one) so that no water overlaps them.  dense_titr_frac > 0
    additionally turns that fraction of *water* atoms into single-atom sites
    (BASELINE config 5).  n_paa > 0 adds that many poly(acrylic acid) repeat units, one site each, bonded
    into chains of `chain_len` units that run along x across two emptied lattice rows (BASELINE config 4).

    md_safe: a start that real dynamics can run from.  The default placement lets
    solutes sit in adjacent lattice rows, where their 8-atom bodies interpenetrate -- harmless for
    prescribed motion, an LJ-core explosion within five steps of an integrator.  With md_safe solutes
    are at least two lattice rows apart in y and z and the four water slots beside them are emptied too.
    """
    rng = np.random.default_rng(seed)
    spacing = (1.0 / 0.0334) ** (1.0 / 3.0)
    nx, ny, nz = _lattice_dims(n_atoms)
    dims = np.array([nx, ny, nz])
    L = dims * spacing
    nsol = n_acid + n_amine

    occupied = np.zeros((nx, ny, nz), dtype=np.int8)    # 0 water, 1 solute centre, 2 emptied
    centres = []
    if nsol:
        stride = 2 if md_safe else 1
        cand = [(i, j, k) for i in range(1, nx - 1, 3) for j in range(0, ny - stride + 1, stride)
                for k in range(0, nz - stride + 1, stride)]
        if len(cand) < nsol:
            raise ValueError("box too small for %d solutes" % nsol)
        pick = rng.choice(len(cand), size=nsol, replace=False)
        pick.sort()
        for p in pick:
            i, j, k = cand[p]
            occupied[i, j, k] = 1
            occupied[i - 1, j, k] = 2
            occupied[i + 1, j, k] = 2
            if md_safe:
                for dj, dk in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                    occupied[i, (j + dj) % ny, (k + dk) % nz] = 2
            centres.append((i, j, k))
    # chains: `chain_len` slots along x in rows j and j+1, one free slot between chains of a row, every other
    # z layer (the carboxyl groups of chains in adjacent layers would interpenetrate)
    chains = []
    if n_paa:
        if n_paa % chain_len or chain_len + 1 > nx:
            raise ValueError("n_paa must be a multiple of chain_len and a chain must fit the box")
        ccand = [(i, j, k) for i in range(0, nx - chain_len + 1, chain_len + 1) for j in range(0, ny - 1, 2)
                 for k in range(0, nz, 2) if not occupied[i:i + chain_len, j:j + 2, k].any()]
        if len(ccand) < n_paa // chain_len:
            raise ValueError("box too small for %d chains" % (n_paa // chain_len))
        pick = rng.choice(len(ccand), size=n_paa // chain_len, replace=False)
        pick.sort()
        for p in pick:
            i, j, k = ccand[p]
            occupied[i:i + chain_len, j:j + 2, k] = 2
            chains.append((i, j, k))
    wi, wj, wk = np.nonzero(occupied == 0)
    nwat = wi.size

    # --- waters -------------------------------------------------------------
    cw = (np.stack([wi, wj, wk], axis=1) + 0.5) * spacing
    cw += rng.uniform(-jitter, jitter, size=cw.shape)
    R = _random_rotations(rng, nwat)
    xw = cw[:, None, :] + np.einsum("nij,aj->nai", R, _WATER_LOCAL)     # (nwat,3,3)
    n_w_atoms = 3 * nwat
    x = [xw.reshape(-1, 3)]
    q = [np.tile(_WATER_Q, nwat)]
    typ = [np.tile(_WATER_TYPE, nwat)]
    mol = [np.repeat(np.arange(1, nwat + 1, dtype=np.int32), 3)]
    maxspecial = 7 if nsol else 2
    wt = _special_template(3, _WATER_BONDS)
    if n_paa:
        cx, ctype, cbonds = _paa_chain(chain_len, spacing)
        ct = _special_template(9 * chain_len, cbonds)
        maxspecial = max(maxspecial, max(len(a) + len(b) + len(c) for a, b, c in ct))

    n = n_w_atoms + 8 * nsol + 9 * n_paa
    nspecial = np.zeros((n, 3), dtype=np.int32)
    special = np.zeros((n, maxspecial), dtype=np.int32)
    base = np.arange(nwat, dtype=np.int32) * 3 + 1                       # tag of atom 0 of each water
    for a in range(3):
        l12, l13, l14 = wt[a]
        rows = np.arange(nwat) * 3 + a
        nspecial[rows, 0] = len(l12)
        nspecial[rows, 1] = len(l12) + len(l13)
        nspecial[rows, 2] = len(l12) + len(l13) + len(l14)
        for c, b in enumerate(l12 + l13 + l14):
            special[rows, c] = base + b

    mask = np.full(n, GROUP_ALL_BIT, dtype=np.int32)
    mask[0:3] |= GROUP_W_BIT                    # first water = the fix's 3-atom water group

    # --- solutes ------------------------------------------------------------
    titr_tag, titr_site, qA, qB, pK = [], [], [], [], []
    at = _special_template(8, _ACID_BONDS)
    mt = _special_template(8, _AMINE_BONDS)
    kinds = np.array([0] * n_acid + [1] * n_amine)
    if nsol:
        kinds = kinds[rng.permutation(nsol)]
    off = n_w_atoms
    for s, (i, j, k) in enumerate(centres):
        kind = int(kinds[s])
        local = _ACID_LOCAL if kind == 0 else _AMINE_LOCAL
        ang = rng.uniform(0, 2 * np.pi)
        ca, sa = np.cos(ang), np.sin(ang)
        Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]])
        c0 = (np.array([i, j, k]) + 0.5) * spacing + rng.uniform(-0.2, 0.2, size=3)
        x.append(c0 + local @ Rx.T)
        a_q, b_q = (_ACID_QA, _ACID_QB) if kind == 0 else (_AMINE_QA, _AMINE_QB)
        q.append(a_q.copy())
        typ.append(_ACID_TYPE if kind == 0 else _AMINE_TYPE)
        mol.append(np.full(8, nwat + 1 + s, dtype=np.int32))
        tmpl = at if kind == 0 else mt
        for a in range(8):
            l12, l13, l14 = tmpl[a]
            row = off + a
            nspecial[row] = (len(l12), len(l12) + len(l13), len(l12) + len(l13) + len(l14))
            for c, b in enumerate(l12 + l13 + l14):
                special[row, c] = off + b + 1
        for a in (_ACID_HGROUP if kind == 0 else _AMINE_HGROUP):
            mask[off + a] |= GROUP_H_BIT
        for a in range(8):
            if a_q[a] != b_q[a]:
                titr_tag.append(off + a + 1)
                titr_site.append(s)
                qA.append(a_q[a])
                qB.append(b_q[a])
        pK.append(PK_CARBOXYL if kind == 0 else PK_AMINE)
        off += 8
    # --- poly(acrylic acid) chains: one molecule per chain, one site per repeat unit ---------------
    nchain_atoms = 9 * chain_len
    for c, (i, j, k) in enumerate(chains):
        c0 = np.array([(i + 0.5) * spacing, (j + 1.0) * spacing - _PAA_Y_MID, (k + 0.5) * spacing])
        x.append(c0 + cx)
        q.append(np.tile(_PAA_QA, chain_len))
        typ.append(ctype)
        mol.append(np.full(nchain_atoms, nwat + 1 + nsol + c, dtype=np.int32))
        for a in range(nchain_atoms):
            l12, l13, l14 = ct[a]
            nspecial[off + a] = (len(l12), len(l12) + len(l13), len(l12) + len(l13) + len(l14))
            for cc, b in enumerate(l12 + l13 + l14):
                special[off + a, cc] = off + b + 1
        for m in range(chain_len):
            site = nsol + c * chain_len + m
            for a in _PAA_HGROUP:
                mask[off + 9 * m + a] |= GROUP_H_BIT
            for a in range(9):
                if _PAA_QA[a] != _PAA_QB[a]:
                    titr_tag.append(off + 9 * m + a + 1)
                    titr_site.append(site)
                    qA.append(_PAA_QA[a])
                    qB.append(_PAA_QB[a])
            pK.append(PK_CARBOXYL)
        off += nchain_atoms

    x = np.concatenate(x).astype(np.float64)
    q = np.concatenate(q).astype(np.float64)
    typ = np.concatenate(typ).astype(np.int32)
    mol = np.concatenate(mol).astype(np.int32)
    tag = np.arange(1, n + 1, dtype=np.int32)

    nsites = nsol + n_paa
    # --- dense single-atom sites on water atoms (config 5) --------------------
    if dense_titr_frac > 0.0:
        nd = int(round(dense_titr_frac * n))
        pick = rng.choice(n_w_atoms, size=nd, replace=False)
        pick.sort()
        dq = rng.uniform(-0.15, 0.15, size=nd)
        for c, p in enumerate(pick):
            titr_tag.append(int(p) + 1)
            titr_site.append(nsites + c)
            qA.append(q[p])
            qB.append(q[p] + dq[c])
            pK.append(PK_CARBOXYL if c % 2 == 0 else PK_AMINE)
        mask[pick[typ[pick] == 2]] |= GROUP_H_BIT
        nsites += nd

    # wrap into the periodic box, atom by atom (LAMMPS remaps atoms, not molecules)
    x -= np.floor(x / L) * L
    x = np.minimum(x, np.nextafter(L, 0.0))

    lam0 = np.zeros(nsites)
    half = nsites // 2
    lam0[:half] = 0.5
    lam0[half:] = (np.arange(nsites - half) % 2).astype(np.float64)
    if nsites == 1:
        lam0[:] = 0.5

    titr_tag = np.array(titr_tag, dtype=np.int32)
    titr_site = np.array(titr_site, dtype=np.int32)
    qA = np.array(qA, dtype=np.float64)
    qB = np.array(qB, dtype=np.float64)
    pK = np.array(pK, dtype=np.float64)

    if shuffle:
        perm = rng.permutation(n)
        x, q, typ, tag, mask, mol = x[perm], q[perm], typ[perm], tag[perm], mask[perm], mol[perm]
        nspecial, special = nspecial[perm], special[perm]
    tag_to_index = np.zeros(n + 1, dtype=np.int64)
    tag_to_index[tag] = np.arange(n)

    # charges at lambda0
    if titr_tag.size:
        q[tag_to_index[titr_tag]] = qA + lam0[titr_site] * (qB - qA)

    eps = np.zeros(NTYPES + 1)
    sig = np.zeros(NTYPES + 1)
    for t, (_, e, s_) in _TYPES.items():
        eps[t], sig[t] = e, s_
    epsilon, sigma = mix_geometric(eps, sig)

    return Box(
        name=name, boxlo=np.zeros(3), boxhi=L.astype(np.float64),
        x=np.ascontiguousarray(x), q=np.ascontiguousarray(q), type=np.ascontiguousarray(typ),
        tag=np.ascontiguousarray(tag), mask=np.ascontiguousarray(mask),
        molecule=np.ascontiguousarray(mol), nspecial=np.ascontiguousarray(nspecial),
        special=np.ascontiguousarray(special), maxspecial=maxspecial, ntypes=NTYPES,
        epsilon=epsilon, sigma=sigma, style=style, cut_lj=cut, cut_coul=cut, alpha=alpha,
        special_lj=np.array([1.0, 0.0, 0.0, special_14]),
        special_coul=np.array([1.0, 0.0, 0.0, special_14]),
        skin=skin, nsites=nsites, pK=pK, titr_tag=titr_tag, titr_site=titr_site, qA=qA, qB=qB,
        lambda0=lam0, v0=np.zeros(nsites), pH=pH, T=T, dt=1.0,
        meta={"seed": seed, "lattice": (nx, ny, nz), "n_water": int(nwat), "n_acid": n_acid,
              "n_amine": n_amine, "n_paa": n_paa, "chain_len": chain_len if n_paa else 0,
              "tag_to_index": tag_to_index},
    )


# ---- the configurations BASELINE.json names ---------------------
def config(k, scale=1.0, **kw):
    """BASELINE.json configs[k-1].  `scale` < 1 shrinks atom and site counts
    proportionally (parity tests run the large configs at reduced size)."""
    if k == 1:
        return make_box("cfg1", n_atoms=3000, n_acid=1, style=STYLE_COUL_CUT, seed=1, pH=4.8, **kw)
    if k == 2:
        return make_box("cfg2", n_atoms=int(32000 * scale), n_acid=max(1, int(10 * scale)),
                        n_amine=max(1, int(10 * scale)), style=STYLE_COUL_DSF, seed=2,
                        pH=kw.pop("pH", 7.0), **kw)
    if k == 3:
        ns = max(2, int(2000 * scale))
        return make_box("cfg3", n_atoms=int(1_000_000 * scale), n_acid=ns // 2, n_amine=ns - ns // 2,
                        style=STYLE_COUL_DSF, seed=3, pH=kw.pop("pH", 7.0), **kw)
    if k == 4:
        # 4M atoms, 50 000 titratable repeat units of poly(acrylic acid) in bonded 25-unit chains, in water
        clen = kw.pop("chain_len", 25)
        ns = max(clen, int(50_000 * scale) // clen * clen)
        # a repeat unit (9 atoms) takes the place of two waters (6 atoms): aim the lattice at the stated total
        return make_box("cfg4", n_atoms=int(4_000_000 * scale) - 3 * ns, n_acid=0, n_amine=0, n_paa=ns, chain_len=clen,
                        style=STYLE_COUL_DSF, seed=4, pH=kw.pop("pH", 4.8), **kw)
    if k == 5:
        return make_box("cfg5", n_atoms=int(512_000 * scale), n_acid=0, n_amine=0,
                        style=STYLE_COUL_DSF, seed=5, dense_titr_frac=0.10,
                        pH=kw.pop("pH", 7.0), **kw)
    raise ValueError("unknown config %r" % (k,))


def lj_end_state_types(box):
    """B-state atom types for the LJ-end-state tests (docs/SPEC.md): a titratable proton loses its LJ site in the
    deprotonated state (HO / HN -> HW, epsilon 0), the hydroxyl oxygen becomes a carbonyl oxygen, a water oxygen
    (config 5's one-atom sites) a hydroxyl oxygen, a water hydrogen gains the small HO site; everything else keeps one
    identity (0).  Aligned with box.titr_tag."""
    swap = {7: 2, 10: 2, 6: 5, 1: 6, 2: 7}
    t = box.type[box.meta["tag_to_index"][box.titr_tag]]
    return np.array([swap.get(int(v), 0) for v in t], dtype=np.int32)


def jiggle_params(box, amp=0.45, period_lo=60.0, period_hi=140.0, seed=12345):
    """Prescribed rigid-molecule motion x_i(t) = x_i(0) + A_m sin(w_m t + p_m) used by the
    tests and the bench as the stand-in for the host MD integrator (which is LAMMPS's
    job, not the fix's).  Keyed on molecule id so every rank derives the same motion."""
    nm = int(box.molecule.max()) + 1
    rng = np.random.default_rng(seed)
    A = rng.uniform(-amp, amp, size=(nm, 3))
    w = 2 * np.pi / rng.uniform(period_lo, period_hi, size=(nm, 3))
    p = rng.uniform(0, 2 * np.pi, size=(nm, 3))
    return A, w, p


def jiggle_positions(box, params, t, x0=None):
    """Positions at time t (fs) under `jiggle_params`; not wrapped (drift is < 1 A)."""
    A, w, p = params
    x0 = box.x if x0 is None else x0
    m = box.molecule
    return x0 + A[m] * (np.sin(w[m] * t + p[m]) - np.sin(p[m]))


def harness_jiggle(x0, amp, t):
    """The atom motion of src/harness.cpp's `jiggle AMP` option, replayed for the oracle:
    x_i(t) = x0_i + AMP sin(2 pi t / T_i + phi_i + d), T_i = 60 + i % 80, phi_i = 0.37 i, d = 0, 1, 2."""
    i = np.arange(x0.shape[0], dtype=np.float64)
    w = 2.0 * np.pi / (60.0 + (np.arange(x0.shape[0]) % 80))
    ph = 0.37 * i
    return x0 + amp * np.sin((w * t + ph)[:, None] + np.arange(3.0)[None, :])


def write_harness_input(box, path_bin, path_sites=None, lj_typeB=None):
    """Binary box + text site table for src/cph_harness (the drop-in fix driven through the
    LAMMPS shim).  Layout documented in src/harness.cpp."""
    pK0 = float(box.pK[0]) if box.nsites else 0.0
    with open(path_bin, "wb") as fh:
        np.array([box.n, box.ntypes, box.maxspecial, box.style, box.nsites, box.titr_tag.size,
                  GROUP_H_BIT, GROUP_W_BIT], dtype=np.int32).tofile(fh)
        hd = np.concatenate([box.boxlo, box.boxhi, [box.cut_lj, box.cut_coul, box.alpha, box.skin],
                             box.special_lj, box.special_coul, [box.pH, box.T, box.dt, pK0]]).astype(np.float64)
        assert hd.size == 22
        hd.tofile(fh)
        for arr, dt in ((box.x, np.float64), (box.q, np.float64), (box.type, np.int32), (box.tag, np.int32),
                        (box.mask, np.int32), (box.molecule, np.int32), (box.nspecial, np.int32),
                        (box.special, np.int32), (box.epsilon, np.float64), (box.sigma, np.float64)):
            np.ascontiguousarray(arr, dtype=dt).tofile(fh)
    if path_sites is not None:
        with open(path_sites, "w") as fh:
            fh.write("%d %d\n" % (box.nsites, box.titr_tag.size))
            for s in range(box.nsites):
                fh.write("%.17g %.17g\n" % (box.pK[s], box.lambda0[s]))
            for t in range(box.titr_tag.size):
                fh.write("%d %d %.17g %.17g%s\n" % (box.titr_tag[t], box.titr_site[t], box.qA[t], box.qB[t],
                                                    "" if lj_typeB is None else " %d" % lj_typeB[t]))


# ---- flexible molecules: bonded topology and masses -----------------------
_MASS = {1: 15.9994, 2: 1.008, 3: 12.011, 4: 12.011, 5: 15.9994, 6: 15.9994, 7: 1.008, 8: 1.008,
         9: 14.0067, 10: 1.008}
# SPC/Fw flexible water in LAMMPS' E = K (r - r0)^2 convention; every other bond / angle type takes
# its equilibrium value from the solute template and a generic stiffness
_WATER_BOND = (529.581, 1.012)
_WATER_ANGLE = (37.95, np.deg2rad(113.24))
_SOLUTE_BOND_K = 300.0
_SOLUTE_ANGLE_K = 50.0


@dataclass
class Topology:
    """Per-atom incident bond / angle lists in LAMMPS' `newton_bond off` layout: every bond is stored
    with both of its atoms and every angle with all three (atom->num_bond, bond_type, bond_atom,
    num_angle, angle_type, angle_atom1/2/3), partner ids are tags."""
    bond_k: np.ndarray         # (nbondtypes+1,)
    bond_r0: np.ndarray
    angle_k: np.ndarray        # (nangletypes+1,)
    angle_theta0: np.ndarray   # radians
    maxbond: int
    num_bond: np.ndarray       # (n,)
    bond_type: np.ndarray      # (n,maxbond)
    bond_atom: np.ndarray      # (n,maxbond) partner tag
    maxangle: int
    num_angle: np.ndarray
    angle_type: np.ndarray     # (n,maxangle)
    angle_atom1: np.ndarray
    angle_atom2: np.ndarray    # centre
    angle_atom3: np.ndarray
    mass: np.ndarray           # (ntypes+1,)

    @property
    def nbondtypes(self):
        return self.bond_k.size - 1

    @property
    def nangletypes(self):
        return self.angle_k.size - 1


def _type_tables():
    """Bond types keyed on the unordered atom-type pair, angle types on (end, centre, end)."""
    bonds, angles = {}, {}
    spacing = (1.0 / 0.0334) ** (1.0 / 3.0)
    for local, types, blist in ((_WATER_LOCAL, _WATER_TYPE, _WATER_BONDS), (_ACID_LOCAL, _ACID_TYPE, _ACID_BONDS),
                                (_AMINE_LOCAL, _AMINE_TYPE, _AMINE_BONDS), _paa_chain(3, spacing)):
        water = local is _WATER_LOCAL
        adj = [[] for _ in range(len(types))]
        for a, b in blist:
            adj[a].append(b)
            adj[b].append(a)
            key = (min(types[a], types[b]), max(types[a], types[b]))
            if key not in bonds:
                r0 = float(np.linalg.norm(local[a] - local[b]))
                bonds[key] = _WATER_BOND if water else (_SOLUTE_BOND_K, r0)
        for c, nb in enumerate(adj):
            for u in range(len(nb)):
                for w in range(u + 1, len(nb)):
                    a, b = nb[u], nb[w]
                    key = (min(types[a], types[b]), int(types[c]), max(types[a], types[b]))
                    if key not in angles:
                        d1, d2 = local[a] - local[c], local[b] - local[c]
                        th = float(np.arccos(np.dot(d1, d2) / np.linalg.norm(d1) / np.linalg.norm(d2)))
                        angles[key] = _WATER_ANGLE if water else (_SOLUTE_ANGLE_K, th)
    return bonds, angles


def topology(box):
    """Bonded topology of `box`, derived from its 1-2 special lists (a 1-2 partner IS a bond)."""
    n = box.n
    t2i = box.meta["tag_to_index"]
    btab, atab = _type_tables()
    bkeys, akeys = sorted(btab), sorted(atab)
    bond_k = np.array([0.0] + [btab[k][0] for k in bkeys])
    bond_r0 = np.array([0.0] + [btab[k][1] for k in bkeys])
    angle_k = np.array([0.0] + [atab[k][0] for k in akeys])
    angle_t0 = np.array([0.0] + [atab[k][1] for k in akeys])
    blook = np.zeros((NTYPES + 1, NTYPES + 1), dtype=np.int32)
    for idx, (a, b) in enumerate(bkeys):
        blook[a, b] = blook[b, a] = idx + 1
    alook = np.zeros((NTYPES + 1, NTYPES + 1, NTYPES + 1), dtype=np.int32)
    for idx, (a, c, b) in enumerate(akeys):
        alook[a, c, b] = alook[b, c, a] = idx + 1

    num_bond = box.nspecial[:, 0].astype(np.int32).copy()
    maxbond = max(1, int(num_bond.max()) if n else 1)
    bond_atom = np.zeros((n, maxbond), dtype=np.int32)
    bond_type = np.zeros((n, maxbond), dtype=np.int32)
    pidx = np.zeros((n, maxbond), dtype=np.int64)
    for m in range(maxbond):
        rows = np.nonzero(num_bond > m)[0]
        tg = box.special[rows, m]
        bond_atom[rows, m] = tg
        pidx[rows, m] = t2i[tg]
        bond_type[rows, m] = blook[box.type[rows], box.type[pidx[rows, m]]]
    assert (bond_type[np.arange(maxbond)[None, :] < num_bond[:, None]] > 0).all(), "bond without a type"

    # incident angles: centred on the atom (pairs of its partners), or through a partner (atom is an end)
    triples = []                                     # (rows, a1, a2, a3) tag arrays
    for m1 in range(maxbond):
        for m2 in range(m1 + 1, maxbond):
            rows = np.nonzero(num_bond > m2)[0]
            triples.append((rows, bond_atom[rows, m1], box.tag[rows], bond_atom[rows, m2]))
    for m in range(maxbond):
        rows_m = np.nonzero(num_bond > m)[0]
        c = pidx[rows_m, m]
        for m2 in range(maxbond):
            ok = (num_bond[c] > m2)
            far = np.where(ok, bond_atom[c, np.minimum(m2, maxbond - 1)], 0)
            ok &= far != box.tag[rows_m]
            rows = rows_m[ok]
            triples.append((rows, box.tag[rows], box.tag[c[ok]], far[ok]))
    num_angle = np.zeros(n, dtype=np.int32)
    for rows, _, _, _ in triples:
        np.add.at(num_angle, rows, 1)
    maxangle = max(1, int(num_angle.max()) if n else 1)
    a1 = np.zeros((n, maxangle), dtype=np.int32)
    a2 = np.zeros((n, maxangle), dtype=np.int32)
    a3 = np.zeros((n, maxangle), dtype=np.int32)
    at = np.zeros((n, maxangle), dtype=np.int32)
    cnt = np.zeros(n, dtype=np.int64)
    for rows, t1, t2, t3 in triples:
        if rows.size == 0:
            continue
        k = cnt[rows]
        a1[rows, k], a2[rows, k], a3[rows, k] = t1, t2, t3
        at[rows, k] = alook[box.type[t2i[t1]], box.type[t2i[t2]], box.type[t2i[t3]]]
        cnt[rows] += 1
    assert (cnt == num_angle).all()
    assert (at[np.arange(maxangle)[None, :] < num_angle[:, None]] > 0).all(), "angle without a type"
    mass = np.zeros(NTYPES + 1)
    for t, mval in _MASS.items():
        mass[t] = mval
    return Topology(bond_k=bond_k, bond_r0=bond_r0, angle_k=angle_k, angle_theta0=angle_t0, maxbond=maxbond,
                    num_bond=num_bond, bond_type=np.ascontiguousarray(bond_type),
                    bond_atom=np.ascontiguousarray(bond_atom), maxangle=maxangle, num_angle=num_angle,
                    angle_type=np.ascontiguousarray(at), angle_atom1=np.ascontiguousarray(a1),
                    angle_atom2=np.ascontiguousarray(a2), angle_atom3=np.ascontiguousarray(a3), mass=mass)


def thermal_velocities(box, topo, T=None, seed=777):
    """Maxwell-Boltzmann velocities [A/fs] at temperature T with the net momentum removed."""
    T = box.T if T is None else T
    rng = np.random.default_rng(seed)
    m = topo.mass[box.type]
    # (1/2) m v^2 mvv2e = (1/2) kB T per component;  mvv2e = 1 / ftm2v in `units real`
    sd = np.sqrt(BOLTZ * T * FTM2V / m)
    v = rng.normal(size=(box.n, 3)) * sd[:, None]
    v -= (m[:, None] * v).sum(axis=0) / m.sum()
    return np.ascontiguousarray(v)
