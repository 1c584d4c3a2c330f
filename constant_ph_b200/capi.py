"""ctypes binding of the C ABI in include/cph_b200.h.

`Engine("cph")` drives libcph_b200.so (the CUDA product) and is the only engine this package
knows how to load.  The call table is generic over the symbol prefix, so that test
infrastructure can register a second implementation of the same ABI to check against
(`register_library`; the checker's own loader does that when ITS module is imported by tests/,
smoke() or bench.py -- nothing in this package imports, builds, loads or names it).  There is
no fallback: if libcph_b200.so is missing or no CUDA device is present, Engine("cph") raises.

Method names are the C names without the prefix; argument meaning follows the header
(which cites the reference line each entry point stands in for).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HOST, DEVICE = 0, 1
PAIR_COUL_CUT, PAIR_COUL_DSF, PAIR_COUL_LONG = 0, 1, 2
KSPACE_NONE, KSPACE_EWALD = 0, 1
DUDL_REFERENCE, DUDL_CHARGE = 0, 1
INTEGRATE_REFERENCE, INTEGRATE_VV = 0, 1
BIAS_EXACT, BIAS_AS_WRITTEN = 0, 1
FSCALE_LAMBDA, FSCALE_ONE_MINUS = 0, 1

# Donnini-2016 Table S2 constants loaded by FixConstantPH::init (fix_constant_pH.cpp:86-96)
BIAS_DEFAULT = dict(w=200.0, s=0.3, hbar=4.0, k=2.533, a=0.034041, b=0.005238, r=16.458,
                    m=0.1507, d=2.0, m_lambda=20.0)

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# CPH_B200_LIB selects another build of the SAME CUDA library (kernel tuning variants under constant_ph_b200/csrc/)
CUDA_LIB = os.environ.get("CPH_B200_LIB") or os.path.join(_ROOT, "constant_ph_b200", "csrc", "libcph_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)


class CphError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cph error %d: %s" % (code, msg))
        self.code = code


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


_libs = {}
_registered = {}     # prefix -> loader(variant) of another implementation of the same ABI (test infrastructure)


def register_library(prefix, loader):
    """Make Engine(prefix) available.  `loader(variant)` returns a ctypes library exporting the ABI of
    include/cph_b200.h under `<prefix>_` names.  Called by checkers, never by this package."""
    if prefix == "cph":
        raise ValueError("the product library is not replaceable")
    _registered[prefix] = loader


def load_library(prefix, variant=False):
    key = (prefix, variant)
    if key in _libs:
        return _libs[key]
    if prefix == "cph":
        if not os.path.exists(CUDA_LIB):
            raise CphError(-3, "libcph_b200.so is not built (run __graft_entry__.build()); "
                               "there is no CPU fallback")
        lib = C.CDLL(CUDA_LIB, mode=C.RTLD_GLOBAL)
    elif prefix in _registered:
        lib = _registered[prefix](variant)
    else:
        raise CphError(-1, "no engine %r: this package ships only the CUDA library" % (prefix,))
    lib_last = getattr(lib, prefix + "_last_error")
    lib_last.restype = C.c_char_p
    lib_last.argtypes = [C.c_void_p]
    _libs[key] = lib
    return lib


class Engine:
    """One handle (one rank / one GPU)."""

    def __init__(self, prefix="cph", device=0, variant=False):
        self.prefix = prefix
        self.lib = load_library(prefix, variant)
        self.h = C.c_void_p()
        self.nlocal = 0
        self.nsites = 1
        rc = self._fn("create")(C.c_int(device), C.byref(self.h))
        if rc != 0:
            msg = self._fn("last_error")(None)
            raise CphError(rc, (msg or b"").decode())

    # -- plumbing -------------------------------------------------------------
    def _fn(self, name):
        return getattr(self.lib, "%s_%s" % (self.prefix, name))

    def _call(self, name, *args):
        rc = self._fn(name)(self.h, *args)
        if rc != 0:
            msg = self._fn("last_error")(self.h)
            raise CphError(rc, (msg or b"").decode())

    def close(self):
        if self.h:
            self._fn("destroy")(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration ----------------------------------------------------------
    def set_units(self, qqrd2e, boltz, ftm2v):
        self._call("set_units", C.c_double(qqrd2e), C.c_double(boltz), C.c_double(ftm2v))

    def set_pair(self, style, ntypes, epsilon, sigma, cut_lj, cut_lj_global, cut_coul, alpha,
                 special_lj, special_coul):
        e, s, c = _f64(epsilon), _f64(sigma), _f64(cut_lj)
        sl, sc = _f64(special_lj), _f64(special_coul)
        self._call("set_pair", C.c_int(style), C.c_int(ntypes), _d(e), _d(s), _d(c),
                   C.c_double(cut_lj_global), C.c_double(cut_coul), C.c_double(alpha), _d(sl), _d(sc))

    def set_domain(self, boxlo, boxhi, periodic=(1, 1, 1), sublo=None, subhi=None,
                   procgrid=(1, 1, 1), myloc=(0, 0, 0), skin=2.0):
        lo, hi = _f64(boxlo), _f64(boxhi)
        slo = _f64(boxlo if sublo is None else sublo)
        shi = _f64(boxhi if subhi is None else subhi)
        per, pg, ml = _i32(periodic), _i32(procgrid), _i32(myloc)
        self._call("set_domain", _d(lo), _d(hi), _i(per), _d(slo), _d(shi), _i(pg), _i(ml),
                   C.c_double(skin))

    def set_fix(self, nevery, groupHbit, groupWbit, pK, pH, T):
        self._call("set_fix", C.c_int(nevery), C.c_int(groupHbit), C.c_int(groupWbit),
                   C.c_double(pK), C.c_double(pH), C.c_double(T))

    def set_bias(self, mode=BIAS_EXACT, **kw):
        p = dict(BIAS_DEFAULT)
        p.update(kw)
        self._call("set_bias", *[C.c_double(p[k]) for k in
                                 ("w", "s", "hbar", "k", "a", "b", "r", "m", "d", "m_lambda")],
                   C.c_int(mode))

    def set_mode(self, dudl=DUDL_CHARGE, integrator=INTEGRATE_REFERENCE, fscale=FSCALE_LAMBDA):
        self._call("set_mode", C.c_int(dudl), C.c_int(integrator), C.c_int(fscale))

    def set_thermostat(self, tau):
        self._call("set_thermostat", C.c_double(tau))

    def set_excluded_policy(self, drop):
        self._call("set_excluded_policy", C.c_int(1 if drop else 0))

    def set_extra_partition(self, dHA, dHB):
        self._call("set_extra_partition", C.c_double(dHA), C.c_double(dHB))

    def set_extra_dudl(self, dudl):
        d = _f64(dudl)
        self._call("set_extra_dudl", C.c_int(int(d.size)), _d(d))

    def set_kspace(self, style, g_ewald=0.0, kmax=(0, 0, 0)):
        """kspace_style ewald on the device (after set_domain); pair style PAIR_COUL_LONG is its real-space part."""
        self._call("set_kspace", C.c_int(style), C.c_double(g_ewald), C.c_int(int(kmax[0])), C.c_int(int(kmax[1])),
                   C.c_int(int(kmax[2])))

    def get_kspace_energy(self):
        e = C.c_double(0.0)
        self._call("get_kspace_energy", C.byref(e))
        return float(e.value)

    def set_coordinate(self, theta=True):
        self._call("set_coordinate", C.c_int(1 if theta else 0))

    def set_water_buffer(self, enable=True):
        self._call("set_water_buffer", C.c_int(1 if enable else 0))

    def set_sites(self, nsites, pK, titr_tag, titr_site, qA, qB):
        pK, qA, qB = _f64(pK), _f64(qA), _f64(qB)
        tt, ts = _i32(titr_tag), _i32(titr_site)
        ntitr = 0 if tt is None else int(tt.size)
        self._call("set_sites", C.c_int(nsites), _d(pK), C.c_int(ntitr), _i(tt), _i(ts), _d(qA), _d(qB))
        self.nsites = max(1, int(nsites))

    def set_lj_states(self, typeB):
        """B-state atom type of every titratable atom of set_sites (0 = one LJ identity); before set_atoms."""
        tb = _i32(typeB)
        self._call("set_lj_states", C.c_int(int(tb.size)), _i(tb))

    def set_lambda(self, lam, v=None):
        l, vv = _f64(lam), _f64(v)
        self._call("set_lambda", _d(l), _d(vv))

    # -- rank group ------------------------------------------------------------------
    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        rc = self._fn("comm_unique_id")(buf)
        if rc != 0:
            raise CphError(rc, "ncclGetUniqueId failed")
        return buf.raw

    def comm_init_nccl(self, nranks, rank, id128):
        self._call("comm_init_nccl", C.c_int(nranks), C.c_int(rank), C.c_char_p(id128))

    # -- atoms -------------------------------------------------------------------------
    def set_atoms(self, x, q, type, tag, mask, molecule=None, nspecial=None, special=None, maxspecial=0):
        x, q = _f64(x), _f64(q)
        ty, tg, mk = _i32(type), _i32(tag), _i32(mask)
        mo, ns, sp = _i32(molecule), _i32(nspecial), _i32(special)
        n = int(q.size)
        self._call("set_atoms", C.c_int(HOST), C.c_int(n), _d(x), _d(q), _i(ty), _i(tg), _i(mk), _i(mo),
                   _i(ns), _i(sp), C.c_int(maxspecial))
        self.nlocal = n

    # -- per step ------------------------------------------------------------------------
    def set_x(self, x, where=HOST):
        if where == HOST:
            x = _f64(x)
            self._call("set_x", C.c_int(HOST), _d(x))
        else:
            self._call("set_x", C.c_int(DEVICE), C.cast(C.c_void_p(int(x)), _dp))

    def check_rebuild(self):
        flag = C.c_int(0)
        self._call("check_rebuild", C.byref(flag))
        return int(flag.value)

    def forward(self):
        self._call("forward")

    def pair_pass(self, eflag=1):
        self._call("pair_pass", C.c_int(eflag))

    def site_reduce(self):
        self._call("site_reduce")

    def integrate_lambda(self, dt):
        self._call("integrate_lambda", C.c_double(dt))

    def initial_integrate(self, dt):
        self._call("initial_integrate", C.c_double(dt))

    def final_integrate(self, dt):
        self._call("final_integrate", C.c_double(dt))

    def apply_charges(self):
        self._call("apply_charges")

    def set_force(self):
        self._call("set_force")

    def post_force(self, ntimestep, dt, x=None, f=None, where=HOST):
        """x / f: numpy arrays (HOST) or raw device addresses (DEVICE); either may be None."""
        if where == HOST:
            x = _f64(x)
            if f is not None:
                assert f.dtype == np.float64 and f.flags.c_contiguous
            xp, fp = _d(x), _d(f)
        else:
            xp = None if x is None else C.cast(C.c_void_p(int(x)), _dp)
            fp = None if f is None else C.cast(C.c_void_p(int(f)), _dp)
        self._call("post_force", C.c_int64(int(ntimestep)), C.c_double(dt), C.c_int(where), xp, fp)

    def setup(self, ntimestep, x=None, f=None, where=HOST):
        """setup(): post_force without the lambda step (see cph_setup)."""
        if where == HOST:
            x = _f64(x)
            if f is not None:
                assert f.dtype == np.float64 and f.flags.c_contiguous
            xp, fp = _d(x), _d(f)
        else:
            xp = None if x is None else C.cast(C.c_void_p(int(x)), _dp)
            fp = None if f is None else C.cast(C.c_void_p(int(f)), _dp)
        self._call("setup", C.c_int64(int(ntimestep)), C.c_int(where), xp, fp)

    def set_force_mode(self, accumulate):
        self._call("set_force_mode", C.c_int(1 if accumulate else 0))

    # -- results -----------------------------------------------------------------------------
    def _get_atoms(self, name, width):
        out = np.empty(self.nlocal * width, dtype=np.float64)
        self._call(name, C.c_int(HOST), _d(out))
        return out.reshape(self.nlocal, width) if width > 1 else out

    def get_forces(self):
        return self._get_atoms("get_forces", 3)

    def get_eatom(self):
        return self._get_atoms("get_eatom", 1)

    def get_phi(self):
        return self._get_atoms("get_phi", 1)

    def get_q(self):
        return self._get_atoms("get_q", 1)

    # -- f2: bonded terms and atom dynamics ----------------------------------------------
    def set_bonded(self, bond_k, bond_r0, angle_k, angle_theta0):
        bk, br, ak, at = _f64(bond_k), _f64(bond_r0), _f64(angle_k), _f64(angle_theta0)
        self._call("set_bonded", C.c_int(bk.size - 1), _d(bk), _d(br), C.c_int(ak.size - 1), _d(ak), _d(at))

    def set_topology(self, topo, rows=None):
        """`topo` is a synth.Topology; `rows` selects the owned atoms (same selection and order as set_atoms)."""
        pick = (lambda a: a) if rows is None else (lambda a: a[rows])
        arrs = [_i32(pick(a)) for a in (topo.num_bond, topo.bond_type, topo.bond_atom, topo.num_angle,
                                        topo.angle_type, topo.angle_atom1, topo.angle_atom2, topo.angle_atom3)]
        nb, bt, ba, na, at, a1, a2, a3 = arrs
        self._call("set_topology", C.c_int(nb.shape[0]), C.c_int(topo.maxbond), _i(nb), _i(bt), _i(ba),
                   C.c_int(topo.maxangle), _i(na), _i(at), _i(a1), _i(a2), _i(a3))

    def get_bonded_energy(self):
        out = np.zeros(2)
        self._call("get_bonded_energy", _d(out))
        return out

    def set_mass(self, mass):
        m = _f64(mass)
        self._call("set_mass", C.c_int(m.size - 1), _d(m))

    def set_v(self, v):
        v = _f64(v)
        self._call("set_v", C.c_int(HOST), _d(v))

    def md_initial_integrate(self, dt):
        self._call("md_initial_integrate", C.c_double(dt))

    def md_final_integrate(self, dt):
        self._call("md_final_integrate", C.c_double(dt))

    def get_x(self):
        return self._get_atoms("get_x", 3)

    def get_v(self):
        return self._get_atoms("get_v", 3)

    def get_scalars(self):
        out = np.zeros(8)
        self._call("get_scalars", _d(out))
        return dict(HA=out[0], HB=out[1], evdwl=out[2], ecoul=out[3], H_lambda=out[4], ke=out[5],
                    maxdisp2=out[6], thermostat=out[7])

    def get_sites(self):
        S = self.nsites
        names = ("lambda", "v_lambda", "dudl", "hdiff", "f_lambda", "f", "df", "U", "dU")
        arrs = [np.zeros(S) for _ in names]
        self._call("get_sites", *[_d(a) for a in arrs])
        return dict(zip(names, arrs))

    def compute_scalar(self):
        out = C.c_double(0)
        self._call("compute_scalar", C.byref(out))
        return out.value

    def compute_vector(self, i):
        out = C.c_double(0)
        self._call("compute_vector", C.c_int(i), C.byref(out))
        return out.value

    def memory_usage(self):
        out = C.c_double(0)
        self._call("memory_usage", C.byref(out))
        return out.value

    def get_counts(self):
        out = np.zeros(8, dtype=np.int64)
        self._call("get_counts", out.ctypes.data_as(_lp))
        return dict(nlocal=int(out[0]), nghost=int(out[1]), neighbors=int(out[2]), maxneigh=int(out[3]),
                    special_pairs=int(out[4]), builds=int(out[5]), titr_owned=int(out[6]), nsites=int(out[7]))

    def get_halo_mode(self):
        m = C.c_int(0)
        self._call("get_halo_mode", C.byref(m))
        return ("none", "nccl", "peer", "peer+mailbox")[m.value]

    def get_inner_counts(self):
        """(entries of the pruned inner rows, the same padded to 32 per row)."""
        out = np.zeros(2, dtype=np.int64)
        self._call("get_inner_counts", out.ctypes.data_as(_lp))
        return int(out[0]), int(out[1])

    def get_site_map(self):
        out = np.empty(self.nlocal, dtype=np.int32)
        self._call("get_site_map", _i(out))
        return out

    def get_neighbors(self):
        """(numneigh[nlocal], keys) with keys sorted within each atom's row."""
        num = np.zeros(self.nlocal, dtype=np.int32)
        self._call("get_neighbors", _i(num), None, C.c_int64(0))
        total = int(num.sum(dtype=np.int64))
        keys = np.zeros(max(total, 1), dtype=np.int64)
        self._call("get_neighbors", _i(num), keys.ctypes.data_as(_lp), C.c_int64(total))
        return num, keys[:total]

    # -- restart --------------------------------------------------------------------------------
    def pack_restart(self):
        n = C.c_int(0)
        self._call("restart_size", C.byref(n))
        buf = np.zeros(n.value)
        self._call("pack_restart", _d(buf))
        return buf

    def unpack_restart(self, buf):
        buf = _f64(buf)
        self._call("unpack_restart", _d(buf), C.c_int(int(buf.size)))

    # -- timing -----------------------------------------------------------------------------------
    def sync(self):
        self._call("sync")

    def stream(self):
        s = C.c_void_p()
        self._call("stream", C.byref(s))
        return s.value or 0

    def timer_start(self):
        self._call("timer_start")

    def timer_stop(self):
        ms = C.c_double(0)
        self._call("timer_stop", C.byref(ms))
        return ms.value

    def profile(self, enable=True):
        self._call("profile", C.c_int(1 if enable else 0))

    def profile_get(self, which):
        ms = C.c_double(0)
        n = C.c_int64(0)
        self._call("profile_get", C.c_int(which), C.byref(ms), C.byref(n))
        return ms.value, int(n.value)

    # not part of include/cph_b200.h: exported only by checker implementations (closed-form known answers)
    def bias_terms(self, lam):
        out = np.zeros(4)
        self._call("bias_terms", C.c_double(lam), _d(out))
        return dict(f=out[0], df=out[1], U=out[2], dU=out[3])


def bench_fp64_peak(device=0):
    """Measured sustained DFMA rate of the device: (warp-level DFMA instructions / s, TFLOP/s)."""
    lib = load_library("cph")
    a, b = C.c_double(0), C.c_double(0)
    rc = lib.cph_bench_fp64_peak(C.c_int(device), C.byref(a), C.byref(b))
    if rc != 0:
        raise CphError(rc, "cph_bench_fp64_peak failed")
    return a.value, b.value


def bench_seed_error(device=0):
    """Worst relative error of the 1/sqrt and 1/x seeds and of their refined forms (fastmath.cuh)."""
    lib = load_library("cph")
    out = np.zeros(4)
    rc = lib.cph_bench_seed_error(C.c_int(device), _d(out))
    if rc != 0:
        raise CphError(rc, "cph_bench_seed_error failed")
    return dict(rsqrt_seed=out[0], rcp_seed=out[1], rsqrt=out[2], rcp=out[3], refine_order=int(lib.cph_refine_order()))


def configure(eng, box, nevery=1, dudl=DUDL_CHARGE, integrator=INTEGRATE_REFERENCE,
              fscale=FSCALE_LAMBDA, bias_mode=BIAS_EXACT, implicit_site=False, ftm2v=None,
              sublo=None, subhi=None, procgrid=(1, 1, 1), myloc=(0, 0, 0), owned=None, bias=None,
              water_buffer=False, theta=False, cut_lj=None, cut_coul=None, thermostat=0.0,
              topology=None, velocities=None, drop_excluded=False, lj_typeB=None, kspace=None):
    """Push a synth.Box into an engine: the calls FixConstantPH's constructor/init/setup make.

    implicit_site=True reproduces the reference's single global lambda over the hydrogen
    group (nsites = 0, pK from the fix arguments, fix_constant_pH.cpp:47).
    owned: index array of the atoms this rank owns (None = all).
    bias: overrides of the init() constants (fix_constant_pH.cpp:86-96), e.g. m_lambda.
    topology: a synth.Topology -> bonded terms on; velocities: (n,3) -> fix-nve dynamics on.
    lj_typeB: B-state atom type per titratable atom (LJ end states, docs/SPEC.md), aligned with box.titr_tag.
    kspace: dict(g_ewald=..., kmax=(kx, ky, kz)) -> kspace_style ewald on; box.style should be PAIR_COUL_LONG with
    box.alpha = g_ewald (what pair lj/cut/coul/long's init_style takes from force->kspace)."""
    from . import synth
    eng.set_units(synth.QQRD2E, synth.BOLTZ, synth.FTM2V if ftm2v is None else ftm2v)
    # cut_lj: optional (ntypes+1)^2 table of per-type-pair LJ cutoffs (pair_coeff ... cut_lj); cut_coul: override
    eng.set_pair(box.style, box.ntypes, box.epsilon, box.sigma, cut_lj, box.cut_lj,
                 box.cut_coul if cut_coul is None else cut_coul, box.alpha, box.special_lj, box.special_coul)
    eng.set_domain(box.boxlo, box.boxhi, (1, 1, 1), sublo, subhi, procgrid, myloc, box.skin)
    if kspace:
        eng.set_kspace(KSPACE_EWALD, kspace["g_ewald"], kspace["kmax"])
    pK0 = float(box.pK[0]) if box.nsites else 0.0
    eng.set_fix(nevery, synth.GROUP_H_BIT, synth.GROUP_W_BIT, pK0, box.pH, box.T)
    eng.set_bias(bias_mode, **(bias or {}))
    eng.set_mode(dudl, integrator, fscale)
    if drop_excluded:
        eng.set_excluded_policy(True)
    if water_buffer:
        eng.set_water_buffer(True)
    if theta:
        eng.set_coordinate(True)
    if thermostat:
        eng.set_thermostat(thermostat)
    if implicit_site:
        eng.set_sites(0, None, None, None, None, None)
        eng.set_lambda(box.lambda0[:1], box.v0[:1])
    else:
        eng.set_sites(box.nsites, box.pK, box.titr_tag, box.titr_site, box.qA, box.qB)
        if lj_typeB is not None:
            eng.set_lj_states(lj_typeB)
        eng.set_lambda(box.lambda0, box.v0)
    sel = slice(None) if owned is None else owned
    eng.set_atoms(box.x[sel], box.q[sel], box.type[sel], box.tag[sel], box.mask[sel], box.molecule[sel],
                  box.nspecial[sel], box.special[sel], box.maxspecial)
    if topology is not None:
        eng.set_bonded(topology.bond_k, topology.bond_r0, topology.angle_k, topology.angle_theta0)
        eng.set_topology(topology, None if owned is None else owned)
        if velocities is not None:
            eng.set_mass(topology.mass)
            eng.set_v(velocities[sel])
    return eng
