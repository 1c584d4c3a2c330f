"""the k-space source of compute_Hs (fix_constant_pH.cpp:241-244) on the device.

`kspace_style ewald` (reciprocal sum) + pair `lj/cut/coul/long` (its real-space part).  The reference holds no
vectors for it; the oracle is pinned by what an Ewald sum must reproduce whatever the code: the Madelung constant
of rock salt, independence of the splitting parameter, forces = -grad E, and the charge derivative equal to the
lambda +- dlambda re-evaluation north_star describes.  The GPU tests compare the CUDA path with that oracle."""
import dataclasses

import numpy as np
import pytest

from constant_ph_b200 import capi, synth

MADELUNG_NACL = 1.7475645946331822        # nearest-neighbour convention


def rock_salt(cells=4, a=2.82, g_ewald=0.36, cut=10.5):
    """cells^3 conventional cells of NaCl (8 ions each), charges +-1, no LJ, no titration sites."""
    pts = np.array([[i, j, k] for i in range(2 * cells) for j in range(2 * cells) for k in range(2 * cells)])
    n = pts.shape[0]
    q = np.where(pts.sum(axis=1) % 2 == 0, 1.0, -1.0)
    L = 2 * cells * a
    z1 = np.zeros(1)
    return synth.Box(name="nacl", boxlo=np.zeros(3), boxhi=np.full(3, L), x=pts * a + 0.25 * a, q=q,
                     type=np.where(q > 0, 1, 2).astype(np.int32), tag=np.arange(1, n + 1, dtype=np.int32),
                     mask=np.ones(n, dtype=np.int32), molecule=np.zeros(n, dtype=np.int32),
                     nspecial=np.zeros((n, 3), dtype=np.int32), special=np.zeros((n, 1), dtype=np.int32), maxspecial=0,
                     ntypes=2, epsilon=np.zeros((3, 3)), sigma=np.ones((3, 3)), style=capi.PAIR_COUL_LONG,
                     cut_lj=cut, cut_coul=cut, alpha=g_ewald, special_lj=np.array([1.0, 0, 0, 0]),
                     special_coul=np.array([1.0, 0, 0, 0]), skin=0.7, nsites=0, pK=z1, titr_tag=np.zeros(0, np.int32),
                     titr_site=np.zeros(0, np.int32), qA=np.zeros(0), qB=np.zeros(0), lambda0=np.full(1, 0.5), v0=z1)


def ewald_box(cfg=1, scale=1.0, g_ewald=0.30):
    """A BASELINE box with its pair style switched to lj/cut/coul/long (alpha = g_ewald)."""
    box = synth.config(cfg, scale=scale)
    return dataclasses.replace(box, style=capi.PAIR_COUL_LONG, alpha=g_ewald)


def total_energy(eng):
    eng.pair_pass(1); eng.site_reduce()
    s = eng.get_scalars()
    return s["evdwl"] + s["ecoul"]


# ---- oracle: what any Ewald sum must reproduce (CPU) ------------------------------------------------------------
def test_oracle_ewald_reproduces_the_madelung_constant_of_rock_salt(built):
    a = 2.82
    box = rock_salt(a=a)
    o = capi.configure(capi.Engine("orc"), box, implicit_site=True, kspace=dict(g_ewald=box.alpha, kmax=(11, 11, 11)))
    e = total_energy(o)
    per_ion = e / box.n
    exact = -0.5 * MADELUNG_NACL * synth.QQRD2E / a       # every ion carries half of its pair energies
    assert abs(per_ion - exact) <= 2e-6 * abs(exact)
    assert np.abs(o.get_forces()).max() <= 1e-6            # a perfect lattice is force-free
    # per-atom view: phi_i = dE/dq_i is the Madelung potential, the same on every site up to sign
    phi = o.get_phi()
    assert np.allclose(phi * box.q, 2 * exact, rtol=2e-6)
    assert abs(0.5 * (box.q * phi).sum() - e) <= 1e-10 * abs(e)


def test_oracle_ewald_total_does_not_depend_on_the_splitting_parameter(built):
    tot = []
    for g in (0.36, 0.40):
        box = ewald_box(1, g_ewald=g)
        o = capi.configure(capi.Engine("orc"), box, kspace=dict(g_ewald=g, kmax=(16, 16, 16)))
        o.pair_pass(1); o.site_reduce()
        s = o.get_scalars()
        tot.append((s["ecoul"], o.get_kspace_energy()))
    # real + reciprocal + self: the sum stays put (to the truncation errors of the two parts) ...
    parts = abs(tot[0][1]) + abs(tot[0][0] - tot[0][1])
    assert abs(tot[0][0] - tot[1][0]) <= 2e-6 * parts
    assert abs(tot[0][1] - tot[1][1]) > 1e-2 * parts                        # ... while the split itself moved


def test_oracle_ewald_forces_are_the_energy_gradient_and_dudl_the_lambda_derivative(built):
    box = ewald_box(1)
    kw = dict(kspace=dict(g_ewald=box.alpha, kmax=(7, 7, 7)))
    o = capi.configure(capi.Engine("orc"), box, **kw)
    e0 = total_energy(o)
    f = o.get_forces().copy()
    dudl = o.get_sites()["dudl"].copy()
    assert np.abs(f.sum(axis=0)).max() <= 1e-8 * np.abs(f).max()            # translation invariance
    # central differences on a titratable atom and on a water atom
    row = {int(t): i for i, t in enumerate(box.tag)}
    h = 1e-4
    for i in (row[int(box.titr_tag[0])], 17):
        for d in range(3):
            es = []
            for sign in (+1, -1):
                x = box.x.copy(); x[i, d] += sign * h
                o.set_x(x)
                es.append(total_energy(o))
            # 5e-5: LAMMPS' polynomial erfc (1.5e-7 absolute) is differentiated exactly in the force, so the pair
            # force is not the exact gradient of the pair energy -- the same offset shows with the k-space part off
            assert abs(-(es[0] - es[1]) / (2 * h) - f[i, d]) <= 5e-5 * np.abs(f[i]).max()
    o.set_x(box.x)
    # north_star: dU/dlambda from the energies at lambda +- dlambda; E is quadratic in lambda -> exact
    es = []
    for sign in (+1, -1):
        lam = box.lambda0.copy(); lam[0] += sign * 0.05
        o.set_lambda(lam); o.apply_charges()
        es.append(total_energy(o))
    assert abs((es[0] - es[1]) / 0.1 - dudl[0]) <= 1e-8 * max(1.0, abs(dudl[0]))
    assert abs(e0) > 0


def test_oracle_ewald_is_invariant_under_translation_and_atom_order(built):
    """Moving every atom by the same vector (periodic box) or handing the atoms over in another order must not change
    the energy, and forces must follow their atoms -- for the pair part and the reciprocal sum together."""
    box = ewald_box(1)
    kw = dict(kspace=dict(g_ewald=box.alpha, kmax=(6, 6, 6)))
    o = capi.configure(capi.Engine("orc"), box, **kw)
    e0 = total_energy(o)
    f0 = o.get_forces().copy()
    d0 = o.get_sites()["dudl"].copy()
    # rigid shift, atoms wrapped back into the box
    L = box.boxhi - box.boxlo
    shifted = dataclasses.replace(box, x=box.boxlo + np.mod(box.x + np.array([3.7, -11.2, 0.9]) - box.boxlo, L))
    o1 = capi.configure(capi.Engine("orc"), shifted, **kw)
    assert abs(total_energy(o1) - e0) <= 1e-9 * max(1.0, abs(e0))
    assert np.abs(o1.get_forces() - f0).max() <= 1e-9 * np.abs(f0).max()
    assert np.allclose(o1.get_sites()["dudl"], d0, rtol=1e-9, atol=1e-9)
    # another atom order (tags travel with the atoms, so sites and special lists still resolve)
    perm = np.random.default_rng(5).permutation(box.n)
    mixed = dataclasses.replace(box, x=box.x[perm], q=box.q[perm], type=box.type[perm], tag=box.tag[perm],
                                mask=box.mask[perm], molecule=box.molecule[perm], nspecial=box.nspecial[perm],
                                special=box.special[perm])
    o2 = capi.configure(capi.Engine("orc"), mixed, **kw)
    assert abs(total_energy(o2) - e0) <= 1e-10 * max(1.0, abs(e0))
    assert np.abs(o2.get_forces() - f0[perm]).max() <= 1e-10 * np.abs(f0).max()
    assert np.allclose(o2.get_sites()["dudl"], d0, rtol=1e-10, atol=1e-10)


def test_kspace_argument_errors(built):
    box = ewald_box(1)
    o = capi.Engine("orc")
    with pytest.raises(capi.CphError):
        o.set_kspace(capi.KSPACE_EWALD, 0.3, (5, 5, 5))                      # before set_domain
    o = capi.configure(capi.Engine("orc"), box)
    with pytest.raises(capi.CphError):
        o.set_kspace(capi.KSPACE_EWALD, -1.0, (5, 5, 5))
    with pytest.raises(capi.CphError):
        o.set_kspace(7, 0.3, (5, 5, 5))


# ---- CUDA path against the oracle ------------------------------------------------------------------------------
def engines(box, **kw):
    gpu = capi.configure(capi.Engine("cph", device=0), box, **kw)
    orc = capi.configure(capi.Engine("orc"), box, **kw)
    return gpu, orc


def close(a, b, rtol=1e-10):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    s = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max()
    assert err <= rtol * s, "max abs err %.3e vs scale %.3e (rel %.3e)" % (err, s, err / s)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,scale,kmax", [(1, 1.0, (7, 7, 7)), (2, 0.25, (9, 8, 10))])
def test_ewald_pass_matches_oracle(built, cfg, scale, kmax):
    """Forces, per-atom energy, phi_i, energies, HA/HB and every site's dU/dlambda with the reciprocal sum on (1e-10)."""
    box = ewald_box(cfg, scale)
    gpu, orc = engines(box, kspace=dict(g_ewald=box.alpha, kmax=kmax))
    for eng in (gpu, orc):
        eng.pair_pass(1); eng.site_reduce()
    close(gpu.get_forces(), orc.get_forces())
    close(gpu.get_eatom(), orc.get_eatom())
    close(gpu.get_phi(), orc.get_phi())
    sg, so = gpu.get_scalars(), orc.get_scalars()
    for k in ("HA", "HB", "evdwl", "ecoul"):
        assert abs(sg[k] - so[k]) <= 1e-10 * abs(so[k]), (k, sg[k], so[k])
    assert abs(gpu.get_kspace_energy() - orc.get_kspace_energy()) <= 1e-10 * abs(so["ecoul"])
    close(gpu.get_sites()["dudl"], orc.get_sites()["dudl"])


@pytest.mark.gpu
def test_ewald_madelung_constant_on_the_device(built):
    a = 2.82
    box = rock_salt(a=a)
    gpu = capi.configure(capi.Engine("cph", device=0), box, implicit_site=True,
                         kspace=dict(g_ewald=box.alpha, kmax=(11, 11, 11)))
    exact = -0.5 * MADELUNG_NACL * synth.QQRD2E / a
    assert abs(total_energy(gpu) / box.n - exact) <= 2e-6 * abs(exact)
    assert np.abs(gpu.get_forces()).max() <= 1e-6


@pytest.mark.gpu
def test_ewald_lambda_trajectory_with_moving_atoms(built):
    """200 steps of lambda dynamics under lj/cut/coul/long + ewald with atoms moving (list rebuilds, prunes)."""
    box = ewald_box(2, 0.25)
    params = synth.jiggle_params(box, amp=0.9, period_lo=40.0, period_hi=90.0)
    gpu, orc = engines(box, bias=dict(m_lambda=2000.0), kspace=dict(g_ewald=box.alpha, kmax=(8, 8, 8)))
    f = np.zeros((box.n, 3))
    lam = {id(gpu): [], id(orc): []}
    for step in range(200):
        x = synth.jiggle_positions(box, params, step * box.dt)
        for eng in (gpu, orc):
            eng.post_force(step, box.dt, x, f)
            lam[id(eng)].append(eng.get_sites()["lambda"].copy())
    lg, lo = np.array(lam[id(gpu)]), np.array(lam[id(orc)])
    assert np.abs(lg - lo).max() <= 1e-8
    assert np.abs(lo[-1] - lo[0]).max() > 1e-4
    assert gpu.get_counts()["builds"] == orc.get_counts()["builds"] > 1
    close(gpu.get_forces(), orc.get_forces(), rtol=1e-9)


@pytest.mark.gpu
def test_ewald_argument_errors_on_the_device(built):
    box = ewald_box(1)
    g = capi.Engine("cph", device=0)
    with pytest.raises(capi.CphError):
        g.set_kspace(capi.KSPACE_EWALD, 0.3, (5, 5, 5))                      # before set_domain
    g = capi.configure(capi.Engine("cph", device=0), box)
    with pytest.raises(capi.CphError):
        g.set_kspace(capi.KSPACE_EWALD, -1.0, (5, 5, 5))
    with pytest.raises(capi.CphError):
        g.set_kspace(capi.KSPACE_EWALD, 0.3, (0, 5, 5))


# ---- the drop-in fix: keyword `ewald KX KY KZ` -------------------------------------------------------------------
@pytest.fixture(scope="module")
def ewald_files(built, tmp_path_factory):
    d = tmp_path_factory.mktemp("harness_ewald")
    box = ewald_box(1)
    b, s = str(d / "box.bin"), str(d / "sites.txt")
    synth.write_harness_input(box, b, s)
    return box, b, s


def test_fix_ewald_keyword_errors(ewald_files):
    """Constructor-level checks run before any device work (like cpp:36-54)."""
    from test_fix_dropin import run
    _, b, s = ewald_files
    r = run([b, 1, "sites", s, "ewald", 0, 5, 5])
    assert r.returncode == 2 and "Illegal fix constant_pH ewald value 0" in r.stderr
    r = run([b, 1, "sites", s, "ewald", 5, 5])
    assert r.returncode == 2 and "missing argument" in r.stderr


@pytest.mark.gpu
def test_fix_with_device_ewald_matches_oracle(ewald_files):
    """pair lj/cut/coul/long + `kspace_modify compute no` + fix keyword ewald: the fix takes g_ewald from force->kspace
    and runs the reciprocal sum in the library; lambda and H_lambda trajectories against the oracle."""
    from test_fix_dropin import run, parse, oracle_trajectory
    box, b, s = ewald_files
    nsteps = 60
    r = run([b, nsteps, "jiggle", 0.5, "sites", s, "mlambda", 2000, "ewald", 7, 6, 7])
    assert r.returncode == 0, r.stderr
    rows, extra = parse(r.stdout)
    kw = dict(bias=dict(m_lambda=2000.0), kspace=dict(g_ewald=box.alpha, kmax=(7, 6, 7)))
    lam, H, f = oracle_trajectory(box, "charge", kw, nsteps, jiggle=0.5)
    assert np.abs(rows[:, 2:] - lam).max() <= 1e-8
    assert np.abs(rows[:, 1] - H).max() <= 1e-8 * np.abs(H).max()
    assert abs(extra["FORCES_ABS_SUM"] - np.abs(f).sum()) <= 1e-9 * np.abs(f).sum()
    # the same input without the keyword is refused only if the host would ALSO compute the sum; with
    # `kspace_modify compute no` and no keyword nothing computes it -- the pair part alone must differ
    r0 = run([b, 2, "sites", s, "mlambda", 2000])
    assert r0.returncode == 0, r0.stderr
    rows0, _ = parse(r0.stdout)
    assert abs(rows0[0, 1] - rows[0, 1]) > 1e-6 * abs(rows[0, 1])


# ---- frozen oracle outputs (tests/golden/ewald_golden.json, written by tests/golden/make_golden.py --ewald) -------
def _golden_cases():
    import json
    import os
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import make_golden
    return json.load(open(os.path.join(here, "ewald_golden.json")))["cases"], make_golden.ewald_case_engine


def _check_against_golden(eng, case, rtol):
    s, t, c = eng.get_scalars(), eng.get_sites(), eng.get_counts()
    # E_coul is the small difference of large parts (pair + reciprocal + self): the k-space energy sets the scale
    slack = 1e-2 * rtol * abs(case["e_kspace"])
    for k, v in case["scalars"].items():
        assert abs(s[k] - v) <= rtol * abs(v) + slack, (case["name"], k, s[k], v)
    assert abs(eng.get_kspace_energy() - case["e_kspace"]) <= rtol * abs(case["e_kspace"])
    assert np.abs(t["lambda"] - np.array(case["lambda"])).max() <= 1e-10
    assert np.allclose(t["dudl"], case["dudl"], rtol=rtol, atol=slack)
    f = eng.get_forces()
    assert abs(np.abs(f).sum() - case["f_abs_sum"]) <= rtol * case["f_abs_sum"]
    assert c["neighbors"] == case["neighbors"] and c["special_pairs"] == case["special_pairs"]


def test_oracle_ewald_matches_golden_fixture(built):
    cases, make = _golden_cases()
    for case in cases:
        _, o = make(case)
        _check_against_golden(o, case, 1e-11)


@pytest.mark.gpu
def test_cuda_ewald_matches_golden_fixture(built):
    """The CUDA path alone against the committed values: no oracle in the loop."""
    cases, make = _golden_cases()
    for case in cases:
        _, g = make(case, prefix="cph", device=0)
        _check_against_golden(g, case, 1e-10)
