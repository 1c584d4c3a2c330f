"""bonded terms (bond_style harmonic, angle_style harmonic) joining the per-atom energy
the reference partitions (fix_constant_pH.cpp:221-229, 264-267), and fix-nve atom dynamics on the device.

CPU tests pin the oracle's restatement on what can be checked without LAMMPS: forces are the exact
gradient of the energies it reports, per-atom shares add up to the totals, the integrator conserves energy.
GPU tests compare the CUDA path with the oracle through the C ABI."""
import numpy as np
import pytest

from constant_ph_b200 import capi, synth

MASSIVE = dict(m_lambda=2000.0)


def oracle(box, topo, **kw):
    return capi.configure(capi.Engine("orc"), box, bias=MASSIVE, topology=topo, **kw)


def test_topology_generator_is_consistent():
    box = synth.config(1)
    topo = synth.topology(box)
    nb = np.arange(topo.maxbond)[None, :] < topo.num_bond[:, None]
    # every bond is listed with both atoms, every angle with all three
    assert topo.num_bond.sum() % 2 == 0 and topo.num_angle.sum() % 3 == 0
    assert topo.num_bond.sum() // 2 == 2 * box.meta["n_water"] + 7          # 2 per water + acetic acid
    assert topo.num_angle.sum() // 3 == box.meta["n_water"] + 10
    t2i = box.meta["tag_to_index"]
    i, m = np.nonzero(nb)
    j = t2i[topo.bond_atom[i, m]]
    back = (topo.bond_atom[j] == box.tag[i][:, None]) & (np.arange(topo.maxbond)[None, :] < topo.num_bond[j][:, None])
    assert back.any(axis=1).all()
    # the water model is SPC/Fw, solute terms start at their equilibrium values
    assert topo.bond_k[1] == pytest.approx(529.581) and topo.bond_r0[1] == pytest.approx(1.012)


def test_chain_box_topology_runs_along_the_backbone():
    """BASELINE config 4: poly(acrylic acid) repeat units bonded into chains.  8 bonds inside a unit plus one link
    to the next unit; special lists are symmetric and cross unit boundaries; every unit is one site of net charge
    0 (protonated) / -1 (deprotonated)."""
    box = synth.config(4, scale=0.008, chain_len=10)
    nm, cl = box.meta["n_paa"], box.meta["chain_len"]
    assert (nm, cl, box.nsites) == (400, 10, 400)
    topo = synth.topology(box)
    assert topo.num_bond.sum() // 2 == 2 * box.meta["n_water"] + 8 * nm + (nm // cl) * (cl - 1)
    t2i = box.meta["tag_to_index"]
    # special partners are mutual, class by class
    edges = [set(), set(), set()]
    for i in range(3 * box.meta["n_water"], box.n):
        lo = 0
        for c in range(3):
            for k in range(lo, box.nspecial[i, c]):
                edges[c].add((int(box.tag[i]), int(box.special[i, k])))
            lo = box.nspecial[i, c]
    for c in range(3):
        assert all((b, a) in edges[c] for a, b in edges[c])
    first = 3 * box.meta["n_water"]
    ch_tag, next_ch2 = int(box.tag[first + 3]), int(box.tag[first + 9])
    assert (ch_tag, next_ch2) in edges[0]                                     # the link bond
    assert box.maxspecial == 20 and box.nspecial[:, 2].max() == 20
    q0 = box.charges_at(np.zeros(box.nsites))[first:].reshape(nm, 9).sum(axis=1)
    q1 = box.charges_at(np.ones(box.nsites))[first:].reshape(nm, 9).sum(axis=1)
    assert np.allclose(q0, 0.0, atol=1e-12) and np.allclose(q1, -1.0, atol=1e-12)
    assert len(np.unique(box.molecule[first:])) == nm // cl                   # a chain is one molecule


def test_md_safe_start_has_no_interpenetrating_molecules():
    """The headline boxes let neighbouring solutes overlap (harmless for prescribed motion, an LJ-core explosion
    for an integrator: profiles/r1_scaling_and_bench.md).  md_safe keeps every pair of LJ-carrying atoms of
    different molecules apart."""
    from scipy.spatial import cKDTree
    worst = {}
    for safe in (False, True):
        # a box crowded with solutes, so that the default placement certainly puts some in adjacent rows
        box = synth.make_box("crowded", n_atoms=30000, n_acid=300, n_amine=300, style=synth.STYLE_COUL_DSF,
                             seed=9, jitter=0.1, md_safe=safe)
        L = box.boxhi - box.boxlo
        x = box.x - np.floor(box.x / L) * L
        x = np.minimum(x, np.nextafter(L, 0.0))
        heavy = np.nonzero(np.diag(box.epsilon)[box.type] > 0.05)[0]          # atoms with a real LJ core
        pairs = cKDTree(x[heavy], boxsize=L).query_pairs(2.6, output_type="ndarray")
        a, b = heavy[pairs[:, 0]], heavy[pairs[:, 1]]
        inter = box.molecule[a] != box.molecule[b]
        d = x[a[inter]] - x[b[inter]]
        d -= L * np.round(d / L)
        worst[safe] = np.linalg.norm(d, axis=1).min() if inter.any() else np.inf
    assert worst[False] < 1.6            # the default placement does produce overlapping solutes
    assert worst[True] > 2.2             # nothing closer than a hydrogen-bonded O...O contact


def test_oracle_bonded_forces_are_the_energy_gradient(built):
    box = synth.config(2, scale=0.1)
    topo = synth.topology(box)
    rng = np.random.default_rng(5)
    x0 = box.x + rng.normal(scale=0.03, size=box.x.shape)       # off equilibrium
    plain = capi.configure(capi.Engine("orc"), box, bias=MASSIVE)
    full = oracle(box, topo)
    f_pair, f_full = np.zeros((box.n, 3)), np.zeros((box.n, 3))
    plain.set_x(x0); plain.pair_pass(1)
    full.set_x(x0); full.pair_pass(1)
    f_pair[:], f_full[:] = plain.get_forces(), full.get_forces()
    fb = f_full - f_pair
    eb = full.get_bonded_energy()
    assert eb[0] > 0 and eb[1] > 0
    # per-atom shares add up: sum eatom = pair energy + bond + angle
    assert full.get_eatom().sum() == pytest.approx(plain.get_eatom().sum() + eb.sum(), rel=1e-12)
    assert np.abs(fb.sum(axis=0)).max() < 1e-9                 # Newton's third law
    # central differences of E_bond + E_angle on a few atoms of each kind
    probe = list(range(0, 9)) + list(range(box.n - 16, box.n))
    h = 1e-5
    for i in probe:
        for c in range(3):
            e = []
            for sgn in (+1, -1):
                x = x0.copy(); x[i, c] += sgn * h
                full.set_x(x); full.pair_pass(1)
                e.append(full.get_bonded_energy().sum())
            assert -(e[0] - e[1]) / (2 * h) == pytest.approx(fb[i, c], rel=2e-6, abs=2e-6)


def test_oracle_dynamics_conserve_energy(built):
    box = synth.config(2, scale=0.1, md_safe=True)
    topo = synth.topology(box)
    v0 = synth.thermal_velocities(box, topo, T=100.0)
    orc = capi.configure(capi.Engine("orc"), box, bias=dict(m_lambda=1e12), topology=topo, velocities=v0)
    dt = 0.25
    m = topo.mass[box.type][:, None]

    def total():
        s = orc.get_scalars()
        ke = 0.5 * (m * orc.get_v() ** 2).sum() / synth.FTM2V
        return s["evdwl"] + s["ecoul"] + orc.get_bonded_energy().sum() + ke, ke

    orc.post_force(0, dt)
    e0, ke0 = total()
    es = []
    for step in range(1, 161):
        orc.md_initial_integrate(dt)
        orc.post_force(step, dt)
        orc.md_final_integrate(dt)
        es.append(total()[0])
    es = np.array(es)
    # the lattice start is far from equilibrium: kinetic energy moves by far more than the total drifts
    assert abs(total()[1] - ke0) > 50 * np.abs(es - e0).max()
    assert np.abs(es - e0).max() < 2e-3 * abs(ke0)


def _pair(box, topo, **kw):
    cph = capi.configure(capi.Engine("cph"), box, bias=MASSIVE, topology=topo, **kw)
    orc = capi.configure(capi.Engine("orc"), box, bias=MASSIVE, topology=topo, **kw)
    return cph, orc


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["cfg1_coul_cut", "cfg2_dsf", "cfg2_shuffled"])
def test_bonded_pass_matches_oracle(built, cfg):
    box = {"cfg1_coul_cut": lambda: synth.config(1), "cfg2_dsf": lambda: synth.config(2, scale=0.25),
           "cfg2_shuffled": lambda: synth.config(2, scale=0.25, shuffle=True)}[cfg]()
    topo = synth.topology(box)
    rng = np.random.default_rng(11)
    x = box.x + rng.normal(scale=0.03, size=box.x.shape)
    cph, orc = _pair(box, topo)
    fc, fo = np.zeros((box.n, 3)), np.zeros((box.n, 3))
    cph.post_force(0, box.dt, x, fc)
    orc.post_force(0, box.dt, x, fo)
    assert _rel(fc, fo) <= 1e-10
    assert _rel(cph.get_eatom(), orc.get_eatom()) <= 1e-10
    assert _rel(cph.get_bonded_energy(), orc.get_bonded_energy()) <= 1e-12
    sc, so = cph.get_scalars(), orc.get_scalars()
    keys = ("HA", "HB", "evdwl", "ecoul")                      # HA, HB now carry the bonded energy
    assert _rel([sc[k] for k in keys], [so[k] for k in keys]) <= 1e-10
    assert _rel(cph.get_sites()["dudl"], orc.get_sites()["dudl"]) <= 1e-10
    # and they differ from a run without the topology by exactly E_bond + E_angle
    bare = capi.configure(capi.Engine("cph"), box, bias=MASSIVE)
    bare.post_force(0, box.dt, x, None)
    assert sc["HA"] - bare.get_scalars()["HA"] == pytest.approx(cph.get_bonded_energy().sum(), rel=1e-9)


@pytest.mark.gpu
def test_reference_mode_partition_with_bonded_energy(built):
    """cpp:264-267 with bond/angle eatom in H_atom: HB - HA drives the single reference lambda."""
    box = synth.config(1)
    topo = synth.topology(box)
    kw = dict(dudl=capi.DUDL_REFERENCE, implicit_site=True, nevery=2)
    cph, orc = _pair(box, topo, **kw)
    params = synth.jiggle_params(box)
    fc, fo = np.zeros((box.n, 3)), np.zeros((box.n, 3))
    for step in range(30):
        x = synth.jiggle_positions(box, params, step * box.dt)
        cph.post_force(step, box.dt, x, fc)
        orc.post_force(step, box.dt, x, fo)
        assert _rel(fc, fo) <= 1e-10
    assert abs(cph.get_sites()["lambda"][0] - orc.get_sites()["lambda"][0]) <= 1e-8
    assert _rel(cph.get_sites()["hdiff"], orc.get_sites()["hdiff"]) <= 1e-10


@pytest.mark.gpu
def test_device_dynamics_follow_the_oracle(built):
    """fix nve on the device + lambda dynamics, positions resident in HBM, list rebuilds inside the run."""
    box = synth.config(2, scale=0.25, md_safe=True)          # a start with no interpenetrating solutes
    topo = synth.topology(box)
    v0 = synth.thermal_velocities(box, topo, T=300.0)
    cph, orc = _pair(box, topo, velocities=v0)
    dt, L = 0.5, box.boxhi - box.boxlo
    cph.post_force(0, dt)
    orc.post_force(0, dt)
    nsteps = 240
    for step in range(1, nsteps + 1):
        for e in (cph, orc):
            e.md_initial_integrate(dt)
            e.post_force(step, dt)
            e.md_final_integrate(dt)
        if step % 60 == 0:
            d = cph.get_x() - orc.get_x()
            d -= L * np.round(d / L)                           # both remap into the box, at their own rebuilds
            assert np.abs(d).max() <= 1e-8
            assert np.abs(cph.get_v() - orc.get_v()).max() <= 1e-9
            assert np.abs(cph.get_sites()["lambda"] - orc.get_sites()["lambda"]).max() <= 1e-8
    assert cph.get_counts()["builds"] >= 3                            # the list was rebuilt along the way
    assert _rel(cph.get_forces(), orc.get_forces()) <= 1e-8
    assert _rel(cph.get_bonded_energy(), orc.get_bonded_energy()) <= 1e-9
    assert np.abs(cph.get_v()).max() < 0.5                     # A/fs: nothing blew up


@pytest.mark.gpu
def test_topology_errors(built):
    box = synth.config(1)
    topo = synth.topology(box)
    cph = capi.configure(capi.Engine("cph"), box, bias=MASSIVE)
    with pytest.raises(capi.CphError) as e:
        cph.set_topology(topo)                                  # coefficients first
    assert e.value.code == -2
    cph.set_bonded(topo.bond_k, topo.bond_r0, topo.angle_k, topo.angle_theta0)
    bad = synth.topology(box)
    bad.bond_atom[5, 0] = int(box.tag[-1])                      # a partner that is no special neighbour
    with pytest.raises(capi.CphError) as e:
        cph.set_topology(bad)
    assert "special" in str(e.value)
    with pytest.raises(capi.CphError):
        cph.set_v(np.zeros((box.n, 3)))                         # masses first
