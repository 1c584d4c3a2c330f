"""CPU tests of the oracle: closed-form known answers (the only values the reference pins,
SURVEY.md §4), internal consistency of the restated pair arithmetic, and the committed
golden fixtures.  No GPU."""
import json
import os

import numpy as np
import pytest

from constant_ph_b200 import capi, synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def orc(built):
    return capi.Engine("orc")


# ---- known answers from the reference's formulae (fix_constant_pH.cpp:86-94, 122, 132-136, 143) -----
def test_bias_potential_known_answers(orc):
    orc.set_bias(capi.BIAS_EXACT)
    for lam, want in [(0.0, -1.95926), (1.0, -1.95926), (0.5, 2.00000), (-0.12, 47.71667), (1.12, 47.71667)]:
        assert abs(orc.bias_terms(lam)["U"] - want) < 6e-6
    for lam, want in [(0.0, 10.12944), (0.02, 44.18287), (0.97, -47.87055), (1.05, 158.05091), (-0.05, -158.05091)]:
        assert abs(orc.bias_terms(lam)["dU"] - want) < 6e-6
    b = orc.bias_terms(0.5)
    assert b["f"] == 0.5 and abs(b["df"] - 12.5) < 1e-12
    assert abs(orc.bias_terms(0.3)["df"] - 2.27e-3) < 1e-5


def test_exact_derivatives_match_finite_differences(orc):
    orc.set_bias(capi.BIAS_EXACT)
    for lam in np.linspace(-0.1, 1.1, 25):
        h = 1e-6
        up, dn, mid = orc.bias_terms(lam + h), orc.bias_terms(lam - h), orc.bias_terms(lam)
        assert abs((up["U"] - dn["U"]) / (2 * h) - mid["dU"]) < 2e-5 * max(1.0, abs(mid["dU"]))
        assert abs((up["f"] - dn["f"]) / (2 * h) - mid["df"]) < 1e-6 * max(1.0, abs(mid["df"]))


def test_as_written_mode_reproduces_the_reference_lines(orc):
    """cpp:123 and cpp:137-141 verbatim, including their arithmetic slips."""
    orc.set_bias(capi.BIAS_AS_WRITTEN)
    b = orc.bias_terms(0.5)
    assert abs(b["df"] - 200.0) < 1e-9                       # 50*exp(0)/(0.5^2)
    lam = -0.05
    p = capi.BIAS_DEFAULT
    U1 = -p["k"] * np.exp(-(lam - 1 - p["b"]) ** 2 / (2 * p["a"] ** 2))
    U2 = -p["k"] * np.exp(-(lam + p["b"]) ** 2 / (2 * p["a"] ** 2))
    U3 = p["d"] * np.exp(-(lam - 0.5) ** 2 / (2 * p["s"] ** 2))
    dU = (-((lam - 1 - p["b"]) / (2 * p["a"] ** 2)) * U1 - ((lam + p["b"]) / (2 * p["a"] ** 2)) * U2
          - ((lam - 0.5) / p["s"] ** 2) * U3
          - 0.5 * p["w"] * p["r"] * 2 * np.exp(-p["r"] ** 2 * (lam + 0.5) ** 2) / np.sqrt(np.pi)
          + 0.5 * p["w"] * p["r"] * 2 * np.exp(-p["r"] ** 2 * (lam - 1 - p["m"]) ** 2) / np.sqrt(np.pi))
    assert abs(orc.bias_terms(lam)["dU"] - dU) < 1e-10
    assert abs(dU - (-18.33)) < 0.01                         # the value SURVEY.md D15 measured
    orc.set_bias(capi.BIAS_EXACT)


def test_integrator_step_is_cpp_111_116():
    """One reference step with a hand-computed force: lambda += v t + a t^2/2, v += a t."""
    box = synth.config(1)
    o = capi.configure(capi.Engine("orc"), box, nevery=3, ftm2v=1.0)
    o.set_lambda(np.array([0.4]), np.array([0.01]))
    o.apply_charges()
    o.pair_pass(1); o.site_reduce()
    dudl = o.get_sites()["dudl"][0]
    b = o.bias_terms(0.4)
    ph = synth.BOLTZ * box.T * np.log(10.0) * (box.pK[0] - box.pH)
    F = -(dudl + b["df"] * ph + b["dU"])
    a = F / 20.0
    t = 3 * box.dt
    o.integrate_lambda(t)
    s = o.get_sites()
    assert abs(s["f_lambda"][0] - F) < 1e-12 * abs(F)
    assert abs(s["lambda"][0] - (0.4 + 0.01 * t + 0.5 * a * t * t)) < 1e-12
    assert abs(s["v_lambda"][0] - (0.01 + a * t)) < 1e-12


def test_nevery_gate_and_argument_errors():
    o = capi.Engine("orc")
    with pytest.raises(capi.CphError):
        o.set_fix(0, 2, 4, 4.76, 4.8, 300.0)
    box = synth.config(1)
    o = capi.configure(capi.Engine("orc"), box, nevery=4)
    lam = []
    for step in range(9):
        o.post_force(step, box.dt, box.x, None)
        lam.append(o.get_sites()["lambda"][0])
    lam = np.array(lam)
    changed = np.nonzero(np.diff(np.concatenate([[box.lambda0[0]], lam])) != 0)[0]
    assert set(changed.tolist()) <= {0, 4, 8} and 4 in changed      # cpp:69


# ---- internal consistency of the restated pair arithmetic ----------------------------------------------
@pytest.mark.parametrize("cfg,scale", [(1, 1.0), (2, 0.2), (4, 0.008)])
def test_energy_partition_and_newton(cfg, scale):
    box = synth.config(cfg, scale=scale, **({"chain_len": 10} if cfg == 4 else {}))
    o = capi.configure(capi.Engine("orc"), box)
    o.pair_pass(1); o.site_reduce()
    s = o.get_scalars()
    e = o.get_eatom()
    f = o.get_forces()
    assert abs(e.sum() - (s["evdwl"] + s["ecoul"])) < 1e-10 * abs(s["evdwl"] + s["ecoul"])
    assert abs(s["HA"] - e.sum()) < 1e-11 * abs(s["HA"])
    H = (box.mask & synth.GROUP_H_BIT) != 0
    assert abs((s["HA"] - s["HB"]) - e[H].sum()) < 1e-9 * max(1.0, abs(e[H].sum()))      # cpp:264-267
    assert np.abs(f.sum(axis=0)).max() < 1e-9 * np.abs(f).max()
    # E_coul = 1/2 sum q_i phi_i  (phi = dE/dq including the dsf self term)
    q, phi = o.get_q(), o.get_phi()
    assert abs(0.5 * (q * phi).sum() - s["ecoul"]) < 1e-10 * abs(s["ecoul"])


@pytest.mark.parametrize("cfg,scale", [(1, 1.0), (2, 0.2), (4, 0.008)])
def test_analytic_dudl_equals_perturbed_charge_reevaluation(cfg, scale):
    """north_star: dU/dlambda from re-evaluating the pair energy at lambda +- dlambda.  E is
    quadratic in each lambda_s, so the central difference equals the analytic value.  Config 4: bonded chains,
    neighbouring sites are 1-3 / 1-4 partners of each other."""
    box = synth.config(cfg, scale=scale, **({"chain_len": 10} if cfg == 4 else {}))
    o = capi.configure(capi.Engine("orc"), box)
    o.pair_pass(1); o.site_reduce()
    dudl = o.get_sites()["dudl"].copy()
    for site in range(min(box.nsites, 3)):
        es = []
        for sign in (+1, -1):
            lam = box.lambda0.copy(); lam[site] += sign * 0.05
            o.set_lambda(lam); o.apply_charges(); o.pair_pass(1); o.site_reduce()
            sc = o.get_scalars(); es.append(sc["evdwl"] + sc["ecoul"])
        fd = (es[0] - es[1]) / 0.1
        assert abs(fd - dudl[site]) < 1e-8 * max(1.0, abs(dudl[site]))


def test_host_kspace_terms_enter_the_site_sums_once():
    """cpp:241-244 seen by the charge derivative (cph_set_extra_dudl): a host-tallied dE/dlambda_s is added to the
    next site reduce and then forgotten; a wrong site count is refused."""
    box = synth.config(2, scale=0.2)
    o = capi.configure(capi.Engine("orc"), box)
    o.pair_pass(1); o.site_reduce()
    base = o.get_sites()["dudl"].copy()
    extra = 0.3 * (1 + np.arange(box.nsites))
    o.set_extra_dudl(extra)
    o.site_reduce()
    assert np.allclose(o.get_sites()["dudl"], base + extra, rtol=0, atol=1e-12)
    o.site_reduce()
    assert np.array_equal(o.get_sites()["dudl"], base)
    with pytest.raises(capi.CphError):
        o.set_extra_dudl(extra[:-1])


def test_lj_end_states_limits_and_derivatives():
    """docs/SPEC.md, LJ end states.  lambda = 0: the plain run; lambda = 1: the run with the B types written into
    atom->type; in between: dU/dlambda equals the central difference of the total energy (E is a polynomial of
    degree <= 2 in every lambda_s) and the force on an end-state atom equals -dE/dx."""
    box = synth.config(2, scale=0.2)
    typeB = synth.lj_end_state_types(box)
    assert (typeB > 0).sum() >= 2 * box.nsites - 2
    pos = box.meta["tag_to_index"][box.titr_tag]

    def energy(o):
        o.pair_pass(1); o.site_reduce()
        sc = o.get_scalars()
        return sc["evdwl"] + sc["ecoul"], sc

    # end points
    for lam_all, swapped in ((0.0, False), (1.0, True)):
        lam = np.full(box.nsites, lam_all)
        plain_box = synth.config(2, scale=0.2)
        if swapped:
            plain_box.type[pos[typeB > 0]] = typeB[typeB > 0]
        plain = capi.configure(capi.Engine("orc"), plain_box)
        plain.set_lambda(lam); plain.apply_charges()
        mixed = capi.configure(capi.Engine("orc"), box, lj_typeB=typeB)
        mixed.set_lambda(lam); mixed.apply_charges()
        (ep, sp), (em, sm) = energy(plain), energy(mixed)
        assert abs(ep - em) <= 1e-12 * abs(ep)
        assert np.abs(plain.get_forces() - mixed.get_forces()).max() <= 1e-11 * np.abs(plain.get_forces()).max()
    # derivative in lambda
    o = capi.configure(capi.Engine("orc"), box, lj_typeB=typeB)
    e0, _ = energy(o)
    dudl = o.get_sites()["dudl"].copy()
    plain = capi.configure(capi.Engine("orc"), box)
    _ = energy(plain)
    assert np.abs(dudl - plain.get_sites()["dudl"]).max() > 1e-3          # the LJ term is there
    for site in range(min(box.nsites, 4)):
        es = []
        for sign in (+1, -1):
            lam = box.lambda0.copy(); lam[site] += sign * 0.05
            o.set_lambda(lam); o.apply_charges()
            es.append(energy(o)[0])
        fd = (es[0] - es[1]) / 0.1
        assert abs(fd - dudl[site]) < 1e-8 * max(1.0, abs(dudl[site]))
    # force on an end-state atom
    o.set_lambda(box.lambda0); o.apply_charges()
    energy(o)
    f = o.get_forces()
    i = int(pos[np.nonzero(typeB > 0)[0][0]])
    h = 1e-4
    for k in range(3):
        es = []
        for sign in (+1, -1):
            x = box.x.copy(); x[i, k] += sign * h
            o.set_x(x)
            es.append(energy(o)[0])
        assert abs(-(es[0] - es[1]) / (2 * h) - f[i, k]) < 2e-5 * max(1.0, abs(f[i, k]))


def test_neighbor_list_is_symmetric_and_complete():
    box = synth.config(1)
    o = capi.configure(capi.Engine("orc"), box)
    num, keys = o.get_neighbors()
    assert num.sum() == keys.size == o.get_counts()["neighbors"]
    # brute force count of pairs within cutoff+skin under minimum image (special pairs with zero weights dropped)
    L = box.boxhi - box.boxlo
    sub = np.arange(0, box.n, 37)
    d = box.x[sub][:, None, :] - box.x[None, :, :]
    d -= L * np.round(d / L)
    within = (d ** 2).sum(-1) < (box.cut_coul + box.skin) ** 2
    brute = within.sum(1) - 1
    nsp = np.array([box.nspecial[i, 1] for i in sub])      # 1-2 and 1-3 partners have weight 0 -> dropped
    assert np.array_equal(num[sub], brute - nsp)


def test_restart_roundtrip_and_site_map():
    box = synth.config(2, scale=0.2)
    a = capi.configure(capi.Engine("orc"), box)
    sm = a.get_site_map()
    want = np.full(box.n, -1, dtype=np.int32)
    want[box.meta["tag_to_index"][box.titr_tag]] = box.titr_site
    assert np.array_equal(sm, want)
    for step in range(5):
        a.post_force(step, box.dt, box.x, None)
    buf = a.pack_restart()
    assert buf.size == 2 + 3 * box.nsites
    b = capi.configure(capi.Engine("orc"), box)
    b.unpack_restart(buf)
    assert np.array_equal(a.get_sites()["lambda"], b.get_sites()["lambda"])
    assert np.array_equal(a.get_q(), b.get_q())


# ---- committed golden fixtures (tests/golden/make_golden.py) -----------------------------------------------
def test_oracle_matches_golden_fixtures():
    g = json.load(open(os.path.join(GOLDEN, "oracle_golden.json")))
    for case in g["cases"]:
        box = synth.config(case["config"], scale=case["scale"])
        topo = synth.topology(box) if case.get("bonded") else None
        o = capi.configure(capi.Engine("orc"), box, topology=topo, **{k: v for k, v in case["kw"].items()})
        for step in range(case["steps"]):
            o.post_force(step, box.dt, box.x, None)
        s, t = o.get_scalars(), o.get_sites()
        if topo is not None:
            assert np.allclose(o.get_bonded_energy(), case["bonded_energy"], rtol=1e-12, atol=0)
        for k, v in case["scalars"].items():
            assert abs(s[k] - v) <= 1e-11 * max(1.0, abs(v)), (case["name"], k, s[k], v)
        assert np.allclose(t["lambda"], case["lambda"], rtol=0, atol=1e-12)
        assert np.allclose(t["dudl"], case["dudl"], rtol=1e-11, atol=1e-11)
        f = o.get_forces()
        assert abs(np.abs(f).sum() - case["f_abs_sum"]) <= 1e-11 * case["f_abs_sum"]
        c = o.get_counts()
        assert c["neighbors"] == case["neighbors"] and c["special_pairs"] == case["special_pairs"]


def test_water_buffer_keeps_total_charge_and_dudl_consistent():
    """modify_water (h:58, TODO at cpp:268): with the buffer on, the box charge is independent of
    lambda and the analytic dU/dlambda (including the buffer term) equals the re-evaluated energy."""
    box = synth.config(2, scale=0.2)
    o = capi.configure(capi.Engine("orc"), box, water_buffer=True)
    totals = []
    for v in (0.0, 0.3, 1.0):
        o.set_lambda(np.full(box.nsites, v)); o.apply_charges()
        totals.append(o.get_q().sum())
    assert max(totals) - min(totals) < 1e-12
    W = (box.mask & synth.GROUP_W_BIT) != 0
    assert W.sum() == 3 and not np.allclose(o.get_q()[W], box.q[W])
    o.set_lambda(box.lambda0); o.apply_charges(); o.pair_pass(1); o.site_reduce()
    dudl = o.get_sites()["dudl"].copy()
    for site in range(3):
        es = []
        for sign in (+1, -1):
            lam = box.lambda0.copy(); lam[site] += sign * 0.05
            o.set_lambda(lam); o.apply_charges(); o.pair_pass(1); o.site_reduce()
            sc = o.get_scalars(); es.append(sc["evdwl"] + sc["ecoul"])
        assert abs((es[0] - es[1]) / 0.1 - dudl[site]) < 1e-8 * max(1.0, abs(dudl[site]))


def test_theta_coordinate_keeps_lambda_in_range_and_conserves_energy():
    """north_star's lambda/theta variables: with lambda = sin^2(theta) the site coordinate cannot leave
    [0, 1]; velocity-Verlet on theta conserves the extended energy with frozen atoms."""
    box = synth.config(2, scale=0.2)
    o = capi.configure(capi.Engine("orc"), box, theta=True, integrator=capi.INTEGRATE_VV, bias=dict(m_lambda=2000.0))
    o.post_force(0, box.dt, box.x, None); o.final_integrate(0.0)
    H, lam = [], []
    for step in range(1, 300):
        o.initial_integrate(box.dt); o.post_force(step, box.dt, box.x, None); o.final_integrate(box.dt)
        H.append(o.compute_scalar()); lam.append(o.get_sites()["lambda"].copy())
    lam, H = np.array(lam), np.array(H)
    assert lam.min() >= 0.0 and lam.max() <= 1.0 and lam.max() - lam.min() > 0.5
    assert np.abs(H - H[0]).max() < 1e-2 * max(1.0, o.get_scalars()["ke"])
    # restart carries theta (version 2) and refuses the other coordinate
    buf = o.pack_restart()
    assert buf[0] == 2.0
    p = capi.configure(capi.Engine("orc"), box, bias=dict(m_lambda=2000.0))
    with pytest.raises(capi.CphError):
        p.unpack_restart(buf)


def test_nose_hoover_lambda_thermostat():
    """f3: Nose-Hoover on the site velocities (velocity-Verlet form).  The site kinetic energy settles at
    kT/2 per site and H_lambda + thermostat energy is conserved with frozen atoms."""
    box = synth.config(2, scale=0.2)
    o = capi.configure(capi.Engine("orc"), box, integrator=capi.INTEGRATE_VV, bias=dict(m_lambda=2000.0),
                       thermostat=50.0, theta=True)
    o.post_force(0, box.dt, box.x, None); o.final_integrate(0.0)
    H, K = [], []
    for step in range(1, 2500):
        o.initial_integrate(box.dt); o.post_force(step, box.dt, box.x, None); o.final_integrate(box.dt)
        s = o.get_scalars(); H.append(s["H_lambda"] + s["thermostat"]); K.append(s["ke"])
    H, K = np.array(H), np.array(K)
    kT = synth.BOLTZ * box.T
    assert abs(K[800:].mean() / box.nsites - 0.5 * kT) < 0.15 * 0.5 * kT
    assert np.abs(H - H[0]).max() < 5e-3
    buf = o.pack_restart()
    assert buf.size == 5 + 3 * box.nsites and buf[-3] != 0.0
