"""Regenerate tests/golden/oracle_golden.json.

The reference ships no golden vectors and cannot be run, so these
fixtures are OUTPUTS OF THE ORACLE on seeded synthetic boxes, frozen so that a change to
the oracle (or to the box generator) cannot go unnoticed.  They do not pin the oracle to
the reference; the closed-form known answers in tests/test_oracle.py do that for the only
part the reference defines (bias potential, switch function, integrator step).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from constant_ph_b200 import capi, synth  # noqa: E402
import oracle.binding  # noqa: E402,F401  -- registers capi.Engine("orc")

CASES = [
    dict(name="cfg1_coul_cut_charge", config=1, scale=1.0, steps=3, kw=dict(bias=dict(m_lambda=2000.0))),
    dict(name="cfg1_reference_mode", config=1, scale=1.0, steps=3,
         kw=dict(dudl=capi.DUDL_REFERENCE, implicit_site=True, bias=dict(m_lambda=2000.0))),
    dict(name="cfg2_dsf_charge", config=2, scale=0.2, steps=3, kw=dict(bias=dict(m_lambda=2000.0))),
    # SPC/Fw bonds and angles join the forces and the partitioned per-atom energy
    dict(name="cfg1_reference_mode_bonded", config=1, scale=1.0, steps=3, bonded=True,
         kw=dict(dudl=capi.DUDL_REFERENCE, implicit_site=True, bias=dict(m_lambda=2000.0))),
    dict(name="cfg2_dsf_charge_bonded", config=2, scale=0.2, steps=3, bonded=True, kw=dict(bias=dict(m_lambda=2000.0))),
]


# The configuration every BENCH / SCALE number is quoted on (BASELINE configs[2], full size), on the first frames of
# bench.py's own trajectory; values after step 0 and after step 3.  bench.py checks its engines against this file
# before the timed loop at every rank count, tests/test_gpu_parity.py does the same under pytest.
CFG3_STEPS = (0, 3)
CFG3_KW = dict(bias=dict(m_lambda=2000.0))


def cfg3_workload(scale=1.0):
    """Box + prescribed-motion parameters of bench.py's workload (same call, same seeds)."""
    box = synth.config(3, scale=scale)
    params = synth.jiggle_params(box, amp=0.45, period_lo=60.0, period_hi=140.0)
    return box, params


def snapshot(eng):
    s, t, c = eng.get_scalars(), eng.get_sites(), eng.get_counts()
    f = eng.get_forces()
    return {"scalars": {k: float(s[k]) for k in ("HA", "HB", "evdwl", "ecoul", "H_lambda")},
            "dudl": [float(v) for v in t["dudl"]], "lambda": [float(v) for v in t["lambda"]],
            "f_abs_sum": float(np.abs(f).sum()), "f_max": float(np.abs(f).max()),
            "neighbors": c["neighbors"], "special_pairs": c["special_pairs"], "nlocal": c["nlocal"]}


def make_cfg3_full(path, scale=1.0):
    box, params = cfg3_workload(scale)
    o = capi.Engine("orc")
    o.lib.orc_set_threads(os.cpu_count() or 1)
    capi.configure(o, box, **CFG3_KW)
    rec = {"generated_by": "tests/golden/make_golden.py (oracle output; the reference holds no vectors)",
           "config": 3, "scale": scale, "atoms": int(box.n), "sites": int(box.nsites),
           "kw": CFG3_KW, "jiggle": dict(amp=0.45, period_lo=60.0, period_hi=140.0), "steps": {}}
    for step in range(max(CFG3_STEPS) + 1):
        o.post_force(step, box.dt, synth.jiggle_positions(box, params, step * box.dt), None)
        if step in CFG3_STEPS:
            rec["steps"][str(step)] = snapshot(o)
    with open(path, "w") as fh:
        json.dump(rec, fh)
    print("wrote", path)


# The other bench workloads (bench.py --config 4 / 5): the same record for the box each rank count runs (config 4 is
# weak-scaled: one box per rank count).  Site values are SAMPLED (every k-th site, at most ~2048) to keep the files
# small; the sums over all sites are stored beside them.
def make_bench_golden(path, config, nranks):
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    args = argparse.Namespace(config=config, atoms=1_000_000, pH=7.0)
    box, params = bench.workload(args, nranks)
    o = capi.Engine("orc")
    o.lib.orc_set_threads(os.cpu_count() or 1)
    capi.configure(o, box, **CFG3_KW)
    stride = max(1, -(-box.nsites // 2048))
    idx = list(range(0, box.nsites, stride))
    rec = {"generated_by": "tests/golden/make_golden.py --bench %d %d (oracle output)" % (config, nranks),
           "config": config, "nranks": nranks, "atoms": int(box.n), "sites": int(box.nsites), "kw": CFG3_KW,
           "site_index": idx, "steps": {}}
    for step in range(max(CFG3_STEPS) + 1):
        o.post_force(step, box.dt, synth.jiggle_positions(box, params, step * box.dt), None)
        if step in CFG3_STEPS:
            snap = snapshot(o)
            d, lam = np.array(snap["dudl"]), np.array(snap["lambda"])
            snap["dudl_sum"], snap["dudl_sq_sum"] = float(d.sum()), float((d * d).sum())
            snap["dudl"], snap["lambda"] = [float(v) for v in d[idx]], [float(v) for v in lam[idx]]
            rec["steps"][str(step)] = snap
    with open(path, "w") as fh:
        json.dump(rec, fh)
    print("wrote", path, box.n, box.nsites)


# lj/cut/coul/long + kspace_style ewald on configs 1 and 2 (oracle output after 3 steps of lambda
# dynamics with frozen atoms), so the oracle's k-space part cannot drift either.  tests/test_kspace.py reads it.
EWALD_CASES = [
    dict(name="cfg1_ewald", config=1, scale=1.0, steps=3, g_ewald=0.30, kmax=[7, 7, 7]),
    dict(name="cfg2_ewald", config=2, scale=0.25, steps=3, g_ewald=0.30, kmax=[9, 8, 10]),
]


def ewald_case_engine(case, prefix="orc", **engine_kw):
    import dataclasses
    box = synth.config(case["config"], scale=case["scale"])
    box = dataclasses.replace(box, style=capi.PAIR_COUL_LONG, alpha=case["g_ewald"])
    eng = capi.configure(capi.Engine(prefix, **engine_kw), box, bias=dict(m_lambda=2000.0),
                         kspace=dict(g_ewald=case["g_ewald"], kmax=case["kmax"]))
    for step in range(case["steps"]):
        eng.post_force(step, box.dt, box.x, None)
    return box, eng


def make_ewald_golden(path):
    out = {"generated_by": "tests/golden/make_golden.py --ewald (oracle output; the reference holds no vectors)", "cases": []}
    for case in EWALD_CASES:
        box, o = ewald_case_engine(case)
        rec = dict(case)
        rec.update(snapshot(o))
        rec["e_kspace"] = o.get_kspace_energy()
        out["cases"].append(rec)
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path)


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    if "--ewald" in sys.argv:
        make_ewald_golden(os.path.join(here, "ewald_golden.json"))
        return
    if "--bench" in sys.argv:
        k = sys.argv.index("--bench")
        config, nranks = int(sys.argv[k + 1]), int(sys.argv[k + 2])
        make_bench_golden(os.path.join(here, "bench_cfg%d_n%d_golden.json" % (config, nranks)), config, nranks)
        return
    if "--cfg3" in sys.argv or "--all" in sys.argv:
        make_cfg3_full(os.path.join(here, "cfg3_full_golden.json"))
        if "--cfg3" in sys.argv:
            return
    out = {"generated_by": "tests/golden/make_golden.py", "cases": []}
    for case in CASES:
        box = synth.config(case["config"], scale=case["scale"])
        topo = synth.topology(box) if case.get("bonded") else None
        o = capi.configure(capi.Engine("orc"), box, topology=topo, **case["kw"])
        for step in range(case["steps"]):
            o.post_force(step, box.dt, box.x, None)
        s, t, c = o.get_scalars(), o.get_sites(), o.get_counts()
        rec = dict(case)
        rec["scalars"] = {k: float(s[k]) for k in ("HA", "HB", "evdwl", "ecoul", "H_lambda")}
        rec["lambda"] = [float(v) for v in t["lambda"]]
        rec["dudl"] = [float(v) for v in t["dudl"]]
        rec["f_abs_sum"] = float(np.abs(o.get_forces()).sum())
        if topo is not None:
            rec["bonded_energy"] = [float(v) for v in o.get_bonded_energy()]
        rec["neighbors"] = c["neighbors"]
        rec["special_pairs"] = c["special_pairs"]
        out["cases"].append(rec)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote %d cases" % len(out["cases"]))


if __name__ == "__main__":
    main()
