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


def main():
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
