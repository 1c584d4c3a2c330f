#!/usr/bin/env python
"""Time the device Ewald pass (cph_profile slot 10: structure factors + per-atom sums) on BASELINE configs 1 and 2 under
lj/cut/coul/long, for the default kernels (table structure factors + row-walking per-atom sums), the table kernels
(CPH_EWALD=tables) and the direct ones (CPH_EWALD=direct), and cross-check them against each other.  Prints one JSON line.  Needs a B200:  python tools/ewald_timing.py > profiles/<name>.json"""
import dataclasses
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from constant_ph_b200 import capi, synth  # noqa: E402


def run(box, kmax, mode, passes=10):
    if mode in ("direct", "tables"):
        os.environ["CPH_EWALD"] = mode
    else:
        os.environ.pop("CPH_EWALD", None)
    eng = capi.configure(capi.Engine("cph", device=0), box, kspace=dict(g_ewald=box.alpha, kmax=kmax))
    for _ in range(3):
        eng.pair_pass(1)
    eng.sync()
    eng.profile(True)
    for _ in range(passes):
        eng.pair_pass(1)
    eng.sync()
    ms_s, n_k = eng.profile_get(10)      # structure factors
    ms_a, _ = eng.profile_get(11)        # per-atom sums
    ms_k = ms_s + ms_a
    ms_p, n_p = eng.profile_get(0)
    eng.profile(False)
    eng.site_reduce()
    out = dict(f=eng.get_forces().copy(), phi=eng.get_phi().copy(), ek=eng.get_kspace_energy(),
               kspace_ms=ms_k / max(n_k, 1), sfac_ms=ms_s / max(n_k, 1), atom_ms=ms_a / max(n_k, 1),
               pair_ms=ms_p / max(n_p, 1))
    eng.close()
    return out


def main():
    if "--profile-run" in sys.argv:     # a short run of the default kernels on config 2 for an ncu launch list
        box = dataclasses.replace(synth.config(2), style=capi.PAIR_COUL_LONG, alpha=0.30)
        r = run(box, (22, 22, 22), "rows", passes=3)
        print(json.dumps(dict(kspace_ms=r["kspace_ms"])))
        return
    if "--tune" in sys.argv:            # launch-shape sweep (CPH_EWALD_TUNE) on config 2
        box = dataclasses.replace(synth.config(2), style=capi.PAIR_COUL_LONG, alpha=0.30)
        out = {"rows": {}, "sfac": {}}
        for blk in (64, 128, 256):
            for bps in (4, 6, 8, 12):
                os.environ["CPH_EWALD_TUNE"] = "%d,%d,256,32,4" % (blk, bps)
                out["rows"]["%d,%d" % (blk, bps)] = round(run(box, (22, 22, 22), "rows", passes=4)["atom_ms"], 4)
        for blk in (128, 256):
            for tile in (16, 32):
                for bps in (4, 6, 8):
                    os.environ["CPH_EWALD_TUNE"] = "128,4,%d,%d,%d" % (blk, tile, bps)
                    out["sfac"]["%d,%d,%d" % (blk, tile, bps)] = round(run(box, (22, 22, 22), "rows", passes=4)["sfac_ms"], 4)
        os.environ.pop("CPH_EWALD_TUNE", None)
        print(json.dumps(out))
        return
    res = {}
    peak_warp_dfma, _ = capi.bench_fp64_peak(0)
    for name, cfg, kmax in (("config1_3k_atoms", 1, (7, 7, 7)), ("config2_32k_atoms", 2, (22, 22, 22))):
        box = synth.config(cfg)
        box = dataclasses.replace(box, style=capi.PAIR_COUL_LONG, alpha=0.30)
        a = run(box, kmax, "rows")           # default: table structure factors + row-walking per-atom kernel
        t = run(box, kmax, "tables")         # table kernels for both
        b = run(box, kmax, "direct")
        L = box.boxhi - box.boxlo
        unitk = 2 * np.pi / L
        gsq = max((unitk * np.array(kmax)) ** 2) * 1.00001
        g = np.stack(np.meshgrid(np.arange(0, kmax[0] + 1), np.arange(-kmax[1], kmax[1] + 1),
                                 np.arange(-kmax[2], kmax[2] + 1), indexing="ij"), -1).reshape(-1, 3)
        half = (g[:, 0] > 0) | ((g[:, 0] == 0) & (g[:, 1] > 0)) | ((g[:, 0] == 0) & (g[:, 1] == 0) & (g[:, 2] > 0))
        K = int((half & (((g * unitk) ** 2).sum(1) <= gsq)).sum())
        pairs = 2.0 * box.n * K             # (atom, wave vector) evaluations of both kernels of a pass
        res[name] = dict(
            atoms=box.n, kmax=list(kmax), wave_vectors=K,
            default_ms_per_pass=a["kspace_ms"], default_structure_factors_ms=a["sfac_ms"], default_atom_sums_ms=a["atom_ms"], tables_ms_per_pass=t["kspace_ms"], direct_ms_per_pass=b["kspace_ms"],
            pair_pass_ms=a["pair_ms"], speedup_over_direct=b["kspace_ms"] / a["kspace_ms"],
            tables_force_rel_diff=float(np.abs(t["f"] - b["f"]).max() / np.abs(b["f"]).max()),
            atom_wavevector_evaluations_per_s=pairs / (a["kspace_ms"] * 1e-3),
            # fp64 instructions per evaluation: ~10 in the structure-factor loop, ~17 in the row-walking loop
            fp64_frac_of_measured_peak=13.5 * pairs / 32.0 / (a["kspace_ms"] * 1e-3) / peak_warp_dfma,
            force_rel_diff_between_variants=float(np.abs(a["f"] - b["f"]).max() / np.abs(b["f"]).max()),
            phi_rel_diff_between_variants=float(np.abs(a["phi"] - b["phi"]).max() / np.abs(b["phi"]).max()),
            e_kspace=a["ek"], e_kspace_rel_diff=abs(a["ek"] - b["ek"]) / abs(b["ek"]))
    res["fp64_peak_warp_dfma_per_s"] = peak_warp_dfma
    print(json.dumps(res))


if __name__ == "__main__":
    main()
