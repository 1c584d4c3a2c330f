#!/usr/bin/env python
"""End-to-end time of one Verlet step THROUGH THE DROP-IN FIX (src/cph_harness: FixConstantPH::post_force with
LAMMPS-style pageable atom->x / atom->f, forces accumulated into atom->f) on the bench workload's box.
   python tools/harness_e2e.py [--atoms 1000000] [--steps 60] [--jiggle 0.45]
Prints one JSON line; compare with bench.py's e2e (page-locked buffers through the C ABI)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from constant_ph_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--atoms", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--jiggle", type=float, default=0.26,
                    help="per-dimension amplitude; 0.26 A per dimension = 0.45 A in total, the bench workload's amplitude")
    args = ap.parse_args()
    box = synth.config(3, scale=args.atoms / 1_000_000.0)
    with tempfile.TemporaryDirectory() as d:
        b, s = os.path.join(d, "box.bin"), os.path.join(d, "sites.txt")
        synth.write_harness_input(box, b, s)
        cmd = [os.path.join(ROOT, "src", "cph_harness"), b, str(args.steps), "jiggle", str(args.jiggle), "timing",
               "sites", s, "mlambda", "2000"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    if r.returncode != 0:
        print(json.dumps({"error": r.stderr[-500:]}))
        return 1
    ms = [float(l.split()[1]) for l in r.stdout.splitlines() if l.startswith("TIMING_MS_PER_STEP")][0]
    print(json.dumps({"metric": "fix_post_force_ms_per_step", "atoms": int(box.n), "steps": args.steps,
                      "ms_per_step": ms, "steps_per_s": 1e3 / ms, "omp_threads": os.environ.get("OMP_NUM_THREADS"),
                      "note": "FixConstantPH::post_force through src/cph_harness: pageable atom->x in, pair forces "
                              "ADDED to pageable atom->f, atoms moving (jiggle), list rebuilds and prunes included"}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
