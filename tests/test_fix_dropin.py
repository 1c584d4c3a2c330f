"""The drop-in FixConstantPH (src/fix_constant_pH.{h,cpp}) compiled against the LAMMPS shim and
driven in Verlet order by src/cph_harness.  CPU tests cover what the reference's constructor
does before any device work (cpp:36-54); GPU tests compare trajectories with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from constant_ph_b200 import capi, synth

HARNESS = os.path.join(ROOT, "src", "cph_harness")


@pytest.fixture(scope="module")
def box_files(built, tmp_path_factory):
    d = tmp_path_factory.mktemp("harness")
    box = synth.config(1)
    b, s = str(d / "box.bin"), str(d / "sites.txt")
    synth.write_harness_input(box, b, s)
    return box, b, s


def run(args, **kw):
    return subprocess.run([HARNESS] + [str(a) for a in args], capture_output=True, text=True, timeout=600, **kw)


def test_fix_builds_against_the_shim(built):
    assert os.path.exists(HARNESS)
    # the fix source names only upstream LAMMPS headers and the C ABI
    src = open(os.path.join(ROOT, "src", "fix_constant_pH.cpp")).read()
    incs = [l.split('"')[1] for l in src.splitlines() if l.startswith('#include "')]
    assert set(incs) <= {"fix_constant_pH.h", "angle.h", "atom.h", "bond.h", "comm.h", "dihedral.h", "domain.h",
                         "error.h", "force.h", "group.h", "improper.h", "kspace.h", "memory.h", "neighbor.h",
                         "pair.h", "update.h", "cph_b200.h"}
    assert "lammps_shim" not in src


def test_constructor_argument_errors(box_files):
    """cpp:36-54 with the defects the survey lists resolved (D4: nevery <= 0, D6: unknown keyword)."""
    _, b, s = box_files
    r = run([b, 1, "nevery", 0])
    assert r.returncode == 2 and "Illegal fix constant pH every value 0" in r.stderr
    r = run([b, 1, "nosuchkeyword", 1])
    assert r.returncode == 2 and "Unknown fix constant_pH keyword" in r.stderr
    r = run([b, 1, "dudl"])
    assert r.returncode == 2 and "missing argument" in r.stderr
    r = run([b, 1, "sites", "/nonexistent/file"])
    assert r.returncode == 2 and "Cannot open" in r.stderr


def test_no_gpu_is_a_loud_error(box_files):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _, b, s = box_files
    r = run([b, 1, "sites", s])
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


def parse(out):
    rows, extra = [], {}
    for line in out.splitlines():
        t = line.split()
        if not t:
            continue
        if t[0].isupper():
            extra[t[0]] = float(t[1])
        else:
            rows.append([float(v) for v in t])
    return np.array(rows), extra


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["charge", "reference", "vv", "buffer", "theta", "bonded", "thermostat", "biasconst"])
def test_fix_trajectory_matches_oracle(box_files, mode):
    box, b, s = box_files
    nsteps = 120
    if mode == "charge":
        args = ["sites", s, "mlambda", 2000]
        kw = dict(bias=dict(m_lambda=2000.0))
    elif mode == "biasconst":
        # settable bias constants (the reference hard-codes Donnini's table in init(), cpp:86-94)
        args = ["sites", s, "mlambda", 2000, "bias_h", 2.5, "bias_w", 150, "bias_d", 1.5]
        kw = dict(bias=dict(m_lambda=2000.0, hbar=2.5, w=150.0, d=1.5))
    elif mode == "thermostat":
        args = ["sites", s, "mlambda", 2000, "integrator", "vv", "coordinate", "theta", "tlambda", 40]
        kw = dict(bias=dict(m_lambda=2000.0), integrator=capi.INTEGRATE_VV, theta=True, thermostat=40.0)
    elif mode == "bonded":
        # reference mode + a host-side bonded energy source folded into HA/HB (cpp:221-253)
        args = ["nevery", 2, "bonded", 0.01, "mlambda", 2000, "lambda0", 0.5]
        kw = dict(bias=dict(m_lambda=2000.0), dudl=capi.DUDL_REFERENCE, implicit_site=True, nevery=2)
    elif mode == "theta":
        args = ["sites", s, "mlambda", 2000, "coordinate", "theta"]
        kw = dict(bias=dict(m_lambda=2000.0), theta=True)
    elif mode == "buffer":
        args = ["sites", s, "mlambda", 2000, "buffer", "yes"]
        kw = dict(bias=dict(m_lambda=2000.0), water_buffer=True)
    elif mode == "vv":
        args = ["sites", s, "mlambda", 2000, "integrator", "vv"]
        kw = dict(bias=dict(m_lambda=2000.0), integrator=capi.INTEGRATE_VV)
    else:
        args = ["nevery", 3, "mlambda", 2000, "fscale", "oneminus", "lambda0", 0.5]
        kw = dict(bias=dict(m_lambda=2000.0), dudl=capi.DUDL_REFERENCE, implicit_site=True, nevery=3,
                  fscale=capi.FSCALE_ONE_MINUS)
    r = run([b, nsteps] + args)
    assert r.returncode == 0, r.stderr
    rows, extra = parse(r.stdout)
    assert rows.shape[0] == nsteps + 1

    orc = capi.configure(capi.Engine("orc"), box, **kw)
    lam, H = [], []
    f = np.zeros((box.n, 3))
    nev = kw.get("nevery", 1)

    def extras(step):
        if mode != "bonded" or step % nev:
            return
        e = 0.01 * (1 + np.arange(box.n) % 7)
        Hm = (box.mask & synth.GROUP_H_BIT) != 0
        orc.set_extra_partition(float(e.sum()), float(e[~Hm].sum()))

    extras(0)
    orc.post_force(0, box.dt, box.x, f)                       # setup()
    lam.append(orc.get_sites()["lambda"].copy()); H.append(orc.compute_scalar())
    for step in range(1, nsteps + 1):
        if mode in ("vv", "thermostat"):
            orc.initial_integrate(box.dt * nev)
        extras(step)
        orc.post_force(step, box.dt, box.x, f)
        if mode in ("vv", "thermostat"):
            orc.final_integrate(box.dt * nev)
        lam.append(orc.get_sites()["lambda"].copy()); H.append(orc.compute_scalar())
    lam, H = np.array(lam), np.array(H)
    assert np.abs(rows[:, 2:] - lam).max() <= 1e-8
    assert np.abs(rows[:, 1] - H).max() <= 1e-8 * np.abs(H).max()
    assert abs(extra["FORCES_ABS_SUM"] - np.abs(f).sum()) <= 1e-9 * np.abs(f).sum()
    assert extra["RESTART_BYTES"] == 8 * (2 + 3 * max(1, lam.shape[1]) + (3 if mode == "thermostat" else 0))
    assert extra["MEMORY_USAGE"] > 1e6
