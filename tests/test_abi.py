"""The C-ABI shared library loads and exports every symbol include/cph_b200.h declares.
No compute calls: this runs without a GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cph_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cph_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hook_surface():
    syms = declared_symbols()
    for name in ("cph_create", "cph_destroy", "cph_set_fix", "cph_set_atoms", "cph_post_force", "cph_pair_pass",
                 "cph_site_reduce", "cph_integrate_lambda", "cph_initial_integrate", "cph_final_integrate",
                 "cph_compute_scalar", "cph_compute_vector", "cph_memory_usage", "cph_pack_restart",
                 "cph_unpack_restart", "cph_comm_init_nccl"):
        assert name in syms


def test_library_exports_every_declared_symbol(built):
    from constant_ph_b200 import capi
    lib = ctypes.CDLL(capi.CUDA_LIB)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.cph_version() >= 100


def test_no_cpu_fallback(built):
    """Without a CUDA device the product refuses to run; it never routes to the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from constant_ph_b200 import capi
    with pytest.raises(capi.CphError) as e:
        capi.Engine("cph")
    assert e.value.code == -3
    assert "no CPU fallback" in str(e.value)


def test_product_sources_never_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under constant_ph_b200/csrc or src/ may name it."""
    bad = []
    for sub in ("constant_ph_b200/csrc", "src"):
        d = os.path.join(ROOT, sub)
        if not os.path.isdir(d):
            continue
        for fn in os.listdir(d):
            if fn.endswith((".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(d, fn)).read()
                if "orc_" in txt or "cph_oracle" in txt:
                    bad.append(fn)
    # the Python side of the package: no import of, path into, or symbol of the checker either
    pkg = os.path.join(ROOT, "constant_ph_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            if "import oracle" in txt or "oracle/" in txt or "libcph_oracle" in txt or "orc_" in txt or '"oracle"' in txt:
                bad.append(fn)
    assert not bad, bad


def test_only_the_cuda_engine_ships_with_the_package():
    """capi.Engine knows one library; a second engine exists only after test infrastructure registers it."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from constant_ph_b200 import capi\n"
            "try:\n"
            "    capi.Engine('orc')\n"
            "except capi.CphError as e:\n"
            "    print('refused', e.code)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.stdout.strip() == "refused -1", (r.stdout, r.stderr[-500:])
