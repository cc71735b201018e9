"""CPU: the C-ABI library loads, exports every symbol include/upol_b200.h declares, and the product
fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from u_pol_b200 import _lib as L

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "upol_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(upol_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(L.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/upol_b200.h but not exported"
    assert set(names) == set(L.EXPORTED_SYMBOLS), "ctypes binding and header disagree"


def test_constant_and_version():
    lib = L.lib()
    assert lib.upol_version() >= 100
    assert lib.upol_one_4pi_eps0() == 138.93545764438198      # polarization.py:19-25
    o = L.MinOptions()
    lib.upol_min_options_default(ctypes.byref(o))
    assert o.tol == 1e-4 and o.maxiter == 500                  # polarization.py:178, jaxopt default
    assert lib.upol_error_string(-2) == b"invalid argument"


def test_argument_errors_without_touching_the_device():
    lib = L.lib()
    h = ctypes.c_void_p()
    assert lib.upol_system_create(ctypes.byref(h), 0, 0, None, None, None, None, None, 0, None, None, None) == -2
    assert lib.upol_uind_energy_grad_dev(None, None, 0, None, None, None) == -2
    assert lib.upol_minimize_dev(None, None, None, None, None) == -2
    assert lib.upol_system_destroy(None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from u_pol_b200 import polarization as pol
    from u_pol_b200 import synth, util
    from u_pol_b200.engine import UindSystem

    assert L.lib().upol_device_count() == 0
    s = synth.water_box(2, seed=1)
    with pytest.raises(L.UpolError):
        UindSystem.from_drude_system(s)
    Rij, Dij = util.get_Rij_Dij(s)
    Qic, Qis, Qjc, Qjs = util.get_QiQj(s)
    k, u = util.get_pol_params(s)
    with pytest.raises(L.UpolError):
        pol.Uind(Rij, Dij, Qis, Qjs, Qic, Qjc, u, k)
    with pytest.raises(L.UpolError):
        pol.drudeOpt(Rij, np.zeros(Dij.size), Qis, Qjs, Qic, Qjc, u, k)
    assert pol.set_constants() == 138.93545764438198       # pure host call


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "u_pol_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "uind_oracle" not in src, f


def test_header_is_valid_c_and_plain_c_program_links(tmp_path):
    """include/upol_b200.h is a C header (not only C++): a C99 program that uses it compiles with
    -Wall -Werror, links against the shared library and runs (without a GPU it stops after the
    constant, as the library has no CPU fallback)."""
    import shutil
    import subprocess
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "c_api_example")
    libdir = os.path.dirname(L.LIB_PATH)
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_api_example.c"), "-L", libdir, "-lupol_b200",
           f"-Wl,-rpath,{libdir}", "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "ONE_4PI_EPS0 = 138.93545764438" in run.stdout
    if torch.cuda.is_available():
        assert "minimised U_ind = -" in run.stdout
