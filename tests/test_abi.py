"""The C-ABI library loads on a CPU-only box and exports every symbol include/cfd_b200.h declares.
No compute call is made here (there is no GPU); the product has no CPU fallback, which is also checked."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry
from cfd_demo_b200 import _abi, model
from cfd_demo_b200.types import SimulationParams, default_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(model.LIB_PATH):
        entry.build()
    return model.load_library()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "cfd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cfd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_reference_api():
    names = declared_functions()
    for must in ["cfd_model_create", "cfd_model_update", "cfd_model_set_params", "cfd_model_get_snapshot",
                 "cfd_model_get_residuals", "cfd_model_destroy", "cfd_last_error"]:
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.cfd_abi_version() == _abi.CFD_ABI_VERSION


def test_struct_layouts_match_the_header(lib):
    # sizes follow from the header's member lists (natural alignment)
    assert C.sizeof(_abi.CfdGrid) == 48
    assert C.sizeof(_abi.CfdParams) == 28
    assert C.sizeof(_abi.CfdSolverConsts) == 80  # ABI 3: + cg_relative, adaptive_substeps
    assert C.sizeof(_abi.CfdOptions) == 32 + 80
    assert C.sizeof(_abi.CfdResiduals) == 8 + 5 * 4 + 4 + 8 + 3 * 8 + 5 * 8 + 3 * 8  # ABI 3: + p_rel, rhs_rms, first_solve_iterations
    o = model.default_options()
    assert (o.precision, o.device, o.rank, o.world_size) == (64, -1, 0, 1)
    c = o.consts
    assert (c.ramp_up_steps, c.jacobi_iterations, c.outer_rounds) == (100, 50, 20)
    assert (c.jacobi_omega, c.pressure_tolerance, c.outer_tolerance, c.cfl) == (0.75, 1e-4, 1e-4, 0.2)
    assert (c.cg_tolerance, c.mg_omega, c.mg_smoothing, c.mg_warm_start) == (1e-8, 0.8, 2, 3)
    assert (c.cg_relative, c.adaptive_substeps) == (0, 0)  # the reference's behaviour is the default


def test_argument_validation_needs_no_gpu(lib):
    from cfd_demo_b200.types import Grid
    with pytest.raises(model.CfdError) as e:
        model.Model(Grid.uniform(20, 8, 1.0, 1.0), SimulationParams())  # nx % 8 != 0: the reference panics
    assert e.value.code == _abi.CFD_ERR_INVALID_ARGUMENT
    with pytest.raises(model.CfdError):
        model.Model(Grid.uniform(16, 2, 1.0, 1.0), SimulationParams())
    with pytest.raises(model.CfdError) as e:
        model.Model(Grid.uniform(65536, 65536, 1.0, 1.0), SimulationParams())  # 2^32 cells: 32-bit field indices
    assert e.value.code == _abi.CFD_ERR_INVALID_ARGUMENT and "too large" in str(e.value)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(model.CfdError) as e:
        model.Model(default_grid(), SimulationParams())
    assert e.value.code == _abi.CFD_ERR_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cfd_demo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".rs")):
                text = open(os.path.join(dirpath, f)).read()
                assert "cfd_oracle" not in text and "cpu_oracle" not in text and "import oracle" not in text, f


def test_cpp_host_mirror_builds_and_fails_loudly_without_gpu(lib):
    """host/cfd_model.hpp (C++ mirror of the reference API) links against the C ABI; without a GPU the headless
    driver must exit non-zero with the library's error (no CPU fallback)."""
    import subprocess
    import torch
    exe = os.path.join(ROOT, "cfd_demo_b200", "host", "cfd_headless")
    assert os.path.exists(exe)
    r = subprocess.run([exe, "2"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "step 2" in r.stdout
    else:
        assert r.returncode != 0 and "cfd_b200" in r.stderr


@pytest.mark.parametrize("ny,world", [(96, 2), (41, 2), (264, 4), (4096, 8), (32768, 8), (16384, 3), (600, 2), (100, 5)])
def test_strip_partition_is_contiguous_and_pairs_up_on_four_multigrid_levels(lib, ny, world):
    """Host logic of the N > 1 path (no GPU needed): the strips tile [0, ny) without gaps, rank 0 / the last rank own
    the ring rows, and every interior boundary sits at array row 1 + 16k — so the unknown rows (array row - 1) of a strip
    are closed under the pairing (2J, 2J+1) -> J on multigrid levels 0 to 3 (restriction and prolongation of the
    MGCG strips then need no communication)."""
    rows = [model.strip_rows(ny, world, r) for r in range(world)]
    assert rows[0][0] == 0 and rows[-1][1] == ny
    for (a0, a1), (b0, b1) in zip(rows, rows[1:]):
        assert a1 == b0 and a1 > a0
    sizes = [b - a for a, b in rows]
    assert min(sizes) >= 16 and max(sizes) - min(sizes) <= 32 + 2
    for a, _ in rows[1:]:
        assert (a - 1) % 16 == 0
    # closure: for every level l <= 3, each level-(l+1) row has all its children in one strip
    owner = np.zeros(ny - 2, dtype=int)
    for r, (a, b) in enumerate(rows):
        owner[max(a, 1) - 1:min(b, ny - 1) - 1] = r
    cur = owner
    for _ in range(4):
        n = len(cur)
        pairs = cur[:n - n % 2].reshape(-1, 2)
        assert (pairs[:, 0] == pairs[:, 1]).all()
        cur = np.concatenate([pairs[:, 0], cur[n - n % 2:]])


def test_python_constants_match_the_header():
    """Field ids, flags and enum values of cfd_demo_b200/_abi.py are the header's #defines (the ctypes mirror is written
    by hand: a drifted constant would read the wrong field without any error)."""
    import re
    from cfd_demo_b200 import _abi
    text = open(os.path.join(ROOT, "include", "cfd_b200.h")).read()
    defines = {m.group(1): int(m.group(2).rstrip("u"), 0) for m in re.finditer(r"#define\s+(CFD_\w+)\s+(\d+u?)\b", text)}
    checked = 0
    for name, value in vars(_abi).items():
        if not name.isupper() or not isinstance(value, int):
            continue
        for key in ("CFD_" + name, "CFD_" + name.replace("SOLVER_", "SOLVER_").replace("SCENARIO_", "SCENARIO_")):
            if key in defines:
                assert defines[key] == value, (name, value, defines[key])
                checked += 1
                break
    assert checked >= 20, checked
    assert defines["CFD_FIELD_COUNT"] == max(v for k, v in vars(_abi).items() if k.startswith("FIELD_") and isinstance(v, int)) + 1


def test_every_environment_hook_of_the_library_is_documented():
    """INTEGRATION.md lists the library's tuning / A-B hooks: every getenv() in csrc/ must appear there (doc drift check)."""
    csrc = os.path.join(ROOT, "cfd_demo_b200", "csrc")
    src = ""
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(csrc, name)) as f:
                src += f.read()
    hooks = set(re.findall(r'getenv\("([A-Z0-9_]+)"\)', src))
    hooks |= set(re.findall(r'tile_hook\("([A-Z0-9_]+)"', src))
    assert len(hooks) >= 15, hooks
    with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
        doc = f.read()
    missing = sorted(h for h in hooks if h not in doc)
    assert not missing, missing
