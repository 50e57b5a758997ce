"""CPU tests of the C++ host layer (impec2d_b200/host): the reference's class surface over the C ABI, driven by the
examples/*.cfg files, linked against the CPU-emulated kernel library (tests only)."""
import os
import re
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "impec2d_b200", "host")
EXE = os.path.join(ROOT, "impec2d_b200", "impec2d_emu")


@pytest.fixture(scope="module", autouse=True)
def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "impec2d_b200", "csrc"), "emu"])
    subprocess.check_call(["make", "-s", "-C", HOST, "emu"])


def cfg_variant(tmp_path, name, subs):
    text = open(os.path.join(ROOT, "examples", name)).read()
    for pat, rep in subs:
        text, n = re.subn(pat, rep, text)
        assert n >= 1, pat
    text = text.replace('"/tmp', '"' + str(tmp_path))
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_all_example_configs_parse(tmp_path):
    """Every shipped example parses (the driver gets as far as its first step or test level)."""
    for name in ("parallel.cfg", "parallel_dispersion.cfg", "diagonal_dispersion.cfg"):
        subs = [(r"x_cells\s*=\s*\d+", "x_cells = 16"), (r"y_cells\s*=\s*\d+", "y_cells = %d" % (8 if "diagonal" in name else 16))]
        cfg = cfg_variant(tmp_path, name, subs)
        out = subprocess.run([EXE, cfg, "--steps", "1"], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "[1]\tITS_e:" in out.stdout


def test_hyperbolic_test_cfg(tmp_path):
    """hyperbolic_test.cfg through Impec2DSolver::TestAdvection: L1 norms of the oracle (tests/test_oracle.py) at 50^2, 100^2."""
    cfg = cfg_variant(tmp_path, "hyperbolic_test.cfg", [(r"levels\s*=\s*4", "levels = 2")])
    out = subprocess.run([EXE, cfg], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    rows = np.loadtxt(str(tmp_path / "hyperbolic_norms.out"))
    assert np.exp(rows[0, 1]) == pytest.approx(1.851837e-03, rel=1e-6)
    assert np.exp(rows[1, 1]) == pytest.approx(5.158563e-04, rel=1e-6)


def test_elliptic_test_cfg(tmp_path):
    """elliptic_test.cfg through EllipticSolver::TestNeumann: second-order decay of the pressure and velocity errors."""
    cfg = cfg_variant(tmp_path, "elliptic_test.cfg", [(r"levels\s*=\s*6", "levels = 3")])
    out = subprocess.run([EXE, cfg], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    rows = np.loadtxt(str(tmp_path / "elliptic_norms.out"))
    for col in (1, 2, 3):
        order = (rows[2, col] - rows[1, col]) / (rows[2, 0] - rows[1, 0])
        assert 1.9 < order < 2.1, (col, order)


def test_impes_run_diagnostics(tmp_path):
    """parallel.cfg at 32^2 for 3 steps: per-step invariants the reference logs (mass balance, no-flow walls, div v = q)."""
    cfg = cfg_variant(tmp_path, "parallel.cfg", [(r"x_cells\s*=\s*\d+", "x_cells = 32"), (r"y_cells\s*=\s*\d+", "y_cells = 32")])
    out = subprocess.run([EXE, cfg, "--steps", "3"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    diag = np.loadtxt(str(tmp_path / "diagnostics.dat"))
    assert diag.shape == (3, 8)
    assert np.all(diag[:, 1] > 5) and np.all(diag[:, 1] < 40)         # PCG iterations
    assert np.all(diag[:, 5] < 1e-10)                                  # mass-balance error
    assert np.all(diag[:, 6] < 1e-10)                                  # boundary velocity
    assert np.all(diag[:, 7] < 1e-8)                                   # divergence error
    assert os.path.exists(str(tmp_path / "concentration.dat")) and os.path.exists(str(tmp_path / "pressure.dat"))


def test_bad_config_is_reported(tmp_path):
    p = tmp_path / "bad.cfg"
    p.write_text("application: { general: { x_cells = 10; }; };")
    out = subprocess.run([EXE, str(p)], capture_output=True, text=True)
    assert out.returncode == 1 and "not found" in out.stderr
