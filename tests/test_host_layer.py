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


def test_parallel_run_writes_projected_fields(tmp_path):
    """A PARALLEL run logs c and p on the diagonal grid of round(N/sqrt 2) cells (Impec2DSolver::Project):
    the result files hold dN^2 points spaced L/(sqrt 2 dN).  (Values are checked bit for bit against the oracle in the parity tests;
    upstream's interpolation extrapolates in the lower half of a cell, so the projected field may overshoot [0, 1].)"""
    N = 32
    cfg = cfg_variant(tmp_path, "parallel.cfg", [(r"x_cells\s*=\s*\d+", "x_cells = %d" % N), (r"y_cells\s*=\s*\d+", "y_cells = %d" % N)])
    out = subprocess.run([EXE, cfg, "--steps", "3"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    text = open(cfg).read()
    rdir = re.search(r'results_dir\s*=\s*"([^"]+)"', text).group(1)
    dN = int(round(N / np.sqrt(2.)))
    for name in ("concentration.dat", "pressure.dat"):
        rows = np.loadtxt(os.path.join(rdir, name))
        assert rows.shape == (dN * dN, 3), (name, rows.shape)
        L = 1.41421356237
        d = L / (np.sqrt(2.) * dN)
        xs = np.unique(np.round(rows[:, 0], 9))
        assert len(xs) == dN and abs((xs[1] - xs[0]) - d) < 1e-6
    conc = np.loadtxt(os.path.join(rdir, "concentration.dat"))[:, 2]
    assert np.isfinite(conc).all() and conc.max() > 0.5 and conc.min() > -0.5
    # the velocity stays on the native grid
    assert np.loadtxt(os.path.join(rdir, "velocity_u.dat")).shape[0] >= N * N


def test_driver_checkpoint_restart(tmp_path):
    """impec2d --checkpoint / --restart: 2 steps in one go == 1 step, checkpoint, restart, 1 more step
    (same concentration file, same last diagnostics line)."""
    subs = [(r"x_cells\s*=\s*\d+", "x_cells = 24"), (r"y_cells\s*=\s*\d+", "y_cells = 24")]
    da = tmp_path / "a"; db = tmp_path / "b"; da.mkdir(); db.mkdir()
    cfg_a = cfg_variant(da, "parallel.cfg", subs); cfg_b = cfg_variant(db, "parallel.cfg", subs)
    ck = str(tmp_path / "run.ckpt")
    out = subprocess.run([EXE, cfg_a, "--steps", "2"], capture_output=True, text=True); assert out.returncode == 0, out.stdout + out.stderr
    last_a = [l for l in out.stdout.splitlines() if l.startswith("[2]")]
    out = subprocess.run([EXE, cfg_b, "--steps", "1", "--checkpoint", ck], capture_output=True, text=True); assert out.returncode == 0, out.stdout + out.stderr
    out = subprocess.run([EXE, cfg_b, "--steps", "1", "--restart", ck], capture_output=True, text=True); assert out.returncode == 0, out.stdout + out.stderr
    assert "Restarted from" in out.stdout
    last_b = [l for l in out.stdout.splitlines() if l.startswith("[2]")]
    assert last_a and last_a == last_b, (last_a, last_b)
    ca = np.loadtxt(str(da / "concentration.dat")); cb = np.loadtxt(str(db / "concentration.dat"))
    assert np.array_equal(ca, cb)


def build_api_check(tmp_path, emulated=True):
    exe = str(tmp_path / "host_api_check")
    host_lib = "impec2d_host_emu" if emulated else "impec2d_host"
    kern_lib = "impec2d_emu" if emulated else "impec2d_b200"
    lib_dir = os.path.join(ROOT, "impec2d_b200")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-I", HOST, "-o", exe, os.path.join(ROOT, "tests", "host_api_check.cpp"),
                           "-L", lib_dir, "-l" + host_lib, "-l" + kern_lib, "-Wl,-rpath," + lib_dir])
    return exe


def test_class_surface_called_directly(tmp_path):
    """The C++ classes themselves (not the driver): CalculateIntercellFlux_F/G consistent with Advect_X/Y to the bit for MUSCL and
    upwind (src/HyperbolicSplitSolver.h:19-22, .cpp:113-159), and the loop body of Impec2DSolver::Solve through EllipticSolver::Solve,
    CalcVelocities, MaxStableTimestep and HyperbolicSolver::Solve."""
    exe = build_api_check(tmp_path)
    cfg = cfg_variant(tmp_path, "parallel.cfg", [(r"x_cells\s*=\s*\d+", "x_cells = 48"), (r"y_cells\s*=\s*\d+", "y_cells = 48")])
    out = subprocess.run([exe, cfg], capture_output=True, text=True)
    assert out.returncode == 0 and "host_api_check ok" in out.stdout, out.stdout + out.stderr


def test_horizontal_front_with_pulsing_injection(tmp_path):
    """source_type "horizontal" through the driver: Dirichlet row, InitHorizontalFront re-evaluated every step
    (src/Impec2DSolver.cpp:437-438).  With mode = "pulsing" the rate jumps by q_high_multiple in the second half of each period 1/q_freq,
    so the CFL time step drops by that factor there; mass stays balanced throughout."""
    cfg = cfg_variant(tmp_path, "parallel.cfg", [(r"x_cells\s*=\s*\d+", "x_cells = 64"), (r"y_cells\s*=\s*\d+", "y_cells = 48"),
                                                 (r'source_type\s*=\s*"parallel"', 'source_type = "horizontal"'), (r'(\s)mode\s*=\s*"\w+"', r'\1mode = "pulsing"'),
                                                 (r"x_ub\s*=\s*[0-9.]+", "x_ub = 1.0"), (r"y_ub\s*=\s*[0-9.]+", "y_ub = 1.0"),
                                                 (r"q_high_multiple\s*=\s*[0-9.]+", "q_high_multiple = 4.")])
    out = subprocess.run([EXE, cfg, "--steps", "8"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    rows = [l for l in out.stdout.splitlines() if l.startswith("[")]
    assert len(rows) == 8
    t = [float(re.search(r"\st: ([0-9.eE+-]+)", l).group(1)) for l in rows]
    dt = [float(re.search(r"dt: ([0-9.eE+-]+)", l).group(1)) for l in rows]
    em = [float(re.search(r"E_m: ([0-9.eE+-]+)", l).group(1)) for l in rows]
    # q_freq = 20: low rate while fmod(t, 0.05) <= 0.025, 4 x the rate afterwards -> the time step of a step that STARTS in the high
    # half is about a quarter of the first one
    high = [k for k in range(1, 8) if (t[k - 1] % 0.05) > 0.025]
    assert high, t
    assert all(dt[k] < 0.4 * dt[0] for k in high) and dt[1] > 0.8 * dt[0], (t, dt)
    # (the reference's mass-balance diagnostic assumes the base rate, src/Impec2DSolver.cpp:391-415: exact while the rate is low, off afterwards)
    low = [k for k in range(8) if k not in high and all(h > k for h in high)]
    assert all(em[k] < 1e-6 for k in low), em


@pytest.mark.parametrize("name,n1,n2,world", [("diagonal_dispersion.cfg", 64, 32, 2), ("parallel.cfg", 64, 64, 2), ("parallel.cfg", 64, 64, 4)])
def test_cpp_driver_on_slabs(tmp_path, name, n1, n2, world):
    """Impec2DSolver(cfg, myid, num_procs) (src/Impec2DSolver.h:54, src/Driver.cpp:58-87) as `world` PROCESSES of the emulated build, one
    slab each (ranks from IMPEC_RANK / IMPEC_NRANKS, the processes connected through files in IMPEC_EMU_COMM_DIR: csrc/emu_filecomm.inl),
    against the same driver as one process: the diagnostics lines and every result file -- the dispersion run's native-grid fields, the
    parallel run's c and p PROJECTED onto the diagonal grid by the collective Impec2DSolver::Project -- agree to 1e-9.  The GPU twin over
    NCCL is tests/test_gpu_multirank.py::test_cpp_driver_on_two_gpus."""
    outs = {}
    subs = [(r"x_cells\s*=\s*\d+", "x_cells = %d" % n1), (r"y_cells\s*=\s*\d+", "y_cells = %d" % n2)]
    for w in (1, world):
        d = tmp_path / ("w%d" % w); d.mkdir()
        comm = d / "comm"; comm.mkdir()
        cfg = cfg_variant(d, name, subs)
        procs = []
        for r in range(w):
            env = dict(os.environ, IMPEC_RANK=str(r), IMPEC_NRANKS=str(w), IMPEC_EMU_COMM_DIR=str(comm), IMPEC_DIST_MIN_CELLS="0", OMP_NUM_THREADS="1")
            procs.append(subprocess.Popen([EXE, cfg, "--steps", "3"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
        logs = [p.communicate(timeout=600)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(l[-2000:] for l in logs)
        assert "[3]\tITS_e:" in logs[0], logs[0][-2000:]
        assert not os.listdir(str(comm)), "the file communicator left files behind"
        res = [os.path.join(dp, f) for dp, _, fs in os.walk(str(d)) for f in fs if f.endswith(".dat")]
        outs[w] = {os.path.basename(f): np.loadtxt(f) for f in res}
        outs[w]["__steps__"] = [l for l in logs[0].splitlines() if re.match(r"\[\d+\]\tITS_e:", l)]
    assert {"concentration.dat", "pressure.dat", "velocity_u.dat", "velocity_v.dat", "diagnostics.dat"} <= set(outs[world]), sorted(outs[world])
    assert len(outs[1]["__steps__"]) == 3 and len(outs[world]["__steps__"]) == 3
    for name_, a in outs[1].items():
        if name_ == "__steps__":
            continue
        b = outs[world][name_]
        assert a.shape == b.shape, (name_, a.shape, b.shape)
        assert np.abs(a - b).max() <= 1e-9 * max(np.abs(a).max(), 1.0), (name_, np.abs(a - b).max())


def test_driver_checkpoint_restart_on_slabs(tmp_path):
    """`impec2d --checkpoint / --restart` as two processes: every rank writes / reads its own file (RankFile); 2 steps in one go ==
    1 step, checkpoint, restart, 1 more step -- same diagnostics line, same gathered concentration file."""
    subs = [(r"x_cells\s*=\s*\d+", "x_cells = 32"), (r"y_cells\s*=\s*\d+", "y_cells = 16")]
    ck = str(tmp_path / "run.ckpt")

    def run(d, extra):
        comm = d / ("comm%d" % len(os.listdir(str(d)))); comm.mkdir()
        procs = []
        for r in range(2):
            env = dict(os.environ, IMPEC_RANK=str(r), IMPEC_NRANKS="2", IMPEC_EMU_COMM_DIR=str(comm), IMPEC_DIST_MIN_CELLS="0", OMP_NUM_THREADS="1")
            procs.append(subprocess.Popen([EXE, cfgs[d]] + extra, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
        logs = [p.communicate(timeout=600)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(l[-2000:] for l in logs)
        return logs[0]

    da = tmp_path / "a"; db = tmp_path / "b"; da.mkdir(); db.mkdir()
    cfgs = {da: cfg_variant(da, "diagonal_dispersion.cfg", subs), db: cfg_variant(db, "diagonal_dispersion.cfg", subs)}
    last_a = [l for l in run(da, ["--steps", "2"]).splitlines() if l.startswith("[2]")]
    run(db, ["--steps", "1", "--checkpoint", ck])
    out = run(db, ["--steps", "1", "--restart", ck])
    assert "Restarted from" in out
    last_b = [l for l in out.splitlines() if l.startswith("[2]")]
    assert last_a and last_a == last_b, (last_a, last_b)
    find = lambda d: [os.path.join(dp, f) for dp, _, fs in os.walk(str(d)) for f in fs if f == "concentration.dat"][0]
    assert np.array_equal(np.loadtxt(find(da)), np.loadtxt(find(db)))


def test_elliptic_test_cfg_on_slabs(tmp_path):
    """elliptic_test.cfg (EllipticSolver::TestNeumann, src/EllipticSolver.cpp:191-239) as two processes: a new slab context per refinement
    level, norms over the gathered fields on rank 0 -- the same norm file as the one-process run."""
    norms = {}
    for w in (1, 2):
        d = tmp_path / ("w%d" % w); d.mkdir(); comm = d / "comm"; comm.mkdir()
        cfg = cfg_variant(d, "elliptic_test.cfg", [(r"levels\s*=\s*6", "levels = 3")])
        procs = []
        for r in range(w):
            env = dict(os.environ, IMPEC_RANK=str(r), IMPEC_NRANKS=str(w), IMPEC_EMU_COMM_DIR=str(comm), IMPEC_DIST_MIN_CELLS="0", OMP_NUM_THREADS="1")
            procs.append(subprocess.Popen([EXE, cfg], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
        logs = [p.communicate(timeout=600)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(l[-2000:] for l in logs)
        assert not os.listdir(str(comm))
        norms[w] = np.loadtxt(str(d / "elliptic_norms.out"))
    assert norms[1].shape == norms[2].shape == (3, 4)
    assert np.abs(norms[1] - norms[2]).max() <= 1e-6, np.abs(norms[1] - norms[2]).max()


def test_hyperbolic_test_cfg_on_slabs(tmp_path):
    """hyperbolic_test.cfg (Impec2DSolver::TestAdvection) as two processes: the periodic ghost rows come from the other rank (ranks closed to a
    ring); the transport path is explicit, so the norm file equals the one-process file to the last digit.  48^2 / 96^2 instead of the shipped
    50^2: a slab needs an even number of rows."""
    norms = {}
    subs = [(r"levels\s*=\s*4", "levels = 2"), (r"initial_x_cells\s*=\s*\d+", "initial_x_cells = 48"), (r"initial_y_cells\s*=\s*\d+", "initial_y_cells = 48")]
    for w in (1, 2):
        d = tmp_path / ("w%d" % w); d.mkdir(); comm = d / "comm"; comm.mkdir()
        cfg = cfg_variant(d, "hyperbolic_test.cfg", subs)
        procs = []
        for r in range(w):
            env = dict(os.environ, IMPEC_RANK=str(r), IMPEC_NRANKS=str(w), IMPEC_EMU_COMM_DIR=str(comm), IMPEC_DIST_MIN_CELLS="0", OMP_NUM_THREADS="1")
            procs.append(subprocess.Popen([EXE, cfg], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
        logs = [p.communicate(timeout=600)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(l[-2000:] for l in logs)
        assert not os.listdir(str(comm))
        norms[w] = np.loadtxt(str(d / "hyperbolic_norms.out"))
    assert norms[1].shape == (2, 2) and np.array_equal(norms[1], norms[2]), (norms[1], norms[2])
    order = (norms[1][1, 1] - norms[1][0, 1]) / (norms[1][1, 0] - norms[1][0, 0])
    assert 1.7 < order < 2.0, order
