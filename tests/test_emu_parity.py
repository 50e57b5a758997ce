"""CPU-side parity: the CUDA kernel SOURCE executed by the emulator (impec2d_b200/csrc/emu.cpp) against the
oracle.  Validates index logic and host orchestration in this GPU-less container; the same checks run on the
B200 in tests/test_gpu_parity.py.  The emulated library is never used by the product."""
import os
import subprocess
import pytest
import oracle as O
import impec2d_b200 as I
import parity_cases as pc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module", autouse=True)
def emu_lib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "..", "impec2d_b200", "csrc"), "emu"])
    return I.load(emulated=True)


def test_ghost_fills(): pc.check_ghost_fills(True)
def test_velocity_roundtrip(): pc.check_velocity_roundtrip(True)
def test_single_sweeps_bitexact(): pc.check_single_sweeps_bitexact(True)
def test_intercell_fluxes_bitexact(): pc.check_intercell_fluxes(True)
def test_single_sweeps_multitile_bitexact(): pc.check_single_sweeps_bitexact(True, N=300)      # > 64 rows (k_muscl_x chunks) and > 256 columns (k_muscl_y tiles)
def test_advection_muscl_multitile_bitexact(): pc.check_advection_bitexact(True, N=290, steps=1)
def test_sweep_kernels_agree(): pc.check_sweep_variants(True, N=150)
def test_advection_muscl_bitexact(): pc.check_advection_bitexact(True, N=24, steps=2)
def test_advection_upwind_bitexact(): pc.check_advection_bitexact(True, N=24, steps=2, method=O.UPWIND)
def test_advection_porosity_bitexact(): pc.check_advection_bitexact(True, N=16, steps=1, porosity=0.37)


@pytest.mark.parametrize("split", [O.ALTERNATING, O.X_FIRST, O.Y_FIRST, O.STRANG])
def test_advection_other_splittings_bitexact(split): pc.check_advection_bitexact(True, N=16, steps=2, split=split)


def test_projection_parallel_to_diagonal(): pc.check_projection(True)
def test_checkpoint_restart(tmp_path): pc.check_checkpoint_restart(True, tmp_path)
def test_source_step(): pc.check_source_step(True, N=32)
def test_scattered_sources(): pc.check_scattered_sources(True)
def test_vcycle_parallel(): pc.check_vcycle(True, 32, 32)
def test_vcycle_rect_diagonal(): pc.check_vcycle(True, 32, 16, stype=O.DIAGONAL)
def test_impes_step_parallel(): pc.check_impes_step(True, 32, 32, steps=1)
def test_impes_step_diagonal_rect(): pc.check_impes_step(True, 32, 16, stype=O.DIAGONAL, steps=1)
def test_elliptic_manufactured_5pt(): pc.check_elliptic_manufactured(True, 20, 30, testno=4, nu=1.0)
def test_elliptic_manufactured_9pt(): pc.check_elliptic_manufactured(True, 20, 30, testno=2, nu=2.0 / 3.0)
def test_fused_vs_unfused_cycle(): pc.check_fused_vs_unfused(True, 256, 384)
def test_dispersion_parallel(): pc.check_dispersion(True, 32, 32)
def test_dispersion_diagonal_rect(): pc.check_dispersion(True, 32, 16, stype=O.DIAGONAL)
def test_matvec_march_vs_tiles(): pc.check_matvec_variants(True)
def test_matvec_march_odd_columns(): pc.check_matvec_variants(True, 20, 31, eps=1e-9)
def test_march_vs_tile_kernels(): pc.check_march_vs_tiles(True)
def test_sgs_march_vs_sweeps(): pc.check_sgs_march(True)
def test_fast_initial_residual(): pc.check_fast_initial_residual(True)
def test_small_vs_tile_kernels(): pc.check_small_vs_tiles(True, 256, 384)
# round 2
def test_generators_bitexact(): pc.check_generators(True, N=48)
def test_one_cell_wells(): pc.check_one_cell_wells(True, N=32)
def test_explicit_diffusive_dt_bitexact(): pc.check_explicit_diffusive_dt(True, N=40)
def test_diagnostic_values(): pc.check_diagnostic_values(True, N=40)
def test_hyperbolic_test_config(): pc.check_hyperbolic_test_config(True, 64, steps=2)
def test_integral_diagnostics_device_vs_host(): pc.check_integral_diagnostics(True, N=48, steps=2)
def test_fused_operator_build_bitexact(): pc.check_fused_build(True, 40, 67)


# source_type "horizontal" (SURVEY 8f rank 4): Dirichlet rows, horizontal-front source with pulsing / linear injection
def test_horizontal_front(): pc.check_horizontal_front(True, 48, 32)
def test_horizontal_front_fp64_cycle_same_iterations(): pc.check_horizontal_front(True, 48, 32, links_fp32=0)
def test_horizontal_front_dispersion(): pc.check_horizontal_front(True, 48, 32, diffusion=1)
def test_horizontal_generators(): pc.check_horizontal_generators(True)
def test_horizontal_front_marching_kernels(monkeypatch):
    monkeypatch.setenv("IMPEC_MARCH_MIN_CELLS", "0")
    pc.check_horizontal_front(True, 128, 128, diffusion=1, steps=1)


# dispersion with alpha_l != alpha_t: non-symmetric Crank-Nicolson operator (matrix-free 9-point operator + BiCGStab)
def test_dispersion_anisotropic_tensor(): pc.check_dispersion(True, 40, 32, stype=O.DIAGONAL, steps=2, alpha_l=1.0, alpha_t=0.2)
def test_dispersion_anisotropic_tensor_parallel(): pc.check_dispersion(True, 32, 32, stype=O.PARALLEL, steps=1, alpha_l=0.3, alpha_t=1.0)


# edge cases: odd and tiny grids (an odd extent leaves a single multigrid level; 4 x 4 is the smallest grid the ABI accepts)
@pytest.mark.parametrize("N1,N2,stype", [(30, 21, O.DIAGONAL), (21, 30, O.DIAGONAL), (6, 6, O.PARALLEL), (4, 4, O.DIAGONAL), (50, 36, O.DIAGONAL)])
def test_impes_step_ragged_and_tiny_grids(N1, N2, stype): pc.check_impes_step(True, N1, N2, stype=stype, steps=1)
def test_horizontal_front_tiny_grid(): pc.check_horizontal_front(True, 12, 8, steps=1)
def test_one_dimensional_flow_with_dispersion(): pc.check_one_dimensional_flow_with_dispersion(True)


# marching kernels (forced onto small levels) on extents that are even but not multiples of 4 / of the chunk height / of a warp's 56-60 columns
@pytest.mark.parametrize("N1,N2", [(30, 22), (70, 66), (34, 130)])
def test_marching_kernels_awkward_extents(monkeypatch, N1, N2):
    monkeypatch.setenv("IMPEC_MARCH_MIN_CELLS", "0")
    pc.check_impes_step(True, N1, N2, stype=O.DIAGONAL, steps=1)
    pc.check_march_vs_tiles(True, N1, N2, stype=O.DIAGONAL)


def test_deferred_solution_update_bitexact(): pc.check_deferred_solution_update(True, 48, 80)


def test_random_shapes_fuzz():
    """Three cases of tools/emu_fuzz.py (random workload, grid extents and parameters; the tool itself was run over 120 cases)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("emu_fuzz", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "emu_fuzz.py"))
    fuzz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fuzz)
    assert fuzz.main(20261018, 3) == 0
