"""Dry-runs, on the CPU, the bodies of the GPU tests that were added or changed after the last hardware run, with the
oracle-backed stand-in of tests/shadow_backend.py in place of the `gpu` fixture.  Catches Python-level mistakes in
those tests (they would otherwise only surface on the GPU box, where `-x` stops the suite at the first failure);
says nothing about the product."""
import pytest

import test_gpu_parity as G
import test_gpu_z_unvalidated as Z
from shadow_backend import ShadowBackend


@pytest.mark.parametrize("name", ["B361", "C342", "D361", "E342", "F361", "F342"])
def test_dry_run_distributed_load_decks(name):
    G.test_solve_golden_decks_distributed_loads(ShadowBackend(), name)


@pytest.mark.parametrize("name", ["A361", "A342"])
def test_dry_run_golden_decks(name):
    G.test_solve_golden_decks(ShadowBackend(), name)


@pytest.mark.parametrize("meshname,precond,fused", [("hex12", 3, 1), ("tet4", 1, 0)])
def test_dry_run_solve_matches_oracle(meshname, precond, fused):
    G.test_solve_matches_oracle(ShadowBackend(), meshname, precond, fused)


def test_dry_run_ic_inverse_variants():
    Z.test_ic_inverse_selftest_picks_a_kernel_that_matches_the_oracle(ShadowBackend())
    Z.test_ic_inverse_register_kernel_matches_the_oracle(ShadowBackend(), 1)
    Z.test_ic_inverse_register_kernel_matches_the_oracle(ShadowBackend(), 2)


@pytest.mark.parametrize("name", ["A361", "B342", "F361", "F342"])
def test_dry_run_linear_update_and_nodal_stress(name):
    import test_gpu_z_unvalidated as Z
    Z.test_linear_update_and_nodal_stress_golden_decks(ShadowBackend(), name)


def test_dry_run_linear_update_mixed():
    import test_gpu_z_unvalidated as Z
    Z.test_linear_update_mixed_types_and_sections(ShadowBackend())


def test_dry_run_mixed_types_and_sections():
    import test_gpu_z_unvalidated as Z
    Z.test_assembly_mixed_types_and_sections(ShadowBackend())


@pytest.mark.parametrize("name", ["I361", "I342", "cylinder"])
def test_dry_run_nonlinear_newton(name, monkeypatch):
    """Body of the opt-in nonlinear GPU test with the device calls of the Newton-driver mirror replaced by the oracle."""
    import test_gpu_z_unvalidated as Z
    from shadow_backend import patch_nonlinear
    Z.test_nlgeom_newton_matches_oracle_and_golden(patch_nonlinear(monkeypatch), name)
