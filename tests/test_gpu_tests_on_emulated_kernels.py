"""Runs the bodies of GPU tests on the CPU with tests/emul_backend.py in place of the `gpu` fixture: assembly, Dirichlet
elimination, strain / stress update and nodal stress execute the product's OWN kernels (assembly.cuh compiled unchanged
for the host, warps as lock-step fibers), only the Krylov solve is the oracle's.  Covers the kernels that have not
run on hardware yet (register/shuffle IC condensation, linear strain / stress recovery, total-Lagrangian tangent and
update incl. the whole Newton loop of the reference's exI decks) and, as a control, ones that have."""
import numpy as np
import pytest

import test_gpu_parity as G
import test_gpu_z_unvalidated as Z
from emul_backend import EmulBackend


@pytest.mark.parametrize("name", ["hex5_ic", "hex4_fi", "hex4_bbar", "tet3", "hex1", "hex_ragged"])
def test_emulated_assembly_matches_oracle(name):
    G.test_assembly_matches_oracle(EmulBackend(), name)


def test_emulated_gather_equals_scatter():
    G.test_assembly_gather_equals_scatter_and_is_reproducible(EmulBackend(), "hex_ragged")


def test_emulated_ic_inverse_variants():
    Z.test_ic_inverse_selftest_picks_a_kernel_that_matches_the_oracle(EmulBackend())
    Z.test_ic_inverse_register_kernel_matches_the_oracle(EmulBackend(), 1)
    Z.test_ic_inverse_register_kernel_matches_the_oracle(EmulBackend(), 2)


def test_emulated_bc_elimination():
    G.test_bc_elimination_matches_oracle(EmulBackend())


def test_emulated_mixed_types_and_sections():
    Z.test_assembly_mixed_types_and_sections(EmulBackend())


@pytest.mark.parametrize("name", ["A361", "A342", "B361", "E342", "F361", "F342"])
def test_emulated_linear_update_and_nodal_stress_golden_decks(name):
    Z.test_linear_update_and_nodal_stress_golden_decks(EmulBackend(), name)


def test_emulated_linear_update_mixed():
    Z.test_linear_update_mixed_types_and_sections(EmulBackend())


def test_emulated_tl_tangent_and_update():
    Z.test_tl_tangent_and_update_match_oracle_elementwise(EmulBackend())


@pytest.mark.parametrize("name", ["I361", "I342"])
def test_emulated_nonlinear_newton_loop_reproduces_the_reference_logs(name):
    """The product's Newton driver with the total-Lagrangian tangent / update KERNELS (host-run) and the oracle's CG:
    same Newton iteration counts as the oracle's loop, the reference's exI golden log at every sub-step, and the
    strain / stress / Mises lines of the final state."""
    Z.test_nlgeom_newton_matches_oracle_and_golden(EmulBackend(), name)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("name", ["hex5_ic", "tet3", "hex1", "hex_ragged"])
def test_emulated_spmv_bitwise(name, variant):
    """The SpMV kernels (native, SELL, pair-interleaved SELL incl. the repack kernel) on the host: bit-identical to the
    oracle, as on the GPU (same fma chain per row)."""
    G.test_spmv_bitwise(EmulBackend(), name, variant)


@pytest.mark.parametrize("meshname", ["tet3", "hex_ragged"])
def test_emulated_sell_c_sigma(meshname):
    Z.test_sell_sigma_is_bitwise_identical_per_row(EmulBackend(), meshname)
