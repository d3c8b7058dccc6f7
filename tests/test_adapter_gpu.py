"""The same adapter scenarios through the real CUDA backend (C ABI -> kernels)."""
import asyncio

import pytest

import adapter_scenarios as S

pytestmark = pytest.mark.gpu


def test_database_scenario(native_lib):
    asyncio.run(S.scenario_test_database(None))


def test_parity_with_oracle_manager(native_lib):
    asyncio.run(S.scenario_parity_with_oracle(None, n=3000, dim=256))


def test_error_convention(native_lib):
    asyncio.run(S.scenario_errors(None))


def test_client_shim_matchtext(native_lib):
    asyncio.run(S.scenario_client_shim(None))
