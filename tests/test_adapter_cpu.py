"""Host logic of the QdrantManager-compatible adapter, on CPU: the device is replaced by an oracle-backed fake."""
import asyncio

import adapter_scenarios as S
from helpers import ExactTieDevice, FakeDevice


def test_database_scenario():
    asyncio.run(S.scenario_test_database(FakeDevice))


def test_parity_with_oracle_manager():
    asyncio.run(S.scenario_parity_with_oracle(FakeDevice, n=1200, dim=64))


def test_error_convention():
    asyncio.run(S.scenario_errors(FakeDevice))


def test_client_shim_matchtext():
    asyncio.run(S.scenario_client_shim(FakeDevice))


def test_reindex_churn_reuses_rows():
    asyncio.run(S.scenario_reindex_churn(FakeDevice))


def test_mass_delete_compacts():
    asyncio.run(S.scenario_mass_delete_compacts(FakeDevice))


def test_random_operation_sequences_match_the_oracle():
    for seed in (1, 2):
        asyncio.run(S.scenario_random_ops(FakeDevice, seed))


def test_exact_ties_follow_the_id():
    asyncio.run(S.scenario_exact_ties_follow_the_id(ExactTieDevice))
