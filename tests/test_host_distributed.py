"""Multi-process CPU tests (torch.distributed gloo, world_size 2 and 4) of the host logic of the N>1 path:
Comm collectives, Distributor plans, BlockMaps, Import/Export lists and fillComplete layout, each compared
bit-exactly with the CPU oracle / the reference's 4-rank golden vectors."""
import pytest

import dist_cases


@pytest.mark.parametrize("world", [2, 4])
def test_comm_and_distributor(world):
    dist_cases.run("case_comm_and_distributor", world)


@pytest.mark.parametrize("world", [2, 4])
def test_blockmaps(world):
    dist_cases.run("case_blockmaps", world)


@pytest.mark.parametrize("world", [2, 4])
def test_shifted_import_export(world):
    dist_cases.run("case_shifted_plans", world)


@pytest.mark.parametrize("world,kind,N", [(2, "7pt", 6), (4, "7pt", 9), (2, "27pt", 5), (4, "5pt", 11), (4, "powerlaw", 900), (2, "powerlaw", 301)])
def test_fill_complete_matches_oracle(world, kind, N):
    dist_cases.run("case_fill_complete", world, kind, N, 150)


def test_laplacian_1d_appendix_b():
    dist_cases.run("case_laplacian_1d_golden", 4)


@pytest.mark.parametrize("world", [2, 4])
def test_irregular_matrix_with_permutes_and_remotes(world):
    dist_cases.run("case_irregular_matrix", world)


@pytest.mark.parametrize("world", [2, 4])
def test_nontrivial_export_plan(world):
    dist_cases.run("case_export_plan", world)


@pytest.mark.parametrize("world,kind,N", [(2, "7pt", 6), (4, "27pt", 5), (4, "powerlaw", 900), (3, "5pt", 11)])
def test_device_fill_complete_glue_on_the_cpu_model(world, kind, N):
    dist_cases.run("case_device_fill_glue", world, kind, N, 150)


@pytest.mark.parametrize("world", [2, 4])
def test_overlapping_row_map_fill_complete_and_export_plan(world):
    dist_cases.run("case_overlapping_rows_fill", world)
