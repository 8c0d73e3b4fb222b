"""Hyper-parameter search helpers: the grid KATs of the reference (hyperparamgrid.rs:89-148) on CPU, the objective
(objective.rs:8-52) against the README's HPO result on the GPU."""
import os
import random

import pytest

from serenade_b200.hpo import HyperParamGrid, exhaustive_grid_search, objective

GRID = {"sample_size": [500, 750, 1000, 2500, 5000], "k": [50, 100, 500, 1000, 1500],
        "last_items_in_session": [1, 2, 3, 5, 10]}


def test_grid_known_answers():
    one = HyperParamGrid({"sample_size": [1000], "k": [500], "last_items_in_session": [10]}).get_all_combinations()
    assert one == [{"sample_size": 1000, "k": 500, "last_items_in_session": 10}]          # should_get_expected_results
    g = HyperParamGrid(GRID)
    assert g.get_qty_combinations() == 5 * 5 * 5                                          # should_determine_qty_combinations
    combos = g.get_all_combinations()
    assert len(combos) == 125 and len(combos[0]) == 3                                      # should_get_all_combinations
    assert len({tuple(sorted(c.items())) for c in combos}) == 125
    assert len(g.get_n_random_combinations(100000000)) == 125                              # should_get_n_random_combinations
    ten = g.get_n_random_combinations(10, random.Random(1))
    assert len(ten) == 10 and len({tuple(sorted(c.items())) for c in ten}) == 10


@pytest.mark.gpu
def test_objective_reproduces_readme_hpo_result(sb, toy_dir):
    """README.md:64-71: best parameters m=1502, k=288, idf=2, last_items=4 → MRR@20 0.3197 on valid.txt and 0.3401 on
    test.txt (tie order of the shipped binary unpinned → ±0.004)."""
    train = os.path.join(toy_dir, "train.txt")
    v = objective(train, os.path.join(toy_dir, "valid.txt"), 1502, 288, 4, 2.0, max_len=15)
    assert v == pytest.approx(0.3197, abs=0.004)
    t = objective(train, os.path.join(toy_dir, "test.txt"), 1502, 288, 4, 2.0, max_len=15)
    assert t == pytest.approx(0.3401, abs=0.004)


@pytest.mark.gpu
def test_exhaustive_grid_search_picks_the_best_trial(sb, toy_dir):
    train, valid = os.path.join(toy_dir, "train.txt"), os.path.join(toy_dir, "valid.txt")
    best, best_value, records = exhaustive_grid_search(train, valid, [100, 1502], [50, 288], [1, 4], [1], max_len=15)
    assert len(records) == 8 and best_value == max(r[-1] for r in records)
    assert best_value == pytest.approx(objective(train, valid, best["n_most_recent_sessions"], best["neighborhood_size_k"],
                                                 best["last_items_in_session"], best["idf_weighting"], max_len=15))
    assert best["last_items_in_session"] == 4 and best_value > 0.30
