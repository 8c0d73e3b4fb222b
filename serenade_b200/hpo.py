"""The hyper-parameter search call pattern of the reference on the B200 path.

`objective()` mirrors src/objective.rs:8-52 — every trial REBUILDS the index from the training TSV (with the trial's
m and idf_weighting) and replays every prefix of every validation session through `predict`, returning Mrr@20; here
the rebuild runs on the device and the replay is one batched `predict_batch`.  `HyperParamGrid` restates
src/hyperparameter/hyperparamgrid.rs:6-82, `exhaustive_grid_search()` the loop of src/bin/exhaustive_grid_search.rs:
54-96 and `random_search()` that of src/bin/hyperparameter_search.rs (n random combinations of the grid).  The TPE
driver (src/bin/tpe_hyperparameter_optm.rs) sits on the third-party `tpe` crate and is not rebuilt; its objective is
this same function.
"""
import itertools
import random

from .evaluate import evaluate, read_test_sessions
from .vmis import VMISIndex, read_sessions_csv


def objective(path_to_training, test_data_file, n_most_recent_sessions, neighborhood_size_k, last_items_in_session,
              idf_weighting, enable_business_logic=False, device=0, max_len=0, test_sessions=None,
              training_sessions=None):
    """objective.rs:8-52 → Mrr@20 (qty_max_reco_results = 20, :21).  `training_sessions` (from `read_sessions_csv`)
    skips re-parsing the training file: the index is still REBUILT for every trial, as in the reference (:17)."""
    if training_sessions is not None:
        index = VMISIndex.from_sessions(*training_sessions, int(n_most_recent_sessions), max_len, float(idf_weighting),
                                        device=device)
    else:
        index = VMISIndex.new_from_csv(path_to_training, int(n_most_recent_sessions), float(idf_weighting),
                                       max_len=max_len, device=device)
    try:
        sessions = test_sessions if test_sessions is not None else read_test_sessions(test_data_file)   # :19
        res = evaluate(index, sessions, int(neighborhood_size_k), int(n_most_recent_sessions), how_many=20,
                       max_items_in_session=int(last_items_in_session), length=20,
                       enable_business_logic=enable_business_logic)
        return res["mrr"]
    finally:
        index.close()


class HyperParamGrid:
    """hyperparamgrid.rs:6-82"""

    def __init__(self, param_grid):
        self.param_grid = dict(param_grid)

    def get_qty_combinations(self):                                                   # :71-82
        total = 0
        for values in self.param_grid.values():
            total = len(values) if total == 0 else total * len(values)
        return total

    def get_all_combinations(self):                                                   # :27-46
        keys = list(self.param_grid)
        return [dict(zip(keys, combo)) for combo in itertools.product(*(self.param_grid[k] for k in keys))]

    def get_n_random_combinations(self, n, rng=None):                                 # :17-25
        combos = self.get_all_combinations()
        (rng or random).shuffle(combos)
        return combos[:n]


def _search(train_data_path, test_data_path, combos, enable_business_logic, device, max_len):
    sessions = read_test_sessions(test_data_path)
    training = read_sessions_csv(train_data_path)                  # parsed once; every trial rebuilds the index from it
    best, best_value, records = None, float("-inf"), []
    for iteration, c in enumerate(combos):
        v = objective(train_data_path, test_data_path, c["n_most_recent_sessions"], c["neighborhood_size_k"],
                      c["last_items_in_session"], c["idf_weighting"], enable_business_logic, device, max_len, sessions,
                      training)
        records.append((iteration, c["n_most_recent_sessions"], c["neighborhood_size_k"], c["last_items_in_session"],
                        c["idf_weighting"], v))
        if v > best_value:                                                            # exhaustive_grid_search.rs:84-90
            best, best_value = dict(c), v
    return best, best_value, records


def exhaustive_grid_search(train_data_path, test_data_path, n_most_recent_sessions_choices, neighborhood_size_k_choices,
                           last_items_in_session_choices, idf_weighting_choices, enable_business_logic=False, device=0,
                           max_len=0):
    """exhaustive_grid_search.rs:54-96 → (best parameters, best Mrr@20, records as written to the results CSV)."""
    combos = [dict(n_most_recent_sessions=m, neighborhood_size_k=k, last_items_in_session=l, idf_weighting=w)
              for m in n_most_recent_sessions_choices for k in neighborhood_size_k_choices
              for l in last_items_in_session_choices for w in idf_weighting_choices]
    return _search(train_data_path, test_data_path, combos, enable_business_logic, device, max_len)


def random_search(train_data_path, test_data_path, grid, n, enable_business_logic=False, device=0, max_len=0, rng=None):
    """hyperparameter_search.rs: n random combinations of a HyperParamGrid with the four model parameters."""
    return _search(train_data_path, test_data_path, grid.get_n_random_combinations(n, rng), enable_business_logic, device,
                   max_len)
