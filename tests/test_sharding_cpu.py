"""CPU, world_size 2 over gloo: the multi-GPU path of the NN sweep (contiguous query shards,
replicated map, all-gather of int32 indices).  Each rank answers its shard with the oracle; the
gathered result must equal the unsharded answer on every rank."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_queries, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import oracle_lib

    synth = importlib.import_module("visual-odometry_b200.synth")
    sharding = importlib.import_module("visual-odometry_b200.sharding")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    m = synth.nn_map_rows_np(0, 4000)
    q, _ = synth.nn_queries_np(n_queries, 4000)
    lo, hi = sharding.shard_bounds(n_queries, world, rank)
    idx_shard = torch.from_numpy(oracle_lib.nn_best_match(m, q[lo:hi], 0.1)[0])
    idx_all = torch.empty(n_queries, dtype=torch.int32)
    sharding.gather_indices(dist, idx_shard, idx_all, n_queries)
    full = oracle_lib.nn_best_match(m, q, 0.1)[0]
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([np.array_equal(idx_all.numpy(), full)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_queries", [1000, 1001])  # equal and ragged shards
def test_sharded_queries_allgather(tmp_path, n_queries):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_queries, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(os.path.join(tmp_path, f"ok_{r}.npy"))[0]


def test_shard_bounds_cover_everything():
    sharding = importlib.import_module("visual-odometry_b200.sharding")
    for q in (0, 1, 7, 100000, 100003):
        for w in (1, 2, 4, 8):
            b = [sharding.shard_bounds(q, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == q
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(sharding.shard_counts(q, w)) - min(sharding.shard_counts(q, w)) <= 1
