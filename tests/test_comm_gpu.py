"""The multi-GPU entry points of the C ABI (include/vo_b200.h section 1b): one process, map
replicated, queries sharded, indices gathered with ncclAllGather.  With a single visible GPU the
communicator degenerates to one shard (still through vo_comm_*); with more, the sharded answers must
equal the single-GPU ones.  The CPU side of the N>1 logic (shard bounds, ragged gathers) is covered
by tests/test_sharding_cpu.py on gloo."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "visual-odometry_b200", "host", "bin", "nn_sharded")


def test_sharded_api_matches_single_gpu(vo, oracle, synth):
    n = min(vo.device_count(), 4)
    M, Q = 50_000, 10_001  # ragged shards
    m = synth.nn_map_rows_np(0, M)
    q, target = synth.nn_queries_np(Q, M)
    sh = vo.ShardedNN(n)
    assert sh.size() == n
    sh.set_map(m)
    got = sh.best_match(q, 0.1)
    sh.close()
    oi, _ = oracle.nn_best_match(m, q, 0.1)
    assert np.array_equal(got, oi)
    with pytest.raises(vo.VoError):
        vo.ShardedNN(0)
    with pytest.raises(vo.VoError):
        vo.ShardedNN(vo.device_count() + 1)


def test_cpp_caller_of_the_sharded_abi(vo):
    if not os.path.exists(APP):
        pytest.skip("host/bin/nn_sharded not built")
    n = min(vo.device_count(), 8)
    out = subprocess.run([APP, str(n), "300000", "20000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:] + out.stdout[-500:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["n_gpus"] == n and r["sharded_differs_from_single"] == 0 and r["planted_wrong"] == 0
