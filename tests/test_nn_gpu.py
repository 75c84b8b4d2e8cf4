"""GPU parity: vo_nn_* vs the oracle's bruteForceBestMatch — indices must be BIT-EXACT."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _vec11(app):
    return np.concatenate([np.arange(len(app), dtype=np.float32)[:, None], app], axis=1)


@pytest.fixture(scope="module")
def bundled():
    return np.load(os.path.join(HERE, "golden", "bundled_frames.npz"))


def test_bundled_frames_match_id_truth_and_oracle(vo, oracle, bundled):
    frames = list(bundled["frames"])
    nn = vo.NNIndex(0)
    for a, b in zip(frames[:-1], frames[1:]):
        if b != a + 1:
            continue
        m, q = _vec11(bundled[f"app_{a}"]), _vec11(bundled[f"app_{b}"])
        nn.set_map(m)
        idx, d2 = nn.best_match(q, 0.1, want_d2=True)
        pos = {int(v): i for i, v in enumerate(bundled[f"ids_{a}"])}
        truth = np.array([pos.get(int(v), -1) for v in bundled[f"ids_{b}"]], np.int32)
        oi, od = oracle.nn_best_match(m, q, 0.1)
        assert np.array_equal(idx, truth)
        assert np.array_equal(idx, oi)
        assert np.array_equal(d2[idx >= 0], od[idx >= 0])
    nn.close()


@pytest.mark.parametrize("M,Q", [(1, 1), (127, 124), (128, 129), (1000, 257), (5000, 3000),
                                 (20000, 9000), (100003, 2048)])
def test_planted_random_vs_oracle(vo, oracle, synth, M, Q):
    m = synth.nn_map_rows_np(0, M)
    q, target = synth.nn_queries_np(Q, M)
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx, d2 = nn.best_match(q, 0.1, want_d2=True)
    nn.close()
    oi, od = oracle.nn_best_match(m, q, 0.1)
    assert np.array_equal(idx, oi)
    hit = oi >= 0
    assert np.array_equal(d2[hit], od[hit])
    exact = (np.arange(Q) % 4) < 2
    assert np.array_equal(idx[exact], target[exact])


# remainder size -> the (queries per thread, threads) instantiation nn_launch_filter picks for it
WIDE_REMAINDERS = [(1, (2, 256)), (300, (2, 256)), (600, (2, 384)), (1200, (4, 384)), (2000, (6, 384)),
                   (2800, (8, 384)), (3600, (10, 384)), (4400, (12, 384))]
WIDE_M = 38_000  # 297 tiles of 128 rows >= 2 x 148 SMs: large enough for the wide register tile


@pytest.mark.parametrize("rem,variant", WIDE_REMAINDERS)
def test_wide_kernel_and_every_remainder_variant_vs_oracle(vo, oracle, synth, monkeypatch, rem, variant):
    """The FP32 filter at headline shape (VO_NN_FORCE_PATH=ffma; by default a map this large goes
    to the tensor-core filter): batches above 8192 queries against a map of >= 2 tiles per SM run
    whole 4608-query tiles through nn_filter_kernel<12,384> plus ONE remainder launch whose register
    tile depends on the remainder size.  The launch log proves which instantiations answered;
    indices and d2 must equal the oracle's (brute_force_search.h:22-41)."""
    monkeypatch.setenv("VO_NN_FORCE_PATH", "ffma")
    Q = 2 * 4608 + rem
    m = synth.nn_map_rows_np(0, WIDE_M)
    q, target = synth.nn_queries_np(Q, WIDE_M)
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx, d2 = nn.best_match(q, 0.1, want_d2=True)
    launches = nn.last_launches()
    nn.close()
    assert [l[:2] for l in launches] == [(12, 384), variant], launches
    assert launches[0][2] == 2  # two full query tiles
    oi, od = oracle.nn_best_match(m, q, 0.1)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2[oi >= 0], od[oi >= 0])
    exact = (np.arange(Q) % 4) < 2
    assert np.array_equal(idx[exact], target[exact])


@pytest.mark.parametrize("Q,variant", [(3000, (2, 256)), (5000, (4, 256)), (700, (2, 64)),
                                       (9216 + 212, (4, 256))])
def test_small_map_routing_vs_oracle(vo, oracle, synth, Q, variant):
    """Maps of fewer than 2 tiles per SM (frame-to-frame association) take the narrow register
    tiles whatever the query count."""
    M = 4099
    m = synth.nn_map_rows_np(0, M)
    q, _ = synth.nn_queries_np(Q, M)
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx, d2 = nn.best_match(q, 0.1, want_d2=True)
    launches = nn.last_launches()
    nn.close()
    assert [l[:2] for l in launches] == [variant], launches
    oi, od = oracle.nn_best_match(m, q, 0.1)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2[oi >= 0], od[oi >= 0])


def _worst_case(kind, rng, M, Q):
    m = rng.uniform(-1, 1, (M, 11)).astype(np.float32)
    q = rng.uniform(-1, 1, (Q, 11)).astype(np.float32)
    if kind == "constant_prefix":
        m[:, 1:6] = 0.25  # the partial distance of every (query,row) pair is identical
        q[:, 1:6] = 0.25
        q[::2, 6:] = m[rng.randint(0, M, (Q + 1) // 2), 6:]  # half the queries have an exact match
    else:
        centre = rng.uniform(-1, 1, 11).astype(np.float32)
        m = (centre + rng.uniform(-0.03, 0.03, (M, 11))).astype(np.float32)
        q = (centre + rng.uniform(-0.03, 0.03, (Q, 11))).astype(np.float32)
    return m, q


@pytest.mark.parametrize("kind", ["constant_prefix", "clustered"])
def test_wide_kernel_worst_case_data_stays_exact(vo, oracle, monkeypatch, kind):
    """FP32 filter (forced).  No-pruning data (every row passes the partial-distance filter, every tile is re-scanned for
    every query) through the WIDE kernel and a remainder variant: only the number of re-scans may
    change, never the answers."""
    monkeypatch.setenv("VO_NN_FORCE_PATH", "ffma")
    rng = np.random.RandomState(13)
    M, Q = WIDE_M, 2 * 4608 + 600
    m, q = _worst_case(kind, rng, M, Q)
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx, d2 = nn.best_match(q, 0.1, want_d2=True)
    launches = nn.last_launches()
    nn.close()
    assert [l[:2] for l in launches] == [(12, 384), (2, 384)], launches
    oi, od = oracle.nn_best_match(m, q, 0.1)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2[oi >= 0], od[oi >= 0])
    assert (oi >= 0).sum() > Q // 3


@pytest.mark.parametrize("kind", ["constant_prefix", "clustered"])
def test_partial_filter_worst_cases_stay_exact(vo, oracle, kind):
    """The streaming filter only looks at the first 5 of the 10 dimensions (a lower bound of the
    distance); data on which that bound prunes nothing — every row equal to the query in those
    dimensions, or all rows in one tight cluster — must only cost re-scans, never exactness."""
    rng = np.random.RandomState(11)
    M, Q = 6000, 900
    m, q = _worst_case(kind, rng, M, Q)
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx, d2 = nn.best_match(q, 0.1, want_d2=True)
    nn.close()
    oi, od = oracle.nn_best_match(m, q, 0.1)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2[oi >= 0], od[oi >= 0])
    assert (oi >= 0).sum() > Q // 3


# ---- the tensor-core filter (csrc/nn_tc.cu): maps >= 32768 rows, batches >= 2048 queries ------------
def _is_tc(launches):
    return len(launches) == 1 and launches[0][0] == 0  # queries per thread == 0 marks nn_tc_filter_kernel


@pytest.mark.parametrize("M,Q", [(32768, 2048), (38000, 2049), (38000, 5000), (50001, 12500),
                                 (33000, 30000), (70000, 2 * 4608 + 600)])
def test_tensor_core_filter_vs_oracle(vo, oracle, synth, M, Q):
    """tcgen05 f16 filter + exact FP32 re-rank: indices AND d2 equal to the oracle's; the launch log
    proves the tensor-core kernel answered.  Sizes cover 1..N query groups, ragged last query tiles
    and ragged last map tiles."""
    m = synth.nn_map_rows_np(0, M)
    q, target = synth.nn_queries_np(Q, M)
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx, d2 = nn.best_match(q, 0.1, want_d2=True)
    launches, rescans = nn.last_launches(), nn.last_rescans()
    nn.close()
    assert _is_tc(launches), launches
    oi, od = oracle.nn_best_match(m, q, 0.1)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2[oi >= 0], od[oi >= 0])
    exact = (np.arange(Q) % 4) < 2
    assert np.array_equal(idx[exact], target[exact])
    # the 10-D filter is sharp: little beyond the true matches is ever re-ranked
    assert (oi >= 0).sum() <= rescans <= 3 * Q


@pytest.mark.parametrize("kind", ["constant_prefix", "clustered"])
def test_tensor_core_filter_worst_case_data_stays_exact(vo, oracle, kind):
    """data on which the FP32 partial-distance filter prunes nothing; the full-distance tensor-core
    filter must stay exact on it as well (and re-scan only what is close in all ten dimensions)"""
    rng = np.random.RandomState(17)
    M, Q = WIDE_M, 2 * 4608 + 600
    m, q = _worst_case(kind, rng, M, Q)
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx, d2 = nn.best_match(q, 0.1, want_d2=True)
    launches = nn.last_launches()
    nn.close()
    assert _is_tc(launches), launches
    oi, od = oracle.nn_best_match(m, q, 0.1)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2[oi >= 0], od[oi >= 0])
    assert (oi >= 0).sum() > Q // 3


def test_tensor_core_filter_matches_at_the_edge_of_the_radius(vo, oracle):
    """every query has exactly one row at a distance within a few ulp of the radius (inside or
    outside), in a random direction: the f16 filter's margin must never hide an acceptable row, and
    the strict `<` of brute_force_search.h:35 must be decided by the reference-order distance."""
    rng = np.random.RandomState(19)
    M, Q = 40000, 4096
    m = rng.uniform(-1, 1, (M, 11)).astype(np.float32)
    q = rng.uniform(-1, 1, (Q, 11)).astype(np.float32)
    rows = rng.choice(M, Q, replace=False)
    d = rng.normal(size=(Q, 10))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    scale = 0.1 * (1.0 + rng.randint(-8, 9, Q) * 2.0 ** -22)
    q[:, 1:] = (m[rows, 1:].astype(np.float64) + d * scale[:, None]).astype(np.float32)
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx, d2 = nn.best_match(q, 0.1, want_d2=True)
    assert _is_tc(nn.last_launches())
    nn.close()
    oi, od = oracle.nn_best_match(m, q, 0.1)
    assert 0.2 * Q < (oi >= 0).sum() < 0.8 * Q  # both sides of the edge are populated
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2[oi >= 0], od[oi >= 0])


@pytest.mark.parametrize("scale,norm", [(8.0, 1.5), (0.05, 0.02), (1.0, 0.35)])
def test_tensor_core_filter_f16_accumulators_other_magnitudes(vo, oracle, scale, norm):
    """the accumulators of the tensor filter are f16 and compared as 16-bit codes: coordinates far from
    the unit cube move both the size of an f16 ulp at the threshold and the tensor core's internal
    accumulation error (tools/tc_probe2.cu).  Every query has a row at 0, 0.5, 0.999.. or 1.001..
    radii; indices and d2 must still equal the oracle's."""
    rng = np.random.RandomState(29)
    M, Q = 40000, 4096
    m = (scale * rng.uniform(-1, 1, (M, 11))).astype(np.float32)
    q = (scale * rng.uniform(-1, 1, (Q, 11))).astype(np.float32)
    rows = rng.choice(M, Q, replace=False)
    d = rng.normal(size=(Q, 10))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    frac = np.choose(np.arange(Q) % 4, [0.0, 0.5, 1.0 - 3e-6, 1.0 + 3e-6])
    q[:, 1:] = (m[rows, 1:].astype(np.float64) + d * (norm * frac)[:, None]).astype(np.float32)
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx, d2 = nn.best_match(q, norm, want_d2=True)
    assert _is_tc(nn.last_launches())
    nn.close()
    oi, od = oracle.nn_best_match(m, q, norm)
    assert (oi >= 0).sum() >= Q // 2
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2[oi >= 0], od[oi >= 0])


@pytest.mark.parametrize("Q", [1, 100, 129, 300, 1100])
def test_tensor_core_filter_forced_on_small_batches(vo, oracle, synth, monkeypatch, Q):
    """VO_NN_FORCE_PATH=tc: batches of 1..9 query tiles (one, two, three ... resident tiles per group,
    i.e. accumulator buffers that are used by every map tile, every second one, or unevenly)"""
    monkeypatch.setenv("VO_NN_FORCE_PATH", "tc")
    M = 33000
    m = synth.nn_map_rows_np(0, M)
    q, target = synth.nn_queries_np(Q, M)
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx, d2 = nn.best_match(q, 0.1, want_d2=True)
    assert _is_tc(nn.last_launches()), nn.last_launches()
    nn.close()
    oi, od = oracle.nn_best_match(m, q, 0.1)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2[oi >= 0], od[oi >= 0])


def test_tensor_core_filter_ties_large_radius_and_odd_queries(vo, oracle, monkeypatch):
    rng = np.random.RandomState(23)
    M = 33000
    m = rng.uniform(-1, 1, (M, 11)).astype(np.float32)
    dup = rng.choice(20000, 64, replace=False)
    m[20000 + np.arange(64)] = m[dup]            # duplicate rows: the lowest index must win
    m[32990:33000] = m[dup[0]]                   # and an 11-way tie ending in the ragged last tile
    q = rng.uniform(-1, 1, (2100, 11)).astype(np.float32)
    q[:64] = m[dup]
    q[64] = m[dup[0]]
    q[65, 1:] = np.nan                           # non-finite query: no row can match
    q[66, 3] = np.inf
    q[67, 1:] = 300.0                            # |q|^2 too large for f16: exact re-scan of every block
    q[68, 1:] = 0.0
    nn = vo.NNIndex(0)
    nn.set_map(m)
    for norm in (0.1, 0.8):                      # 0.8: hundreds of candidates per query, bound tightening
        idx, d2 = nn.best_match(q, norm, want_d2=True)
        assert _is_tc(nn.last_launches())
        oi, od = oracle.nn_best_match(m, q, norm)
        assert np.array_equal(idx, oi)
        assert np.array_equal(d2[oi >= 0], od[oi >= 0])
        assert np.array_equal(idx[:64], dup) and idx[64] == dup[0]
        assert idx[65] == idx[66] == idx[67] == -1
    nn.close()
    # a radius for which the f16 margin would swamp the bound falls back to the FP32 filter
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx = nn.best_match(q, 0.01)
    assert not _is_tc(nn.last_launches())
    assert np.array_equal(idx, oracle.nn_best_match(m, q, 0.01)[0])
    nn.close()


def test_both_filters_agree_at_scale(vo, synth, monkeypatch):
    """1e5 queries x 1e6 rows through the tensor-core filter and through the FP32 filter: identical
    indices and d2 (both are decided by the same exact re-rank)."""
    import torch

    M, Q = 1_000_000, 100_000
    dev = torch.device("cuda:0")
    m = synth.nn_map_torch(M, dev)
    qn, _ = synth.nn_queries_np(Q, M)
    q = torch.from_numpy(qn).to(dev)
    out = {}
    for path in ("tc", "ffma"):
        monkeypatch.setenv("VO_NN_FORCE_PATH", path)
        idx = torch.empty(Q, dtype=torch.int32, device=dev)
        d2 = torch.empty(Q, dtype=torch.float32, device=dev)
        nn = vo.NNIndex(0)
        nn.set_stream(torch.cuda.current_stream().cuda_stream)
        nn.set_map_device(m.data_ptr(), M, 11, 1)
        nn.best_match_device(q.data_ptr(), Q, 11, 0.1, idx.data_ptr(), d2.data_ptr())
        torch.cuda.synchronize()
        assert _is_tc(nn.last_launches()) == (path == "tc")
        out[path] = (idx.cpu().numpy(), d2.cpu().numpy())
        nn.close()
    assert np.array_equal(out["tc"][0], out["ffma"][0])
    hit = out["tc"][0] >= 0
    assert np.array_equal(out["tc"][1][hit], out["ffma"][1][hit])


@pytest.mark.parametrize("nq", [1, 3, 8])
def test_few_queries_single_launch_path(vo, oracle, nq):
    """what the reference's main does (vo_complete.cpp:37-38: one bestMatchFull per measurement): up
    to 8 queries against a frame-sized map are answered by ONE kernel launch (queries as kernel
    parameters, answers through mapped host memory); same bits as the oracle, incl. ties and misses"""
    rng = np.random.RandomState(31 + nq)
    m = rng.uniform(-1, 1, (900, 11)).astype(np.float32)
    m[500] = m[20]                                   # duplicate: the lower index wins
    q = rng.uniform(-1, 1, (nq, 11)).astype(np.float32)
    q[0] = m[500]
    if nq > 1:
        q[1, 1:] = m[77, 1:] + np.float32(0.01)      # a near match
    nn = vo.NNIndex(0)
    nn.set_map(m)
    for norm in (0.1, 3.0):
        idx, d2 = nn.best_match(q, norm, want_d2=True)
        assert nn.last_launches() == [(1, 256, nq, 1)]
        oi, od = oracle.nn_best_match(m, q, norm)
        assert np.array_equal(idx, oi) and idx[0] == 20
        assert np.array_equal(d2[oi >= 0], od[oi >= 0])
    nn.close()


def test_large_radius_true_argmin(vo, oracle):
    """radius large enough that EVERY row is a candidate: exercises the bound-tightening path."""
    rng = np.random.RandomState(5)
    m = rng.uniform(-1, 1, (3000, 11)).astype(np.float32)
    q = rng.uniform(-1, 1, (700, 11)).astype(np.float32)
    nn = vo.NNIndex(0)
    nn.set_map(m)
    for norm in (0.8, 2.0, 100.0):
        idx, d2 = nn.best_match(q, norm, want_d2=True)
        oi, od = oracle.nn_best_match(m, q, norm)
        assert np.array_equal(idx, oi)
        assert np.array_equal(d2[oi >= 0], od[oi >= 0])
    nn.close()


def test_ties_lowest_index_and_no_match_and_edge(vo, oracle):
    rng = np.random.RandomState(3)
    m = rng.uniform(-1, 1, (4096, 11)).astype(np.float32)
    dup = rng.choice(2000, 64, replace=False)
    m[2000 + np.arange(64)] = m[dup]      # 64 duplicate pairs
    m[4000:4010] = m[dup[0]]              # and a 12-way tie
    q = np.concatenate([m[dup], m[4000:4001], np.full((3, 11), 7.0, np.float32)])
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx = nn.best_match(q, 0.1)
    oi, _ = oracle.nn_best_match(m, q, 0.1)
    assert np.array_equal(idx, oi)
    assert np.array_equal(idx[:64], dup)
    assert idx[64] == dup[0]
    assert idx[-3:].tolist() == [-1, -1, -1]
    # d2 exactly == norm^2 is NOT a match (strict '<', brute_force_search.h:35)
    one = np.zeros((1, 11), np.float32)
    qq = np.zeros((1, 11), np.float32)
    qq[0, 1] = 0.5
    nn.set_map(one)
    assert nn.best_match(qq, 0.5).tolist() == [-1]
    assert nn.best_match(qq, float(np.nextafter(np.float32(0.5), np.float32(1)))).tolist() == [0]
    nn.close()


def test_near_ties_within_rounding(vo, oracle):
    """rows a few ulp apart around a query: the winner must follow the reference's summation
    order, not the FMA filter's."""
    rng = np.random.RandomState(9)
    base = rng.uniform(-1, 1, 10).astype(np.float32)
    rows = np.tile(base, (512, 1))
    rows += (rng.randint(-3, 4, rows.shape) * np.float32(2.0 ** -24)).astype(np.float32)
    m = _vec11(rows)
    q = _vec11(np.tile(base, (64, 1)) + (rng.randint(-3, 4, (64, 10)) * np.float32(2.0 ** -23)).astype(np.float32))
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx, d2 = nn.best_match(q, 0.1, want_d2=True)
    nn.close()
    oi, od = oracle.nn_best_match(m, q, 0.1)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2, od)


def test_empty_inputs(vo):
    nn = vo.NNIndex(0)
    nn.set_map(np.zeros((0, 11), np.float32))
    assert nn.best_match(np.zeros((5, 11), np.float32), 0.1).tolist() == [-1] * 5
    nn.set_map(np.zeros((10, 11), np.float32))
    assert nn.best_match(np.zeros((0, 11), np.float32), 0.1).shape == (0,)
    nn.close()


@pytest.mark.parametrize("dim", [2, 3, 5, 16])
def test_general_dimension(vo, oracle, dim):
    rng = np.random.RandomState(dim)
    m = rng.uniform(-1, 1, (2000, dim + 1)).astype(np.float32)
    q = np.concatenate([m[rng.choice(2000, 100)], rng.uniform(-1, 1, (100, dim + 1)).astype(np.float32)])
    nn = vo.NNIndex(0)
    nn.set_map(m)
    idx, d2 = nn.best_match(q, 0.3, want_d2=True)
    nn.close()
    oi, od = oracle.nn_best_match(m, q, 0.3)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2[oi >= 0], od[oi >= 0])


def test_radius_search(vo, oracle):
    rng = np.random.RandomState(4)
    m = rng.uniform(-0.3, 0.3, (3000, 11)).astype(np.float32)
    q = rng.uniform(-0.3, 0.3, (37, 11)).astype(np.float32)
    nn = vo.NNIndex(0)
    nn.set_map(m)
    counts, lst = nn.radius_search(q, 0.6, 256)
    nn.close()
    oc, ol = oracle.nn_radius_search(m, q, 0.6, 256)
    assert np.array_equal(counts, oc)
    for i in range(len(q)):
        k = min(counts[i], 256)
        assert np.array_equal(lst[i, :k], ol[i, :k])
    got = vo.bruteForceSearch(m, q[0], 0.6)
    assert np.array_equal(got, ol[0, :oc[0]])
    assert vo.bruteForceBestMatch(m, q[0], 0.6) == oracle.nn_best_match(m, q[:1], 0.6)[0][0]


def test_device_resident_path_and_linearity_at_scale(vo, synth):
    """Full-size property test (oracle-free): 1e5 queries vs a 1e6-row device-generated map;
    every exact planted copy must come back as its own row (or a lower-index duplicate at d2=0),
    and answers must not depend on how the query batch is split."""
    import torch

    M, Q = 1_000_000, 100_000
    dev = torch.device("cuda:0")
    m = synth.nn_map_torch(M, dev)
    qn, target = synth.nn_queries_np(Q, M)
    q = torch.from_numpy(qn).to(dev)
    idx = torch.empty(Q, dtype=torch.int32, device=dev)
    d2 = torch.empty(Q, dtype=torch.float32, device=dev)
    nn = vo.NNIndex(0)
    nn.set_stream(torch.cuda.current_stream().cuda_stream)
    nn.set_map_device(m.data_ptr(), M, 11, 1)
    nn.best_match_device(q.data_ptr(), Q, 11, 0.1, idx.data_ptr(), d2.data_ptr())
    torch.cuda.synchronize()
    got = idx.cpu().numpy()
    exact = (np.arange(Q) % 4) < 2
    assert np.array_equal(got[exact], target[exact])
    noisy = (np.arange(Q) % 4) == 2
    assert np.array_equal(got[noisy], target[noisy])
    assert np.all(got[(np.arange(Q) % 4) == 3] == -1)
    # split invariance
    idx2 = torch.empty(Q, dtype=torch.int32, device=dev)
    h = Q // 3
    nn.best_match_device(q.data_ptr(), h, 11, 0.1, idx2.data_ptr())
    nn.best_match_device(q[h:].data_ptr(), Q - h, 11, 0.1, idx2[h:].data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(idx, idx2)
    nn.close()
