"""CPU: the oracle (oracle/qa_oracle.py) against the golden fixtures made from the reference."""
import numpy as np
import pytest

from oracle import qa_oracle as orc
from tests import golden_util as G


@pytest.mark.parametrize("name", G.kat_cases())
def test_formats_bit_exact(name):
    x, outs = G.kat(name)
    with np.errstate(all="ignore"):
        for fmt in G.FORMATS:
            got = G.bits(orc.quantize(x, fmt))
            assert np.array_equal(got, outs[fmt]), f"{name}/{fmt}"


def test_appendix_c_hex_vectors():
    # SURVEY.md Appendix C, bfp4 row of the 16-element known-answer group
    x, outs = G.kat("appendix_c_row")
    want = "3f800000 3f800000 3f800000 3f000000 3f000000 bf400000 00000000 00000000 bfe00000 3fe00000 " \
           "00000000 00000000 00000000 3e800000 be800000 00000000"
    assert [f"{v:08x}" for v in outs["bfp4"]] == want.split()
    assert [f"{v:08x}" for v in G.bits(orc.quantize(x, "bfp4"))] == want.split()


def test_numpy_arith_restatements_match_numpy():
    rng = np.random.default_rng(5)
    a = rng.standard_normal((50, 1024)).astype(np.float32)
    b = (a + rng.standard_normal((50, 1024)).astype(np.float32) * 0.01).astype(np.float32)
    assert np.array_equal(orc.np_pairwise_sum(a), np.add.reduce(a, axis=1))
    for n in (7, 100, 130, 1000):
        v = rng.standard_normal((3, n)).astype(np.float32)
        assert np.array_equal(orc.np_pairwise_sum(v), np.add.reduce(v, axis=1))
    d64 = a.astype(np.float64) * 1e3
    assert np.array_equal(orc.np_pairwise_sum(d64), np.add.reduce(d64, axis=1))
    from threadpoolctl import threadpool_info
    arch = {d.get("architecture") for d in threadpool_info() if d.get("internal_api") == "openblas"}
    if arch and arch != {"SkylakeX"}:
        pytest.skip(f"np.dot uses a different OpenBLAS kernel here: {arch}")
    want = np.array([np.dot(a[i], b[i]) for i in range(a.shape[0])], dtype=np.float32)
    assert np.array_equal(orc.np_sdot_f32(a, b), want)
    rows = orc.pearson_f32_rows(a, b)
    want = np.array([orc.pearson_f32(a[i], b[i]) for i in range(a.shape[0])], dtype=np.float32)
    assert np.array_equal(rows, want)


@pytest.mark.parametrize("name", G.algo_case_names())
def test_tile_scores_bit_exact(name):
    x = G.algo_input(name)
    z = G.npz("algo_small.npz")
    for metric in ("pcc", "mae", "atol"):
        sc = orc.padded_tile_scores(x, G.MIXED, metric)
        for fmt in G.MIXED:
            want = z[f"{name}__tilescore__{fmt}__{metric}"]
            assert np.array_equal(sc[fmt].view(np.uint32), want.view(np.uint32)), (name, fmt, metric)


def _formats_for(params):
    raw = params.get("formats")
    if raw:
        return [p.strip() for p in raw.split(",")]
    return list(G.MIXED)


@pytest.mark.parametrize("name", G.algo_case_names())
def test_algorithms_match_reference(name):
    x = G.algo_input(name)
    table = orc.tile_stat_table(x)
    for key, m, want_assign, want_y in G.algo_runs(name):
        p, fmts = m["params"], _formats_for(m["params"])
        if m["algo"] == "mixed-tile-greedy":
            a, counts = orc.greedy_assign(table, fmts, p["metric"], p["threshold"], p["seed"])
        elif m["algo"] == "mixed-tile-threshold":
            sc = orc.padded_tile_scores(x, fmts, p["metric"])
            a, counts = orc.threshold_assign(sc, fmts, p["metric"], p["threshold"], table["th"], table["tw"])
        else:
            a, counts, samples = orc.random_assign(x, fmts, p["metric"], p["threshold"], p["iters"], p["seed"])
            for s, w in zip(samples, m["samples"]):
                assert s["counts"] == w["counts"] and s["total_bytes"] == w["total_bytes"]
                assert s["pcc"] == w["pcc"] and s["mae"] == w["mae"] and s["atol"] == w["atol"]
        assert np.array_equal(a, want_assign), key
        assert counts == m["counts"], key
        assert orc.total_bytes(counts) == m["tile_bytes"], key
        y = orc.apply_assignment(x, a)
        assert np.array_equal(G.bits(y), want_y.reshape(-1)), key


def test_cfg1_shape_goldens():
    from quantization_analysis_b200 import synthetic
    import hashlib
    meta, z = G.js("cfg1_q_a_proj.json"), G.npz("cfg1_q_a_proj.npz")
    x = synthetic.randn_f32_np((1536, 7168), 0)
    assert hashlib.sha256(x.tobytes()).hexdigest() == meta["input_sha256"]
    for fmt in G.FORMATS:
        y = orc.quantize(x, fmt)
        assert hashlib.sha256(y.tobytes()).hexdigest() == meta["none"][fmt]["y_sha256"]
        ex = orc.exact_metrics_f64(x, y)
        # the reference's float32 pcc is only ~1e-5 accurate; mae/atol are tight
        assert abs(ex["pcc"] - meta["none"][fmt]["pcc_f32"]) < 5e-5
        assert ex["mae"] == pytest.approx(meta["none"][fmt]["mae_f32"], rel=2e-6, abs=0)
        assert ex["atol"] == meta["none"][fmt]["atol_f32"]
    table = orc.tile_stat_table(x)
    a, counts = orc.greedy_assign(table, G.MIXED, "pcc", 0.999, 123)
    assert np.array_equal(a, z["greedy_pcc0999_seed123"])
    assert counts == meta["greedy_pcc0999_seed123"]["counts"]


def test_pcg64_restatement_matches_numpy_streams():
    z = G.npz("numpy_rng.npz")
    r = orc.Pcg64(123)
    assert np.array_equal(r.permutation(10752), z["s123__perm_10752"])
    assert np.array_equal(r.permutation(777), z["s123__perm_777"])
    assert np.array_equal(r.integers(4, 1000), z["s123__int4_1000"])
    assert np.array_equal(r.integers(3, 1000), z["s123__int3_1000"])
    assert np.array_equal(r.permutation(4096), z["s123__perm_4096"])
    assert np.array_equal(r.integers(2, 777), z["s123__int2_777"])
    assert np.array_equal(r.permutation(2), z["s123__perm_2"])
    assert np.array_equal(r.permutation(1), z["s123__perm_1"])


def test_scalar_proxies_match_reference_goldens():
    """mxfp4 / nvfp4 scalar proxies: oracle vs the reference on every bf16 pattern and 20 000 random float32 values."""
    z = G.npz("scalar_proxies.npz")
    x16 = (np.arange(65536, dtype=np.uint32) << 16).view(np.float32)
    xr = z["rand__in"].view(np.float32)
    for fmt in ("mxfp4", "nvfp4"):
        with np.errstate(all="ignore"):
            got16, gotr = orc.quantize(x16, fmt), orc.quantize(xr, fmt)
        w16 = z[f"bf16__{fmt}"]
        nan = np.isnan(w16.view(np.float32))
        assert np.array_equal(np.isnan(got16), nan)
        assert np.array_equal(got16.view(np.uint32)[~nan], w16[~nan]), fmt
        assert np.array_equal(gotr.view(np.uint32), z[f"rand__{fmt}"]), fmt


def test_fp8_block_dequant_matches_reference_goldens():
    """fp8 e4m3fn + scale_inv block dequantization (hf_model_utils.py:199-215): oracle vs the reference through torch."""
    z = G.npz("fp8_dequant.npz")
    for tag in ("all", "rag"):
        got = orc.fp8_block_dequant(z[f"{tag}__w"], z[f"{tag}__s"])
        want = z[f"{tag}__out"].view(np.float32).reshape(got.shape)
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan)
        assert np.array_equal(got.view(np.uint32)[~nan], want.view(np.uint32)[~nan]), tag


def test_restated_whole_tensor_orders_match_numpy():
    """oracle.np_sdot_f32 / pearson_f32_restated / wq_scores_restated (explicit restatement of NumPy's pairwise sums and
    OpenBLAS's sdot incl. the 32-element block and the double-accumulated tail) == np.dot / np.mean on this image, for
    every tail case.  The CUDA scorer is checked against the restatement, so this pins it to the reference's arithmetic."""
    rng = np.random.default_rng(12)
    for n in [1, 2, 7, 8, 31, 32, 33, 63, 64, 65, 95, 96, 100, 129, 1000, 1023, 1056, 4113, 10007, 50083, 96 * 160]:
        x = (rng.standard_normal(n) * 0.02).astype(np.float32)
        y = orc.quantize(x.reshape(1, -1), "bfp4").reshape(-1)
        assert np.float32(np.dot(x, y)).view(np.uint32) == np.float32(orc.np_sdot_f32(x, y)).view(np.uint32), n
        assert orc.pearson_f32(x, y) == orc.pearson_f32_restated(x, y), n
        d = np.abs(x - y)
        assert float(np.mean(d)) == orc.wq_scores_restated(x, y)["mae"], n


def test_pairwise_plan_reproduces_numpy_add_reduce():
    """qa_pairwise_plan_build (host side of the C ABI): leaves + tree evaluated in NumPy == np.add.reduce."""
    from quantization_analysis_b200 import engine
    rng = np.random.default_rng(1)
    for n in [1, 5, 8, 100, 128, 129, 257, 1000, 4113, 50083, 1536 * 7]:
        a = (rng.standard_normal(n) * 0.02).astype(np.float32)
        plan = engine.pairwise_plan_host(n)
        nl, lv, nn, ni = (int(v) for v in plan[:4])
        loff = plan[4:4 + lv + 1]
        ls = plan[4 + lv + 1:4 + lv + 1 + nl]
        ll = plan[4 + lv + 1 + nl:4 + lv + 1 + 2 * nl]
        lf = plan[4 + lv + 1 + 2 * nl:4 + lv + 1 + 2 * nl + ni]
        rt = plan[4 + lv + 1 + 2 * nl + ni:]
        assert int(ll.sum()) == n and int(ll.max()) <= 128 and nn == nl + ni
        v = np.zeros(nn, np.float32)
        for i in range(nl):
            v[i] = orc.np_pairwise_sum(a[ls[i] * 8: ls[i] * 8 + ll[i]])
        for L in range(lv):
            idx = np.arange(loff[L], loff[L + 1])
            v[nl + idx] = v[lf[idx]] + v[rt[idx]]
        assert v[nn - 1] == np.add.reduce(a), n


def test_cli_goldens_present_and_strip_times():
    from tests import cli_util as U
    for case in U.WQ_CASES:
        assert (U.CLI_GOLDEN / "wq" / case / "table.txt").exists(), case
    for case in U.SWEEP_CASES:
        assert len(list((U.CLI_GOLDEN / "sweep" / case).rglob("sweep_results.csv"))) == 1, case
    a = "  none  BF16   1.00000  0.000e+00  0.000e+00    0.008  0.000"
    b = "  none  BF16   1.00000  0.000e+00  0.000e+00    1.234  0.000"
    assert U.strip_times(a) == U.strip_times(b) and U.strip_times(a) != a
