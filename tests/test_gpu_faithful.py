"""GPU parity of the whole-tensor NumPy-float32-faithful scorer (qa_tensor_scores_f32), the column-group (`transpose`)
reconstruction and the generic pair tile scorer: bit-for-bit against numbers produced by the unmodified reference
(tests/golden) and against the oracle's explicit restatement of NumPy's / OpenBLAS's summation orders."""
import numpy as np
import pytest
import torch

from oracle import qa_oracle as orc
from tests import golden_util as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qa():
    from quantization_analysis_b200 import _lib, engine, quantization_formats as qf
    from quantization_analysis_b200 import compression_algorithms as ca
    _lib.lib()
    return {"engine": engine, "qf": qf, "ca": ca}


def _same(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


@pytest.mark.parametrize("n", [1, 5, 8, 31, 64, 100, 129, 1000, 1023, 4113, 10007, 50083, 96 * 160, 64 * 256 + 17])
def test_tensor_scores_match_restated_orders(qa, n):
    """Every tail case of the pairwise tree (n < 8, n % 8, irregular splits) and of sdot (64-blocks, the 32-block, the
    double-accumulated n % 32 tail), float32 and bf16 operands."""
    eng = qa["engine"]
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 0.02).astype(np.float32)
    y = orc.quantize(x.reshape(1, -1), "bfp4").reshape(-1)
    want = orc.wq_scores_restated(x, y)
    got = eng.tensor_scores_f32(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda())[0]
    assert _same(got[0], want["pcc"]) and _same(got[1], want["mae"]) and _same(got[2], want["atol"]), (n, got, want)
    # bf16 operands (what the device pipeline feeds it) give the same numbers as their float32 images
    xb = torch.from_numpy(x).cuda().to(torch.bfloat16)
    yb = torch.from_numpy(orc.quantize(xb.float().cpu().numpy().reshape(1, -1), "bfp8").reshape(-1)).cuda().to(torch.bfloat16)
    want = orc.wq_scores_restated(xb.float().cpu().numpy(), yb.float().cpu().numpy())
    got = eng.tensor_scores_f32(xb, yb)[0]
    assert _same(got[0], want["pcc"]) and _same(got[1], want["mae"]) and _same(got[2], want["atol"]), (n, "bf16")


def test_tensor_scores_batch_zero_and_degenerate(qa):
    eng = qa["engine"]
    rng = np.random.default_rng(5)
    n = 3000
    x = (rng.standard_normal(n) * 0.02).astype(np.float32)
    ys = np.stack([orc.quantize(x.reshape(1, -1), f).reshape(-1) for f in ("bf16", "bfp8", "bfp4", "bfp2")] + [x])
    got = eng.tensor_scores_f32(torch.from_numpy(x).cuda(), torch.from_numpy(ys).cuda())
    for i in range(ys.shape[0]):
        w = orc.wq_scores_restated(x, ys[i])
        assert _same(got[i, 0], w["pcc"]) and _same(got[i, 1], w["mae"]) and _same(got[i, 2], w["atol"]), i
    # fp0 (y = None = zeros): denominator 0 -> pcc 0.0 (metrics.py:14-15), mae = mean|x|, atol = max|x|
    z = eng.tensor_scores_f32(torch.from_numpy(x).cuda(), None)[0]
    w = orc.wq_scores_restated(x, np.zeros_like(x))
    assert z[0] == 0.0 and _same(z[1], w["mae"]) and _same(z[2], w["atol"])
    # constant tensors: denominator 0 and no error -> 1.0
    c = np.full(777, 0.25, np.float32)
    assert eng.tensor_scores_f32(torch.from_numpy(c).cuda(), torch.from_numpy(c).cuda())[0, 0] == 1.0
    # NaN / inf propagate like np.max / np.mean do
    bad = x.copy()
    bad[17] = np.inf
    got = eng.tensor_scores_f32(torch.from_numpy(bad).cuda(), torch.from_numpy(x).cuda())[0]
    with np.errstate(all="ignore"):
        w = orc.wq_scores_restated(bad, x)
    assert np.isnan(got[0]) == np.isnan(w["pcc"]) and got[2] == w["atol"] == np.inf


def test_cfg1_reference_float32_scores_bit_for_bit(qa):
    """The numbers the reference prints for config 1 (`none` on [1536,7168]): pcc 0.9999874830245972 for bfp8 etc."""
    from quantization_analysis_b200 import synthetic
    eng = qa["engine"]
    meta = G.js("cfg1_q_a_proj.json")
    x = synthetic.randn_bf16_cpu((1536, 7168), 0).cuda()
    p = eng.prepare_rows(x)
    recon = eng.quant_recon(p, G.MIXED)
    ys = torch.stack([recon[f] for f in G.MIXED])
    got = eng.tensor_scores_f32(x, ys)
    for i, f in enumerate(G.MIXED):
        m = meta["none"][f]
        assert float(got[i, 0]) == m["pcc_f32"] and float(got[i, 1]) == m["mae_f32"] and float(got[i, 2]) == m["atol_f32"], f
    z = eng.tensor_scores_f32(x, None)[0]
    m = meta["none"]["fp0"]
    assert float(z[0]) == m["pcc_f32"] and float(z[1]) == m["mae_f32"] and float(z[2]) == m["atol_f32"]


@pytest.mark.parametrize("name", G.algo_case_names())
def test_algo_results_reference_float32_scores(qa, name):
    """pcc / mae / atol the reference computed (wq:684-687) for every recorded algorithm run, from x and its y."""
    eng = qa["engine"]
    x = G.algo_input(name)
    for key, m, _a, ybits in G.algo_runs(name):
        y = G.f32_from_bits(ybits, x.shape)
        got = eng.tensor_scores_f32(torch.from_numpy(x.reshape(-1)).cuda(), torch.from_numpy(y.reshape(-1)).cuda())[0]
        assert float(got[0]) == m["pcc_f32"] and float(got[1]) == m["mae_f32"] and float(got[2]) == m["atol_f32"], key


def test_metrics_module_is_reference_float32(qa):
    from quantization_analysis_b200.compression_algorithms import metrics
    x = G.algo_input("het_96x160")
    y = orc.quantize(x, "bfp4")
    assert metrics.pearson_corr(x, y) == orc.pearson_f32(x, y)
    for k in ("pcc", "mae", "atol"):
        assert metrics.metric_value(x, y, k) == orc.metric_f32(x, y, k)
    assert metrics.pearson_corr(np.zeros((0,), np.float32), np.zeros((0,), np.float32)) == 1.0
    with pytest.raises(ValueError):
        metrics.metric_value(x, y, "psnr")


def _transpose_cases():
    z = G.npz("transpose_small.npz")
    return sorted(k[:-4] for k in z if k.endswith("__in"))


@pytest.mark.parametrize("name", _transpose_cases())
def test_transpose_algorithm_bit_exact_vs_reference(qa, name):
    """compression_algorithms/transpose.py:13-33: shared exponent along axis 0, via the column-group kernel."""
    z = G.npz("transpose_small.npz")
    x = G.f32_from_bits(z[f"{name}__in"], z[f"{name}__shape"])
    res = qa["ca"].create_algorithm("transpose", {}).run(x, G.FORMATS, qa["ca"].quantizer.Quantizer("emulation"), None)
    assert [r.fmt for r in res] == [f.upper() for f in G.FORMATS] and all(r.compression == "transpose" for r in res)
    for r in res:
        assert r.y.shape == x.shape and r.y.dtype == np.float32
        assert np.array_equal(G.bits(r.y), z[f"{name}__{r.fmt.lower()}"].reshape(-1)), (name, r.fmt)
    if x.ndim >= 2:       # device tensors in, device tensors out
        xt = torch.from_numpy(x).cuda()
        rt = qa["ca"].create_algorithm("transpose", {}).run(xt, ["bfp4"], None, None)[0]
        assert rt.y.is_cuda and np.array_equal(G.bits(rt.y.float().cpu().numpy()), z[f"{name}__bfp4"].reshape(-1))


def test_transpose_equals_row_kernel_on_transposed_input(qa):
    """Property at a larger size: column groups of x == row groups of x.T (the reference's definition)."""
    rng = np.random.default_rng(8)
    x = (rng.standard_normal((1000, 777)) * 0.02).astype(np.float32)
    res = qa["ca"].create_algorithm("transpose", {}).run(x, ["bfp8", "bfp4", "bfp2"], None, None)
    for r in res:
        want = qa["qf"].quantize_weight_values(np.ascontiguousarray(x.T), r.fmt.lower()).T
        assert np.array_equal(r.y, want), r.fmt


def test_tile_metrics_accepts_arbitrary_operands(qa):
    """tile_utils.py:46-57 on operands that are NOT a quantization of each other, incl. inf / nan tiles."""
    from quantization_analysis_b200.compression_algorithms import tile_utils
    rng = np.random.default_rng(2)
    ref = (rng.standard_normal((9, 32, 32)) * 0.1).astype(np.float32)
    q = (ref + rng.standard_normal((9, 32, 32)).astype(np.float32) * 1e-3).astype(np.float32)
    q[3] = ref[3]                 # identical tile
    q[4] = 0.0
    ref[5] = 0.5
    q[5] = 0.5                    # constant: denominator 0, no error -> 1.0
    ref[6, 3, 7] = np.inf         # inf - finite = inf; pcc nan
    q[7, 1, 1] = np.nan
    with np.errstate(all="ignore"):
        for metric in ("pcc", "mae", "atol"):
            want = orc.tile_scores_f32(ref, q, metric)
            got = tile_utils.tile_metrics(ref, q, metric)
            assert got.dtype == np.float32 and got.shape == (9,)
            assert np.array_equal(got.view(np.uint32)[:6], want.view(np.uint32)[:6]), metric
            assert np.array_equal(np.isnan(got), np.isnan(want)), metric
            assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)]), metric
