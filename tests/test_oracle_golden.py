"""Pins the CPU oracle (oracle/proto_oracle.py) to the golden fixtures that
tests/golden/make_golden.py produced by running the REAL reference.

Tolerances: prototypes / class sums 1e-5 * max|ref|;
distances 1e-5 of the raw distance scale; soft predictions 1e-5 absolute;
labels bit-exact off near-ties (top-2 margin < 1e-6 or |m - thresh| < 1e-6).
The oracle keeps the reference's evaluation order, so in practice it is
bit-identical and the tests assert that where it holds.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import proto_oracle as po

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
OPS = sorted(glob.glob(os.path.join(GOLDEN, "ops_*.npz")))


def T(a):
    return torch.from_numpy(np.asarray(a))


def load_handler(z):
    h = po.OracleHandler(ma_lambda=float(z["ma_lambda"]), tau=float(z["tau"]), thresh=float(z["thresh"]),
                         distance_metric=str(z["metric"]))
    h.prototypes, h.squared_mean, h.counter = T(z["protos"]).clone(), T(z["sq_mean"]).clone(), T(z["counter"]).clone()
    return h


@pytest.mark.parametrize("path", OPS, ids=[os.path.basename(p)[:-4] for p in OPS])
def test_ops_match_reference(path):
    z = np.load(path)
    h = load_handler(z)
    feat, prior, out = T(z["feat"]), T(z["prior"]), T(z["out"])
    mon = po.OracleMonitor(200, 0.003, "hamming")
    dist = h.distance_measure(feat)
    labels = h.pseudo_labels(feat, prior, confidence_monitor=mon)
    soft = h.pseudo_labels(feat, prior, soft=True)
    assert torch.equal(dist, T(z["ref_dist"]))
    assert torch.equal(labels, T(z["ref_labels"]))
    assert labels.dtype == torch.int64 and labels.shape == (feat.shape[0] * feat.shape[2] * feat.shape[3], 1)
    np.testing.assert_allclose(soft.numpy(), z["ref_soft"], rtol=0, atol=1e-7)
    assert float(mon.window["prototypes"][0]) == pytest.approx(float(z["ref_stat_proto"]), abs=1e-7)
    assert po.mean_max(prior) == pytest.approx(float(z["ref_stat_prior"]), abs=1e-7)
    s1, cnt = h.get_proto_array(feat, out)
    s2, _ = h.get_proto_array(feat ** 2, out)
    for got, ref in ((s1, z["ref_sum"]), (s2, z["ref_sumsq"])):
        np.testing.assert_allclose(got.numpy(), ref, rtol=0, atol=1e-5 * np.abs(ref).max())
    assert torch.equal(cnt, T(z["ref_count"]))
    np.testing.assert_allclose(h.global_var().numpy(), z["ref_global_std"], rtol=1e-6, equal_nan=True)
    np.testing.assert_allclose(h.prototype_var().numpy(), z["ref_class_std"], rtol=1e-6, equal_nan=True)
    h.ma(feat, out)
    for got, ref in ((h.prototypes, z["ref_ma_protos"]), (h.squared_mean, z["ref_ma_sq_mean"])):
        np.testing.assert_allclose(got.numpy(), ref, rtol=0, atol=1e-5 * np.abs(ref).max())


def test_append_sequence():
    z = np.load(os.path.join(GOLDEN, "append_seq.npz"))
    h = po.OracleHandler(distance_metric="mahalanobis")
    for i in range(3):
        h.append(T(z[f"feat{i}"]), T(z[f"out{i}"]))
    h.append(T(z["rows3"]), T(z["hot3"]))
    for got, ref in ((h.prototypes, z["ref_protos"]), (h.squared_mean, z["ref_sq_mean"])):
        np.testing.assert_allclose(got.numpy(), ref, rtol=0, atol=1e-5 * np.abs(ref).max())
    assert torch.equal(h.counter, T(z["ref_counter"]))


@pytest.mark.parametrize("tag", ["hamming", "median", "mean"])
def test_monitor_and_selectors(tag):
    z = np.load(os.path.join(GOLDEN, f"monitor_trace_{tag}.npz"))
    mon = po.OracleMonitor(int(z["limit"]), float(z["exp_const"]), str(z["dev_func"]))
    sel = po.OracleHybridSelect(0, tuple(z["gray"]), float(z["dev_thresh"]))
    vsel = po.OracleDevSelect(0, float(z["vthresh"]))
    cur, cur_dev, dev, med, ema, vcur, pct = [], [], [], [], [], [], []
    for v in z["conf"]:
        mon.add({"prior static": float(v)})
        d = mon.dev_avg("prior static")
        sel.evaluate(mon.avg("prior static"), d)
        vsel.evaluate(d)
        cur.append(sel.current); cur_dev.append(sel.current_dev); dev.append(d)
        med.append(mon.avg("prior static")); ema.append(mon.exp("prior static")); vcur.append(vsel.current)
        pct.append(po.hswitch_percentage(mon.avg("prior static"), True))
    assert np.array_equal(cur, z["ref_current"])
    assert np.array_equal(cur_dev, z["ref_current_dev"])
    assert np.array_equal(vcur, z["ref_vcurrent"])
    assert np.array_equal(np.array(dev, dtype=np.float64), z["ref_dev"])
    assert np.array_equal(np.array(med), z["ref_median"])
    assert np.array_equal(np.array(ema), z["ref_exp"])
    assert np.array_equal(np.array(pct), z["ref_pct"])
    empty = po.OracleMonitor(5)
    assert empty.avg("nope") == z["ref_missing_avg"] and empty.exp("nope") == z["ref_missing_exp"]
    assert empty.dev_avg("nope") == z["ref_missing_dev"]


def test_sequence_of_steps():
    z = np.load(os.path.join(GOLDEN, "sequence_ma.npz"))
    steps, d = int(z["steps"]), int(z["d"])
    first = po.synth_case(4999, 1, d, 9, 11)
    h = po.OracleHandler(ma_lambda=0.95, tau=1.0, thresh=0.3, distance_metric="mahalanobis")
    h.prototypes, h.squared_mean, h.counter = first["protos"].clone(), first["sq_mean"].clone(), first["counter"].clone()
    mon = po.OracleMonitor(50, 0.003, "hamming")
    for i in range(steps):
        case = po.synth_case(5000 + i, 1, d, 9, 11, protos=first["protos"] + 0.002 * i, counter=first["counter"])
        labels, soft = po.fused_step(h, case["feat"], case["prior"], case["out"], mon)
        assert int((labels != 255).sum()) == int(z["ref_npl"][i])
        assert int((labels.flatten() * torch.arange(1, labels.numel() + 1)).sum()) == int(z["ref_label_hash"][i])
        assert float(mon.window["prototypes"][-1]) == pytest.approx(float(z["ref_stat"][i]), abs=1e-7)
    np.testing.assert_allclose(h.prototypes.numpy(), z["ref_protos"], rtol=0, atol=1e-5 * np.abs(z["ref_protos"]).max())
    np.testing.assert_allclose(h.squared_mean.numpy(), z["ref_sq_mean"], rtol=0, atol=1e-5 * np.abs(z["ref_sq_mean"]).max())
    assert mon.dev_avg("prototypes") == pytest.approx(float(z["ref_dev_proto"]), abs=1e-9)


def test_switch_statistics_and_entropy():
    z = np.load(os.path.join(GOLDEN, "stats_logits.npz"))
    la, lb, lc = T(z["la"]), T(z["lb"]), T(z["lc"])
    got = [po.mean_max_softmax(x) for x in (la, lb, lc)]
    np.testing.assert_allclose(got, z["ref_conf"], atol=1e-7)
    mix = 0.25 * la.softmax(1) + 1.0 * lb.softmax(1)
    np.testing.assert_allclose(mix.numpy(), z["ref_mix"], atol=1e-7)
    np.testing.assert_allclose(po.normalised_entropy(la.softmax(1)).numpy(), z["ref_entropy"], atol=1e-7)


def test_degenerate_rows_match_reference():
    """Zero prior row (0 / 0), NaN feature vector, negative tau: the restatement gives what the real reference gave
    (labels bit-exact, NaN rows in the same places, the rest within 1e-7) -- golden/edge_nan_rows.npz."""
    z = np.load(os.path.join(GOLDEN, "edge_nan_rows.npz"))
    for tag, tau in (("pos", 1.0), ("neg", -0.8)):
        h = po.OracleHandler(ma_lambda=0.9995, tau=tau, thresh=0.3, distance_metric="mahalanobis")
        h.prototypes, h.squared_mean, h.counter = T(z["protos"]).clone(), T(z["sq_mean"]).clone(), T(z["counter"]).clone()
        labels = h.pseudo_labels(T(z["feat"]), T(z["prior"]))
        soft = h.pseudo_labels(T(z["feat"]), T(z["prior"]), soft=True)
        assert torch.equal(labels, T(z[f"ref_labels_{tag}"]))
        ref = z[f"ref_soft_{tag}"]
        assert np.array_equal(np.isnan(soft.numpy()), np.isnan(ref))
        assert sorted(np.isnan(ref).all(axis=1).nonzero()[0].tolist()) == [7, 13]
        assert labels[7].item() == 0 and labels[13].item() == 0
        np.testing.assert_allclose(np.nan_to_num(soft.numpy()), np.nan_to_num(ref), rtol=0, atol=1e-7)


def test_bad_metric_raises():
    with pytest.raises(ValueError):
        po.OracleHandler(distance_metric="cosine")


def test_fp64_truth_bounds_reference_error():
    """The fp32 reference's own error against an fp64 evaluation: documents the
    noise floor the tolerances above were derived from."""
    z = np.load(os.path.join(GOLDEN, "ops_mahal_d256_legacy.npz"))
    feat, P, S, c = (T(z[k]).double() for k in ("feat", "protos", "sq_mean", "counter"))
    truth = po.raw_distance(feat, P, po.pooled_std(P, S, c))
    shifted = po.shift_by_row_min(truth)
    err = (shifted.float() - T(z["ref_dist"])).abs() / truth.float()
    assert err.max() < 5e-6


def test_widened_rows_match_reference():
    """update_ema (prototypes.py:407-416) and the evaluation inner loop (adaptation_model.py:143-160, func.py:77-85) as the
    REAL reference computed them (tests/golden/make_golden.py::widened_case): the restatements are bit-identical."""
    import numpy as np
    z = np.load(os.path.join(GOLDEN, "widened_rows.npz"))
    n_p, n_b = int(z["n_params"]), int(z["n_buffers"])
    pq = [torch.from_numpy(z[f"q{i}"]) for i in range(n_p)]
    pk = [torch.from_numpy(z[f"k{i}_before"]) for i in range(n_p)]
    bq = [torch.from_numpy(z[f"bq{i}"]) for i in range(n_b)]
    bk = [torch.from_numpy(z[f"bk{i}_before"]) for i in range(n_b)]
    for _ in range(3):
        pk, bk = po.update_ema(pq, pk, bq, bk, float(z["ema_update"]))
    for i in range(n_p):
        assert torch.equal(pk[i], torch.from_numpy(z[f"k{i}_after3"]))
    for i in range(n_b):
        want = torch.from_numpy(z[f"bk{i}_after3"])
        assert bk[i].dtype == want.dtype and torch.equal(bk[i], want)
    pred, labels = torch.from_numpy(z["eval_pred"]), torch.from_numpy(z["eval_labels"])
    prob, arg, hist = po.eval_confusion(pred, labels, 19, tuple(labels.shape[1:]))
    assert torch.equal(prob, torch.from_numpy(z["eval_prob"]))
    assert np.array_equal(arg.numpy(), z["eval_argmax"]) and np.array_equal(hist, z["eval_hist"])
    iu = np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist) + np.finfo(float).eps)
    assert np.array_equal(iu, z["eval_iu"])


@pytest.mark.parametrize("fixture", ["method_hswitch_soft.npz", "method_hswitch_hard.npz"])
def test_hswitch_share_restatement_matches_reference_trace(fixture):
    """The h-switch ramp / threshold of the product (switching.static_share) and of the oracle reproduce the share the real
    hswitch_proDA recorded at every step, given the reference's own static confidences."""
    from onda_b200.switching import Monitor, static_share
    z = np.load(os.path.join(GOLDEN, fixture))
    mon = Monitor(int(z["limit"]), float(z["exp_const"]), "hamming")
    for i, conf in enumerate(z["ref_stat_prior_static"]):
        mon.add({"prior static": float(conf)})
        share = static_share(mon.avg("prior static"), bool(int(z["soft_trans"])), float(z["switch_prior_thresh"]))
        assert share == float(z["ref_share"][i]), i
        assert po.hswitch_percentage(mon.avg("prior static"), bool(int(z["soft_trans"])), float(z["switch_prior_thresh"])) == share


@pytest.mark.parametrize("reg", ["MRKLD", "MRENT"])
def test_target_losses_restatement_matches_reference(reg):
    """oracle.target_losses against the REAL cross_entropy_2d / rce / regular_loss and their autograd gradient."""
    z = np.load(os.path.join(GOLDEN, "target_losses.npz"))
    out = torch.from_numpy(z["out"]).requires_grad_(True)
    got = po.target_losses(out, torch.from_numpy(z["labels"]), float(z["alpha"]), float(z["beta"]), float(z["reg_weight"]), reg)
    got["total"].backward()
    for k in ("ce", "rce", "reg", "total"):
        assert got[k].item() == pytest.approx(float(z[f"ref_{reg}_{k}"]), rel=1e-6, abs=1e-7), k
    np.testing.assert_allclose(out.grad.numpy(), z[f"ref_{reg}_grad"], rtol=1e-5, atol=1e-8)
