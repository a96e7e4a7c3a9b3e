"""GPU parity tests: the CUDA path, called through the C ABI via the drop-in handler, against
(1) the golden fixtures produced by the real reference and (2) the CPU oracle on seeded inputs.

Tolerances, all written out here:
  distances        |d' - ref| <= 1e-5 * (raw distance + row minimum)   (raw-scale relative)
  soft predictions |r - ref|  <= 1e-5 absolute
  class sums / prototypes / squared means   <= 1e-5 * max|ref|
  labels           bit-exact except rows whose reference top-2 margin < 1e-6 or |max - thresh| < 1e-6
  statistics       <= 1e-6 absolute
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import proto_oracle as po

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
OPS = sorted(glob.glob(os.path.join(GOLDEN, "ops_*.npz")))
IMPLS = ["simt", "tcgen05"]


def need_shape(impl, D, C=19):
    """The tcgen05 kernel covers what onda_impl_supported reports (D % 64 == 0, C <= 32); everything else is the
    CUDA-core kernel."""
    from onda_b200 import _native as nat
    if impl == "tcgen05" and not nat.load().onda_impl_supported(1, D, 128, C, nat.IMPL["tcgen05"]):
        pytest.skip(f"tcgen05 kernel does not cover D={D}, C={C}")


def dev():
    return torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.asarray(a))


def make_handler(protos, sq_mean, counter, metric, tau=1.0, thresh=0.3, ma_lambda=0.9995, impl="auto", **kw):
    from onda_b200 import prototype_handler
    h = prototype_handler(ma_lambda=ma_lambda, tau=tau, thresh=thresh, distance_metric=metric, impl=impl, **kw)
    h.prototypes = protos.clone().to(dev())
    h.squared_mean = sq_mean.clone().to(dev())
    h.counter = counter.clone().to(dev())
    return h


def make_oracle(protos, sq_mean, counter, metric, tau=1.0, thresh=0.3, ma_lambda=0.9995):
    o = po.OracleHandler(ma_lambda=ma_lambda, tau=tau, thresh=thresh, distance_metric=metric)
    o.prototypes, o.squared_mean, o.counter = protos.clone(), sq_mean.clone(), counter.clone()
    return o


def check_labels(labels, ref_labels, ref_soft, thresh):
    labels = labels.cpu().flatten()
    ref_labels = ref_labels.flatten()
    top2 = ref_soft.topk(2, dim=1)[0]
    exempt = ((top2[:, 0] - top2[:, 1]) < 1e-6) | ((top2[:, 0] - thresh).abs() < 1e-6) | torch.isnan(top2[:, 0])
    bad = (labels != ref_labels) & ~exempt
    assert int(bad.sum()) == 0, f"{int(bad.sum())} labels differ off near-ties (exempt rows: {int(exempt.sum())})"
    return int(((labels != ref_labels) & exempt).sum())


def check_dist(dist, ref_shifted, raw_truth):
    scale = raw_truth + raw_truth.min(dim=1, keepdim=True)[0]
    err = (dist.cpu().double() - ref_shifted.double()).abs()
    assert bool((err <= 1e-5 * scale + 1e-7).all()), f"distance error {float((err / (scale + 1e-12)).max()):.3e} (raw-scale relative)"


def raw_truth64(feat, protos, sq_mean, counter, metric):
    f, P = feat.double(), protos.double()
    sigma = po.pooled_std(P, sq_mean.double(), counter.double()) if metric == "mahalanobis" else None
    return po.raw_distance(f, P, sigma)


def close_rel_max(got, ref, tol=1e-5):
    got, ref = got.detach().cpu().double(), torch.as_tensor(ref).double()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= tol * scale + 1e-30, f"max error {err:.3e} vs {tol:.0e} * {scale:.3e}"


# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("path", OPS, ids=[os.path.basename(p)[:-4] for p in OPS])
def test_golden_ops(path, impl):
    from onda_b200 import Monitor
    z = np.load(path)
    need_shape(impl, z["protos"].shape[1])
    metric, tau, thresh, lam = str(z["metric"]), float(z["tau"]), float(z["thresh"]), float(z["ma_lambda"])
    protos, sq_mean, counter = T(z["protos"]), T(z["sq_mean"]), T(z["counter"])
    h = make_handler(protos, sq_mean, counter, metric, tau, thresh, lam, impl)
    feat, prior, out = (T(z[k]).to(dev()) for k in ("feat", "prior", "out"))
    raw = raw_truth64(T(z["feat"]), protos, sq_mean, counter, metric)

    dist = h.distance_measure(feat)
    assert dist.shape == z["ref_dist"].shape and dist.dtype == torch.float32
    check_dist(dist, T(z["ref_dist"]), raw)

    mon = Monitor(200, 0.003, "hamming")
    labels = h.pseudo_labels(feat, prior, confidence_monitor=mon)
    soft = h.pseudo_labels(feat, prior, soft=True)
    assert labels.shape == z["ref_labels"].shape and labels.dtype == torch.int64
    assert soft.shape == z["ref_soft"].shape and soft.dtype == torch.float32
    assert float((soft.cpu() - T(z["ref_soft"])).abs().max()) <= 1e-5
    check_labels(labels, T(z["ref_labels"]), T(z["ref_soft"]), np.float32(thresh))
    assert mon.current_dict["prototypes"][0] == pytest.approx(float(z["ref_stat_proto"]), abs=1e-6)
    assert h.last_stats["prior"] == pytest.approx(float(z["ref_stat_prior"]), abs=1e-6)
    assert h.last_stats["pseudolabel confidence"] == pytest.approx(float(z["ref_stat_pl"]), abs=1e-6)
    assert h.last_stats["pseudolabel_pixel_num"] == float((T(z["ref_labels"]) != 255).sum())

    s1, cnt = h.get_proto_array(feat, out)
    s2, _ = h.get_proto_array(feat ** 2, out)
    close_rel_max(s1, z["ref_sum"])
    close_rel_max(s2, z["ref_sumsq"])
    assert torch.equal(cnt.cpu(), T(z["ref_count"]))
    np.testing.assert_allclose(h.global_var().cpu().numpy(), z["ref_global_std"], rtol=2e-6, equal_nan=True)
    np.testing.assert_allclose(h.prototype_var().cpu().numpy(), z["ref_class_std"], rtol=2e-6, atol=1e-7, equal_nan=True)

    h.ma(feat, out)
    close_rel_max(h.prototypes, z["ref_ma_protos"])
    close_rel_max(h.squared_mean, z["ref_ma_sq_mean"])
    assert torch.equal(h.counter.cpu(), counter)  # ma leaves the counter alone


@pytest.mark.parametrize("impl", IMPLS)
def test_fused_equals_separate_calls(impl):
    case = po.synth_case(21, 2, 128, 11, 19)
    hs = [make_handler(case["protos"], case["sq_mean"], case["counter"], "mahalanobis", impl=impl, fuse_hard_soft=f)
          for f in (True, False)]
    feat, prior, out = (case[k].to(dev()) for k in ("feat", "prior", "out"))
    labels_f, soft_f = hs[0].pseudo_labels_fused(feat, prior, out)
    hs[0].ma(feat, out)
    labels_s = hs[1].pseudo_labels(feat, prior)
    soft_s = hs[1].pseudo_labels(feat, prior, soft=True)
    hs[1].ma(feat, out)
    assert torch.equal(labels_f, labels_s) and torch.equal(soft_f, soft_s)
    if impl == "simt":
        assert torch.equal(hs[0].prototypes, hs[1].prototypes) and torch.equal(hs[0].squared_mean, hs[1].squared_mean)
    else:   # the separate ma() call has no distance output and runs the CUDA-core class-sum kernel: same sums, other order
        close_rel_max(hs[0].prototypes, hs[1].prototypes.cpu(), 1e-6)
        close_rel_max(hs[0].squared_mean, hs[1].squared_mean.cpu(), 1e-6)


def test_hard_then_soft_reuses_launch_only_for_same_tensors():
    from onda_b200 import _native as nat
    case = po.synth_case(22, 1, 32, 9, 9)
    h = make_handler(case["protos"], case["sq_mean"], case["counter"], "euclidean")
    feat, prior = case["feat"].to(dev()), case["prior"].to(dev())
    lib = nat.load()
    h.pseudo_labels(feat, prior)
    n0 = lib.onda_launch_count()
    soft = h.pseudo_labels(feat, prior, soft=True)
    assert lib.onda_launch_count() == n0          # served by the hard call's launch
    prior2 = prior.clone()
    h.pseudo_labels(feat, prior)
    soft2 = h.pseudo_labels(feat, prior2, soft=True)  # different tensor object -> recomputed
    assert lib.onda_launch_count() > n0
    assert torch.equal(soft, soft2)
    h.pseudo_labels(feat, prior)
    prior.mul_(0.5)                                  # in-place change bumps the version -> recomputed
    n1 = lib.onda_launch_count()
    h.pseudo_labels(feat, prior, soft=True)
    assert lib.onda_launch_count() > n1


def test_append_sequence_golden():
    from onda_b200 import prototype_handler
    z = np.load(os.path.join(GOLDEN, "append_seq.npz"))
    h = prototype_handler(distance_metric="mahalanobis")
    for i in range(3):
        h.append(T(z[f"feat{i}"]).to(dev()), T(z[f"out{i}"]).to(dev()))
    h.append(T(z["rows3"]).to(dev()), T(z["hot3"]).to(dev()))   # (M, D) rows + int64 one-hot
    close_rel_max(h.prototypes, z["ref_protos"])
    close_rel_max(h.squared_mean, z["ref_sq_mean"])
    assert torch.equal(h.counter.cpu(), T(z["ref_counter"]))


SHAPES = [
    # (seed, B, D, h, w, C, metric)
    (31, 1, 256, 65, 129, 19, "mahalanobis"),     # BASELINE config 1 at the real feature width
    (32, 1, 2048, 65, 129, 19, "mahalanobis"),    # BASELINE config 1 (D = 2048)
    (33, 1, 2048, 65, 129, 19, "euclidean"),
    (34, 3, 256, 33, 57, 19, "euclidean"),
    (35, 2, 50, 9, 17, 19, "mahalanobis"),        # D not a multiple of 8
    (36, 2, 96, 13, 11, 7, "mahalanobis"),        # fewer classes
    (37, 1, 64, 17, 9, 25, "mahalanobis"),        # more than 20 classes (32-wide path)
    (38, 5, 24, 1, 1, 19, "euclidean"),           # 5 pixels in total
    (39, 2, 320, 3, 51, 19, "mahalanobis"),       # HW = 153
    (40, 2, 128, 21, 13, 19, "mahalanobis"),      # tcgen05 range: D = 128, ragged last tile
    (42, 3, 192, 5, 31, 7, "euclidean"),          # D = 192 (6 chunks), few classes, H*W = 3 mod 4
    (43, 1, 128, 9, 14, 25, "mahalanobis"),       # 25 classes (32-wide epilogue), 126 pixels: a single partial tile
    (44, 7, 160, 37, 53, 19, "mahalanobis"),      # D = 160: not a multiple of 64 (CUDA-core kernel), 13727 pixels
    (45, 2, 64, 16, 24, 19, "mahalanobis"),       # D = 64 (2 chunks), even H*W (all rows share one shift)
    (46, 3, 256, 10, 13, 19, "mahalanobis"),      # H*W = 130 = 2 mod 4, a 2-pixel last tile
]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("seed,B,D,h,w,C,metric", SHAPES)
def test_oracle_parity_shapes(seed, B, D, h, w, C, metric, impl):
    need_shape(impl, D, C)
    case = po.synth_case(seed, B, D, h, w, c=C)
    hd = make_handler(case["protos"], case["sq_mean"], case["counter"], metric, impl=impl)
    orc = make_oracle(case["protos"], case["sq_mean"], case["counter"], metric)
    feat, prior, out = (case[k].to(dev()) for k in ("feat", "prior", "out"))
    ref_shift = orc.distance_measure(case["feat"])
    raw = raw_truth64(case["feat"], case["protos"], case["sq_mean"], case["counter"], metric)
    check_dist(hd.distance_measure(feat), ref_shift, raw)
    ref_labels, ref_soft = po.fused_step(orc, case["feat"], case["prior"], case["out"])
    labels, soft = hd.pseudo_labels_fused(feat, prior, out)
    hd.ma(feat, out)
    assert float((soft.cpu() - ref_soft).abs().max()) <= 1e-5
    check_labels(labels, ref_labels, ref_soft, np.float32(0.3))
    close_rel_max(hd.prototypes, orc.prototypes)
    close_rel_max(hd.squared_mean, orc.squared_mean)
    q, _ = po.rectify(ref_shift, po.to_rows(case["prior"]), 1.0)
    assert hd.last_stats["prototypes"] == pytest.approx(q.max(dim=1)[0].mean().item(), abs=1e-6)
    assert hd.last_stats["pixels"] == B * h * w


@pytest.mark.parametrize("impl", IMPLS)
def test_two_runs_are_bit_identical(impl):
    case = po.synth_case(41, 4, 256, 33, 65)
    feat, prior, out = (case[k].to(dev()) for k in ("feat", "prior", "out"))
    results = []
    for _ in range(3):
        h = make_handler(case["protos"], case["sq_mean"], case["counter"], "mahalanobis", impl=impl)
        labels, soft = h.pseudo_labels_fused(feat, prior, out)
        s1, cnt = h.get_proto_array(feat, out)
        h.ma(feat, out)
        results.append((labels, soft, s1, cnt, h.prototypes.clone(), h.squared_mean.clone()))
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert torch.equal(a, b)


@pytest.mark.parametrize("impl", IMPLS)
def test_sequence_of_steps_golden(impl):
    """260 pseudo-label -> ma steps; compares the final prototypes and the per-step traces."""
    from onda_b200 import Monitor
    z = np.load(os.path.join(GOLDEN, "sequence_ma.npz"))
    steps, d = int(z["steps"]), int(z["d"])
    need_shape(impl, d)
    first = po.synth_case(4999, 1, d, 9, 11)
    h = make_handler(first["protos"], first["sq_mean"], first["counter"], "mahalanobis", ma_lambda=0.95, impl=impl)
    mon = Monitor(50, 0.003, "hamming")
    mismatched = 0
    for i in range(steps):
        case = po.synth_case(5000 + i, 1, d, 9, 11, protos=first["protos"] + 0.002 * i, counter=first["counter"])
        feat, prior, out = (case[k].to(dev()) for k in ("feat", "prior", "out"))
        labels, soft = h.pseudo_labels_fused(feat, prior, out, confidence_monitor=mon)
        h.ma(feat, out)
        assert mon.current_dict["prototypes"][-1] == pytest.approx(float(z["ref_stat"][i]), abs=1e-6)
        lab = labels.cpu().flatten()
        mismatched += int(int((lab * torch.arange(1, lab.numel() + 1)).sum()) != int(z["ref_label_hash"][i]))
    assert mismatched <= 2, f"{mismatched} of {steps} steps had a label differing from the reference"
    close_rel_max(h.prototypes, z["ref_protos"])
    close_rel_max(h.squared_mean, z["ref_sq_mean"])
    assert mon.dev_avg("prototypes") == pytest.approx(float(z["ref_dev_proto"]), abs=1e-6)


def test_prior_mix_golden():
    z = np.load(os.path.join(GOLDEN, "stats_logits.npz"))
    case = po.synth_case(61, 2, 8, 11, 17)
    h = make_handler(case["protos"], case["sq_mean"], case["counter"], "euclidean")
    la, lb, lc = (T(z[k]).to(dev()) for k in ("la", "lb", "lc"))
    _, conf, _ = h.prior_mix([la, lb, lc], [0, 0, 0], write_prior=False)
    np.testing.assert_allclose(conf, z["ref_conf"], atol=1e-6)
    mix, conf, mix_conf = h.prior_mix([la, lb], [0.25, 1.0])
    assert mix.shape == la.shape
    assert float((mix.cpu() - T(z["ref_mix"])).abs().max()) <= 1e-6
    assert mix_conf == pytest.approx(float(z["ref_mix_conf"]), abs=1e-6)
    pct = float(z["pct"])
    hmix, conf, hconf = h.prior_mix([la, lb, lc], [pct * 0.0, pct * 1.0, (1 - pct) * 1.0])
    assert float((hmix.cpu() - T(z["ref_hmix"])).abs().max()) <= 1e-6
    assert hconf == pytest.approx(float(z["ref_hmix_conf"]), abs=1e-6)


class _Spec(dict):
    def __getattr__(self, k):
        return self[k] if k in self else {}


class _FakeSegModel(torch.nn.Module):
    def __init__(self, z, name):
        super().__init__()
        self.stem = torch.nn.Conv2d(3, 24, 3, stride=8, padding=1)
        self.head = torch.nn.Conv2d(24, 19, 1)
        with torch.no_grad():
            self.stem.weight.copy_(T(z[f"w_{name}_stem_weight"])); self.stem.bias.copy_(T(z[f"w_{name}_stem_bias"]))
            self.head.weight.copy_(T(z[f"w_{name}_head_weight"])); self.head.bias.copy_(T(z[f"w_{name}_head_bias"]))

    def forward(self, x):
        feat = self.stem(x)
        return None, {"feat": feat, "out": self.head(feat)}


def test_hybrid_method_golden():
    """hybrid_proDA.prototype_predictions + ma over 56 steps against the real reference run:
    identical selector trace, labels, soft predictions and final prototypes."""
    from onda_b200 import Monitor, HybridSelect, methods
    z = np.load(os.path.join(GOLDEN, "method_hybrid.npz"))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    class Method:
        pass
    m = Method()
    m.device = dev()
    m.cfg_spec = _Spec(EMA_LAMBDA=0, STATIC_LAMBDA=1, DYNAMIC_LAMBDA=1)
    m.intensity_ma = Monitor(int(z["limit"]), float(z["exp_const"]), "hamming")
    m.model_select = HybridSelect(HybridSelect.static, tuple(z["gray"]), float(z["dev_thresh"]))
    m.ema_model, m.static_model, m.dynamic_model = (_FakeSegModel(z, n).to(dev()).eval() for n in ("ema", "static", "dynamic"))
    m.prototypes = make_handler(T(z["init_protos"]), T(z["init_sq_mean"]), T(z["init_counter"]), "mahalanobis",
                                tau=float(z["tau"]), thresh=float(z["thresh"]), ma_lambda=float(z["ma_lambda"]))
    select, exempt_total = [], 0
    for i in range(z["images"].shape[0]):
        pred = methods.hybrid_prototype_predictions(m, {"image": T(z["images"][i])})
        m.prototypes.ma(pred["ema_model"]["feat"], pred["ema_model"]["out"])
        select.append(m.model_select.current)
        ref_soft = T(z["ref_soft"][i])
        assert float((pred["soft_predictions"].cpu() - ref_soft).abs().max()) <= 2e-5  # conv backbone on GPU adds ~1e-6
        exempt_total += check_labels_loose(pred["pseudolabels"], T(z["ref_labels"][i]), ref_soft, np.float32(z["thresh"]))
        assert m.intensity_ma.current_dict["prior static"][-1] == pytest.approx(float(z["ref_stat_prior_static"][i]), abs=2e-6)
    assert select == list(z["ref_select"])
    close_rel_max(m.prototypes.prototypes, z["ref_final_protos"])
    close_rel_max(m.prototypes.squared_mean, z["ref_final_sq_mean"])


def _variant_images(z):
    rs = np.random.RandomState(int(z["image_seed"]))
    imgs = [torch.from_numpy(rs.standard_normal((1, 3, 32, 56)).astype(np.float32) * np.float32(sc)) for sc in z["image_scales"]]
    assert float(torch.stack(imgs).double().sum()) == pytest.approx(float(z["image_checksum"]), abs=1e-6)
    return imgs


VARIANTS = [
    # fixture, methods.<function>, selector
    ("method_hswitch_soft.npz", "hswitch_prototype_predictions", None),
    ("method_hswitch_hard.npz", "hswitch_prototype_predictions", None),
    ("method_vswitch.npz", "vswitch_prototype_predictions", "dev"),
    ("method_base_mix.npz", "base_prototype_predictions", None),
    ("method_base_rule.npz", "base_prototype_predictions", None),
]


@pytest.mark.parametrize("fixture,fn,selector", VARIANTS, ids=[v[0][7:-4] for v in VARIANTS])
def test_method_variants_golden(fixture, fn, selector):
    """hswitch_proDA / vswitch_proDA / online_proDA ``prototype_predictions`` + ``ma`` over 44 steps against runs of the
    real reference classes (prototypes_hswitch.py:26-85, prototypes_vswitch.py:36-87, prototypes.py:208-273): same
    static share / selector trace, labels, soft predictions, Monitor entries and final prototypes."""
    from onda_b200 import Monitor, DevSelect, methods
    z = np.load(os.path.join(GOLDEN, fixture))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    class Method:
        pass
    m = Method()
    m.device = dev()
    m.cfg_spec = _Spec(EMA_LAMBDA=float(z["ema_lambda"]), STATIC_LAMBDA=float(z["static_lambda"]),
                       DYNAMIC_LAMBDA=float(z["dynamic_lambda"]), SOFT_TRANS=bool(int(z["soft_trans"])),
                       SWITCH_PRIOR_THRESH=float(z["switch_prior_thresh"]))
    m.intensity_ma = Monitor(int(z["limit"]), float(z["exp_const"]), "hamming")
    if selector == "dev":
        m.model_select = DevSelect(DevSelect.static, float(z["switch_prior_thresh"]))
    m.ema_model, m.static_model, m.dynamic_model = (_FakeSegModel(z, n).to(dev()).eval() for n in ("ema", "static", "dynamic"))
    m.prototypes = make_handler(T(z["init_protos"]), T(z["init_sq_mean"]), T(z["init_counter"]), "mahalanobis",
                                tau=float(z["tau"]), thresh=float(z["thresh"]), ma_lambda=float(z["ma_lambda"]))
    predict = getattr(methods, fn)
    select, share = [], []
    for i, img in enumerate(_variant_images(z)):
        pred = predict(m, {"image": img, "label": 0})
        m.prototypes.ma(pred["ema_model"]["feat"], pred["ema_model"]["out"])
        cur = m.intensity_ma.current_dict
        select.append(m.model_select.current if selector else -1)
        share.append(float(cur["percentage_static"][-1]) if "percentage_static" in cur else -1.0)
        ref_soft = T(z["ref_soft"][i])
        assert float((pred["soft_predictions"].cpu() - ref_soft).abs().max()) <= 2e-5  # conv backbone on GPU adds ~1e-6
        check_labels_loose(pred["pseudolabels"], T(z["ref_labels"][i]), ref_soft, np.float32(z["thresh"]))
        for key in ("prior static", "prior", "prototypes", "pseudolabel confidence"):
            assert cur[key][-1] == pytest.approx(float(z["ref_stat_" + key.replace(" ", "_")][i]), abs=3e-6), (i, key)
    assert select == list(z["ref_select"])
    # the h-switch ramp amplifies the confidence noise of the GPU convolution by 25/3
    assert np.allclose(share, z["ref_share"], atol=5e-5)
    close_rel_max(m.prototypes.prototypes, z["ref_final_protos"])
    close_rel_max(m.prototypes.squared_mean, z["ref_final_sq_mean"])


def check_labels_loose(labels, ref_labels, ref_soft, thresh, margin=1e-4):
    """Label check for inputs that went through a cuDNN convolution (feature noise ~1e-6)."""
    labels, ref_labels = labels.cpu().flatten(), ref_labels.flatten()
    top2 = ref_soft.topk(2, dim=1)[0]
    exempt = ((top2[:, 0] - top2[:, 1]) < margin) | ((top2[:, 0] - thresh).abs() < margin)
    bad = (labels != ref_labels) & ~exempt
    assert int(bad.sum()) == 0
    return int(exempt.sum())


# --------------------------------------------------------------------------------------
# full BASELINE sizes against the oracle (1.4 s of CPU per 270k pixels at D = 256, ~10 s at D = 2048)
# --------------------------------------------------------------------------------------
FULL = [
    (61, 32, 256, 65, 129),      # the bench workload: BASELINE configs[2] on one GPU
    (62, 8, 256, 129, 257),      # BASELINE configs[3]: full resolution
    (63, 4, 2048, 65, 129),      # BASELINE configs[0] width, four images
]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("seed,B,D,h,w", FULL)
def test_full_size_oracle_parity(seed, B, D, h, w, impl):
    """Labels, soft predictions, statistics (entropy included) and the post-ma state at the sizes the bench times."""
    need_shape(impl, D)
    case = po.synth_case(seed, B, D, h, w)
    hd = make_handler(case["protos"], case["sq_mean"], case["counter"], "mahalanobis", impl=impl)
    orc = make_oracle(case["protos"], case["sq_mean"], case["counter"], "mahalanobis")
    feat, prior, out = (case[k].to(dev()) for k in ("feat", "prior", "out"))
    ref_labels, ref_soft = po.fused_step(orc, case["feat"], case["prior"], case["out"])
    labels, soft = hd.pseudo_labels_fused(feat, prior, out)
    hd.ma(feat, out)
    assert float((soft.cpu() - ref_soft).abs().max()) <= 1e-5
    check_labels(labels, ref_labels, ref_soft, np.float32(0.3))
    close_rel_max(hd.prototypes, orc.prototypes)
    close_rel_max(hd.squared_mean, orc.squared_mean)
    st = hd.last_stats
    assert st["pixels"] == B * h * w
    assert st["pseudolabel confidence"] == pytest.approx(ref_soft.max(dim=1)[0].mean().item(), abs=1e-6)
    assert st["pseudolabel_pixel_num"] == pytest.approx(float((ref_labels != 255).sum()), abs=2)
    # prob_2_entropy (framework/utils/func.py:71-74) of the rectified posterior, summed over classes, batch mean
    ref_entropy = po.normalised_entropy(ref_soft).sum(dim=1).double().mean().item()
    assert st["entropy"] == pytest.approx(ref_entropy, abs=2e-6)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("B,D", [(32, 256), (4, 2048)])
def test_step_is_bit_reproducible(B, D, impl):
    """Many tiles per SM, several launches, tile_schedule="fixed": labels, soft predictions and the post-ma state are
    the same bits every time (one writer per accumulator, fixed-order combine: north_star's deterministic class sums).
    The default dynamic schedule gives up exactly this (the last bits of the class sums) for 4-5 % of the kernel time."""
    need_shape(impl, D)
    case = po.synth_case(77, B, D, 65, 129)
    feat, prior, out = (case[k].to(dev()) for k in ("feat", "prior", "out"))
    runs = []
    for _ in range(4):
        hd = make_handler(case["protos"], case["sq_mean"], case["counter"], "mahalanobis", impl=impl)
        hd.tile_schedule = "fixed"
        labels, soft = hd.pseudo_labels_fused(feat, prior, out)
        hd.ma(feat, out)
        runs.append((labels.clone(), soft.clone(), hd.prototypes.clone(), hd.squared_mean.clone()))
    for r in runs[1:]:
        for a, b in zip(runs[0], r):
            assert torch.equal(a, b)


_CHAIN_SCRIPT = r"""
import hashlib, sys, torch
sys.path.insert(0, {root!r})
from oracle import proto_oracle as po
from onda_b200 import prototype_handler
dev = torch.device("cuda:0")
case = po.synth_case(91, 6, 256, 65, 129)
h = prototype_handler(ma_lambda=0.9995, tau=1.0, thresh=0.3, distance_metric="mahalanobis", impl="tcgen05", tile_schedule="fixed")
h.prototypes, h.squared_mean, h.counter = (case[k].clone().to(dev) for k in ("protos", "sq_mean", "counter"))
feat, prior, out = (case[k].to(dev) for k in ("feat", "prior", "out"))
m = hashlib.sha256()
for step in range(6):                    # back-to-back steps: every kernel of the chain follows its predecessor directly
    labels, soft = h.pseudo_labels_fused(feat, prior, out)
    h.ma(feat, out)
g = torch.cuda.CUDAGraph()               # and the same chain captured and replayed
with torch.cuda.graph(g):
    labels, soft = h.pseudo_labels_fused(feat, prior, out)
    h.ma(feat, out)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
for t in (labels, soft, h.prototypes, h.squared_mean):
    m.update(t.cpu().numpy().tobytes())
print("HASH", m.hexdigest())
"""


def test_programmatic_launch_chain_changes_no_bit():
    """The step's kernels are launched with the programmatic-stream-serialization attribute (a kernel may start while its
    predecessor runs and waits in griddepcontrol.wait before touching global memory).  Eleven chained steps -- six eager,
    five replays of a captured graph -- must give the same bits as plain launches (ONDA_PDL=0), and the same bits twice."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = _CHAIN_SCRIPT.format(root=root)
    hashes = []
    for pdl in ("1", "0", "1"):
        env = dict(os.environ, ONDA_PDL=pdl)
        out = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        hashes.append([ln for ln in out.stdout.splitlines() if ln.startswith("HASH")][-1])
    assert hashes[0] == hashes[1] == hashes[2]


# --------------------------------------------------------------------------------------
# full BASELINE sizes: size-independent properties
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("B,D,h,w", [(32, 256, 65, 129), (8, 256, 129, 257), (8, 2048, 65, 129)])
def test_full_size_properties(B, D, h, w, impl):
    need_shape(impl, D)
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + D)
    d = dev()
    C = 19
    protos = torch.randn(C, D, generator=g, device=d) * 2.5
    sq_mean = protos ** 2 + torch.rand(C, D, generator=g, device=d) * 1.5 + 0.5
    counter = torch.floor(torch.rand(C, generator=g, device=d) * 6.9e4 + 1e3)
    lab = torch.randint(0, C, (B, (h + 7) // 8, (w + 7) // 8), generator=g, device=d)
    lab = lab.repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :h, :w]
    feat = torch.randn(B, D, h, w, generator=g, device=d) * 2.5 + 0.5 * protos[lab].permute(0, 3, 1, 2)
    hot = torch.nn.functional.one_hot(lab, C).permute(0, 3, 1, 2).float()
    out = torch.randn(B, C, h, w, generator=g, device=d) * 3 + 4 * hot
    prior = (torch.randn(B, C, h, w, generator=g, device=d) * 3 + 4 * hot).softmax(1)
    hd = make_handler(protos.cpu(), sq_mean.cpu(), counter.cpu(), "mahalanobis", impl=impl)
    N = B * h * w
    labels, soft = hd.pseudo_labels_fused(feat, prior, out)
    s1, cnt = hd.get_proto_array(feat, out)
    s2, _ = hd.get_proto_array(feat ** 2, out)
    # rows of the rectified posterior sum to one; labels are its argmax or 255 exactly below the threshold
    assert float((soft.sum(1) - 1).abs().max()) <= 1e-5
    m, arg = soft.max(1)
    lab_out = labels.flatten()
    keep = lab_out != 255
    assert bool((lab_out[keep] == arg[keep]).all()) and bool((m[keep] >= np.float32(0.3)).all())
    assert bool((m[~keep] < np.float32(0.3)).all())
    assert hd.last_stats["pixels"] == N
    # class sums: counts add up to N and match a bincount; the sums are linear (sum over classes = sum over pixels)
    y = out.argmax(1).flatten()
    assert torch.equal(cnt, torch.bincount(y, minlength=C).float()) and float(cnt.sum()) == N
    tot = feat.double().sum(dim=(0, 2, 3))
    tot2 = (feat.double() ** 2).sum(dim=(0, 2, 3))
    assert float((s1.double().sum(0) - tot).abs().max()) <= 1e-5 * float(feat.abs().sum(dim=(0, 2, 3)).max())
    assert float((s2.double().sum(0) - tot2).abs().max()) <= 1e-5 * float(tot2.max())
    # sharding invariance: two half-batches give the same labels/soft and the same sums up to fp32 order
    half = B // 2
    la, sa = hd.pseudo_labels_fused(feat[:half], prior[:half], out[:half])
    a1, ac = hd.get_proto_array(feat[:half], out[:half])
    lb, sb = hd.pseudo_labels_fused(feat[half:], prior[half:], out[half:])
    b1, bc = hd.get_proto_array(feat[half:], out[half:])
    assert torch.equal(torch.cat([la, lb]), labels) and torch.equal(torch.cat([sa, sb]), soft)
    assert torch.equal(ac + bc, cnt)
    assert float((a1 + b1 - s1).abs().max()) <= 1e-5 * float(s1.abs().max())
    # distances: the row minimum is exactly zero and the public value is non-negative
    dist = hd.distance_measure(feat[:1])
    assert float(dist.min(1)[0].abs().max()) == 0.0 and float(dist.min()) >= 0.0


def test_error_behaviour_and_persistence(tmp_path):
    from onda_b200 import prototype_handler
    with pytest.raises(ValueError):
        prototype_handler(distance_metric="cosine")            # prototype_handler.py:29
    case = po.synth_case(51, 1, 16, 5, 7)
    h = make_handler(case["protos"], case["sq_mean"], case["counter"], "mahalanobis")
    feat = case["feat"].to(dev())
    with pytest.raises(AttributeError):
        h.pseudo_labels(feat, prior=None)                       # the reference dereferences prior.device (:142)
    with pytest.raises(RuntimeError):
        h.pseudo_labels(case["feat"], case["prior"])            # CPU tensors: no fallback
    assert prototype_handler(confidence_regularization_threshold={}).confidence_regularization_threshold == 1
    loc = str(tmp_path / "proto.pickle")
    h.save(loc)
    h2 = prototype_handler(distance_metric="mahalanobis", thresh=0.3)
    assert h2.load(loc) is True and h2.load(str(tmp_path / "missing.pickle")) is False
    h.thresh = 0.3
    a = h.pseudo_labels(feat, case["prior"].to(dev()))
    b = h2.pseudo_labels(feat, case["prior"].to(dev()))
    assert torch.equal(a, b)
    import pickle
    with open(loc, "rb") as f:
        tup = pickle.load(f)
    assert len(tup) == 3 and tup[0].shape == (19, 16)           # the reference's 3-tuple format


def test_tau_regularisation_side_effect():
    """confidence_monitor median above the threshold bumps tau by 0.001 (prototype_handler.py:151-156)."""
    from onda_b200 import Monitor
    case = po.synth_case(52, 1, 16, 5, 7)
    h = make_handler(case["protos"], case["sq_mean"], case["counter"], "euclidean")
    h.confidence_regularization_threshold = 0.0
    mon = Monitor(10)
    feat, prior = case["feat"].to(dev()), case["prior"].to(dev())
    h.pseudo_labels(feat, prior, confidence_monitor=mon)
    assert h.tau == pytest.approx(1.001) and mon.current_dict["tau"] == [h.tau]
    mon.eval()
    h.pseudo_labels(feat, prior, confidence_monitor=mon)
    assert h.tau == pytest.approx(1.001)                        # frozen monitor: no side effects


def test_fused_call_follows_the_tau_bump_like_two_reference_calls():
    """With the confidence regulariser firing, the reference's hard call raises tau and its soft call that follows
    already uses the new value (prototype_handler.py:151-156): pseudo_labels_fused must return that pair."""
    from onda_b200 import Monitor
    case = po.synth_case(53, 1, 32, 6, 9)
    h = make_handler(case["protos"], case["sq_mean"], case["counter"], "mahalanobis")
    orc = make_oracle(case["protos"], case["sq_mean"], case["counter"], "mahalanobis")
    h.confidence_regularization_threshold = orc.confidence_regularization_threshold = 0.0
    mon, omon = Monitor(10), po.OracleMonitor(10)
    feat, prior = case["feat"].to(dev()), case["prior"].to(dev())
    ref_labels = orc.pseudo_labels(case["feat"], case["prior"], confidence_monitor=omon)
    ref_soft = orc.pseudo_labels(case["feat"], case["prior"], soft=True)
    labels, soft = h.pseudo_labels_fused(feat, prior, None, confidence_monitor=mon)
    assert h.tau == pytest.approx(1.001) and orc.tau == pytest.approx(1.001)
    assert float((soft.cpu() - ref_soft).abs().max()) <= 1e-5
    check_labels(labels, ref_labels, po.rectify(orc.distance_measure(case["feat"]), po.to_rows(case["prior"]), 1.0)[1], np.float32(0.3))


def test_zero_rectified_row_keeps_label_zero_like_the_reference():
    """A pixel whose prior is zero for every class: the reference divides 0 / 0, gets a NaN row, and torch.max returns
    (NaN, index 0), which is not below the threshold -> label 0 (prototype_handler.py:159-166)."""
    case = po.synth_case(54, 1, 32, 4, 5)
    case["prior"][0, :, 1, 2] = 0.0
    h = make_handler(case["protos"], case["sq_mean"], case["counter"], "mahalanobis")
    orc = make_oracle(case["protos"], case["sq_mean"], case["counter"], "mahalanobis")
    ref_labels, ref_soft = po.fused_step(orc, case["feat"], case["prior"], case["out"])
    labels, soft = h.pseudo_labels_fused(case["feat"].to(dev()), case["prior"].to(dev()), case["out"].to(dev()))
    n = 1 * 5 + 2
    assert int(ref_labels[n]) == 0 and bool(torch.isnan(ref_soft[n]).all())
    assert int(labels[n]) == 0 and bool(torch.isnan(soft[n]).all())
    ok = torch.ones(ref_labels.numel(), dtype=torch.bool); ok[n] = False
    assert float((soft.cpu()[ok] - ref_soft[ok]).abs().max()) <= 1e-5


@pytest.mark.parametrize("impl", IMPLS)
def test_degenerate_rows_golden(impl):
    """golden/edge_nan_rows.npz (outputs of the real reference): a zero prior row, a NaN feature vector, and a negative
    tau.  NaN rows in the same places with label 0, labels equal off near-ties, the rest within 1e-5."""
    z = np.load(os.path.join(GOLDEN, "edge_nan_rows.npz"))
    need_shape(impl, z["protos"].shape[1])
    for tag, tau in (("pos", 1.0), ("neg", -0.8)):
        h = make_handler(T(z["protos"]), T(z["sq_mean"]), T(z["counter"]), "mahalanobis", tau=tau, impl=impl)
        feat, prior = T(z["feat"]).to(dev()), T(z["prior"]).to(dev())
        labels = h.pseudo_labels(feat, prior).cpu()
        soft = h.pseudo_labels(feat, prior, soft=True).cpu()
        ref_labels, ref_soft = T(z[f"ref_labels_{tag}"]), T(z[f"ref_soft_{tag}"])
        nan_rows = torch.isnan(ref_soft).all(dim=1)
        assert torch.equal(torch.isnan(soft).all(dim=1), nan_rows) and int(nan_rows.sum()) == 2
        assert bool((labels.flatten()[nan_rows] == 0).all())
        ok = ~nan_rows
        assert float((soft[ok] - ref_soft[ok]).abs().max()) <= 1e-5
        check_labels(labels[ok], ref_labels[ok], ref_soft[ok], 0.3)


LABEL_MAPS = ["one_class", "two_classes_odd_split", "aligned_runs_of_32", "stripes_of_5", "random", "ragged_tail"]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("kind", LABEL_MAPS)
def test_class_sums_for_label_maps(kind, impl):
    """Class sums / counts for label maps that stress the summation scheme of the tensor-core kernel: classes that span
    one, two or all four 32-entry ranges of a tile's class-sorted order (head partials), runs that end exactly on a
    range boundary, and a last tile with fewer than 128 pixels.  Tolerance 1e-5 * max|ref|; counts exact."""
    D, C = 256, 19
    need_shape(impl, D, C)
    B, h, w = (1, 8, 16 * 3) if kind != "ragged_tail" else (1, 7, 53)       # 384 px = 3 full tiles; 371 px = ragged
    n = B * h * w
    g = torch.Generator().manual_seed(77)
    if kind == "one_class":
        y = torch.full((n,), 3)
    elif kind == "two_classes_odd_split":
        y = torch.where(torch.arange(n) % 128 < 71, 5, 11)
    elif kind == "aligned_runs_of_32":
        y = (torch.arange(n) // 32) % C
    elif kind == "stripes_of_5":
        y = (torch.arange(n) // 5) % C
    else:
        y = torch.randint(0, C, (n,), generator=g)
    y = y[torch.randperm(n, generator=g)] if kind in ("two_classes_odd_split", "stripes_of_5") else y
    out = torch.randn(n, C, generator=g) * 0.1
    out[torch.arange(n), y] += 5.0
    out = out.reshape(B, h, w, C).permute(0, 3, 1, 2).contiguous()
    feat = torch.randn(B, D, h, w, generator=g) * 2.0 + 0.3
    case = po.synth_case(5, 1, D, 4, 4, c=C)
    hd = make_handler(case["protos"], case["sq_mean"], case["counter"], "mahalanobis", impl=impl)
    sums, cnt = hd.get_proto_array(feat.to(dev()), out.to(dev()))
    ref_s, ref_c = po.class_sums(feat, out)
    assert torch.equal(cnt.cpu(), ref_c)
    assert (sums.cpu() - ref_s).abs().max() <= 1e-5 * ref_s.abs().max()
    # the squared sums go through ma(): compare the blended state
    o = make_oracle(case["protos"], case["sq_mean"], case["counter"], "mahalanobis")
    prior = torch.full((B, C, h, w), 1.0 / C)
    hd.pseudo_labels_fused(feat.to(dev()), prior.to(dev()), out.to(dev()))
    hd.ma(feat.to(dev()), out.to(dev()))
    o.ma(feat, out)
    assert (hd.prototypes.cpu() - o.prototypes).abs().max() <= 1e-5 * o.prototypes.abs().max()
    assert (hd.squared_mean.cpu() - o.squared_mean).abs().max() <= 1e-5 * o.squared_mean.abs().max()


def test_step_log_stats_match_reference_expressions():
    """prototypes.py:341-352 -- agreement with the student's argmax, non-ignored label count, mean squared prototype.
    Counts are exact (so the fp32 mean of 0/1 values is bit-identical); the prototype mean within 1e-6 relative."""
    case = po.synth_case(21, 3, 64, 9, 14)
    hd = make_handler(case["protos"], case["sq_mean"], case["counter"], "mahalanobis")
    o = make_oracle(case["protos"], case["sq_mean"], case["counter"], "mahalanobis")
    labels = hd.pseudo_labels(case["feat"].to(dev()), case["prior"].to(dev()))
    ref_labels = o.pseudo_labels(case["feat"], case["prior"])
    g = torch.Generator().manual_seed(3)
    student = case["out"] + torch.randn(case["out"].shape, generator=g) * 2.0
    got = hd.step_log_stats(labels, student.to(dev()))
    want = po.step_log_stats(labels.cpu(), student, case["protos"])
    assert got["pseudolabel_pixel_num"] == want["pseudolabel_pixel_num"]
    assert got["output & prototype agreement"] == want["output & prototype agreement"]
    assert abs(got["mean_prototype_intensity_values"] - want["mean_prototype_intensity_values"]) <= 1e-6 * want["mean_prototype_intensity_values"]
    assert 0 < got["pseudolabel_pixel_num"] < labels.numel() and torch.equal(labels.cpu(), ref_labels)
    with pytest.raises(ValueError):
        hd.step_log_stats(labels[:5], student.to(dev()))


def test_weight_ema_is_bit_identical_to_the_reference_expression():
    """update_ema (prototypes.py:407-416) as one launch: parameters of awkward sizes and alignments, int64 / bool / empty
    buffers; bit-exact against `k.clone()*a + q.clone()*(1-a)` evaluated by torch on the CPU; three consecutive updates."""
    from onda_b200 import WeightEma, update_ema
    g = torch.Generator().manual_seed(5)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            big = torch.randn(70001 + 3, generator=g)
            self.w_big = torch.nn.Parameter(big[3:])                        # 4-byte aligned only, several chunks + tail
            self.w_mid = torch.nn.Parameter(torch.randn(8192 * 3, generator=g))   # exactly three chunks
            self.w_small = torch.nn.Parameter(torch.randn(7, generator=g))
            self.w_one = torch.nn.Parameter(torch.randn(1, generator=g))
            self.bn = torch.nn.BatchNorm2d(5)                               # float buffers + int64 num_batches_tracked
            self.register_buffer("flags", torch.tensor([True, False, True]))
            self.register_buffer("odd_bytes", torch.arange(37, dtype=torch.uint8))
            self.register_buffer("empty", torch.zeros(0))

    q, k = Net(), Net()
    with torch.no_grad():
        for p in list(q.parameters()) + list(k.parameters()):
            p.copy_(torch.randn(p.shape, generator=g) * 3)
        q.bn.running_mean.copy_(torch.randn(5, generator=g))
        q.bn.num_batches_tracked.fill_(12345678901)
    pq = [p.data.clone() for p in q.parameters()]
    pk = [p.data.clone() for p in k.parameters()]
    bq = [b.data.clone() for b in q.buffers()]
    bk = [b.data.clone() for b in k.buffers()]
    q, k = q.to(dev()), k.to(dev())
    # keep the misalignment of w_big on the device too
    for net in (q, k):
        padded = torch.empty(70001 + 3, device=dev())
        padded[3:].copy_(net.w_big.data)
        net.w_big.data = padded[3:]
    plan = WeightEma(q, k)
    assert plan.n_chunks >= 9 + 3 + 2
    for a in (0.999, 0.5, 0.9995):
        pk, bk = po.update_ema(pq, pk, bq, bk, a)
        plan.update(a)
    for got, want in zip(list(k.parameters()) + list(k.buffers()), pk + bk):
        assert got.dtype == want.dtype and torch.equal(got.data.cpu(), want)
    for got, want in zip(q.parameters(), pq):                                # the trained model is untouched
        assert torch.equal(got.data.cpu(), want)
    update_ema(q, k, 0.25)                                                   # functional form: builds and caches its own plan
    pk, bk = po.update_ema(pq, pk, bq, bk, 0.25)
    assert all(torch.equal(got.data.cpu(), want) for got, want in zip(k.parameters(), pk))
    with pytest.raises(RuntimeError, match="CUDA"):
        WeightEma(Net(), Net())
    # update_dynamic (prototypes.py:99-102: dynamic_model = deepcopy(model)): every parameter and buffer copied into the
    # existing module in one launch, the source untouched; a second call reuses the cached chunk table
    from onda_b200 import update_dynamic
    dyn = Net().to(dev())
    for _ in range(2):
        assert update_dynamic(q, dyn) is dyn
        for got, want in zip(list(dyn.parameters()) + list(dyn.buffers()), list(q.parameters()) + list(q.buffers())):
            assert got.dtype == want.dtype and torch.equal(got.data, want.data)
            assert got.numel() == 0 or got.data_ptr() != want.data_ptr()
        with torch.no_grad():
            q.w_small.add_(1.0)                                                 # the next snapshot must see the change
    assert all(torch.equal(got.data.cpu(), want) for got, want in zip(list(q.parameters())[1:2], pq[1:2]))


@pytest.mark.parametrize("h,w,H,W", [(9, 17, 65, 129), (10, 13, 37, 50), (33, 65, 257, 513), (6, 7, 6, 7)])
def test_confusion_meter_matches_reference_evaluation(h, w, H, W):
    """da_model.evaluate's counters (adaptation_model.py:143-160, func.py:77-79) from one fused kernel.  The per-pixel
    prediction equals the reference's argmax of softmax(interp(pred)) except where that softmax's top-2 margin is below
    1e-6 (interpolation rounding); the matrix is exactly fast_hist of our own prediction, and equal to the reference's
    when no exempt pixel flipped.  Labels carry 255 and -1 (ignored)."""
    import numpy as np
    from onda_b200 import ConfusionMeter
    g = torch.Generator().manual_seed(1000 + h * w)
    B, C = 3, 19
    meter = ConfusionMeter(C, dev())
    total_ref = np.zeros((C, C), dtype=np.int64)
    for batch in range(2):
        pred = torch.randn(B, C, h, w, generator=g) * 3
        pred[0, 4] = pred[0, 3]                                   # exact ties between two classes: first index wins
        labels = torch.randint(0, C, (B, H, W), generator=g)
        labels[torch.rand(B, H, W, generator=g) < 0.1] = 255
        labels[torch.rand(B, H, W, generator=g) < 0.02] = -1
        prob, ref_pred, ref_hist = po.eval_confusion(pred, labels, C, (H, W))
        before = meter.hist().copy()
        got = meter.update(pred.to(dev()), labels if batch == 0 else labels.to(dev()), return_prediction=True).cpu().long()
        top2 = prob.topk(2, dim=1)[0]
        exempt = (top2[:, 0] - top2[:, 1]) < 1e-6
        assert int(((got != ref_pred) & ~exempt).sum()) == 0
        a, b_ = labels.numpy().flatten(), got.numpy().flatten()
        k = (a >= 0) & (a < C)
        own = np.bincount(C * a[k] + b_[k], minlength=C * C).reshape(C, C)
        assert np.array_equal(meter.hist() - before, own)          # exact counting, accumulated over batches
        if int((got != ref_pred).sum()) == 0:
            assert np.array_equal(own, ref_hist)
        total_ref += ref_hist
    if np.array_equal(meter.hist(), total_ref):
        iu_ref = np.diag(total_ref) / (total_ref.sum(1) + total_ref.sum(0) - np.diag(total_ref) + np.finfo(float).eps)
        assert np.array_equal(meter.per_class_iu(), iu_ref)
    meter.reset()
    assert meter.hist().sum() == 0
    with pytest.raises(ValueError):
        meter.update(torch.zeros(1, 5, 4, 4, device=dev()), torch.zeros(1, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        meter.update(torch.zeros(1, C, 4, 4), torch.zeros(1, 8, 8))


def test_append_source_labels_matches_reference_gather_path():
    """calculate_prototypes with STARTING_PROTO == "source" (prototypes.py:142-154): nearest-resized labels, 255 masked out,
    class sums straight from the NCHW map -- against the reference's mask-gather + one-hot + append on the oracle."""
    g = torch.Generator().manual_seed(71)
    B, D, h, w, H, W, C = 3, 96, 17, 23, 129, 180, 19
    feats = [torch.randn(B, D, h, w, generator=g) * 2 for _ in range(3)]
    labs = []
    for _ in range(3):
        lab = torch.randint(0, C, (B, H, W), generator=g)
        lab[torch.rand(B, H, W, generator=g) < 0.2] = 255
        labs.append(lab)
    from onda_b200 import prototype_handler
    hd = prototype_handler(distance_metric="mahalanobis")
    orc = po.OracleHandler(distance_metric="mahalanobis")
    for feat, lab in zip(feats, labs):
        # the reference's own lines, on the CPU oracle
        labels_clone = torch.nn.functional.interpolate(lab.unsqueeze(1).float(), size=(h, w)).view(-1)
        mask = labels_clone != 255
        feat_clone = feat.permute(1, 0, 2, 3).reshape(D, -1)
        rows = feat_clone[:, mask].permute(1, 0)
        onehot = torch.nn.functional.one_hot(labels_clone[mask].long(), C)
        orc.append(rows, onehot)
        hd.append_source_labels(feat.to(dev()), lab.to(dev()), num_classes=C)
    assert torch.equal(hd.counter.cpu(), orc.counter)
    close_rel_max(hd.prototypes, orc.prototypes)
    close_rel_max(hd.squared_mean, orc.squared_mean)


@pytest.mark.parametrize("reg", ["MRKLD", "MRENT"])
def test_target_losses_golden(reg):
    """The fused CE + RCE + regulariser kernel (forward and gradient) against the real reference loss functions."""
    from onda_b200 import target_losses
    z = np.load(os.path.join(GOLDEN, "target_losses.npz"))
    out = T(z["out"]).to(dev()).requires_grad_(True)
    got = target_losses(out, T(z["labels"]).to(dev()), float(z["alpha"]), float(z["beta"]), float(z["reg_weight"]), reg)
    got["Total target loss"].backward()
    for ours, k in (("ce_loss", "ce"), ("rce_loss", "rce"), ("regularization_loss", "reg"), ("Total target loss", "total")):
        assert float(got[ours]) == pytest.approx(float(z[f"ref_{reg}_{k}"]), rel=2e-6), k
    gref = T(z[f"ref_{reg}_grad"])
    assert float((out.grad.cpu() - gref).abs().max()) <= 1e-5 * float(gref.abs().max())


def test_target_losses_full_size_against_oracle():
    """Config-3 sized logits (32 x 19 x 65 x 129) with the labels and the device-side label count of a fused pass."""
    from onda_b200 import target_losses
    case = po.synth_case(91, 8, 64, 65, 129)
    hd = make_handler(case["protos"], case["sq_mean"], case["counter"], "mahalanobis")
    labels, _ = hd.pseudo_labels_fused(case["feat"].to(dev()), case["prior"].to(dev()), case["out"].to(dev()))
    g = torch.Generator().manual_seed(92)
    student = torch.randn(8, 19, 65, 129, generator=g) * 3
    ref_in = student.clone().requires_grad_(True)
    ref = po.target_losses(ref_in, labels.cpu().view(8, 65, 129), 0.1, 1.0, 0.1, "MRKLD")
    ref["total"].backward()
    out = student.to(dev()).requires_grad_(True)
    got = target_losses(out, labels, 0.1, 1.0, 0.1, "MRKLD", n_valid=hd.last_pixel_count())
    (2.0 * got["Total target loss"]).backward()                       # upstream gradient != 1
    assert float(got["pseudolabel_pixel_num"]) == float(ref["n_valid"])
    assert float(got["output & prototype agreement"]) == pytest.approx(float(ref["agreement"]), abs=1e-7)
    for ours, k in (("ce_loss", "ce"), ("rce_loss", "rce"), ("regularization_loss", "reg"), ("Total target loss", "total")):
        assert float(got[ours]) == pytest.approx(float(ref[k]), rel=1e-5), k
    assert float((out.grad.cpu() - 2.0 * ref_in.grad).abs().max()) <= 1e-5 * float(2.0 * ref_in.grad.abs().max())
