"""Generate the golden fixtures in this directory from the REAL reference.

Run in the authoring container only (it imports theo2021/OnDA from
/root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Every ``*.npz`` written here stores the inputs *and* what the unmodified
reference (`framework.domain_adaptation.methods.prototype_handler`,
`framework.utils.monitoring.Monitor`, the `model_select` classes and
`hybrid_proDA.prototype_predictions`) returned for them on CPU, fp32,
torch 2.11.  The tests pin both the oracle restatement (`oracle/`) and the CUDA
path to these files.  Nothing here is imported by the product.
"""
from __future__ import annotations

import io
import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("ONDA_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from framework.domain_adaptation.methods.prototype_handler import prototype_handler  # noqa: E402
from framework.utils.monitoring import Monitor  # noqa: E402
from framework.utils.func import prob_2_entropy  # noqa: E402
import framework.domain_adaptation.methods.prototypes_hybrid_switch as ref_hybrid  # noqa: E402
import framework.domain_adaptation.methods.prototypes_vswitch as ref_vswitch  # noqa: E402
import framework.domain_adaptation.methods.prototypes_hswitch as ref_hswitch  # noqa: E402
import framework.domain_adaptation.methods.prototypes as ref_base  # noqa: E402

from oracle.proto_oracle import synth_case  # noqa: E402  (seeded input generator only)

torch.set_num_threads(4)


def npz(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


# ---------------------------------------------------------------------------
def legacy_pickle():
    """The shipped prototypes.pickle: 2-tuple (P[19,256], counter[19]) saved from CUDA."""

    class CpuUnpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module == "torch.storage" and name == "_load_from_bytes":
                return lambda b: torch.load(io.BytesIO(b), map_location="cpu", weights_only=False)
            return super().find_class(module, name)

    with open(os.path.join(REF, "prototypes.pickle"), "rb") as f:
        protos, counter = CpuUnpickler(f).load()
    return protos.float().contiguous(), counter.float().contiguous()


def make_handler(metric, case, ma_lambda, tau, thresh):
    h = prototype_handler(ma_lambda=ma_lambda, tau=tau, thresh=thresh, distance_metric=metric)
    h.prototypes = case["protos"].clone()
    h.squared_mean = case["sq_mean"].clone()
    h.counter = case["counter"].clone()
    return h


def ops_case(name, seed, b, d, h, w, metric, tau=1.0, thresh=0.3, ma_lambda=0.9995,
             protos=None, counter=None, tweak=None):
    case = synth_case(seed, b, d, h, w, protos=protos, counter=counter)
    if tweak is not None:
        tweak(case)
    hd = make_handler(metric, case, ma_lambda, tau, thresh)
    mon = Monitor(200, 0.003, "hamming")
    dist = hd.distance_measure(case["feat"])
    labels = hd.pseudo_labels(case["feat"], case["prior"], confidence_monitor=mon)
    soft = hd.pseudo_labels(case["feat"], case["prior"], soft=True)
    stat_proto = float(mon.current_dict["prototypes"][0])
    s1, cnt = hd.get_proto_array(case["feat"], case["out"])
    s2, _ = hd.get_proto_array(case["feat"] ** 2, case["out"])
    gvar = hd.global_var()
    pvar = hd.prototype_var()
    hd.ma(case["feat"], case["out"])
    npz(name, feat=case["feat"], prior=case["prior"], out=case["out"], protos=case["protos"],
        sq_mean=case["sq_mean"], counter=case["counter"],
        metric=np.array(metric), tau=np.float64(tau), thresh=np.float64(thresh),
        ma_lambda=np.float64(ma_lambda),
        ref_dist=dist, ref_labels=labels, ref_soft=soft, ref_stat_proto=np.float64(stat_proto),
        ref_stat_prior=np.float64(case["prior"].max(axis=1)[0].mean().item()),
        ref_stat_pl=np.float64(soft.max(axis=1)[0].mean().item()),
        ref_sum=s1, ref_sumsq=s2, ref_count=cnt, ref_global_std=gvar, ref_class_std=pvar,
        ref_ma_protos=hd.prototypes, ref_ma_sq_mean=hd.squared_mean)


def tweak_edge(case):
    """Adversarial rows: exact ties, an all-one-class image, an empty class,
    a prior exactly at the threshold boundary, zeroed channels."""
    feat, out, prior = case["feat"], case["out"], case["prior"]
    # pixel 0: sits exactly on prototype 3 (distance 0 to it)
    feat[0, :, 0, 0] = case["protos"][3]
    # pixel 1: equidistant (mirror) from prototypes 5 and 7 -> near tie on distance
    feat[0, :, 0, 1] = 0.5 * (case["protos"][5] + case["protos"][7])
    # flat prior rows (all classes equal) and an exactly tied logit row (first index wins)
    prior[0, :, 0, 2] = 1.0 / prior.shape[1]
    out[0, :, 0, 3] = 0.0
    out[0, 4, 0, 4] = out[0, 9, 0, 4] = 50.0
    # class 11 never wins the EMA-logit argmax (empty class in ma)
    out[:, 11] = -100.0
    # second image: a single class everywhere
    if out.shape[0] > 1:
        out[1] = -5.0
        out[1, 2] = 5.0
    # a prior that is one-hot (drives r to exactly 1.0 / 0.0)
    prior[0, :, 1, 0] = 0.0
    prior[0, 6, 1, 0] = 1.0


def nan_rows_case():
    """Rows where the reference's arithmetic degenerates (prototype_handler.py:159-166): a prior that is zero for every
    class (0 / 0 -> NaN row, torch.max gives (NaN, 0), `NaN < thresh` is False -> label 0), a NaN feature vector (NaN
    distances -> NaN row), and a negative tau (the softmax prefers the FARTHEST prototype).  Stored under a name the
    ops_* globs do not match: the CPU oracle is pinned to it; the CUDA path is compared with the oracle on such rows in
    tests/test_gpu_parity.py::test_zero_rectified_row_keeps_label_zero_like_the_reference."""
    out = {}
    for tag, tau in (("pos", 1.0), ("neg", -0.8)):
        case = synth_case(61, 1, 32, 4, 5)
        case["prior"][0, :, 1, 2] = 0.0
        case["feat"][0, :, 2, 3] = float("nan")
        hd = make_handler("mahalanobis", case, 0.9995, tau, 0.3)
        labels = hd.pseudo_labels(case["feat"], case["prior"])
        soft = hd.pseudo_labels(case["feat"], case["prior"], soft=True)
        out.update({f"ref_labels_{tag}": labels, f"ref_soft_{tag}": soft})
    npz("edge_nan_rows.npz", feat=case["feat"], prior=case["prior"], protos=case["protos"], sq_mean=case["sq_mean"],
        counter=case["counter"], **out)


def append_case():
    g = torch.Generator().manual_seed(77)
    hd = prototype_handler(distance_metric="mahalanobis")
    steps = []
    for i in range(3):
        case = synth_case(100 + i, 2, 48, 9, 13)
        hd.append(case["feat"], case["out"])
        steps.append((case["feat"], case["out"]))
    # the 2-D "source" path: (M,D) rows with an int64 one-hot (prototypes.py:142-153)
    m, d, c = 333, 48, 19
    rows = torch.randn(m, d, generator=g) * 2
    lab = torch.randint(0, c, (m,), generator=g)
    hot = torch.nn.functional.one_hot(lab, c)  # int64
    hd.append(rows, hot)
    npz("append_seq.npz",
        feat0=steps[0][0], out0=steps[0][1], feat1=steps[1][0], out1=steps[1][1],
        feat2=steps[2][0], out2=steps[2][1], rows3=rows, hot3=hot,
        ref_protos=hd.prototypes, ref_sq_mean=hd.squared_mean, ref_counter=hd.counter)


def confidence_stream(n, seed):
    """A scripted 'prior static' confidence stream that crosses the gray area both ways."""
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    base = 0.865 + 0.06 * np.sin(2 * np.pi * t / 700.0) - 0.00004 * t
    drop = np.where((t > 900) & (t < 1300), -0.05, 0.0)
    return (base + drop + rng.normal(0, 0.004, n)).astype(np.float64)


def monitor_trace():
    n = 1800
    conf = confidence_stream(n, 5)
    for tag, args in (("hamming", (200, 0.003, "hamming")), ("median", (60, 0.01, "median")),
                      ("mean", (31, 0.05, "mean"))):
        mon = Monitor(*args)
        sel = ref_hybrid.model_select(ref_hybrid.model_select.static, (0.83, 0.9), 0.0002)
        vsel = ref_vswitch.model_select(ref_vswitch.model_select.static, 0.00028)
        cur, cur_dev, dev, med, ema, vcur, pct = [], [], [], [], [], [], []
        for v in conf:
            mon.add({"prior static": float(v)})
            d = mon.dev_avg("prior static")
            sel.evaluate(mon.avg("prior static"), d)
            vsel.evaluate(d)
            cur.append(sel.current)
            cur_dev.append(sel.current_dev)
            dev.append(d)
            med.append(mon.avg("prior static"))
            ema.append(mon.exp("prior static"))
            vcur.append(vsel.current)
            vl = mon.avg("prior static")
            vt = torch.tensor(vl, dtype=torch.float32)    # hswitch_proDA keeps 0-dim float32 tensors in its Monitor
            pct.append(float(max(min(vt * (25.0 / 3) - (41.0 / 6), 1), 0)))  # prototypes_hswitch.py:47
        npz(f"monitor_trace_{tag}.npz", conf=conf, limit=np.int64(args[0]), exp_const=np.float64(args[1]),
            dev_func=np.array(args[2]), gray=np.array([0.83, 0.9]), dev_thresh=np.float64(0.0002),
            vthresh=np.float64(0.00028),
            ref_current=np.array(cur), ref_current_dev=np.array(cur_dev), ref_dev=np.array(dev, dtype=np.float64),
            ref_median=np.array(med, dtype=np.float64), ref_exp=np.array(ema, dtype=np.float64),
            ref_vcurrent=np.array(vcur), ref_pct=np.array(pct, dtype=np.float64),
            ref_missing_avg=np.float64(Monitor(5).avg("nope")), ref_missing_exp=np.float64(Monitor(5).exp("nope")),
            ref_missing_dev=np.float64(Monitor(5).dev_avg("nope")))


def sequence_case():
    """260 pseudo-label -> ma steps with drifting inputs; inputs are regenerated
    from seeds by the tests (synth_case(seed=5000+i, 1, 32, 9, 11))."""
    steps, d = 260, 32
    first = synth_case(4999, 1, d, 9, 11)
    hd = prototype_handler(ma_lambda=0.95, tau=1.0, thresh=0.3, distance_metric="mahalanobis")
    hd.prototypes, hd.squared_mean, hd.counter = first["protos"].clone(), first["sq_mean"].clone(), first["counter"].clone()
    mon = Monitor(50, 0.003, "hamming")
    stat, npl, lab_hash = [], [], []
    for i in range(steps):
        case = synth_case(5000 + i, 1, d, 9, 11, protos=first["protos"] + 0.002 * i, counter=first["counter"])
        labels = hd.pseudo_labels(case["feat"], case["prior"], confidence_monitor=mon)
        soft = hd.pseudo_labels(case["feat"], case["prior"], soft=True)
        mon.add({"pseudolabel confidence": soft.max(axis=1)[0].mean()})
        hd.ma(case["feat"], case["out"])
        stat.append(float(mon.current_dict["prototypes"][-1]))
        npl.append(int((labels != 255).sum()))
        lab_hash.append(int((labels.flatten() * torch.arange(1, labels.numel() + 1)).sum()))
    npz("sequence_ma.npz", steps=np.int64(steps), d=np.int64(d), ref_protos=hd.prototypes,
        ref_sq_mean=hd.squared_mean, ref_stat=np.array(stat), ref_npl=np.array(npl),
        ref_label_hash=np.array(lab_hash, dtype=np.int64),
        ref_dev_proto=np.float64(mon.dev_avg("prototypes")), ref_med_proto=np.float64(mon.avg("prototypes")))


# ---------------------------------------------------------------------------
# the method class itself: hybrid_proDA.prototype_predictions with a fake model
# ---------------------------------------------------------------------------
class AttrDict(dict):
    """Stand-in for addict.Dict (not installed): missing keys read as {}."""

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return self[k]

    def __setattr__(self, k, v):
        self[k] = v

    def __missing__(self, k):
        v = AttrDict()
        self[k] = v
        return v


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    return d


class FakeSegModel(torch.nn.Module):
    """forward(x) -> (None, {"feat": (B,D,h,w), "out": (B,19,h,w)}), deep-copyable."""

    def __init__(self, d=24, c=19, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.stem = torch.nn.Conv2d(3, d, 3, stride=8, padding=1)
        self.head = torch.nn.Conv2d(d, c, 1)
        with torch.no_grad():
            self.stem.weight.copy_(torch.randn(self.stem.weight.shape, generator=g) * 0.8)
            self.stem.bias.copy_(torch.randn(d, generator=g) * 0.3)
            self.head.weight.copy_(torch.randn(self.head.weight.shape, generator=g) * 1.2)
            self.head.bias.zero_()

    def forward(self, x):
        feat = self.stem(x)
        return None, {"feat": feat, "out": self.head(feat)}

    def optim_parameters(self, lr):
        return [{"params": self.parameters(), "lr": lr}]


def method_case():
    import yaml
    with open(os.path.join(REF, "configs", "hybrid_switch.yml")) as f:
        cfg = to_attr(yaml.safe_load(f))
    cfg.OTHERS.DEVICE = "cpu"
    cfg.NUM_CLASSES = 19
    cfg.OTHERS.SNAPSHOT_DIR = "/tmp/onda_golden_snap"
    spec = cfg.METHOD.ADAPTATION.PROTO_ONLINE_HYBRIDSWITCH
    spec.set_ = "golden"
    spec.LOAD_PROTO = AttrDict()
    spec.AVG_MONITOR_SIZE = 12           # fill the window quickly
    spec.GRAY_AREA = [0.66, 0.78]
    spec.DEV_THRESH = 0.002
    model = FakeSegModel()
    method = ref_hybrid.hybrid_proDA(model, cfg, spec)
    for m in (method.ema_model, method.dynamic_model, method.static_model, method.model):
        m.eval()
    # perturb the dynamic model so the two priors differ
    with torch.no_grad():
        method.dynamic_model.head.weight.mul_(1.5)
    d, c = 24, 19
    g = torch.Generator().manual_seed(9)
    first = synth_case(31, 1, d, 5, 5)
    method.prototypes.prototypes = first["protos"].clone() * 0.2
    method.prototypes.squared_mean = (first["protos"] * 0.2) ** 2 + 1.0
    method.prototypes.counter = first["counter"].clone()
    steps = 56
    images, labs, softs, sel, stat = [], [], [], [], {k: [] for k in ("prior static", "prior", "prototypes", "pseudolabel confidence")}
    for i in range(steps):
        scale = 0.07 + 0.3 * (0.5 + 0.5 * np.sin(i / 5.0))   # logit sharpness -> confidence wave
        img = torch.randn(1, 3, 32, 56, generator=g) * float(scale)
        pred = method.prototype_predictions({"image": img})
        method.prototypes.ma(pred["ema_model"]["feat"], pred["ema_model"]["out"])
        images.append(img)
        labs.append(pred["pseudolabels"])
        softs.append(pred["soft_predictions"])
        sel.append(method.model_select.current)
        for k in stat:
            stat[k].append(float(method.intensity_ma.current_dict[k][-1]))
    state = {}
    for nm, m in (("ema", method.ema_model), ("static", method.static_model), ("dynamic", method.dynamic_model)):
        for k, v in m.state_dict().items():
            state[f"w_{nm}_{k.replace('.', '_')}"] = v
    npz("method_hybrid.npz", images=torch.stack(images), ref_labels=torch.stack(labs), ref_soft=torch.stack(softs),
        ref_select=np.array(sel), init_protos=first["protos"] * 0.2, init_sq_mean=(first["protos"] * 0.2) ** 2 + 1.0,
        init_counter=first["counter"], ref_final_protos=method.prototypes.prototypes,
        ref_final_sq_mean=method.prototypes.squared_mean,
        gray=np.array(spec.GRAY_AREA), dev_thresh=np.float64(spec.DEV_THRESH), limit=np.int64(spec.AVG_MONITOR_SIZE),
        exp_const=np.float64(spec.EXP_MONITOR_CONST), ma_lambda=np.float64(spec.MA_LAMBDA), tau=np.float64(spec.TAU),
        thresh=np.float64(spec.PSEUDO_THRESH),
        **{f"ref_stat_{k.replace(' ', '_')}": np.array(v) for k, v in stat.items()}, **state)


def method_variant_case(name, yml, key, cls, steps=44, wave=(0.07, 0.3), overrides=None):
    """The other three method classes (a11): the REAL hswitch_proDA / vswitch_proDA / online_proDA
    ``prototype_predictions`` followed by ``ma`` over a stream whose logit sharpness (hence the static confidence)
    rises and falls, so the switch takes both branches."""
    import yaml
    with open(os.path.join(REF, "configs", yml)) as f:
        cfg = to_attr(yaml.safe_load(f))
    cfg.OTHERS.DEVICE = "cpu"
    cfg.NUM_CLASSES = 19
    cfg.OTHERS.SNAPSHOT_DIR = "/tmp/onda_golden_snap"
    spec = cfg.METHOD.ADAPTATION[key]
    spec.set_ = "golden"
    spec.LOAD_PROTO = AttrDict()
    spec.AVG_MONITOR_SIZE = 12
    for k, v in (overrides or {}).items():
        spec[k] = v
    method = cls(FakeSegModel(), cfg, spec)
    for m in (method.ema_model, method.dynamic_model, method.static_model, method.model):
        m.eval()
    with torch.no_grad():
        method.dynamic_model.head.weight.mul_(1.5)
    seed = 10 + len(name)
    rs = np.random.RandomState(seed)      # legacy stream: stable across numpy versions, so only the seed is stored
    first = synth_case(31, 1, 24, 5, 5)
    method.prototypes.prototypes = first["protos"].clone() * 0.2
    method.prototypes.squared_mean = (first["protos"] * 0.2) ** 2 + 1.0
    method.prototypes.counter = first["counter"].clone()
    keys = ["prior static", "prior", "prototypes", "pseudolabel confidence"]
    images, labs, softs, sel, stat, share = [], [], [], [], {k: [] for k in keys}, []
    for i in range(steps):
        scale = wave[0] + wave[1] * (0.5 + 0.5 * np.sin(i / 5.0))
        img = torch.from_numpy(rs.standard_normal((1, 3, 32, 56)).astype(np.float32) * np.float32(scale))
        pred = method.prototype_predictions({"image": img, "label": 0})
        method.prototypes.ma(pred["ema_model"]["feat"], pred["ema_model"]["out"])
        images.append(img)
        labs.append(pred["pseudolabels"])
        softs.append(pred["soft_predictions"])
        sel.append(method.model_select.current if hasattr(method, "model_select") else -1)
        cur = method.intensity_ma.current_dict
        share.append(float(cur["percentage_static"][-1]) if "percentage_static" in cur else -1.0)
        for k in keys:
            stat[k].append(float(cur[k][-1]))
    state = {}
    for nm, m in (("ema", method.ema_model), ("static", method.static_model), ("dynamic", method.dynamic_model)):
        for k, v in m.state_dict().items():
            state[f"w_{nm}_{k.replace('.', '_')}"] = v
    n_dyn = len(method.intensity_ma.current_dict.get("prior dynamic", []))
    print(name, "static conf range", min(stat["prior static"]), max(stat["prior static"]), "select", sel, "share", share,
          "dynamic forwards recorded", n_dyn)
    thr = spec.SWITCH_PRIOR_THRESH
    npz(name, image_seed=np.int64(seed), image_scales=np.array([wave[0] + wave[1] * (0.5 + 0.5 * np.sin(i / 5.0)) for i in range(steps)]),
        image_checksum=np.float64(torch.stack(images).double().sum().item()),
        ref_labels=torch.stack(labs), ref_soft=torch.stack(softs),
        ref_select=np.array(sel), ref_share=np.array(share), init_protos=first["protos"] * 0.2,
        init_sq_mean=(first["protos"] * 0.2) ** 2 + 1.0, init_counter=first["counter"],
        ref_final_protos=method.prototypes.prototypes, ref_final_sq_mean=method.prototypes.squared_mean,
        limit=np.int64(spec.AVG_MONITOR_SIZE), exp_const=np.float64(spec.EXP_MONITOR_CONST),
        ma_lambda=np.float64(spec.MA_LAMBDA), tau=np.float64(spec.TAU), thresh=np.float64(spec.PSEUDO_THRESH),
        ema_lambda=np.float64(spec.EMA_LAMBDA), static_lambda=np.float64(spec.STATIC_LAMBDA),
        dynamic_lambda=np.float64(spec.DYNAMIC_LAMBDA), soft_trans=np.int64(1 if spec.SOFT_TRANS is True else 0),
        switch_prior_thresh=np.float64(thr if not isinstance(thr, dict) else 0.0),
        **{f"ref_stat_{k.replace(' ', '_')}": np.array(v) for k, v in stat.items()}, **state)


def method_variants():
    H, V, B = ref_hswitch.hswitch_proDA, ref_vswitch.vswitch_proDA, ref_base.online_proDA
    # h-switch, soft transition: the ramp 25/3 * median - 41/6 sweeps (0.82, 0.94)
    method_variant_case("method_hswitch_soft.npz", "confidence_switch.yml", "PROTO_ONLINE_HSWITCH", H, wave=(0.25, 0.9))
    # h-switch, hard transition at SWITCH_PRIOR_THRESH, with an EMA share in the mix
    method_variant_case("method_hswitch_hard.npz", "confidence_switch.yml", "PROTO_ONLINE_HSWITCH", H, wave=(0.25, 0.9),
                        overrides={"SOFT_TRANS": False, "SWITCH_PRIOR_THRESH": 0.88, "EMA_LAMBDA": 0.25})
    # v-switch: derivative of the static confidence against SWITCH_PRIOR_THRESH (its threshold_c)
    method_variant_case("method_vswitch.npz", "confidence_der_switch.yml", "PROTO_ONLINE_VSWITCH", V,
                        overrides={"SWITCH_PRIOR_THRESH": 0.002})
    # base class: additive three-way mix (threshold 0) and the replace / skip rule (threshold > 0)
    method_variant_case("method_base_mix.npz", "dynamic_model.yml", "PROTO_ONLINE", B,
                        overrides={"STATIC_LAMBDA": 0.5, "EMA_LAMBDA": 0.25, "DYNAMIC_LAMBDA": 0.25})
    method_variant_case("method_base_rule.npz", "dynamic_model.yml", "PROTO_ONLINE", B, wave=(0.25, 0.9),
                        overrides={"STATIC_LAMBDA": 1, "SWITCH_PRIOR_THRESH": 0.88})


def losses_case():
    """f1: the REAL cross_entropy_2d / rce (framework/utils/loss.py:16-45, 88-112) and regular_loss
    (prototypes.py:29-39) on hard pseudo-labels, with the autograd gradient of the weighted total of
    pseudolabel_loss (prototypes.py:313-328) with respect to the student logits."""
    from framework.utils.loss import cross_entropy_2d, rce
    g = torch.Generator().manual_seed(81)
    B, C, h, w = 2, 19, 9, 13
    out = (torch.randn(B, C, h, w, generator=g) * 2.5).requires_grad_(True)
    lab = torch.randint(0, C, (B, h, w), generator=g)
    lab[torch.rand(B, h, w, generator=g) < 0.3] = 255
    res = {}
    for reg in ("MRKLD", "MRENT"):
        out.grad = None
        ce = cross_entropy_2d(out, lab.long(), False)
        r = rce(out, lab, "cpu", soft=False)
        rg = ref_base.regular_loss(reg, out)
        total = 0.1 * ce + 1.0 * r + 0.1 * rg                 # RCE_ALPHA, RCE_BETA, REGULARIZER_WEIGHT of the YAMLs
        total.backward()
        res[reg] = (ce.item(), r.item(), rg.item(), total.item(), out.grad.clone())
    npz("target_losses.npz", out=out.detach(), labels=lab, alpha=np.float64(0.1), beta=np.float64(1.0), reg_weight=np.float64(0.1),
        **{f"ref_{k}_{n}": v for k, (a, b, c, d, gr) in res.items() for n, v in (("ce", a), ("rce", b), ("reg", c), ("total", d), ("grad", gr))})


def stats_case():
    """Switch statistics and the entropy map on raw logits (K4 parity)."""
    case = synth_case(61, 2, 8, 11, 17)
    g = torch.Generator().manual_seed(62)
    la, lb, lc = (torch.randn(2, 19, 11, 17, generator=g) * s for s in (3.0, 1.0, 6.0))
    pa, pb, pc = la.softmax(1), lb.softmax(1), lc.softmax(1)
    mix = 0.25 * pa + 1.0 * pb
    pct = 0.4
    hmix = pct * (0.0 * pa + 1.0 * pb) + (1 - pct) * 1.0 * pc
    npz("stats_logits.npz", la=la, lb=lb, lc=lc,
        ref_conf=np.array([p.max(axis=1)[0].mean().item() for p in (pa, pb, pc)]),
        ref_mix=mix, ref_mix_conf=np.float64(mix.max(axis=1)[0].mean().item()),
        ref_hmix=hmix, ref_hmix_conf=np.float64(hmix.max(axis=1)[0].mean().item()), pct=np.float64(pct),
        ref_entropy=prob_2_entropy(pa))


def widened_case():
    """Rows widened beyond the prototype path: the reference's own update_ema (prototypes.py:407-416) on a real
    hybrid_proDA instance, and the inner loop of da_model.evaluate (adaptation_model.py:143-160) with the reference's
    own ``interp`` module, ``fast_hist`` and ``per_class_iu``."""
    import yaml
    from framework.utils.func import fast_hist, per_class_iu
    with open(os.path.join(REF, "configs", "hybrid_switch.yml")) as f:
        cfg = to_attr(yaml.safe_load(f))
    cfg.OTHERS.DEVICE = "cpu"
    cfg.NUM_CLASSES = 19
    cfg.OTHERS.SNAPSHOT_DIR = "/tmp/onda_golden_snap"
    cfg.SCHEME.RESOLUTION = [56, 32]              # (W, H) of the full-resolution label maps of this fixture
    spec = cfg.METHOD.ADAPTATION.PROTO_ONLINE_HYBRIDSWITCH
    spec.set_ = "golden"
    spec.LOAD_PROTO = AttrDict()

    class FakeWithBuffers(FakeSegModel):
        def __init__(self, seed=0):
            super().__init__(seed=seed)
            self.bn = torch.nn.BatchNorm2d(24)

    model = FakeWithBuffers(seed=3)
    method = ref_hybrid.hybrid_proDA(model, cfg, spec)
    g = torch.Generator().manual_seed(77)
    with torch.no_grad():                          # a trained model that has moved away from its EMA copy
        for p in method.model.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
        method.model.bn.running_mean.add_(torch.randn(24, generator=g))
        method.model.bn.running_var.mul_(1.3)
        method.model.bn.num_batches_tracked.fill_(41)
    out = {"ema_update": np.float64(spec.EMA_UPDATE)}
    for i, (q, k) in enumerate(zip(method.model.parameters(), method.ema_model.parameters())):
        out[f"q{i}"], out[f"k{i}_before"] = q.data.clone(), k.data.clone()
    for i, (q, k) in enumerate(zip(method.model.buffers(), method.ema_model.buffers())):
        out[f"bq{i}"], out[f"bk{i}_before"] = q.data.clone(), k.data.clone()
    for step in range(3):                          # three consecutive updates
        method.update_ema()
    for i, k in enumerate(method.ema_model.parameters()):
        out[f"k{i}_after3"] = k.data.clone()
    for i, k in enumerate(method.ema_model.buffers()):
        out[f"bk{i}_after3"] = k.data.clone()
    out["n_params"] = np.int64(len(list(method.model.parameters())))
    out["n_buffers"] = np.int64(len(list(method.model.buffers())))
    # evaluation inner loop with the reference's own objects
    pred = torch.randn(3, 19, 4, 7, generator=g) * 3
    labels = torch.randint(0, 19, (3, 32, 56), generator=g)
    labels[torch.rand(3, 32, 56, generator=g) < 0.1] = 255
    prob = method.interp(pred).softmax(axis=1)
    counter = 0
    preds = []
    for item_pred, label in zip(prob, labels):
        lab = label.numpy()
        item_pred_labels = item_pred.permute(1, 2, 0).argmax(dim=2).cpu().numpy()
        counter = counter + fast_hist(lab.flatten(), item_pred_labels.flatten(), cfg.NUM_CLASSES)
        preds.append(item_pred_labels)
    out.update(eval_pred=pred, eval_labels=labels, eval_prob=prob, eval_argmax=np.stack(preds), eval_hist=counter,
               eval_iu=per_class_iu(counter))
    npz("widened_rows.npz", **out)


def main():
    widened_case()
    p_legacy, c_legacy = legacy_pickle()
    npz("prototypes_legacy.npz", protos=p_legacy, counter=c_legacy)
    ops_case("ops_euclid_small.npz", 11, 2, 48, 9, 13, "euclidean")
    ops_case("ops_mahal_small.npz", 12, 2, 48, 9, 13, "mahalanobis")
    ops_case("ops_mahal_tau.npz", 13, 1, 40, 7, 21, "mahalanobis", tau=0.37, thresh=0.55, ma_lambda=0.9)
    ops_case("ops_mahal_edge.npz", 14, 2, 32, 6, 9, "mahalanobis", tweak=tweak_edge)
    ops_case("ops_euclid_edge.npz", 15, 2, 32, 6, 9, "euclidean", thresh=0.0, tweak=tweak_edge)
    ops_case("ops_mahal_d256_legacy.npz", 16, 1, 256, 17, 21, "mahalanobis", protos=p_legacy, counter=c_legacy)
    ops_case("ops_mahal_d2048.npz", 17, 1, 2048, 9, 10, "mahalanobis")
    append_case()
    nan_rows_case()
    monitor_trace()
    sequence_case()
    stats_case()
    method_case()
    method_variants()
    losses_case()


if __name__ == "__main__":
    main()
