"""BASELINE configs[4]: a multi-domain stream (clear -> rain -> fog, intensity steps) with per-domain prototype EMA,
hybrid switch statistics and the per-domain ``save`` / ``load`` of the prototype state
(framework/domain_adaptation/methods/prototypes.py:68-72, 124-126; configs/hybrid_switch_fog.yml).

The CUDA path (methods.hybrid_prototype_predictions + handler.ma through the C ABI) runs the whole stream next to the
CPU oracle (oracle.hybrid_prior + OracleHandler, both pinned to the real reference by tests/golden): identical selector
trace, labels off near-ties, soft predictions, Monitor entries and final prototypes -- including across a process-like
restart in the middle of the stream (fresh handler + load of the pickle written at the domain boundary)."""
import os

import numpy as np
import pytest
import torch

from oracle import proto_oracle as po

pytestmark = pytest.mark.gpu

C, D, H, W = 19, 64, 9, 17
DOMAINS = ["clear", "rain_25mm", "rain_50mm", "rain_100mm", "fog_150m", "fog_75m", "rain_50mm_again", "clear_again"]
STEPS = 14        # per domain; the Monitor window (12) fills inside the first domain


class _Spec(dict):
    def __getattr__(self, k):
        return self[k] if k in self else {}


class _Replay(torch.nn.Module):
    """A 'network' that returns precomputed (feat, out) for the image index it is given: the backbone is out of scope."""

    def __init__(self, table):
        super().__init__()
        self.table = table

    def forward(self, image):
        return None, self.table[int(image.flatten()[0].item())]


def _stream():
    """Per step: EMA feat/out, static and dynamic logits whose sharpness follows the domain's severity, so the static
    confidence crosses the gray area both ways over the stream."""
    first = po.synth_case(900, 1, D, H, W)
    steps, k = [], 0
    for di, _ in enumerate(DOMAINS):
        severity = [0.0, 0.35, 0.6, 0.9, 0.8, 1.0, 0.6, 0.0][di]
        for s in range(STEPS):
            sharp = 4.5 - 3.6 * severity + 0.15 * np.sin(k / 3.0)
            case = po.synth_case(1000 + k, 1, D, H, W, protos=first["protos"] + 0.01 * di, counter=first["counter"], sharp=float(sharp))
            g = torch.Generator().manual_seed(5000 + k)
            dyn = case["prior_logits"] * 1.4 + torch.randn(1, C, H, W, generator=g) * 0.5
            steps.append({"feat": case["feat"], "out": case["out"], "static": case["prior_logits"], "dynamic": dyn, "domain": di})
            k += 1
    return first, steps


def test_multi_domain_stream_with_per_domain_save_and_load(tmp_path):
    from onda_b200 import prototype_handler, Monitor, HybridSelect, methods
    dev = torch.device("cuda:0")
    first, steps = _stream()
    params = dict(ma_lambda=0.98, tau=1, thresh=0.3, distance_metric="mahalanobis")
    gray, dev_thresh, limit = (0.55, 0.75), 0.002, 12

    # ---- oracle run
    orc = po.OracleHandler(**params)
    orc.prototypes, orc.squared_mean, orc.counter = first["protos"].clone(), first["sq_mean"].clone(), first["counter"].clone()
    omon, osel = po.OracleMonitor(limit, 0.003, "hamming"), po.OracleHybridSelect(0, gray, dev_thresh)
    ref = []
    for st in steps:
        prior = po.hybrid_prior(st["out"], st["static"], lambda: st["dynamic"], omon, osel, 0, 1, 1)
        labels = orc.pseudo_labels(st["feat"], prior, confidence_monitor=omon)
        soft = orc.pseudo_labels(st["feat"], prior, soft=True)
        orc.ma(st["feat"], st["out"])
        ref.append((labels, soft, osel.current, omon.window["prior static"][-1]))
    assert len({r[2] for r in ref}) == 2, "the scripted stream must exercise both selector states"

    # ---- CUDA run, with a save at every domain boundary and a restart from the pickle after domain 3
    class Method:
        pass

    def fresh_handler():
        return prototype_handler(**params)

    m = Method()
    m.device = dev
    m.cfg_spec = _Spec(EMA_LAMBDA=0, STATIC_LAMBDA=1, DYNAMIC_LAMBDA=1)
    m.intensity_ma = Monitor(limit, 0.003, "hamming")
    m.model_select = HybridSelect(HybridSelect.static, gray, dev_thresh)
    ema_tab = [{"feat": st["feat"].to(dev), "out": st["out"].to(dev)} for st in steps]
    m.ema_model = _Replay(ema_tab)
    m.static_model = _Replay([{"feat": None, "out": st["static"].to(dev)} for st in steps])
    m.dynamic_model = _Replay([{"feat": None, "out": st["dynamic"].to(dev)} for st in steps])
    m.prototypes = fresh_handler()
    m.prototypes.prototypes, m.prototypes.squared_mean, m.prototypes.counter = (first[k].to(dev) for k in ("protos", "sq_mean", "counter"))
    exempt = 0
    for i, st in enumerate(steps):
        if i > 0 and st["domain"] != steps[i - 1]["domain"]:
            path = os.path.join(tmp_path, f"proto_{DOMAINS[steps[i - 1]['domain']]}.pickle")
            m.prototypes.save(path)                                   # prototypes.py:124-126
            if st["domain"] == 4:                                     # "restart": a new handler picks the state up (:68-70)
                m.prototypes = fresh_handler()
                assert m.prototypes.load(path) is True
        pred = methods.hybrid_prototype_predictions(m, {"image": torch.tensor([float(i)])})
        m.prototypes.ma(pred["ema_model"]["feat"], pred["ema_model"]["out"])
        labels, soft, sel, conf = ref[i]
        assert m.model_select.current == sel, f"selector differs at step {i}"
        assert m.intensity_ma.current_dict["prior static"][-1] == pytest.approx(conf, abs=2e-6)
        assert float((pred["soft_predictions"].cpu() - soft).abs().max()) <= 1e-5
        lab = pred["pseudolabels"].cpu().flatten()
        top2 = soft.topk(2, dim=1)[0]
        near = ((top2[:, 0] - top2[:, 1]) < 1e-6) | ((top2[:, 0] - 0.3).abs() < 1e-6)
        assert int(((lab != labels.flatten()) & ~near).sum()) == 0
        exempt += int(near.sum())
    scale = float(orc.prototypes.abs().max())
    assert float((m.prototypes.prototypes.cpu() - orc.prototypes).abs().max()) <= 1e-5 * scale
    assert float((m.prototypes.squared_mean.cpu() - orc.squared_mean).abs().max()) <= 1e-5 * float(orc.squared_mean.abs().max())
    assert sorted(os.listdir(tmp_path)) == sorted(f"proto_{d}.pickle" for d in DOMAINS[:-1])
