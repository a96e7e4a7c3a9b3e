"""``prototype_predictions`` of OnDA's prototype method classes, on the fused kernels.

Each function takes the method object itself (duck-typed: ``ema_model``, ``static_model``,
``dynamic_model``, ``cfg_spec``, ``intensity_ma``, ``prototypes``, ``device``, ``record_ece``,
and ``model_select`` for the v/hybrid switches) and returns the same dict as the reference:
``{"ema_model": {...}, "pseudolabels": (N,1) int64, "soft_predictions": (N,C) float32}``.
They can be bound onto the reference classes unchanged, e.g.

    from onda_b200 import methods
    hybrid_proDA.prototype_predictions = methods.hybrid_prototype_predictions

The network forwards stay stock PyTorch.  What changes: the softmax/max/mean chains become one
``prior_mix`` kernel per decision point, the two ``pseudo_labels`` calls become one fused pass
which also stages the class sums for the ``ma`` call that follows in ``pseudolabel_loss``
(prototypes.py:291-294), and Monitor entries are Python floats.

References: base prototypes.py:208-273; h-switch prototypes_hswitch.py:26-85; v-switch
prototypes_vswitch.py:36-87; hybrid prototypes_hybrid_switch.py:45-101.
"""
from __future__ import annotations

import torch

from .switching import below_f32, static_share


def _missing(v):
    """addict-style configs read absent keys as an empty dict."""
    return isinstance(v, dict) and len(v) == 0


def _record_ece(method, name, logits, label):
    """ECE is off in every shipped YAML (ECE_SKIP: True); when on, hand the softmax over."""
    if getattr(method, "ece_record", False):
        method.record_ece(name, logits.softmax(axis=1), label)


def _ema_and_static(method, batch):
    image = batch["image"].to(method.device)
    spec = method.cfg_spec
    _, pred_ema = method.ema_model(image)
    out_static = None
    if spec.STATIC_LAMBDA > 0:
        _, pred_static = method.static_model(image)
        out_static = pred_static["out"]
    return image, pred_ema, out_static


def _finish(method, batch, pred_ema, prior, prior_conf):
    """The part common to all variants: prior stat, fused pseudo-labels, confidence stat."""
    monitor = method.intensity_ma
    monitor.add({"prior": prior_conf})
    labels, soft = method.prototypes.pseudo_labels_fused(
        pred_ema["feat"], prior, pred_ema["out"], confidence_monitor=monitor)
    if not monitor.freeze:
        monitor.add({"pseudolabel confidence": method.prototypes.last_stats["pseudolabel confidence"]})
    return {"ema_model": pred_ema, "pseudolabels": labels, "soft_predictions": soft}


def hybrid_prototype_predictions(method, batch):
    """hybrid_proDA.prototype_predictions (prototypes_hybrid_switch.py:45-101)."""
    with torch.no_grad():
        if "label" not in batch:
            batch["label"] = 0
        spec, monitor, handler = method.cfg_spec, method.intensity_ma, method.prototypes
        image, pred_ema, out_static = _ema_and_static(method, batch)
        prior, conf, prior_conf = handler.prior_mix(
            [pred_ema["out"], out_static], [spec.EMA_LAMBDA, spec.STATIC_LAMBDA if out_static is not None else 0.0])
        monitor.add({"prior EMA": conf[0]})
        _record_ece(method, "ema", pred_ema["out"], batch["label"])
        if out_static is not None:
            monitor.add({"prior static": conf[1]})
            _record_ece(method, "static", out_static, batch["label"])
        if not _missing(spec.EXP_PR_STATIC) and spec.EXP_PR_STATIC:
            static_conf = monitor.exp("prior static")
        else:
            static_conf = monitor.avg("prior static")
        select = method.model_select
        select.evaluate(static_conf, monitor.dev_avg("prior static"))
        if select.current == select.dynamic and spec.DYNAMIC_LAMBDA > 0:
            _, pred_dyn = method.dynamic_model(image)
            prior, conf, prior_conf = handler.prior_mix([None, None, pred_dyn["out"]], [0.0, 0.0, spec.DYNAMIC_LAMBDA])
            monitor.add({"prior dynamic": conf[2]})
            _record_ece(method, "dynamic", pred_dyn["out"], batch["label"])
    return _finish(method, batch, pred_ema, prior, prior_conf)


def vswitch_prototype_predictions(method, batch):
    """vswitch_proDA.prototype_predictions (prototypes_vswitch.py:36-87)."""
    with torch.no_grad():
        spec, monitor, handler = method.cfg_spec, method.intensity_ma, method.prototypes
        image, pred_ema, out_static = _ema_and_static(method, batch)
        prior, conf, prior_conf = handler.prior_mix(
            [pred_ema["out"], out_static], [spec.EMA_LAMBDA, spec.STATIC_LAMBDA if out_static is not None else 0.0])
        monitor.add({"prior EMA": conf[0]})
        _record_ece(method, "ema", pred_ema["out"], batch.get("label", 0))
        if out_static is not None:
            monitor.add({"prior static": conf[1]})
            _record_ece(method, "static", out_static, batch.get("label", 0))
        select = method.model_select
        select.evaluate(monitor.dev_avg("prior static"))
        if select.current == select.dynamic and spec.DYNAMIC_LAMBDA > 0:
            _, pred_dyn = method.dynamic_model(image)
            prior, conf, prior_conf = handler.prior_mix([None, None, pred_dyn["out"]], [0.0, 0.0, spec.DYNAMIC_LAMBDA])
            monitor.add({"prior dynamic": conf[2]})
            _record_ece(method, "dynamic", pred_dyn["out"], batch.get("label", 0))
    return _finish(method, batch, pred_ema, prior, prior_conf)


def hswitch_prototype_predictions(method, batch):
    """hswitch_proDA.prototype_predictions (prototypes_hswitch.py:26-85)."""
    with torch.no_grad():
        spec, monitor, handler = method.cfg_spec, method.intensity_ma, method.prototypes
        image, pred_ema, out_static = _ema_and_static(method, batch)
        lam_s = spec.STATIC_LAMBDA if out_static is not None else 0.0
        # statistics first: the static share depends on them
        _, conf, _ = handler.prior_mix([pred_ema["out"], out_static], [spec.EMA_LAMBDA, lam_s], write_prior=False)
        monitor.add({"prior EMA": conf[0]})
        _record_ece(method, "ema", pred_ema["out"], batch.get("label", 0))
        if out_static is not None:
            monitor.add({"prior static": conf[1]})
            _record_ece(method, "static", out_static, batch.get("label", 0))
        share = static_share(monitor.avg("prior static"), spec.SOFT_TRANS,
                             0.0 if _missing(spec.SWITCH_PRIOR_THRESH) else spec.SWITCH_PRIOR_THRESH)
        monitor.add({"percentage_static": share})
        out_dyn = None
        if spec.DYNAMIC_LAMBDA > 0 and share < 1:
            _, pred_dyn = method.dynamic_model(image)
            out_dyn = pred_dyn["out"]
            _record_ece(method, "dynamic", out_dyn, batch.get("label", 0))
        # prior = (l_e p_e + l_s p_s) * share + ((1 - share) * l_d) * p_d, rounded in the reference's order (:36-68)
        prior, conf, prior_conf = handler.prior_mix(
            [pred_ema["out"], out_static, out_dyn],
            [spec.EMA_LAMBDA, lam_s, (1 - share) * spec.DYNAMIC_LAMBDA if out_dyn is not None else 0.0], scale01=share)
        if out_dyn is not None:
            monitor.add({"prior dynamic": conf[2]})
    return _finish(method, batch, pred_ema, prior, prior_conf)


def base_prototype_predictions(method, batch):
    """online_proDA.prototype_predictions (prototypes.py:208-273)."""
    with torch.no_grad():
        spec, monitor, handler = method.cfg_spec, method.intensity_ma, method.prototypes
        image, pred_ema, out_static = _ema_and_static(method, batch)
        lam_s = spec.STATIC_LAMBDA if out_static is not None else 0.0
        prior, conf, prior_conf = handler.prior_mix([pred_ema["out"], out_static], [spec.EMA_LAMBDA, lam_s])
        monitor.add({"prior EMA": conf[0]})
        _record_ece(method, "ema", pred_ema["out"], batch.get("label", 0))
        if out_static is not None:
            monitor.add({"prior static": conf[1]})
            _record_ece(method, "static", out_static, batch.get("label", 0))
        thresh = 0 if _missing(spec.SWITCH_PRIOR_THRESH) else spec.SWITCH_PRIOR_THRESH
        calculate_dyn, replace_dyn = True, False
        if thresh > 0 and below_f32(monitor.avg("prior static"), thresh):
            replace_dyn = True
        elif thresh > 0:
            calculate_dyn = False
        if spec.DYNAMIC_LAMBDA > 0 and calculate_dyn:
            _, pred_dyn = method.dynamic_model(image)
            _record_ece(method, "dynamic", pred_dyn["out"], batch.get("label", 0))
            if replace_dyn:
                prior, conf, prior_conf = handler.prior_mix([None, None, pred_dyn["out"]], [0.0, 0.0, spec.DYNAMIC_LAMBDA])
            else:
                prior, conf, prior_conf = handler.prior_mix(
                    [pred_ema["out"], out_static, pred_dyn["out"]], [spec.EMA_LAMBDA, lam_s, spec.DYNAMIC_LAMBDA])
            monitor.add({"prior dynamic": conf[2]})
    return _finish(method, batch, pred_ema, prior, prior_conf)


PROTOTYPE_PREDICTIONS = {
    # keys follow framework/handlers/adaptation_method_handler.py:1-8
    "PROTO_ONLINE": base_prototype_predictions,
    "PROTO_ONLINE_HSWITCH": hswitch_prototype_predictions,
    "PROTO_ONLINE_VSWITCH": vswitch_prototype_predictions,
    "PROTO_ONLINE_HYBRIDSWITCH": hybrid_prototype_predictions,
}
