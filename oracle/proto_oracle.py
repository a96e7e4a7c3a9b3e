"""CPU oracle for the OnDA prototype pseudo-labelling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``onda_b200/`` may import this module.
The only legitimate users are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` (there as the thing
that checks, or as the reported CPU baseline -- never as the product path).

What it is: a from-scratch restatement, on CPU in fp32 torch (the reference's
arithmetic is plain ATen), of the algorithm in
``framework/domain_adaptation/methods/prototype_handler.py`` and of the host-side
switch logic in ``framework/utils/monitoring.py`` and the ``prototypes_*switch``
method modules of theo2021/OnDA.  Every function cites the reference lines it
follows.  It deliberately keeps the reference's evaluation order (class loop
with full N x D temporaries, divide-then-norm, softmax -> multiply by prior ->
renormalise) so that (1) rounding matches the reference as closely as a
re-implementation can and (2) timing it gives a representative CPU baseline.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the *real*
reference from /root/reference (possible only in the authoring container) and
stores its outputs for seeded inputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against every one of
those fixtures (the reference itself ships no tests or golden vectors for this
path -- SURVEY.md section 4).

``dtype=torch.float64`` can be passed to most functions to obtain a
high-precision "truth" used by the tests to bound *both* implementations'
rounding error.
"""
from __future__ import annotations

import math
import pickle
from statistics import median as _median

import numpy as np
import torch

IGNORE_LABEL = 255


# --------------------------------------------------------------------------
# layout helpers
# --------------------------------------------------------------------------
def to_rows(t: torch.Tensor) -> torch.Tensor:
    """NCHW -> (N, ch) pixel-major rows, n = (b*h + y)*w + x; 2-D passes through.

    Reference: prototype_handler.transform, prototype_handler.py:105-109.
    """
    if t.dim() == 2:
        return t
    ch = t.shape[1]
    return t.permute(0, 2, 3, 1).reshape(-1, ch)


# --------------------------------------------------------------------------
# prototype statistics
# --------------------------------------------------------------------------
def pooled_std(protos: torch.Tensor, sq_mean: torch.Tensor, counter: torch.Tensor) -> torch.Tensor:
    """Count-weighted pooled per-channel standard deviation, shape (D,).

    Reference: prototype_handler.global_var, prototype_handler.py:53-60.  The
    evaluation order (multiply by c_k, divide by the scalar sum, then add over
    classes) is the reference's.
    """
    total = counter.sum()
    pooled_sq = (sq_mean.T * counter / total).T.sum(dim=0)
    pooled_mean = (protos.T * counter / total).T.sum(dim=0)
    return torch.sqrt(pooled_sq - pooled_mean ** 2)


def class_std(protos: torch.Tensor, sq_mean: torch.Tensor) -> torch.Tensor:
    """Per-class per-channel std.  Reference: prototype_var, prototype_handler.py:49-51."""
    return torch.sqrt(sq_mean - protos ** 2)


# --------------------------------------------------------------------------
# distances
# --------------------------------------------------------------------------
def raw_distance(feat: torch.Tensor, protos: torch.Tensor, sigma: torch.Tensor | None = None) -> torch.Tensor:
    """Un-shifted distance matrix (N, C): ||x_n - P_k||_2, or ||(x_n - P_k)/sigma||_2.

    Reference: the class loops at prototype_handler.py:117-120 (mahalanobis) and
    :132-135 (euclidean).  Kept as a loop over classes with an N x D temporary
    per class, like the reference.  The result buffer is float32 whatever the
    input dtype in the reference (torch.ones(...), :113/:129); here it follows
    ``feat.dtype`` so a float64 truth can be produced.
    """
    rows = to_rows(feat)
    n, c = rows.shape[0], protos.shape[0]
    out = torch.ones(n, c, dtype=rows.dtype)
    for k in range(c):
        delta = rows - protos[k]
        if sigma is not None:
            delta = delta / sigma
        out[:, k] = torch.norm(delta, 2, dim=1)
    return out


def shift_by_row_min(dist: torch.Tensor) -> torch.Tensor:
    """d - min_k d per row.  Reference: prototype_handler.py:124-125 and :137-138."""
    return (dist.T - dist.min(dim=1)[0]).T


def distance(feat, protos, sq_mean=None, counter=None, metric="euclidean"):
    """Public ``distance`` / ``mahalanobis_distance`` value (row-min shifted).

    Reference: prototype_handler.py:111-138.
    """
    sigma = pooled_std(protos, sq_mean, counter) if metric == "mahalanobis" else None
    return shift_by_row_min(raw_distance(feat, protos, sigma))


# --------------------------------------------------------------------------
# rectification and labels
# --------------------------------------------------------------------------
def rectify(shifted: torch.Tensor, prior_rows: torch.Tensor, tau: float):
    """(q, r): q = softmax_k(-d'/tau), r = q*prior / sum_k q*prior.

    Reference: prototype_handler.pseudo_labels, prototype_handler.py:147 and
    :159-160 (multiply in place, then renormalise by the row sum).
    """
    q = (-shifted / tau).softmax(dim=1)
    r = q * prior_rows
    r = r / r.sum(dim=1, keepdim=True)
    return q, r


def hard_labels(r: torch.Tensor, thresh: float) -> torch.Tensor:
    """(N,1) int64 labels, 255 where the winning probability is below ``thresh``.

    Reference: prototype_handler.py:163-166.  torch.max returns the first maximal
    index and, for rows containing NaN, the first NaN (so NaN rows keep label 0:
    ``nan < thresh`` is False).
    """
    m, labels = r.max(dim=1, keepdim=True)
    labels = labels.clone()
    labels[m < thresh] = IGNORE_LABEL
    return labels


# --------------------------------------------------------------------------
# class sums and prototype updates
# --------------------------------------------------------------------------
def first_argmax_onehot(rows: torch.Tensor) -> torch.Tensor:
    """One-hot (float32) of the first maximal column per row.

    Reference: prototype_handler.onehot, prototype_handler.py:83-86.
    """
    hot = torch.zeros_like(rows).float()
    return hot.scatter(1, rows.argmax(dim=1, keepdim=True), 1)


def class_sums(feat: torch.Tensor, out: torch.Tensor):
    """(sum over pixels of class k of x_n  [C, D],  pixel count per class [C]).

    Reference: prototype_handler.get_proto_array, prototype_handler.py:76-81
    (one-hot transposed times the feature rows).
    """
    rows = to_rows(feat)
    hot = first_argmax_onehot(to_rows(out))
    return hot.T.to(rows.dtype) @ rows, hot.sum(dim=0)


def _ones_where_empty(x: torch.Tensor) -> torch.Tensor:
    """Reference: the ``mask`` lambda, prototype_handler.py:21."""
    return torch.where(x > 0, x, torch.ones_like(x))


def ema_update(protos, sq_mean, feat, out, ma_lambda: float):
    """Returns the new (prototypes, squared_mean) after one moving-average step.

    Reference: prototype_handler.ma, prototype_handler.py:88-99.  Classes with no
    pixel keep their value (rho = lambda**0 = 1); ``counter`` is untouched.
    """
    s1, cnt = class_sums(feat, out)
    s2, _ = class_sums(feat ** 2, out)
    rho = ma_lambda ** (cnt > 0).float()
    safe = _ones_where_empty(cnt)
    new_p = (protos.T * rho).T + ((1 - rho) * (s1.T / safe)).T
    new_s = (sq_mean.T * rho).T + ((1 - rho) * (s2.T / safe)).T
    return new_p, new_s


def append_update(protos, sq_mean, counter, feat, out):
    """Cumulative-mean update used to initialise prototypes.

    Reference: prototype_handler.append, prototype_handler.py:62-74.  State may
    be the int 0 (fresh handler), in which case zeros are allocated.
    Returns the new (prototypes, squared_mean, counter).
    """
    s1, cnt = class_sums(feat, out)
    s2, _ = class_sums(feat ** 2, out)
    counter = counter + cnt
    safe = _ones_where_empty(counter)
    if isinstance(protos, int):
        protos = torch.zeros_like(s1, dtype=torch.float)
        sq_mean = torch.zeros_like(s1, dtype=torch.float)
    d1 = s1 - (protos.T * cnt).T
    d2 = s2 - (sq_mean.T * cnt).T
    return protos + (d1.T / safe).T, sq_mean + (d2.T / safe).T, counter


# --------------------------------------------------------------------------
# a handler with the reference's duck type (state + methods)
# --------------------------------------------------------------------------
class OracleHandler:
    """Stateful restatement with the reference class's public surface.

    Reference: class prototype_handler, prototype_handler.py:8-166.
    """

    def __init__(self, ma_lambda=0.9999, tau=1, thresh=0, distance_metric="euclidean",
                 confidence_regularization_threshold=1):
        if distance_metric not in ("euclidean", "mahalanobis"):
            raise ValueError("unexpected value for attribute distance_metric")  # :29
        self.prototypes = 0
        self.squared_mean = 0
        self.counter = 0
        self.ma_lambda = ma_lambda
        self.tau = tau
        self.thresh = thresh
        self.metric = distance_metric
        # an addict-missing key arrives as an empty dict and means "1" (:30-35)
        self.confidence_regularization_threshold = (
            1 if isinstance(confidence_regularization_threshold, dict)
            else confidence_regularization_threshold)

    # -- persistence (:37-47) ------------------------------------------------
    def save(self, loc="prototypes.pickle"):
        with open(loc, "wb") as f:
            pickle.dump((self.prototypes, self.squared_mean, self.counter), f)

    def load(self, loc="prototypes.pickle"):
        import os
        if not os.path.exists(loc):
            return False
        with open(loc, "rb") as f:
            self.prototypes, self.squared_mean, self.counter = pickle.load(f)
        return True

    # -- math ---------------------------------------------------------------
    def global_var(self):
        return pooled_std(self.prototypes, self.squared_mean, self.counter)

    def prototype_var(self):
        return class_std(self.prototypes, self.squared_mean)

    def distance_measure(self, feat):
        return distance(feat, self.prototypes, self.squared_mean, self.counter, self.metric)

    def get_proto_array(self, feat, out):
        return class_sums(feat, out)

    def ma(self, feat, out):
        self.prototypes, self.squared_mean = ema_update(
            self.prototypes, self.squared_mean, feat, out, self.ma_lambda)

    def append(self, feat, out):
        self.prototypes, self.squared_mean, self.counter = append_update(
            self.prototypes, self.squared_mean, self.counter, feat, out)

    def pseudo_labels(self, feat, prior, soft=False, confidence_monitor=None):
        """Reference: prototype_handler.py:140-166 (including the Monitor side effects)."""
        shifted = self.distance_measure(feat)
        q, r = rectify(shifted, to_rows(prior), self.tau)
        if confidence_monitor is not None and not confidence_monitor.freeze:
            confidence_monitor.add({"prototypes": q.max(dim=1)[0].mean()})
            if confidence_monitor.avg("prototypes") > self.confidence_regularization_threshold:
                self.tau += 0.001
                confidence_monitor.add({"tau": self.tau})
        if soft:
            return r
        return hard_labels(r, self.thresh)


def fused_step(handler: OracleHandler, feat, prior, out, monitor=None):
    """One unit of bench work: hard labels + soft predictions + EMA update.

    Reference call sites: prototypes_hybrid_switch.py:89-93 (the two
    pseudo_labels calls) followed by prototypes.py:292-294 (ma).
    """
    labels = handler.pseudo_labels(feat, prior, confidence_monitor=monitor)
    soft = handler.pseudo_labels(feat, prior, soft=True)
    handler.ma(feat, out)
    return labels, soft


# --------------------------------------------------------------------------
# switch statistics and selectors (host-side logic)
# --------------------------------------------------------------------------
def mean_max_softmax(logits: torch.Tensor) -> float:
    """conf(z) = mean over pixels of max_k softmax(z).  Reference:
    prototypes_hybrid_switch.py:53-54, :60-63, :79-82."""
    return logits.softmax(dim=1).max(dim=1)[0].mean().item()


def mean_max(prob: torch.Tensor) -> float:
    """mean over pixels of max_k prob (no renormalisation).  Reference:
    prototypes_hybrid_switch.py:88 and :94-96."""
    return prob.max(dim=1)[0].mean().item()


def normalised_entropy(prob: torch.Tensor) -> torch.Tensor:
    """-p*log2(p+1e-30)/log2(C) per element.  Reference: prob_2_entropy,
    framework/utils/func.py:71-74."""
    c = prob.shape[1]
    return -torch.mul(prob, torch.log2(prob + 1e-30)) / np.log2(c)


class OracleMonitor:
    """Sliding-window statistics.  Reference: Monitor, framework/utils/monitoring.py:7-96."""

    def __init__(self, limit=None, exp_const=0.01, dev_func="hamming"):
        self.window = {}
        self.ema = {}
        self.limit = limit
        self.exp_const = exp_const
        self.freeze = False
        self.weights = np.hamming(limit - 1)          # :24
        self.weights_sum = np.sum(self.weights)
        self.dev_func = dev_func

    def eval(self):
        self.freeze = True

    def train(self):
        self.freeze = False

    def _level(self, vals):
        if self.dev_func == "median":
            return _median(vals)
        if self.dev_func == "mean":
            return np.mean(np.array(vals))
        return np.sum(self.weights * np.array(vals)) / self.weights_sum  # :31-33

    def add(self, values, reset=False):
        if self.freeze:                                # :42-43
            return 0
        for key, val in values.items():
            if key not in self.window or reset:
                self.window[key] = [val]
                self.ema[key] = val
            else:
                self.window[key].append(val)
                if self.limit is not None and len(self.window[key]) > self.limit:
                    self.window[key].pop(0)
                self.ema[key] = (1 - self.exp_const) * self.ema[key] + self.exp_const * val

    def dev_avg(self, item):
        """Weighted first difference; 0 until the window is full (:64-73)."""
        if item not in self.window:
            return 0
        w = self.window[item]
        if len(w) < self.limit:
            return 0
        return self._level(w[1:]) - self._level(w[:-1])

    def exp(self, item):
        return self.ema.get(item, 1)                   # :75-81

    def avg(self, item):
        return _median(self.window[item]) if item in self.window else 1   # :83-89


class OracleHybridSelect:
    """Reference: model_select, prototypes_hybrid_switch.py:5-34."""
    static, dynamic = 0, 1

    def __init__(self, start=0, gray_area=(0.84, 0.88), dev_threshold=0.0002):
        self.current = start
        self.current_dev = start
        self.freeze = False
        self.gray_area = gray_area
        self.dev_threshold = dev_threshold

    def evaluate(self, confidence, dev_value):
        if self.freeze:
            return
        if dev_value > self.dev_threshold:
            self.current_dev = self.static
        elif dev_value < -self.dev_threshold:
            self.current_dev = self.dynamic
        if confidence < self.gray_area[0]:
            self.current = self.dynamic
        elif confidence > self.gray_area[1]:
            self.current = self.static
        else:
            self.current = self.current_dev


class OracleDevSelect:
    """Reference: model_select, prototypes_vswitch.py:5-25."""
    static, dynamic = 0, 1

    def __init__(self, start=0, threshold_c=0.00028):
        self.current = start
        self.freeze = False
        self.threshold = threshold_c

    def evaluate(self, dev_value):
        if self.freeze:
            return
        if dev_value > self.threshold:
            self.current = self.static
        elif dev_value < -self.threshold:
            self.current = self.dynamic


def hswitch_percentage(median_static: float, soft_trans: bool, switch_thresh: float = 0.0):
    """Reference: prototypes_hswitch.py:45-55."""
    # the reference evaluates this on a 0-dim float32 tensor (Monitor.avg of tensors): float32 arithmetic
    v = torch.tensor(float(median_static), dtype=torch.float32)
    if soft_trans:
        return float(max(min(v * (25.0 / 3) - (41.0 / 6), 1), 0))
    return int(v > switch_thresh)


def hybrid_prior(logits_ema, logits_static, logits_dynamic, monitor: OracleMonitor,
                 select: OracleHybridSelect, ema_lambda, static_lambda, dynamic_lambda,
                 exp_pr_static=False):
    """The prior mix and Monitor traffic of hybrid_proDA.prototype_predictions.

    Reference: prototypes_hybrid_switch.py:52-88.  ``logits_dynamic`` is a
    callable returning the dynamic model's logits (only evaluated when the
    selector says dynamic, like the reference's extra network forward).
    """
    p_ema = logits_ema.softmax(dim=1)
    monitor.add({"prior EMA": p_ema.max(dim=1)[0].mean()})
    prior = ema_lambda * p_ema
    if static_lambda > 0:
        p_static = logits_static.softmax(dim=1)
        monitor.add({"prior static": p_static.max(dim=1)[0].mean().item()})
        prior = prior + static_lambda * p_static
    conf = monitor.exp("prior static") if exp_pr_static else monitor.avg("prior static")
    select.evaluate(conf, monitor.dev_avg("prior static"))
    if select.current == select.dynamic and dynamic_lambda > 0:
        p_dyn = logits_dynamic().softmax(dim=1)
        monitor.add({"prior dynamic": p_dyn.max(dim=1)[0].mean()})
        prior = dynamic_lambda * p_dyn
    monitor.add({"prior": prior.max(dim=1)[0].mean()})
    return prior


# --------------------------------------------------------------------------
# seeded synthetic inputs shared by tests and bench
# --------------------------------------------------------------------------
def step_log_stats(pseudolabels: torch.Tensor, student_out: torch.Tensor, protos: torch.Tensor) -> dict:
    """The per-step log reductions, as written in framework/domain_adaptation/methods/prototypes.py:341-352."""
    batch_size, _, w, h = student_out.shape
    return {
        "pseudolabel_pixel_num": float(((pseudolabels >= 0) * (pseudolabels != 255)).float().sum()),
        "output & prototype agreement": float((pseudolabels.reshape(batch_size, w, h) == student_out.argmax(axis=1)).float().mean()),
        "mean_prototype_intensity_values": float((protos ** 2).mean()),
    }


def target_losses(out: torch.Tensor, labels: torch.Tensor, alpha: float, beta: float, reg_weight: float, regularizer: str):
    """CE + RCE + regulariser on hard pseudo-labels, as online_proDA.pseudolabel_loss combines them
    (prototypes.py:313-328).  ``out`` (B, C, h, w) logits, ``labels`` (B, h, w) int64 with 255 = ignore.
    Restates cross_entropy_2d (framework/utils/loss.py:30-45), rce (loss.py:88-112) and regular_loss
    (prototypes.py:29-39); differentiable through torch autograd."""
    b, c, h, w = out.shape
    mask = (labels >= 0) & (labels != 255)                                    # loss.py:36
    rows = out.permute(0, 2, 3, 1)[mask]                                      # :39-41: the selected pixels' logit rows
    ce = torch.nn.functional.cross_entropy(rows, labels[mask].long())         # :44, mean over the selection
    p = out.softmax(dim=1)                                                    # loss.py:89
    clone = labels.long().clone()
    clone[clone == 255] = c                                                   # :101
    onehot = torch.nn.functional.one_hot(clone, c + 1).float().permute(0, 3, 1, 2)[:, :-1]
    onehot = torch.clamp(onehot, min=1e-4, max=1.0)                           # :104-106
    m = (labels != 255).float()
    rce_loss = -((p * torch.log(onehot)).sum(dim=1) * m).sum() / (m.sum() + 1e-6)   # :107-109
    logp = torch.nn.functional.log_softmax(out, dim=1)                        # prototypes.py:31
    if regularizer == "MRENT":
        reg = (logp.exp() * logp).sum() / (b * h * w)                         # :33-35
    elif regularizer == "MRKLD":
        reg = -logp.sum() / (b * c * h * w)                                   # :36-39
    else:
        reg = out.sum() * 0
    total = alpha * ce + beta * rce_loss + reg_weight * reg
    agree = (labels == out.argmax(dim=1)).float().mean()                      # prototypes.py:346-347
    return {"ce": ce, "rce": rce_loss, "reg": reg, "total": total, "agreement": agree, "n_valid": mask.sum()}


def update_ema(params_q, params_k, buffers_q, buffers_k, ema_update: float):
    """The model-weight EMA, as written in framework/domain_adaptation/methods/prototypes.py:407-416; returns the new
    (parameter list, buffer list) of the EMA model."""
    new_params = [k.clone() * ema_update + q.clone() * (1.0 - ema_update) for q, k in zip(params_q, params_k)]
    new_buffers = [q.clone() for q, _ in zip(buffers_q, buffers_k)]
    return new_params, new_buffers


def eval_confusion(pred: torch.Tensor, labels: torch.Tensor, n: int, size):
    """Prediction map and confusion counts of one batch, as written in
    framework/domain_adaptation/methods/adaptation_model.py:94-98,143-160 and framework/utils/func.py:77-79.
    Returns (softmaxed upsampled prediction (B, n, H, W), per-pixel argmax (B, H, W) int64, hist (n, n) int64)."""
    import numpy as np
    interp = torch.nn.Upsample(size=size, mode="bilinear", align_corners=True)
    prob = interp(pred).softmax(axis=1)
    hist = np.zeros((n, n), dtype=np.int64)
    preds = []
    for item_pred, label in zip(prob, labels):
        a = label.numpy().flatten()
        b = item_pred.permute(1, 2, 0).argmax(dim=2).cpu().numpy().flatten()
        k = (a >= 0) & (a < n)
        hist += np.bincount(n * a[k].astype(int) + b[k], minlength=n ** 2).reshape(n, n)
        preds.append(torch.from_numpy(b.reshape(label.shape)))
    return prob, torch.stack(preds), hist


def synth_case(seed: int, b: int, d: int, h: int, w: int, c: int = 19, protos=None,
               counter=None, sharp: float = 4.0):
    """Cityscapes-shaped synthetic inputs: blocky label map, class-separable
    features with a Dropout2d pattern, logits peaked on the true class.

    Returns a dict of CPU float32 tensors: feat, out, prior, protos, sq_mean,
    counter.  Deterministic for a given torch build (CPU generator).
    """
    g = torch.Generator().manual_seed(seed)
    if protos is None:
        protos = torch.randn(c, d, generator=g) * 2.5
    if counter is None:
        counter = torch.floor(torch.rand(c, generator=g) * 6.9e4 + 1e3)
    sq_mean = protos ** 2 + (torch.rand(c, d, generator=g) * 1.5 + 0.5)
    by, bx = (h + 7) // 8, (w + 7) // 8
    blocks = torch.randint(0, c, (b, by, bx), generator=g)
    labels = blocks.repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :h, :w]
    feat = torch.randn(b, d, h, w, generator=g) * 2.5
    feat = feat + 0.5 * protos[labels].permute(0, 3, 1, 2)
    keep = (torch.rand(b, d, 1, 1, generator=g) >= 0.1).float() / 0.9
    feat = (feat * keep).contiguous()
    hot = torch.nn.functional.one_hot(labels, c).permute(0, 3, 1, 2).float()
    out = (torch.randn(b, c, h, w, generator=g) * 3 + sharp * hot).contiguous()
    prior_logits = torch.randn(b, c, h, w, generator=g) * 3 + sharp * hot
    prior = prior_logits.softmax(dim=1).contiguous()
    return {"feat": feat, "out": out, "prior": prior, "prior_logits": prior_logits.contiguous(),
            "protos": protos.contiguous(), "sq_mean": sq_mean.contiguous(),
            "counter": counter.contiguous()}
