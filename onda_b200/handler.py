"""``prototype_handler``: drop-in for OnDA's class of the same name, running on sm_100a kernels.

Mirrors the public surface of
``framework/domain_adaptation/methods/prototype_handler.py:8-166`` (constructor arguments,
mutable attributes, method names, argument meaning, return layouts, error behaviour) so that

    from onda_b200 import prototype_handler

can replace the reference import in ``methods/prototypes.py:16`` unchanged.  All tensor math
runs in ``libonda_b200.so`` through the C ABI of ``include/onda_b200.h``; PyTorch is used only
to own device memory and streams.  There is no CPU path: tensors must live on a CUDA device.

Beyond the reference surface it adds a fused entry, ``pseudo_labels_fused``, which returns the
hard labels and the soft predictions from ONE pass over ``feat`` (the reference needs two
``pseudo_labels`` calls) and, when the EMA logits are passed too, also stages the class sums so
that the following ``ma(feat, out)`` is just the tiny blend kernel.
"""
from __future__ import annotations

import os
import pickle
import weakref

import torch

from . import _native as nat
from . import sharding


def _stream_ptr(device):
    return nat.C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _on(device):
    """The C ABI launches on the CURRENT device: make the tensors' device current for the duration of the calls (the
    handler may live on cuda:1, or be used from a process that never called torch.cuda.set_device)."""
    return torch.cuda.device(device)


class _CpuUnpickler(pickle.Unpickler):
    """pickle.load for tensors that were saved from a CUDA device, on a machine without one."""

    def find_class(self, module, name):
        if module == "torch.storage" and name == "_load_from_bytes":
            import io
            return lambda b: torch.load(io.BytesIO(b), map_location="cpu", weights_only=False)
        return super().find_class(module, name)


class prototype_handler:
    """Class prototypes with distance-based pseudo-labelling and EMA / cumulative updates.

    Reference: prototype_handler.py:9-35 for the constructor.  Extra keyword arguments
    (not in the reference): ``impl`` selects the kernel ("auto" | "simt" | "tcgen05"),
    ``process_group`` makes ``ma``/``append`` sum their class statistics over the ranks of a
    ``torch.distributed`` group so every rank keeps identical prototypes (``allreduce="nccl"`` uses
    ``dist.all_reduce``, ``"oneshot"`` the library's own one-launch NVLink peer-memory kernel), ``fuse_hard_soft``
    lets a ``pseudo_labels(soft=True)`` call that directly follows the hard call on the very
    same tensors reuse that launch.
    """

    def __init__(self, ma_lambda=0.9999, tau=1, thresh=0, distance_metric="euclidean",
                 confidence_regularization_threshold=1, impl="auto", process_group=None, fuse_hard_soft=True,
                 allreduce="nccl", tile_schedule="dynamic"):
        self.prototypes = 0  # classes x features once appended / loaded (prototype_handler.py:17)
        self.squared_mean = 0
        self.counter = 0
        self.ma_lambda = ma_lambda
        self.mask = lambda x: torch.where(x > 0, x, torch.ones_like(x))
        self.tau = tau
        self.thresh = thresh
        if distance_metric == "euclidean":
            self.distance_measure = self.distance
        elif distance_metric == "mahalanobis":
            self.distance_measure = self.mahalanobis_distance
        else:
            raise ValueError("unexpected value for attribute distance_metric")
        self.distance_metric = distance_metric
        if isinstance(confidence_regularization_threshold, dict):  # addict-missing key
            self.confidence_regularization_threshold = 1
        else:
            self.confidence_regularization_threshold = confidence_regularization_threshold
        if impl not in nat.IMPL:
            raise ValueError(f"unknown impl {impl!r}")
        self.impl = impl
        self.process_group = process_group
        if allreduce not in ("nccl", "oneshot"):
            raise ValueError(f"unknown allreduce {allreduce!r}")
        self.allreduce = allreduce     # "nccl": torch.distributed all_reduce; "oneshot": own NVLink peer-memory kernel
        self._symm = None              # (tensor, handle, n, slot_floats) of the one-shot all-reduce
        self._ar_calls = 0
        if tile_schedule not in ("dynamic", "fixed"):
            raise ValueError(f"unknown tile_schedule {tile_schedule!r}")
        # tcgen05 kernel: "dynamic" draws the tiles of an SM from a device counter (fastest; which SM accumulates which
        # tile, and so the last bits of the class sums, vary from run to run); "fixed" is the round-robin schedule
        # whose class sums are bit-reproducible.
        self.tile_schedule = tile_schedule
        self.fuse_hard_soft = fuse_hard_soft
        self._stats_src = None     # (sums, C, D) of the last fused pass
        self._local_stats = None   # same, never replaced by the all-reduced buffer
        self._stats_cache = None
        self._lib = nat.load()
        self._epoch = 0            # bumped whenever our kernels rewrite the state in place
        self._table = {}           # metric -> (key, tensor)
        self._bufs = {}            # (kind, shape key) -> tensor
        self._memo = None          # last fused launch, for the hard -> soft reuse
        self._pending = None       # class sums staged by pseudo_labels_fused for ma()
        self._deferred_monitor = None

    # ------------------------------------------------------------------ persistence
    def save(self, loc="prototypes.pickle"):
        """Pickle (prototypes, squared_mean, counter) -- prototype_handler.py:37-38."""
        with open(loc, "wb") as f:
            pickle.dump((self.prototypes, self.squared_mean, self.counter), f)

    def load(self, loc="prototypes.pickle"):
        """Load the 3-tuple written by ``save``; returns False if the file is absent (:40-47).

        Also reads the legacy 2-tuple ``(prototypes, counter)`` of the ``prototypes.pickle`` shipped with the
        reference (which the reference's own ``load`` cannot unpack): ``squared_mean`` then stays uninitialised (only
        the Euclidean metric works) until the first ``append``, which starts the second moments from zero.  Tensors pickled on a CUDA device are
        mapped to the CPU when no CUDA device is present.
        """
        if os.path.exists(loc):
            with open(loc, "rb") as f:
                try:
                    state = pickle.load(f)
                except RuntimeError:          # CUDA storages without a CUDA device
                    f.seek(0)
                    state = _CpuUnpickler(f).load()
            if len(state) == 2:
                self.prototypes, self.counter = state
                self.squared_mean = 0
            else:
                self.prototypes, self.squared_mean, self.counter = state
            print("Prototypes loaded!")
            return True
        return False

    # ------------------------------------------------------------------ helpers
    def transform(self, matrix):
        """NCHW -> (N, channels) rows; 2-D passes through (:105-109).  A view operation."""
        if matrix.dim() == 2:
            return matrix
        _, channels, _, _ = matrix.size()
        return matrix.permute(0, 2, 3, 1).reshape(-1, channels)

    def onehot(self, matrix):
        """One-hot of the first maximal column per row (:83-86)."""
        hot = torch.zeros_like(matrix).float()
        return hot.scatter(1, matrix.argmax(axis=1, keepdim=True), 1)

    def _buf(self, kind, shape, dtype, device, zero=False):
        key = (kind, tuple(shape), dtype, str(device))
        t = self._bufs.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
            self._bufs[key] = t
        return t

    @staticmethod
    def _require_cuda(t, name):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if not t.is_cuda:
            raise RuntimeError(f"onda_b200 runs on CUDA devices only: {name} is on {t.device} (there is no CPU path)")

    def _nchw(self, t, name, dtype_ok=(torch.float32,)):
        """Dense fp32 (B, ch, HW) view of a 4-D NCHW or 2-D (M, ch) input."""
        self._require_cuda(t, name)
        if t.dim() == 2:      # (M, ch) rows -> one image of M pixels (prototype_handler.py:106-107)
            t = t.t().unsqueeze(0)
        elif t.dim() == 4:
            t = t.reshape(t.shape[0], t.shape[1], -1)
        else:
            raise ValueError(f"{name} must be (B, ch, H, W) or (M, ch), got {tuple(t.shape)}")
        if t.dtype != torch.float32:
            t = t.float()
        return t.contiguous()

    def _state(self, device, need_stats):
        """State tensors as dense fp32 on ``device`` (moves / converts a freshly loaded pickle once)."""
        if isinstance(self.prototypes, int):
            raise AttributeError("prototypes are not initialised: call append() or load() first")
        for name in ("prototypes", "squared_mean", "counter"):
            t = getattr(self, name)
            if not isinstance(t, torch.Tensor):
                if name == "prototypes" or need_stats:
                    raise AttributeError(f"{name} is not initialised")
                continue
            if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                setattr(self, name, t.to(device=device, dtype=torch.float32).contiguous())
        return self.prototypes, self.squared_mean, self.counter

    def _table_key(self, P, S, cnt, need_stats, device):
        C, D = P.shape
        return (self._epoch, P.data_ptr(), P._version,
                S.data_ptr() if need_stats else 0, S._version if need_stats else 0,
                cnt.data_ptr() if need_stats else 0, cnt._version if need_stats else 0, C, D, str(device))

    def _distance_table(self, metric, device):
        need_stats = metric == "mahalanobis"
        P, S, cnt = self._state(device, need_stats)
        C, D = P.shape
        if C > nat.MAX_CLASSES:
            raise ValueError(f"{C} classes unsupported (max {nat.MAX_CLASSES})")
        key = self._table_key(P, S, cnt, need_stats, device)
        hit = self._table.get(metric)
        if hit is not None and hit[0] == key:
            return hit[1]
        table = self._buf(("table", metric), (self._lib.onda_table_floats(C, D),), torch.float32, device, zero=True)
        with _on(device):
            nat.check(self._lib.onda_build_distance_table(
                nat.ptr(P), nat.ptr(S) if need_stats else None, nat.ptr(cnt) if need_stats else None,
                C, D, nat.METRIC[metric], nat.ptr(table), _stream_ptr(device)), "onda_build_distance_table")
        self._table[metric] = (key, table)
        return table

    def _launch(self, feat3, prior3, logits3, metric, want_labels, want_soft, want_dist, C):
        """One fused pass.  Returns (labels, soft, dist, sums); unused outputs are None."""
        device = feat3.device
        B, D, HW = feat3.shape
        N = B * HW
        need_table = want_labels or want_soft or want_dist
        table = self._distance_table(metric, device) if need_table else None
        labels = torch.empty((N, 1), dtype=torch.int64, device=device) if want_labels else None
        soft = torch.empty((N, C), dtype=torch.float32, device=device) if want_soft else None
        dist = torch.empty((N, C), dtype=torch.float32, device=device) if want_dist else None
        n_sums = self._lib.onda_sums_floats(C, D)
        sums, guard = None, (None, None, 0)
        if self.process_group is not None and self.allreduce == "oneshot" and logits3 is not None:
            sums = self._symm_slot(n_sums, device)      # written straight into peer-visible memory: no staging copy
            if sums is not None:                        # ... once the peers have read what it held (their "done" words)
                guard = self._symm[4][self._ar_calls & 1][4]
        if sums is None:
            sums = torch.empty((n_sums,), dtype=torch.float32, device=device)
        impl = nat.IMPL[self.impl]
        with _on(device):
            wbytes = self._lib.onda_fused_workspace_bytes(B, D, HW, C, impl)
            work = self._buf("work", (wbytes,), torch.uint8, device)
            self._lib.onda_set_tile_schedule(1 if self.tile_schedule == "dynamic" else 0)
            nat.check(self._lib.onda_pseudolabel_fused_guarded(
                nat.ptr(feat3), nat.ptr(prior3), nat.ptr(logits3), nat.ptr(table), B, D, HW, C,
                float(self.tau), float(self.thresh), nat.ptr(labels), nat.ptr(soft), nat.ptr(dist), nat.ptr(sums),
                nat.ptr(work), wbytes, impl, guard[0], guard[1], guard[2], _stream_ptr(device)), "onda_pseudolabel_fused")
        return labels, soft, dist, sums

    def _num_classes(self, other=None):
        if isinstance(self.prototypes, torch.Tensor):
            return self.prototypes.shape[0]
        return other

    # ------------------------------------------------------------------ statistics of the state
    def prototype_var(self):
        """Per-class std sqrt(squared_mean - prototypes**2) (:49-51)."""
        self._require_cuda(self.prototypes, "prototypes")
        P, S, _ = self._state(self.prototypes.device, True)
        out = torch.empty_like(P)
        with _on(P.device):
            nat.check(self._lib.onda_prototype_std(nat.ptr(P), nat.ptr(S), P.shape[0], P.shape[1], nat.ptr(out),
                                                   _stream_ptr(P.device)), "onda_prototype_std")
        return out

    def global_var(self):
        """Count-weighted pooled per-channel std (:53-60)."""
        self._require_cuda(self.prototypes, "prototypes")
        device = self.prototypes.device
        table = self._distance_table("mahalanobis", device)
        C, D = self.prototypes.shape
        out = torch.empty((D,), dtype=torch.float32, device=device)
        with _on(device):
            nat.check(self._lib.onda_table_global_std(nat.ptr(table), C, D, nat.ptr(out), _stream_ptr(device)),
                      "onda_table_global_std")
        return out

    # ------------------------------------------------------------------ distances
    def _distance(self, feat, metric):
        feat3 = self._nchw(feat, "feat")
        C = self._num_classes()
        if feat3.shape[0] * feat3.shape[2] == 0:
            return torch.empty((0, C), dtype=torch.float32, device=feat.device)
        _, _, dist, _ = self._launch(feat3, None, None, metric, False, False, True, C)
        return dist

    def mahalanobis_distance(self, feat):
        """(N, C) pooled-variance Mahalanobis distances, shifted by the row minimum (:111-125)."""
        return self._distance(feat, "mahalanobis")

    def distance(self, feat):
        """(N, C) Euclidean distances, shifted by the row minimum (:127-138)."""
        return self._distance(feat, "euclidean")

    # ------------------------------------------------------------------ pseudo labels
    def _monitor_side_effects(self, confidence_monitor, stats):
        """The Monitor traffic of pseudo_labels (:148-156)."""
        if confidence_monitor is None or confidence_monitor.freeze:
            return
        confidence_monitor.add({"prototypes": stats["prototypes"]})
        if confidence_monitor.avg("prototypes") > self.confidence_regularization_threshold:
            self.tau += 0.001
            confidence_monitor.add({"tau": self.tau})

    def _stats_from(self, sums, C, D):
        tail = sums[2 * C * D + C:].tolist()   # one small D2H copy (the reference syncs here too)
        return sharding.stats_from_tail(tail)

    @property
    def last_stats(self):
        """Batch statistics of the most recent fused pass (means over all pixels; over all ranks
        once ``ma`` has all-reduced them).  Read lazily: the device-to-host copy happens here, so read it before the
        next-but-one fused pass -- the device buffer behind it (one of two alternating ones when the ranks exchange
        through peer memory) is reused then."""
        src = self._stats_src
        if src is not None and self._stats_cache is None:
            self._stats_cache = self._stats_from(*src)
        return self._stats_cache or {}

    def last_pixel_count(self):
        """Device tensor (1 float) with the number of non-ignored pseudo-labels of the most recent fused pass of THIS rank
        (the ONDA_STAT_PL_PIXELS slot of its statistics) -- what ``target_losses(..., n_valid=...)`` divides by, without
        a device-to-host copy.  None before the first pass."""
        src = self._local_stats
        if src is None:
            return None
        sums, C, D = src
        return sums[2 * C * D + C + nat.STAT_PL_PIXELS: 2 * C * D + C + nat.STAT_PL_PIXELS + 1]

    def pseudo_labels_fused(self, feat, prior, out=None, confidence_monitor=None, want_labels=True, want_soft=True):
        """Hard labels AND soft predictions from one pass over ``feat``.

        Equivalent to the reference's ``pseudo_labels(feat, prior, confidence_monitor=m)`` followed by
        ``pseudo_labels(feat, prior, soft=True)`` (prototypes_hybrid_switch.py:89-93).  If ``out``
        (the EMA logits) is given, the class sums / counts that ``ma(feat, out)`` needs are produced
        by the same pass and kept until that call (prototypes.py:292-294).  Returns
        ``(labels (N,1) int64, soft (N,C) float32)``; ``self.last_stats`` holds the batch means.

        With a ``process_group`` and ``out`` given, the Monitor side effects are applied by the
        following ``ma`` call, after the single all-reduce, so that every rank records the same
        (global) confidence; otherwise they are applied here like the reference does.
        """
        if prior is None:
            raise AttributeError("'NoneType' object has no attribute 'device'")  # prototype_handler.py:142
        feat3 = self._nchw(feat, "feat")
        prior3 = self._nchw(prior, "prior")
        logits3 = self._nchw(out, "out") if out is not None else None
        C = self._num_classes()
        B, D, HW = feat3.shape
        if prior3.shape != (B, C, HW):
            raise ValueError(f"prior shape {tuple(prior.shape)} does not match feat {tuple(feat.shape)} / {C} classes")
        if B * HW == 0:
            empty_l = torch.empty((0, 1), dtype=torch.int64, device=feat.device)
            return empty_l, torch.empty((0, C), dtype=torch.float32, device=feat.device)
        tau_used = self.tau
        labels, soft, _, sums = self._launch(feat3, prior3, logits3, self.distance_metric, want_labels, want_soft,
                                             False, C)
        self._stats_src, self._stats_cache = (sums, C, D), None
        self._local_stats = (sums, C, D) if want_labels else self._local_stats
        monitor_live = confidence_monitor is not None and not confidence_monitor.freeze
        defer = self.process_group is not None and logits3 is not None
        if logits3 is not None:
            self._pending = (weakref.ref(feat), feat._version, weakref.ref(out), out._version, self._epoch, sums,
                             confidence_monitor if (defer and monitor_live) else None)
        if monitor_live and not defer:
            if self.process_group is not None:
                import torch.distributed as dist
                dist.all_reduce(sums[2 * C * D + C:], op=dist.ReduceOp.SUM, group=self.process_group)
            self._monitor_side_effects(confidence_monitor, self.last_stats)
            if self.tau != tau_used and want_soft:
                # the confidence regulariser raised tau (prototype_handler.py:151-156): the reference's soft call that
                # follows the hard one already sees the new value, so the soft predictions are recomputed with it
                stats_keep = (self._stats_src, self._stats_cache)
                _, soft, _, _ = self._launch(feat3, prior3, None, self.distance_metric, False, True, False, C)
                self._stats_src, self._stats_cache = stats_keep
                tau_used = self.tau
        self._memo = (weakref.ref(feat), feat._version, weakref.ref(prior), prior._version, self._epoch,
                      tau_used, self.thresh, labels, soft, sums)
        return labels, soft

    def pseudo_labels(self, feat, prior=None, soft=False, confidence_monitor=None):
        """(N,1) int64 labels with 255 = ignore, or (N,C) float32 soft predictions (:140-166)."""
        if prior is None:
            raise AttributeError("'NoneType' object has no attribute 'device'")  # the reference dereferences it (:142)
        if soft and confidence_monitor is None and self._memo is not None:
            f, fv, p, pv, epoch, tau, thresh, _, soft_t, _ = self._memo
            if (soft_t is not None and f() is feat and p() is prior and fv == feat._version and pv == prior._version
                    and epoch == self._epoch and tau == self.tau and thresh == self.thresh):
                self._memo = None
                return soft_t
        if soft:
            if confidence_monitor is not None:
                _, s = self.pseudo_labels_fused(feat, prior, None, confidence_monitor, want_labels=False)
            else:
                _, s = self.pseudo_labels_fused(feat, prior, None, None, want_labels=False)
            self._memo = None
            return s
        labels, _ = self.pseudo_labels_fused(feat, prior, None, confidence_monitor, want_soft=self.fuse_hard_soft)
        return labels

    # ------------------------------------------------------------------ prior mix / switch statistics
    def prior_mix(self, logits, coefs, write_prior=True, scale01=1.0):
        """Softmax-mix of up to three logit maps and their batch confidences, in one pass.

        ``logits``: sequence of up to 3 tensors (B, C, h, w) or None; ``coefs``: their weights.
        Returns ``(prior, conf, prior_conf)`` where ``prior = (coefs[0]*softmax(logits[0], 1) + coefs[1]*softmax(logits[1],
        1)) * scale01 + coefs[2]*softmax(logits[2], 1)`` (every step rounded like the reference's tensor expression;
        ``scale01`` is the h-switch's ``percentage_static``)
        as a (B, C, h, w) tensor (None if ``write_prior`` is False), ``conf[i] = mean_n max_k
        softmax(logits[i])`` (None for missing inputs) and ``prior_conf = mean_n max_k prior``.
        Replaces the softmax/max/mean chains of prototypes_hybrid_switch.py:52-88 (and the
        h-switch / v-switch / base variants).  Confidences are Python floats, global over the
        ``process_group`` if one is set (so every rank takes the same switch decision).
        """
        logits = list(logits) + [None] * (3 - len(logits))
        coefs = list(coefs) + [0.0] * (3 - len(coefs))
        ref = next((t for t in logits if t is not None), None)
        if ref is None:
            raise ValueError("prior_mix needs at least one logits tensor")
        shape = tuple(ref.shape)
        dense = [self._nchw(t, "logits") if t is not None else None for t in logits]
        B, C, HW = next(t for t in dense if t is not None).shape
        for t in dense:
            if t is not None and t.shape != (B, C, HW):
                raise ValueError("prior_mix inputs must share one shape")
        device = ref.device
        prior = torch.empty((B, C, HW), dtype=torch.float32, device=device) if write_prior else None
        stats = torch.empty((nat.NUM_STATS,), dtype=torch.float32, device=device)
        with _on(device):
            wbytes = self._lib.onda_prior_workspace_bytes(B, C, HW)
        work = self._buf("prior_work", (wbytes,), torch.uint8, device, zero=True)
        with _on(device):
          nat.check(self._lib.onda_prior_mix_stats(
            nat.ptr(dense[0]), nat.ptr(dense[1]), nat.ptr(dense[2]), float(coefs[0]), float(coefs[1]), float(coefs[2]),
            float(scale01), B, C, HW, nat.ptr(prior), nat.ptr(stats), nat.ptr(work), wbytes, _stream_ptr(device)), "onda_prior_mix_stats")
        if self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.process_group)
        host = stats.tolist()
        n = host[4]
        inv = 1.0 / n if n > 0 else float("nan")
        conf = [host[i] * inv if dense[i] is not None else None for i in range(3)]
        if prior is not None:
            prior = prior.view(shape) if len(shape) == 4 else prior[0].t()
        return prior, conf, host[3] * inv

    def step_log_stats(self, pseudolabels, student_out):
        """The per-step log reductions of the method classes (prototypes.py:341-352) in one launch.

        ``pseudolabels``: (N, 1) / (N,) / (B, h, w) int64 with 255 = ignore; ``student_out``: (B, C, h, w) logits of
        the model being trained.  Returns ``{"pseudolabel_pixel_num", "output & prototype agreement",
        "mean_prototype_intensity_values"}`` as Python floats (the reference logs them, nothing reads them back).
        """
        self._require_cuda(pseudolabels, "pseudolabels")
        logits3 = self._nchw(student_out, "student_out")
        B, C, HW = logits3.shape
        if pseudolabels.dtype != torch.int64 or pseudolabels.numel() != B * HW:
            raise ValueError(f"pseudolabels must be int64 with {B * HW} entries, got {pseudolabels.dtype} {tuple(pseudolabels.shape)}")
        self._require_cuda(self.prototypes, "prototypes")
        device = logits3.device
        P, _, _ = self._state(device, False)
        labels = pseudolabels.contiguous()
        out4 = torch.empty((4,), dtype=torch.float32, device=device)
        with _on(device):
            wbytes = self._lib.onda_step_log_workspace_bytes()
            work = self._buf("log_work", (wbytes,), torch.uint8, device, zero=True)
            nat.check(self._lib.onda_step_log_stats(nat.ptr(labels), nat.ptr(logits3), nat.ptr(P), B, C, HW, P.shape[1],
                                                    nat.ptr(out4), nat.ptr(work), wbytes, _stream_ptr(device)),
                      "onda_step_log_stats")
        if self.process_group is not None:
            import torch.distributed as dist
            keep = out4[2].clone()                   # the prototypes are replicated, not sharded
            dist.all_reduce(out4, op=dist.ReduceOp.SUM, group=self.process_group)
            out4[2] = keep
        agree, valid, sq, n = out4.tolist()
        f32 = torch.tensor([agree, n, sq, float(P.numel())], dtype=torch.float32)
        return {"pseudolabel_pixel_num": valid,
                "output & prototype agreement": float(f32[0] / f32[1]),          # .float().mean() of exact 0/1 values
                "mean_prototype_intensity_values": float(f32[2] / f32[3])}

    # ------------------------------------------------------------------ class sums and updates
    def _class_sums(self, feat, out):
        pend = self._pending
        if pend is not None:
            f, fv, o, ov, epoch, sums, monitor = pend
            self._pending = None
            if f() is feat and o() is out and fv == feat._version and ov == out._version and epoch == self._epoch:
                self._deferred_monitor = monitor
                return sums, feat.shape[1], out.shape[1], feat.device
        feat3 = self._nchw(feat, "feat")
        logits3 = self._nchw(out, "out")
        B, D, HW = feat3.shape
        C = logits3.shape[1]
        if logits3.shape != (B, C, HW):
            raise ValueError(f"out shape {tuple(out.shape)} does not match feat {tuple(feat.shape)}")
        if C > nat.MAX_CLASSES:
            raise ValueError(f"{C} classes unsupported (max {nat.MAX_CLASSES})")
        if B * HW == 0:
            n = self._lib.onda_sums_floats(C, D)
            return torch.zeros((n,), dtype=torch.float32, device=feat.device), D, C, feat.device
        _, _, _, sums = self._launch(feat3, None, logits3, self.distance_metric, False, False, False, C)
        return sums, D, C, feat3.device

    def _allreduce(self, sums):
        if self.process_group is None:
            return sums
        if self.allreduce == "oneshot" and sums.is_cuda:
            return self._allreduce_oneshot(sums)
        sharding.allreduce_sums(sums, self.process_group)
        return sums

    def _symm_init(self, n, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        slot_floats = (n + 63) // 64 * 64
        buf = symm_mem.empty(2 * slot_floats + 64, dtype=torch.float32, device=device)
        hdl = symm_mem.rendezvous(buf, self.process_group)
        buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=self.process_group)
        world, rank = dist.get_world_size(self.process_group), dist.get_rank(self.process_group)
        ptr_t = nat.C.c_void_p * world
        peer_ptrs = [int(p) for p in hdl.buffer_ptrs]
        # per slot: (view of this rank's input, peers' input pointers, peers' flag pointers) -- built once, the step
        # itself must not spend host time on it
        # epochs of the exchange fused into ma(): one device counter per slot, bumped by the kernel itself, so the
        # call has no per-step host argument and can be replayed from a CUDA graph
        self._epoch_ctr = torch.ones((2,), dtype=torch.int32, device=device)
        slots = []
        for slot in (0, 1):
            view = buf[slot * slot_floats: slot * slot_floats + n]
            bufs = ptr_t(*[p + 4 * slot * slot_floats for p in peer_ptrs])
            flags = ptr_t(*[p + 4 * (2 * slot_floats + 32 * slot) for p in peer_ptrs])
            flags_fused = ptr_t(*[p + 4 * (2 * slot_floats + 32 * slot + 8) for p in peer_ptrs])   # own words: own epochs
            done = ptr_t(*[p + 4 * (2 * slot_floats + 32 * slot + 16) for p in peer_ptrs])         # "has read my slot" words
            # the writer's guard: this rank's own done words, the slot's epoch counter, the number of ranks
            guard = (nat.C.c_void_p(peer_ptrs[rank] + 4 * (2 * slot_floats + 32 * slot + 16)),
                     nat.C.c_void_p(self._epoch_ctr.data_ptr() + 4 * slot), world)
            slots.append((view, bufs, flags, flags_fused, guard, done))
        self._symm = (buf, hdl, n, slot_floats, slots, rank, world)
        self._ar_calls = 0

    def _symm_slot(self, n, device):
        """The peer-visible input slot of the next one-shot all-reduce (two slots, alternating per call)."""
        if self._symm is None or self._symm[2] != n:
            try:
                self._symm_init(n, device)
            except Exception as exc:      # no peer-mapped memory on this system (every rank fails alike): use NCCL
                import warnings
                warnings.warn(f"onda_b200: symmetric memory unavailable ({exc!r}); all-reducing with NCCL instead")
                self.allreduce, self._symm = "nccl", None
                return None
        return self._symm[4][self._ar_calls & 1][0]

    def _oneshot_args(self, sums, fused=False):
        """Stage ``sums`` in this call's symmetric slot (no copy when the fused pass wrote it there) and return
        (rank, world, peer slot pointers, peer flag pointers, epoch) of the exchange; for the exchange fused into
        ``ma`` the last item is the device pointer of the slot's epoch counter instead."""
        slot_view = self._symm_slot(sums.numel(), sums.device)
        if slot_view is None:
            return None
        _, _, _, _, slots, rank, world = self._symm
        slot = self._ar_calls & 1
        _, bufs, flags, flags_fused, _, done = slots[slot]
        if sums.data_ptr() != slot_view.data_ptr():
            slot_view.copy_(sums)                       # input was produced elsewhere: stage it
        self._ar_calls += 1
        if fused:
            return rank, world, bufs, flags_fused, (done, self._epoch_ctr.data_ptr() + 4 * slot)
        return rank, world, bufs, flags, self._ar_calls

    def _allreduce_oneshot(self, sums):
        """Sum over the ranks with onda_allreduce_oneshot: inputs in symmetric (peer-mapped) memory, two slots."""
        args = self._oneshot_args(sums)
        if args is None:                  # fell back to NCCL
            sharding.allreduce_sums(sums, self.process_group)
            return sums
        rank, world, bufs, flags, epoch = args
        n = sums.numel()
        out = torch.empty((n,), dtype=torch.float32, device=sums.device)
        with _on(sums.device):
            nat.check(self._lib.onda_allreduce_oneshot(nat.ptr(out), n, rank, world, bufs, flags, epoch,
                                                       _stream_ptr(sums.device)), "onda_allreduce_oneshot")
        return out

    def get_proto_array(self, feat, out):
        """(class sums (C, D), pixel counts (C,)) keyed by argmax of ``out`` (:76-81)."""
        sums, D, C, _ = self._class_sums(feat, out)
        return sums[:C * D].view(C, D).clone(), sums[2 * C * D:2 * C * D + C].clone()

    def ma(self, feat, out):
        """Moving-average prototype update (:88-99).  In place; ``counter`` is untouched."""
        self._deferred_monitor = None
        sums, D, C, device = self._class_sums(feat, out)
        P, S, _ = self._state(device, False)
        if not isinstance(S, torch.Tensor):
            raise AttributeError("squared_mean is not initialised")
        if P.shape != (C, D):
            raise ValueError(f"feat/out give a {C}x{D} update but prototypes are {tuple(P.shape)}")
        metric = self.distance_metric
        need_stats = metric == "mahalanobis"
        cnt = self.counter if isinstance(self.counter, torch.Tensor) else None
        fused_exchange = (self.process_group is not None and self.allreduce == "oneshot" and sums.is_cuda
                          and not (need_stats and cnt is None))
        args = self._oneshot_args(sums, fused=True) if fused_exchange else None
        fused_exchange = args is not None
        if fused_exchange:
            # all-reduce + blend + next distance table in ONE launch: the kernel reads the peers' slots over NVLink
            rank, world, bufs, flags, (done, epoch_ctr) = args
            reduced = self._buf(("reduced", self._ar_calls & 1), (sums.numel(),), torch.float32, device)
            table = self._buf(("table", metric), (self._lib.onda_table_floats(C, D),), torch.float32, device, zero=True)
            with _on(device):
              nat.check(self._lib.onda_ema_update_and_table_allreduce(
                nat.ptr(P), nat.ptr(S), nat.ptr(cnt) if need_stats else None, nat.ptr(reduced), C, D,
                float(self.ma_lambda), nat.METRIC[metric], nat.ptr(table), rank, world, bufs, flags, done, 0, epoch_ctr,
                _stream_ptr(device)), "onda_ema_update_and_table_allreduce")
            sums = reduced
        else:
            sums = self._allreduce(sums)
        if self.process_group is not None:       # the statistics tail is global (all ranks) only now
            self._stats_src, self._stats_cache = (sums, C, D), None
        if self._deferred_monitor is not None:
            self._monitor_side_effects(self._deferred_monitor, self.last_stats)
            self._deferred_monitor = None
        if fused_exchange:
            self._epoch += 1
            self._table = {metric: (self._table_key(P, S, cnt, need_stats, device), table)}
            return
        if need_stats and cnt is None:
            with _on(device):
                nat.check(self._lib.onda_ema_update(nat.ptr(P), nat.ptr(S), nat.ptr(sums), C, D, float(self.ma_lambda),
                                                    _stream_ptr(device)), "onda_ema_update")
            self._epoch += 1
            return
        # blend and rebuild the distance table for the next step in one launch
        table = self._buf(("table", metric), (self._lib.onda_table_floats(C, D),), torch.float32, device, zero=True)
        with _on(device):
            nat.check(self._lib.onda_ema_update_and_table(
                nat.ptr(P), nat.ptr(S), nat.ptr(cnt) if need_stats else None, nat.ptr(sums), C, D, float(self.ma_lambda),
                nat.METRIC[metric], nat.ptr(table), _stream_ptr(device)), "onda_ema_update_and_table")
        self._epoch += 1
        self._table = {metric: (self._table_key(P, S, cnt, need_stats, device), table)}

    def append_source_labels(self, feat, labels, num_classes=None):
        """``calculate_prototypes`` with ``STARTING_PROTO == "source"`` in one pass (prototypes.py:142-154): ``labels``
        (B, H, W) ground truth at any resolution is nearest-resized to the feature map like the reference does
        (``F.interpolate(labels.unsqueeze(1).float(), size=(h, w))``), pixels labelled 255 (or anything outside
        ``[0, num_classes)``) are skipped, and the class sums are taken straight from the NCHW ``feat`` -- no mask-gather,
        no transposition copy, no one-hot matrix -- before the cumulative ``append`` update (:62-74)."""
        self._require_cuda(feat, "feat")
        if feat.dim() != 4:
            raise ValueError(f"feat must be (B, D, h, w), got {tuple(feat.shape)}")
        B, D, h, w = feat.shape
        C = num_classes if num_classes is not None else self._num_classes()
        if C is None:
            raise ValueError("num_classes is needed before the prototypes exist")
        if C > nat.MAX_CLASSES:
            raise ValueError(f"{C} classes unsupported (max {nat.MAX_CLASSES})")
        labels = torch.as_tensor(labels)
        if labels.dim() != 3 or labels.shape[0] != B:
            raise ValueError(f"labels must be (B, H, W) with B={B}, got {tuple(labels.shape)}")
        device = feat.device
        ids = torch.nn.functional.interpolate(labels.to(device).unsqueeze(1).float(), size=(h, w)).view(-1).to(torch.int64)
        feat3 = self._nchw(feat, "feat")
        n = self._lib.onda_sums_floats(C, D)
        sums = torch.empty((n,), dtype=torch.float32, device=device)
        with _on(device):
            wbytes = self._lib.onda_fused_workspace_bytes(B, D, h * w, C, nat.IMPL["simt"])
            work = self._buf("work", (wbytes,), torch.uint8, device)
            nat.check(self._lib.onda_class_sums_labelled(nat.ptr(feat3), nat.ptr(ids), B, D, h * w, C, nat.ptr(sums),
                                                         nat.ptr(work), wbytes, _stream_ptr(device)), "onda_class_sums_labelled")
        self._append_sums(sums, D, C, device)

    def append(self, feat, out):
        """Cumulative-mean update used to initialise the prototypes (:62-74)."""
        sums, D, C, device = self._class_sums(feat, out)
        self._append_sums(sums, D, C, device)

    def _append_sums(self, sums, D, C, device):
        sums = self._allreduce(sums)
        if isinstance(self.prototypes, int):     # first call allocates the state (:68-70)
            self.prototypes = torch.zeros((C, D), dtype=torch.float32, device=device)
            self.squared_mean = torch.zeros((C, D), dtype=torch.float32, device=device)
        if not isinstance(self.squared_mean, torch.Tensor):     # state loaded from the legacy 2-tuple pickle: no second moments yet
            self.squared_mean = torch.zeros((C, D), dtype=torch.float32, device=device)
        if not isinstance(self.counter, torch.Tensor):
            self.counter = torch.full((C,), float(self.counter), dtype=torch.float32, device=device)
        P, S, cnt = self._state(device, True)
        if P.shape != (C, D):
            raise ValueError(f"feat/out give a {C}x{D} update but prototypes are {tuple(P.shape)}")
        with _on(device):
            nat.check(self._lib.onda_append_update(nat.ptr(P), nat.ptr(S), nat.ptr(cnt), nat.ptr(sums), C, D,
                                                   _stream_ptr(device)), "onda_append_update")
        self._epoch += 1
