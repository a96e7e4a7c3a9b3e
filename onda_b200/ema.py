"""Model-weight EMA as one launch (``update_ema``, framework/domain_adaptation/methods/prototypes.py:407-416).

The reference walks ``zip(model.parameters(), ema_model.parameters())`` in Python, clones both tensors and runs three
elementwise kernels per parameter, then copies every buffer: ~1 500 launches and ~0.7 GB of traffic per step for
DeepLabV2-ResNet50.  Here the two models are cut once into a table of chunks (``onda_ema_chunk``) and every step is a
single kernel that streams them: ``k = k*a + q*(1-a)`` with both products rounded before the add (bit-identical to the
reference expression), buffers copied byte for byte.  Updates are in place (the reference rebinds ``param_k.data``).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as nat

CHUNK_BYTES = 32768      # ONDA_EMA_CHUNK_BYTES


class _Chunk(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("count", C.c_uint32), ("mode", C.c_uint32)]


def _rows(src, dst, mode):
    """Chunk rows (src ptr, dst ptr, count, mode) of one tensor pair."""
    if src.shape != dst.shape or src.dtype != dst.dtype:
        raise ValueError(f"model and EMA model disagree: {tuple(src.shape)} {src.dtype} vs {tuple(dst.shape)} {dst.dtype}")
    if not (src.is_cuda and dst.is_cuda):
        raise RuntimeError("onda_b200 runs on CUDA devices only (there is no CPU path)")
    if not (src.is_contiguous() and dst.is_contiguous()):
        raise ValueError("parameters and buffers must be contiguous")
    if mode == 0 and src.dtype != torch.float32:
        raise TypeError(f"weight EMA expects float32 parameters, got {src.dtype}")
    nbytes = src.numel() * src.element_size()
    unit = 4 if mode == 0 else 1
    rows, off = [], 0
    while off < nbytes:
        n = min(CHUNK_BYTES, nbytes - off)
        rows.append((src.data_ptr() + off, dst.data_ptr() + off, n // unit, mode))
        off += n
    return rows


class WeightEma:
    """The chunk table of a (model, ema_model) pair; ``update(a)`` is one launch.  ``copy_only=True`` makes every row a
    byte copy: the table of a snapshot (``update_dynamic``)."""

    def __init__(self, model, ema_model, copy_only=False):
        self._lib = nat.load()
        params = list(zip(model.parameters(), ema_model.parameters()))
        buffers = list(zip(model.buffers(), ema_model.buffers()))
        self._keys = [(q.data_ptr(), k.data_ptr()) for q, k in params + buffers]
        self._pairs = params + buffers          # keeps the tensors alive
        rows = []
        for q, k in params:
            if q.numel():
                rows += _rows(q.data, k.data, 1 if copy_only else 0)
        for q, k in buffers:
            if q.numel():
                rows += _rows(q.data, k.data, 1)
        self.n_chunks = len(rows)
        self.device = params[0][1].device if params else (buffers[0][1].device if buffers else torch.device("cuda"))
        table = (_Chunk * max(1, self.n_chunks))(*[_Chunk(*r) for r in rows])
        raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8) if self.n_chunks else torch.zeros(16, dtype=torch.uint8)
        self._table = raw.to(self.device)
        self.param_bytes = sum(q.numel() * 4 for q, _ in params)
        self.buffer_bytes = sum(q.numel() * q.element_size() for q, _ in buffers)

    def still_valid(self, model, ema_model):
        pairs = list(zip(model.parameters(), ema_model.parameters())) + list(zip(model.buffers(), ema_model.buffers()))
        return len(pairs) == len(self._keys) and all((q.data_ptr(), k.data_ptr()) == key for (q, k), key in zip(pairs, self._keys))

    def update(self, ema_update):
        keep = float(ema_update)
        take = 1.0 - keep                              # computed in double like the reference, rounded to fp32 by the call
        with torch.cuda.device(self.device):          # the C ABI launches on the current device
            stream = nat.C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            nat.check(self._lib.onda_weight_ema_update(nat.ptr(self._table), self.n_chunks, keep, take, stream),
                      "onda_weight_ema_update")


_cache = {}


def update_ema(model, ema_model, ema_update):
    """Drop-in body of ``online_proDA.update_ema``: ``update_ema(self.model, self.ema_model, self.cfg_spec.EMA_UPDATE)``."""
    key = (id(model), id(ema_model))
    plan = _cache.get(key)
    if plan is None or not plan.still_valid(model, ema_model):
        plan = _cache[key] = WeightEma(model, ema_model)
    plan.update(ema_update)


_snap_cache = {}


def update_dynamic(model, dynamic_model):
    """The copy inside ``online_proDA.update_dynamic`` (prototypes.py:99-102: ``self.dynamic_model = deepcopy(self.model)``,
    fired by ``evaluate_update_dynamic`` :396-405 when the confidence derivative leaves its band): every parameter and
    buffer of ``model`` is copied into the EXISTING ``dynamic_model`` in one launch (byte for byte) instead of a module
    deep copy (allocation plus one copy kernel per tensor).  The caller keeps the reference's follow-up
    (``models_default_config()``: train/eval modes).  Returns ``dynamic_model``."""
    key = (id(model), id(dynamic_model))
    plan = _snap_cache.get(key)
    if plan is None or not plan.still_valid(model, dynamic_model):
        plan = _snap_cache[key] = WeightEma(model, dynamic_model, copy_only=True)
    plan.update(0.0)          # the blend factors are unused by copy rows
    return dynamic_model
