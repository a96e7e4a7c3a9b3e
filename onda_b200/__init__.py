"""onda_b200: B200-native implementation of OnDA's prototype pseudo-labelling hot path.

Public surface (mirrors the reference's names):

* ``prototype_handler`` -- drop-in for ``framework.domain_adaptation.methods.prototype_handler``
* ``Monitor``, ``HybridSelect``, ``DevSelect``, ``static_share`` -- host-side switch logic
* ``onda_b200.methods`` -- ``prototype_predictions`` for the base / h-switch / v-switch / hybrid
  method classes, to be bound onto the reference's ``online_proDA`` subclasses
* ``update_ema`` / ``WeightEma`` -- the model-weight EMA of ``online_proDA.update_ema`` as one launch;
  ``update_dynamic`` -- the model snapshot of ``online_proDA.update_dynamic`` as one launch
* ``ConfusionMeter`` -- upsample + argmax + confusion matrix of ``da_model.evaluate`` in one kernel
* ``target_losses`` -- CE + RCE + MRKLD/MRENT of ``online_proDA.pseudolabel_loss`` on the student logits, forward and
  gradient in one kernel (a ``torch.autograd.Function``)

Importing the package loads ``libonda_b200.so`` lazily (on first handler construction); if the
library has not been built the constructor raises -- there is no CPU or PyTorch fallback.
"""
from .switching import Monitor, HybridSelect, DevSelect, static_share  # noqa: F401
from .handler import prototype_handler  # noqa: F401
from .ema import update_ema, update_dynamic, WeightEma  # noqa: F401
from .evaluation import ConfusionMeter  # noqa: F401
from .losses import target_losses  # noqa: F401

__all__ = ["prototype_handler", "Monitor", "HybridSelect", "DevSelect", "static_share", "update_ema", "update_dynamic", "WeightEma", "ConfusionMeter", "target_losses"]
