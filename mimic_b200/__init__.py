"""mimic_b200 — B200-native (sm_100a) implementation of MimIC's hot path: shifted self-attention,
layer-wise alignment loss, LM-loss cross-entropy and the record-hook path, behind the reference's
`AttnApproximator` / `ShiftModel` API.  Importing the package loads libmimic_b200.so (and raises if
it is missing and cannot be built): there is no CPU or eager fallback."""
from . import _lib  # noqa: F401  (fail loudly, first)
from . import ops
from .model_base import HookedModel, try_inject_params
from .shift_encoder import (AttnApproxHandle, AttnApproximator, AttnFFNShift, BaseHookEncoder, MultiheadLinear, RecordPlan,
                            ShiftStrategy)
from .shift_model import ShiftModel, Strategy

__all__ = ["ops", "HookedModel", "try_inject_params", "AttnApproxHandle", "AttnApproximator", "AttnFFNShift", "BaseHookEncoder",
           "MultiheadLinear", "RecordPlan", "ShiftStrategy", "ShiftModel", "Strategy"]
