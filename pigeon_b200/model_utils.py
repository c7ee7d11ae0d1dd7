"""Types and helpers shared by the host mirrors (the counterpart of the reference's models/utils.py)."""
from __future__ import annotations

from collections import namedtuple
from typing import Mapping

import torch

# Field names and order of the reference's ModelOutput (models/utils.py:7-9); its consumers read the fields by name
# (training/train_eval_loop.py:80-112).
_MODEL_OUTPUT_FIELDS = ('loss', 'loss_clf', 'loss_reg', 'loss_climate', 'loss_month',
                        'preds_LLH', 'preds_geocell', 'preds_mt', 'preds_climate', 'preds_month',
                        'top5_geocells', 'embedding')
ModelOutput = namedtuple('ModelOutput', _MODEL_OUTPUT_FIELDS)

# `.values` / `.indices`, the shape of torch.topk's result that the evaluation loop reads (train_eval_loop.py:101-112)
TopK = namedtuple('topk', ('values', 'indices'))


def load_state_dict(self, state_dict: Mapping[str, torch.Tensor], embedder: bool = False) -> None:
    """Best-effort load by parameter name, same behaviour as reference models/utils.py:24-45: with `embedder` the leading
    component of keys that mention 'base_model' is dropped, keys the module does not have are reported on stdout (same
    message) and skipped, serialized `Parameter`s are unwrapped.  Afterwards the packed kernel copies are invalidated."""
    target = self.state_dict()
    for key, value in state_dict.items():
        if embedder and 'base_model' in key:
            key = key.split('.', 1)[1]
        dst = target.get(key)
        if dst is None:
            print(f"Parameter {key} not in model's state.")
            continue
        dst.copy_(value.data if isinstance(value, torch.nn.Parameter) else value)
    invalidate = getattr(self, '_weights_changed', None)
    if invalidate is not None:
        invalidate()
