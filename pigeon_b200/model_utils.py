"""API types shared by the modules — mirrors reference models/utils.py."""
from collections import namedtuple
from typing import Dict

import torch
from torch.nn.parameter import Parameter

# models/utils.py:7-9 — same field order, consumers index by name (training/train_eval_loop.py:80-112)
ModelOutput = namedtuple('ModelOutput', 'loss loss_clf loss_reg loss_climate loss_month \
                         preds_LLH preds_geocell preds_mt preds_climate preds_month \
                         top5_geocells embedding')

# what torch.topk returns (`.values`, `.indices`), read at training/train_eval_loop.py:101-102,110-112
TopK = namedtuple('topk', 'values indices')


def load_state_dict(self, state_dict: Dict, embedder: bool = False):
    """models/utils.py:24-45 — copy parameters by name wherever possible; unknown keys are printed, never raised."""
    own_state = self.state_dict()
    for name, param in state_dict.items():
        if embedder and 'base_model' in name:
            name = '.'.join(name.split('.')[1:])
        if name not in own_state:
            print(f'Parameter {name} not in model\'s state.')
            continue
        if isinstance(param, Parameter):
            param = param.data
        own_state[name].copy_(param)
    if hasattr(self, '_weights_changed'):
        self._weights_changed()
