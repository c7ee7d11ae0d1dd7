"""CLIPEmbedding — CLIP ViT-L/14-336 image embedder, B200 execution behind the reference's interface.

Mirror of reference models/clip_embedder.py: `CLIPEmbedding(model_name, device='cuda', load_checkpoint=False,
panorama=False)`, `forward(image) -> Tensor[N, 1024]` = mean over all tokens of last_hidden_state
(:63-65), no_grad, not trainable.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor

from .config import CLIP_MODEL
from .model_utils import load_state_dict
from .super_guessr import CLIPVisionTower, as_tower
from .vit_engine import VitDims


class CLIPEmbedding(torch.nn.Module):
    def __init__(self, model_name: str, device: str = 'cuda', load_checkpoint: bool = False, panorama: bool = False,
                 clip_model=None, processor=None):
        """Arguments as in the reference (:11-23).

        Extension for offline use (the reference downloads `CLIP_MODEL` from the hub, :25-26): `clip_model=` a
        `CLIPVisionTower` or HF `CLIPVisionModel` to wrap, `processor=` a callable like HF `CLIPProcessor`."""
        super().__init__()
        self.device = device
        if clip_model is None:
            from transformers import CLIPVisionModel
            clip_model = CLIPVisionModel.from_pretrained(CLIP_MODEL)
        # :25 `CLIPProcessor.from_pretrained(CLIP_MODEL)` — here the GPU pre-processor (bit-identical output, no hub access),
        # built on first use so that a model fed pre-processed tensors never needs it
        self.processor = processor
        self.clip_model = as_tower(clip_model)
        self.panorama = panorama

        if load_checkpoint:                                                  # :29-32
            state_dict = torch.load(model_name, map_location=torch.device('cuda'))
            load_state_dict(self.clip_model.base_model, state_dict, embedder=True)
            print('Loaded embedder from checkpoint:', model_name)

        if type(device) == str:                                              # :34-37
            self.clip_model = self.clip_model.to(self.device)
        else:
            self.clip_model = self.clip_model.cuda(self.device)
        self.eval()

    def _get_embedding(self, image) -> Tensor:
        """:42-66 — accepts a pre-processed tensor or PIL image(s)."""
        with torch.no_grad():
            if isinstance(image, Tensor) == False:
                if self.processor is None:
                    from .preprocess import ClipImageProcessor
                    dev = next(self.clip_model.parameters()).device
                    self.processor = ClipImageProcessor(size=self.clip_model.dims.image_size, device=dev, dtype=torch.float16)
                inputs = self.processor(images=image, return_tensors='pt')
                pixel_values = inputs['pixel_values']
            else:
                pixel_values = image
            if type(self.device) == str:
                pixel_values = pixel_values.to(self.device)
            else:
                pixel_values = pixel_values.cuda(self.device)
            return self.clip_model.embed(pixel_values)                       # ViT forward + torch.mean(dim=1), fused

    def forward(self, image: Dict) -> Tensor:
        return self._get_embedding(image)
