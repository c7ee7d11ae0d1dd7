#!/usr/bin/env python
"""Entry points of the hot path with the reference's CLI (run.py:21-93): `embed`, `evaluate` and `finetune`.

    python run.py evaluate HEAD.model -l DATASET_DIR [-b BASE.model] [--yfcc]
    python run.py embed CLIP.model -l DATASET_DIR
    python run.py finetune openai/clip-vit-large-patch14-336 -l DATASET_DIR [-m] [--heading]     # tower + head (run.py:165-169)
    python run.py finetune - -l EMBEDDING_DATASET_DIR --embeddings [-m]                          # head on embeddings (:171-175)

`pretrain` (CLIP contrastive pre-training through the HF Trainer) is outside the path and raises NotImplementedError, like
the reference does for its unsupported resume modes (run.py:166-173,189).  Datasets are HF `DatasetDict`s on disk, as in
the reference (run.py:143-162); none are shipped with either repository.
"""
import argparse
import logging

import torch

from pigeon_b200 import CLIPEmbedding, ProtoRefiner, SuperGuessr
from pigeon_b200 import config as cfg
from pigeon_b200.loops import embed_images, evaluate_model

logger = logging.getLogger('run')


def evaluate(model: str, dataset, yfcc: bool, landmarks: bool = False, base_model: str = None, heading: bool = False,
             refine: bool = True, metrics=None):
    """reference evaluation/evaluate.py:10-86."""
    embed_model = None
    if base_model is not None:
        from transformers import CLIPVisionModel
        embed_model = CLIPVisionModel.from_pretrained(cfg.CLIP_MODEL)                      # :36
        if base_model != cfg.CLIP_MODEL:
            from pigeon_b200 import load_state_dict
            load_state_dict(embed_model, torch.load(base_model, map_location='cpu'))       # :37-40
    full_model = SuperGuessr(embed_model, panorama=True, hierarchical=False, multi_task=False, heading=heading,
                             freeze_base=True, yfcc=yfcc, num_candidates=50)               # :42-44
    full_model.to('cuda')
    # :45-46 — the PIGEOTTO head first, then `model` on top: entries `model` does not carry keep the base head's values
    full_model.load_state(cfg.CLIP_PRETRAINED_HEAD_YFCC)
    full_model.load_state(model)
    refiner = None
    if refine:                                                                             # :50-80
        proto_model_path = cfg.PROTO_MODEL_YFCC_PATH if yfcc else cfg.PROTO_MODEL_PATH
        proto_path = cfg.PROTO_PATH_YFCC if yfcc else cfg.PROTO_PATH
        dataset_path = cfg.DATASET_PATH_YFCC if yfcc else cfg.DATASET_PATH
        if landmarks:                                                                      # :57-60: YFCC + landmarks bank
            proto_path = cfg.PROTO_PATH_LANDMARKS
            proto_model_path = cfg.PROTO_MODEL_LANDMARKS_PATH
            dataset_path = [cfg.DATASET_PATH_YFCC, cfg.DATASET_PATH_LANDMARKS]
        protos = None
        try:
            protos = torch.load(proto_model_path, map_location='cpu', weights_only=False).protos
        except FileNotFoundError:
            pass
        if protos is None:
            refiner = ProtoRefiner(20, False, 10000, proto_path=proto_path, dataset_path=dataset_path, temperature=1)
            torch.save(refiner, proto_model_path)
        else:
            refiner = ProtoRefiner(40, False, 100000, proto_path=proto_path, dataset_path=dataset_path, protos=protos,
                                   temperature=0.6, verbose=False)
        print(refiner)
    return evaluate_model(full_model, dataset, metrics, None, refiner)


def main():
    p = argparse.ArgumentParser(description='PIGEON hot path on B200 (embed / evaluate)')
    p.add_argument('mode', choices=['pretrain', 'finetune', 'embed', 'evaluate'])
    p.add_argument('name', help='model / prediction-head path')
    p.add_argument('-l', '--load', default=None, help='comma-separated HF dataset directories')
    p.add_argument('-b', '--base', default=None, help='base (vision tower) checkpoint')
    p.add_argument('-r', '--refine', action='store_true', default=True)
    p.add_argument('--yfcc', action='store_true')
    p.add_argument('--landmarks', action='store_true')
    p.add_argument('--heading', action='store_true')
    p.add_argument('-m', '--multitask', action='store_true')
    p.add_argument('--embeddings', action='store_true', help='finetune: train the head on pre-computed embeddings')
    args = p.parse_args()
    if args.mode == 'pretrain':
        raise NotImplementedError('"pretrain" (CLIP contrastive pre-training) is outside the B200 hot path')
    if args.load is None:
        raise NotImplementedError('A dataset must be given with -l (the reference regenerates it from un-shipped raw data).')
    from datasets import DatasetDict
    dataset = DatasetDict.load_from_disk(args.load.split(',')[0])
    if args.mode == 'finetune':
        from pigeon_b200.training import finetune_model, finetune_on_embeddings
        ds = dataset.with_format('torch')
        if args.embeddings:
            finetune_on_embeddings(ds, multi_task=args.multitask, heading=args.heading, yfcc=args.yfcc)   # run.py:175
        else:
            finetune_model(args.name, ds, multi_task=args.multitask, heading=args.heading, yfcc=args.yfcc)  # run.py:169
    elif args.mode == 'embed':
        model = CLIPEmbedding(args.name, load_checkpoint=True, panorama=not args.yfcc)    # run.py:126-129
        embed_images(model, dataset)
    else:
        split = dataset['test'] if 'test' in dataset else dataset[list(dataset.keys())[-1]]
        print(evaluate(args.name, split.with_format('torch'), args.yfcc, args.landmarks, base_model=args.base,
                       heading=args.heading, refine=args.refine))


if __name__ == '__main__':
    torch.multiprocessing.set_start_method('spawn', force=True)                           # run.py:192
    main()
