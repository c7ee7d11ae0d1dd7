"""Constants of the hot path, same names and values as the reference's config.py (file:line cited per block).
Only what the hot path and the fine-tune step read is mirrored; dataset-creation paths are out of scope."""
from types import SimpleNamespace

# OpenAI's pretrained implementation                                   (config.py:5-7)
CLIP_MODEL = 'openai/clip-vit-large-patch14-336'
CLIP_EMBED_DIM = 1024

# Geocells path                                                        (config.py:34-36)
GEOCELL_PATH = 'data/geocells_2203.csv'       # PIGEON
GEOCELL_PATH_YFCC = 'data/geocells_yfcc.csv'  # PIGEOTTO

# Haversine smoothing constant                                         (config.py:54-56)
LABEL_SMOOTHING_CONSTANT = 65

# Models                                                               (config.py:58-69)
CURRENT_SAVE_PATH = 'saved_models/WorldCLIP_head_landmarks.model'
PRETRAINED_CLIP = 'saved_models/StreetviewCLIP.model'
CLIP_PRETRAINED_HEAD = 'saved_models/New_Base_smooth_avg_MT_Geo_SV.model'
PRETRAINED_CLIP_YFCC = 'saved_models/WorldCLIP.model'
CLIP_PRETRAINED_HEAD_YFCC = 'saved_models/WorldCLIP_head.model'
CLIP_PRETRAINED_HEAD_YFCC_LANDMARKS = 'saved_models/WorldCLIP_head_landmarks.model'

# Embedding                                                            (config.py:70-71)
EMBED_BATCH_SIZE_PER_GPU = 512

# Cluster refinement model                                             (config.py:73-90)
PROTO_PATH = 'data/data_prototypes_2203.csv'
DATASET_PATH = 'data/hf_SVCLIP_2203'
PROTO_MODEL_PATH = 'saved_models/refiner/proto.refiner'
PROTO_PATH_YFCC = 'data/data_prototypes_YFCC.csv'
DATASET_PATH_YFCC = 'data/hf_YFCC'
PROTO_MODEL_YFCC_PATH = 'saved_models/refiner/proto_YFCC.refiner'
PROTO_PATH_LANDMARKS = 'data/data_prototypes_landmarks.csv'
DATASET_PATH_LANDMARKS = 'data/hf_landmarks'
PROTO_MODEL_LANDMARKS_PATH = 'saved_models/refiner/proto_landmarks.refiner'

# Evaluation batch size (TRAIN_ARGS.per_device_eval_batch_size)        (config.py:98)
EVAL_BATCH_SIZE = 256

# Fine-tuning arguments (the fields training/train_eval_loop.py reads from TRAIN_ARGS)   (config.py:94-109)
TRAIN_ARGS = SimpleNamespace(output_dir='saved_models', per_device_train_batch_size=256, per_device_eval_batch_size=256,
                             num_train_epochs=1000, learning_rate=2e-5, logging_steps=1, gradient_accumulation_steps=1,
                             seed=330)
