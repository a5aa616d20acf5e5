"""few_shot_seg_cwt_b200 — B200-native per-episode few-shot segmentation head of CWT.

Hand-written sm_100a CUDA kernels behind a C ABI (include/cwt_b200.h, lib/libcwt_b200.so),
wrapped with the reference's own Python call surface:

    MultiHeadAttentionOne          src/model/transformer.py:33-83
    fit_classifier / inner_loop    src/test.py:164-187, src/model/pspnet.py:189-205
    batch_intersectionAndUnionGPU  src/util.py:237-277
    intersectionAndUnionGPU        src/util.py:280-308
    validate_transformer           src/test.py:103-254
    meta_train_step                src/train.py:233-267
    do_epoch                       src/train.py:166-290
    train_worker / test_worker     src/train.py:90-163, src/test.py:55-100 (checkpoints in the reference's format)

Importing the package never needs a GPU; calling an op without the built library or without a CUDA
tensor raises (there is no CPU fallback).
"""
from . import _lib, synthetic  # noqa: F401
from .transforms import ValTransform, find_new_hw, resize_pad_normalize  # noqa: F401
from .classifier import CosCls, draw_initial_weights, get_classifier, increment_inner_loop, inner_loop  # noqa: F401
from .episodic import (HeadOutput, HeadPipeline, HostPipeline, IoUTable, bind_host_to_gpu, do_epoch, episode_head, meta_train_step, query_loss, run_sweep,  # noqa: F401
                       transformer_params, validate_transformer)
from .metrics import (batch_intersection_union_int, batch_intersectionAndUnionGPU,  # noqa: F401
                      intersection_union_int, intersectionAndUnionGPU)
from .ops import (feat_times_rows, normalize_features, fit_classifier, fit_classifier_bias, fit_classifier_dice, fit_coscls, fit_multiclass, intersection_union, label_counts, logits_iou,  # noqa: F401
                  query_loss_grad, rows_times_feat, transformer_backward, transformer_forward,
                  upsample_argmax_iou)
from .ops import fit_status, raise_for_status  # noqa: F401
from .hostformat import CompressedEpisodeBatch, compress_batch, compress_map, expand_map, expand_map_reference  # noqa: F401
from .transformer import MultiHeadAttentionOne  # noqa: F401
from .drivers import (build_transformer, freeze_backbone, get_model_dir_trans, get_optimizer, load_backbone_weights_for_test,  # noqa: F401
                      load_backbone_weights_for_train, load_transformer_checkpoint, save_transformer_checkpoint, test_worker,
                      train_worker)

__all__ = [
    "MultiHeadAttentionOne", "fit_classifier", "fit_classifier_bias", "fit_classifier_dice", "fit_coscls", "fit_multiclass", "ValTransform", "find_new_hw", "resize_pad_normalize", "inner_loop", "increment_inner_loop", "get_classifier", "CosCls", "draw_initial_weights",
    "batch_intersectionAndUnionGPU", "intersectionAndUnionGPU", "batch_intersection_union_int",
    "intersection_union_int", "validate_transformer", "episode_head", "run_sweep", "IoUTable", "HostPipeline", "HeadPipeline", "bind_host_to_gpu",
    "meta_train_step", "do_epoch", "query_loss", "transformer_forward", "transformer_backward", "logits_iou",
    "upsample_argmax_iou", "intersection_union", "label_counts", "query_loss_grad",
    "rows_times_feat", "feat_times_rows", "synthetic", "fit_status", "raise_for_status",
    "train_worker", "test_worker", "save_transformer_checkpoint", "load_transformer_checkpoint", "get_model_dir_trans",
    "get_optimizer", "build_transformer", "freeze_backbone", "load_backbone_weights_for_test", "load_backbone_weights_for_train",
    "CompressedEpisodeBatch", "compress_batch", "compress_map", "expand_map", "expand_map_reference",
]


def register_torch_ops() -> None:
    """Register ``torch.ops.cwt_b200.*`` custom ops in front of the C ABI."""
    from .ops import _register_custom_ops
    _register_custom_ops()
