"""``main_worker``-level drivers around the head: what src/train.py:40-163 and src/test.py:29-100 do once the
backbone, the loaders and ``args`` exist — build the transformer and its optimizer, load / save checkpoints in the
reference's own format, run ``do_epoch`` / ``validate_transformer``, keep ``best.pth`` / ``final.pth``.

The PSPNet backbone (``get_model``), the datasets and the yaml / argparse front end stay the reference's: the caller passes
``model`` and the loaders in, exactly the objects the reference builds at src/train.py:53,101-102 and src/test.py:45,94."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from .episodic import do_epoch, validate_transformer
from .transformer import MultiHeadAttentionOne


def get_model_dir_trans(args) -> str:
    """Directory of the transformer checkpoints, as the reference lays it out (src/util.py:167-179)."""
    return os.path.join(args.model_dir, args.train_name, f'split={args.train_split}', 'model', f'shot_{args.shot}',
                        f'transformer_{args.arch}{args.layers}')


def get_optimizer(args, parameters) -> torch.optim.Optimizer:
    """src/optimizer.py:8-19: SGD (momentum, weight decay, nesterov) or Adam over the given parameter groups."""
    if args.main_optim == 'SGD':
        return torch.optim.SGD(parameters, momentum=args.momentum, weight_decay=args.weight_decay, nesterov=args.nesterov)
    if args.main_optim == 'Adam':
        return torch.optim.Adam(parameters, weight_decay=args.weight_decay)
    raise ValueError(f"main_optim {args.main_optim!r} (the reference knows 'SGD' and 'Adam')")


def build_transformer(args, device) -> MultiHeadAttentionOne:
    """``MultiHeadAttentionOne(args.heads, d, d, d, dropout=0.5)`` with d = bottleneck_dim (src/train.py:94-96, src/test.py:55-56)."""
    d = args.bottleneck_dim
    return MultiHeadAttentionOne(args.heads, d, d, d, dropout=0.5).to(device)


def save_transformer_checkpoint(path: str, epoch: int, transformer, optimizer) -> None:
    """``{'epoch', 'state_dict', 'optimizer'}`` — the reference's checkpoint (src/train.py:138-163); the state-dict names and
    shapes are the reference module's, so its own src/test.py loads the file."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    torch.save({'epoch': epoch, 'state_dict': transformer.state_dict(), 'optimizer': optimizer.state_dict()}, path)


def load_transformer_checkpoint(path: str, transformer, optimizer: Optional[torch.optim.Optimizer] = None) -> int:
    """Load a checkpoint written by the reference (or by :func:`save_transformer_checkpoint`) — src/test.py:82-89. Returns its epoch."""
    assert os.path.isfile(path), path                                    # src/test.py:84
    checkpoint = torch.load(path, map_location="cpu")
    transformer.load_state_dict(checkpoint['state_dict'])
    if optimizer is not None and 'optimizer' in checkpoint:
        optimizer.load_state_dict(checkpoint['optimizer'])
    return int(checkpoint.get('epoch', -1))


def load_backbone_weights_for_test(model, path: str, verbose: bool = True) -> bool:
    """src/test.py:60-79: copy a stage-1 checkpoint into the backbone by POSITION (zip of the two key lists), skipping the
    classifier and entries whose shapes differ. Returns False when the file does not exist (the reference prints and goes on)."""
    if not os.path.isfile(path):
        if verbose:
            print("=> no weight found at '{}'".format(path))
        return False
    if verbose:
        print("=> loading weight '{}'".format(path))
    pre_weight = torch.load(path, map_location="cpu")['state_dict']
    pre_dict = model.state_dict()
    for index, (key1, key2) in enumerate(zip(pre_dict.keys(), pre_weight.keys())):
        if 'classifier' not in key1 and index < len(pre_dict.keys()):
            if pre_dict[key1].shape == pre_weight[key2].shape:
                pre_dict[key1] = pre_weight[key2]
            elif verbose:
                print('Pre-trained {} shape and model {} shape: {}, {}'.format(key2, key1, pre_weight[key2].shape, pre_dict[key1].shape))
    model.load_state_dict(pre_dict, strict=True)
    if verbose:
        print("=> loaded weight '{}'".format(path))
    return True


def load_backbone_weights_for_train(model, path: str, verbose: bool = True) -> bool:
    """src/train.py:55-72: copy a stage-1 checkpoint (saved from a DataParallel model: keys prefixed 'module.') into the
    backbone by NAME, skipping the classifier and 'gamma' entries and shapes that differ."""
    if not os.path.isfile(path):
        if verbose:
            print("=> no weight found at '{}'".format(path))
        return False
    pre_weight = torch.load(path, map_location="cpu")['state_dict']
    pre_dict = model.state_dict()
    for key in pre_dict.keys():
        if 'classifier' not in key and 'gamma' not in key:
            if pre_dict[key].shape == pre_weight['module.' + key].shape:
                pre_dict[key] = pre_weight['module.' + key]
            elif verbose:
                print('Mismatched shape {}: {}, {}'.format(key, pre_weight['module.' + key].shape, pre_dict[key].shape))
    model.load_state_dict(pre_dict, strict=True)
    return True


def freeze_backbone(model) -> None:
    """src/train.py:74-88: the backbone is frozen for the transformer stage."""
    for name in ("layer0", "layer1", "layer2", "layer3", "layer4", "ppm", "bottleneck"):
        part = getattr(model, name, None)
        if part is not None:
            for param in part.parameters():
                param.requires_grad = False


def train_worker(args, model, train_loader, val_loader, device=None, verbose: bool = True) -> Tuple[float, MultiHeadAttentionOne]:
    """The body of src/train.py:main (90-163) behind the loaders: transformer + optimizer, ``epochs`` x { do_epoch,
    validate_transformer, keep the best }, ``best.pth`` / ``final.pth`` in the reference's checkpoint format.
    Returns (max validation mIoU, the transformer)."""
    device = torch.device(device if device is not None else "cuda")
    transformer = build_transformer(args, device)
    optimizer_transformer = get_optimizer(args, [dict(params=transformer.parameters(), lr=args.trans_lr * args.scale_lr)])
    trans_save_dir = get_model_dir_trans(args)
    max_val_mIoU = 0.
    if getattr(args, "debug", False):
        iter_per_epoch = 5
    else:
        iter_per_epoch = args.iter_per_epoch if args.iter_per_epoch <= len(train_loader) else len(train_loader)
    log_iter = iter_per_epoch
    if verbose:
        print('==> Start training')
    for epoch in range(args.epochs):
        do_epoch(args=args, train_loader=train_loader, iter_per_epoch=iter_per_epoch, model=model, transformer=transformer,
                 optimizer_trans=optimizer_transformer, epoch=epoch, log_iter=log_iter, verbose=verbose)
        val_Iou, _val_loss = validate_transformer(args=args, val_loader=val_loader, model=model, transformer=transformer,
                                                  verbose=verbose)
        if float(val_Iou) > max_val_mIoU:                                # model selection (src/train.py:129-147)
            max_val_mIoU = float(val_Iou)
            if args.save_models:
                filename_transformer = os.path.join(trans_save_dir, 'best.pth')
                if verbose:
                    print('Saving checkpoint to: ' + filename_transformer)
                save_transformer_checkpoint(filename_transformer, epoch, transformer, optimizer_transformer)
        if verbose:
            print("=> Max_mIoU = {:.3f}".format(max_val_mIoU))
    if args.save_models:                                                 # last epoch (src/train.py:151-159)
        save_transformer_checkpoint(os.path.join(trans_save_dir, 'final.pth'), args.epochs, transformer, optimizer_transformer)
    return max_val_mIoU, transformer


def test_worker(args, model, val_loader, device=None, verbose: bool = True) -> Tuple[float, float]:
    """The body of src/test.py:main_worker (55-100) behind the loader: build the transformer, load ``{ckpt_used}.pth`` from the
    reference's directory layout when ``args.ckpt_used`` is set, run ``validate_transformer``. Returns (mIoU, loss)."""
    device = torch.device(device if device is not None else "cuda")
    transformer = build_transformer(args, device)
    if getattr(args, "ckpt_used", None) is not None:
        filepath = os.path.join(get_model_dir_trans(args), f'{args.ckpt_used}.pth')
        if verbose:
            print("=> loading transformer weight '{}'".format(filepath))
        load_transformer_checkpoint(filepath, transformer)
        if verbose:
            print("=> loaded transformer weight '{}'".format(filepath))
    elif verbose:
        print("=> Not loading anything")
    return validate_transformer(args=args, val_loader=val_loader, model=model, transformer=transformer, verbose=verbose)
