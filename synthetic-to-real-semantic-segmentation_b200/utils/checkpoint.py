"""Checkpoint I/O in the reference's layouts, so that files written by the reference's
scripts drive the B200 path and vice versa:

* train_adapt.py:204-209 / :96-109 -- {'epoch', 'state_dict', 'optimizer', 'best_pred'} for the segmentation network
  and its SGD optimizer (the discriminator is not stored by the reference);
* train.py:242-253 / :123-146 -- {'epoch', '{backbone,assp,y,d}_model_state_dict', 'task_optimizer', 'd_optimizer',
  'd_inv_optimizer', 'c_optimizer', 'best_pred'}.

Module state_dicts keep the reference's key names and NCHW/OIHW fp32 shapes (parameters ARE the fp32 masters; the bf16
packed filter copies are derived data, re-packed automatically after a load because the parameter version changes).
The fused optimizers read and write torch.optim-format state (optim._FusedBase.state_dict).  `module.` prefixes left by
nn.DataParallel (the reference saves `model.module.state_dict()`, but checkpoints saved from the wrapper exist in the
wild) are stripped on load.
"""
import os

import torch


def _strip_module(sd):
    if sd and all(k.startswith('module.') for k in sd):
        return {k[len('module.'):]: v for k, v in sd.items()}
    return sd


def adapt_state(model, optimizer, epoch, best_pred):
    """The dict train_adapt.py:204-209 hands to Saver.save_checkpoint."""
    return {'epoch': epoch + 1, 'state_dict': model.state_dict(), 'optimizer': optimizer.state_dict(),
            'best_pred': best_pred}


def load_adapt(path_or_state, model, optimizer=None, ft=False):
    """train_adapt.py:96-109: returns (start_epoch, best_pred)."""
    ck = path_or_state
    if not isinstance(ck, dict):
        if not os.path.isfile(ck):
            raise RuntimeError("=> no checkpoint found at '{}'".format(ck))
        ck = torch.load(ck, map_location='cpu')
    model.load_state_dict(_strip_module(ck['state_dict']))
    if not ft and optimizer is not None:
        optimizer.load_state_dict(ck['optimizer'])
    return (0 if ft else ck['epoch']), ck['best_pred']


def feature_state(backbone_model, assp_model, y_model, d_model, task_optimizer, d_optimizer, d_inv_optimizer, epoch,
                  best_pred, c_optimizer=None):
    """The dict train.py:242-253 stores (c_optimizer never steps in the reference; its state is kept when given)."""
    return {'epoch': epoch + 1,
            'backbone_model_state_dict': backbone_model.state_dict(),
            'assp_model_state_dict': assp_model.state_dict(),
            'y_model_state_dict': y_model.state_dict(),
            'd_model_state_dict': d_model.state_dict(),
            'task_optimizer': task_optimizer.state_dict(),
            'd_optimizer': d_optimizer.state_dict(),
            'd_inv_optimizer': d_inv_optimizer.state_dict(),
            'c_optimizer': c_optimizer.state_dict() if c_optimizer is not None else {'state': {}, 'param_groups': []},
            'best_pred': best_pred}


def load_feature(path_or_state, backbone_model, assp_model, y_model, d_model, task_optimizer=None, d_optimizer=None,
                 d_inv_optimizer=None, ft=False):
    """train.py:123-146: returns (start_epoch, best_pred)."""
    ck = path_or_state
    if not isinstance(ck, dict):
        if not os.path.isfile(ck):
            raise RuntimeError("=> no checkpoint found at '{}'".format(ck))
        ck = torch.load(ck, map_location='cpu')
    for mod, key in ((backbone_model, 'backbone_model_state_dict'), (assp_model, 'assp_model_state_dict'),
                     (y_model, 'y_model_state_dict'), (d_model, 'd_model_state_dict')):
        mod.load_state_dict(_strip_module(ck[key]))
    if not ft:
        for opt, key in ((task_optimizer, 'task_optimizer'), (d_optimizer, 'd_optimizer'), (d_inv_optimizer, 'd_inv_optimizer')):
            if opt is not None:
                opt.load_state_dict(ck[key])
    return (0 if ft else ck['epoch']), ck.get('best_pred', 0.0)
