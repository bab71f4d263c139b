"""Checkpoint I/O with the reference's file format (``utils.py:7-30``): one ``torch.save`` dict with the
keys ``optimizer``, ``model``, ``ema``, ``step``.  The reference goes through ``tf.io.gfile``; plain
``os`` calls are used here (local paths), everything else - ``strict=False`` model loading, returning
the unchanged state with a warning when the file does not exist - is the same behaviour."""
import logging
import os

import torch


def restore_checkpoint(ckpt_dir, state, device):
  """``state`` is a dict with ``optimizer`` (may be ``None`` for sampling-only use), ``model``, ``ema``, ``step``."""
  if not os.path.exists(ckpt_dir):
    os.makedirs(os.path.dirname(ckpt_dir) or '.', exist_ok=True)
    logging.warning(f"No checkpoint found at {ckpt_dir}. Returned the same state as input")
    return state
  try:
    # the checkpoint dict holds tensors, lists and scalars only: refuse to unpickle anything else
    loaded_state = torch.load(ckpt_dir, map_location=device, weights_only=True)
  except Exception as err:   # pragma: no cover - legacy files with pickled objects (e.g. a numpy scalar 'step')
    logging.warning(f"{ckpt_dir}: weights_only load failed ({type(err).__name__}); falling back to a full unpickle")
    loaded_state = torch.load(ckpt_dir, map_location=device, weights_only=False)
  if state.get('optimizer') is not None:
    state['optimizer'].load_state_dict(loaded_state['optimizer'])
  state['model'].load_state_dict(loaded_state['model'], strict=False)
  state['ema'].load_state_dict(loaded_state['ema'])
  state['step'] = loaded_state['step']
  return state


def save_checkpoint(ckpt_dir, state, dataparallel_prefix=True):
  """``dataparallel_prefix=True`` (default) writes the model keys with the ``module.`` prefix the reference's
  checkpoints carry (its ``create_model`` wraps the network in ``nn.DataParallel``, ``models/utils.py:93``, and its
  ``restore_checkpoint`` loads with ``strict=False``: un-prefixed keys would silently load nothing there).
  This package's ``restore_checkpoint`` / ``NCSNpp.load_state_dict`` accept both spellings."""
  model_sd = state['model'].state_dict()
  if dataparallel_prefix and not any(k.startswith('module.') for k in model_sd):
    model_sd = {'module.' + k: v for k, v in model_sd.items()}
  saved_state = {
    'optimizer': state['optimizer'].state_dict() if state.get('optimizer') is not None else {},
    'model': model_sd,
    'ema': state['ema'].state_dict(),
    'step': state['step'],
  }
  torch.save(saved_state, ckpt_dir)
