"""Exponential moving average of a parameter list, interface-compatible with the reference's
``models/ema.py:9-107`` (``update``, ``copy_to``, ``store``, ``restore``, ``state_dict`` /
``load_state_dict`` with the keys ``decay``, ``num_updates``, ``shadow_params``), so a reference
checkpoint's ``'ema'`` entry loads unchanged (``utils.restore_checkpoint``, run_lib.py:276-284).

Shadow parameters are positional: entry ``i`` belongs to the ``i``-th trainable parameter of
``model.parameters()``.  The engine-backed ``NCSNpp`` registers its parameters in the reference's
order with the same ``requires_grad`` flags (the Fourier projection ``W`` is frozen in both), which
``tests/test_host_cpu.py`` pins against the reference's own parameter list.

One addition: the engine keeps a packed copy of the weights on the device, so writing averaged
values into a model's parameters (``copy_to`` / ``restore``) also tells the owning module to repack
(``NCSNpp.invalidate_weights``); with the reference's class one has to call that by hand."""
import torch


def _trainable(parameters):
  return [p for p in parameters if p.requires_grad]


def _notify_owners(parameters):
  seen = set()
  for p in parameters:
    ref = getattr(p, '_b200_owner', None)
    owner = ref() if ref is not None else None
    if owner is not None and id(owner) not in seen:
      seen.add(id(owner))
      owner.invalidate_weights()


class ExponentialMovingAverage:
  """``shadow <- shadow - (1 - d) * (shadow - param)`` with ``d = min(decay, (1 + n) / (10 + n))`` after
  ``n`` updates when ``use_num_updates`` (models/ema.py:34-51)."""

  def __init__(self, parameters, decay, use_num_updates=True):
    if decay < 0.0 or decay > 1.0:
      raise ValueError('Decay must be between 0 and 1')
    self.decay = decay
    self.num_updates = 0 if use_num_updates else None
    self.shadow_params = [p.clone().detach() for p in _trainable(parameters)]
    self.collected_params = []

  def update(self, parameters):
    decay = self.decay
    if self.num_updates is not None:
      self.num_updates += 1
      decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
    keep = 1.0 - decay
    with torch.no_grad():
      for shadow, p in zip(self.shadow_params, _trainable(parameters)):
        shadow.sub_(keep * (shadow - p))

  def copy_to(self, parameters):
    """Write the averages into ``parameters`` (the model then samples with EMA weights)."""
    params = _trainable(list(parameters))
    with torch.no_grad():
      for shadow, p in zip(self.shadow_params, params):
        p.data.copy_(shadow.data)
    _notify_owners(params)

  def store(self, parameters):
    """Remember the current values so ``restore`` can put them back after an EMA evaluation."""
    self.collected_params = [p.clone() for p in parameters]

  def restore(self, parameters):
    params = list(parameters)
    with torch.no_grad():
      for saved, p in zip(self.collected_params, params):
        p.data.copy_(saved.data)
    _notify_owners(params)

  def state_dict(self):
    return dict(decay=self.decay, num_updates=self.num_updates, shadow_params=self.shadow_params)

  def load_state_dict(self, state_dict):
    self.decay = state_dict['decay']
    self.num_updates = state_dict['num_updates']
    self.shadow_params = state_dict['shadow_params']
