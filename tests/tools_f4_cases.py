"""The FIR parameterisations of tests/golden/f4_op_grads.npz (tools/make_golden_f4.py:FIR_CASES)."""
FIR_CASES = [('up2',), ('down2',), ('same',), ('generic',), ('crop',)]


import torch


class StandIn(torch.nn.Module):
  """The deterministic stand-in score model of the SMLD goldens (tools/make_golden_f4.py:StandIn)."""

  def forward(self, x, labels):
    return 0.3 * torch.flip(x, dims=(3,)) + (0.001 * labels.float())[:, None, None, None] * x
