"""Dump the REAL reference's NCSNpp parameter list (name, shape, requires_grad, in `parameters()` order) for the golden
configurations into tests/golden/param_order.json.  The EMA shadow list of a reference checkpoint is positional over
this list (models/ema.py:27-28), so the engine-backed module must register its parameters identically.
Build container only (reads /root/reference):  python tools/make_param_order.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg   # noqa: E402


def main():
  _, _, ncsnpp, _, _ = mg.import_reference()
  out = {}
  for name, (cfg, _) in mg.golden_configs().items():
    torch.manual_seed(0)
    ref = ncsnpp.NCSNpp(cfg)
    out[name] = [[n, list(p.shape), bool(p.requires_grad)] for n, p in ref.named_parameters()]
  path = os.path.join(mg.OUT, 'param_order.json')
  json.dump(out, open(path, 'w'))
  print('wrote', path, {k: len(v) for k, v in out.items()})


if __name__ == '__main__':
  main()
