"""world_size-2 gloo tests (CPU) of the sharding host logic: one parameter broadcast, per-rank seeds,
ragged shard sizes, gather order.  The N>1 GPU path runs the same code over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from score_sde_pytorch_b200 import distributed as bdist


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, total, q):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                    LOCAL_RANK=str(rank))
  r, w, _ = bdist.init_distributed('gloo')
  assert (r, w) == (rank, world)
  torch.manual_seed(100 + rank)                   # ranks start from different weights
  model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3))
  model[1].running_mean.fill_(float(rank))
  bdist.broadcast_parameters(model, src=0)
  flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()] + [b.reshape(-1).float() for b in model.buffers()])
  gathered = [torch.empty_like(flat) for _ in range(world)]
  dist.all_gather(gathered, flat)
  same = all(torch.equal(gathered[0], g) for g in gathered)

  def factory(b):
    def fn(m):
      return torch.randn(b, 2) + 0 * rank, b      # consumes the rank-seeded CPU generator
    return fn
  s, nfe = bdist.sharded_pc_sample(factory, model, total, seed=7)
  torch.manual_seed(bdist.rank_seed(7, rank))
  expect = torch.randn(s.shape[0], 2)
  all_s = bdist.gather_samples(s, dst=0)
  q.put((rank, same, s.shape[0], torch.equal(s, expect), None if all_s is None else all_s.clone()))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize('total', [5, 4])
def test_two_rank_broadcast_shard_gather(total):
  world = 2
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  sizes = bdist.shard_batch(total, world)
  assert [r[2] for r in res] == sizes and sum(sizes) == total
  assert all(r[1] for r in res), 'parameters differ after broadcast'
  assert all(r[3] for r in res), 'per-rank seed rule violated'
  full = res[0][4]
  assert res[1][4] is None and full.shape[0] == total
  # rank order and per-rank seeds: rank r's rows are randn under seed 7 + r
  off = 0
  for r, n in enumerate(sizes):
    torch.manual_seed(7 + r)
    assert torch.equal(full[off:off + n], torch.randn(n, 2))
    off += n


def test_shard_batch_edges():
  assert bdist.shard_batch(0, 4) == [0, 0, 0, 0]
  assert bdist.shard_batch(3, 4) == [1, 1, 1, 0]
  assert bdist.shard_batch(1024, 8) == [128] * 8
  assert bdist.rank_seed(1, 3) == 4
