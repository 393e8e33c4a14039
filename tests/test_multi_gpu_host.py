"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_global, out_dir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from assistive_gym_b200.feeding_batch import FeedingBatch
    from assistive_gym_b200.sharding import all_gather_rewards, sample_block, shard_range
    fb = FeedingBatch()
    lo, hi = shard_range(rank, world, n_global)
    s = sample_block(fb, lo, hi)
    # stand-in for the per-rank reward tensor: a deterministic function of the global env id
    rew = torch.tensor(s['plane_friction'] + np.arange(lo, hi), dtype=torch.float32)
    allr = all_gather_rewards(rew)
    np.save(os.path.join(out_dir, 'rank%d.npy' % rank), allr.numpy())
    np.save(os.path.join(out_dir, 'head%d.npy' % rank), s['head_deg'])
    dist.destroy_process_group()


def test_two_rank_sharding_is_partition_invariant(tmp_path):
    n_global = 8
    mp.spawn(_worker, args=(2, 29541, n_global, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / 'rank0.npy'), np.load(tmp_path / 'rank1.npy')
    assert np.array_equal(r0, r1) and r0.shape == (n_global,)
    # single-process result for the whole batch equals the concatenation of the two shards
    sys.path.insert(0, ROOT)
    from assistive_gym_b200.feeding_batch import FeedingBatch
    from assistive_gym_b200.sharding import sample_block
    s = sample_block(FeedingBatch(), 0, n_global)
    assert np.allclose(r0, s['plane_friction'] + np.arange(n_global), atol=1e-6)
    heads = np.concatenate([np.load(tmp_path / 'head0.npy'), np.load(tmp_path / 'head1.npy')])
    assert np.array_equal(heads, s['head_deg'])


def test_shard_range():
    from assistive_gym_b200.sharding import shard_range
    assert [shard_range(r, 4, 32768) for r in range(4)] == [(0, 8192), (8192, 16384), (16384, 24576), (24576, 32768)]
