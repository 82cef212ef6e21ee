"""Multi-rank path on CPU (gloo, world_size 2): particle sharding + end-of-run reductions.

Each rank tracks its shard of the beam (host build of the device code, tests/hostsim) and the
partial statistics / loss histograms are all-reduced exactly as bench.py does over NCCL; the
result must equal the single-process run of the whole beam."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, n_total, turns, out_dir):
    for pp in (os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), 'oracle'), HERE):
        if pp not in sys.path:
            sys.path.insert(0, pp)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import common
    import hostsim
    from xtrack_b200 import sharding
    line = common.load_line('sps')
    full = common.gaussian_particles(line, n_total, 41, common.SIGMAS['sps'], scale=6.0)
    lo, hi = sharding.shard_range(n_total, rank, world)
    import xtrack_b200 as xb
    ref = line.particle_ref
    kw = {cc: full.get(cc)[lo:hi] for cc in sharding.COORDS}
    p = xb.Particles(p0c=float(ref.get('p0c')[0]), mass0=ref.mass0, q0=ref.q0,
                     particle_id=np.arange(lo, hi), **kw)
    hostsim.build_hostsim_tracker(line)
    line.track(p, num_turns=turns)
    stats = torch.from_numpy(sharding.partial_stats_host(p))
    sharding.all_reduce_stats(stats)
    lost = p.get('state') <= 0
    hist = torch.from_numpy(np.bincount(p.get('at_element')[lost], minlength=len(line) + 1))
    sharding.all_reduce_histogram(hist)
    lost_all = sharding.gather_lost_particles(p, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, 'stats.npy'), stats.numpy())
        np.save(os.path.join(out_dir, 'hist.npy'), hist.numpy())
        np.savez(os.path.join(out_dir, 'lost.npz'), **{nn: lost_all.get(nn) for nn in
                 ('particle_id', 'state', 'at_element', 'at_turn', 'x', 'y', 's')})
    else:
        assert lost_all is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path):
    sys.path.insert(0, HERE)
    import common
    import hostsim
    from xtrack_b200 import sharding
    n_total, turns, world = 1001, 5, 2
    assert sharding.shard_range(n_total, 0, world) == (0, 501)
    assert sharding.shard_range(n_total, 1, world) == (501, 1001)
    assert sharding.shard_range(3, 3, 4) == (3, 3)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_total, turns, str(tmp_path)), nprocs=world, join=True)
    stats = np.load(tmp_path / 'stats.npy')
    hist = np.load(tmp_path / 'hist.npy')

    line = common.load_line('sps')
    full = common.gaussian_particles(line, n_total, 41, common.SIGMAS['sps'], scale=6.0)
    hostsim.build_hostsim_tracker(line)
    line.track(full, num_turns=turns)
    ref = sharding.partial_stats_host(full)
    assert stats[0] == ref[0] and stats[1] == ref[1]
    assert ref[1] > 20                       # the wide beam loses particles on the apertures
    np.testing.assert_allclose(stats[2:], ref[2:], rtol=1e-12, atol=1e-18)
    lost = full.get('state') <= 0
    assert np.array_equal(hist, np.bincount(full.get('at_element')[lost], minlength=len(line) + 1))
    # the lost-particle records gathered on rank 0 are those of the single run
    got = np.load(tmp_path / 'lost.npz')
    one = sharding.lost_particles(full)
    assert len(got['particle_id']) == int(ref[1]) and np.all(np.diff(got['particle_id']) > 0)
    for nn in ('particle_id', 'state', 'at_element', 'at_turn', 'x', 'y', 's'):
        assert np.array_equal(got[nn], one.get(nn)), nn
    bs = sharding.beam_statistics(stats)
    assert bs['n_alive'] + bs['n_lost'] == n_total
    assert 0 < bs['sigma']['x'] < 2e-2
