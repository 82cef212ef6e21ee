"""Particle sharding over the GPUs of one box and the end-of-run reductions.

The tracking path has no collective step: particles are independent, each
rank (one process per GPU) tracks a contiguous block of `particle_id`s against its own copy
of the lowered lattice.  The only communication is the final reduction of the per-GPU
partial statistics (`xtb_reduce_stats`: N_alive, N_lost, first and second moments) and of the
loss histogram by element (`xtb_loss_histogram`) -- a few hundred doubles / int64 over NCCL.
The reference has no multi-GPU path; this replaces "run N independent jobs and merge".
"""
import numpy as np
import torch
import torch.distributed as dist

N_STATS = 29      # n_alive, n_lost, sum[6], sum2[21]  (xtb_stats_t, include/xtb200.h)
COORDS = ('x', 'px', 'y', 'py', 'zeta', 'delta')


def shard_range(n_total, rank, world_size):
    """[start, stop) of the particle_id block of `rank`: contiguous blocks of ceil(n/world)."""
    per = -(-int(n_total) // int(world_size))
    start = min(rank * per, n_total)
    return start, min(start + per, n_total)


def all_reduce_stats(partial):
    """Sum of the per-rank partials (tensor of N_STATS float64 on the rank's device)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM)
    return partial


def all_reduce_histogram(hist):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def beam_statistics(stats):
    """Means, covariance matrix and counts from the (all-reduced) partial sums."""
    ss = stats.detach().cpu().numpy() if isinstance(stats, torch.Tensor) else np.asarray(stats)
    n_alive, n_lost = int(round(ss[0])), int(round(ss[1]))
    out = {'n_alive': n_alive, 'n_lost': n_lost}
    if n_alive == 0:
        return out
    mean = ss[2:8] / n_alive
    cov = np.zeros((6, 6))
    kk = 8
    for ii in range(6):
        for jj in range(ii, 6):
            cov[ii, jj] = cov[jj, ii] = ss[kk] / n_alive - mean[ii] * mean[jj]
            kk += 1
    out['mean'] = dict(zip(COORDS, mean))
    out['cov'] = cov
    out['sigma'] = dict(zip(COORDS, np.sqrt(np.clip(np.diag(cov), 0, None))))
    return out


def partial_stats_host(particles):
    """The partial sums of `xtb_reduce_stats` computed on the host from a Particles object
    (used where the particles are already on the host, e.g. after an e2e round trip)."""
    st = particles.get('state')
    alive = st > 0
    out = np.zeros(N_STATS)
    out[0] = alive.sum()
    out[1] = ((st <= 0) & (st > -999999999)).sum()
    vv = [particles.get(cc)[alive] for cc in COORDS]
    out[2:8] = [v.sum() for v in vv]
    kk = 8
    for ii in range(6):
        for jj in range(ii, 6):
            out[kk] = (vv[ii] * vv[jj]).sum()
            kk += 1
    return out


def lost_particles(particles):
    """The lost particles of `particles` (state <= 0, allocated slots) as a host `Particles`,
    ordered by particle_id: where (`at_element`, `s` -- refined or not), when (`at_turn`) and
    with which coordinates each of them stopped.  `to_dict()` / `to_pandas()` give the record."""
    from .particles import LAST_INVALID_STATE
    host = particles.copy(_device='cpu')
    st = host.get('state')
    keep = (st <= 0) & (st > LAST_INVALID_STATE)
    out = host.filter(keep)
    out.sort(by='particle_id', interleave_lost_particles=True)
    return out


def gather_lost_particles(particles, dst=0):
    """The lost particles of ALL ranks on rank `dst` (a host `Particles` ordered by particle_id;
    None on the other ranks): the optional end-of-run gather of SURVEY.md section 8(e).  Only the lost
    ones travel -- a few KB for a dynamic-aperture run."""
    from .particles import Particles
    mine = lost_particles(particles)
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return mine
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, mine.to_dict())
    if dist.get_rank() != dst:
        return None
    merged = Particles.merge([Particles.from_dict(dd) for dd in parts if len(dd['state']) > 0]
                             or [mine])
    merged.sort(by='particle_id', interleave_lost_particles=True)
    return merged
