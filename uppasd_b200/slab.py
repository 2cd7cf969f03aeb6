"""Host-side plumbing of the multi-GPU modes (one process per GPU, torch.distributed for the rendezvous only).

The reference is single-device (no MPI / NCCL anywhere); both modes are introduced by this build:

* ensemble sharding  -- `ensemble_shard`: which of the Mensemble ensembles a rank owns; no data-path communication;
* slab decomposition -- `slab_planes`, `ring_neighbours`, `connect_ring`: a supercell cut into z-slabs, boundary
  spins stored straight into the ring neighbours' halo slots by the stage kernels (NVLink peer memory).  The only
  host-side exchange is the one-off all-gather of the CUDA IPC handles, done with whatever process group the
  caller has (NCCL on the GPU box, gloo in the CPU tests), and small reductions of per-slab observables.
"""
import numpy as np


def ensemble_shard(mensemble, world, rank):
    """(first, count): contiguous block of ensembles of `rank`; remainders go to the lowest ranks."""
    if not (0 <= rank < world):
        raise ValueError('rank %d outside 0..%d' % (rank, world - 1))
    base, rem = divmod(mensemble, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def slab_planes(n3, world, rank, halo):
    """(z0, nz): the planes of slab `rank`.  N3 must split evenly and every slab must be at least `halo` thick
    (a boundary spin is then needed by the adjacent slab only)."""
    if n3 % world != 0:
        raise ValueError('N3 = %d is not a multiple of the %d slabs' % (n3, world))
    nz = n3 // world
    if nz < halo:
        raise ValueError('slab thickness %d is smaller than the halo depth %d' % (nz, halo))
    return rank * nz, nz


def halo_depth(cell_shift, nslot):
    """interaction range along z in cell planes = max |dz| over the stencil entries in use (neighbourmap.f90:187-201)."""
    cs = np.asarray(cell_shift)
    h = 0
    for i0, n in enumerate(np.asarray(nslot)):
        if n > 0:
            h = max(h, int(np.abs(cs[i0, :n, 2]).max()))
    return max(h, 1)


def ring_neighbours(world, rank):
    """(lower, upper) ranks along z with periodic wrap."""
    return (rank - 1) % world, (rank + 1) % world


def exchange_handles(handle, world, rank, dist=None, group=None):
    """all-gathers the per-rank IPC handle blobs; returns the list indexed by rank."""
    if world == 1:
        return [handle]
    t = dist.get_backend(group)
    import torch
    dev = torch.device('cuda', torch.cuda.current_device()) if t == 'nccl' else torch.device('cpu')
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [bytes(x.cpu().tolist()) for x in out]


def connect_ring(engine, world, rank, dist=None, group=None):
    """after engine.commit(): maps the ring neighbours' buffers into this process."""
    if world == 1:
        engine.slab_connect_local(engine, engine)
        return
    handles = exchange_handles(engine.slab_export(), world, rank, dist, group)
    lo, hi = ring_neighbours(world, rank)
    engine.slab_connect_ipc(handles[lo], handles[hi])
    dist.barrier(group=group)


def allreduce_sum(x, world, dist=None, group=None):
    """sum of a small host array over the slabs (per-ensemble sum M, energies)."""
    if world == 1:
        return np.asarray(x, dtype=np.float64)
    import torch
    t = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if t == 'nccl' else torch.device('cpu')
    v = torch.tensor(np.asarray(x, dtype=np.float64), device=dev)
    dist.all_reduce(v, group=group)
    return v.cpu().numpy()
