"""Host-side logic of the multi-GPU modes on CPU: partitioning arithmetic and the rendezvous plumbing over a
world_size-2 gloo group (the data path itself -- peer stores from the stage kernels -- is covered by
tests/test_gpu_slab.py)."""
import os
import socket

import numpy as np
import pytest

from uppasd_b200 import slab


def test_ensemble_shard_covers_every_ensemble_once():
    for m in (1, 7, 8, 13):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                first, cnt = slab.ensemble_shard(m, w, r)
                seen += list(range(first, first + cnt))
            assert seen == list(range(m))
    with pytest.raises(ValueError):
        slab.ensemble_shard(8, 2, 2)


def test_slab_planes_and_ring():
    assert [slab.slab_planes(256, 8, r, 2) for r in range(8)] == [(32 * r, 32) for r in range(8)]
    with pytest.raises(ValueError):
        slab.slab_planes(10, 4, 0, 2)
    with pytest.raises(ValueError):
        slab.slab_planes(4, 4, 0, 2)        # one-plane slabs cannot serve a two-plane halo
    assert slab.ring_neighbours(8, 0) == (7, 1)
    assert slab.ring_neighbours(2, 1) == (0, 0)
    assert slab.ring_neighbours(1, 0) == (0, 0)


def test_halo_depth_of_the_bcc_table():
    import bench
    from uppasd_b200 import lattice
    B = bench.BCC
    ns, ca, cs, sh = lattice.stencil(B['cell'], B['bas'], B['atype'], np.array([4]), B['shells'][None], 1, np.ones((1, 4), dtype=int))
    assert slab.halo_depth(cs, ns) == 2     # shell (1.5,.5,.5) from the body-centre atom reaches two cell planes
    assert slab.halo_depth(cs[:, :8], np.array([8, 8])) == 1


class _FakeEngine:
    """records what connect_ring would hand to the C ABI"""

    def __init__(self, rank):
        self.rank, self.got = rank, None

    def slab_export(self):
        return bytes([self.rank]) * 192

    def slab_connect_ipc(self, lower, upper):
        self.got = (lower[0], upper[0], len(lower), len(upper))

    def slab_connect_local(self, lower, upper):
        self.got = ('local', lower is self, upper is self)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    e = _FakeEngine(rank)
    slab.connect_ring(e, world, rank, dist)
    tot = slab.allreduce_sum(np.array([[1.0 + rank, 2.0], [3.0, 4.0 * rank]]), world, dist)
    first, cnt = slab.ensemble_shard(5, world, rank)
    q.put((rank, e.got, tot.tolist(), first, cnt))
    dist.destroy_process_group()


def test_ring_rendezvous_over_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    # two slabs: each rank's lower and upper neighbour is the other rank, handle blobs arrive intact
    assert res[0][1] == (1, 1, 192, 192) and res[1][1] == (0, 0, 192, 192)
    assert res[0][2] == res[1][2] == [[3.0, 4.0], [6.0, 4.0]]
    assert (res[0][3], res[0][4], res[1][3], res[1][4]) == (0, 3, 3, 2)


def test_single_slab_connects_to_itself():
    e = _FakeEngine(0)
    slab.connect_ring(e, 1, 0)
    assert e.got == ('local', True, True)


def test_reference_arm_under_a_two_rank_launcher_uses_all_host_cores():
    """`bench.py --impl reference` launched the way the driver launches it for N > 1 (torch.distributed.run, which exports
    OMP_NUM_THREADS=1 to its workers): rank 0 alone prints ONE JSON line, timed on every host core (explicit
    omp_set_num_threads), the other rank exits 0 without work."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, OMP_NUM_THREADS='1')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', str(port), os.path.join(root, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                        '--steps', '3', '--warmup', '1', '--cpu-ncell', '16', '16', '16'],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['gpu_launches'] == 0
    assert d['cpu_baseline']['cores'] == (os.cpu_count() or 1) and d['cpu_baseline']['kind'] == 'port'
    assert d['e2e']['value'] == d['value'] > 0 and d['e2e']['h2d_bytes_per_step'] == 0
