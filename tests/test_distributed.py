"""world_size-2 gloo test of the only multi-rank logic of the path: pair sharding + the final all-gather of poses."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffphore_b200.distributed import shard_range, all_gather_poses


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_range(5, rank, world)                         # 5 pairs over 2 ranks -> 3 + 2
    atoms = [10 + p for p in range(lo, hi)]
    pos = torch.cat([torch.full((n, 3), float(p)) for p, n in zip(range(lo, hi), atoms)])
    all_pos, all_cnt = all_gather_poses(pos, atoms)
    q.put((rank, [p.shape[0] for p in all_pos], [c.tolist() for c in all_cnt], float(all_pos[1 - rank][0, 0])))
    dist.destroy_process_group()


def test_shard_and_gather_poses_gloo():
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    for rank, sizes, counts, other_first in res:
        assert sizes == [10 + 11 + 12, 13 + 14] and counts == [[10, 11, 12], [13, 14]]
        assert other_first == (3.0 if rank == 0 else 0.0)


def _cli_worker(rank, world, port, q):
    """Rank side of `torchrun src/inference.py`: round-robin shard of the pairs, local results, one gather of the table."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'src'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import inference
    assert inference._ranks() == (rank, world, rank)
    names = [f't__lig{i}' for i in range(5)]
    mine = names[rank::world]
    local = {'name': mine, 'fitscore': [[0.1 * int(n[-1]), -2.0] for n in mine], 'run_time': [float(rank)] * len(mine)}
    q.put((rank, inference.gather_results(local, names, world)))


def test_cli_rank_sharding_and_results_gather_gloo():
    """src/inference.py under torchrun: pairs dealt round-robin, merged results in input order on every rank."""
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_cli_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    for rank, merged in res:
        assert merged['name'] == [f't__lig{i}' for i in range(5)]
        assert merged['run_time'] == [0.0, 1.0, 0.0, 1.0, 0.0]
        assert [f[0] for f in merged['fitscore']] == [0.1 * i for i in range(5)]
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'src'))
    import inference
    empty = {'name': [], 'fitscore': [], 'run_time': []}
    assert inference.merge_rank_results([empty, empty], ['a__b']) == empty          # a rank (or a job) without pairs
