"""Multi-GPU plumbing: the path shards embarrassingly (every (pair, sample) trajectory is independent),
so ranks own contiguous blocks of pairs and the ONLY collective is one all-gather of the final poses
(NCCL over NVLink on GPUs; gloo in the CPU tests).  The reference has no distributed code at all."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank`; blocks differ by at most one item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_poses(pos, atoms_per_graph, group=None):
    """pos [n_atoms_local, 3] + atoms_per_graph [n_graphs_local] -> (list of per-rank pos tensors, list of per-rank
    atom-count tensors) on every rank.  Ragged sizes are exchanged first, payloads are padded to the largest rank."""
    world = dist.get_world_size(group)
    dev = pos.device
    counts = torch.as_tensor(atoms_per_graph, dtype=torch.int64, device=dev)
    sizes = torch.tensor([pos.shape[0], counts.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    max_atoms = int(max(s[0] for s in all_sizes))
    max_graphs = int(max(s[1] for s in all_sizes))
    pad_pos = torch.zeros(max_atoms, 3, dtype=pos.dtype, device=dev)
    pad_pos[:pos.shape[0]] = pos
    pad_cnt = torch.zeros(max_graphs, dtype=torch.int64, device=dev)
    pad_cnt[:counts.shape[0]] = counts
    out_pos = [torch.empty_like(pad_pos) for _ in range(world)]
    out_cnt = [torch.empty_like(pad_cnt) for _ in range(world)]
    dist.all_gather(out_pos, pad_pos, group=group)
    dist.all_gather(out_cnt, pad_cnt, group=group)
    return ([p[:int(s[0])] for p, s in zip(out_pos, all_sizes)], [c[:int(s[1])] for c, s in zip(out_cnt, all_sizes)])
