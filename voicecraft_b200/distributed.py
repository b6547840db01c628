"""Batch sharding by utterance across the GPUs of one node.

The decode path has no cross-utterance dependency, so the multi-GPU story is: one process per GPU
(torchrun), every rank holds a full weight replica and its own KV pool, utterances are partitioned across
ranks, and there is NO collective inside the decode loop.  The only exchange is the final gather of the
variable-length token matrices (padded all_gather over NCCL / NVLink; gloo in the CPU tests).
The reference has no inference-side equivalent (single device, inference_tts_scale.py:42-105).
"""
from typing import List, Sequence

import torch
import torch.distributed as dist


def partition(lengths: Sequence[int], world_size: int, rank: int) -> List[int]:
    """Indices of the utterances this rank decodes: longest-processing-time greedy over expected lengths, so the
    slowest rank (which sets the wall time) is as short as possible.  Deterministic on every rank."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    owner = [0] * len(lengths)
    for i in order:
        r = min(range(world_size), key=lambda q: (loads[q], q))
        owner[i] = r
        loads[r] += int(lengths[i])
    return [i for i in range(len(lengths)) if owner[i] == rank]


def _exchange_device(local, group, device):
    """Device the collective runs on: NCCL needs CUDA tensors on EVERY rank, including one that decoded nothing
    (partition() leaves ranks empty when there are fewer utterances than ranks)."""
    if device is not None:
        return torch.device(device)
    if dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return local[0].device if local else torch.device("cpu")


last_gather_bytes = 0       # bytes this rank contributed to the last gather_token_lists (bench.py reports it)


def gather_token_lists(local: List[torch.Tensor], local_ids: List[int], n_total: int, group=None, device=None) -> List[torch.Tensor]:
    """All ranks end up with the full list (indexed by global utterance id) of [K, T_i] int64 token matrices.
    Works on CPU tensors with gloo and CUDA tensors with NCCL; a rank may hold no utterance at all."""
    global last_gather_bytes
    last_gather_bytes = 0
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = [None] * n_total
        for i, t in zip(local_ids, local):
            out[i] = t
        return out
    world = dist.get_world_size(group)
    dev = _exchange_device(local, group, device)
    local = [t.to(dev) for t in local]
    K = local[0].shape[0] if local else 0
    meta = torch.tensor([len(local), K, max((t.shape[1] for t in local), default=0)], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    n_max = int(max(m[0] for m in metas))
    K = int(max(m[1] for m in metas))
    t_max = int(max(m[2] for m in metas))
    pack = torch.full((n_max, K, t_max), -1, dtype=torch.int64, device=dev)
    info = torch.full((n_max, 2), -1, dtype=torch.int64, device=dev)          # (global id, length)
    for j, (i, t) in enumerate(zip(local_ids, local)):
        pack[j, :, : t.shape[1]] = t
        info[j, 0], info[j, 1] = i, t.shape[1]
    packs = [torch.empty_like(pack) for _ in range(world)]
    infos = [torch.empty_like(info) for _ in range(world)]
    dist.all_gather(packs, pack, group=group)
    dist.all_gather(infos, info, group=group)
    last_gather_bytes = (pack.numel() + info.numel() + meta.numel()) * 8
    out = [None] * n_total
    for p, inf in zip(packs, infos):
        for j in range(n_max):
            gid, ln = int(inf[j, 0]), int(inf[j, 1])
            if gid >= 0:
                out[gid] = p[j, :, :ln].clone()
    return out
