"""Data-parallel sampling over the GPUs of one box: prompts are independent through the whole DDIM loop
(GroupNorm / LayerNorm / attention are per-sample; the CFG pair of a prompt stays on one rank), so each
rank runs its own loop on a contiguous shard and the ONLY collective is one all-gather of the final latents
(NCCL over NVLink; gloo in the CPU tests).  Precedent in the reference: rank-strided ownership without a
gather, eval/evaluate_gen.py:55.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced: the first n % world ranks own one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_conditioning(cond: Optional[Dict[str, list]], lo: int, hi: int, memo: Optional[dict] = None):
    """Slice every tensor of a reference-style conditioning dict ({'c_crossattn': [T], 'example_pair': [T],
    'query': [T]}) along the batch dim.  ``memo`` (shared between the conditional and the unconditional dict of one
    shard) keeps tensor IDENTITY: entries that were the same object before slicing are the same object after it, which
    is what lets the sampler encode a shared example_pair / query once per CFG pair."""
    if cond is None:
        return None
    memo = {} if memo is None else memo

    def cut(t):
        key = (id(t), lo, hi)
        if key not in memo:
            memo[key] = (t, t[lo:hi])          # the source is kept alive so that its id cannot be reused
        return memo[key][1]
    return {k: [cut(t) for t in v] for k, v in cond.items()}


def all_gather_latents(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """One all-gather of the final latents; ragged shards are padded to the largest shard."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (total + world - 1) // world
    padded = local
    if local.shape[0] < per:
        padded = torch.cat([local, local.new_zeros((per - local.shape[0],) + tuple(local.shape[1:]))])
    padded = padded.contiguous()
    outs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(outs, padded, group=group)
    parts = []
    for r, o in enumerate(outs):
        lo, hi = shard_bounds(total, r, world)
        parts.append(o[: hi - lo])
    return torch.cat(parts)


def sample_sharded(sample_fn: Callable, batch_size: int, conditioning, unconditional_conditioning=None,
                   x_T: Optional[torch.Tensor] = None, group=None, chunk: Optional[int] = None):
    """Run ``sample_fn(local_batch, cond, un_cond, x_T) -> latents`` on this rank's shard (optionally in
    chunks of ``chunk`` prompts) and all-gather the results in prompt order.

    ``x_T`` (if given) is the FULL-batch noise, so every rank denoises exactly the rows the single-GPU run would.
    In fp32 mode the gathered result equals the single-GPU run bit for bit (kernels are deterministic and per-sample);
    in bf16 mode a prompt's result depends on its position inside the rank-local batch (tile boundaries change the fp32
    summation order), so sharded and unsharded runs agree to bf16 rounding noise, not bit for bit."""
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    if batch_size < world:
        # checked on EVERY rank before any work or collective: a rank without prompts must not leave the others
        # waiting inside the all-gather
        raise ValueError(f"batch {batch_size} < world size {world}: every rank needs at least one prompt")
    lo, hi = shard_bounds(batch_size, rank, world)
    outs = []
    step = (hi - lo) if not chunk else chunk
    for a in range(lo, hi, max(step, 1)):
        b = min(hi, a + step)
        memo = {}
        outs.append(sample_fn(b - a, shard_conditioning(conditioning, a, b, memo),
                              shard_conditioning(unconditional_conditioning, a, b, memo),
                              None if x_T is None else x_T[a:b]))
    return all_gather_latents(torch.cat(outs), batch_size, group)
