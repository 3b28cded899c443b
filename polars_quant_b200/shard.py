"""Symbol sharding for the multi-GPU driver: symbols are independent on the indicator
path, so G GPUs split a panel by contiguous symbol ranges and never exchange data.  Ranges are
made of whole 32-symbol blocks (the unit of the tiled device layout and of one CTA), so no block
straddles two GPUs and every rank but the last gets the same number of blocks +-1."""
from __future__ import annotations

BLOCK = 32   # symbols per device block (suite_kernel.cuh SYM)


def symbol_range(n_symbols: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of the symbols owned by `rank` out of `world` GPUs."""
    if n_symbols < 0 or world <= 0 or not 0 <= rank < world:
        raise ValueError("bad shard request n_symbols=%r world=%r rank=%r" % (n_symbols, world, rank))
    n_blocks = -(-n_symbols // BLOCK)
    base, extra = divmod(n_blocks, world)
    b_lo = rank * base + min(rank, extra)
    b_hi = b_lo + base + (1 if rank < extra else 0)
    return min(b_lo * BLOCK, n_symbols), min(b_hi * BLOCK, n_symbols)


def all_ranges(n_symbols: int, world: int) -> list[tuple[int, int]]:
    return [symbol_range(n_symbols, world, r) for r in range(world)]


def shard_columns(names: list[str], world: int, rank: int) -> list[str]:
    """The `{symbol}` names (sorted order = panel order) a rank owns."""
    lo, hi = symbol_range(len(names), world, rank)
    return names[lo:hi]


def run_sharded(close, high, low, volume, params=None, starts=None, dist=None, device=None):
    """Multi-GPU driver: each rank (one process per GPU) runs the fused suite on its own symbol range
    of the row-major host panel [n_symbols, n_bars] and returns (lo, hi, outputs) for that range.
    `dist` = an initialised torch.distributed module (or None for a single process); it is used only
    to learn rank/world -- there is no collective on the data path."""
    import os
    from . import _native as N
    from .panel import Panel, get_engine
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    n_symbols, n_bars = close.shape
    lo, hi = symbol_range(n_symbols, world, rank)
    if hi == lo:
        return lo, hi, {}
    if device is None:
        # one process per GPU: the rank's own device (torchrun's LOCAL_RANK), never everyone on GPU 0
        n_dev = max(1, N.lib().pqb_device_count())
        device = int(os.environ.get("LOCAL_RANK", rank)) % n_dev
    eng = get_engine(device)
    p = Panel(hi - lo, n_bars, engine=eng)
    p.set_fields(close[lo:hi], None if high is None else high[lo:hi], None if low is None else low[lo:hi],
                 None if volume is None else volume[lo:hi], starts=None if starts is None else starts[lo:hi])
    p.run_host(params)
    out = {k: (v.copy(), ok.copy()) for k, (v, ok) in p.outputs().items()}
    p.close()
    return lo, hi, out
