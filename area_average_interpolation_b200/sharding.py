"""Work assignment for one-process-per-GPU runs: no collective on the data path.

* one large image  -> contiguous canvas row bands balanced by covered pixels (``aai_partition_rows``) plus the
  source halo each band needs (``aai_band_source_window``);
* a batch of images -> whole images, contiguous blocks per rank.

The end-to-end distribution of ONE large host image over the ranks (each source row uploaded once, halos over NVLink) is
the peer group of the C ABI (``aai_peer_*`` in include/aai.h, ``PeerGroup`` in this package).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

from . import Plan, band_source_window, covered_pixels, partition_rows


@dataclass(frozen=True)
class Band:
    rank: int
    row0: int
    row1: int
    src_x0: int
    src_x1: int
    src_y0: int
    src_y1: int
    covered: int

    @property
    def rows(self) -> int:
        return self.row1 - self.row0


def band_for_rank(plan: Plan, rank: int, world_size: int, empty_weight: Optional[float] = None) -> Band:
    """``empty_weight``: cost of an empty canvas pixel relative to a covered one (``band_empty_weight(plan, mode, arith)``
    for the kernel that will run; default: the FP32 overlap kernel's)."""
    bounds = partition_rows(plan, world_size, empty_weight)
    r0, r1 = bounds[rank], bounds[rank + 1]
    x0, x1, y0, y1 = band_source_window(plan, r0, r1)
    return Band(rank, r0, r1, x0, x1, y0, y1, covered_pixels(plan, r0, r1))


def all_bands(plan: Plan, world_size: int, empty_weight: Optional[float] = None) -> List[Band]:
    return [band_for_rank(plan, r, world_size, empty_weight) for r in range(world_size)]


def batch_slice(n_images: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of whole images for ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(n_images, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)
