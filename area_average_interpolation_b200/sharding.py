"""Work assignment for one-process-per-GPU runs: no collective on the data path.

* one large image  -> contiguous canvas row bands balanced by covered pixels (``aai_partition_rows``) plus the
  source halo each band needs (``aai_band_source_window``);
* a batch of images -> whole images, contiguous blocks per rank.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import ctypes as C

from . import (Image, Plan, band_source_window, covered_pixels, image_alloc, image_copy_rows, image_free, image_upload,
               ipc_close, ipc_export, ipc_open, partition_rows)


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


def band_for_rank(plan: Plan, rank: int, world_size: int) -> Band:
    bounds = partition_rows(plan, world_size)
    r0, r1 = bounds[rank], bounds[rank + 1]
    x0, x1, y0, y1 = band_source_window(plan, r0, r1)
    return Band(rank, r0, r1, x0, x1, y0, y1, covered_pixels(plan, r0, r1))


def all_bands(plan: Plan, world_size: int) -> List[Band]:
    return [band_for_rank(plan, r, world_size) for r in range(world_size)]


def batch_slice(n_images: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of whole images for ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(n_images, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class PeerSource:
    """Source distribution for a one-process-per-GPU run over ONE large image (NVLink halo option).

    With plain row bands every rank would upload its whole halo from the host -- about half of the source per rank
    for a rotated image, i.e. N/2 times the image over PCIe in total.  Here every source row crosses PCIe exactly once:
    rank r uploads rows [H r/N, H (r+1)/N) into its own full-size device image, and pulls the other rows of its halo
    from their owners' device images with peer copies over NVLink (CUDA IPC; no NCCL on the data path -- the process
    group is only used once to exchange the IPC handles and for the per-step barriers).

    Per step:  upload_owned() -> sync() -> pull_halo() -> [kernels, download] -> sync().
    """

    def __init__(self, plan: Plan, dtype: int, channels: int, rank: int, world: int, device: int, band: "Band",
                 all_gather_object):
        self.plan, self.rank, self.world, self.device, self.band = plan, rank, world, device, band
        h = plan.src_h
        self.owned = [(h * p // world, h * (p + 1) // world) for p in range(world)]
        self.full = image_alloc(device, plan.src_w, h, dtype, channels)
        handles = all_gather_object(ipc_export(self.full.data))
        self._opened = []
        self.peers = []
        for p in range(world):
            if p == rank:
                self.peers.append(self.full)
                continue
            ptr = ipc_open(handles[p], device)
            self._opened.append(ptr)
            self.peers.append(Image(ptr, self.full.pitch_bytes, self.full.width, self.full.height, 0, h, dtype, channels))

    def owned_rows(self):
        return self.owned[self.rank]

    def upload_owned(self, host_img: Image, stream: int = 0) -> None:
        """host_img holds (at least) this rank's owned rows."""
        o0, o1 = self.owned[self.rank]
        view = Image(self.full.data + o0 * self.full.pitch_bytes, self.full.pitch_bytes, self.full.width, self.full.height,
                     o0, o1 - o0, self.full.dtype, self.full.channels)
        image_upload(view, host_img, self.device, stream)

    def pull_halo(self, stream: int = 0) -> int:
        """Peer-copies the rows of this rank's halo that other ranks own; returns the bytes moved over NVLink."""
        moved = 0
        for p in range(self.world):
            if p == self.rank:
                continue
            lo, hi = max(self.band.src_y0, self.owned[p][0]), min(self.band.src_y1, self.owned[p][1])
            if hi > lo:
                image_copy_rows(self.full, self.peers[p], lo, hi, self.device, stream)
                moved += (hi - lo) * self.full.pitch_bytes
        return moved

    def close(self) -> None:
        for ptr in self._opened:
            ipc_close(ptr, self.device)
        self._opened = []
        image_free(self.full, self.device)
