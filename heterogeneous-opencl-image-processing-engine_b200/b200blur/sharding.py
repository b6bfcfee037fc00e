"""Host-side work distribution (layer L4 of SURVEY.md): whole-image shards (Approach 1) and row bands with halo
(Approach 2), generalised from the reference's two-device CPU/GPU ratio split to G equal GPUs.

Pure integer arithmetic; the same numbers come out of the C ABI (b200blur_partition), this module just gives them
names.  Reference: heterogeneous_blur.c:446-458 (image split), split_image_blur.c:144-166, :511-517 (row split and
halo pointers).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple


def even_split(n_items: int, n_parts: int, part: int) -> Tuple[int, int]:
    """(begin, count) of `part`; the first n_items % n_parts parts get one extra item.  Same as b200blur_partition."""
    if n_items < 0 or n_parts < 1 or not (0 <= part < n_parts):
        raise ValueError("bad partition request")
    q, r = divmod(n_items, n_parts)
    return part * q + min(part, r), q + (1 if part < r else 0)


def image_shard(n_images: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Approach 1 on G GPUs: GPU `rank` blurs images [begin, begin+count) of the stream."""
    return even_split(n_images, world_size, rank)


def band_rows(height: int, n_bands: int, band: int) -> Tuple[int, int]:
    """Approach 2 on G GPUs: band `band` owns rows [row0, row0+rows) of every image."""
    return even_split(height, n_bands, band)


@dataclass(frozen=True)
class BandPlan:
    band: int
    row0: int          # first owned row
    rows: int          # owned (= output) rows
    has_top: bool      # a neighbour above supplies one halo row (row0-1)
    has_bottom: bool   # a neighbour below supplies one halo row (row0+rows)

    @property
    def input_rows(self) -> int:
        """Rows the band's kernel sees -- the `height` argument of split_image_blur.c:401/:414."""
        return self.rows + int(self.has_top) + int(self.has_bottom)


def plan_bands(height: int, n_bands: int) -> List[BandPlan]:
    """All bands of an image.  Empty bands (n_bands > height) are dropped, so halos always come from a real row."""
    plans = []
    for k in range(n_bands):
        r0, n = band_rows(height, n_bands, k)
        if n == 0:
            continue
        plans.append((k, r0, n))
    out = []
    for i, (k, r0, n) in enumerate(plans):
        out.append(BandPlan(band=k, row0=r0, rows=n, has_top=i > 0, has_bottom=i < len(plans) - 1))
    return out
