"""resample_in_space: pick rectify / affine / reproject (``spatial.py:40-168``)."""

from __future__ import annotations

from collections.abc import Iterable

from .constants import LOG
from .dataset import from_any
from .gridmapping import GridMapping
from .utils import _can_apply_affine_transform


def resample_in_space(
    source_ds,
    target_gm: GridMapping | None = None,
    source_gm: GridMapping | None = None,
    variables: str | Iterable[str] | None = None,
    interp_methods=None,
    agg_methods=None,
    recover_nans=False,
    fill_values=None,
    tile_size: int | tuple[int, int] | None = None,
):
    """Resample the spatial dimensions of *source_ds* to *target_gm*.

    Decision rules of the reference (spatial.py:121-168): irregular source ->
    rectify; no target for a regular source -> warning, source returned; grids
    close -> source returned; same CRS (or both geographic) -> affine; else
    reproject.
    """
    from .affine import affine_transform_dataset
    from .rectify import rectify_dataset
    from .reproject import reproject_dataset

    if source_gm is None:
        source_gm = GridMapping.from_dataset(from_any(source_ds))

    kwargs = dict(source_gm=source_gm, variables=variables, interp_methods=interp_methods,
                  agg_methods=agg_methods, recover_nans=recover_nans, fill_values=fill_values)
    if not source_gm.is_regular:
        return rectify_dataset(source_ds, target_gm=target_gm, tile_size=tile_size, **kwargs)
    if target_gm is None:
        LOG.warning(
            "If source grid mapping is regular `target_gm` must be given. "
            "Source dataset is returned."
        )
        return source_ds
    GridMapping.assert_regular(target_gm, name="target_gm")
    if source_gm.is_close(target_gm):
        return source_ds
    if _can_apply_affine_transform(source_gm, target_gm):
        return affine_transform_dataset(source_ds, target_gm, **kwargs)
    return reproject_dataset(source_ds, target_gm, **kwargs)
