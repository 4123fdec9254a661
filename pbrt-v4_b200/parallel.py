"""Multi-GPU film tiling: one process per GPU, disjoint scanline bands, one reduce of the film accumulators at frame end.

The sampler is a stateless function of (pixel, sample index, dimension) and the filter is importance-sampled into its own
pixel (samplers.h:252-356, film.h:239-255), so bands are independent; the only exchange is the frame-end sum of the
zero-padded per-rank accumulators {rgbSum[3], weightSum} (float64, order-independent because each pixel is non-zero on
exactly one rank).  Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""


def partition_bands(yres, world, rank, band=16, y0=0):
    """Rows [y0, yres) in blocks of `band` rows dealt round-robin to the ranks -> list of (ya, yb) for this rank."""
    out = []
    for i, y in enumerate(range(y0, yres, band)):
        if i % world == rank:
            out.append((y, min(y + band, yres)))
    return out


def reduce_film(film, dist=None, dst=0):
    """Sum the per-rank film accumulator tensors onto rank `dst` (in place).  `film`: [yres, xres, 4] float64."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return film
    dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


def band_rows(yres, world, rank, band=16, y0=0):
    """Film rows of `rank` under the round-robin band partition, as a list (the rows bpt_render_bands renders for that part)."""
    return [y for a, b in partition_bands(yres, world, rank, band, y0) for y in range(a, b)]


def gather_film(film, dist=None, band=16, dst=0, y0=0):
    """Frame-end exchange by tiles instead of a full-frame sum ("or equivalently grouped ncclSend/ncclRecv of the
    tiles"): every rank packs the rows it rendered (1/N of the film) and rank `dst` places them into its film.  Moves
    (N-1)/N x one film over NVLink instead of reducing N zero-padded full films, and is bit-identical to the reduce because every
    pixel is non-zero on exactly one rank.  `film`: [yres, xres, 4] float64, modified in place on `dst`."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return film
    world, rank = dist.get_world_size(), dist.get_rank()
    yres = film.shape[0]
    rows = [band_rows(yres, world, r, band, y0) for r in range(world)]
    nmax = max(len(r) for r in rows)
    mine = torch.zeros((nmax,) + tuple(film.shape[1:]), dtype=film.dtype, device=film.device)
    if rows[rank]:
        mine[:len(rows[rank])] = film[torch.as_tensor(rows[rank], device=film.device)]
    if rank == dst:
        bufs = [torch.empty_like(mine) for _ in range(world)]
        dist.gather(mine, bufs, dst=dst)
        for r in range(world):
            if r != dst and rows[r]:
                film[torch.as_tensor(rows[r], device=film.device)] = bufs[r][:len(rows[r])]
    else:
        dist.gather(mine, None, dst=dst)
    return film


def render_tiled(engine, spp, dist=None, band=16, first_sample=0):
    """Render this rank's bands and reduce; returns the film tensor (valid on rank 0)."""
    import torch
    x, y = engine.dims()
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    engine.film_clear()
    if world == 1:
        engine.render(first_sample, spp)
    else:
        # all of this rank's bands in one call: they share the wavefront passes (long ray queues, as on one GPU)
        engine.render_bands(first_sample, spp, band, world, rank)
    film = torch.zeros((y, x, 4), dtype=torch.float64, device="cuda")
    engine.film_to_device_ptr(film.data_ptr())
    gather_film(film, dist, band)
    if rank == 0 and world > 1:
        engine.film_from_device_ptr(film.data_ptr())
    return film
