"""Host-side box arithmetic shared by the benchmark drivers (pure Python).

``chop_1d`` restates how AMReX cuts one direction of the level-0 domain into
grids (AmrMesh::MakeBaseGrids + BoxList::maxSize, SURVEY.md appendix C
[AMReX, unverified]): coarsen by 2 when the extent is even, split into
ceil(len/chunk) nearly equal pieces after stripping common factors of two,
larger pieces first counted from the HIGH end, refine back.
"""


def max_size_pieces(length, chunk):
    """Sizes of the pieces BoxList::maxSize produces for one direction, low to high."""
    if length <= chunk:
        return [length]
    ratio, bs, nlen = 1, chunk, length
    while bs % 2 == 0 and nlen % 2 == 0:
        ratio *= 2
        bs //= 2
        nlen //= 2
    numblk = nlen // bs + (1 if nlen % bs else 0)
    sz, extra = nlen // numblk, nlen % numblk
    # pieces are chopped off the high end: k = 0 is the top-most piece
    from_high = [((sz + 1) if k < extra else sz) * ratio for k in range(numblk - 1)]
    low = length - sum(from_high)
    return [low] + from_high[::-1]


def chop_1d(n, max_grid=32):
    """Piece boundaries [0, e1, ..., n] of one direction of the base grids."""
    fac = 2 if n % 2 == 0 else 1
    sizes = [s * fac for s in max_size_pieces(n // fac, max(max_grid // fac, 1))]
    edges = [0]
    for s in sizes:
        edges.append(edges[-1] + s)
    assert edges[-1] == n
    return edges


def slab_partition(nz, nranks):
    """Contiguous z-slabs, one per rank: [(k_lo, k_hi_inclusive)]; remainder to low ranks."""
    base, extra = divmod(nz, nranks)
    out, k = [], 0
    for r in range(nranks):
        h = base + (1 if r < extra else 0)
        out.append((k, k + h - 1))
        k += h
    return out
