"""Slab partition of a genome across ranks -- pure host logic, no CUDA.

The chromosomes, concatenated in chromsSorted order, are cut into `world`
contiguous slabs of (nearly) equal size.  A rank owns the pieces of the
chromosomes that fall inside its slab; a piece whose chromosome continues on the
neighbouring rank gets `halo` readable cells on that side, which the ranks fill
with `exchange_halos()` (NCCL send/recv through torch.distributed on GPUs, gloo on
CPU in the tests) before a windowed operator of radius <= halo runs.
"""

ALIGN = 64


def _round_up(x, a):
    return (x + a - 1) // a * a


def partition(sorted_lengths, world, rank, halo):
    """-> (segs, buffer_cells) for `rank`.

    sorted_lengths: chromosome lengths in layout (chromsSorted) order.
    segs: list of (sorted_index, lo, hi, dlo, dhi, pos0) -- one per owned piece,
    buffer cell ranges as in gdsp_seg (lo multiple of 64).
    """
    total = sum(sorted_lengths)
    cut0 = total * rank // world
    cut1 = total * (rank + 1) // world
    segs = []
    g0 = 0                       # genome coordinate of the current chromosome's first base
    pos = 0                      # next free buffer cell
    for si, length in enumerate(sorted_lengths):
        a, b = max(cut0, g0), min(cut1, g0 + length)
        if a < b:
            pos0 = a - g0
            left = halo if pos0 > 0 else 0                    # chromosome continues on rank-1
            right = halo if (pos0 + (b - a)) < length else 0  # ... on rank+1
            left = min(left, pos0)
            right = min(right, length - (pos0 + (b - a)))
            lo = _round_up(pos + left, ALIGN)
            hi = lo + (b - a)
            segs.append((si, lo, hi, lo - left, hi + right, pos0))
            pos = hi + right
        g0 += length
    buffer_cells = _round_up(pos, ALIGN) + ALIGN
    return segs, buffer_cells


def halo_plan(sorted_lengths, world, rank, halo):
    """-> list of (peer, send_lo, send_hi, recv_lo, recv_hi) in buffer cells for `rank`.

    Only the first piece can continue to the left and only the last piece to the
    right (the slab is contiguous in genome coordinates)."""
    segs, _ = partition(sorted_lengths, world, rank, halo)
    plan = []
    if not segs:
        return plan
    si, lo, hi, dlo, dhi, pos0 = segs[0]
    if dlo < lo:                     # rank-1 holds the cells just before pos0
        n = lo - dlo
        plan.append((rank - 1, lo, lo + min(n, hi - lo), dlo, lo))
    si, lo, hi, dlo, dhi, pos0 = segs[-1]
    if dhi > hi:
        n = dhi - hi
        plan.append((rank + 1, hi - min(n, hi - lo), hi, hi, dhi))
    return plan


def exchange_halos(buf, plan, dist):
    """Fill the halo cells of `buf` (1-D tensor) from the neighbouring ranks.

    Every rank sends the owned cells adjacent to a cut and receives the
    neighbour's; sends and receives are posted together (ncclGroup semantics)."""
    if not plan:
        return
    ops, recv_views = [], []
    for peer, s_lo, s_hi, r_lo, r_hi in plan:
        ops.append(dist.P2POp(dist.isend, buf[s_lo:s_hi], peer))
        ops.append(dist.P2POp(dist.irecv, buf[r_lo:r_hi], peer))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
