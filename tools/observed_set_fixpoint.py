"""Reference model of how the CUDA path solves `fast`'s order-dependent observed-voxel set in parallel,
small enough to read in one sitting and checked against the sequential definition in tests/test_fixpoint_prototype.py.

Sequential definition (fast.cpp:110-122 + ApproxHashSet::replaceHash, A.4): rays are processed in rank order; ray r walks its
voxels s = 0, 1, ...; at each step  collided = (table[slot] == value);  table[slot] = value;  the ray stops at the first step where
more than `max_collisions` collisions are consecutive, and every step before that updates its voxel.  U[r] = number of updated
voxels.  The table persists across frames.

Parallel formulation: a candidate (r, s) is PERFORMED iff s < U[r] or s is the breaking step itself (the breaking step still
executed replaceHash).  The slot always ends up holding the value of its latest visitor, so (r, s) collides iff the latest
performed candidate before it, in (rank, step) order, on the same slot carries the same value (none: the persistent table
decides).  U[r] depends only on rays of lower rank and on r's own earlier steps: a triangular system with a UNIQUE solution,
which plain Jacobi sweeps - re-evaluate every ray against the previous sweep's U - reach from ANY start in at most R + 1 sweeps
(in practice ~6-10 for a 640x480 frame, because a ray's outcome only depends on the few rays that share slots with it).
"""
from typing import Dict, List, Sequence, Tuple

MASK = (1 << 20) - 1


def sequential(rays: Sequence[Sequence[int]], table: Dict[int, int], max_collisions: int) -> Tuple[List[int], Dict[int, int]]:
    """rays[r] = values (hash + offset) of ray r's voxels in walking order. Returns (U, table after the frame)."""
    table = dict(table)
    U = []
    for vals in rays:
        run, n = 0, 0
        for v in vals:
            k = v & MASK
            if table.get(k) == v:
                run += 1
            else:
                run = 0
            table[k] = v
            if run > max_collisions:
                break
            n += 1
        U.append(n)
    return U, table


def visits(vals: Sequence[int], u: int) -> int:
    """Number of steps of a ray that execute replaceHash when it updates u voxels: the breaking step, if any, still does."""
    return min(len(vals), u + 1)


def jacobi_sweep(rays, table, max_collisions, U_prev):
    """One parallel sweep: every ray is evaluated independently against the PREVIOUS estimate of all lower-ranked rays."""
    # per slot: performed candidates of the previous estimate as (rank, step, value), in (rank, step) order
    by_slot: Dict[int, List[Tuple[int, int, int]]] = {}
    for r, vals in enumerate(rays):
        for s in range(visits(vals, U_prev[r])):
            by_slot.setdefault(vals[s] & MASK, []).append((r, s, vals[s]))
    U_new = []
    for r, vals in enumerate(rays):
        run, n = 0, 0
        own: Dict[int, int] = {}                       # slot -> value of this ray's own latest earlier step (always performed)
        for s, v in enumerate(vals):
            k = v & MASK
            if k in own:
                prev = own[k]
            else:
                prev = table.get(k)
                for (r2, s2, v2) in by_slot.get(k, ()):   # latest performed visit by a lower-ranked ray
                    if r2 >= r:
                        break
                    prev = v2
            run = run + 1 if prev == v else 0
            own[k] = v
            if run > max_collisions:
                break
            n += 1
        U_new.append(n)
    return U_new


def solve(rays, table, max_collisions, U_start, max_sweeps=None):
    """Jacobi iteration to the fixpoint; returns (U, sweeps)."""
    U = list(U_start)
    limit = max_sweeps or len(rays) + 2
    for sweep in range(1, limit + 1):
        nxt = jacobi_sweep(rays, table, max_collisions, U)
        if nxt == U:
            return U, sweep
        U = nxt
    raise RuntimeError("no fixpoint within R + 2 sweeps: the system would not be triangular")
