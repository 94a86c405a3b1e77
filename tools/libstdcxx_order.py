"""Closed form of libstdc++'s std::unordered_map iteration order, as a data-parallel algorithm (prototype for the device
kernel that will let `merged` reproduce the reference's bundle order, merged.cpp:210-231, instead of the canonical
first-insertion order).  Checked against the real container in tests/test_unordered_map_order.py.

libstdc++'s _Hashtable keeps ONE singly linked list of all nodes; a bucket stores the node *before* its first node.
  * insert into a non-empty bucket  -> the node becomes the FIRST of that bucket's run (runs are LIFO);
  * insert into an empty bucket     -> the node goes to the FRONT of the whole list (a new run, in front of all others);
  * rehash                          -> walk the list front to back and re-insert every node by the two rules above.
So, between two rehashes, the list is: runs ordered by creation time, newest first; inside a run, newest first.  A whole phase
(rehash of what is there + all insertions until the next rehash) is therefore ONE sort of the nodes by
      (creation time of the node's bucket, descending ; arrival time of the node inside the phase, descending)
where for the rehash "time" is the position in the old list.  The sizes at which libstdc++ rehashes (13, 29, 59, 127, ...:
_Prime_rehash_policy, next prime >= 2x) are not re-derived here: they are read from the container of the platform's own
libstdc++ (bucket_count() after every insertion), which is also what defines the reference's behaviour on that platform.

Work: every phase sorts all nodes present at its end; bucket counts at least double, so the total is < 3 N keys for N
insertions - a handful of radix sorts over geometrically growing prefixes on the device.
"""
import numpy as np


def iteration_order(hashes: np.ndarray, bucket_count_after_insert: np.ndarray) -> np.ndarray:
    """hashes[i]: hash code (size_t) of the i-th DISTINCT key in insertion order; bucket_count_after_insert[i]: the
    container's bucket_count() right after inserting key i.  Returns the node indices in iteration order."""
    h = np.asarray(hashes, dtype=np.uint64)
    bc = np.asarray(bucket_count_after_insert, dtype=np.uint64)
    n = len(h)
    order = np.zeros(0, dtype=np.int64)        # current list, front first
    start = 0
    while start < n:
        B = bc[start]
        end = start
        while end < n and bc[end] == B:
            end += 1
        old = len(order)
        nodes = np.concatenate([order, np.arange(start, end, dtype=np.int64)])
        arrival = np.arange(len(nodes), dtype=np.int64)            # old list position, then insertion time: "later = bigger"
        bucket = (h[nodes] % B).astype(np.int64)
        # creation time of a bucket's run in this phase = arrival of its first node
        first = np.full(int(B), np.iinfo(np.int64).max, dtype=np.int64)
        np.minimum.at(first, bucket, arrival)
        ctime = first[bucket]
        # rehash members keep walking order semantics: a node met later in the walk is put in front of its run -> arrival desc
        perm = np.lexsort((-arrival, -ctime))                      # primary: ctime desc, secondary: arrival desc
        order = nodes[perm]
        del old
        start = end
    return order


def phases(bucket_count_after_insert):
    """[(first insert index, one-past-last, bucket count)] of the phases, for sizing device scratch."""
    bc = np.asarray(bucket_count_after_insert)
    cuts = np.flatnonzero(np.diff(bc)) + 1
    starts = np.concatenate([[0], cuts])
    ends = np.concatenate([cuts, [len(bc)]])
    return [(int(s), int(e), int(bc[s])) for s, e in zip(starts, ends)]
