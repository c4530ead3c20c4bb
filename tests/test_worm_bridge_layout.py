"""Host-side check of the node numbering worm_bridge_fill (pimc_worm.cuh) relies on: the recursion of sample_middle
(mc_qworm.cc:240-287: midpoint rint((it0 + it2)/2), left half first) visits every interior point of a bridge exactly once,
and in its own pre-order the left child of node k is node k + 1 and the right child node k + (it1 - it0).  The device lets
every interior point find its node by walking down from the root with that rule; here the walk is compared with the
recursion for every bridge length the device supports and every alignment of the left end."""


def recursion(it0, it2, depth=0, out=None):
    if out is None:
        out = []
    if it2 - it0 < 2:
        return out
    it1 = int(round(0.5 * (it0 + it2)))          # Python rounds halves to even, like rint
    out.append((it0, it1, it2, depth))
    recursion(it0, it1, depth + 1, out)
    recursion(it1, it2, depth + 1, out)
    return out


def walk(it0r, it2r, j):
    a, b, k, dep = it0r, it2r, 0, 0
    target = it0r + j
    while True:
        mid = int(round(0.5 * (a + b)))
        if mid == target:
            return k, (a, mid, b, dep)
        if target < mid:
            b = mid; k += 1
        else:
            k += mid - a; a = mid
        dep += 1


def test_every_interior_point_finds_its_preorder_node():
    for L in range(2, 65):                       # WORM_MAXM = 64
        for it0 in range(0, 9):
            nodes = recursion(it0, it0 + L)
            assert len(nodes) == L - 1
            assert sorted(n[1] for n in nodes) == list(range(it0 + 1, it0 + L))
            seen = set()
            for j in range(1, L):
                k, node = walk(it0, it0 + L, j)
                assert nodes[k] == node, (L, it0, j)
                seen.add(k)
            assert seen == set(range(L - 1))
            # a level only depends on shallower ones: both ends of a node are bridge ends or midpoints of smaller depth
            depth_of = {n[1]: n[3] for n in nodes}
            for a, m, b, d in nodes:
                for e in (a, b):
                    assert e in (it0, it0 + L) or depth_of[e] < d
