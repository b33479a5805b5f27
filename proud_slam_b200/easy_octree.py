"""``grid.build_octree`` (reference ``third_party/sparse_voxels/src/octree.cpp:12-160``, NSVF's
EasyOctree): CPU builder of a dense-indexed octree over integer points.  API surface only -- the SLAM
path builds its map with ``svo.Octree`` and ``build_easy_octree`` (voxel_helpers.py:494) has no caller.

Rows: terminal nodes keep the index of their point (0..M-1); internal nodes are numbered downwards from
``total-1`` (the root) in breadth-first order.  ``children[:, 8] = 2^(depth+1)`` (1 for terminals)."""
from collections import deque

import numpy as np
import torch


class _Node:
    __slots__ = ("center", "depth", "index", "children")

    def __init__(self, center, depth, index):
        self.center, self.depth, self.index, self.children = center, depth, index, [None] * 8


def build_octree(center, points, depth):
    c0 = np.asarray(center.cpu(), dtype=np.int64).reshape(3)
    pts = np.asarray(points.cpu(), dtype=np.int64).reshape(-1, 3)
    root = _Node(c0, int(depth), -1)
    for k, pt in enumerate(pts):
        node = root
        while True:
            diff = (pt > node.center).astype(np.int64)
            idx = int(diff[0] + 2 * diff[1] + 4 * diff[2])
            if node.depth == 0:
                node.children[idx] = _Node(pt, -1, k)       # a later point in the same cell replaces the earlier one
                break
            if node.children[idx] is None:
                node.children[idx] = _Node(node.center + (2 * diff - 1) * (1 << (node.depth - 1)), node.depth - 1, -1)
            node = node.children[idx]
    total = 0
    stack = [root]
    while stack:
        n = stack.pop()
        total += 1
        stack.extend(ch for ch in n.children if ch is not None)
    centers = np.zeros((total, 3), np.int32)
    children = -np.ones((total, 9), np.int32)
    nxt = total - 1
    root.index = nxt
    queue = deque([root])
    while queue:
        n = queue.popleft()
        for i, ch in enumerate(n.children):
            if ch is not None:
                if ch.depth > -1:
                    nxt -= 1
                    ch.index = nxt
                queue.append(ch)
                children[n.index, i] = ch.index
        children[n.index, 8] = 1 << (n.depth + 1)
        centers[n.index] = n.center
    return torch.from_numpy(centers), torch.from_numpy(children)
