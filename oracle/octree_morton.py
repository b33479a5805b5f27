"""oracle/octree_morton.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A second, independent restatement of the reference's CPU octree (``third_party/sparse_octree/src/octree.cpp``), written
straight from its Morton-key code path instead of from coordinates with cleared low bits (which is what both
``oracle/octree_oracle.c`` and the product's ``csrc/octree_host.cpp`` do).  The reference's ``svo`` extension cannot be built in
this image (PCL / Eigen / glog / TBB headers); this file follows its arithmetic line by line so that the two
coordinate-based implementations are pinned against the key-based formulation:

  expand / compact / compute_morton / encode / decode   utils.h:79-124 (21 bits per axis, MASK[20] applied by ``encode``)
  MASK                                                  utils.h:45-75 (MASK[0] = 0x7000..., MASK[i] = MASK[i-1] | MASK[0] >> 3i)
  Octant.index_ = creation order                        octree.h:41, 111 (static counter)
  insert                                                octree.cpp:104-294: 8 corners (incr_x/y/z :12-14), key = encode(x, y, z),
                                                        all_keys, level loop with childid from the coordinate bits, a new
                                                        octant's code = key & MASK[d + shift], shift = MAX_BITS - max_level - 1,
                                                        leaf type SURFACE for corner 0 else FEATURE, promotion on a later corner 0
  try_insert                                            octree.cpp:385-416 (overlap ratio of corner keys with all_keys)
  find_octant / has_voxel                               octree.cpp:419-473
  get_leaf_voxels                                       octree.cpp:480-511 (depth-first, children 0..7, decode(code))
  count_nodes / count_leaf_nodes                        octree.cpp:535-556, 712-735
  get_centres_and_children                              octree.cpp:561-687 (BFS over non-FEATURE octants, rows = index_)

The per-octant point-cloud payload (subspaces, Hilbert colour codes) feeds only a branch that ``render_rays`` has commented
out (render_helpers.py:481) and is left out.  Pure Python: use it on small scenes.
"""
import numpy as np

MAX_BITS = 21
_M64 = (1 << 64) - 1
MASK = [0x7000000000000000]
for _i in range(1, 21):
    MASK.append(MASK[-1] | (MASK[0] >> (_i * 3)))
INCR_X = [0, 0, 0, 0, 1, 1, 1, 1]
INCR_Y = [0, 0, 1, 1, 0, 0, 1, 1]
INCR_Z = [0, 1, 0, 1, 0, 1, 0, 1]
NONLEAF, SURFACE, FEATURE = -1, 0, 1


def expand(value):
    x = value & 0x1FFFFF
    x = (x | x << 32) & 0x1F00000000FFFF
    x = (x | x << 16) & 0x1F0000FF0000FF
    x = (x | x << 8) & 0x100F00F00F00F00F
    x = (x | x << 4) & 0x10C30C30C30C30C3
    x = (x | x << 2) & 0x1249249249249249
    return x & _M64


def compact(value):
    x = value & 0x1249249249249249
    x = (x | x >> 2) & 0x10C30C30C30C30C3
    x = (x | x >> 4) & 0x100F00F00F00F00F
    x = (x | x >> 8) & 0x1F0000FF0000FF
    x = (x | x >> 16) & 0x1F00000000FFFF
    x = (x | x >> 32) & 0x1FFFFF
    return x


def encode(x, y, z):
    return (expand(x) | (expand(y) << 1) | (expand(z) << 2)) & MASK[MAX_BITS - 1]


def decode(code):
    return compact(code >> 0), compact(code >> 1), compact(code >> 2)


class Octant:
    __slots__ = ("code", "side", "index", "is_leaf", "type", "child")


class Octree:
    def __init__(self, grid_dim):
        self.next_index = 0                       # Octant::next_index_
        self.size = int(grid_dim)
        self.max_level = int(np.log2(self.size))
        self.all_keys = set()
        self.root = self._new()
        self.root.side = self.size

    def _new(self):
        o = Octant()
        o.code, o.side, o.is_leaf, o.type, o.child = 0, 0, False, NONLEAF, [None] * 8
        o.index = self.next_index
        self.next_index += 1
        return o

    def insert(self, pts):
        shift = MAX_BITS - self.max_level - 1
        for px, py, pz in np.asarray(pts, dtype=np.int64).reshape(-1, 3).tolist():
            for j in range(8):
                x, y, z = px + INCR_X[j], py + INCR_Y[j], pz + INCR_Z[j]
                key = encode(x, y, z)
                self.all_keys.add(key)
                n, edge = self.root, self.size // 2
                for d in range(1, self.max_level + 1):
                    childid = int((x & edge) > 0) + 2 * int((y & edge) > 0) + 4 * int((z & edge) > 0)
                    tmp = n.child[childid]
                    if tmp is None:
                        tmp = self._new()
                        tmp.code = key & MASK[d + shift]
                        tmp.side = edge
                        tmp.is_leaf = d == self.max_level
                        tmp.type = (SURFACE if j == 0 else FEATURE) if tmp.is_leaf else NONLEAF
                        n.child[childid] = tmp
                    elif tmp.type == FEATURE and j == 0:
                        tmp.type = SURFACE
                    n = tmp
                    edge //= 2

    def try_insert(self, pts):
        tmp = set()
        for px, py, pz in np.asarray(pts, dtype=np.int64).reshape(-1, 3).tolist():
            for j in range(8):
                tmp.add(encode(px + INCR_X[j], py + INCR_Y[j], pz + INCR_Z[j]))
        # (the reference collects the intersection in a std::set<int>: keys of a 256^3 grid fit in 24 bits)
        return len({int(k) & 0xFFFFFFFF for k in (self.all_keys & tmp)}) / len(tmp)

    def find_octant(self, x, y, z):
        n, edge = self.root, self.size // 2
        for _ in range(1, self.max_level + 1):
            n = n.child[int((x & edge) > 0) + 2 * int((y & edge) > 0) + 4 * int((z & edge) > 0)]
            if n is None:
                return None
            edge //= 2
        return n

    def has_voxel(self, pt):
        return self.find_octant(int(pt[0]), int(pt[1]), int(pt[2])) is not None

    def get_leaf_voxels(self):
        out = []

        def rec(n):
            if n is None:
                return
            if n.is_leaf and n.type == SURFACE:
                out.append(decode(n.code))
                return
            for c in n.child:
                rec(c)
        rec(self.root)
        return np.asarray(out, np.float32).reshape(-1, 3)

    def count_nodes(self):
        def rec(n):
            return 0 if n is None else 1 + sum(rec(c) for c in n.child)
        return rec(self.root)

    def count_leaf_nodes(self):
        def rec(n):
            if n is None:
                return 0
            return 1 if n.type == SURFACE else sum(rec(c) for c in n.child)
        return rec(self.root)

    def get_centres_and_children(self):
        total = self.count_nodes()
        voxels = np.zeros((total, 4), np.float32)
        children = -np.ones((total, 8), np.float32)
        features = -np.ones((total, 8), np.int32)
        queue = [self.root]
        while queue:
            node = queue.pop(0)
            x, y, z = decode(node.code)
            voxels[node.index] = (x, y, z, float(node.side))
            if node.type == SURFACE:
                for i in range(8):
                    v = self.find_octant(x + INCR_X[i], y + INCR_Y[i], z + INCR_Z[i])
                    if v is not None:
                        features[node.index, i] = v.index
            for i in range(8):
                c = node.child[i]
                if c is not None and c.type != FEATURE:
                    queue.append(c)
                    children[node.index, i] = float(c.index)
        return voxels, children, features
