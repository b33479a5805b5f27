// Host-side sparse voxel octree that produces the tensors the render path reads
// (the reference's `torch.classes.svo.Octree`, third_party/sparse_octree/src).
//
// Behaviour kept from the reference:
//   insert     octree.cpp:104-294   for every voxel and each of its 8 corners
//                                   (corner j = (j>>2&1, j>>1&1, j&1), octree.cpp:12-14) walk
//                                   the levels and create missing octants; an octant's row id
//                                   is its creation order (octree.h:41), the root is row 0; a
//                                   leaf created by corner 0 is a SURFACE voxel, by another
//                                   corner a FEATURE vertex that a later corner-0 visit promotes
//   flatten    octree.cpp:561-687   BFS from the root over non-FEATURE octants: voxels[row] =
//                                   (lower corner xyz, side), children[row] = child rows (-1 for
//                                   missing or FEATURE children), features[row] = rows of the 8
//                                   corner leaves of a SURFACE voxel; unvisited rows keep 0 / -1
//   find       octree.cpp:419-439
// The per-octant point-cloud payload (pcd_xyz / pcd_color) feeds only the
// `get_features_pcd` branch that the reference has commented out of render_rays
// (render_helpers.py:481), so it is not stored; the Python class returns zeros
// of the reference's shapes for it.
//
// Not a pointer tree: octants live in one growable array indexed by row id, child
// links are row ids, so flatten is two linear passes and the structure can be
// uploaded as-is.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/proud_slam_b200.h"

namespace {

enum : int8_t { kNonLeaf = -1, kSurface = 0, kFeature = 1 };

struct Octant {
    int32_t child[8];
    int32_t x, y, z;   // lower corner, voxel units
    int32_t side;
    int8_t type;
};

struct Tree {
    std::vector<Octant> nodes;
    int size = 0, max_level = 0;

    int new_node(int x, int y, int z, int side, int8_t type)
    {
        Octant o;
        for (int &c : o.child) c = -1;
        o.x = x; o.y = y; o.z = z; o.side = side; o.type = type;
        nodes.push_back(o);
        return (int)nodes.size() - 1;
    }

    static int child_id(int x, int y, int z, unsigned edge)
    {
        return ((x & edge) > 0) + 2 * ((y & edge) > 0) + 4 * ((z & edge) > 0);
    }

    void insert_corner(int x, int y, int z, bool first_corner)
    {
        int n = 0;
        unsigned edge = (unsigned)size / 2;
        for (int d = 1; d <= max_level; edge /= 2, ++d) {
            const int cid = child_id(x, y, z, edge);
            int c = nodes[n].child[cid];
            if (c < 0) {
                const bool leaf = (d == max_level);
                const int keep = ~((int)edge - 1);
                c = new_node(x & keep, y & keep, z & keep, (int)edge,
                             leaf ? (first_corner ? kSurface : kFeature) : kNonLeaf);
                nodes[n].child[cid] = c;
            } else if (nodes[c].type == kFeature && first_corner) {
                nodes[c].type = kSurface;
            }
            n = c;
        }
    }

    int find(int x, int y, int z) const
    {
        int n = 0;
        unsigned edge = (unsigned)size / 2;
        for (int d = 1; d <= max_level; edge /= 2, ++d) {
            n = nodes[n].child[child_id(x, y, z, edge)];
            if (n < 0) return -1;
        }
        return n;
    }
};

}  // namespace

extern "C" void *pslam_octree_new(int grid_dim)
{
    if (grid_dim < 2) return nullptr;
    Tree *t = new Tree();
    t->size = grid_dim;
    int lv = 0;
    while ((1 << (lv + 1)) <= grid_dim) ++lv;   // log2(size), octree.cpp:55
    t->max_level = lv;
    t->new_node(0, 0, 0, grid_dim, kNonLeaf);   // root = row 0
    return t;
}

extern "C" void pslam_octree_free(void *h) { delete static_cast<Tree *>(h); }

extern "C" int pslam_octree_count(void *h) { return h ? (int)static_cast<Tree *>(h)->nodes.size() : -1; }

extern "C" int pslam_octree_count_leaves(void *h)
{
    if (!h) return -1;
    int n = 0;
    for (const Octant &o : static_cast<Tree *>(h)->nodes) n += (o.type == kSurface);
    return n;
}

extern "C" int pslam_octree_insert(void *h, const int *vox, int m)
{
    if (!h || (!vox && m > 0) || m < 0) return PSLAM_E_ARG;
    Tree *t = static_cast<Tree *>(h);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < 8; ++j)
            t->insert_corner(vox[i * 3] + ((j >> 2) & 1), vox[i * 3 + 1] + ((j >> 1) & 1), vox[i * 3 + 2] + (j & 1), j == 0);
    return 0;
}

extern "C" int pslam_octree_has_voxel(void *h, int x, int y, int z)
{
    if (!h) return -1;
    const Tree *t = static_cast<Tree *>(h);
    return t->find(x, y, z) >= 0;   // any leaf at that coordinate, SURFACE voxel or FEATURE corner (octree.cpp:441-473)
}

extern "C" int pslam_octree_flatten(void *h, float *voxels, float *children, int *features)
{
    if (!h || !voxels || !children || !features) return PSLAM_E_ARG;
    const Tree *t = static_cast<Tree *>(h);
    const size_t N = t->nodes.size();
    memset(voxels, 0, sizeof(float) * 4 * N);
    for (size_t i = 0; i < N * 8; ++i) { children[i] = -1.0f; features[i] = -1; }
    // reachable = root and every non-FEATURE octant (each has exactly one parent), which is
    // what the reference's BFS visits; row order is irrelevant because rows are addressed by id
    for (size_t r = 0; r < N; ++r) {
        const Octant &o = t->nodes[r];
        if (o.type == kFeature) continue;
        float *v = voxels + r * 4;
        v[0] = (float)o.x; v[1] = (float)o.y; v[2] = (float)o.z; v[3] = (float)o.side;
        if (o.type == kSurface)
            for (int i = 0; i < 8; ++i)
                features[r * 8 + i] = t->find(o.x + ((i >> 2) & 1), o.y + ((i >> 1) & 1), o.z + (i & 1));
        for (int i = 0; i < 8; ++i) {
            const int c = o.child[i];
            if (c >= 0 && t->nodes[c].type != kFeature) children[r * 8 + i] = (float)c;
        }
    }
    return 0;
}

// rows of SURFACE voxels as (x, y, z) voxel coordinates; returns how many exist
extern "C" int pslam_octree_leaf_voxels(void *h, int *out, int cap)
{
    if (!h) return -1;
    int n = 0;
    for (const Octant &o : static_cast<Tree *>(h)->nodes)
        if (o.type == kSurface) {
            if (out && n < cap) { out[n * 3] = o.x; out[n * 3 + 1] = o.y; out[n * 3 + 2] = o.z; }
            ++n;
        }
    return n;
}
