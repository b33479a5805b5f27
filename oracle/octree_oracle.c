/*
 * oracle/octree_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Dependency-free restatement of the part of the reference's CPU octree that
 * defines the tensors the render path reads (the reference's own build needs
 * PCL/Eigen/glog/TBB, none of which exist in this image -- SURVEY.md 8(c)):
 *
 *   octree_insert   <- third_party/sparse_octree/src/octree.cpp:104-294
 *                      (8 corner keys per voxel, level loop, SURFACE/FEATURE typing;
 *                      the per-octant point-cloud payload is out of scope and omitted)
 *   octree_flatten  <- octree.cpp:561-687 (BFS from the root, rows indexed by
 *                      Octant::index_ = creation order, octree.h:41)
 *   find            <- octree.cpp:419-439
 *   node corner     <- utils.h:79-124 (Morton code truncated at the node's
 *                      level, then decoded = coordinate with the low bits cleared)
 *
 * A pointer tree is kept on purpose (like the reference) so that the product's
 * array/hash octree in proud_slam_b200/csrc/octree_host.cpp is an independent
 * implementation checked against this one.
 *
 * Parity unpinned by reference tests (it has none); pinned by construction
 * rules only, and by the reference Python rendering the result (make_golden.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { T_NONLEAF = -1, T_SURFACE = 0, T_FEATURE = 1 };

typedef struct Node {
    struct Node *child[8];
    int index, side, type, is_leaf;
    int cx, cy, cz; /* lower corner in voxel units */
} Node;

typedef struct {
    Node *root;
    Node **by_index; /* creation order */
    int count, cap;
    int size, max_level;
} Tree;

static Node *new_node(Tree *t)
{
    Node *n = (Node *)calloc(1, sizeof(Node));
    n->index = t->count;
    n->type = T_NONLEAF;
    if (t->count == t->cap) {
        t->cap = t->cap ? t->cap * 2 : 1024;
        t->by_index = (Node **)realloc(t->by_index, sizeof(Node *) * (size_t)t->cap);
    }
    t->by_index[t->count++] = n;
    return n;
}

void *octree_new(int grid_dim)
{
    Tree *t = (Tree *)calloc(1, sizeof(Tree));
    t->size = grid_dim;
    int lv = 0;
    while ((1 << (lv + 1)) <= grid_dim) ++lv; /* log2, octree.cpp:55 */
    t->max_level = lv;
    t->root = new_node(t); /* row 0 */
    t->root->side = grid_dim;
    return t;
}

void octree_free(void *h)
{
    Tree *t = (Tree *)h;
    for (int i = 0; i < t->count; ++i) free(t->by_index[i]);
    free(t->by_index);
    free(t);
}

int octree_count(void *h) { return ((Tree *)h)->count; }

/* vox: [m,3] int32 voxel coordinates (floor(point / voxel_size), mapping.py:264). */
void octree_insert(void *h, const int *vox, int m)
{
    Tree *t = (Tree *)h;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < 8; ++j) { /* corner j = (j>>2&1, j>>1&1, j&1), octree.cpp:12-14 */
            const int x = vox[i * 3 + 0] + ((j >> 2) & 1);
            const int y = vox[i * 3 + 1] + ((j >> 1) & 1);
            const int z = vox[i * 3 + 2] + (j & 1);
            Node *n = t->root;
            unsigned edge = (unsigned)t->size / 2;
            for (int d = 1; d <= t->max_level; edge /= 2, ++d) {
                const int cid = ((x & edge) > 0) + 2 * ((y & edge) > 0) + 4 * ((z & edge) > 0);
                Node *c = n->child[cid];
                if (!c) {
                    c = new_node(t);
                    c->side = (int)edge;
                    c->is_leaf = (d == t->max_level);
                    c->type = c->is_leaf ? (j == 0 ? T_SURFACE : T_FEATURE) : T_NONLEAF;
                    const int keep = ~((int)edge - 1);
                    c->cx = x & keep; c->cy = y & keep; c->cz = z & keep;
                    n->child[cid] = c;
                } else if (c->type == T_FEATURE && j == 0) {
                    c->type = T_SURFACE;
                }
                n = c;
            }
        }
}

static Node *find(Tree *t, int x, int y, int z)
{
    Node *n = t->root;
    unsigned edge = (unsigned)t->size / 2;
    for (int d = 1; d <= t->max_level; edge /= 2, ++d) {
        n = n->child[((x & edge) > 0) + 2 * ((y & edge) > 0) + 4 * ((z & edge) > 0)];
        if (!n) return 0;
    }
    return n;
}

/* voxels [N,4] f32 (zeros), children [N,8] f32 (-1), features [N,8] i32 (-1);
 * rows of FEATURE leaves are never visited and keep those defaults. */
void octree_flatten(void *h, float *voxels, float *children, int *features)
{
    Tree *t = (Tree *)h;
    const int N = t->count;
    memset(voxels, 0, sizeof(float) * 4 * (size_t)N);
    for (long i = 0; i < (long)N * 8; ++i) { children[i] = -1.0f; features[i] = -1; }
    Node **queue = (Node **)malloc(sizeof(Node *) * (size_t)N);
    int qh = 0, qt = 0;
    queue[qt++] = t->root;
    while (qh < qt) {
        Node *n = queue[qh++];
        float *v = voxels + (size_t)n->index * 4;
        v[0] = (float)n->cx; v[1] = (float)n->cy; v[2] = (float)n->cz; v[3] = (float)n->side;
        if (n->type == T_SURFACE)
            for (int i = 0; i < 8; ++i) {
                Node *c = find(t, n->cx + ((i >> 2) & 1), n->cy + ((i >> 1) & 1), n->cz + (i & 1));
                if (c) features[(size_t)n->index * 8 + i] = c->index;
            }
        for (int i = 0; i < 8; ++i) {
            Node *c = n->child[i];
            if (c && c->type != T_FEATURE) {
                queue[qt++] = c;
                children[(size_t)n->index * 8 + i] = (float)c->index;
            }
        }
    }
    free(queue);
}
