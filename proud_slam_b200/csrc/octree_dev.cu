// Device-side sparse voxel octree: Octree::insert and get_centres_and_children of the reference's torch.classes.svo.Octree
// (third_party/sparse_octree/src/octree.cpp:104-294, 561-687; used at src/mapping.py:258-295, 301-406) on the GPU, producing
// the SAME tensors with the SAME row ids as the sequential host code (octree_host.cpp), so the map the render path reads is
// built where it is consumed: no host tree, no re-flatten + upload per keyframe.
//
// Row ids are creation order in the reference (octree.h:41).  Sequential insertion visits item q = 8 i + j (voxel i, corner j)
// after item q - 1 and creates the missing octants of its root-to-leaf path top-down, so an octant that does not exist yet is
// created by the FIRST item whose path contains it, and the creation order of a batch is the order of (first item, depth).
// That is computable in parallel:
//   k_octree_touch   one thread per (item, level): the octant's key (its path of child ids behind a leading 1 bit) goes into a
//                    hash table; octants not yet in the tree take the minimum item index that needs them (atomicMin), leaves
//                    visited by a corner-0 item are marked SURFACE (a FEATURE leaf is promoted, octree.cpp:150-160)
//   k_octree_pending collects the new octants with the sort key (first item, depth); a radix sort (cub) ranks them
//   k_octree_create  rank r becomes row n_nodes + r: coordinates and side from the key, type from the SURFACE mark
//   k_octree_link    every new octant looks its parent up and enters itself as that parent's child
//   k_octree_flatten per row: (lower corner, side), child rows (FEATURE children hidden, octree.cpp:600-640), and for a SURFACE
//                    voxel the rows of its 8 corner leaves (hash look-ups) -- straight into device tensors
// Child id = x bit + 2 y bit + 4 z bit of the level's edge (octree.cpp:12-14, 125-135), the leaf of corner j of voxel v is the
// leaf at v + (j>>2&1, j>>1&1, j&1).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace pslam {
namespace {

enum : int { kNonLeaf = -1, kSurface = 0, kFeature = 1 };
constexpr int kMaxProbe = 256;
constexpr int kFirstNone = 0x7fffffff;

struct DevTree {
    int size = 0, max_level = 0;
    int cap = 0, hcap = 0, n_nodes = 0;
    unsigned long long *hkey = nullptr;     // [hcap] 0 = empty
    int *hrow = nullptr, *hfirst = nullptr, *hsurf = nullptr;   // [hcap] row id (-1: new in this batch), first item, SURFACE mark
    int *child = nullptr, *xyzs = nullptr, *type = nullptr;     // [cap][8], [cap][4], [cap]
    unsigned long long *nkey = nullptr;     // [cap]
    unsigned long long *sort_k[2] = {nullptr, nullptr};         // [cap] (first item << 5 | depth), double buffer of the sort
    int *sort_v[2] = {nullptr, nullptr};                        // [cap] hash slot
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    int *counters = nullptr;                // device [4]: 0 new octants, 1 hash overflow
};

__device__ __forceinline__ unsigned long long octant_key(int x, int y, int z, int size, int d)
{
    unsigned long long k = 1ull;
    unsigned edge = (unsigned)size >> 1;
    for (int t = 1; t <= d; ++t, edge >>= 1)
        k = (k << 3) | (unsigned long long)(((x & edge) > 0) + 2 * ((y & edge) > 0) + 4 * ((z & edge) > 0));
    return k;
}
__device__ __forceinline__ unsigned hash_key(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}
// slot of `key`, inserting it when absent (insert) -- -1: not found / table full
__device__ __forceinline__ int hash_slot(unsigned long long *hkey, int hcap, unsigned long long key, bool insert)
{
    unsigned s = hash_key(key) & (unsigned)(hcap - 1);
    for (int p = 0; p < kMaxProbe; ++p, s = (s + 1) & (unsigned)(hcap - 1)) {
        unsigned long long k = hkey[s];
        if (k == key) return (int)s;
        if (k == 0ull) {
            if (!insert) return -1;
            k = atomicCAS(hkey + s, 0ull, key);
            if (k == 0ull || k == key) return (int)s;
        }
    }
    return -1;
}

__global__ void k_octree_fill(int *a, int v, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = v;
}

__global__ void k_octree_touch(const int *__restrict__ vox, int m, int size, int L, unsigned long long *hkey, int hcap, int *hrow, int *hfirst,
                               int *hsurf, int *type, int *counters)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m * 8 * L) return;
    const int d = (int)(t % L) + 1;
    const int q = (int)(t / L), i = q >> 3, j = q & 7;
    const int x = vox[i * 3] + ((j >> 2) & 1), y = vox[i * 3 + 1] + ((j >> 1) & 1), z = vox[i * 3 + 2] + (j & 1);
    const unsigned long long key = octant_key(x, y, z, size, d);
    const int s = hash_slot(hkey, hcap, key, true);
    if (s < 0) { atomicOr(counters + 1, 1); return; }
    const int row = hrow[s];
    if (row < 0) {
        atomicMin(hfirst + s, q);
        if (d == L && j == 0) hsurf[s] = 1;
    } else if (d == L && j == 0) {
        type[row] = kSurface;       // a corner-0 visit promotes a FEATURE leaf (and leaves a SURFACE one as it is)
    }
}

__global__ void k_octree_pending(const unsigned long long *__restrict__ hkey, const int *__restrict__ hrow, const int *__restrict__ hfirst,
                                 int hcap, int cap, unsigned long long *sort_k, int *sort_v, int *counters)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= hcap) return;
    const unsigned long long key = hkey[s];
    if (key == 0ull || hrow[s] >= 0) return;
    const int at = atomicAdd(counters, 1);
    if (at >= cap) return;                      // (the host grows the arrays and repeats the pass)
    const int d = (63 - __clzll((long long)key)) / 3;
    sort_k[at] = ((unsigned long long)(unsigned)hfirst[s] << 5) | (unsigned long long)d;
    sort_v[at] = s;
}

__global__ void k_octree_create(const int *__restrict__ slots, int n_new, int n_nodes, int size, int L, const unsigned long long *__restrict__ hkey,
                                int *hrow, const int *__restrict__ hsurf, int *child, int *xyzs, int *type, unsigned long long *nkey)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_new) return;
    const int s = slots[r], row = n_nodes + r;
    const unsigned long long key = hkey[s];
    const int d = (63 - __clzll((long long)key)) / 3;
    int x = 0, y = 0, z = 0;
    for (int t = 1; t <= d; ++t) {
        const int digit = (int)((key >> (3 * (d - t))) & 7ull), edge = size >> t;
        x |= (digit & 1) ? edge : 0; y |= (digit & 2) ? edge : 0; z |= (digit & 4) ? edge : 0;
    }
    hrow[s] = row;
    nkey[row] = key;
    xyzs[row * 4] = x; xyzs[row * 4 + 1] = y; xyzs[row * 4 + 2] = z; xyzs[row * 4 + 3] = size >> d;
    type[row] = d == L ? (hsurf[s] ? kSurface : kFeature) : kNonLeaf;
#pragma unroll
    for (int c = 0; c < 8; ++c) child[row * 8 + c] = -1;
}

__global__ void k_octree_link(int n_new, int n_nodes, const unsigned long long *__restrict__ nkey, unsigned long long *hkey, int hcap,
                              const int *__restrict__ hrow, int *child)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_new) return;
    const int row = n_nodes + r;
    const unsigned long long key = nkey[row];
    const int s = hash_slot(hkey, hcap, key >> 3, false);      // the parent exists: an older octant or one created in this batch
    if (s >= 0 && hrow[s] >= 0) child[hrow[s] * 8 + (int)(key & 7ull)] = row;
}

__global__ void k_octree_rehash(const unsigned long long *__restrict__ nkey, int n_nodes, unsigned long long *hkey, int hcap, int *hrow, int *counters)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_nodes) return;
    const int s = hash_slot(hkey, hcap, nkey[row], true);
    if (s < 0) { atomicOr(counters + 1, 1); return; }
    hrow[s] = row;
}

__global__ void k_octree_flatten(int n_nodes, int size, int L, const int *__restrict__ child, const int *__restrict__ xyzs, const int *__restrict__ type,
                                 unsigned long long *hkey, int hcap, const int *__restrict__ hrow, float *__restrict__ voxels,
                                 float *__restrict__ children, int *__restrict__ features)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_nodes * 8) return;
    const int row = t >> 3, i = t & 7;
    const int ty = type[row];
    const int x = xyzs[row * 4], y = xyzs[row * 4 + 1], z = xyzs[row * 4 + 2], side = xyzs[row * 4 + 3];
    if (i < 4) voxels[row * 4 + i] = ty == kFeature ? 0.0f : (float)(i == 0 ? x : (i == 1 ? y : (i == 2 ? z : side)));
    float ch = -1.0f;
    int ft = -1;
    if (ty != kFeature) {
        const int c = child[row * 8 + i];
        if (c >= 0 && type[c] != kFeature) ch = (float)c;
        if (ty == kSurface) {
            const int s = hash_slot(hkey, hcap, octant_key(x + ((i >> 2) & 1), y + ((i >> 1) & 1), z + (i & 1), size, L), false);
            ft = s >= 0 ? hrow[s] : -1;
        }
    }
    children[t] = ch;
    features[t] = ft;
}

#define OCT_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) { set_error("device octree: %s", cudaGetErrorString(e__)); return (int)e__; } \
    } while (0)

int fill(int *a, int v, size_t n, cudaStream_t st)
{
    if (n == 0) return 0;
    k_octree_fill<<<(int)(n / 256 + 1 < 4096 ? n / 256 + 1 : 4096), 256, 0, st>>>(a, v, n);
    PSLAM_CHECK_LAUNCH("octree_fill");
    return 0;
}

int alloc_hash(DevTree *t, int hcap, cudaStream_t st)
{
    cudaFree(t->hkey); cudaFree(t->hrow); cudaFree(t->hfirst); cudaFree(t->hsurf);
    t->hcap = hcap;
    OCT_CUDA(cudaMalloc(&t->hkey, sizeof(unsigned long long) * hcap));
    OCT_CUDA(cudaMalloc(&t->hrow, sizeof(int) * hcap));
    OCT_CUDA(cudaMalloc(&t->hfirst, sizeof(int) * hcap));
    OCT_CUDA(cudaMalloc(&t->hsurf, sizeof(int) * hcap));
    OCT_CUDA(cudaMemsetAsync(t->hkey, 0, sizeof(unsigned long long) * hcap, st));
    OCT_CUDA(cudaMemsetAsync(t->hrow, 0xff, sizeof(int) * hcap, st));
    OCT_CUDA(cudaMemsetAsync(t->hsurf, 0, sizeof(int) * hcap, st));
    return fill(t->hfirst, kFirstNone, (size_t)hcap, st);
}

template <class T>
int grow(T *&p, size_t old_n, size_t new_n, cudaStream_t st)
{
    T *q = nullptr;
    OCT_CUDA(cudaMalloc(&q, sizeof(T) * new_n));
    if (p && old_n) OCT_CUDA(cudaMemcpyAsync(q, p, sizeof(T) * old_n, cudaMemcpyDeviceToDevice, st));
    OCT_CUDA(cudaStreamSynchronize(st));
    cudaFree(p);
    p = q;
    return 0;
}

int grow_nodes(DevTree *t, int cap, cudaStream_t st)
{
    if (int rc = grow(t->child, (size_t)t->n_nodes * 8, (size_t)cap * 8, st)) return rc;
    if (int rc = grow(t->xyzs, (size_t)t->n_nodes * 4, (size_t)cap * 4, st)) return rc;
    if (int rc = grow(t->type, (size_t)t->n_nodes, (size_t)cap, st)) return rc;
    if (int rc = grow(t->nkey, (size_t)t->n_nodes, (size_t)cap, st)) return rc;
    for (int b = 0; b < 2; ++b) {
        cudaFree(t->sort_k[b]); cudaFree(t->sort_v[b]);
        OCT_CUDA(cudaMalloc(&t->sort_k[b], sizeof(unsigned long long) * cap));
        OCT_CUDA(cudaMalloc(&t->sort_v[b], sizeof(int) * cap));
    }
    cudaFree(t->cub_tmp);
    t->cub_tmp = nullptr;
    cub::DoubleBuffer<unsigned long long> dk(t->sort_k[0], t->sort_k[1]);
    cub::DoubleBuffer<int> dv(t->sort_v[0], t->sort_v[1]);
    OCT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t->cub_bytes, dk, dv, cap, 0, 40, st));
    OCT_CUDA(cudaMalloc(&t->cub_tmp, t->cub_bytes));
    t->cap = cap;
    return 0;
}

// the hash table is rebuilt from the rows (pending entries of an abandoned pass are dropped with the old table)
int rehash(DevTree *t, int hcap, cudaStream_t st)
{
    for (;;) {
        if (int rc = alloc_hash(t, hcap, st)) return rc;
        OCT_CUDA(cudaMemsetAsync(t->counters, 0, sizeof(int) * 4, st));
        if (t->n_nodes > 0) {
            k_octree_rehash<<<ceil_div(t->n_nodes, 256), 256, 0, st>>>(t->nkey, t->n_nodes, t->hkey, t->hcap, t->hrow, t->counters);
            PSLAM_CHECK_LAUNCH("octree_rehash");
        }
        int c[4];
        OCT_CUDA(cudaMemcpyAsync(c, t->counters, sizeof(c), cudaMemcpyDeviceToHost, st));
        OCT_CUDA(cudaStreamSynchronize(st));
        if (!c[1]) return 0;
        hcap *= 2;
    }
}

}  // namespace
}  // namespace pslam

using namespace pslam;

extern "C" void *pslam_doctree_new(int grid_dim, int capacity_hint)
{
    if (grid_dim < 2) { set_error("device octree: grid_dim must be >= 2"); return nullptr; }
    DevTree *t = new DevTree();
    t->size = grid_dim;
    int lv = 0;
    while ((1 << (lv + 1)) <= grid_dim) ++lv;   // log2(size), octree.cpp:55
    t->max_level = lv;
    if (lv > 21) { set_error("device octree: grid_dim too large for 64-bit path keys"); delete t; return nullptr; }
    cudaStream_t st = nullptr;
    int cap = capacity_hint > 1024 ? capacity_hint : 1024;
    int hcap = 1;
    while (hcap < 4 * cap) hcap <<= 1;
    if (cudaMalloc(&t->counters, sizeof(int) * 4) != cudaSuccess || grow_nodes(t, cap, st) || alloc_hash(t, hcap, st)) {
        set_error("device octree: allocation failed");
        return nullptr;
    }
    // root = row 0 (key 1)
    const int root_child[8] = {-1, -1, -1, -1, -1, -1, -1, -1}, root_xyzs[4] = {0, 0, 0, grid_dim}, root_type = kNonLeaf;
    const unsigned long long root_key = 1ull;
    cudaMemcpy(t->child, root_child, sizeof(root_child), cudaMemcpyHostToDevice);
    cudaMemcpy(t->xyzs, root_xyzs, sizeof(root_xyzs), cudaMemcpyHostToDevice);
    cudaMemcpy(t->type, &root_type, sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(t->nkey, &root_key, sizeof(root_key), cudaMemcpyHostToDevice);
    t->n_nodes = 1;
    if (rehash(t, t->hcap, st)) return nullptr;
    return t;
}

extern "C" void pslam_doctree_free(void *h)
{
    DevTree *t = static_cast<DevTree *>(h);
    if (!t) return;
    cudaFree(t->hkey); cudaFree(t->hrow); cudaFree(t->hfirst); cudaFree(t->hsurf);
    cudaFree(t->child); cudaFree(t->xyzs); cudaFree(t->type); cudaFree(t->nkey);
    for (int b = 0; b < 2; ++b) { cudaFree(t->sort_k[b]); cudaFree(t->sort_v[b]); }
    cudaFree(t->cub_tmp); cudaFree(t->counters);
    delete t;
}

extern "C" int pslam_doctree_count(void *h) { return h ? static_cast<DevTree *>(h)->n_nodes : -1; }

/* Octree::insert, octree.cpp:104-294; vox [m,3] int32 voxel coordinates ON THE DEVICE (0 <= v < grid_dim - 1).  Synchronises
 * the stream once per pass (the count of new octants sizes the arrays, as the host tree's vector growth does). */
extern "C" int pslam_doctree_insert(void *h, const int *vox, int m, pslam_stream_t stream)
{
    DevTree *t = static_cast<DevTree *>(h);
    PSLAM_CHECK_ARG(t && (vox || m == 0) && m >= 0, PSLAM_E_ARG, "device octree insert: bad argument");
    if (m == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int L = t->max_level;
    const long long threads = (long long)m * 8 * L;
    for (int attempt = 0; attempt < 12; ++attempt) {
        OCT_CUDA(cudaMemsetAsync(t->counters, 0, sizeof(int) * 4, st));
        k_octree_touch<<<(int)((threads + 255) / 256), 256, 0, st>>>(vox, m, t->size, L, t->hkey, t->hcap, t->hrow, t->hfirst, t->hsurf, t->type, t->counters);
        PSLAM_CHECK_LAUNCH("octree_touch");
        k_octree_pending<<<ceil_div(t->hcap, 256), 256, 0, st>>>(t->hkey, t->hrow, t->hfirst, t->hcap, t->cap, t->sort_k[0], t->sort_v[0], t->counters);
        PSLAM_CHECK_LAUNCH("octree_pending");
        int c[4];
        OCT_CUDA(cudaMemcpyAsync(c, t->counters, sizeof(c), cudaMemcpyDeviceToHost, st));
        OCT_CUDA(cudaStreamSynchronize(st));
        const int n_new = c[0];
        const bool hash_full = c[1] != 0 || (long long)(t->n_nodes + n_new) * 2 > t->hcap;
        const bool nodes_full = t->n_nodes + n_new > t->cap || n_new > t->cap;
        if (hash_full || nodes_full) {
            // grow and repeat the pass: the rebuilt table holds the committed rows only, nothing of this batch was committed
            if (nodes_full) {
                int cap = t->cap;
                while (cap < t->n_nodes + n_new) cap *= 2;
                if (int rc = grow_nodes(t, cap, st)) return rc;
            }
            int hcap = t->hcap;
            while ((long long)(t->n_nodes + n_new) * 4 > hcap || (c[1] && hcap <= t->hcap)) hcap *= 2;
            if (int rc = rehash(t, hcap, st)) return rc;
            continue;
        }
        if (n_new == 0) return 0;
        cub::DoubleBuffer<unsigned long long> dk(t->sort_k[0], t->sort_k[1]);
        cub::DoubleBuffer<int> dv(t->sort_v[0], t->sort_v[1]);
        size_t bytes = t->cub_bytes;
        OCT_CUDA(cub::DeviceRadixSort::SortPairs(t->cub_tmp, bytes, dk, dv, n_new, 0, 40, st));
        k_octree_create<<<ceil_div(n_new, 256), 256, 0, st>>>(dv.Current(), n_new, t->n_nodes, t->size, L, t->hkey, t->hrow, t->hsurf, t->child, t->xyzs,
                                                             t->type, t->nkey);
        PSLAM_CHECK_LAUNCH("octree_create");
        k_octree_link<<<ceil_div(n_new, 256), 256, 0, st>>>(n_new, t->n_nodes, t->nkey, t->hkey, t->hcap, t->hrow, t->child);
        PSLAM_CHECK_LAUNCH("octree_link");
        t->n_nodes += n_new;
        return 0;
    }
    set_error("device octree insert: the hash table kept overflowing");
    return PSLAM_E_RANGE;
}

/* get_centres_and_children, octree.cpp:561-687, into DEVICE tensors: voxels [N,4] f32, children [N,8] f32, features [N,8] i32 */
extern "C" int pslam_doctree_flatten(void *h, float *voxels, float *children, int *features, pslam_stream_t stream)
{
    DevTree *t = static_cast<DevTree *>(h);
    PSLAM_CHECK_ARG(t && voxels && children && features, PSLAM_E_ARG, "device octree flatten: bad argument");
    k_octree_flatten<<<ceil_div(t->n_nodes * 8, 256), 256, 0, (cudaStream_t)stream>>>(t->n_nodes, t->size, t->max_level, t->child, t->xyzs, t->type,
                                                                                     t->hkey, t->hcap, t->hrow, voxels, children, features);
    PSLAM_CHECK_LAUNCH("octree_flatten");
    return 0;
}

/* number of SURFACE voxels is not kept on the host: count_leaf_nodes reads the type array */
extern "C" int pslam_doctree_types(void *h, int *types_dev_out, pslam_stream_t stream)
{
    DevTree *t = static_cast<DevTree *>(h);
    PSLAM_CHECK_ARG(t && types_dev_out, PSLAM_E_ARG, "device octree types: bad argument");
    cudaError_t e = cudaMemcpyAsync(types_dev_out, t->type, sizeof(int) * t->n_nodes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("device octree types: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
