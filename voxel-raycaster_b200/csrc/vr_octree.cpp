/* vr_octree.cpp -- see vr_octree.h. */
#include "vr_octree.h"

#include <string.h>

#include <algorithm>
#include <functional>
#include <stdexcept>

namespace {

inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

/* ------------------------------------------------------------------------------------------------
 * reference-format generator
 * ---------------------------------------------------------------------------------------------- */
struct TNode {
    uint8_t valid = 0, leaf = 0;
    bool has_block = false;           /* false for 2^3-level descriptors (children are voxels) */
    uint32_t kid[8];
    uint64_t subtree = 0;             /* entries (descriptors + far slots) below this node      */
    uint8_t far_mask = 0;             /* bit j: j-th valid child is reached through a far ptr   */
};

struct RefGen {
    const int8_t *data;
    int dim;
    std::vector<TNode> nodes;
    static constexpr uint32_t kNone = 0xFFFFFFFFu;

    inline int8_t vox(int x, int y, int z) const {
        return data[(size_t)x + (size_t)dim * ((size_t)y + (size_t)dim * (size_t)z)];
    }
    /* returns node id, or kNone when the cube is uniformly empty */
    uint32_t build(int px, int py, int pz, int size) {
        const int h = size / 2;
        TNode n;
        for (int i = 0; i < 8; i++) n.kid[i] = kNone;
        if (size == 2) {
            for (int i = 0; i < 8; i++)
                if (vox(px + (i & 1), py + ((i >> 1) & 1), pz + ((i >> 2) & 1))) n.valid |= (uint8_t)(1u << i);
            n.leaf = 0xFF;
            if (!n.valid) return kNone;
            nodes.push_back(n);
            return (uint32_t)nodes.size() - 1;
        }
        for (int i = 0; i < 8; i++) {
            const uint32_t k = build(px + (i & 1) * h, py + ((i >> 1) & 1) * h, pz + ((i >> 2) & 1) * h, h);
            if (k == kNone) n.leaf |= (uint8_t)(1u << i);
            else { n.valid |= (uint8_t)(1u << i); n.kid[i] = k; }
        }
        if (!n.valid) return kNone;
        n.has_block = true;
        /* block = one descriptor per valid child + far slots; far decision is pessimistic in nfar */
        int m = __builtin_popcount(n.valid);
        uint64_t before = 0;
        int j = 0, nfar = 0;
        for (int i = 0; i < 8; i++) {
            if (n.kid[i] == kNone) continue;
            const TNode &c = nodes[n.kid[i]];
            if (c.has_block && (uint64_t)(m - j) + (uint64_t)m + before > 0x7fffull) { n.far_mask |= (uint8_t)(1u << j); nfar++; }
            before += c.subtree;
            j++;
        }
        n.subtree = (uint64_t)m + (uint64_t)nfar + before;
        nodes.push_back(n);
        return (uint32_t)nodes.size() - 1;
    }

    void place(uint32_t id, uint64_t q, std::vector<uint64_t> &buf) const {
        const TNode &n = nodes[id];
        const int m = __builtin_popcount(n.valid);
        const int nfar = __builtin_popcount(n.far_mask);
        uint64_t cursor = q + (uint64_t)m + (uint64_t)nfar;
        int j = 0, k = 0;
        for (int i = 0; i < 8; i++) {
            if (n.kid[i] == kNone) continue;
            const TNode &c = nodes[n.kid[i]];
            uint64_t desc = ((uint64_t)c.valid << 16) | ((uint64_t)c.leaf << 24);
            if (c.has_block) {
                const uint64_t at = q + (uint64_t)j;
                if (n.far_mask & (1u << j)) {
                    const uint64_t slot = q + (uint64_t)m + (uint64_t)k;
                    buf[slot] = cursor;                       /* absolute index (Octree.cpp:281) */
                    desc |= 0x8000ull | (slot - at);
                    k++;
                } else {
                    desc |= (cursor - at);
                }
                buf[at] = desc;
                place(n.kid[i], cursor, buf);
                cursor += c.subtree;
            } else {
                buf[q + (uint64_t)j] = desc;
            }
            j++;
        }
    }
};

/* ------------------------------------------------------------------------------------------------
 * native 64-tree emit from a mask pyramid
 * ---------------------------------------------------------------------------------------------- */
struct Pyramid {
    int levels = 0;
    std::vector<int> g;                              /* grid edge per level */
    std::vector<std::vector<uint64_t>> M;            /* masks per level     */
    inline uint64_t &at(int l, int x, int y, int z) { return M[l][(size_t)x + (size_t)g[l] * ((size_t)y + (size_t)g[l] * (size_t)z)]; }
};

void pyramid_init(Pyramid &p, int dim) {
    int L = 1;
    while ((1 << (2 * L)) < dim) L++;
    p.levels = L;
    p.g.assign(L, 1);
    p.M.assign(L, {});
    for (int l = L - 1; l >= 0; l--) {
        const int cell = 1 << (2 * (L - l));         /* voxels covered by a node of level l */
        p.g[l] = (dim + cell - 1) / cell;
        p.M[l].assign((size_t)p.g[l] * p.g[l] * p.g[l], 0ull);
    }
}

void pyramid_reduce(Pyramid &p) {
    for (int l = p.levels - 2; l >= 0; l--) {
        const int gc = p.g[l + 1];
#pragma omp parallel for schedule(static)
        for (int z = 0; z < p.g[l]; z++)
            for (int y = 0; y < p.g[l]; y++)
                for (int x = 0; x < p.g[l]; x++) {
                    uint64_t m = 0;
                    for (int ci = 0; ci < 64; ci++) {
                        const int cx = 4 * x + (ci & 3), cy = 4 * y + ((ci >> 2) & 3), cz = 4 * z + (ci >> 4);
                        if (cx < gc && cy < gc && cz < gc && p.at(l + 1, cx, cy, cz)) m |= 1ull << ci;
                    }
                    p.at(l, x, y, z) = m;
                }
    }
}

/* type_of(x,y,z) -> voxel value for a set leaf bit */
void pyramid_emit(Pyramid &p, int dim, const std::function<uint8_t(int, int, int)> &type_of, vr_native_tree &out) {
    out.nodes.clear();
    out.leaf_types.clear();
    out.levels = p.levels;
    out.dim = dim;
    out.solid_voxels = 0;
    struct Coord { int x, y, z; };
    std::vector<Coord> cur{{0, 0, 0}}, next;
    for (int l = 0; l < p.levels; l++) {
        const size_t level_start = out.nodes.size();
        const size_t next_start = level_start + cur.size();
        next.clear();
        for (const Coord &c : cur) {
            const uint64_t m = p.at(l, c.x, c.y, c.z);
            vr_node n;
            n.mask_lo = (uint32_t)m;
            n.mask_hi = (uint32_t)(m >> 32);
            n.aux = vr_node_planes(m);
            if (l == p.levels - 1) {
                n.child_base = (uint32_t)out.leaf_types.size();
                for (int ci = 0; ci < 64; ci++)
                    if ((m >> ci) & 1ull) {
                        out.leaf_types.push_back(type_of(4 * c.x + (ci & 3), 4 * c.y + ((ci >> 2) & 3), 4 * c.z + (ci >> 4)));
                        out.solid_voxels++;
                    }
            } else {
                n.child_base = (uint32_t)(next_start + next.size());
                for (int ci = 0; ci < 64; ci++)
                    if ((m >> ci) & 1ull) next.push_back({4 * c.x + (ci & 3), 4 * c.y + ((ci >> 2) & 3), 4 * c.z + (ci >> 4)});
            }
            out.nodes.push_back(n);
        }
        cur.swap(next);
    }
    if (out.leaf_types.empty()) out.leaf_types.push_back(0);   /* keep the device buffer non-empty */
    if (out.leaf_types.size() > 0xFFFFFFFFull || out.nodes.size() > 0x7FFFFFFFull)
        throw std::length_error("64-tree: more solid voxels or nodes than its 32-bit child pointers can address");
}

}  // namespace

bool vr_ref_octree_generate(const int8_t *data, int dim, std::vector<uint64_t> &out, uint64_t *root_index) {
    if (!data || dim < 2 || !is_pow2(dim)) return false;
    RefGen g{data, dim, {}};
    const uint32_t root = g.build(0, 0, 0, dim);
    out.clear();
    if (root == RefGen::kNone) {
        /* uniformly empty map: a root whose eight children are collapsed-empty */
        out.push_back((0xFFull << 24) | 1ull);
        if (root_index) *root_index = 0;
        return true;
    }
    const TNode &r = g.nodes[root];
    out.assign(1 + r.subtree, 0ull);
    out[0] = ((uint64_t)r.valid << 16) | ((uint64_t)r.leaf << 24) | 1ull;     /* Octree.cpp:27 */
    if (r.has_block) g.place(root, 1, out);
    if (root_index) *root_index = 0;
    return true;
}

int vr_ref_octree_query(const uint64_t *desc, uint64_t len, uint64_t root_index, int octdim, const int pos[3],
                        int sub_oct_pos[3], int *resolution) {
    uint64_t idx = root_index;
    int dimension = octdim;
    int res = dimension / 2;
    int o[3] = {0, 0, 0};
    sub_oct_pos[0] = sub_oct_pos[1] = sub_oct_pos[2] = 0;
    *resolution = res;
    if (!desc || idx >= len) return 0;
    uint64_t cd = desc[idx];
    while (dimension > 1) {
        const int half = dimension / 2;
        int ci = 0;
        for (int a = 0; a < 3; a++)
            if (pos[a] >= half + o[a]) { ci |= 1 << a; o[a] += half; }
        sub_oct_pos[0] = o[0]; sub_oct_pos[1] = o[1]; sub_oct_pos[2] = o[2];
        *resolution = res;
        const bool valid = (cd >> (16 + ci)) & 1ull, leaf = (cd >> (24 + ci)) & 1ull;
        if (!valid) return 0;
        if (leaf) return 1;                                   /* kernel:199-205: no halving */
        dimension = half;
        res /= 2;
        *resolution = res;
        const uint64_t count = (uint64_t)__builtin_popcountll((cd >> 16) & ((2ull << ci) - 1ull)) - 1ull;
        if (cd & 0x8000ull) {
            const uint64_t far = idx + (cd & 0x7fffull);
            if (far >= len) return 0;
            idx = desc[far] + count;
        } else {
            idx = idx + (cd & 0x7fffull) + count;
        }
        if (idx >= len) return 0;
        cd = desc[idx];
    }
    return 1;
}

bool vr_native_from_dense(const int8_t *map, int dim, vr_native_tree &out) {
    if (!map || dim < 1 || !is_pow2(dim)) return false;
    Pyramid p;
    pyramid_init(p, dim);
    const int L = p.levels, gl = p.g[L - 1];
#pragma omp parallel for schedule(static)
    for (int bz = 0; bz < gl; bz++)
        for (int by = 0; by < gl; by++)
            for (int bx = 0; bx < gl; bx++) {
                uint64_t m = 0;
                for (int cz = 0; cz < 4; cz++)
                    for (int cy = 0; cy < 4; cy++) {
                        const int y = 4 * by + cy, z = 4 * bz + cz;
                        if (y >= dim || z >= dim) continue;
                        const int8_t *row = map + (size_t)dim * ((size_t)y + (size_t)dim * (size_t)z) + 4 * (size_t)bx;
                        for (int cx = 0; cx < 4 && 4 * bx + cx < dim; cx++)
                            if (row[cx] == 5 || row[cx] == 6) m |= 1ull << (cx | (cy << 2) | (cz << 4));
                    }
                p.at(L - 1, bx, by, bz) = m;
            }
    pyramid_reduce(p);
    pyramid_emit(p, dim, [&](int x, int y, int z) {
        return (uint8_t)map[(size_t)x + (size_t)dim * ((size_t)y + (size_t)dim * (size_t)z)];
    }, out);
    return true;
}

bool vr_native_from_ref(const uint64_t *desc, uint64_t len, uint64_t root_index, int dim, const int8_t *types,
                        vr_native_tree &out) {
    if (!desc || root_index >= len || dim < 2 || !is_pow2(dim)) return false;
    Pyramid p;
    pyramid_init(p, dim);
    const int L = p.levels;
    bool ok = true;
    std::function<void(uint64_t, int, int, int, int)> walk = [&](uint64_t idx, int ox, int oy, int oz, int size) {
        const uint64_t cd = desc[idx];
        const int h = size / 2;
        for (int i = 0; i < 8; i++) {
            if (!((cd >> (16 + i)) & 1ull)) continue;
            const int cx = ox + (i & 1) * h, cy = oy + ((i >> 1) & 1) * h, cz = oz + ((i >> 2) & 1) * h;
            if (((cd >> (24 + i)) & 1ull) || h == 1) {       /* solid leaf cube of edge h */
                for (int z = cz; z < cz + h; z++)
                    for (int y = cy; y < cy + h; y++)
                        for (int x = cx; x < cx + h; x++)
                            p.at(L - 1, x >> 2, y >> 2, z >> 2) |= 1ull << ((x & 3) | ((y & 3) << 2) | ((z & 3) << 4));
                continue;
            }
            const uint64_t count = (uint64_t)__builtin_popcountll((cd >> 16) & ((2ull << i) - 1ull)) - 1ull;
            uint64_t child;
            if (cd & 0x8000ull) {
                const uint64_t far = idx + (cd & 0x7fffull);
                if (far >= len) { ok = false; return; }
                child = desc[far] + count;
            } else {
                child = idx + (cd & 0x7fffull) + count;
            }
            if (child >= len) { ok = false; return; }
            walk(child, cx, cy, cz, h);
        }
    };
    walk(root_index, 0, 0, 0, dim);
    if (!ok) return false;
    pyramid_reduce(p);
    pyramid_emit(p, dim, [&](int x, int y, int z) -> uint8_t {
        if (!types) return 5;
        const int8_t v = types[(size_t)x + (size_t)dim * ((size_t)y + (size_t)dim * (size_t)z)];
        return (v == 5 || v == 6) ? (uint8_t)v : (uint8_t)5;
    }, out);
    return true;
}

bool vr_native_from_columns(const int32_t *lo, const int32_t *hi, int dim, uint8_t type, vr_native_tree &out) {
    if (!lo || !hi || dim < 1 || !is_pow2(dim)) return false;
    int L = 1;
    while ((1 << (2 * L)) < dim) L++;
    const int gl = (dim + 3) / 4;                         /* leaf bricks per axis */
    struct Leaf { uint64_t key, mask; };
    /* BFS order inside a level == lexicographic order of the root-to-node child-slot path; the path of a leaf
     * brick (bx,by,bz) packs one 6-bit slot per level */
    auto key_of = [L](int bx, int by, int bz) {
        uint64_t k = 0;
        for (int j = L - 2; j >= 0; j--) {
            const int sx = (bx >> (2 * j)) & 3, sy = (by >> (2 * j)) & 3, sz = (bz >> (2 * j)) & 3;
            k = (k << 6) | (uint64_t)(sx | (sy << 2) | (sz << 4));
        }
        return k;
    };
    std::vector<std::vector<Leaf>> rows((size_t)gl);
#pragma omp parallel for schedule(dynamic, 4)
    for (int by = 0; by < gl; by++) {
        std::vector<Leaf> &acc = rows[(size_t)by];
        for (int bx = 0; bx < gl; bx++) {
            int zmin = dim, zmax = -1;
            int clo[16], chi[16];
            for (int c = 0; c < 16; c++) {
                const int x = 4 * bx + (c & 3), y = 4 * by + (c >> 2);
                clo[c] = 1; chi[c] = 0;
                if (x >= dim || y >= dim) continue;
                int a = lo[(size_t)x + (size_t)dim * (size_t)y], b = hi[(size_t)x + (size_t)dim * (size_t)y];
                if (a < 0) a = 0;
                if (b > dim - 1) b = dim - 1;
                if (a > b) continue;
                clo[c] = a; chi[c] = b;
                if (a < zmin) zmin = a;
                if (b > zmax) zmax = b;
            }
            for (int bz = zmin >> 2; zmax >= 0 && bz <= (zmax >> 2); bz++) {
                uint64_t m = 0;
                for (int c = 0; c < 16; c++)
                    for (int cz = 0; cz < 4; cz++) {
                        const int z = 4 * bz + cz;
                        if (z >= clo[c] && z <= chi[c]) m |= 1ull << (c | (cz << 4));
                    }
                if (m) acc.push_back({key_of(bx, by, bz), m});
            }
        }
    }
    std::vector<Leaf> level;
    {
        size_t total = 0;
        for (auto &r : rows) total += r.size();
        level.reserve(total);
        for (auto &r : rows) { level.insert(level.end(), r.begin(), r.end()); std::vector<Leaf>().swap(r); }
    }
    std::sort(level.begin(), level.end(), [](const Leaf &a, const Leaf &b) { return a.key < b.key; });
    /* levels[l] = (key, mask) of the nodes of level l, sorted; built bottom-up by grouping on key >> 6 */
    std::vector<std::vector<Leaf>> levels((size_t)L);
    levels[(size_t)L - 1].swap(level);
    for (int l = L - 2; l >= 0; l--) {
        const std::vector<Leaf> &kids = levels[(size_t)l + 1];
        std::vector<Leaf> &mine = levels[(size_t)l];
        for (const Leaf &k : kids) {
            const uint64_t pk = k.key >> 6;
            if (mine.empty() || mine.back().key != pk) mine.push_back({pk, 0});
            mine.back().mask |= 1ull << (k.key & 63);
        }
    }
    if (levels[0].empty()) levels[0].push_back({0, 0});          /* empty map: a root without children */
    out.nodes.clear();
    out.leaf_types.clear();
    out.levels = L;
    out.dim = dim;
    out.solid_voxels = 0;
    size_t start = 0;
    for (int l = 0; l < L; l++) {
        const std::vector<Leaf> &cur = levels[(size_t)l];
        const size_t next_start = start + cur.size();
        size_t child = 0;
        for (const Leaf &n : cur) {
            vr_node v;
            v.mask_lo = (uint32_t)n.mask;
            v.mask_hi = (uint32_t)(n.mask >> 32);
            v.aux = vr_node_planes(n.mask);
            const int pc = __builtin_popcountll(n.mask);
            if (l == L - 1) {
                v.child_base = (uint32_t)out.solid_voxels;
                out.solid_voxels += (uint64_t)pc;
            } else {
                v.child_base = (uint32_t)(next_start + child);
                child += (size_t)pc;
            }
            out.nodes.push_back(v);
        }
        start = next_start;
    }
    out.leaf_types.assign(out.solid_voxels ? (size_t)out.solid_voxels : 1, out.solid_voxels ? type : (uint8_t)0);
    return true;
}

size_t vr_native_collapse_solid(vr_native_tree &t) {
    const size_t n = t.nodes.size();
    const int L = t.levels;
    if (L < 1 || n == 0) return 0;
    auto mask_of = [&](size_t i) { return (uint64_t)t.nodes[i].mask_lo | ((uint64_t)t.nodes[i].mask_hi << 32); };
    /* level l = nodes [start[l], start[l + 1]) */
    std::vector<size_t> start((size_t)L + 1, 0);
    start[1] = 1;
    for (size_t i = 0; i < n; i++)
        if (t.nodes[i].child_base & VR_NODE_SOLID) return 0;      /* collapsed already (a loaded file, a received tree) */
    for (int l = 0; l + 1 < L; l++) {
        size_t kids = 0;
        for (size_t i = start[l]; i < start[l + 1] && i < n; i++) kids += (size_t)__builtin_popcountll(mask_of(i));
        start[l + 2] = start[l + 1] + kids;
    }
    if (start[L] != n) return 0;                                  /* not a complete BFS tree: leave it */
    /* (1) bottom-up: type of the node's cube if it is solid, else -1 */
    std::vector<int16_t> st(n, -1);
    size_t solid = 0;
    for (size_t i = start[L - 1]; i < start[L]; i++) {
        if (mask_of(i) != ~0ull) continue;
        const uint8_t *ty = t.leaf_types.data() + t.nodes[i].child_base;
        bool same = true;
        for (int k = 1; k < 64 && same; k++) same = ty[k] == ty[0];
        if (same) { st[i] = ty[0]; solid++; }
    }
    for (int l = L - 2; l >= 0; l--)
        for (size_t i = start[l]; i < start[l + 1]; i++) {
            if (mask_of(i) != ~0ull) continue;
            const size_t c = t.nodes[i].child_base;
            bool same = st[c] >= 0;
            for (int k = 1; k < 64 && same; k++) same = st[c + k] == st[c];
            if (same) { st[i] = st[c]; solid++; }
        }
    if (!solid) return 0;
    /* (2) top-down: a node is kept unless its parent is solid (or was dropped itself) */
    std::vector<uint8_t> keep(n, 0);
    keep[0] = 1;
    for (int l = 0; l + 1 < L; l++)
        for (size_t i = start[l]; i < start[l + 1]; i++) {
            const int pc = __builtin_popcountll(mask_of(i));
            const uint8_t k = keep[i] && st[i] < 0;
            for (int j = 0; j < pc; j++) keep[t.nodes[i].child_base + (size_t)j] = k;
        }
    /* (3) re-pack in the same order */
    std::vector<uint32_t> at(n, 0);
    uint32_t cnt = 0;
    for (size_t i = 0; i < n; i++) { at[i] = cnt; cnt += keep[i]; }
    std::vector<vr_node> nodes;
    std::vector<uint8_t> types;
    nodes.reserve(cnt);
    for (int l = 0; l < L; l++)
        for (size_t i = start[l]; i < start[l + 1]; i++) {
            if (!keep[i]) continue;
            vr_node v = t.nodes[i];
            if (st[i] >= 0) {
                v.child_base = VR_NODE_SOLID | (uint32_t)(uint8_t)st[i];
            } else if (l == L - 1) {
                const int pc = __builtin_popcountll(mask_of(i));
                const uint8_t *ty = t.leaf_types.data() + t.nodes[i].child_base;
                v.child_base = (uint32_t)types.size();
                types.insert(types.end(), ty, ty + pc);
            } else if (mask_of(i)) {
                v.child_base = at[t.nodes[i].child_base];
            }
            nodes.push_back(v);
        }
    if (types.empty()) types.push_back(0);
    const size_t removed = n - nodes.size();
    t.nodes.swap(nodes);
    t.leaf_types.swap(types);
    return removed;
}

int vr_native_query(const vr_native_tree &t, int x, int y, int z, int *cell_shift) {
    int s = 2 * (t.levels - 1);
    uint32_t idx = 0;
    for (;;) {
        const vr_node &n = t.nodes[idx];
        if (n.child_base & VR_NODE_SOLID) { if (cell_shift) *cell_shift = 0; return (int)(int8_t)(n.child_base & 0xffu); }
        const uint64_t m = (uint64_t)n.mask_lo | ((uint64_t)n.mask_hi << 32);
        const int ci = ((x >> s) & 3) | (((y >> s) & 3) << 2) | (((z >> s) & 3) << 4);
        if (!((m >> ci) & 1ull)) { if (cell_shift) *cell_shift = s; return 0; }
        const uint32_t rank = (uint32_t)__builtin_popcountll(m & ((1ull << ci) - 1ull));
        if (s == 0) { if (cell_shift) *cell_shift = 0; return (int)(int8_t)t.leaf_types[n.child_base + rank]; }
        idx = n.child_base + rank;
        s -= 2;
    }
}

/* ---- top grid of the closed-form walk ------------------------------------------------------------------------- */
/* every block classified by a descent to the slot of edge 1 << g that is the block: grid = node entry (bit 31) or 0,
 * cell = log2 edge of the aligned empty octree cell around an empty block */
static bool grid_classify(const vr_node *nodes, int levels, int dim, std::vector<uint32_t> &grid, std::vector<uint8_t> &cell, int *g_out, int *bits_out) {
    const int root_shift = 2 * (levels - 1);
    if (root_shift < 2 || dim < 8) return false;
    const int g = vr_grid_shift_for(root_shift, dim);
    const int G = dim >> g;
    if (G < 1) return false;
    int bits = 0;
    while ((1 << bits) < G) bits++;
    grid.assign((size_t)G * G * G, 0u);
    cell.assign((size_t)G * G * G, 0);
    for (int bz = 0; bz < G; bz++)
        for (int by = 0; by < G; by++)
            for (int bx = 0; bx < G; bx++) {
                const int x = bx << g, y = by << g, z = bz << g;
                const size_t i = (size_t)bx + (size_t)G * ((size_t)by + (size_t)G * bz);
                uint32_t idx = 0, entry = 0;
                for (int s = root_shift;; s -= 2) {
                    const vr_node &n = nodes[idx];
                    if (n.child_base & VR_NODE_SOLID) { entry = 0x80000000u | idx; break; }     /* the block lies in a solid node */
                    const unsigned long long m = (unsigned long long)n.mask_lo | ((unsigned long long)n.mask_hi << 32);
                    const int ci = ((x >> s) & 3) | (((y >> s) & 3) << 2) | (((z >> s) & 3) << 4);
                    if (!((m >> ci) & 1ull)) {
                        int cs = s + ((((m >> (ci & 0x2A)) & 0x00330033ull) == 0ull) ? 1 : 0);
                        while ((1 << cs) > dim) cs--;                  /* (a root wider than the map) */
                        cell[i] = (uint8_t)cs;
                        break;
                    }
                    idx = n.child_base + (uint32_t)__builtin_popcountll(m & ((1ull << ci) - 1ull));
                    if (s == g) { entry = 0x80000000u | idx; break; }
                }
                grid[i] = entry;
            }
    *g_out = g;
    *bits_out = bits;
    return true;
}

bool vr_native_grid(const vr_node *nodes, int levels, int dim, std::vector<uint32_t> &grid, int *grid_shift, int *grid_bits) {
    std::vector<uint8_t> cell;
    int g = 0, bits = 0;
    if (!grid_classify(nodes, levels, dim, grid, cell, &g, &bits)) return false;
    const int G = dim >> g;
    /* (2) Chebyshev distance to the nearest non-empty block or to the outside of the map, minus one, by repeated erosion
     * with the 3x3x3 cube: an empty block gets radius r when its 26 neighbours all exist and have radius >= r - 1 */
    for (uint32_t r = 1; r <= VR_GRID_MAX_RADIUS; r++) {
        bool any = false;
        for (int bz = 1; bz < G - 1; bz++)
            for (int by = 1; by < G - 1; by++)
                for (int bx = 1; bx < G - 1; bx++) {
                    const size_t i = (size_t)bx + (size_t)G * ((size_t)by + (size_t)G * bz);
                    if (grid[i] != r - 1) continue;
                    bool ok = true;
                    for (int dz = -1; dz <= 1 && ok; dz++)
                        for (int dy = -1; dy <= 1 && ok; dy++)
                            for (int dx = -1; dx <= 1; dx++) {
                                const uint32_t e = grid[(size_t)(bx + dx) + (size_t)G * ((size_t)(by + dy) + (size_t)G * (bz + dz))];
                                if (e < r - 1 || (e & 0x80000000u)) { ok = false; break; }
                            }
                    if (ok) { grid[i] = r; any = true; }
                }
        if (!any) break;
    }
    /* (3) final form of the empty entries (vr_types.h): the wider of the two empty cells known around the block */
    for (size_t i = 0; i < grid.size(); i++) {
        if (grid[i] & 0x80000000u) continue;
        const uint32_t d = grid[i];
        grid[i] = (((2u * d + 1u) << g) > (1u << cell[i])) ? ((uint32_t)g | ((d << g) << 8)) : (uint32_t)cell[i];
    }
    *grid_shift = g;
    *grid_bits = bits;
    return true;
}

/* ---- directed top grids: one table per direction octant ------------------------------------------------------ */
/* A ray only ever needs the empty space AHEAD of it.  For octant o (bit a set = the ray moves towards lower coordinates
 * on axis a) the entry of an empty block b holds the largest cube of empty blocks that has b in its rear corner and
 * extends in the ray's direction on every axis: E(b) = 1 + min E(b + d), d in {0,1}^3 \ 0 taken along the direction of
 * travel (the 3-D maximal-square recurrence; blocks outside the map count as 0, so a cube never leaves the map), capped
 * at VR_GRID_MAX_CUBE.  It contains the centred cube of the undirected grid (radius r  =>  E >= r + 1) and is far wider
 * for rays that move away from a surface.  Entry format = vr_native_grid's: m = block edge - 1, ext = (E - 1) blocks;
 * the aligned octree cell around the block is kept where it reaches at least as far on every axis and further on one.
 * Table o starts at entry o << (3 * grid_bits). */
bool vr_native_grid_directed(const vr_node *nodes, int levels, int dim, std::vector<uint32_t> &grid, int *grid_shift, int *grid_bits) {
    std::vector<uint32_t> base;
    std::vector<uint8_t> cell;
    int g = 0, bits = 0;
    if (!grid_classify(nodes, levels, dim, base, cell, &g, &bits)) return false;
    const int G = dim >> g;
    if ((1 << bits) != G) return false;                                    /* the octant number is XORed into the key */
    const size_t n3 = (size_t)G * G * G;
    grid.assign(8 * n3, 0u);
#pragma omp parallel for schedule(dynamic, 1)
    for (int o = 0; o < 8; o++) {
        std::vector<uint8_t> E(n3);
        const int sx = (o & 1) ? -1 : 1, sy = (o & 2) ? -1 : 1, sz = (o & 4) ? -1 : 1;
        uint32_t *out = grid.data() + (size_t)o * n3;
        for (int kz = G - 1; kz >= 0; kz--)
            for (int ky = G - 1; ky >= 0; ky--)
                for (int kx = G - 1; kx >= 0; kx--) {                      /* k = mirrored block coordinate, descending */
                    const int bx = sx < 0 ? G - 1 - kx : kx, by = sy < 0 ? G - 1 - ky : ky, bz = sz < 0 ? G - 1 - kz : kz;
                    const size_t i = (size_t)bx + (size_t)G * ((size_t)by + (size_t)G * bz);
                    if (base[i] & 0x80000000u) { E[i] = 0; out[i] = base[i]; continue; }
                    uint32_t mn = VR_GRID_MAX_CUBE;
                    if (kx == G - 1 || ky == G - 1 || kz == G - 1) mn = 0;
                    else
                        for (int d = 1; d < 8; d++) {
                            const size_t j = (size_t)(bx + ((d & 1) ? sx : 0)) + (size_t)G * ((size_t)(by + ((d & 2) ? sy : 0)) + (size_t)G * (bz + ((d & 4) ? sz : 0)));
                            if (E[j] < mn) mn = E[j];
                        }
                    const uint32_t e = mn + 1 > VR_GRID_MAX_CUBE ? VR_GRID_MAX_CUBE : mn + 1;
                    E[i] = (uint8_t)e;
                    out[i] = vr_grid_directed_entry(e, cell[i], g, kx, ky, kz);
                }
    }
    *grid_shift = g;
    *grid_bits = bits;
    return true;
}
