// wide_bvh.cpp — collapse of the binary SAH tree into 8-wide nodes with quantised child boxes (see wide_bvh.h).
#include "wide_bvh.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>

namespace b200rt_wide {
namespace {

struct Item {
  uint32_t ref;  // FNODE index into the binary nodes, or LEAF1..4 start into the leaf-order triangle arrays
  float mn[3], mx[3];
};

inline float area(const Item& it) {
  const float dx = it.mx[0] - it.mn[0], dy = it.mx[1] - it.mn[1], dz = it.mx[2] - it.mn[2];
  return 2.f * (dx * dy + dy * dz + dz * dx);
}
inline bool expandable(uint32_t ref) {
  const uint32_t k = B200RT_REF_KIND(ref);
  return k == B200RT_KIND_FNODE || k == B200RT_KIND_LEAF4;  // a 4-triangle leaf is two 2-triangle leaves here
}
inline int leaf_count(uint32_t ref) { return (int)(B200RT_REF_KIND(ref) - B200RT_KIND_LEAF1) + 1; }

// Box of `count` triangles from their records {v0, a1 = v0 - v1, a2 = v0 - v2, n}, grown like the SAH builder's.
Item leaf_item(uint32_t start, int count, const float* geom, float grow) {
  Item it;
  it.ref = B200RT_REF(B200RT_KIND_LEAF1 + (uint32_t)(count - 1), start);
  for (int a = 0; a < 3; a++) { it.mn[a] = INFINITY; it.mx[a] = -INFINITY; }
  for (int k = 0; k < count; k++) {
    const float* g = geom + 12 * (size_t)(start + k);
    for (int a = 0; a < 3; a++) {
      const float v0 = g[a], v1 = g[a] - g[3 + a], v2 = g[a] - g[6 + a];
      it.mn[a] = std::min(it.mn[a], std::min(v0, std::min(v1, v2)));
      it.mx[a] = std::max(it.mx[a], std::max(v0, std::max(v1, v2)));
    }
  }
  for (int a = 0; a < 3; a++) { it.mn[a] -= grow; it.mx[a] += grow; }
  return it;
}

void expand(const Item& it, const b200rt_bvh_node* bin, const float* geom, float grow, Item out[2]) {
  const uint32_t idx = B200RT_REF_INDEX(it.ref);
  if (B200RT_REF_KIND(it.ref) == B200RT_KIND_FNODE) {
    const b200rt_bvh_node& n = bin[idx];
    out[0].ref = n.left; out[1].ref = n.right;
    for (int a = 0; a < 3; a++) {
      out[0].mn[a] = n.lmin[a]; out[0].mx[a] = n.lmax[a];
      out[1].mn[a] = n.rmin[a]; out[1].mx[a] = n.rmax[a];
    }
  } else {  // LEAF4
    out[0] = leaf_item(idx, 2, geom, grow);
    out[1] = leaf_item(idx + 2, 2, geom, grow);
  }
}

struct Work {
  uint32_t node;
  Item item;
  int depth;
};

}  // namespace

int build(const b200rt_bvh_node* bin, uint32_t root, const float* geom, float grow, std::vector<Node>& nodes,
          std::vector<int32_t>& tri_order, Stats* stats) {
  const uint32_t rk = B200RT_REF_KIND(root);
  if (rk == B200RT_KIND_NONE) return -1;
  Item root_item;
  root_item.ref = root;
  if (rk >= B200RT_KIND_LEAF1 && rk <= B200RT_KIND_LEAF4) root_item = leaf_item(B200RT_REF_INDEX(root), leaf_count(root), geom, grow);
  else for (int a = 0; a < 3; a++) { root_item.mn[a] = 0.f; root_item.mx[a] = 0.f; }  // an inner node's own box is the union of its children
  const int root_index = (int)nodes.size();
  nodes.push_back(Node());
  std::vector<Work> work;
  work.push_back(Work{(uint32_t)root_index, root_item, 1});
  int max_depth = 0, n_leaf = 0, n_inner = 0;
  while (!work.empty()) {
    const Work w = work.back();
    work.pop_back();
    max_depth = std::max(max_depth, w.depth);
    Item ch[8];
    int nc = 0;
    if (expandable(w.item.ref)) { expand(w.item, bin, geom, grow, ch); nc = 2; }
    else { ch[0] = w.item; nc = 1; }
    while (nc < 8) {
      int best = -1;
      float best_area = -1.f;
      for (int i = 0; i < nc; i++)
        if (expandable(ch[i].ref)) {
          const float a = area(ch[i]);
          if (a > best_area) { best_area = a; best = i; }
        }
      if (best < 0) break;
      Item two[2];
      expand(ch[best], bin, geom, grow, two);
      ch[best] = two[0];
      ch[nc++] = two[1];
    }
    // the node's frame: union of the children
    float lo[3], hi[3];
    for (int a = 0; a < 3; a++) {
      lo[a] = ch[0].mn[a]; hi[a] = ch[0].mx[a];
      for (int i = 1; i < nc; i++) { lo[a] = std::min(lo[a], ch[i].mn[a]); hi[a] = std::max(hi[a], ch[i].mx[a]); }
    }
    // slots: child c goes where its offset from the node centre points (bit 2 = +x, bit 1 = +y, bit 0 = +z), so that a
    // ray visits slots in the order (slot ^ octant) front to back without sorting — greedy assignment by best gain
    int slot_of[8], child_in[8];
    for (int i = 0; i < 8; i++) { slot_of[i] = -1; child_in[i] = -1; }
    float cost[8][8];
    for (int c = 0; c < nc; c++)
      for (int s = 0; s < 8; s++) {
        float v = 0.f;
        for (int a = 0; a < 3; a++) {
          const float off = 0.5f * (ch[c].mn[a] + ch[c].mx[a]) - 0.5f * (lo[a] + hi[a]);
          v += ((s >> (2 - a)) & 1) ? off : -off;
        }
        cost[c][s] = v;
      }
    for (int round = 0; round < nc; round++) {
      int bc = -1, bs = -1;
      float bv = -INFINITY;
      for (int c = 0; c < nc; c++) {
        if (slot_of[c] >= 0) continue;
        for (int s = 0; s < 8; s++) {
          if (child_in[s] >= 0) continue;
          if (cost[c][s] > bv) { bv = cost[c][s]; bc = c; bs = s; }
        }
      }
      slot_of[bc] = bs; child_in[bs] = bc;
    }
    Node nd;
    std::memset(&nd, 0, sizeof(nd));
#if RT_WIDE_BF16
    for (int a = 0; a < 3; a++) nd.p[a] = lo[a];
    // bfloat16 of a non-negative float: the upper half of its bits; truncation rounds down, +0xffff first rounds up
    auto bf_to_f = [](uint32_t h) { const uint32_t u = h << 16; float f; std::memcpy(&f, &u, 4); return f; };
    auto f_bits = [](float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; };
#else
    double scale[3];
    for (int a = 0; a < 3; a++) {
      nd.p[a] = lo[a];
      const double ext = (double)hi[a] - (double)lo[a];
      int e = ext > 0.0 ? (int)std::ceil(std::log2(ext / 255.0)) : -100;
      e = std::max(-100, std::min(100, e));
      while (std::ceil(ext / std::ldexp(1.0, e)) > 255.0) e++;
      nd.e[a] = (uint8_t)(e + 127);
      scale[a] = std::ldexp(1.0, e);
    }
#endif
    nd.child_base = (uint32_t)nodes.size();
    nd.tri_base = (uint32_t)tri_order.size();
    for (int s = 0; s < 8; s++) {
      const int c = child_in[s];
#if RT_WIDE_BF16
      if (c < 0) {  // empty slot: meta 0, and a box that is inside out (min = 3.4e38, max = 0)
        for (int a = 0; a < 3; a++) nd.planes[a][s] = 0x00007f7fu;
        continue;
      }
      for (int a = 0; a < 3; a++) {
        const double vl = std::max(0.0, (double)ch[c].mn[a] - (double)lo[a]);
        const double vh = std::max(0.0, (double)ch[c].mx[a] - (double)lo[a]);
        uint32_t ql = f_bits((float)vl) >> 16;
        uint32_t qh = (f_bits((float)vh) + 0xffffu) >> 16;
        // (float)v may have rounded the wrong way: step until the decoded plane is on the safe side of the exact one
        while (ql > 0u && (double)lo[a] + (double)bf_to_f(ql) > (double)ch[c].mn[a]) ql--;
        while (qh < 0x7f7fu && (double)lo[a] + (double)bf_to_f(qh) < (double)ch[c].mx[a]) qh++;
        nd.planes[a][s] = (qh << 16) | ql;
      }
#else
      if (c < 0) {  // empty slot: meta 0, and a box that is inside out
        for (int a = 0; a < 3; a++) { nd.qlo[a][s] = 255; nd.qhi[a][s] = 0; }
        continue;
      }
      for (int a = 0; a < 3; a++) {
        double ql = std::floor(((double)ch[c].mn[a] - (double)lo[a]) / scale[a]);
        double qh = std::ceil(((double)ch[c].mx[a] - (double)lo[a]) / scale[a]);
        ql = std::max(0.0, std::min(255.0, ql));
        qh = std::max(0.0, std::min(255.0, qh));
        while (ql > 0.0 && (double)lo[a] + ql * scale[a] > (double)ch[c].mn[a]) ql -= 1.0;
        while (qh < 255.0 && (double)lo[a] + qh * scale[a] < (double)ch[c].mx[a]) qh += 1.0;
        nd.qlo[a][s] = (uint8_t)ql; nd.qhi[a][s] = (uint8_t)qh;
      }
#endif
      if (expandable(ch[c].ref)) {
        nd.imask |= (uint8_t)(1u << s);
        nd.meta[s] = (uint8_t)(0x20u | (24u + (uint32_t)s));
        const uint32_t ci = (uint32_t)nodes.size();
        nodes.push_back(Node());
        work.push_back(Work{ci, ch[c], w.depth + 1});
        n_inner++;
      } else {
        const int cnt = leaf_count(ch[c].ref);   // 1..3
        const uint32_t off = (uint32_t)tri_order.size() - nd.tri_base;
        const uint32_t start = B200RT_REF_INDEX(ch[c].ref);
        for (int k = 0; k < cnt; k++) tri_order.push_back((int32_t)(start + (uint32_t)k));
        nd.meta[s] = (uint8_t)((((1u << cnt) - 1u) << 5) | off);
        n_leaf++;
      }
    }
    nodes[w.node] = nd;
  }
  if (stats) {
    stats->nodes += (int)nodes.size() - root_index;
    stats->depth = std::max(stats->depth, max_depth);
    stats->leaf_children += n_leaf;
    stats->inner_children += n_inner;
  }
  return root_index;
}

bool flatten_top(const B200SceneDesc* d, std::vector<float>& flat) {
  flat.clear();
  bool ok = d->semantics != B200RT_SEM_HW1;
  std::function<void(uint32_t, const float*, bool)> walk = [&](uint32_t ref, const float* box, bool own) {
    const uint32_t kind = B200RT_REF_KIND(ref);
    if (kind == B200RT_KIND_NONE || !ok) return;
    if (kind == B200RT_KIND_INNER) {
      const uint32_t idx = B200RT_REF_INDEX(ref);
      if ((int)idx >= d->n_nodes) { ok = false; return; }
      const b200rt_bvh_node& n = d->nodes[idx];
      const uint32_t ch[2] = {n.left, n.right};
      const float* cb[2] = {n.lmin, n.rmin};  // lmin,lmax / rmin,rmax are contiguous: 6 floats each
      for (int c = 0; c < 2; c++) {
        const uint32_t ck = B200RT_REF_KIND(ch[c]);
        if (ck == B200RT_KIND_INNER || ck == B200RT_KIND_INSTANCE) walk(ch[c], cb[c], true);  // its own box, as stored in the parent
        else walk(ch[c], box, own);                                                          // primitives: the node's own gate
      }
      return;
    }
    float rec[8] = {box ? box[0] : 0.f, box ? box[1] : 0.f, box ? box[2] : 0.f, 0.f,
                    box ? box[3] : 0.f, box ? box[4] : 0.f, box ? box[5] : 0.f, 0.f};
    std::memcpy(&rec[3], &ref, 4);
    const uint32_t gated = (own && box) ? 1u : 0u;
    std::memcpy(&rec[7], &gated, 4);
    flat.insert(flat.end(), rec, rec + 8);
  };
  if (ok) {
    const uint32_t rk = B200RT_REF_KIND(d->root);
    // BVH::intersect and Mesh_instance::intersect test the root's box; a lone primitive has no gate at all
    walk(d->root, d->root_box, rk == B200RT_KIND_INNER || rk == B200RT_KIND_INSTANCE);
  }
  return ok;
}

}  // namespace b200rt_wide
