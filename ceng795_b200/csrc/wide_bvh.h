// wide_bvh.h — the tree the default traversal walks inside a mesh: 8 children per node, child boxes quantised to
// 8 bits per plane in the node's own frame, 80 bytes per node (host builder side; the device reads the same bytes
// as five float4, csrc/rt_wide.cuh).
//
// Why a second, wider tree.  The reference's own BVH (spatial median, round-robin axis, one primitive per leaf,
// HW7/src/Bounding_volume_hierarchy.cpp:3-29) decides WHICH primitive a ray returns — through the per-triangle gate
// (tri_gate) and the tie-break key, DESIGN.md 4b — but says nothing about HOW candidates are found.  The binary SAH
// tree of csrc/fast_bvh.cpp needs ~35 dependent 64-byte fetches per ray on cfg2; collapsing it 8 wide cuts the
// dependent steps by ~3x and the bytes per ray by ~2x, and a traversal stack of node GROUPS (one entry per level) is
// short enough to live in shared memory (the binary walk kept 2 x 64 words per thread in local memory).
//
// Conservative by construction: a child's quantised box is floor / ceil of the (already grown) box the SAH builder
// stored for it, so it can only be larger.  Nothing here is an accept / reject decision of the reference.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/b200rt.h"

namespace b200rt_wide {

// Two encodings of the child boxes, chosen when the library is compiled (RT_WIDE_BF16; both were measured on the device,
// profiles/r02_tuning.txt).  0: 8-bit planes in a per-node power-of-two grid, 80 bytes.  1: planes as bfloat16 offsets
// from the node's origin, (hi << 16) | lo per slot and axis, 128 bytes = one cache line: `lo` becomes a float by a
// multiply by 65536 and `hi` by using the word as it is (its low half, the `lo` bits, can only enlarge the box), both on
// the FMA pipe — the byte extraction of encoding 0 runs on the ALU pipe, which is the busy one in the traversal kernel.
#ifndef RT_WIDE_BF16
#define RT_WIDE_BF16 0   // measured: no faster (3.59-3.65 against 3.54 ms per cfg2 frame): the ALU pipe drops from 62 to 48 %
                         // busy, but 128-byte nodes and 24 more live registers put it back into long-scoreboard stalls
#endif
#if RT_WIDE_BF16
struct Node {            // 128 bytes = 8 x float4
  float p[3];            // origin of the node's frame (its box minimum)
  uint8_t e[3];          // unused in this encoding
  uint8_t imask;         // bit s set: the child in slot s is an inner node
  uint32_t child_base;   // inner children are stored contiguously in slot order from here
  uint32_t tri_base;     // triangles of the leaf children: [tri_base, tri_base + 24) in the wide triangle order
  uint8_t meta[8];       // 0 empty; inner: 0x20 | (24 + slot); leaf: unary count (1,3,7) << 5 | offset from tri_base
  uint32_t planes[3][8]; // axis, slot: (bf16(max - p) rounded up) << 16 | bf16(min - p) rounded down
};
static_assert(sizeof(Node) == 128, "wide node must be 128 bytes");
#else
struct Node {            // 80 bytes = 5 x float4
  float p[3];            // origin of the node's quantisation frame (its box minimum)
  uint8_t e[3];          // biased exponents: plane = p + q * 2^(e-127)
  uint8_t imask;         // bit s set: the child in slot s is an inner node
  uint32_t child_base;   // inner children are stored contiguously in slot order from here
  uint32_t tri_base;     // triangles of the leaf children: [tri_base, tri_base + 24) in the wide triangle order
  uint8_t meta[8];       // 0 empty; inner: 0x20 | (24 + slot); leaf: unary count (1,3,7) << 5 | offset from tri_base
  uint8_t qlo[3][8];     // x, y, z minimum planes of the 8 slots
  uint8_t qhi[3][8];     // x, y, z maximum planes
};
static_assert(sizeof(Node) == 80, "wide node must be 80 bytes");
#endif

struct Stats {
  int nodes, depth, leaf_children, inner_children;
};

// Collapses the binary tree under `root` (an FNODE / LEAF1..4 ref into bin_nodes / the leaf-order triangle arrays
// fast_tri_geom, 12 floats per triangle) into 8-wide nodes appended to `nodes`.  `tri_order` receives, in the order
// the wide leaves want their triangles, the leaf-order POSITIONS (indices into fast_tri_geom / fast_tri_ref) — the
// caller gathers geometry and back-references with it.  `grow` is added to boxes that have to be recomputed from the
// triangles (halves of a 4-triangle leaf, a mesh whose root is a leaf).  Returns the index of the mesh's root node in
// `nodes`, or -1 when root is NONE.
int build(const b200rt_bvh_node* bin_nodes, uint32_t root, const float* fast_tri_geom, float grow,
          std::vector<Node>& nodes, std::vector<int32_t>& tri_order, Stats* stats);

// The top-level objects as a flat list, each with the innermost box of the reference's gate chain above it: an
// instance's own world box (Mesh_instance::intersect tests it itself, HW3/include/Mesh.h:59-62), for a primitive the
// box of the BVH node that holds it.  Every outer box of the chain contains that one and the slab test is monotone in
// the box, so "innermost gate passes" == "the reference reaches the object".  8 floats per object:
// {min.xyz, as_float(ref)}, {max.xyz, as_float(1 if gated, 0 if entered unconditionally)}.  Returns false when the
// description cannot be flattened (HW1: no tree; a bad node index).
bool flatten_top(const B200SceneDesc* d, std::vector<float>& flat);

}  // namespace b200rt_wide
