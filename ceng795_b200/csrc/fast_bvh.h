// fast_bvh.h — device-oriented BVH over a mesh's triangles (host builder side).
//
// The reference's own tree (spatial median, round-robin axis, one primitive per
// leaf — HW7/src/Bounding_volume_hierarchy.cpp:3-29) decides WHICH primitive a
// ray returns (reachability gates + DFS tie-break), but it is a poor structure
// to traverse.  This builder makes a binned-SAH binary BVH over the same
// triangles (<= 4 per leaf, triangles re-stored in leaf order) that the CUDA
// traversal walks instead; every candidate it finds is then admitted only if
// the reference's innermost gate for that triangle passes, so
// the answer is the reference's.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/b200rt.h"

namespace b200rt_fast {

// boxes6: n tight triangle boxes (min xyz, max xyz).  Triangle i of the mesh is
// global triangle tri_base + i.  `grow` is added to every stored child box on
// all sides (conservative).  Appends nodes to `nodes`, global triangle ids in
// leaf order to `order`.  Returns the root child ref (FNODE or LEAF1..4), or
// B200RT_REF_NONE when n == 0.
uint32_t build(const float* boxes6, int n, int tri_base, float grow, std::vector<b200rt_bvh_node>& nodes,
               std::vector<int32_t>& order);

}  // namespace b200rt_fast
