// arv2_internal.h -- types shared by the host front end, the BVH builders, the CUDA
// kernels and the C ABI of libarv2.so.  Not part of the public interface.
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <climits>
#include <string>
#include <vector>

#include "../../include/arv2.h"

namespace arv2 {

struct Vec3 { float x, y, z; };

// struct TriangleMesh / OptixModel (OR/OptixModel.h:9-32) flattened: the renderer
// only ever needs positions, the mesh a triangle belongs to and the mesh's
// material name.
struct HostScene {
    std::vector<float> tri_verts;          // [T][3][3]
    std::vector<int32_t> tri_mesh;         // [T]
    std::vector<std::string> mesh_material;
    std::vector<std::string> mtl_names;    // materials of the MTL, file order
    int64_t n_tris() const { return (int64_t)tri_mesh.size(); }
};

struct HostReceiver {
    std::vector<float> left, right;        // [n][3][3] untransformed templates
};

// ---- device BVH layout --------------------------------------------------------
// Binary BVH, 64 B per inner node (4 x float4), children boxes stored in the
// parent so one node fetch decides both descents:
//   q0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)
//   q1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//   q2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//   q3 = (int c0, int c1, -, -)   c >= 0: inner node index
//                                 c <  0: leaf, ~c = (first_tri << 3) | (count - 1)
struct BvhNode { float q[16]; };
constexpr int kMaxLeafTris = 4;
constexpr int kLeafShift = 3;
// An absent child is a degenerate box far outside any scene: with the min/max slab
// test an inverted (+inf,-inf) box would read as "everything", this one as a miss.
constexpr float kEmptyBox = 3.0e38f;

struct HostBvh {
    std::vector<BvhNode> nodes;     // node 0 is the root (always an inner node)
    std::vector<int32_t> order;     // leaf-order position -> input triangle index
    float lo[3], hi[3];             // root bounds (padded)
};

// Binned-SAH top-down builder (host). `tri_verts` = [n][3][3].
void build_bvh_sah(const float* tri_verts, int64_t n, HostBvh* out, int n_threads);
// Recompute every box of an existing tree for moved vertices (same topology/order).
void refit_bvh(const float* tri_verts, int64_t n, HostBvh* bvh);
// Conservative padding applied to every box (relative to the scene extent).
inline float bvh_pad(float extent) { return extent * 1e-5f + 1e-6f; }

// ---- device layout -------------------------------------------------------------------------
// Nodes: BvhNode as above, [0] = two-level top node (child 0 = scene tree at node 1, child 1 =
// receiver tree), so a receiver move rewrites only the receiver's ~85 KB.
// Triangles, 64 B = two 256-bit loads, in leaf order:
//   (P1, id) (P2, material) (P3, Ng.x) (Ng.y, Ng.z, 0, 0)
// material >= 0 = mesh index into keep[mesh][band], -1 / -2 = receiver ears; Ng = the unit
// normal of the arithmetic contract, precomputed once on the host.
// 4-wide node (experiment, host/bvh4.cpp): 128 B = 8 x float4: lo.x[4] hi.x[4] lo.y[4] hi.y[4] lo.z[4] hi.z[4],
// child codes[4] (>= 0: kWideBit | wide node index, < 0: leaf as above), packed split axes (a0 | aL << 2 | aR << 4).
struct Bvh4Node { float q[32]; };
constexpr int32_t kWideBit = 1 << 29;
// returns the depth of the wide tree (0 for an empty input)
int collapse_bvh4(const HostBvh& bvh2, std::vector<Bvh4Node>* out);
// Quantised binary node (experiment, -DARV2_QNODES=1): 32 B = one sector.  w[3*child + axis] = lo | hi << 16, two 15-bit
// planes on a scene-wide grid (plane = qc + (2^23 + 256 q) * qk); child codes as in Bvh4Node.  Node i mirrors binary node i.
struct Q16Node { uint32_t w[6]; int32_t child[2]; };
void quantise_bvh2(const HostBvh& bvh2, std::vector<Q16Node>* out, float qk[3], float qinvk[3], float qc[3]);
constexpr int kTraversalStack = 64;      // per-lane stack entries of the kernels (trace.cu)
int bvh2_depth(const HostBvh& bvh2);
void make_tri_record(const float* v9, int32_t id, int32_t material, float* out16);

// ---- host front end -------------------------------------------------------------
int load_obj(const std::string& path, HostScene* out, std::string* err);
void place_receiver_half(const std::vector<float>& tmpl, const float cam[3], float rotation_deg,
                         float* out);
float material_absorption(const std::string& name, const arv2_material* mats, int n);
int parse_config(const std::string& json, arv2_config* out, std::string* err);

int wav_read(const std::string& path, float** samples, size_t* n, int32_t* rate, int32_t* channels, std::string* err);
int wav_write_stereo_normalized(const std::string& path, const float* l, const float* r, size_t n, int32_t rate, std::string* err);

void set_error(const std::string& msg);

} // namespace arv2
