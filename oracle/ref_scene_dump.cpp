// oracle/ref_scene_dump.cpp -- TEST INFRASTRUCTURE (pins the oracle; never part of the product).
// Driver around the reference's OWN scene front end, compiled from /root/reference where it lies (oracle/Makefile
// target `ref` builds OR/OptixModel.cpp, OR/HalfSphere.cpp, OR/Sphere.cpp unmodified, against oracle/ref_stubs/ for
// the absent OptiX / Win32 / GL headers): loadOBJ (OR/OptixModel.cpp:75-151) and, when a receiver is given,
// placeReceiver / place_receiver_half (OR/OptixModel.cpp:153-257, glm::rotate + mat4 * vec4).
//
//   ref_scene_dump scene.obj [leftHalf.obj rightHalf.obj cam_x cam_y cam_z rotation_deg]
//
// Prints, per TriangleMesh of the model in model->meshes order: "mesh <material_name>" and one line
// "t x1 y1 z1 x2 y2 z2 x3 y3 z3" per triangle (mesh->vertex[mesh->index[i]]), floats as hex bit patterns.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "OptixModel.h"

// the order OR/OptixModel.cpp:10-33 defines for tinyobj::index_t (see oracle/ref_stubs/ref_prelude.h)
namespace tinyobj {
bool operator<(const index_t& a, const index_t& b)
{
    if (a.vertex_index != b.vertex_index) return a.vertex_index < b.vertex_index;
    if (a.normal_index != b.normal_index) return a.normal_index < b.normal_index;
    return a.texcoord_index < b.texcoord_index;
}
}

static unsigned bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv)
{
    if (argc != 2 && argc != 8) { fprintf(stderr, "usage: ref_scene_dump scene.obj [left.obj right.obj x y z rotation_deg]\n"); return 2; }
    try {
        OptixModel* model = loadOBJ(argv[1]);
        if (argc == 8) {
            HalfSphere left(argv[2]), right(argv[3]);
            Sphere sphere(&left, &right);
            placeReceiver(sphere, model, vec3f((float)atof(argv[4]), (float)atof(argv[5]), (float)atof(argv[6])), (float)atof(argv[7]));
        }
        for (const TriangleMesh* mesh : model->meshes) {
            printf("mesh %s\n", mesh->material_name.c_str());
            for (const vec3i& idx : mesh->index) {
                printf("t");
                const int v[3] = {idx.x, idx.y, idx.z};
                for (int k = 0; k < 3; ++k) {
                    const vec3f& p = mesh->vertex[v[k]];
                    printf(" %08x %08x %08x", bits(p.x), bits(p.y), bits(p.z));
                }
                printf("\n");
            }
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
