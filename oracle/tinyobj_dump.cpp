// Shim around the reference's vendored tiny_obj_loader.h (compiled from
// /root/reference where it lies; see Makefile target `ref`).  Prints the flat
// triangle list exactly as OR/OptixModel.cpp:75-151 (loadOBJ) would hand it to
// the renderer: one mesh per (shape, material id) in std::set order, faces in
// file order.  Output: "mesh <name>\n" then "t x1 y1 z1 x2 y2 z2 x3 y3 z3" with
// floats printed as hex bit patterns.  Used to generate tests/golden/*.mesh.
#define TINYOBJLOADER_IMPLEMENTATION
#include "tiny_obj_loader.h"
#include <cstdio>
#include <cstring>
#include <set>
#include <string>

static unsigned bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: tinyobj_dump file.obj\n"); return 2; }
    const std::string objFile = argv[1];
    const std::string mtlDir = objFile.substr(0, objFile.rfind('/') + 1);
    tinyobj::attrib_t attributes;
    std::vector<tinyobj::shape_t> shapes;
    std::vector<tinyobj::material_t> materials;
    std::string err;
    bool ok = tinyobj::LoadObj(&attributes, &shapes, &materials, &err, &err, objFile.c_str(), mtlDir.c_str(), true);
    if (!ok) { fprintf(stderr, "load failed: %s\n", err.c_str()); return 1; }
    printf("materials %zu\n", materials.size());
    for (auto& m : materials) printf("material %s\n", m.name.c_str());
    for (size_t s = 0; s < shapes.size(); ++s) {
        auto& shape = shapes[s];
        std::set<int> ids(shape.mesh.material_ids.begin(), shape.mesh.material_ids.end());
        for (int id : ids) {
            std::string name = id >= 0 ? materials[id].name : std::string();
            printf("mesh %s\n", name.c_str());
            for (size_t f = 0; f < shape.mesh.material_ids.size(); ++f) {
                if (shape.mesh.material_ids[f] != id) continue;
                printf("t");
                for (int k = 0; k < 3; ++k) {
                    int vi = shape.mesh.indices[3 * f + k].vertex_index;
                    for (int a = 0; a < 3; ++a) printf(" %08x", bits(attributes.vertices[3 * vi + a]));
                }
                printf("\n");
            }
        }
    }
    return 0;
}
