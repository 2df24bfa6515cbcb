// scene.cpp -- host front end of libarv2: OBJ/MTL scene loading, receiver placement
// and the material-name -> absorption rule.
//
// Behavioural contract (what, not how) taken from the reference:
//   loadOBJ             OR/OptixModel.cpp:75-151   one mesh per (shape, material id),
//                                                   ids ascending, faces in file order
//   tinyobjloader 2.0rc prebuild/common/3rdParty/tiny_obj_loader.h:2134-2640 (shape /
//                       group / usemtl flushing), :1385-1593 (polygon ear clipping)
//   placeReceiver       OR/OptixModel.cpp:153-257
//   getMaterialAbsorption OR/AudioRenderer.cpp:34-56
// The output is pinned bit-for-bit against the reference's own code compiled where it lies: the loader against
// tiny_obj_loader.h and loadOBJ (tests/test_host_cpu.py::test_cpp_obj_loader_matches_reference_tinyobj, tests/golden/
// meshes.json, ref_loadobj.json), the placement against placeReceiver (tests/test_pin_cpu.py, tests/golden/ref_placement.npz).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <set>
#include <sstream>

#include "../arv2_internal.h"

namespace arv2 {

namespace {

struct Face { std::vector<int> v; };
struct ShapeTri { int a, b, c, mat; };

bool is_blank(char c) { return c == ' ' || c == '\t'; }

// MTL: a material is emitted on `newmtl` only when the previous one was named; the
// trailing material is always emitted, so an empty file yields one unnamed entry.
bool read_mtl_names(const std::string& path, std::vector<std::string>* names)
{
    std::ifstream in(path);
    if (!in) return false;
    std::string line, cur;
    while (std::getline(in, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
        size_t p = line.find_first_not_of(" \t");
        if (p == std::string::npos) continue;
        if (line.compare(p, 6, "newmtl") == 0 && line.size() > p + 6 && is_blank(line[p + 6])) {
            if (!cur.empty()) names->push_back(cur);
            cur = line.substr(p + 7);
            size_t a = cur.find_first_not_of(" \t");
            size_t b = cur.find_last_not_of(" \t");
            cur = (a == std::string::npos) ? std::string() : cur.substr(a, b - a + 1);
        }
    }
    names->push_back(cur);
    return true;
}

bool inside_tri2d(const float vx[3], const float vy[3], float tx, float ty)
{
    bool c = false;
    for (int i = 0, j = 2; i < 3; j = i++) {
        if ((vy[i] > ty) != (vy[j] > ty)) {
            const float xi = (vx[j] - vx[i]) * (ty - vy[i]) / (vy[j] - vy[i]) + vx[i];
            if (tx < xi) c = !c;
        }
    }
    return c;
}

// Ear clipping in the 2-D projection that drops the dominant normal axis of the
// first non-degenerate corner; float arithmetic throughout (tinyobj real_t = float).
void triangulate(const Face& f, const std::vector<float>& v, int mat, std::vector<ShapeTri>* out)
{
    const size_t n = f.v.size();
    if (n < 3) return;
    if (n == 3) { out->push_back({f.v[0], f.v[1], f.v[2], mat}); return; }
    auto P = [&](int id, int ax) { return v[(size_t)id * 3 + ax]; };
    int ax0 = 1, ax1 = 2;
    const float eps = std::numeric_limits<float>::epsilon();
    for (size_t k = 0; k < n; ++k) {
        const int i0 = f.v[k % n], i1 = f.v[(k + 1) % n], i2 = f.v[(k + 2) % n];
        const float e0x = P(i1, 0) - P(i0, 0), e0y = P(i1, 1) - P(i0, 1), e0z = P(i1, 2) - P(i0, 2);
        const float e1x = P(i2, 0) - P(i1, 0), e1y = P(i2, 1) - P(i1, 1), e1z = P(i2, 2) - P(i1, 2);
        const float cx = std::fabs(e0y * e1z - e0z * e1y);
        const float cy = std::fabs(e0z * e1x - e0x * e1z);
        const float cz = std::fabs(e0x * e1y - e0y * e1x);
        if (cx > eps || cy > eps || cz > eps) {
            if (!(cx > cy && cx > cz)) {
                ax0 = 0;
                if (cz > cx && cz > cy) ax1 = 1;
            }
            break;
        }
    }
    float area = 0.f;
    for (size_t k = 0; k < n; ++k) {
        const int i0 = f.v[k], i1 = f.v[(k + 1) % n];
        area += (P(i0, ax0) * P(i1, ax1) - P(i0, ax1) * P(i1, ax0)) * 0.5f;
    }
    std::vector<int> rem = f.v;
    size_t guess = 0, budget = n, last = n;
    while (rem.size() > 3 && budget > 0) {
        const size_t m = rem.size();
        if (guess >= m) guess -= m;
        if (last != m) { last = m; budget = m; } else { --budget; }
        int ind[3];
        float vx[3], vy[3];
        for (int k = 0; k < 3; ++k) {
            ind[k] = rem[(guess + k) % m];
            vx[k] = P(ind[k], ax0);
            vy[k] = P(ind[k], ax1);
        }
        const float e0x = vx[1] - vx[0], e0y = vy[1] - vy[0];
        const float e1x = vx[2] - vx[1], e1y = vy[2] - vy[1];
        const float cr = e0x * e1y - e0y * e1x;
        if (cr * area < 0.f) { ++guess; continue; }       // reflex corner
        bool blocked = false;
        for (size_t o = 3; o < m && !blocked; ++o) {
            const int id = rem[(guess + o) % m];
            blocked = inside_tri2d(vx, vy, P(id, ax0), P(id, ax1));
        }
        if (blocked) { ++guess; continue; }
        out->push_back({ind[0], ind[1], ind[2], mat});
        rem.erase(rem.begin() + (long)((guess + 1) % m));
    }
    if (rem.size() == 3) out->push_back({rem[0], rem[1], rem[2], mat});
}

} // namespace

int load_obj(const std::string& path, HostScene* out, std::string* err)
{
    std::ifstream in(path);
    if (!in) { *err = "Could not read OBJ model from " + path; return ARV2_ERR_IO; }
    const std::string dir = path.substr(0, path.rfind('/') + 1);

    std::vector<float> v;
    std::vector<std::string> mtl;
    std::map<std::string, int> mtl_index;
    std::vector<std::vector<ShapeTri>> shapes;
    std::vector<ShapeTri> shape;
    std::vector<Face> pending;
    int material = -1;

    auto flush = [&]() {
        const bool any = !pending.empty();
        for (const Face& f : pending) triangulate(f, v, material, &shape);
        return any;
    };

    std::string line;
    while (std::getline(in, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
        const size_t p = line.find_first_not_of(" \t");
        if (p == std::string::npos) continue;
        const char* tok = line.c_str() + p;
        if (tok[0] == '#') continue;
        if (tok[0] == 'v' && is_blank(tok[1])) {
            const char* s = tok + 2;
            float xyz[3] = {0.f, 0.f, 0.f};
            for (int k = 0; k < 3; ++k) {
                char* e = nullptr;
                const double d = std::strtod(s, &e);
                if (e == s) break;
                xyz[k] = (float)d;
                s = e;
            }
            v.insert(v.end(), xyz, xyz + 3);
        } else if (tok[0] == 'f' && is_blank(tok[1])) {
            Face f;
            const char* s = tok + 2;
            for (;;) {
                while (is_blank(*s)) ++s;
                if (*s == '\0') break;
                char* e = nullptr;
                long idx = std::strtol(s, &e, 10);
                if (e == s || idx == 0) { *err = "Failed parse `f' line in " + path; return ARV2_ERR_IO; }
                const int nv = (int)(v.size() / 3);
                f.v.push_back(idx > 0 ? (int)idx - 1 : nv + (int)idx);
                s = e;
                while (*s && !is_blank(*s)) ++s;    // skip /vt/vn
            }
            for (int id : f.v)
                if (id < 0 || (size_t)id * 3 + 2 >= v.size()) { *err = "face index out of range in " + path; return ARV2_ERR_IO; }
            pending.push_back(std::move(f));
        } else if (std::strncmp(tok, "usemtl", 6) == 0 && is_blank(tok[6])) {
            const std::string name(tok + 7);
            auto it = mtl_index.find(name);
            const int id = it == mtl_index.end() ? -1 : it->second;
            if (id != material) { flush(); pending.clear(); material = id; }
        } else if (std::strncmp(tok, "mtllib", 6) == 0 && is_blank(tok[6])) {
            std::stringstream ss(tok + 7);
            std::string fn;
            while (std::getline(ss, fn, ' ')) {
                if (fn.empty()) continue;
                std::vector<std::string> names;
                if (read_mtl_names(dir + fn, &names)) {
                    for (const std::string& nm : names) {
                        mtl_index.insert({nm, (int)mtl.size()});
                        mtl.push_back(nm);
                    }
                    break;
                }
            }
        } else if (tok[0] == 'g' && is_blank(tok[1])) {
            flush();
            if (!shape.empty()) shapes.push_back(shape);
            shape.clear(); pending.clear();
        } else if (tok[0] == 'o' && is_blank(tok[1])) {
            if (flush()) shapes.push_back(shape);
            shape.clear(); pending.clear();
        }
    }
    if (flush() || !shape.empty()) shapes.push_back(shape);

    if (mtl.empty()) { *err = "could not parse materials ..."; return ARV2_ERR_IO; }

    out->tri_verts.clear(); out->tri_mesh.clear(); out->mesh_material.clear();
    out->mtl_names = mtl;
    for (const auto& shp : shapes) {
        std::set<int> ids;
        for (const ShapeTri& t : shp) ids.insert(t.mat);
        for (int id : ids) {
            const int mesh = (int)out->mesh_material.size();
            bool any = false;
            for (const ShapeTri& t : shp) {
                if (t.mat != id) continue;
                any = true;
                for (int vi : {t.a, t.b, t.c})
                    out->tri_verts.insert(out->tri_verts.end(), v.begin() + (size_t)vi * 3, v.begin() + (size_t)vi * 3 + 3);
                out->tri_mesh.push_back(mesh);
            }
            if (any) out->mesh_material.push_back(id >= 0 ? mtl[id] : std::string());
        }
    }
    return ARV2_OK;
}

// v' = cam + M v, M = rotation about +Y by -rotation_deg, evaluated in the order of a
// column-major mat4 * vec4 product: (col0*x + col1*y) + (col2*z + col3*1).
// cos/sin are taken in fp64 and narrowed (arithmetic contract).
void place_receiver_half(const std::vector<float>& tmpl, const float cam[3], float rotation_deg, float* out)
{
    const float rad = rotation_deg * 0.01745329251994329576923690768489f;
    const float a = -rad;
    const float c = (float)std::cos((double)a);
    const float s = (float)std::sin((double)a);
    const float k = c + (1.0f - c);
    const float zero = 0.0f;
    const size_t nv = tmpl.size() / 3;
    for (size_t i = 0; i < nv; ++i) {
        const float x = tmpl[3 * i], y = tmpl[3 * i + 1], z = tmpl[3 * i + 2];
        const float xr = (c * x + zero * y) + (s * z + zero);
        const float yr = (zero * x + k * y) + (zero * z + zero);
        const float zr = ((-s) * x + zero * y) + (c * z + zero);
        out[3 * i] = cam[0] + xr;
        out[3 * i + 1] = cam[1] + yr;
        out[3 * i + 2] = cam[2] + zr;
    }
}

float material_absorption(const std::string& name, const arv2_material* mats, int n)
{
    if (name == "receiver_left") return -1.f;
    if (name == "receiver_right") return -2.f;
    for (int i = 0; i < n; ++i)
        if (mats[i].name && name == mats[i].name) return mats[i].mat_absorption[0];
    return 0.5f;
}

} // namespace arv2
