// scene_loader.hpp -- C++ host side: scene containers + TOML scene loader + viewport matrix, feeding the C ABI.
//
// Restates the reference loader `scene::load` (reference src/scene.cpp:483-618), the containers of src/scene.hpp:8-25 /
// src/soa.toml and `camera::viewport` (src/camera.hpp:122-137) with the same defaults, clamps, aliases and error
// messages, without toml++ / muu (neither is available here).  The Python twin is rt_b200/scene.py + camera.py; the two
// are checked against each other (tests/test_host_cpp.py).  Quirks kept on purpose: named colours are binarised
// (colour.hpp:72-98), colour arrays start from zero with alpha 1 (scene.cpp:347-356), spp / bounces clamp to [1,1000]
// (scene.cpp:531-532), dielectric-class materials keep their IOR in `reflectivity` (scene.cpp:546-556).
#pragma once
#include "toml_lite.hpp"
#include <rtcu.h>

#include <array>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace rtb {

struct camera {
    std::array<float, 3> position{ 0.0f, 1.0f, 0.0f };    // camera.hpp:55
    std::array<float, 3> direction{ 0.0f, 0.0f, -1.0f };  // forward = -Z
    double vfov = 3.14159265358979323846 / 4.0;           // camera.hpp:54 (private in the reference, no setter)
    double near_clip = 0.01, far_clip = 1000.0;           // camera.hpp:57-58
};

// rt::scene with the soagen tables as plain columns (src/soa.toml)
struct scene {
    unsigned samples_per_pixel = 30; // scene.hpp:10
    unsigned max_bounces = 10;       // scene.hpp:11
    std::string path;
    rtb::camera camera;
    std::vector<rtcu_material> materials;
    std::vector<std::string> material_names;
    std::vector<std::array<float, 4>> spheres; // spheres.value(): {cx,cy,cz,radius}
    std::vector<uint32_t> sphere_material;
    std::vector<std::array<float, 4>> planes;  // planes.value(): {nx,ny,nz,d}
    std::vector<uint32_t> plane_material;
    std::vector<std::array<float, 6>> boxes;   // centre, extents (drawn by the rasterizer only; never hit by the ray tracers, mg_ray_tracer.cpp:89-93)
    std::vector<uint32_t> box_material;

    rtcu_scene descriptor() const
    {
        rtcu_scene d{};
        d.spheres = spheres.empty() ? nullptr : spheres[0].data();
        d.sphere_material = sphere_material.data();
        d.n_spheres = static_cast<uint32_t>(spheres.size());
        d.planes = planes.empty() ? nullptr : planes[0].data();
        d.plane_material = plane_material.data();
        d.n_planes = static_cast<uint32_t>(planes.size());
        d.materials = materials.data();
        d.n_materials = static_cast<uint32_t>(materials.size());
        d.boxes = boxes.empty() ? nullptr : boxes[0].data();
        d.box_material = box_material.data();
        d.n_boxes = static_cast<uint32_t>(boxes.size());
        return d;
    }
};

struct scene_error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

namespace detail {

using toml_lite::node;

inline const char* const material_type_names[8] = { "lambert", "metal", "dielectric", "air", "vacuum", "water", "ice", "diamond" };

struct named_colour { const char* name; uint32_t rgb; };
inline const named_colour colour_table[] = {
#include "colour_table.inc"
};

// colours::<name> as the reference constructs it: colour{uint32} clamps the *byte value* to [0,1] (colour.hpp:72-98)
inline std::array<float, 4> colour_by_name(const std::string& name)
{
    for (const auto& c : colour_table)
        if (name == c.name)
        {
            auto comp = [](uint32_t byte) { return byte ? 1.0f : 0.0f; };
            return { comp((c.rgb >> 16) & 0xFF), comp((c.rgb >> 8) & 0xFF), comp(c.rgb & 0xFF), 1.0f };
        }
    throw scene_error("unknown colour alias '" + name + "'");
}

// `node.value<float>()` + the infinity / NaN check of scene.cpp:89-102.  toml++ (un-vendored: subprojects/tomlplusplus.wrap @
// f1a38d23) converts permissively; its rules as recalled, UNVERIFIED like the muu arithmetic: an integer converts when it lies in
// [-2^24, 2^24] (the whole numbers a float holds exactly), a finite float when it lies inside float's range, inf / nan pass through
// (and are then refused by the reference's own check); booleans, strings ... have no mapping.
inline float finite_float(const node& n, const char* what)
{
    auto no_mapping = [&]() { return scene_error(std::string("No mapping from TOML ") + n.type_name() + " to float (" + what + ")"); };
    if (n.kind == node::integer)
    {
        if (n.i < -(int64_t{ 1 } << 24) || n.i > (int64_t{ 1 } << 24)) throw no_mapping();
        return static_cast<float>(n.i);
    }
    if (n.kind != node::floating) throw no_mapping();
    if (std::isnan(n.f) || std::isinf(n.f)) throw scene_error("Infinities and NaNs are not allowed.");
    if (n.f < -static_cast<double>(FLT_MAX) || n.f > static_cast<double>(FLT_MAX)) throw no_mapping();
    return static_cast<float>(n.f);
}

struct alias { const char* name; float v[3]; };
inline const alias vector_aliases[] = { // scene.cpp:113-144; right-handed, forward = -Z
    { "origin", { 0, 0, 0 } }, { "zero", { 0, 0, 0 } }, { "one", { 1, 1, 1 } }, { "forward", { 0, 0, -1 } }, { "back", { 0, 0, 1 } },
    { "backward", { 0, 0, 1 } }, { "up", { 0, 1, 0 } }, { "down", { 0, -1, 0 } }, { "left", { -1, 0, 0 } }, { "right", { 1, 0, 0 } },
    { "x", { 1, 0, 0 } }, { "x_axis", { 1, 0, 0 } }, { "y", { 0, 1, 0 } }, { "y_axis", { 0, 1, 0 } }, { "z", { 0, 0, 1 } }, { "z_axis", { 0, 0, 1 } },
};

// scene.cpp:113-166: alias string, scalar broadcast, array of <= 3 with missing components keeping the default
inline std::array<float, 3> vector3(const node* n, std::array<float, 3> def, const char* what)
{
    if (!n) return def;
    if (n->kind == node::string)
    {
        for (const auto& a : vector_aliases)
            if (n->s == a.name) return { a.v[0], a.v[1], a.v[2] };
        throw scene_error("unknown vector alias '" + n->s + "'");
    }
    if (n->kind == node::integer || n->kind == node::floating)
    {
        const float s = n->kind == node::integer ? static_cast<float>(n->i) : static_cast<float>(n->f);
        return { s, s, s };
    }
    if (n->kind != node::array || n->items.size() > 3)
        throw scene_error(std::string("No mapping from TOML ") + n->type_name() + " to vector<float, 3> (" + what + ")");
    for (size_t i = 0; i < n->items.size(); i++) def[i] = finite_float(*n->items[i], what);
    return def;
}

// scene.cpp:184-356
inline std::array<float, 4> colour(const node* n, std::array<float, 4> def)
{
    if (!n) return def;
    if (n->kind == node::string) return colour_by_name(n->s);
    if (n->kind != node::array || n->items.size() > 4) throw scene_error(std::string("No mapping from TOML ") + n->type_name() + " to colour");
    std::array<float, 4> out{ 0, 0, 0, 0 };
    for (size_t i = 0; i < n->items.size(); i++) out[i] = finite_float(*n->items[i], "colour");
    if (n->items.size() < 4) out[3] = 1.0f;
    return out;
}

// `node.value<unsigned>()` (toml++, rules as recalled, UNVERIFIED): an integer inside [0, UINT_MAX] (out-of-range values have no
// mapping -- they do not wrap), a float holding a whole number inside that range, a boolean as 0 / 1
inline unsigned unsigned_value(const node* n, unsigned def, const char* what)
{
    if (!n) return def;
    auto no_mapping = [&]() { return scene_error(std::string("No mapping from TOML ") + n->type_name() + " to unsigned (" + what + ")"); };
    int64_t v;
    if (n->kind == node::integer) v = n->i;
    else if (n->kind == node::boolean) v = n->b ? 1 : 0;
    else if (n->kind == node::floating)
    {
        if (!std::isfinite(n->f) || n->f < -9.2e18 || n->f > 9.2e18 || static_cast<double>(static_cast<int64_t>(n->f)) != n->f) throw no_mapping();
        v = static_cast<int64_t>(n->f);
    }
    else throw no_mapping();
    if (v < 0 || v > static_cast<int64_t>(UINT32_MAX)) throw no_mapping();
    return static_cast<unsigned>(v);
}

// scene.cpp:381-404 (magic_enum by integer or by name)
inline uint32_t material_type(const node* n)
{
    if (!n) return RTCU_LAMBERT;
    if (n->kind == node::integer)
    {
        if (n->i < 0 || n->i > 7) throw scene_error("integer value " + std::to_string(n->i) + " was not a member of enum material_type");
        return static_cast<uint32_t>(n->i);
    }
    if (n->kind == node::string)
    {
        for (uint32_t t = 0; t < 8; t++)
            if (n->s == material_type_names[t]) return t;
        throw scene_error("string value '" + n->s + "' was not a member of enum material_type");
    }
    throw scene_error(std::string("No mapping from TOML ") + n->type_name() + " to material_type");
}

inline const std::vector<toml_lite::node_ptr>& table_array(const node& cfg, const char* key)
{
    static const std::vector<toml_lite::node_ptr> empty;
    const node* n = cfg.get(key);
    if (!n) return empty;
    if (n->kind != node::array) throw scene_error(std::string("expected array at key '") + key + "', got " + n->type_name());
    // an element that is not a table reads as an empty one, as in the reference: its lookups go through toml::node_view's
    // operator[] (scene.cpp:420-430), which yields an empty view for a non-table parent, so every field takes its default
    // (node::get finds nothing in a non-table node either)
    return n->items;
}

} // namespace detail

// scene::load on TOML text (scene.cpp:527-618)
inline scene load_scene_text(const std::string& text, const std::string& path = "")
{
    using namespace detail;
    toml_lite::node_ptr root;
    try
    {
        root = toml_lite::parse(text);
    }
    catch (const toml_lite::parse_error& e)
    {
        throw scene_error(e.what());
    }
    const node& cfg = *root;
    scene s;
    s.path = path;
    auto clamp_u = [](unsigned v) { return v < 1u ? 1u : (v > 1000u ? 1000u : v); };
    s.samples_per_pixel = clamp_u(unsigned_value(cfg.get("samples_per_pixel"), 30u, "samples_per_pixel"));
    s.max_bounces = clamp_u(unsigned_value(cfg.get("max_bounces"), 10u, "max_bounces"));

    if (const node* cam = cfg.get("camera"))
    {
        if (cam->kind != node::table) throw scene_error(std::string("expected table at key 'camera', got ") + cam->type_name());
        s.camera.position = vector3(cam->get("position"), { 0, 1, 0 }, "camera.position");
        s.camera.direction = vector3(cam->get("direction"), { 0, 0, -1 }, "camera.direction");
    }

    for (const auto& t : table_array(cfg, "materials"))
    {
        rtcu_material m{};
        m.type = material_type(t->get("type"));
        float reflectiveness;
        switch (m.type) // scene.cpp:546-556
        {
            case RTCU_METAL: reflectiveness = 0.8f; break;
            case RTCU_DIELECTRIC: reflectiveness = 1.52f; break;
            case RTCU_AIR: reflectiveness = 1.000293f; break;
            case RTCU_VACUUM: reflectiveness = 1.0f; break;
            case RTCU_ICE: reflectiveness = 1.31f; break;
            case RTCU_WATER: reflectiveness = 1.333f; break;
            default: reflectiveness = 0.5f;
        }
        std::string name;
        if (const node* n = t->get("name"))
        {
            if (n->kind != node::string) throw scene_error("No mapping from TOML value to string (name)");
            name = n->s;
        }
        const auto albedo = colour(t->get("albedo"), colour_by_name("fuchsia"));
        std::memcpy(m.albedo, albedo.data(), sizeof m.albedo);
        m.roughness = t->get("roughness") ? finite_float(*t->get("roughness"), "roughness") : (m.type == RTCU_DIELECTRIC ? 0.0f : 0.5f);
        m.reflectivity = t->get("reflectivity") ? finite_float(*t->get("reflectivity"), "reflectivity") : reflectiveness;
        s.materials.push_back(m);
        s.material_names.push_back(name);
    }
    if (s.materials.empty()) // scene.cpp:565-566
    {
        rtcu_material m{};
        m.type = RTCU_LAMBERT;
        const auto c = colour_by_name("fuchsia");
        std::memcpy(m.albedo, c.data(), sizeof m.albedo);
        m.roughness = 0.05f;
        m.reflectivity = 0.5f;
        s.materials.push_back(m);
        s.material_names.emplace_back();
    }
    auto material_of = [&](const node& t) -> uint32_t // scene.cpp:568-574
    {
        const unsigned m = unsigned_value(t.get("material"), 0u, "material");
        if (m >= s.materials.size()) throw scene_error("material index " + std::to_string(m) + " out-of-range");
        return m;
    };
    for (const auto& t : table_array(cfg, "planes")) // scene.cpp:576-585
    {
        const auto pos = vector3(t->get("position"), { 0, 0, 0 }, "plane.position");
        auto n = vector3(t->get("normal"), { 0, 1, 0 }, "plane.normal");
        // DESIGN.md SPEC S1 / S2 (the arithmetic the muu stand-in, the oracle and the kernels share): dot3(a, b) = fma(a.z, b.z,
        // fma(a.y, b.y, a.x * b.x)), normalize(v) = v * (1 / sqrt(dot3(v, v))); a zero normal becomes NaN, as in the reference
        auto dot3 = [](const std::array<float, 3>& a, const std::array<float, 3>& b) {
            const volatile float xx = a[0] * b[0]; // one IEEE multiply, never contracted into the fma below
            return std::fmaf(a[2], b[2], std::fmaf(a[1], b[1], xx));
        };
        const volatile float len = std::sqrt(dot3(n, n));
        const volatile float inv = 1.0f / len;
        for (auto& c : n)
        {
            const volatile float scaled = c * inv;
            c = scaled;
        }
        const float d = -dot3(n, pos); // muu plane{position, normal}: dot(n,p) + d == 0
        s.planes.push_back({ n[0], n[1], n[2], d });
        s.plane_material.push_back(material_of(*t));
    }
    for (const auto& t : table_array(cfg, "spheres")) // scene.cpp:587-597
    {
        const auto pos = vector3(t->get("position"), { 0, 1, -3 }, "sphere.position");
        const float radius = t->get("radius") ? finite_float(*t->get("radius"), "radius") : 0.5f;
        s.spheres.push_back({ pos[0], pos[1], pos[2], radius });
        s.sphere_material.push_back(material_of(*t));
    }
    for (const auto& t : table_array(cfg, "boxes")) // scene.cpp:599-615
    {
        const auto pos = vector3(t->get("position"), { 0, 1, -3 }, "box.position");
        const auto ext = vector3(t->get("extents"), { 0.5f, 0.5f, 0.5f }, "box.extents");
        s.boxes.push_back({ pos[0], pos[1], pos[2], ext[0], ext[1], ext[2] });
        s.box_material.push_back(material_of(*t));
    }
    return s;
}

// scene::load(file) with the relative-path search of scene.cpp:479-525
// scene::load (scene.cpp:483-525): "-" reads standard input; a relative path is searched under the six prefixes; only regular
// files count
inline scene load_scene(const std::string& path)
{
    namespace fs = std::filesystem;
    if (path.empty()) throw scene_error("no scene file path provided");
    if (path == "-")
    {
        std::ostringstream ss;
        ss << std::cin.rdbuf();
        return load_scene_text(ss.str(), "");
    }
    static const char* const prefixes[] = { "scenes/", "../scenes/", "../../scenes/", "", "../", "../../" };
    std::vector<fs::path> candidates;
    if (fs::path(path).is_absolute()) candidates.emplace_back(path);
    else
        for (const char* p : prefixes) candidates.push_back(*p ? fs::path(p) / path : fs::path(path));
    for (const auto& c : candidates)
    {
        std::error_code ec;
        if (!fs::is_regular_file(c, ec)) continue;
        std::ifstream f(c, std::ios::binary);
        if (!f) continue;
        std::ostringstream ss;
        ss << f.rdbuf();
        return load_scene_text(ss.str(), c.string());
    }
    throw scene_error("scene path '" + path + "' did not exist or was not a file");
}

// scene::load_first_available (scene.cpp:620-643; what the app loads when --scene is not given, main.cpp:121-125): the first
// regular *.toml file of the first search directory that has one, in the directory's own iteration order
inline scene load_first_available()
{
    namespace fs = std::filesystem;
    static const char* const prefixes[] = { "scenes/", "../scenes/", "../../scenes/", "", "../", "../../" };
    for (const char* p : prefixes)
    {
        std::error_code ec;
        const fs::path dir(p); // the empty prefix is not a directory (fs::status("") is not_found): the working directory is not searched
        if (!fs::is_directory(dir, ec)) continue;
        for (fs::directory_iterator it(dir, ec), end; !ec && it != end; it.increment(ec))
        {
            const fs::path& file = it->path();
            if (!file.has_stem() || !file.has_extension() || file.extension() != ".toml") continue;
            if (!fs::is_regular_file(file, ec)) continue;
            return load_scene(file.string());
        }
    }
    throw scene_error("no scene files found");
}

// ---- camera::viewport -> inverse_view_projection, column-major (camera.hpp:122-137); same construction as rt_b200/camera.py:
// right-handed, forward = -Z, depth 0..1, computed in double and rounded once
inline void inverse_view_projection(const camera& cam, unsigned width, unsigned height, float out[16])
{
    auto norm = [](double v[3]) { const double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); for (int i = 0; i < 3; i++) v[i] /= l; };
    auto cross = [](const double a[3], const double b[3], double r[3]) { r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0]; };
    double f[3] = { cam.direction[0], cam.direction[1], cam.direction[2] };
    norm(f);
    const double back[3] = { -f[0], -f[1], -f[2] };
    double up[3] = { 0, 1, 0 };
    if (std::fabs(f[1]) >= 0.9999) { up[0] = 0; up[1] = 0; up[2] = f[1] < 0 ? 1 : -1; }
    double right[3], up2[3];
    cross(up, back, right);
    norm(right);
    cross(back, right, up2);
    // world = translate(pos) * rot; view = inverse(world); VP = P * view; result = inverse(VP) = world * inverse(P)
    double world[4][4] = { { right[0], up2[0], back[0], cam.position[0] }, { right[1], up2[1], back[1], cam.position[1] },
                           { right[2], up2[2], back[2], cam.position[2] }, { 0, 0, 0, 1 } };
    const double t = 1.0 / std::tan(cam.vfov / 2.0), aspect = double(width) / double(height), n = cam.near_clip, fa = cam.far_clip;
    const double a = fa / (n - fa), b = n * fa / (n - fa);
    // P = [[t/aspect,0,0,0],[0,t,0,0],[0,0,a,b],[0,0,-1,0]];  P^-1 = [[aspect/t,0,0,0],[0,1/t,0,0],[0,0,0,-1],[0,0,1/b,a/b]]
    const double pinv[4][4] = { { aspect / t, 0, 0, 0 }, { 0, 1.0 / t, 0, 0 }, { 0, 0, 0, -1 }, { 0, 0, 1.0 / b, a / b } };
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++)
        {
            double acc = 0;
            for (int k = 0; k < 4; k++) acc += world[r][k] * pinv[k][c];
            out[c * 4 + r] = static_cast<float>(acc);
        }
}

} // namespace rtb
