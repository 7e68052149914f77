// rt_headless.cpp -- a headless C++ host for the B200 path tracer: the part of `rt --scene <toml> --renderer <name>`
// (reference src/main.cpp:327-379) that leads to one `render` call, without the SDL window.
//
//   rt_headless --scene scenes/dielectric.toml [--renderer cuda_path_tracer] [--size WxH] [--spp N] [--bounces N]
//               [--mode sm|mg] [--seed N] [--out image.ppm] [--dump-scene] [--dump-view] [--list] [--device N] [--gpus N]
//
// Scene loading restates scene.cpp (scene_loader.hpp); rendering goes through the C ABI of include/rtcu.h exactly like
// plugin/cuda_path_tracer.cpp does.  Errors follow main.cpp:370-379: "error: <what>" on stderr, exit code 1.
// --dump-scene / --dump-view print the flattened scene / the inverse view-projection as JSON and need no GPU.
#include "scene_loader.hpp"

#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace {

const char* const renderer_names[] = { "cuda_path_tracer", "cuda_rasterizer" };

// main.cpp:68-81: exact name, then prefix
const char* find_renderer(const std::string& name)
{
    for (const char* r : renderer_names)
        if (name == r) return r;
    for (const char* r : renderer_names)
        if (!name.empty() && std::string(r).rfind(name, 0) == 0) return r;
    return nullptr;
}

// a float as a JSON number that round-trips (9 significant digits); non-finite values -- a zero-length plane normal normalises to
// NaN in the reference as well -- use the tokens Python's json module reads
std::string json_number(float v)
{
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-Infinity" : "Infinity";
    if (v == 0.0f) return std::signbit(v) ? "-0.0" : "0"; // "-0" would be read back as the integer 0
    char b[48];
    std::snprintf(b, sizeof b, "%.9g", v);
    return b;
}

void dump_scene(const rtb::scene& s)
{
    auto arr = [](const float* v, size_t n) { std::string o = "["; for (size_t i = 0; i < n; i++) { o += i ? ", " : ""; o += json_number(v[i]); } return o + "]"; };
    std::printf("{\"samples_per_pixel\": %u, \"max_bounces\": %u, \"camera\": {\"position\": %s, \"direction\": %s},\n", s.samples_per_pixel, s.max_bounces,
                arr(s.camera.position.data(), 3).c_str(), arr(s.camera.direction.data(), 3).c_str());
    std::printf(" \"materials\": [");
    for (size_t i = 0; i < s.materials.size(); i++)
    {
        const auto& m = s.materials[i];
        std::printf("%s{\"type\": %u, \"albedo\": %s, \"roughness\": %s, \"reflectivity\": %s}", i ? ", " : "", m.type, arr(m.albedo, 4).c_str(),
                    json_number(m.roughness).c_str(), json_number(m.reflectivity).c_str());
    }
    std::printf("],\n \"spheres\": [");
    for (size_t i = 0; i < s.spheres.size(); i++) std::printf("%s%s", i ? ", " : "", arr(s.spheres[i].data(), 4).c_str());
    std::printf("],\n \"sphere_material\": [");
    for (size_t i = 0; i < s.sphere_material.size(); i++) std::printf("%s%u", i ? ", " : "", s.sphere_material[i]);
    std::printf("],\n \"planes\": [");
    for (size_t i = 0; i < s.planes.size(); i++) std::printf("%s%s", i ? ", " : "", arr(s.planes[i].data(), 4).c_str());
    std::printf("],\n \"plane_material\": [");
    for (size_t i = 0; i < s.plane_material.size(); i++) std::printf("%s%u", i ? ", " : "", s.plane_material[i]);
    std::printf("],\n \"boxes\": [");
    for (size_t i = 0; i < s.boxes.size(); i++) std::printf("%s%s", i ? ", " : "", arr(s.boxes[i].data(), 6).c_str());
    std::printf("]}\n");
}

int run(int argc, char** argv)
{
    std::string scene_path, renderer = "cuda_path_tracer", out_path, mode = "sm";
    unsigned width = 800, height = 600; // main.cpp:153
    long spp = -1, bounces = -1;
    unsigned long long seed = 0x5EED;
    int device = 0, gpus = 1;
    bool list = false, want_dump_scene = false, want_dump_view = false;
    for (int i = 1; i < argc; i++)
    {
        const std::string a = argv[i];
        auto value = [&]() -> std::string
        {
            if (i + 1 >= argc) throw std::runtime_error("option '" + a + "' needs a value");
            return argv[++i];
        };
        // a whole-string number in [lo, hi], or an error that names the option (std::stol alone says "stol")
        auto number = [&](long long lo, unsigned long long hi) -> unsigned long long
        {
            const std::string v = value();
            char* end = nullptr;
            errno = 0;
            const bool negative = !v.empty() && v[0] == '-';
            const unsigned long long u = negative ? 0 : std::strtoull(v.c_str(), &end, 0);
            const long long sgn = negative ? std::strtoll(v.c_str(), &end, 0) : 0;
            if (v.empty() || *end || errno == ERANGE || (negative ? sgn < lo : u > hi) || (!negative && lo > 0 && u < (unsigned long long)lo))
                throw std::runtime_error("bad " + a + " '" + v + "' (expected a whole number in [" + std::to_string(lo) + ", " + std::to_string(hi) + "])");
            return negative ? (unsigned long long)sgn : u;
        };
        if (a == "-l" || a == "--list") list = true;
        else if (a == "-s" || a == "--scene") scene_path = value();
        else if (a == "-r" || a == "--renderer") renderer = value();
        else if (a == "--size")
        {
            const std::string v = value();
            if (std::sscanf(v.c_str(), "%ux%u", &width, &height) != 2 || !width || !height) throw std::runtime_error("bad --size '" + v + "' (expected WxH)");
        }
        else if (a == "--spp") spp = (long)number(1, 1u << 30);
        else if (a == "--bounces") bounces = (long)number(1, 1u << 30);
        else if (a == "--mode") mode = value();
        else if (a == "--seed") seed = number(0, ~0ull);
        else if (a == "--out") out_path = value();
        else if (a == "--device") device = (int)number(0, 1023);
        else if (a == "--gpus") gpus = (int)number(1, 8);
        else if (a == "--dump-scene") want_dump_scene = true;
        else if (a == "--dump-view") want_dump_view = true;
        else throw std::runtime_error("unknown option '" + a + "'");
    }
    if (list) // main.cpp:355-360
    {
        for (const char* r : renderer_names) std::printf("%s\n", r);
        return 0;
    }
    const char* found = find_renderer(renderer);
    if (!found) throw std::runtime_error("unknown renderer '" + renderer + "'");
    if (mode != "sm" && mode != "mg") throw std::runtime_error("bad --mode '" + mode + "' (sm or mg)");

    rtb::scene scene = scene_path.empty() ? rtb::load_first_available() : rtb::load_scene(scene_path); // main.cpp:121-125
    // samples_per_pixel / max_bounces are public fields of rt::scene (scene.hpp:10-11): a harness may set them past the
    // loader's clamp (BASELINE config 5 uses 4096 spp)
    if (spp > 0) scene.samples_per_pixel = static_cast<unsigned>(spp);
    if (bounces > 0) scene.max_bounces = static_cast<unsigned>(bounces);
    std::fprintf(stderr, "loaded scene '%s': %zu spheres, %zu planes, %zu boxes, %zu materials\n", scene.path.c_str(), scene.spheres.size(),
                 scene.planes.size(), scene.boxes.size(), scene.materials.size());
    if (want_dump_scene)
    {
        dump_scene(scene);
        return 0;
    }
    rtcu_view v{};
    rtb::inverse_view_projection(scene.camera, width, height, v.inv_view_proj);
    if (want_dump_view)
    {
        std::printf("[");
        for (int i = 0; i < 16; i++) std::printf("%s%.9g", i ? ", " : "", v.inv_view_proj[i]);
        std::printf("]\n");
        return 0;
    }
    v.width = width;
    v.height = height;
    v.samples_per_pixel = scene.samples_per_pixel;
    v.max_bounces = scene.max_bounces;
    v.sample_begin = 0;
    v.sample_end = scene.samples_per_pixel;
    v.tile_x1 = width;
    v.tile_y1 = height;
    v.seed = seed;
    v.material_mode = mode == "mg" ? RTCU_MODE_MG : RTCU_MODE_SM;
    v.flags = static_cast<uint32_t>(RTCU_ACCEL_AUTO) | static_cast<uint32_t>(RTCU_PIPE_AUTO);

    if (gpus < 1 || gpus > 8) throw std::runtime_error("--gpus must be 1..8");
    std::vector<rtcu_ctx*> ctxs;
    struct closer { std::vector<rtcu_ctx*>& c; ~closer() { for (auto* x : c) rtcu_destroy(x); } } guard{ ctxs };
    const rtcu_scene desc = scene.descriptor();
    for (int g = 0; g < gpus; g++)
    {
        rtcu_ctx* ctx = rtcu_create(device + g);
        if (!ctx) throw std::runtime_error(std::string(found) + ": " + rtcu_last_error());
        ctxs.push_back(ctx);
        if (rtcu_upload_scene(ctx, &desc) != RTCU_OK) throw std::runtime_error(std::string(found) + ": " + rtcu_last_error());
    }
    std::vector<uint32_t> pixels(static_cast<size_t>(width) * height, 0x000000FFu); // cleared to black like main.cpp:318
    const bool raster = std::string(found) == "cuda_rasterizer"; // the preview is one ray per pixel: a single device
    const int rc = raster      ? rtcu_rasterize(ctxs[0], &v, pixels.data(), nullptr, nullptr)
                   : gpus == 1 ? rtcu_render(ctxs[0], &v, pixels.data(), nullptr)
                               : rtcu_render_multi(ctxs.data(), static_cast<uint32_t>(gpus), &v, pixels.data(), nullptr);
    if (rc != RTCU_OK) throw std::runtime_error(std::string(found) + ": " + rtcu_last_error());
    rtcu_stats st{};
    rtcu_get_stats(ctxs[0], &st);
    if (raster)
        std::fprintf(stderr, "%s: %ux%u: %.3f ms on the device, %.1f Mpixels/s\n", found, width, height, st.ms_render, double(width) * height / (st.ms_render * 1e3));
    else
        std::fprintf(stderr, "%s: %ux%u, %u spp, depth %u: %.3f ms on the device, %.1f Msamples/s, %llu segments\n", found, width, height, scene.samples_per_pixel,
                     scene.max_bounces, st.ms_render, double(width) * height * scene.samples_per_pixel / (st.ms_render * 1e3), static_cast<unsigned long long>(st.segments));
    if (!out_path.empty())
    {
        FILE* f = std::fopen(out_path.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot write '" + out_path + "'");
        std::fprintf(f, "P6\n%u %u\n255\n", width, height);
        for (const uint32_t px : pixels)
        {
            const unsigned char rgb[3] = { static_cast<unsigned char>(px >> 24), static_cast<unsigned char>(px >> 16), static_cast<unsigned char>(px >> 8) };
            std::fwrite(rgb, 1, 3, f);
        }
        std::fclose(f);
    }
    return 0;
}

} // namespace

int main(int argc, char** argv)
{
    try
    {
        return run(argc, argv);
    }
    catch (const std::exception& e) // main.cpp:370-379
    {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
