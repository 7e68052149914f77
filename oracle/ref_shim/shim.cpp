// shim.cpp -- glue that lets the reference's own renderer sources run headless and deterministically.  TEST
// INFRASTRUCTURE (oracle/_ref): compiled together with /root/reference/src/renderers/{mg,sm}_ray_tracer.cpp and
// src/renderer.cpp, which are used where they lie and never copied.
//
//  * rt::detail::random_float() (declared in the reference's random.hpp:27-28, defined in random.cpp which is NOT
//    compiled) is replaced by the counter-based stream of SPEC S9: Philox4x32-7, counter (pixel, sample, block, retry).
//    The reference calls it without any context, so the stand-in math library reports the call sites that delimit
//    the coordinates: thread_pool::for_range (pixel), vec3::operator+= in the per-pixel worker (sample end),
//    transform_position(depth 1) (primary ray built -> block 1), ray::at after draws (a scatter event closed).
//  * refbin_render builds an rt::scene through the reference's own soagen tables and camera, looks the renderer up in
//    the reference's registry by name and calls renderer_interface::render.
#include "scene.hpp"
#include "image.hpp"
#include "colour.hpp"
#include "random.hpp"
#include "renderer.hpp"
#include <muu/thread_pool.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <memory>
#include <string>

extern "C" void rtref_philox_stream(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]); // oracle/rtref.c
extern "C" float rtref_u01(uint32_t x);

namespace
{
	struct rng_state
	{
		uint32_t pixel = 0, sample = 0, block = 0, draws = 0;
		// one Philox block serves the three draws of an event: cache it (keeps the CPU timing arm fair)
		uint32_t cached_ctr[4] = { ~0u, ~0u, ~0u, ~0u };
		uint32_t cached_out[4] = {};
	};
	thread_local rng_state g_rng;
	uint64_t g_seed = 0;
	char g_last_error[512] = "";
	uint32_t g_sample_begin = 0;
}

namespace muu::shim
{
	unsigned row_step = 1, row_width = 0;
	void on_pixel_begin(unsigned pixel_index) noexcept
	{
		g_rng.pixel = pixel_index;
		g_rng.sample = g_sample_begin;
		g_rng.block = 0;
		g_rng.draws = 0;
	}
	void on_sample_end() noexcept
	{
		g_rng.sample++;
		g_rng.block = 0;
		g_rng.draws = 0;
	}
	void on_primary_done() noexcept
	{
		g_rng.block = 1;
		g_rng.draws = 0;
	}
	void on_ray_at() noexcept
	{
		if (g_rng.draws)
		{
			g_rng.block++;
			g_rng.draws = 0;
		}
	}
}

namespace rt::detail
{
	// draw n of the current event: lane n % 3 of block (pixel, sample, block, retry = n / 3) -- a unit-vector redraw
	// (random.hpp:57-66) consumes the next retry block; the jitter (2 draws) and the dielectric draw (1) use retry 0
#ifndef REFBIN_REFERENCE_RNG // the `_mt` flavour links the reference's own src/random.cpp (thread_local mt19937) instead
	float random_float() noexcept
	{
		const uint32_t n = g_rng.draws++;
		const uint32_t ctr[4] = { g_rng.pixel, g_rng.sample, g_rng.block, n / 3u };
		if (std::memcmp(ctr, g_rng.cached_ctr, sizeof ctr) != 0)
		{
			const uint32_t key[2] = { static_cast<uint32_t>(g_seed), static_cast<uint32_t>(g_seed >> 32) };
			rtref_philox_stream(ctr, key, g_rng.cached_out);
			std::memcpy(g_rng.cached_ctr, ctr, sizeof ctr);
		}
		return rtref_u01(g_rng.cached_out[n % 3u]);
	}
#endif
}

extern "C"
{
	struct refbin_material
	{
		uint32_t type;
		float albedo[4];
		float roughness;
		float reflectivity;
	};
	struct refbin_scene
	{
		const float* spheres;
		const uint32_t* sphere_material;
		uint32_t n_spheres;
		const float* planes;
		const uint32_t* plane_material;
		uint32_t n_planes;
		const refbin_material* materials;
		uint32_t n_materials;
		const float* boxes; // n x {cx,cy,cz,ex,ey,ez}; only the rasterizer tests them
		const uint32_t* box_material;
		uint32_t n_boxes;
	};

	const char* refbin_last_error() { return g_last_error; }

	// names of the renderers registered by the compiled reference sources, '\n'-separated
	int refbin_list(char* buf, uint32_t size)
	{
		std::string s;
		for (const auto& r : rt::renderers::all())
		{
			s += std::string{ r.name };
			s += '\n';
		}
		if (size)
		{
			std::strncpy(buf, s.c_str(), size - 1);
			buf[size - 1] = 0;
		}
		return static_cast<int>(rt::renderers::all().size());
	}

	// Renders with the reference's renderer `name` ("mg_ray_tracer" / "sm_ray_tracer").  Returns 0, or -1 when the
	// renderer is unknown.  inv_view_proj_out receives the matrix the reference's camera produced (column-major) so the
	// oracle and the CUDA path can be driven with exactly the same input.
	static void fill_scene(rt::scene& scene, const refbin_scene* sc, const float cam_pos[3], const float cam_dir[3], uint32_t spp, uint32_t max_bounces);

	// A renderer that lives across frames, as in the application (main.cpp:49-53: one instance per --renderer choice): what a
	// plugin caches between calls (device buffers, the page-locked image, the uploaded scene) is only visible this way.
	void* refbin_open(const char* name)
	{
		const auto* desc = rt::renderers::find_by_name(name);
		if (!desc)
		{
			std::snprintf(g_last_error, sizeof g_last_error, "no renderer named %s", name);
			return nullptr;
		}
		try
		{
			return desc->create();
		}
		catch (const std::exception& e)
		{
			std::snprintf(g_last_error, sizeof g_last_error, "%s", e.what());
			return nullptr;
		}
	}
	void refbin_close(void* renderer) { delete static_cast<rt::renderer_interface*>(renderer); }
	// one frame into the caller's buffer (as back_buffer's image would be: any host memory, 64-byte aligned or not)
	int refbin_render_with(void* renderer, const refbin_scene* sc, const float cam_pos[3], const float cam_dir[3], uint32_t width, uint32_t height,
						   uint32_t spp, uint32_t max_bounces, uint64_t seed, uint32_t* rgba8, int threads)
	{
		if (!renderer || !rgba8)
			return -1;
		rt::scene scene;
		fill_scene(scene, sc, cam_pos, cam_dir, spp, max_bounces);
		g_seed				 = seed;
		g_sample_begin		 = 0;
		muu::shim::row_step	 = 1;
		muu::shim::row_width = width;
		rt::image_view pixels{ rgba8, rt::vec2u{ width, height } };
		muu::thread_pool pool{ threads > 0 ? static_cast<unsigned>(threads) : 0u };
		static_cast<rt::renderer_interface*>(renderer)->render(scene, pixels, pool);
		return 0;
	}

	int refbin_render(const refbin_scene* sc, const float cam_pos[3], const float cam_dir[3], uint32_t width, uint32_t height, uint32_t spp,
					  uint32_t max_bounces, uint64_t seed, const char* name, uint32_t* rgba8, int threads, uint32_t row_step,
					  float inv_view_proj_out[16])
	{
		const auto* desc = rt::renderers::find_by_name(name);
		if (!desc)
			return -1;
		rt::scene scene;
		fill_scene(scene, sc, cam_pos, cam_dir, spp, max_bounces);
		if (inv_view_proj_out)
		{
			const auto view = scene.camera.viewport(rt::vec2u{ width, height });
			for (int c = 0; c < 4; c++)
				for (int r = 0; r < 4; r++)
					inv_view_proj_out[c * 4 + r] = view.inverse_view_projection(static_cast<size_t>(r), static_cast<size_t>(c));
		}
		if (!rgba8)
			return 0;
		g_seed					= seed;
		g_sample_begin			= 0;
		muu::shim::row_step		= row_step ? row_step : 1;
		muu::shim::row_width	= width;
		std::unique_ptr<rt::renderer_interface> renderer;
		try
		{
			renderer.reset(desc->create()); // a plugin constructor may throw (main.cpp:329-379 catches it in the real app)
		}
		catch (const std::exception& e)
		{
			std::snprintf(g_last_error, sizeof g_last_error, "%s", e.what());
			return -2;
		}
		rt::image_view pixels{ rgba8, rt::vec2u{ width, height } };
		muu::thread_pool pool{ threads > 0 ? static_cast<unsigned>(threads) : 0u };
		renderer->render(scene, pixels, pool);
		return 0;
	}

	static void fill_scene(rt::scene& scene, const refbin_scene* sc, const float cam_pos[3], const float cam_dir[3], uint32_t spp, uint32_t max_bounces)
	{
		scene.samples_per_pixel = spp;
		scene.max_bounces		= max_bounces;
		scene.camera.pose(rt::vec3{ cam_pos[0], cam_pos[1], cam_pos[2] }, rt::vec3{ cam_dir[0], cam_dir[1], cam_dir[2] });
		for (uint32_t i = 0; i < sc->n_materials; i++)
		{
			const auto& m = sc->materials[i];
			rt::colour albedo;
			std::memcpy(albedo.values, m.albedo, sizeof(float) * 4);
			scene.materials.push_back(std::string{}, static_cast<rt::material_type>(m.type), albedo, m.roughness, m.reflectivity);
		}
		for (uint32_t i = 0; i < sc->n_planes; i++)
		{
			const float* p = sc->planes + 4 * i;
			const rt::plane pl{ rt::vec3{ p[0], p[1], p[2] }, p[3] };
			scene.planes.push_back(pl, sc->plane_material[i], p[0], p[1], p[2], p[3]);
		}
		for (uint32_t i = 0; i < sc->n_spheres; i++)
		{
			const float* s = sc->spheres + 4 * i;
			const rt::sphere sp{ rt::vec3{ s[0], s[1], s[2] }, s[3] };
			scene.spheres.push_back(sp, sc->sphere_material[i], s[0], s[1], s[2], s[3]);
		}
		for (uint32_t i = 0; i < sc->n_boxes; i++)
		{
			const float* b = sc->boxes + 6 * i;
			const rt::box bx{ rt::vec3{ b[0], b[1], b[2] }, rt::vec3{ b[3], b[4], b[5] } };
			scene.boxes.push_back(bx, sc->box_material[i], b[0], b[1], b[2], b[3], b[4], b[5]);
		}
	}
}
