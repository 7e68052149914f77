// cuda_path_tracer.cpp -- renderer plugin for marzer/rt: drop this file into src/renderers/, add it to
// src/renderers/meson.build next to mg_ray_tracer.cpp, and link librtcu.so.
//
// `rt --scene <toml> --renderer cuda_path_tracer` then renders through the B200 library instead of the CPU
// loops of mg_ray_tracer.cpp:178-205 / sm_ray_tracer.cpp:263-289, and `--renderer cuda_rasterizer` replaces the
// preview of rasterizer.cpp:22-88.  The plugin only flattens rt::scene into the POD descriptor of include/rtcu.h and
// calls the C ABI; scene.hpp, image.cpp and the window stay untouched.
//
// Behaviour at the boundary (reference src/renderer.hpp:9-14, src/main.cpp:315-321):
//  * render() is noexcept and has no error channel: on failure it logs to stderr and leaves the image as the
//    caller cleared it (black, main.cpp:318).
//  * the constructor runs inside create() under main's try/catch (main.cpp:329-379), so it may throw.
//  * rt::scene has no dirty flag and is replaced wholesale on reload (main.cpp:123-125): the scene columns are
//    compared by content against the last upload and re-uploaded when they differ.
//  * RT_CUDA_MATERIAL_MODE=mg selects mg_ray_tracer's scatter table (no dielectric); default is sm_ray_tracer's.
//  * every B200 of the box is used (RT_CUDA_DEVICES=n caps it): a frame with at least 16 samples per pixel and device is split
//    by sample range over the devices (rtcu_render_multi: one context per device, the fp32 sums added over NVLink inside the
//    resolve kernel of device 0); smaller frames, and boxes without peer access, render on device 0.
//  * RT_CUDA_PROGRESSIVE=1: a frame whose scene, camera, size and sampling equal the previous frame's is not traced again from
//    sample 0 but REFINED -- the next samples_per_pixel samples are added to the fp32 sums the library keeps on the device and
//    the image is resolved over all samples so far (RTCU_FLAG_ACCUMULATE).  Off by default: the reference's app renders only
//    when something changed (window.cpp:213-217), and a drop-in must return the same image for the same call.
//  * image_view memory is pageable (image.cpp:9-13) and lives as long as the window's back buffer: the plugin opts in to
//    rtcu_set_output_pinning, so the library page-locks it once per (pointer, size) and writes later frames into it directly --
//    no staging copy follows the kernels.
#ifdef RTCU_PLUGIN_STUB_CHECK
	#include "rt_stub.hpp" // minimal stand-ins for the accessors used below (compile check without muu)
#else
	#include "../scene.hpp"
	#include "../image.hpp"
	#include "../colour.hpp"
	#include "../renderer.hpp"
MUU_DISABLE_WARNINGS;
	#include <muu/thread_pool.h>
MUU_ENABLE_WARNINGS;
#endif

#include <rtcu.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

using namespace rt;

namespace
{
	// what both renderers share: the context, the scene upload and the view
	rtcu_stats g_last_stats{}; // of the most recent frame drawn by any renderer of this file (rt_cuda_last_stats)

	struct cuda_renderer : renderer_interface
	{
		rtcu_ctx* ctx_ = nullptr;	  // device 0: every frame's destination
		std::vector<rtcu_ctx*> ctxs_; // all devices in use, ctxs_[0] == ctx_
		bool multi_ok_ = true;		  // cleared when rtcu_render_multi fails (no peer access): device 0 alone from then on

		// last uploaded scene columns (content comparison, see header comment)
		std::vector<float> spheres_, planes_, boxes_;
		std::vector<uint32_t> sphere_mat_, plane_mat_, box_mat_;
		std::vector<rtcu_material> materials_;
		bool uploaded_ = false;
		uint32_t scene_uploads_ = 0, uploads_seen_ = 0; // counts scene changes (a progressive render restarts on one)

		explicit cuda_renderer(const char* name, bool all_devices = false)
		{
			ctx_ = rtcu_create(0);
			if (!ctx_)
				throw std::runtime_error{ std::string{ name } + ": " + rtcu_last_error() };
			ctxs_.push_back(ctx_);
			// the application's image lives as long as its back buffer and arrives frame after frame: let the library page-lock
			// it once per (pointer, size) and write into it directly (rtcu.h, rtcu_set_output_pinning)
			rtcu_set_output_pinning(ctx_, 1);
			int devices = all_devices ? rtcu_device_count() : 1;
			if (const char* cap = std::getenv("RT_CUDA_DEVICES"); cap && std::atoi(cap) >= 1)
				devices = std::min(devices, std::atoi(cap));
			for (int d = 1; d < std::min(devices, 8); d++)
			{
				rtcu_ctx* c = rtcu_create(d);
				if (!c) // a device that cannot be used is not an error: render on the ones before it
				{
					std::fprintf(stderr, "%s: device %d not used: %s\n", name, d, rtcu_last_error());
					break;
				}
				ctxs_.push_back(c);
			}
		}

		~cuda_renderer() noexcept override
		{
			for (rtcu_ctx* c : ctxs_)
				rtcu_destroy(c);
		}

		template <typename T>
		static bool assign_if_changed(std::vector<T>& cache, const T* src, size_t count)
		{
			if (cache.size() == count && (count == 0 || std::memcmp(cache.data(), src, count * sizeof(T)) == 0))
				return false;
			cache.assign(src, src + count);
			return true;
		}

		bool sync_scene(const rt::scene& scene) noexcept
		{
			// spheres.value() is muu::bounding_sphere<float>[] = {center.xyz, radius}: already a float4 array
			static_assert(sizeof(scene.spheres.value()[0]) == 4 * sizeof(float));
			static_assert(sizeof(scene.planes.value()[0]) == 4 * sizeof(float));

			bool changed = !uploaded_;
			changed |= assign_if_changed(spheres_, reinterpret_cast<const float*>(scene.spheres.value()), scene.spheres.size() * 4);
			changed |= assign_if_changed(sphere_mat_, scene.spheres.material(), scene.spheres.size());
			changed |= assign_if_changed(planes_, reinterpret_cast<const float*>(scene.planes.value()), scene.planes.size() * 4);
			changed |= assign_if_changed(plane_mat_, scene.planes.material(), scene.planes.size());
			// boxes.value() is muu::bounding_box<float>[] = {center.xyz, extents.xyz}; only the rasterizer draws them
			static_assert(sizeof(scene.boxes.value()[0]) == 6 * sizeof(float));
			changed |= assign_if_changed(boxes_, reinterpret_cast<const float*>(scene.boxes.value()), scene.boxes.size() * 6);
			changed |= assign_if_changed(box_mat_, scene.boxes.material(), scene.boxes.size());

			std::vector<rtcu_material> mats(scene.materials.size());
			for (size_t i = 0; i < mats.size(); i++)
			{
				mats[i].type = static_cast<uint32_t>(scene.materials.type()[i]);
				std::memcpy(mats[i].albedo, &scene.materials.albedo()[i], sizeof(float) * 4);
				mats[i].roughness	 = scene.materials.roughness()[i];
				mats[i].reflectivity = scene.materials.reflectivity()[i];
			}
			changed |= assign_if_changed(materials_, mats.data(), mats.size());
			if (!changed)
				return true;

			rtcu_scene desc{};
			desc.spheres		 = spheres_.data();
			desc.sphere_material = sphere_mat_.data();
			desc.n_spheres		 = static_cast<uint32_t>(sphere_mat_.size());
			desc.planes			 = planes_.data();
			desc.plane_material	 = plane_mat_.data();
			desc.n_planes		 = static_cast<uint32_t>(plane_mat_.size());
			desc.materials		 = materials_.data();
			desc.n_materials	 = static_cast<uint32_t>(materials_.size());
			desc.boxes			 = boxes_.data();
			desc.box_material	 = box_mat_.data();
			desc.n_boxes		 = static_cast<uint32_t>(box_mat_.size());
			// validated and built (BVH) once, copied to every device in use
			uploaded_ = rtcu_upload_scene_multi(ctxs_.data(), static_cast<uint32_t>(ctxs_.size()), &desc) == RTCU_OK;
			scene_uploads_++;
			return uploaded_;
		}

		static rtcu_view make_view(const rt::scene& scene, const image_view& pixels) noexcept
		{
			const auto view = scene.camera.viewport(pixels.size());

			rtcu_view v{};
			for (int c = 0; c < 4; c++)
				for (int r = 0; r < 4; r++)
					v.inv_view_proj[c * 4 + r] = view.inverse_view_projection(static_cast<size_t>(r), static_cast<size_t>(c));
			v.width				= pixels.size().x;
			v.height			= pixels.size().y;
			v.samples_per_pixel = scene.samples_per_pixel;
			v.max_bounces		= scene.max_bounces;
			v.sample_begin		= 0;
			v.sample_end		= scene.samples_per_pixel;
			v.tile_x0			= 0;
			v.tile_y0			= 0;
			v.tile_x1			= v.width;
			v.tile_y1			= v.height;
			v.seed				= 0x5EEDull;
			v.flags				= static_cast<uint32_t>(RTCU_ACCEL_AUTO) | static_cast<uint32_t>(RTCU_PIPE_AUTO);
			return v;
		}
	};

	struct cuda_path_tracer final : cuda_renderer
	{
		uint32_t material_mode_ = RTCU_MODE_SM;
		bool progressive_		= false;
		rtcu_view last_view_{}; // the frame the sums on the device belong to
		uint32_t samples_done_ = 0;

		cuda_path_tracer() : cuda_renderer{ "cuda_path_tracer", true }
		{
			if (const char* mode = std::getenv("RT_CUDA_MATERIAL_MODE"); mode && std::strcmp(mode, "mg") == 0)
				material_mode_ = RTCU_MODE_MG;
			if (const char* prog = std::getenv("RT_CUDA_PROGRESSIVE"); prog && prog[0] == '1')
				progressive_ = true;
		}

		void render(const rt::scene& scene, image_view& pixels, muu::thread_pool& /*threads*/) noexcept override
		{
			if (!sync_scene(scene))
			{
				std::fprintf(stderr, "cuda_path_tracer: %s\n", rtcu_last_error());
				return;
			}
			const bool scene_changed = scene_uploads_ != uploads_seen_;
			uploads_seen_			 = scene_uploads_;
			rtcu_view v				 = make_view(scene, pixels);
			v.material_mode			 = material_mode_;

			if (progressive_)
			{
				// the same scene through the same view as the sums on the device: add the next samples instead of starting over
				const bool same = !scene_changed && samples_done_ > 0 && std::memcmp(&v, &last_view_, sizeof v) == 0;
				last_view_ = v;
				if (!same)
					samples_done_ = 0;
				v.sample_begin		= samples_done_;
				v.sample_end		= samples_done_ + scene.samples_per_pixel;
				v.samples_per_pixel = v.sample_end;
				if (same)
					v.flags |= RTCU_FLAG_ACCUMULATE;
				if (rtcu_render(ctx_, &v, pixels.data(), nullptr) != RTCU_OK)
				{
					std::fprintf(stderr, "cuda_path_tracer: %s\n", rtcu_last_error());
					samples_done_ = 0;
					return;
				}
				samples_done_ = v.sample_end;
				rtcu_get_stats(ctx_, &g_last_stats);
				return;
			}

			// sample-range split over the devices while every device keeps at least 16 samples per pixel (below that the
			// per-device kernels are too short to pay for the exchange)
			uint32_t devices = multi_ok_ ? static_cast<uint32_t>(ctxs_.size()) : 1u;
			while (devices > 1 && scene.samples_per_pixel < 16u * devices)
				devices--;
			if (devices > 1)
			{
				if (rtcu_render_multi(ctxs_.data(), devices, &v, pixels.data(), nullptr) == RTCU_OK)
				{
					rtcu_get_stats(ctx_, &g_last_stats);
					return;
				}
				std::fprintf(stderr, "cuda_path_tracer: %s; rendering on device 0 from now on\n", rtcu_last_error());
				multi_ok_ = false;
			}
			if (rtcu_render(ctx_, &v, pixels.data(), nullptr) != RTCU_OK)
				std::fprintf(stderr, "cuda_path_tracer: %s\n", rtcu_last_error());
			else
				rtcu_get_stats(ctx_, &g_last_stats);
		}
	};

	// the preview renderer (rasterizer.cpp:22-88): one ray per pixel, N.L shading; also draws scene.boxes
	struct cuda_rasterizer final : cuda_renderer
	{
		cuda_rasterizer() : cuda_renderer{ "cuda_rasterizer" } {}

		void render(const rt::scene& scene, image_view& pixels, muu::thread_pool& /*threads*/) noexcept override
		{
			if (!sync_scene(scene))
			{
				std::fprintf(stderr, "cuda_rasterizer: %s\n", rtcu_last_error());
				return;
			}
			const rtcu_view v = make_view(scene, pixels);
			if (rtcu_rasterize(ctx_, &v, pixels.data(), nullptr, nullptr) != RTCU_OK)
				std::fprintf(stderr, "cuda_rasterizer: %s\n", rtcu_last_error());
			else
				rtcu_get_stats(ctx_, &g_last_stats);
		}
	};

	REGISTER_RENDERER(cuda_path_tracer);
	REGISTER_RENDERER(cuda_rasterizer);
}

// statistics of the most recent frame (device time, copy time, devices' kernel launches): a diagnostics hook for an imgui
// overlay or a test; not part of the renderer interface
extern "C" void rt_cuda_last_stats(rtcu_stats* out)
{
	if (out)
		*out = g_last_stats;
}
