// rt_stub.hpp -- COMPILE-CHECK ONLY.  Minimal stand-ins for the parts of marzer/rt (and muu) that
// plugin/cuda_path_tracer.cpp touches, so the plugin TU can be syntax- and type-checked in an environment
// where the reference cannot be built (muu is not vendored).  Nothing here is shipped or linked.
// Shapes follow: src/scene.hpp:8-25, src/soa.hpp:157-199 (column accessors), src/camera.hpp:6-49,:122-137,
// src/image.hpp:112-163, src/renderer.hpp:9-41.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string_view>

#define MUU_DISABLE_WARNINGS static_assert(true)
#define MUU_ENABLE_WARNINGS static_assert(true)

namespace muu
{
	class thread_pool
	{};
}

namespace rt
{
	using std::size_t;
	using std::uint32_t;
	struct vec2u { unsigned x, y; };
	struct vec3 { float x, y, z; };
	struct sphere { vec3 center; float radius; };
	struct plane { vec3 normal; float d; };
	struct box { vec3 center; vec3 extents; };
	struct colour { float r, g, b, a; };
	enum class material_type : unsigned { lambert, metal, dielectric, air, vacuum, water, ice, diamond };

	struct mat4
	{
		float m[4][4]; // m[column][row]
		float operator()(size_t r, size_t c) const noexcept { return m[c][r]; }
	};

	struct viewport
	{
		mat4 inverse_view_projection;
	};

	class camera
	{
	  public:
		rt::viewport viewport(vec2u) const noexcept { return {}; }
	};

	class materials
	{
	  public:
		size_t size() const noexcept { return 0; }
		const material_type* type() const noexcept { return nullptr; }
		const colour* albedo() const noexcept { return nullptr; }
		const float* roughness() const noexcept { return nullptr; }
		const float* reflectivity() const noexcept { return nullptr; }
	};
	class spheres
	{
	  public:
		size_t size() const noexcept { return 0; }
		const sphere* value() const noexcept { return nullptr; }
		const unsigned* material() const noexcept { return nullptr; }
	};
	class planes
	{
	  public:
		size_t size() const noexcept { return 0; }
		const plane* value() const noexcept { return nullptr; }
		const unsigned* material() const noexcept { return nullptr; }
	};

	class boxes
	{
	  public:
		size_t size() const noexcept { return 0; }
		const box* value() const noexcept { return nullptr; }
		const unsigned* material() const noexcept { return nullptr; }
	};

	struct scene
	{
		unsigned samples_per_pixel = 30;
		unsigned max_bounces	   = 10;
		rt::camera camera;
		rt::materials materials;
		rt::planes planes;
		rt::spheres spheres;
		rt::boxes boxes;
	};

	class image_view
	{
		uint32_t* data_ = nullptr;
		vec2u size_{};

	  public:
		const vec2u& size() const noexcept { return size_; }
		uint32_t* data() const noexcept { return data_; }
	};

	struct renderer_interface
	{
		virtual void render(const scene&, image_view&, muu::thread_pool&) noexcept = 0;
		virtual ~renderer_interface() noexcept = default;
	};

	namespace renderers
	{
		struct description
		{
			using create_func = renderer_interface*();
			std::string_view key;
			std::string_view name;
			create_func* create;
		};
		inline void install(const description&) {}
	}
}

#define RT_STUB_STR2(x) #x
#define RT_STUB_STR(x) RT_STUB_STR2(x)
#define REGISTER_RENDERER(T)                                                                                           \
	[[maybe_unused]] static const int register_val_impl_##T =                                                         \
		(::rt::renderers::install({ .key = __FILE__ ":" RT_STUB_STR(__LINE__) ":" #T, .name = #T,                     \
									.create = []() -> renderer_interface* { return new T; } }),                       \
		 0)
