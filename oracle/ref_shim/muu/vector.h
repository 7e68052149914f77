// STAND-IN for muu/vector.h -- the operations marzer/rt's renderers call, with the arithmetic order of the oracle's
// numbered SPEC (oracle/rtref.c S1-S3, S8).  The real muu (git wrap @06dbcecb) is absent, so these bodies are this
// repo's restatement of muu's published semantics, NOT muu's code.  TEST INFRASTRUCTURE.
#pragma once
#include "preprocessor.h"
#include <cmath>
#include <cstddef>
#include <type_traits>

namespace muu
{
	class thread_pool; // muu/fwd.h declares it; renderer.hpp:11 names it before including thread_pool.h
	inline namespace literals
	{}

	template <typename From, typename To>
	inline constexpr bool allow_implicit_bit_cast = false;

	template <typename To, typename From>
	[[nodiscard]] constexpr To bit_cast(const From& from) noexcept
	{
		static_assert(sizeof(To) == sizeof(From));
		return __builtin_bit_cast(To, from);
	}

	template <typename T>
	using remove_cvref = std::remove_cv_t<std::remove_reference_t<T>>;
	template <typename T, bool = std::is_enum_v<T>>
	struct remove_enum_ { using type = T; };
	template <typename T>
	struct remove_enum_<T, true> { using type = std::underlying_type_t<T>; };
	template <typename T>
	using remove_enum = typename remove_enum_<T>::type;
	template <typename T>
	inline constexpr bool is_floating_point = std::is_floating_point_v<T>;

	template <typename T>
	[[nodiscard]] constexpr auto unwrap(T val) noexcept
	{
		if constexpr (std::is_enum_v<T>)
			return static_cast<std::underlying_type_t<T>>(val);
		else
			return val;
	}

	template <typename T>
	[[nodiscard]] constexpr const T& clamp(const T& v, const T& lo, const T& hi) noexcept
	{
		return v < lo ? lo : (hi < v ? hi : v);
	}

	template <typename T>
	struct constants
	{
		static constexpr T pi			  = static_cast<T>(3.14159265358979323846264338327950288L);
		static constexpr T pi_over_four	  = static_cast<T>(3.14159265358979323846264338327950288L / 4.0L);
		static constexpr T default_epsilon = static_cast<T>(1e-5L); // UNVERIFIED
	};

	// the fused multiply-add of the SPEC (strict build: -ffp-contract=off, so only these fuse)
	constexpr float shim_fma(float a, float b, float c) noexcept
	{
		return __builtin_fmaf(a, b, c);
	}

	template <typename T, std::size_t N>
	struct vector;

	// hooks implemented in oracle/ref_shim/shim.cpp: they give the reference's context-free RNG calls the
	// (pixel, sample, block) coordinates of the counter-based stream (SPEC S9)
	namespace shim
	{
		void on_sample_end() noexcept;	 // `colour += trace(...)` in the per-pixel worker
		void on_primary_done() noexcept; // screen_to_world(pos, 1.0f)
		void on_ray_at() noexcept;		 // ray::at() closes a scatter event
	}

	template <typename T, std::size_t N>
	struct vector_constants;

	template <typename T>
	struct vector<T, 2>
	{
		T x{}, y{};
		constexpr vector() noexcept = default;
		constexpr vector(T x_, T y_) noexcept : x{ x_ }, y{ y_ } {}
		explicit constexpr vector(T s) noexcept : x{ s }, y{ s } {}
		template <typename U>
		explicit constexpr vector(const vector<U, 2>& o) noexcept : x{ static_cast<T>(o.x) }, y{ static_cast<T>(o.y) }
		{}
		friend constexpr vector operator+(vector a, vector b) noexcept { return { a.x + b.x, a.y + b.y }; }
		friend constexpr bool operator==(vector a, vector b) noexcept { return a.x == b.x && a.y == b.y; }
	};

	template <typename T>
	struct vector<T, 3>
	{
		T x{}, y{}, z{};

		using constants = vector_constants<T, 3>;

		constexpr vector() noexcept = default;
		constexpr vector(T x_, T y_, T z_) noexcept : x{ x_ }, y{ y_ }, z{ z_ } {}
		explicit constexpr vector(T s) noexcept : x{ s }, y{ s }, z{ s } {}
		template <typename U>
		requires(allow_implicit_bit_cast<U, vector>)
		constexpr vector(const U& o) noexcept : vector{ muu::bit_cast<vector>(o) }
		{}

		constexpr T& operator[](std::size_t i) noexcept { return i == 0 ? x : (i == 1 ? y : z); }
		constexpr const T& operator[](std::size_t i) const noexcept { return i == 0 ? x : (i == 1 ? y : z); }

		friend constexpr vector operator+(vector a, vector b) noexcept { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
		friend constexpr vector operator-(vector a, vector b) noexcept { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
		friend constexpr vector operator-(vector a) noexcept { return { -a.x, -a.y, -a.z }; }
		friend constexpr vector operator*(vector a, vector b) noexcept { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
		friend constexpr vector operator*(vector a, T s) noexcept { return { a.x * s, a.y * s, a.z * s }; }
		friend constexpr vector operator*(T s, vector a) noexcept { return { a.x * s, a.y * s, a.z * s }; }
		friend constexpr vector operator/(vector a, T s) noexcept { return { a.x / s, a.y / s, a.z / s }; }
		friend constexpr bool operator==(vector a, vector b) noexcept { return a.x == b.x && a.y == b.y && a.z == b.z; }
		// only the per-pixel worker uses += (mg_ray_tracer.cpp:193): it marks the end of a sample for the RNG shim
		vector& operator+=(vector b) noexcept
		{
			x += b.x; y += b.y; z += b.z;
			shim::on_sample_end();
			return *this;
		}
		constexpr vector& operator/=(T s) noexcept
		{
			x /= s; y /= s; z /= s; // SPEC S11: sum / float(spp)
			return *this;
		}

		// SPEC S1
		[[nodiscard]] static constexpr T dot(vector a, vector b) noexcept { return shim_fma(a.z, b.z, shim_fma(a.y, b.y, a.x * b.x)); }
		[[nodiscard]] constexpr T dot(vector b) const noexcept { return dot(*this, b); }
		[[nodiscard]] constexpr T length() const noexcept { return __builtin_sqrtf(dot(*this, *this)); }
		[[nodiscard]] static constexpr T distance(vector a, vector b) noexcept { return (b - a).length(); }
		[[nodiscard]] static constexpr vector min(vector a, vector b) noexcept
		{
			return { a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z };
		}
		// SPEC S2
		[[nodiscard]] static constexpr vector normalize(vector v) noexcept
		{
			const T inv = T{ 1 } / __builtin_sqrtf(dot(v, v));
			return { v.x * inv, v.y * inv, v.z * inv };
		}
		constexpr vector& normalize() noexcept { return *this = normalize(*this); }
		[[nodiscard]] static constexpr vector direction(vector from, vector to) noexcept { return normalize(to - from); }
		[[nodiscard]] constexpr bool approx_zero(T eps = muu::constants<T>::default_epsilon) const noexcept
		{
			return __builtin_fabsf(x) < eps && __builtin_fabsf(y) < eps && __builtin_fabsf(z) < eps;
		}
		// SPEC S8: start * (1 - alpha) + finish * alpha
		[[nodiscard]] static constexpr vector lerp(vector a, vector b, T t) noexcept
		{
			const T w = T{ 1 } - t;
			return { shim_fma(b.x, t, a.x * w), shim_fma(b.y, t, a.y * w), shim_fma(b.z, t, a.z * w) };
		}
		[[nodiscard]] static constexpr vector cross(vector a, vector b) noexcept
		{
			return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
		}
	};

	template <typename T>
	struct vector<T, 4>
	{
		T x{}, y{}, z{}, w{};
		using constants = vector_constants<T, 4>;
		constexpr vector() noexcept = default;
		constexpr vector(T x_, T y_, T z_, T w_) noexcept : x{ x_ }, y{ y_ }, z{ z_ }, w{ w_ } {}
		explicit constexpr vector(T s) noexcept : x{ s }, y{ s }, z{ s }, w{ s } {}
		constexpr vector(vector<T, 3> v, T w_) noexcept : x{ v.x }, y{ v.y }, z{ v.z }, w{ w_ } {}
		template <typename U>
		explicit constexpr vector(const vector<U, 4>& o) noexcept
			: x{ static_cast<T>(o.x) }, y{ static_cast<T>(o.y) }, z{ static_cast<T>(o.z) }, w{ static_cast<T>(o.w) }
		{}
		template <typename U>
		requires(allow_implicit_bit_cast<U, vector>)
		constexpr vector(const U& o) noexcept : vector{ muu::bit_cast<vector>(o) }
		{}
		friend constexpr vector operator+(vector a, vector b) noexcept { return { a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w }; }
		friend constexpr vector operator*(vector a, vector b) noexcept { return { a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w }; }
		friend constexpr vector operator*(vector a, T s) noexcept { return { a.x * s, a.y * s, a.z * s, a.w * s }; }
		constexpr vector& operator/=(T s) noexcept
		{
			x /= s; y /= s; z /= s; w /= s;
			return *this;
		}
		[[nodiscard]] static constexpr vector clamp(vector v, vector lo, vector hi) noexcept
		{
			// NaN clamps to lo like fminf(fmaxf(c, 0), 1) (SPEC S11)
			auto c = [](T a, T l, T h) { return __builtin_fminf(__builtin_fmaxf(a, l), h); };
			return { c(v.x, lo.x, hi.x), c(v.y, lo.y, hi.y), c(v.z, lo.z, hi.z), c(v.w, lo.w, hi.w) };
		}
	};

	template <typename T>
	struct vector_constants<T, 3>
	{
		static constexpr vector<T, 3> zero{ T{}, T{}, T{} };
		static constexpr vector<T, 3> one{ T{ 1 }, T{ 1 }, T{ 1 } };
		static constexpr vector<T, 3> x_axis{ T{ 1 }, T{}, T{} };
		static constexpr vector<T, 3> y_axis{ T{}, T{ 1 }, T{} };
		static constexpr vector<T, 3> z_axis{ T{}, T{}, T{ 1 } };
		static constexpr vector<T, 3> right{ T{ 1 }, T{}, T{} };
		static constexpr vector<T, 3> up{ T{}, T{ 1 }, T{} };
		static constexpr vector<T, 3> forward{ T{}, T{}, T{ -1 } }; // right-handed, -Z forward (UNVERIFIED)
		static constexpr vector<T, 3> backward{ T{}, T{}, T{ 1 } };
		static constexpr vector<T, 3> left{ T{ -1 }, T{}, T{} };
		static constexpr vector<T, 3> down{ T{}, T{ -1 }, T{} };
	};
	template <typename T>
	struct vector_constants<T, 4>
	{
		static constexpr vector<T, 4> zero{ T{}, T{}, T{}, T{} };
		static constexpr vector<T, 4> one{ T{ 1 }, T{ 1 }, T{ 1 }, T{ 1 } };
	};

	template <typename T, std::size_t N>
	[[nodiscard]] constexpr bool infinity_or_nan(const vector<T, N>&) noexcept
	{
		return false;
	}
}
