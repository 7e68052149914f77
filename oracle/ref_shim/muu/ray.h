// STAND-IN for muu/ray.h: ray::at and ray::hits(plane / bounding_sphere) as restated in oracle/rtref.c (SPEC S3-S5).
// muu's own formula is UNVERIFIED (Game-Physics-Cookbook form).  TEST INFRASTRUCTURE.
#pragma once
#include "bounding_box.h"
#include "bounding_sphere.h"
#include "plane.h"
#include <optional>
namespace muu
{
	template <typename T>
	struct ray
	{
		vector<T, 3> origin{};
		vector<T, 3> direction{};
		constexpr ray() noexcept = default;
		constexpr ray(vector<T, 3> o, vector<T, 3> d) noexcept : origin{ o }, direction{ d } {}

		// SPEC S3
		[[nodiscard]] vector<T, 3> at(T t) const noexcept
		{
			shim::on_ray_at();
			return { shim_fma(direction.x, t, origin.x), shim_fma(direction.y, t, origin.y), shim_fma(direction.z, t, origin.z) };
		}
		// rasterizer.cpp:48 passes the optional itself: `hit_pos = r.at(hit)` (only reached when it holds a value)
		[[nodiscard]] vector<T, 3> at(const std::optional<T>& t) const noexcept { return at(*t); }
		// SPEC S13 (rasterizer only): slab test against a centre/extents box, Game-Physics-Cookbook form like S4/S5:
		// t1/t2 per axis by IEEE division, tmin = max of the three near values, tmax = min of the far ones;
		// tmax < 0 or tmin > tmax -> miss; origin inside (tmin < 0) -> tmax, else tmin.  UNVERIFIED against muu.
		[[nodiscard]] constexpr std::optional<T> hits(const bounding_box<T>& b) const noexcept
		{
			const T lo[3] = { b.center.x - b.extents.x, b.center.y - b.extents.y, b.center.z - b.extents.z };
			const T hi[3] = { b.center.x + b.extents.x, b.center.y + b.extents.y, b.center.z + b.extents.z };
			const T o[3] = { origin.x, origin.y, origin.z }, d[3] = { direction.x, direction.y, direction.z };
			T tmin = -__builtin_inff(), tmax = __builtin_inff();
			for (int k = 0; k < 3; k++)
			{
				const T t1 = (lo[k] - o[k]) / d[k], t2 = (hi[k] - o[k]) / d[k];
				tmin = __builtin_fmaxf(tmin, __builtin_fminf(t1, t2));
				tmax = __builtin_fminf(tmax, __builtin_fmaxf(t1, t2));
			}
			if (tmax < T{} || tmin > tmax) return {};
			return tmin < T{} ? tmax : tmin;
		}
		// SPEC S5
		[[nodiscard]] constexpr std::optional<T> hits(const plane<T>& p) const noexcept
		{
			using v3 = vector<T, 3>;
			const T nd = v3::dot(direction, p.normal);
			if (nd >= T{}) return {};
			const T t = (-p.d - v3::dot(origin, p.normal)) / nd;
			if (t < T{}) return {};
			return t;
		}
		// SPEC S4
		[[nodiscard]] constexpr std::optional<T> hits(const bounding_sphere<T>& s) const noexcept
		{
			using v3 = vector<T, 3>;
			const v3 e = s.center - origin;
			const T e2 = v3::dot(e, e);
			const T r2 = s.radius * s.radius;
			const T a = v3::dot(e, direction);
			const T disc = r2 - shim_fma(-a, a, e2);
			if (disc < T{}) return {};
			const T f = __builtin_sqrtf(disc);
			return (e2 < r2) ? a + f : a - f;
		}
	};
}
