#pragma once
// STAND-IN for muu/bounding_sphere.h (see vector.h).  TEST INFRASTRUCTURE.
#include "vector.h"
namespace muu
{
	template <typename T>
	struct bounding_sphere
	{
		vector<T, 3> center{};
		T radius{};
		constexpr bounding_sphere() noexcept = default;
		constexpr bounding_sphere(vector<T, 3> c, T r) noexcept : center{ c }, radius{ r } {}
	};
}
