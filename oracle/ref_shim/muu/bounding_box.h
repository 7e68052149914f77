#pragma once
// STAND-IN for muu/bounding_box.h (see vector.h).  TEST INFRASTRUCTURE.
#include "vector.h"
namespace muu
{
	template <typename T>
	struct bounding_box
	{
		vector<T, 3> center{};
		vector<T, 3> extents{};
		constexpr bounding_box() noexcept = default;
		constexpr bounding_box(vector<T, 3> c, vector<T, 3> e) noexcept : center{ c }, extents{ e } {}
	};
}
