// STAND-IN for muu/plane.h (see vector.h).  TEST INFRASTRUCTURE.
#pragma once
#include "vector.h"
namespace muu
{
	template <typename T>
	struct plane
	{
		vector<T, 3> normal{};
		T d{};
		constexpr plane() noexcept = default;
		constexpr plane(vector<T, 3> n, T d_) noexcept : normal{ n }, d{ d_ } {}
		// plane through `position` with unit `direction`: dot(n, p) + d == 0
		constexpr plane(vector<T, 3> position, vector<T, 3> direction) noexcept
			: normal{ direction }, d{ -vector<T, 3>::dot(direction, position) }
		{}
	};
}
