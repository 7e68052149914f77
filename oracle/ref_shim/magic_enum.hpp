// STAND-IN for magic_enum.hpp: the renderers only call enum_count<rt::material_type>() (mg_ray_tracer.cpp:144), whose
// answer is the 8 enumerators of common.hpp:105-115.  TEST INFRASTRUCTURE.
#pragma once
#include <cstddef>
namespace magic_enum
{
	template <typename E>
	[[nodiscard]] constexpr std::size_t enum_count() noexcept
	{
		return 8;
	}
}
