#pragma once
// STAND-IN (empty) for muu/assume_aligned.h; nothing on the renderer path uses it.  TEST INFRASTRUCTURE.
#include "vector.h"
namespace muu { template <std::size_t N, typename T> [[nodiscard]] constexpr T* assume_aligned(T* p) noexcept { return p; } }
