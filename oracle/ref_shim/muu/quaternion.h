#pragma once
// STAND-IN (empty) for muu/quaternion.h; nothing on the renderer path uses it.  TEST INFRASTRUCTURE.
#include "vector.h"
namespace muu { template <typename T> struct quaternion { T s{}; vector<T, 3> v{}; }; }
