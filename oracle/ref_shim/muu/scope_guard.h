#pragma once
// STAND-IN (empty) for muu/scope_guard.h; nothing on the renderer path uses it.  TEST INFRASTRUCTURE.
#include "vector.h"
