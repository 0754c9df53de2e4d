// colli2de-b200 — public constants.  API-compatible with the reference's
// include/colli2de/Constants.hpp:9-10 (same names, same values).
#pragma once

#include <cstddef>

namespace c2d
{

inline constexpr std::size_t MAX_POLYGON_VERTICES = 8;
inline constexpr float PI = 3.14159265358979323846f;

} // namespace c2d
