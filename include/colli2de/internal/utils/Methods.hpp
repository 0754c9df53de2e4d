// colli2de-b200 — float_equals, the epsilon comparison every operator== of the
// value types is built on (reference: internal/utils/Methods.hpp:6-14).
#pragma once

#include <cmath>

inline constexpr bool float_equals(float a, float b, float epsilon = 1e-6f)
{
    const float d = a - b;
    return (d < 0.0f ? -d : d) < epsilon;
}

inline constexpr bool float_equals(double a, double b, double epsilon = 1e-6)
{
    const double d = a - b;
    return (d < 0.0 ? -d : d) < epsilon;
}
