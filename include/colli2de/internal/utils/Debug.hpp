// colli2de-b200 — precondition macro.  Same convention as the reference
// (internal/utils/Debug.hpp:11-37): with DBG or ENABLE_ASSERT defined a failed
// precondition throws std::runtime_error("Assertion failed..."), otherwise the
// check compiles away.
#pragma once

#include <stdexcept>
#include <string>

#ifdef DBG
#define DEBUG_ONLY(code) code
#else
#define DEBUG_ONLY(code)
#endif

#if defined(DBG) || defined(ENABLE_ASSERT)
namespace c2d::detail
{
[[noreturn]] inline void assertionFailed()
{
    throw std::runtime_error("Assertion failed.");
}
[[noreturn]] inline void assertionFailed(const std::string& message)
{
    throw std::runtime_error("Assertion failed: " + message);
}
} // namespace c2d::detail
#define C2D_DEBUG_ASSERT(condition, ...)                                                                               \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(condition))                                                                                              \
            ::c2d::detail::assertionFailed(__VA_ARGS__);                                                               \
    } while (0)
#else
#define C2D_DEBUG_ASSERT(...)
#endif
