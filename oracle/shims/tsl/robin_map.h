// Shim for tsl::robin_map (v1.4.0, fetched by the reference's CMake, absent
// here): alias to std::unordered_map.  Registry.hpp only uses
// find/at/emplace/erase/size/clear/iteration (Registry.hpp:157,167,266-270,
// 363-379); iteration order affects serialisation only, never query results.
#pragma once
#include <functional>
#include <unordered_map>
namespace tsl
{
template <class K, class V, class H = std::hash<K>, class E = std::equal_to<K>>
using robin_map = std::unordered_map<K, V, H, E>;
}
