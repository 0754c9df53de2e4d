// colli2de-b200 — Ray, InfiniteRay and RaycastHit of the public API.
//
// Source compatible with the reference's include/colli2de/Ray.hpp.  Kept:
//   * hit points: start + (end - start) * t  (finite) / start + direction * t (infinite)  (Ray.hpp:34-65)
//   * RaycastHit ordering is on (entryTime, exitTime) ONLY (Ray.hpp:77-80): a std::set of hits
//     collapses two hits with identical times into one; operator== compares ids.
#pragma once

#include <colli2de/Vec2.hpp>

#include <type_traits>
#include <utility>

namespace c2d
{

struct Ray
{
    Vec2 start;
    Vec2 end;
};

struct InfiniteRay
{
    Vec2 start;
    Vec2 direction;
};

template <typename RayType>
concept IsRay = std::is_same_v<RayType, Ray> || std::is_same_v<RayType, InfiniteRay>;

template <typename IdType>
struct RaycastHit
{
    IdType id;
    Vec2 entry;
    Vec2 exit;
    float entryTime;
    float exitTime;

    static RaycastHit fromRay(IdType id, Ray ray, float entryTime, float exitTime)
    {
        const Vec2 span = ray.end - ray.start;
        return RaycastHit{id, ray.start + span * entryTime, ray.start + span * exitTime, entryTime, exitTime};
    }
    static RaycastHit fromRay(IdType id, InfiniteRay ray, float entryTime, float exitTime)
    {
        return RaycastHit{
            id, ray.start + ray.direction * entryTime, ray.start + ray.direction * exitTime, entryTime, exitTime};
    }
    template <IsRay RayType>
    static RaycastHit fromRay(IdType id, RayType ray, std::pair<float, float> times)
    {
        return fromRay(id, ray, times.first, times.second);
    }

    bool operator==(const RaycastHit& o) const { return id == o.id; }
    bool operator!=(const RaycastHit& o) const { return !(*this == o); }
    bool operator<(const RaycastHit& o) const
    {
        return entryTime < o.entryTime || (entryTime == o.entryTime && exitTime < o.exitTime);
    }
    bool operator>(const RaycastHit& o) const
    {
        return entryTime > o.entryTime || (entryTime == o.entryTime && exitTime > o.exitTime);
    }
    bool operator<=(const RaycastHit& o) const { return !(*this > o); }
    bool operator>=(const RaycastHit& o) const { return !(*this < o); }
};

} // namespace c2d
