// snapshot_check.cpp — TEST INFRASTRUCTURE: the host-only snapshot codec of include/colli2de/internal/utils/Snapshot.hpp
// behind a C interface, so the CPU test suite can feed it snapshots written by the reference (oracle/_ref) and hand
// the re-encoded bytes back to the reference's deserialize.  No CUDA, no GPU library.
#include <colli2de/internal/utils/Snapshot.hpp>

#include <cstdlib>
#include <sstream>
#include <string>

namespace
{
std::string g_error;
}

extern "C" {

const char* snap_last_error(void)
{
    return g_error.c_str();
}

// Decodes a Registry<uint32_t, None | Grid> snapshot and encodes it again (the trees are rebuilt).  Returns 0 on success;
// *out is malloc'ed, release it with snap_free.
int snap_reencode(const void* bytes, uint64_t size, int grid, void** out, uint64_t* out_size)
{
    try
    {
        std::stringstream in(std::string(static_cast<const char*>(bytes), size), std::ios::in | std::ios::binary);
        const auto state = c2d::snapshot::read<uint32_t>(in, grid != 0);
        std::stringstream ss(std::ios::in | std::ios::out | std::ios::binary);
        c2d::snapshot::write(ss, state);
        const std::string blob = ss.str();
        *out = std::malloc(blob.size() ? blob.size() : 1);
        std::memcpy(*out, blob.data(), blob.size());
        *out_size = blob.size();
        g_error.clear();
        return 0;
    }
    catch (const std::exception& e)
    {
        g_error = e.what();
        return 1;
    }
}

// Summary of a decoded snapshot: counts[0] entities, [1] shape slots, [2] live shapes, [3] free ids, [4] bullets with a
// previous transform, [5] inactive live shapes.
int snap_summary(const void* bytes, uint64_t size, int grid, uint64_t* counts6)
{
    try
    {
        std::stringstream in(std::string(static_cast<const char*>(bytes), size), std::ios::in | std::ios::binary);
        const auto state = c2d::snapshot::read<uint32_t>(in, grid != 0);
        counts6[0] = state.entities.size();
        counts6[1] = state.shapes.size();
        counts6[2] = counts6[4] = counts6[5] = 0;
        for (const auto& s : state.shapes)
        {
            counts6[2] += s.alive ? 1 : 0;
            counts6[5] += (s.alive && !s.active) ? 1 : 0;
        }
        counts6[3] = state.freeShapeIds.size();
        for (const auto& e : state.entities)
            counts6[4] += e.hasPrev ? 1 : 0;
        g_error.clear();
        return 0;
    }
    catch (const std::exception& e)
    {
        g_error = e.what();
        return 1;
    }
}

void snap_free(void* p)
{
    std::free(p);
}
}
