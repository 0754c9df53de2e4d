// c2d_batch.cu — stateless batched entry points over the narrowphase device functions
// (include/c2d_gpu.h, "stateless batches").  They exist so that the reference's known-answer tests of the
// free functions (tests/collision/test_collision.cpp, tests/geometry/test_SweepShapes.cpp,
// test_RaycastShapes.cpp, test_SweepRaycast.cpp) can be replayed against the SAME device code the
// per-frame pipeline runs.  Host buffers in, host buffers out; one thread per item.
#include "c2d_internal.hpp"
#include "c2d_ray.cuh"

#include "../../include/c2d_gpu.h"

#include <cmath>
#include <string>
#include <vector>

using namespace c2d_dev;

namespace
{

struct BatchShapes
{
    const uint8_t* type;
    const float* geom; // n x 32
    const uint8_t* count;
    const float* xf5;    // n x 5
    const float* xf5End; // n x 5 or nullptr
    const float* trig;   // n x 2: libm sinf/cosf of the start angle
};

__device__ __forceinline__ ShapeRef batch_shape(const BatchShapes& b, uint32_t i)
{
    ShapeRef s;
    s.type = b.type[i];
    s.count = b.count[i];
    s.g = b.geom + static_cast<size_t>(i) * 32;
    return s;
}
__device__ __forceinline__ XF batch_xf(const float* xf5, uint32_t i)
{
    XF t;
    t.tx = xf5[5 * i];
    t.ty = xf5[5 * i + 1];
    t.s = xf5[5 * i + 3];
    t.c = xf5[5 * i + 4];
    return t;
}
__device__ __forceinline__ Motion batch_motion(const BatchShapes& b, uint32_t i)
{
    if (!b.xf5End)
        return motion_fixed(batch_xf(b.xf5, i));
    return motion_between(batch_xf(b.xf5, i), b.xf5[5 * i + 2], b.trig[2 * i], b.trig[2 * i + 1],
                          batch_xf(b.xf5End, i), b.xf5End[5 * i + 2]);
}

__global__ void k_batch_collide(uint32_t n, BatchShapes a, BatchShapes b, float* manifolds17, uint8_t* flags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const ShapeRef A = batch_shape(a, i), B = batch_shape(b, i);
    const XF xa = batch_xf(a.xf5, i), xb = batch_xf(b.xf5, i);
    Manifold m;
    if (manifolds17)
    {
        narrow_test<true>(A, xa, B, xb, m);
        manifold_store(m, manifolds17 + static_cast<size_t>(i) * 17);
    }
    if (flags)
        flags[i] = narrow_test<false>(A, xa, B, xb, m) ? 1 : 0;
}

__global__ void k_batch_sweep(uint32_t n, BatchShapes a, BatchShapes b, float* fractions, float* manifolds17,
                              uint8_t* hits)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const ShapeRef A = batch_shape(a, i), B = batch_shape(b, i);
    const Motion ma = batch_motion(a, i), mb = batch_motion(b, i);
    Manifold m;
    manifold_clear(m);
    float fraction = -1.0f;
    const bool hit = narrow_sweep(A, ma, B, mb, fraction, m);
    if (!hit)
    {
        manifold_clear(m);
        fraction = -1.0f;
    }
    if (hits)
        hits[i] = hit ? 1 : 0;
    if (fractions)
        fractions[i] = fraction;
    if (manifolds17)
        manifold_store(m, manifolds17 + static_cast<size_t>(i) * 17);
}

__global__ void k_batch_raycast(uint32_t n, BatchShapes s, const float* ray4, const uint8_t* infinite, float* times2,
                                uint8_t* hits)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const ShapeRef S = batch_shape(s, i);
    RayQ r;
    r.start = mk(ray4[4 * i], ray4[4 * i + 1]);
    r.infinite = infinite && infinite[i];
    const V2 second = mk(ray4[4 * i + 2], ray4[4 * i + 3]);
    r.d = r.infinite ? second : (second - r.start);
    float t0 = 0.0f, t1 = 0.0f;
    bool hit;
    if (s.xf5End)
        hit = ray_sweep(S, batch_motion(s, i), r, second, t0, t1);
    else
        hit = ray_shape(S, batch_xf(s.xf5, i), r, second, t0, t1);
    hits[i] = hit ? 1 : 0;
    times2[2 * i] = hit ? t0 : 0.0f;
    times2[2 * i + 1] = hit ? t1 : 0.0f;
}

__global__ void k_batch_aabb(uint32_t n, BatchShapes s, float* out4)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const Box b = compute_aabb(s.type[i], s.count[i], s.geom + static_cast<size_t>(i) * 32, batch_xf(s.xf5, i));
    out4[4 * i] = b.minx;
    out4[4 * i + 1] = b.miny;
    out4[4 * i + 2] = b.maxx;
    out4[4 * i + 3] = b.maxy;
}

// ---- host plumbing -----------------------------------------------------------------------------------
struct Arena
{
    std::vector<void*> ptrs;
    cudaError_t err = cudaSuccess;
    ~Arena()
    {
        for (void* p : ptrs)
            cudaFree(p);
    }
    template <typename T>
    T* up(const T* host, size_t count)
    {
        if (!host || err != cudaSuccess)
            return nullptr;
        T* d = nullptr;
        err = cudaMalloc(reinterpret_cast<void**>(&d), std::max<size_t>(count, 1) * sizeof(T));
        if (err != cudaSuccess)
            return nullptr;
        ptrs.push_back(d);
        err = cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice);
        return d;
    }
    template <typename T>
    T* alloc(size_t count)
    {
        if (err != cudaSuccess)
            return nullptr;
        T* d = nullptr;
        err = cudaMalloc(reinterpret_cast<void**>(&d), std::max<size_t>(count, 1) * sizeof(T));
        if (err != cudaSuccess)
            return nullptr;
        ptrs.push_back(d);
        return d;
    }
    template <typename T>
    void down(T* host, const T* dev, size_t count)
    {
        if (host && dev && err == cudaSuccess)
            err = cudaMemcpy(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost);
    }
};

thread_local std::string g_batchError;

int batch_fail(const std::string& msg);

BatchShapes upload(Arena& ar, uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt, const float* xf5,
                   const float* xf5End)
{
    BatchShapes b{};
    b.type = ar.up(type, n);
    b.geom = ar.up(geom, static_cast<size_t>(n) * 32);
    b.count = ar.up(cnt, n);
    b.xf5 = ar.up(xf5, static_cast<size_t>(n) * 5);
    b.xf5End = xf5End ? ar.up(xf5End, static_cast<size_t>(n) * 5) : nullptr;
    b.trig = nullptr;
    if (xf5End)
    {
        // Transform(Vec2, float) evaluates sinf/cosf with the host's libm (Transform.hpp:23-28)
        std::vector<float> trig(static_cast<size_t>(n) * 2);
        for (uint32_t i = 0; i < n; ++i)
        {
            trig[2 * i] = std::sin(xf5[5 * i + 2]);
            trig[2 * i + 1] = std::cos(xf5[5 * i + 2]);
        }
        b.trig = ar.up(trig.data(), trig.size());
    }
    return b;
}

int finish(Arena& ar, const char* what)
{
    if (ar.err == cudaSuccess)
        ar.err = cudaGetLastError();
    if (ar.err == cudaSuccess)
        ar.err = cudaDeviceSynchronize();
    if (ar.err != cudaSuccess)
        return batch_fail(std::string(what) + ": " + cudaGetErrorString(ar.err));
    return 0;
}

} // namespace

// defined in c2d_gpu.cu's namespace via c2d_gpu_last_error(NULL): share the thread-local message
extern "C" void c2d_gpu_internal_set_error(const char* msg);
namespace
{
int batch_fail(const std::string& msg)
{
    c2d_gpu_internal_set_error(msg.c_str());
    return 1;
}
} // namespace

extern "C" {

int c2d_gpu_collide_batch(uint32_t n, const uint8_t* typeA, const float* geomA, const uint8_t* cntA, const float* xfA5,
                          const uint8_t* typeB, const float* geomB, const uint8_t* cntB, const float* xfB5,
                          void* out_manifold68, uint8_t* out_are_colliding)
{
    if (n == 0)
        return 0;
    Arena ar;
    const BatchShapes a = upload(ar, n, typeA, geomA, cntA, xfA5, nullptr);
    const BatchShapes b = upload(ar, n, typeB, geomB, cntB, xfB5, nullptr);
    float* dm = out_manifold68 ? ar.alloc<float>(static_cast<size_t>(n) * 17) : nullptr;
    uint8_t* df = out_are_colliding ? ar.alloc<uint8_t>(n) : nullptr;
    if (ar.err == cudaSuccess)
        k_batch_collide<<<(n + 127) / 128, 128>>>(n, a, b, dm, df);
    if (int rc = finish(ar, "c2d_gpu_collide_batch"))
        return rc;
    ar.down(static_cast<float*>(out_manifold68), dm, static_cast<size_t>(n) * 17);
    ar.down(out_are_colliding, df, n);
    return ar.err == cudaSuccess ? 0 : batch_fail("c2d_gpu_collide_batch: download failed");
}

int c2d_gpu_sweep_batch(uint32_t n, int mode, const uint8_t* typeA, const float* geomA, const uint8_t* cntA,
                        const float* xfA5, const float* xfA5_end, const uint8_t* typeB, const float* geomB,
                        const uint8_t* cntB, const float* xfB5, const float* xfB5_end, float* out_fraction,
                        void* out_manifold68, uint8_t* out_hit)
{
    if (n == 0)
        return 0;
    if (mode != 1 && mode != 2)
        return batch_fail("c2d_gpu_sweep_batch: mode must be 1 or 2");
    Arena ar;
    const BatchShapes a = upload(ar, n, typeA, geomA, cntA, xfA5, xfA5_end);
    const BatchShapes b = upload(ar, n, typeB, geomB, cntB, xfB5, mode == 2 ? xfB5_end : nullptr);
    float* dfr = ar.alloc<float>(n);
    float* dm = ar.alloc<float>(static_cast<size_t>(n) * 17);
    uint8_t* dh = ar.alloc<uint8_t>(n);
    if (ar.err == cudaSuccess)
        k_batch_sweep<<<(n + 127) / 128, 128>>>(n, a, b, dfr, dm, dh);
    if (int rc = finish(ar, "c2d_gpu_sweep_batch"))
        return rc;
    ar.down(out_fraction, dfr, n);
    ar.down(static_cast<float*>(out_manifold68), dm, static_cast<size_t>(n) * 17);
    ar.down(out_hit, dh, n);
    return ar.err == cudaSuccess ? 0 : batch_fail("c2d_gpu_sweep_batch: download failed");
}

int c2d_gpu_raycast_shape_batch(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt,
                                const float* xf5, const float* xf5_end, const float* ray4, const uint8_t* infinite,
                                float* out_times2, uint8_t* out_hit)
{
    if (n == 0)
        return 0;
    Arena ar;
    const BatchShapes s = upload(ar, n, type, geom, cnt, xf5, xf5_end);
    const float* dr = ar.up(ray4, static_cast<size_t>(n) * 4);
    const uint8_t* di = infinite ? ar.up(infinite, n) : nullptr;
    float* dt = ar.alloc<float>(static_cast<size_t>(n) * 2);
    uint8_t* dh = ar.alloc<uint8_t>(n);
    if (ar.err == cudaSuccess)
        k_batch_raycast<<<(n + 127) / 128, 128>>>(n, s, dr, di, dt, dh);
    if (int rc = finish(ar, "c2d_gpu_raycast_shape_batch"))
        return rc;
    ar.down(out_times2, dt, static_cast<size_t>(n) * 2);
    ar.down(out_hit, dh, n);
    return ar.err == cudaSuccess ? 0 : batch_fail("c2d_gpu_raycast_shape_batch: download failed");
}

int c2d_gpu_compute_aabb_batch(uint32_t n, const uint8_t* type, const float* geom, const uint8_t* cnt,
                               const float* xf5, float* out_aabb4)
{
    if (n == 0)
        return 0;
    Arena ar;
    const BatchShapes s = upload(ar, n, type, geom, cnt, xf5, nullptr);
    float* d = ar.alloc<float>(static_cast<size_t>(n) * 4);
    if (ar.err == cudaSuccess)
        k_batch_aabb<<<(n + 127) / 128, 128>>>(n, s, d);
    if (int rc = finish(ar, "c2d_gpu_compute_aabb_batch"))
        return rc;
    ar.down(out_aabb4, d, static_cast<size_t>(n) * 4);
    return ar.err == cudaSuccess ? 0 : batch_fail("c2d_gpu_compute_aabb_batch: download failed");
}

} // extern "C"
