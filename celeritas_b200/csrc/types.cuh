//---------------------------------------------------------------------------//
// Basic types and small-vector math shared by every kernel.
//
// Arithmetic notes (they matter for bit-level agreement with the reference's
// host build, which is compiled without FMA contraction):
//  * this code is compiled with --fmad=false, so a*b+c is two roundings;
//  * where the reference calls fma() explicitly (dot_product, axpy, the
//    interpolator: /root/reference/src/corecel/math/ArrayUtils.hh:78-100,
//    /root/reference/src/corecel/grid/Interpolator.hh) we call fma() too.
//---------------------------------------------------------------------------//
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace b200
{
using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;
using real = double;

constexpr u32 INVALID = 0xffffffffu;

#define B2_HD __host__ __device__ __forceinline__
#define B2_D __device__ __forceinline__
#define B2_NOINLINE __device__ __noinline__
// Outlining level: 0 = transcendentals only, 1 = + universe dispatch (univ_*),
// 2 = + track-level navigation (geo_*). Levels >= 1 take addresses inside the kernel
// parameters, which therefore are declared __grid_constant__.
// Measured on the TestEm3 benchmark (profiles/README_r01.md): level 0 141.7 ms per pass,
// level 1 146.4 ms, level 2 152.1 ms -- the calls cost more (spills around them, lost
// cross-call scheduling) than the smaller code gains, so level 0 is the default.
#ifndef B2_OUTLINE_LEVEL
#    define B2_OUTLINE_LEVEL 0
#endif
// Kernel parameters are ALWAYS __grid_constant__: the big-volume geometry path (orange.cuh)
// is out of line at every level and takes `GeoParams const&`; without the qualifier every
// thread copied the whole ParamsView (1.6 kB) to its stack at kernel entry (TestEm3 pass
// 92 -> 153 ms, profiles/README_r02.md).
#define B2_GRID_CONSTANT __grid_constant__
#if B2_OUTLINE_LEVEL >= 1
#    define B2_UNIV_FN B2_NOINLINE inline
#else
#    define B2_UNIV_FN B2_D
#endif
#if B2_OUTLINE_LEVEL >= 2
#    define B2_GEO_FN B2_NOINLINE inline
#else
#    define B2_GEO_FN B2_D
#endif

// State-layout experiments (north-star (e): "coalesced, vectorised HBM gathers"), measured in
// profiles/README_r02.md; both change only how the same bytes are arranged
// Measured (TestEm3 bench, A/B in one gpurun call): packed RNG words 8.56e8 -> 8.71e8
// track-steps/s (charged along-step 222 -> 211 us, pre-step 71.7 -> 69.0 us at the saturated
// iteration): ON. Packed position/direction on top of it: 8.70e8, no change: OFF.
#ifndef B2_RNG_PACKED
#    define B2_RNG_PACKED 1
#endif
#ifndef B2_POSDIR_PACKED
#    define B2_POSDIR_PACKED 0
#endif
// Pre-step with the value-grid tables staged in shared memory (kernels.cu); needs plain loads
// for the table columns (B2_RO_LDG=0: the non-coherent path cannot address shared memory)
#ifndef B2_SMEM_GRIDS
#    define B2_SMEM_GRIDS 0
#endif
#if B2_SMEM_GRIDS && !defined(B2_RO_LDG)
#    define B2_RO_LDG 0
#endif

// Read-only column of the problem description (ParamsView). Element loads go through
// the non-coherent path (LDG.E.CONSTANT): the compiler may then move table loads across
// the stores to the track state, which it cannot prove not to alias.
// Measured (gpurun_out/variants_ro.log, TestEm3 bench, two A/B pairs): 4447 of 18.9 k loads
// become LDG.E.CONSTANT; along-step 47.1 -> 46.6 ms, pass 92.2 ms either way: the kernels wait
// on dependent loads, not on load/store ordering. Kept (no cost). B2_RO_LDG=0 restores plain loads.
#ifndef B2_RO_LDG
#    define B2_RO_LDG 1
#endif
template<class T>
struct RO
{
    T const* p;
    RO() = default;
    B2_HD RO(T const* q) : p(q) {}
    B2_HD operator T const*() const { return p; }
    template<class I>
    B2_D T operator[](I i) const
    {
#if B2_RO_LDG && defined(__CUDA_ARCH__)
        return __ldg(p + i);
#else
        return p[i];
#endif
    }
};

// Transcendentals are called out of line. Everything else in the track loop is
// force-inlined into a handful of large kernels; with ~70 call sites the inlined
// libdevice bodies (40-150 instructions each) made the charged along-step ~600 KB
// of SASS and a quarter of its stall samples were instruction-fetch misses
// (profiles/README_r01.md). Unqualified exp/log/sin/cos inside namespace b200
// resolve to these.
// (device pass only: in the host pass the names fall through to ::exp etc.)
#ifdef __CUDA_ARCH__
B2_NOINLINE inline real exp(real x) { return ::exp(x); }
B2_NOINLINE inline real log(real x) { return ::log(x); }
B2_NOINLINE inline real sin(real x) { return ::sin(x); }
B2_NOINLINE inline real cos(real x) { return ::cos(x); }
#endif

//! Track status: same numbering as the reference's TrackStatus
//! (/root/reference/src/celeritas/Types.hh:113-122)
enum TrackStatus : u8
{
    ST_INACTIVE = 0,
    ST_INITIALIZING = 1,
    ST_ALIVE = 2,
    ST_ERRORED = 3,
    ST_KILLED = 4
};

//! Track ordering (reference TrackOrder, Types.hh:151-171)
enum TrackOrder : u32
{
    ORDER_NONE = 0,
    ORDER_INIT_CHARGE = 1,
    ORDER_REINDEX_SHUFFLE = 2,  // not supported (a libstdc++ std::shuffle of the slot map)
    // reindex_*: SortTracksAction (track/SortTracksAction.cc:46-131) keeps a permutation of
    // all track slots sorted by a key, see csrc/kernels_sort.cu
    ORDER_REINDEX_STATUS = 3,
    ORDER_REINDEX_PARTICLE_TYPE = 4,
    ORDER_REINDEX_ALONG_STEP_ACTION = 5,
    ORDER_REINDEX_STEP_LIMIT_ACTION = 6,
    ORDER_REINDEX_BOTH_ACTION = 7,
    ORDER_SIZE_,
    // not a track order: key of the stable partition that compacts the step/hit output
    SORT_KEY_HITS = 100
};

struct Real3
{
    real v[3];
    B2_HD real& operator[](int i) { return v[i]; }
    B2_HD real const& operator[](int i) const { return v[i]; }
};

B2_HD Real3 make_real3(real x, real y, real z)
{
    Real3 r;
    r.v[0] = x;
    r.v[1] = y;
    r.v[2] = z;
    return r;
}

B2_HD real ipow2(real x)
{
    return x * x;
}

//! fma-chained dot product (ArrayUtils.hh:91-100)
B2_HD real dot(Real3 const& a, Real3 const& b)
{
    real r = 0;
    r = fma(a[0], b[0], r);
    r = fma(a[1], b[1], r);
    r = fma(a[2], b[2], r);
    return r;
}

//! y += a x with fma (ArrayUtils.hh:77-85)
B2_HD void axpy(real a, Real3 const& x, Real3& y)
{
    y[0] = fma(a, x[0], y[0]);
    y[1] = fma(a, x[1], y[1]);
    y[2] = fma(a, x[2], y[2]);
}

B2_HD real norm(Real3 const& a)
{
    return sqrt(dot(a, a));
}

B2_HD Real3 make_unit_vector(Real3 const& a)
{
    real scale = 1 / norm(a);
    return make_real3(a[0] * scale, a[1] * scale, a[2] * scale);
}

//! Direction from polar cosine and azimuth (ArrayUtils.hh:167-173)
B2_HD Real3 from_spherical(real costheta, real phi)
{
    real sintheta = sqrt(1 - costheta * costheta);
    return make_real3(sintheta * cos(phi), sintheta * sin(phi), costheta);
}

//! Rotate `dir` (given relative to +z) into the frame whose z axis is `rot`
//! (ArrayUtils.hh:215-263)
B2_HD Real3 rotate(Real3 const& dir, Real3 const& rot)
{
    real sintheta = sqrt(1 - ipow2(rot[2]));
    real cosphi, sinphi;
    if (sintheta >= 0.005)
    {
        real inv = 1 / sintheta;
        cosphi = rot[0] * inv;
        sinphi = rot[1] * inv;
    }
    else if (sintheta > 0)
    {
        cosphi = rot[0] / sqrt(ipow2(rot[0]) + ipow2(rot[1]));
        sinphi = sqrt(1 - ipow2(cosphi));
    }
    else
    {
        cosphi = 1;
        sinphi = 0;
    }
    Real3 r = make_real3(
        (rot[2] * dir[0] + sintheta * dir[2]) * cosphi - sinphi * dir[1],
        (rot[2] * dir[0] + sintheta * dir[2]) * sinphi + cosphi * dir[1],
        -sintheta * dir[0] + rot[2] * dir[2]);
    return make_unit_vector(r);
}

B2_HD real real_inf()
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double(0x7ff0000000000000LL);
#else
    return __builtin_huge_val();
#endif
}

B2_HD real real_max()
{
    return 1.7976931348623157e308;
}

//! Linear interpolation through (xl,yl),(xr,yr) as the reference computes it
//! (Interpolator.hh: slope = (yr-yl)/(xr-xl); y = fma(slope, x-xl, yl))
B2_HD real lerp_points(real xl, real yl, real xr, real yr, real x)
{
    real slope = (yr - yl) / (xr - xl);
    return fma(slope, x - xl, yl);
}

namespace constants
{
constexpr real pi = 3.14159265358979323846;
constexpr real c_light = 2.99792458e10;  // cm/s (CGS native)
}  // namespace constants
}  // namespace b200
