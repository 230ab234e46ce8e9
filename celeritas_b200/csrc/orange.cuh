//---------------------------------------------------------------------------//
// ORANGE navigation on structure-of-arrays track state.
//
// Implements the behaviour of the reference's OrangeTrackView
// (/root/reference/src/orange/OrangeTrackView.hh:265-842) and
// SimpleUnitTracker (/root/reference/src/orange/univ/SimpleUnitTracker.hh:162-675)
// for multi-level geometries of "simple unit" universes: point location through
// a bounding-interval hierarchy, distance-to-boundary over a volume's faces
// with simple / complex (internal surfaces) / background handling, boundary
// crossing through surface connectivity, and the quadric surface family.
//
// Per-thread scratch (face senses, intersection distances) lives in registers /
// local memory sized by the compile-time caps below instead of the reference's
// per-track global scratch arrays (OrangeTrackView.hh:1042-1067).
//---------------------------------------------------------------------------//
#pragma once

#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>

#include "views.cuh"

namespace b200
{
// Volumes with at most this many faces AND intersections take the register path: face senses
// in one u32, candidate intersections in small per-thread arrays.
constexpr int ORANGE_MAX_FACES = 32;
constexpr int ORANGE_MAX_ISECT = 32;
// Larger volumes (background volumes whose faces are all the surfaces of their unit, CMS-scale
// mother volumes with hundreds of faces) take the "big volume" path below: no per-thread
// arrays sized by the face count, only the sense words (1 bit per face). The reference sizes
// per-track global scratch from max_faces / max_intersections at run time
// (orange/OrangeData.hh:348-544, OrangeTrackView.hh:1042-1067); here the only limit is:
// (1024 faces = 128 B of sense words in the stack frame of the out-of-line search)
constexpr int ORANGE_BIG_MAX_FACES = 1024;
constexpr int ORANGE_BIG_SENSE_WORDS = ORANGE_BIG_MAX_FACES / 32;
// Lanes of a warp that reach the big-volume search together share the work (1) or every lane
// searches on its own (0: measurement knob, same results)
#ifndef B2_ORANGE_BIG_COOP
#    define B2_ORANGE_BIG_COOP 1
#endif

struct Propagation
{
    real distance;
    bool boundary;
    bool looping;
};

//---------------------------------------------------------------------------//
// SURFACES
//---------------------------------------------------------------------------//
//! Sense of a quadric value: -1 inside, 0 on, +1 outside (NaN -> outside)
B2_D int real_to_sense(real q)
{
    return static_cast<int>(!(q <= 0)) - static_cast<int>(q < 0);
}

struct SurfaceRef
{
    u8 type;
    real const* d;
};

B2_D SurfaceRef get_surface(GeoParams const& g, SimpleUnit const& u, u32 local_surface)
{
    SurfaceRef s;
    s.type = g.surface_types[u.surf_begin + local_surface];
    s.d = g.reals + g.real_ids[u.real_id_begin + local_surface];
    return s;
}

B2_D int surface_sense_all(SurfaceRef const& s, Real3 const& pos)
{
    real const* d = s.d;
    switch (s.type)
    {
        case SURF_PX: return real_to_sense(pos[0] - d[0]);
        case SURF_PY: return real_to_sense(pos[1] - d[0]);
        case SURF_PZ: return real_to_sense(pos[2] - d[0]);
        case SURF_CXC: return real_to_sense(ipow2(pos[1]) + ipow2(pos[2]) - d[0]);
        case SURF_CYC: return real_to_sense(ipow2(pos[0]) + ipow2(pos[2]) - d[0]);
        case SURF_CZC: return real_to_sense(ipow2(pos[0]) + ipow2(pos[1]) - d[0]);
        case SURF_SC: return real_to_sense(dot(pos, pos) - d[0]);
        case SURF_CX:
        {
            real u = pos[1] - d[0], v = pos[2] - d[1];
            return real_to_sense(ipow2(u) + ipow2(v) - d[2]);
        }
        case SURF_CY:
        {
            real u = pos[0] - d[0], v = pos[2] - d[1];
            return real_to_sense(ipow2(u) + ipow2(v) - d[2]);
        }
        case SURF_CZ:
        {
            real u = pos[0] - d[0], v = pos[1] - d[1];
            return real_to_sense(ipow2(u) + ipow2(v) - d[2]);
        }
        case SURF_P:
        {
            Real3 n = make_real3(d[0], d[1], d[2]);
            return real_to_sense(dot(n, pos) - d[3]);
        }
        case SURF_S:
        {
            Real3 t = make_real3(pos[0] - d[0], pos[1] - d[1], pos[2] - d[2]);
            return real_to_sense(dot(t, t) - d[3]);
        }
        case SURF_KX:
        case SURF_KY:
        case SURF_KZ:
        {
            int T = s.type - SURF_KX;
            int U = (T == 0) ? 1 : 0;
            int V = (T == 2) ? 1 : 2;
            real x = pos[T] - d[T], y = pos[U] - d[U], z = pos[V] - d[V];
            return real_to_sense((-d[3] * ipow2(x)) + ipow2(y) + ipow2(z));
        }
        case SURF_SQ:
        {
            real x = pos[0], y = pos[1], z = pos[2];
            return real_to_sense((d[0] * ipow2(x) + d[1] * ipow2(y) + d[2] * ipow2(z))
                                 + (d[3] * x + d[4] * y + d[5] * z) + (d[6]));
        }
        case SURF_GQ:
        {
            real x = pos[0], y = pos[1], z = pos[2];
            real r = (d[0] * x + d[3] * y + d[5] * z + d[6]) * x
                     + (d[1] * y + d[4] * z + d[7]) * y + (d[2] * z + d[8]) * z + d[9];
            return real_to_sense(r);
        }
        default: return 1;
    }
}

// Quadratic solver (reference surf/detail/QuadraticSolver.hh:60-230)
struct Roots
{
    real r[2];
};

constexpr real SQRT_QUADRATIC = 1e-5;
constexpr real MIN_A = SQRT_QUADRATIC * SQRT_QUADRATIC;

B2_D Roots quad_solve_off(real a_inv, real hba, real c)
{
    c *= a_inv;
    real b2_4 = ipow2(hba);
    Roots res;
    if (b2_4 > c)
    {
        real t2 = sqrt(b2_4 - c);
        res.r[0] = -hba - t2;
        res.r[1] = -hba + t2;
        if (res.r[1] <= 0)
        {
            res.r[0] = real_inf();
            res.r[1] = real_inf();
        }
        else if (res.r[0] <= 0)
        {
            res.r[0] = real_inf();
        }
    }
    else if (b2_4 == c)
    {
        res.r[0] = -hba;
        res.r[1] = real_inf();
        if (res.r[0] <= 0)
            res.r[0] = real_inf();
    }
    else
    {
        res.r[0] = real_inf();
        res.r[1] = real_inf();
    }
    return res;
}

B2_D Roots quad_solve_on(real hba)
{
    Roots res;
    res.r[0] = -2 * hba;
    res.r[1] = real_inf();
    if (res.r[0] <= 0)
        res.r[0] = real_inf();
    return res;
}

//! Solve a x^2 + 2 half_b x + c = 0 given normalised-by-a helper
B2_D Roots quad_solve(real a, real half_b, real c, bool on_surface)
{
    real a_inv = 1 / a;
    real hba = half_b * a_inv;
    return on_surface ? quad_solve_on(hba) : quad_solve_off(a_inv, hba, c);
}

B2_D Roots quad_solve_general(real a, real half_b, real c, bool on_surface)
{
    if (fabs(a) >= MIN_A)
        return quad_solve(a, half_b, c, on_surface);
    Roots res;
    res.r[0] = real_inf();
    res.r[1] = real_inf();
    if (!on_surface)
    {
        // travelling along the surface: linear equation
        if (fabs(half_b) > MIN_A)
        {
            res.r[0] = -c / (2 * half_b);
            if (res.r[0] < 0)
                res.r[0] = real_inf();
        }
    }
    return res;
}

//! Number of possible intersections for a surface type
B2_D int surface_num_isect(u8 type)
{
    return (type <= SURF_PZ || type == SURF_P) ? 1 : 2;
}

//! Distances to a surface along (pos, dir); roots are +inf when absent
B2_D Roots surface_intersect_all(SurfaceRef const& s, Real3 const& pos, Real3 const& dir, bool on_surface)
{
    real const* d = s.d;
    Roots none;
    none.r[0] = real_inf();
    none.r[1] = real_inf();
    switch (s.type)
    {
        case SURF_PX:
        case SURF_PY:
        case SURF_PZ:
        {
            int T = s.type;
            real n_dir = dir[T];
            if (!on_surface && n_dir != 0)
            {
                real dist = (d[0] - pos[T]) / n_dir;
                if (dist > 0)
                    none.r[0] = dist;
            }
            return none;
        }
        case SURF_CXC:
        case SURF_CYC:
        case SURF_CZC:
        {
            int T = s.type - SURF_CXC;
            int U = (T == 0) ? 1 : 0;
            int V = (T == 2) ? 1 : 2;
            real a = 1 - ipow2(dir[T]);
            if (a < MIN_A)
                return none;
            real u = pos[U], v = pos[V];
            real half_b = dir[U] * u + dir[V] * v;
            return quad_solve(a, half_b, ipow2(u) + ipow2(v) - d[0], on_surface);
        }
        case SURF_SC:
        {
            return quad_solve(1, dot(pos, dir), dot(pos, pos) - d[0], on_surface);
        }
        case SURF_CX:
        case SURF_CY:
        case SURF_CZ:
        {
            int T = s.type - SURF_CX;
            int U = (T == 0) ? 1 : 0;
            int V = (T == 2) ? 1 : 2;
            real a = 1 - ipow2(dir[T]);
            if (a < MIN_A)
                return none;
            real u = pos[U] - d[0], v = pos[V] - d[1];
            real half_b = dir[U] * u + dir[V] * v;
            return quad_solve(a, half_b, ipow2(u) + ipow2(v) - d[2], on_surface);
        }
        case SURF_P:
        {
            Real3 n = make_real3(d[0], d[1], d[2]);
            real n_dir = dot(n, dir);
            if (!on_surface && n_dir != 0)
            {
                real n_pos = dot(n, pos);
                real dist = (d[3] - n_pos) / n_dir;
                if (dist > 0)
                    none.r[0] = dist;
            }
            return none;
        }
        case SURF_S:
        {
            Real3 t = make_real3(pos[0] - d[0], pos[1] - d[1], pos[2] - d[2]);
            return quad_solve(1, dot(t, dir), dot(t, t) - d[3], on_surface);
        }
        case SURF_KX:
        case SURF_KY:
        case SURF_KZ:
        {
            int T = s.type - SURF_KX;
            int U = (T == 0) ? 1 : 0;
            int V = (T == 2) ? 1 : 2;
            real x = pos[T] - d[T], y = pos[U] - d[U], z = pos[V] - d[V];
            real u = dir[T], v = dir[U], w = dir[V];
            real tsq = d[3];
            real a = (-tsq * ipow2(u)) + ipow2(v) + ipow2(w);
            real half_b = (-tsq * x * u) + (y * v) + (z * w);
            real c = (-tsq * ipow2(x)) + ipow2(y) + ipow2(z);
            return quad_solve_general(a, half_b, c, on_surface);
        }
        case SURF_SQ:
        {
            real x = pos[0], y = pos[1], z = pos[2];
            real u = dir[0], v = dir[1], w = dir[2];
            real a = (d[0] * u) * u + (d[1] * v) * v + (d[2] * w) * w;
            real b = (2 * d[0] * x + d[3]) * u + (2 * d[1] * y + d[4]) * v
                     + (2 * d[2] * z + d[5]) * w;
            real c = (d[0] * x + d[3]) * x + (d[1] * y + d[4]) * y + (d[2] * z + d[5]) * z
                     + d[6];
            return quad_solve_general(a, b / 2, c, on_surface);
        }
        case SURF_GQ:
        {
            real x = pos[0], y = pos[1], z = pos[2];
            real u = dir[0], v = dir[1], w = dir[2];
            real a = (d[0] * u + d[3] * v) * u + (d[1] * v + d[4] * w) * v
                     + (d[2] * w + d[5] * u) * w;
            real b = (2 * d[0] * x + d[3] * y + d[5] * z + d[6]) * u
                     + (2 * d[1] * y + d[3] * x + d[4] * z + d[7]) * v
                     + (2 * d[2] * z + d[4] * y + d[5] * x + d[8]) * w;
            real c = ((d[0] * x + d[3] * y + d[6]) * x + (d[1] * y + d[4] * z + d[7]) * y
                      + (d[2] * z + d[5] * x + d[8]) * z + d[9]);
            return quad_solve_general(a, b / 2, c, on_surface);
        }
        default: return none;
    }
}

//! Outward normal at pos
B2_D Real3 surface_normal(SurfaceRef const& s, Real3 const& pos)
{
    real const* d = s.d;
    switch (s.type)
    {
        case SURF_PX: return make_real3(1, 0, 0);
        case SURF_PY: return make_real3(0, 1, 0);
        case SURF_PZ: return make_real3(0, 0, 1);
        case SURF_CXC: return make_unit_vector(make_real3(0, pos[1], pos[2]));
        case SURF_CYC: return make_unit_vector(make_real3(pos[0], 0, pos[2]));
        case SURF_CZC: return make_unit_vector(make_real3(pos[0], pos[1], 0));
        case SURF_SC: return make_unit_vector(pos);
        case SURF_CX: return make_unit_vector(make_real3(0, pos[1] - d[0], pos[2] - d[1]));
        case SURF_CY: return make_unit_vector(make_real3(pos[0] - d[0], 0, pos[2] - d[1]));
        case SURF_CZ: return make_unit_vector(make_real3(pos[0] - d[0], pos[1] - d[1], 0));
        case SURF_P: return make_real3(d[0], d[1], d[2]);
        case SURF_S:
            return make_unit_vector(make_real3(pos[0] - d[0], pos[1] - d[1], pos[2] - d[2]));
        case SURF_KX:
        case SURF_KY:
        case SURF_KZ:
        {
            int T = s.type - SURF_KX;
            Real3 n = make_real3(pos[0] - d[0], pos[1] - d[1], pos[2] - d[2]);
            n[T] *= -d[3];
            return make_unit_vector(n);
        }
        case SURF_SQ:
        {
            return make_unit_vector(make_real3(2 * d[0] * pos[0] + d[3],
                                               2 * d[1] * pos[1] + d[4],
                                               2 * d[2] * pos[2] + d[5]));
        }
        case SURF_GQ:
        {
            real x = pos[0], y = pos[1], z = pos[2];
            return make_unit_vector(
                make_real3(2 * d[0] * x + d[3] * y + d[5] * z + d[6],
                           2 * d[1] * y + d[3] * x + d[4] * z + d[7],
                           2 * d[2] * z + d[4] * y + d[5] * x + d[8]));
        }
        default: return make_real3(0, 0, 1);
    }
}

B2_D bool surface_simple_safety(u8 type)
{
    return type <= SURF_SC || type == SURF_P || type == SURF_S;
}

//! Safety distance to one surface (CalcSafetyDistance, SurfaceFunctors.hh)
B2_D real surface_safety_all(SurfaceRef const& s, Real3 const& pos)
{
    if (!surface_simple_safety(s.type))
        return 0;
    Real3 dir = surface_normal(s, pos);
    if (isnan(dir[0]))
        return real_inf();
    int sense = surface_sense_all(s, pos);
    if (sense > 0)
    {
        dir[0] *= -1;
        dir[1] *= -1;
        dir[2] *= -1;
    }
    else if (sense == 0)
    {
        return 0;
    }
    Roots r = surface_intersect_all(s, pos, dir, false);
    if (surface_num_isect(s.type) == 1)
        return r.r[0];
    return r.r[1] < r.r[0] ? r.r[1] : r.r[0];
}

// Surface dispatch by frequency. Every call site of the three functions below inlines what
// it calls: with all 17 quadric types inline, surface code was 53 % of the fused step's
// 61 k SASS instructions (quad_solve_off alone 10 %), and a kernel waiting on instruction
// fetch pays for code it never runs (profiles/README_r02.md). B2_SURF_OUTLINE:
//   0  everything inline (the round-1 layout)
//   1  axis-aligned planes, centred axis-aligned cylinders and general planes inline (what
//      TestEm3, simple-CMS and the CMS-scale geometry are made of); spheres, off-axis
//      cylinders, cones, simple and general quadrics are ONE out-of-line copy per kernel
//   2  only the axis-aligned planes inline
// Same arithmetic either way (the out-of-line functions call the full implementations).
#ifndef B2_SURF_OUTLINE
#    define B2_SURF_OUTLINE 1
#endif

B2_D bool surface_is_inline_type(u8 type)
{
#if B2_SURF_OUTLINE == 0
    return true;
#elif B2_SURF_OUTLINE == 1
    return type <= SURF_CZC || type == SURF_P;
#else
    return type <= SURF_PZ;
#endif
}

B2_NOINLINE inline int surface_sense_general(u32 type, real const* d, real x, real y, real z)
{
    SurfaceRef s;
    s.type = static_cast<u8>(type);
    s.d = d;
    return surface_sense_all(s, make_real3(x, y, z));
}

B2_NOINLINE inline Roots surface_intersect_general(
    u32 type, real const* d, real x, real y, real z, real dx, real dy, real dz, bool on_surface)
{
    SurfaceRef s;
    s.type = static_cast<u8>(type);
    s.d = d;
    return surface_intersect_all(s, make_real3(x, y, z), make_real3(dx, dy, dz), on_surface);
}

B2_NOINLINE inline real surface_safety_general(u32 type, real const* d, real x, real y, real z)
{
    SurfaceRef s;
    s.type = static_cast<u8>(type);
    s.d = d;
    return surface_safety_all(s, make_real3(x, y, z));
}

B2_D real select_axis(Real3 const& v, int axis)
{
    return axis == 0 ? v[0] : axis == 1 ? v[1] : v[2];
}

B2_D int surface_sense(SurfaceRef const& s, Real3 const& pos)
{
#if B2_SURF_OUTLINE == 0
    return surface_sense_all(s, pos);
#else
    real const* d = s.d;
    if (s.type <= SURF_PZ)
        return real_to_sense(select_axis(pos, s.type) - d[0]);
#    if B2_SURF_OUTLINE == 1
    if (s.type <= SURF_CZC)
    {
        int const T = s.type - SURF_CXC;
        real const u = T == 0 ? pos[1] : pos[0];
        real const v = T == 2 ? pos[1] : pos[2];
        return real_to_sense(ipow2(u) + ipow2(v) - d[0]);
    }
    if (s.type == SURF_P)
    {
        Real3 n = make_real3(d[0], d[1], d[2]);
        return real_to_sense(dot(n, pos) - d[3]);
    }
#    endif
    return surface_sense_general(s.type, d, pos[0], pos[1], pos[2]);
#endif
}

B2_D Roots surface_intersect(SurfaceRef const& s, Real3 const& pos, Real3 const& dir, bool on_surface)
{
#if B2_SURF_OUTLINE == 0
    return surface_intersect_all(s, pos, dir, on_surface);
#else
    real const* d = s.d;
    if (s.type <= SURF_PZ)
    {
        Roots none;
        none.r[0] = real_inf();
        none.r[1] = real_inf();
        real const n_dir = select_axis(dir, s.type);
        if (!on_surface && n_dir != 0)
        {
            real dist = (d[0] - select_axis(pos, s.type)) / n_dir;
            if (dist > 0)
                none.r[0] = dist;
        }
        return none;
    }
#    if B2_SURF_OUTLINE == 1
    if (s.type <= SURF_CZC)
    {
        Roots none;
        none.r[0] = real_inf();
        none.r[1] = real_inf();
        int const T = s.type - SURF_CXC;
        int const U = (T == 0) ? 1 : 0;
        int const V = (T == 2) ? 1 : 2;
        real const a = 1 - ipow2(select_axis(dir, T));
        if (a < MIN_A)
            return none;
        real const u = select_axis(pos, U), v = select_axis(pos, V);
        real const half_b = select_axis(dir, U) * u + select_axis(dir, V) * v;
        return quad_solve(a, half_b, ipow2(u) + ipow2(v) - d[0], on_surface);
    }
    if (s.type == SURF_P)
    {
        Roots none;
        none.r[0] = real_inf();
        none.r[1] = real_inf();
        Real3 n = make_real3(d[0], d[1], d[2]);
        real n_dir = dot(n, dir);
        if (!on_surface && n_dir != 0)
        {
            real n_pos = dot(n, pos);
            real dist = (d[3] - n_pos) / n_dir;
            if (dist > 0)
                none.r[0] = dist;
        }
        return none;
    }
#    endif
    return surface_intersect_general(
        s.type, d, pos[0], pos[1], pos[2], dir[0], dir[1], dir[2], on_surface);
#endif
}

B2_D real surface_safety(SurfaceRef const& s, Real3 const& pos)
{
#if B2_SURF_OUTLINE == 0
    return surface_safety_all(s, pos);
#else
    if (s.type <= SURF_PZ)
    {
        // CalcSafetyDistance for an axis-aligned plane: the distance along the normal
        real const p = select_axis(pos, s.type);
        int const sense = real_to_sense(p - s.d[0]);
        if (sense == 0)
            return 0;
        real const n_dir = sense > 0 ? real(-1) : real(1);
        real const dist = (s.d[0] - p) / n_dir;
        return dist > 0 ? dist : real_inf();
    }
#    if B2_SURF_OUTLINE == 1
    if (s.type <= SURF_CZC || s.type == SURF_P)
    {
        // surface_safety_all for these types, with the inline sense / intersection
        Real3 dir;
        if (s.type == SURF_P)
        {
            dir = make_real3(s.d[0], s.d[1], s.d[2]);
        }
        else
        {
            int const T = s.type - SURF_CXC;
            dir = make_unit_vector(make_real3(T == 0 ? real(0) : pos[0],
                                              T == 1 ? real(0) : pos[1],
                                              T == 2 ? real(0) : pos[2]));
        }
        if (isnan(dir[0]))
            return real_inf();
        int const sense = surface_sense(s, pos);
        if (sense > 0)
        {
            dir[0] *= -1;
            dir[1] *= -1;
            dir[2] *= -1;
        }
        else if (sense == 0)
        {
            return 0;
        }
        Roots const r = surface_intersect(s, pos, dir, false);
        if (s.type == SURF_P)
            return r.r[0];
        return r.r[1] < r.r[0] ? r.r[1] : r.r[0];
    }
#    endif
    return surface_safety_general(s.type, s.d, pos[0], pos[1], pos[2]);
#endif
}

//---------------------------------------------------------------------------//
// VOLUMES, LOGIC
//---------------------------------------------------------------------------//
struct VolumeRef
{
    u32 face_begin;
    u32 num_faces;
    u32 logic_begin;
    u32 logic_end;
    u32 flags;
    u32 max_isect;
};

B2_D VolumeRef get_volume(GeoParams const& g, SimpleUnit const& u, u32 local_volume)
{
    u32 rec = u.vol_begin + local_volume;
    VolumeRef v;
    v.face_begin = g.vol_face_begin[rec];
    v.num_faces = g.vol_face_end[rec] - v.face_begin;
    v.logic_begin = g.vol_logic_begin[rec];
    v.logic_end = g.vol_logic_end[rec];
    v.flags = g.vol_flags[rec];
    v.max_isect = g.vol_max_isect[rec];
    return v;
}

B2_D u32 volume_surface(GeoParams const& g, VolumeRef const& v, u32 face)
{
    return g.local_surface_ids[v.face_begin + face];
}

//! Face index of a local surface in this volume (sorted face list) or INVALID
B2_D u32 volume_find_face(GeoParams const& g, VolumeRef const& v, u32 surface)
{
    u32 lo = 0, len = v.num_faces;
    while (len > 0)
    {
        u32 half = len >> 1;
        u32 mid = lo + half;
        if (g.local_surface_ids[v.face_begin + mid] < surface)
        {
            lo = mid + 1;
            len -= half + 1;
        }
        else
            len = half;
    }
    if (lo == v.num_faces || g.local_surface_ids[v.face_begin + lo] != surface)
        return INVALID;
    return lo;
}

//! Evaluate RPN logic over a bitmask of face senses (bit i set = outside)
B2_D bool eval_logic(GeoParams const& g, VolumeRef const& v, u32 senses)
{
    u32 stack = 0;
    for (u32 i = v.logic_begin; i < v.logic_end; ++i)
    {
        u32 tok = g.logic_ints[i];
        if (tok < LOGIC_BEGIN)
        {
            stack = (stack << 1) | ((senses >> tok) & 1u);
        }
        else if (tok == LOGIC_TRUE)
        {
            stack = (stack << 1) | 1u;
        }
        else if (tok == LOGIC_OR)
        {
            stack = (stack >> 1) | (stack & 1u);
        }
        else if (tok == LOGIC_AND)
        {
            u32 t = stack & 1u;
            stack = (stack >> 1) & (t | ~u32(1));
        }
        else if (tok == LOGIC_NOT)
        {
            stack ^= 1u;
        }
    }
    return stack & 1u;
}

struct OnFace
{
    u32 face;   // INVALID when not on a face
    u8 sense;   // 0 inside, 1 outside
};

//! Senses of all faces of a volume as a bitmask; reports first "on" face
//! (SenseCalculator.hh)
B2_D u32 calc_senses(GeoParams const& g,
                     SimpleUnit const& u,
                     VolumeRef const& v,
                     Real3 const& pos,
                     OnFace& face)
{
    u32 senses = 0;
    for (u32 f = 0; f < v.num_faces; ++f)
    {
        u32 cur;
        if (f != face.face)
        {
            int ss = surface_sense(get_surface(g, u, volume_surface(g, v, f)), pos);
            cur = ss >= 0;
            if (face.face == INVALID && ss == 0)
            {
                face.face = f;
                face.sense = cur;
            }
        }
        else
        {
            cur = face.sense;
        }
        senses |= cur << f;
    }
    return senses;
}

//! RPN logic over face senses held as words (bit f of word f / 32: face f is outside)
B2_D bool eval_logic_words(GeoParams const& g, VolumeRef const& v, u32 const* words)
{
    u32 stack = 0;
    for (u32 i = v.logic_begin; i < v.logic_end; ++i)
    {
        u32 tok = g.logic_ints[i];
        if (tok < LOGIC_BEGIN)
        {
            stack = (stack << 1) | ((words[tok >> 5] >> (tok & 31u)) & 1u);
        }
        else if (tok == LOGIC_TRUE)
        {
            stack = (stack << 1) | 1u;
        }
        else if (tok == LOGIC_OR)
        {
            stack = (stack >> 1) | (stack & 1u);
        }
        else if (tok == LOGIC_AND)
        {
            u32 t = stack & 1u;
            stack = (stack >> 1) & (t | ~u32(1));
        }
        else if (tok == LOGIC_NOT)
        {
            stack ^= 1u;
        }
    }
    return stack & 1u;
}

//! Point-in-volume test for a volume of any size (SenseCalculator + LogicEvaluator). `face`
//! names a face whose sense is known (or INVALID) and receives the first face found to be
//! exactly "on"; `probe` (or INVALID) is a face whose sense is returned in `probe_sense`.
B2_NOINLINE inline bool volume_contains_big(GeoParams const& g,
                                            SimpleUnit const& u,
                                            VolumeRef const& v,
                                            Real3 const& pos,
                                            OnFace& face,
                                            u32 probe,
                                            u8& probe_sense)
{
    u32 words[ORANGE_BIG_SENSE_WORDS];
    u32 const nwords = (v.num_faces + 31u) >> 5;
    for (u32 w = 0; w < nwords; ++w)
        words[w] = 0;
    for (u32 f = 0; f < v.num_faces; ++f)
    {
        u32 cur;
        if (f != face.face)
        {
            int ss = surface_sense(get_surface(g, u, volume_surface(g, v, f)), pos);
            cur = ss >= 0;
            if (face.face == INVALID && ss == 0)
            {
                face.face = f;
                face.sense = cur;
            }
        }
        else
        {
            cur = face.sense;
        }
        words[f >> 5] |= cur << (f & 31u);
    }
    if (probe != INVALID)
        probe_sense = (words[probe >> 5] >> (probe & 31u)) & 1u;
    return eval_logic_words(g, v, words);
}

//! Point-in-volume test, any volume size
B2_D bool volume_contains(GeoParams const& g,
                          SimpleUnit const& u,
                          VolumeRef const& v,
                          Real3 const& pos,
                          OnFace& face)
{
    if (v.num_faces <= u32(ORANGE_MAX_FACES))
    {
        u32 senses = calc_senses(g, u, v, pos, face);
        return eval_logic(g, v, senses);
    }
    u8 unused = 0;
    return volume_contains_big(g, u, v, pos, face, INVALID, unused);
}

//---------------------------------------------------------------------------//
// BIH TRAVERSAL (reference detail/BIHTraverser.hh:103-295)
//---------------------------------------------------------------------------//
B2_D bool bbox_contains(float const* bb, Real3 const& p)
{
    // BoundingBox<float> is_inside: lower <= p <= upper per axis
    for (int ax = 0; ax < 3; ++ax)
    {
        if (!(p[ax] >= bb[ax] && p[ax] <= bb[3 + ax]))
            return false;
    }
    return true;
}

template<class F>
B2_D u32 bih_find_volume(GeoParams const& g, SimpleUnit const& u, Real3 const& pos, F&& is_inside)
{
    u32 const leaf_offset = u.inner_count;
    u32 previous = INVALID;
    u32 current = 0;
    do
    {
        u32 next;
        if (current >= leaf_offset)
        {
            // leaf: test its volumes
            u32 leaf = u.leaf_begin + (current - leaf_offset);
            u32 vb = g.bih_leaf_vol_begin[leaf];
            u32 ve = g.bih_leaf_vol_end[leaf];
            for (u32 i = vb; i < ve; ++i)
            {
                u32 id = g.bih_local_volume_ids[i];
                if (bbox_contains(g.bih_bboxes + 6 * (u.bbox_begin + id), pos) && is_inside(id))
                    return id;
            }
            next = previous;
        }
        else
        {
            u32 node = u.inner_begin + current;
            u32 parent = g.bih_inner_parent[node];
            u32 axis = g.bih_inner_axis[node];
            u32 lchild = g.bih_inner_left_child[node];
            u32 rchild = g.bih_inner_right_child[node];
            real pp = pos[axis];
            if (previous == parent)
            {
                // visit left if the point is below the left plane, else right
                if (pp < g.bih_inner_left_pos[node])
                    next = lchild;
                else
                    next = rchild;
            }
            else if (previous == lchild)
            {
                if (g.bih_inner_right_pos[node] < pp)
                    next = rchild;
                else
                    next = parent;
            }
            else
            {
                next = parent;
            }
        }
        previous = current;
        current = next;
    } while (current != INVALID);

    for (u32 i = 0; i < u.inf_count; ++i)
    {
        u32 id = g.bih_local_volume_ids[u.inf_begin + i];
        if (is_inside(id))
            return id;
    }
    return INVALID;
}

//---------------------------------------------------------------------------//
// SIMPLE UNIT TRACKER
//---------------------------------------------------------------------------//
struct LocalState
{
    Real3 pos;
    Real3 dir;
    u32 volume;
    u32 surface;     // local surface id or INVALID
    u8 sense;
};

struct Initialization
{
    u32 volume;
    u32 surface;
    u8 sense;
};

struct Intersection
{
    u32 surface;  // INVALID = none
    u8 sense;
    real distance;
};

B2_D Initialization unit_initialize(GeoParams const& g, SimpleUnit const& u, Real3 const& pos)
{
    bool on_surface = false;
    auto is_inside = [&](u32 id) -> bool {
        VolumeRef vol = get_volume(g, u, id);
        OnFace face{INVALID, 0};
        bool const inside = volume_contains(g, u, vol, pos, face);
        on_surface = (face.face != INVALID);
        return inside;
    };
    u32 id = bih_find_volume(g, u, pos, is_inside);
    if (on_surface)
        id = INVALID;
    else if (id == INVALID)
        id = u.background;
    return Initialization{id, INVALID, 0};
}

B2_D Initialization unit_cross_boundary(GeoParams const& g, SimpleUnit const& u, LocalState const& st)
{
    u32 on_surf = INVALID;
    u8 on_sense = 0;
    auto is_inside = [&](u32 id) -> bool {
        if (id == st.volume)
            return false;
        VolumeRef vol = get_volume(g, u, id);
        OnFace face{volume_find_face(g, vol, st.surface), st.sense};
        if (volume_contains(g, u, vol, st.pos, face))
        {
            on_surf = (face.face != INVALID) ? volume_surface(g, vol, face.face) : INVALID;
            on_sense = face.sense;
            return true;
        }
        return false;
    };
    u32 conn = u.conn_begin + st.surface;
    u32 nb = g.conn_begin[conn], ne = g.conn_end[conn];
    if (ne - nb < 3)
    {
        for (u32 i = nb; i < ne; ++i)
        {
            u32 id = g.local_volume_ids[i];
            if (is_inside(id))
                return Initialization{id, on_surf, on_sense};
        }
    }
    else
    {
        u32 id = bih_find_volume(g, u, st.pos, is_inside);
        if (id != INVALID)
            return Initialization{id, on_surf, on_sense};
    }
    return Initialization{u.background, st.surface, st.sense};
}

//---------------------------------------------------------------------------//
// BIG VOLUMES: WARP-COOPERATIVE DISTANCE SEARCH
//
// SimpleUnitTracker::intersect_impl (univ/SimpleUnitTracker.hh:390-455) fills per-track
// scratch arrays with every valid intersection of every face, sorts them by distance, and
// walks them in order (complex_intersect :506-560, background_intersect :585-640). For a
// volume with hundreds of faces that is hundreds of dependent quadric solves in one thread
// plus a sort through global memory. Here the lanes of a warp that arrive at the search
// together AND hold the same track (cooperative_groups: coalesced_threads, partitioned by
// query; see unit_intersect_big for which lanes those are and why not more) share the search:
// the group deals the FACES out to its lanes (face f -> lane f mod n), every lane solves its
// faces, and the group reduces to the next crossing in (distance, face, root) order. The ordered walk of the reference becomes "extract the next
// minimum, strictly after the previous one": no intersection array, no sort, no scratch
// memory; a walk normally ends after one or two crossings. Sense words are built the same
// way (one OR-reduction per 32 faces), and a background volume's neighbour tests are dealt
// out by neighbour. Results do not depend on which lanes happen to cooperate: every
// per-face value is computed by the same instruction sequence, the reductions are exact
// integer minima, and ties are broken by (face, root), the order of the reference's
// intersection array.
//---------------------------------------------------------------------------//
namespace cg = cooperative_groups;

//! A group of one lane: the same search without cooperation
struct SoloGroup
{
    B2_D u32 size() const { return 1; }
    B2_D u32 thread_rank() const { return 0; }
};
B2_D u64 group_min(SoloGroup const&, u64 v) { return v; }
B2_D u32 group_min(SoloGroup const&, u32 v) { return v; }
B2_D u32 group_or(SoloGroup const&, u32 v) { return v; }
B2_D u64 group_min(cg::coalesced_group const& grp, u64 v)
{
    return cg::reduce(grp, v, cg::less<u64>());
}
B2_D u32 group_min(cg::coalesced_group const& grp, u32 v)
{
    return cg::reduce(grp, v, cg::less<u32>());
}
B2_D u32 group_or(cg::coalesced_group const& grp, u32 v)
{
    return cg::reduce(grp, v, cg::bit_or<u32>());
}

//! One track's search by all lanes of `grp` (every argument is uniform over the group)
template<class Group>
B2_D Intersection big_volume_search(Group const& grp,
                                    GeoParams const& g,
                                    SimpleUnit const& u,
                                    LocalState const& st,
                                    bool limited,
                                    real max_dist,
                                    u32* words)
{
    u32 const n = grp.size();
    u32 const rank = grp.thread_rank();
    VolumeRef const vol = get_volume(g, u, st.volume);
    u32 const on_face = (st.surface != INVALID) ? volume_find_face(g, vol, st.surface) : INVALID;
    constexpr u64 none = ~u64(0);

    // Next valid crossing strictly after (prev_bits, prev_key) in (distance, face, root)
    // order. Valid distances are positive, so their bit patterns order like the doubles.
    auto next_crossing = [&](u64 prev_bits, u32 prev_key, u64& out_bits, u32& out_key) {
        u64 best_bits = none;
        u32 best_key = INVALID;
        for (u32 f = rank; f < vol.num_faces; f += n)
        {
            SurfaceRef const sr = get_surface(g, u, volume_surface(g, vol, f));
            bool const on = (f == on_face);
            int const nroots = surface_num_isect(sr.type);
            if (nroots == 1 && on)
                continue;
            Roots const r = surface_intersect(sr, st.pos, st.dir, on);
            for (int k = 0; k < nroots; ++k)
            {
                real const d = r.r[k];
                bool const valid = limited ? (d <= max_dist) : (d < real_max());
                if (!valid)
                    continue;
                u64 const bits = static_cast<u64>(__double_as_longlong(d));
                u32 const key = 2u * f + u32(k);
                bool const after = bits > prev_bits || (bits == prev_bits && key > prev_key);
                bool const better = bits < best_bits || (bits == best_bits && key < best_key);
                if (after && better)
                {
                    best_bits = bits;
                    best_key = key;
                }
            }
        }
        out_bits = group_min(grp, best_bits);
        out_key = group_min(grp, best_bits == out_bits ? best_key : INVALID);
        return out_bits != none;
    };

    Intersection result{INVALID, 0, real_inf()};
    u64 bits = 0;
    u32 key = 0;
    if (!(vol.flags & (VOL_INTERNAL_SURFACES | VOL_IMPLICIT)))
    {
        // simple_intersect: the nearest crossing leaves the volume
        if (next_crossing(0, 0, bits, key))
        {
            u32 const surface = volume_surface(g, vol, key >> 1);
            result.surface = surface;
            result.sense = (surface == st.surface)
                               ? st.sense
                               : static_cast<u8>(surface_sense(get_surface(g, u, surface), st.pos)
                                                 >= 0);
            result.distance = __longlong_as_double(static_cast<long long>(bits));
        }
    }
    else if (vol.flags & VOL_INTERNAL_SURFACES)
    {
        // complex_intersect: cross surfaces in order until the logic says "outside"
        // (`words`: ORANGE_BIG_SENSE_WORDS of caller-provided scratch)
        u32 const nwords = (vol.num_faces + 31u) >> 5;
        for (u32 w = 0; w < nwords; ++w)
        {
            u32 mine = 0;
            u32 const end = (32u * w + 32u < vol.num_faces) ? 32u * w + 32u : vol.num_faces;
            for (u32 f = 32u * w + rank; f < end; f += n)
            {
                u32 const cur
                    = (f == on_face)
                          ? u32(st.sense)
                          : u32(surface_sense(get_surface(g, u, volume_surface(g, vol, f)), st.pos)
                                >= 0);
                mine |= cur << (f & 31u);
            }
            words[w] = group_or(grp, mine);
        }
        while (next_crossing(bits, key, bits, key))
        {
            u32 const f = key >> 1;
            words[f >> 5] ^= 1u << (f & 31u);
            u32 const new_sense = (words[f >> 5] >> (f & 31u)) & 1u;
            if (!eval_logic_words(g, vol, words))
            {
                result.surface = volume_surface(g, vol, f);
                result.sense = static_cast<u8>(new_sense ^ 1u);
                result.distance = __longlong_as_double(static_cast<long long>(bits));
                break;
            }
        }
    }
    else
    {
        // background_intersect: the first crossing that enters a neighbouring volume
        real bump = g.tol_abs;
        for (int ax = 0; ax < 3; ++ax)
        {
            real t = g.tol_rel * fabs(st.pos[ax]);
            bump = t > bump ? t : bump;
        }
        while (result.surface == INVALID && next_crossing(bits, key, bits, key))
        {
            real const dist = __longlong_as_double(static_cast<long long>(bits));
            u32 const surface = volume_surface(g, vol, key >> 1);
            Real3 pos = st.pos;
            axpy(dist + bump, st.dir, pos);
            u32 const conn = u.conn_begin + surface;
            u32 const nb = g.conn_begin[conn], ne = g.conn_end[conn];
            // neighbours dealt out to the lanes; the first one (in connectivity order)
            // that contains the bumped point wins
            u32 found = INVALID;
            for (u32 i = nb + rank; i < ne && found == INVALID; i += n)
            {
                u32 const vid = g.local_volume_ids[i];
                VolumeRef const nv = get_volume(g, u, vid);
                u32 const nf = volume_find_face(g, nv, surface);
                OnFace face{INVALID, 0};
                u8 sense = 0;
                bool inside;
                if (nv.num_faces <= u32(ORANGE_MAX_FACES))
                {
                    u32 const senses = calc_senses(g, u, nv, pos, face);
                    sense = (senses >> nf) & 1u;
                    inside = eval_logic(g, nv, senses);
                }
                else
                {
                    inside = volume_contains_big(g, u, nv, pos, face, nf, sense);
                }
                if (inside)
                    found = ((i - nb) << 1) | u32(sense);
            }
            found = group_min(grp, found);
            if (found != INVALID)
            {
                result.distance = dist;
                result.surface = surface;
                result.sense = static_cast<u8>((found & 1u) ^ 1u);
            }
        }
    }
    if (limited && result.surface == INVALID)
        result.distance = max_dist;
    return result;
}

//! Distance to boundary in a volume too large for the register path. Out of line: the step
//! kernels' register budgets are set by the small-volume path.
//!
//! Which lanes cooperate: the ones that hold the SAME query, i.e. the lanes of a warp that
//! carries one track in all its lanes (the device-resident loop's one-warp-per-track mode,
//! csrc/tail.cu). Lanes with different tracks each search on their own, side by side.
//! Measured on this geometry class (profiles/README_r02.md, "big volumes"): serving the
//! different tracks of a full warp one after the other with 32-way parallelism has the same
//! critical path as 32 lanes searching side by side plus the reductions, and is 2.3 x slower
//! (k_geo_trace, 113-face volume: 92.7 ms against 39.7 ms for 262 144 rays); borrowing lanes
//! only pays when they would otherwise repeat the same work.
B2_NOINLINE inline Intersection unit_intersect_big(GeoParams const& g,
                                                   SimpleUnit const& u,
                                                   LocalState const& st,
                                                   bool limited,
                                                   real max_dist)
{
    u32 words[ORANGE_BIG_SENSE_WORDS];
#if B2_ORANGE_BIG_COOP
    // Partition the lanes that are here by (a hash of) their query
    cg::coalesced_group const here = cg::coalesced_threads();
    u64 h = static_cast<u64>(__double_as_longlong(st.pos[0]));
    h = h * 0x9e3779b97f4a7c15ull + static_cast<u64>(__double_as_longlong(st.pos[1]));
    h = h * 0x9e3779b97f4a7c15ull + static_cast<u64>(__double_as_longlong(st.pos[2]));
    h = h * 0x9e3779b97f4a7c15ull + static_cast<u64>(__double_as_longlong(st.dir[0]));
    h = h * 0x9e3779b97f4a7c15ull + static_cast<u64>(__double_as_longlong(st.dir[1]));
    h = h * 0x9e3779b97f4a7c15ull + static_cast<u64>(__double_as_longlong(max_dist));
    h = h * 0x9e3779b97f4a7c15ull + (u64(st.volume) << 32 | u64(st.surface));
    cg::coalesced_group const grp = cg::labeled_partition(here, static_cast<u32>(h ^ (h >> 32)));
    if (grp.size() > 1)
    {
        // every member must hold exactly the query of the first one (hash collisions)
        u32 const my_unit = static_cast<u32>(&u - g.simple_units);
        bool same = true;
        for (int k = 0; k < 3; ++k)
        {
            same = same && grp.shfl(st.pos[k], 0) == st.pos[k]
                   && grp.shfl(st.dir[k], 0) == st.dir[k];
        }
        same = same && grp.shfl(st.volume, 0) == st.volume
               && grp.shfl(st.surface, 0) == st.surface
               && grp.shfl(u32(st.sense), 0) == u32(st.sense) && grp.shfl(my_unit, 0) == my_unit
               && grp.shfl(u32(limited), 0) == u32(limited)
               && (grp.shfl(max_dist, 0) == max_dist || !limited);
        if (grp.all(same))
            return big_volume_search(grp, g, u, st, limited, max_dist, words);
    }
#endif
    return big_volume_search(SoloGroup{}, g, u, st, limited, max_dist, words);
}

//! Distance to the boundary of the current volume. max_dist < 0 = unlimited.
B2_D Intersection unit_intersect(GeoParams const& g,
                                 SimpleUnit const& u,
                                 LocalState const& st,
                                 bool limited,
                                 real max_dist)
{
    VolumeRef vol = get_volume(g, u, st.volume);
    if (vol.num_faces > u32(ORANGE_MAX_FACES) || vol.max_isect > u32(ORANGE_MAX_ISECT))
        return unit_intersect_big(g, u, st, limited, max_dist);
    u32 on_face = (st.surface != INVALID) ? volume_find_face(g, vol, st.surface) : INVALID;
    bool const simple = !(vol.flags & (VOL_INTERNAL_SURFACES | VOL_IMPLICIT));

    if (simple)
    {
        // Nearest intersection over the faces, kept in registers. Same choice as the
        // general path below (first of the smallest distances in face order) without
        // its per-thread arrays: those are indexed by a per-thread count, so every
        // lane's access landed in a different local-memory sector (ncu: 1.1 useful
        // bytes per 32-byte sector, profiles/README_r01.md).
        real best_dist = real_inf();
        u32 best_face = INVALID;
        for (u32 f = 0; f < vol.num_faces; ++f)
        {
            SurfaceRef s = get_surface(g, u, volume_surface(g, vol, f));
            bool on = (f == on_face);
            int nroots = surface_num_isect(s.type);
            if (nroots == 1 && on)
                continue;
            Roots r = surface_intersect(s, st.pos, st.dir, on);
            for (int k = 0; k < nroots; ++k)
            {
                real d = r.r[k];
                bool valid = limited ? (d <= max_dist) : (d < real_max());
                if (valid && (best_face == INVALID || d < best_dist))
                {
                    best_dist = d;
                    best_face = f;
                }
            }
        }
        Intersection result{INVALID, 0, real_inf()};
        if (best_face != INVALID)
        {
            u32 surface = volume_surface(g, vol, best_face);
            u8 cur_sense;
            if (surface == st.surface)
                cur_sense = st.sense;
            else
                cur_sense = surface_sense(get_surface(g, u, surface), st.pos) >= 0;
            result.surface = surface;
            result.sense = cur_sense;
            result.distance = best_dist;
        }
        if (limited && result.surface == INVALID)
            result.distance = max_dist;
        return result;
    }

    real dist[ORANGE_MAX_ISECT];
    u8 face_of[ORANGE_MAX_ISECT];
    u32 num_isect = 0;
    for (u32 f = 0; f < vol.num_faces; ++f)
    {
        SurfaceRef s = get_surface(g, u, volume_surface(g, vol, f));
        bool on = (f == on_face);
        int nroots = surface_num_isect(s.type);
        if (nroots == 1 && on)
            continue;
        Roots r = surface_intersect(s, st.pos, st.dir, on);
        for (int k = 0; k < nroots; ++k)
        {
            real d = r.r[k];
            bool valid = limited ? (d <= max_dist) : (d < real_max());
            if (valid)
            {
                dist[num_isect] = d;
                face_of[num_isect] = f;
                ++num_isect;
            }
        }
    }
    Intersection result{INVALID, 0, real_inf()};
    if (num_isect == 0)
    {
        // fallthrough
    }
    else if (simple)
    {
        u32 best = 0;
        for (u32 i = 1; i < num_isect; ++i)
            if (dist[i] < dist[best])
                best = i;
        u32 surface = volume_surface(g, vol, face_of[best]);
        u8 cur_sense;
        if (surface == st.surface)
            cur_sense = st.sense;
        else
            cur_sense = surface_sense(get_surface(g, u, surface), st.pos) >= 0;
        result.surface = surface;
        result.sense = cur_sense;
        result.distance = dist[best];
    }
    else
    {
        // Sort intersection indices by distance (insertion sort: tiny N)
        u8 order[ORANGE_MAX_ISECT];
        for (u32 i = 0; i < num_isect; ++i)
        {
            u32 j = i;
            while (j > 0 && dist[i] < dist[order[j - 1]])
            {
                order[j] = order[j - 1];
                --j;
            }
            order[j] = i;
        }
        if (vol.flags & VOL_INTERNAL_SURFACES)
        {
            OnFace face{on_face, st.sense};
            u32 senses = calc_senses(g, u, vol, st.pos, face);
            for (u32 k = 0; k < num_isect; ++k)
            {
                u32 isect = order[k];
                u32 f = face_of[isect];
                senses ^= (1u << f);
                u32 new_sense = (senses >> f) & 1u;
                if (!eval_logic(g, vol, senses))
                {
                    result.surface = volume_surface(g, vol, f);
                    result.sense = new_sense ^ 1u;
                    result.distance = dist[isect];
                    break;
                }
            }
        }
        else if (vol.flags & VOL_IMPLICIT)
        {
            // Background volume: faces are *all* unit surfaces (face id ==
            // local surface id); test neighbours just past each crossing
            real bump = g.tol_abs;
            for (int ax = 0; ax < 3; ++ax)
            {
                real t = g.tol_rel * fabs(st.pos[ax]);
                bump = t > bump ? t : bump;
            }
            for (u32 k = 0; k < num_isect && result.surface == INVALID; ++k)
            {
                u32 isect = order[k];
                u32 surface = volume_surface(g, vol, face_of[isect]);
                Real3 pos = st.pos;
                axpy(dist[isect] + bump, st.dir, pos);
                u32 conn = u.conn_begin + surface;
                for (u32 i = g.conn_begin[conn]; i < g.conn_end[conn]; ++i)
                {
                    u32 vid = g.local_volume_ids[i];
                    VolumeRef nv = get_volume(g, u, vid);
                    OnFace face{INVALID, 0};
                    u32 senses = calc_senses(g, u, nv, pos, face);
                    if (eval_logic(g, nv, senses))
                    {
                        u32 nf = volume_find_face(g, nv, surface);
                        result.distance = dist[isect];
                        result.surface = surface;
                        result.sense = ((senses >> nf) & 1u) ^ 1u;
                        break;
                    }
                }
            }
        }
    }
    if (limited && result.surface == INVALID)
        result.distance = max_dist;
    return result;
}

B2_D real unit_safety(GeoParams const& g, SimpleUnit const& u, Real3 const& pos, u32 volid)
{
    VolumeRef vol = get_volume(g, u, volid);
    if (!(vol.flags & VOL_SIMPLE_SAFETY))
        return 0;
    real result = real_inf();
    for (u32 f = 0; f < vol.num_faces; ++f)
    {
        real d = surface_safety(get_surface(g, u, volume_surface(g, vol, f)), pos);
        result = d < result ? d : result;
    }
    return result;
}

B2_D u32 unit_daughter(GeoParams const& g, SimpleUnit const& u, u32 volid)
{
    return g.vol_daughter[u.vol_begin + volid];
}

//---------------------------------------------------------------------------//
// RECT ARRAY TRACKER (reference univ/RectArrayTracker.hh:120-360)
// Record (16 u32): daughter_begin, daughter_count, dims[3], {grid begin, end} x 3,
// surface indexer offsets[4], pad. Local volume = (ix * ny + iy) * nz + iz; local
// surfaces are the grid planes, numbered per axis through the ragged offsets.
//---------------------------------------------------------------------------//
struct RectArrayRef
{
    u32 const* r;
    real const* reals;
    B2_D u32 dim(int ax) const { return r[2 + ax]; }
    B2_D real const* grid(int ax) const { return reals + r[5 + 2 * ax]; }
    B2_D u32 grid_size(int ax) const { return r[6 + 2 * ax] - r[5 + 2 * ax]; }
    B2_D u32 surf_offset(int i) const { return r[11 + i]; }
    B2_D void coords(u32 vol, u32 c[3]) const
    {
        c[2] = vol % dim(2);
        vol = (vol - c[2]) / dim(2);
        c[1] = vol % dim(1);
        vol = (vol - c[1]) / dim(1);
        c[0] = vol;
    }
    B2_D u32 index(u32 const c[3]) const { return (c[0] * dim(1) + c[1]) * dim(2) + c[2]; }
    B2_D u32 surface_axis(u32 surf) const
    {
        u32 i = 0;
        while (surf >= surf_offset(i + 1))
            ++i;
        return i;
    }
};

B2_D RectArrayRef get_rect_array(GeoParams const& g, u32 index)
{
    return RectArrayRef{g.rect_arrays + 16 * index, g.reals};
}

B2_D Initialization rect_initialize(RectArrayRef const& ra, Real3 const& pos)
{
    u32 c[3];
    for (int ax = 0; ax < 3; ++ax)
    {
        real const* grid = ra.grid(ax);
        u32 const n = ra.grid_size(ax);
        real const p = pos[ax];
        if (p < grid[0] || p > grid[n - 1])
            return Initialization{INVALID, INVALID, 0};
        // NonuniformGrid::find
        u32 lo = 0, len = n;
        while (len > 0)
        {
            u32 half = len >> 1;
            u32 mid = lo + half;
            if (grid[mid] < p)
            {
                lo = mid + 1;
                len -= half + 1;
            }
            else
                len = half;
        }
        if (p != grid[lo])
            --lo;
        if (grid[lo] == p)
            return Initialization{INVALID, INVALID, 0};
        c[ax] = lo;
    }
    return Initialization{ra.index(c), INVALID, 0};
}

B2_D Initialization rect_cross_boundary(RectArrayRef const& ra, LocalState const& st)
{
    u32 c[3];
    ra.coords(st.volume, c);
    u32 ax = ra.surface_axis(st.surface);
    // sense outside (1) = increasing coordinate
    c[ax] += (st.sense == 1) ? 1u : u32(-1);
    return Initialization{ra.index(c), st.surface, st.sense};
}

B2_D Intersection rect_intersect(RectArrayRef const& ra, LocalState const& st, bool limited, real max_dist)
{
    u32 c[3];
    ra.coords(st.volume, c);
    Intersection result{INVALID, 0, real_inf()};
    for (int ax = 0; ax < 3; ++ax)
    {
        real dir = st.dir[ax];
        if (dir == 0)
            continue;
        u32 target_coord = c[ax] + (dir > 0 ? 1u : 0u);
        real target_value = ra.grid(ax)[target_coord];
        real dist = (target_value - st.pos[ax]) / st.dir[ax];
        bool valid = limited ? (dist <= max_dist) : (dist < real_max());
        if (dist > 0 && valid && dist < result.distance)
        {
            result.distance = dist;
            result.sense = dir > 0 ? 0 : 1;
            result.surface = ra.surf_offset(ax) + target_coord;
        }
    }
    if (limited && result.surface == INVALID)
        result.distance = max_dist;
    return result;
}

B2_D real rect_safety(RectArrayRef const& ra, Real3 const& pos, u32 vol)
{
    u32 c[3];
    ra.coords(vol, c);
    real min_dist = real_inf();
    for (int ax = 0; ax < 3; ++ax)
    {
        real const* grid = ra.grid(ax);
        for (u32 i = 0; i < 2; ++i)
        {
            real d = fabs(pos[ax] - grid[c[ax] + i]);
            min_dist = d < min_dist ? d : min_dist;
        }
    }
    return min_dist;
}

//---------------------------------------------------------------------------//
// UNIVERSE DISPATCH (reference univ/TrackerVisitor.hh:63-79)
//
// These are the out-of-line entry points of the geometry: everything above is
// inlined into them once, and the (many) call sites in the step kernels share the
// code. `g` must be addressable: kernels declare their ParamsView __grid_constant__.
//---------------------------------------------------------------------------//
B2_UNIV_FN Initialization univ_initialize(GeoParams const& g, u32 uid, Real3 const& pos)
{
    u32 idx = g.universe_index[uid];
    if (g.universe_type[uid] == UNIV_SIMPLE)
        return unit_initialize(g, g.simple_units[idx], pos);
    return rect_initialize(get_rect_array(g, idx), pos);
}

B2_UNIV_FN Initialization univ_cross_boundary(GeoParams const& g, u32 uid, LocalState const& st)
{
    u32 idx = g.universe_index[uid];
    if (g.universe_type[uid] == UNIV_SIMPLE)
        return unit_cross_boundary(g, g.simple_units[idx], st);
    return rect_cross_boundary(get_rect_array(g, idx), st);
}

B2_UNIV_FN Intersection univ_intersect(GeoParams const& g, u32 uid, LocalState const& st, bool limited, real max_dist)
{
    u32 idx = g.universe_index[uid];
    if (g.universe_type[uid] == UNIV_SIMPLE)
        return unit_intersect(g, g.simple_units[idx], st, limited, max_dist);
    return rect_intersect(get_rect_array(g, idx), st, limited, max_dist);
}

B2_UNIV_FN real univ_safety(GeoParams const& g, u32 uid, Real3 const& pos, u32 vol)
{
    u32 idx = g.universe_index[uid];
    if (g.universe_type[uid] == UNIV_SIMPLE)
        return unit_safety(g, g.simple_units[idx], pos, vol);
    return rect_safety(get_rect_array(g, idx), pos, vol);
}

B2_D u32 univ_daughter(GeoParams const& g, u32 uid, u32 vol)
{
    u32 idx = g.universe_index[uid];
    if (g.universe_type[uid] == UNIV_SIMPLE)
        return unit_daughter(g, g.simple_units[idx], vol);
    return g.rect_arrays[16 * idx] + vol;
}

B2_UNIV_FN Real3 univ_normal(GeoParams const& g, u32 uid, Real3 const& pos, u32 surf)
{
    u32 idx = g.universe_index[uid];
    if (g.universe_type[uid] == UNIV_SIMPLE)
        return surface_normal(get_surface(g, g.simple_units[idx], surf), pos);
    Real3 n = make_real3(0, 0, 0);
    n[get_rect_array(g, idx).surface_axis(surf)] = 1;
    return n;
}

//---------------------------------------------------------------------------//
// TRANSFORMS
//---------------------------------------------------------------------------//
B2_D void transform_down(GeoParams const& g, u32 transform_id, Real3& pos, Real3& dir)
{
    u8 type = g.transform_type[transform_id];
    real const* d = g.reals + g.transform_offset[transform_id];
    if (type == TRANSFORM_TRANSLATION)
    {
        pos[0] -= d[0];
        pos[1] -= d[1];
        pos[2] -= d[2];
    }
    else if (type == TRANSFORM_TRANSFORMATION)
    {
        // r_d = R^T (r_p - t): rot stored row-major in d[0..8], tra in d[9..11]
        Real3 t = make_real3(pos[0] - d[9], pos[1] - d[10], pos[2] - d[11]);
        Real3 np, nd;
        for (int i = 0; i < 3; ++i)
        {
            np[i] = 0;
            nd[i] = 0;
            for (int j = 0; j < 3; ++j)
            {
                np[i] = fma(d[3 * j + i], t[j], np[i]);
                nd[i] = fma(d[3 * j + i], dir[j], nd[i]);
            }
        }
        pos = np;
        dir = nd;
    }
}

B2_D Real3 rotate_up(GeoParams const& g, u32 transform_id, Real3 const& dir)
{
    u8 type = g.transform_type[transform_id];
    if (type != TRANSFORM_TRANSFORMATION)
        return dir;
    real const* d = g.reals + g.transform_offset[transform_id];
    Real3 nd;
    for (int i = 0; i < 3; ++i)
    {
        nd[i] = 0;
        for (int j = 0; j < 3; ++j)
            nd[i] = fma(d[3 * i + j], dir[j], nd[i]);
    }
    return nd;
}

//---------------------------------------------------------------------------//
// TRACK VIEW
//---------------------------------------------------------------------------//
// Out-of-line entry points of the track-level navigation (defined after GeoTrack).
// The step kernels call these from many places (MSC safety, linear and field
// propagation, boundary crossing, scattering); sharing one copy per kernel keeps the
// kernels within reach of the instruction cache. `g` and `s` are references into the
// kernels' __grid_constant__ parameters.
B2_GEO_FN Propagation
geo_find_next_step(GeoParams const& g, StateView const& s, u32 slot, bool limited, real max_step);
B2_GEO_FN real geo_find_safety(GeoParams const& g, StateView const& s, u32 slot);
B2_GEO_FN void
geo_set_dir(GeoParams const& g, StateView const& s, u32 slot, real dx, real dy, real dz);
B2_GEO_FN bool geo_cross_boundary(GeoParams const& g, StateView const& s, u32 slot);
B2_GEO_FN bool geo_initialize(GeoParams const& g,
                                       StateView const& s,
                                       u32 slot,
                                       Real3 const& pos,
                                       Real3 const& dir);
B2_GEO_FN void
geo_move_internal_pos(GeoParams const& g, StateView const& s, u32 slot, real x, real y, real z);

// COOP = warp-cooperative mode: all 32 lanes of a warp hold the SAME track (same slot, same
// registers, same control flow) and split the per-face work of the distance and safety
// searches among them (coop_find_next_step / coop_find_safety below). Everything else is
// executed redundantly by every lane: loads are broadcasts, stores write the same value to
// the same address. Used for iterations with fewer tracks than warps (csrc/tail.cu).
template<bool COOP>
struct GeoTrackT;
template<bool COOP>
B2_D Propagation coop_find_next_step(GeoTrackT<COOP>& t, bool limited, real max_step);
template<bool COOP>
B2_D real coop_find_safety(GeoTrackT<COOP>& t);

template<bool COOP>
struct GeoTrackT
{
    GeoParams const& g;
    StateView const& s;
    u32 slot;
    bool failed;

    B2_D GeoTrackT(ParamsView const& p, StateView const& st, u32 sl)
        : g(p.geo), s(st), slot(sl), failed(false)
    {
    }
    B2_D GeoTrackT(GeoParams const& gp, StateView const& st, u32 sl)
        : g(gp), s(st), slot(sl), failed(false)
    {
    }

    // --- per-level accessors
    B2_D u32 lidx(u32 level) const { return level * s.num_slots + slot; }
    // Layout of geo_pos / geo_dir (24 B per level and slot either way):
    //   B2_POSDIR_PACKED 0  three columns [3][level][slot]
    //   B2_POSDIR_PACKED 1  double2 {x, y}[level][slot] followed by z[level][slot]
    B2_D static Real3 load3(real const* base, u32 n, u32 i)
    {
#if B2_POSDIR_PACKED
        double2 const xy = reinterpret_cast<double2 const*>(base)[i];
        return make_real3(xy.x, xy.y, base[2 * size_t(n) + i]);
#else
        return make_real3(base[i], base[n + i], base[2 * n + i]);
#endif
    }
    B2_D static void store3(real* base, u32 n, u32 i, Real3 const& v)
    {
#if B2_POSDIR_PACKED
        reinterpret_cast<double2*>(base)[i] = make_double2(v[0], v[1]);
        base[2 * size_t(n) + i] = v[2];
#else
        base[i] = v[0];
        base[n + i] = v[1];
        base[2 * n + i] = v[2];
#endif
    }
    B2_D Real3 pos(u32 level) const
    {
        return load3(s.geo_pos, s.num_slots * s.max_depth, lidx(level));
    }
    B2_D Real3 dir(u32 level) const
    {
        return load3(s.geo_dir, s.num_slots * s.max_depth, lidx(level));
    }
    B2_D void set_pos(u32 level, Real3 const& v) const
    {
        store3(s.geo_pos, s.num_slots * s.max_depth, lidx(level), v);
    }
    B2_D void set_dir(u32 level, Real3 const& v) const
    {
        store3(s.geo_dir, s.num_slots * s.max_depth, lidx(level), v);
    }
    B2_D u32 vol(u32 level) const { return s.geo_vol[lidx(level)]; }
    B2_D u32 univ(u32 level) const { return s.geo_univ[lidx(level)]; }

    B2_D u32 level() const { return s.geo_level[slot]; }
    B2_D Real3 pos() const { return pos(0); }
    B2_D Real3 dir() const { return dir(0); }
    B2_D bool is_on_boundary() const { return s.geo_surface_level[slot] != INVALID; }
    B2_D bool is_outside() const { return vol(0) == 0; }

    B2_D u32 volume_id() const
    {
        u32 lev = level();
        return g.universe_volume_offset[univ(lev)] + vol(lev);
    }
    B2_D u32 surface_id() const
    {
        u32 sl = s.geo_surface_level[slot];
        if (sl == INVALID)
            return INVALID;
        return g.universe_surface_offset[univ(sl)] + s.geo_surf[slot];
    }

    B2_D void clear_next() const
    {
        s.geo_next_step[slot] = 0;
        s.geo_next_surf[slot] = INVALID;
    }
    B2_D void clear_surface() const { s.geo_surface_level[slot] = INVALID; }

    B2_D SimpleUnit const& unit_of(u32 universe) const
    {
        return g.simple_units[g.universe_index[universe]];
    }

    B2_D u32 daughter_of(u32 universe, u32 volume) const
    {
        return univ_daughter(g, universe, volume);
    }

    //! Locate a track from scratch (OrangeTrackView::operator=(Initializer))
    B2_D void initialize(Real3 const& ipos, Real3 const& idir)
    {
        failed = geo_initialize(g, s, slot, ipos, idir);
    }
    B2_D void initialize_impl(Real3 const& ipos, Real3 const& idir)
    {
        failed = false;
        Real3 lpos = ipos, ldir = idir;
        u32 uid = 0;
        u32 lev = 0;
        u32 daughter;
        do
        {
            Initialization tinit = univ_initialize(g, uid, lpos);
            if (tinit.volume == INVALID || tinit.surface != INVALID)
            {
                failed = true;
                tinit.volume = 0;
            }
            s.geo_vol[lidx(lev)] = tinit.volume;
            s.geo_univ[lidx(lev)] = uid;
            set_pos(lev, lpos);
            set_dir(lev, ldir);
            daughter = univ_daughter(g, uid, tinit.volume);
            if (daughter != INVALID)
            {
                transform_down(g, g.daughter_transform[daughter], lpos, ldir);
                uid = g.daughter_universe[daughter];
                ++lev;
            }
        } while (daughter != INVALID);
        s.geo_level[slot] = lev;
        s.geo_boundary[slot] = 1;
        clear_surface();
        clear_next();
    }

    //! Locate a track whose volume hierarchy is already known (it was born at its
    //! parent's position): identical state to initialize() without the BIH searches
    B2_D void initialize_known(Real3 const& ipos,
                               Real3 const& idir,
                               u32 lev_max,
                               u32 const* vols,
                               u32 const* univs,
                               u32 stride)
    {
        failed = false;
        Real3 lpos = ipos, ldir = idir;
        for (u32 lev = 0; lev <= lev_max; ++lev)
        {
            u32 const v = vols[lev * stride];
            u32 const uid = univs[lev * stride];
            s.geo_vol[lidx(lev)] = v;
            s.geo_univ[lidx(lev)] = uid;
            set_pos(lev, lpos);
            set_dir(lev, ldir);
            if (lev < lev_max)
            {
                u32 daughter = daughter_of(uid, v);
                transform_down(g, g.daughter_transform[daughter], lpos, ldir);
            }
        }
        s.geo_level[slot] = lev_max;
        s.geo_boundary[slot] = 1;
        clear_surface();
        clear_next();
    }

    //! Copy another slot's location with a new direction (DetailedInitializer)
    B2_D void initialize_from(u32 other, Real3 const& newdir)
    {
        failed = false;
        u32 lev = s.geo_level[other];
        if (other != slot)
        {
            s.geo_level[slot] = lev;
            s.geo_surface_level[slot] = s.geo_surface_level[other];
            s.geo_surf[slot] = s.geo_surf[other];
            s.geo_sense[slot] = s.geo_sense[other];
            s.geo_boundary[slot] = s.geo_boundary[other];
            u32 n = s.num_slots * s.max_depth;
            for (u32 l = 0; l <= lev; ++l)
            {
                u32 src = l * s.num_slots + other;
                u32 dst = lidx(l);
                store3(s.geo_pos, n, dst, load3(s.geo_pos, n, src));
                store3(s.geo_dir, n, dst, load3(s.geo_dir, n, src));
                s.geo_vol[dst] = s.geo_vol[src];
                s.geo_univ[dst] = s.geo_univ[src];
            }
        }
        clear_next();
        set_dir_all_levels(newdir);
    }

    B2_D void set_dir_all_levels(Real3 const& newdir)
    {
        Real3 ldir = newdir;
        u32 lev = level();
        for (u32 l = 0; l < lev; ++l)
        {
            set_dir(l, ldir);
            u32 daughter = daughter_of(univ(l), vol(l));
            Real3 dummy = make_real3(0, 0, 0);
            transform_down(g, g.daughter_transform[daughter], dummy, ldir);
        }
        set_dir(lev, ldir);
    }

    B2_D LocalState local_state(u32 lev) const
    {
        LocalState st;
        st.pos = pos(lev);
        st.dir = dir(lev);
        st.volume = vol(lev);
        if (lev == s.geo_surface_level[slot])
        {
            st.surface = s.geo_surf[slot];
            st.sense = s.geo_sense[slot];
        }
        else
        {
            st.surface = INVALID;
            st.sense = 0;
        }
        return st;
    }

    //! Distance to next boundary over all levels (find_next_step[_impl])
    B2_D Propagation find_next_step(bool limited, real max_step)
    {
        if constexpr (COOP)
            return coop_find_next_step(*this, limited, max_step);
        else
            return geo_find_next_step(g, s, slot, limited, max_step);
    }
    B2_D Propagation find_next_step_impl(bool limited, real max_step)
    {
        if (s.geo_boundary[slot] == 0)
        {
            // reentrant: already "at" the next boundary
            return Propagation{0, true, false};
        }
        Intersection isect = univ_intersect(g, 0, local_state(0), limited, max_step);
        u32 min_level = 0;
        u32 lev = level();
        for (u32 l = 1; l <= lev; ++l)
        {
            Intersection li
                = univ_intersect(g, univ(l), local_state(l), true, isect.distance);
            if (li.distance < isect.distance)
            {
                isect = li;
                min_level = l;
            }
        }
        s.geo_next_step[slot] = isect.distance;
        s.geo_next_surf[slot] = isect.surface;
        s.geo_next_sense[slot] = isect.sense;
        if (isect.surface != INVALID)
            s.geo_next_level[slot] = min_level;
        return Propagation{isect.distance, isect.surface != INVALID, false};
    }

    B2_D void move_to_boundary()
    {
        real dist = s.geo_next_step[slot];
        u32 lev = level();
        for (u32 l = 0; l <= lev; ++l)
        {
            Real3 p = pos(l);
            axpy(dist, dir(l), p);
            set_pos(l, p);
        }
        s.geo_surface_level[slot] = s.geo_next_level[slot];
        s.geo_surf[slot] = s.geo_next_surf[slot];
        s.geo_sense[slot] = s.geo_next_sense[slot];
        clear_next();
    }

    B2_D void move_internal(real dist)
    {
        u32 lev = level();
        for (u32 l = 0; l <= lev; ++l)
        {
            Real3 p = pos(l);
            axpy(dist, dir(l), p);
            set_pos(l, p);
        }
        s.geo_next_step[slot] = s.geo_next_step[slot] - dist;
        clear_surface();
    }

    B2_D void move_internal_pos(Real3 const& newpos)
    {
        geo_move_internal_pos(g, s, slot, newpos[0], newpos[1], newpos[2]);
    }
    B2_D void move_internal_pos_impl(Real3 const& newpos)
    {
        Real3 lpos = newpos;
        u32 lev = level();
        for (u32 l = 0; l < lev; ++l)
        {
            set_pos(l, lpos);
            u32 daughter = daughter_of(univ(l), vol(l));
            Real3 dummy = make_real3(0, 0, 0);
            // translate only: rotate position with a throwaway direction
            transform_down(g, g.daughter_transform[daughter], lpos, dummy);
        }
        set_pos(lev, lpos);
        clear_surface();
        clear_next();
    }

    B2_D void cross_boundary()
    {
        if (geo_cross_boundary(g, s, slot))
            failed = true;
    }
    B2_D void cross_boundary_impl()
    {
        if (s.geo_boundary[slot] == 0)
        {
            s.geo_boundary[slot] = 1;
            return;
        }
        s.geo_sense[slot] ^= 1;
        s.geo_boundary[slot] = 1;

        u32 lev = s.geo_surface_level[slot];
        u32 universe = univ(lev);
        LocalState local;
        local.pos = pos(lev);
        local.dir = dir(lev);
        local.volume = vol(lev);
        local.surface = s.geo_surf[slot];
        local.sense = s.geo_sense[slot];

        Initialization ci = univ_cross_boundary(g, universe, local);
        u32 volume = ci.volume;
        if (volume == INVALID)
        {
            failed = true;
            volume = 0;
        }
        s.geo_vol[lidx(lev)] = volume;
        u32 daughter = daughter_of(universe, volume);
        while (daughter != INVALID)
        {
            ++lev;
            transform_down(g, g.daughter_transform[daughter], local.pos, local.dir);
            universe = g.daughter_universe[daughter];
            Initialization ti = univ_initialize(g, universe, local.pos);
            volume = ti.volume;
            if (volume == INVALID)
            {
                failed = true;
                volume = 0;
            }
            daughter = daughter_of(universe, volume);
            s.geo_vol[lidx(lev)] = volume;
            s.geo_univ[lidx(lev)] = universe;
            set_pos(lev, local.pos);
            set_dir(lev, local.dir);
        }
        s.geo_level[slot] = lev;
    }

    B2_D void set_dir(Real3 const& newdir)
    {
        geo_set_dir(g, s, slot, newdir[0], newdir[1], newdir[2]);
    }
    B2_D void set_dir_impl(Real3 const& newdir)
    {
        if (is_on_boundary())
        {
            u32 sl = s.geo_surface_level[slot];
            Real3 normal = univ_normal(g, univ(sl), pos(sl), s.geo_surf[slot]);
            for (int l = int(level()) - 1; l >= 0; --l)
            {
                u32 daughter = daughter_of(univ(l), vol(l));
                normal = rotate_up(g, g.daughter_transform[daughter], normal);
            }
            if ((dot(normal, newdir) >= 0) != (dot(normal, dir()) >= 0))
                s.geo_boundary[slot] ^= 1;
        }
        set_dir_all_levels(newdir);
        clear_next();
    }

    B2_D real find_safety()
    {
        if constexpr (COOP)
            return coop_find_safety(*this);
        else
            return geo_find_safety(g, s, slot);
    }
    B2_D real find_safety_impl()
    {
        real min_safety = real_inf();
        u32 lev = level();
        for (u32 l = 0; l <= lev; ++l)
        {
            real sd = univ_safety(g, univ(l), pos(l), vol(l));
            min_safety = sd < min_safety ? sd : min_safety;
        }
        return min_safety;
    }
};
using GeoTrack = GeoTrackT<false>;

//---------------------------------------------------------------------------//
// WARP-COOPERATIVE DISTANCE AND SAFETY SEARCH
//
// The reference walks the universe levels of a track one after the other and, inside each
// level, the faces of the current volume one after the other
// (OrangeTrackView::find_next_step_impl, OrangeTrackView.hh; SimpleUnitTracker::
// intersect_impl / simple_intersect, univ/SimpleUnitTracker.hh:390-507): for a lone track
// that is (levels x faces) dependent chains of four loads and a quadratic solve. Here the
// (level, face) pairs of ALL levels are dealt out to the lanes of the warp that holds the
// track, every lane solves one face, and two warp-wide integer minima pick the winner.
//
// Same result as the serial search: at level 0 a root is valid if d <= max_step (limited)
// or d < real_max; at deeper levels the reference passes the shallower levels' best distance
// as the limit and takes a deeper intersection only if it is STRICTLY closer, so the result
// is the lexicographic minimum of (distance, level, face) over the roots that are strictly
// below that limit; "first of equal" inside a level is the lowest face, across levels the
// shallowest level. Lanes are numbered in (level, face) order, so the tie-break is the lowest
// lane. Levels that are rect arrays or volumes with internal surfaces / implicit
// (background) volumes are ONE item: the lane runs the serial per-universe search for it.
// More than 32 items: every lane runs the serial search (redundantly).
//---------------------------------------------------------------------------//
B2_D u32 coop_lane()
{
    return threadIdx.x & 31u;
}

//! Lane holding the smallest non-negative double among the valid lanes (ties: lowest lane);
//! 32 if no lane is valid. Two integer warp minima (REDUX) instead of a shuffle tree.
B2_D u32 coop_argmin(real d, bool valid)
{
    constexpr unsigned full = 0xffffffffu;
    u64 const bits = static_cast<u64>(__double_as_longlong(d));
    u32 const hi = valid ? static_cast<u32>(bits >> 32) : 0xffffffffu;
    u32 const min_hi = __reduce_min_sync(full, hi);
    bool const in_hi = valid && hi == min_hi;
    u32 const lo = in_hi ? static_cast<u32>(bits) : 0xffffffffu;
    u32 const min_lo = __reduce_min_sync(full, lo);
    unsigned const winners = __ballot_sync(full, in_hi && lo == min_lo);
    return winners ? static_cast<u32>(__ffs(winners) - 1) : 32u;
}

B2_D real coop_broadcast(real v, u32 lane)
{
    return __shfl_sync(0xffffffffu, v, lane);
}

template<bool COOP>
B2_D Propagation coop_find_next_step(GeoTrackT<COOP>& t, bool limited, real max_step)
{
    GeoParams const& g = t.g;
    StateView const& s = t.s;
    u32 const slot = t.slot;
    if (s.geo_boundary[slot] == 0)
        return Propagation{0, true, false};  // reentrant: already "at" the next boundary

    u32 const lane = coop_lane();
    u32 const lev = t.level();
    // deal the (level, face) items out to the lanes
    u32 my_level = INVALID, my_item = 0;
    bool my_whole = false;
    u32 total = 0;
    for (u32 l = 0; l <= lev; ++l)
    {
        u32 const uid = t.univ(l);
        u32 count = 1;
        bool whole = true;
        if (g.universe_type[uid] == UNIV_SIMPLE)
        {
            SimpleUnit const& u = g.simple_units[g.universe_index[uid]];
            u32 const rec = u.vol_begin + t.vol(l);
            if (!(g.vol_flags[rec] & (VOL_INTERNAL_SURFACES | VOL_IMPLICIT)))
            {
                count = g.vol_face_end[rec] - g.vol_face_begin[rec];
                whole = false;
            }
        }
        if (lane >= total && lane < total + count)
        {
            my_level = l;
            my_item = lane - total;
            my_whole = whole;
        }
        total += count;
    }
    if (total > 32)
        return t.find_next_step_impl(limited, max_step);

    // one root search per lane
    real const bound = limited ? max_step : real_inf();
    real d = real_inf();
    bool valid = false;
    u32 w_surface = INVALID;
    u32 w_sense = 0;
    if (my_level != INVALID)
    {
        LocalState const st = t.local_state(my_level);
        u32 const uid = t.univ(my_level);
        if (my_whole)
        {
            Intersection const is = (my_level == 0)
                                        ? univ_intersect(g, uid, st, limited, max_step)
                                        : univ_intersect(g, uid, st, false, real_inf());
            d = is.distance;
            w_surface = is.surface;
            w_sense = is.sense;
            valid = is.surface != INVALID && (my_level == 0 || d < bound);
        }
        else
        {
            SimpleUnit const& u = g.simple_units[g.universe_index[uid]];
            VolumeRef const vol = get_volume(g, u, st.volume);
            u32 const on_face
                = (st.surface != INVALID) ? volume_find_face(g, vol, st.surface) : INVALID;
            SurfaceRef const sr = get_surface(g, u, volume_surface(g, vol, my_item));
            bool const on = (my_item == on_face);
            int const nroots = surface_num_isect(sr.type);
            if (!(nroots == 1 && on))
            {
                Roots const r = surface_intersect(sr, st.pos, st.dir, on);
                for (int k = 0; k < nroots; ++k)
                {
                    real const dk = r.r[k];
                    bool const ok = (my_level == 0)
                                        ? (limited ? (dk <= max_step) : (dk < real_max()))
                                        : (dk < bound);
                    if (ok && (!valid || dk < d))
                    {
                        d = dk;
                        valid = true;
                    }
                }
            }
        }
    }
    u32 const winner = coop_argmin(d, valid);

    Intersection isect{INVALID, 0, bound};
    u32 min_level = 0;
    if (winner < 32)
    {
        constexpr unsigned full = 0xffffffffu;
        isect.distance = coop_broadcast(d, winner);
        min_level = __shfl_sync(full, my_level, winner);
        u32 const item = __shfl_sync(full, my_item, winner);
        bool const whole = __shfl_sync(full, static_cast<int>(my_whole), winner) != 0;
        u32 const ws = __shfl_sync(full, w_surface, winner);
        u32 const wn = __shfl_sync(full, w_sense, winner);
        if (whole)
        {
            isect.surface = ws;
            isect.sense = static_cast<u8>(wn);
        }
        else
        {
            // surface and current sense of the winning face (every lane, same values)
            LocalState const st = t.local_state(min_level);
            SimpleUnit const& u = g.simple_units[g.universe_index[t.univ(min_level)]];
            VolumeRef const vol = get_volume(g, u, st.volume);
            u32 const surface = volume_surface(g, vol, item);
            isect.surface = surface;
            isect.sense = (surface == st.surface)
                              ? st.sense
                              : static_cast<u8>(surface_sense(get_surface(g, u, surface), st.pos)
                                                >= 0);
        }
    }
    s.geo_next_step[slot] = isect.distance;
    s.geo_next_surf[slot] = isect.surface;
    s.geo_next_sense[slot] = isect.sense;
    if (isect.surface != INVALID)
        s.geo_next_level[slot] = min_level;
    return Propagation{isect.distance, isect.surface != INVALID, false};
}

//! Safety distance over all levels: min over levels of the per-universe safety
//! (OrangeTrackView::find_safety; SimpleUnitTracker::safety, univ/SimpleUnitTracker.hh)
template<bool COOP>
B2_D real coop_find_safety(GeoTrackT<COOP>& t)
{
    GeoParams const& g = t.g;
    u32 const lane = coop_lane();
    u32 const lev = t.level();
    u32 my_level = INVALID, my_item = 0;
    bool my_whole = false;
    u32 total = 0;
    for (u32 l = 0; l <= lev; ++l)
    {
        u32 const uid = t.univ(l);
        u32 count = 1;
        bool whole = true;
        if (g.universe_type[uid] == UNIV_SIMPLE)
        {
            SimpleUnit const& u = g.simple_units[g.universe_index[uid]];
            u32 const rec = u.vol_begin + t.vol(l);
            u32 const nf = g.vol_face_end[rec] - g.vol_face_begin[rec];
            if ((g.vol_flags[rec] & VOL_SIMPLE_SAFETY) && nf > 0)
            {
                count = nf;
                whole = false;
            }
        }
        if (lane >= total && lane < total + count)
        {
            my_level = l;
            my_item = lane - total;
            my_whole = whole;
        }
        total += count;
    }
    if (total > 32)
        return t.find_safety_impl();
    real d = real_inf();
    if (my_level != INVALID)
    {
        u32 const uid = t.univ(my_level);
        Real3 const pos = t.pos(my_level);
        if (my_whole)
        {
            d = univ_safety(g, uid, pos, t.vol(my_level));
        }
        else
        {
            SimpleUnit const& u = g.simple_units[g.universe_index[uid]];
            VolumeRef const vol = get_volume(g, u, t.vol(my_level));
            d = surface_safety(get_surface(g, u, volume_surface(g, vol, my_item)), pos);
        }
    }
    // distances are >= 0 or +inf (never NaN: surface_safety maps a NaN normal to +inf)
    u32 const winner = coop_argmin(d, my_level != INVALID);
    return winner < 32 ? coop_broadcast(d, winner) : real_inf();
}

B2_GEO_FN Propagation
geo_find_next_step(GeoParams const& g, StateView const& s, u32 slot, bool limited, real max_step)
{
    return GeoTrack(g, s, slot).find_next_step_impl(limited, max_step);
}

B2_GEO_FN real geo_find_safety(GeoParams const& g, StateView const& s, u32 slot)
{
    return GeoTrack(g, s, slot).find_safety_impl();
}

B2_GEO_FN void
geo_set_dir(GeoParams const& g, StateView const& s, u32 slot, real dx, real dy, real dz)
{
    GeoTrack(g, s, slot).set_dir_impl(make_real3(dx, dy, dz));
}

B2_GEO_FN bool geo_cross_boundary(GeoParams const& g, StateView const& s, u32 slot)
{
    GeoTrack t(g, s, slot);
    t.cross_boundary_impl();
    return t.failed;
}

B2_GEO_FN bool geo_initialize(GeoParams const& g,
                                       StateView const& s,
                                       u32 slot,
                                       Real3 const& pos,
                                       Real3 const& dir)
{
    GeoTrack t(g, s, slot);
    t.initialize_impl(pos, dir);
    return t.failed;
}

B2_GEO_FN void
geo_move_internal_pos(GeoParams const& g, StateView const& s, u32 slot, real x, real y, real z)
{
    GeoTrack(g, s, slot).move_internal_pos_impl(make_real3(x, y, z));
}
}  // namespace b200
