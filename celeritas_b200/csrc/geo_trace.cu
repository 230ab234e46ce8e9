//---------------------------------------------------------------------------//
// Ray tracing through the ORANGE geometry on fixed ray sets.
//
// For every ray: locate, then repeat {find_next_step, move_to_boundary,
// cross_boundary} recording the volume id, the surface id crossed and the
// distance of each segment until the ray leaves the geometry. This is how the
// reference pins its navigation (track() helpers in test/orange/OrangeJson.test.cc
// and the celer-geo ray tracer); the ids must match the reference bit for bit.
// Also exposes safety distances at the ray origins.
//---------------------------------------------------------------------------//
#include "../../include/celeritas_b200.h"
#include "orange.cuh"
#include "views.cuh"

namespace b200
{
__global__ void k_geo_trace(B2_GRID_CONSTANT ParamsView const p,
                            B2_GRID_CONSTANT StateView const s,
                            real const* __restrict__ pos,
                            real const* __restrict__ dir,
                            u32 num_rays,
                            u32 max_segments,
                            u32* __restrict__ out_volume,
                            u32* __restrict__ out_surface,
                            real* __restrict__ out_distance,
                            u32* __restrict__ out_count,
                            real* __restrict__ out_safety)
{
    u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= num_rays)
        return;
    u32 const slot = tid;
    GeoTrack geo(p, s, slot);
    geo.initialize(make_real3(pos[3 * tid], pos[3 * tid + 1], pos[3 * tid + 2]),
                   make_real3(dir[3 * tid], dir[3 * tid + 1], dir[3 * tid + 2]));
    u32 n = 0;
    if (geo.failed)
    {
        out_count[tid] = INVALID;
        out_safety[tid] = -1;
        return;
    }
    out_safety[tid] = geo.is_outside() ? real(-1) : geo.find_safety();
    while (!geo.is_outside() && n < max_segments)
    {
        u32 const vol = geo.volume_id();
        Propagation prop = geo.find_next_step(false, 0);
        if (!prop.boundary)
        {
            out_volume[tid * max_segments + n] = vol;
            out_surface[tid * max_segments + n] = INVALID;
            out_distance[tid * max_segments + n] = prop.distance;
            ++n;
            break;
        }
        geo.move_to_boundary();
        out_volume[tid * max_segments + n] = vol;
        out_surface[tid * max_segments + n] = geo.surface_id();
        out_distance[tid * max_segments + n] = prop.distance;
        ++n;
        geo.cross_boundary();
        if (geo.failed)
        {
            n |= 0x80000000u;
            break;
        }
    }
    out_count[tid] = n;
}
}  // namespace b200

using namespace b200;

extern "C" int b200_geo_trace(B200ParamsView const* params,
                              B200StateView const* state,
                              double const* d_pos,
                              double const* d_dir,
                              uint32_t num_rays,
                              uint32_t max_segments,
                              uint32_t* d_volume,
                              uint32_t* d_surface,
                              double* d_distance,
                              uint32_t* d_count,
                              double* d_safety,
                              cudaStream_t stream)
{
    ParamsView const& p = *reinterpret_cast<ParamsView const*>(params);
    StateView const& s = *reinterpret_cast<StateView const*>(state);
    if (num_rays > s.num_slots)
        return B200_ERR_INVALID_ARGUMENT;
    k_geo_trace<<<(num_rays + 127) / 128, 128, 0, stream>>>(p,
                                                             s,
                                                             d_pos,
                                                             d_dir,
                                                             num_rays,
                                                             max_segments,
                                                             d_volume,
                                                             d_surface,
                                                             d_distance,
                                                             d_count,
                                                             d_safety);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : static_cast<int>(e);
}
