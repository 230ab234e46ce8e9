//---------------------------------------------------------------------------//
// Host wrapper of the ray-trace kernel: host buffers in, host buffers out.
//---------------------------------------------------------------------------//
#include <memory>
#include <string>
#include <vector>

#include "../../include/celeritas_b200.h"
#include "CoreParams.hh"
#include "CoreState.hh"
#include "DeviceMemory.hh"
#include "Handles.hh"

using namespace celeritas_b200;


extern "C" int b200_geo_trace_host(B200Params const* params,
                                   double const* pos,
                                   double const* dir,
                                   uint32_t num_rays,
                                   uint32_t max_segments,
                                   uint32_t* volume,
                                   uint32_t* surface,
                                   double* distance,
                                   uint32_t* count,
                                   double* safety)
{
    if (!params || !pos || !dir || num_rays == 0 || max_segments == 0)
        return B200_ERR_INVALID_ARGUMENT;
    try
    {
        CoreState state(params->params, 0, num_rays);
        DeviceArena arena;
        std::vector<double> hp(pos, pos + 3 * size_t(num_rays)), hd(dir, dir + 3 * size_t(num_rays));
        double const* d_pos = arena.upload(hp);
        double const* d_dir = arena.upload(hd);
        size_t const total = size_t(num_rays) * max_segments;
        uint32_t* d_vol = arena.alloc_fill<uint32_t>(total, 0xff);
        uint32_t* d_surf = arena.alloc_fill<uint32_t>(total, 0xff);
        double* d_dist = arena.alloc<double>(total);
        uint32_t* d_count = arena.alloc<uint32_t>(num_rays);
        double* d_safety = arena.alloc<double>(num_rays);
        int rc = b200_geo_trace(
            reinterpret_cast<B200ParamsView const*>(&params->params->view()),
            reinterpret_cast<B200StateView const*>(&state.view()),
            d_pos,
            d_dir,
            num_rays,
            max_segments,
            d_vol,
            d_surf,
            d_dist,
            d_count,
            d_safety,
            state.stream());
        if (rc != 0)
            return rc;
        B2_CUDA_CALL(cudaStreamSynchronize(state.stream()));
        B2_CUDA_CALL(cudaMemcpy(volume, d_vol, total * 4, cudaMemcpyDeviceToHost));
        B2_CUDA_CALL(cudaMemcpy(surface, d_surf, total * 4, cudaMemcpyDeviceToHost));
        B2_CUDA_CALL(cudaMemcpy(distance, d_dist, total * 8, cudaMemcpyDeviceToHost));
        B2_CUDA_CALL(cudaMemcpy(count, d_count, num_rays * 4, cudaMemcpyDeviceToHost));
        B2_CUDA_CALL(cudaMemcpy(safety, d_safety, num_rays * 8, cudaMemcpyDeviceToHost));
        return B200_OK;
    }
    catch (CudaError const& e)
    {
        return e.code;
    }
    catch (std::exception const&)
    {
        return B200_ERR_RUNTIME;
    }
}
