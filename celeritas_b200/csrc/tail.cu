//---------------------------------------------------------------------------//
// C-ABI launchers of the device-resident step loop (kernel: tail_loop.cuh).
//---------------------------------------------------------------------------//
#include <algorithm>
#include <cstdlib>

#include "launch_util.cuh"
#include "tail_args.cuh"

using namespace b200;

extern "C" {
//! Largest cooperative grid of the tail kernel that is resident at once on this device
int b200_tail_max_blocks(B200ParamsView const* params, int* out)
{
    int device = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e == cudaSuccess)
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess)
    {
        FieldParams const& f = PV(params).model.field;
        e = !f.enabled ? tail_blocks_per_sm_nofield(&per_sm)
            : f.rz_values ? tail_blocks_per_sm_rzfield(&per_sm)
                          : tail_blocks_per_sm_field(&per_sm);
    }
    if (e != cudaSuccess)
        return static_cast<int>(e);
    *out = sms * per_sm;
    return 0;
}

int b200_step_tail_loop(B200ParamsView const* params,
                        B200StateView const* state,
                        uint32_t num_blocks,
                        uint32_t max_iterations,
                        uint32_t exit_active,
                        uint32_t* ring,
                        uint32_t* done,
                        cudaStream_t stream)
{
    StateView const& s = SV(state);
    if (!s.run_vac_prefix || !s.run_vac_mask || !s.run_scan || !s.tail_reset_list || !s.tail_ctrl || !ring || !done
        || num_blocks == 0 || max_iterations == 0 || s.num_slots % 32u != 0
        || PV(params).model.has_extra_models)
        return B200_ERR_INVALID_ARGUMENT;
    static u32 const coop = [] {
        char const* env = std::getenv("B200_TAIL_COOP");
        return env ? static_cast<u32>(std::atoi(env)) : 1u;
    }();
    static u32 const coop_max_env = [] {
        char const* env = std::getenv("B200_TAIL_COOP_MAX");
        return env ? static_cast<u32>(std::atoi(env)) : 16u;
    }();
    // never more tracks than warps in the grid
    u32 const coop_max = std::min<u32>(coop_max_env, num_blocks * (BLOCK / 32));
    TailArgs args{max_iterations, exit_active, ring, done, coop, coop_max};
    ParamsView const& p = PV(params);
    FieldParams const& f = p.model.field;
    cudaError_t const e = !f.enabled    ? tail_launch_nofield(p, s, args, num_blocks, stream)
                          : f.rz_values ? tail_launch_rzfield(p, s, args, num_blocks, stream)
                                        : tail_launch_field(p, s, args, num_blocks, stream);
    B2_COUNT(1);
    return e == cudaSuccess ? check_launch() : static_cast<int>(e);
}
}  // extern "C"
