//---------------------------------------------------------------------------//
// Whole-step fused kernel for small iterations and its launcher.
// (own translation unit: these are the largest kernels of the library and compile in
// parallel with kernels.cu; device code in step_device.cuh)
//---------------------------------------------------------------------------//
#include "launch_util.cuh"
#include "step_device.cuh"

namespace b200
{
template<bool FIELD>
__global__ void __launch_bounds__(BLOCK, B2_FUSED_MIN_BLOCKS)
    k_step_fused(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    step_fused_track<FIELD>(p, s, thread_id());
}

}  // namespace b200

using namespace b200;

extern "C" {
int b200_step_fused(B200ParamsView const* params, B200StateView const* state, cudaStream_t stream)
{
    StateView const& s = SV(state);
    unsigned const grid = grid_for(active_hint(s));
    if (PV(params).model.field.enabled)
        k_step_fused<true><<<grid, BLOCK, 0, stream>>>(PV(params), s);
    else
        k_step_fused<false><<<grid, BLOCK, 0, stream>>>(PV(params), s);
    B2_COUNT(1);
    return check_launch();
}

}  // extern "C"
