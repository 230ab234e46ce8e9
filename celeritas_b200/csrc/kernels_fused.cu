//---------------------------------------------------------------------------//
// Whole-step fused kernel for small iterations and its launcher.
// (own translation unit: these are the largest kernels of the library and compile in
// parallel with kernels.cu; device code in step_device.cuh)
//---------------------------------------------------------------------------//
#include <algorithm>
#include <cstdlib>

#include "launch_util.cuh"
#include "step_device.cuh"

namespace b200
{
// Tracks are dealt out to the warps of the launch ROUND-ROBIN: with `num_warps` warps, lane l
// of warp w works on active track l * num_warps + w. A fused whole step is ~15 k dependent
// instructions per track and tracks of one warp take different branches (one scatters, one
// radiates, one crosses a boundary ...), so a warp's time is roughly the SUM over its tracks.
// With few tracks it therefore pays to give every track its own warp (measured inside the
// device-resident loop, profiles/README_r02.md: 16..128 tracks 85 -> 46 us per iteration
// on TestEm3, 366 -> 257 us on the CMS-scale stand-in); with many, num_warps = tracks / 32 and
// the mapping is the dense one.
template<int FIELD>
__global__ void __launch_bounds__(BLOCK, B2_FUSED_MIN_BLOCKS)
    k_step_fused(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s, u32 num_warps)
{
    u32 const gtid = thread_id();
    u32 const warp = gtid >> 5;
    if (warp < num_warps)
        step_fused_track<FIELD>(p, s, (gtid & 31u) * num_warps + warp);
}

}  // namespace b200

using namespace b200;

extern "C" {
int b200_step_fused(B200ParamsView const* params, B200StateView const* state, cudaStream_t stream)
{
    StateView const& s = SV(state);
    // built without the interactors beyond the core list (run_interaction<false>)
    if (PV(params).model.has_extra_models)
        return B200_ERR_INVALID_ARGUMENT;
    // Warps to spread the tracks over: one track per warp up to B200_SPREAD_WARPS warps
    // (default 8 per SM), never fewer than tracks / 32
    static u32 const spread_warps = [] {
        if (char const* env = std::getenv("B200_SPREAD_WARPS"))
            return static_cast<u32>(std::strtoul(env, nullptr, 10));
        int device = 0, sms = 148;
        if (cudaGetDevice(&device) == cudaSuccess)
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        return static_cast<u32>(8 * sms);
    }();
    // Above ~4 tracks per warp the spread mapping loses to the dense one: its lanes gather
    // from scattered slots (measured, TestEm3 iterations of 4096..16384 tracks: 143 -> 156 us)
    static u32 const spread_max_tracks = [] {
        char const* env = std::getenv("B200_SPREAD_MAX_TRACKS");
        return env ? static_cast<u32>(std::strtoul(env, nullptr, 10)) : 4096u;
    }();
    u32 const n = active_hint(s);
    u32 const dense = (n + 31u) / 32u;
    u32 const num_warps = std::max<u32>(
        n <= spread_max_tracks ? std::max<u32>(dense, std::min<u32>(n, spread_warps)) : dense, 1u);
    unsigned const grid = grid_for(num_warps * 32u);
    if (PV(params).model.field.enabled && PV(params).model.field.rz_values)
        k_step_fused<2><<<grid, BLOCK, 0, stream>>>(PV(params), s, num_warps);
    else if (PV(params).model.field.enabled)
        k_step_fused<1><<<grid, BLOCK, 0, stream>>>(PV(params), s, num_warps);
    else
        k_step_fused<0><<<grid, BLOCK, 0, stream>>>(PV(params), s, num_warps);
    B2_COUNT(1);
    return check_launch();
}

}  // extern "C"
