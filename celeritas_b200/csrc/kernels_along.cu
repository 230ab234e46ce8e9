//---------------------------------------------------------------------------//
// Along-step kernels (one launch per charge class, or four phase kernels) and launchers.
// (own translation unit: these are the largest kernels of the library and compile in
// parallel with kernels.cu; device code in step_device.cuh)
//---------------------------------------------------------------------------//
#include "launch_util.cuh"
#include "step_device.cuh"

namespace b200
{
template<int FIELD, bool SELECT>
__global__ void __launch_bounds__(B2_ALONG_BLOCK, (FIELD ? B2_ALONG_FIELD_MIN_BLOCKS : ALONG_MIN_BLOCKS) * BLOCK / B2_ALONG_BLOCK) k_along_step_charged(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 tid = thread_id();
    u32 slot = INVALID;
    if (tid < s.counters[CTR_NUM_CHARGED])
        slot = s.track_slots[tid];
    if (slot != INVALID)
    {
        prefetch_along_step_state<true>(s, slot);
        if (s.status[slot] == ST_ALIVE)
            along_step<true, FIELD>(p, s, slot);
    }
    if (SELECT)
        select_and_append(p, s, slot);
}

#define B2_ALONG_PHASE_KERNEL(NAME, PHASE, MIN_BLOCKS)                                   \
    __global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)                                 \
        NAME(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)    \
    {                                                                                    \
        u32 tid = thread_id();                                                           \
        if (tid >= s.counters[CTR_NUM_CHARGED])                                          \
            return;                                                                      \
        u32 slot = s.track_slots[tid];                                                   \
        if (s.status[slot] != ST_ALIVE)                                                  \
            return;                                                                      \
        PHASE(p, s, slot);                                                               \
    }
B2_ALONG_PHASE_KERNEL(k_along_msc_limit, along_phase_msc_limit, B2_PHASE_MIN_BLOCKS)
B2_ALONG_PHASE_KERNEL(k_along_propagate_linear, along_phase_propagate<0>, B2_PHASE_MIN_BLOCKS)
B2_ALONG_PHASE_KERNEL(k_along_propagate_field, along_phase_propagate<1>, B2_PROPAGATE_FIELD_MIN_BLOCKS)
B2_ALONG_PHASE_KERNEL(k_along_propagate_rzfield, along_phase_propagate<2>, B2_PROPAGATE_FIELD_MIN_BLOCKS)
B2_ALONG_PHASE_KERNEL(k_along_msc_apply, along_phase_msc_apply, B2_PHASE_MIN_BLOCKS)
B2_ALONG_PHASE_KERNEL(k_along_finish, along_phase_finish, B2_PHASE_MIN_BLOCKS)
#undef B2_ALONG_PHASE_KERNEL

template<bool SELECT>
__global__ void __launch_bounds__(B2_ALONG_BLOCK, B2_NEUTRAL_MIN_BLOCKS * BLOCK / B2_ALONG_BLOCK) k_along_step_neutral(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 tid = thread_id();
    u32 slot = INVALID;
    if (tid < s.counters[CTR_NUM_NEUTRAL])
        slot = s.track_slots[s.num_slots - 1 - tid];
    if (slot != INVALID)
    {
        prefetch_along_step_state<false>(s, slot);
        if (s.status[slot] == ST_ALIVE)
            along_step<false, 0>(p, s, slot);
    }
    if (SELECT)
        select_and_append(p, s, slot);
}

}  // namespace b200

using namespace b200;

extern "C" {
int b200_step_along_step(B200ParamsView const* params, B200StateView const* state, cudaStream_t stream)
{
    StateView const& s = SV(state);
    u32 nc = s.hint_charged < s.num_slots ? s.hint_charged : s.num_slots;
    u32 nn = s.hint_neutral < s.num_slots ? s.hint_neutral : s.num_slots;
    bool const split = PV(params).model.field.enabled
                           ? (B2_ALONG_SPLIT_FIELD_THRESHOLD != 0
                              && nc >= B2_ALONG_SPLIT_FIELD_THRESHOLD)
                           : (B2_ALONG_SPLIT_THRESHOLD != 0 && nc >= B2_ALONG_SPLIT_THRESHOLD);
    if (split)
    {
        ParamsView const& p = PV(params);
        unsigned const grid = grid_for(nc);
        if (p.model.msc.enabled)
            k_along_msc_limit<<<grid, BLOCK, 0, stream>>>(p, s);
        if (p.model.field.enabled && p.model.field.rz_values)
            k_along_propagate_rzfield<<<grid, BLOCK, 0, stream>>>(p, s);
        else if (p.model.field.enabled)
            k_along_propagate_field<<<grid, BLOCK, 0, stream>>>(p, s);
        else
            k_along_propagate_linear<<<grid, BLOCK, 0, stream>>>(p, s);
        if (p.model.msc.enabled)
            k_along_msc_apply<<<grid, BLOCK, 0, stream>>>(p, s);
        k_along_finish<<<grid, BLOCK, 0, stream>>>(p, s);
        B2_COUNT(p.model.msc.enabled ? 4 : 2);
    }
    else if (nc > 0)
    {
        if (PV(params).model.field.enabled && PV(params).model.field.rz_values)
            k_along_step_charged<2, false><<<(nc + B2_ALONG_BLOCK - 1) / B2_ALONG_BLOCK, B2_ALONG_BLOCK, 0, stream>>>(PV(params), s);
        else if (PV(params).model.field.enabled)
            k_along_step_charged<1, false><<<(nc + B2_ALONG_BLOCK - 1) / B2_ALONG_BLOCK, B2_ALONG_BLOCK, 0, stream>>>(PV(params), s);
        else
            k_along_step_charged<0, false><<<(nc + B2_ALONG_BLOCK - 1) / B2_ALONG_BLOCK, B2_ALONG_BLOCK, 0, stream>>>(PV(params), s);
        B2_COUNT(1);
    }
    if (nn > 0)
    {
        k_along_step_neutral<false><<<(nn + B2_ALONG_BLOCK - 1) / B2_ALONG_BLOCK, B2_ALONG_BLOCK, 0, stream>>>(PV(params), s);
        B2_COUNT(1);
    }
    return check_launch();
}

int b200_step_along_select(B200ParamsView const* params, B200StateView const* state, cudaStream_t stream)
{
    StateView const& s = SV(state);
    if (!s.interact_list)
        return B200_ERR_INVALID_ARGUMENT;
    u32 nc = s.hint_charged < s.num_slots ? s.hint_charged : s.num_slots;
    u32 nn = s.hint_neutral < s.num_slots ? s.hint_neutral : s.num_slots;
    if (nc > 0)
    {
        if (PV(params).model.field.enabled && PV(params).model.field.rz_values)
            k_along_step_charged<2, true><<<grid_for(nc), BLOCK, 0, stream>>>(PV(params), s);
        else if (PV(params).model.field.enabled)
            k_along_step_charged<1, true><<<grid_for(nc), BLOCK, 0, stream>>>(PV(params), s);
        else
            k_along_step_charged<0, true><<<grid_for(nc), BLOCK, 0, stream>>>(PV(params), s);
        B2_COUNT(1);
    }
    if (nn > 0)
    {
        k_along_step_neutral<true><<<grid_for(nn), BLOCK, 0, stream>>>(PV(params), s);
        B2_COUNT(1);
    }
    return check_launch();
}

}  // extern "C"
