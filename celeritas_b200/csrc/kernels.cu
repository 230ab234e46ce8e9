//---------------------------------------------------------------------------//
// Step-action kernels and their C-ABI launchers.
//
// One kernel per step action of the reference's loop
// (/root/reference/src/celeritas/global/ActionSequence.cc:77-138).
//
// B200-specific structure:
//  * All counters (CoreStateCounters) are resident in device memory: kernels size
//    themselves from them, the host only reads them back once per iteration.
//  * Per-step kernels do not run over all track slots. The end-of-step pass builds
//    DENSE lists of active slots, charged tracks from the front of `track_slots`,
//    neutral tracks from the back; thread i of a kernel works on the i-th active
//    slot. Warps are therefore fully populated and charge-coherent, which is what
//    the reference's TrackOrder::init_charge / SortTracksAction aim for
//    (/root/reference/src/celeritas/track/SortTracksAction.cc:46-131) without a
//    radix sort: the lists fall out of the block scans that the vacancy compaction
//    needs anyway.
//  * Results are per-slot deterministic: thread->slot mapping never changes what a
//    slot computes (RNG state, physics and geometry are all per slot).
//---------------------------------------------------------------------------//
#include <atomic>
#include <cstdio>

#include "../../include/celeritas_b200.h"
#include "along_step.cuh"
#include "interact.cuh"
#include "orange.cuh"
#include "physics.cuh"
#include "rng.cuh"
#include "views.cuh"

namespace b200
{
constexpr int BLOCK = 128;
#ifndef B2_ALONG_MIN_BLOCKS
#    define B2_ALONG_MIN_BLOCKS 8
#endif
// Resident blocks per SM asked of the fused whole-step kernel (small iterations)
#ifndef B2_FUSED_MIN_BLOCKS
#    define B2_FUSED_MIN_BLOCKS 2
#endif
// Resident blocks per SM asked of the phase kernels of the split charged along-step
#ifndef B2_PHASE_MIN_BLOCKS
#    define B2_PHASE_MIN_BLOCKS 6
#endif
// Charged tracks from which the along-step runs as four phase kernels (0 = never)
// Measured (profiles/README_r01.md): the split is NOT faster (56.2 vs 54.3 ms per pass in
// the along-step), so it is off; the phase kernels stay for profiling single phases.
#ifndef B2_ALONG_SPLIT_THRESHOLD
#    define B2_ALONG_SPLIT_THRESHOLD 0
#endif
// With a magnetic field the split IS faster (CMS-scale stand-in, saturated iterations:
// 3.84 -> 3.06 ns per track-step; profiles/README_r01.md): the propagation phase is a substep
// loop over Dormand-Prince trials and boundary searches through several universe levels,
// and as one kernel with MSC and energy loss the charged along-step (14 k instructions) spends
// 44 stall cycles per issued instruction waiting for instruction fetch (ncu). Charged
// tracks from which the along-step of a FIELD problem runs as phase kernels (0 = never):
#ifndef B2_ALONG_SPLIT_FIELD_THRESHOLD
#    define B2_ALONG_SPLIT_FIELD_THRESHOLD 1
#endif
// Resident blocks per SM asked of the field-propagation phase kernel. Measured at
// saturation (gpurun_out/variants_cms2.log): 8 / 6 / 4 / 3 blocks (64 / 80 / 128 / 158
// registers; 3.4 kB / 2.1 kB / 0.3 kB / 0 of spill loads) = 2.52 / 2.46 / 2.37 / 2.37 ns per
// track-step
#ifndef B2_PROPAGATE_FIELD_MIN_BLOCKS
#    define B2_PROPAGATE_FIELD_MIN_BLOCKS 4
#endif
constexpr int ALONG_MIN_BLOCKS = B2_ALONG_MIN_BLOCKS;
// Threads per block of the along-step kernels (the same register budget per SM: the
// resident-block request scales with BLOCK / B2_ALONG_BLOCK)
#ifndef B2_ALONG_BLOCK
#    define B2_ALONG_BLOCK 128
#endif
// The charged along-step WITH the field propagator (Dormand-Prince driver) needs more
// registers than the field-free one
#ifndef B2_ALONG_FIELD_MIN_BLOCKS
#    define B2_ALONG_FIELD_MIN_BLOCKS 8
#endif
// Resident blocks per SM asked of the other large-iteration kernels. Measured, ms per pass
// (profiles/README_r01.md): pre-step 15.3 -> 12.9, neutral along-step -1.9, end passes
// 20.8 -> 16.7 when capped at 64 registers (8 blocks of 128 threads)
#ifndef B2_PRE_MIN_BLOCKS
#    define B2_PRE_MIN_BLOCKS 8
#endif
#ifndef B2_NEUTRAL_MIN_BLOCKS
#    define B2_NEUTRAL_MIN_BLOCKS 8
#endif
#ifndef B2_INTERACT_MIN_BLOCKS
#    define B2_INTERACT_MIN_BLOCKS 8
#endif
#ifndef B2_TAIL_MIN_BLOCKS
#    define B2_TAIL_MIN_BLOCKS 8
#endif
#ifndef B2_END_MIN_BLOCKS
#    define B2_END_MIN_BLOCKS 8
#endif
// Software prefetch of the per-slot state at kernel entry (see prefetch_l2): 1 = to L2,
// 2 = to L1. Measured: no effect either way (99.9 / 100.0 / 99.8 ms per pass for 0 / 1 / 2,
// profiles/README_r01.md), so it is off.
#ifndef B2_PREFETCH
#    define B2_PREFETCH 0
#endif

// Size of StateView::interact_count (CoreState allocates this many counters)
constexpr u32 MAX_INTERACT_MODELS_RESET = 16;

B2_D u32 thread_id()
{
    return blockIdx.x * blockDim.x + threadIdx.x;
}

//---------------------------------------------------------------------------//
// The step kernels are bound by the latency of dependent loads of per-slot state
// (ncu: ~60 % of stall samples are long-scoreboard, spread evenly over ~40 fields;
// L2 hit rate 44 % because the state of 2^20 slots is three times the L2). A thread
// knows its slot at entry, so it asks for every line it is going to touch right away:
// the later loads then find their sectors in (or on the way to) L2.
//---------------------------------------------------------------------------//
B2_D void prefetch_l2(void const* ptr)
{
#if B2_PREFETCH == 2
    asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
#elif B2_PREFETCH
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#else
    (void)ptr;
#endif
}

template<bool CHARGED>
B2_D void prefetch_along_step_state(StateView const& s, u32 slot)
{
    u32 const n = s.num_slots;
    prefetch_l2(s.step_length + slot);
    prefetch_l2(s.energy + slot);
    prefetch_l2(s.particle_id + slot);
    prefetch_l2(s.material_id + slot);
    prefetch_l2(s.post_step_action + slot);
    prefetch_l2(s.interaction_mfp + slot);
    prefetch_l2(s.macro_xs + slot);
    prefetch_l2(s.time + slot);
    prefetch_l2(s.num_steps + slot);
    // geometry (level 0; deeper levels are rare and follow on demand)
    u32 const ng = n * s.max_depth;
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        prefetch_l2(s.geo_pos + k * ng + slot);
        prefetch_l2(s.geo_dir + k * ng + slot);
    }
    prefetch_l2(s.geo_vol + slot);
    prefetch_l2(s.geo_univ + slot);
    prefetch_l2(s.geo_level + slot);
    prefetch_l2(s.geo_surface_level + slot);
    prefetch_l2(s.geo_surf + slot);
    prefetch_l2(s.geo_sense + slot);
    prefetch_l2(s.geo_boundary + slot);
    if (CHARGED)
    {
        prefetch_l2(s.dedx_range + slot);
        prefetch_l2(s.energy_deposition + slot);
#pragma unroll
        for (int k = 0; k < 3; ++k)
            prefetch_l2(s.msc_range + k * n + slot);
#pragma unroll
        for (int k = 0; k < 6; ++k)
            prefetch_l2(s.rng + k * n + slot);
    }
}

B2_D void prefetch_pre_step_state(StateView const& s, u32 slot)
{
    u32 const n = s.num_slots;
    prefetch_l2(s.interaction_mfp + slot);
    prefetch_l2(s.particle_id + slot);
    prefetch_l2(s.energy + slot);
    prefetch_l2(s.material_id + slot);
    prefetch_l2(s.geo_level + slot);
    prefetch_l2(s.geo_vol + slot);
    prefetch_l2(s.geo_univ + slot);
#pragma unroll
    for (int k = 0; k < 6; ++k)
        prefetch_l2(s.rng + k * n + slot);
}

//! i-th active slot: charged from the front, neutral from the back
B2_D u32 active_slot(StateView const& s, u32 tid)
{
    u32 const nc = s.counters[CTR_NUM_CHARGED];
    if (tid < nc)
        return s.track_slots[tid];
    tid -= nc;
    if (tid < s.counters[CTR_NUM_NEUTRAL])
        return s.track_slots[s.num_slots - 1 - tid];
    return INVALID;
}

//---------------------------------------------------------------------------//
// generate: primaries -> track initializers
// (track/detail/ProcessPrimariesExecutor.hh:56-76)
//---------------------------------------------------------------------------//
__global__ void k_extend_from_primaries(StateView s,
                                        B200Primary const* __restrict__ primaries,
                                        u32 const* __restrict__ rank_in_event,
                                        u32 const* __restrict__ neutral_inclusive,
                                        u32 n)
{
    u32 tid = thread_id();
    if (tid >= n)
        return;
    // counters[NUM_INITIALIZERS] has not yet been incremented
    u32 idx = s.counters[CTR_NUM_INITIALIZERS] + tid;
    if (idx >= s.init_capacity)
    {
        s.counters[CTR_ERROR] = B200_ERR_INITIALIZER_CAPACITY;
        return;
    }
    B200Primary const& pr = primaries[tid];
    s.ti_track_id[idx] = s.track_counters[pr.event_id] + rank_in_event[tid];
    s.ti_parent_id[idx] = INVALID;
    s.ti_event_id[idx] = pr.event_id;
    s.ti_time[idx] = pr.time;
    s.ti_particle_id[idx] = pr.particle_id;
    s.ti_energy[idx] = pr.energy;
    s.ti_level[idx] = INVALID;
    for (int k = 0; k < 3; ++k)
    {
        s.ti_pos[k * s.init_capacity + idx] = pr.pos[k];
        s.ti_dir[k * s.init_capacity + idx] = pr.dir[k];
    }
    if (s.ti_neutral_prefix && neutral_inclusive)
    {
        // init_charge: neutral initializers in [0, idx] (host-side inclusive count of the
        // neutral primaries up to this one, on top of the queue's count so far)
        s.ti_neutral_prefix[idx + 1]
            = s.ti_neutral_prefix[s.counters[CTR_NUM_INITIALIZERS]] + neutral_inclusive[tid];
    }
}

__global__ void k_primaries_finalize(StateView s,
                                     u32 const* __restrict__ event_ids,
                                     u32 const* __restrict__ event_counts,
                                     u32 num_events,
                                     u32 n)
{
    u32 tid = thread_id();
    if (tid < num_events)
        s.track_counters[event_ids[tid]] += event_counts[tid];
    if (tid == 0)
    {
        s.counters[CTR_NUM_INITIALIZERS] += n;
        s.counters[CTR_NUM_GENERATED] += n;
    }
}

//---------------------------------------------------------------------------//
// start: initialize tracks in vacant slots
// (track/detail/InitTracksExecutor.hh:71-175)
//---------------------------------------------------------------------------//
__global__ void __launch_bounds__(BLOCK) k_initialize_tracks(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 tid = thread_id();
    u32 const num_init = s.counters[CTR_NUM_INITIALIZERS];
    u32 const num_vac = s.counters[CTR_NUM_VACANCIES];
    u32 const num_new = num_init < num_vac ? num_init : num_vac;
    if (tid >= num_new)
        return;
    u32 ti = num_init - tid - 1;
    u32 slot;
    if (p.scalars.track_order == ORDER_INIT_CHARGE)
    {
        // The reference stable-partitions the num_new initializers about to start into
        // neutral | charged and walks them from the back: charged tracks take the highest
        // vacancies, neutral tracks the lowest (InitTracksExecutor.hh:71-96,
        // detail/Utils.hh:88-98, TrackInitAlgorithms.cc:80-96). With the running neutral
        // count of the queue that is, for the initializer of rank r among the starting
        // neutral (charged) ones: vacancies[r] (vacancies[num_vac - num_charged + r]).
        u32 const first = num_init - num_new;
        u32 const neutral_before_first = s.ti_neutral_prefix[first];
        u32 const num_neutral = s.ti_neutral_prefix[num_init] - neutral_before_first;
        u32 const neutral_rank = s.ti_neutral_prefix[ti] - neutral_before_first;
        bool const is_neutral = p.particle.charge[s.ti_particle_id[ti]] == 0;
        if (is_neutral)
            slot = s.vacancies[neutral_rank];
        else
            slot = s.vacancies[num_vac - (num_new - num_neutral) + ((ti - first) - neutral_rank)];
    }
    else
    {
        slot = s.vacancies[num_vac - tid - 1];
    }

    // sim
    s.track_id[slot] = s.ti_track_id[ti];
    s.parent_id[slot] = s.ti_parent_id[ti];
    s.event_id[slot] = s.ti_event_id[ti];
    s.time[slot] = s.ti_time[ti];
    s.num_steps[slot] = 0;
    s.num_looping_steps[slot] = 0;
    s.status[slot] = ST_INITIALIZING;
    s.step_length[slot] = 0;
    s.post_step_action[slot] = INVALID;
    s.along_step_action[slot] = INVALID;
    // particle
    u32 const pid = s.ti_particle_id[ti];
    s.particle_id[slot] = pid;
    s.energy[slot] = s.ti_energy[ti];
    // append to the dense active lists
    if (p.particle.charge[pid] != 0)
    {
        u32 pos = atomicAdd(&s.counters[CTR_NUM_CHARGED], 1u);
        s.track_slots[pos] = slot;
    }
    else
    {
        u32 pos = atomicAdd(&s.counters[CTR_NUM_NEUTRAL], 1u);
        s.track_slots[s.num_slots - 1 - pos] = slot;
    }
    // geometry
    Real3 pos, dir;
    for (int k = 0; k < 3; ++k)
    {
        pos[k] = s.ti_pos[k * s.init_capacity + ti];
        dir[k] = s.ti_dir[k * s.init_capacity + ti];
    }
    GeoTrack geo(p, s, slot);
    u32 const known_level = s.ti_level[ti];
    if (known_level != INVALID)
    {
        geo.initialize_known(
            pos, dir, known_level, s.ti_vol + ti, s.ti_univ + ti, s.init_capacity);
    }
    else
    {
        geo.initialize(pos, dir);
    }
    bool errored = geo.failed || geo.is_outside();
    u32 matid = INVALID;
    if (!errored)
    {
        matid = p.geo.volume_material[geo.volume_id()];
        errored = (matid == INVALID);
    }
    if (errored)
    {
        // apply_errored (CoreTrackView.hh:340-347)
        s.status[slot] = ST_ERRORED;
        s.along_step_action[slot] = INVALID;
        s.post_step_action[slot] = p.scalars.tracking_cut_action;
        return;
    }
    s.material_id[slot] = matid;
    // physics = {} : reset
    s.interaction_mfp[slot] = 0;
    s.msc_range[slot] = 0;
    s.msc_range[s.num_slots + slot] = 0;
    s.msc_range[2 * s.num_slots + slot] = 0;
}

__global__ void k_initialize_finalize(StateView s)
{
    u32 const num_init = s.counters[CTR_NUM_INITIALIZERS];
    u32 const num_vac = s.counters[CTR_NUM_VACANCIES];
    u32 const num_new = num_init < num_vac ? num_init : num_vac;
    s.counters[CTR_NUM_INITIALIZERS] = num_init - num_new;
    s.counters[CTR_NUM_VACANCIES] = num_vac - num_new;
    s.counters[CTR_NUM_ACTIVE] = s.num_slots - (num_vac - num_new);
    s.counters[CTR_NUM_NEW_TRACKS] = num_new;
    // recomputed by this step's end pass (atomicMin over the blocks that hold tracks)
    s.counters[CTR_FIRST_BUSY_BLOCK] = INVALID;
    // per-model interaction lists are rebuilt by this step's discrete select
    if (s.interact_count)
    {
        for (u32 m = 0; m < MAX_INTERACT_MODELS_RESET; ++m)
            s.interact_count[m] = 0;
    }
    // whole-run tallies kept on the device: track-steps and step iterations
    s.step_counters[0] += s.num_slots - (num_vac - num_new);
    s.step_counters[1] += 1;
}

//---------------------------------------------------------------------------//
// pre: physics step limits (phys/detail/PreStepExecutor.hh:45-115)
//---------------------------------------------------------------------------//
B2_D void do_pre_step(ParamsView const& p, StateView const& s, u32 slot)
{
    u8 status = s.status[slot];
    s.energy_deposition[slot] = 0;
    for (int i = 0; i < MAX_SECONDARIES; ++i)
        s.sec_particle[i * s.num_slots + slot] = INVALID;
    s.element[slot] = INVALID;
    // pre-step volume for detector scoring (StepGatherExecutor<pre> runs for every
    // non-inactive track, errored ones included: their energy is deposited by the
    // tracking cut in the volume they are in; a track that started outside the geometry
    // is in the exterior volume, which never is a detector)
    if (s.pre_volume)
    {
        GeoTrack geo(p, s, slot);
        s.pre_volume[slot] = geo.volume_id();
    }
    if (status == ST_ERRORED)
        return;
    s.status[slot] = ST_ALIVE;

    if (!(s.interaction_mfp[slot] > 0))
    {
        Rng rng;
        rng.load(s, slot);
        s.interaction_mfp[slot] = sample_exponential(rng);
        rng.store(s, slot);
    }
    Particle particle = load_particle(p, s, slot);
    PhysTrack phys(p, particle.id, s.material_id[slot]);
    StepLimit limit = calc_physics_step_limit(p, s, slot, particle, phys);
    s.step_length[slot] = limit.step;
    s.post_step_action[slot] = limit.action;
    s.along_step_action[slot] = (particle.charge == 0) ? p.scalars.along_step_neutral_action
                                                       : p.scalars.along_step_user_action;
}

__global__ void __launch_bounds__(BLOCK, B2_PRE_MIN_BLOCKS) k_pre_step(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 slot = active_slot(s, thread_id());
    if (slot != INVALID)
    {
        prefetch_pre_step_state(s, slot);
        do_pre_step(p, s, slot);
    }
}

//---------------------------------------------------------------------------//
// pre-post: discrete select (phys/detail/DiscreteSelectExecutor.hh:37-63)
//---------------------------------------------------------------------------//
B2_D void do_discrete_select(ParamsView const& p, StateView const& s, u32 slot)
{
    if (s.status[slot] != ST_ALIVE)
        return;
    if (s.post_step_action[slot] != p.phys.model_to_action - 2)
        return;
    s.interaction_mfp[slot] = 0;
    Particle particle = load_particle(p, s, slot);
    PhysTrack phys(p, particle.id, s.material_id[slot]);
    Rng rng;
    rng.load(s, slot);
    u32 action = select_discrete_interaction(p, s, slot, particle, phys, rng);
    rng.store(s, slot);
    s.post_step_action[slot] = action;
}

// Select (if the track's step ended at a discrete interaction) and append the interacting
// track to its model's slot list. Called by EVERY thread of the block (slot == INVALID for
// the ones without a track): the appends are aggregated per block in shared memory.
constexpr u32 MAX_INTERACT_MODELS = 16;

B2_D void select_and_append(ParamsView const& p, StateView const& s, u32 slot)
{
    __shared__ u32 count[MAX_INTERACT_MODELS];
    __shared__ u32 base[MAX_INTERACT_MODELS];
    bool const build_lists = s.interact_list != nullptr;
    if (build_lists)
    {
        if (threadIdx.x < MAX_INTERACT_MODELS)
            count[threadIdx.x] = 0;
        __syncthreads();
    }
    u32 model = INVALID;
    if (slot != INVALID)
    {
        do_discrete_select(p, s, slot);
        if (build_lists && s.status[slot] == ST_ALIVE)
        {
            u32 m = s.post_step_action[slot] - p.phys.model_to_action;
            if (m < p.phys.num_models)
                model = m;
        }
    }
    if (!build_lists)
        return;
    u32 rank = 0;
    if (model != INVALID)
        rank = atomicAdd(&count[model], 1u);
    __syncthreads();
    if (threadIdx.x < p.phys.num_models && count[threadIdx.x] > 0)
        base[threadIdx.x] = atomicAdd(&s.interact_count[threadIdx.x], count[threadIdx.x]);
    __syncthreads();
    if (model != INVALID)
        s.interact_list[size_t(model) * s.num_slots + base[model] + rank] = slot;
}

//---------------------------------------------------------------------------//
// along-step: one launch per charge class over its dense list
//---------------------------------------------------------------------------//
// SELECT: the discrete-process selection (order pre_post, the next action in the sequence)
// is done by the same thread right after its along-step, and the selected interactions are
// appended to the per-model lists: one launch and one pass over the active tracks less.
template<bool FIELD, bool SELECT>
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

// The same charged along-step as four phase kernels (see along_step.cuh). Used for large
// iterations; small ones stay fused, where launch latency matters more than occupancy.
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
B2_ALONG_PHASE_KERNEL(k_along_propagate_linear, along_phase_propagate<false>, B2_PHASE_MIN_BLOCKS)
B2_ALONG_PHASE_KERNEL(k_along_propagate_field, along_phase_propagate<true>, B2_PROPAGATE_FIELD_MIN_BLOCKS)
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
            along_step<false, false>(p, s, slot);
    }
    if (SELECT)
        select_and_append(p, s, slot);
}

// Besides selecting, the launch sorts the interacting tracks BY MODEL into per-model slot
// lists (block-aggregated appends), so that the interaction kernel runs warps in which
// every lane executes the same interactor. Launched over all active tracks, only ~4 of
// 32 lanes were active per instruction (ncu: profiles/README_r01.md).
__global__ void __launch_bounds__(BLOCK) k_discrete_select(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    select_and_append(p, s, active_slot(s, thread_id()));
}

//---------------------------------------------------------------------------//
// post: every EM model (dispatch on the selected action id)
//---------------------------------------------------------------------------//
B2_D void do_interact(ParamsView const& p, StateView const& s, u32 slot)
{
    if (s.status[slot] != ST_ALIVE)
        return;
    u32 action = s.post_step_action[slot];
    if (action < p.phys.model_to_action || action >= p.phys.model_to_action + p.phys.num_models)
        return;
    Rng rng;
    rng.load(s, slot);
    run_interaction(p, s, slot, action, rng);
    rng.store(s, slot);
}

__global__ void __launch_bounds__(BLOCK) k_interact(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 slot = active_slot(s, thread_id());
    if (slot != INVALID)
        do_interact(p, s, slot);
}

//! Interactions over the per-model lists built by k_discrete_select: thread t works on
//! the t-th interacting track in model order
__global__ void __launch_bounds__(BLOCK, B2_INTERACT_MIN_BLOCKS) k_interact_lists(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 tid = thread_id();
    u32 const num_models = p.phys.num_models;
    u32 model = 0;
    u32 count = 0;
    for (; model < num_models; ++model)
    {
        // every model's segment starts on a warp boundary: no warp mixes interactors
        count = s.interact_count[model];
        u32 const padded = (count + 31u) & ~31u;
        if (tid < padded)
            break;
        tid -= padded;
    }
    if (model == num_models || tid >= count)
        return;
    do_interact(p, s, s.interact_list[size_t(model) * s.num_slots + tid]);
}

// (geo/detail/BoundaryExecutor.hh:41-84)
B2_D void do_boundary(ParamsView const& p, StateView const& s, u32 slot)
{
    if (s.status[slot] != ST_ALIVE || s.post_step_action[slot] != p.scalars.boundary_action)
        return;
    GeoTrack geo(p, s, slot);
    geo.cross_boundary();
    bool errored = geo.failed;
    if (!errored && !geo.is_outside())
    {
        u32 matid = p.geo.volume_material[geo.volume_id()];
        if (matid == INVALID)
            errored = true;
        else
            s.material_id[slot] = matid;
    }
    else if (!errored)
    {
        s.status[slot] = ST_KILLED;
    }
    if (errored)
    {
        s.status[slot] = ST_ERRORED;
        s.along_step_action[slot] = INVALID;
        s.post_step_action[slot] = p.scalars.tracking_cut_action;
    }
}

__global__ void __launch_bounds__(BLOCK) k_boundary(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 slot = active_slot(s, thread_id());
    if (slot != INVALID)
        do_boundary(p, s, slot);
}

// (phys/detail/TrackingCutExecutor.hh:48-83)
B2_D void do_tracking_cut(ParamsView const& p, StateView const& s, u32 slot)
{
    u8 status = s.status[slot];
    if (status == ST_INACTIVE || status == ST_KILLED)
        return;
    if (s.post_step_action[slot] != p.scalars.tracking_cut_action)
        return;
    u32 pid = s.particle_id[slot];
    real deposited = s.energy[slot];
    if (particle_is_antiparticle(p, pid))
        deposited += 2 * p.particle.mass[pid];
    s.energy_deposition[slot] += deposited;
    s.energy[slot] = 0;
    s.status[slot] = ST_KILLED;
}

__global__ void __launch_bounds__(BLOCK) k_tracking_cut(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 slot = active_slot(s, thread_id());
    if (slot != INVALID)
        do_tracking_cut(p, s, slot);
}

// user_post: tallies (user/detail/SimpleCaloExecutor.hh:48-67)
// Per-detector sums are first accumulated in shared memory (one copy per block) and
// flushed with one global atomic per touched bin, instead of one contended global
// atomic per depositing track.
constexpr u32 TALLY_SMEM_BINS = 1024;

__global__ void __launch_bounds__(BLOCK) k_tally(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s, u32 num_det)
{
    __shared__ real bins[TALLY_SMEM_BINS];
    bool const use_smem = num_det <= TALLY_SMEM_BINS;
    if (use_smem)
    {
        for (u32 i = threadIdx.x; i < num_det; i += BLOCK)
            bins[i] = 0;
        __syncthreads();
    }
    u32 slot = active_slot(s, thread_id());
    if (slot != INVALID && s.status[slot] != ST_INACTIVE)
    {
        real edep = s.energy_deposition[slot];
        if (edep != 0)
        {
            u32 det = s.calo_detector_of_volume[s.pre_volume[slot]];
            if (det != INVALID)
            {
                if (use_smem)
                    atomicAdd(&bins[det], edep);
                else
                    atomicAdd(&s.calo_edep[det], edep);
            }
        }
    }
    if (use_smem)
    {
        __syncthreads();
        for (u32 i = threadIdx.x; i < num_det; i += BLOCK)
        {
            real v = bins[i];
            if (v != 0)
                atomicAdd(&s.calo_edep[i], v);
        }
    }
}

//---------------------------------------------------------------------------//
// diagnostics
// post: tally the post-step action of every track that took this step
//   (user/detail/ActionDiagnosticExecutor.hh:30-65)
// user_post: tally the number of steps of every track killed this step
//   (user/detail/StepDiagnosticExecutor.hh:28-60)
// Counts are first gathered per block in shared memory (a handful of bins are hot).
//---------------------------------------------------------------------------//
constexpr u32 DIAG_SMEM_BINS = 1024;

template<bool STEPS>
__global__ void __launch_bounds__(BLOCK) k_diagnostic(StateView s, u32 num_particles)
{
    __shared__ u32 bins[DIAG_SMEM_BINS];
    u32* const out = STEPS ? s.diag_step_counts : s.diag_action_counts;
    u32 const nb = STEPS ? s.diag_step_bins : s.diag_action_bins;
    u32 const total = nb * num_particles;
    bool const use_smem = total <= DIAG_SMEM_BINS;
    if (use_smem)
    {
        for (u32 i = threadIdx.x; i < total; i += BLOCK)
            bins[i] = 0;
        __syncthreads();
    }
    u32 slot = active_slot(s, thread_id());
    if (slot != INVALID)
    {
        u8 status = s.status[slot];
        u32 bin = INVALID;
        if (!STEPS && status != ST_INACTIVE)
        {
            bin = s.particle_id[slot] * nb + s.post_step_action[slot];
        }
        if (STEPS && status == ST_KILLED)
        {
            u32 n = s.num_steps[slot];
            bin = s.particle_id[slot] * nb + (n < nb - 1 ? n : nb - 1);
        }
        if (bin != INVALID)
            atomicAdd(use_smem ? &bins[bin] : &out[bin], 1u);
    }
    if (use_smem)
    {
        __syncthreads();
        for (u32 i = threadIdx.x; i < total; i += BLOCK)
        {
            u32 v = bins[i];
            if (v != 0)
                atomicAdd(&out[i], v);
        }
    }
}

//---------------------------------------------------------------------------//
// Whole step of one track in one launch: pre-step, along-step, discrete select,
// interaction, boundary, tracking cut, tallies and diagnostics, in action order.
//
// Every one of those actions touches only its own slot (plus atomic tallies), so running
// them back to back per thread gives exactly the per-action results. This is the path for
// SMALL iterations (shower tails): there the cost of a step is not throughput but the
// latency of ten dependent launches, each of which starts with cold instruction and data
// caches -- measured floor 213 us per iteration for a single 1 GeV shower, of which the
// per-action kernels account for ~165 us (profiles/README_r01.md).
//---------------------------------------------------------------------------//
template<bool FIELD>
__global__ void __launch_bounds__(BLOCK, B2_FUSED_MIN_BLOCKS)
    k_step_fused(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 const tid = thread_id();
    u32 const slot = active_slot(s, tid);
    if (slot == INVALID)
        return;
    bool const charged = tid < s.counters[CTR_NUM_CHARGED];
    // pre
    do_pre_step(p, s, slot);
    // along
    if (s.status[slot] == ST_ALIVE)
    {
        if (charged)
            along_step<true, FIELD>(p, s, slot);
        else
            along_step<false, false>(p, s, slot);
    }
    // pre_post, post
    do_discrete_select(p, s, slot);
    do_interact(p, s, slot);
    do_boundary(p, s, slot);
    do_tracking_cut(p, s, slot);
    u8 const status = s.status[slot];
    if (s.diag_action_counts && status != ST_INACTIVE)
    {
        atomicAdd(&s.diag_action_counts[s.particle_id[slot] * s.diag_action_bins
                                        + s.post_step_action[slot]],
                  1u);
    }
    // user_post
    if (s.calo_edep && status != ST_INACTIVE)
    {
        real edep = s.energy_deposition[slot];
        if (edep != 0)
        {
            u32 det = s.calo_detector_of_volume[s.pre_volume[slot]];
            if (det != INVALID)
                atomicAdd(&s.calo_edep[det], edep);
        }
    }
    if (s.diag_step_counts && status == ST_KILLED)
    {
        u32 const nb = s.diag_step_bins;
        u32 const n = s.num_steps[slot];
        atomicAdd(&s.diag_step_counts[s.particle_id[slot] * nb + (n < nb - 1 ? n : nb - 1)], 1u);
    }
}

//---------------------------------------------------------------------------//
// The tail of a step in one launch for LARGE iterations: boundary crossing, tracking
// cut, action diagnostic (order post), calorimeter tally and step diagnostic (order
// user_post) are consecutive in the action sequence, each touches only its own slot, and
// all but the boundary crossing are a few instructions per track: as separate launches
// they each re-read the slot lists and the status and post-step action of every track.
// Tallies go through per-block shared-memory bins as in k_tally / k_diagnostic.
//---------------------------------------------------------------------------//
__global__ void __launch_bounds__(BLOCK, B2_TAIL_MIN_BLOCKS)
    k_post_tail(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    __shared__ real calo_bins[TALLY_SMEM_BINS];
    __shared__ u32 action_bins[DIAG_SMEM_BINS];
    __shared__ u32 step_bins[DIAG_SMEM_BINS];
    u32 const num_det = s.calo_edep ? s.num_detectors : 0;
    u32 const num_particles = p.particle.num_particles;
    u32 const num_action = s.diag_action_counts ? s.diag_action_bins * num_particles : 0;
    u32 const num_step = s.diag_step_counts ? s.diag_step_bins * num_particles : 0;
    bool const calo_smem = num_det <= TALLY_SMEM_BINS;
    bool const action_smem = num_action <= DIAG_SMEM_BINS;
    bool const step_smem = num_step <= DIAG_SMEM_BINS;
    static_assert(TALLY_SMEM_BINS == DIAG_SMEM_BINS, "one loop zeroes and flushes all bins");
    u32 used_bins = calo_smem ? num_det : 0;
    if (action_smem && num_action > used_bins)
        used_bins = num_action;
    if (step_smem && num_step > used_bins)
        used_bins = num_step;
    for (u32 i = threadIdx.x; i < used_bins; i += BLOCK)
    {
        calo_bins[i] = 0;
        action_bins[i] = 0;
        step_bins[i] = 0;
    }
    __syncthreads();

    u32 const slot = active_slot(s, thread_id());
    if (slot != INVALID)
    {
        do_boundary(p, s, slot);
        do_tracking_cut(p, s, slot);
        u8 const status = s.status[slot];
        if (num_action && status != ST_INACTIVE)
        {
            u32 bin = s.particle_id[slot] * s.diag_action_bins + s.post_step_action[slot];
            atomicAdd(action_smem ? &action_bins[bin] : &s.diag_action_counts[bin], 1u);
        }
        if (num_det && status != ST_INACTIVE)
        {
            real edep = s.energy_deposition[slot];
            if (edep != 0)
            {
                u32 det = s.calo_detector_of_volume[s.pre_volume[slot]];
                if (det != INVALID)
                    atomicAdd(calo_smem ? &calo_bins[det] : &s.calo_edep[det], edep);
            }
        }
        if (num_step && status == ST_KILLED)
        {
            u32 const nb = s.diag_step_bins;
            u32 const n = s.num_steps[slot];
            u32 bin = s.particle_id[slot] * nb + (n < nb - 1 ? n : nb - 1);
            atomicAdd(step_smem ? &step_bins[bin] : &s.diag_step_counts[bin], 1u);
        }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < used_bins; i += BLOCK)
    {
        if (calo_smem && i < num_det)
        {
            real v = calo_bins[i];
            if (v != 0)
                atomicAdd(&s.calo_edep[i], v);
        }
        if (action_smem && i < num_action)
        {
            u32 v = action_bins[i];
            if (v != 0)
                atomicAdd(&s.diag_action_counts[i], v);
        }
        if (step_smem && i < num_step)
        {
            u32 v = step_bins[i];
            if (v != 0)
                atomicAdd(&s.diag_step_counts[i], v);
        }
    }
}

//---------------------------------------------------------------------------//
// end: secondaries -> initializers, vacancy compaction, dense active lists
// (track/detail/LocateAliveExecutor.hh:60-106,
//  track/detail/ProcessSecondariesExecutor.hh:69-183,
//  track/detail/TrackInitAlgorithms.cu:34-78)
//
// Pass 1 (per block): classify each slot, block-level exclusive scans of five
//   quantities packed into two words, block totals to scratch.
// Pass 2 (one block): scan of block totals -> block offsets, global counters.
// Pass 3 (per block): write compacted vacancies, track initializers and the
//   dense charged/neutral lists of the slots that stay active.
//---------------------------------------------------------------------------//
template<int B, class T>
B2_D T block_exclusive_scan(T value, T* total)
{
    __shared__ T warp_sums[B / 32];
    __shared__ T block_total;
    u32 lane = threadIdx.x & 31;
    u32 warp = threadIdx.x >> 5;
    T incl = value;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1)
    {
        T n = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off)
            incl += n;
    }
    if (lane == 31)
        warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        T w = lane < B / 32 ? warp_sums[lane] : T(0);
        T wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1)
        {
            T n = __shfl_up_sync(0xffffffffu, wi, off);
            if (lane >= off)
                wi += n;
        }
        if (lane < B / 32)
            warp_sums[lane] = wi - w;
        if (lane == 31)
            block_total = wi;
    }
    __syncthreads();
    T result = warp_sums[warp] + incl - value;
    *total = block_total;
    __syncthreads();
    return result;
}

struct SlotEnd
{
    u32 is_vacant;
    u32 num_sec;      // secondaries that become initializers
    u32 num_sec_all;  // including one that reuses the slot in place
    u32 num_sec_neutral;  // of num_sec, how many are neutral (init_charge bookkeeping)
    u32 charged;      // stays active with a charged particle
    u32 neutral;      // stays active with a neutral particle
    bool reuse_slot;  // first secondary replaces a dead parent in place
    bool inactive;    // status == inactive at the end of the step
};

// The classification is computed once (pass 1) and handed to pass 3 as one byte per slot:
// pass 3 then needs a single coalesced load instead of the chain status -> secondaries ->
// particle -> charge that it took to classify.
B2_D u8 pack_class(SlotEnd const& e)
{
    return u8(e.is_vacant | (e.charged << 1) | (u32(e.inactive) << 2) | (e.num_sec << 3)
              | (e.num_sec_neutral << 5) | (u32(e.reuse_slot) << 7));
}

B2_D SlotEnd unpack_class(u8 c)
{
    SlotEnd e;
    e.is_vacant = c & 1u;
    e.charged = (c >> 1) & 1u;
    e.inactive = (c >> 2) & 1u;
    e.num_sec = (c >> 3) & 3u;
    e.num_sec_neutral = (c >> 5) & 3u;
    e.reuse_slot = (c >> 7) & 1u;
    e.num_sec_all = e.num_sec + (e.reuse_slot ? 1u : 0u);
    e.neutral = (!e.is_vacant && !e.charged) ? 1u : 0u;
    return e;
}
static_assert(MAX_SECONDARIES <= 3, "two bits per secondary count in the class byte");

B2_D SlotEnd classify_slot(ParamsView const& p, StateView const& s, u32 slot)
{
    SlotEnd r{0, 0, 0, 0, 0, 0, false, false};
    if (slot >= s.num_slots)
        return r;
    u8 status = s.status[slot];
    r.inactive = (status == ST_INACTIVE);
    u32 first_sec = INVALID;
    bool const by_charge = p.scalars.track_order == ORDER_INIT_CHARGE;
    if (status != ST_INACTIVE)
    {
        for (int i = MAX_SECONDARIES - 1; i >= 0; --i)
        {
            u32 sp = s.sec_particle[i * s.num_slots + slot];
            if (sp != INVALID)
            {
                ++r.num_sec;
                first_sec = sp;
                if (by_charge && p.particle.charge[sp] == 0)
                    ++r.num_sec_neutral;
            }
        }
    }
    r.num_sec_all = r.num_sec;
    u32 active_particle = INVALID;
    if (status == ST_ALIVE)
    {
        active_particle = s.particle_id[slot];
    }
    else if (r.num_sec > 0 && p.scalars.track_order != ORDER_INIT_CHARGE)
    {
        --r.num_sec;
        r.reuse_slot = true;
        active_particle = first_sec;
    }
    else
    {
        r.is_vacant = 1;
    }
    if (active_particle != INVALID)
    {
        bool charged = p.particle.charge[active_particle] != 0;
        r.charged = charged;
        r.neutral = !charged;
    }
    return r;
}

// Packed scan words: A = vacant | charged << 10 | neutral << 20 (each <= BLOCK),
//   B = num_sec | num_sec_all << 10 | num_sec_neutral << 20 (each <= MAX_SECONDARIES * BLOCK)
static_assert(BLOCK <= 512 && MAX_SECONDARIES * BLOCK < 1024, "packed scan field widths");

B2_D u32 pack_secondaries(SlotEnd const& e)
{
    return e.num_sec | (e.num_sec_all << 10) | (e.num_sec_neutral << 20);
}

__global__ void __launch_bounds__(BLOCK) k_end_pass1(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 slot = s.slot_begin + thread_id();
    SlotEnd e = classify_slot(p, s, slot);
    u32 ta, tb;
    block_exclusive_scan<BLOCK, u32>(e.is_vacant | (e.charged << 10) | (e.neutral << 20), &ta);
    block_exclusive_scan<BLOCK, u32>(pack_secondaries(e), &tb);
    // Lowest block that held a track during this step: next step's passes start there
    bool const busy = slot < s.num_slots && !e.inactive;
    if (slot < s.num_slots)
        s.slot_class[slot] = pack_class(e);
    bool const any_busy = __syncthreads_or(busy);
    if (threadIdx.x == 0)
    {
        if (any_busy)
            atomicMin(&s.counters[CTR_FIRST_BUSY_BLOCK], slot / BLOCK);
        u32 const nb = gridDim.x;
        s.block_scratch[blockIdx.x] = ta & 0x3ffu;
        s.block_scratch[nb + blockIdx.x] = (ta >> 10) & 0x3ffu;
        s.block_scratch[2 * nb + blockIdx.x] = (ta >> 20) & 0x3ffu;
        s.block_scratch[3 * nb + blockIdx.x] = tb & 0x3ffu;
        s.block_scratch[4 * nb + blockIdx.x] = (tb >> 10) & 0x3ffu;
        s.block_scratch[5 * nb + blockIdx.x] = (tb >> 20) & 0x3ffu;
    }
}

__global__ void __launch_bounds__(1024) k_end_pass2(StateView s, u32 num_blocks)
{
    // Six blocks, one per scanned quantity: each thread owns a run of consecutive
    // block totals; the last block to finish publishes the global counters.
    constexpr int B = 1024;
    u32 const a = blockIdx.x;
    u32 const per = (num_blocks + B - 1) / B;
    u32 const begin = threadIdx.x * per;
    u32 const end = begin + per < num_blocks ? begin + per : num_blocks;
    u32* const scratch = s.block_scratch + a * num_blocks;
    u32 local = 0;
    u32 total;
    constexpr u32 MAX_PER = 16;
    if (per <= MAX_PER)
    {
        // The thread's run is held in registers: all loads are in flight together
        // (a load/add/store loop costs one memory round trip per element)
        u32 v[MAX_PER];
#pragma unroll
        for (u32 k = 0; k < MAX_PER; ++k)
            v[k] = (begin + k < end) ? scratch[begin + k] : 0u;
#pragma unroll
        for (u32 k = 0; k < MAX_PER; ++k)
            local += v[k];
        u32 run = block_exclusive_scan<B, u32>(local, &total);
#pragma unroll
        for (u32 k = 0; k < MAX_PER; ++k)
        {
            if (begin + k < end)
                scratch[begin + k] = run;
            run += v[k];
        }
    }
    else
    {
        for (u32 i = begin; i < end; ++i)
            local += scratch[i];
        u32 run = block_exclusive_scan<B, u32>(local, &total);
        for (u32 i = begin; i < end; ++i)
        {
            u32 v = scratch[i];
            scratch[i] = run;
            run += v;
        }
    }
    __shared__ bool is_last;
    if (threadIdx.x == 0)
    {
        s.counters[CTR_SCAN_TOTALS + a] = total;
        __threadfence();
        u32 done = atomicAdd(&s.counters[CTR_SCAN_DONE], 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0)
    {
        __threadfence();
        s.counters[CTR_SCAN_DONE] = 0;
        u32 carry[6];
        for (int k = 0; k < 6; ++k)
            carry[k] = reinterpret_cast<u32 volatile*>(s.counters)[CTR_SCAN_TOTALS + k];
        // slots below slot_begin are all vacant
        u32 num_vac = carry[0] + s.slot_begin;
        u32 num_sec = carry[3];
        s.counters[CTR_NUM_VACANCIES] = num_vac;
        s.counters[CTR_NUM_CHARGED] = carry[1];
        s.counters[CTR_NUM_NEUTRAL] = carry[2];
        s.counters[CTR_NUM_SECONDARIES] = num_sec;
        u32 num_init = s.counters[CTR_NUM_INITIALIZERS] + num_sec;
        s.counters[CTR_NUM_INITIALIZERS] = num_init;
        s.counters[CTR_NUM_ALIVE] = s.num_slots - num_vac;
        if (num_init > s.init_capacity)
            s.counters[CTR_ERROR] = B200_ERR_INITIALIZER_CAPACITY;
        // Single event in flight: track ids are assigned in slot order from the
        // scan (what the reference's sequential host loop produces)
        if (s.single_event != INVALID)
        {
            s.counters[CTR_TRACK_ID_BASE] = s.track_counters[s.single_event];
            s.track_counters[s.single_event] += carry[4];
        }
        // Publish the step's counters to the host: nothing after this point changes them
        if (s.host_counters)
        {
            u32 volatile* host = s.host_counters;
            for (u32 k = 0; k < CTR_SIZE; ++k)
                host[k] = reinterpret_cast<u32 volatile*>(s.counters)[k];
            __threadfence_system();
            host[CTR_SIZE] = s.iteration_seq;
        }
    }
}

__global__ void __launch_bounds__(BLOCK, B2_END_MIN_BLOCKS) k_end_pass3(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 slot = s.slot_begin + thread_id();
    // classification of pass 1 (the same launch sequence; nothing changed in between)
    SlotEnd e = unpack_class(slot < s.num_slots ? s.slot_class[slot] : u8(0));
    u32 ta, tb;
    u32 const nb = gridDim.x;
    // Everything that does not depend on the scans is loaded BEFORE their barriers, so
    // that these round trips overlap with the scans instead of queueing up behind them
    // (the kernel is latency bound: 38 long-scoreboard stall cycles per issue, ncu)
    u32 const block_vac = s.block_scratch[blockIdx.x];
    u32 const block_chg = s.block_scratch[nb + blockIdx.x];
    u32 const block_neu = s.block_scratch[2 * nb + blockIdx.x];
    u32 const block_sec = s.block_scratch[3 * nb + blockIdx.x];
    u32 const block_all = s.block_scratch[4 * nb + blockIdx.x];
    u32 const block_neutral_sec = s.block_scratch[5 * nb + blockIdx.x];
    u32 const device_error = s.counters[CTR_ERROR];
    u32 const num_init = s.counters[CTR_NUM_INITIALIZERS];
    u32 const num_sec_total = s.counters[CTR_NUM_SECONDARIES];
    u32 event = 0, parent_track = 0;
    real time = 0;
    if (e.num_sec_all > 0)
    {
        event = s.event_id[slot];
        parent_track = s.track_id[slot];
        time = s.time[slot];
    }
    u32 sa = block_exclusive_scan<BLOCK, u32>(
        e.is_vacant | (e.charged << 10) | (e.neutral << 20), &ta);
    u32 sb = block_exclusive_scan<BLOCK, u32>(pack_secondaries(e), &tb);
    // vacancies[i] = i for the (all vacant) slots below slot_begin
    u32 vac_off = s.slot_begin + (sa & 0x3ffu) + block_vac;
    u32 chg_off = ((sa >> 10) & 0x3ffu) + block_chg;
    u32 neu_off = ((sa >> 20) & 0x3ffu) + block_neu;
    u32 sec_off = (sb & 0x3ffu) + block_sec;
    u32 all_off = ((sb >> 10) & 0x3ffu) + block_all;
    // neutral initializers created by lower slots in this step (init_charge only)
    u32 neutral_off = ((sb >> 20) & 0x3ffu) + block_neutral_sec;
    if (slot >= s.num_slots)
        return;
    if (device_error != 0)
        return;
    if (e.is_vacant)
        s.vacancies[vac_off] = slot;
    if (e.charged)
        s.track_slots[chg_off] = slot;
    if (e.neutral)
        s.track_slots[s.num_slots - 1 - neu_off] = slot;

    if (e.inactive)
    {
        // The reference's pre-step resets the step limit of inactive slots
        // (PreStepExecutor.hh:47-57); inactive slots are never visited by the dense
        // kernels here, so do it once, when the slot is first seen inactive
        if (s.post_step_action[slot] != INVALID || s.along_step_action[slot] != INVALID)
        {
            s.step_length[slot] = real_inf();
            s.post_step_action[slot] = INVALID;
            s.along_step_action[slot] = INVALID;
        }
        return;
    }

    // Initializers created this step occupy [num_init - num_sec, num_init)
    // in slot order (exclusive scan of the per-slot counts)
    u32 out = num_init - num_sec_total + sec_off;
    u32 neutral_run = 0;
    if (s.ti_neutral_prefix && e.num_sec_all > 0)
        neutral_run = s.ti_neutral_prefix[num_init - num_sec_total] + neutral_off;
    bool initialized = false;
    u32 const n = s.num_slots;
    u32 const cap = s.init_capacity;

    if (e.num_sec_all > 0)
    {
        GeoTrack geo(p, s, slot);
        Real3 const pos = geo.pos();
        u32 const lev = geo.level();

        // Track ids: per-event counter (reference: atomic_add, detail/Utils.hh:107-116);
        // deterministic slot-order ids when a single event is in flight
        u32 id_base;
        if (s.single_event != INVALID)
            id_base = s.counters[CTR_TRACK_ID_BASE] + all_off;
        else
        {
            // One atomic per (warp, event) instead of one per track: with merged events
            // a step creates ~1e5 secondaries on a few dozen counters. Lanes of the same
            // event take consecutive ids in lane order (counts are 1 or 2 per lane).
            unsigned const active = __activemask();
            unsigned const group = __match_any_sync(active, event);
            unsigned const lane = threadIdx.x & 31u;
            unsigned const lower = group & ((1u << lane) - 1u);
            unsigned const ones = __ballot_sync(active, e.num_sec_all == 1);
            unsigned const twos = __ballot_sync(active, e.num_sec_all == 2);
            static_assert(MAX_SECONDARIES == 2, "ballot-based offsets assume 1 or 2");
            u32 const offset = __popc(lower & ones) + 2 * __popc(lower & twos);
            u32 const total = __popc(group & ones) + 2 * __popc(group & twos);
            int const leader = __ffs(group) - 1;
            u32 base = 0;
            if (int(lane) == leader)
                base = atomicAdd(&s.track_counters[event], total);
            base = __shfl_sync(group, base, leader);
            id_base = base + offset;
        }

        for (int i = 0; i < MAX_SECONDARIES; ++i)
        {
            u32 spid = s.sec_particle[i * n + slot];
            if (spid == INVALID)
                continue;
            real senergy = s.sec_energy[i * n + slot];
            Real3 sdir = make_real3(s.sec_dir[(i * 3 + 0) * n + slot],
                                    s.sec_dir[(i * 3 + 1) * n + slot],
                                    s.sec_dir[(i * 3 + 2) * n + slot]);
            u32 new_id = id_base++;
            if (!initialized && e.reuse_slot)
            {
                // The first secondary takes over the dead parent's slot
                s.track_id[slot] = new_id;
                s.parent_id[slot] = parent_track;
                s.num_steps[slot] = 0;
                s.num_looping_steps[slot] = 0;
                s.status[slot] = ST_INITIALIZING;
                s.step_length[slot] = 0;
                s.post_step_action[slot] = INVALID;
                s.along_step_action[slot] = INVALID;
                geo.initialize_from(slot, sdir);
                s.particle_id[slot] = spid;
                s.energy[slot] = senergy;
                s.interaction_mfp[slot] = 0;
                s.msc_range[slot] = 0;
                s.msc_range[n + slot] = 0;
                s.msc_range[2 * n + slot] = 0;
                initialized = true;
            }
            else
            {
                s.ti_track_id[out] = new_id;
                s.ti_parent_id[out] = parent_track;
                s.ti_event_id[out] = event;
                s.ti_time[out] = time;
                s.ti_particle_id[out] = spid;
                s.ti_energy[out] = senergy;
                for (int k = 0; k < 3; ++k)
                {
                    s.ti_pos[k * cap + out] = pos[k];
                    s.ti_dir[k * cap + out] = sdir[k];
                }
                // the parent's volume hierarchy at this point
                s.ti_level[out] = lev;
                for (u32 l = 0; l <= lev; ++l)
                {
                    s.ti_vol[l * cap + out] = s.geo_vol[l * n + slot];
                    s.ti_univ[l * cap + out] = s.geo_univ[l * n + slot];
                }
                if (s.ti_neutral_prefix)
                {
                    // running count of neutral initializers in queue order
                    if (p.particle.charge[spid] == 0)
                        ++neutral_run;
                    s.ti_neutral_prefix[out + 1] = neutral_run;
                }
                ++out;
            }
        }
    }
    // a vacant slot that is not yet inactive holds a killed track
    if (!initialized && e.is_vacant && s.status[slot] == ST_KILLED)
        s.status[slot] = ST_INACTIVE;
}

//---------------------------------------------------------------------------//
// reseed (random/RngReseed.cu:29-74)
//---------------------------------------------------------------------------//
__global__ void k_reseed(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s, u64 event_id)
{
    u32 slot = thread_id();
    if (slot >= s.num_slots)
        return;
    Rng rng;
    rng.initialize(p.rng, p.rng.seed, event_id * u64(s.num_slots) + slot, 0);
    rng.store(s, slot);
}

__global__ void k_reset_generated(StateView s)
{
    s.counters[CTR_NUM_GENERATED] = 0;
}

__global__ void k_kill_active(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 slot = thread_id();
    if (slot >= s.num_slots)
        return;
    if (s.status[slot] == ST_INACTIVE)
        return;
    s.status[slot] = ST_ERRORED;
    s.along_step_action[slot] = INVALID;
    s.post_step_action[slot] = p.scalars.tracking_cut_action;
}
}  // namespace b200

//---------------------------------------------------------------------------//
// C-ABI launchers
//---------------------------------------------------------------------------//
using namespace b200;

namespace
{
std::atomic<uint64_t> g_launches{0};
#define B2_COUNT(n) g_launches.fetch_add(n, std::memory_order_relaxed)
inline unsigned grid_for(u32 n)
{
    return n == 0 ? 1u : (n + BLOCK - 1) / BLOCK;
}
inline int check_launch()
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : static_cast<int>(e);
}
inline ParamsView const& PV(B200ParamsView const* p)
{
    return *reinterpret_cast<ParamsView const*>(p);
}
inline StateView const& SV(B200StateView const* s)
{
    return *reinterpret_cast<StateView const*>(s);
}
//! Threads needed to cover the active list (host upper bound, capped by slots)
inline u32 active_hint(StateView const& s)
{
    return s.hint_active < s.num_slots ? s.hint_active : s.num_slots;
}
}  // namespace

extern "C" {
uint64_t b200_launch_count(void)
{
    return g_launches.load();
}

int b200_step_extend_from_primaries(B200StateView const* state,
                                    B200Primary const* d_primaries,
                                    uint32_t const* d_rank_in_event,
                                    uint32_t const* d_event_ids,
                                    uint32_t const* d_event_counts,
                                    uint32_t const* d_neutral_inclusive,
                                    uint32_t num_events,
                                    uint32_t n,
                                    cudaStream_t stream)
{
    if (n == 0)
        return 0;
    if (SV(state).ti_neutral_prefix && !d_neutral_inclusive)
        return B200_ERR_INVALID_ARGUMENT;
    k_extend_from_primaries<<<grid_for(n), BLOCK, 0, stream>>>(
        SV(state), d_primaries, d_rank_in_event, d_neutral_inclusive, n);
    k_primaries_finalize<<<grid_for(num_events), BLOCK, 0, stream>>>(
        SV(state), d_event_ids, d_event_counts, num_events, n);
    B2_COUNT(2);
    return check_launch();
}

int b200_step_initialize_tracks(B200ParamsView const* params,
                                B200StateView const* state,
                                cudaStream_t stream)
{
    StateView const& s = SV(state);
    u32 n = s.hint_new < s.num_slots ? s.hint_new : s.num_slots;
    if (n > 0)
    {
        k_initialize_tracks<<<grid_for(n), BLOCK, 0, stream>>>(PV(params), s);
        B2_COUNT(1);
    }
    k_initialize_finalize<<<1, 1, 0, stream>>>(s);
    B2_COUNT(1);
    return check_launch();
}

int b200_step_pre_step(B200ParamsView const* params, B200StateView const* state, cudaStream_t stream)
{
    StateView const& s = SV(state);
    k_pre_step<<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(PV(params), s);
    B2_COUNT(1);
    return check_launch();
}

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
        if (p.model.field.enabled)
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
        if (PV(params).model.field.enabled)
            k_along_step_charged<true, false><<<(nc + B2_ALONG_BLOCK - 1) / B2_ALONG_BLOCK, B2_ALONG_BLOCK, 0, stream>>>(PV(params), s);
        else
            k_along_step_charged<false, false><<<(nc + B2_ALONG_BLOCK - 1) / B2_ALONG_BLOCK, B2_ALONG_BLOCK, 0, stream>>>(PV(params), s);
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
        if (PV(params).model.field.enabled)
            k_along_step_charged<true, true><<<grid_for(nc), BLOCK, 0, stream>>>(PV(params), s);
        else
            k_along_step_charged<false, true><<<grid_for(nc), BLOCK, 0, stream>>>(PV(params), s);
        B2_COUNT(1);
    }
    if (nn > 0)
    {
        k_along_step_neutral<true><<<grid_for(nn), BLOCK, 0, stream>>>(PV(params), s);
        B2_COUNT(1);
    }
    return check_launch();
}

int b200_step_discrete_select(B200ParamsView const* params,
                              B200StateView const* state,
                              cudaStream_t stream)
{
    StateView const& s = SV(state);
    k_discrete_select<<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(PV(params), s);
    B2_COUNT(1);
    return check_launch();
}

int b200_step_interact(B200ParamsView const* params, B200StateView const* state, cudaStream_t stream)
{
    StateView const& s = SV(state);
    if (s.interact_list)
    {
        // lists built by this step's b200_step_discrete_select; segments are padded to
        // whole warps
        u32 const bound = active_hint(s) + 32 * PV(params).phys.num_models;
        k_interact_lists<<<grid_for(bound), BLOCK, 0, stream>>>(PV(params), s);
    }
    else
    {
        k_interact<<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(PV(params), s);
    }
    B2_COUNT(1);
    return check_launch();
}

int b200_step_boundary(B200ParamsView const* params, B200StateView const* state, cudaStream_t stream)
{
    StateView const& s = SV(state);
    k_boundary<<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(PV(params), s);
    B2_COUNT(1);
    return check_launch();
}

int b200_step_tracking_cut(B200ParamsView const* params,
                           B200StateView const* state,
                           cudaStream_t stream)
{
    StateView const& s = SV(state);
    k_tracking_cut<<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(PV(params), s);
    B2_COUNT(1);
    return check_launch();
}

int b200_step_tally(B200ParamsView const* params, B200StateView const* state, cudaStream_t stream)
{
    StateView const& s = SV(state);
    if (!s.calo_edep)
        return 0;
    k_tally<<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(PV(params), s, s.num_detectors);
    B2_COUNT(1);
    return check_launch();
}

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

int b200_step_post_tail(B200ParamsView const* params, B200StateView const* state, cudaStream_t stream)
{
    StateView const& s = SV(state);
    k_post_tail<<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(PV(params), s);
    B2_COUNT(1);
    return check_launch();
}

int b200_step_action_diagnostic(B200ParamsView const* params,
                                B200StateView const* state,
                                cudaStream_t stream)
{
    StateView const& s = SV(state);
    if (!s.diag_action_counts)
        return B200_ERR_INVALID_ARGUMENT;
    k_diagnostic<false><<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(
        s, PV(params).particle.num_particles);
    B2_COUNT(1);
    return check_launch();
}

int b200_step_step_diagnostic(B200ParamsView const* params,
                              B200StateView const* state,
                              cudaStream_t stream)
{
    StateView const& s = SV(state);
    if (!s.diag_step_counts)
        return B200_ERR_INVALID_ARGUMENT;
    k_diagnostic<true><<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(
        s, PV(params).particle.num_particles);
    B2_COUNT(1);
    return check_launch();
}

int b200_step_extend_from_secondaries(B200ParamsView const* params,
                                      B200StateView const* state,
                                      cudaStream_t stream)
{
    StateView const& s = SV(state);
    if (s.slot_begin % BLOCK != 0 || s.slot_begin > s.num_slots)
        return B200_ERR_INVALID_ARGUMENT;
    unsigned nb = grid_for(s.num_slots - s.slot_begin);
    k_end_pass1<<<nb, BLOCK, 0, stream>>>(PV(params), s);
    k_end_pass2<<<6, 1024, 0, stream>>>(s, nb);
    k_end_pass3<<<nb, BLOCK, 0, stream>>>(PV(params), s);
    B2_COUNT(3);
    return check_launch();
}

int b200_reseed(B200ParamsView const* params,
                B200StateView const* state,
                uint64_t event_id,
                cudaStream_t stream)
{
    StateView const& s = SV(state);
    k_reseed<<<grid_for(s.num_slots), BLOCK, 0, stream>>>(PV(params), s, event_id);
    B2_COUNT(1);
    return check_launch();
}

int b200_reset_generated(B200StateView const* state, cudaStream_t stream)
{
    k_reset_generated<<<1, 1, 0, stream>>>(SV(state));
    B2_COUNT(1);
    return check_launch();
}

int b200_kill_active(B200ParamsView const* params, B200StateView const* state, cudaStream_t stream)
{
    StateView const& s = SV(state);
    k_kill_active<<<grid_for(s.num_slots), BLOCK, 0, stream>>>(PV(params), s);
    B2_COUNT(1);
    return check_launch();
}
}  // extern "C"
