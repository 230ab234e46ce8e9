//---------------------------------------------------------------------------//
// Device functions of the step actions (shared by kernels.cu and tail.cu).
//
// One function per step action of the reference's loop
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
#pragma once

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

//! Nanoseconds of the device-wide timer
B2_D u64 global_timer_ns()
{
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
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

//! L2 prefetch of a slot's RNG state (layout: Rng::load)
B2_D void prefetch_rng(StateView const& s, u32 slot)
{
#if B2_RNG_PACKED
    prefetch_l2(s.rng + 4 * size_t(slot));
    prefetch_l2(s.rng + 4 * size_t(s.num_slots) + 2 * size_t(slot));
#else
#pragma unroll
    for (int k = 0; k < 6; ++k)
        prefetch_l2(s.rng + size_t(k) * s.num_slots + slot);
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
#if B2_POSDIR_PACKED
    prefetch_l2(s.geo_pos + 2 * size_t(slot));
    prefetch_l2(s.geo_pos + 2 * size_t(ng) + slot);
    prefetch_l2(s.geo_dir + 2 * size_t(slot));
    prefetch_l2(s.geo_dir + 2 * size_t(ng) + slot);
#else
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        prefetch_l2(s.geo_pos + k * ng + slot);
        prefetch_l2(s.geo_dir + k * ng + slot);
    }
#endif
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
        prefetch_rng(s, slot);
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
    prefetch_rng(s, slot);
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


//---------------------------------------------------------------------------//
// start: initialize tracks in vacant slots
// (track/detail/InitTracksExecutor.hh:71-175)
//---------------------------------------------------------------------------//
//! Start the tid-th of the num_new = min(num_init, num_vac) tracks that start this step
//! `vacancy_at(k)` returns the k-th vacant slot in slot order: the sorted vacancy array of
//! the per-action path, or the per-run prefix search of the device-resident loop (tail.cu)
template<class VacancyAt>
B2_D void initialize_track(ParamsView const& p,
                           StateView const& s,
                           u32 tid,
                           u32 num_init,
                           u32 num_vac,
                           u32 num_new,
                           VacancyAt&& vacancy_at)
{
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
            slot = vacancy_at(neutral_rank);
        else
            slot = vacancy_at(num_vac - (num_new - num_neutral) + ((ti - first) - neutral_rank));
    }
    else
    {
        slot = vacancy_at(num_vac - tid - 1);
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


//! Counter bookkeeping after the starts of a step (one thread)
B2_D void initialize_finalize(StateView const& s, u32 num_init, u32 num_vac)
{
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
        if (s.hit_pre)
        {
            // StepGatherExecutor<pre> (user/detail/StepGatherExecutor.hh:117-149): the
            // pre-step point of every track that steps
            size_t const n = s.num_slots;
            Real3 const pos = geo.pos(), dir = geo.dir();
            s.hit_pre[slot] = s.time[slot];
            for (int k = 0; k < 3; ++k)
            {
                s.hit_pre[(1 + k) * n + slot] = pos[k];
                s.hit_pre[(4 + k) * n + slot] = dir[k];
            }
            s.hit_pre[7 * n + slot] = s.energy[slot];
        }
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

// The same charged along-step as four phase kernels (see along_step.cuh). Used for large
// iterations; small ones stay fused, where launch latency matters more than occupancy.


// Besides selecting, the launch sorts the interacting tracks BY MODEL into per-model slot
// lists (block-aggregated appends), so that the interaction kernel runs warps in which
// every lane executes the same interactor. Launched over all active tracks, only ~4 of
// 32 lanes were active per instruction (ncu: profiles/README_r01.md).

//---------------------------------------------------------------------------//
// post: every EM model (dispatch on the selected action id)
//---------------------------------------------------------------------------//
template<bool EXTRA>
B2_D void do_interact(ParamsView const& p, StateView const& s, u32 slot)
{
    if (s.status[slot] != ST_ALIVE)
        return;
    u32 action = s.post_step_action[slot];
    if (action < p.phys.model_to_action || action >= p.phys.model_to_action + p.phys.num_models)
        return;
    Rng rng;
    rng.load(s, slot);
    run_interaction<EXTRA>(p, s, slot, action, rng);
    rng.store(s, slot);
}


//! Interactions over the per-model lists built by k_discrete_select: thread t works on
//! the t-th interacting track in model order

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


// user_post: tallies (user/detail/SimpleCaloExecutor.hh:48-67)
// Per-detector sums are first accumulated in shared memory (one copy per block) and
// flushed with one global atomic per touched bin, instead of one contended global
// atomic per depositing track.
constexpr u32 TALLY_SMEM_BINS = 1024;


//---------------------------------------------------------------------------//
// diagnostics
// post: tally the post-step action of every track that took this step
//   (user/detail/ActionDiagnosticExecutor.hh:30-65)
// user_post: tally the number of steps of every track killed this step
//   (user/detail/StepDiagnosticExecutor.hh:28-60)
// Counts are first gathered per block in shared memory (a handful of bins are hot).
//---------------------------------------------------------------------------//
constexpr u32 DIAG_SMEM_BINS = 1024;


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
//! COOP: called by all 32 lanes of a warp with the same `tid` (warp-cooperative navigation,
//! orange.cuh); everything but the face searches runs redundantly, tallies by lane 0 only.
template<int FIELD, bool COOP = false>
B2_D void step_fused_slot(ParamsView const& p, StateView const& s, u32 slot, bool charged)
{
    bool const tally_lane = !COOP || (threadIdx.x & 31u) == 0;
    // pre
    do_pre_step(p, s, slot);
    // along
    if (s.status[slot] == ST_ALIVE)
    {
        if (charged)
            along_step<true, FIELD, COOP>(p, s, slot);
        else
            along_step<false, false, COOP>(p, s, slot);
    }
    // pre_post, post
    do_discrete_select(p, s, slot);
    do_interact<false>(p, s, slot);
    do_boundary(p, s, slot);
    do_tracking_cut(p, s, slot);
    u8 const status = s.status[slot];
    if (tally_lane && s.diag_action_counts && status != ST_INACTIVE)
    {
        atomicAdd(&s.diag_action_counts[s.particle_id[slot] * s.diag_action_bins
                                        + s.post_step_action[slot]],
                  1u);
    }
    // user_post
    if (tally_lane && s.calo_edep && status != ST_INACTIVE)
    {
        real edep = s.energy_deposition[slot];
        if (edep != 0)
        {
            u32 det = s.calo_detector_of_volume[s.pre_volume[slot]];
            if (det != INVALID)
                atomicAdd(&s.calo_edep[det], edep);
        }
    }
    if (tally_lane && s.diag_step_counts && status == ST_KILLED)
    {
        u32 const nb = s.diag_step_bins;
        u32 const n = s.num_steps[slot];
        atomicAdd(&s.diag_step_counts[s.particle_id[slot] * nb + (n < nb - 1 ? n : nb - 1)], 1u);
    }
}

template<int FIELD>
B2_D void step_fused_track(ParamsView const& p, StateView const& s, u32 tid)
{
    u32 const slot = active_slot(s, tid);
    if (slot == INVALID)
        return;
    step_fused_slot<FIELD, false>(p, s, slot, tid < s.counters[CTR_NUM_CHARGED]);
}

//---------------------------------------------------------------------------//
// Warp-cooperative step of ONE track (fewer tracks than warps: csrc/tail.cu).
//
// All 32 lanes of the warp run the whole step of the same track and share the per-face
// work of its distance and safety searches (coop_find_next_step, orange.cuh). The lanes
// are NOT guaranteed to stay in lock-step between those searches, and the step is full of
// read-modify-write sequences on the track's state (s.num_steps[slot] += 1, the boundary
// crossing, the RNG words): run on the shared state, a lane that is a few instructions
// ahead would feed its results to the lanes behind it (measured: tracks crossed a
// boundary twice). So every lane works on a PRIVATE one-slot copy of the track's state
// (ShadowSlot, in local memory; the same StateView code runs on it with num_slots = 1,
// slot = 0) and lane 0 writes the result back. Tallies are atomics on the real arrays, by
// lane 0 only.
//---------------------------------------------------------------------------//
constexpr u32 SHADOW_MAX_DEPTH = 8;
constexpr u32 SHADOW_MAX_PROCESSES = 8;

struct ShadowSlot
{
    real time, step_length, energy, geo_next_step;
    real geo_pos[3 * SHADOW_MAX_DEPTH], geo_dir[3 * SHADOW_MAX_DEPTH];
    real interaction_mfp, macro_xs, energy_deposition, dedx_range;
    real msc_range[3], msc_true_path, msc_geom_path, msc_alpha;
    real per_process_xs[SHADOW_MAX_PROCESSES];
    real sec_energy[MAX_SECONDARIES], sec_dir[MAX_SECONDARIES * 3];
    u32 track_id, parent_id, event_id, num_steps, num_looping_steps;
    u32 post_step_action, along_step_action, particle_id, material_id;
    u32 geo_level, geo_surface_level, geo_surf, geo_next_level, geo_next_surf;
    u32 geo_vol[SHADOW_MAX_DEPTH], geo_univ[SHADOW_MAX_DEPTH];
    u32 element, sec_particle[MAX_SECONDARIES], pre_volume;
    alignas(16) u32 rng[8];  // six words, laid out as Rng::load expects for num_slots = 1
    u8 status, geo_sense, geo_boundary, geo_next_sense, msc_is_displaced;
};

B2_D bool shadow_supported(StateView const& s)
{
    // the copy loops below assume one column per component (stride num_slots); the RNG
    // words go through Rng::load / store
    if (B2_POSDIR_PACKED)
        return false;
    return s.max_depth <= SHADOW_MAX_DEPTH && s.max_processes <= SHADOW_MAX_PROCESSES;
}

//! The per-slot fields the step reads or writes: X(field, count) with `count` entries of
//! stride num_slots (level-major / component-major columns)
#define B2_SHADOW_FIELDS(X, D, P)                                                          \
    X(time, 1) X(step_length, 1) X(energy, 1) X(geo_next_step, 1) X(geo_pos, 3 * (D))      \
    X(geo_dir, 3 * (D)) X(interaction_mfp, 1) X(macro_xs, 1) X(energy_deposition, 1)       \
    X(dedx_range, 1) X(msc_range, 3) X(msc_true_path, 1) X(msc_geom_path, 1)               \
    X(msc_alpha, 1) X(per_process_xs, (P)) X(sec_energy, MAX_SECONDARIES)                  \
    X(sec_dir, MAX_SECONDARIES * 3) X(track_id, 1) X(parent_id, 1) X(event_id, 1)          \
    X(num_steps, 1) X(num_looping_steps, 1) X(post_step_action, 1)                         \
    X(along_step_action, 1) X(particle_id, 1) X(material_id, 1) X(geo_level, 1)            \
    X(geo_surface_level, 1) X(geo_surf, 1) X(geo_next_level, 1) X(geo_next_surf, 1)        \
    X(geo_vol, (D)) X(geo_univ, (D)) X(element, 1) X(sec_particle, MAX_SECONDARIES)        \
    X(status, 1) X(geo_sense, 1) X(geo_boundary, 1) X(geo_next_sense, 1)         \
    X(msc_is_displaced, 1)

template<class T>
B2_D T* shadow_ptr(T& x)
{
    return &x;
}
template<class T, size_t N>
B2_D T* shadow_ptr(T (&x)[N])
{
    return x;
}

//! View of the private copy: per-slot columns point into `buf`, everything else
//! (counters, tallies, lists) is the real thing
B2_D StateView shadow_view(StateView const& s, ShadowSlot& buf)
{
    StateView v = s;
    v.num_slots = 1;
#define B2_X(field, count) v.field = shadow_ptr(buf.field);
    B2_SHADOW_FIELDS(B2_X, 0, 0)
#undef B2_X
    v.rng = buf.rng;
    v.pre_volume = s.pre_volume ? &buf.pre_volume : nullptr;
    return v;
}

B2_D void shadow_load(StateView const& s, u32 slot, ShadowSlot& buf)
{
    u32 const n = s.num_slots;
    u32 const D = s.max_depth, P = s.max_processes;
#define B2_X(field, count)                          \
    for (u32 i = 0; i < (count); ++i)               \
        shadow_ptr(buf.field)[i] = s.field[size_t(i) * n + slot];
    B2_SHADOW_FIELDS(B2_X, D, P)
#undef B2_X
    {
        // with one slot both RNG layouts are the six words in order
        Rng r;
        r.load(s, slot);
        for (int k = 0; k < 5; ++k)
            buf.rng[k] = r.x[k];
        buf.rng[5] = r.d;
    }
    if (s.pre_volume)
        buf.pre_volume = s.pre_volume[slot];
}

B2_D void shadow_store(StateView const& s, u32 slot, ShadowSlot const& buf)
{
    u32 const n = s.num_slots;
    u32 const D = s.max_depth, P = s.max_processes;
#define B2_X(field, count)                          \
    for (u32 i = 0; i < (count); ++i)               \
        s.field[size_t(i) * n + slot] = shadow_ptr(const_cast<ShadowSlot&>(buf).field)[i];
    B2_SHADOW_FIELDS(B2_X, D, P)
#undef B2_X
    {
        Rng r;
        for (int k = 0; k < 5; ++k)
            r.x[k] = buf.rng[k];
        r.d = buf.rng[5];
        r.store(s, slot);
    }
    if (s.pre_volume)
        s.pre_volume[slot] = buf.pre_volume;
}

//! Whole step of the tid-th active track by ONE thread on a private copy of its state
//! (experiment: is it the private copy or the shared searches that makes the cooperative
//! step faster? see profiles/README_r02.md)
template<int FIELD>
B2_D void step_fused_track_shadow(ParamsView const& p, StateView const& s, u32 tid)
{
    u32 const slot = active_slot(s, tid);
    if (slot == INVALID)
        return;
    bool const charged = tid < s.counters[CTR_NUM_CHARGED];
    ShadowSlot buf;
    shadow_load(s, slot, buf);
    StateView const v = shadow_view(s, buf);
    step_fused_slot<FIELD, false>(p, v, 0, charged);
    shadow_store(s, slot, buf);
}

//! Whole step of the tid-th active track by the calling WARP (all 32 lanes)
template<int FIELD>
B2_D void step_fused_track_coop(ParamsView const& p, StateView const& s, u32 tid)
{
    u32 const slot = active_slot(s, tid);
    if (slot == INVALID)
        return;
    bool const charged = tid < s.counters[CTR_NUM_CHARGED];
    ShadowSlot buf;
    shadow_load(s, slot, buf);
    StateView const v = shadow_view(s, buf);
    step_fused_slot<FIELD, true>(p, v, 0, charged);
    __syncwarp();
    if ((threadIdx.x & 31u) == 0)
        shadow_store(s, slot, buf);
}


//---------------------------------------------------------------------------//
// The tail of a step in one launch for LARGE iterations: boundary crossing, tracking
// cut, action diagnostic (order post), calorimeter tally and step diagnostic (order
// user_post) are consecutive in the action sequence, each touches only its own slot, and
// all but the boundary crossing are a few instructions per track: as separate launches
// they each re-read the slot lists and the status and post-step action of every track.
// Tallies go through per-block shared-memory bins as in k_tally / k_diagnostic.
//---------------------------------------------------------------------------//

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

//! Pass 1 for the 128 slots of (virtual) block `vb` of `nb`, starting at slot_begin
B2_D void end_pass1_block(ParamsView const& p, StateView const& s, u32 slot_begin, u32 vb, u32 nb)
{
    u32 slot = slot_begin + vb * BLOCK + threadIdx.x;
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
        s.block_scratch[vb] = ta & 0x3ffu;
        s.block_scratch[nb + vb] = (ta >> 10) & 0x3ffu;
        s.block_scratch[2 * nb + vb] = (ta >> 20) & 0x3ffu;
        s.block_scratch[3 * nb + vb] = tb & 0x3ffu;
        s.block_scratch[4 * nb + vb] = (tb >> 10) & 0x3ffu;
        s.block_scratch[5 * nb + vb] = (tb >> 20) & 0x3ffu;
    }
}


//! Global counters of the step from the six scan totals (one thread, after the scans)
B2_D void end_pass2_finish(StateView const& s, u32 slot_begin)
{
    u32 carry[6];
    for (int k = 0; k < 6; ++k)
        carry[k] = reinterpret_cast<u32 volatile*>(s.counters)[CTR_SCAN_TOTALS + k];
    // slots below slot_begin are all vacant
    u32 num_vac = carry[0] + slot_begin;
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
}


//! The reference's pre-step resets the step limit of inactive slots
//! (PreStepExecutor.hh:47-57); inactive slots are never visited by the dense kernels
//! here, so it is done once, when the slot is first seen inactive
B2_D void reset_inactive_slot(StateView const& s, u32 slot)
{
    if (s.post_step_action[slot] != INVALID || s.along_step_action[slot] != INVALID)
    {
        s.step_length[slot] = real_inf();
        s.post_step_action[slot] = INVALID;
        s.along_step_action[slot] = INVALID;
    }
}

//! End of step for one slot that was NOT inactive during the step: dense-list entry,
//! secondaries -> initializers (or in-place reuse of a dead parent's slot), killed ->
//! inactive. The offsets are the slot's exclusive prefix sums in slot order. Returns whether
//! the slot became inactive (its step limit is reset when it is next seen inactive).
B2_D bool end_slot_active(ParamsView const& p,
                          StateView const& s,
                          u32 slot,
                          SlotEnd const& e,
                          u32 chg_off,
                          u32 neu_off,
                          u32 sec_off,
                          u32 all_off,
                          u32 neutral_off,
                          u32 num_init,
                          u32 num_sec_total,
                          u32 event,
                          u32 parent_track,
                          real time)
{
    if (e.charged)
        s.track_slots[chg_off] = slot;
    if (e.neutral)
        s.track_slots[s.num_slots - 1 - neu_off] = slot;


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
    {
        s.status[slot] = ST_INACTIVE;
        return true;
    }
    return false;
}

//! Pass 3 for the 128 slots of (virtual) block `vb` of `nb`, starting at slot_begin
B2_D void end_pass3_block(ParamsView const& p, StateView const& s, u32 slot_begin, u32 vb, u32 nb)
{
    u32 slot = slot_begin + vb * BLOCK + threadIdx.x;
    // classification of pass 1 (the same launch sequence; nothing changed in between)
    SlotEnd e = unpack_class(slot < s.num_slots ? s.slot_class[slot] : u8(0));
    u32 ta, tb;
    // Everything that does not depend on the scans is loaded BEFORE their barriers, so
    // that these round trips overlap with the scans instead of queueing up behind them
    // (the kernel is latency bound: 38 long-scoreboard stall cycles per issue, ncu)
    u32 const block_vac = s.block_scratch[vb];
    u32 const block_chg = s.block_scratch[nb + vb];
    u32 const block_neu = s.block_scratch[2 * nb + vb];
    u32 const block_sec = s.block_scratch[3 * nb + vb];
    u32 const block_all = s.block_scratch[4 * nb + vb];
    u32 const block_neutral_sec = s.block_scratch[5 * nb + vb];
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
    u32 vac_off = slot_begin + (sa & 0x3ffu) + block_vac;
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
    if (e.inactive)
    {
        reset_inactive_slot(s, slot);
        return;
    }
    end_slot_active(p,
                    s,
                    slot,
                    e,
                    chg_off,
                    neu_off,
                    sec_off,
                    all_off,
                    neutral_off,
                    num_init,
                    num_sec_total,
                    event,
                    parent_track,
                    time);
}

//---------------------------------------------------------------------------//
// reseed (random/RngReseed.cu:29-74)
//---------------------------------------------------------------------------//


}  // namespace b200
