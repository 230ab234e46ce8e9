//---------------------------------------------------------------------------//
// Step-action kernels and their C-ABI launchers: one kernel per step action of the
// reference's loop (SURVEY.md section 2.3;
// /root/reference/src/celeritas/global/ActionSequence.cc:77-138). The per-track device code
// is in step_device.cuh.
//---------------------------------------------------------------------------//
#include <atomic>
#include <cstdio>

#include "launch_util.cuh"
#include "step_device.cuh"

namespace b200
{
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

__global__ void __launch_bounds__(BLOCK) k_initialize_tracks(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 tid = thread_id();
    u32 const num_init = s.counters[CTR_NUM_INITIALIZERS];
    u32 const num_vac = s.counters[CTR_NUM_VACANCIES];
    u32 const num_new = num_init < num_vac ? num_init : num_vac;
    if (tid < num_new)
        initialize_track(p, s, tid, num_init, num_vac, num_new, [&s](u32 k) { return s.vacancies[k]; });
}

__global__ void k_initialize_finalize(StateView s)
{
    initialize_finalize(s, s.counters[CTR_NUM_INITIALIZERS], s.counters[CTR_NUM_VACANCIES]);
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

#if B2_SMEM_GRIDS
//---------------------------------------------------------------------------//
// Experiment (north-star (c): "physics tables staged in shared memory or via TMA"): the
// pre-step as a PERSISTENT kernel whose blocks first copy the value-grid tables (values and
// node energies of every cross-section / energy-loss / range grid: 2 x 23 kB for the TestEm3
// problem) into shared memory with two 1-D bulk copies (cp.async.bulk, completion on an
// mbarrier) and then walk the active list; every XsCalculator / RangeCalculator lookup of
// do_pre_step then reads shared memory instead of L1/L2. Measured against the plain kernel
// in profiles/README_r02.md.
//---------------------------------------------------------------------------//
constexpr u32 STAGED_BLOCK = 256;

B2_D u32 smem_addr(void const* ptr)
{
    return static_cast<u32>(__cvta_generic_to_shared(ptr));
}

__global__ void __launch_bounds__(STAGED_BLOCK, 2)
    k_pre_step_staged(B2_GRID_CONSTANT ParamsView const p,
                      B2_GRID_CONSTANT StateView const s,
                      u32 reals_bytes,
                      u32 energy_bytes)
{
    extern __shared__ __align__(128) unsigned char staged[];
    __shared__ __align__(8) u64 bar;
    real* const sm_reals = reinterpret_cast<real*>(staged);
    real* const sm_energy = reinterpret_cast<real*>(staged + reals_bytes);
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)),
                     "r"(reals_bytes + energy_bytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_addr(sm_reals)),
            "l"(static_cast<real const*>(p.phys.reals)), "r"(reals_bytes), "r"(smem_addr(&bar))
            : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_addr(sm_energy)),
            "l"(static_cast<real const*>(p.phys.grid_energy)), "r"(energy_bytes),
            "r"(smem_addr(&bar))
            : "memory");
    }
    // wait for phase 0 of the barrier
    {
        u32 done = 0;
        while (!done)
        {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n"
                : "=r"(done)
                : "r"(smem_addr(&bar)), "r"(0u)
                : "memory");
        }
    }
    // the same problem description with the two table columns redirected to shared memory
    ParamsView q = p;
    q.phys.reals = RO<real>(sm_reals);
    q.phys.grid_energy = RO<real>(sm_energy);
    u32 const nact = s.counters[CTR_NUM_CHARGED] + s.counters[CTR_NUM_NEUTRAL];
    for (u32 tid = thread_id(); tid < nact; tid += gridDim.x * STAGED_BLOCK)
    {
        u32 const slot = active_slot(s, tid);
        if (slot != INVALID)
        {
            prefetch_pre_step_state(s, slot);
            do_pre_step(q, s, slot);
        }
    }
}
#endif




__global__ void __launch_bounds__(BLOCK) k_discrete_select(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    select_and_append(p, s, active_slot(s, thread_id()));
}

template<bool EXTRA>
__global__ void __launch_bounds__(BLOCK) k_interact(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 slot = active_slot(s, thread_id());
    if (slot != INVALID)
        do_interact<EXTRA>(p, s, slot);
}

template<bool EXTRA>
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
    do_interact<EXTRA>(p, s, s.interact_list[size_t(model) * s.num_slots + tid]);
}

__global__ void __launch_bounds__(BLOCK) k_boundary(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 slot = active_slot(s, thread_id());
    if (slot != INVALID)
        do_boundary(p, s, slot);
}

__global__ void __launch_bounds__(BLOCK) k_tracking_cut(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 slot = active_slot(s, thread_id());
    if (slot != INVALID)
        do_tracking_cut(p, s, slot);
}

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

__global__ void __launch_bounds__(BLOCK) k_end_pass1(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    end_pass1_block(p, s, s.slot_begin, blockIdx.x, gridDim.x);
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
        end_pass2_finish(s, s.slot_begin);
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
    end_pass3_block(p, s, s.slot_begin, blockIdx.x, gridDim.x);
}

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

namespace b200
{
std::atomic<uint64_t> g_launches{0};
}

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
#if B2_SMEM_GRIDS
    {
        // table sizes [bytes], rounded up to the 16-byte granularity of a bulk copy
        // (the loader pads both columns, CoreParams.cc)
        ParamsView const& p = PV(params);
        u32 const reals_bytes = (p.phys_reals_count * 8u + 15u) & ~15u;
        u32 const energy_bytes = (p.phys_energy_count * 8u + 15u) & ~15u;
        static int sms = 0;
        if (sms == 0)
        {
            int device = 0;
            cudaGetDevice(&device);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
            cudaFuncSetAttribute(k_pre_step_staged, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 100 * 1024);
        }
        if (reals_bytes + energy_bytes <= 100 * 1024)
        {
            u32 const want = (active_hint(s) + STAGED_BLOCK - 1) / STAGED_BLOCK;
            u32 const grid = want < u32(2 * sms) ? (want ? want : 1u) : u32(2 * sms);
            k_pre_step_staged<<<grid, STAGED_BLOCK, reals_bytes + energy_bytes, stream>>>(
                p, s, reals_bytes, energy_bytes);
            B2_COUNT(1);
            return check_launch();
        }
    }
#endif
    k_pre_step<<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(PV(params), s);
    B2_COUNT(1);
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
        if (PV(params).model.has_extra_models)
            k_interact_lists<true><<<grid_for(bound), BLOCK, 0, stream>>>(PV(params), s);
        else
            k_interact_lists<false><<<grid_for(bound), BLOCK, 0, stream>>>(PV(params), s);
    }
    else if (PV(params).model.has_extra_models)
    {
        k_interact<true><<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(PV(params), s);
    }
    else
    {
        k_interact<false><<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(PV(params), s);
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
    // StepGatherAction<post> (user/detail/StepGatherAction.cc:70-99): SimpleCalo's tally,
    // then the step/hit output for the callbacks that want whole steps
    if (s.calo_edep)
    {
        k_tally<<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(PV(params), s, s.num_detectors);
        B2_COUNT(1);
    }
    if (s.hit_pre)
        return b200_step_gather_hits(params, state, stream);
    return check_launch();
}


int b200_step_post_tail(B200ParamsView const* params, B200StateView const* state, cudaStream_t stream)
{
    StateView const& s = SV(state);
    k_post_tail<<<grid_for(active_hint(s)), BLOCK, 0, stream>>>(PV(params), s);
    B2_COUNT(1);
    if (s.hit_pre)
        return b200_step_gather_hits(params, state, stream);
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
