//---------------------------------------------------------------------------//
// SortTracksAction: TrackOrder::reindex_* (reference track/SortTracksAction.cc:46-131,
// track/detail/TrackSortUtils.cu:66-207).
//
// The reference keeps `track_slots`, a permutation of all track slots, sorted by a small
// integer key (the along-step action id, the step-limit action id, the particle id, or
// active/inactive) with thrust::sort_by_key / thrust::partition (radix sort over 32-bit
// keys: several passes over keys and values, temporary storage from a pool), then finds the
// first thread of every action with one more kernel plus a device-to-host copy and a stream
// synchronisation (count_tracks_per_action + backfill_action_count).
//
// Keys here are tiny (a few dozen actions at most), so this is ONE counting sort:
//   k_sort_count    per block, a shared-memory histogram of the keys of its slots
//   k_sort_scan     one block: exclusive scan of the [key][block] count matrix in key-major
//                   order, which is at once the global offset of every key (the reference's
//                   action_thread_offsets, already back-filled) and every block's base
//   k_sort_scatter  every slot goes to base[key][block] + its rank inside the block; ranks
//                   come from warp match/ballot, so equal keys stay in slot order (stable)
// Two reads of the key column and one write of the permutation; no temporary storage, no
// host synchronisation. The permutation is a member of the same class as the reference's
// (its sorts are not stable): same key sequence, same set of slots per key
// (tests/test_gpu_track_order.py).
//---------------------------------------------------------------------------//
#include "launch_util.cuh"
#include "orange.cuh"

namespace b200
{
constexpr u32 SORT_BLOCK = 256;
constexpr u32 SORT_MAX_KEYS = 255;  // + 1 "no key" bin; shared-memory histograms

//! Sort key of a slot for the given single-key order; `nkeys` = the "no key" bin
B2_D u32 sort_key(StateView const& s, u32 slot, u32 order, u32 nkeys)
{
    u32 key;
    switch (order)
    {
        case SORT_KEY_HITS: {
            // StepGatherExecutor<post> (user/detail/StepGatherExecutor.hh:60-115): a step is
            // recorded if the track stepped, started the step in a sensitive volume and
            // (optionally) deposited energy
            if (s.status[slot] == ST_INACTIVE)
                return 1u;
            u32 const vol = s.pre_volume[slot];
            if (vol == INVALID || s.hit_detector_of_volume[vol] == INVALID)
                return 1u;
            if (s.hit_nonzero_edep && s.energy_deposition[slot] == 0)
                return 1u;
            return 0u;
        }
        case ORDER_REINDEX_STATUS:
            // thrust::partition(IsNotInactive): active slots first
            return s.status[slot] != ST_INACTIVE ? 0u : 1u;
        case ORDER_REINDEX_PARTICLE_TYPE:
            // every slot, as the reference: an inactive slot sorts by the particle that last
            // held it (never used: invalid id, last)
            key = s.particle_id[slot];
            break;
        case ORDER_REINDEX_ALONG_STEP_ACTION:
            // The action sorts run after pre-step, which has reset the actions of every
            // inactive slot in the reference (PreStepExecutor.hh:47-57); here a slot is
            // reset when an end-of-step pass first sees it inactive (reset_inactive_slot)
            if (s.status[slot] == ST_INACTIVE)
                return nkeys;
            key = s.along_step_action[slot];
            break;
        default:
            if (s.status[slot] == ST_INACTIVE)
                return nkeys;
            key = s.post_step_action[slot];
            break;
    }
    // invalid ids (inactive slots, reset by pre-step) are the largest key in the reference
    return key < nkeys ? key : nkeys;
}

__global__ void __launch_bounds__(SORT_BLOCK)
    k_sort_count(B2_GRID_CONSTANT StateView const s, u32 order, u32 nkeys)
{
    __shared__ u32 hist[SORT_MAX_KEYS + 1];
    for (u32 k = threadIdx.x; k <= nkeys; k += SORT_BLOCK)
        hist[k] = 0;
    __syncthreads();
    u32 const slot = blockIdx.x * SORT_BLOCK + threadIdx.x;
    bool const have = slot < s.num_slots;
    u32 const key = have ? sort_key(s, slot, order, nkeys) : INVALID;
    // one atomic per distinct key of the warp
    unsigned const peers = __match_any_sync(0xffffffffu, key);
    if (have && (threadIdx.x & 31u) == u32(__ffs(peers) - 1))
        atomicAdd(&hist[key], u32(__popc(peers)));
    __syncthreads();
    for (u32 k = threadIdx.x; k <= nkeys; k += SORT_BLOCK)
        s.sort_block_counts[size_t(k) * gridDim.x + blockIdx.x] = hist[k];
}

__global__ void __launch_bounds__(1024)
    k_sort_scan(B2_GRID_CONSTANT StateView const s, u32 nkeys, u32 nblocks)
{
    // exclusive scan of counts[key][block] in key-major order (in place)
    u32 const total = (nkeys + 1) * nblocks;
    u32 const per = (total + 1023u) / 1024u;
    u32 const begin = threadIdx.x * per;
    u32 const end = begin + per < total ? begin + per : total;
    u32 local = 0;
    for (u32 i = begin; i < end; ++i)
        local += s.sort_block_counts[i];
    u32 sum;
    u32 run = block_exclusive_scan<1024, u32>(local, &sum);
    for (u32 i = begin; i < end; ++i)
    {
        u32 const v = s.sort_block_counts[i];
        s.sort_block_counts[i] = run;
        run += v;
    }
    __syncthreads();
    // first index of every key (+ the end): action_thread_offsets, back-filled
    for (u32 k = threadIdx.x; k <= nkeys + 1; k += 1024)
        s.sort_offsets[k] = k <= nkeys ? s.sort_block_counts[size_t(k) * nblocks] : sum;
}

__global__ void __launch_bounds__(SORT_BLOCK)
    k_sort_scatter(B2_GRID_CONSTANT StateView const s, u32 order, u32 nkeys)
{
    constexpr u32 NWARPS = SORT_BLOCK / 32;
    // per (warp, key): slots with that key in earlier warps of the block
    __shared__ u32 warp_count[NWARPS][SORT_MAX_KEYS + 1];
    for (u32 i = threadIdx.x; i < NWARPS * (SORT_MAX_KEYS + 1); i += SORT_BLOCK)
        (&warp_count[0][0])[i] = 0;
    __syncthreads();
    u32 const slot = blockIdx.x * SORT_BLOCK + threadIdx.x;
    u32 const warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    bool const have = slot < s.num_slots;
    u32 const key = have ? sort_key(s, slot, order, nkeys) : INVALID;
    unsigned const peers = __match_any_sync(0xffffffffu, key);
    u32 const rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (have && lane == u32(__ffs(peers) - 1))
        warp_count[warp][key] = __popc(peers);
    __syncthreads();
    if (have)
    {
        u32 before = 0;
        for (u32 w = 0; w < warp; ++w)
            before += warp_count[w][key];
        u32 const base = s.sort_block_counts[size_t(key) * gridDim.x + blockIdx.x];
        s.sort_slots[base + before + rank_in_warp] = slot;
    }
}
//! Compact hit records in slot order: record i is slot sort_slots[i], i < sort_offsets[1]
//! (the reference: thrust::copy_if of the slots with a detector + gather_step_kernel,
//! user/DetectorSteps.cu:36-93, 150-200)
__global__ void __launch_bounds__(SORT_BLOCK)
    k_hits_gather(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s)
{
    u32 const count = s.sort_offsets[1];
    u32 const i = blockIdx.x * SORT_BLOCK + threadIdx.x;
    if (i == 0)
        *s.hit_count = count;
    if (i >= count)
        return;
    size_t const n = s.num_slots;
    u32 const slot = s.sort_slots[i];
    s.hit_u32[i] = s.hit_detector_of_volume[s.pre_volume[slot]];
    s.hit_u32[n + i] = s.track_id[slot];
    s.hit_u32[2 * n + i] = s.event_id[slot];
    s.hit_u32[3 * n + i] = s.parent_id[slot];
    s.hit_u32[4 * n + i] = s.num_steps[slot];
    s.hit_u32[5 * n + i] = s.particle_id[slot];
    s.hit_f64[i] = s.step_length[slot];
    s.hit_f64[n + i] = s.energy_deposition[slot];
    for (u32 k = 0; k < 8; ++k)
        s.hit_f64[(2 + k) * n + i] = s.hit_pre[k * n + slot];
    GeoTrack geo(p, s, slot);
    Real3 const pos = geo.pos(), dir = geo.dir();
    s.hit_f64[10 * n + i] = s.time[slot];
    for (u32 k = 0; k < 3; ++k)
    {
        s.hit_f64[(11 + k) * n + i] = pos[k];
        s.hit_f64[(14 + k) * n + i] = dir[k];
    }
    s.hit_f64[17 * n + i] = s.energy[slot];
}
}  // namespace b200

using namespace b200;

//! Step/hit output of the step that just ran (order user_post, after the tallies): stable
//! partition of the slots by "is a hit", then the gather. No-op without sensitive volumes.
extern "C" int b200_step_gather_hits(B200ParamsView const* params,
                                     B200StateView const* state,
                                     cudaStream_t stream)
{
    StateView const& s = SV(state);
    if (!s.hit_pre)
        return 0;
    if (!s.sort_slots || !s.sort_offsets || !s.sort_block_counts || !s.pre_volume)
        return B200_ERR_INVALID_ARGUMENT;
    u32 const nblocks = (s.num_slots + SORT_BLOCK - 1) / SORT_BLOCK;
    k_sort_count<<<nblocks, SORT_BLOCK, 0, stream>>>(s, SORT_KEY_HITS, 1);
    k_sort_scan<<<1, 1024, 0, stream>>>(s, 1, nblocks);
    k_sort_scatter<<<nblocks, SORT_BLOCK, 0, stream>>>(s, SORT_KEY_HITS, 1);
    k_hits_gather<<<nblocks, SORT_BLOCK, 0, stream>>>(PV(params), s);
    B2_COUNT(4);
    return check_launch();
}

extern "C" int b200_step_sort_tracks(B200ParamsView const* params,
                                     B200StateView const* state,
                                     uint32_t track_order,
                                     cudaStream_t stream)
{
    StateView const& s = SV(state);
    if (!s.sort_slots || !s.sort_offsets || !s.sort_block_counts)
        return B200_ERR_INVALID_ARGUMENT;
    u32 nkeys;
    switch (track_order)
    {
        case ORDER_REINDEX_STATUS: nkeys = 1; break;  // 0 active, 1 inactive
        case ORDER_REINDEX_PARTICLE_TYPE: nkeys = PV(params).particle.num_particles; break;
        case ORDER_REINDEX_ALONG_STEP_ACTION:
        case ORDER_REINDEX_STEP_LIMIT_ACTION: nkeys = s.num_sort_keys; break;
        default: return B200_ERR_INVALID_ARGUMENT;
    }
    if (nkeys > SORT_MAX_KEYS || nkeys > s.num_sort_keys)
        return B200_ERR_INVALID_ARGUMENT;
    u32 const nblocks = (s.num_slots + SORT_BLOCK - 1) / SORT_BLOCK;
    k_sort_count<<<nblocks, SORT_BLOCK, 0, stream>>>(s, track_order, nkeys);
    k_sort_scan<<<1, 1024, 0, stream>>>(s, nkeys, nblocks);
    k_sort_scatter<<<nblocks, SORT_BLOCK, 0, stream>>>(s, track_order, nkeys);
    B2_COUNT(3);
    return check_launch();
}
