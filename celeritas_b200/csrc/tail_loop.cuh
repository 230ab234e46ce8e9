//---------------------------------------------------------------------------//
// Device-resident step loop for SMALL iterations (shower tails, looping tracks): the kernel.
// Instantiated once per translation unit (tail_field.cu, tail_nofield.cu: the two largest
// kernels of the library compile side by side); launchers in tail.cu.
//
// The reference's loop (app/celer-sim/Transporter.cc:84-179, global/Stepper.cc:124-140)
// returns to the host after every step iteration: the host reads the track counters,
// decides whether any track is left and launches the ~20 kernels of the next iteration.
// With a handful of tracks left that handshake, not the physics, is the cost of an
// iteration (TestEm3: 97 of 249 iterations of a pass hold fewer than 16 k tracks and take
// 146 us each whatever their size; CMS-scale: 619 of 707; profiles/README_r01.md).
//
// k_tail_loop runs up to `max_iterations` WHOLE step iterations in one cooperative launch:
//
//   A  start queued tracks in vacant slots            (InitializeTracksAction)
//   B  the whole step of every active track            (pre-step ... tallies, as k_step_fused)
//   C  end of step: locate alive, secondaries -> initializers, dense lists, counters
//      (ExtendFromSecondariesAction)
//
// separated by grid barriers, and stops when no track is left, when the next iteration would
// exceed `exit_active` tracks (the per-action kernels are the better tool there), on a device
// error, or after `max_iterations`. After every iteration the step's counters are written
// to a ring in mapped host memory: the Transporter's per-iteration arrays are the same as
// with one host round trip per iteration.
//
// End of step here costs O(num_slots / 32) word operations plus O(active tracks): every
// thread owns RUNS of 32 consecutive slots. A run whose 32 status bytes are all zero
// (inactive: in a tail that is nearly every run) contributes 32 vacancies and nothing else.
// The sorted vacancy ARRAY of the per-action path (2 MB per 2^19 slots, rewritten by its
// pass 3 every iteration) is not maintained inside the loop: vacancy k is found by a
// binary search over the per-run vacancy prefix (run_vac_prefix, 4 B per run) and a bit
// select in that run's vacancy mask; the array is rebuilt once when the loop exits.
// Results per slot are identical to the per-action path (same classification, same
// slot-ordered prefix sums): tests/test_gpu_tail.py.
//---------------------------------------------------------------------------//
#pragma once

#include <cooperative_groups.h>

#include "launch_util.cuh"
#include "step_device.cuh"
#include "tail_args.cuh"

namespace b200
{
namespace cg = cooperative_groups;

constexpr u32 RUN = 32;                  // slots per run (one thread)
constexpr u32 SUPER = BLOCK * RUN;       // slots per super-block (one block scan)
constexpr u32 TAIL_RING_WORDS = B200_TAIL_RING_WORDS;

// Resident blocks per SM asked of the loop's kernel. Without a request (or with 1) ptxas settles on
// 64 registers and spills 2.7 kB; with 2 it takes 134 (198 with field) and spills nothing
#ifndef B2_LOOP_MIN_BLOCKS
#    define B2_LOOP_MIN_BLOCKS 2
#endif
#if B2_LOOP_MIN_BLOCKS > 0
#    define B2_LOOP_BOUNDS __launch_bounds__(BLOCK, B2_LOOP_MIN_BLOCKS)
#else
#    define B2_LOOP_BOUNDS __launch_bounds__(BLOCK)
#endif


//! Per-run totals packed for the block scans: 16 bits per quantity
//! (a run holds <= 32 slots and <= 64 secondaries; a super-block 128 runs)
struct RunTotals
{
    u64 a;  // vacant | charged << 16 | neutral << 32
    u64 b;  // num_sec | num_sec_all << 16 | num_sec_neutral << 32
};

B2_D void accumulate(RunTotals& t, SlotEnd const& e)
{
    t.a += u64(e.is_vacant) | (u64(e.charged) << 16) | (u64(e.neutral) << 32);
    t.b += u64(e.num_sec) | (u64(e.num_sec_all) << 16) | (u64(e.num_sec_neutral) << 32);
}

//! Run (inside its super-block of BLOCK runs) that thread `t` of the block looks after in the
//! classification passes: lane j of warp w <-> run (BLOCK / 32) j + w
B2_D u32 interleaved_run(u32 t)
{
    return (t & 31u) * (BLOCK / 32u) + (t >> 5);
}

//! True if the 32 status bytes of a run are all ST_INACTIVE (= 0)
B2_D bool run_is_idle(StateView const& s, u32 first_slot)
{
    // num_slots is a multiple of RUN (checked by the launcher): aligned 32-byte load
    uint4 const* w = reinterpret_cast<uint4 const*>(s.status + first_slot);
    uint4 x = w[0], y = w[1];
    return (x.x | x.y | x.z | x.w | y.x | y.y | y.z | y.w) == 0u;
}

//! k-th vacancy in slot order (k < number of vacancies) from the per-run prefix
B2_D u32 tail_vacancy(StateView const& s, u32 k)
{
    u32 const nruns = s.num_slots / RUN;
    // last run r with run_vac_prefix[r] <= k
    u32 lo = 0, hi = nruns;  // invariant: prefix[lo] <= k < prefix[hi]
    while (hi - lo > 1)
    {
        u32 mid = (lo + hi) >> 1;
        if (s.run_vac_prefix[mid] <= k)
            lo = mid;
        else
            hi = mid;
    }
    u32 const j = k - s.run_vac_prefix[lo];
    // (j+1)-th set bit of the run's vacancy mask (written by the previous end of step: not
    // touched by the tracks that start concurrently)
    u32 const bit = __fns(s.run_vac_mask[lo], 0, j + 1);
    return bit < RUN ? lo * RUN + bit : INVALID;
}

enum TailExit : u32
{
    TAIL_EXIT_DONE = 0,        // no track alive, no initializer queued
    TAIL_EXIT_MAX_ITERATIONS,  // ran max_iterations
    TAIL_EXIT_TOO_MANY,        // next iteration exceeds exit_active tracks
    TAIL_EXIT_ERROR            // device error flag set (CTR_ERROR)
};

template<int FIELD>
__global__ void B2_LOOP_BOUNDS
    k_tail_loop(B2_GRID_CONSTANT ParamsView const p, B2_GRID_CONSTANT StateView const s, TailArgs const a)
{
    cg::grid_group grid = cg::this_grid();
    u32 const nthreads = gridDim.x * BLOCK;
    u32 const gtid = blockIdx.x * BLOCK + threadIdx.x;
    u32 const n = s.num_slots;
    u32 const nsuper = (n + SUPER - 1) / SUPER;
    u32 const nruns = n / RUN;

    u32 it = 0;
    u32 reason = TAIL_EXIT_MAX_ITERATIONS;
    for (; it < a.max_iterations; ++it)
    {
        //// Snapshot of the counters: stable since the last grid barrier ////
        u32 const num_init = s.counters[CTR_NUM_INITIALIZERS];
        u32 const num_vac = s.counters[CTR_NUM_VACANCIES];
        u32 const num_new = num_init < num_vac ? num_init : num_vac;
        u32 const device_error_in = s.counters[CTR_ERROR];
        u32 const alive = n - num_vac;
        if (device_error_in != 0)
        {
            reason = TAIL_EXIT_ERROR;
            break;
        }
        if (alive == 0 && num_init == 0)
        {
            reason = TAIL_EXIT_DONE;
            break;
        }
        if (alive + num_new > a.exit_active)
        {
            reason = TAIL_EXIT_TOO_MANY;
            break;
        }
        u32 const pending = it & 1u;
        u64 const t_begin = (gtid == 0) ? global_timer_ns() : 0;

        //// A: start tracks ////
        if (gtid == 0)
            s.counters[CTR_NUM_GENERATED] = 0;
        if (it == 0)
        {
            // the sorted vacancy array is valid on entry
            for (u32 t = gtid; t < num_new; t += nthreads)
                initialize_track(p, s, t, num_init, num_vac, num_new,
                                 [&s](u32 k) { return s.vacancies[k]; });
        }
        else
        {
            for (u32 t = gtid; t < num_new; t += nthreads)
                initialize_track(p, s, t, num_init, num_vac, num_new,
                                 [&s](u32 k) { return tail_vacancy(s, k); });
        }
        grid.sync();
        u64 const t_started = (gtid == 0) ? global_timer_ns() : 0;
        if (gtid == 0)
        {
            initialize_finalize(s, num_init, num_vac);
            s.tail_ctrl[pending ^ 1u] = 0;  // filled by this iteration's end of step
        }
        // Slots that became inactive in the previous iteration and were not taken by a track
        // that just started: reset their step limit now (what the reference's pre-step does
        // for every inactive slot). Nothing else touches an inactive slot during the step.
        if (it > 0)
        {
            u32 const count = s.tail_ctrl[pending];
            u32 const* list = s.tail_reset_list + size_t(pending) * n;
            for (u32 i = gtid; i < count; i += nthreads)
            {
                u32 const slot = list[i];
                if (s.status[slot] == ST_INACTIVE)
                    reset_inactive_slot(s, slot);
            }
        }

        //// B: the whole step of every active track ////
        {
            u32 const nact = s.counters[CTR_NUM_CHARGED] + s.counters[CTR_NUM_NEUTRAL];
            u32 const nwarps = nthreads / 32u;
            if (a.coop == 2 && shadow_supported(s))
            {
                // (experiment) one thread per track on a private copy of its state
                for (u32 t = gtid; t < nact; t += nthreads)
                    step_fused_track_shadow<FIELD>(p, s, t);
            }
            else if (a.coop == 1 && nact <= a.coop_max && shadow_supported(s))
            {
                // a handful of tracks: one WARP per track, the lanes share its distance
                // and safety searches (coop_find_next_step, orange.cuh)
                u32 const warp = gtid >> 5;
                if (warp < nact)
                    step_fused_track_coop<FIELD>(p, s, warp);
            }
            else
            {
                // tracks dealt out round-robin to all warps of the grid (see k_step_fused):
                // tracks that share a warp serialise on their different branches
                u32 const warp = gtid >> 5;
                for (u32 t = (gtid & 31u) * nwarps + warp; t < nact; t += nthreads)
                    step_fused_track<FIELD>(p, s, t);
            }
        }
        grid.sync();
        u64 const t_stepped = (gtid == 0) ? global_timer_ns() : 0;

        //// C1: classify, per-run totals, per-super-block scans ////
        for (u32 sb = blockIdx.x; sb < nsuper; sb += gridDim.x)
        {
            // Lane j of warp w looks after run 4 j + w of the super-block: neighbouring runs
            // (tracks cluster in slot space, by charge) belong to different warps
            u32 const run = sb * BLOCK + interleaved_run(threadIdx.x);
            u32 const first = run * RUN;
            RunTotals t{0, 0};
            bool const idle = run < nruns ? run_is_idle(s, first) : true;
            if (run < nruns && idle)
                t.a = RUN;  // 32 vacancies
            // Runs that held a track: the WARP classifies their 32 slots together (lane =
            // slot), one run after the other; in a tail nearly every run is idle
            constexpr unsigned full = 0xffffffffu;
            u32 const lane = threadIdx.x & 31u;
            unsigned busy = __ballot_sync(full, run < nruns && !idle);
            while (busy)
            {
                u32 const r = __ffs(busy) - 1;
                busy &= busy - 1;
                u32 const slot = __shfl_sync(full, run, r) * RUN + lane;
                SlotEnd const e = classify_slot(p, s, slot);
                s.slot_class[slot] = pack_class(e);
                auto count = [&](u32 v) {
                    return u64(__popc(__ballot_sync(full, v & 1u))
                               + 2 * __popc(__ballot_sync(full, v & 2u)));
                };
                u64 const vac = count(e.is_vacant), chg = count(e.charged), neu = count(e.neutral);
                u64 const ns = count(e.num_sec), nsa = count(e.num_sec_all),
                          nsn = count(e.num_sec_neutral);
                unsigned const live = __ballot_sync(full, !e.inactive);
                if (lane == r)
                {
                    t.a = vac | (chg << 16) | (neu << 32);
                    t.b = ns | (nsa << 16) | (nsn << 32);
                    if (live)
                        atomicMin(&s.counters[CTR_FIRST_BUSY_BLOCK],
                                  (slot - lane + (__ffs(live) - 1)) / BLOCK);
                }
            }
            // back to thread i <-> run i of the super-block for the scans
            __shared__ u64 totals_a[BLOCK], totals_b[BLOCK];
            totals_a[interleaved_run(threadIdx.x)] = t.a;
            totals_b[interleaved_run(threadIdx.x)] = t.b;
            __syncthreads();
            t.a = totals_a[threadIdx.x];
            t.b = totals_b[threadIdx.x];
            __syncthreads();
            u64 ta, tb;
            u64 const ea = block_exclusive_scan<BLOCK, u64>(t.a, &ta);
            u64 const eb = block_exclusive_scan<BLOCK, u64>(t.b, &tb);
            u32 const scan_run = sb * BLOCK + threadIdx.x;
            if (scan_run < nruns)
            {
                // exclusive prefix of the run INSIDE its super-block (completed in C3)
                s.run_scan[2 * size_t(scan_run)] = ea;
                s.run_scan[2 * size_t(scan_run) + 1] = eb;
            }
            if (threadIdx.x == 0)
            {
                s.block_scratch[sb] = u32(ta & 0xffffu);
                s.block_scratch[nsuper + sb] = u32((ta >> 16) & 0xffffu);
                s.block_scratch[2 * nsuper + sb] = u32((ta >> 32) & 0xffffu);
                s.block_scratch[3 * nsuper + sb] = u32(tb & 0xffffu);
                s.block_scratch[4 * nsuper + sb] = u32((tb >> 16) & 0xffffu);
                s.block_scratch[5 * nsuper + sb] = u32((tb >> 32) & 0xffffu);
            }
        }
        grid.sync();
        u64 const t_classified = (gtid == 0) ? global_timer_ns() : 0;

        //// C2: scan of the super-block totals, global counters (one block) ////
        if (blockIdx.x == 0)
        {
            u32 const per = (nsuper + BLOCK - 1) / BLOCK;
            for (u32 q = 0; q < 6; ++q)
            {
                u32* const scratch = s.block_scratch + q * nsuper;
                u32 const begin = threadIdx.x * per;
                u32 const end = begin + per < nsuper ? begin + per : nsuper;
                u32 local = 0;
                for (u32 i = begin; i < end; ++i)
                    local += scratch[i];
                u32 total;
                u32 runv = block_exclusive_scan<BLOCK, u32>(local, &total);
                for (u32 i = begin; i < end; ++i)
                {
                    u32 v = scratch[i];
                    scratch[i] = runv;
                    runv += v;
                }
                if (threadIdx.x == 0)
                    s.counters[CTR_SCAN_TOTALS + q] = total;
            }
            __syncthreads();
            if (threadIdx.x == 0)
            {
                end_pass2_finish(s, 0);
                s.tail_ctrl[pending] = 0;  // consumed during the step
                // this iteration's result for the host
                u32 volatile* out = a.ring + size_t(it) * TAIL_RING_WORDS;
                out[0] = s.counters[CTR_NUM_GENERATED];
                out[1] = s.counters[CTR_NUM_INITIALIZERS];
                out[2] = s.counters[CTR_NUM_VACANCIES];
                out[3] = s.counters[CTR_NUM_ACTIVE];
                out[4] = s.counters[CTR_NUM_SECONDARIES];
                out[5] = s.counters[CTR_NUM_ALIVE];
                out[6] = s.counters[CTR_NUM_CHARGED];
                out[7] = s.counters[CTR_NUM_NEUTRAL];
                out[8] = s.counters[CTR_FIRST_BUSY_BLOCK];
                out[9] = s.counters[CTR_ERROR];
                u64 const now = global_timer_ns();
                out[10] = u32(now);
                out[11] = u32(now >> 32);
                // phase durations [ns]: starts, step, classification, scan of totals
                out[12] = u32(t_started - t_begin);
                out[13] = u32(t_stepped - t_started);
                out[14] = u32(t_classified - t_stepped);
                out[15] = u32(now - t_classified);
            }
        }
        grid.sync();

        //// C3: dense lists, initializers from secondaries, per-run vacancy prefix ////
        {
            u32 const device_error = s.counters[CTR_ERROR];
            u32 const num_init_after = s.counters[CTR_NUM_INITIALIZERS];
            u32 const num_sec_total = s.counters[CTR_NUM_SECONDARIES];
            // every warp walks the runs of its 32 lanes that held a track, lane = slot (all
            // lanes of the warp take part, also those without a run of their own)
            for (u32 sb = blockIdx.x; sb < nsuper; sb += gridDim.x)
            {
                constexpr unsigned full = 0xffffffffu;
                u32 const lane = threadIdx.x & 31u;
                u32 const run = sb * BLOCK + interleaved_run(threadIdx.x);
                bool const have = run < nruns;
                u32 const first = run * RUN;
                bool const idle = have ? run_is_idle(s, first) : true;
                u64 const ea = have ? s.run_scan[2 * size_t(run)] : 0;
                u64 const eb = have ? s.run_scan[2 * size_t(run) + 1] : 0;
                if (have)
                {
                    s.run_vac_prefix[run] = u32(ea & 0xffffu) + s.block_scratch[sb];
                    if (run == nruns - 1)
                        s.run_vac_prefix[nruns] = s.counters[CTR_NUM_VACANCIES];
                }
                if (have && idle && device_error == 0)
                {
                    s.run_vac_mask[run] = 0xffffffffu;
                    if (it == 0)
                    {
                        // slots that the per-action path left killed -> inactive in its
                        // last iteration have not had their step limit reset yet
                        for (u32 i = 0; i < RUN; ++i)
                            reset_inactive_slot(s, first + i);
                    }
                }
                u32 const my_chg = u32((ea >> 16) & 0xffffu) + s.block_scratch[nsuper + sb];
                u32 const my_neu = u32((ea >> 32) & 0xffffu) + s.block_scratch[2 * nsuper + sb];
                u32 const my_sec = u32(eb & 0xffffu) + s.block_scratch[3 * nsuper + sb];
                u32 const my_all = u32((eb >> 16) & 0xffffu) + s.block_scratch[4 * nsuper + sb];
                u32 const my_nsn = u32((eb >> 32) & 0xffffu) + s.block_scratch[5 * nsuper + sb];
                unsigned busy = __ballot_sync(full, have && !idle && device_error == 0);
                while (busy)
                {
                    u32 const r = __ffs(busy) - 1;
                    busy &= busy - 1;
                    u32 const brun = __shfl_sync(full, run, r);
                    u32 const slot = brun * RUN + lane;
                    SlotEnd const e = unpack_class(s.slot_class[slot]);
                    unsigned const lower = (1u << lane) - 1u;
                    auto before = [&](u32 v) {
                        return u32(__popc(__ballot_sync(full, v & 1u) & lower)
                                   + 2 * __popc(__ballot_sync(full, v & 2u) & lower));
                    };
                    u32 const chg_off = __shfl_sync(full, my_chg, r) + before(e.charged);
                    u32 const neu_off = __shfl_sync(full, my_neu, r) + before(e.neutral);
                    u32 const sec_off = __shfl_sync(full, my_sec, r) + before(e.num_sec);
                    u32 const all_off = __shfl_sync(full, my_all, r) + before(e.num_sec_all);
                    u32 const neutral_off
                        = __shfl_sync(full, my_nsn, r) + before(e.num_sec_neutral);
                    unsigned const vac_mask = __ballot_sync(full, e.is_vacant);
                    if (lane == r)
                        s.run_vac_mask[brun] = vac_mask;
                    if (e.inactive)
                    {
                        if (it == 0)
                            reset_inactive_slot(s, slot);
                    }
                    else
                    {
                        u32 event = 0, parent_track = 0;
                        real time = 0;
                        if (e.num_sec_all > 0)
                        {
                            event = s.event_id[slot];
                            parent_track = s.track_id[slot];
                            time = s.time[slot];
                        }
                        bool const now_inactive = end_slot_active(p,
                                                                  s,
                                                                  slot,
                                                                  e,
                                                                  chg_off,
                                                                  neu_off,
                                                                  sec_off,
                                                                  all_off,
                                                                  neutral_off,
                                                                  num_init_after,
                                                                  num_sec_total,
                                                                  event,
                                                                  parent_track,
                                                                  time);
                        if (now_inactive)
                        {
                            u32 const at = atomicAdd(&s.tail_ctrl[pending ^ 1u], 1u);
                            s.tail_reset_list[size_t(pending ^ 1u) * n + at] = slot;
                        }
                    }
                }
            }
        }
        grid.sync();
    }

    //// Exit: hand the state back to the per-action path ////
    if (it > 0)
    {
        // rebuild the sorted vacancy array from the per-run prefix and masks
        for (u32 run = gtid; run < nruns; run += nthreads)
        {
            u32 const first = run * RUN;
            u32 out = s.run_vac_prefix[run];
            u32 mask = s.run_vac_mask[run];
            while (mask)
            {
                u32 const i = __ffs(mask) - 1;
                mask &= mask - 1;
                s.vacancies[out++] = first + i;
            }
        }
    }
    if (gtid == 0)
    {
        u32 volatile* done = a.done;
        done[1] = reason;
        __threadfence_system();
        done[0] = it;
    }
}

//! Resident blocks per SM and cooperative launch of one instantiation
template<int FIELD>
inline cudaError_t tail_blocks_per_sm(int* per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, k_tail_loop<FIELD>, BLOCK, 0);
}

template<int FIELD>
inline cudaError_t tail_launch(ParamsView const& p,
                               StateView const& s,
                               TailArgs const& a,
                               u32 num_blocks,
                               cudaStream_t stream)
{
    TailArgs args = a;
    void* kargs[] = {const_cast<ParamsView*>(&p), const_cast<StateView*>(&s), &args};
    return cudaLaunchCooperativeKernel(reinterpret_cast<void const*>(&k_tail_loop<FIELD>),
                                       dim3(num_blocks), dim3(BLOCK), kargs, 0, stream);
}
}  // namespace b200
