//---------------------------------------------------------------------------//
// Allocate and initialise per-stream state.
//---------------------------------------------------------------------------//
#include "CoreState.hh"

#include "../../include/celeritas_b200.h"

#include <atomic>
#include <cstring>
#include <numeric>
#include <random>

using namespace b200;

namespace celeritas_b200
{
namespace
{
//! Position of word k (x0..x4, weyl) of slot i in StateView::rng (layout: csrc/rng.cuh)
inline size_t rng_index(int k, size_t i, size_t n)
{
#if B2_RNG_PACKED
    return k < 4 ? 4 * i + k : 4 * n + 2 * i + (k - 4);
#else
    return size_t(k) * n + i;
#endif
}
}  // namespace

CoreState::CoreState(std::shared_ptr<CoreParams const> params,
                     uint32_t stream_id,
                     uint32_t num_track_slots)
    : params_(std::move(params)), stream_id_(stream_id)
{
    if (num_track_slots == 0)
        throw std::runtime_error("num_track_slots must be positive");
    params_->freeze();
    B2_CUDA_CALL(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    ParamsView const& p = params_->view();
    StateView& s = view_;
    uint32_t const n = num_track_slots;
    uint32_t const D = p.geo.max_depth;
    uint32_t const P = p.phys.max_processes;
    s.num_slots = n;
    s.max_depth = D;
    s.max_processes = P;

    s.status = arena_.alloc<u8>(n);
    s.track_id = arena_.alloc_fill<u32>(n, 0xff);
    s.parent_id = arena_.alloc_fill<u32>(n, 0xff);
    s.event_id = arena_.alloc_fill<u32>(n, 0xff);
    s.num_steps = arena_.alloc<u32>(n);
    s.num_looping_steps = arena_.alloc<u32>(n);
    s.time = arena_.alloc<real>(n);
    s.step_length = arena_.alloc<real>(n);
    s.post_step_action = arena_.alloc_fill<u32>(n, 0xff);
    s.along_step_action = arena_.alloc_fill<u32>(n, 0xff);
    s.particle_id = arena_.alloc_fill<u32>(n, 0xff);
    s.energy = arena_.alloc<real>(n);
    s.material_id = arena_.alloc_fill<u32>(n, 0xff);

    s.geo_level = arena_.alloc<u32>(n);
    s.geo_surface_level = arena_.alloc_fill<u32>(n, 0xff);
    s.geo_surf = arena_.alloc_fill<u32>(n, 0xff);
    s.geo_sense = arena_.alloc<u8>(n);
    s.geo_boundary = arena_.alloc<u8>(n);
    s.geo_next_level = arena_.alloc_fill<u32>(n, 0xff);
    s.geo_next_step = arena_.alloc<real>(n);
    s.geo_next_surf = arena_.alloc_fill<u32>(n, 0xff);
    s.geo_next_sense = arena_.alloc<u8>(n);
    s.geo_pos = arena_.alloc<real>(size_t(3) * D * n);
    s.geo_dir = arena_.alloc<real>(size_t(3) * D * n);
    s.geo_vol = arena_.alloc<u32>(size_t(D) * n);
    s.geo_univ = arena_.alloc<u32>(size_t(D) * n);

    s.interaction_mfp = arena_.alloc<real>(n);
    s.macro_xs = arena_.alloc<real>(n);
    s.energy_deposition = arena_.alloc<real>(n);
    s.dedx_range = arena_.alloc<real>(n);
    s.msc_range = arena_.alloc<real>(size_t(3) * n);
    s.msc_is_displaced = arena_.alloc<u8>(n);
    s.msc_true_path = arena_.alloc<real>(n);
    s.msc_geom_path = arena_.alloc<real>(n);
    s.msc_alpha = arena_.alloc<real>(n);
    s.per_process_xs = arena_.alloc<real>(size_t(P ? P : 1) * n);
    s.element = arena_.alloc_fill<u32>(n, 0xff);
    s.sec_particle = arena_.alloc_fill<u32>(size_t(MAX_SECONDARIES) * n, 0xff);
    s.sec_energy = arena_.alloc<real>(size_t(MAX_SECONDARIES) * n);
    s.sec_dir = arena_.alloc<real>(size_t(MAX_SECONDARIES) * 3 * n);

    // RNG: same host-side seeding as the reference's initialize_xorwow
    // (/root/reference/src/celeritas/random/XorwowRngData.cc:28-58)
    {
        std::vector<std::seed_seq::result_type> host_seeds{params_->rng_seed()};
        if (stream_id != 0)
            host_seeds.push_back(stream_id);
        std::seed_seq seed_seq(host_seeds.begin(), host_seeds.end());
        std::mt19937 rng(seed_seq);
        std::uniform_int_distribution<uint32_t> sample;
        std::vector<uint32_t> soa(size_t(6) * n);
        for (uint32_t i = 0; i < n; ++i)
        {
            for (int k = 0; k < 6; ++k)
                soa[rng_index(k, i, n)] = sample(rng);
        }
        s.rng = const_cast<u32*>(arena_.upload(soa));
    }

    // Track initialization
    {
        std::vector<uint32_t> seq(n);
        std::iota(seq.begin(), seq.end(), 0u);
        s.vacancies = const_cast<u32*>(arena_.upload(seq));
        s.secondary_counts = arena_.alloc<u32>(size_t(n) + 1);
        s.parents = arena_.alloc_fill<u32>(n, 0xff);
        s.indices = arena_.alloc<u32>(n);
        s.track_counters = arena_.alloc<u32>(params_->max_events());
        uint32_t cap = params_->init_capacity();
        s.init_capacity = cap;
        s.ti_track_id = arena_.alloc<u32>(cap);
        s.ti_parent_id = arena_.alloc<u32>(cap);
        s.ti_event_id = arena_.alloc<u32>(cap);
        s.ti_particle_id = arena_.alloc<u32>(cap);
        s.ti_time = arena_.alloc<real>(cap);
        s.ti_energy = arena_.alloc<real>(cap);
        s.ti_pos = arena_.alloc<real>(size_t(3) * cap);
        s.ti_dir = arena_.alloc<real>(size_t(3) * cap);
        s.ti_level = arena_.alloc_fill<u32>(cap, 0xff);
        s.ti_vol = arena_.alloc<u32>(size_t(D) * cap);
        s.ti_univ = arena_.alloc<u32>(size_t(D) * cap);
        s.track_slots = arena_.alloc<u32>(n);
    }
    {
        std::vector<uint32_t> ctr(CTR_SIZE, 0);
        ctr[CTR_NUM_VACANCIES] = n;
        s.counters = const_cast<u32*>(arena_.upload(ctr));
        uint32_t num_blocks = (n + 127) / 128;
        s.block_scratch = arena_.alloc<u32>(size_t(6) * num_blocks);
        s.slot_class = arena_.alloc<u8>(n);
        if (p.scalars.track_order == ORDER_INIT_CHARGE)
            s.ti_neutral_prefix = arena_.alloc<u32>(size_t(params_->init_capacity()) + 1);
        s.single_event = INVALID;
        s.step_counters = arena_.alloc<u64>(4);
    }
    // Scoring
    if (params_->num_detectors() > 0)
    {
        s.pre_volume = arena_.alloc_fill<u32>(n, 0xff);
        s.calo_detector_of_volume = params_->detector_of_volume();
        s.calo_edep = arena_.alloc<real>(params_->num_detectors());
        s.num_detectors = params_->num_detectors();
    }
    // Per-model lists of interacting slots (16 = MAX_INTERACT_MODELS in kernels.cu)
    if (p.phys.num_models > 0 && p.phys.num_models <= 16)
    {
        s.interact_list = arena_.alloc<u32>(size_t(p.phys.num_models) * n);
        s.interact_count = arena_.alloc<u32>(16);
    }
    // Step/hit output
    bool const hits = params_->hit_detector_of_volume() != nullptr;
    if (hits)
    {
        if (!s.pre_volume)
            s.pre_volume = arena_.alloc_fill<u32>(n, 0xff);
        s.hit_detector_of_volume = params_->hit_detector_of_volume();
        s.hit_nonzero_edep = params_->hits_nonzero_edep() ? 1u : 0u;
        s.hit_pre = arena_.alloc<real>(size_t(8) * n);
        s.hit_u32 = arena_.alloc<u32>(size_t(6) * n);
        s.hit_f64 = arena_.alloc<real>(size_t(18) * n);
        s.hit_count = arena_.alloc<u32>(1);
    }
    // TrackOrder::reindex_* and the hit compaction: the sorted slot permutation
    // (csrc/kernels_sort.cu)
    if (p.scalars.track_order >= ORDER_REINDEX_STATUS || hits)
    {
        uint32_t const nkeys = std::max<uint32_t>(
            std::max<uint32_t>(uint32_t(params_->actions().size()), p.particle.num_particles), 1u);
        uint32_t const sort_blocks = (n + 255) / 256;
        std::vector<uint32_t> seq(n);
        std::iota(seq.begin(), seq.end(), 0u);
        s.sort_slots = const_cast<u32*>(arena_.upload(seq));
        s.sort_offsets = arena_.alloc<u32>(size_t(nkeys) + 2);
        s.sort_block_counts = arena_.alloc<u32>((size_t(nkeys) + 1) * sort_blocks);
        s.num_sort_keys = nkeys;
    }
    // Device-resident step loop (csrc/tail.cu); needs whole runs of 32 slots
    if (n % 32 == 0)
    {
        uint32_t const nruns = n / 32;
        s.run_vac_prefix = arena_.alloc<u32>(size_t(nruns) + 1);
        s.run_vac_mask = arena_.alloc<u32>(nruns);
        s.run_scan = arena_.alloc<u64>(size_t(2) * nruns);
        s.tail_reset_list = arena_.alloc<u32>(size_t(2) * n);
        s.tail_ctrl = arena_.alloc<u32>(2);
        size_t const ring_bytes
            = size_t(tail_ring_capacity) * B200_TAIL_RING_WORDS * sizeof(uint32_t);
        B2_CUDA_CALL(cudaHostAlloc(
            reinterpret_cast<void**>(&h_tail_ring_), ring_bytes, cudaHostAllocMapped));
        B2_CUDA_CALL(cudaHostAlloc(reinterpret_cast<void**>(&h_tail_done_),
                                   2 * sizeof(uint32_t),
                                   cudaHostAllocMapped));
        std::memset(h_tail_ring_, 0, ring_bytes);
        std::memset(h_tail_done_, 0, 2 * sizeof(uint32_t));
        void* mapped = nullptr;
        B2_CUDA_CALL(cudaHostGetDevicePointer(&mapped, h_tail_ring_, 0));
        d_tail_ring_ = static_cast<uint32_t*>(mapped);
        B2_CUDA_CALL(cudaHostGetDevicePointer(&mapped, h_tail_done_, 0));
        d_tail_done_ = static_cast<uint32_t*>(mapped);
    }
    B2_CUDA_CALL(cudaHostAlloc(reinterpret_cast<void**>(&h_counters_),
                               (CTR_SIZE + 1) * sizeof(uint32_t),
                               cudaHostAllocMapped));
    std::memset(h_counters_, 0, (CTR_SIZE + 1) * sizeof(uint32_t));
    {
        void* mapped = nullptr;
        B2_CUDA_CALL(cudaHostGetDevicePointer(&mapped, h_counters_, 0));
        s.host_counters = static_cast<u32*>(mapped);
    }
    B2_CUDA_CALL(cudaDeviceSynchronize());
}

CoreState::~CoreState()
{
    if (h_counters_)
        cudaFreeHost(h_counters_);
    if (h_tail_ring_)
        cudaFreeHost(h_tail_ring_);
    if (h_tail_done_)
        cudaFreeHost(h_tail_done_);
    if (stream_)
        cudaStreamDestroy(stream_);
}

CoreStateCounters CoreState::sync_counters()
{
    B2_CUDA_CALL(cudaMemcpyAsync(h_counters_,
                                 view_.counters,
                                 CTR_SIZE * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost,
                                 stream_));
    B2_CUDA_CALL(cudaStreamSynchronize(stream_));
    return this->unpack_counters();
}

CoreStateCounters CoreState::wait_counters()
{
    // The end-of-step scan writes the counters and then this iteration's sequence number
    // into mapped host memory (k_end_pass2)
    uint32_t volatile* flag = h_counters_ + CTR_SIZE;
    uint32_t const want = iteration_seq_;
    for (uint64_t spins = 0; *flag != want; ++spins)
    {
        if ((spins & 0x3ff) == 0x3ff)
        {
            // stream idle (or failed) without the flag: the kernels did not publish
            cudaError_t q = cudaStreamQuery(stream_);
            if (q == cudaSuccess)
            {
                if (*flag == want)
                    break;
                return this->sync_counters();
            }
            if (q != cudaErrorNotReady)
                B2_CUDA_CALL(q);
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return this->unpack_counters();
}

CoreStateCounters CoreState::unpack_tail_entry(uint32_t index)
{
    uint32_t const* e = h_tail_ring_ + size_t(index) * B200_TAIL_RING_WORDS;
    CoreStateCounters c;
    c.num_generated = e[0];
    c.num_initializers = e[1];
    c.num_vacancies = e[2];
    c.num_active = e[3];
    c.num_secondaries = e[4];
    c.num_alive = e[5];
    c.num_charged = e[6];
    c.num_neutral = e[7];
    c.first_busy_block = e[8];
    last_error_ = e[9];
    return c;
}

CoreStateCounters CoreState::unpack_counters()
{
    CoreStateCounters c;
    c.num_generated = h_counters_[CTR_NUM_GENERATED];
    c.num_initializers = h_counters_[CTR_NUM_INITIALIZERS];
    c.num_vacancies = h_counters_[CTR_NUM_VACANCIES];
    c.num_active = h_counters_[CTR_NUM_ACTIVE];
    c.num_secondaries = h_counters_[CTR_NUM_SECONDARIES];
    c.num_alive = h_counters_[CTR_NUM_ALIVE];
    c.num_charged = h_counters_[CTR_NUM_CHARGED];
    c.num_neutral = h_counters_[CTR_NUM_NEUTRAL];
    c.first_busy_block = h_counters_[CTR_FIRST_BUSY_BLOCK];
    last_error_ = h_counters_[CTR_ERROR];
    return c;
}

void CoreState::get_field(std::string const& f, void* out)
{
    StateView const& s = view_;
    size_t const n = s.num_slots;
    B2_CUDA_CALL(cudaStreamSynchronize(stream_));
    auto copy = [&](void const* src, size_t bytes) {
        B2_CUDA_CALL(cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
    };
    // Fetch status for masking inactive slots in geometry queries
    std::vector<u8> status(n);
    B2_CUDA_CALL(cudaMemcpy(status.data(), s.status, n, cudaMemcpyDeviceToHost));

    if (f == "status") copy(s.status, n);
    else if (f == "track_id") copy(s.track_id, 4 * n);
    else if (f == "parent_id") copy(s.parent_id, 4 * n);
    else if (f == "event_id") copy(s.event_id, 4 * n);
    else if (f == "num_steps") copy(s.num_steps, 4 * n);
    else if (f == "num_looping_steps") copy(s.num_looping_steps, 4 * n);
    else if (f == "time") copy(s.time, 8 * n);
    else if (f == "step_length") copy(s.step_length, 8 * n);
    else if (f == "post_step_action") copy(s.post_step_action, 4 * n);
    else if (f == "along_step_action") copy(s.along_step_action, 4 * n);
    else if (f == "particle_id") copy(s.particle_id, 4 * n);
    else if (f == "energy") copy(s.energy, 8 * n);
    else if (f == "material_id") copy(s.material_id, 4 * n);
    else if (f == "interaction_mfp") copy(s.interaction_mfp, 8 * n);
    else if (f == "macro_xs") copy(s.macro_xs, 8 * n);
    else if (f == "energy_deposition") copy(s.energy_deposition, 8 * n);
    else if (f == "dedx_range") copy(s.dedx_range, 8 * n);
    else if (f == "geo_level")
    {
        copy(s.geo_level, 4 * n);
        auto* o = static_cast<uint32_t*>(out);
        for (size_t i = 0; i < n; ++i)
            if (status[i] == ST_INACTIVE)
                o[i] = INVALID;
    }
    else if (f == "rng")
    {
        std::vector<uint32_t> soa(6 * n);
        B2_CUDA_CALL(cudaMemcpy(soa.data(), s.rng, 24 * n, cudaMemcpyDeviceToHost));
        auto* o = static_cast<uint32_t*>(out);
        for (size_t i = 0; i < n; ++i)
            for (int k = 0; k < 6; ++k)
                o[6 * i + k] = soa[rng_index(k, i, n)];
    }
    else if (f == "pos" || f == "dir")
    {
        // level 0 only, as [slot][3]
        size_t const stride = size_t(s.max_depth) * n;
        std::vector<double> soa(3 * stride);
        B2_CUDA_CALL(cudaMemcpy(soa.data(),
                                f == "pos" ? s.geo_pos : s.geo_dir,
                                soa.size() * 8,
                                cudaMemcpyDeviceToHost));
        auto* o = static_cast<double*>(out);
        for (size_t i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k)
            {
#if B2_POSDIR_PACKED
                size_t const at = k < 2 ? 2 * i + k : 2 * stride + i;
#else
                size_t const at = k * stride + i;
#endif
                o[3 * i + k] = status[i] == ST_INACTIVE ? 0 : soa[at];
            }
    }
    else if (f == "volume_id" || f == "surface_id")
    {
        GeoParams const& g = params_->view().geo;
        std::vector<uint32_t> level(n), vol(size_t(s.max_depth) * n),
            univ(size_t(s.max_depth) * n), slevel(n), surf(n);
        B2_CUDA_CALL(cudaMemcpy(level.data(), s.geo_level, 4 * n, cudaMemcpyDeviceToHost));
        B2_CUDA_CALL(cudaMemcpy(vol.data(), s.geo_vol, 4 * vol.size(), cudaMemcpyDeviceToHost));
        B2_CUDA_CALL(cudaMemcpy(univ.data(), s.geo_univ, 4 * univ.size(), cudaMemcpyDeviceToHost));
        B2_CUDA_CALL(cudaMemcpy(slevel.data(), s.geo_surface_level, 4 * n, cudaMemcpyDeviceToHost));
        B2_CUDA_CALL(cudaMemcpy(surf.data(), s.geo_surf, 4 * n, cudaMemcpyDeviceToHost));
        std::vector<uint32_t> voff(g.num_universes + 1), soff(g.num_universes + 1);
        B2_CUDA_CALL(cudaMemcpy(voff.data(), g.universe_volume_offset, 4 * voff.size(), cudaMemcpyDeviceToHost));
        B2_CUDA_CALL(cudaMemcpy(soff.data(), g.universe_surface_offset, 4 * soff.size(), cudaMemcpyDeviceToHost));
        auto* o = static_cast<uint32_t*>(out);
        for (size_t i = 0; i < n; ++i)
        {
            if (status[i] == ST_INACTIVE)
            {
                o[i] = INVALID;
            }
            else if (f == "volume_id")
            {
                size_t li = size_t(level[i]) * n + i;
                o[i] = voff[univ[li]] + vol[li];
            }
            else
            {
                if (slevel[i] == INVALID)
                    o[i] = INVALID;
                else
                    o[i] = soff[univ[size_t(slevel[i]) * n + i]] + surf[i];
            }
        }
    }
    else if (f == "sort_slots" || f == "sort_offsets")
    {
        if (!s.sort_slots)
            throw std::runtime_error("the track order of this problem does not sort tracks");
        if (f == "sort_slots")
            copy(s.sort_slots, size_t(4) * n);
        else
            copy(s.sort_offsets, size_t(4) * (s.num_sort_keys + 2));
    }
    else if (f == "interact_count")
    {
        // [16] interacting tracks per model in the last per-action step
        if (!s.interact_count)
            throw std::runtime_error("this state has no interaction lists");
        copy(s.interact_count, 4 * 16);
    }
    else if (f == "interact_list")
    {
        // [num_models][num_slots] slots of the interacting tracks, model-major
        if (!s.interact_list)
            throw std::runtime_error("this state has no interaction lists");
        copy(s.interact_list, 4 * n * params_->view().phys.num_models);
    }
    else if (f == "track_slots")
    {
        // dense active lists: charged from the front, neutral from the back
        copy(s.track_slots, 4 * n);
    }
    else
    {
        throw std::runtime_error("unknown state field '" + f + "'");
    }
}

uint32_t CoreState::hits_count()
{
    if (!view_.hit_count)
        throw std::runtime_error("the problem has no sensitive volumes (hits.volumes)");
    B2_CUDA_CALL(cudaStreamSynchronize(stream_));
    uint32_t n = 0;
    B2_CUDA_CALL(cudaMemcpy(&n, view_.hit_count, sizeof(n), cudaMemcpyDeviceToHost));
    return n;
}

void CoreState::hits_get(std::string const& field, void* out)
{
    uint32_t const count = this->hits_count();
    if (count == 0)
        return;
    size_t const n = view_.num_slots;
    static char const* const u32_fields[]
        = {"detector", "track_id", "event_id", "parent_id", "track_step_count", "particle"};
    for (size_t k = 0; k < 6; ++k)
    {
        if (field == u32_fields[k])
        {
            B2_CUDA_CALL(cudaMemcpy(out, view_.hit_u32 + k * n, size_t(4) * count,
                                    cudaMemcpyDeviceToHost));
            return;
        }
    }
    // double columns: 0 step_length, 1 edep, pre {2 time, 3-5 pos, 6-8 dir, 9 energy},
    // post {10 time, 11-13 pos, 14-16 dir, 17 energy}
    auto scalar = [&](size_t column) {
        B2_CUDA_CALL(cudaMemcpy(out, view_.hit_f64 + column * n, size_t(8) * count,
                                cudaMemcpyDeviceToHost));
    };
    auto vec3 = [&](size_t column) {
        std::vector<double> soa(size_t(3) * count);
        for (size_t k = 0; k < 3; ++k)
            B2_CUDA_CALL(cudaMemcpy(soa.data() + k * count, view_.hit_f64 + (column + k) * n,
                                    size_t(8) * count, cudaMemcpyDeviceToHost));
        auto* o = static_cast<double*>(out);
        for (size_t i = 0; i < count; ++i)
            for (size_t k = 0; k < 3; ++k)
                o[3 * i + k] = soa[k * count + i];
    };
    if (field == "step_length") scalar(0);
    else if (field == "energy_deposition") scalar(1);
    else if (field == "pre_time") scalar(2);
    else if (field == "pre_pos") vec3(3);
    else if (field == "pre_dir") vec3(6);
    else if (field == "pre_energy") scalar(9);
    else if (field == "post_time") scalar(10);
    else if (field == "post_pos") vec3(11);
    else if (field == "post_dir") vec3(14);
    else if (field == "post_energy") scalar(17);
    else
        throw std::runtime_error("unknown hit field '" + field + "'");
}

void CoreState::calo_get(double* out)
{
    if (!view_.calo_edep)
        throw std::runtime_error("no detectors registered");
    B2_CUDA_CALL(cudaStreamSynchronize(stream_));
    B2_CUDA_CALL(cudaMemcpy(
        out, view_.calo_edep, params_->num_detectors() * sizeof(double), cudaMemcpyDeviceToHost));
}

void CoreState::calo_clear()
{
    if (view_.calo_edep)
        B2_CUDA_CALL(cudaMemsetAsync(
            view_.calo_edep, 0, params_->num_detectors() * sizeof(double), stream_));
}

void CoreState::enable_action_diagnostic(uint32_t num_actions)
{
    if (view_.diag_action_counts)
        return;
    view_.diag_action_bins = num_actions;
    view_.diag_action_counts
        = arena_.alloc<u32>(size_t(num_actions) * params_->view().particle.num_particles);
}

void CoreState::enable_step_diagnostic(uint32_t max_step_bin)
{
    if (view_.diag_step_counts)
        return;
    if (max_step_bin == 0)
        throw std::runtime_error("nonpositive step diagnostic 'max' bin");
    // two extra bins for underflow and overflow (user/StepDiagnostic.cc:65-71)
    view_.diag_step_bins = max_step_bin + 2;
    view_.diag_step_counts
        = arena_.alloc<u32>(size_t(max_step_bin + 2) * params_->view().particle.num_particles);
}

void CoreState::diagnostic_get(bool steps, uint32_t* out)
{
    u32 const* src = steps ? view_.diag_step_counts : view_.diag_action_counts;
    size_t bins = steps ? view_.diag_step_bins : view_.diag_action_bins;
    if (!src)
        throw std::runtime_error("diagnostic is not enabled");
    B2_CUDA_CALL(cudaStreamSynchronize(stream_));
    B2_CUDA_CALL(cudaMemcpy(out,
                            src,
                            bins * params_->view().particle.num_particles * sizeof(u32),
                            cudaMemcpyDeviceToHost));
}

void CoreState::diagnostics_clear()
{
    size_t const np = params_->view().particle.num_particles;
    if (view_.diag_action_counts)
        B2_CUDA_CALL(cudaMemsetAsync(
            view_.diag_action_counts, 0, np * view_.diag_action_bins * sizeof(u32), stream_));
    if (view_.diag_step_counts)
        B2_CUDA_CALL(cudaMemsetAsync(
            view_.diag_step_counts, 0, np * view_.diag_step_bins * sizeof(u32), stream_));
}

uint64_t CoreState::num_tracks()
{
    std::vector<uint32_t> counters(params_->max_events());
    B2_CUDA_CALL(cudaStreamSynchronize(stream_));
    B2_CUDA_CALL(cudaMemcpy(counters.data(),
                            view_.track_counters,
                            counters.size() * sizeof(uint32_t),
                            cudaMemcpyDeviceToHost));
    return std::accumulate(counters.begin(), counters.end(), uint64_t(0));
}

void CoreState::reset()
{
    // reference: CoreState::reset (global/CoreState.cc:144-155): no tracks, no
    // initializers, every slot vacant
    uint32_t const n = view_.num_slots;
    B2_CUDA_CALL(cudaStreamSynchronize(stream_));
    B2_CUDA_CALL(cudaMemset(view_.status, 0, n));
    std::vector<uint32_t> seq(n);
    std::iota(seq.begin(), seq.end(), 0u);
    B2_CUDA_CALL(cudaMemcpy(view_.vacancies, seq.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    std::vector<uint32_t> ctr(CTR_SIZE, 0);
    ctr[CTR_NUM_VACANCIES] = n;
    B2_CUDA_CALL(
        cudaMemcpy(view_.counters, ctr.data(), CTR_SIZE * sizeof(uint32_t), cudaMemcpyHostToDevice));
    last_error_ = 0;
}
}  // namespace celeritas_b200
