//---------------------------------------------------------------------------//
// Action sequence construction and the stepping loop.
//---------------------------------------------------------------------------//
#include "Stepper.hh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>

using namespace b200;

namespace celeritas_b200
{
namespace
{
void check_rc(int rc, char const* what)
{
    if (rc != 0)
        throw std::runtime_error(std::string(what) + " failed with code " + std::to_string(rc));
}

B200ParamsView const* pv(CoreParams const& p)
{
    return reinterpret_cast<B200ParamsView const*>(&p.view());
}
B200StateView const* sv(CoreState const& s)
{
    return reinterpret_cast<B200StateView const*>(&s.view());
}
}  // namespace

//---------------------------------------------------------------------------//
void KernelAction::step(CoreParams const& params, CoreState& state) const
{
    check_rc(launch_(pv(params), sv(state), state.stream()), label_.c_str());
}

//---------------------------------------------------------------------------//
// Primaries action: stages host primaries and launches the generate kernels
class ExtendFromPrimariesAction final : public StepActionInterface
{
  public:
    ExtendFromPrimariesAction(uint32_t id, std::string label) : id_(id), label_(std::move(label)) {}
    uint32_t action_id() const override { return id_; }
    std::string const& label() const override { return label_; }
    StepActionOrder order() const override { return StepActionOrder::generate; }
    void step(CoreParams const&, CoreState&) const override {}

  private:
    uint32_t id_;
    std::string label_;
};

struct Stepper::Staging
{
    uint32_t capacity{0};
    uint32_t count{0};
    uint32_t num_events{0};
    B200Primary* h_primaries{nullptr};  // pinned
    // pinned: rank[cap], event_ids[cap], event_counts[cap], neutral_inclusive[cap]
    uint32_t* h_aux{nullptr};
    B200Primary* d_primaries{nullptr};
    uint32_t* d_aux{nullptr};

    void reserve(uint32_t n)
    {
        if (n <= capacity)
            return;
        release();
        capacity = std::max<uint32_t>(n, 1024);
        B2_CUDA_CALL(cudaMallocHost(reinterpret_cast<void**>(&h_primaries),
                                    capacity * sizeof(B200Primary)));
        B2_CUDA_CALL(cudaMallocHost(reinterpret_cast<void**>(&h_aux), 4 * capacity * sizeof(uint32_t)));
        B2_CUDA_CALL(cudaMalloc(reinterpret_cast<void**>(&d_primaries), capacity * sizeof(B200Primary)));
        B2_CUDA_CALL(cudaMalloc(reinterpret_cast<void**>(&d_aux), 4 * capacity * sizeof(uint32_t)));
    }
    void release()
    {
        if (h_primaries) cudaFreeHost(h_primaries);
        if (h_aux) cudaFreeHost(h_aux);
        if (d_primaries) cudaFree(d_primaries);
        if (d_aux) cudaFree(d_aux);
        h_primaries = nullptr;
        h_aux = nullptr;
        d_primaries = nullptr;
        d_aux = nullptr;
        capacity = 0;
    }
    ~Staging() { release(); }
};

//---------------------------------------------------------------------------//
ActionSequence::ActionSequence(CoreParams const& params, Options options)
{
    using Order = StepActionOrder;
    auto const& view = params.view();
    uint32_t const model_begin = view.phys.model_to_action;
    uint32_t const model_end = model_begin + view.phys.num_models;
    bool have_interact = false;
    bool have_tally = false;
    bool have_sort = false;

    // Full action table: the problem's actions, then the diagnostics in the order
    // celer-sim registers them (app/celer-sim/Runner.cc:616-633), then user actions
    std::vector<ActionRecord> records = params.actions();
    auto has_label = [&records](std::string const& label) {
        return std::any_of(records.begin(), records.end(), [&](ActionRecord const& r) {
            return r.label == label;
        });
    };
    if (options.action_diagnostic && !has_label("action-diagnostic"))
    {
        records.push_back(
            {uint32_t(records.size()), "action-diagnostic", static_cast<uint32_t>(Order::post)});
    }
    if (options.step_diagnostic_bins && !has_label("step-diagnostic"))
    {
        records.push_back(
            {uint32_t(records.size()), "step-diagnostic", static_cast<uint32_t>(Order::user_post)});
    }
    step_diagnostic_bins_ = options.step_diagnostic_bins;
    for (SPAction const& user : options.user_actions)
    {
        if (!user || user->action_id() != records.size())
            throw std::runtime_error("user action ids must continue the action table");
        records.push_back({user->action_id(), user->label(), static_cast<uint32_t>(user->order())});
        actions_.push_back(user);
    }
    for (ActionRecord const& r : records)
        labels_.push_back(r.label);

    for (ActionRecord const& a : records)
    {
        if (a.id >= records.size() - options.user_actions.size())
            break;  // user actions are already in the list
        if (a.order == INVALID)
            continue;  // implicit action: no kernel
        Order order = static_cast<Order>(a.order);
        SPAction act;
        if (a.label == "extend-from-primaries")
        {
            act = std::make_shared<ExtendFromPrimariesAction>(a.id, a.label);
        }
        else if (a.label == "initialize-tracks")
        {
            act = std::make_shared<KernelAction>(a.id, a.label, order, &b200_step_initialize_tracks);
        }
        else if (a.label == "pre-step")
        {
            act = std::make_shared<KernelAction>(a.id, a.label, order, &b200_step_pre_step);
        }
        else if (a.label.rfind("along-step-", 0) == 0)
        {
            // Neutral and charged along-step are one launch: register once,
            // under the user (charged) action when present, else the neutral
            if (a.id == view.scalars.along_step_user_action)
            {
                act = std::make_shared<KernelAction>(a.id, a.label, order, &b200_step_along_step);
            }
        }
        else if (a.label == "physics-discrete-select")
        {
            act = std::make_shared<KernelAction>(a.id, a.label, order, &b200_step_discrete_select);
        }
        else if (a.label.rfind("sort-tracks-", 0) == 0)
        {
            // TrackOrder::reindex_*: labels of SortTracksAction::label()
            // (track/SortTracksAction.cc:83-98)
            uint32_t key = INVALID;
            if (a.label == "sort-tracks-status")
                key = ORDER_REINDEX_STATUS;
            else if (a.label == "sort-tracks-start")
                key = ORDER_REINDEX_PARTICLE_TYPE;
            else if (a.label == "sort-tracks-along-step")
                key = ORDER_REINDEX_ALONG_STEP_ACTION;
            else if (a.label == "sort-tracks-post-step")
                key = ORDER_REINDEX_STEP_LIMIT_ACTION;
            else
                throw std::runtime_error("unknown sort action '" + a.label + "'");
            act = std::make_shared<SortTracksAction>(a.id, a.label, order, key);
            have_sort = true;
        }
        else if (a.id >= model_begin && a.id < model_end)
        {
            // All discrete models share one launch that dispatches on action id
            if (!have_interact)
            {
                act = std::make_shared<KernelAction>(
                    a.id, "interact[" + a.label + ",...]", order, &b200_step_interact);
                have_interact = true;
            }
        }
        else if (a.label == "geo-boundary")
        {
            act = std::make_shared<KernelAction>(a.id, a.label, order, &b200_step_boundary);
        }
        else if (a.label == "tracking-cut")
        {
            act = std::make_shared<KernelAction>(a.id, a.label, order, &b200_step_tracking_cut);
        }
        else if (a.label == "extend-from-secondaries")
        {
            act = std::make_shared<KernelAction>(
                a.id, a.label, order, &b200_step_extend_from_secondaries);
        }
        else if (a.label == "action-diagnostic")
        {
            act = std::make_shared<KernelAction>(a.id, a.label, order, &b200_step_action_diagnostic);
            action_diagnostic_ = true;
        }
        else if (a.label == "step-diagnostic")
        {
            if (step_diagnostic_bins_ == 0)
                throw std::runtime_error(
                    "the problem has a step diagnostic: set step_diagnostic_bins");
            act = std::make_shared<KernelAction>(a.id, a.label, order, &b200_step_step_diagnostic);
        }
        else if (a.label.rfind("step-gather-", 0) == 0)
        {
            // pre-step gather is folded into pre-step; post-step gather + calo
            // are one tally launch at user_post
            if (order == Order::user_post && !have_tally)
            {
                act = std::make_shared<KernelAction>(a.id, "tally[" + a.label + "]", order, &b200_step_tally);
                have_tally = true;
            }
        }
        else
        {
            throw std::runtime_error("no B200 kernel for step action '" + a.label + "'");
        }
        if (act)
            actions_.push_back(std::move(act));
    }
    std::stable_sort(actions_.begin(), actions_.end(), [](SPAction const& a, SPAction const& b) {
        if (a->order() != b->order())
            return a->order() < b->order();
        return a->action_id() < b->action_id();
    });

    // The fused small-iteration launch covers exactly the built-in actions between
    // `pre` and `user_post`; any user action in that range turns it off
    fusable_ = true;
    for (SPAction const& user : options.user_actions)
    {
        if (user->order() >= Order::pre && user->order() <= Order::user_post)
            fusable_ = false;
    }
    // A sort action has to see the state between two groups of the fused launch; the
    // step/hit output is gathered after every iteration by the tally launch
    if (have_sort || params.hit_detector_of_volume())
        fusable_ = false;
    // The fused step and the device-resident loop are built with the core interactors only
    if (params.view().model.has_extra_models)
        fusable_ = false;
    fuse_threshold_ = options.fuse_threshold ? options.fuse_threshold : default_fuse_threshold;
    if (char const* env = std::getenv("B200_FUSE_THRESHOLD"))
        fuse_threshold_ = static_cast<uint32_t>(std::strtoul(env, nullptr, 10));
    if (fuse_threshold_ == 0xffffffffu)
        fusable_ = false;
    // The device-resident loop runs the same fused step: same precondition
    tail_threshold_ = options.tail_threshold ? options.tail_threshold : default_tail_threshold;
    if (char const* env = std::getenv("B200_TAIL_THRESHOLD"))
        tail_threshold_ = static_cast<uint32_t>(std::strtoul(env, nullptr, 10));
    if (tail_threshold_ == 0xffffffffu || !fusable_)
        tail_threshold_ = 0;

    // The run [geo-boundary, tracking-cut, action-diagnostic?, tally?, step-diagnostic?] of
    // consecutive built-in actions is one launch (b200_step_post_tail) when nothing else
    // sits in between
    tail_begin_ = tail_end_ = 0;
    for (size_t i = 0; i < actions_.size(); ++i)
    {
        if (actions_[i]->label() != "geo-boundary")
            continue;
        size_t j = i + 1;
        auto is_tail = [](std::string const& label) {
            return label == "tracking-cut" || label == "action-diagnostic"
                   || label == "step-diagnostic" || label.rfind("tally[", 0) == 0;
        };
        while (j < actions_.size() && is_tail(actions_[j]->label())
               && dynamic_cast<KernelAction const*>(actions_[j].get()))
            ++j;
        if (j - i >= 2)
        {
            tail_begin_ = i;
            tail_end_ = j;
        }
        break;
    }
    if (std::getenv("B200_NO_POST_TAIL"))
        tail_begin_ = tail_end_ = 0;

    // along-step directly followed by the discrete select: one launch per charge class
    // does both (b200_step_along_select)
    along_select_ = actions_.size();
    for (size_t i = 0; i + 1 < actions_.size(); ++i)
    {
        if (actions_[i]->label().rfind("along-step-", 0) == 0
            && actions_[i + 1]->label() == "physics-discrete-select"
            && view.phys.num_models > 0 && view.phys.num_models <= 16)
        {
            along_select_ = i;
        }
    }
    // Measured: not faster (105.8 vs 105.2 ms per pass on one stream, 95.9 vs 94.5 on two;
    // profiles/README_r01.md), so it is opt-in
    if (!std::getenv("B200_ALONG_SELECT"))
        along_select_ = actions_.size();
}

void SortTracksAction::step(CoreParams const& params, CoreState& state) const
{
    check_rc(b200_step_sort_tracks(pv(params), sv(state), track_order_, state.stream()),
             label_.c_str());
}

ActionSequence::~ActionSequence()
{
    for (auto& p : pending_)
    {
        cudaEventDestroy(p.start);
        cudaEventDestroy(p.stop);
    }
    for (auto e : pool_)
        cudaEventDestroy(e);
}

void ActionSequence::step(CoreParams const& params, CoreState& state)
{
    if (fusable_ && !action_times_ && state.view().hint_active <= fuse_threshold_)
    {
        // Small iteration: one launch for everything between `pre` and `user_post`
        for (auto const& a : actions_)
            if (a->order() < StepActionOrder::pre)
                a->step(params, state);
        check_rc(b200_step_fused(pv(params), sv(state), state.stream()), "step_fused");
        for (auto const& a : actions_)
            if (a->order() > StepActionOrder::user_post)
                a->step(params, state);
        return;
    }
    if (!action_times_)
    {
        for (size_t i = 0; i < actions_.size(); ++i)
        {
            if (i == tail_begin_ && tail_end_ > tail_begin_)
            {
                check_rc(b200_step_post_tail(pv(params), sv(state), state.stream()),
                         "step_post_tail");
                i = tail_end_ - 1;
                continue;
            }
            if (i == along_select_)
            {
                check_rc(b200_step_along_select(pv(params), sv(state), state.stream()),
                         "step_along_select");
                ++i;  // the discrete select is done
                continue;
            }
            actions_[i]->step(params, state);
        }
        return;
    }
    auto get_event = [this] {
        cudaEvent_t e;
        if (!pool_.empty())
        {
            e = pool_.back();
            pool_.pop_back();
        }
        else
        {
            B2_CUDA_CALL(cudaEventCreate(&e));
        }
        return e;
    };
    for (uint32_t i = 0; i < actions_.size(); ++i)
    {
        Pending p{i, get_event(), get_event()};
        B2_CUDA_CALL(cudaEventRecord(p.start, state.stream()));
        actions_[i]->step(params, state);
        B2_CUDA_CALL(cudaEventRecord(p.stop, state.stream()));
        pending_.push_back(p);
    }
}

void ActionSequence::collect_times()
{
    accum_time_.resize(actions_.size(), 0.0);
    for (auto& p : pending_)
    {
        float ms = 0;
        B2_CUDA_CALL(cudaEventElapsedTime(&ms, p.start, p.stop));
        accum_time_[p.action] += ms * 1e-3;
        pool_.push_back(p.start);
        pool_.push_back(p.stop);
    }
    pending_.clear();
}

//---------------------------------------------------------------------------//
Stepper::Stepper(StepperInput input) : params_(std::move(input.params))
{
    if (!params_)
        throw std::runtime_error("Stepper requires params");
    actions_ = std::make_shared<ActionSequence>(*params_, std::move(input.actions));
    actions_->action_times(input.action_times);
    state_ = std::make_unique<CoreState>(params_, input.stream_id, input.num_track_slots);
    if (actions_->action_diagnostic())
        state_->enable_action_diagnostic(actions_->labels().size());
    if (actions_->step_diagnostic_bins())
        state_->enable_step_diagnostic(actions_->step_diagnostic_bins());
    staging_ = std::make_unique<Staging>();
    last_.num_vacancies = input.num_track_slots;
}

Stepper::~Stepper() = default;

void Stepper::insert(B200Primary const* primaries, uint32_t n)
{
    if (staging_->count != 0)
        throw std::runtime_error("multiple consecutive primary insertions");
    if (n == 0)
        return;
    // reference: CELER_VALIDATE in ExtendFromPrimariesAction::insert
    // (track/ExtendFromPrimariesAction.cc:107-113)
    if (uint64_t(n) + last_.num_initializers > params_->init_capacity())
    {
        throw std::runtime_error("insufficient initializer capacity ("
                                 + std::to_string(params_->init_capacity()) + ") with size ("
                                 + std::to_string(last_.num_initializers) + ") for primaries ("
                                 + std::to_string(n) + ")");
    }
    staging_->reserve(n);
    Staging& st = *staging_;
    std::copy(primaries, primaries + n, st.h_primaries);
    // Deterministic track ids: rank of each primary among earlier primaries of
    // the same event, plus per-event totals to advance the device counters
    uint32_t* rank = st.h_aux;
    uint32_t* ev_ids = st.h_aux + st.capacity;
    uint32_t* ev_counts = st.h_aux + 2 * st.capacity;
    std::map<uint32_t, uint32_t> counts;
    for (uint32_t i = 0; i < n; ++i)
    {
        if (primaries[i].event_id >= params_->max_events())
            throw std::runtime_error("primary event id exceeds max_events");
        rank[i] = counts[primaries[i].event_id]++;
    }
    // TrackOrder::init_charge: running count of neutral primaries (see k_initialize_tracks)
    {
        uint32_t* neutral_inclusive = st.h_aux + 3 * st.capacity;
        uint32_t running = 0;
        for (uint32_t i = 0; i < n; ++i)
        {
            if (params_->particle_is_neutral(primaries[i].particle_id))
                ++running;
            neutral_inclusive[i] = running;
        }
    }
    uint32_t ne = 0;
    for (auto const& kv : counts)
    {
        ev_ids[ne] = kv.first;
        ev_counts[ne] = kv.second;
        ++ne;
    }
    st.count = n;
    st.num_events = ne;
    // Track which events are in flight: with exactly one, secondaries get
    // deterministic slot-ordered track ids
    for (auto const& kv : counts)
        events_in_flight_.insert(kv.first);
    state_->single_event(events_in_flight_.size() == 1 ? *events_in_flight_.begin() : INVALID);
}

void Stepper::step_async()
{
    this->begin_iteration();
    actions_->step(*params_, *state_);
}

void Stepper::begin_iteration()
{
    CoreState& state = *state_;
    cudaStream_t stream = state.stream();
    Staging& st = *staging_;
    {
        // Grid sizing: exact counts from the previous iteration's counters. Every
        // queued initializer (plus staged primaries) may start if a slot is vacant.
        uint64_t const n = state.size();
        uint64_t queued = uint64_t(last_.num_initializers) + st.count;
        uint64_t fresh = std::min<uint64_t>(queued, last_.num_vacancies);
        // The end-of-step passes start at the first block that can hold a track: the
        // lowest busy block of the last step, lowered by the tracks about to start
        // (they take the highest vacancies; at worst all of them lie just below it)
        uint64_t const block = 128;
        uint64_t busy_begin = last_.first_busy_block == INVALID
                                  ? n
                                  : std::min<uint64_t>(n, last_.first_busy_block * block);
        uint64_t slot_begin = busy_begin > fresh ? (busy_begin - fresh) / block * block : 0;
        // init_charge gives neutral tracks the LOWEST vacancies: the whole range is live
        if (params_->view().scalars.track_order == ORDER_INIT_CHARGE)
            slot_begin = 0;
        state.launch_hints(std::min<uint64_t>(n, last_.num_alive + fresh),
                           std::min<uint64_t>(n, last_.num_charged + fresh),
                           std::min<uint64_t>(n, last_.num_neutral + fresh),
                           fresh,
                           slot_begin);
    }
    check_rc(b200_reset_generated(sv(state), stream), "reset_generated");
    if (st.count > 0)
    {
        B2_CUDA_CALL(cudaMemcpyAsync(st.d_primaries,
                                     st.h_primaries,
                                     st.count * sizeof(B200Primary),
                                     cudaMemcpyHostToDevice,
                                     stream));
        B2_CUDA_CALL(cudaMemcpyAsync(st.d_aux,
                                     st.h_aux,
                                     4 * st.capacity * sizeof(uint32_t),
                                     cudaMemcpyHostToDevice,
                                     stream));
        check_rc(b200_step_extend_from_primaries(sv(state),
                                                 st.d_primaries,
                                                 st.d_aux,
                                                 st.d_aux + st.capacity,
                                                 st.d_aux + 2 * st.capacity,
                                                 st.d_aux + 3 * st.capacity,
                                                 st.num_events,
                                                 st.count,
                                                 stream),
                 "extend_from_primaries");
        st.count = 0;
    }
}

StepperResult Stepper::end_iteration()
{
    return this->finish_iteration(state_->wait_counters());
}

StepperResult Stepper::finish_iteration(CoreStateCounters const& c)
{
    last_ = c;
    if (uint32_t err = state_->last_device_error())
    {
        if (err == B200_ERR_INITIALIZER_CAPACITY)
            throw std::runtime_error(
                "insufficient capacity (" + std::to_string(params_->init_capacity())
                + ") for track initializers");
        throw std::runtime_error("device error " + std::to_string(err));
    }
    StepperResult r;
    r.generated = c.num_generated;
    r.active = c.num_active;
    r.alive = c.num_alive;
    r.queued = c.num_initializers;
    if (!r)
    {
        events_in_flight_.clear();
        state_->single_event(INVALID);
    }
    return r;
}

StepperResult Stepper::operator()()
{
    this->step_async();
    // With per-action timing the event pairs must have completed: full synchronisation.
    // Otherwise take the counters as soon as the end-of-step scan has published them.
    CoreStateCounters c = (actions_->action_times() || std::getenv("B200_FULL_SYNC"))
                              ? state_->sync_counters()
                              : state_->wait_counters();
    actions_->collect_times();
    return this->finish_iteration(c);
}

//---------------------------------------------------------------------------//
// Device-resident loop (csrc/tail.cu)
//---------------------------------------------------------------------------//
bool Stepper::tail_eligible() const
{
    uint32_t const threshold = actions_->tail_threshold();
    if (threshold == 0 || actions_->action_times() || staging_->count != 0)
        return false;
    if (!state_->tail_ring_device())
        return false;
    if (last_.num_alive == 0 && last_.num_initializers == 0)
        return false;
    // tracks in the next iteration: the ones alive plus the queued ones that find a slot
    uint64_t const next_active
        = uint64_t(last_.num_alive) + std::min(last_.num_initializers, last_.num_vacancies);
    return next_active <= threshold;
}

uint32_t Stepper::run_tail(uint32_t max_iterations,
                           std::vector<StepperResult>* results,
                           std::vector<double>* seconds)
{
    using Clock = std::chrono::steady_clock;
    CoreState& state = *state_;
    if (tail_blocks_ == 0)
    {
        int max_blocks = 0;
        check_rc(b200_tail_max_blocks(pv(*params_), &max_blocks), "tail_max_blocks");
        int device = 0, sms = 0;
        B2_CUDA_CALL(cudaGetDevice(&device));
        B2_CUDA_CALL(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        // One block on every other SM by default: the loop is latency bound (a few tracks),
        // and while it is resident a second stream's kernels keep the rest of the GPU
        // (measured with two streams, profiles/README_r02.md: CMS-scale pass 379 ms without
        // the loop, 381 ms with 148 blocks, 352 ms with 74, 354 ms with 32)
        int blocks = std::min(max_blocks, std::max(1, sms / 2));
        if (char const* env = std::getenv("B200_TAIL_BLOCKS"))
            blocks = std::min<int>(max_blocks, std::max<int>(1, std::atoi(env)));
        if (blocks <= 0)
            throw std::runtime_error("the device-resident loop does not fit on this device");
        tail_blocks_ = static_cast<uint32_t>(blocks);
    }
    uint32_t const chunk = std::min<uint32_t>(max_iterations, CoreState::tail_ring_capacity);
    uint32_t const exit_active
        = std::min<uint64_t>(uint64_t(2) * actions_->tail_threshold(), state.size());
    uint32_t volatile* done = state.tail_done_host();
    done[0] = 0xffffffffu;
    done[1] = 0xffffffffu;
    auto const start = Clock::now();
    check_rc(b200_step_tail_loop(pv(*params_),
                                 sv(state),
                                 tail_blocks_,
                                 chunk,
                                 exit_active,
                                 state.tail_ring_device(),
                                 state.tail_done_device(),
                                 state.stream()),
             "step_tail_loop");
    B2_CUDA_CALL(cudaStreamSynchronize(state.stream()));
    double const elapsed = std::chrono::duration<double>(Clock::now() - start).count();
    uint32_t const n = done[0];
    if (n == 0xffffffffu)
        throw std::runtime_error("the device-resident loop did not report back");
    ++tail_launches_;
    tail_iterations_ += n;
    uint32_t const* ring = state.tail_ring_host();
    auto stamp = [ring](uint32_t i) {
        uint32_t const* e = ring + size_t(i) * B200_TAIL_RING_WORDS;
        return uint64_t(e[10]) | (uint64_t(e[11]) << 32);
    };
    if (char const* dump = std::getenv("B200_TAIL_DUMP"))
    {
        // phase durations of every iteration (profiling aid): active, A, B, C1, C2 [ns]
        if (FILE* f = std::fopen(dump, "a"))
        {
            for (uint32_t i = 0; i < n; ++i)
            {
                uint32_t const* e = ring + size_t(i) * B200_TAIL_RING_WORDS;
                uint64_t total = i > 0 ? stamp(i) - stamp(i - 1) : 0;
                std::fprintf(f, "%u %u %u %u %u %llu\n", e[3], e[12], e[13], e[14], e[15],
                             static_cast<unsigned long long>(total));
            }
            std::fclose(f);
        }
    }
    for (uint32_t i = 0; i < n; ++i)
    {
        CoreStateCounters c = state.unpack_tail_entry(i);
        results->push_back(this->finish_iteration(c));
        if (seconds)
        {
            // device timer between consecutive end-of-step scans; the first iteration gets
            // what is left of the launch's host-side duration
            double dt = i > 0 ? double(stamp(i) - stamp(i - 1)) * 1e-9
                              : std::max(0.0, elapsed - double(stamp(n - 1) - stamp(0)) * 1e-9);
            seconds->push_back(dt);
        }
    }
    return n;
}

uint32_t Stepper::advance(uint32_t max_iterations,
                          std::vector<StepperResult>* results,
                          std::vector<double>* seconds)
{
    using Clock = std::chrono::steady_clock;
    if (!results)
        throw std::runtime_error("advance needs a result vector");
    uint32_t done = 0;
    while (done < max_iterations)
    {
        if (this->tail_eligible())
        {
            uint32_t const n = this->run_tail(max_iterations - done, results, seconds);
            done += n;
            if (n > 0)
            {
                if (!results->back())
                    break;
                continue;
            }
        }
        auto const start = Clock::now();
        StepperResult r = (*this)();
        results->push_back(r);
        if (seconds)
            seconds->push_back(std::chrono::duration<double>(Clock::now() - start).count());
        ++done;
        if (!r)
            break;
    }
    return done;
}

StepperResult Stepper::operator()(B200Primary const* primaries, uint32_t n)
{
    this->insert(primaries, n);
    return (*this)();
}

void Stepper::warm_up()
{
    // reference: Stepper::warm_up (Stepper.cc:104-115) requires an empty state
    CoreStateCounters c = state_->sync_counters();
    if (c.num_alive != 0 || c.num_initializers != 0 || staging_->count != 0)
        throw std::runtime_error("cannot warm up when state has active tracks");
    (*this)();
}

void Stepper::kill_active()
{
    check_rc(b200_kill_active(pv(*params_), sv(*state_), state_->stream()), "kill_active");
}

void Stepper::reset_state()
{
    state_->reset();
    staging_->count = 0;
    events_in_flight_.clear();
    state_->single_event(INVALID);
    last_ = {};
    last_.num_vacancies = state_->size();
}

void Stepper::reseed(uint64_t event_id)
{
    check_rc(b200_reseed(pv(*params_), sv(*state_), event_id, state_->stream()), "reseed");
    // reference: Stepper::reseed also zeroes the track-id counters (Stepper.cc:193-201)
    B2_CUDA_CALL(cudaMemsetAsync(state_->view().track_counters,
                                 0,
                                 params_->max_events() * sizeof(uint32_t),
                                 state_->stream()));
}
}  // namespace celeritas_b200
