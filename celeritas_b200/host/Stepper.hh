//---------------------------------------------------------------------------//
// Step actions, the ordered action sequence and the Stepper.
//
// Same shape as the reference's plugin surface:
//   StepActionInterface::step(CoreParams const&, CoreState&)
//     (/root/reference/src/corecel/sys/ActionInterface.hh:175-186)
//   ActionSequence::step, actions sorted by (order, action id)
//     (/root/reference/src/celeritas/global/ActionSequence.cc:37-138,
//      /root/reference/src/corecel/sys/ActionGroups.t.hh:23-54)
//   Stepper::operator()(primaries) / operator()() / warm_up / reseed / kill_active
//     (/root/reference/src/celeritas/global/Stepper.cc:66-201)
// Each concrete action is a thin adapter over one C-ABI launcher
// (include/celeritas_b200.h); there is no host implementation of any action.
//---------------------------------------------------------------------------//
#pragma once

#include <functional>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../../include/celeritas_b200.h"
#include "CoreParams.hh"
#include "CoreState.hh"

namespace celeritas_b200
{
class StepActionInterface
{
  public:
    virtual ~StepActionInterface() = default;
    virtual uint32_t action_id() const = 0;
    virtual std::string const& label() const = 0;
    virtual StepActionOrder order() const = 0;
    //! Launch on the state's stream (asynchronous)
    virtual void step(CoreParams const&, CoreState&) const = 0;
};

//! Adapter over a `int f(params view, state view, stream)` launcher
class KernelAction final : public StepActionInterface
{
  public:
    using Launcher = int (*)(B200ParamsView const*, B200StateView const*, cudaStream_t);
    KernelAction(uint32_t id, std::string label, StepActionOrder order, Launcher f)
        : id_(id), label_(std::move(label)), order_(order), launch_(f)
    {
    }
    uint32_t action_id() const override { return id_; }
    std::string const& label() const override { return label_; }
    StepActionOrder order() const override { return order_; }
    void step(CoreParams const& params, CoreState& state) const override;

  private:
    uint32_t id_;
    std::string label_;
    StepActionOrder order_;
    Launcher launch_;
};

//! Sort or partition the track-slot permutation (reference: SortTracksAction,
//! track/SortTracksAction.cc:46-131); one per TrackOrder::reindex_* key
class SortTracksAction final : public StepActionInterface
{
  public:
    SortTracksAction(uint32_t id, std::string label, StepActionOrder order, uint32_t track_order)
        : id_(id), label_(std::move(label)), order_(order), track_order_(track_order)
    {
    }
    uint32_t action_id() const override { return id_; }
    std::string const& label() const override { return label_; }
    StepActionOrder order() const override { return order_; }
    uint32_t track_order() const { return track_order_; }
    void step(CoreParams const& params, CoreState& state) const override;

  private:
    uint32_t id_;
    std::string label_;
    StepActionOrder order_;
    uint32_t track_order_;
};

class ActionSequence
{
  public:
    using SPAction = std::shared_ptr<StepActionInterface const>;
    struct Options
    {
        bool action_diagnostic{false};
        uint32_t step_diagnostic_bins{0};
        //! Extra user step actions; ids continue the action table
        std::vector<SPAction> user_actions;
        //! Iterations with at most this many active tracks run pre..user_post as one
        //! fused launch (0: default, 0xffffffff: never)
        uint32_t fuse_threshold{0};
        //! Stepper::advance runs iterations with at most this many active tracks inside
        //! the device-resident loop (0: default, 0xffffffff: never)
        uint32_t tail_threshold{0};
    };
    //! Default of Options::tail_threshold
    //! (measured, profiles/README_r02.md: above a few hundred tracks the loop's end of step
    //! costs what the host round trip does)
    static constexpr uint32_t default_tail_threshold = 256;
    //! Default of Options::fuse_threshold (measured: profiles/README_r01.md)
    // Swept on the final code (profiles/fuse_sweep_r02*.log): a plateau from 16 384 to
    // 32 768 tracks (TestEm3 89.4 ms, CMS-scale 337 ms per pass) against 90.7 / 342.8 ms at
    // the 65 536 that was best before the per-action kernels lost a quarter of their code
    static constexpr uint32_t default_fuse_threshold = 24576;
    //! Build the B200 adapters for every step action in the problem's table
    explicit ActionSequence(CoreParams const& params) : ActionSequence(params, Options{}) {}
    ActionSequence(CoreParams const& params, Options options);
    ~ActionSequence();
    void step(CoreParams const& params, CoreState& state);
    std::vector<SPAction> const& actions() const { return actions_; }
    //! Labels of ALL actions (explicit and implicit) by action id
    std::vector<std::string> const& labels() const { return labels_; }
    bool action_diagnostic() const { return action_diagnostic_; }
    uint32_t step_diagnostic_bins() const { return step_diagnostic_bins_; }
    bool action_times() const { return action_times_; }
    //! Whether small iterations can take the fused path, and up to how many tracks
    bool fusable() const { return fusable_; }
    uint32_t fuse_threshold() const { return fuse_threshold_; }
    //! Active-track bound of the device-resident loop (0: the loop is off)
    uint32_t tail_threshold() const { return tail_threshold_; }

    //! Per-action device timing with CUDA events on the state's stream
    //! (reference option: StepperInput::action_times, ActionSequence.cc:99-121)
    void action_times(bool enable) { action_times_ = enable; }
    //! Fold finished event pairs into the accumulators (stream must be idle)
    void collect_times();
    //! Accumulated device seconds per action since construction
    std::vector<double> const& accum_time() const { return accum_time_; }

  private:
    std::vector<SPAction> actions_;
    std::vector<std::string> labels_;
    bool action_diagnostic_{false};
    uint32_t step_diagnostic_bins_{0};
    bool action_times_{false};
    bool fusable_{false};
    uint32_t fuse_threshold_{0};
    uint32_t tail_threshold_{0};
    size_t tail_begin_{0}, tail_end_{0};  // [begin, end) of the boundary..diagnostics run
    size_t along_select_{0};  // index of the along-step action when the select follows it
    std::vector<double> accum_time_;
    struct Pending
    {
        uint32_t action;
        cudaEvent_t start, stop;
    };
    std::vector<Pending> pending_;
    std::vector<cudaEvent_t> pool_;
};

struct StepperResult
{
    uint32_t generated{};
    uint32_t queued{};
    uint32_t active{};
    uint32_t alive{};
    explicit operator bool() const { return queued > 0 || alive > 0; }
};

struct StepperInput
{
    std::shared_ptr<CoreParams const> params;
    uint32_t stream_id{0};
    uint32_t num_track_slots{0};
    bool action_times{false};
    ActionSequence::Options actions;
};

class Stepper
{
  public:
    explicit Stepper(StepperInput input);
    ~Stepper();

    void warm_up();
    StepperResult operator()();
    StepperResult operator()(B200Primary const* primaries, uint32_t n);
    void kill_active();
    void reseed(uint64_t event_id);
    //! Drop all tracks and initializers (reference: Stepper::reset_state)
    void reset_state();

    ActionSequence const& actions() const { return *actions_; }
    ActionSequence& action_sequence() { return *actions_; }
    CoreState& state() { return *state_; }
    CoreParams const& params() const { return *params_; }

    //! Up to `max_iterations` iterations without new primaries, stopping early when no track
    //! is left: Stepper::operator()() repeated. While few tracks are left the iterations run
    //! inside the device-resident loop (b200_step_tail_loop: no host round trip between
    //! them); results and state are the same either way. Appends one StepperResult (and,
    //! if asked, one duration in seconds) per iteration; returns the number taken.
    uint32_t advance(uint32_t max_iterations,
                     std::vector<StepperResult>* results,
                     std::vector<double>* seconds = nullptr);
    //! Iterations taken inside the device-resident loop / launches of it so far
    uint64_t tail_iterations() const { return tail_iterations_; }
    uint64_t tail_launches() const { return tail_launches_; }

    //! Enqueue one iteration without reading anything back
    void step_async();
    //! Stage primaries for the next iteration (host buffers)
    void insert(B200Primary const* primaries, uint32_t n);

    //!@{
    //! One iteration driven by an EXTERNAL action sequence (the reference's own
    //! ActionSequence calling one b200_step_* launcher per action):
    //! begin_iteration sizes the iteration's grids from the previous counters and moves
    //! staged primaries into the initializer queue (extend-from-primaries); the caller then
    //! launches the step actions on state().stream(); end_iteration waits for the counters
    //! that extend-from-secondaries publishes and returns the iteration's result.
    void begin_iteration();
    StepperResult end_iteration();
    //!@}

  private:
    std::shared_ptr<CoreParams const> params_;
    std::shared_ptr<ActionSequence> actions_;
    std::unique_ptr<CoreState> state_;

    // staging for primaries (pinned host + device)
    struct Staging;
    std::unique_ptr<Staging> staging_;
    std::set<uint32_t> events_in_flight_;
    CoreStateCounters last_{};  // counters at the end of the previous iteration
    uint32_t tail_blocks_{0};   // cooperative grid of the device-resident loop (0: unknown)
    uint64_t tail_iterations_{0};
    uint64_t tail_launches_{0};

    bool tail_eligible() const;
    uint32_t run_tail(uint32_t max_iterations,
                      std::vector<StepperResult>* results,
                      std::vector<double>* seconds);
    StepperResult finish_iteration(CoreStateCounters const& c);
};
}  // namespace celeritas_b200
