//---------------------------------------------------------------------------//
// Reference-side binding of the B200 step actions (see B200Actions.hh).
//---------------------------------------------------------------------------//
#include "B200Actions.hh"

#include <algorithm>

#include "corecel/Assert.hh"
#include "corecel/data/AuxParamsRegistry.hh"
#include "corecel/data/AuxStateVec.hh"
#include "corecel/sys/Device.hh"
#include "corecel/sys/Stream.hh"
#include "celeritas/track/TrackInitParams.hh"

namespace celeritas_b200_adapter
{
namespace
{
//! C-ABI error convention -> the reference's (CELER_VALIDATE -> RuntimeError)
void check(int rc, char const* what)
{
    CELER_VALIDATE(rc == B200_OK, << "B200 " << what << " failed (" << rc
                                   << "): " << b200_last_error());
}

//! Placeholder for the reference's implicit actions (labels and ids only, no kernel)
class ImplicitAction final : public celeritas::ConcreteAction
{
  public:
    using ConcreteAction::ConcreteAction;
};

constexpr char const aux_label[] = "b200-track-state";
}  // namespace

//---------------------------------------------------------------------------//
B200Problem::B200Problem(void const* image, std::size_t size)
{
    check(b200_params_create_from_memory(image, size, &params_), "params_create_from_memory");
}

B200Problem::~B200Problem()
{
    b200_params_destroy(params_);
}

//---------------------------------------------------------------------------//
B200AuxState::B200AuxState(B200Problem const& problem, StreamId stream, size_type num_track_slots)
{
    // One launch per action, driven by the reference's ActionSequence: the library's own
    // fused launch and device-resident loop stay off
    B200StepperOptions options{};
    options.stream_id = stream.get();
    options.num_track_slots = num_track_slots;
    options.fuse_threshold = 0xffffffffu;
    options.tail_threshold = 0xffffffffu;
    check(b200_stepper_create_opts(problem.get(), &options, &stepper_), "stepper_create");
}

B200AuxState::~B200AuxState()
{
    b200_stepper_destroy(stepper_);
}

auto B200AuxParams::create_state(MemSpace m, StreamId stream, size_type size) const -> UPState
{
    CELER_VALIDATE(m == MemSpace::device, << "the B200 track loop has no host implementation");
    return std::make_unique<B200AuxState>(*problem_, stream, size);
}

//---------------------------------------------------------------------------//
B200StepAction::B200StepAction(ActionId id,
                               std::string label,
                               std::string description,
                               StepActionOrder order,
                               B200Role role,
                               Launcher launch,
                               std::shared_ptr<B200Problem const> problem,
                               AuxId state_id)
    : ConcreteAction(id, std::move(label), std::move(description))
    , order_(order)
    , role_(role)
    , launch_(launch)
    , problem_(std::move(problem))
    , state_id_(state_id)
{
    CELER_EXPECT(problem_ && state_id_);
    CELER_EXPECT(launch_ || role_ == B200Role::begin_iteration || role_ == B200Role::merged);
}

void B200StepAction::step(CoreParams const&, CoreStateHost&) const
{
    CELER_NOT_CONFIGURED("B200 host execution");
}

void B200StepAction::step(CoreParams const&, CoreStateDevice& state) const
{
    auto& b2 = celeritas::get<B200AuxState>(state.aux(), state_id_);
    switch (role_)
    {
        case B200Role::merged:
            return;
        case B200Role::begin_iteration:
            check(b200_stepper_begin_iteration(b2.handle()), "begin_iteration");
            return;
        case B200Role::launch:
            check(launch_(problem_->view(), b2.view(), b2.stream()), this->label().data());
            return;
        case B200Role::end_iteration: {
            check(launch_(problem_->view(), b2.view(), b2.stream()), this->label().data());
            // The reference's ExtendFromSecondariesAction leaves the iteration's counters in
            // CoreState::counters() (ExtendFromSecondariesAction.cc:55-100); so does this
            B200StepperResult r{};
            check(b200_stepper_end_iteration(b2.handle(), &r), "end_iteration");
            auto& c = state.counters();
            c.num_generated = r.generated;
            c.num_initializers = r.queued;
            c.num_active = r.active;
            c.num_alive = r.alive;
            c.num_vacancies = state.size() - r.alive;
            return;
        }
    }
}

void B200StepAction::insert(CoreStateDevice& state,
                            celeritas::Span<celeritas::Primary const> primaries) const
{
    CELER_EXPECT(role_ == B200Role::begin_iteration);
    std::vector<B200Primary> staged(primaries.size());
    for (std::size_t i = 0; i < primaries.size(); ++i)
    {
        celeritas::Primary const& p = primaries[i];
        B200Primary& q = staged[i];
        q.particle_id = p.particle_id.get();
        q.event_id = p.event_id.get();
        q.energy = p.energy.value();
        for (int k = 0; k < 3; ++k)
        {
            q.pos[k] = p.position[k];
            q.dir[k] = p.direction[k];
        }
        q.time = p.time;
    }
    auto& b2 = celeritas::get<B200AuxState>(state.aux(), state_id_);
    check(b200_stepper_insert(b2.handle(), staged.data(), staged.size()), "insert");
}

//---------------------------------------------------------------------------//
B200Registry make_b200_registry(CoreParams const& core,
                                std::shared_ptr<B200Problem const> problem,
                                AuxId state_id)
{
    using Order = StepActionOrder;
    auto const& ref = *core.action_reg();
    auto const& scalars = core.host_ref().scalars;
    ActionId::size_type const model_begin = b200_params_model_action_begin(problem->get());
    ActionId::size_type const model_end = model_begin + b200_params_num_models(problem->get());

    B200Registry result;
    result.actions = std::make_shared<celeritas::ActionRegistry>();
    bool have_interact = false;
    bool have_tally = false;
    for (ActionId::size_type i = 0; i < ref.num_actions(); ++i)
    {
        ActionId const id{i};
        auto const& base = ref.action(id);
        std::string const label{base->label()};
        std::string const descr{base->description()};
        auto const* step = dynamic_cast<celeritas::CoreStepActionInterface const*>(base.get());
        CELER_ASSERT(result.actions->next_id() == id);
        if (!step)
        {
            result.actions->insert(std::make_shared<ImplicitAction>(id, label, descr));
            continue;
        }
        Order const order = step->order();
        B200Role role = B200Role::launch;
        B200StepAction::Launcher launch = nullptr;
        if (label == "extend-from-primaries")
        {
            role = B200Role::begin_iteration;
        }
        else if (label == "initialize-tracks")
        {
            launch = &b200_step_initialize_tracks;
        }
        else if (label == "pre-step")
        {
            launch = &b200_step_pre_step;
        }
        else if (label.rfind("along-step-", 0) == 0)
        {
            // neutral and charged along-step are one launch, under the user action's id
            if (id == scalars.along_step_user_action)
                launch = &b200_step_along_step;
            else
                role = B200Role::merged;
        }
        else if (label == "physics-discrete-select")
        {
            launch = &b200_step_discrete_select;
        }
        else if (i >= model_begin && i < model_end)
        {
            // every discrete model: one launch over the per-model track lists
            if (!have_interact)
                launch = &b200_step_interact;
            else
                role = B200Role::merged;
            have_interact = true;
        }
        else if (label == "geo-boundary")
        {
            launch = &b200_step_boundary;
        }
        else if (label == "tracking-cut")
        {
            launch = &b200_step_tracking_cut;
        }
        else if (label.rfind("step-gather-", 0) == 0)
        {
            // pre-step gather is part of pre-step; post-step gather + SimpleCalo are one launch
            if (order == Order::user_post && !have_tally)
            {
                launch = &b200_step_tally;
                have_tally = true;
            }
            else
            {
                role = B200Role::merged;
            }
        }
        else if (label == "extend-from-secondaries")
        {
            role = B200Role::end_iteration;
            launch = &b200_step_extend_from_secondaries;
        }
        else
        {
            CELER_VALIDATE(false, << "no B200 kernel for step action '" << label << "'");
        }
        auto action = std::make_shared<B200StepAction>(
            id, label, descr, order, role, launch, problem, state_id);
        if (role == B200Role::begin_iteration)
            result.primaries = action;
        result.actions->insert(std::shared_ptr<B200StepAction const>(action));
    }
    CELER_VALIDATE(result.primaries, << "primary generator was not added to the stepping loop");
    CELER_ENSURE(result.actions->num_actions() == ref.num_actions());
    return result;
}

//---------------------------------------------------------------------------//
B200StepperAdapter::B200StepperAdapter(Input input, std::shared_ptr<B200Problem const> problem)
    : params_(std::move(input.params)), problem_(std::move(problem))
{
    CELER_EXPECT(params_ && problem_);
    // The SoA track state is auxiliary state of the reference's CoreState: register its
    // params once per problem (steppers of other streams share it)
    auto& aux = *params_->aux_reg();
    state_id_ = aux.find(aux_label);
    if (!state_id_)
    {
        state_id_ = aux.next_id();
        aux.insert(std::make_shared<B200AuxParams>(state_id_, problem_));
    }
    registry_ = make_b200_registry(*params_, problem_, state_id_);

    ActionSequenceT::Options opts;
    opts.action_times = input.action_times;
    actions_ = std::make_shared<ActionSequenceT>(*registry_.actions, opts);

    // Create state, including aux data
    state_ = std::make_shared<celeritas::CoreState<MemSpace::device>>(
        *params_, input.stream_id, input.num_track_slots);
    actions_->begin_run(*params_, *state_);
}

B200StepperAdapter::~B200StepperAdapter() = default;

B200AuxState& B200StepperAdapter::b200_state() const
{
    return celeritas::get<B200AuxState>(state_->aux(), state_id_);
}

void B200StepperAdapter::warm_up()
{
    CELER_VALIDATE(state_->counters().num_active == 0,
                   << "cannot warm up when state has active tracks");
    state_->warming_up(true);
    try
    {
        actions_->step(*params_, *state_);
    }
    catch (...)
    {
        state_->warming_up(false);
        throw;
    }
    state_->warming_up(false);
    CELER_ENSURE(state_->counters().num_active == 0);
}

auto B200StepperAdapter::operator()() -> result_type
{
    auto& counters = state_->counters();
    counters.num_generated = 0;
    actions_->step(*params_, *state_);

    result_type result;
    result.generated = counters.num_generated;
    result.active = counters.num_active;
    result.alive = counters.num_alive;
    result.queued = counters.num_initializers;
    return result;
}

auto B200StepperAdapter::operator()(SpanConstPrimary primaries) -> result_type
{
    CELER_EXPECT(!primaries.empty());
    auto max_id = std::max_element(primaries.begin(),
                                   primaries.end(),
                                   [](celeritas::Primary const& left, celeritas::Primary const& right) {
                                       return left.event_id < right.event_id;
                                   });
    CELER_VALIDATE(max_id->event_id < params_->init()->max_events(),
                   << "event number " << max_id->event_id.unchecked_get()
                   << " exceeds max_events=" << params_->init()->max_events());
    registry_.primaries->insert(*state_, primaries);
    return (*this)();
}

void B200StepperAdapter::kill_active()
{
    check(b200_stepper_kill_active(b200_state().handle()), "kill_active");
}

void B200StepperAdapter::reseed(celeritas::UniqueEventId event_id)
{
    check(b200_stepper_reseed(b200_state().handle(), event_id.get()), "reseed");
}

std::uint64_t B200StepperAdapter::launch_count() const
{
    return b200_stepper_launch_count(b200_state().handle());
}
}  // namespace celeritas_b200_adapter
