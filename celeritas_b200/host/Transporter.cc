//---------------------------------------------------------------------------//
// Transport loop over whole events (see Transporter.hh).
//---------------------------------------------------------------------------//
#include "Transporter.hh"

#include <algorithm>
#include <chrono>

namespace celeritas_b200
{
Transporter::Transporter(std::shared_ptr<Stepper> stepper, TransporterInput input)
    : stepper_(std::move(stepper)), input_(input)
{
    if (!stepper_)
        throw std::runtime_error("Transporter requires a stepper");
}

TransporterResult Transporter::operator()(B200Primary const* primaries, uint32_t n)
{
    using Clock = std::chrono::steady_clock;
    TransporterResult result;
    Stepper& step = *stepper_;
    auto append = [&](StepperResult const& c, double seconds) {
        if (input_.store_track_counts)
        {
            result.generated.push_back(c.generated);
            result.initializers.push_back(c.queued);
            result.active.push_back(c.active);
            result.alive.push_back(c.alive);
        }
        if (input_.store_step_times)
            result.step_times.push_back(seconds);
        ++result.num_step_iterations;
        result.num_steps += c.active;
        result.max_queued = std::max<uint64_t>(result.max_queued, c.queued);
    };

    uint64_t remaining_steps = input_.max_steps;
    auto const start = Clock::now();
    StepperResult counts = step(primaries, n);
    append(counts, std::chrono::duration<double>(Clock::now() - start).count());
    std::vector<StepperResult> batch;
    std::vector<double> batch_seconds;
    while (counts)
    {
        // The reference's loop, `if (max_steps && --remaining_steps == 0) break;` before
        // every further iteration: at most remaining_steps - 1 more may run
        if (input_.max_steps != 0 && remaining_steps <= 1)
        {
            // Exceeded the step count: abort the transport loop
            break;
        }
        uint64_t const budget = input_.max_steps != 0 ? remaining_steps - 1 : 0xffffffffull;
        batch.clear();
        batch_seconds.clear();
        // Stepper::advance is operator()() repeated; iterations with few tracks run inside
        // the device-resident loop, without a host round trip in between
        uint32_t const taken
            = step.advance(static_cast<uint32_t>(std::min<uint64_t>(budget, 0xffffffffull)),
                           &batch,
                           input_.store_step_times ? &batch_seconds : nullptr);
        if (input_.max_steps != 0)
            remaining_steps -= taken;
        for (uint32_t i = 0; i < taken; ++i)
            append(batch[i], input_.store_step_times ? batch_seconds[i] : 0.0);
        counts = batch.back();
    }
    result.num_tracks = step.state().num_tracks();
    result.num_aborted = uint64_t(counts.alive) + counts.queued;
    result.num_track_slots = step.state().size();
    if (result.num_aborted > 0)
        step.reset_state();
    return result;
}

void Transporter::accum_action_times(std::map<std::string, double>* result) const
{
    ActionSequence const& seq = stepper_->actions();
    if (!seq.action_times())
        return;
    auto const& actions = seq.actions();
    auto const& times = seq.accum_time();
    for (size_t i = 0; i < actions.size() && i < times.size(); ++i)
        (*result)[actions[i]->label()] += times[i];
}
}  // namespace celeritas_b200
