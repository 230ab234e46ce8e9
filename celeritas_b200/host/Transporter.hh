//---------------------------------------------------------------------------//
// Transport all tracks of one event (or several merged events) to completion.
//
// Same loop and the same tallies as the reference's celer-sim Transporter
// (/root/reference/app/celer-sim/Transporter.cc:84-179, Transporter.hh:45-78):
// first iteration with the primaries, then step until no track is alive or queued
// or `max_steps` ITERATIONS have been taken.
//---------------------------------------------------------------------------//
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "Stepper.hh"

namespace celeritas_b200
{
struct TransporterInput
{
    uint64_t max_steps{0};           //!< step iterations per call; 0 = unlimited
    bool store_track_counts{false};  //!< keep per-iteration counts
    bool store_step_times{false};    //!< keep per-iteration wall times
};

struct TransporterResult
{
    // Per-step diagnostics (empty unless requested)
    std::vector<uint32_t> generated;
    std::vector<uint32_t> initializers;
    std::vector<uint32_t> active;
    std::vector<uint32_t> alive;
    std::vector<double> step_times;  //!< [s]

    // Always-on tallies
    uint64_t num_track_slots{0};
    uint64_t num_step_iterations{0};
    uint64_t num_steps{0};
    uint64_t num_aborted{0};
    uint64_t num_tracks{0};
    uint64_t max_queued{0};
};

class Transporter
{
  public:
    Transporter(std::shared_ptr<Stepper> stepper, TransporterInput input);
    //! Transport the primaries and all their secondaries
    TransporterResult operator()(B200Primary const* primaries, uint32_t n);
    //! Add this stepper's accumulated per-action device times, keyed by action label
    void accum_action_times(std::map<std::string, double>* result) const;
    Stepper& stepper() { return *stepper_; }

  private:
    std::shared_ptr<Stepper> stepper_;
    TransporterInput input_;
};
}  // namespace celeritas_b200
