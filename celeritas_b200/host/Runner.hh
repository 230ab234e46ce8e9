//---------------------------------------------------------------------------//
// celer-sim front end: run a celer-sim JSON input on the B200 track loop.
//
// Mirrors the reference's standalone app for the part that drives the hot path:
//   RunnerInput  (/root/reference/app/celer-sim/RunnerInput.hh:40-140,
//                 RunnerInputIO.json.cc:40-139): same keys, defaults and checks
//   Runner       (app/celer-sim/Runner.cc:123-330): events from primary_options,
//                 per-stream transporters, warm-up, action times
//   run()        (app/celer-sim/celer-sim.cc:74-143): merged / per-event loop
//   RunnerOutput (app/celer-sim/RunnerOutput.cc:37-115): the "result"/"runner" JSON
//
// The problem definition (geometry + physics tables) is not built here: it comes from
// a problem image ("image_file") exported by the reference-side adapter. Options that
// change the tables (physics_options, brem_combined, step_limiter, ...) are therefore
// fixed at export time; options that only affect the run are applied here.
//---------------------------------------------------------------------------//
#pragma once

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "PrimaryGenerator.hh"
#include "Transporter.hh"

namespace celeritas_b200
{
struct RunnerInput
{
    std::string image_file;     //!< B200 problem image (extension)
    std::string geometry_file;  //!< recorded only
    std::string physics_file;   //!< recorded only
    std::string event_file;     //!< HepMC3/ROOT input: not supported
    PrimaryGeneratorOptions primary_options;
    std::vector<std::string> simple_calo;
    bool action_diagnostic{false};
    bool step_diagnostic{false};
    int step_diagnostic_bins{1000};
    bool write_track_counts{true};
    bool write_step_times{true};
    unsigned int seed{0};
    uint32_t num_track_slots{0};
    uint64_t max_steps{static_cast<uint32_t>(-1)};
    uint32_t initializer_capacity{0};
    double secondary_stack_factor{0};
    bool use_device{false};
    bool action_times{false};
    bool merge_events{false};
    bool default_stream{false};
    bool warm_up{false};
    double field[3]{0, 0, 0};
    bool has_field_key{false};
    double step_limiter{0};
    bool brem_combined{false};
    std::string track_order{"none"};
    std::string base_dir;  //!< directory relative paths are resolved against

    //! Parse + validate (throws std::runtime_error)
    static RunnerInput from_json_string(std::string const& text);
    std::string to_json_string() const;
};

struct SimulationResult
{
    double total_time{0};
    double setup_time{0};
    double warmup_time{0};
    std::map<std::string, double> action_times;
    std::vector<TransporterResult> events;
    uint32_t num_streams{1};
};

//! Run the input and return the full celer-sim style JSON report
std::string celer_sim_run(std::string const& input_json);
}  // namespace celeritas_b200
