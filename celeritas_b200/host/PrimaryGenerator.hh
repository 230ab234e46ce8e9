//---------------------------------------------------------------------------//
// Primary generation from celer-sim "primary_options".
//
// Same behaviour as the reference's PrimaryGenerator
// (/root/reference/src/celeritas/phys/PrimaryGenerator.cc:30-110) and its option
// parsing (phys/PrimaryGeneratorOptions.cc:20-140,
// phys/PrimaryGeneratorOptionsIO.json.cc:67-134): one std::mt19937 seeded with
// `seed`; for each primary the energy, position and direction are sampled in that
// order; primary i of an event has particle pdg[i % pdg.size()]; time 0.
// Host code (this is the event *source*, not the track loop).
//---------------------------------------------------------------------------//
#pragma once

#include <cstdint>
#include <random>
#include <string>
#include <vector>

#include "../../include/celeritas_b200.h"

namespace celeritas_b200
{
enum class DistributionSelection
{
    delta,
    isotropic,
    box,
    size_
};

struct DistributionOptions
{
    DistributionSelection distribution{DistributionSelection::size_};
    std::vector<double> params;
    explicit operator bool() const { return distribution != DistributionSelection::size_; }
};

struct PrimaryGeneratorOptions
{
    unsigned int seed{0};
    std::vector<int> pdg;
    uint32_t num_events{0};
    uint32_t primaries_per_event{0};
    DistributionOptions energy;
    DistributionOptions position;
    DistributionOptions direction;

    explicit operator bool() const
    {
        return !pdg.empty() && num_events > 0 && primaries_per_event > 0 && energy && position
               && direction;
    }
    //! Parse the JSON object (throws std::runtime_error with the reference's messages)
    static PrimaryGeneratorOptions from_json_string(std::string const& text);
    std::string to_json_string() const;
};

char const* to_cstring(DistributionSelection value);

class PrimaryGenerator
{
  public:
    //! particle_ids[i] is the particle id of options.pdg[i]
    PrimaryGenerator(PrimaryGeneratorOptions const& options, std::vector<uint32_t> particle_ids);

    //! Primaries of the next event (empty when all events have been generated)
    std::vector<B200Primary> operator()();
    uint32_t num_events() const { return num_events_; }
    uint32_t primaries_per_event() const { return primaries_per_event_; }

  private:
    uint32_t num_events_;
    uint32_t primaries_per_event_;
    DistributionOptions energy_, position_, direction_;
    std::vector<uint32_t> particle_id_;
    uint32_t event_count_{0};
    std::mt19937 rng_;

    double uniform(double a, double b);
};
}  // namespace celeritas_b200
