//---------------------------------------------------------------------------//
// Primary generation from celer-sim "primary_options" (see PrimaryGenerator.hh).
//---------------------------------------------------------------------------//
#include "PrimaryGenerator.hh"

#include <cmath>
#include <sstream>
#include <stdexcept>

#include <nlohmann/json.hpp>

namespace celeritas_b200
{
namespace
{
using json = nlohmann::json;

DistributionSelection selection_from_string(std::string const& s)
{
    for (int i = 0; i < static_cast<int>(DistributionSelection::size_); ++i)
    {
        auto v = static_cast<DistributionSelection>(i);
        if (s == to_cstring(v))
            return v;
    }
    throw std::runtime_error("invalid distribution type '" + s + "'");
}

DistributionOptions distribution_from_json(json const& j, bool scalar_ok)
{
    DistributionOptions opts;
    if (j.is_object())
    {
        opts.distribution = selection_from_string(j.at("distribution").get<std::string>());
        if (j.contains("params"))
            j.at("params").get_to(opts.params);
    }
    else
    {
        // Bare value(s): a delta distribution
        opts.distribution = DistributionSelection::delta;
        if (scalar_ok)
            opts.params = {j.get<double>()};
        else
            j.get_to(opts.params);
    }
    return opts;
}

json distribution_to_json(DistributionOptions const& opts)
{
    if (!opts)
        return json::object();
    return json{{"distribution", to_cstring(opts.distribution)}, {"params", opts.params}};
}

// Number of parameters each distribution needs (PrimaryGeneratorOptions.cc:22-50)
void check_params_size(char const* sampler, std::size_t dimension, DistributionOptions const& o)
{
    std::size_t required = 0;
    switch (o.distribution)
    {
        case DistributionSelection::delta: required = dimension; break;
        case DistributionSelection::isotropic: required = 0; break;
        case DistributionSelection::box: required = 2 * dimension; break;
        default: throw std::runtime_error(std::string("unset distribution for ") + sampler);
    }
    if (o.params.size() != required)
    {
        std::ostringstream os;
        os << sampler << " input parameters have " << o.params.size() << " elements but the '"
           << to_cstring(o.distribution) << "' distribution needs exactly " << required;
        throw std::runtime_error(os.str());
    }
}

void check_allowed(char const* sampler,
                   DistributionOptions const& o,
                   std::initializer_list<DistributionSelection> allowed)
{
    for (auto a : allowed)
        if (a == o.distribution)
            return;
    throw std::runtime_error(std::string("invalid distribution type '")
                             + to_cstring(o.distribution) + "' for " + sampler);
}
}  // namespace

char const* to_cstring(DistributionSelection value)
{
    switch (value)
    {
        case DistributionSelection::delta: return "delta";
        case DistributionSelection::isotropic: return "isotropic";
        case DistributionSelection::box: return "box";
        default: return "<invalid>";
    }
}

PrimaryGeneratorOptions PrimaryGeneratorOptions::from_json_string(std::string const& text)
{
    json j = json::parse(text);
    if (auto it = j.find("_format"); it != j.end())
    {
        if (it->get<std::string>() != "primary-generator")
            throw std::runtime_error("invalid format for \"primary-generator\" input: \""
                                     + it->get<std::string>() + "\"");
    }
    if (auto it = j.find("_units"); it != j.end())
    {
        // native unit system of this library: CGS lengths, MeV (as the reference's default)
        if (it->get<std::string>() != "cgs")
            throw std::runtime_error(
                "incompatible unit system in primary-generator JSON file: constructed with "
                + it->get<std::string>() + " units, but current executable requires cgs");
    }
    PrimaryGeneratorOptions opts;
    if (auto it = j.find("seed"); it != j.end())
        it->get_to(opts.seed);
    auto const& pdg = j.at("pdg");
    if (pdg.is_array())
        pdg.get_to(opts.pdg);
    else
        opts.pdg = {pdg.get<int>()};
    for (int p : opts.pdg)
        if (p == 0)
            throw std::runtime_error("invalid PDG number 0");
    j.at("num_events").get_to(opts.num_events);
    j.at("primaries_per_event").get_to(opts.primaries_per_event);
    opts.energy = distribution_from_json(j.at("energy"), true);
    opts.position = distribution_from_json(j.at("position"), false);
    opts.direction = distribution_from_json(j.at("direction"), false);
    return opts;
}

std::string PrimaryGeneratorOptions::to_json_string() const
{
    json j = {{"_format", "primary-generator"},
              {"_units", "cgs"},
              {"seed", seed},
              {"pdg", pdg},
              {"num_events", num_events},
              {"primaries_per_event", primaries_per_event},
              {"energy", distribution_to_json(energy)},
              {"position", distribution_to_json(position)},
              {"direction", distribution_to_json(direction)}};
    return j.dump();
}

//---------------------------------------------------------------------------//
PrimaryGenerator::PrimaryGenerator(PrimaryGeneratorOptions const& o,
                                   std::vector<uint32_t> particle_ids)
    : num_events_(o.num_events)
    , primaries_per_event_(o.primaries_per_event)
    , energy_(o.energy)
    , position_(o.position)
    , direction_(o.direction)
    , particle_id_(std::move(particle_ids))
{
    if (!o)
        throw std::runtime_error("incomplete primary generator options");
    if (particle_id_.size() != o.pdg.size())
        throw std::runtime_error("one particle id is needed per PDG number");
    check_params_size("energy", 1, energy_);
    check_allowed("energy", energy_, {DistributionSelection::delta});
    check_params_size("position", 3, position_);
    check_allowed("position", position_, {DistributionSelection::delta, DistributionSelection::box});
    check_params_size("direction", 3, direction_);
    check_allowed(
        "direction", direction_, {DistributionSelection::delta, DistributionSelection::isotropic});
    rng_.seed(o.seed);
}

// UniformRealDistribution over a std engine
// (random/distribution/UniformRealDistribution.hh:71-77, GenerateCanonical.hh:84-90)
double PrimaryGenerator::uniform(double a, double b)
{
    double const delta = b - a;
    return std::fma(delta, std::generate_canonical<double, 53>(rng_), a);
}

std::vector<B200Primary> PrimaryGenerator::operator()()
{
    if (event_count_ == num_events_)
        return {};
    std::vector<B200Primary> result(primaries_per_event_);
    for (uint32_t i = 0; i < primaries_per_event_; ++i)
    {
        B200Primary& p = result[i];
        p.particle_id = particle_id_[i % particle_id_.size()];
        p.energy = energy_.params[0];
        if (position_.distribution == DistributionSelection::delta)
        {
            for (int k = 0; k < 3; ++k)
                p.pos[k] = position_.params[k];
        }
        else
        {
            // UniformBoxDistribution.hh:66-76: x, y, z in turn
            for (int k = 0; k < 3; ++k)
                p.pos[k] = this->uniform(position_.params[k], position_.params[3 + k]);
        }
        if (direction_.distribution == DistributionSelection::delta)
        {
            for (int k = 0; k < 3; ++k)
                p.dir[k] = direction_.params[k];
        }
        else
        {
            // IsotropicDistribution.hh:58-65, from_spherical (corecel/math/ArrayUtils.hh)
            double const costheta = this->uniform(-1, 1);
            double const phi = this->uniform(0, 2 * 3.14159265358979323846);
            double const sintheta = std::sqrt(1 - costheta * costheta);
            p.dir[0] = sintheta * std::cos(phi);
            p.dir[1] = sintheta * std::sin(phi);
            p.dir[2] = costheta;
        }
        p.time = 0;
        p.event_id = event_count_;
    }
    ++event_count_;
    return result;
}
}  // namespace celeritas_b200
