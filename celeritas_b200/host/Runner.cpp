//---------------------------------------------------------------------------//
// celer-sim front end (see Runner.hh).
//---------------------------------------------------------------------------//
#include "Runner.hh"
#include "OrangeBuilder.hh"
#include "RootImport.hh"

#include <fstream>

#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>

#include <nlohmann/json.hpp>

#include "CoreParams.hh"
#include "Handles.hh"
#include "Stepper.hh"

namespace celeritas_b200
{
namespace
{
using json = nlohmann::json;

template<class T>
void load_option(json const& j, char const* key, T& value)
{
    if (auto it = j.find(key); it != j.end())
        it->get_to(value);
}

template<class T>
void load_required(json const& j, char const* key, T& value)
{
    auto it = j.find(key);
    if (it == j.end())
        throw std::runtime_error(std::string("missing required key '") + key + "'");
    it->get_to(value);
}

template<class T>
void load_deprecated(json const& j, char const* old_key, char const* new_key, T& value)
{
    if (auto it = j.find(old_key); it != j.end())
    {
        (void)new_key;  // the reference warns and keeps going
        it->get_to(value);
    }
}

std::string resolve(std::string const& base, std::string const& path)
{
    if (path.empty() || path[0] == '/' || base.empty())
        return path;
    return base + "/" + path;
}

class Stopwatch
{
  public:
    Stopwatch() : start_(std::chrono::steady_clock::now()) {}
    double operator()() const
    {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - start_).count();
    }

  private:
    std::chrono::steady_clock::time_point start_;
};

json null_if_empty(json&& array)
{
    return array.empty() ? json(nullptr) : std::move(array);
}

// app/celer-sim/RunnerOutput.cc:37-115
json runner_output(SimulationResult const& result)
{
    json active = json::array(), alive = json::array(), generated = json::array(),
         initializers = json::array(), num_track_slots = json::array(),
         num_step_iterations = json::array(), num_tracks = json::array(),
         num_steps = json::array(), num_aborted = json::array(), max_queued = json::array(),
         step_times = json::array();
    for (TransporterResult const& event : result.events)
    {
        if (!event.active.empty())
        {
            active.push_back(event.active);
            alive.push_back(event.alive);
            generated.push_back(event.generated);
            initializers.push_back(event.initializers);
        }
        num_track_slots.push_back(event.num_track_slots);
        num_step_iterations.push_back(event.num_step_iterations);
        num_tracks.push_back(event.num_tracks);
        num_steps.push_back(event.num_steps);
        num_aborted.push_back(event.num_aborted);
        max_queued.push_back(event.max_queued);
        if (!event.step_times.empty())
            step_times.push_back(event.step_times);
    }
    json times = {{"steps", null_if_empty(std::move(step_times))},
                  {"actions", result.action_times},
                  {"total", result.total_time},
                  {"setup", result.setup_time},
                  {"warmup", result.warmup_time}};
    return json{{"_index", {"event", "step"}},
                {"active", null_if_empty(std::move(active))},
                {"alive", null_if_empty(std::move(alive))},
                {"generated", null_if_empty(std::move(generated))},
                {"initializers", null_if_empty(std::move(initializers))},
                {"num_track_slots", std::move(num_track_slots)},
                {"num_step_iterations", std::move(num_step_iterations)},
                {"num_tracks", std::move(num_tracks)},
                {"num_steps", std::move(num_steps)},
                {"num_aborted", std::move(num_aborted)},
                {"max_queued", std::move(max_queued)},
                {"num_streams", result.num_streams},
                {"time", std::move(times)}};
}
}  // namespace

//---------------------------------------------------------------------------//
RunnerInput RunnerInput::from_json_string(std::string const& text)
{
    json j = json::parse(text);
    RunnerInput v;
    if (auto it = j.find("_format"); it != j.end())
    {
        if (it->get<std::string>() != "celer-sim")
            throw std::runtime_error("invalid format for \"celer-sim\" input: \""
                                     + it->get<std::string>() + "\"");
    }
    load_deprecated(j, "hepmc3_filename", "event_file", v.event_file);
    load_deprecated(j, "event_filename", "event_file", v.event_file);
    load_deprecated(j, "geometry_filename", "geometry_file", v.geometry_file);
    load_deprecated(j, "physics_filename", "physics_file", v.physics_file);
    if (v.geometry_file.empty())
        load_required(j, "geometry_file", v.geometry_file);
    load_option(j, "physics_file", v.physics_file);
    load_option(j, "event_file", v.event_file);
    load_option(j, "image_file", v.image_file);
    load_option(j, "base_dir", v.base_dir);
    for (char const* key : {"primary_gen_options", "primary_options"})
    {
        if (auto it = j.find(key); it != j.end())
            v.primary_options = PrimaryGeneratorOptions::from_json_string(it->dump());
    }
    load_deprecated(j, "step_diagnostic_maxsteps", "step_diagnostic_bins", v.step_diagnostic_bins);
    if (auto it = j.find("simple_calo"); it != j.end())
    {
        // Labels are "name" or "name@ext" strings
        v.simple_calo.clear();
        for (auto const& item : *it)
            v.simple_calo.push_back(item.get<std::string>());
    }
    load_option(j, "action_diagnostic", v.action_diagnostic);
    load_option(j, "step_diagnostic", v.step_diagnostic);
    load_option(j, "step_diagnostic_bins", v.step_diagnostic_bins);
    load_option(j, "write_track_counts", v.write_track_counts);
    load_option(j, "write_step_times", v.write_step_times);
    load_deprecated(j, "max_num_tracks", "num_track_slots", v.num_track_slots);
    load_deprecated(j, "sync", "action_times", v.action_times);
    load_option(j, "seed", v.seed);
    load_option(j, "num_track_slots", v.num_track_slots);
    load_option(j, "max_steps", v.max_steps);
    load_required(j, "initializer_capacity", v.initializer_capacity);
    load_required(j, "secondary_stack_factor", v.secondary_stack_factor);
    load_required(j, "use_device", v.use_device);
    load_option(j, "action_times", v.action_times);
    load_option(j, "merge_events", v.merge_events);
    load_option(j, "default_stream", v.default_stream);
    if (auto it = j.find("warm_up"); it != j.end())
        it->get_to(v.warm_up);
    else if (v.use_device)
        v.warm_up = true;
    for (char const* key : {"mag_field", "field"})
    {
        if (auto it = j.find(key); it != j.end())
        {
            auto f = it->get<std::vector<double>>();
            if (f.size() != 3)
                throw std::runtime_error("'field' needs three components");
            std::copy(f.begin(), f.end(), v.field);
            v.has_field_key = true;
        }
    }
    load_option(j, "step_limiter", v.step_limiter);
    load_option(j, "brem_combined", v.brem_combined);
    if (auto it = j.find("track_order"); it != j.end())
        it->get_to(v.track_order);
    else if (v.use_device)
        v.track_order = "init_charge";

    bool const no_field = v.field[0] == 0 && v.field[1] == 0 && v.field[2] == 0;
    if (v.event_file.empty() != static_cast<bool>(v.primary_options))
        throw std::runtime_error(
            "either a event filename or options to generate primaries must be provided (but "
            "not both)");
    if (no_field && j.contains("field_options"))
        throw std::runtime_error("'field_options' cannot be specified without providing 'field'");
    for (char const* key : {"mctruth_file", "mctruth_filter", "slot_diagnostic_prefix"})
    {
        if (j.contains(key) && !j.at(key).empty())
            throw std::runtime_error(std::string("'") + key
                                     + "' output is outside the scope of this library");
    }
    return v;
}

std::string RunnerInput::to_json_string() const
{
    json j = {{"_format", "celer-sim"},
              {"image_file", image_file},
              {"geometry_file", geometry_file},
              {"physics_file", physics_file},
              {"event_file", event_file},
              {"simple_calo", simple_calo},
              {"action_diagnostic", action_diagnostic},
              {"step_diagnostic", step_diagnostic},
              {"step_diagnostic_bins", step_diagnostic_bins},
              {"write_track_counts", write_track_counts},
              {"write_step_times", write_step_times},
              {"seed", seed},
              {"num_track_slots", num_track_slots},
              {"max_steps", max_steps},
              {"initializer_capacity", initializer_capacity},
              {"secondary_stack_factor", secondary_stack_factor},
              {"use_device", use_device},
              {"action_times", action_times},
              {"merge_events", merge_events},
              {"default_stream", default_stream},
              {"warm_up", warm_up},
              {"field", {field[0], field[1], field[2]}},
              {"step_limiter", step_limiter},
              {"brem_combined", brem_combined},
              {"track_order", track_order}};
    if (primary_options)
        j["primary_options"] = json::parse(primary_options.to_json_string());
    return j.dump();
}

//---------------------------------------------------------------------------//
std::string celer_sim_run(std::string const& input_json)
{
    RunnerInput inp = RunnerInput::from_json_string(input_json);
    json output;
    output["input"] = json::parse(inp.to_json_string());

    //// Runner construction (app/celer-sim/Runner.cc:123-174) ////
    Stopwatch get_setup_time;
    if (!inp.use_device)
        throw std::runtime_error(
            "use_device=false: this library has no host track loop (run the reference)");
    if (inp.image_file.empty())
        throw std::runtime_error(
            "missing 'image_file': export the problem with the reference-side adapter "
            "(INTEGRATION.md)");
    if (!inp.event_file.empty())
        throw std::runtime_error(
            "event_file input (HepMC3/ROOT) is not supported: use primary_options");
    if (inp.num_track_slots == 0)
        throw std::runtime_error("nonpositive num_track_slots=0");
    if (inp.max_steps == 0)
        throw std::runtime_error("nonpositive max_steps=0");
    if (!(inp.secondary_stack_factor > 0))
        throw std::runtime_error("nonpositive secondary_stack_factor");
    if (inp.step_diagnostic && inp.step_diagnostic_bins <= 0)
        throw std::runtime_error("nonpositive step diagnostic 'max' bin");
    uint32_t track_order = b200::ORDER_NONE;
    if (inp.track_order == "none" || inp.track_order == "unsorted")
        track_order = b200::ORDER_NONE;
    else if (inp.track_order == "init_charge")
        track_order = b200::ORDER_INIT_CHARGE;
    else if (inp.track_order == "reindex_status")
        track_order = b200::ORDER_REINDEX_STATUS;
    else if (inp.track_order == "reindex_particle_type")
        track_order = b200::ORDER_REINDEX_PARTICLE_TYPE;
    else if (inp.track_order == "reindex_along_step_action")
        track_order = b200::ORDER_REINDEX_ALONG_STEP_ACTION;
    else if (inp.track_order == "reindex_step_limit_action")
        track_order = b200::ORDER_REINDEX_STEP_LIMIT_ACTION;
    else if (inp.track_order == "reindex_both_action")
        track_order = b200::ORDER_REINDEX_BOTH_ACTION;
    else if (inp.track_order == "reindex_shuffle")
        track_order = b200::ORDER_REINDEX_SHUFFLE;  // same per-slot results as "none"
    else
        throw std::runtime_error("track_order '" + inp.track_order + "' is not supported");

    // The problem image carries materials, physics tables and the action table; the GEOMETRY
    // is built here from `geometry_file` when that is an ORANGE JSON file that exists
    // (host/OrangeBuilder.cpp: the reference's OrangeParams construction), and must be the
    // geometry the image's volume -> material map was exported for.
    b200::Image image = b200::Image::read(resolve(inp.base_dir, inp.image_file));
    {
        std::string const geo_path = resolve(inp.base_dir, inp.geometry_file);
        bool const is_org_json = geo_path.size() > 9
                                 && geo_path.compare(geo_path.size() - 9, 9, ".org.json") == 0;
        if (is_org_json && std::ifstream(geo_path).good())
        {
            b200::Image const geo = build_orange_image(geo_path);
            if (geo.get_string("geo.volume_labels") != image.get_string("geo.volume_labels"))
                throw std::runtime_error(
                    "geometry_file '" + inp.geometry_file
                    + "' is not the geometry the problem image was exported for (volume "
                      "labels differ)");
            for (auto const& kv : geo.entries())
                if (kv.first.rfind("geo.", 0) == 0)
                    image.put_entry(kv.first, kv.second);
        }
    }
    // `physics_file`: a reference physics export (ROOT, decoded by host/RootImport.cpp) or
    // its JSON form is read when it exists and must describe the particles and elements the
    // image's tables were built for (the tables themselves come from the image)
    std::string physics_check;
    {
        std::string const phys_path = resolve(inp.base_dir, inp.physics_file);
        auto ends_with = [&](char const* ext) {
            size_t const n = std::strlen(ext);
            return phys_path.size() > n && phys_path.compare(phys_path.size() - n, n, ext) == 0;
        };
        if ((ends_with(".root") || ends_with(".json")) && !inp.physics_file.empty()
            && std::ifstream(phys_path).good())
        {
            json data;
            if (ends_with(".root"))
                data = json::parse(b200::import_root_to_json(phys_path));
            else
                data = json::parse(std::ifstream(phys_path));
            if (!data.contains("particles") || !data.contains("elements"))
                throw std::runtime_error("physics_file '" + inp.physics_file
                                         + "' is not a celeritas::ImportData export");
            auto const pdg = image.get<uint32_t>("particle.pdg");
            auto const mass = image.get<double>("particle.mass");
            for (size_t i = 0; i < pdg.size(); ++i)
            {
                bool found = false;
                for (auto const& p : data.at("particles"))
                {
                    double const m = p.at("mass").get<double>();
                    found = found
                            || (p.at("pdg").get<int>() == static_cast<int32_t>(pdg[i])
                                && std::fabs(m - mass[i]) <= 1e-9 * std::fabs(m));
                }
                if (!found)
                    throw std::runtime_error(
                        "physics_file '" + inp.physics_file
                        + "' is not the physics the problem image was exported for (particle "
                        + std::to_string(static_cast<int32_t>(pdg[i])) + ")");
            }
            for (uint32_t z : image.get<uint32_t>("mat.element_z"))
            {
                bool found = false;
                for (auto const& e : data.at("elements"))
                    found = found || e.at("atomic_number").get<uint32_t>() == z;
                if (!found)
                    throw std::runtime_error(
                        "physics_file '" + inp.physics_file
                        + "' is not the physics the problem image was exported for (element Z="
                        + std::to_string(z) + ")");
            }
            physics_check = std::to_string(data.at("particles").size()) + " particles, "
                            + std::to_string(data.at("elements").size()) + " elements, "
                            + std::to_string(data.at("processes").size()) + " processes";
        }
    }
    std::shared_ptr<CoreParams> params = CoreParams::from_image(image);
    bool const no_field = inp.field[0] == 0 && inp.field[1] == 0 && inp.field[2] == 0;
    if (!no_field)
        params->uniform_field_tesla(inp.field);
    else if (inp.has_field_key && params->has_uniform_field())
        throw std::runtime_error("the problem image has a uniform-field along-step action: "
                                 "'field' must be nonzero");
    if (!inp.simple_calo.empty() && inp.simple_calo != params->detector_volumes())
        throw std::runtime_error(
            "'simple_calo' differs from the detector volumes the problem image was exported "
            "with");
    params->rng_seed(inp.seed);
    // Run options celer-sim applies when it builds the problem (Runner.cc:323-440). The
    // track order is a property of the loop, not of the tables: it is applied here. The
    // step limiter and the bremsstrahlung model choice are baked into the exported physics
    // tables: they must agree with the image.
    params->track_order(track_order);
    {
        double const image_limiter = params->view().phys.fixed_step_limiter;
        if (inp.step_limiter != image_limiter)
            throw std::runtime_error("'step_limiter' (" + std::to_string(inp.step_limiter)
                                     + ") differs from the value the problem image was "
                                       "exported with ("
                                     + std::to_string(image_limiter) + ")");
        if (inp.brem_combined != params->has_action("brems-combined"))
            throw std::runtime_error(
                std::string("'brem_combined' is ") + (inp.brem_combined ? "true" : "false")
                + " but the problem image was exported "
                + (inp.brem_combined ? "without" : "with") + " the combined bremsstrahlung model");
    }

    // Events (Runner::build_events, Runner.cc:452-500)
    std::vector<uint32_t> particle_ids;
    for (int pdg : inp.primary_options.pdg)
    {
        uint32_t id = params->find_particle(pdg);
        if (id == b200::INVALID)
            throw std::runtime_error("PDG " + std::to_string(pdg)
                                     + " is not a particle of this problem");
        particle_ids.push_back(id);
    }
    PrimaryGenerator generate(inp.primary_options, particle_ids);
    std::vector<std::vector<B200Primary>> events;
    if (inp.merge_events)
        events.resize(1);
    for (auto event = generate(); !event.empty(); event = generate())
    {
        if (inp.merge_events)
            events.front().insert(events.front().end(), event.begin(), event.end());
        else
            events.push_back(std::move(event));
    }
    // One stream per process (one process per GPU)
    params->init_capacity(inp.initializer_capacity);
    params->max_events(generate.num_events());

    StepperInput sinp;
    sinp.params = params;
    sinp.stream_id = 0;
    sinp.num_track_slots = inp.num_track_slots;
    sinp.action_times = inp.action_times;
    sinp.actions.action_diagnostic = inp.action_diagnostic;
    sinp.actions.step_diagnostic_bins = inp.step_diagnostic ? inp.step_diagnostic_bins : 0;
    auto stepper = std::make_shared<Stepper>(std::move(sinp));
    TransporterInput tinp;
    tinp.max_steps = inp.max_steps == static_cast<uint32_t>(-1) ? 0 : inp.max_steps;
    tinp.store_track_counts = inp.write_track_counts;
    tinp.store_step_times = inp.write_step_times;
    Transporter transport(stepper, tinp);

    SimulationResult result;
    result.setup_time = get_setup_time();
    result.events.resize(events.size());
    result.num_streams = 1;

    //// run() (app/celer-sim/celer-sim.cc:106-140) ////
    if (inp.warm_up)
    {
        Stopwatch get_warmup_time;
        stepper->warm_up();
        result.warmup_time = get_warmup_time();
    }
    Stopwatch get_transport_time;
    for (size_t e = 0; e < events.size(); ++e)
        result.events[e] = transport(events[e].data(), events[e].size());
    transport.accum_action_times(&result.action_times);
    result.total_time = get_transport_time();

    //// Output ////
    json res;
    res["runner"] = runner_output(result);
    CoreState& state = stepper->state();
    uint32_t const num_particles = params->particle_names().size();
    if (stepper->actions().action_diagnostic())
    {
        // user/ActionDiagnostic.cc:120-130: counts[particle][action]
        uint32_t const nb = stepper->actions().labels().size();
        std::vector<uint32_t> counts(size_t(nb) * num_particles);
        state.diagnostic_get(false, counts.data());
        json per_particle = json::array();
        json nonzero = json::object();
        for (uint32_t p = 0; p < num_particles; ++p)
        {
            per_particle.push_back(
                std::vector<uint32_t>(counts.begin() + size_t(p) * nb, counts.begin() + size_t(p + 1) * nb));
            for (uint32_t a = 0; a < nb; ++a)
            {
                if (uint32_t c = counts[size_t(p) * nb + a])
                    nonzero[stepper->actions().labels()[a] + " " + params->particle_names()[p]] = c;
            }
        }
        res["action-diagnostic"] = {{"actions", std::move(per_particle)},
                                    {"_index", {"particle", "action"}},
                                    {"_nonzero", std::move(nonzero)}};
    }
    if (uint32_t bins = stepper->actions().step_diagnostic_bins())
    {
        // user/StepDiagnostic.cc:103-111: counts[particle][num_steps]
        uint32_t const nb = bins + 2;
        std::vector<uint32_t> counts(size_t(nb) * num_particles);
        state.diagnostic_get(true, counts.data());
        json per_particle = json::array();
        for (uint32_t p = 0; p < num_particles; ++p)
            per_particle.push_back(
                std::vector<uint32_t>(counts.begin() + size_t(p) * nb, counts.begin() + size_t(p + 1) * nb));
        res["step-diagnostic"] = {{"steps", std::move(per_particle)},
                                  {"_index", {"particle", "num_steps"}}};
    }
    if (params->num_detectors() > 0)
    {
        // user/SimpleCalo.cc:163-189
        std::vector<double> edep(params->num_detectors());
        state.calo_get(edep.data());
        std::vector<int> ids;
        for (std::string const& name : params->detector_volumes())
        {
            auto const& labels = params->volume_labels();
            ids.push_back(std::find(labels.begin(), labels.end(), name) - labels.begin());
        }
        res["simple_calo"] = {{"volume_ids", ids},
                              {"volume_labels", params->detector_volumes()},
                              {"energy_deposition", edep},
                              {"_units", {{"energy_deposition", "MeV"}}}};
    }
    output["result"] = std::move(res);
    json labels = json::array();
    for (auto const& a : stepper->actions().actions())
        labels.push_back(a->label());
    output["internal"] = {{"actions", {{"label", stepper->actions().labels()}}},
                          {"step_actions", std::move(labels)},
                          {"kernel_launches", b200_launch_count()},
                          {"device_bytes", {{"params", params->device_bytes()},
                                            {"state", state.device_bytes()}}}};
    output["system"] = {{"device", "B200 (sm_100a)"}, {"library", "celeritas_b200"}};
    if (!physics_check.empty())
        output["internal"]["physics_file"] = physics_check;
    return output.dump(1);
}
}  // namespace celeritas_b200

//---------------------------------------------------------------------------//
// C-ABI
//---------------------------------------------------------------------------//
extern "C" int b200_primaries_generate(B200Params const* params,
                                       char const* primary_options_json,
                                       B200Primary* out,
                                       uint64_t capacity,
                                       uint64_t* count,
                                       uint32_t* primaries_per_event)
{
    using namespace celeritas_b200;
    if (!params || !primary_options_json || !count)
        return B200_ERR_INVALID_ARGUMENT;
    try
    {
        auto opts = PrimaryGeneratorOptions::from_json_string(primary_options_json);
        std::vector<uint32_t> ids;
        for (int pdg : opts.pdg)
        {
            uint32_t id = params->params->find_particle(pdg);
            if (id == b200::INVALID)
                throw std::runtime_error("PDG " + std::to_string(pdg)
                                         + " is not a particle of this problem");
            ids.push_back(id);
        }
        PrimaryGenerator generate(opts, ids);
        *count = uint64_t(opts.num_events) * opts.primaries_per_event;
        if (primaries_per_event)
            *primaries_per_event = opts.primaries_per_event;
        if (!out)
            return B200_OK;
        if (capacity < *count)
            return B200_ERR_INVALID_ARGUMENT;
        for (auto event = generate(); !event.empty(); event = generate())
            out = std::copy(event.begin(), event.end(), out);
        return B200_OK;
    }
    catch (std::exception const& e)
    {
        celeritas_b200::set_last_error(e.what());
        return B200_ERR_RUNTIME;
    }
}

extern "C" int b200_celer_sim_run(char const* input_json, char** report)
{
    if (!input_json || !report)
        return B200_ERR_INVALID_ARGUMENT;
    *report = nullptr;
    try
    {
        std::string text = celeritas_b200::celer_sim_run(input_json);
        char* buffer = static_cast<char*>(std::malloc(text.size() + 1));
        if (!buffer)
            throw std::runtime_error("out of memory");
        std::memcpy(buffer, text.c_str(), text.size() + 1);
        *report = buffer;
        return B200_OK;
    }
    catch (celeritas_b200::CudaError const& e)
    {
        celeritas_b200::set_last_error(e.what());
        return e.code;
    }
    catch (std::exception const& e)
    {
        celeritas_b200::set_last_error(e.what());
        return B200_ERR_RUNTIME;
    }
}
