//---------------------------------------------------------------------------//
// TEST INFRASTRUCTURE ONLY. C API over the reference's own host Stepper so
// that Python tests (ctypes) and bench.py's CPU baseline can drive it.
// Reference surface used: Stepper<MemSpace::host>
// (/root/reference/src/celeritas/global/Stepper.hh:82-190).
//---------------------------------------------------------------------------//
#include <chrono>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>
#include <omp.h>
#include <nlohmann/json.hpp>

#include "corecel/io/Logger.hh"
#include "corecel/sys/Device.hh"
#include "corecel/Config.hh"
#if CELERITAS_USE_CUDA
#    include <cuda_runtime_api.h>
#endif
#include "corecel/sys/ActionRegistry.hh"
#include "corecel/sys/Environment.hh"
#include "celeritas/geo/GeoTrackView.hh"
#include "celeritas/global/CoreState.hh"
#include "celeritas/global/Stepper.hh"
#include "celeritas/phys/ParticleParams.hh"
#include "celeritas/phys/Primary.hh"
#include "celeritas/phys/PrimaryGenerator.hh"
#include "celeritas/phys/PrimaryGeneratorOptions.hh"
#include "celeritas/phys/PrimaryGeneratorOptionsIO.json.hh"
#include "celeritas/random/RngEngine.hh"

#include "celeritas/user/DetectorSteps.hh"
#include "celeritas/user/StepData.hh"

#include "Problem.hh"

using namespace celeritas;
using json = nlohmann::json;

namespace
{
thread_local std::string g_last_error;

struct RefStepper
{
    celerref::Problem* problem;
    std::unique_ptr<Stepper<MemSpace::host>> step;
};

//! Same POD layout as B200Primary in include/celeritas_b200.h
struct CPrimary
{
    uint32_t particle_id;
    uint32_t event_id;
    double energy;
    double pos[3];
    double dir[3];
    double time;
};

template<class F>
int guarded(F&& f)
{
    try
    {
        f();
        return 0;
    }
    catch (std::exception const& e)
    {
        g_last_error = e.what();
        return 1;
    }
}

std::vector<Primary> to_primaries(CPrimary const* p, uint32_t n)
{
    std::vector<Primary> result(n);
    for (uint32_t i = 0; i < n; ++i)
    {
        result[i].particle_id = ParticleId{p[i].particle_id};
        result[i].energy = units::MevEnergy{p[i].energy};
        result[i].position = {p[i].pos[0], p[i].pos[1], p[i].pos[2]};
        result[i].direction = {p[i].dir[0], p[i].dir[1], p[i].dir[2]};
        result[i].time = p[i].time;
        result[i].event_id = EventId{p[i].event_id};
    }
    return result;
}
}  // namespace

extern "C" {
//---------------------------------------------------------------------------//
char const* celerref_last_error()
{
    return g_last_error.c_str();
}

void* celerref_problem_create(char const* config_json)
{
    void* result = nullptr;
    guarded([&] {
        // Quiet the reference's logger unless asked
        if (celeritas::getenv("CELER_LOG").empty())
        {
            celeritas::world_logger().level(LogLevel::warning);
            celeritas::self_logger().level(LogLevel::warning);
        }
#if CELERITAS_USE_CUDA
        // The reference's CUDA build: params are mirrored to the device at construction,
        // so the device must be active first (app/celer-sim/celer-sim.cc:185-195)
        if (!celeritas::device())
        {
            celeritas::activate_device();
        }
#endif
        auto p = celerref::build_problem(json::parse(config_json));
        result = p.release();
    });
    return result;
}

void celerref_problem_destroy(void* p)
{
    delete static_cast<celerref::Problem*>(p);
}

int celerref_export_image(void* p, char const* path)
{
    return guarded(
        [&] { celerref::export_image(*static_cast<celerref::Problem*>(p), path); });
}

//! Number of actions, and label of action i
uint32_t celerref_num_actions(void* p)
{
    return static_cast<celerref::Problem*>(p)->core->action_reg()->num_actions();
}

//---------------------------------------------------------------------------//
void* celerref_stepper_create_stream(void* problem, uint32_t num_track_slots, uint32_t stream_id);

void* celerref_stepper_create(void* problem, uint32_t num_track_slots)
{
    return celerref_stepper_create_stream(problem, num_track_slots, 0);
}

// Stepper on a given stream (the problem's max_streams must exceed it): the RNG states
// of a stream are seeded from {seed, stream_id} (random/XorwowRngData.cc:28-58)
void* celerref_stepper_create_stream(void* problem, uint32_t num_track_slots, uint32_t stream_id)
{
    void* result = nullptr;
    guarded([&] {
        auto* p = static_cast<celerref::Problem*>(problem);
        StepperInput inp;
        inp.params = p->core;
        inp.stream_id = StreamId{stream_id};
        inp.num_track_slots = num_track_slots;
        auto s = std::make_unique<RefStepper>();
        s->problem = p;
        s->step = std::make_unique<Stepper<MemSpace::host>>(std::move(inp));
        result = s.release();
    });
    return result;
}

void celerref_stepper_destroy(void* s)
{
    delete static_cast<RefStepper*>(s);
}

//! One step iteration. counts = {generated, queued, active, alive}
int celerref_step(void* stepper,
                  CPrimary const* primaries,
                  uint32_t num_primaries,
                  uint32_t* counts)
{
    return guarded([&] {
        auto* s = static_cast<RefStepper*>(stepper);
        StepperResult r;
        if (num_primaries > 0)
        {
            auto prim = to_primaries(primaries, num_primaries);
            r = (*s->step)(make_span(prim));
        }
        else
        {
            r = (*s->step)();
        }
        counts[0] = r.generated;
        counts[1] = r.queued;
        counts[2] = r.active;
        counts[3] = r.alive;
    });
}

int celerref_reseed(void* stepper, uint64_t event_id)
{
    return guarded([&] {
        static_cast<RefStepper*>(stepper)->step->reseed(
            UniqueEventId{static_cast<UniqueEventId::size_type>(event_id)});
    });
}

//! Stepper::kill_active (global/Stepper.cc:177-182)
int celerref_kill_active(void* stepper)
{
    return guarded([&] { static_cast<RefStepper*>(stepper)->step->kill_active(); });
}

//! Copy a per-slot state field into `out` (caller sizes it)
int celerref_state_get(void* stepper, char const* field, void* out)
{
    return guarded([&] {
        auto* s = static_cast<RefStepper*>(stepper);
        auto const& state = s->step->state_ref();
        auto const& params = s->problem->core->host_ref();
        size_type n = state.size();
        std::string f = field;
        auto* o32 = static_cast<uint32_t*>(out);
        auto* o64 = static_cast<double*>(out);
        auto* o8 = static_cast<uint8_t*>(out);
        for (size_type i = 0; i < n; ++i)
        {
            TrackSlotId ts{i};
            if (f == "status")
                o8[i] = static_cast<uint8_t>(state.sim.status[ts]);
            else if (f == "track_id")
                o32[i] = state.sim.track_ids[ts].unchecked_get();
            else if (f == "parent_id")
                o32[i] = state.sim.parent_ids[ts].unchecked_get();
            else if (f == "event_id")
                o32[i] = state.sim.event_ids[ts].unchecked_get();
            else if (f == "num_steps")
                o32[i] = state.sim.num_steps[ts];
            else if (f == "num_looping_steps")
                o32[i] = state.sim.num_looping_steps.empty()
                             ? 0
                             : state.sim.num_looping_steps[ts];
            else if (f == "time")
                o64[i] = state.sim.time[ts];
            else if (f == "step_length")
                o64[i] = state.sim.step_length[ts];
            else if (f == "post_step_action")
                o32[i] = state.sim.post_step_action[ts].unchecked_get();
            else if (f == "along_step_action")
                o32[i] = state.sim.along_step_action[ts].unchecked_get();
            else if (f == "track_slots")
                // thread -> slot permutation kept by SortTracksAction (TrackOrder::reindex_*)
                o32[i] = state.track_slots[ThreadId{i}];
            else if (f == "particle_id")
                o32[i] = state.particles.particle_id[ts].unchecked_get();
            else if (f == "energy")
                o64[i] = state.particles.particle_energy[ts];
            else if (f == "material_id")
                o32[i] = state.materials.state[ts].material_id.unchecked_get();
            else if (f == "interaction_mfp")
                o64[i] = state.physics.state[ts].interaction_mfp;
            else if (f == "macro_xs")
                o64[i] = state.physics.state[ts].macro_xs;
            else if (f == "energy_deposition")
                o64[i] = state.physics.state[ts].energy_deposition;
            else if (f == "dedx_range")
                o64[i] = state.physics.state[ts].dedx_range;
            else if (f == "rng")
            {
                auto const& r = state.rng.state[ts];
                for (int k = 0; k < 5; ++k)
                    o32[6 * i + k] = r.xorstate[k];
                o32[6 * i + 5] = r.weylstate;
            }
            else if (f == "pos" || f == "dir" || f == "volume_id"
                     || f == "surface_id" || f == "geo_level")
            {
                bool inactive = state.sim.status[ts] == TrackStatus::inactive;
                GeoTrackView geo(params.geometry, state.geometry, ts);
                if (f == "pos")
                    for (int k = 0; k < 3; ++k)
                        o64[3 * i + k] = inactive ? 0 : geo.pos()[k];
                else if (f == "dir")
                    for (int k = 0; k < 3; ++k)
                        o64[3 * i + k] = inactive ? 0 : geo.dir()[k];
                else if (f == "volume_id")
                    o32[i] = inactive ? 0xffffffffu
                                      : geo.volume_id().unchecked_get();
                else if (f == "surface_id")
                    o32[i] = inactive ? 0xffffffffu
                                      : geo.surface_id().unchecked_get();
                else
                    o32[i] = inactive ? 0xffffffffu
                                      : geo.level().unchecked_get();
            }
            else
            {
                CELER_VALIDATE(false, << "unknown state field '" << f << "'");
            }
        }
    });
}

//! Per-detector energy deposition accumulated by SimpleCalo [MeV]
int celerref_calo_get(void* problem, double* out)
{
    return guarded([&] {
        auto* p = static_cast<celerref::Problem*>(problem);
        CELER_VALIDATE(p->calo, << "no simple_calo in this problem");
        auto v = p->calo->calc_total_energy_deposition();
        std::copy(v.begin(), v.end(), out);
    });
}

//! Hits of the last step (HitRecorder): count, then one named field at a time.
//! Fields: detector track_id event_id parent_id track_step_count particle (u32),
//! step_length energy_deposition pre_time pre_energy post_time post_energy (f64),
//! pre_pos pre_dir post_pos post_dir (f64 x 3)
uint32_t celerref_hits_count(void* problem, uint32_t stream)
{
    auto* p = static_cast<celerref::Problem*>(problem);
    return p->hits ? p->hits->last(stream).size() : 0;
}

int celerref_hits_get(void* problem, uint32_t stream, char const* field, void* out)
{
    return guarded([&] {
        auto* p = static_cast<celerref::Problem*>(problem);
        CELER_VALIDATE(p->hits, << "no hit_volumes in this problem");
        DetectorStepOutput const& h = p->hits->last(stream);
        std::string f = field;
        auto* o32 = static_cast<uint32_t*>(out);
        auto* o64 = static_cast<double*>(out);
        auto ids = [&](auto const& v) {
            CELER_VALIDATE(v.size() == h.size(), << "field '" << f << "' was not collected");
            for (size_type i = 0; i < v.size(); ++i)
                o32[i] = v[i].unchecked_get();
        };
        auto point = [&](DetectorStepPointOutput const& pt, std::string const& name) {
            if (name == "time")
                std::copy(pt.time.begin(), pt.time.end(), o64);
            else if (name == "energy")
                for (size_type i = 0; i < pt.energy.size(); ++i)
                    o64[i] = pt.energy[i].value();
            else if (name == "pos" || name == "dir")
            {
                auto const& v = name == "pos" ? pt.pos : pt.dir;
                for (size_type i = 0; i < v.size(); ++i)
                    for (int k = 0; k < 3; ++k)
                        o64[3 * i + k] = v[i][k];
            }
            else
                CELER_VALIDATE(false, << "unknown hit field '" << f << "'");
        };
        if (f == "detector")
            ids(h.detector);
        else if (f == "track_id")
            ids(h.track_id);
        else if (f == "event_id")
            ids(h.event_id);
        else if (f == "parent_id")
            ids(h.parent_id);
        else if (f == "particle")
            ids(h.particle);
        else if (f == "track_step_count")
            std::copy(h.track_step_count.begin(), h.track_step_count.end(), o32);
        else if (f == "step_length")
            std::copy(h.step_length.begin(), h.step_length.end(), o64);
        else if (f == "energy_deposition")
            for (size_type i = 0; i < h.energy_deposition.size(); ++i)
                o64[i] = h.energy_deposition[i].value();
        else if (f.rfind("pre_", 0) == 0)
            point(h.points[StepPoint::pre], f.substr(4));
        else if (f.rfind("post_", 0) == 0)
            point(h.points[StepPoint::post], f.substr(5));
        else
            CELER_VALIDATE(false, << "unknown hit field '" << f << "'");
    });
}

int celerref_calo_clear(void* problem)
{
    return guarded([&] {
        auto* p = static_cast<celerref::Problem*>(problem);
        if (p->calo)
            p->calo->clear();
    });
}

//---------------------------------------------------------------------------//
/*!
 * Transport whole events the way celer-sim's Transporter does
 * (/root/reference/app/celer-sim/Transporter.cc:84-179): one Stepper per
 * OpenMP thread, events dealt to threads, `num_steps += active` per iteration.
 *
 * primaries are grouped by event: event e owns [offsets[e], offsets[e+1]).
 * result = {num_steps, num_step_iterations, num_tracks(generated+secondaries
 * is not tracked by the reference; primaries only), max_queued}; returns wall
 * seconds of the transport loop (setup and one warm-up step excluded).
 */
// Diagnostic tallies summed over streams: counts[particle][bin]; *num_bins = bins per
// particle. `out` may be null to query the size.
int celerref_diagnostic_get(void* problem, int steps, uint32_t* out, uint32_t* num_bins)
{
    return guarded([&] {
        auto* p = static_cast<celerref::Problem*>(problem);
        std::vector<std::vector<size_type>> counts;
        if (steps)
        {
            CELER_VALIDATE(p->step_diag, << "no step diagnostic");
            counts = p->step_diag->calc_steps();
        }
        else
        {
            CELER_VALIDATE(p->action_diag, << "no action diagnostic");
            counts = p->action_diag->calc_actions();
        }
        *num_bins = counts.empty() ? 0 : counts.front().size();
        if (out)
        {
            for (auto const& row : counts)
                out = std::copy(row.begin(), row.end(), out);
        }
    });
}

int celerref_num_particles(void* problem)
{
    auto* p = static_cast<celerref::Problem*>(problem);
    return p->core->particle()->size();
}

// Action labels by id, newline separated
int celerref_action_labels(void* problem, char* out, uint32_t capacity)
{
    return guarded([&] {
        auto* p = static_cast<celerref::Problem*>(problem);
        std::string text;
        auto const& reg = *p->core->action_reg();
        for (auto i : range(ActionId{reg.num_actions()}))
        {
            text += reg.id_to_label(i);
            text += '\n';
        }
        CELER_VALIDATE(text.size() < capacity, << "buffer too small");
        std::copy(text.begin(), text.end(), out);
        out[text.size()] = 0;
    });
}

// Primaries from celer-sim "primary_options" JSON with the reference's PrimaryGenerator
int celerref_generate_primaries(void* problem,
                                char const* options_json,
                                CPrimary* out,
                                uint64_t capacity,
                                uint64_t* count)
{
    return guarded([&] {
        auto* p = static_cast<celerref::Problem*>(problem);
        PrimaryGeneratorOptions opts;
        nlohmann::json::parse(options_json).get_to(opts);
        auto generate = PrimaryGenerator::from_options(p->core->particle(), opts);
        uint64_t n = 0;
        for (auto event = generate(); !event.empty(); event = generate())
        {
            for (Primary const& pr : event)
            {
                if (out && n < capacity)
                {
                    CPrimary& c = out[n];
                    c.particle_id = pr.particle_id.unchecked_get();
                    c.event_id = pr.event_id.unchecked_get();
                    c.energy = pr.energy.value();
                    for (int k = 0; k < 3; ++k)
                    {
                        c.pos[k] = pr.position[k];
                        c.dir[k] = pr.direction[k];
                    }
                    c.time = pr.time;
                }
                ++n;
            }
        }
        *count = n;
    });
}

double celerref_run_events(void* problem,
                           CPrimary const* primaries,
                           uint32_t const* offsets,
                           uint32_t num_events,
                           uint32_t num_track_slots,
                           int num_threads,
                           uint64_t* result)
{
    double elapsed = -1;
    guarded([&] {
        auto* p = static_cast<celerref::Problem*>(problem);
        if (num_threads <= 0)
            num_threads = omp_get_max_threads();
        num_threads = std::min<int>(num_threads, p->core->max_streams());
        std::vector<std::unique_ptr<Stepper<MemSpace::host>>> steppers(
            num_threads);
        for (int t = 0; t < num_threads; ++t)
        {
            StepperInput inp;
            inp.params = p->core;
            inp.stream_id = StreamId{static_cast<size_type>(t)};
            inp.num_track_slots = num_track_slots;
            steppers[t] = std::make_unique<Stepper<MemSpace::host>>(inp);
            steppers[t]->warm_up();
        }
        uint64_t num_steps = 0, num_iters = 0, max_queued = 0, num_prim = 0;
        std::string err;
        auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(num_threads) schedule(dynamic, 1) \
    reduction(+ : num_steps, num_iters, num_prim) reduction(max : max_queued)
        for (uint32_t e = 0; e < num_events; ++e)
        {
            try
            {
                auto& step = *steppers[omp_get_thread_num()];
                auto prim = to_primaries(primaries + offsets[e],
                                         offsets[e + 1] - offsets[e]);
                step.reseed(UniqueEventId{prim.front().event_id.get()});
                num_prim += prim.size();
                auto counts = step(make_span(prim));
                while (true)
                {
                    num_steps += counts.active;
                    ++num_iters;
                    max_queued = std::max<uint64_t>(max_queued, counts.queued);
                    if (!counts)
                        break;
                    counts = step();
                }
            }
            catch (std::exception const& ex)
            {
#pragma omp critical
                err = ex.what();
            }
        }
        auto t1 = std::chrono::steady_clock::now();
        CELER_VALIDATE(err.empty(), << err);
        elapsed = std::chrono::duration<double>(t1 - t0).count();
        result[0] = num_steps;
        result[1] = num_iters;
        result[2] = num_prim;
        result[3] = max_queued;
    });
    return elapsed;
}

//! Transport all primaries merged onto ONE device state (celer-sim `merge_events`,
//! app/celer-sim/Transporter.cc:75-135) with the reference's own CUDA kernels; returns the
//! wall time of the transport loop after warm_up (the reference's `time.total` definition).
double celerref_run_merged_device(void* problem,
                                  CPrimary const* primaries,
                                  uint32_t num_primaries,
                                  uint32_t num_track_slots,
                                  uint64_t* result)
{
    double elapsed = -1;
#if CELERITAS_USE_CUDA
    guarded([&] {
        auto* p = static_cast<celerref::Problem*>(problem);
        StepperInput inp;
        inp.params = p->core;
        inp.stream_id = StreamId{0};
        inp.num_track_slots = num_track_slots;
        Stepper<MemSpace::device> step(inp);
        step.warm_up();
        auto prim = to_primaries(primaries, num_primaries);
        uint64_t num_steps = 0, num_iters = 0, max_queued = 0;
        cudaDeviceSynchronize();
        auto t0 = std::chrono::steady_clock::now();
        auto counts = step(make_span(prim));
        while (true)
        {
            num_steps += counts.active;
            ++num_iters;
            max_queued = std::max<uint64_t>(max_queued, counts.queued);
            if (!counts)
                break;
            counts = step();
        }
        cudaDeviceSynchronize();
        auto t1 = std::chrono::steady_clock::now();
        elapsed = std::chrono::duration<double>(t1 - t0).count();
        result[0] = num_steps;
        result[1] = num_iters;
        result[2] = num_primaries;
        result[3] = max_queued;
    });
#else
    (void)problem; (void)primaries; (void)num_primaries; (void)num_track_slots; (void)result;
    g_last_error = "celerref_run_merged_device: not a CUDA build (make -C oracle ref_cuda)";
#endif
    return elapsed;
}
}  // extern "C"

//---------------------------------------------------------------------------//
// Element data readers (build-time helper for tools/make_physics.py): parse the
// reference's bundled G4EMLOW excerpts (test/celeritas/data/br29, pe-*-19.dat)
// with the reference's own readers and return them as JSON.
//---------------------------------------------------------------------------//
#include "celeritas/io/LivermorePEReader.hh"
#include "celeritas/io/SeltzerBergerReader.hh"

extern "C" int celerref_element_data_json(char const* dir, int z_sb, int z_pe, char* out, size_t size)
{
    return guarded([&] {
        json j;
        {
            SeltzerBergerReader read_sb(dir);
            auto t = read_sb(AtomicNumber{z_sb});
            j["sb"] = {{"x", t.x}, {"y", t.y}, {"value", t.value}};
        }
        {
            LivermorePEReader read_pe(dir);
            auto pe = read_pe(AtomicNumber{z_pe});
            auto vec = [](ImportPhysicsVector const& v) {
                return json{{"vector_type", static_cast<int>(v.vector_type)}, {"x", v.x}, {"y", v.y}};
            };
            json shells = json::array();
            for (auto const& s : pe.shells)
            {
                shells.push_back({{"binding_energy", s.binding_energy},
                                  {"param_lo", s.param_lo},
                                  {"param_hi", s.param_hi},
                                  {"xs", s.xs},
                                  {"energy", s.energy}});
            }
            j["livermore_pe"] = {{"xs_lo", vec(pe.xs_lo)},
                                 {"xs_hi", vec(pe.xs_hi)},
                                 {"thresh_lo", pe.thresh_lo},
                                 {"thresh_hi", pe.thresh_hi},
                                 {"shells", shells}};
        }
        std::string s = j.dump();
        CELER_VALIDATE(s.size() + 1 <= size, << "output buffer too small");
        std::memcpy(out, s.c_str(), s.size() + 1);
    });
}

//---------------------------------------------------------------------------//
// Ray tracing with the reference's OrangeTrackView on the host (same protocol as
// celeritas_b200/csrc/geo_trace.cu): per ray, the sequence of (volume id, surface id
// crossed, segment length) until the ray leaves the geometry, and the safety at
// the origin.
//---------------------------------------------------------------------------//
#include "corecel/data/CollectionStateStore.hh"
#include "orange/OrangeParams.hh"
#include "orange/OrangeTrackView.hh"

extern "C" int celerref_geo_trace(void* problem,
                                  double const* pos,
                                  double const* dir,
                                  uint32_t num_rays,
                                  uint32_t max_segments,
                                  uint32_t* volume,
                                  uint32_t* surface,
                                  double* distance,
                                  uint32_t* count,
                                  double* safety)
{
    return guarded([&] {
        auto* p = static_cast<celerref::Problem*>(problem);
        std::shared_ptr<OrangeParams const> geo
            = p->geo ? p->geo : p->core->geometry();
        auto const& params = geo->host_ref();
        CollectionStateStore<OrangeStateData, MemSpace::host> store(params, num_rays);
        for (uint32_t i = 0; i < num_rays; ++i)
        {
            OrangeTrackView trk(params, store.ref(), TrackSlotId{i});
            trk = GeoTrackInitializer{{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]},
                                      {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]}};
            if (trk.failed())
            {
                count[i] = 0xffffffffu;
                safety[i] = -1;
                continue;
            }
            safety[i] = trk.is_outside() ? -1 : trk.find_safety();
            uint32_t n = 0;
            while (!trk.is_outside() && n < max_segments)
            {
                uint32_t vol = trk.volume_id().unchecked_get();
                auto prop = trk.find_next_step();
                if (!prop.boundary)
                {
                    volume[i * max_segments + n] = vol;
                    surface[i * max_segments + n] = 0xffffffffu;
                    distance[i * max_segments + n] = prop.distance;
                    ++n;
                    break;
                }
                trk.move_to_boundary();
                volume[i * max_segments + n] = vol;
                surface[i * max_segments + n] = trk.surface_id().unchecked_get();
                distance[i * max_segments + n] = prop.distance;
                ++n;
                trk.cross_boundary();
                if (trk.failed())
                {
                    n |= 0x80000000u;
                    break;
                }
            }
            count[i] = n;
        }
    });
}
