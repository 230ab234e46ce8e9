//---------------------------------------------------------------------------//
// TEST INFRASTRUCTURE ONLY. C API over the compiled drop-in
// (celeritas_b200/adapter/B200Actions.{hh,cc}): the B200 step actions registered behind the
// reference's CoreStepActionInterface, executed by the reference's own ActionSequence on the
// reference's own CoreState<device> (the reference's CUDA build, oracle/_ref/
// libcelerref_cuda.so), with the problem handed across in memory
// (b200_params_create_from_memory). Built by `make -C oracle dropin` into
// oracle/_ref/libcelerref_dropin.so; driven by tests/dropin_run.py.
//---------------------------------------------------------------------------//
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "corecel/sys/ActionRegistry.hh"
#include "celeritas/global/ActionSequence.hh"
#include "celeritas/global/CoreState.hh"
#include "celeritas/global/Stepper.hh"
#include "celeritas/phys/Primary.hh"

#include "../ref_harness/Problem.hh"
#include "B200Actions.hh"

using namespace celeritas;
namespace adapter = celeritas_b200_adapter;

namespace
{
thread_local std::string g_error;

//! Same POD layout as B200Primary
struct CPrimary
{
    uint32_t particle_id;
    uint32_t event_id;
    double energy;
    double pos[3];
    double dir[3];
    double time;
};

struct DropIn
{
    celerref::Problem* problem;
    std::shared_ptr<adapter::B200Problem const> uploaded;
    std::unique_ptr<adapter::B200StepperAdapter> step;
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
        g_error = e.what();
        return 1;
    }
}
}  // namespace

extern "C" {
char const* celerref_dropin_last_error()
{
    return g_error.c_str();
}

//! Reference CoreParams (already built, on the device) -> image in memory -> B200 stepper
//! behind the reference's StepperInterface
void* celerref_dropin_create(void* problem, uint32_t num_track_slots, uint32_t stream_id)
{
    void* result = nullptr;
    guarded([&] {
        auto* p = static_cast<celerref::Problem*>(problem);
        auto d = std::make_unique<DropIn>();
        d->problem = p;
        std::vector<unsigned char> const image = celerref::export_image_bytes(*p);
        d->uploaded = std::make_shared<adapter::B200Problem>(image.data(), image.size());
        StepperInput inp;
        inp.params = p->core;
        inp.stream_id = StreamId{stream_id};
        inp.num_track_slots = num_track_slots;
        d->step = std::make_unique<adapter::B200StepperAdapter>(std::move(inp), d->uploaded);
        result = d.release();
    });
    return result;
}

void celerref_dropin_destroy(void* d)
{
    delete static_cast<DropIn*>(d);
}

//! One step iteration through StepperInterface. counts = {generated, queued, active, alive}
int celerref_dropin_step(void* dropin, CPrimary const* primaries, uint32_t n, uint32_t* counts)
{
    return guarded([&] {
        StepperInterface& step = *static_cast<DropIn*>(dropin)->step;
        StepperResult r;
        if (n > 0)
        {
            std::vector<Primary> prim(n);
            for (uint32_t i = 0; i < n; ++i)
            {
                prim[i].particle_id = ParticleId{primaries[i].particle_id};
                prim[i].energy = units::MevEnergy{primaries[i].energy};
                prim[i].position = {primaries[i].pos[0], primaries[i].pos[1], primaries[i].pos[2]};
                prim[i].direction = {primaries[i].dir[0], primaries[i].dir[1], primaries[i].dir[2]};
                prim[i].time = primaries[i].time;
                prim[i].event_id = EventId{primaries[i].event_id};
            }
            r = step(make_span(prim));
        }
        else
        {
            r = step();
        }
        counts[0] = r.generated;
        counts[1] = r.queued;
        counts[2] = r.active;
        counts[3] = r.alive;
    });
}

int celerref_dropin_warm_up(void* dropin)
{
    return guarded([&] { static_cast<DropIn*>(dropin)->step->warm_up(); });
}

int celerref_dropin_reseed(void* dropin, uint64_t event_id)
{
    return guarded([&] {
        static_cast<DropIn*>(dropin)->step->reseed(
            UniqueEventId{static_cast<UniqueEventId::size_type>(event_id)});
    });
}

//! Labels of the step actions in the order the reference's ActionSequence runs them, newline
//! separated; returns the number of bytes needed
int celerref_dropin_sequence_labels(void* dropin, char* out, uint32_t capacity)
{
    std::string text;
    for (auto const& sp : static_cast<DropIn*>(dropin)->step->actions().actions().step())
        text += std::string(sp->label()) + "\n";
    if (out && capacity > 0)
    {
        std::strncpy(out, text.c_str(), capacity - 1);
        out[capacity - 1] = '\0';
    }
    return static_cast<int>(text.size() + 1);
}

//! Kernel launches issued by libceleritas_b200.so for this stepper
uint64_t celerref_dropin_launch_count(void* dropin)
{
    return static_cast<DropIn*>(dropin)->step->launch_count();
}

//! Copy a per-slot field of the B200 track state to the host (b200_state_get)
int celerref_dropin_state_get(void* dropin, char const* field, void* out)
{
    return guarded([&] {
        auto& st = static_cast<DropIn*>(dropin)->step->b200_state();
        int rc = b200_state_get(b200_stepper_state(st.handle()), field, out);
        CELER_VALIDATE(rc == 0, << b200_last_error());
    });
}

//! The reference CoreState's own counters after the last iteration:
//! {num_generated, num_initializers, num_vacancies, num_active, num_alive}
int celerref_dropin_counters(void* dropin, uint32_t* out)
{
    return guarded([&] {
        auto const& c = static_cast<DropIn*>(dropin)->step->state().counters();
        out[0] = c.num_generated;
        out[1] = c.num_initializers;
        out[2] = c.num_vacancies;
        out[3] = c.num_active;
        out[4] = c.num_alive;
    });
}
}  // extern "C"
