//---------------------------------------------------------------------------//
// C-ABI (include/celeritas_b200.h) over the host-side objects.
//---------------------------------------------------------------------------//
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/celeritas_b200.h"
#include "CoreParams.hh"
#include "CoreState.hh"
#include "Handles.hh"
#include "OrangeBuilder.hh"
#include "RootImport.hh"
#include "Stepper.hh"
#include "Transporter.hh"

using namespace celeritas_b200;


namespace
{
thread_local std::string g_error;

template<class F>
int guarded(F&& f)
{
    try
    {
        f();
        return B200_OK;
    }
    catch (CudaError const& e)
    {
        g_error = e.what();
        return e.code;
    }
    catch (std::exception const& e)
    {
        g_error = e.what();
        return B200_ERR_RUNTIME;
    }
}
}  // namespace

void celeritas_b200::set_last_error(std::string const& message)
{
    g_error = message;
}

extern "C" {
void b200_string_free(char* s)
{
    std::free(s);
}

char const* b200_last_error(void)
{
    return g_error.c_str();
}

int b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
        return 0;
    return n;
}

//---------------------------------------------------------------------------//
int b200_params_create_from_image(char const* image_path, B200Params** out)
{
    if (!image_path || !out)
        return B200_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (b200_device_count() == 0)
    {
        g_error = "no CUDA device available (this library has no CPU path)";
        return B200_ERR_NO_DEVICE;
    }
    return guarded([&] {
        auto p = std::make_unique<B200Params>();
        p->params = CoreParams::from_image(image_path);
        *out = p.release();
    });
}

int b200_params_create_from_memory(void const* image, size_t size, B200Params** out)
{
    if (!image || !out)
        return B200_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (b200_device_count() == 0)
    {
        g_error = "no CUDA device available (this library has no CPU path)";
        return B200_ERR_NO_DEVICE;
    }
    return guarded([&] {
        auto p = std::make_unique<B200Params>();
        p->params = CoreParams::from_image(b200::Image::parse(image, size));
        *out = p.release();
    });
}

int b200_orange_build_image(char const* org_json_path, void** image, size_t* size)
{
    if (!org_json_path || !image || !size)
        return B200_ERR_INVALID_ARGUMENT;
    *image = nullptr;
    *size = 0;
    return guarded([&] {
        std::vector<unsigned char> const bytes = build_orange_image(org_json_path).serialize();
        void* out = std::malloc(bytes.size());
        if (!out)
            throw std::runtime_error("out of memory");
        std::memcpy(out, bytes.data(), bytes.size());
        *image = out;
        *size = bytes.size();
    });
}

int b200_import_root(char const* root_path, char** import_data_json)
{
    if (!root_path || !import_data_json)
        return B200_ERR_INVALID_ARGUMENT;
    *import_data_json = nullptr;
    return guarded([&] {
        std::string const text = b200::import_root_to_json(root_path);
        char* out = static_cast<char*>(std::malloc(text.size() + 1));
        if (!out)
            throw std::runtime_error("out of memory");
        std::memcpy(out, text.c_str(), text.size() + 1);
        *import_data_json = out;
    });
}

int b200_params_create_from_org_json(char const* org_json_path, B200Params** out)
{
    if (!org_json_path || !out)
        return B200_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (b200_device_count() == 0)
    {
        g_error = "no CUDA device available (this library has no CPU path)";
        return B200_ERR_NO_DEVICE;
    }
    return guarded([&] {
        auto p = std::make_unique<B200Params>();
        p->params = CoreParams::from_image(build_orange_image(org_json_path));
        *out = p.release();
    });
}

void b200_params_destroy(B200Params* params)
{
    delete params;
}

B200ParamsView const* b200_params_view(B200Params const* params)
{
    return reinterpret_cast<B200ParamsView const*>(&params->params->view());
}

uint32_t b200_params_num_actions(B200Params const* params)
{
    return params->params->actions().size();
}

char const* b200_params_action_label(B200Params const* params, uint32_t action_id)
{
    auto const& a = params->params->actions();
    return action_id < a.size() ? a[action_id].label.c_str() : "";
}

uint32_t b200_params_num_volumes(B200Params const* params)
{
    return params->params->volume_labels().size();
}

char const* b200_params_volume_label(B200Params const* params, uint32_t volume_id)
{
    auto const& v = params->params->volume_labels();
    return volume_id < v.size() ? v[volume_id].c_str() : "";
}

uint32_t b200_params_num_detectors(B200Params const* params)
{
    return params->params->num_detectors();
}

uint32_t b200_params_num_particles(B200Params const* params)
{
    return params->params->particle_names().size();
}

uint32_t b200_params_num_models(B200Params const* params)
{
    return params->params->view().phys.num_models;
}

uint32_t b200_params_model_action_begin(B200Params const* params)
{
    return params->params->view().phys.model_to_action;
}

uint32_t b200_params_max_depth(B200Params const* params)
{
    return params->params->view().geo.max_depth;
}

uint32_t b200_params_find_particle(B200Params const* params, int pdg)
{
    return params->params->find_particle(pdg);
}

//---------------------------------------------------------------------------//
int b200_state_create(B200Params const* params,
                      uint32_t stream_id,
                      uint32_t num_track_slots,
                      B200State** out)
{
    if (!params || !out)
        return B200_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    return guarded([&] {
        auto s = std::make_unique<B200State>();
        s->owned = std::make_unique<CoreState>(params->params, stream_id, num_track_slots);
        s->state = s->owned.get();
        *out = s.release();
    });
}

void b200_state_destroy(B200State* state)
{
    delete state;
}

B200StateView const* b200_state_view(B200State const* state)
{
    return reinterpret_cast<B200StateView const*>(&state->state->view());
}

int b200_state_get(B200State* state, char const* field, void* out)
{
    return guarded([&] { state->state->get_field(field, out); });
}

int b200_state_calo_get(B200State* state, double* out)
{
    return guarded([&] { state->state->calo_get(out); });
}

int b200_state_calo_clear(B200State* state)
{
    return guarded([&] { state->state->calo_clear(); });
}

//---------------------------------------------------------------------------//
int b200_stepper_create_opts(B200Params const* params,
                             B200StepperOptions const* options,
                             B200Stepper** out)
{
    if (!params || !options || !out)
        return B200_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    return guarded([&] {
        auto s = std::make_unique<B200Stepper>();
        StepperInput inp;
        inp.params = params->params;
        inp.stream_id = options->stream_id;
        inp.num_track_slots = options->num_track_slots;
        inp.action_times = options->action_times != 0;
        inp.actions.action_diagnostic = options->action_diagnostic != 0;
        inp.actions.step_diagnostic_bins = options->step_diagnostic_bins;
        inp.actions.fuse_threshold = options->fuse_threshold;
        inp.actions.tail_threshold = options->tail_threshold;
        s->stepper = std::make_shared<Stepper>(std::move(inp));
        s->state_handle.state = &s->stepper->state();
        s->launches_at_create = b200_launch_count();
        *out = s.release();
    });
}

int b200_stepper_create(B200Params const* params,
                        uint32_t stream_id,
                        uint32_t num_track_slots,
                        B200Stepper** out)
{
    B200StepperOptions options{};
    options.stream_id = stream_id;
    options.num_track_slots = num_track_slots;
    return b200_stepper_create_opts(params, &options, out);
}

uint32_t b200_stepper_num_actions(B200Stepper const* stepper)
{
    return stepper->stepper->actions().labels().size();
}

char const* b200_stepper_action_label(B200Stepper const* stepper, uint32_t action_id)
{
    auto const& labels = stepper->stepper->actions().labels();
    return action_id < labels.size() ? labels[action_id].c_str() : "";
}

int b200_stepper_action_diagnostic_get(B200Stepper* stepper, uint32_t* counts)
{
    if (!stepper->stepper->actions().action_diagnostic())
        return B200_ERR_INVALID_ARGUMENT;
    return guarded([&] { stepper->stepper->state().diagnostic_get(false, counts); });
}

int b200_stepper_step_diagnostic_get(B200Stepper* stepper, uint32_t* counts)
{
    if (!stepper->stepper->actions().step_diagnostic_bins())
        return B200_ERR_INVALID_ARGUMENT;
    return guarded([&] { stepper->stepper->state().diagnostic_get(true, counts); });
}

uint32_t b200_stepper_step_diagnostic_bins(B200Stepper const* stepper)
{
    return stepper->stepper->actions().step_diagnostic_bins();
}

int b200_stepper_diagnostics_clear(B200Stepper* stepper)
{
    return guarded([&] { stepper->stepper->state().diagnostics_clear(); });
}

void b200_stepper_destroy(B200Stepper* stepper)
{
    delete stepper;
}

B200State* b200_stepper_state(B200Stepper* stepper)
{
    return &stepper->state_handle;
}

int b200_stepper_step(B200Stepper* stepper,
                      B200Primary const* primaries,
                      uint32_t num_primaries,
                      B200StepperResult* result)
{
    return guarded([&] {
        StepperResult r = num_primaries ? (*stepper->stepper)(primaries, num_primaries)
                                        : (*stepper->stepper)();
        if (result)
        {
            result->generated = r.generated;
            result->queued = r.queued;
            result->active = r.active;
            result->alive = r.alive;
        }
    });
}

int b200_stepper_insert(B200Stepper* stepper, B200Primary const* primaries, uint32_t num_primaries)
{
    if (!stepper || (num_primaries && !primaries))
        return B200_ERR_INVALID_ARGUMENT;
    return guarded([&] { stepper->stepper->insert(primaries, num_primaries); });
}

int b200_stepper_begin_iteration(B200Stepper* stepper)
{
    if (!stepper)
        return B200_ERR_INVALID_ARGUMENT;
    return guarded([&] { stepper->stepper->begin_iteration(); });
}

int b200_stepper_end_iteration(B200Stepper* stepper, B200StepperResult* result)
{
    if (!stepper)
        return B200_ERR_INVALID_ARGUMENT;
    return guarded([&] {
        StepperResult r = stepper->stepper->end_iteration();
        if (result)
        {
            result->generated = r.generated;
            result->queued = r.queued;
            result->active = r.active;
            result->alive = r.alive;
        }
    });
}

cudaStream_t b200_stepper_stream(B200Stepper* stepper)
{
    return stepper ? stepper->stepper->state().stream() : nullptr;
}

int b200_stepper_advance(B200Stepper* stepper,
                         uint32_t max_iterations,
                         B200StepperResult* results,
                         uint32_t* num_done)
{
    if (!stepper || !results || !num_done)
        return B200_ERR_INVALID_ARGUMENT;
    *num_done = 0;
    return guarded([&] {
        std::vector<StepperResult> batch;
        uint32_t const n = stepper->stepper->advance(max_iterations, &batch);
        for (uint32_t i = 0; i < n; ++i)
        {
            results[i].generated = batch[i].generated;
            results[i].queued = batch[i].queued;
            results[i].active = batch[i].active;
            results[i].alive = batch[i].alive;
        }
        *num_done = n;
    });
}

int b200_stepper_hits_count(B200Stepper* stepper, uint32_t* count)
{
    if (!stepper || !count)
        return B200_ERR_INVALID_ARGUMENT;
    return guarded([&] { *count = stepper->stepper->state().hits_count(); });
}

int b200_stepper_hits_get(B200Stepper* stepper, char const* field, void* out)
{
    if (!stepper || !field || !out)
        return B200_ERR_INVALID_ARGUMENT;
    return guarded([&] { stepper->stepper->state().hits_get(field, out); });
}

uint64_t b200_stepper_tail_iterations(B200Stepper const* stepper)
{
    return stepper->stepper->tail_iterations();
}

int b200_stepper_warm_up(B200Stepper* stepper)
{
    return guarded([&] { stepper->stepper->warm_up(); });
}

int b200_stepper_reseed(B200Stepper* stepper, uint64_t event_id)
{
    return guarded([&] { stepper->stepper->reseed(event_id); });
}

int b200_stepper_kill_active(B200Stepper* stepper)
{
    return guarded([&] { stepper->stepper->kill_active(); });
}

uint32_t b200_stepper_num_step_actions(B200Stepper const* stepper)
{
    return stepper->stepper->actions().actions().size();
}

char const* b200_stepper_step_action_label(B200Stepper const* stepper, uint32_t i)
{
    auto const& a = stepper->stepper->actions().actions();
    return i < a.size() ? a[i]->label().c_str() : "";
}

int b200_set_device(int device)
{
    return guarded([&] { B2_CUDA_CALL(cudaSetDevice(device)); });
}

int b200_stepper_set_action_times(B200Stepper* stepper, int enable)
{
    return guarded([&] { stepper->stepper->action_sequence().action_times(enable != 0); });
}

double b200_stepper_action_time(B200Stepper const* stepper, uint32_t i)
{
    auto const& t = stepper->stepper->actions().accum_time();
    return i < t.size() ? t[i] : 0.0;
}

uint64_t b200_stepper_launch_count(B200Stepper const* stepper)
{
    return b200_launch_count() - stepper->launches_at_create;
}

//---------------------------------------------------------------------------//
// Transport whole events (reference: app/celer-sim/Transporter.cc:84-179)
int b200_run_events(B200Stepper* handle,
                    B200Primary const* primaries,
                    uint32_t const* offsets,
                    uint32_t num_events,
                    int merge_events,
                    uint64_t max_steps,
                    B200RunResult* result)
{
    if (!handle || !primaries || !offsets)
        return B200_ERR_INVALID_ARGUMENT;
    return guarded([&] {
        TransporterInput tinp;
        tinp.max_steps = max_steps;
        Transporter transport(handle->stepper, tinp);
        Stepper& step = *handle->stepper;
        cudaStream_t stream = step.state().stream();
        cudaEvent_t ev0, ev1;
        B2_CUDA_CALL(cudaEventCreate(&ev0));
        B2_CUDA_CALL(cudaEventCreate(&ev1));
        B200RunResult r{};
        B2_CUDA_CALL(cudaEventRecord(ev0, stream));
        auto run = [&](B200Primary const* p, uint32_t n) {
            step.reseed(p[0].event_id);
            TransporterResult t = transport(p, n);
            r.num_primaries += n;
            r.num_steps += t.num_steps;
            r.num_step_iterations += t.num_step_iterations;
            r.num_tracks += t.num_tracks;
            r.num_aborted += t.num_aborted;
            r.max_queued = std::max<uint64_t>(r.max_queued, t.max_queued);
        };
        if (merge_events)
        {
            run(primaries, offsets[num_events]);
        }
        else
        {
            for (uint32_t e = 0; e < num_events; ++e)
            {
                uint32_t n = offsets[e + 1] - offsets[e];
                if (n != 0)
                    run(primaries + offsets[e], n);
            }
        }
        B2_CUDA_CALL(cudaEventRecord(ev1, stream));
        B2_CUDA_CALL(cudaEventSynchronize(ev1));
        float ms = 0;
        B2_CUDA_CALL(cudaEventElapsedTime(&ms, ev0, ev1));
        r.seconds = ms * 1e-3;
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
        if (result)
            *result = r;
    });
}

//---------------------------------------------------------------------------//
// Several streams at once: one host thread per Stepper, as celer-sim runs one
// OpenMP thread per stream (app/celer-sim/celer-sim.cc:120-137). On the device the
// streams overlap: the latency-bound small iterations of one shower tail fill in
// under the throughput-bound iterations of another.
int b200_run_events_streams(B200Stepper* const* handles,
                            uint32_t num_streams,
                            B200Primary const* primaries,
                            uint32_t const* offsets,
                            uint32_t num_events,
                            int merge_events,
                            uint64_t max_steps,
                            B200RunResult* results,
                            double* seconds)
{
    if (!handles || num_streams == 0 || !primaries || !offsets || !results)
        return B200_ERR_INVALID_ARGUMENT;
    for (uint32_t k = 0; k < num_streams; ++k)
        if (!handles[k])
            return B200_ERR_INVALID_ARGUMENT;
    return guarded([&] {
        int device = 0;
        B2_CUDA_CALL(cudaGetDevice(&device));
        // Event e belongs to stream e % num_streams (static, so runs are reproducible)
        std::vector<std::vector<B200Primary>> merged(num_streams);
        if (merge_events)
        {
            for (uint32_t e = 0; e < num_events; ++e)
            {
                auto& dst = merged[e % num_streams];
                dst.insert(dst.end(), primaries + offsets[e], primaries + offsets[e + 1]);
            }
        }
        cudaStream_t stream0 = handles[0]->stepper->state().stream();
        cudaEvent_t ev0, ev1;
        B2_CUDA_CALL(cudaEventCreate(&ev0));
        B2_CUDA_CALL(cudaEventCreate(&ev1));
        std::vector<std::string> errors(num_streams);
        std::vector<int> codes(num_streams, 0);
        B2_CUDA_CALL(cudaEventRecord(ev0, stream0));
        auto work = [&](uint32_t k) {
            try
            {
                B2_CUDA_CALL(cudaSetDevice(device));
                TransporterInput tinp;
                tinp.max_steps = max_steps;
                Transporter transport(handles[k]->stepper, tinp);
                Stepper& step = *handles[k]->stepper;
                B200RunResult r{};
                auto run = [&](B200Primary const* p, uint32_t n) {
                    step.reseed(p[0].event_id);
                    TransporterResult t = transport(p, n);
                    r.num_primaries += n;
                    r.num_steps += t.num_steps;
                    r.num_step_iterations += t.num_step_iterations;
                    r.num_tracks += t.num_tracks;
                    r.num_aborted += t.num_aborted;
                    r.max_queued = std::max<uint64_t>(r.max_queued, t.max_queued);
                };
                if (merge_events)
                {
                    if (!merged[k].empty())
                        run(merged[k].data(), merged[k].size());
                }
                else
                {
                    for (uint32_t e = k; e < num_events; e += num_streams)
                    {
                        uint32_t n = offsets[e + 1] - offsets[e];
                        if (n != 0)
                            run(primaries + offsets[e], n);
                    }
                }
                results[k] = r;
            }
            catch (CudaError const& e)
            {
                errors[k] = e.what();
                codes[k] = e.code;
            }
            catch (std::exception const& e)
            {
                errors[k] = e.what();
                codes[k] = B200_ERR_RUNTIME;
            }
        };
        std::vector<std::thread> threads;
        for (uint32_t k = 1; k < num_streams; ++k)
            threads.emplace_back(work, k);
        work(0);
        for (auto& t : threads)
            t.join();
        // Every stream is idle here (each transport ends with a synchronised iteration)
        B2_CUDA_CALL(cudaEventRecord(ev1, stream0));
        B2_CUDA_CALL(cudaEventSynchronize(ev1));
        float ms = 0;
        B2_CUDA_CALL(cudaEventElapsedTime(&ms, ev0, ev1));
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
        for (uint32_t k = 0; k < num_streams; ++k)
        {
            if (codes[k] != 0)
                throw std::runtime_error("stream " + std::to_string(k) + ": " + errors[k]);
            results[k].seconds = ms * 1e-3;
        }
        if (seconds)
            *seconds = ms * 1e-3;
    });
}
}  // extern "C"
