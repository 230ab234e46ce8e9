/*---------------------------------------------------------------------------*
 * celeritas_b200: C-ABI of the B200-native per-step track loop.
 *
 * Plain C, pointers and sizes only. Everything the reference's step loop does
 * on the device is reachable from here; there is no CPU fallback behind any
 * entry point (they fail with a CUDA error code when no device is present).
 *
 * The entry points mirror the reference's plugin surface for the hot path:
 *
 *   reference (file:line under /root/reference/src)               | here
 *   ---------------------------------------------------------------+---------------------------
 *   CoreParams (celeritas/global/CoreParams.hh:42-146)             | b200_params_*
 *   CoreState<device> (celeritas/global/CoreState.hh:73-182)       | b200_state_*
 *   StepActionInterface::step(params, state)                       | b200_step_<action>
 *     (corecel/sys/ActionInterface.hh:175-186), one per action:    |
 *     ExtendFromPrimariesAction (track/ExtendFromPrimariesAction.cc:183-231)   | b200_step_extend_from_primaries
 *     InitializeTracksAction (track/InitializeTracksAction.cc:55-97)           | b200_step_initialize_tracks
 *     PreStepAction (phys/detail/PreStepExecutor.hh:45-115)                    | b200_step_pre_step
 *     AlongStep{Neutral,GeneralLinear,UniformMsc}Action                        | b200_step_along_step
 *       (global/alongstep/AlongStep.hh:50-58)                                  |
 *     DiscreteSelectAction (phys/detail/DiscreteSelectExecutor.hh:37-63)       | b200_step_discrete_select
 *     every EM model's step() (em/model/XModel.cu via InteractionApplier)      | b200_step_interact
 *     BoundaryAction (geo/detail/BoundaryExecutor.hh:41-84)                    | b200_step_boundary
 *     TrackingCutAction (phys/detail/TrackingCutExecutor.hh:48-83)             | b200_step_tracking_cut
 *     StepGather/SimpleCalo (user/detail/SimpleCaloExecutor.hh:48-67)          | b200_step_tally
 *     ExtendFromSecondariesAction (track/ExtendFromSecondariesAction.cc:55-100)| b200_step_extend_from_secondaries
 *   Stepper<device> (celeritas/global/Stepper.hh:82-190)           | b200_stepper_*
 *   reseed_rng (celeritas/random/RngReseed.cu:29-74)               | b200_reseed
 *   celer-sim Runner/Transporter (app/celer-sim/Transporter.cc:84-179) | b200_run_events
 *
 * Return convention: 0 on success; a positive cudaError_t value for CUDA
 * failures; B200_ERR_* (>= 10000) for library errors. b200_last_error() gives
 * a message for the calling thread.
 *---------------------------------------------------------------------------*/
#ifndef CELERITAS_B200_H
#define CELERITAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

enum
{
    B200_OK = 0,
    B200_ERR_INVALID_ARGUMENT = 10001,
    B200_ERR_IMAGE = 10002,
    B200_ERR_INITIALIZER_CAPACITY = 10003, /* reference: CELER_VALIDATE in ExtendFromSecondariesAction.cc:89-95 */
    B200_ERR_GEOMETRY_LIMITS = 10004,
    B200_ERR_NO_DEVICE = 10005,
    B200_ERR_RUNTIME = 10006
};

/* One primary particle (reference: celeritas::Primary, phys/Primary.hh). */
typedef struct B200Primary
{
    uint32_t particle_id;
    uint32_t event_id;
    double energy; /* MeV */
    double pos[3]; /* cm */
    double dir[3];
    double time; /* s */
} B200Primary;

/* Result of one step iteration (reference: StepperResult, Stepper.hh:57-70). */
typedef struct B200StepperResult
{
    uint32_t generated;
    uint32_t queued;
    uint32_t active;
    uint32_t alive;
} B200StepperResult;

/* Whole-run tallies (reference: TransporterResult, app/celer-sim/Transporter.hh). */
typedef struct B200RunResult
{
    uint64_t num_steps;           /* sum over iterations of active tracks */
    uint64_t num_step_iterations;
    uint64_t num_primaries;
    uint64_t max_queued;
    double seconds;               /* device time of the transport loop */
    uint64_t num_tracks;          /* tracks created (sum of the per-event track counters) */
    uint64_t num_aborted;         /* alive + queued tracks left when max_steps was hit */
} B200RunResult;

/* Stepper construction options (reference: StepperInput, Stepper.hh:42-54, plus the
 * celer-sim diagnostics of app/celer-sim/Runner.cc:616-633). */
typedef struct B200StepperOptions
{
    uint32_t stream_id;
    uint32_t num_track_slots;
    int action_times;              /* time every action with CUDA events */
    int action_diagnostic;         /* add an ActionDiagnostic unless the problem has one */
    uint32_t step_diagnostic_bins; /* >0: add a StepDiagnostic with this 'max' bin */
    /* Iterations with at most this many active tracks run pre-step..tally as one fused
     * launch (b200_step_fused). 0: library default; 0xffffffff: never fuse. */
    uint32_t fuse_threshold;
    /* b200_stepper_advance and b200_run_events run iterations with at most this many active
     * tracks inside the device-resident loop (b200_step_tail_loop). 0: library default;
     * 0xffffffff: never (one host round trip per iteration, as the reference). */
    uint32_t tail_threshold;
} B200StepperOptions;

/* Opaque views holding device pointers (layout: celeritas_b200/csrc/views.cuh). */
typedef struct B200ParamsView B200ParamsView;
typedef struct B200StateView B200StateView;
/* Opaque host-side objects. */
typedef struct B200Params B200Params;
typedef struct B200State B200State;
typedef struct B200Stepper B200Stepper;

char const* b200_last_error(void);
int b200_device_count(void);

/*--- problem parameters ---------------------------------------------------*/
/* Load a flattened problem image (celeritas_b200/host/Image.hh) into HBM. */
int b200_params_create_from_image(char const* image_path, B200Params** out);
/* The same from an image held in host memory (Image::serialize): the hand-off from a live
 * CoreParams::host_ref() on the reference side (src/celeritas/global/CoreParams.hh:155-172)
 * without a file in between; the bytes are not referenced after the call returns. */
int b200_params_create_from_memory(void const* image, size_t size, B200Params** out);
/* Native ORANGE construction (reference: OrangeParams(OrangeInput&&),
 * src/orange/OrangeParams.cc:137-209 with UnitInserter / RectArrayInserter / BIHBuilder): the
 * geometry image of an .org.json file, column for column what the reference builds. The image
 * bytes are malloc'ed (free with b200_string_free); host-only, needs no GPU. */
int b200_orange_build_image(char const* org_json_path, void** image, size_t* size);
/* Physics-data reader (reference: RootImporter::operator(), src/celeritas/ext/RootImporter.cc,
 * reading what RootExporter.cc:47-76 wrote): the `celeritas::ImportData` entry
 * (src/celeritas/io/ImportData.hh:55-112) of a reference physics export (.root), decoded
 * without the ROOT library and returned as a JSON document with the reference's member
 * names (optical_* members are skipped). The string is malloc'ed (free with
 * b200_string_free); host-only, needs no GPU. */
int b200_import_root(char const* root_path, char** import_data_json);
/* Geometry-only problem (navigation, b200_geo_trace) straight from an .org.json file. */
int b200_params_create_from_org_json(char const* org_json_path, B200Params** out);
void b200_params_destroy(B200Params* params);
B200ParamsView const* b200_params_view(B200Params const* params);
/* Problem metadata */
uint32_t b200_params_num_actions(B200Params const* params);
char const* b200_params_action_label(B200Params const* params, uint32_t action_id);
uint32_t b200_params_num_volumes(B200Params const* params);
char const* b200_params_volume_label(B200Params const* params, uint32_t volume_id);
uint32_t b200_params_num_detectors(B200Params const* params);
uint32_t b200_params_find_particle(B200Params const* params, int pdg); /* 0xffffffff if absent */
uint32_t b200_params_num_particles(B200Params const* params);
/* Discrete models are the actions [model_action_begin, model_action_begin + num_models) */
uint32_t b200_params_num_models(B200Params const* params);
uint32_t b200_params_model_action_begin(B200Params const* params);
/* Deepest universe nesting of the geometry (OrangeParams::max_depth, src/orange/OrangeParams.hh) */
uint32_t b200_params_max_depth(B200Params const* params);

/*--- per-stream state -------------------------------------------------------*/
int b200_state_create(B200Params const* params,
                      uint32_t stream_id,
                      uint32_t num_track_slots,
                      B200State** out);
void b200_state_destroy(B200State* state);
B200StateView const* b200_state_view(B200State const* state);
/* Copy one per-slot field to the host (names as in oracle/celerref.py FIELDS). */
int b200_state_get(B200State* state, char const* field, void* out);
/* Per-detector energy deposition [MeV] */
int b200_state_calo_get(B200State* state, double* out);
int b200_state_calo_clear(B200State* state);

/*--- step actions (asynchronous on `stream`) --------------------------------*/
int b200_step_extend_from_primaries(B200StateView const* state,
                                    B200Primary const* d_primaries,
                                    uint32_t const* d_rank_in_event,
                                    uint32_t const* d_event_ids,
                                    uint32_t const* d_event_counts,
                                    uint32_t const* d_neutral_inclusive, /* [n] number of neutral primaries in [0, i]; required for TrackOrder::init_charge problems, else may be NULL */
                                    uint32_t num_events,
                                    uint32_t n,
                                    cudaStream_t stream);
int b200_step_initialize_tracks(B200ParamsView const*, B200StateView const*, cudaStream_t);
int b200_step_pre_step(B200ParamsView const*, B200StateView const*, cudaStream_t);
int b200_step_along_step(B200ParamsView const*, B200StateView const*, cudaStream_t);
int b200_step_discrete_select(B200ParamsView const*, B200StateView const*, cudaStream_t);
int b200_step_interact(B200ParamsView const*, B200StateView const*, cudaStream_t);
int b200_step_boundary(B200ParamsView const*, B200StateView const*, cudaStream_t);
int b200_step_tracking_cut(B200ParamsView const*, B200StateView const*, cudaStream_t);
int b200_step_tally(B200ParamsView const*, B200StateView const*, cudaStream_t);
/* Step/hit output of the step that just ran (reference: StepCollector + DetectorSteps,
 * src/celeritas/user/DetectorSteps.cu:150-200): compact records, in slot order, of the tracks
 * that started the step in a sensitive volume ("hits.volumes" of the image). Called by
 * b200_step_tally / b200_step_post_tail; read back with b200_stepper_hits_count / _get. */
int b200_step_gather_hits(B200ParamsView const*, B200StateView const*, cudaStream_t);
int b200_step_extend_from_secondaries(B200ParamsView const*, B200StateView const*, cudaStream_t);
/* SortTracksAction::step (src/celeritas/track/SortTracksAction.cc:100-131) for problems whose
 * track order is one of TrackOrder::reindex_* (src/celeritas/Types.hh:151-171): rebuilds the
 * permutation of all track slots sorted by the key of `track_order` (3 = reindex_status,
 * 4 = reindex_particle_type, 5 = reindex_along_step_action, 6 = reindex_step_limit_action; the
 * reference's enum values) and the first index of every key (action_thread_offsets,
 * back-filled). Read back with b200_state_get "sort_slots" / "sort_offsets". */
int b200_step_sort_tracks(B200ParamsView const*, B200StateView const*, uint32_t track_order,
                          cudaStream_t);
/* ActionDiagnostic (user/detail/ActionDiagnosticExecutor.hh:30-65, order post) and
 * StepDiagnostic (user/detail/StepDiagnosticExecutor.hh:28-60, order user_post); the
 * state must have been created with the diagnostic enabled (B200StepperOptions). */
/* All of pre-step .. tally/diagnostics for every active track in ONE launch (identical
 * results: those actions only touch their own slot). The stepper uses it for small
 * iterations, where launch latency dominates; see B200StepperOptions::fuse_threshold. */
int b200_step_fused(B200ParamsView const*, B200StateView const*, cudaStream_t);
/* boundary + tracking-cut + action diagnostic + tally + step diagnostic (consecutive in the
 * action sequence) in one launch; what the stepper uses for them in large iterations. */
int b200_step_post_tail(B200ParamsView const*, B200StateView const*, cudaStream_t);
/* along-step immediately followed, in the same threads, by the discrete-process selection
 * (the next action in the sequence) and the per-model list build: replaces
 * b200_step_along_step + b200_step_discrete_select in large iterations. */
int b200_step_along_select(B200ParamsView const*, B200StateView const*, cudaStream_t);
int b200_step_action_diagnostic(B200ParamsView const*, B200StateView const*, cudaStream_t);
int b200_step_step_diagnostic(B200ParamsView const*, B200StateView const*, cudaStream_t);
/* Device-resident step loop for small iterations: up to `max_iterations` WHOLE step
 * iterations (start tracks, the fused step, extend-from-secondaries) in ONE cooperative
 * launch of `num_blocks` blocks, i.e. Stepper::operator() (global/Stepper.cc:124-140)
 * repeated without returning to the host. The loop leaves when no track is left, when the
 * next iteration could hold more than `exit_active` tracks, on a device error, or after
 * max_iterations. `ring` ([max_iterations][B200_TAIL_RING_WORDS]) and `done` ([2]) are
 * device pointers to MAPPED HOST memory: the counters of every iteration (generated,
 * initializers, vacancies, active, secondaries, alive, charged, neutral, first busy block,
 * error, device timer lo/hi) and {iterations done, exit reason (B200TailExit)}. */
#define B200_TAIL_RING_WORDS 16
typedef enum B200TailExit
{
    B200_TAIL_EXIT_DONE = 0,
    B200_TAIL_EXIT_MAX_ITERATIONS = 1,
    B200_TAIL_EXIT_TOO_MANY = 2,
    B200_TAIL_EXIT_ERROR = 3
} B200TailExit;
int b200_step_tail_loop(B200ParamsView const*,
                        B200StateView const*,
                        uint32_t num_blocks,
                        uint32_t max_iterations,
                        uint32_t exit_active,
                        uint32_t* ring,
                        uint32_t* done,
                        cudaStream_t);
/* Largest grid of the loop's kernel that is resident at once on the current device */
int b200_tail_max_blocks(B200ParamsView const*, int* num_blocks);
int b200_reseed(B200ParamsView const*, B200StateView const*, uint64_t event_id, cudaStream_t);
int b200_reset_generated(B200StateView const*, cudaStream_t);
int b200_kill_active(B200ParamsView const*, B200StateView const*, cudaStream_t);

/*--- ORANGE navigation on fixed ray sets -------------------------------------------*/
/* Reference: OrangeTrackView::{operator=(Initializer), find_next_step, move_to_boundary,
 * cross_boundary, find_safety, volume_id, surface_id} (orange/OrangeTrackView.hh:265-842).
 * Per ray: up to max_segments records {volume id, surface id crossed, distance};
 * count[i] = number of segments (bit 31 set: a crossing failed; 0xffffffff: could not
 * locate the origin); safety[i] = safety distance at the origin (-1 if outside).
 * Device-pointer version (asynchronous) and host-buffer convenience version. */
int b200_geo_trace(B200ParamsView const* params,
                   B200StateView const* state,
                   double const* d_pos,
                   double const* d_dir,
                   uint32_t num_rays,
                   uint32_t max_segments,
                   uint32_t* d_volume,
                   uint32_t* d_surface,
                   double* d_distance,
                   uint32_t* d_count,
                   double* d_safety,
                   cudaStream_t stream);
int b200_geo_trace_host(B200Params const* params,
                        double const* pos,
                        double const* dir,
                        uint32_t num_rays,
                        uint32_t max_segments,
                        uint32_t* volume,
                        uint32_t* surface,
                        double* distance,
                        uint32_t* count,
                        double* safety);

/* Total kernel launches issued by this library in this process */
uint64_t b200_launch_count(void);

/*--- stepper (owns a state and the ordered action sequence) ------------------*/
int b200_stepper_create(B200Params const* params,
                        uint32_t stream_id,
                        uint32_t num_track_slots,
                        B200Stepper** out);
int b200_stepper_create_opts(B200Params const* params,
                             B200StepperOptions const* options,
                             B200Stepper** out);
void b200_stepper_destroy(B200Stepper* stepper);
B200State* b200_stepper_state(B200Stepper* stepper);
/* Diagnostic tallies: counts[particle][bin]. Action diagnostic: bin = action id, num_bins
 * = b200_stepper_num_actions(); step diagnostic: num_bins = step_diagnostic_bins + 2.
 * Returns B200_ERR_INVALID_ARGUMENT if the diagnostic is not enabled. */
uint32_t b200_stepper_num_actions(B200Stepper const* stepper);
char const* b200_stepper_action_label(B200Stepper const* stepper, uint32_t action_id);
int b200_stepper_action_diagnostic_get(B200Stepper* stepper, uint32_t* counts);
int b200_stepper_step_diagnostic_get(B200Stepper* stepper, uint32_t* counts);
uint32_t b200_stepper_step_diagnostic_bins(B200Stepper const* stepper);
int b200_stepper_diagnostics_clear(B200Stepper* stepper);
/* One step iteration; host primaries may be NULL/0. Synchronises to read counters. */
int b200_stepper_step(B200Stepper* stepper,
                      B200Primary const* primaries,
                      uint32_t num_primaries,
                      B200StepperResult* result);
/* Up to `max_iterations` step iterations without primaries. While few tracks are left
 * they run inside the device-resident loop (b200_step_tail_loop), otherwise one by one as
 * b200_stepper_step does; either way results[i] is iteration i's StepperResult and the
 * state afterwards is the same. Stops early when no track is left. *num_done <=
 * max_iterations iterations were taken. */
int b200_stepper_advance(B200Stepper* stepper,
                         uint32_t max_iterations,
                         B200StepperResult* results,
                         uint32_t* num_done);
/* One iteration driven by an EXTERNAL action sequence: the reference's own ActionSequence
 * (src/celeritas/global/ActionSequence.cc:77-138) calling one adapter per action
 * (celeritas_b200/adapter/B200Actions.cc). insert = ExtendFromPrimariesAction::insert
 * (track/ExtendFromPrimariesAction.cc:103-131), stages host primaries; begin_iteration sizes
 * the iteration's launches from the previous counters and runs extend-from-primaries; the
 * caller then calls the b200_step_* launchers with b200_stepper_state()'s view on
 * b200_stepper_stream(); end_iteration (after b200_step_extend_from_secondaries) returns the
 * iteration's counters as soon as the device has published them. */
int b200_stepper_insert(B200Stepper* stepper, B200Primary const* primaries, uint32_t num_primaries);
int b200_stepper_begin_iteration(B200Stepper* stepper);
int b200_stepper_end_iteration(B200Stepper* stepper, B200StepperResult* result);
cudaStream_t b200_stepper_stream(B200Stepper* stepper);
/* Hits of the last step iteration (reference: DetectorStepOutput, user/DetectorSteps.hh:38-72).
 * Fields: detector track_id event_id parent_id track_step_count particle (uint32), step_length
 * energy_deposition pre_time pre_energy post_time post_energy (double), pre_pos pre_dir post_pos
 * post_dir (double[3] per hit). `out` holds *count elements of the field. */
int b200_stepper_hits_count(B200Stepper* stepper, uint32_t* count);
int b200_stepper_hits_get(B200Stepper* stepper, char const* field, void* out);
/* Iterations that ran inside the device-resident loop since the stepper was created */
uint64_t b200_stepper_tail_iterations(B200Stepper const* stepper);
int b200_stepper_warm_up(B200Stepper* stepper);
int b200_stepper_reseed(B200Stepper* stepper, uint64_t event_id);
int b200_stepper_kill_active(B200Stepper* stepper);
uint32_t b200_stepper_num_step_actions(B200Stepper const* stepper);
char const* b200_stepper_step_action_label(B200Stepper const* stepper, uint32_t i);
/* Per-action device timing with CUDA events on the stepper's stream (reference:
 * StepperInput::action_times, celeritas/global/ActionSequence.cc:99-121). */
int b200_stepper_set_action_times(B200Stepper* stepper, int enable);
double b200_stepper_action_time(B200Stepper const* stepper, uint32_t i); /* seconds */
/* Select the CUDA device for objects created afterwards by this thread. */
int b200_set_device(int device);
/* Kernel launches issued by this stepper so far */
uint64_t b200_stepper_launch_count(B200Stepper const* stepper);

/*--- whole events (celer-sim Transporter loop) -------------------------------*/
/* Event e owns primaries [offsets[e], offsets[e+1]); host buffers. max_steps limits the
 * step ITERATIONS per transport call as in app/celer-sim/Transporter.cc:133-141
 * (0: unlimited); tracks left over are counted in num_aborted and the state is reset. */
int b200_run_events(B200Stepper* stepper,
                    B200Primary const* primaries,
                    uint32_t const* offsets,
                    uint32_t num_events,
                    int merge_events,
                    uint64_t max_steps,
                    B200RunResult* result);

/* The same over several streams at once, one host thread per stepper (reference: one
 * OpenMP thread per stream, app/celer-sim/celer-sim.cc:120-137). Event e is transported
 * by stepper e % num_streams; with merge_events every stepper transports its events as
 * one batch. results[num_streams]; *seconds = device time of the whole call. */
int b200_run_events_streams(B200Stepper* const* steppers,
                            uint32_t num_streams,
                            B200Primary const* primaries,
                            uint32_t const* offsets,
                            uint32_t num_events,
                            int merge_events,
                            uint64_t max_steps,
                            B200RunResult* results,
                            double* seconds);

/*--- celer-sim front end ------------------------------------------------------*/
/* Primaries from a celer-sim "primary_options" JSON object (reference:
 * celeritas/phys/PrimaryGenerator.cc:30-110, PrimaryGeneratorOptionsIO.json.cc): same
 * std::mt19937 stream, same sampling order, so the primaries are identical to the
 * reference's. `out` receives num_events * primaries_per_event records (query the count
 * with out == NULL). */
int b200_primaries_generate(B200Params const* params,
                            char const* primary_options_json,
                            B200Primary* out,
                            uint64_t capacity,
                            uint64_t* count,
                            uint32_t* primaries_per_event);
/* Run a celer-sim input (app/celer-sim/RunnerInputIO.json.cc:40-139; the problem comes
 * from "image_file") and write the celer-sim-style JSON report
 * (app/celer-sim/RunnerOutput.cc:37-115, plus the diagnostics' and SimpleCalo's output)
 * into a newly allocated string: release it with b200_string_free(). */
int b200_celer_sim_run(char const* input_json, char** report);
void b200_string_free(char* s);

#ifdef __cplusplus
}
#endif
#endif /* CELERITAS_B200_H */
