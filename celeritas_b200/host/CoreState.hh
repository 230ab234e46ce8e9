//---------------------------------------------------------------------------//
// CoreState: all mutable per-stream data, resident in HBM as SoA.
//
// Mirrors the reference's CoreState<MemSpace::device>
// (/root/reference/src/celeritas/global/CoreState.hh:73-182): one per stream,
// sized by num_track_slots; owns track state, the initializer queue, RNG
// state and the (here device-resident) CoreStateCounters.
//---------------------------------------------------------------------------//
#pragma once

#include <memory>

#include "../csrc/views.cuh"
#include "CoreParams.hh"
#include "DeviceMemory.hh"

namespace celeritas_b200
{
//! Host copy of the device counters (reference: track/CoreStateCounters.hh:24-47)
struct CoreStateCounters
{
    uint32_t num_generated{0};
    uint32_t num_initializers{0};
    uint32_t num_vacancies{0};
    uint32_t num_active{0};
    uint32_t num_secondaries{0};
    uint32_t num_alive{0};
    uint32_t num_charged{0};  //!< active charged tracks in the dense list
    uint32_t num_neutral{0};  //!< active neutral tracks in the dense list
    //! First 128-slot block that held a track during the last step (INVALID: none)
    uint32_t first_busy_block{0};
};

class CoreState
{
  public:
    CoreState(std::shared_ptr<CoreParams const> params, uint32_t stream_id, uint32_t num_track_slots);
    ~CoreState();
    CoreState(CoreState const&) = delete;
    CoreState& operator=(CoreState const&) = delete;

    b200::StateView const& view() const { return view_; }
    //! Declare which single event is in flight (INVALID: several / unknown)
    void single_event(uint32_t event_id) { view_.single_event = event_id; }
    //! Upper bounds used to size the next iteration's grids
    void launch_hints(uint32_t active,
                      uint32_t charged,
                      uint32_t neutral,
                      uint32_t fresh,
                      uint32_t slot_begin = 0)
    {
        view_.hint_active = active;
        view_.hint_charged = charged;
        view_.hint_neutral = neutral;
        view_.hint_new = fresh;
        view_.slot_begin = slot_begin;
        view_.iteration_seq = ++iteration_seq_;
    }
    uint32_t size() const { return view_.num_slots; }
    uint32_t stream_id() const { return stream_id_; }
    cudaStream_t stream() const { return stream_; }
    CoreParams const& params() const { return *params_; }

    //! Copy device counters to the host (synchronises the stream)
    CoreStateCounters sync_counters();
    //! Counters of the step iteration launched after the last launch_hints(), as soon as
    //! the end-of-step scan has published them (the stream may still be running the last
    //! pass); falls back to sync_counters() if the kernels did not publish
    CoreStateCounters wait_counters();
    //! Nonzero if a kernel flagged an error (B200_ERR_*)
    uint32_t last_device_error() const { return last_error_; }

    //! Copy a named per-slot field to host memory
    void get_field(std::string const& name, void* out);
    //! Step/hit output of the last step iteration (csrc/kernels_sort.cu: k_hits_gather)
    uint32_t hits_count();
    void hits_get(std::string const& field, void* out);
    void calo_get(double* out);
    void calo_clear();

    //! Allocate the [particle][action] tally of the action diagnostic
    void enable_action_diagnostic(uint32_t num_actions);
    //! Allocate the [particle][max_bin + 2] tally of the step diagnostic
    void enable_step_diagnostic(uint32_t max_step_bin);
    //! Copy a diagnostic tally to the host (synchronises the stream)
    void diagnostic_get(bool steps, uint32_t* out);
    void diagnostics_clear();
    //! Sum of the per-event track-id counters: tracks created since the last reseed
    uint64_t num_tracks();
    //! Back to the freshly constructed state (reference: CoreState::reset)
    void reset();

    size_t device_bytes() const { return arena_.bytes(); }

    //!@{
    //! Device-resident step loop (csrc/tail.cu): mapped host ring of per-iteration counters
    static constexpr uint32_t tail_ring_capacity = 1024;
    uint32_t* tail_ring_device() const { return d_tail_ring_; }
    uint32_t* tail_done_device() const { return d_tail_done_; }
    uint32_t const* tail_ring_host() const { return h_tail_ring_; }
    uint32_t volatile* tail_done_host() const { return h_tail_done_; }
    //! Counters of one ring entry, as sync_counters() would have returned them
    CoreStateCounters unpack_tail_entry(uint32_t index);
    //!@}

  private:
    std::shared_ptr<CoreParams const> params_;
    uint32_t stream_id_;
    cudaStream_t stream_{nullptr};
    DeviceArena arena_;
    b200::StateView view_{};
    uint32_t* h_counters_{nullptr};  // pinned, mapped: [CTR_SIZE + 1]
    uint32_t* h_tail_ring_{nullptr};  // pinned, mapped: [tail_ring_capacity][RING_WORDS]
    uint32_t* h_tail_done_{nullptr};  // pinned, mapped: [2]
    uint32_t* d_tail_ring_{nullptr};
    uint32_t* d_tail_done_{nullptr};
    uint32_t iteration_seq_{0};
    CoreStateCounters unpack_counters();
    uint32_t last_error_{0};
};
}  // namespace celeritas_b200
