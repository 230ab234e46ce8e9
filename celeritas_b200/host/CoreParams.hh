//---------------------------------------------------------------------------//
// CoreParams: immutable problem data resident in HBM.
//
// Mirrors the role of the reference's CoreParams
// (/root/reference/src/celeritas/global/CoreParams.hh:42-146): it owns the
// device copies of geometry, materials, particles, cutoffs, physics tables,
// model data, RNG and sim parameters, plus the action table (ids, labels,
// step order) that the action sequence is built from.
//---------------------------------------------------------------------------//
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "../csrc/views.cuh"
#include "DeviceMemory.hh"
#include "Image.hh"

namespace celeritas_b200
{
//! Ordering of step actions (reference StepActionOrder, ActionInterface.hh:29-46)
enum class StepActionOrder : uint32_t
{
    generate,
    start,
    user_start,
    sort_start,
    pre,
    user_pre,
    sort_pre,
    along,
    sort_along,
    pre_post,
    sort_pre_post,
    post,
    user_post,
    end,
    size_
};

struct ActionRecord
{
    uint32_t id;
    std::string label;
    uint32_t order;  // StepActionOrder or 0xffffffff for implicit actions
};

class CoreParams
{
  public:
    static std::shared_ptr<CoreParams> from_image(std::string const& path);
    static std::shared_ptr<CoreParams> from_image(b200::Image const& img);

    b200::ParamsView const& view() const { return view_; }
    std::vector<ActionRecord> const& actions() const { return actions_; }
    std::vector<std::string> const& volume_labels() const { return volume_labels_; }
    std::vector<std::string> const& particle_names() const { return particle_names_; }
    std::vector<int> const& particle_pdg() const { return particle_pdg_; }
    std::vector<std::string> const& detector_volumes() const { return detector_volumes_; }
    //! Device table: global volume id -> detector id (empty if no detectors)
    uint32_t const* detector_of_volume() const { return d_detector_of_volume_; }
    uint32_t num_detectors() const { return detector_volumes_.size(); }
    //! Sensitive volumes of the step/hit output (detector id = position) and their map
    std::vector<std::string> const& hit_volumes() const { return hit_volumes_; }
    uint32_t const* hit_detector_of_volume() const { return d_hit_detector_of_volume_; }
    bool hits_nonzero_edep() const { return hits_nonzero_edep_; }

    uint32_t init_capacity() const { return init_capacity_; }
    uint32_t max_events() const { return max_events_; }
    uint32_t rng_seed() const { return view_.rng.seed; }
    uint32_t find_particle(int pdg) const;
    //! Whether the problem's action table has an action with this label
    bool has_action(std::string const& label) const;
    bool particle_is_neutral(uint32_t particle_id) const
    {
        return particle_id < particle_charge_.size() && particle_charge_[particle_id] == 0;
    }
    size_t device_bytes() const { return arena_.bytes(); }

    //!@{
    //! Run options that celer-sim takes from its input rather than from the problem
    //! (app/celer-sim/Runner.cc:412-440); set before any state is created
    void rng_seed(uint32_t seed) { view_.rng.seed = seed; }
    void init_capacity(uint32_t capacity);
    void max_events(uint32_t num_events);
    //! Track order (b200::TrackOrder); slot assignment of starting tracks depends on it
    void track_order(uint32_t order);
    uint32_t track_order() const { return view_.scalars.track_order; }
    //! True once a CoreState has been built on these params: the run options above are
    //! frozen from then on (CoreState sizes its arrays from them)
    void freeze() const { frozen_ = true; }
    //! Uniform field [T]; only for problems built with the uniform-field along-step
    void uniform_field_tesla(double const (&field)[3]);
    bool has_uniform_field() const { return view_.model.field.enabled != 0; }
    //!@}

  private:
    CoreParams() = default;
    void load(b200::Image const& img);

    DeviceArena arena_;
    b200::ParamsView view_{};
    std::vector<ActionRecord> actions_;
    std::vector<std::string> volume_labels_;
    std::vector<std::string> particle_names_;
    std::vector<int> particle_pdg_;
    std::vector<double> particle_charge_;
    std::vector<std::string> detector_volumes_;
    uint32_t const* d_detector_of_volume_{nullptr};
    std::vector<std::string> hit_volumes_;
    uint32_t const* d_hit_detector_of_volume_{nullptr};
    bool hits_nonzero_edep_{false};
    uint32_t init_capacity_{0};
    uint32_t max_events_{0};
    mutable bool frozen_{false};
};
}  // namespace celeritas_b200
