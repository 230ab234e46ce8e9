//---------------------------------------------------------------------------//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// A "Problem" is the reference's CoreParams built the way celer-sim builds it
// (/root/reference/app/celer-sim/Runner.cc:281-442), except that the physics
// ImportData comes from the JSON fixtures under data/physics (decoded from the
// reference's own .root exports by tools/rootlite.py) instead of ROOT/Geant4,
// neither of which exists in this image.
//---------------------------------------------------------------------------//
#pragma once

#include <memory>
#include <string>
#include <vector>
#include <nlohmann/json.hpp>

#include "celeritas/em/params/FluctuationParams.hh"
#include "celeritas/em/params/UrbanMscParams.hh"
#include "celeritas/field/RZMapFieldInput.hh"
#include "celeritas/field/UniformFieldData.hh"
#include "orange/OrangeParams.hh"
#include "celeritas/global/CoreParams.hh"
#include "celeritas/io/ImportData.hh"
#include "celeritas/user/ActionDiagnostic.hh"
#include "celeritas/user/SimpleCalo.hh"
#include "celeritas/user/StepDiagnostic.hh"
#include "celeritas/user/DetectorSteps.hh"
#include "celeritas/user/StepCollector.hh"
#include "celeritas/user/StepInterface.hh"

namespace celerref
{
//! Step/hit output of the reference, kept per step: a StepInterface whose callback runs the
//! reference's own compaction (copy_steps, user/DetectorSteps.{cc,cu}) and keeps the result
//! ("hit_volumes": [...] and "hits_nonzero_edep": bool in the problem configuration)
class HitRecorder final : public celeritas::StepInterface
{
  public:
    HitRecorder(std::vector<celeritas::VolumeId> volumes, bool nonzero_edep)
        : volumes_(std::move(volumes)), nonzero_edep_(nonzero_edep)
    {
    }
    Filters filters() const final;
    celeritas::StepSelection selection() const final;
    void process_steps(HostStepState) final;
    void process_steps(DeviceStepState) final;
    //! Hits of the last step of the given stream
    celeritas::DetectorStepOutput const& last(unsigned stream = 0) const;

  private:
    std::vector<celeritas::VolumeId> volumes_;
    bool nonzero_edep_;
    std::vector<celeritas::DetectorStepOutput> last_{8};
};

struct Problem
{
    nlohmann::json config;
    celeritas::ImportData imported;
    std::shared_ptr<celeritas::CoreParams> core;
    //! Geometry-only problems ("problem": "geometry"): no physics, core is null
    std::shared_ptr<celeritas::OrangeParams const> geo;
    std::shared_ptr<celeritas::SimpleCalo> calo;
    std::shared_ptr<celeritas::StepCollector> collector;
    std::shared_ptr<HitRecorder> hits;
    std::vector<std::string> hit_volumes;
    bool hits_nonzero_edep{false};
    std::vector<std::string> calo_volumes;
    //! celer-sim diagnostics ("action_diagnostic": true, "step_diagnostic_bins": N)
    std::shared_ptr<celeritas::ActionDiagnostic> action_diag;
    std::shared_ptr<celeritas::StepDiagnostic> step_diag;
    std::shared_ptr<celeritas::UrbanMscParams const> msc;
    std::shared_ptr<celeritas::FluctuationParams const> fluct;
    celeritas::UniformFieldParams field;
    //! "field_map": RZ field map file (field/RZMapFieldInput.hh) instead of a uniform field
    celeritas::RZMapFieldInput rz_field;
    bool has_rz_field{false};
    bool has_msc{false};
    bool has_fluct{false};
    bool has_field{false};
};

// Fill ImportData from the rootlite JSON schema
void import_from_json(nlohmann::json const& j, celeritas::ImportData* out);

// Build from a JSON configuration string
std::unique_ptr<Problem> build_problem(nlohmann::json const& config);

// Write the flattened problem image consumed by the B200 loader
void export_image(Problem const& p, std::string const& path);
// The same image as one byte string (b200_params_create_from_memory)
std::vector<unsigned char> export_image_bytes(Problem const& p);
}  // namespace celerref
