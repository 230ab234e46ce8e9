//---------------------------------------------------------------------------//
// TEST INFRASTRUCTURE ONLY. Builds the reference's CoreParams from JSON.
// Follows /root/reference/app/celer-sim/Runner.cc:281-442 (build_core_params)
// and, for "simple-compton", /root/reference/test/celeritas/SimpleTestBase.cc:37-216.
//---------------------------------------------------------------------------//
#include "Problem.hh"

#include <fstream>
#include <map>

#include "corecel/io/Logger.hh"
#include "corecel/io/OutputRegistry.hh"
#include "corecel/sys/ActionRegistry.hh"
#include "corecel/data/AuxParamsRegistry.hh"
#include "geocel/UnitUtils.hh"
#include "orange/OrangeParams.hh"
#include "celeritas/Quantities.hh"
#include "celeritas/Units.hh"
#include "celeritas/em/params/FluctuationParams.hh"
#include "celeritas/em/params/UrbanMscParams.hh"
#include "celeritas/em/params/WentzelOKVIParams.hh"
#include "celeritas/em/process/ComptonProcess.hh"
#include "celeritas/field/UniformFieldData.hh"
#include "celeritas/geo/GeoMaterialParams.hh"
#include "celeritas/geo/GeoParams.hh"
#include "celeritas/global/alongstep/AlongStepGeneralLinearAction.hh"
#include "celeritas/global/alongstep/AlongStepNeutralAction.hh"
#include "celeritas/field/RZMapFieldInputIO.json.hh"
#include "celeritas/global/alongstep/AlongStepRZMapFieldMscAction.hh"
#include "celeritas/global/alongstep/AlongStepUniformMscAction.hh"
#include "celeritas/io/detail/ImportDataConverter.hh"
#include "celeritas/mat/MaterialParams.hh"
#include "celeritas/phys/CutoffParams.hh"
#include "celeritas/phys/ImportedProcessAdapter.hh"
#include "celeritas/phys/PDGNumber.hh"
#include "celeritas/phys/ParticleParams.hh"
#include "celeritas/phys/PhysicsParams.hh"
#include "celeritas/phys/ProcessBuilder.hh"
#include "celeritas/random/RngParams.hh"
#include "celeritas/track/SimParams.hh"
#include "celeritas/track/TrackInitParams.hh"

using namespace celeritas;
using json = nlohmann::json;

namespace celerref
{
namespace
{
//---------------------------------------------------------------------------//
ImportPhysicsVector vec_from_json(json const& j)
{
    ImportPhysicsVector v;
    v.vector_type = static_cast<ImportPhysicsVectorType>(
        j.at("vector_type").get<int>());
    v.x = j.at("x").get<std::vector<double>>();
    v.y = j.at("y").get<std::vector<double>>();
    return v;
}

ImportPhysicsTable table_from_json(json const& j)
{
    ImportPhysicsTable t;
    t.table_type = static_cast<ImportTableType>(j.at("table_type").get<int>());
    t.x_units = static_cast<ImportUnits>(j.at("x_units").get<int>());
    t.y_units = static_cast<ImportUnits>(j.at("y_units").get<int>());
    for (auto const& v : j.at("physics_vectors"))
    {
        t.physics_vectors.push_back(vec_from_json(v));
    }
    return t;
}

ImportPhysics2DVector vec2d_from_json(json const& j)
{
    ImportPhysics2DVector v;
    v.x = j.at("x").get<std::vector<double>>();
    v.y = j.at("y").get<std::vector<double>>();
    v.value = j.at("value").get<std::vector<double>>();
    return v;
}

std::string resolve(json const& cfg, std::string const& key)
{
    std::string p = cfg.at(key).get<std::string>();
    if (!p.empty() && p[0] != '/' && cfg.contains("base_dir"))
    {
        p = cfg.at("base_dir").get<std::string>() + "/" + p;
    }
    return p;
}
}  // namespace

//---------------------------------------------------------------------------//
void import_from_json(json const& j, ImportData* out)
{
    ImportData& d = *out;
    for (auto const& e : j.at("isotopes"))
    {
        ImportIsotope r;
        r.name = e.at("name");
        r.atomic_number = e.at("atomic_number");
        r.atomic_mass_number = e.at("atomic_mass_number");
        r.binding_energy = e.at("binding_energy");
        r.proton_loss_energy = e.at("proton_loss_energy");
        r.neutron_loss_energy = e.at("neutron_loss_energy");
        r.nuclear_mass = e.at("nuclear_mass");
        d.isotopes.push_back(r);
    }
    for (auto const& e : j.at("elements"))
    {
        ImportElement r;
        r.name = e.at("name");
        r.atomic_number = e.at("atomic_number");
        r.atomic_mass = e.at("atomic_mass");
        for (auto const& f : e.at("isotopes_fractions"))
        {
            r.isotopes_fractions.push_back(
                {f.at("first").get<unsigned>(), f.at("second").get<double>()});
        }
        d.elements.push_back(r);
    }
    for (auto const& e : j.at("geo_materials"))
    {
        ImportGeoMaterial r;
        r.name = e.at("name");
        r.state = static_cast<ImportMaterialState>(e.at("state").get<int>());
        r.temperature = e.at("temperature");
        r.number_density = e.at("number_density");
        for (auto const& c : e.at("elements"))
        {
            r.elements.push_back({c.at("element_id").get<unsigned>(),
                                  c.at("number_fraction").get<double>()});
        }
        d.geo_materials.push_back(r);
    }
    for (auto const& e : j.at("phys_materials"))
    {
        ImportPhysMaterial r;
        r.geo_material_id = e.at("geo_material_id");
        // optical physics is outside the EM track loop: never attach optical data
        r.optical_material_id = ImportPhysMaterial::unspecified;
        for (auto const& c : e.at("pdg_cutoffs"))
        {
            ImportProductionCut pc;
            pc.energy = c.at("second").at("energy");
            pc.range = c.at("second").at("range");
            r.pdg_cutoffs[c.at("first").get<int>()] = pc;
        }
        d.phys_materials.push_back(r);
    }
    for (auto const& e : j.at("regions"))
    {
        ImportRegion r;
        r.name = e.at("name");
        r.field_manager = e.at("field_manager");
        r.production_cuts = e.at("production_cuts");
        r.user_limits = e.at("user_limits");
        d.regions.push_back(r);
    }
    for (auto const& e : j.at("volumes"))
    {
        ImportVolume r;
        r.geo_material_id = e.at("geo_material_id");
        r.region_id = e.at("region_id");
        r.phys_material_id = e.at("phys_material_id");
        r.name = e.at("name");
        r.solid_name = e.at("solid_name");
        d.volumes.push_back(r);
    }
    for (auto const& e : j.at("particles"))
    {
        ImportParticle r;
        r.name = e.at("name");
        r.pdg = e.at("pdg");
        r.mass = e.at("mass");
        r.charge = e.at("charge");
        r.spin = e.at("spin");
        r.lifetime = e.at("lifetime");
        r.is_stable = e.at("is_stable");
        d.particles.push_back(r);
    }
    for (auto const& e : j.at("processes"))
    {
        ImportProcess r;
        r.particle_pdg = e.at("particle_pdg");
        r.secondary_pdg = e.at("secondary_pdg");
        r.process_type
            = static_cast<ImportProcessType>(e.at("process_type").get<int>());
        r.process_class
            = static_cast<ImportProcessClass>(e.at("process_class").get<int>());
        for (auto const& m : e.at("models"))
        {
            ImportModel im;
            im.model_class
                = static_cast<ImportModelClass>(m.at("model_class").get<int>());
            for (auto const& mm : m.at("materials"))
            {
                ImportModelMaterial imm;
                imm.energy = mm.at("energy").get<std::vector<double>>();
                imm.micro_xs
                    = mm.at("micro_xs").get<std::vector<std::vector<double>>>();
                im.materials.push_back(std::move(imm));
            }
            r.models.push_back(std::move(im));
        }
        for (auto const& t : e.at("tables"))
        {
            r.tables.push_back(table_from_json(t));
        }
        d.processes.push_back(std::move(r));
    }
    for (auto const& e : j.at("msc_models"))
    {
        ImportMscModel r;
        r.particle_pdg = e.at("particle_pdg");
        r.model_class
            = static_cast<ImportModelClass>(e.at("model_class").get<int>());
        r.xs_table = table_from_json(e.at("xs_table"));
        d.msc_models.push_back(std::move(r));
    }
    if (j.contains("sb_data"))
    {
        for (auto const& [k, v] : j.at("sb_data").items())
        {
            d.sb_data[std::stoi(k)] = vec2d_from_json(v);
        }
    }
    if (j.contains("livermore_pe_data"))
    {
        for (auto const& [k, v] : j.at("livermore_pe_data").items())
        {
            ImportLivermorePE pe;
            pe.xs_lo = vec_from_json(v.at("xs_lo"));
            pe.xs_hi = vec_from_json(v.at("xs_hi"));
            pe.thresh_lo = v.at("thresh_lo");
            pe.thresh_hi = v.at("thresh_hi");
            for (auto const& s : v.at("shells"))
            {
                ImportLivermoreSubshell sh;
                sh.binding_energy = s.at("binding_energy");
                sh.param_lo = s.at("param_lo").get<std::vector<double>>();
                sh.param_hi = s.at("param_hi").get<std::vector<double>>();
                sh.xs = s.at("xs").get<std::vector<double>>();
                sh.energy = s.at("energy").get<std::vector<double>>();
                pe.shells.push_back(std::move(sh));
            }
            d.livermore_pe_data[std::stoi(k)] = std::move(pe);
        }
    }
    {
        auto const& e = j.at("em_params");
        auto& p = d.em_params;
        p.energy_loss_fluct = e.at("energy_loss_fluct");
        p.lpm = e.at("lpm");
        p.integral_approach = e.at("integral_approach");
        p.linear_loss_limit = e.at("linear_loss_limit");
        p.lowest_electron_energy = e.at("lowest_electron_energy");
        p.auger = e.at("auger");
        p.msc_step_algorithm = static_cast<MscStepLimitAlgorithm>(
            e.at("msc_step_algorithm").get<int>());
        p.msc_range_factor = e.at("msc_range_factor");
        p.msc_safety_factor = e.at("msc_safety_factor");
        p.msc_lambda_limit = e.at("msc_lambda_limit");
        p.msc_theta_limit = e.at("msc_theta_limit");
        p.apply_cuts = e.at("apply_cuts");
        p.screening_factor = e.at("screening_factor");
        p.angle_limit_factor = e.at("angle_limit_factor");
        p.form_factor
            = static_cast<NuclearFormFactorType>(e.at("form_factor").get<int>());
    }
    {
        auto const& e = j.at("trans_params");
        for (auto const& [k, v] : e.at("looping").items())
        {
            ImportLoopingThreshold lt;
            lt.threshold_trials = v.at("threshold_trials");
            lt.important_energy = v.at("important_energy");
            d.trans_params.looping[std::stoi(k)] = lt;
        }
        d.trans_params.max_substeps = e.at("max_substeps");
    }
    d.units = j.at("units").get<std::string>();
}

//---------------------------------------------------------------------------//
namespace
{
// SimpleTestBase recipe: two boxes, hand-made Compton tables, no Geant4 data
void build_simple_compton(Problem& p, CoreParams::Input& params)
{
    using namespace celeritas::units;
    json const& cfg = p.config;
    params.geometry = std::make_shared<GeoParams>(resolve(cfg, "geometry_file"));
    {
        MaterialParams::Input inp;
        inp.elements = {{AtomicNumber{13}, AmuMass{27}, {}, "Al"}};
        inp.materials = {{native_value_from(MolCcDensity{0.1}),
                          293.0,
                          MatterState::solid,
                          {{ElementId{0}, 1.0}},
                          "Al"},
                         {0, 0, MatterState::unspecified, {}, "hard vacuum"}};
        params.material = std::make_shared<MaterialParams>(std::move(inp));
    }
    {
        GeoMaterialParams::Input input;
        input.geometry = params.geometry;
        input.materials = params.material;
        input.volume_to_mat = {MaterialId{0}, MaterialId{1}, MaterialId{}};
        input.volume_labels
            = {Label{"inner"}, Label{"world"}, Label{"[EXTERIOR]"}};
        params.geomaterial
            = std::make_shared<GeoMaterialParams>(std::move(input));
    }
    {
        using namespace constants;
        ParticleParams::Input defs;
        defs.push_back({"gamma",
                        pdg::gamma(),
                        zero_quantity(),
                        zero_quantity(),
                        stable_decay_constant});
        defs.push_back({"electron",
                        pdg::electron(),
                        MevMass{0.5},
                        ElementaryCharge{-1},
                        stable_decay_constant});
        params.particle = std::make_shared<ParticleParams>(std::move(defs));
    }
    {
        CutoffParams::Input input;
        input.materials = params.material;
        input.particles = params.particle;
        input.cutoffs = {
            {pdg::gamma(),
             {{MevEnergy{0.01}, 0.1 * millimeter},
              {MevEnergy{100}, 100 * centimeter}}},
            {pdg::electron(),
             {{MevEnergy{1000}, 1000 * centimeter},
              {MevEnergy{1000}, 1000 * centimeter}}},
        };
        params.cutoff = std::make_shared<CutoffParams>(std::move(input));
    }
    {
        PhysicsParams::Input input;
        input.options.secondary_stack_factor
            = cfg.value("secondary_stack_factor", 3.0);

        ImportProcess compton_data;
        compton_data.particle_pdg = pdg::gamma().get();
        compton_data.secondary_pdg = pdg::electron().get();
        compton_data.process_type = ImportProcessType::electromagnetic;
        compton_data.process_class = ImportProcessClass::compton;
        {
            ImportModel kn_model;
            kn_model.model_class = ImportModelClass::klein_nishina;
            kn_model.materials.resize(params.material->size());
            for (ImportModelMaterial& imm : kn_model.materials)
            {
                imm.energy = {1e-4, 1e8};
            }
            compton_data.models.push_back(std::move(kn_model));
        }
        {
            ImportPhysicsTable lambda;
            lambda.table_type = ImportTableType::lambda;
            lambda.x_units = ImportUnits::mev;
            lambda.y_units = ImportUnits::len_inv;
            lambda.physics_vectors = {
                {ImportPhysicsVectorType::log, {1e-4, 1.0}, {1e1, 1e0}},
                {ImportPhysicsVectorType::log, {1e-4, 1.0}, {1e-10, 1e-10}},
            };
            compton_data.tables.push_back(std::move(lambda));
        }
        {
            ImportPhysicsTable lambdap;
            lambdap.table_type = ImportTableType::lambda_prim;
            lambdap.x_units = ImportUnits::mev;
            lambdap.y_units = ImportUnits::len_mev_inv;
            lambdap.physics_vectors = {
                {ImportPhysicsVectorType::log,
                 {1.0, 1e4, 1e8},
                 {1e0, 1e-2, 1e-4}},
                {ImportPhysicsVectorType::log,
                 {1.0, 1e4, 1e8},
                 {1e-10, 1e-10, 1e-10}},
            };
            compton_data.tables.push_back(std::move(lambdap));
        }
        {
            celeritas::detail::ImportDataConverter convert{
                celeritas::UnitSystem::cgs};
            convert(&compton_data);
        }
        auto process_data = std::make_shared<ImportedProcesses>(
            std::vector<ImportProcess>{std::move(compton_data)});
        input.particles = params.particle;
        input.materials = params.material;
        input.processes = {std::make_shared<ComptonProcess>(input.particles,
                                                            process_data)};
        input.action_registry = params.action_reg.get();
        params.physics = std::make_shared<PhysicsParams>(std::move(input));
    }
    {
        SimParams::Input input;
        input.particles = params.particle;
        params.sim = std::make_shared<SimParams>(input);
    }
    {
        WentzelOKVIParams::Options options;
        params.wentzel
            = std::make_shared<WentzelOKVIParams>(params.material, options);
    }
    {
        auto result = std::make_shared<AlongStepNeutralAction>(
            params.action_reg->next_id());
        params.action_reg->insert(result);
    }
}

//---------------------------------------------------------------------------//
// celer-sim recipe with JSON-imported physics
void build_imported(Problem& p, CoreParams::Input& params)
{
    json const& cfg = p.config;
    {
        std::ifstream f(resolve(cfg, "physics_file"));
        CELER_VALIDATE(f, << "cannot open physics file");
        json j = json::parse(f);
        import_from_json(j, &p.imported);
        convert_to_native(&p.imported);
    }
    ImportData& imported = p.imported;
    if (cfg.contains("eloss_fluctuation"))
    {
        imported.em_params.energy_loss_fluct = cfg.at("eloss_fluctuation");
    }
    if (cfg.value("disable_msc", false))
    {
        imported.msc_models.clear();
    }

    params.geometry = std::make_shared<GeoParams>(resolve(cfg, "geometry_file"));
    params.material = MaterialParams::from_import(imported);
    params.geomaterial = GeoMaterialParams::from_import(
        imported, params.geometry, params.material);
    params.particle = ParticleParams::from_import(imported);
    params.cutoff = CutoffParams::from_import(
        imported, params.particle, params.material);
    params.wentzel = WentzelOKVIParams::from_import(imported, params.material);

    params.physics = [&] {
        PhysicsParams::Input input;
        input.particles = params.particle;
        input.materials = params.material;
        input.action_registry = params.action_reg.get();
        input.options.fixed_step_limiter = cfg.value("step_limiter", 0.0);
        input.options.secondary_stack_factor
            = cfg.value("secondary_stack_factor", 3.0);
        input.options.linear_loss_limit = imported.em_params.linear_loss_limit;
        input.options.lowest_electron_energy = PhysicsParamsOptions::Energy(
            imported.em_params.lowest_electron_energy);

        ProcessBuilder::Options opts;
        opts.brem_combined = cfg.value("brem_combined", false);
        ProcessBuilder build_process(
            imported, params.particle, params.material, opts);
        for (auto pc :
             ProcessBuilder::get_all_process_classes(imported.processes))
        {
            input.processes.push_back(build_process(pc));
            CELER_ASSERT(input.processes.back());
        }
        return std::make_shared<PhysicsParams>(std::move(input));
    }();

    bool eloss = imported.em_params.energy_loss_fluct;
    auto msc = UrbanMscParams::from_import(
        *params.particle, *params.material, imported);
    p.has_msc = static_cast<bool>(msc);
    p.has_fluct = eloss;
    std::vector<double> field = cfg.value("field", std::vector<double>{0, 0, 0});
    p.has_field = (field[0] != 0 || field[1] != 0 || field[2] != 0);
    p.msc = msc;
    if (eloss)
    {
        p.fluct = std::make_shared<FluctuationParams>(*params.particle,
                                                      *params.material);
    }
    if (cfg.contains("field_map"))
    {
        // Magnetic field map in r-z (field/RZMapField.hh), e.g. the reference's
        // test/celeritas/data/cms-tiny.field.json: AlongStepRZMapFieldMscAction
        CELER_VALIDATE(!p.has_field, << "'field' and 'field_map' are mutually exclusive");
        std::ifstream in(cfg.at("field_map").get<std::string>());
        CELER_VALIDATE(in, << "cannot open field map '"
                           << cfg.at("field_map").get<std::string>() << "'");
        in >> p.rz_field;
        p.has_rz_field = true;
        auto along_step = std::make_shared<AlongStepRZMapFieldMscAction>(
            params.action_reg->next_id(), p.rz_field, p.fluct, msc);
        params.action_reg->insert(along_step);
    }
    else if (!p.has_field)
    {
        auto along_step = std::make_shared<AlongStepGeneralLinearAction>(
            params.action_reg->next_id(), p.fluct, msc);
        params.action_reg->insert(along_step);
    }
    else
    {
        UniformFieldParams field_params;
        for (int i = 0; i < 3; ++i)
        {
            field_params.field[i]
                = native_value_from(units::FieldTesla{field[i]});
        }
        p.field = field_params;
        auto along_step = std::make_shared<AlongStepUniformMscAction>(
            params.action_reg->next_id(), field_params, p.fluct, msc);
        params.action_reg->insert(along_step);
    }
    params.sim = SimParams::from_import(
        imported, params.particle, FieldDriverOptions{}.max_substeps);
}
}  // namespace

//---------------------------------------------------------------------------//
std::unique_ptr<Problem> build_problem(json const& config)
{
    auto p = std::make_unique<Problem>();
    p->config = config;
    json const& cfg = p->config;

    if (cfg.value("problem", std::string("imported")) == "geometry")
    {
        // ORANGE only: used by the navigation ray-trace tests
        p->geo = std::make_shared<OrangeParams>(resolve(cfg, "geometry_file"));
        return p;
    }

    CoreParams::Input params;
    params.action_reg = std::make_shared<ActionRegistry>();
    params.output_reg = std::make_shared<OutputRegistry>();
    params.aux_reg = std::make_shared<AuxParamsRegistry>();

    std::string kind = cfg.value("problem", std::string("imported"));
    if (kind == "simple-compton")
    {
        build_simple_compton(*p, params);
    }
    else
    {
        build_imported(*p, params);
    }

    params.rng = std::make_shared<RngParams>(cfg.value("seed", 20220904u));
    params.max_streams = cfg.value("max_streams", 1u);
    {
        TrackInitParams::Input input;
        input.capacity = cfg.value("initializer_capacity", 4096u);
        input.max_events = cfg.value("max_events", 4096u);
        std::string order = cfg.value("track_order", std::string("none"));
        static std::map<std::string, TrackOrder> const orders{
            {"none", TrackOrder::none},
            {"init_charge", TrackOrder::init_charge},
            {"reindex_shuffle", TrackOrder::reindex_shuffle},
            {"reindex_status", TrackOrder::reindex_status},
            {"reindex_particle_type", TrackOrder::reindex_particle_type},
            {"reindex_along_step_action", TrackOrder::reindex_along_step_action},
            {"reindex_step_limit_action", TrackOrder::reindex_step_limit_action},
            {"reindex_both_action", TrackOrder::reindex_both_action}};
        auto found = orders.find(order);
        CELER_VALIDATE(found != orders.end(), << "unsupported track_order " << order);
        input.track_order = found->second;
        params.init = std::make_shared<TrackInitParams>(std::move(input));
    }
    p->core = std::make_shared<CoreParams>(std::move(params));

    if (cfg.contains("hit_volumes"))
    {
        // Sensitive detectors with full step output (instead of a calorimeter tally: one
        // volume cannot feed two step interfaces, user/detail/StepParams.cc:60-70)
        CELER_VALIDATE(!cfg.contains("simple_calo"),
                       << "'hit_volumes' and 'simple_calo' are mutually exclusive");
        std::vector<VolumeId> vols;
        for (auto const& s : cfg.at("hit_volumes"))
        {
            p->hit_volumes.push_back(s.get<std::string>());
            VolumeId v = p->core->geometry()->volumes().find_unique(s.get<std::string>());
            CELER_VALIDATE(v, << "no volume '" << s.get<std::string>() << "'");
            vols.push_back(v);
        }
        p->hits_nonzero_edep = cfg.value("hits_nonzero_edep", false);
        p->hits = std::make_shared<HitRecorder>(std::move(vols), p->hits_nonzero_edep);
        StepCollector::VecInterface ifaces{p->hits};
        p->collector = std::make_shared<StepCollector>(p->core->geometry(),
                                                       std::move(ifaces),
                                                       p->core->aux_reg().get(),
                                                       p->core->action_reg().get());
    }
    if (cfg.contains("simple_calo"))
    {
        std::vector<Label> labels;
        for (auto const& s : cfg.at("simple_calo"))
        {
            p->calo_volumes.push_back(s.get<std::string>());
            // "name@ext" names a volume by its full label (Label::default_sep)
            labels.push_back(Label::from_separator(s.get<std::string>()));
        }
        p->calo = std::make_shared<SimpleCalo>(
            labels, *p->core->geometry(), p->core->max_streams());
        StepCollector::VecInterface ifaces{p->calo};
        p->collector = std::make_shared<StepCollector>(
            p->core->geometry(),
            std::move(ifaces),
            p->core->aux_reg().get(),
            p->core->action_reg().get());
    }
    // Diagnostics as celer-sim adds them (app/celer-sim/Runner.cc:616-633); registered
    // after the step collector so that the other action ids stay those of the
    // diagnostic-free problem
    if (cfg.value("action_diagnostic", false))
    {
        p->action_diag = ActionDiagnostic::make_and_insert(*p->core);
    }
    if (cfg.value("step_diagnostic_bins", 0) > 0)
    {
        p->step_diag = StepDiagnostic::make_and_insert(
            *p->core, cfg.at("step_diagnostic_bins").get<size_type>());
    }
    return p;
}
//---------------------------------------------------------------------------//
auto HitRecorder::filters() const -> Filters
{
    Filters f;
    for (std::size_t i = 0; i < volumes_.size(); ++i)
        f.detectors[volumes_[i]] = DetectorId{static_cast<DetectorId::size_type>(i)};
    f.nonzero_energy_deposition = nonzero_edep_;
    return f;
}

StepSelection HitRecorder::selection() const
{
    StepSelection sel = StepSelection::all();
    sel.action_id = false;
    for (auto& point : sel.points)
        point.volume_id = false;
    return sel;
}

void HitRecorder::process_steps(HostStepState state)
{
    copy_steps(&last_.at(state.stream_id.get()), state.steps);
}

void HitRecorder::process_steps(DeviceStepState state)
{
    copy_steps(&last_.at(state.stream_id.get()), state.steps);
}

DetectorStepOutput const& HitRecorder::last(unsigned stream) const
{
    return last_.at(stream);
}
}  // namespace celerref