//---------------------------------------------------------------------------//
// Load a problem image into HBM and build the device ParamsView.
//---------------------------------------------------------------------------//
#include "CoreParams.hh"

#include <cmath>
#include <cstring>
#include <map>
#include <sstream>
#include <tuple>

#include "../csrc/orange.cuh"

using namespace b200;

namespace celeritas_b200
{
namespace
{
std::vector<std::string> split_lines(std::string const& s)
{
    std::vector<std::string> out;
    std::istringstream is(s);
    std::string line;
    while (std::getline(is, line))
        out.push_back(line);
    return out;
}

// Node energies of the log-uniform grids, computed once on the host exactly as the
// reference computes them at every lookup: E_i = std::exp(log_front + log_delta * i)
// (corecel/grid/UniformGrid.hh operator[], celeritas/grid/XsCalculator.hh:139-152).
// Identical (front, delta, size) grids share one table.
struct NodeEnergyPool
{
    std::map<std::tuple<uint64_t, uint64_t, uint32_t>, uint32_t> index;
    std::vector<double> values;

    uint32_t get(double front, double delta, uint32_t size)
    {
        uint64_t fb, db;
        std::memcpy(&fb, &front, 8);
        std::memcpy(&db, &delta, 8);
        auto key = std::make_tuple(fb, db, size);
        auto it = index.find(key);
        if (it != index.end())
            return it->second;
        uint32_t offset = values.size();
        for (uint32_t i = 0; i < size; ++i)
        {
            // two roundings (no contraction), as in the reference's host build
            volatile double scaled = delta * i;
            values.push_back(std::exp(front + scaled));
        }
        index.emplace(key, offset);
        return offset;
    }
};
}  // namespace

std::shared_ptr<CoreParams> CoreParams::from_image(std::string const& path)
{
    return from_image(Image::read(path));
}

std::shared_ptr<CoreParams> CoreParams::from_image(Image const& img)
{
    std::shared_ptr<CoreParams> p(new CoreParams);
    p->load(img);
    return p;
}

uint32_t CoreParams::find_particle(int pdg) const
{
    for (size_t i = 0; i < particle_pdg_.size(); ++i)
        if (particle_pdg_[i] == pdg)
            return i;
    return 0xffffffffu;
}

bool CoreParams::has_action(std::string const& label) const
{
    for (ActionRecord const& a : actions_)
        if (a.label == label)
            return true;
    return false;
}

namespace
{
//! The image is input: a truncated, stale or hand-edited file must fail at load, not turn
//! into out-of-bounds reads inside the kernels. Checks the index columns against the columns
//! they point into (INVALID = 0xffffffff is allowed where the format uses it as "none").
void validate_image(Image const& img)
{
    auto fail = [](std::string const& what) {
        throw std::runtime_error("inconsistent problem image: " + what);
    };
    auto u32 = [&](char const* n) { return img.get<uint32_t>(n); };
    auto below = [&](char const* name, std::vector<uint32_t> const& v, size_t limit,
                     bool allow_invalid) {
        for (uint32_t x : v)
            if (!(x < limit || (allow_invalid && x == 0xffffffffu)))
                fail(std::string(name) + " holds " + std::to_string(x) + " (limit "
                     + std::to_string(limit) + ")");
    };
    auto same = [&](char const* a, size_t na, char const* b, size_t nb) {
        if (na != nb)
            fail(std::string(a) + " and " + b + " differ in length (" + std::to_string(na)
                 + ", " + std::to_string(nb) + ")");
    };
    auto ranges = [&](char const* name, std::vector<uint32_t> const& b,
                      std::vector<uint32_t> const& e, size_t limit) {
        same(name, b.size(), name, e.size());
        for (size_t i = 0; i < b.size(); ++i)
            if (b[i] > e[i] || e[i] > limit)
                fail(std::string(name) + " range " + std::to_string(i) + " = ["
                     + std::to_string(b[i]) + ", " + std::to_string(e[i]) + ") exceeds "
                     + std::to_string(limit));
    };

    //// geometry
    {
        size_t const nsurf_ids = u32("geo.local_surface_ids").size();
        size_t const nvol_ids = u32("geo.local_volume_ids").size();
        size_t const nlogic = u32("geo.logic_ints").size();
        size_t const nreals = img.get<double>("geo.reals").size();
        auto utype = img.get<uint8_t>("geo.universe_type");
        same("geo.universe_type", utype.size(), "geo.universe_index", u32("geo.universe_index").size());
        if (u32("geo.universe_surface_offset").size() != utype.size() + 1
            || u32("geo.universe_volume_offset").size() != utype.size() + 1)
            fail("geo.universe_*_offset must have one entry per universe plus one");
        ranges("geo.vol_face", u32("geo.vol_face_begin"), u32("geo.vol_face_end"), nsurf_ids);
        ranges("geo.vol_logic", u32("geo.vol_logic_begin"), u32("geo.vol_logic_end"), nlogic);
        ranges("geo.conn", u32("geo.conn_begin"), u32("geo.conn_end"), nvol_ids);
        below("geo.real_ids", u32("geo.real_ids"), nreals, false);
        same("geo.real_ids", u32("geo.real_ids").size(), "geo.surface_types",
             img.get<uint8_t>("geo.surface_types").size());
        size_t const nvol = u32("geo.vol_flags").size();
        same("geo.vol_flags", nvol, "geo.vol_face_begin", u32("geo.vol_face_begin").size());
        same("geo.vol_flags", nvol, "geo.vol_daughter", u32("geo.vol_daughter").size());
        size_t const ndaughters = u32("geo.daughter_universe").size();
        below("geo.vol_daughter", u32("geo.vol_daughter"), ndaughters, true);
        below("geo.daughter_universe", u32("geo.daughter_universe"), utype.size(), false);
        same("geo.daughter_universe", ndaughters, "geo.daughter_transform",
             u32("geo.daughter_transform").size());
        below("geo.daughter_transform", u32("geo.daughter_transform"),
              u32("geo.transform_offset").size(), false);
        if (u32("geo.simple_units").size() % 16 != 0 || u32("geo.rect_arrays").size() % 16 != 0)
            fail("geo.simple_units / geo.rect_arrays must be rows of 16");
        if (img.get<float>("geo.bih_bboxes").size() % 6 != 0)
            fail("geo.bih_bboxes must be rows of 6");
    }
    if (!img.has("phys.dims"))
        return;

    //// materials, physics tables
    {
        size_t const nelem = u32("mat.element_z").size();
        below("mat.elcomp_element", u32("mat.elcomp_element"), nelem, false);
        ranges("mat.material_elcomp", u32("mat.material_elcomp_begin"),
               u32("mat.material_elcomp_end"), u32("mat.elcomp_element").size());
        size_t const nmat = u32("mat.material_elcomp_begin").size();
        below("geomat.volume_material", u32("geomat.volume_material"), nmat, true);

        auto dims = u32("phys.dims");
        if (dims.size() < 4)
            fail("phys.dims");
        size_t const npart = dims[0], maxproc = dims[1], nmatp = dims[2], nmodels = dims[3];
        if (nmatp != nmat)
            fail("phys.dims and mat.* disagree on the number of materials");
        auto gsize = u32("phys.grid_size");
        auto goff = u32("phys.grid_value_offset");
        size_t const ngrid = gsize.size();
        size_t const nreals = img.get<double>("phys.reals").size();
        same("phys.grid_size", ngrid, "phys.grid_value_offset", goff.size());
        same("phys.grid_size", ngrid, "phys.grid_prime", u32("phys.grid_prime").size());
        same("phys.grid_size", ngrid, "phys.grid_log_front",
             img.get<double>("phys.grid_log_front").size());
        same("phys.grid_size", ngrid, "phys.grid_log_delta",
             img.get<double>("phys.grid_log_delta").size());
        for (size_t g = 0; g < ngrid; ++g)
            if (gsize[g] < 2 || size_t(goff[g]) + gsize[g] > nreals)
                fail("phys grid " + std::to_string(g) + " exceeds phys.reals");
        if (u32("phys.pp_grid").size() != 3 * npart * maxproc * nmat)
            fail("phys.pp_grid must be [3][particle][process][material]");
        below("phys.pp_grid", u32("phys.pp_grid"), ngrid, true);
        below("phys.elsel_grid", u32("phys.elsel_grid"), ngrid, true);
        if (u32("phys.pp_process").size() != npart * maxproc
            || u32("phys.pp_model_begin").size() != npart * maxproc
            || u32("phys.pp_num").size() != npart)
            fail("phys.pp_* must be [particle][process]");
        below("phys.pp_num", u32("phys.pp_num"), maxproc + 1, false);
        below("phys.pmid_model", u32("phys.pmid_model"), nmodels, false);
        below("phys.pm_pmid", u32("phys.pm_pmid"), u32("phys.pmid_model").size(), false);
    }
}
}  // namespace

void CoreParams::load(Image const& img)
{
    validate_image(img);
    auto U32 = [&](char const* n) { return arena_.upload(img.get<uint32_t>(n)); };
    auto F64 = [&](char const* n) { return arena_.upload(img.get<double>(n)); };
    auto F32 = [&](char const* n) { return arena_.upload(img.get<float>(n)); };
    auto U8 = [&](char const* n) { return arena_.upload(img.get<uint8_t>(n)); };

    // Geometry-only images (navigation tests) carry no physics
    bool const geo_only = !img.has("phys.dims");

    //// CORE ////
    if (!geo_only)
    {
        auto a = img.get<uint32_t>("core.actions");
        view_.scalars.boundary_action = a.at(0);
        view_.scalars.propagation_limit_action = a.at(1);
        view_.scalars.tracking_cut_action = a.at(2);
        view_.scalars.along_step_user_action = a.at(3);
        view_.scalars.along_step_neutral_action = a.at(4);
        auto labels = split_lines(img.get_string("core.action_labels"));
        auto order = img.get<uint32_t>("core.action_order");
        for (uint32_t i = 0; i < labels.size(); ++i)
            actions_.push_back({i, labels[i], order.at(i)});
        auto init = img.get<uint32_t>("init.scalars");
        init_capacity_ = init.at(0);
        max_events_ = init.at(1);
        view_.scalars.track_order = init.at(2);
        // reindex_shuffle (CoreTrackData.cc:52-56, detail/TrackSlotUtils.cc:21-32) permutes
        // the reference's thread -> slot map once at construction: which thread works on a
        // slot changes, the slot's results do not. Here threads walk dense lists, so the
        // order is accepted and every per-slot result equals the reference's
        // (tests/test_gpu_track_order.py::test_reindex_shuffle_is_slot_identical).
        if (init.at(2) >= ORDER_SIZE_)
            throw std::runtime_error("unsupported track_order in image");
    }

    //// GEOMETRY ////
    {
        GeoParams& g = view_.geo;
        auto sc = img.get<uint32_t>("geo.scalars");
        g.max_depth = sc.at(0);
        g.max_faces = sc.at(1);
        g.max_intersections = sc.at(2);
        // Volumes beyond ORANGE_MAX_FACES / ORANGE_MAX_ISECT take the big-volume path
        // (csrc/orange.cuh), whose only limit is the size of its sense-word array
        if (g.max_faces > ORANGE_BIG_MAX_FACES)
        {
            throw std::runtime_error(
                "geometry exceeds the face limit of the big-volume path ("
                + std::to_string(ORANGE_BIG_MAX_FACES) + "): max_faces="
                + std::to_string(g.max_faces)
                + " max_intersections=" + std::to_string(g.max_intersections));
        }
        auto tol = img.get<double>("geo.tol");
        g.tol_rel = tol.at(0);
        g.tol_abs = tol.at(1);
        auto utype = img.get<uint8_t>("geo.universe_type");
        g.num_universes = utype.size();
        for (auto t : utype)
            if (t != UNIV_SIMPLE && t != UNIV_RECT_ARRAY)
                throw std::runtime_error("unsupported universe type in image");
        if (utype.empty() || utype[0] != UNIV_SIMPLE)
            throw std::runtime_error("the global universe must be a simple unit");
        g.universe_type = U8("geo.universe_type");
        g.universe_index = U32("geo.universe_index");
        g.universe_surface_offset = U32("geo.universe_surface_offset");
        g.universe_volume_offset = U32("geo.universe_volume_offset");
        static_assert(sizeof(SimpleUnit) == 16 * sizeof(uint32_t), "unit row layout");
        g.simple_units
            = reinterpret_cast<SimpleUnit const*>(U32("geo.simple_units"));
        g.rect_arrays = U32("geo.rect_arrays");
        g.local_surface_ids = U32("geo.local_surface_ids");
        g.local_volume_ids = U32("geo.local_volume_ids");
        g.real_ids = U32("geo.real_ids");
        {
            auto li = img.get<uint32_t>("geo.logic_ints");
            std::vector<uint16_t> l16(li.begin(), li.end());
            g.logic_ints = arena_.upload(l16);
        }
        g.reals = F64("geo.reals");
        g.surface_types = U8("geo.surface_types");
        for (auto t : img.get<uint8_t>("geo.surface_types"))
            if (t >= SURF_INV)
                throw std::runtime_error("involute surfaces are not supported");
        g.vol_face_begin = U32("geo.vol_face_begin");
        g.vol_face_end = U32("geo.vol_face_end");
        g.vol_logic_begin = U32("geo.vol_logic_begin");
        g.vol_logic_end = U32("geo.vol_logic_end");
        g.vol_max_isect = U32("geo.vol_max_isect");
        g.vol_flags = U32("geo.vol_flags");
        g.vol_daughter = U32("geo.vol_daughter");
        g.conn_begin = U32("geo.conn_begin");
        g.conn_end = U32("geo.conn_end");
        g.daughter_universe = U32("geo.daughter_universe");
        g.daughter_transform = U32("geo.daughter_transform");
        g.transform_type = U8("geo.transform_type");
        g.transform_offset = U32("geo.transform_offset");
        g.bih_bboxes = F32("geo.bih_bboxes");
        g.bih_local_volume_ids = U32("geo.bih_local_volume_ids");
        g.bih_inner_parent = U32("geo.bih_inner_parent");
        g.bih_inner_axis = U32("geo.bih_inner_axis");
        g.bih_inner_left_pos = F32("geo.bih_inner_left_pos");
        g.bih_inner_left_child = U32("geo.bih_inner_left_child");
        g.bih_inner_right_pos = F32("geo.bih_inner_right_pos");
        g.bih_inner_right_child = U32("geo.bih_inner_right_child");
        g.bih_leaf_parent = U32("geo.bih_leaf_parent");
        g.bih_leaf_vol_begin = U32("geo.bih_leaf_vol_begin");
        g.bih_leaf_vol_end = U32("geo.bih_leaf_vol_end");
        if (!geo_only)
            g.volume_material = U32("geomat.volume_material");
        volume_labels_ = split_lines(img.get_string("geo.volume_labels"));
    }

    if (geo_only)
        return;

    //// MATERIALS ////
    {
        MatParams& m = view_.mat;
        m.num_elements = img.get<uint32_t>("mat.element_z").size();
        m.num_materials = img.get<uint32_t>("mat.material_state").size();
        m.max_element_components = img.get_scalar<uint32_t>("mat.max_element_components");
        m.element_z = U32("mat.element_z");
        m.element_reals = F64("mat.element_reals");
        m.elcomp_element = U32("mat.elcomp_element");
        m.elcomp_fraction = F64("mat.elcomp_fraction");
        m.material_elcomp_begin = U32("mat.material_elcomp_begin");
        m.material_elcomp_end = U32("mat.material_elcomp_end");
        m.material_reals = F64("mat.material_reals");
    }

    //// PARTICLES / CUTOFFS ////
    {
        ParticleParams& pp = view_.particle;
        pp.num_particles = img.get<double>("particle.mass").size();
        pp.mass = F64("particle.mass");
        pp.charge = F64("particle.charge");
        particle_charge_ = img.get<double>("particle.charge");
        pp.decay_constant = F64("particle.decay_constant");
        pp.matter = U8("particle.matter");
        particle_names_ = split_lines(img.get_string("particle.names"));
        for (auto v : img.get<uint32_t>("particle.pdg"))
            particle_pdg_.push_back(static_cast<int>(v));

        CutoffParams& c = view_.cutoff;
        auto sc = img.get<uint32_t>("cutoff.scalars");
        c.num_particles = sc.at(0);
        c.num_materials = sc.at(1);
        c.apply_post_interaction = sc.at(2);
        c.id_gamma = sc.at(3);
        c.id_electron = sc.at(4);
        c.id_positron = sc.at(5);
        c.energy = F64("cutoff.energy");
        c.range = F64("cutoff.range");
        c.id_to_index = U32("cutoff.id_to_index");
    }

    //// PHYSICS ////
    {
        PhysParams& p = view_.phys;
        auto dims = img.get<uint32_t>("phys.dims");
        p.num_particles = dims.at(0);
        p.max_processes = dims.at(1);
        p.num_materials = dims.at(2);
        p.num_models = dims.at(3);
        auto f = img.get<double>("phys.scalars_f64");
        p.min_range = f.at(0);
        p.max_step_over_range = f.at(1);
        p.min_eprime_over_e = f.at(2);
        p.lowest_electron_energy = f.at(3);
        p.linear_loss_limit = f.at(4);
        p.fixed_step_limiter = f.at(5);
        p.lambda_limit = f.at(6);
        p.range_factor = f.at(7);
        p.safety_factor = f.at(8);
        auto u = img.get<uint32_t>("phys.scalars_u32");
        p.model_to_action = u.at(0);
        p.step_limit_algorithm = u.at(2);
        p.fixed_step_action = u.at(3);
        p.grid_size = U32("phys.grid_size");
        p.grid_log_front = F64("phys.grid_log_front");
        p.grid_log_back = F64("phys.grid_log_back");
        p.grid_log_delta = F64("phys.grid_log_delta");
        p.grid_prime = U32("phys.grid_prime");
        p.grid_value_offset = U32("phys.grid_value_offset");
        {
            // padded to whole 16-byte units: a bulk copy to shared memory moves multiples of 16
            auto reals = img.get<double>("phys.reals");
            view_.phys_reals_count = reals.size();
            if (reals.size() % 2)
                reals.push_back(0);
            p.reals = arena_.upload(reals);
        }
        {
            NodeEnergyPool pool;
            auto size = img.get<uint32_t>("phys.grid_size");
            auto front = img.get<double>("phys.grid_log_front");
            auto delta = img.get<double>("phys.grid_log_delta");
            std::vector<uint32_t> offsets(size.size());
            for (size_t g = 0; g < size.size(); ++g)
                offsets[g] = pool.get(front[g], delta[g], size[g]);
            p.grid_energy_offset = arena_.upload(offsets);
            view_.phys_energy_count = pool.values.size();
            if (pool.values.size() % 2)
                pool.values.push_back(0);
            p.grid_energy = arena_.upload(pool.values);
        }
        p.pp_num = U32("phys.pp_num");
        p.pp_eloss_ppid = U32("phys.pp_eloss_ppid");
        p.pp_has_at_rest = U32("phys.pp_has_at_rest");
        p.pp_process = U32("phys.pp_process");
        p.pp_grid = U32("phys.pp_grid");
        p.pp_integral = U8("phys.pp_integral");
        p.pp_energy_max_xs = F64("phys.pp_energy_max_xs");
        p.pp_model_begin = U32("phys.pp_model_begin");
        p.pp_model_count = U32("phys.pp_model_count");
        p.pm_energy_begin = U32("phys.pm_energy_begin");
        p.pm_energy = F64("phys.pm_energy");
        p.pm_pmid = U32("phys.pm_pmid");
        p.pmid_model = U32("phys.pmid_model");
        p.elsel_begin = U32("phys.elsel_begin");
        p.elsel_count = U32("phys.elsel_count");
        p.elsel_grid = U32("phys.elsel_grid");
        auto hw = img.get<uint32_t>("phys.hardwired");
        p.hw_photoelectric = hw.at(0);
        p.hw_livermore_pe = hw.at(1);
        p.hw_positron_annihilation = hw.at(2);
        p.hw_eplusgg = hw.at(3);
        p.hw_photoelectric_table_thresh
            = img.get_scalar<double>("phys.photoelectric_table_thresh");
    }

    //// MODELS ////
    {
        ModelParams& m = view_.model;
        m.kn.action = INVALID;
        if (img.has("model.kn.ids"))
        {
            auto ids = img.get<uint32_t>("model.kn.ids");
            m.kn.electron = ids.at(0);
            m.kn.gamma = ids.at(1);
            m.kn.inv_electron_mass = img.get_scalar<double>("model.kn.inv_electron_mass");
            m.kn.action = img.get_scalar<uint32_t>("model.kn.action");
        }
        m.mb.action = INVALID;
        if (img.has("model.mb.ids"))
        {
            auto ids = img.get<uint32_t>("model.mb.ids");
            m.mb.action = ids.at(0);
            m.mb.electron = ids.at(1);
            m.mb.positron = ids.at(2);
            m.mb.electron_mass = img.get_scalar<double>("model.mb.electron_mass");
        }
        m.epgg.action = INVALID;
        if (img.has("model.epgg.ids"))
        {
            auto ids = img.get<uint32_t>("model.epgg.ids");
            m.epgg.action = ids.at(0);
            m.epgg.positron = ids.at(1);
            m.epgg.gamma = ids.at(2);
            m.epgg.electron_mass = img.get_scalar<double>("model.epgg.electron_mass");
        }
        m.bh.action = INVALID;
        if (img.has("model.bh.ids"))
        {
            auto ids = img.get<uint32_t>("model.bh.ids");
            m.bh.action = ids.at(0);
            m.bh.electron = ids.at(1);
            m.bh.positron = ids.at(2);
            m.bh.gamma = ids.at(3);
            m.bh.enable_lpm = ids.at(4);
            m.bh.electron_mass = img.get_scalar<double>("model.bh.electron_mass");
        }
        m.sb.action = INVALID;
        if (img.has("model.sb.ids"))
        {
            auto ids = img.get<uint32_t>("model.sb.ids");
            m.sb.action = ids.at(0);
            m.sb.electron = ids.at(1);
            m.sb.positron = ids.at(2);
            m.sb.gamma = ids.at(3);
            m.sb.electron_mass = img.get_scalar<double>("model.sb.electron_mass");
            m.sb.elements = U32("model.sb.elements");
            m.sb.sizes = U32("model.sb.sizes");
            m.sb.reals = F64("model.sb.reals");
        }
        m.rb.action = INVALID;
        if (img.has("model.rb.ids"))
        {
            auto ids = img.get<uint32_t>("model.rb.ids");
            m.rb.action = ids.at(0);
            m.rb.electron = ids.at(1);
            m.rb.positron = ids.at(2);
            m.rb.gamma = ids.at(3);
            m.rb.enable_lpm = ids.at(4);
            m.rb.electron_mass = img.get_scalar<double>("model.rb.electron_mass");
            m.rb.elem_data = F64("model.rb.elem_data");
        }
        m.pe.action = INVALID;
        if (img.has("model.pe.ids"))
        {
            auto ids = img.get<uint32_t>("model.pe.ids");
            m.pe.action = ids.at(0);
            m.pe.electron = ids.at(1);
            m.pe.gamma = ids.at(2);
            m.pe.inv_electron_mass = img.get_scalar<double>("model.pe.inv_electron_mass");
            m.pe.elements = U32("model.pe.elements");
            m.pe.element_thresh = F64("model.pe.element_thresh");
            m.pe.shells = U32("model.pe.shells");
            m.pe.shell_reals = F64("model.pe.shell_reals");
            m.pe.reals = F64("model.pe.reals");
        }
        // Combined bremsstrahlung model (celer-sim `brem_combined`): uses the sb / rb data
        m.cb.action = INVALID;
        m.cb.sb_upper_limit = 1e3;  // em/interactor/detail/PhysicsConstants.hh:62-65
        if (img.has("model.cb.action"))
        {
            if (!img.has("model.sb.ids") || !img.has("model.rb.ids"))
                throw std::runtime_error("combined bremsstrahlung needs the sb and rb data");
            m.cb.action = img.get_scalar<uint32_t>("model.cb.action");
        }
        m.msc.enabled = 0;
        if (img.has("msc.ids"))
        {
            auto ids = img.get<uint32_t>("msc.ids");
            auto pr = img.get<double>("msc.params");
            m.msc.enabled = 1;
            m.msc.electron = ids.at(0);
            m.msc.positron = ids.at(1);
            m.msc.electron_mass = pr.at(0);
            m.msc.tau_small = pr.at(1);
            m.msc.tau_big = pr.at(2);
            m.msc.tau_limit = pr.at(3);
            m.msc.safety_tol = pr.at(4);
            m.msc.geom_limit = pr.at(5);
            m.msc.low_energy_limit = pr.at(6);
            m.msc.high_energy_limit = pr.at(7);
            m.msc.material_data = F64("msc.material_data");
            m.msc.par_mat_data = F64("msc.par_mat_data");
            m.msc.xs_grid_u32 = U32("msc.xs_grid_u32");
            m.msc.xs_grid_f64 = F64("msc.xs_grid_f64");
            m.msc.reals = F64("msc.reals");
            {
                NodeEnergyPool pool;
                auto gu = img.get<uint32_t>("msc.xs_grid_u32");
                auto gf = img.get<double>("msc.xs_grid_f64");
                std::vector<uint32_t> offsets(gu.size() / 3);
                for (size_t e = 0; e < offsets.size(); ++e)
                    offsets[e] = pool.get(gf[3 * e], gf[3 * e + 2], gu[3 * e]);
                m.msc.xs_grid_energy_offset = arena_.upload(offsets);
                m.msc.grid_energy = arena_.upload(pool.values);
            }
        }
        m.rayleigh.action = INVALID;
        if (img.has("model.rayleigh.ids"))
        {
            auto ids = img.get<uint32_t>("model.rayleigh.ids");
            auto consts = img.get<double>("model.rayleigh.consts");
            auto reals = img.get<double>("model.rayleigh.params");
            if (reals.size() != size_t(9) * view_.mat.num_elements)
                throw std::runtime_error(
                    "inconsistent problem image: model.rayleigh.params is not 9 per element");
            m.rayleigh.action = ids.at(0);
            m.rayleigh.gamma = ids.at(1);
            m.rayleigh.hc_factor = consts.at(0);
            m.rayleigh.mev = consts.at(1);
            m.rayleigh.params = arena_.upload(reals);
        }
        m.muioni.bragg_action = m.muioni.icru73qo_action = INVALID;
        m.muioni.bethe_bloch_action = m.muioni.mu_bethe_bloch_action = INVALID;
        if (img.has("model.muioni.actions"))
        {
            auto actions = img.get<uint32_t>("model.muioni.actions");
            auto reals = img.get<double>("model.muioni.reals");
            if (actions.size() != 4 || reals.size() != 3)
                throw std::runtime_error("inconsistent problem image: model.muioni.*");
            m.muioni.bragg_action = actions[0];
            m.muioni.icru73qo_action = actions[1];
            m.muioni.bethe_bloch_action = actions[2];
            m.muioni.mu_bethe_bloch_action = actions[3];
            m.muioni.electron = img.get<uint32_t>("model.muioni.electron").at(0);
            m.muioni.electron_mass = reals[0];
            m.muioni.proton_mass = reals[1];
            m.muioni.alpha_over_twopi = reals[2];
        }
        m.mubrems.action = INVALID;
        if (img.has("model.mubrems.ids"))
        {
            auto ids = img.get<uint32_t>("model.mubrems.ids");
            auto reals = img.get<double>("model.mubrems.reals");
            if (ids.size() != 4 || reals.size() != 3)
                throw std::runtime_error("inconsistent problem image: model.mubrems.*");
            m.mubrems.action = ids[0];
            m.mubrems.gamma = ids[1];
            m.mubrems.mu_minus = ids[2];
            m.mubrems.mu_plus = ids[3];
            m.mubrems.electron_mass = reals[0];
            m.mubrems.sqrt_euler = reals[1];
            m.mubrems.dcs_factor = reals[2];
        }
        m.coulomb.action = INVALID;
        if (img.has("model.coulomb.ids"))
        {
            auto ids = img.get<uint32_t>("model.coulomb.ids");
            auto reals = img.get<double>("model.coulomb.reals");
            auto mott = img.get<double>("model.coulomb.mott");
            auto prefactor = img.get<double>("model.coulomb.nuclear_form_prefactor");
            auto inv_mass = img.get<double>("model.coulomb.inv_mass_cbrt_sq");
            auto el_range = img.get<uint32_t>("mat.element_isocomp_range");
            auto ic_iso = img.get<uint32_t>("mat.isocomp_isotope");
            auto ic_frac = img.get<double>("mat.isocomp_fraction");
            auto iso_za = img.get<uint32_t>("mat.isotope_za");
            auto iso_mass = img.get<double>("mat.isotope_nuclear_mass");
            size_t const ne = view_.mat.num_elements;
            size_t const ni = iso_mass.size();
            bool ok = ids.size() == 5 && reals.size() == 7 && mott.size() == 60 * ne
                      && el_range.size() == 2 * ne && ic_frac.size() == ic_iso.size()
                      && iso_za.size() == 2 * ni && prefactor.size() == ni
                      && (ids[3] == 0 || inv_mass.size() == view_.mat.num_materials)
                      && ids[4] <= 3;
            for (size_t e = 0; ok && e < ne; ++e)
                ok = el_range[2 * e] < el_range[2 * e + 1]
                     && el_range[2 * e + 1] <= ic_iso.size();
            for (size_t i = 0; ok && i < ic_iso.size(); ++i)
                ok = ic_iso[i] < ni;
            if (!ok)
                throw std::runtime_error(
                    "inconsistent problem image: model.coulomb.* / isotope columns");
            m.coulomb.action = ids[0];
            m.coulomb.electron = ids[1];
            m.coulomb.positron = ids[2];
            m.coulomb.is_combined = ids[3];
            m.coulomb.form_factor_type = ids[4];
            m.coulomb.costheta_limit = reals[0];
            m.coulomb.screening_factor = reals[1];
            m.coulomb.a_sq_factor = reals[2];
            m.coulomb.screen_r_sq_elec = reals[3];
            m.coulomb.twopi_mrsq = reals[4];
            m.coulomb.alpha_fine_structure = reals[5];
            m.coulomb.fm_par_hbar = reals[6];
            m.coulomb.nuclear_form_prefactor = arena_.upload(prefactor);
            m.coulomb.mott = arena_.upload(mott);
            m.coulomb.inv_mass_cbrt_sq = arena_.upload(inv_mass);
            m.coulomb.element_isocomp_range = arena_.upload(el_range);
            m.coulomb.isocomp_isotope = arena_.upload(ic_iso);
            m.coulomb.isocomp_fraction = arena_.upload(ic_frac);
            m.coulomb.isotope_za = arena_.upload(iso_za);
            m.coulomb.isotope_nuclear_mass = arena_.upload(iso_mass);
        }
        m.fluct.enabled = 0;
        if (img.has("fluct.urban"))
        {
            m.fluct.enabled = 1;
            m.fluct.electron = img.get_scalar<uint32_t>("fluct.electron");
            m.fluct.electron_mass = img.get_scalar<double>("fluct.electron_mass");
            m.fluct.urban = F64("fluct.urban");
        }
        m.field.enabled = 0;
        m.field.rz_values = nullptr;
        if (img.has("field.uniform") || img.has("field.rz_values"))
        {
            bool const rz = img.has("field.rz_values");
            auto f = rz ? std::vector<double>{0, 0, 0} : img.get<double>("field.uniform");
            if (rz)
            {
                auto grid = img.get<double>("field.rz_grid");
                auto sizes = img.get<uint32_t>("field.rz_sizes");
                auto values = img.get<double>("field.rz_values");
                if (grid.size() != 6 || sizes.size() != 2 || sizes[0] < 2 || sizes[1] < 2
                    || values.size() != size_t(2) * sizes[0] * sizes[1])
                    throw std::runtime_error("inconsistent problem image: field.rz_*");
                for (int i = 0; i < 3; ++i)
                {
                    m.field.rz_z[i] = grid[i];
                    m.field.rz_r[i] = grid[3 + i];
                }
                m.field.rz_size_z = sizes[0];
                m.field.rz_size_r = sizes[1];
                m.field.rz_values = arena_.upload(values);
            }
            auto o = img.get<double>("field.options");
            auto u = img.get<uint32_t>("field.options_u32");
            m.field.enabled = 1;
            for (int i = 0; i < 3; ++i)
                m.field.field[i] = f.at(i);
            m.field.minimum_step = o.at(0);
            m.field.delta_chord = o.at(1);
            m.field.delta_intersection = o.at(2);
            m.field.epsilon_step = o.at(3);
            m.field.epsilon_rel_max = o.at(4);
            m.field.errcon = o.at(5);
            m.field.pgrow = o.at(6);
            m.field.pshrink = o.at(7);
            m.field.safety = o.at(8);
            m.field.max_stepping_increase = o.at(9);
            m.field.max_stepping_decrease = o.at(10);
            m.field.coeffi_per_charge = o.at(11);
            m.field.max_nsteps = u.at(0);
            m.field.max_substeps = u.at(1);
        }
        m.has_extra_models = m.cb.action != INVALID || m.rayleigh.action != INVALID
                             || m.coulomb.action != INVALID
                             || m.mubrems.action != INVALID
                             || m.muioni.bragg_action != INVALID
                             || m.muioni.icru73qo_action != INVALID
                             || m.muioni.bethe_bloch_action != INVALID
                             || m.muioni.mu_bethe_bloch_action != INVALID;
        // Every discrete model action must have an interactor here: an unclaimed one would
        // limit steps through its cross section and then do nothing at the interaction
        for (uint32_t a = view_.phys.model_to_action;
             a < view_.phys.model_to_action + view_.phys.num_models;
             ++a)
        {
            bool const claimed = a == m.kn.action || a == m.mb.action || a == m.epgg.action
                                 || a == m.bh.action || a == m.sb.action || a == m.rb.action
                                 || a == m.pe.action || a == m.cb.action
                                 || a == m.rayleigh.action || a == m.coulomb.action
                                 || a == m.muioni.bragg_action || a == m.muioni.icru73qo_action
                                 || a == m.muioni.bethe_bloch_action
                                 || a == m.muioni.mu_bethe_bloch_action
                                 || a == m.mubrems.action;
            if (!claimed)
                throw std::runtime_error(
                    "no B200 interactor for model action '"
                    + (a < actions_.size() ? actions_[a].label : std::to_string(a)) + "'");
        }
        auto c = img.get<double>("constants");
        m.constants.migdal_constant = c.at(0);
        m.constants.lpm_constant = c.at(1);
        m.constants.r_electron = c.at(2);
        m.constants.alpha_fine_structure = c.at(3);
    }

    //// RNG / SIM ////
    {
        auto r = img.get<uint32_t>("rng.params");
        if (r.size() < 1 + 160 + 160)
            throw std::runtime_error("rng.params: expected seed + 2 x [32][5] jump polynomials");
        view_.rng.seed = r.at(0);
        std::vector<uint32_t> jump(r.begin() + 1, r.begin() + 1 + 160);
        std::vector<uint32_t> jump_sub(r.begin() + 161, r.begin() + 161 + 160);
        view_.rng.jump = arena_.upload(jump);
        view_.rng.jump_subsequence = arena_.upload(jump_sub);

        auto ls = img.get<uint32_t>("sim.looping_steps");
        view_.sim.has_looping = !ls.empty();
        view_.sim.looping_steps = U32("sim.looping_steps");
        view_.sim.looping_energy = F64("sim.looping_energy");
    }

    //// DETECTORS ////
    {
        detector_volumes_ = split_lines(img.get_string("calo.volumes"));
        if (!detector_volumes_.empty())
        {
            std::vector<uint32_t> det(volume_labels_.size(), INVALID);
            for (uint32_t d = 0; d < detector_volumes_.size(); ++d)
            {
                bool found = false;
                for (uint32_t v = 0; v < volume_labels_.size(); ++v)
                {
                    if (volume_labels_[v] == detector_volumes_[d])
                    {
                        det[v] = d;
                        found = true;
                    }
                }
                if (!found)
                    throw std::runtime_error("detector volume '" + detector_volumes_[d]
                                             + "' not found in geometry");
            }
            d_detector_of_volume_ = arena_.upload(det);
        }
    }
    //// STEP / HIT OUTPUT ////
    if (img.has("hits.volumes"))
    {
        hit_volumes_ = split_lines(img.get_string("hits.volumes"));
        hits_nonzero_edep_ = img.get_scalar<uint32_t>("hits.nonzero_edep") != 0;
        std::vector<uint32_t> det(volume_labels_.size(), INVALID);
        for (uint32_t d = 0; d < hit_volumes_.size(); ++d)
        {
            uint32_t found = 0;
            for (uint32_t v = 0; v < volume_labels_.size(); ++v)
            {
                if (volume_labels_[v] == hit_volumes_[d])
                {
                    det[v] = d;
                    ++found;
                }
            }
            if (found != 1)
                throw std::runtime_error("sensitive volume '" + hit_volumes_[d]
                                         + "' is not a unique volume of the geometry");
        }
        if (!hit_volumes_.empty())
            d_hit_detector_of_volume_ = arena_.upload(det);
    }
}

namespace
{
void require_unfrozen(bool frozen, char const* what)
{
    // CoreState sizes track_counters, the initializer queue and ti_neutral_prefix from these
    // at construction: changing them under a live state would overrun its arrays
    if (frozen)
        throw std::runtime_error(std::string(what)
                                 + " cannot change after a CoreState has been created");
}
}  // namespace

void CoreParams::init_capacity(uint32_t capacity)
{
    if (capacity == 0)
        throw std::runtime_error("nonpositive initializer_capacity=0");
    if (capacity != init_capacity_)
        require_unfrozen(frozen_, "initializer_capacity");
    init_capacity_ = capacity;
}

void CoreParams::max_events(uint32_t num_events)
{
    if (num_events == 0)
        throw std::runtime_error("max_events must be positive");
    if (num_events != max_events_)
        require_unfrozen(frozen_, "max_events");
    max_events_ = num_events;
}

void CoreParams::track_order(uint32_t order)
{
    if (order >= ORDER_SIZE_)
        throw std::runtime_error("unsupported track_order");
    // The reindex orders come with SortTracksAction entries in the action table
    // (CoreParams.cc:253-284 of the reference): they are a property of the exported image
    if (order != view_.scalars.track_order
        && (order >= ORDER_REINDEX_STATUS || view_.scalars.track_order >= ORDER_REINDEX_STATUS))
    {
        throw std::runtime_error(
            "track_order differs from the order the problem image was exported with, and one "
            "of them sorts tracks (its sort actions are part of the image's action table)");
    }
    if (order != view_.scalars.track_order)
        require_unfrozen(frozen_, "track_order");
    view_.scalars.track_order = order;
}

void CoreParams::uniform_field_tesla(double const (&field)[3])
{
    if (!this->has_uniform_field() || view_.model.field.rz_values)
        throw std::runtime_error(
            "the problem image was exported without a uniform-field along-step action");
    // native field unit is gauss (reference: units::FieldTesla -> native, Runner.cc:394-398)
    for (int i = 0; i < 3; ++i)
        view_.model.field.field[i] = field[i] * 1e4;
}
}  // namespace celeritas_b200
