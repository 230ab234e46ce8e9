//---------------------------------------------------------------------------//
// Reference-side adapter: flattens the reference's host-side CoreParams
// (/root/reference/src/celeritas/global/CoreTrackData.hh:65-106) into the
// column arrays of a B200 problem image. This is the code a maintainer of the
// reference would add to hand a constructed problem across the C-ABI; here it
// also produces the committed image fixtures under data/images/.
//---------------------------------------------------------------------------//
#include <cmath>
#include <cstdint>
#include <limits>
#include <string>
#include <vector>

#include "corecel/data/Collection.hh"
#include "corecel/sys/ActionRegistry.hh"
#include "orange/OrangeData.hh"
#include "orange/OrangeParams.hh"
#include "celeritas/Quantities.hh"
#include "celeritas/em/interactor/detail/PhysicsConstants.hh"
#include "celeritas/em/model/BetheHeitlerModel.hh"
#include "celeritas/em/model/EPlusGGModel.hh"
#include "celeritas/em/model/KleinNishinaModel.hh"
#include "celeritas/em/model/LivermorePEModel.hh"
#include "celeritas/em/model/MollerBhabhaModel.hh"
#include "celeritas/em/model/RayleighModel.hh"
#include "celeritas/em/model/BetheBlochModel.hh"
#include "celeritas/em/model/BraggModel.hh"
#include "celeritas/em/model/CombinedBremModel.hh"
#include "celeritas/em/model/ICRU73QOModel.hh"
#include "celeritas/em/model/MuBetheBlochModel.hh"
#include "celeritas/em/model/MuBremsstrahlungModel.hh"
#include "celeritas/em/model/CoulombScatteringModel.hh"
#include "celeritas/em/params/WentzelOKVIParams.hh"
#include "celeritas/em/xs/NuclearFormFactors.hh"
#include "celeritas/em/model/RelativisticBremModel.hh"
#include "celeritas/em/model/SeltzerBergerModel.hh"
#include "celeritas/em/params/FluctuationParams.hh"
#include "celeritas/em/params/UrbanMscParams.hh"
#include "celeritas/geo/GeoMaterialParams.hh"
#include "celeritas/geo/GeoParams.hh"
#include "celeritas/global/alongstep/AlongStepGeneralLinearAction.hh"
#include "celeritas/global/alongstep/AlongStepUniformMscAction.hh"
#include "celeritas/mat/MaterialParams.hh"
#include "celeritas/phys/CutoffParams.hh"
#include "celeritas/phys/ParticleParams.hh"
#include "celeritas/phys/PhysicsParams.hh"
#include "celeritas/random/RngParams.hh"
#include "celeritas/track/SimParams.hh"
#include "celeritas/track/TrackInitParams.hh"

#include "../../celeritas_b200/host/Image.hh"
#include "corecel/grid/UniformGridData.hh"

#include "Problem.hh"

using namespace celeritas;

namespace celerref
{
namespace
{
using U32 = std::vector<uint32_t>;
using F64 = std::vector<double>;
using F32 = std::vector<float>;
using U8 = std::vector<uint8_t>;
constexpr uint32_t invalid = 0xffffffffu;

template<class T, Ownership W, MemSpace M, class I>
Span<T const> all(Collection<T, W, M, I> const& c)
{
    if (c.empty())
        return {};
    return c[AllItems<T, M>{}];
}

template<class Id>
uint32_t raw(Id id)
{
    return id ? id.unchecked_get() : invalid;
}

//---------------------------------------------------------------------------//
void export_geometry(HostCRef<OrangeParamsData> const& g, b200::Image& img)
{
    img.put("geo.scalars",
            U32{g.scalars.max_depth,
                g.scalars.max_faces,
                g.scalars.max_intersections,
                g.scalars.max_logic_depth});
    img.put("geo.tol", F64{g.scalars.tol.rel, g.scalars.tol.abs});

    // Universes
    {
        U8 types;
        U32 indices;
        for (auto t : all(g.universe_types))
            types.push_back(static_cast<uint8_t>(t));
        for (auto i : all(g.universe_indices))
            indices.push_back(i);
        img.put("geo.universe_type", types);
        img.put("geo.universe_index", indices);
        U32 so, vo;
        for (auto v : all(g.universe_indexer_data.surfaces))
            so.push_back(v);
        for (auto v : all(g.universe_indexer_data.volumes))
            vo.push_back(v);
        img.put("geo.universe_surface_offset", so);
        img.put("geo.universe_volume_offset", vo);
    }

    // Flat pools (copied verbatim; records index into them)
    {
        U32 lsi, lvi, rid, logic;
        for (auto v : all(g.local_surface_ids))
            lsi.push_back(raw(v));
        for (auto v : all(g.local_volume_ids))
            lvi.push_back(raw(v));
        for (auto v : all(g.real_ids))
            rid.push_back(raw(v));
        for (auto v : all(g.logic_ints))
            logic.push_back(v);
        img.put("geo.local_surface_ids", lsi);
        img.put("geo.local_volume_ids", lvi);
        img.put("geo.real_ids", rid);
        img.put("geo.logic_ints", logic);
        F64 reals(all(g.reals).begin(), all(g.reals).end());
        img.put("geo.reals", reals);
        U8 st;
        for (auto v : all(g.surface_types))
            st.push_back(static_cast<uint8_t>(v));
        img.put("geo.surface_types", st);
    }

    // Volume records
    {
        U32 fb, fe, lb, le, mi, fl, da;
        for (auto const& v : all(g.volume_records))
        {
            fb.push_back(v.faces.begin()->unchecked_get());
            fe.push_back(v.faces.end()->unchecked_get());
            lb.push_back(v.logic.begin()->unchecked_get());
            le.push_back(v.logic.end()->unchecked_get());
            mi.push_back(v.max_intersections);
            fl.push_back(v.flags);
            da.push_back(raw(v.daughter_id));
        }
        img.put("geo.vol_face_begin", fb);
        img.put("geo.vol_face_end", fe);
        img.put("geo.vol_logic_begin", lb);
        img.put("geo.vol_logic_end", le);
        img.put("geo.vol_max_isect", mi);
        img.put("geo.vol_flags", fl);
        img.put("geo.vol_daughter", da);
    }
    // Connectivity records
    {
        U32 nb, ne;
        for (auto const& c : all(g.connectivity_records))
        {
            nb.push_back(c.neighbors.begin()->unchecked_get());
            ne.push_back(c.neighbors.end()->unchecked_get());
        }
        img.put("geo.conn_begin", nb);
        img.put("geo.conn_end", ne);
    }
    // Daughters and transforms
    {
        U32 du, dt;
        for (auto const& d : all(g.daughters))
        {
            du.push_back(raw(d.universe_id));
            dt.push_back(raw(d.transform_id));
        }
        img.put("geo.daughter_universe", du);
        img.put("geo.daughter_transform", dt);
        U8 tt;
        U32 to;
        for (auto const& t : all(g.transforms))
        {
            tt.push_back(static_cast<uint8_t>(t.type));
            to.push_back(raw(t.data_offset));
        }
        img.put("geo.transform_type", tt);
        img.put("geo.transform_offset", to);
    }
    // BIH storage
    {
        F32 bb;
        for (auto const& b : all(g.bih_tree_data.bboxes))
        {
            for (int i = 0; i < 3; ++i)
                bb.push_back(b.lower()[i]);
            for (int i = 0; i < 3; ++i)
                bb.push_back(b.upper()[i]);
        }
        img.put("geo.bih_bboxes", bb);
        U32 lv;
        for (auto v : all(g.bih_tree_data.local_volume_ids))
            lv.push_back(raw(v));
        img.put("geo.bih_local_volume_ids", lv);
        U32 ip, ia, ilc, irc;
        F32 ilp, irp;
        using Edge = celeritas::detail::BIHInnerNode::Edge;
        for (auto const& n : all(g.bih_tree_data.inner_nodes))
        {
            ip.push_back(raw(n.parent));
            ia.push_back(static_cast<uint32_t>(n.axis));
            ilp.push_back(n.bounding_planes[Edge::left].position);
            ilc.push_back(raw(n.bounding_planes[Edge::left].child));
            irp.push_back(n.bounding_planes[Edge::right].position);
            irc.push_back(raw(n.bounding_planes[Edge::right].child));
        }
        img.put("geo.bih_inner_parent", ip);
        img.put("geo.bih_inner_axis", ia);
        img.put("geo.bih_inner_left_pos", ilp);
        img.put("geo.bih_inner_left_child", ilc);
        img.put("geo.bih_inner_right_pos", irp);
        img.put("geo.bih_inner_right_child", irc);
        U32 lp, lb, le;
        for (auto const& n : all(g.bih_tree_data.leaf_nodes))
        {
            lp.push_back(raw(n.parent));
            lb.push_back(n.vol_ids.begin()->unchecked_get());
            le.push_back(n.vol_ids.end()->unchecked_get());
        }
        img.put("geo.bih_leaf_parent", lp);
        img.put("geo.bih_leaf_vol_begin", lb);
        img.put("geo.bih_leaf_vol_end", le);
    }
    // Simple unit records: one row per unit
    {
        U32 rows;
        for (auto const& u : all(g.simple_units))
        {
            rows.push_back(u.surfaces.types.begin()->unchecked_get());
            rows.push_back(u.surfaces.types.end()->unchecked_get());
            rows.push_back(u.surfaces.data_offsets.begin()->unchecked_get());
            rows.push_back(u.connectivity.begin()->unchecked_get());
            // volumes: ItemMap over a contiguous range of VolumeRecordId
            rows.push_back(u.volumes.size() ? raw(u.volumes[LocalVolumeId{0}])
                                            : 0);
            rows.push_back(u.volumes.size());
            rows.push_back(raw(u.background));
            rows.push_back(u.simple_safety ? 1 : 0);
            auto const& t = u.bih_tree;
            rows.push_back(t.bboxes.size() ? raw(t.bboxes[LocalVolumeId{0}])
                                           : 0);
            rows.push_back(t.inner_nodes.begin()->unchecked_get());
            rows.push_back(t.inner_nodes.size());
            rows.push_back(t.leaf_nodes.begin()->unchecked_get());
            rows.push_back(t.leaf_nodes.size());
            rows.push_back(t.inf_volids.begin()->unchecked_get());
            rows.push_back(t.inf_volids.size());
            rows.push_back(0);
        }
        img.put("geo.simple_units", rows);  // 16 u32 per unit
    }
    // Rect arrays
    {
        U32 rows;
        for (auto const& r : all(g.rect_arrays))
        {
            rows.push_back(r.daughters.size()
                               ? raw(r.daughters[LocalVolumeId{0}])
                               : 0);
            rows.push_back(r.daughters.size());
            for (int ax = 0; ax < 3; ++ax)
                rows.push_back(r.dims[ax]);
            for (int ax = 0; ax < 3; ++ax)
            {
                rows.push_back(r.grid[ax].begin()->unchecked_get());
                rows.push_back(r.grid[ax].end()->unchecked_get());
            }
            for (int i = 0; i < 4; ++i)
                rows.push_back(r.surface_indexer_data.offsets[i]);
            rows.push_back(0);
        }
        img.put("geo.rect_arrays", rows);  // 16 u32 per array
    }
}

//---------------------------------------------------------------------------//
void export_materials(HostCRef<MaterialParamsData> const& m, b200::Image& img)
{
    U32 el_z;
    F64 el;  // 6 per element
    for (auto const& e : all(m.elements))
    {
        el_z.push_back(e.atomic_number.unchecked_get());
        el.push_back(e.atomic_mass.value());
        el.push_back(e.cbrt_z);
        el.push_back(e.cbrt_zzp);
        el.push_back(e.log_z);
        el.push_back(e.coulomb_correction);
        el.push_back(e.mass_radiation_coeff);
    }
    img.put("mat.element_z", el_z);
    img.put("mat.element_reals", el);
    U32 ce;
    F64 cf;
    for (auto const& c : all(m.elcomponents))
    {
        ce.push_back(raw(c.element));
        cf.push_back(c.fraction);
    }
    img.put("mat.elcomp_element", ce);
    img.put("mat.elcomp_fraction", cf);
    // Isotopes (mat/MaterialData.hh:30-60): per element a range of (isotope, fraction)
    // components; per isotope Z, A and the nuclear mass
    U32 el_iso, ic_iso, iso_za;
    F64 ic_frac, iso_mass;
    for (auto const& e : all(m.elements))
    {
        el_iso.push_back(e.isotopes.begin()->unchecked_get());
        el_iso.push_back(e.isotopes.end()->unchecked_get());
    }
    for (auto const& c : all(m.isocomponents))
    {
        ic_iso.push_back(raw(c.isotope));
        ic_frac.push_back(c.fraction);
    }
    for (auto const& i : all(m.isotopes))
    {
        iso_za.push_back(i.atomic_number.unchecked_get());
        iso_za.push_back(i.atomic_mass_number.unchecked_get());
        iso_mass.push_back(i.nuclear_mass.value());
    }
    img.put("mat.element_isocomp_range", el_iso);
    img.put("mat.isocomp_isotope", ic_iso);
    img.put("mat.isocomp_fraction", ic_frac);
    img.put("mat.isotope_za", iso_za);
    img.put("mat.isotope_nuclear_mass", iso_mass);
    U32 mb, me, ms;
    F64 mr;  // 8 per material
    for (auto const& r : all(m.materials))
    {
        mb.push_back(r.elements.begin()->unchecked_get());
        me.push_back(r.elements.end()->unchecked_get());
        ms.push_back(static_cast<uint32_t>(r.matter_state));
        mr.push_back(r.number_density);
        mr.push_back(r.temperature);
        mr.push_back(r.zeff);
        mr.push_back(r.density);
        mr.push_back(r.electron_density);
        mr.push_back(r.rad_length);
        mr.push_back(r.mean_exc_energy.value());
        mr.push_back(r.log_mean_exc_energy.value());
    }
    img.put("mat.material_elcomp_begin", mb);
    img.put("mat.material_elcomp_end", me);
    img.put("mat.material_state", ms);
    img.put("mat.material_reals", mr);
    img.put_scalar<uint32_t>("mat.max_element_components",
                             m.max_element_components);
}

//---------------------------------------------------------------------------//
void export_physics(HostCRef<PhysicsParamsData> const& p,
                    uint32_t num_materials,
                    b200::Image& img)
{
    auto const& sc = p.scalars;
    uint32_t const np = p.process_groups.size();
    uint32_t const P = sc.max_particle_processes;
    img.put("phys.dims", U32{np, P, num_materials, sc.num_models});
    img.put("phys.scalars_f64",
            F64{sc.min_range,
                sc.max_step_over_range,
                sc.min_eprime_over_e,
                sc.lowest_electron_energy.value(),
                sc.linear_loss_limit,
                sc.fixed_step_limiter,
                sc.lambda_limit,
                sc.range_factor,
                sc.safety_factor,
                sc.secondary_stack_factor});
    img.put("phys.scalars_u32",
            U32{sc.model_to_action,
                sc.num_models,
                static_cast<uint32_t>(sc.step_limit_algorithm),
                raw(sc.fixed_step_action)});

    // Grids
    {
        U32 gsz, gpr, gvo;
        F64 gfr, gde, gba;
        for (auto const& g : all(p.value_grids))
        {
            gsz.push_back(g.log_energy.size);
            gfr.push_back(g.log_energy.front);
            gba.push_back(g.log_energy.back);
            gde.push_back(g.log_energy.delta);
            gpr.push_back(g.prime_index);
            gvo.push_back(g.value.begin()->unchecked_get());
        }
        img.put("phys.grid_size", gsz);
        img.put("phys.grid_log_front", gfr);
        img.put("phys.grid_log_back", gba);
        img.put("phys.grid_log_delta", gde);
        img.put("phys.grid_prime", gpr);
        img.put("phys.grid_value_offset", gvo);
        F64 reals(all(p.reals).begin(), all(p.reals).end());
        img.put("phys.reals", reals);
    }

    auto grid_of_table = [&](ValueTable const& table, uint32_t idx) -> uint32_t {
        if (!table || idx >= table.grids.size())
            return invalid;
        auto ref = table.grids[idx];
        if (!ref)
            return invalid;
        return raw(p.value_grid_ids[ref]);
    };

    U32 pp_num(np), pp_eloss(np), pp_at_rest(np);
    U32 pp_process(np * P, invalid);
    // [vgt][particle][ppid][material]
    U32 pp_grid(3 * np * P * num_materials, invalid);
    U8 pp_integral(np * P, 0);
    F64 pp_emax(np * P * num_materials, 0.0);
    U32 pp_model_begin(np * P, 0), pp_model_count(np * P, 0);
    // per-ppid flattened model bounds
    F64 pm_energy;  // (count+1) bounds per ppid, concatenated
    U32 pm_energy_begin(np * P, 0);
    U32 pm_pmid;  // particle-model ids, concatenated per ppid

    for (uint32_t ip = 0; ip < np; ++ip)
    {
        ProcessGroup const& pg = p.process_groups[ParticleId{ip}];
        pp_num[ip] = pg.size();
        pp_eloss[ip] = raw(pg.eloss_ppid);
        pp_at_rest[ip] = pg.has_at_rest;
        for (uint32_t pp = 0; pp < pg.size(); ++pp)
        {
            uint32_t row = ip * P + pp;
            pp_process[row] = raw(p.process_ids[pg.processes[pp]]);
            for (int vgt = 0; vgt < 3; ++vgt)
            {
                ValueTable const& table
                    = p.value_tables[pg.tables[ValueGridType(vgt)][pp]];
                for (uint32_t m = 0; m < num_materials; ++m)
                {
                    pp_grid[((vgt * np + ip) * P + pp) * num_materials + m]
                        = grid_of_table(table, m);
                }
            }
            IntegralXsProcess const& ixs = p.integral_xs[pg.integral_xs[pp]];
            if (ixs)
            {
                pp_integral[row] = 1;
                for (uint32_t m = 0; m < num_materials; ++m)
                {
                    pp_emax[row * num_materials + m]
                        = p.reals[ixs.energy_max_xs[m]];
                }
            }
            ModelGroup const& mg = p.model_groups[pg.models[pp]];
            pp_model_begin[row] = pm_pmid.size();
            pp_model_count[row] = mg.model.size();
            pm_energy_begin[row] = pm_energy.size();
            for (auto e : p.reals[mg.energy])
                pm_energy.push_back(e);
            for (auto pmid : p.pmodel_ids[mg.model])
                pm_pmid.push_back(raw(pmid));
        }
    }
    img.put("phys.pp_num", pp_num);
    img.put("phys.pp_eloss_ppid", pp_eloss);
    img.put("phys.pp_has_at_rest", pp_at_rest);
    img.put("phys.pp_process", pp_process);
    img.put("phys.pp_grid", pp_grid);
    img.put("phys.pp_integral", pp_integral);
    img.put("phys.pp_energy_max_xs", pp_emax);
    img.put("phys.pp_model_begin", pp_model_begin);
    img.put("phys.pp_model_count", pp_model_count);
    img.put("phys.pm_energy", pm_energy);
    img.put("phys.pm_energy_begin", pm_energy_begin);
    img.put("phys.pm_pmid", pm_pmid);

    // Particle-model -> model id, and per-(pmid, material) element xs grids
    {
        uint32_t npm = p.model_ids.size();
        U32 model_id(npm);
        U32 elsel_begin(npm * num_materials, invalid);
        U32 elsel_count(npm * num_materials, 0);
        U32 elsel_grid;
        for (uint32_t i = 0; i < npm; ++i)
        {
            ParticleModelId pmid{i};
            model_id[i] = raw(p.model_ids[pmid]);
            ModelXsTable const& mx = p.model_xs[pmid];
            if (!mx)
                continue;
            for (uint32_t m = 0; m < num_materials; ++m)
            {
                auto ref = mx.material[m];
                if (!ref)
                    continue;
                ValueTableId tid = p.value_table_ids[ref];
                if (!tid)
                    continue;
                ValueTable const& table = p.value_tables[tid];
                if (!table)
                    continue;
                elsel_begin[i * num_materials + m] = elsel_grid.size();
                elsel_count[i * num_materials + m] = table.grids.size();
                for (uint32_t e = 0; e < table.grids.size(); ++e)
                    elsel_grid.push_back(grid_of_table(table, e));
            }
        }
        img.put("phys.pmid_model", model_id);
        img.put("phys.elsel_begin", elsel_begin);
        img.put("phys.elsel_count", elsel_count);
        img.put("phys.elsel_grid", elsel_grid);
    }

    // Hardwired models
    {
        auto const& h = p.hardwired;
        // Atomic relaxation emits more secondaries per interaction than the fixed per-slot
        // secondary storage of the B200 state holds
        CELER_VALIDATE(!h.relaxation_data,
                       << "atomic relaxation has no B200 export");
        img.put("phys.hardwired",
                U32{raw(h.photoelectric),
                    raw(h.livermore_pe),
                    raw(h.positron_annihilation),
                    raw(h.eplusgg)});
        img.put_scalar<double>("phys.photoelectric_table_thresh",
                               h.photoelectric ? h.photoelectric_table_thresh.value()
                                               : 0.0);
    }
}

//---------------------------------------------------------------------------//
//! Volume labels of the image: the name; volumes of one universe that share a name keep
//! their extension ("box@1" .. "box@4", corecel/io/Label.cc:59-71) so that a detector can
//! be attached to one of them
template<class GeoParamsT>
std::string image_volume_labels(GeoParamsT const& geo)
{
    auto const& offsets = geo.host_ref().universe_indexer_data.volumes;
    std::string labels;
    for (size_t u = 0; u + 1 < offsets.size(); ++u)
    {
        size_t const begin = offsets[AllItems<size_type>{}][u];
        size_t const end = offsets[AllItems<size_type>{}][u + 1];
        for (size_t v = begin; v < end; ++v)
        {
            Label const& lab = geo.volumes().at(VolumeId(v));
            size_t same = 0;
            for (size_t w = begin; w < end; ++w)
                same += geo.volumes().at(VolumeId(w)).name == lab.name;
            labels += (same > 1 && !lab.ext.empty())
                          ? lab.name + Label::default_sep + lab.ext
                          : lab.name;
            labels += "\n";
        }
    }
    return labels;
}

void export_models(Problem const& prob, b200::Image& img)
{
    PhysicsParams const& phys = *prob.core->physics();
    U32 model_kind(phys.num_models(), 0);
    U32 muioni_actions(4, 0xffffffffu);  // Bragg, ICRU73QO, Bethe-Bloch, mu Bethe-Bloch
    std::string labels;
    auto put_sb = [&img](uint32_t action,
                         uint32_t electron,
                         uint32_t positron,
                         uint32_t gamma,
                         double electron_mass,
                         auto const& t) {
        img.put("model.sb.ids", U32{action, electron, positron, gamma});
        img.put_scalar<double>("model.sb.electron_mass", electron_mass);
        U32 rows;
        for (auto const& el : all(t.elements))
        {
            rows.push_back(el.grid.x.begin()->unchecked_get());
            rows.push_back(el.grid.x.size());
            rows.push_back(el.grid.y.begin()->unchecked_get());
            rows.push_back(el.grid.y.size());
            rows.push_back(el.grid.values.begin()->unchecked_get());
            rows.push_back(el.argmax.begin()->unchecked_get());
            rows.push_back(0);
            rows.push_back(0);
        }
        img.put("model.sb.elements", rows);
        U32 sizes(all(t.sizes).begin(), all(t.sizes).end());
        img.put("model.sb.sizes", sizes);
        F64 reals(all(t.reals).begin(), all(t.reals).end());
        img.put("model.sb.reals", reals);
    };
    auto put_rb = [&img](uint32_t action, auto const& d) {
        img.put("model.rb.ids",
                U32{action,
                    raw(d.ids.electron),
                    raw(d.ids.positron),
                    raw(d.ids.gamma),
                    d.enable_lpm ? 1u : 0u});
        img.put_scalar<double>("model.rb.electron_mass", d.electron_mass.value());
        F64 ed;
        for (auto const& e : all(d.elem_data))
        {
            ed.push_back(e.fz);
            ed.push_back(e.factor1);
            ed.push_back(e.factor2);
            ed.push_back(e.gamma_factor);
            ed.push_back(e.epsilon_factor);
        }
        img.put("model.rb.elem_data", ed);
    };
    for (auto mid : range(ModelId{phys.num_models()}))
    {
        auto const& model = *phys.model(mid);
        labels += std::string(model.label()) + "\n";
        if (auto* kn = dynamic_cast<KleinNishinaModel const*>(&model))
        {
            auto const& d = kn->host_ref();
            img.put("model.kn.ids", U32{raw(d.ids.electron), raw(d.ids.gamma)});
            img.put_scalar<double>("model.kn.inv_electron_mass",
                                   d.inv_electron_mass);
            img.put_scalar<uint32_t>("model.kn.action",
                                     kn->action_id().unchecked_get());
        }
        else if (auto* mb = dynamic_cast<MollerBhabhaModel const*>(&model))
        {
            auto const& d = mb->host_ref();
            img.put("model.mb.ids",
                    U32{mb->action_id().unchecked_get(),
                        raw(d.ids.electron),
                        raw(d.ids.positron)});
            img.put_scalar<double>("model.mb.electron_mass",
                                   d.electron_mass.value());
        }
        else if (auto* ep = dynamic_cast<EPlusGGModel const*>(&model))
        {
            auto const& d = ep->host_ref();
            img.put("model.epgg.ids",
                    U32{ep->action_id().unchecked_get(),
                        raw(d.positron),
                        raw(d.gamma)});
            img.put_scalar<double>("model.epgg.electron_mass",
                                   d.electron_mass.value());
        }
        else if (auto* bh = dynamic_cast<BetheHeitlerModel const*>(&model))
        {
            auto const& d = bh->host_ref();
            img.put("model.bh.ids",
                    U32{bh->action_id().unchecked_get(),
                        raw(d.ids.electron),
                        raw(d.ids.positron),
                        raw(d.ids.gamma),
                        d.enable_lpm ? 1u : 0u});
            img.put_scalar<double>("model.bh.electron_mass",
                                   d.electron_mass.value());
        }
        else if (auto* sb = dynamic_cast<SeltzerBergerModel const*>(&model))
        {
            auto const& d = sb->host_ref();
            put_sb(sb->action_id().unchecked_get(),
                   raw(d.ids.electron),
                   raw(d.ids.positron),
                   raw(d.ids.gamma),
                   d.electron_mass.value(),
                   d.differential_xs);
        }
        else if (auto* rb = dynamic_cast<RelativisticBremModel const*>(&model))
        {
            put_rb(rb->action_id().unchecked_get(), rb->host_ref());
        }
        else if (auto* cb = dynamic_cast<CombinedBremModel const*>(&model))
        {
            // Seltzer-Berger below 1 GeV, relativistic above, behind ONE action
            // (em/interactor/CombinedBremInteractor.hh:132-170): both data sets are
            // exported with no action of their own, plus the combined action id
            auto const& d = cb->host_ref();
            put_sb(0xffffffffu,
                   raw(d.rb_data.ids.electron),
                   raw(d.rb_data.ids.positron),
                   raw(d.rb_data.ids.gamma),
                   d.rb_data.electron_mass.value(),
                   d.sb_differential_xs);
            put_rb(0xffffffffu, d.rb_data);
            img.put_scalar<uint32_t>("model.cb.action", cb->action_id().unchecked_get());
        }
        else if (auto* pe = dynamic_cast<LivermorePEModel const*>(&model))
        {
            auto const& d = pe->host_ref();
            img.put("model.pe.ids",
                    U32{pe->action_id().unchecked_get(),
                        raw(d.ids.electron),
                        raw(d.ids.gamma)});
            img.put_scalar<double>("model.pe.inv_electron_mass",
                                   d.inv_electron_mass);
            U32 el_rows, sh_rows;
            F64 el_thresh, sh_reals;
            for (auto const& el : all(d.xs.elements))
            {
                el_rows.push_back(el.xs_lo.grid.empty()
                                      ? 0
                                      : el.xs_lo.grid.begin()->unchecked_get());
                el_rows.push_back(el.xs_lo.grid.size());
                el_rows.push_back(el.xs_lo.value.empty()
                                      ? 0
                                      : el.xs_lo.value.begin()->unchecked_get());
                el_rows.push_back(el.xs_hi.grid.begin()->unchecked_get());
                el_rows.push_back(el.xs_hi.grid.size());
                el_rows.push_back(el.xs_hi.value.begin()->unchecked_get());
                el_rows.push_back(el.shells.begin()->unchecked_get());
                el_rows.push_back(el.shells.size());
                el_thresh.push_back(el.thresh_lo.value());
                el_thresh.push_back(el.thresh_hi.value());
            }
            for (auto const& sh : all(d.xs.shells))
            {
                sh_rows.push_back(sh.xs.grid.begin()->unchecked_get());
                sh_rows.push_back(sh.xs.grid.size());
                sh_rows.push_back(sh.xs.value.begin()->unchecked_get());
                sh_rows.push_back(0);
                sh_reals.push_back(sh.binding_energy.value());
                for (int k = 0; k < 2; ++k)
                    for (int i = 0; i < 6; ++i)
                        sh_reals.push_back(sh.param[k][i]);
            }
            img.put("model.pe.elements", el_rows);
            img.put("model.pe.element_thresh", el_thresh);
            img.put("model.pe.shells", sh_rows);
            img.put("model.pe.shell_reals", sh_reals);
            F64 reals(all(d.xs.reals).begin(), all(d.xs.reals).end());
            img.put("model.pe.reals", reals);
        }
        else if (dynamic_cast<BraggModel const*>(&model)
                 || dynamic_cast<ICRU73QOModel const*>(&model)
                 || dynamic_cast<BetheBlochModel const*>(&model)
                 || dynamic_cast<MuBetheBlochModel const*>(&model))
        {
            // em/data/MuHadIonizationData.hh: one interactor, four energy samplers
            MuHadIonizationData const* d = nullptr;
            uint32_t which = 0;
            if (auto* m = dynamic_cast<BraggModel const*>(&model))
            {
                d = &m->host_ref();
                which = 0;
            }
            else if (auto* m = dynamic_cast<ICRU73QOModel const*>(&model))
            {
                d = &m->host_ref();
                which = 1;
            }
            else if (auto* m = dynamic_cast<BetheBlochModel const*>(&model))
            {
                d = &m->host_ref();
                which = 2;
            }
            else if (auto* m = dynamic_cast<MuBetheBlochModel const*>(&model))
            {
                d = &m->host_ref();
                which = 3;
            }
            muioni_actions[which] = model.action_id().unchecked_get();
            img.put("model.muioni.electron", U32{raw(d->electron)});
            img.put("model.muioni.reals",
                    F64{d->electron_mass.value(),
                        native_value_to<units::MevMass>(constants::proton_mass)
                            .value(),
                        constants::alpha_fine_structure / (2 * constants::pi)});
            img.put("model.muioni.actions", muioni_actions);
        }
        else if (auto* mub = dynamic_cast<MuBremsstrahlungModel const*>(&model))
        {
            auto const& d = mub->host_ref();
            img.put("model.mubrems.ids",
                    U32{mub->action_id().unchecked_get(),
                        raw(d.gamma),
                        raw(d.mu_minus),
                        raw(d.mu_plus)});
            real_type const me = d.electron_mass.value();
            // em/xs/MuBremsDiffXsCalculator.hh:147,195-197
            img.put("model.mubrems.reals",
                    F64{me,
                        std::sqrt(constants::euler),
                        16 * constants::alpha_fine_structure
                            * constants::na_avogadro
                            * ipow<2>(me * constants::r_electron)});
        }
        else if (auto* cs = dynamic_cast<CoulombScatteringModel const*>(&model))
        {
            // em/data/CoulombScatteringData.hh + em/data/WentzelOKVIData.hh (the shared
            // Wentzel OK&VI data CoreParams carries for this model)
            CELER_VALIDATE(prob.core->wentzel(),
                           << "Coulomb scattering without Wentzel OK&VI data");
            auto const& d = cs->host_ref();
            auto const& w = prob.core->wentzel()->host_ref();
            img.put("model.coulomb.ids",
                    U32{cs->action_id().unchecked_get(),
                        raw(d.ids.electron),
                        raw(d.ids.positron),
                        static_cast<uint32_t>(w.params.is_combined),
                        static_cast<uint32_t>(w.params.form_factor_type)});
            // constants as WentzelHelper / NuclearFormFactors compute them
            // (em/xs/WentzelHelper.hh:253-283, em/xs/NuclearFormFactors.hh)
            constexpr real_type ctf = 0.8853413770001135;
            img.put("model.coulomb.reals",
                    F64{w.params.costheta_limit,
                        w.params.screening_factor,
                        w.params.a_sq_factor,
                        native_value_to<units::MevMomentumSq>(
                            ipow<2>(constants::hbar_planck
                                    / (2 * ctf * constants::a0_bohr)))
                            .value(),
                        2 * constants::pi
                            * ipow<2>(native_value_to<units::MevMass>(
                                          constants::electron_mass)
                                          .value()
                                      * constants::r_electron),
                        constants::alpha_fine_structure,
                        value_as<NuclearFormFactorTraits::InvMomentum>(
                            NuclearFormFactorTraits::fm_par_hbar())});
            F64 prefactor(all(w.nuclear_form_prefactor).begin(),
                          all(w.nuclear_form_prefactor).end());
            img.put("model.coulomb.nuclear_form_prefactor", prefactor);
            F64 mott;  // [element][electron, positron][theta 5][beta 6]
            for (auto const& el : all(w.mott_coeffs))
            {
                for (auto const* mat : {&el.electron, &el.positron})
                    for (auto const& row : *mat)
                        for (real_type v : row)
                            mott.push_back(v);
            }
            img.put("model.coulomb.mott", mott);
            F64 inv_mass(all(w.inv_mass_cbrt_sq).begin(),
                         all(w.inv_mass_cbrt_sq).end());
            img.put("model.coulomb.inv_mass_cbrt_sq", inv_mass);
        }
        else if (auto* ray = dynamic_cast<RayleighModel const*>(&model))
        {
            // em/data/RayleighData.hh: the gamma id and nine form-factor fit parameters
            // (a[3], b[3], n[3]) per element
            auto const& d = ray->host_ref();
            img.put("model.rayleigh.ids",
                    U32{ray->action_id().unchecked_get(), raw(d.gamma)});
            F64 reals;
            for (auto const& el : all(d.params))
            {
                for (int i = 0; i < 3; ++i)
                    reals.push_back(el.a[i]);
                for (int i = 0; i < 3; ++i)
                    reals.push_back(el.b[i]);
                for (int i = 0; i < 3; ++i)
                    reals.push_back(el.n[i]);
            }
            img.put("model.rayleigh.params", reals);
            // RayleighInteractor.hh:180-182: factor = (cm / (c h) * E_native)^2
            img.put("model.rayleigh.consts",
                    F64{units::centimeter
                            / (constants::c_light * constants::h_planck),
                        native_value_from(units::MevEnergy{1})});
        }
        else
        {
            CELER_VALIDATE(false,
                           << "model '" << model.label()
                           << "' has no B200 export");
        }
    }
    img.put_string("model.labels", labels);

    // Urban MSC
    if (prob.msc)
    {
        auto const& d = prob.msc->host_ref();
        img.put("msc.ids", U32{raw(d.ids.electron), raw(d.ids.positron)});
        auto const& pr = d.params;
        img.put("msc.params",
                F64{d.electron_mass.value(),
                    pr.tau_small,
                    pr.tau_big,
                    pr.tau_limit,
                    pr.safety_tol,
                    pr.geom_limit,
                    pr.low_energy_limit.value(),
                    pr.high_energy_limit.value()});
        F64 md, pm, gf;
        U32 gu;
        for (auto const& m : all(d.material_data))
        {
            md.push_back(m.stepmin_coeff[0]);
            md.push_back(m.stepmin_coeff[1]);
            md.push_back(m.theta_coeff[0]);
            md.push_back(m.theta_coeff[1]);
            md.push_back(m.tail_coeff[0]);
            md.push_back(m.tail_coeff[1]);
            md.push_back(m.tail_coeff[2]);
            md.push_back(m.tail_corr);
        }
        for (auto const& m : all(d.par_mat_data))
        {
            pm.push_back(m.scaled_zeff);
            pm.push_back(m.d_over_r);
        }
        for (auto const& g : all(d.xs))
        {
            gu.push_back(g.log_energy.size);
            gu.push_back(g.prime_index);
            gu.push_back(g.value.begin()->unchecked_get());
            gf.push_back(g.log_energy.front);
            gf.push_back(g.log_energy.back);
            gf.push_back(g.log_energy.delta);
        }
        img.put("msc.material_data", md);
        img.put("msc.par_mat_data", pm);
        img.put("msc.xs_grid_u32", gu);
        img.put("msc.xs_grid_f64", gf);
        F64 reals(all(d.reals).begin(), all(d.reals).end());
        img.put("msc.reals", reals);
    }
    // Energy loss fluctuations
    if (prob.fluct)
    {
        auto const& d = prob.fluct->host_ref();
        img.put_scalar<uint32_t>("fluct.electron", raw(d.electron_id));
        img.put_scalar<double>("fluct.electron_mass", d.electron_mass.value());
        F64 u;
        for (auto const& m : all(d.urban))
        {
            for (int i = 0; i < 2; ++i)
                u.push_back(m.binding_energy[i]);
            for (int i = 0; i < 2; ++i)
                u.push_back(m.log_binding_energy[i]);
            for (int i = 0; i < 2; ++i)
                u.push_back(m.oscillator_strength[i]);
        }
        img.put("fluct.urban", u);
    }
    // Uniform field
    if (prob.has_field)
    {
        F64 f{prob.field.field[0], prob.field.field[1], prob.field.field[2]};
        auto const& o = prob.field.options;
        img.put("field.uniform", f);
        img.put("field.options",
                F64{o.minimum_step,
                    o.delta_chord,
                    o.delta_intersection,
                    o.epsilon_step,
                    o.epsilon_rel_max,
                    o.errcon,
                    o.pgrow,
                    o.pshrink,
                    o.safety,
                    o.max_stepping_increase,
                    o.max_stepping_decrease,
                    native_value_from(units::ElementaryCharge{1})
                        / native_value_from(units::MevMomentum{1})});
        img.put("field.options_u32", U32{static_cast<uint32_t>(o.max_nsteps), static_cast<uint32_t>(o.max_substeps)});
    }
    // RZ field map (field/RZMapFieldParams.cc:30-84): uniform grids in z and r, values in
    // native units, element (iz, ir) at iz * num_grid_r + ir
    if (prob.has_rz_field)
    {
        auto const& in = prob.rz_field;
        auto gr = UniformGridData::from_bounds(in.min_r, in.max_r, in.num_grid_r);
        auto gz = UniformGridData::from_bounds(in.min_z, in.max_z, in.num_grid_z);
        img.put("field.rz_grid",
                F64{gz.front, gz.back, gz.delta, gr.front, gr.back, gr.delta});
        img.put("field.rz_sizes", U32{static_cast<uint32_t>(gz.size), static_cast<uint32_t>(gr.size)});
        F64 values;
        for (std::size_t i = 0; i < in.field_z.size(); ++i)
        {
            values.push_back(in.field_z[i]);
            values.push_back(in.field_r[i]);
        }
        img.put("field.rz_values", values);
        auto const& o = in.driver_options;
        img.put("field.options",
                F64{o.minimum_step,
                    o.delta_chord,
                    o.delta_intersection,
                    o.epsilon_step,
                    o.epsilon_rel_max,
                    o.errcon,
                    o.pgrow,
                    o.pshrink,
                    o.safety,
                    o.max_stepping_increase,
                    o.max_stepping_decrease,
                    native_value_from(units::ElementaryCharge{1})
                        / native_value_from(units::MevMomentum{1})});
        img.put("field.options_u32", U32{static_cast<uint32_t>(o.max_nsteps), static_cast<uint32_t>(o.max_substeps)});
    }
    // Physical constants as the reference computes them
    img.put("constants",
            F64{celeritas::detail::migdal_constant(),
                value_as<celeritas::detail::MevPerLen>(
                    celeritas::detail::lpm_constant()),
                constants::r_electron,
                constants::alpha_fine_structure});
}

}  // namespace

//---------------------------------------------------------------------------//
namespace
{
b200::Image build_image(Problem const& prob);
}

void export_image(Problem const& prob, std::string const& path)
{
    build_image(prob).write(path);
}

std::vector<unsigned char> export_image_bytes(Problem const& prob)
{
    return build_image(prob).serialize();
}

namespace
{
b200::Image build_image(Problem const& prob)
{
    b200::Image img;
    img.put_string("config", prob.config.dump());
    if (!prob.core)
    {
        // Geometry-only image
        CELER_VALIDATE(prob.geo, << "empty problem");
        export_geometry(prob.geo->host_ref(), img);
        img.put_string("geo.volume_labels", image_volume_labels(*prob.geo));
        return img;
    }
    CoreParams const& core = *prob.core;
    auto const& ref = core.host_ref();


    // Core scalars and the action table
    {
        auto const& s = ref.scalars;
        img.put("core.actions",
                U32{raw(s.boundary_action),
                    raw(s.propagation_limit_action),
                    raw(s.tracking_cut_action),
                    raw(s.along_step_user_action),
                    raw(s.along_step_neutral_action)});
        std::string labels;
        U32 order;
        auto const& reg = *core.action_reg();
        for (auto aid : range(ActionId{reg.num_actions()}))
        {
            auto const& a = *reg.action(aid);
            labels += std::string(a.label()) + "\n";
            auto* step = dynamic_cast<CoreStepActionInterface const*>(&a);
            order.push_back(step ? static_cast<uint32_t>(step->order())
                                 : invalid);
        }
        img.put_string("core.action_labels", labels);
        img.put("core.action_order", order);
    }

    export_geometry(ref.geometry, img);
    {
        // volume -> material
        U32 vm;
        for (auto m : all(ref.geo_mats.materials))
            vm.push_back(raw(m));
        img.put("geomat.volume_material", vm);
        // volume labels (for detector maps / diagnostics)
        img.put_string("geo.volume_labels", image_volume_labels(*core.geometry()));
    }
    export_materials(ref.materials, img);

    // Particles
    {
        F64 mass, charge, decay;
        U8 matter;
        for (auto v : all(ref.particles.mass))
            mass.push_back(v.value());
        for (auto v : all(ref.particles.charge))
            charge.push_back(v.value());
        for (auto v : all(ref.particles.decay_constant))
            decay.push_back(v);
        for (auto v : all(ref.particles.matter))
            matter.push_back(static_cast<uint8_t>(v));
        img.put("particle.mass", mass);
        img.put("particle.charge", charge);
        img.put("particle.decay_constant", decay);
        img.put("particle.matter", matter);
        U32 pdg;
        std::string names;
        auto const& pp = *core.particle();
        for (auto pid : range(ParticleId{pp.size()}))
        {
            pdg.push_back(static_cast<uint32_t>(pp.id_to_pdg(pid).get()));
            names += pp.id_to_label(pid) + "\n";
        }
        img.put("particle.pdg", pdg);
        img.put_string("particle.names", names);
    }
    // Cutoffs
    {
        auto const& c = ref.cutoffs;
        F64 energy, range_;
        for (auto const& pc : all(c.cutoffs))
        {
            energy.push_back(pc.energy.value());
            range_.push_back(pc.range);
        }
        img.put("cutoff.energy", energy);
        img.put("cutoff.range", range_);
        U32 idx;
        for (auto v : all(c.id_to_index))
            idx.push_back(v);
        img.put("cutoff.id_to_index", idx);
        img.put("cutoff.scalars",
                U32{c.num_particles,
                    c.num_materials,
                    c.apply_post_interaction ? 1u : 0u,
                    raw(c.ids.gamma),
                    raw(c.ids.electron),
                    raw(c.ids.positron)});
    }
    export_physics(ref.physics, ref.materials.materials.size(), img);
    export_models(prob, img);

    // RNG
    {
        U32 rng;
        rng.push_back(ref.rng.seed[0]);
        for (auto const& poly : ref.rng.jump)
            for (auto w : poly)
                rng.push_back(w);
        for (auto const& poly : ref.rng.jump_subsequence)
            for (auto w : poly)
                rng.push_back(w);
        img.put("rng.params", rng);  // seed, jump[32][5], jump_subsequence[32][5]
    }
    // Sim
    {
        U32 loop;
        F64 thresh;
        for (auto const& l : all(ref.sim.looping))
        {
            loop.push_back(l.max_subthreshold_steps);
            loop.push_back(l.max_steps);
            thresh.push_back(l.threshold_energy.value());
        }
        img.put("sim.looping_steps", loop);
        img.put("sim.looping_energy", thresh);
    }
    // Track init
    img.put("init.scalars",
            U32{ref.init.capacity,
                ref.init.max_events,
                static_cast<uint32_t>(ref.init.track_order)});

    // Detector map for SimpleCalo
    {
        std::string names;
        for (auto const& n : prob.calo_volumes)
            names += n + "\n";
        img.put_string("calo.volumes", names);
    }
    // Sensitive volumes of the step/hit output (HitRecorder): detector id = position
    if (prob.hits)
    {
        std::string names;
        for (auto const& n : prob.hit_volumes)
            names += n + "\n";
        img.put_string("hits.volumes", names);
        img.put_scalar<uint32_t>("hits.nonzero_edep", prob.hits_nonzero_edep ? 1u : 0u);
    }

    return img;
}
}  // namespace
}  // namespace celerref
