//---------------------------------------------------------------------------//
// Physics step limits, discrete-process selection and particle kinematics.
//
// Behaviour follows the reference's calc_physics_step_limit /
// select_discrete_interaction
// (/root/reference/src/celeritas/phys/PhysicsStepUtils.hh:36-120,252-308) and
// PhysicsTrackView (/root/reference/src/celeritas/phys/PhysicsTrackView.hh:340-640)
// on flat [particle][process][material] index tables.
//---------------------------------------------------------------------------//
#pragma once

#include "grid.cuh"
#include "rng.cuh"
#include "views.cuh"

namespace b200
{
//---------------------------------------------------------------------------//
// Particle kinematics (ParticleTrackView.hh:303-391)
//---------------------------------------------------------------------------//
struct Particle
{
    u32 id;
    real energy;
    real mass;
    real charge;

    B2_D bool is_stopped() const { return energy == 0; }
    B2_D real total_energy() const { return energy + mass; }
    B2_D real beta_sq() const
    {
        real inv_gamma = mass / (energy + mass);
        return 1 - ipow2(inv_gamma);
    }
    B2_D real speed() const { return sqrt(beta_sq()); }  // units of c
    B2_D real lorentz_factor() const { return 1 + energy / mass; }
    B2_D real momentum_sq() const { return ipow2(energy) + 2 * mass * energy; }
    B2_D real momentum() const { return sqrt(momentum_sq()); }
};

B2_D Particle load_particle(ParamsView const& p, StateView const& s, u32 slot)
{
    Particle q;
    q.id = s.particle_id[slot];
    q.energy = s.energy[slot];
    q.mass = p.particle.mass[q.id];
    q.charge = p.particle.charge[q.id];
    return q;
}

B2_D bool particle_is_antiparticle(ParamsView const& p, u32 pid)
{
    return p.particle.matter[pid] != 0;
}

//---------------------------------------------------------------------------//
// Material helpers
//---------------------------------------------------------------------------//
B2_D u32 material_num_elements(MatParams const& m, u32 mat)
{
    return m.material_elcomp_end[mat] - m.material_elcomp_begin[mat];
}

B2_D real material_real(MatParams const& m, u32 mat, MaterialReal which)
{
    return m.material_reals[mat * MAT_NUM_REALS + which];
}

B2_D real element_real(MatParams const& m, u32 el, ElementReal which)
{
    return m.element_reals[el * EL_NUM_REALS + which];
}

//---------------------------------------------------------------------------//
// Per-track physics view
//---------------------------------------------------------------------------//
struct PhysTrack
{
    ParamsView const& pv;
    PhysParams const& p;
    u32 particle;
    u32 material;

    B2_D PhysTrack(ParamsView const& params, u32 particle_id, u32 material_id)
        : pv(params), p(params.phys), particle(particle_id), material(material_id)
    {
    }

    B2_D u32 num_processes() const { return p.pp_num[particle]; }
    B2_D u32 row(u32 ppid) const { return particle * p.max_processes + ppid; }
    B2_D u32 process(u32 ppid) const { return p.pp_process[row(ppid)]; }
    B2_D u32 eloss_ppid() const { return p.pp_eloss_ppid[particle]; }
    B2_D bool has_at_rest() const { return p.pp_has_at_rest[particle] != 0; }

    B2_D u32 value_grid(int vgt, u32 ppid) const
    {
        return p.pp_grid[((vgt * p.num_particles + particle) * p.max_processes + ppid)
                             * p.num_materials
                         + material];
    }

    //! Model applicable at this energy (GridIdFinder.hh): returns pmid or INVALID
    B2_D u32 find_model(u32 ppid, real energy) const
    {
        u32 r = row(ppid);
        u32 n = p.pp_model_count[r];
        real const* grid = p.pm_energy + p.pm_energy_begin[r];
        u32 const* values = p.pm_pmid + p.pp_model_begin[r];
        // lower_bound over n+1 bounds
        u32 lo = 0, len = n + 1;
        while (len > 0)
        {
            u32 half = len >> 1;
            u32 mid = lo + half;
            if (grid[mid] < energy)
            {
                lo = mid + 1;
                len -= half + 1;
            }
            else
                len = half;
        }
        if (lo == n + 1)
            return INVALID;
        if (lo == 0 && energy != grid[0])
            return INVALID;
        if (lo + 1 == n + 1 || energy != grid[lo])
            --lo;
        return values[lo];
    }

    B2_D u32 model_id(u32 pmid) const { return p.pmid_model[pmid]; }
    B2_D u32 model_to_action(u32 model) const { return model + p.model_to_action; }

    //! Model id if this process computes its xs on the fly at this energy
    B2_D u32 hardwired_model(u32 ppid, real energy) const
    {
        u32 proc = process(ppid);
        if ((proc == p.hw_photoelectric && energy < p.hw_photoelectric_table_thresh)
            || (proc == p.hw_positron_annihilation))
        {
            u32 pmid = find_model(ppid, energy);
            return pmid == INVALID ? INVALID : model_id(pmid);
        }
        return INVALID;
    }

    B2_D real calc_xs(u32 ppid, real energy) const;
    B2_D real calc_max_xs(u32 ppid, real energy) const
    {
        real energy_max_xs = p.pp_energy_max_xs[row(ppid) * p.num_materials + material];
        real energy_xi = energy * p.min_eprime_over_e;
        if (energy_max_xs >= energy_xi && energy_max_xs < energy)
            return calc_xs(ppid, energy_max_xs);
        real a = calc_xs(ppid, energy);
        real b = calc_xs(ppid, energy_xi);
        return a > b ? a : b;
    }

    //! Scaled range -> step (PhysicsTrackView::range_to_step)
    B2_D real range_to_step(real range) const
    {
        real const rho = p.min_range;
        // sqrt_tol for double = 1e-6 (corecel/math/SoftEqual.hh)
        if (range < rho * (1 + 1e-6))
            return range;
        real const alpha = p.max_step_over_range;
        return alpha * range + rho * (1 - alpha) * (2 - rho / range);
    }
};

// Hardwired on-the-fly cross sections are defined with the models
B2_D real calc_hardwired_xs(ParamsView const& pv, u32 model, u32 material, real energy);

B2_D real PhysTrack::calc_xs(u32 ppid, real energy) const
{
    real result = 0;
    u32 hw = hardwired_model(ppid, energy);
    if (hw != INVALID)
    {
        result = calc_hardwired_xs(pv, hw, material, energy);
    }
    else
    {
        u32 grid = value_grid(VGT_MACRO_XS, ppid);
        if (grid != INVALID)
            result = b200::calc_xs(p, grid, energy);
    }
    return result;
}

struct StepLimit
{
    real step;
    u32 action;
};

//! Physics step limit; stores per-process xs, macro xs and range in state
B2_D StepLimit calc_physics_step_limit(ParamsView const& pv,
                                       StateView const& s,
                                       u32 slot,
                                       Particle const& particle,
                                       PhysTrack const& phys)
{
    PhysParams const& p = pv.phys;
    real total = 0;
    u32 const np = phys.num_processes();
    for (u32 ppid = 0; ppid < np; ++ppid)
    {
        real xs;
        if (p.pp_integral[phys.row(ppid)])
            xs = phys.calc_max_xs(ppid, particle.energy);
        else
            xs = phys.calc_xs(ppid, particle.energy);
        total += xs;
        s.per_process_xs[ppid * s.num_slots + slot] = xs;
    }
    s.macro_xs[slot] = total;

    StepLimit limit;
    limit.action = p.model_to_action - 2;  // discrete action
    if (particle.is_stopped())
    {
        limit.step = 0;
    }
    else
    {
        limit.step = s.interaction_mfp[slot] / total;
        u32 eppid = phys.eloss_ppid();
        if (eppid != INVALID)
        {
            u32 grid = phys.value_grid(VGT_RANGE, eppid);
            real range = calc_range(p, grid, particle.energy);
            s.dedx_range[slot] = range;
            real eloss_step = phys.range_to_step(range);
            if (eloss_step <= limit.step)
            {
                limit.step = eloss_step;
                limit.action = p.model_to_action - 3;  // range action
            }
            if (p.fixed_step_limiter > 0 && p.fixed_step_limiter < limit.step)
            {
                limit.step = p.fixed_step_limiter;
                limit.action = p.fixed_step_action;
            }
        }
        else if (np == 0)
        {
            limit.action = INVALID;
        }
    }
    return limit;
}

//! Sample the element for a tabulated micro-xs CDF (TabulatedElementSelector)
B2_D u32 select_element_tabulated(PhysParams const& p, u32 begin, u32 count, real energy, Rng& rng)
{
    u32 i = 0;
    real u = rng.canonical();
    for (; i < count - 1; ++i)
    {
        if (calc_xs(p, p.elsel_grid[begin + i], energy) > u)
            break;
    }
    return i;
}

//! Choose the discrete interaction; returns the post-step action id
B2_D u32 select_discrete_interaction(ParamsView const& pv,
                                     StateView const& s,
                                     u32 slot,
                                     Particle const& particle,
                                     PhysTrack const& phys,
                                     Rng& rng)
{
    PhysParams const& p = pv.phys;
    u32 const np = phys.num_processes();
    // Selector: accumulate until exceeding total * xi
    u32 ppid;
    {
        real accum = -s.macro_xs[slot] * rng.canonical();
        ppid = np - 1;
        for (u32 i = 0; i + 1 < np; ++i)
        {
            accum += s.per_process_xs[i * s.num_slots + slot];
            if (accum > 0)
            {
                ppid = i;
                break;
            }
        }
    }
    if (p.pp_integral[phys.row(ppid)])
    {
        real xs = phys.calc_xs(ppid, particle.energy);
        if (rng.canonical() * s.per_process_xs[ppid * s.num_slots + slot] > xs)
            return p.model_to_action - 1;  // integral rejection
    }
    u32 pmid = phys.find_model(ppid, particle.energy);
    u32 elcomp = INVALID;
    if (material_num_elements(pv.mat, phys.material) == 1)
    {
        elcomp = 0;
    }
    else
    {
        u32 idx = pmid * p.num_materials + phys.material;
        u32 begin = p.elsel_begin[idx];
        if (begin != INVALID)
            elcomp = select_element_tabulated(p, begin, p.elsel_count[idx], particle.energy, rng);
    }
    s.element[slot] = elcomp;
    return phys.model_to_action(phys.model_id(pmid));
}
}  // namespace b200
