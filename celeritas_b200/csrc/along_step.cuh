//---------------------------------------------------------------------------//
// Along-step: propagation, (multiple scattering, energy loss), time and
// mean-free-path bookkeeping for one track.
//
// Order of operations and RNG draws per slot follows the reference's
// AlongStep functor (/root/reference/src/celeritas/global/alongstep/AlongStep.hh:50-58):
// msc step limit -> propagate -> msc scatter -> time -> energy loss -> track
// update.
//---------------------------------------------------------------------------//
#pragma once

#include "orange.cuh"
#include "physics.cuh"

namespace b200
{
//! Straight-line propagation up to `dist` (field/LinearPropagator.hh:58-93)
B2_D Propagation propagate_linear(GeoTrack& geo, real dist)
{
    Propagation result = geo.find_next_step(true, dist);
    if (result.boundary)
        geo.move_to_boundary();
    else
        geo.move_internal(dist);
    return result;
}

//! Apply the propagation result to the step (detail/PropagationApplier.hh:93-192)
B2_D void apply_propagation(ParamsView const& p, StateView const& s, u32 slot, Propagation const& pr)
{
    if (pr.boundary)
    {
        s.step_length[slot] = pr.distance;
        s.post_step_action[slot] = p.scalars.boundary_action;
    }
    else if (pr.distance < s.step_length[slot])
    {
        s.step_length[slot] = pr.distance;
        s.post_step_action[slot] = p.scalars.propagation_limit_action;
    }
}

//! t += step / v (detail/TimeUpdater.hh:28-46)
B2_D void update_time(StateView const& s, u32 slot, Particle const& particle)
{
    real speed = particle.speed() * constants::c_light;
    if (speed > 0)
        s.time[slot] += s.step_length[slot] / speed;
}

//! mfp -= step * xs; ++num_steps (detail/TrackUpdater.hh:32-67)
B2_D void update_track(ParamsView const& p, StateView const& s, u32 slot)
{
    u8 status = s.status[slot];
    if (status == ST_ERRORED)
        return;
    if (status == ST_ALIVE)
    {
        if (s.post_step_action[slot] != p.phys.model_to_action - 2)
        {
            s.interaction_mfp[slot]
                = s.interaction_mfp[slot] - s.step_length[slot] * s.macro_xs[slot];
        }
    }
    s.num_steps[slot] += 1;
}

//! Whole along-step for one alive track
B2_D void along_step(ParamsView const& p, StateView const& s, u32 slot)
{
    Particle particle = load_particle(p, s, slot);
    GeoTrack geo(p, s, slot);
    if (s.step_length[slot] != 0)
    {
        Propagation pr = propagate_linear(geo, s.step_length[slot]);
        apply_propagation(p, s, slot, pr);
    }
    if (s.status[slot] != ST_ERRORED)
        update_time(s, slot, particle);
    update_track(p, s, slot);
}
}  // namespace b200
