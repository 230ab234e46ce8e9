//---------------------------------------------------------------------------//
// Discrete interactions: common result type, the post-interaction applier and
// the individual EM interactors.
//
// Reference: InteractionApplier (/root/reference/src/celeritas/phys/InteractionApplier.hh:104-171),
// Interaction/Secondary (phys/Interaction.hh:31-63, phys/Secondary.hh:23-34).
// Secondaries are written to fixed per-slot storage (MAX_SECONDARIES entries)
// instead of a shared atomically-allocated stack: the EM interactors here emit
// at most two, and fixed storage keeps the writes coalesced and deterministic.
//---------------------------------------------------------------------------//
#pragma once

#include "orange.cuh"
#include "physics.cuh"

namespace b200
{
enum InteractionAction : u8
{
    IA_SCATTERED = 0,
    IA_ABSORBED = 1,
    IA_UNCHANGED = 2,
    IA_FAILED = 3
};

struct SecondaryOut
{
    u32 particle;  // INVALID = empty
    real energy;
    Real3 direction;
};

struct Interaction
{
    real energy;
    Real3 direction;
    real energy_deposition;
    u8 action;
    u32 num_secondaries;
    SecondaryOut sec[MAX_SECONDARIES];

    B2_D Interaction() : energy(0), energy_deposition(0), action(IA_SCATTERED), num_secondaries(0)
    {
        direction = make_real3(0, 0, 0);
        for (int i = 0; i < MAX_SECONDARIES; ++i)
            sec[i].particle = INVALID;
    }
    B2_D static Interaction from_absorption()
    {
        Interaction r;
        r.energy = 0;
        r.action = IA_ABSORBED;
        return r;
    }
    B2_D static Interaction from_unchanged()
    {
        Interaction r;
        r.action = IA_UNCHANGED;
        return r;
    }
};

//! p_in * d_in - p_out * d_out, normalised (calc_exiting_direction)
B2_D Real3 calc_exiting_direction(real inc_mag, Real3 const& inc_dir, real out_mag, Real3 const& out_dir)
{
    Real3 r;
    for (int i = 0; i < 3; ++i)
        r[i] = inc_dir[i] * inc_mag - out_dir[i] * out_mag;
    return make_unit_vector(r);
}

B2_D bool cutoff_applies(ParamsView const& pv, u32 material, SecondaryOut const& sec)
{
    CutoffParams const& c = pv.cutoff;
    if (!(sec.particle == c.id_gamma || sec.particle == c.id_electron
          || sec.particle == c.id_positron))
        return false;
    real e = c.energy[c.num_materials * c.id_to_index[sec.particle] + material];
    return sec.energy < e;
}

//! Write the outcome of an interaction into the track state
B2_D void apply_interaction(ParamsView const& pv,
                            StateView const& s,
                            u32 slot,
                            Interaction& result)
{
    if (result.action == IA_FAILED)
    {
        // step_limit({0, failure_action})
        if (0 < s.step_length[slot])
        {
            s.step_length[slot] = 0;
            s.post_step_action[slot] = pv.phys.model_to_action + pv.phys.num_models;
        }
        return;
    }
    if (result.action == IA_UNCHANGED)
        return;

    s.energy[slot] = result.energy;
    if (result.action != IA_ABSORBED)
    {
        GeoTrack geo(pv, s, slot);
        geo.set_dir(result.direction);
    }
    else
    {
        s.status[slot] = ST_KILLED;
    }
    real deposition = result.energy_deposition;
    u32 const material = s.material_id[slot];
    if (pv.cutoff.apply_post_interaction)
    {
        for (u32 i = 0; i < result.num_secondaries; ++i)
        {
            SecondaryOut& sec = result.sec[i];
            if (sec.particle != INVALID && cutoff_applies(pv, material, sec))
            {
                deposition += sec.energy;
                if (particle_is_antiparticle(pv, sec.particle))
                    deposition += 2 * pv.particle.mass[sec.particle];
                sec.particle = INVALID;
            }
        }
    }
    s.energy_deposition[slot] += deposition;
    u32 const n = s.num_slots;
    for (u32 i = 0; i < MAX_SECONDARIES; ++i)
    {
        SecondaryOut const& sec = result.sec[i];
        bool live = i < result.num_secondaries && sec.particle != INVALID;
        s.sec_particle[i * n + slot] = live ? sec.particle : INVALID;
        if (live)
        {
            s.sec_energy[i * n + slot] = sec.energy;
            for (int k = 0; k < 3; ++k)
                s.sec_dir[(i * 3 + k) * n + slot] = sec.direction[k];
        }
    }
}

//---------------------------------------------------------------------------//
// Klein-Nishina Compton scattering
// (/root/reference/src/celeritas/em/interactor/KleinNishinaInteractor.hh:104-190)
//---------------------------------------------------------------------------//
B2_D Interaction interact_klein_nishina(KleinNishinaParams const& shared,
                                        real inc_energy,
                                        Real3 const& inc_direction,
                                        Rng& rng)
{
    real const inc_energy_per_mecsq = inc_energy * shared.inv_electron_mass;
    real const epsilon_0 = 1 / (1 + 2 * inc_energy_per_mecsq);
    real const f1 = -log(epsilon_0);
    real const f2 = real(0.5) * (1 - ipow2(epsilon_0));
    real const p_f1 = f1 / (f1 + f2);
    ReciprocalDist sample_f1(epsilon_0);
    real const eps0_sq = ipow2(epsilon_0);

    real epsilon, one_minus_costheta, reject_prob;
    do
    {
        real epsilon_sq;
        if (rng.canonical() < p_f1)
        {
            epsilon = sample_f1(rng);
            epsilon_sq = epsilon * epsilon;
        }
        else
        {
            epsilon_sq = sample_uniform(rng, eps0_sq, 1);
            epsilon = sqrt(epsilon_sq);
        }
        one_minus_costheta = (1 - epsilon) / (epsilon * inc_energy_per_mecsq);
        real sintheta_sq = one_minus_costheta * (2 - one_minus_costheta);
        reject_prob = epsilon * sintheta_sq / (1 + epsilon_sq);
    } while (rng.canonical() < reject_prob);

    Interaction result;
    result.energy = epsilon * inc_energy;
    result.direction = sample_exiting_direction(rng, 1 - one_minus_costheta, inc_direction);
    result.num_secondaries = 1;
    real electron_energy = inc_energy - result.energy;
    if (electron_energy < 1e-4)  // secondary_cutoff()
    {
        result.energy_deposition = electron_energy;
        result.sec[0].particle = INVALID;
        return result;
    }
    result.sec[0].particle = shared.electron;
    result.sec[0].energy = electron_energy;
    result.sec[0].direction
        = calc_exiting_direction(inc_energy, inc_direction, result.energy, result.direction);
    return result;
}

//---------------------------------------------------------------------------//
// On-the-fly macroscopic cross sections for "hardwired" models
//---------------------------------------------------------------------------//
B2_D real calc_hardwired_xs(ParamsView const& pv, u32 model, u32 material, real energy)
{
    (void)pv;
    (void)model;
    (void)material;
    (void)energy;
    return 0;
}

//! Dispatch the interaction for a model action
B2_D void run_interaction(ParamsView const& pv, StateView const& s, u32 slot, u32 action, Rng& rng)
{
    Interaction result = Interaction::from_unchanged();
    GeoTrack geo(pv, s, slot);
    Real3 const dir = geo.dir();
    real const energy = s.energy[slot];
    if (action == pv.model.kn.action)
    {
        result = interact_klein_nishina(pv.model.kn, energy, dir, rng);
    }
    apply_interaction(pv, s, slot, result);
}
}  // namespace b200
