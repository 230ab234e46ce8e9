//---------------------------------------------------------------------------//
// Discrete interactions: common result type, the post-interaction applier and
// the EM interactors.
//
// Reference: InteractionApplier (/root/reference/src/celeritas/phys/InteractionApplier.hh:104-171),
// Interaction/Secondary (phys/Interaction.hh:31-63, phys/Secondary.hh:23-34).
// Secondaries are written to fixed per-slot storage (MAX_SECONDARIES entries)
// instead of a shared atomically-allocated stack: the EM interactors here emit
// at most two, and fixed storage keeps the writes coalesced and deterministic.
// Each interactor consumes random numbers in exactly the reference's order.
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

B2_D real cutoff_energy(ParamsView const& pv, u32 material, u32 particle)
{
    CutoffParams const& c = pv.cutoff;
    return c.energy[c.num_materials * c.id_to_index[particle] + material];
}

B2_D bool cutoff_applies(ParamsView const& pv, u32 material, SecondaryOut const& sec)
{
    CutoffParams const& c = pv.cutoff;
    if (!(sec.particle == c.id_gamma || sec.particle == c.id_electron
          || sec.particle == c.id_positron))
        return false;
    return sec.energy < cutoff_energy(pv, material, sec.particle);
}

//! Write the outcome of an interaction into the track state
B2_D void apply_interaction(ParamsView const& pv, StateView const& s, u32 slot, Interaction& result)
{
    if (result.action == IA_FAILED)
    {
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
// Shared samplers
//---------------------------------------------------------------------------//
//! Reject when f < fmax * xi (RejectionSampler.hh)
B2_D bool reject(Rng& rng, real f, real fmax)
{
    return f < fmax * rng.canonical();
}

//! Tsai angular distribution for brems/pair (TsaiUrbanDistribution.hh)
B2_D real sample_tsai_urban(Rng& rng, real energy, real mass)
{
    real const umax = 2 * (1 + energy / mass);
    real u;
    do
    {
        real uu = -log(rng.canonical() * rng.canonical());
        u = uu * (sample_bernoulli(rng, 0.25) ? real(1.6) : real(1.6 / 3));
    } while (u > umax);
    return 1 - 2 * ipow2(u / umax);
}

//! Horner polynomial with fma, as the reference's PolyEvaluator
template<int N>
B2_D real poly(real const (&c)[N], real x)
{
    real r = c[N - 1];
#pragma unroll
    for (int i = N - 2; i >= 0; --i)
        r = fma(x, r, c[i]);
    return r;
}
B2_D real poly_lin(real c0, real c1, real x)
{
    return fma(x, c1, c0);
}
B2_D real poly_quad(real c0, real c1, real c2, real x)
{
    return fma(x, fma(x, c2, c1), c0);
}

//---------------------------------------------------------------------------//
// Klein-Nishina Compton scattering (em/interactor/KleinNishinaInteractor.hh:104-190)
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
// The interactors beyond the north star's list (Rayleigh, Coulomb, muons) are out-of-line
// calls: the fused step and the device-resident loop carry every interactor, and these are
// never run by the benchmark problems; out of line they stay out of those kernels' register
// allocation.
//
// Rayleigh scattering (em/interactor/RayleighInteractor.hh:107-199): coherent scattering off
// the element chosen by the discrete selection; the angle is drawn from a three-term fit of
// the squared form factor, the energy is unchanged and nothing is deposited
//---------------------------------------------------------------------------//
B2_NOINLINE inline Interaction interact_rayleigh(RayleighParams const& shared,
                                   real inc_energy,
                                   Real3 const& inc_direction,
                                   u32 element,
                                   Rng& rng)
{
    real const fit_slice = 0.02;
    real pa[3], pb[3], pn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        pa[i] = shared.params[9 * element + i];
        pb[i] = shared.params[9 * element + 3 + i];
        pn[i] = shared.params[9 * element + 6 + i];
    }

    // evaluate_weight_and_prob
    real const factor = ipow2(shared.hc_factor * (inc_energy * shared.mev));
    real weight[3], prob[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        real const x = fma(factor, pb[i], pb[i]);
        real const n = pn[i];
        weight[i] = (x > fit_slice) ? 1 - exp(-n * log(1 + x))
                                    : n * x * (1 - (n - 1) / 2 * x * (1 - (n - 2) / 3 * x));
        prob[i] = weight[i] * pa[i] / (pb[i] * n);
    }
    real const inv_sum = 1 / (prob[0] + prob[1] + prob[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
        prob[i] = fma(inv_sum, prob[i], real(0));

    real cost;
    do
    {
        // Selector over the three terms (total 1; the last is never accumulated)
        real accum = -rng.canonical();
        int index = 2;
        accum += prob[0];
        if (accum > 0)
            index = 0;
        else
        {
            accum += prob[1];
            if (accum > 0)
                index = 1;
        }
        real const w = index == 0 ? weight[0] : index == 1 ? weight[1] : weight[2];
        real const ninv = 1 / (index == 0 ? pn[0] : index == 1 ? pn[1] : pn[2]);
        real const b = index == 0 ? pb[0] : index == 1 ? pb[1] : pb[2];

        real x;
        real const y = w * rng.canonical();
        if (y < fit_slice)
            x = y * ninv * (1 + real(0.5) * (ninv + 1) * y * (1 - (ninv + 2) * y / 3));
        else
            x = exp(-ninv * log(1 - y)) - 1;
        cost = 1 - 2 * x / (b * factor);
    } while (2 * rng.canonical() > 1 + ipow2(cost) || cost < -1);

    Interaction result;
    result.energy = inc_energy;
    result.direction = sample_exiting_direction(rng, cost, inc_direction);
    return result;
}

//---------------------------------------------------------------------------//
// Single Coulomb scattering off a nucleus or its electrons, Wentzel model
// (em/executor/CoulombScatteringExecutor.hh:38-65, em/interactor/
// CoulombScatteringInteractor.hh:105-176, em/xs/WentzelHelper.hh:139-325,
// em/distribution/WentzelDistribution.hh:172-313, em/xs/MottRatioCalculator.hh:73-97,
// em/xs/NuclearFormFactors.hh:184-296, mat/IsotopeSelector.hh:62-76)
//---------------------------------------------------------------------------//
//! corecel/math/Algorithms.hh fastpow: exp(b log a)
B2_D real exp_b_log_a(real a, real b)
{
    return exp(b * log(a));
}

struct WentzelHelper
{
    real target_z;
    real screening_coefficient;
    real kin_factor;
    real mott_factor;
    real cos_thetamax_elec;
    real cos_thetamax_nuc;

    B2_D WentzelHelper(CoulombParams const& w,
                       Particle const& particle,
                       u32 material,
                       u32 z,
                       real cutoff)
        : target_z(z)
    {
        real const mom_sq = particle.momentum_sq();
        real const beta_sq = particle.beta_sq();
        // Moliere screening coefficient
        {
            real correction = 1;
            real const sq_cbrt_z = exp_b_log_a(target_z, real(2) / 3);
            if (z > 1)
            {
                real const tau = particle.energy / particle.mass;
                real const factor = sqrt(tau / (tau + sq_cbrt_z));
                correction = fmin(target_z * real(1.13),
                                  real(1.13)
                                      + real(3.76) * ipow2(target_z * w.alpha_fine_structure)
                                            * factor / beta_sq);
            }
            screening_coefficient = correction * w.screen_r_sq_elec * sq_cbrt_z / mom_sq
                                    * w.screening_factor;
        }
        kin_factor = w.twopi_mrsq * target_z * ipow2(particle.charge) / (beta_sq * mom_sq);
        mott_factor = particle.id == w.electron ? 1 + real(2e-4) * ipow2(target_z) : real(1);
        // maximum scattering angle off electrons
        {
            real const inc_energy = particle.energy;
            real const max_energy = particle.id == w.electron ? real(0.5) * inc_energy
                                                              : inc_energy;
            real const final_energy = inc_energy - fmin(cutoff, max_energy);
            cos_thetamax_elec = 0;
            if (final_energy > 0)
            {
                real const incident_ratio = 1 + 2 * particle.mass / inc_energy;
                real const final_ratio = 1 + 2 * particle.mass / final_energy;
                real const cos_t_max = sqrt(incident_ratio / final_ratio);
                cos_thetamax_elec = fmin(fmax(cos_t_max, real(0)), real(1));
            }
        }
        // maximum scattering angle off a nucleus
        cos_thetamax_nuc = w.costheta_limit;
        if (w.is_combined)
        {
            cos_thetamax_nuc = fmax(
                w.costheta_limit, 1 - w.a_sq_factor * w.inv_mass_cbrt_sq[material] / mom_sq);
        }
    }

    B2_D real calc_xs_factor(real cos_thetamin, real cos_thetamax) const
    {
        return kin_factor * mott_factor * (cos_thetamin - cos_thetamax)
               / ((1 - cos_thetamin + 2 * screening_coefficient)
                  * (1 - cos_thetamax + 2 * screening_coefficient));
    }
    B2_D real calc_xs_electron(real cos_thetamin, real cos_thetamax) const
    {
        cos_thetamin = fmax(cos_thetamin, cos_thetamax_elec);
        cos_thetamax = fmax(cos_thetamax, cos_thetamax_elec);
        if (cos_thetamin <= cos_thetamax)
            return 0;
        return calc_xs_factor(cos_thetamin, cos_thetamax);
    }
    B2_D real calc_xs_nuclear(real cos_thetamin, real cos_thetamax) const
    {
        return target_z * calc_xs_factor(cos_thetamin, cos_thetamax);
    }
    //! [Fern] eqn 92 with cos(theta) = 1 - 2 mu
    B2_D real sample_costheta(real cos_thetamin, real cos_thetamax, Rng& rng) const
    {
        real const mu1 = real(0.5) * (1 - cos_thetamin);
        real const mu2 = real(0.5) * (1 - cos_thetamax);
        real const w = rng.canonical() * (mu2 - mu1);
        real const sc = screening_coefficient;
        return 1 - 2 * mu1 - 2 * (sc + mu1) * w / (sc + mu2 - w);
    }
};

//! Nuclear form factor at a squared momentum transfer (NuclearFormFactors.hh)
B2_D real nuclear_form_factor(CoulombParams const& w, u32 isotope, real mt_sq)
{
    switch (w.form_factor_type)
    {
        case 1:  // flat: folded uniform-uniform spheres
        {
            real const target_mom = sqrt(mt_sq);
            real const a = real(w.isotope_za[2 * isotope + 1]);
            real const nucl_radius_fm = real(1.2) * exp_b_log_a(a, real(1) / 3);
            auto sphere_ff = [&](real r) {
                real const x = target_mom * (r * w.fm_par_hbar);
                return (3 / (x * x * x)) * fma(-x, cos(x), sin(x));
            };
            return fmin(sphere_ff(nucl_radius_fm) * sphere_ff(real(2)), real(1));
        }
        case 2:  // exponential
            return 1 / ipow2(1 + w.nuclear_form_prefactor[isotope] * mt_sq);
        case 3:  // gaussian
            return exp(-2 * w.nuclear_form_prefactor[isotope] * mt_sq);
        default:
            return 1;
    }
}

B2_NOINLINE inline Interaction interact_coulomb(ParamsView const& pv,
                                  Particle const& particle,
                                  Real3 const& inc_direction,
                                  u32 material,
                                  u32 element,
                                  Rng& rng)
{
    CoulombParams const& w = pv.model.coulomb;

    // IsotopeSelector: the element's isotopes by number fraction
    u32 isotope;
    {
        u32 const begin = w.element_isocomp_range[2 * element];
        u32 const imax = w.element_isocomp_range[2 * element + 1] - begin - 1;
        real cumulative = -rng.canonical();
        u32 i = 0;
        for (; i < imax; ++i)
        {
            cumulative += w.isocomp_fraction[begin + i];
            if (cumulative > 0)
                break;
        }
        isotope = w.isocomp_isotope[begin + i];
    }
    u32 const z = w.isotope_za[2 * isotope];

    WentzelHelper const helper(w, particle, material, z, cutoff_energy(pv, material, w.electron));
    real const cos_thetamin = helper.cos_thetamax_nuc;
    real const cos_thetamax = -1;  // CoulombScatteringData::cos_thetamax()

    // WentzelDistribution
    real cos_theta = 1;
    if (sample_bernoulli(rng,
                         helper.calc_xs_electron(cos_thetamin, cos_thetamax),
                         helper.calc_xs_nuclear(cos_thetamin, cos_thetamax)))
    {
        // off electrons
        real const lo = fmax(cos_thetamin, helper.cos_thetamax_elec);
        real const hi = fmax(cos_thetamax, helper.cos_thetamax_elec);
        cos_theta = helper.sample_costheta(lo, hi, rng);
    }
    else
    {
        // off the nucleus, with rejection for false scattering
        cos_theta = helper.sample_costheta(cos_thetamin, cos_thetamax, rng);
        // Mott / Rutherford ratio: polynomial in (beta - 0.7181228) and sqrt(1 - cos)
        real const beta0 = sqrt(particle.beta_sq()) - real(0.7181228);
        real const fcos_t = sqrt(1 - cos_theta);
        u32 const base = (2 * element + (particle.charge < 0 ? 0u : 1u)) * 30;
        real theta_coeffs[5];
#pragma unroll
        for (int i = 0; i < 5; ++i)
        {
            real c[6];
#pragma unroll
            for (int j = 0; j < 6; ++j)
                c[j] = w.mott[base + 6 * i + j];
            theta_coeffs[i] = poly(c, beta0);
        }
        real const mott_ratio = poly(theta_coeffs, fcos_t);
        real const mt_sq = 2 * particle.momentum_sq() * (1 - cos_theta);
        real const xs = mott_ratio * ipow2(nuclear_form_factor(w, isotope, mt_sq));
        // RejectionSampler(xs, mott_factor): reject when xs < mott_factor * xi
        if (xs < helper.mott_factor * rng.canonical())
            cos_theta = 1;
    }

    Interaction result;
    result.direction = sample_exiting_direction(rng, cos_theta, inc_direction);
    // recoil energy: kinetic energy transferred to the atom
    real const projectile_mass = particle.mass + particle.energy;
    real const target_mass = w.isotope_nuclear_mass[isotope];
    real const recoil_energy = particle.momentum_sq() * (1 - cos_theta)
                               / (target_mass + projectile_mass * (1 - cos_theta));
    result.energy = particle.energy - recoil_energy;
    result.energy_deposition = recoil_energy;
    return result;
}

//---------------------------------------------------------------------------//
// Moller / Bhabha ionisation (em/interactor/MollerBhabhaInteractor.hh,
// em/distribution/{Moller,Bhabha}EnergyDistribution.hh,
// em/interactor/detail/IoniFinalStateHelper.hh)
//---------------------------------------------------------------------------//
B2_D real moller_g(real gamma, real epsilon)
{
    real const two_gamma_term = (2 * gamma - 1) / ipow2(gamma);
    real const complement_frac = 1 - epsilon;
    return 1 - two_gamma_term * epsilon
           + ipow2(epsilon)
                 * (1 - two_gamma_term
                    + (1 - two_gamma_term * complement_frac) / ipow2(complement_frac));
}

B2_D real bhabha_g(real gamma, real epsilon_min, real epsilon_max)
{
    real const y = 1 / (1 + gamma);
    real const y_sq = ipow2(y);
    real const one_minus_2y = 1 - 2 * y;
    real const b1 = 2 - y_sq;
    real const b2 = one_minus_2y * (3 + y_sq);
    real const b4 = one_minus_2y * one_minus_2y * one_minus_2y;
    real const b3 = ipow2(one_minus_2y) + b4;
    real const beta_sq = 1 - (1 / ipow2(gamma));
    real const emax2 = ipow2(epsilon_max);
    return 1
           + (emax2 * emax2 * b4 - (epsilon_min * epsilon_min * epsilon_min) * b3 + emax2 * b2
              - epsilon_min * b1)
                 * beta_sq;
}

B2_D Interaction interact_moller_bhabha(MollerBhabhaParams const& shared,
                                        Particle const& particle,
                                        real electron_cutoff,
                                        Real3 const& inc_direction,
                                        Rng& rng)
{
    real const inc_energy = particle.energy;
    real const inc_momentum = particle.momentum();
    bool const is_electron = (particle.id == shared.electron);
    real const min_frac = electron_cutoff / inc_energy;
    real const gamma = 1 + inc_energy / shared.electron_mass;
    real epsilon;
    if (is_electron)
    {
        real const max_frac = 0.5;
        real const g_denominator = moller_g(gamma, max_frac);
        real const a = 1 / max_frac, b = 1 / min_frac;
        do
        {
            epsilon = 1 / sample_uniform(rng, a, b);
        } while (reject(rng, moller_g(gamma, epsilon), g_denominator));
    }
    else
    {
        real const max_frac = 1;
        real const g_denominator = bhabha_g(gamma, min_frac, max_frac);
        real const a = 1 / max_frac, b = 1 / min_frac;
        do
        {
            epsilon = 1 / sample_uniform(rng, a, b);
        } while (reject(rng, bhabha_g(gamma, epsilon, epsilon), g_denominator));
    }
    real const electron_energy = inc_energy * epsilon;
    real const me = shared.electron_mass;
    // IoniFinalStateHelper
    real momentum = sqrt(electron_energy * (electron_energy + 2 * me));
    real costheta = electron_energy * (inc_energy + me + me) / (momentum * inc_momentum);
    Interaction result;
    result.num_secondaries = 1;
    result.sec[0].energy = electron_energy;
    result.sec[0].direction = sample_exiting_direction(rng, costheta, inc_direction);
    result.sec[0].particle = shared.electron;
    result.energy = inc_energy - electron_energy;
    result.direction
        = calc_exiting_direction(inc_momentum, inc_direction, momentum, result.sec[0].direction);
    return result;
}

//---------------------------------------------------------------------------//
// e+ annihilation to two gammas (em/interactor/EPlusGGInteractor.hh)
//---------------------------------------------------------------------------//
B2_D Interaction interact_eplusgg(EPlusGGParams const& shared,
                                  real inc_energy,
                                  Real3 const& inc_direction,
                                  Rng& rng)
{
    Interaction result = Interaction::from_absorption();
    result.num_secondaries = 2;
    result.sec[0].particle = shared.gamma;
    result.sec[1].particle = shared.gamma;
    if (inc_energy == 0)
    {
        result.sec[0].energy = shared.electron_mass;
        result.sec[1].energy = shared.electron_mass;
        result.sec[0].direction = sample_isotropic(rng);
        result.sec[1].direction = make_real3(-result.sec[0].direction[0],
                                             -result.sec[0].direction[1],
                                             -result.sec[0].direction[2]);
    }
    else
    {
        constexpr real half = 0.5;
        real const tau = inc_energy / shared.electron_mass;
        real const tau2 = tau + 2;
        real const sqgrate = sqrt(tau / tau2) * half;
        ReciprocalDist sample_eps(half - sqgrate, half + sqgrate);
        real epsil;
        do
        {
            epsil = sample_eps(rng);
        } while (sample_bernoulli(
            rng, epsil - (2 * (tau + 1) * epsil - 1) / (epsil * ipow2(tau2))));
        real const cost = (epsil * tau2 - 1) / (epsil * sqrt(tau * tau2));
        real const total_energy = inc_energy + 2 * shared.electron_mass;
        real const gamma_energy = epsil * total_energy;
        real const eplus_moment = sqrt(inc_energy * total_energy);
        result.sec[0].energy = gamma_energy;
        result.sec[0].direction = sample_exiting_direction(rng, cost, inc_direction);
        result.sec[1].energy = total_energy - gamma_energy;
        result.sec[1].direction
            = calc_exiting_direction(eplus_moment, inc_direction, inc_energy, inc_direction);
    }
    return result;
}

//! On-the-fly annihilation macro xs (em/xs/EPlusGGMacroXsCalculator.hh)
B2_D real calc_eplusgg_xs(ParamsView const& pv, u32 material, real energy)
{
    real const me = pv.model.epgg.electron_mass;
    real const e = energy > 1e-6 ? energy : 1e-6;
    real const gamma = e / me;
    real const sqrt_gg2 = sqrt(gamma * (gamma + 2));
    real const el_density = material_real(pv.mat, material, MAT_ELECTRON_DENSITY);
    real const re = pv.model.constants.r_electron;
    return constants::pi * ipow2(re) * el_density
           * (poly_quad(1, 4, 1, gamma + 1) * log(gamma + 1 + sqrt_gg2) - (gamma + 4) * sqrt_gg2)
           / (gamma * ipow2(gamma + 2));
}

//---------------------------------------------------------------------------//
// LPM functions (em/xs/LPMCalculator.hh)
//---------------------------------------------------------------------------//
struct LPMFunctions
{
    real xi, g, phi;
};

B2_D real lpm_phi(real s)
{
    if (s < real(0.01))
        return s * poly_lin(6, -6 * constants::pi, s);
    else if (s < real(1.55))
    {
        real a = poly_quad(0.623, 0.796, 0.658, s);
        real b = poly_quad(-6, -6 * (3 - constants::pi), 1 / a, s);
        return 1 - exp(s * b);
    }
    real s2 = ipow2(s);
    return 1 - real(0.01190476) / (s2 * s2);
}

B2_D real lpm_g(real s, real phi)
{
    if (s < real(0.01))
        return poly_lin(-2 * phi, 12, s);
    else if (s < real(0.415827))
    {
        real const c[5] = {1, 3.936, 4.97, -0.05, 7.5};
        real a = poly(c, s);
        real b = poly_lin(-4, -8 / a, s);
        real psi = 1 - exp(s * b);
        return 3 * psi - 2 * phi;
    }
    else if (s < real(1.9156))
    {
        real const c[5] = {-0.160723, 3.755030, -1.798138, 0.672827, -0.120772};
        return tanh(poly(c, s));
    }
    real s2 = ipow2(s);
    return 1 - real(0.0230655) / (s2 * s2);
}

B2_D LPMFunctions calc_lpm(ParamsView const& pv,
                           u32 material,
                           u32 element,
                           bool dielectric_suppression,
                           real gamma_energy,
                           real epsilon)
{
    real const electron_density = material_real(pv.mat, material, MAT_ELECTRON_DENSITY);
    real const lpm_energy
        = material_real(pv.mat, material, MAT_RAD_LENGTH) * pv.model.constants.lpm_constant;
    real const cbrt_z = element_real(pv.mat, element, EL_CBRT_Z);
    real const sqrt_two = 1.41421356237309504880;

    real const s_prime = sqrt(lpm_energy / (8 * epsilon * gamma_energy * fabs(epsilon - 1)));
    real const s1 = ipow2(cbrt_z / real(184.15));
    real xi = 2;
    if (s_prime > 1)
        xi = 1;
    else if (s_prime > sqrt_two * s1)
    {
        real const log_s1 = log(sqrt_two * s1);
        real const h = log(s_prime) / log_s1;
        xi = 1 + h - real(0.08) * (1 - h) * h * (2 - h) / log_s1;
    }
    real s = s_prime / sqrt(xi);
    if (dielectric_suppression)
    {
        real const k_p_sq
            = electron_density * pv.model.constants.migdal_constant * ipow2(epsilon * gamma_energy);
        s *= (1 + k_p_sq / ipow2(gamma_energy));
        xi = 2;
        if (s > 1)
            xi = 1;
        else if (s > s1)
            xi = 1 + log(s) / log(s1);
    }
    real phi = lpm_phi(s);
    if (xi * phi > 1 || s > real(0.57))
        xi = 1 / phi;
    LPMFunctions r;
    r.phi = phi;
    r.xi = xi;
    r.g = lpm_g(s, phi);
    return r;
}

//---------------------------------------------------------------------------//
// Bethe-Heitler pair production (em/interactor/BetheHeitlerInteractor.hh)
//---------------------------------------------------------------------------//
B2_D real bh_screening_f1(real delta)
{
    return delta > real(1.4) ? real(42.038) - real(8.29) * log(delta + real(0.958))
                             : real(42.184) - delta * (real(7.444) - real(1.623) * delta);
}
B2_D real bh_screening_f2(real delta)
{
    return delta > real(1.4) ? real(42.038) - real(8.29) * log(delta + real(0.958))
                             : real(41.326) - delta * (real(5.848) - real(0.902) * delta);
}

B2_D Interaction interact_bethe_heitler(ParamsView const& pv,
                                        real inc_energy,
                                        Real3 const& inc_direction,
                                        u32 material,
                                        u32 element,
                                        Rng& rng)
{
    BetheHeitlerParams const& shared = pv.model.bh;
    bool const enable_lpm = shared.enable_lpm && inc_energy > 1e5;
    real const epsilon0 = shared.electron_mass / inc_energy;
    real const cbrt_z = element_real(pv.mat, element, EL_CBRT_Z);
    constexpr real half = 0.5;
    real epsilon;
    if (inc_energy < 2)
    {
        epsilon = sample_uniform(rng, epsilon0, half);
    }
    else
    {
        real const delta_min = 4 * 136 / cbrt_z * epsilon0;
        real f_z = real(8) / real(3) * element_real(pv.mat, element, EL_LOG_Z);
        if (inc_energy > 50)
            f_z += 8 * element_real(pv.mat, element, EL_COULOMB);
        real const delta_max = exp((real(42.038) - f_z) / real(8.29)) - real(0.958);
        real const epsilon1 = half - half * sqrt(1 - delta_min / delta_max);
        real const epsilon_min = epsilon0 > epsilon1 ? epsilon0 : epsilon1;
        real const f10 = bh_screening_f1(delta_min) - f_z;
        real const f20 = bh_screening_f2(delta_min) - f_z;
        real const st = ipow2(half - epsilon_min) * f10;
        real const sf = real(1.5) * f20;
        real const p_f1g1 = st / (st + sf);
        real g;
        do
        {
            if (rng.canonical() < p_f1g1)
            {
                epsilon = half - (half - epsilon_min) * cbrt(rng.canonical());
                real delta = 136 / cbrt_z * epsilon0 / (epsilon * (1 - epsilon));
                if (enable_lpm)
                {
                    real phi1, phi2;
                    if (delta > real(1.4))
                    {
                        phi1 = real(21.0190) - real(4.145) * log(delta + real(0.958));
                        phi2 = phi1;
                    }
                    else
                    {
                        phi1 = real(20.806) - delta * (real(3.190) - real(0.5710) * delta);
                        phi2 = real(20.234) - delta * (real(2.126) - real(0.0903) * delta);
                    }
                    LPMFunctions lpm = calc_lpm(pv, material, element, false, inc_energy, epsilon);
                    g = lpm.xi * ((2 * lpm.phi + lpm.g) * phi1 - lpm.g * phi2 - lpm.phi * f_z)
                        / f10;
                }
                else
                {
                    g = (bh_screening_f1(delta) - f_z) / f10;
                }
            }
            else
            {
                epsilon = epsilon_min + (half - epsilon_min) * rng.canonical();
                real delta = 136 / cbrt_z * epsilon0 / (epsilon * (1 - epsilon));
                if (enable_lpm)
                {
                    real phi1, phi2;
                    if (delta > real(1.4))
                    {
                        phi1 = real(21.0190) - real(4.145) * log(delta + real(0.958));
                        phi2 = phi1;
                    }
                    else
                    {
                        phi1 = real(20.806) - delta * (real(3.190) - real(0.5710) * delta);
                        phi2 = real(20.234) - delta * (real(2.126) - real(0.0903) * delta);
                    }
                    LPMFunctions lpm = calc_lpm(pv, material, element, false, inc_energy, epsilon);
                    g = lpm.xi
                        * ((lpm.phi + half * lpm.g) * phi1 + half * lpm.g * phi2
                           - half * (lpm.g + lpm.phi) * f_z)
                        / f20;
                }
                else
                {
                    g = (bh_screening_f2(delta) - f_z) / f20;
                }
            }
        } while (g < rng.canonical());
    }
    Interaction result = Interaction::from_absorption();
    result.num_secondaries = 2;
    result.sec[0].particle = shared.electron;
    result.sec[1].particle = shared.positron;
    result.sec[0].energy = (1 - epsilon) * inc_energy - shared.electron_mass;
    result.sec[1].energy = epsilon * inc_energy - shared.electron_mass;
    if (sample_bernoulli(rng, half))
    {
        real t = result.sec[0].energy;
        result.sec[0].energy = result.sec[1].energy;
        result.sec[1].energy = t;
    }
    real phi = sample_uniform(rng, 0, 2 * constants::pi);
    real cost = sample_tsai_urban(rng, result.sec[0].energy, shared.electron_mass);
    result.sec[0].direction = rotate(from_spherical(cost, phi), inc_direction);
    cost = sample_tsai_urban(rng, result.sec[1].energy, shared.electron_mass);
    result.sec[1].direction = rotate(from_spherical(cost, phi + constants::pi), inc_direction);
    return result;
}

//---------------------------------------------------------------------------//
// Muon / hadron ionisation (em/interactor/MuHadIonizationInteractor.hh:103-146 with the
// delta-ray energy distributions em/distribution/{BraggICRU73QO,BetheBloch,MuBB}
// EnergyDistribution.hh and the maximum energy transfer of em/distribution/detail/Utils.hh)
//---------------------------------------------------------------------------//
enum MuHadSampler { MUHAD_BRAGG_ICRU73QO, MUHAD_BETHE_BLOCH, MUHAD_MU_BETHE_BLOCH };

B2_NOINLINE inline Interaction interact_muhad_ionization(MuHadIonizationParams const& shared,
                                           int sampler,
                                           Particle const& particle,
                                           real electron_cutoff,
                                           Real3 const& inc_direction,
                                           Rng& rng)
{
    real const inc_mass = particle.mass;
    real const me = shared.electron_mass;
    real const beta_sq = particle.beta_sq();
    // calc_max_secondary_energy
    real max_energy;
    {
        real const mass_ratio = me / inc_mass;
        real const tau = particle.energy / inc_mass;
        max_energy = 2 * me * tau * (tau + 2) / (1 + 2 * (tau + 1) * mass_ratio + ipow2(mass_ratio));
    }
    real min_energy = electron_cutoff;
    if (sampler == MUHAD_BRAGG_ICRU73QO)
    {
        // lowest kinetic energy scaled from the proton: ICRU73QO 5 keV, Bragg 0.25 keV
        real const lowest = particle.charge < 0 ? real(5e-3) : real(2.5e-4);
        min_energy = fmin(electron_cutoff, lowest * inc_mass / shared.proton_mass);
    }
    if (min_energy >= max_energy)
        return Interaction::from_unchanged();

    real const total_energy = particle.energy + inc_mass;
    bool const use_rad_correction = sampler == MUHAD_MU_BETHE_BLOCH && particle.energy > 250
                                    && max_energy > real(0.1);
    real envelope = 1;
    if (use_rad_correction)
        envelope = 1 + shared.alpha_over_twopi * ipow2(log(2 * total_energy / inc_mass));

    // InverseSquareDistribution(min, max) with the model's rejection function
    real const product = min_energy * max_energy;
    real energy;
    bool rejected;
    do
    {
        energy = product / sample_uniform(rng, min_energy, max_energy);
        real target = 1 - (beta_sq / max_energy) * energy;
        if (sampler == MUHAD_MU_BETHE_BLOCH)
        {
            target = target + real(0.5) * ipow2(energy / total_energy);
            if (use_rad_correction && energy > real(0.1))
            {
                real const a1 = log(1 + 2 * energy / me);
                real const a3 = log(4 * total_energy * (total_energy - energy) / ipow2(inc_mass));
                target *= (1 + shared.alpha_over_twopi * a1 * (a3 - a1));
            }
            rejected = target < envelope * rng.canonical();
        }
        else
        {
            rejected = target < rng.canonical();  // RejectionSampler(f): fmax = 1
        }
    } while (rejected);

    // IoniFinalStateHelper
    real const inc_momentum = particle.momentum();
    real const momentum = sqrt(energy * (energy + 2 * me));
    real const costheta = energy * (particle.energy + inc_mass + me) / (momentum * inc_momentum);
    Interaction result;
    result.num_secondaries = 1;
    result.sec[0].energy = energy;
    result.sec[0].direction = sample_exiting_direction(rng, costheta, inc_direction);
    result.sec[0].particle = shared.electron;
    result.energy = particle.energy - energy;
    result.direction
        = calc_exiting_direction(inc_momentum, inc_direction, momentum, result.sec[0].direction);
    return result;
}

B2_D Interaction brem_final_state(real inc_energy,
                                  Real3 const& inc_direction,
                                  real inc_momentum,
                                  u32 gamma_id,
                                  real gamma_energy,
                                  real costheta,
                                  Rng& rng);

//---------------------------------------------------------------------------//
// Muon bremsstrahlung (em/interactor/MuBremsstrahlungInteractor.hh:104-181,
// em/xs/MuBremsDiffXsCalculator.hh:108-200)
//---------------------------------------------------------------------------//
struct MuBremsDiffXs
{
    real atomic_number, atomic_mass, inv_cbrt_z, inc_energy, inc_mass, inc_mass_sq;
    real total_energy, electron_mass, d_n, b, b_prime, sqrt_euler, dcs_factor;

    B2_D MuBremsDiffXs(ParamsView const& pv, u32 element, Particle const& particle)
    {
        MuBremsstrahlungParams const& shared = pv.model.mubrems;
        u32 const z = pv.mat.element_z[element];
        atomic_number = real(z);
        atomic_mass = pv.mat.element_reals[EL_NUM_REALS * element + EL_MASS];
        inv_cbrt_z = 1 / pv.mat.element_reals[EL_NUM_REALS * element + EL_CBRT_Z];
        inc_energy = particle.energy;
        inc_mass = particle.mass;
        inc_mass_sq = ipow2(inc_mass);
        total_energy = inc_energy + inc_mass;
        electron_mass = shared.electron_mass;
        sqrt_euler = shared.sqrt_euler;
        dcs_factor = shared.dcs_factor;
        d_n = real(1.54) * pow(atomic_mass, real(0.27));
        if (z == 1)
        {
            b = real(202.4);
            b_prime = 446;
        }
        else
        {
            b = 183;
            b_prime = 1429;
            d_n = pow(d_n, 1 - real(1) / atomic_number);
        }
    }

    B2_D real operator()(real energy) const
    {
        if (energy >= inc_energy)
            return 0;
        real const v = energy / total_energy;
        real const delta = real(0.5) * inc_mass_sq * v / (total_energy - energy);
        real const phi_n = fmax(
            log(b * inv_cbrt_z * (inc_mass + delta * (d_n * sqrt_euler - 2))
                / (d_n * (electron_mass + delta * sqrt_euler * b * inv_cbrt_z))),
            real(0));
        real const energy_max_prime
            = total_energy / (1 + real(0.5) * inc_mass_sq / (electron_mass * total_energy));
        real phi_e = 0;
        if (energy < energy_max_prime)
        {
            real const inv_cbrt_z_sq = ipow2(inv_cbrt_z);
            phi_e = fmax(
                log(b_prime * inv_cbrt_z_sq * inc_mass
                    / ((1 + delta * inc_mass / (ipow2(electron_mass) * sqrt_euler))
                       * (electron_mass + delta * sqrt_euler * b_prime * inv_cbrt_z_sq))),
                real(0));
        }
        return dcs_factor * atomic_number * (atomic_number * phi_n + phi_e)
               * (1 - v * (1 - real(0.75) * v)) / (3 * inc_mass_sq * energy * atomic_mass);
    }
};

B2_NOINLINE inline Interaction interact_mu_bremsstrahlung(ParamsView const& pv,
                                            Particle const& particle,
                                            Real3 const& inc_direction,
                                            u32 material,
                                            u32 element,
                                            Rng& rng)
{
    MuBremsstrahlungParams const& shared = pv.model.mubrems;
    MuBremsDiffXs const calc_dcs(pv, element, particle);
    real const gamma_cutoff = cutoff_energy(pv, material, shared.gamma);
    ReciprocalDist const sample_energy(gamma_cutoff, particle.energy);
    real const envelope = gamma_cutoff * calc_dcs(gamma_cutoff);

    real gamma_energy;
    do
    {
        gamma_energy = sample_energy(rng);
    } while (gamma_energy * calc_dcs(gamma_energy) < envelope * rng.canonical());

    // sample_cos_theta
    real const gamma = particle.lorentz_factor();
    real const r_max_sq = ipow2(gamma * constants::pi * real(0.5)
                                * fmin(real(1), gamma * particle.mass / gamma_energy - 1));
    real const a = rng.canonical() * r_max_sq / (1 + r_max_sq);
    real const costheta = cos(sqrt(a / (1 - a)) / gamma);
    return brem_final_state(particle.energy,
                            inc_direction,
                            particle.momentum(),
                            shared.gamma,
                            gamma_energy,
                            costheta,
                            rng);
}

//---------------------------------------------------------------------------//
// Bremsstrahlung final state (em/interactor/detail/BremFinalStateHelper.hh)
//---------------------------------------------------------------------------//
B2_D Interaction brem_final_state(real inc_energy,
                                  Real3 const& inc_direction,
                                  real inc_momentum,
                                  u32 gamma_id,
                                  real gamma_energy,
                                  real costheta,
                                  Rng& rng)
{
    Interaction result;
    result.num_secondaries = 1;
    result.sec[0].direction = sample_exiting_direction(rng, costheta, inc_direction);
    result.sec[0].particle = gamma_id;
    result.sec[0].energy = gamma_energy;
    result.energy = inc_energy - gamma_energy;
    result.direction = calc_exiting_direction(
        inc_momentum, inc_direction, gamma_energy, result.sec[0].direction);
    return result;
}

//---------------------------------------------------------------------------//
// Seltzer-Berger bremsstrahlung (em/interactor/SeltzerBergerInteractor.hh,
// em/interactor/detail/SBEnergySampler.hh, em/distribution/SBEnergyDistHelper.hh)
//---------------------------------------------------------------------------//
struct SBGrid
{
    real const* x;
    u32 nx;
    real const* y;
    u32 ny;
    real const* values;
    u32 const* argmax;
};

//! Lower bin of `v` in a sorted array (NonuniformGrid::find)
B2_D u32 nonuniform_find(real const* grid, u32 n, real v)
{
    u32 lo = 0, len = n;
    while (len > 0)
    {
        u32 half = len >> 1;
        u32 mid = lo + half;
        if (grid[mid] < v)
        {
            lo = mid + 1;
            len -= half + 1;
        }
        else
            len = half;
    }
    if (v != grid[lo])
        --lo;
    return lo;
}

B2_D Interaction interact_seltzer_berger(ParamsView const& pv,
                                         Particle const& particle,
                                         Real3 const& inc_direction,
                                         u32 material,
                                         u32 element,
                                         Rng& rng)
{
    SeltzerBergerParams const& shared = pv.model.sb;
    real const inc_energy = particle.energy;
    real const gamma_cutoff = cutoff_energy(pv, material, shared.gamma);
    bool const is_electron = particle.id == shared.electron;
    real const density_factor
        = material_real(pv.mat, material, MAT_ELECTRON_DENSITY) * pv.model.constants.migdal_constant;
    real const dens_corr = density_factor * ipow2(particle.total_energy());

    u32 const* row = shared.elements + 8 * element;
    SBGrid g;
    g.x = shared.reals + row[0];
    g.nx = row[1];
    g.y = shared.reals + row[2];
    g.ny = row[3];
    g.values = shared.reals + row[4];
    g.argmax = shared.sizes + row[5];

    // x location: log(E)
    real const logx = log(inc_energy);
    u32 const ix = nonuniform_find(g.x, g.nx, logx);
    real const xfrac = (logx - g.x[ix]) / (g.x[ix + 1] - g.x[ix]);
    // max xs at this energy from the per-row argmax
    real const max_xs = (1 - xfrac) * g.values[ix * g.ny + g.argmax[ix]]
                        + xfrac * g.values[(ix + 1) * g.ny + g.argmax[ix + 1]];
    real const inv_inc_energy = 1 / inc_energy;
    ReciprocalDist sample_esq(ipow2(gamma_cutoff) + dens_corr, ipow2(inc_energy) + dens_corr);

    // positron correction (SBPositronXsCorrector)
    real const pmass = particle.mass;
    real const alpha_z = 2 * constants::pi * pv.model.constants.alpha_fine_structure
                         * pv.mat.element_z[element];
    auto calc_invbeta = [&](real gamma_energy) {
        real e = inc_energy - gamma_energy;
        return (e + pmass) / sqrt(e * (e + 2 * pmass));
    };
    real const cutoff_invbeta = is_electron ? 0 : calc_invbeta(gamma_cutoff);

    real exit_energy, xs;
    do
    {
        exit_energy = sqrt(sample_esq(rng) - dens_corr);
        // bilinear interpolation in (log E, k/E)
        real const yv = exit_energy * inv_inc_energy;
        u32 const iy = nonuniform_find(g.y, g.ny, yv);
        real const yfrac = (yv - g.y[iy]) / (g.y[iy + 1] - g.y[iy]);
        real const v00 = g.values[ix * g.ny + iy];
        real const v01 = g.values[ix * g.ny + iy + 1];
        real const v10 = g.values[(ix + 1) * g.ny + iy];
        real const v11 = g.values[(ix + 1) * g.ny + iy + 1];
        xs = (1 - xfrac) * ((1 - yfrac) * v00 + (yfrac)*v01)
             + (xfrac) * ((1 - yfrac) * v10 + (yfrac)*v11);
        if (!is_electron)
        {
            real delta = cutoff_invbeta - calc_invbeta(exit_energy);
            xs = xs * exp(alpha_z * (delta < 0 ? delta : real(0)));
        }
    } while (reject(rng, xs, max_xs));

    real const costheta = sample_tsai_urban(rng, inc_energy, particle.mass);
    return brem_final_state(
        inc_energy, inc_direction, particle.momentum(), shared.gamma, exit_energy, costheta, rng);
}

//---------------------------------------------------------------------------//
// Relativistic bremsstrahlung (em/interactor/RelativisticBremInteractor.hh,
// em/xs/RBDiffXsCalculator.hh, em/interactor/detail/RBEnergySampler.hh)
//---------------------------------------------------------------------------//
B2_D Interaction interact_relativistic_brem(ParamsView const& pv,
                                            Particle const& particle,
                                            Real3 const& inc_direction,
                                            u32 material,
                                            u32 element,
                                            Rng& rng)
{
    RelativisticBremParams const& shared = pv.model.rb;
    real const* ed = shared.elem_data + 5 * element;
    real const fz = ed[0], factor1 = ed[1], factor2 = ed[2], gamma_factor = ed[3],
               epsilon_factor = ed[4];
    real const total_energy = particle.total_energy();
    real const density_factor
        = material_real(pv.mat, material, MAT_ELECTRON_DENSITY) * pv.model.constants.migdal_constant;
    real const density_corr = density_factor * ipow2(total_energy);
    real const lpm_energy
        = material_real(pv.mat, material, MAT_RAD_LENGTH) * pv.model.constants.lpm_constant;
    real const lpm_threshold = lpm_energy * sqrt(density_factor);
    bool const enable_lpm = shared.enable_lpm && (total_energy > lpm_threshold);
    u32 const z = pv.mat.element_z[element];
    real const log_z = element_real(pv.mat, element, EL_LOG_Z);

    real const gcut = cutoff_energy(pv, material, shared.gamma);
    real const tmin = gcut < particle.energy ? gcut : particle.energy;
    real const tmax = 1e8 < particle.energy ? real(1e8) : particle.energy;
    ReciprocalDist sample_esq(ipow2(tmin) + density_corr, ipow2(tmax) + density_corr);
    real const max_value = factor1 + factor2;

    real gamma_energy, dsigma;
    do
    {
        gamma_energy = sqrt(sample_esq(rng) - density_corr);
        if (enable_lpm)
        {
            real epsilon = total_energy / gamma_energy;
            LPMFunctions lpm = calc_lpm(pv, material, element, true, gamma_energy, epsilon);
            real y = gamma_energy / total_energy;
            real onemy = 1 - y;
            real y2 = real(0.25) * ipow2(y);
            real term = lpm.xi * (y2 * lpm.g + (onemy + 2 * y2) * lpm.phi);
            dsigma = term * factor1 + onemy * factor2;
        }
        else
        {
            real y = gamma_energy / total_energy;
            real onemy = 1 - y;
            real term0 = onemy + real(0.75) * ipow2(y);
            if (z < 5)
            {
                dsigma = term0 * factor1 + onemy * factor2;
            }
            else
            {
                real invz = 1 / static_cast<real>(z);
                real term1 = y / (total_energy - gamma_energy);
                real gam = term1 * gamma_factor;
                real eps = term1 * epsilon_factor;
                real gam2 = ipow2(gam), eps2 = ipow2(eps);
                real phi1 = real(16.863) - 2 * log(1 + real(0.311877) * gam2)
                            + real(2.4) * exp(real(-0.9) * gam) + real(1.6) * exp(real(-1.5) * gam);
                real phi2 = 2 / (3 + real(19.5) * gam + 18 * gam2);
                real psi1 = real(24.34) - 2 * log(1 + real(13.111641) * eps2)
                            + real(2.8) * exp(real(-8) * eps) + real(1.2) * exp(real(-29.2) * eps);
                real psi2 = 2 / (3 + 120 * eps + 1200 * eps2);
                dsigma = term0 * ((real(0.25) * phi1 - fz) + (real(0.25) * psi1 - 2 * log_z / 3) * invz)
                         + real(0.125) * onemy * (phi2 + psi2 * invz);
            }
        }
        dsigma = dsigma > 0 ? dsigma : real(0);
    } while (reject(rng, dsigma, max_value));

    real const costheta = sample_tsai_urban(rng, particle.energy, particle.mass);
    return brem_final_state(particle.energy,
                            inc_direction,
                            particle.momentum(),
                            shared.gamma,
                            gamma_energy,
                            costheta,
                            rng);
}

//---------------------------------------------------------------------------//
// Livermore photoelectric effect (em/interactor/LivermorePEInteractor.hh,
// em/xs/LivermorePEMicroXsCalculator.hh)
//---------------------------------------------------------------------------//
//! Linear interpolation on a nonuniform grid (grid/GenericCalculator.hh)
B2_D real generic_calc(real const* x, real const* y, u32 n, real v)
{
    if (v <= x[0])
        return y[0];
    if (v >= x[n - 1])
        return y[n - 1];
    u32 i = nonuniform_find(x, n, v);
    return lerp_points(x[i], y[i], x[i + 1], y[i + 1], v);
}

struct PEShell
{
    real binding_energy;
    real const* param[2];
    real const* xs_grid;
    real const* xs_value;
    u32 xs_size;
};

B2_D PEShell pe_shell(LivermorePEParams const& pe, u32 shell)
{
    PEShell s;
    real const* r = pe.shell_reals + 13 * shell;
    s.binding_energy = r[0];
    s.param[0] = r + 1;
    s.param[1] = r + 7;
    u32 const* u = pe.shells + 4 * shell;
    s.xs_grid = pe.reals + u[0];
    s.xs_size = u[1];
    s.xs_value = pe.reals + u[2];
    return s;
}

B2_D real poly5(real const* c, real x)
{
    real r = c[5];
#pragma unroll
    for (int i = 4; i >= 0; --i)
        r = fma(x, r, c[i]);
    return r;
}

//! Microscopic xs [b] for one element
B2_D real calc_livermore_micro_xs(LivermorePEParams const& pe, u32 element, real inc_energy)
{
    u32 const* el = pe.elements + 8 * element;
    real const thresh_lo = pe.element_thresh[2 * element];
    real const thresh_hi = pe.element_thresh[2 * element + 1];
    u32 const shell_begin = el[6], shell_count = el[7];
    PEShell back = pe_shell(pe, shell_begin + shell_count - 1);
    real energy = inc_energy > back.binding_energy ? inc_energy : back.binding_energy;
    real inv_energy = 1. / energy;
    real result = 0;
    if (energy >= thresh_lo)
    {
        real const* param = back.param[energy < thresh_hi ? 0 : 1];
        result = inv_energy * poly5(param, inv_energy);
    }
    else
    {
        PEShell front = pe_shell(pe, shell_begin);
        real inv3 = inv_energy * inv_energy * inv_energy;
        if (energy >= front.binding_energy)
            result = inv3 * generic_calc(pe.reals + el[3], pe.reals + el[5], el[4], energy);
        else
            result = inv3 * generic_calc(pe.reals + el[0], pe.reals + el[2], el[1], energy);
    }
    return result;
}

//! Macroscopic xs from on-the-fly micro xs (phys/MacroXsCalculator.hh)
B2_D real calc_livermore_macro_xs(ParamsView const& pv, u32 material, real energy)
{
    MatParams const& m = pv.mat;
    real result = 0;
    for (u32 i = m.material_elcomp_begin[material]; i < m.material_elcomp_end[material]; ++i)
    {
        real micro = calc_livermore_micro_xs(pv.model.pe, m.elcomp_element[i], energy);
        result += micro * m.elcomp_fraction[i];
    }
    // barn -> native (cm^2)
    return (result * 1e-24) * material_real(m, material, MAT_NUMBER_DENSITY);
}

B2_D Interaction interact_livermore_pe(ParamsView const& pv,
                                       real inc_energy,
                                       Real3 const& inc_direction,
                                       u32 element,
                                       Rng& rng)
{
    LivermorePEParams const& pe = pv.model.pe;
    u32 const* el = pe.elements + 8 * element;
    real const thresh_lo = pe.element_thresh[2 * element];
    real const thresh_hi = pe.element_thresh[2 * element + 1];
    u32 const shell_begin = el[6], shell_count = el[7];
    real const inv_energy = 1 / inc_energy;

    // sample subshell
    u32 shell_id = 0;
    real const cutoff = rng.canonical() * calc_livermore_micro_xs(pe, element, inc_energy);
    bool no_shell = false;
    if (inc_energy < thresh_lo)
    {
        real xs = 0;
        real const inv_cube = inv_energy * inv_energy * inv_energy;
        for (; shell_id < shell_count; ++shell_id)
        {
            PEShell sh = pe_shell(pe, shell_begin + shell_id);
            if (inc_energy < sh.binding_energy)
                continue;
            xs += inv_cube * generic_calc(sh.xs_grid, sh.xs_value, sh.xs_size, inc_energy);
            if (xs > cutoff)
                break;
        }
        if (shell_id == shell_count)
            no_shell = true;
    }
    else
    {
        int const pidx = inc_energy < thresh_hi ? 0 : 1;
        u32 const shell_end = shell_count - 1;
        for (; shell_id < shell_end; ++shell_id)
        {
            PEShell sh = pe_shell(pe, shell_begin + shell_id);
            real xs = inv_energy * poly5(sh.param[pidx], inv_energy);
            if (xs > cutoff)
                break;
        }
    }
    if (no_shell)
    {
        Interaction result = Interaction::from_absorption();
        result.energy_deposition = inc_energy;
        return result;
    }
    real const binding_energy = pe_shell(pe, shell_begin + shell_id).binding_energy;

    // photoelectron direction (Sauter-Gavrila)
    Real3 direction;
    if (inc_energy > 100.)
    {
        direction = inc_direction;
    }
    else
    {
        real e = inc_energy > 1.e-6 ? inc_energy : real(1.e-6);
        real energy_per_mecsq = e * pe.inv_electron_mass;
        real gamma = energy_per_mecsq + 1;
        real beta = sqrt(energy_per_mecsq * (gamma + 1)) / gamma;
        real a = (1 - beta) / beta;
        constexpr real half = 0.5;
        real b = half * beta * gamma * energy_per_mecsq * (gamma - 2);
        real g_max = 2 * (1 / a + b);
        real g, nu;
        do
        {
            real u = rng.canonical();
            nu = 2 * a * (2 * u + (a + 2) * sqrt(u)) / (ipow2(a + 2) - 4 * u);
            g = (2 - nu) * (1 / (a + nu) + b);
        } while (g < g_max * rng.canonical());
        direction = sample_exiting_direction(rng, 1 - nu, inc_direction);
    }
    Interaction result = Interaction::from_absorption();
    result.num_secondaries = 1;
    result.sec[0].particle = pe.electron;
    result.sec[0].energy = inc_energy - binding_energy;
    result.sec[0].direction = direction;
    result.energy_deposition = binding_energy;
    return result;
}

//---------------------------------------------------------------------------//
// On-the-fly macroscopic cross sections for "hardwired" models
// (PhysicsTrackView::calc_xs, phys/PhysicsTrackView.hh)
//---------------------------------------------------------------------------//
B2_D real calc_hardwired_xs(ParamsView const& pv, u32 model, u32 material, real energy)
{
    if (model == pv.phys.hw_livermore_pe)
        return calc_livermore_macro_xs(pv, material, energy);
    if (model == pv.phys.hw_eplusgg)
        return calc_eplusgg_xs(pv, material, energy);
    return 0;
}

//! Dispatch the interaction for a model action (the reference's *Executor.hh)
//! EXTRA: with the interactors beyond the north star's list (Rayleigh, Coulomb, muons) and
//! the combined bremsstrahlung model, which no benchmark problem has. The
//! fused step and the device-resident loop are built without them: carried along (even as
//! out-of-line calls) they cost the TestEm3 pass 1.1 % (k_step_fused +4 %, measured).
template<bool EXTRA>
B2_D void run_interaction(ParamsView const& pv, StateView const& s, u32 slot, u32 action, Rng& rng)
{
    ModelParams const& m = pv.model;
    GeoTrack geo(pv, s, slot);
    Real3 const dir = geo.dir();
    Particle particle = load_particle(pv, s, slot);
    u32 const material = s.material_id[slot];
    MatParams const& mat = pv.mat;
    auto element_of = [&](u32 elcomp) {
        return mat.elcomp_element[mat.material_elcomp_begin[material] + elcomp];
    };

    // a model action without an interactor here is a load-time error (CoreParams::load);
    // should one get through, the track fails loudly instead of silently not interacting
    Interaction result;
    result.action = IA_FAILED;
    if (action == m.kn.action)
    {
        result = interact_klein_nishina(m.kn, particle.energy, dir, rng);
    }
    else if (action == m.mb.action)
    {
        result = interact_moller_bhabha(
            m.mb, particle, cutoff_energy(pv, material, m.mb.electron), dir, rng);
    }
    else if (action == m.epgg.action)
    {
        result = interact_eplusgg(m.epgg, particle.energy, dir, rng);
    }
    else if (action == m.bh.action)
    {
        result = interact_bethe_heitler(
            pv, particle.energy, dir, material, element_of(s.element[slot]), rng);
    }
    else if (action == m.sb.action)
    {
        result = interact_seltzer_berger(pv, particle, dir, material, element_of(s.element[slot]), rng);
    }
    else if (action == m.rb.action)
    {
        result = interact_relativistic_brem(
            pv, particle, dir, material, element_of(s.element[slot]), rng);
    }
    else if (EXTRA && action == m.cb.action)
    {
        // (EXTRA: a second inlined copy of both bremsstrahlung samplers, used only by
        // problems built with celer-sim's `brem_combined`)
        // em/interactor/CombinedBremInteractor.hh:132-170: relativistic sampler at and above
        // 1 GeV, Seltzer-Berger below; same angular distribution and final state either way
        // The combined model has no element selector: its executor always interacts with the
        // material's first element (em/executor/CombinedBremExecutor.hh:42-44)
        u32 const element = element_of(0);
        if (particle.energy >= m.cb.sb_upper_limit)
            result = interact_relativistic_brem(pv, particle, dir, material, element, rng);
        else
            result = interact_seltzer_berger(pv, particle, dir, material, element, rng);
    }
    else if (EXTRA
             && (action == m.muioni.bragg_action || action == m.muioni.icru73qo_action
                 || action == m.muioni.bethe_bloch_action
                 || action == m.muioni.mu_bethe_bloch_action))
    {
        int const sampler = action == m.muioni.mu_bethe_bloch_action ? MUHAD_MU_BETHE_BLOCH
                            : action == m.muioni.bethe_bloch_action  ? MUHAD_BETHE_BLOCH
                                                                      : MUHAD_BRAGG_ICRU73QO;
        result = interact_muhad_ionization(
            m.muioni, sampler, particle, cutoff_energy(pv, material, m.muioni.electron), dir, rng);
    }
    else if (EXTRA && action == m.mubrems.action)
    {
        result = interact_mu_bremsstrahlung(
            pv, particle, dir, material, element_of(s.element[slot]), rng);
    }
    else if (EXTRA && action == m.coulomb.action)
    {
        result = interact_coulomb(pv, particle, dir, material, element_of(s.element[slot]), rng);
    }
    else if (EXTRA && action == m.rayleigh.action)
    {
        result = interact_rayleigh(
            m.rayleigh, particle.energy, dir, element_of(s.element[slot]), rng);
    }
    else if (action == m.pe.action)
    {
        u32 elcomp = s.element[slot];
        if (elcomp == INVALID)
        {
            // Select the element on the fly from micro xs (mat/ElementSelector.hh)
            u32 const b = mat.material_elcomp_begin[material];
            u32 const ne = mat.material_elcomp_end[material] - b;
            real material_xs = 0;
            for (u32 i = 0; i < ne; ++i)
            {
                real micro = calc_livermore_micro_xs(m.pe, mat.elcomp_element[b + i], particle.energy);
                material_xs += micro * mat.elcomp_fraction[b + i];
            }
            real accum = -material_xs * rng.canonical();
            u32 i = 0;
            for (; i != ne - 1; ++i)
            {
                real micro = calc_livermore_micro_xs(m.pe, mat.elcomp_element[b + i], particle.energy);
                accum += mat.elcomp_fraction[b + i] * micro;
                if (accum > 0)
                    break;
            }
            elcomp = i;
            s.element[slot] = elcomp;
        }
        result = interact_livermore_pe(pv, particle.energy, dir, element_of(elcomp), rng);
    }
    apply_interaction(pv, s, slot, result);
}
}  // namespace b200
