//---------------------------------------------------------------------------//
// Along-step: Urban multiple scattering, propagation, continuous energy loss
// with fluctuations, time and mean-free-path bookkeeping for one track.
//
// Order of operations and of RNG draws per slot follows the reference's
// AlongStep functor (/root/reference/src/celeritas/global/alongstep/AlongStep.hh:50-58):
// msc step limit -> propagate -> msc scatter -> time -> energy loss -> track
// update.
//---------------------------------------------------------------------------//
#pragma once

#include "field.cuh"
#include "interact.cuh"
#include "orange.cuh"
#include "physics.cuh"

namespace b200
{
//---------------------------------------------------------------------------//
// Normal distribution with the reference's spare-value behaviour
// (random/distribution/NormalDistribution.hh)
//---------------------------------------------------------------------------//
struct NormalDist
{
    real mean, stddev, spare;
    bool has_spare;
    B2_D NormalDist(real m, real s) : mean(m), stddev(s), spare(0), has_spare(false) {}
    B2_D real operator()(Rng& rng)
    {
        if (has_spare)
        {
            has_spare = false;
            return fma(spare, stddev, mean);
        }
        constexpr real twopi = 2 * constants::pi;
        real theta = twopi * rng.canonical();
        real r = sqrt(-2 * log(rng.canonical()));
        spare = r * cos(theta);
        has_spare = true;
        return fma(r * sin(theta), stddev, mean);
    }
};

//! Poisson sampling (random/distribution/PoissonDistribution.hh)
B2_D u32 sample_poisson(Rng& rng, real lambda)
{
    if (lambda <= 16)
    {
        int k = 0;
        real p = exp(lambda);
        do
        {
            ++k;
            p *= rng.canonical();
        } while (p > 1);
        return static_cast<u32>(k - 1);
    }
    NormalDist sample_normal(lambda, sqrt(lambda));
    return static_cast<u32>(sample_normal(rng) + real(0.5));
}

B2_D real fastpow(real a, real b)
{
    return exp(b * log(a));
}

//---------------------------------------------------------------------------//
// URBAN MSC (em/msc/UrbanMsc.hh:75-298 and em/msc/detail/*.hh)
//---------------------------------------------------------------------------//
struct MscHelper
{
    ParamsView const& pv;
    UrbanMscParams const& msc;
    PhysTrack const& phys;
    Particle const& particle;
    u32 idx;      // (material, particle) entry
    real lambda;  // msc mean free path at the pre-step energy
    real range;   // dedx range

    B2_D MscHelper(ParamsView const& p, PhysTrack const& ph, Particle const& pa, real dedx_range)
        : pv(p), msc(p.model.msc), phys(ph), particle(pa), range(dedx_range)
    {
        idx = phys.material * 2 + (particle.id == msc.electron ? 0 : 1);
        lambda = calc_msc_mfp(particle.energy);
    }

    //! 1 / (scaled xs / E^2) (UrbanMscHelper::calc_msc_mfp)
    B2_D real calc_msc_mfp(real energy) const
    {
        u32 const* gu = msc.xs_grid_u32 + 3 * idx;
        real const* gf = msc.xs_grid_f64 + 3 * idx;
        u32 const n = gu[0], prime = gu[1];
        real const* values = msc.reals + gu[2];
        real const front = gf[0], back = gf[1], delta = gf[2];
        real const loge = log(energy);
        real xs;
        if (loge <= front)
        {
            xs = values[0];
            if (0 >= prime)
                xs /= energy;
        }
        else if (loge >= back)
        {
            xs = values[n - 1];
            if (n - 1 >= prime)
                xs /= energy;
        }
        else
        {
            u32 const lower = static_cast<u32>((loge - front) / delta);
            real const* node_energy = msc.grid_energy + msc.xs_grid_energy_offset[idx];
            real const upper_energy = node_energy[lower + 1];
            real upper_xs = values[lower + 1];
            if (lower + 1 == prime)
                upper_xs /= upper_energy;
            xs = lerp_points(node_energy[lower], values[lower], upper_energy, upper_xs, energy);
            if (lower >= prime)
                xs /= energy;
        }
        real xsec = xs / ipow2(energy);
        return 1 / xsec;
    }

    B2_D real scaled_zeff() const { return msc.par_mat_data[2 * idx]; }
    B2_D real max_step() const { return range * msc.par_mat_data[2 * idx + 1]; }

    B2_D real calc_inverse_range(real step) const
    {
        u32 grid = phys.value_grid(VGT_RANGE, phys.eloss_ppid());
        return b200::calc_inverse_range(pv.phys, grid, step);
    }

    //! Energy at the end of a step (UrbanMscHelper::calc_end_energy)
    B2_D real calc_end_energy(real step) const
    {
        if (step <= range * real(0.05))
        {
            u32 grid = phys.value_grid(VGT_ENERGY_LOSS, phys.eloss_ppid());
            real dedx = calc_xs(pv.phys, grid, particle.energy);
            return particle.energy - step * dedx;
        }
        return calc_inverse_range(range - step);
    }
};

B2_D bool msc_is_applicable(ParamsView const& p, StateView const& s, u32 slot, Particle const& particle, real step)
{
    UrbanMscParams const& msc = p.model.msc;
    if (!msc.enabled)
        return false;
    if (step <= msc.geom_limit)
        return false;
    if (s.status[slot] != ST_ALIVE)
        return false;
    if (particle.id != msc.electron && particle.id != msc.positron)
        return false;
    return particle.energy > msc.low_energy_limit && particle.energy < msc.high_energy_limit;
}

struct GeoPath
{
    real step;
    real alpha;
};

//! True path -> geometric path (em/msc/detail/MscStepToGeo.hh)
B2_D GeoPath msc_step_to_geo(MscHelper const& h, real tstep)
{
    UrbanMscParams const& msc = h.msc;
    GeoPath result;
    result.alpha = 0;
    real const min_step = 1e-7;  // 1 nm
    if (tstep < min_step)
    {
        result.step = tstep;
    }
    else if (tstep < h.range * real(0.05))
    {
        result.step = -h.lambda * expm1(-tstep / h.lambda);
    }
    else
    {
        real mfp_slope;
        if (h.particle.energy < msc.electron_mass || tstep == h.range)
        {
            result.alpha = 1 / h.range;
            real v = 1 - result.alpha * tstep;
            mfp_slope = v > 0 ? v : real(0);
        }
        else
        {
            real rfinal = h.range - tstep;
            real endpoint_energy = h.calc_inverse_range(rfinal);
            real lambda1 = h.calc_msc_mfp(endpoint_energy);
            result.alpha = (h.lambda - lambda1) / (h.lambda * tstep);
            mfp_slope = lambda1 / h.lambda;
        }
        real w = 1 + 1 / (result.alpha * h.lambda);
        result.step = (1 - fastpow(mfp_slope, w)) / (result.alpha * w);
    }
    result.step = result.step < tstep ? result.step : tstep;
    return result;
}

//! Geometric path -> true path (em/msc/detail/MscStepFromGeo.hh)
B2_D real msc_step_from_geo(real gstep, real true_step, real alpha, real range, real lambda)
{
    real const min_step = 1e-7;
    if (gstep < min_step)
        return gstep;
    real tstep;
    if (alpha == 0)
    {
        tstep = -lambda * log1p(-gstep / lambda);
        if (tstep < min_step)
            tstep = gstep;
    }
    else
    {
        real w = 1 + 1 / (alpha * lambda);
        real x = alpha * w * gstep;
        x = x < 1 ? x : real(1);
        real temp = 1 - fastpow(1 - x, 1 / w);
        real result = temp / alpha;
        tstep = result < range ? result : range;
    }
    // clamp(tstep, gstep, true_step)
    return tstep < gstep ? gstep : (true_step < tstep ? true_step : tstep);
}

//! Urban msc step limitation (UrbanMsc::limit_step, UrbanMscSafetyStepLimit.hh,
//! UrbanMscMinimalStepLimit.hh)
template<class Geo>
B2_D void msc_limit_step(ParamsView const& p,
                         StateView const& s,
                         u32 slot,
                         Particle const& particle,
                         PhysTrack const& phys,
                         Geo& geo)
{
    UrbanMscParams const& msc = p.model.msc;
    PhysParams const& pp = p.phys;
    u32 const n = s.num_slots;
    real const phys_step = s.step_length[slot];
    real const range = s.dedx_range[slot];
    MscHelper helper(p, phys, particle, range);
    bool displaced = false;
    real true_path;
    real const limit_min_fix = 1e-9;  // 0.01 nm
    do
    {
        if (phys_step <= limit_min_fix)
        {
            true_path = phys_step;
            break;
        }
        real safety = 0;
        bool const on_boundary = geo.is_on_boundary();
        if (!on_boundary)
        {
            real const max_step = helper.max_step();
            safety = geo.find_safety();
            if (safety >= max_step)
            {
                true_path = phys_step;
                break;
            }
        }
        displaced = true;
        Rng rng;
        rng.load(s, slot);
        bool const use_safety_plus = (pp.step_limit_algorithm == 2);
        bool const minimal = (pp.step_limit_algorithm == 0);
        real range_init = s.msc_range[slot];
        real range_factor = s.msc_range[n + slot];
        real limit_min = s.msc_range[2 * n + slot];
        bool const has_range = range_init > 0 && range_factor > 0 && limit_min > 0;
        if (minimal)
        {
            // UrbanMscMinimalStepLimit
            if (!has_range)
            {
                range_init = real_inf();
                range_factor = pp.range_factor;
                limit_min = 10 * limit_min_fix;
            }
            if (on_boundary)
            {
                real m = range > helper.lambda ? range : helper.lambda;
                range_init = range_factor * m;
                range_init = range_init > limit_min ? range_init : limit_min;
            }
            s.msc_range[slot] = range_init;
            s.msc_range[n + slot] = range_factor;
            s.msc_range[2 * n + slot] = limit_min;
            real limit = range_init;
            if (phys_step <= limit)
            {
                true_path = phys_step;
            }
            else if (limit == limit_min)
            {
                true_path = limit_min;
            }
            else
            {
                NormalDist sample_gauss(limit, real(0.1) * (limit - limit_min));
                real sampled = sample_gauss(rng);
                true_path = sampled < limit_min ? limit_min
                                                : (phys_step < sampled ? phys_step : sampled);
            }
        }
        else
        {
            // UrbanMscSafetyStepLimit
            real max_step = phys_step;
            if (!has_range || on_boundary)
            {
                range_factor = pp.range_factor;
                range_init = use_safety_plus ? range
                                             : (range > helper.lambda ? range : helper.lambda);
                if (helper.lambda > pp.lambda_limit)
                {
                    real c = use_safety_plus ? 0.84 : 0.75;
                    range_factor *= c + (1 - c) * helper.lambda / pp.lambda_limit;
                }
                real const* md = msc.material_data + 8 * phys.material;
                real xm = helper.lambda / poly_quad(2, md[0], md[1], particle.energy);
                xm *= helper.scaled_zeff();
                if (particle.energy < 5e-3)
                    xm *= (real(0.5) + real(0.5) * particle.energy / real(5e-3));
                limit_min = xm > limit_min_fix ? xm : limit_min_fix;
                s.msc_range[slot] = range_init;
                s.msc_range[n + slot] = range_factor;
                s.msc_range[2 * n + slot] = limit_min;
            }
            real limit = range;
            if (safety < range)
            {
                real a = range_factor * range_init;
                real b = pp.safety_factor * safety;
                limit = a > b ? a : b;
            }
            limit = limit > limit_min ? limit : limit_min;
            if (use_safety_plus)
            {
                real rho = 1e-3;
                if (range > rho)
                {
                    real alpha = 0.35;
                    real limit_step = alpha * range + rho * (1 - alpha) * (2 - rho / range);
                    max_step = max_step < limit_step ? max_step : limit_step;
                }
            }
            if (max_step <= limit)
            {
                true_path = max_step;
            }
            else if (limit == limit_min)
            {
                true_path = limit_min;
            }
            else
            {
                NormalDist sample_gauss(limit, real(0.1) * (limit - limit_min));
                real sampled = sample_gauss(rng);
                true_path = sampled < limit_min ? limit_min
                                                : (max_step < sampled ? max_step : sampled);
            }
        }
        rng.store(s, slot);
    } while (0);

    bool limited = (true_path < phys_step);
    GeoPath gp = msc_step_to_geo(helper, true_path);
    if (gp.step > helper.lambda)
    {
        gp.step = helper.lambda;
        limited = true;
    }
    s.msc_is_displaced[slot] = displaced;
    s.msc_true_path[slot] = true_path;
    s.msc_geom_path[slot] = gp.step;
    s.msc_alpha[slot] = gp.alpha;
    s.step_length[slot] = gp.step;
    if (limited)
        s.post_step_action[slot] = pp.model_to_action - 4;  // msc action
}

//! Positron theta0 correction (em/msc/detail/UrbanPositronCorrector.hh)
B2_D real urban_positron_correction(real zeff, real y)
{
    real a = poly_lin(0.994, -4.08e-3, zeff);
    real b = poly_quad(7.16, 52.6, 365, 1 / zeff);
    real c = poly_lin(1, -4.47e-3, zeff);
    real d = real(1.21e-3) * zeff;
    real mult = poly_quad(1.41125, -1.86427e-2, 1.84035e-4, zeff);
    constexpr real xl = 0.6, xh = 0.9, e = 113;
    real x = sqrt(y * (y + 2) / ipow2(y + 1));
    real corr;
    if (x < xl)
        corr = a * (1 - exp(-b * x));
    else if (x > xh)
        corr = c + d * exp(e * (x - 1));
    else
    {
        real yl = a * (1 - exp(-b * xl));
        real yh = c + d * exp(e * (xh - 1));
        real y0 = (yh - yl) / (xh - xl);
        real y1 = yl - y0 * xl;
        corr = y0 * x + y1;
    }
    return corr * mult;
}

//! Apply msc: true path, angular deflection, lateral displacement
//! (UrbanMsc::apply_step, em/msc/detail/UrbanMscScatter.hh)
template<class Geo>
B2_D void msc_apply_step(ParamsView const& p,
                         StateView const& s,
                         u32 slot,
                         Particle const& particle,
                         PhysTrack const& phys,
                         Geo& geo)
{
    UrbanMscParams const& msc = p.model.msc;
    PhysParams const& pp = p.phys;
    u32 const n = s.num_slots;
    real const range = s.dedx_range[slot];
    MscHelper helper(p, phys, particle, range);
    bool is_displaced = s.msc_is_displaced[slot];
    real true_path = s.msc_true_path[slot];
    real geom_path = s.msc_geom_path[slot];
    real const alpha = s.msc_alpha[slot];

    u32 const psa = s.post_step_action[slot];
    if (psa == p.scalars.boundary_action || psa == p.scalars.propagation_limit_action)
    {
        geom_path = s.step_length[slot];
        true_path = msc_step_from_geo(geom_path, true_path, alpha, range, helper.lambda);
        is_displaced = false;
    }
    s.step_length[slot] = true_path;

    real safety = 0;
    if (is_displaced)
    {
        real rmax2 = (true_path - geom_path) * (true_path + geom_path);
        real displ = real(0.73) * sqrt(rmax2);
        real dd = displ * (1 + 2 * msc.safety_tol);
        displ = dd > msc.geom_limit ? dd : msc.geom_limit;
        (void)displ;
        safety = geo.find_safety();
        if (safety == 0)
            is_displaced = false;
    }
    s.msc_is_displaced[slot] = is_displaced;
    s.msc_true_path[slot] = true_path;
    s.msc_geom_path[slot] = geom_path;

    // --- UrbanMscScatter constructor
    real const* md = msc.material_data + 8 * phys.material;
    real const inc_energy = particle.energy;
    bool const is_positron = particle.id == msc.positron;
    real limit_min = s.msc_range[2 * n + slot];
    real const rad_length = material_real(p.mat, phys.material, MAT_RAD_LENGTH);
    real const mass = msc.electron_mass;
    real end_energy = 0, tau = 0, xmean = 0, x2mean = 0, theta0 = -1;
    bool skip_sampling = false;
    if (true_path == range)
        skip_sampling = true;
    else if (true_path < msc.geom_limit)
        skip_sampling = true;
    else
    {
        end_energy = helper.calc_end_energy(true_path);
        if (end_energy < 1e-6)
            skip_sampling = true;
        else if (true_path <= helper.lambda * msc.tau_small)
            skip_sampling = true;
    }
    if (!skip_sampling)
    {
        real lambda = helper.lambda;
        real lambda_end = helper.calc_msc_mfp(end_energy);
        real denom;
        if (fabs(lambda - lambda_end) < lambda * real(0.01))
            denom = helper.lambda;
        else
            denom = (lambda - lambda_end) / log(lambda / lambda_end);
        tau = true_path / denom;
        if (tau < msc.tau_big)
        {
            xmean = exp(-tau);
            x2mean = (1 + 2 * exp(real(-2.5) * tau)) / 3;
            if (limit_min == 0)
                limit_min = 1e-8;  // UrbanMscParameters::limit_min() = 10 * limit_min_fix
            limit_min = limit_min < pp.lambda_limit ? limit_min : pp.lambda_limit;
            // compute_theta0
            {
                real tp = limit_min > true_path ? limit_min : true_path;
                real y = tp / rad_length;
                if (is_positron)
                {
                    real zeff = material_real(p.mat, phys.material, MAT_ZEFF);
                    y *= urban_positron_correction(zeff, sqrt(inc_energy * end_energy) / mass);
                }
                real invbetacp = sqrt((inc_energy + mass) * (end_energy + mass)
                                      / (inc_energy * (inc_energy + 2 * mass) * end_energy
                                         * (end_energy + 2 * mass)));
                real t0 = real(13.6) * sqrt(y) * invbetacp;
                t0 *= poly_lin(md[2], md[3], log(y));
                if (true_path < limit_min)
                    t0 *= sqrt(true_path / limit_min);
                theta0 = t0 > 0 ? t0 : real(0);
            }
            if (theta0 < real(1e-8))
            {
                if (!is_displaced)
                    skip_sampling = true;
                else
                    theta0 = 0;
            }
        }
    }
    if (skip_sampling)
        return;  // unchanged

    // --- sample
    Rng rng;
    rng.load(s, slot);
    Real3 const inc_direction = geo.dir();

    auto simple_scattering = [&]() {
        real a = (2 * xmean + 9 * x2mean - 3) / (2 * xmean - 3 * x2mean + 1);
        real p_pow = (a + 2) * xmean / a;
        real result;
        do
        {
            real rdm = rng.canonical();
            result = 2 * (sample_bernoulli(rng, p_pow) ? fastpow(rdm, 1 / (a + 1)) : rdm) - 1;
        } while (fabs(result) > 1);
        return result;
    };
    auto sample_cos_theta = [&]() -> real {
        real xsi;
        {
            real maxtau = true_path < limit_min ? limit_min / helper.lambda : tau;
            real u = fastpow(maxtau, 1 / real(6));
            real radlen_mfp = true_path / (tau * rad_length);
            real r = poly_quad(md[4], md[5], md[6], u) + md[7] * log(radlen_mfp);
            xsi = r > real(1.9) ? r : real(1.9);
        }
        real ea = exp(-xsi);
        real x = ipow2(2 * sin(real(0.5) * theta0));
        real xmean_1 = 1 - x * (1 + (xsi * ea) / (1 - ea));
        if (xmean_1 <= real(0.999) * xmean)
            return simple_scattering();
        real c;
        if (fabs(xsi - 3) < real(0.001))
            c = real(3.001);
        else if (fabs(xsi - 2) < real(0.001))
            c = real(2.001);
        else
            c = xsi;
        real b1 = 2 + (c - xsi) * x;
        real d = fastpow(c * x / b1, c - 1);
        real x0 = 1 - xsi * x;
        real xmean_2 = (x0 + d - (c * x - b1 * d) / (c - 2)) / (1 - d);
        real f2x0 = (c - 1) / (c * (1 - d));
        real prob = f2x0 / (ea / (1 - ea) + f2x0);
        real qprob = xmean / (prob * xmean_1 + (1 - prob) * xmean_2);
        if (rng.canonical() >= qprob)
            return sample_uniform(rng, -1, 1);
        if (rng.canonical() < prob)
        {
            return 1 + log(sample_uniform(rng, ea, 1)) * x;
        }
        else
        {
            real var = (1 - d) * rng.canonical();
            if (var < real(0.01) * d)
            {
                var /= (d * (c - 1));
                return -1 + var * (1 - real(0.5) * var * c) * (2 + (c - xsi) * x);
            }
            else
            {
                return x * (c - xsi - c * fastpow(var + d, -1 / (c - 1))) + 1;
            }
        }
    };

    real costheta;
    if (theta0 <= 0)
        costheta = 1;
    else if (tau >= msc.tau_big)
        costheta = sample_uniform(rng, -1, 1);
    else if (2 * end_energy < inc_energy || theta0 > constants::pi / 6)
        costheta = simple_scattering();
    else
        costheta = sample_cos_theta();

    real phi = sample_uniform(rng, 0, 2 * constants::pi);
    bool displaced_action = false;
    Real3 displacement = make_real3(0, 0, 0);
    if (is_displaced)
    {
        real rmax2 = (true_path - geom_path) * (true_path + geom_path);
        real length = real(0.73) * sqrt(rmax2);
        real lim = (1 - msc.safety_tol) * safety;
        length = length < lim ? length : lim;
        if (length >= msc.geom_limit)
        {
            // sample_displacement_dir
            constexpr real cbeta = 2.160;
            constexpr real cbeta1 = 0.9988703417569197;
            real psi = -log(1 - rng.canonical() * cbeta1) / cbeta;
            real dphi = phi + (sample_bernoulli(rng, 0.5) ? psi : -psi);
            Real3 dd = make_real3(cos(dphi), sin(dphi), 0);
            dd = rotate(dd, inc_direction);
            displacement = make_real3(dd[0] * length, dd[1] * length, dd[2] * length);
            displaced_action = true;
        }
    }
    Real3 direction = rotate(from_spherical(costheta, phi), inc_direction);
    rng.store(s, slot);

    geo.set_dir(direction);
    if (displaced_action)
    {
        Real3 pos = geo.pos();
        pos[0] += displacement[0];
        pos[1] += displacement[1];
        pos[2] += displacement[2];
        geo.move_internal_pos(pos);
    }
}

//---------------------------------------------------------------------------//
// ENERGY LOSS (phys/PhysicsStepUtils.hh:176-233, alongstep/detail/{Mean,Fluct}ELoss.hh,
// em/distribution/EnergyLoss*.hh)
//---------------------------------------------------------------------------//
B2_D real calc_mean_energy_loss(ParamsView const& p,
                                StateView const& s,
                                u32 slot,
                                Particle const& particle,
                                PhysTrack const& phys,
                                real step)
{
    u32 const ppid = phys.eloss_ppid();
    real const pre_step_energy = particle.energy;
    real eloss;
    {
        u32 grid = phys.value_grid(VGT_ENERGY_LOSS, ppid);
        eloss = step * calc_xs(p.phys, grid, pre_step_energy);
    }
    if (eloss >= pre_step_energy * p.phys.linear_loss_limit)
    {
        u32 grid = phys.value_grid(VGT_RANGE, ppid);
        real range = s.dedx_range[slot];
        if (step == range)
            return pre_step_energy;
        eloss = pre_step_energy - calc_inverse_range(p.phys, grid, range - step);
    }
    return eloss;
}

//! Truncated gaussian energy loss (EnergyLossGaussianDistribution.hh)
B2_D real sample_eloss_gaussian(Rng& rng, real mean, real stddev)
{
    real const max_loss = 2 * mean;
    NormalDist sample_normal(mean, stddev);
    real result;
    do
    {
        result = sample_normal(rng);
    } while (result <= 0 || result > max_loss);
    return result;
}

B2_D real sample_fast_urban(Rng& rng, real mean, real stddev)
{
    if (stddev <= 4 * mean)
        return sample_eloss_gaussian(rng, mean, stddev);
    return sample_uniform(rng, 0, 2 * mean);
}

//! Urban energy loss fluctuation model (EnergyLossUrbanDistribution.hh)
B2_D real sample_eloss_urban(ParamsView const& p,
                             u32 material,
                             real unscaled_mean_loss,
                             real max_energy,
                             real two_mebsgs,
                             real beta_sq,
                             Rng& rng)
{
    constexpr real rate = 0.56;
    constexpr real max_collisions = 8;
    constexpr real exc_thresh = 42;
    constexpr real e_0 = 1e-5;            // ionization_energy()
    constexpr real fwhm_min_energy = 1e-3;

    real t = fwhm_min_energy / max_energy;
    real const loss_scaling = real(0.5) * (t < 1 ? t : real(1)) + real(1);
    real const mean_loss = unscaled_mean_loss / loss_scaling;
    real const* up = p.model.fluct.urban + 6 * material;
    real binding_energy[2] = {up[0], up[1]};
    real const log_binding_energy1 = up[3];
    real const osc[2] = {up[4], up[5]};
    real xs_exc[2] = {0, 0};
    real const mean_exc = material_real(p.mat, material, MAT_MEAN_EXC);
    if (max_energy > mean_exc)
    {
        real const w = log(two_mebsgs) - beta_sq;
        real const w_0 = material_real(p.mat, material, MAT_LOG_MEAN_EXC);
        if (w > w_0)
        {
            if (w > log_binding_energy1)
            {
                real const c = mean_loss * (1 - rate) / (w - w_0);
                for (int i = 0; i < 2; ++i)
                    xs_exc[i] = c * osc[i] * (w - up[2 + i]) / up[i];
            }
            else
            {
                xs_exc[0] = mean_loss * (1 - rate) / up[0];
            }
            real scaling = 4;
            if (xs_exc[0] < exc_thresh)
                scaling = real(0.5) + (scaling - real(0.5)) * sqrt(xs_exc[0] / exc_thresh);
            binding_energy[0] *= scaling;
            xs_exc[0] /= scaling;
        }
    }
    real xs_ion = mean_loss * (max_energy - e_0) / (max_energy * e_0 * log(max_energy / e_0));
    if (xs_exc[0] + xs_exc[1] > 0)
        xs_ion *= rate;

    // excitation
    real result = 0;
    {
        real mean = 0, variance = 0;
        for (int i = 0; i < 2; ++i)
        {
            if (xs_exc[i] > max_collisions)
            {
                mean += xs_exc[i] * binding_energy[i];
                variance += xs_exc[i] * ipow2(binding_energy[i]);
            }
            else if (xs_exc[i] > 0)
            {
                u32 nc = sample_poisson(rng, xs_exc[i]);
                if (nc > 0)
                {
                    // UniformRealDistribution(n - 1, n + 1) on unsigned n
                    real a = static_cast<real>(nc - 1), b = static_cast<real>(nc + 1);
                    result += sample_uniform(rng, a, b) * binding_energy[i];
                }
            }
        }
        if (variance > 0)
            result += sample_fast_urban(rng, mean, sqrt(variance));
    }
    // ionisation
    {
        real const energy_ratio = max_energy / e_0;
        real alpha = 1;
        real mean_num_coll = 0;
        if (xs_ion > max_collisions)
        {
            alpha = (xs_ion + max_collisions) * energy_ratio
                    / (max_collisions * energy_ratio + xs_ion);
            real const mean_loss_coll = alpha * log(alpha) / (alpha - 1);
            mean_num_coll = xs_ion * energy_ratio * (alpha - 1) / ((energy_ratio - 1) * alpha);
            real const mean = mean_num_coll * mean_loss_coll * e_0;
            real const stddev = e_0 * sqrt(xs_ion * (alpha - ipow2(mean_loss_coll)));
            result += sample_fast_urban(rng, mean, stddev);
        }
        if (xs_ion > 0 && energy_ratio > alpha)
        {
            u32 nion = sample_poisson(rng, xs_ion - mean_num_coll);
            real const a = alpha / energy_ratio;
            for (u32 k = nion; k > 0; --k)
                result += alpha * e_0 / sample_uniform(rng, a, 1);
        }
    }
    return loss_scaling * result;
}

//! Gamma-distributed loss (EnergyLossGammaDistribution.hh, GammaDistribution.hh)
B2_D real sample_eloss_gamma(Rng& rng, real mean, real var)
{
    real const k = ipow2(mean) / var;
    real const alpha = k, beta = mean / k;
    real const alpha_p = alpha < 1 ? alpha + 1 : alpha;
    real const d = alpha_p - real(1) / 3;
    real const c = 1 / sqrt(9 * d);
    NormalDist sample_normal(0, 1);
    real u, v, z;
    do
    {
        do
        {
            z = sample_normal(rng);
            v = 1 + c * z;
        } while (v <= 0);
        v = v * v * v;
        u = rng.canonical();
    } while (u > 1 - real(0.0331) * ipow2(ipow2(z))
             && log(u) > real(0.5) * ipow2(z) + d * (1 - v + log(v)));
    real result = d * v * beta;
    if (alpha != alpha_p)
        result *= fastpow(rng.canonical(), 1 / alpha);
    return result;
}

//! Energy lost along the step [MeV] (FluctELoss::calc_eloss / MeanELoss::calc_eloss)
B2_D real calc_eloss(ParamsView const& p,
                     StateView const& s,
                     u32 slot,
                     Particle const& particle,
                     PhysTrack const& phys,
                     real step,
                     bool apply_cut)
{
    real const lowest = p.phys.lowest_electron_energy;
    if (apply_cut && particle.energy < lowest)
        return particle.energy;
    real eloss = calc_mean_energy_loss(p, s, slot, particle, phys, step);
    FluctuationParams const& fl = p.model.fluct;
    if (fl.enabled && eloss < particle.energy)
    {
        // EnergyLossHelper
        real const mean_loss = eloss;
        int model = 0;  // none
        real beta_sq = 0, max_energy = 0, two_mebsgs = 0, bohr_var = 0;
        if (!(mean_loss < 1e-5))
        {
            real const gamma = particle.lorentz_factor();
            beta_sq = particle.beta_sq();
            two_mebsgs = 2 * fl.electron_mass * beta_sq * ipow2(gamma);
            real max_energy_transfer;
            real mass_ratio = 1;
            if (particle.id == fl.electron)
            {
                max_energy_transfer = real(0.5) * particle.energy;
            }
            else
            {
                mass_ratio = fl.electron_mass / particle.mass;
                max_energy_transfer = two_mebsgs / (1 + mass_ratio * (2 * gamma + mass_ratio));
            }
            real ecut = cutoff_energy(p, phys.material, fl.electron);
            max_energy = ecut < max_energy_transfer ? ecut : max_energy_transfer;
            if (!(max_energy <= 1e-5))
            {
                real const re = p.model.constants.r_electron;
                bohr_var = 2 * constants::pi * ipow2(re) * fl.electron_mass
                           * material_real(p.mat, phys.material, MAT_ELECTRON_DENSITY)
                           * ipow2(particle.charge) * max_energy * step * (1 / beta_sq - real(0.5));
                if (mass_ratio >= 1 || mean_loss < 10 * max_energy
                    || max_energy_transfer > 2 * max_energy)
                    model = 3;  // urban
                else if (ipow2(mean_loss) >= 4 * bohr_var)
                    model = 2;  // gaussian
                else
                    model = 1;  // gamma
            }
        }
        Rng rng;
        rng.load(s, slot);
        switch (model)
        {
            case 0: break;  // none: eloss stays the mean (EnergyLossDeltaDistribution)
            case 1: eloss = sample_eloss_gamma(rng, mean_loss, bohr_var); break;
            case 2: eloss = sample_eloss_gaussian(rng, mean_loss, sqrt(bohr_var)); break;
            case 3:
                eloss = sample_eloss_urban(
                    p, phys.material, mean_loss, max_energy, two_mebsgs, beta_sq, rng);
                break;
        }
        rng.store(s, slot);
        if (eloss >= particle.energy)
            eloss = apply_cut ? particle.energy : mean_loss;
    }
    if (apply_cut && (particle.energy - eloss <= lowest))
        return particle.energy;
    return eloss;
}

//! Apply continuous energy loss (alongstep/detail/ElossApplier.hh)
B2_D void apply_eloss(ParamsView const& p, StateView const& s, u32 slot, PhysTrack const& phys)
{
    if (s.status[slot] == ST_ERRORED)
        return;
    if (phys.eloss_ppid() == INVALID)
        return;
    Particle particle = load_particle(p, s, slot);
    if (particle.is_stopped())
        return;
    real const step = s.step_length[slot];
    bool const apply_cut = (s.post_step_action[slot] != p.scalars.boundary_action);
    real deposited = calc_eloss(p, s, slot, particle, phys, step, apply_cut);
    if (deposited > 0)
    {
        s.energy_deposition[slot] += deposited;
        // ParticleTrackView::subtract_energy
        particle.energy -= deposited;
        s.energy[slot] = particle.energy;
    }
    if (particle.is_stopped())
    {
        if (!phys.has_at_rest())
        {
            s.status[slot] = ST_KILLED;
            s.post_step_action[slot] = p.phys.model_to_action - 3;  // range action
        }
        else
        {
            s.post_step_action[slot] = p.phys.model_to_action - 2;  // discrete action
        }
    }
}

//---------------------------------------------------------------------------//
// PROPAGATION / TIME / TRACK UPDATE
//---------------------------------------------------------------------------//
//! Straight-line propagation up to `dist` (field/LinearPropagator.hh:58-93)
template<class Geo>
B2_D Propagation propagate_linear(Geo& geo, real dist)
{
    Propagation result = geo.find_next_step(true, dist);
    if (result.boundary)
        geo.move_to_boundary();
    else
        geo.move_internal(dist);
    return result;
}

//! Apply the propagation result to the step (detail/PropagationApplier.hh:93-192)
B2_D void apply_propagation(ParamsView const& p,
                            StateView const& s,
                            u32 slot,
                            Propagation const& pr,
                            bool tracks_can_loop,
                            Particle const& particle)
{
    if (tracks_can_loop)
    {
        // SimTrackView::update_looping
        if (pr.looping)
            s.num_looping_steps[slot] += 1;
        else
            s.num_looping_steps[slot] = 0;
    }
    if (tracks_can_loop && pr.looping)
    {
        s.step_length[slot] = pr.distance;
        bool abandon = false;
        if (p.particle.decay_constant[particle.id] == 0)
        {
            // SimTrackView::is_looping
            u32 nloop = s.num_looping_steps[slot];
            if (particle.energy < p.sim.looping_energy[particle.id])
                abandon = nloop >= p.sim.looping_steps[2 * particle.id];
            else
                abandon = nloop >= p.sim.looping_steps[2 * particle.id + 1];
        }
        s.post_step_action[slot] = abandon ? p.scalars.tracking_cut_action
                                           : p.scalars.propagation_limit_action;
    }
    else if (pr.boundary)
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

//---------------------------------------------------------------------------//
// The charged along-step in four phases, one kernel each (the reference's along-step is
// the same sequence of appliers: global/alongstep/detail/AlongStepImpl / AlongStep.hh:
// 50-58). Everything passed between phases already lives in the per-slot state
// (step_length, post_step_action, msc_{true,geom}_path, msc_alpha, msc_is_displaced), so
// the split changes no result. Smaller kernels need fewer registers (more resident warps
// to hide the latency-bound chains) and keep warps convergent within a phase.
//---------------------------------------------------------------------------//
//! Phase 1: MSC step limit (msc_geom_path = 0 marks "MSC not applied this step")
B2_D void along_phase_msc_limit(ParamsView const& p, StateView const& s, u32 slot)
{
    Particle particle = load_particle(p, s, slot);
    if (msc_is_applicable(p, s, slot, particle, s.step_length[slot]))
    {
        GeoTrack geo(p, s, slot);
        PhysTrack phys(p, particle.id, s.material_id[slot]);
        msc_limit_step(p, s, slot, particle, phys, geo);
    }
    else
    {
        s.msc_geom_path[slot] = 0;
    }
}

//! Phase 2: propagation through the geometry. FIELD is a property of the problem (the
//! along-step action it was built with), so it selects the kernel at launch: 0 = none,
//! 1 = uniform field, 2 = r-z map field.
template<int FIELD>
B2_D void along_phase_propagate(ParamsView const& p, StateView const& s, u32 slot)
{
    if (s.step_length[slot] == 0)
        return;
    Particle particle = load_particle(p, s, slot);
    GeoTrack geo(p, s, slot);
    Propagation pr;
    if constexpr (FIELD != 0)
        pr = propagate_field_impl<FIELD == 2>(p, particle, geo, s.step_length[slot]);
    else
        pr = propagate_linear(geo, s.step_length[slot]);
    apply_propagation(p, s, slot, pr, FIELD != 0, particle);
}

//! Phase 3: MSC scattering and displacement
B2_D void along_phase_msc_apply(ParamsView const& p, StateView const& s, u32 slot)
{
    if (!(s.msc_geom_path[slot] > 0))
        return;
    Particle particle = load_particle(p, s, slot);
    GeoTrack geo(p, s, slot);
    PhysTrack phys(p, particle.id, s.material_id[slot]);
    msc_apply_step(p, s, slot, particle, phys, geo);
}

//! Phase 4: time, energy loss, track bookkeeping
B2_D void along_phase_finish(ParamsView const& p, StateView const& s, u32 slot)
{
    Particle particle = load_particle(p, s, slot);
    PhysTrack phys(p, particle.id, s.material_id[slot]);
    update_time(s, slot, particle);
    apply_eloss(p, s, slot, phys);
    update_track(p, s, slot);
}

//! Whole along-step for one alive track. CHARGED is a compile-time property of the
//! launch (dense per-charge slot lists), so the neutral kernel carries no msc/eloss code.
//! COOP: the warp's 32 lanes all run this track (see GeoTrackT in orange.cuh).
template<bool CHARGED, int FIELD, bool COOP = false>
B2_D void along_step(ParamsView const& p, StateView const& s, u32 slot)
{
    Particle particle = load_particle(p, s, slot);
    GeoTrackT<COOP> geo(p, s, slot);
    PhysTrack phys(p, particle.id, s.material_id[slot]);
    constexpr bool charged = CHARGED;
    constexpr bool use_field = CHARGED && FIELD != 0;

    // msc step limit
    bool use_msc = false;
    if (charged)
    {
        if (msc_is_applicable(p, s, slot, particle, s.step_length[slot]))
        {
            msc_limit_step(p, s, slot, particle, phys, geo);
            use_msc = true;
        }
        else
        {
            s.msc_geom_path[slot] = 0;
        }
    }
    // propagation
    if (s.step_length[slot] != 0)
    {
        Propagation pr;
        if constexpr (use_field)
            pr = propagate_field_impl<FIELD == 2>(p, particle, geo, s.step_length[slot]);
        else
            pr = propagate_linear(geo, s.step_length[slot]);
        apply_propagation(p, s, slot, pr, use_field, particle);
    }
    if (charged)
    {
        // msc scatter
        if (use_msc && s.status[slot] == ST_ALIVE && s.msc_geom_path[slot] > 0)
            msc_apply_step(p, s, slot, particle, phys, geo);
    }
    if (s.status[slot] != ST_ERRORED)
        update_time(s, slot, particle);
    if (charged)
        apply_eloss(p, s, slot, phys);
    update_track(p, s, slot);
}
}  // namespace b200
