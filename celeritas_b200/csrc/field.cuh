//---------------------------------------------------------------------------//
// Charged-particle propagation in a uniform magnetic field.
//
// Dormand-Prince RK5(4)7M stepper with error estimate and dense midpoint
// (/root/reference/src/celeritas/field/DormandPrinceStepper.hh:101-219), the
// chord/accuracy driver (field/FieldDriver.hh:180-426) and the boundary-aware
// propagator (field/FieldPropagator.hh:149-331), for the Lorentz equation of
// motion (field/MagFieldEquation.hh:103-121) in a UniformField.
//---------------------------------------------------------------------------//
#pragma once

#include "orange.cuh"
#include "physics.cuh"

namespace b200
{
struct OdeState
{
    Real3 pos;
    Real3 mom;
};

B2_D void ode_axpy(real a, OdeState const& x, OdeState& y)
{
    axpy(a, x.pos, y.pos);
    axpy(a, x.mom, y.mom);
}

B2_D Real3 cross_product(Real3 const& x, Real3 const& y)
{
    return make_real3(x[1] * y[2] - x[2] * y[1], x[2] * y[0] - x[0] * y[2], x[0] * y[1] - x[1] * y[0]);
}

struct FieldStepperResult
{
    OdeState mid_state;
    OdeState end_state;
    OdeState err_state;
};

// The Dormand-Prince trial step has five call sites in the driver (short step, chord search,
// one_good_step, integrate_step) and is ~1500 SASS instructions with --fmad=false: inlined
// everywhere, the charged along-step kernel with field is 14 k instructions (224 KB), more
// than the SM's instruction caches hold, and ncu shows 44 "no instruction" stall cycles per
// issued instruction at a saturated CMS-scale iteration (profiles/README_r01.md). Out of
// line it is one copy.  B2_FIELD_OUTLINE: 0 = inline, 1 = trial step out of line,
// 2 = + right-hand side out of line.
// MEASURED (CMS-scale stand-in, gpurun_out/variants_cms.log): out of line is SLOWER, 3.84 ->
// 4.25 (level 1) / 4.62 (level 2) ns per track-step at saturation and 479 -> 557 / 724 us per
// tail iteration: the call passes 18 doubles of result through local memory and the trial
// step is only ~4 % of the kernel's code (division and square root slow paths are already
// shared subroutines). Kept as a knob, default inline; what helps is running the
// propagation as its own kernel (B2_ALONG_SPLIT_FIELD_THRESHOLD, kernels.cu).
#ifndef B2_FIELD_OUTLINE
#    define B2_FIELD_OUTLINE 0
#endif
#if B2_FIELD_OUTLINE >= 1
#    define B2_FIELD_STEP_FN B2_NOINLINE inline
#else
#    define B2_FIELD_STEP_FN B2_D
#endif
#if B2_FIELD_OUTLINE >= 2
#    define B2_FIELD_RHS_FN B2_NOINLINE inline
#else
#    define B2_FIELD_RHS_FN B2_D
#endif

//! The magnetic field seen by the equation of motion: a constant vector (UniformField) or
//! the r-z map (RZMapField::operator(), field/RZMapField.hh:67-106). The kind is a
//! compile-time parameter: as a run-time branch inside the right-hand side (seven
//! evaluations per Dormand-Prince trial) the map code cost the UNIFORM-field problems 35 % of
//! their along-step (CMS-scale pass 351 -> 433 ms, profiles/README_r02.md).
template<bool RZ>
struct FieldSource
{
    Real3 uniform;
    FieldParams const* map;

    B2_D Real3 operator()(Real3 const& pos) const
    {
        if constexpr (!RZ)
        {
            return uniform;
        }
        else
        {
            FieldParams const& f = *map;
            Real3 value = make_real3(0, 0, 0);
            real const r = sqrt(ipow2(pos[0]) + ipow2(pos[1]));
            real const z = pos[2];
            if (!(z >= f.rz_z[0] && z <= f.rz_z[1] && r >= f.rz_r[0] && r <= f.rz_r[1]))
                return value;
            // find_interp<UniformGrid> (corecel/grid/FindInterp.hh:43-57)
            u32 const ir = static_cast<u32>((r - f.rz_r[0]) / f.rz_r[2]);
            u32 const iz = static_cast<u32>((z - f.rz_z[0]) / f.rz_z[2]);
            real const r_lo = f.rz_r[0] + f.rz_r[2] * ir, r_hi = f.rz_r[0] + f.rz_r[2] * (ir + 1);
            real const z_lo = f.rz_z[0] + f.rz_z[2] * iz, z_hi = f.rz_z[0] + f.rz_z[2] * (iz + 1);
            real const frac_r = (r - r_lo) / (r_hi - r_lo);
            real const frac_z = (z - z_lo) / (z_hi - z_lo);
            real const* v = f.rz_values;
            u32 const nr = f.rz_size_r;
            real low = v[2 * (iz * nr + ir)];
            real high = v[2 * ((iz + 1) * nr + ir)];
            value[2] = low + (high - low) * frac_z;
            low = v[2 * (iz * nr + ir) + 1];
            high = v[2 * (iz * nr + ir + 1) + 1];
            real const tmp = (r != 0) ? (low + (high - low) * frac_r) / r : low;
            value[0] = tmp * pos[0];
            value[1] = tmp * pos[1];
            return value;
        }
    }
};

//! Right-hand side of the equation of motion
template<bool RZ>
B2_FIELD_RHS_FN OdeState field_rhs(real coeffi, FieldSource<RZ> const& source, OdeState const& y)
{
    real momentum_inv = 1 / sqrt(dot(y.mom, y.mom));
    OdeState r;
    r.pos = make_real3(momentum_inv * y.mom[0], momentum_inv * y.mom[1], momentum_inv * y.mom[2]);
    real c = coeffi * momentum_inv;
    Real3 const field = source(y.pos);
    Real3 x = cross_product(y.mom, field);
    r.mom = make_real3(c * x[0], c * x[1], c * x[2]);
    return r;
}

//! One Dormand-Prince trial step
template<bool RZ>
B2_FIELD_STEP_FN void field_apply_step(real coeffi, FieldSource<RZ> const& field, real step, OdeState const& beg, FieldStepperResult& result)
{
    using R = real;
    constexpr R a11 = 0.2;
    constexpr R a21 = 0.075;
    constexpr R a22 = 0.225;
    constexpr R a31 = 44 / R(45);
    constexpr R a32 = -56 / R(15);
    constexpr R a33 = 32 / R(9);
    constexpr R a41 = 19372 / R(6561);
    constexpr R a42 = -25360 / R(2187);
    constexpr R a43 = 64448 / R(6561);
    constexpr R a44 = -212 / R(729);
    constexpr R a51 = 9017 / R(3168);
    constexpr R a52 = -355 / R(33);
    constexpr R a53 = 46732 / R(5247);
    constexpr R a54 = 49 / R(176);
    constexpr R a55 = -5103 / R(18656);
    constexpr R a61 = 35 / R(384);
    constexpr R a63 = 500 / R(1113);
    constexpr R a64 = 125 / R(192);
    constexpr R a65 = -2187 / R(6784);
    constexpr R a66 = 11 / R(84);
    constexpr R d71 = a61 - 5179 / R(57600);
    constexpr R d73 = a63 - 7571 / R(16695);
    constexpr R d74 = a64 - 393 / R(640);
    constexpr R d75 = a65 + 92097 / R(339200);
    constexpr R d76 = a66 - 187 / R(2100);
    constexpr R d77 = -1 / R(40);
    constexpr R c71 = R(6025192743.) / R(30085553152.);
    constexpr R c73 = R(51252292925.) / R(65400821598.);
    constexpr R c74 = R(-2691868925.) / R(45128329728.);
    constexpr R c75 = R(187940372067.) / R(1594534317056.);
    constexpr R c76 = R(-1776094331.) / R(19743644256.);
    constexpr R c77 = R(11237099.) / R(235043384.);

    OdeState k1 = field_rhs(coeffi, field, beg);
    OdeState state = beg;
    ode_axpy(a11 * step, k1, state);
    OdeState k2 = field_rhs(coeffi, field, state);
    state = beg;
    ode_axpy(a21 * step, k1, state);
    ode_axpy(a22 * step, k2, state);
    OdeState k3 = field_rhs(coeffi, field, state);
    state = beg;
    ode_axpy(a31 * step, k1, state);
    ode_axpy(a32 * step, k2, state);
    ode_axpy(a33 * step, k3, state);
    OdeState k4 = field_rhs(coeffi, field, state);
    state = beg;
    ode_axpy(a41 * step, k1, state);
    ode_axpy(a42 * step, k2, state);
    ode_axpy(a43 * step, k3, state);
    ode_axpy(a44 * step, k4, state);
    OdeState k5 = field_rhs(coeffi, field, state);
    state = beg;
    ode_axpy(a51 * step, k1, state);
    ode_axpy(a52 * step, k2, state);
    ode_axpy(a53 * step, k3, state);
    ode_axpy(a54 * step, k4, state);
    ode_axpy(a55 * step, k5, state);
    OdeState k6 = field_rhs(coeffi, field, state);
    result.end_state = beg;
    ode_axpy(a61 * step, k1, result.end_state);
    ode_axpy(a63 * step, k3, result.end_state);
    ode_axpy(a64 * step, k4, result.end_state);
    ode_axpy(a65 * step, k5, result.end_state);
    ode_axpy(a66 * step, k6, result.end_state);
    OdeState k7 = field_rhs(coeffi, field, result.end_state);
    result.err_state.pos = make_real3(0, 0, 0);
    result.err_state.mom = make_real3(0, 0, 0);
    ode_axpy(d71 * step, k1, result.err_state);
    ode_axpy(d73 * step, k3, result.err_state);
    ode_axpy(d74 * step, k4, result.err_state);
    ode_axpy(d75 * step, k5, result.err_state);
    ode_axpy(d76 * step, k6, result.err_state);
    ode_axpy(d77 * step, k7, result.err_state);
    real half_step = step / real(2);
    result.mid_state = beg;
    ode_axpy(c71 * half_step, k1, result.mid_state);
    ode_axpy(c73 * half_step, k3, result.mid_state);
    ode_axpy(c74 * half_step, k4, result.mid_state);
    ode_axpy(c75 * half_step, k5, result.mid_state);
    ode_axpy(c76 * half_step, k6, result.mid_state);
    ode_axpy(c77 * half_step, k7, result.mid_state);
}


template<bool RZ>
struct FieldDriver
{
    FieldParams const& opt;
    real coeffi;     // charge / momentum unit
    FieldSource<RZ> field;
    real max_chord;

    B2_D FieldDriver(FieldParams const& f, real charge) : opt(f), max_chord(real_inf())
    {
        coeffi = charge * f.coeffi_per_charge;
        field.uniform = make_real3(f.field[0], f.field[1], f.field[2]);
        field.map = &f;
    }

    //! One Dormand-Prince trial step
    B2_D FieldStepperResult apply_step(real step, OdeState const& beg) const
    {
        FieldStepperResult result;
        field_apply_step(coeffi, field, step, beg, result);
        return result;
    }

    //! Relative truncation error squared (detail/FieldUtils.hh: rel_err_sq)
    B2_D real rel_err_sq(OdeState const& err, real step, Real3 const& mom) const
    {
        real errpos2 = dot(err.pos, err.pos);
        real errvel2 = dot(err.mom, err.mom);
        errpos2 /= ipow2(step);
        errvel2 /= dot(mom, mom);
        return errpos2 > errvel2 ? errpos2 : errvel2;
    }

    B2_D real new_step_scale(real err_sq) const
    {
        return opt.safety * exp((real(0.5) * (err_sq > 1 ? opt.pshrink : opt.pgrow)) * log(err_sq));
    }

    struct DriverResult
    {
        OdeState state;
        real step;
    };

    struct Integration
    {
        DriverResult end;
        real proposed_step;
    };

    B2_D Integration one_good_step(real step, OdeState const& state) const
    {
        Integration output;
        bool succeeded = false;
        u32 remaining_steps = opt.max_nsteps;
        real err_sq;
        FieldStepperResult result;
        do
        {
            result = apply_step(step, state);
            err_sq = rel_err_sq(result.err_state, step, state.mom) / ipow2(opt.epsilon_rel_max);
            if (err_sq > 1)
            {
                real sc = new_step_scale(err_sq);
                step *= sc > opt.max_stepping_decrease ? sc : opt.max_stepping_decrease;
            }
            else
            {
                succeeded = true;
            }
        } while (!succeeded && --remaining_steps > 0);
        output.end.state = result.end_state;
        output.end.step = step;
        real sc = new_step_scale(err_sq);
        output.proposed_step = step * (sc < opt.max_stepping_increase ? sc : opt.max_stepping_increase);
        return output;
    }

    B2_D Integration integrate_step(real step, OdeState const& state) const
    {
        Integration output;
        if (step > opt.minimum_step)
        {
            output = one_good_step(step, state);
        }
        else
        {
            FieldStepperResult result = apply_step(step, state);
            output.end.state = result.end_state;
            output.end.step = step;
            real err_sq = rel_err_sq(result.err_state, step, state.mom) / ipow2(opt.epsilon_rel_max);
            output.proposed_step = step * new_step_scale(err_sq);
        }
        return output;
    }

    B2_D DriverResult accurate_advance(real step, OdeState const& state, real hinitial) const
    {
        real end_curve_length = step;
        constexpr real initial_step_tol = 1e-6;
        real h = ((hinitial > initial_step_tol * step) && (hinitial < step)) ? hinitial : step;
        real h_threshold = opt.epsilon_step * step;
        Integration output;
        output.end.state = state;
        output.end.step = 0;
        output.proposed_step = 0;
        bool succeeded = false;
        real curve_length = 0;
        int remaining_steps = static_cast<int>(opt.max_nsteps);
        do
        {
            output = integrate_step(h, output.end.state);
            curve_length += output.end.step;
            if (h < h_threshold || curve_length >= end_curve_length)
            {
                succeeded = true;
            }
            else
            {
                real a = output.proposed_step > opt.minimum_step ? output.proposed_step
                                                                 : opt.minimum_step;
                real b = end_curve_length - curve_length;
                h = a < b ? a : b;
            }
        } while (!succeeded && --remaining_steps > 0);
        output.end.step = curve_length < step ? curve_length : step;
        return output.end;
    }

    //! Advance by up to `step` with a chord within delta_chord of the curve
    B2_D DriverResult advance(real step, OdeState const& state)
    {
        if (step <= opt.minimum_step)
        {
            DriverResult result;
            result.state = apply_step(step, state).end_state;
            result.step = step;
            return result;
        }
        // find_next_chord
        constexpr real dchord_tol = 1e-5 * 0.1;  // 1e-5 mm
        constexpr real min_chord_shrink = 0.5;
        real cstep = step < max_chord ? step : max_chord;
        DriverResult end;
        real err_sq;
        {
            bool succeeded = false;
            int remaining_steps = static_cast<int>(opt.max_nsteps);
            FieldStepperResult result;
            do
            {
                result = apply_step(cstep, state);
                // distance_chord
                Real3 beg_mid, beg_end;
                for (int i = 0; i < 3; ++i)
                {
                    beg_mid[i] = result.mid_state.pos[i] - state.pos[i];
                    beg_end[i] = result.end_state.pos[i] - state.pos[i];
                }
                Real3 cr = cross_product(beg_end, beg_mid);
                real dchord = sqrt(dot(cr, cr) / dot(beg_end, beg_end));
                if (dchord > opt.delta_chord + dchord_tol)
                {
                    real sc = sqrt(opt.delta_chord / dchord);
                    cstep *= sc > min_chord_shrink ? sc : min_chord_shrink;
                }
                else
                {
                    succeeded = true;
                }
            } while (!succeeded && --remaining_steps > 0);
            end.step = cstep;
            end.state = result.end_state;
            err_sq = rel_err_sq(result.err_state, cstep, state.mom) / ipow2(opt.epsilon_rel_max);
        }
        if (end.step < step)
            max_chord = end.step * (1 / min_chord_shrink);
        if (err_sq > 1)
        {
            real next_step = step * new_step_scale(err_sq);
            end = accurate_advance(end.step, state, next_step);
        }
        return end;
    }
};

//! Propagate a charged track in the field up to `step` or the next boundary
template<bool RZ, class Geo>
B2_D Propagation propagate_field_impl(ParamsView const& p, Particle const& particle, Geo& geo, real step)
{
    FieldParams const& opt = p.model.field;
    FieldDriver<RZ> driver(opt, particle.charge);
    OdeState state;
    state.pos = geo.pos();
    {
        real mom = particle.momentum();
        Real3 d = geo.dir();
        state.mom = make_real3(mom * d[0], mom * d[1], mom * d[2]);
    }
    real const delta_intersection = opt.delta_intersection;
    real const minimum_substep = opt.minimum_step;
    real const bump_distance = delta_intersection * real(0.1);

    Propagation result;
    result.boundary = geo.is_on_boundary();
    result.distance = 0;
    result.looping = false;
    real remaining = step;
    int remaining_substeps = static_cast<int>(opt.max_substeps);
    do
    {
        typename FieldDriver<RZ>::DriverResult substep = driver.advance(remaining, state);
        // make_chord
        Real3 cdir = make_real3(substep.state.pos[0] - state.pos[0],
                                substep.state.pos[1] - state.pos[1],
                                substep.state.pos[2] - state.pos[2]);
        real clen = norm(cdir);
        cdir[0] /= clen;
        cdir[1] /= clen;
        cdir[2] /= clen;
        if (clen >= minimum_substep)
            geo.set_dir(cdir);
        Propagation linear_step = geo.find_next_step(true, clen + delta_intersection);
        real const update_length = substep.step * linear_step.distance / clen;
        if (!linear_step.boundary)
        {
            state = substep.state;
            result.boundary = false;
            result.distance += substep.step;
            remaining = step - result.distance;
            geo.move_internal_pos(state.pos);
            --remaining_substeps;
        }
        else if (result.boundary && linear_step.distance < bump_distance)
        {
            remaining = substep.step / 2;
        }
        else
        {
            bool close = false;
            if (!(update_length <= minimum_substep))
            {
                real delta_sq = 0;
                for (int i = 0; i < 3; ++i)
                {
                    delta_sq += ipow2(state.pos[i] - substep.state.pos[i]
                                      + linear_step.distance * cdir[i]);
                }
                close = delta_sq <= ipow2(delta_intersection);
            }
            if (update_length <= minimum_substep || close || clen == 0)
            {
                result.boundary = (linear_step.distance <= clen
                                   || result.distance + update_length <= step || clen == 0);
                if (!result.boundary)
                {
                    state.pos = substep.state.pos;
                    geo.move_internal_pos(substep.state.pos);
                }
                result.distance += update_length < substep.step ? update_length : substep.step;
                state.mom = substep.state.mom;
                remaining = 0;
            }
            else
            {
                remaining = update_length;
            }
        }
    } while (remaining > minimum_substep && remaining_substeps > 0);

    if (remaining_substeps == 0 && result.distance < step)
    {
        result.looping = true;
    }
    else if (result.distance > 0)
    {
        if (result.boundary)
        {
            geo.move_to_boundary();
            state.pos = geo.pos();
        }
        else if (result.distance < step)
        {
            result.distance = step;
        }
    }
    Real3 dir = make_unit_vector(state.mom);
    geo.set_dir(dir);
    if (result.distance == 0)
    {
        result.distance = bump_distance < step ? bump_distance : step;
        result.boundary = false;
        axpy(result.distance, dir, state.pos);
        geo.move_internal_pos(state.pos);
    }
    return result;
}
}  // namespace b200
