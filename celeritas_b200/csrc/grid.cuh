//---------------------------------------------------------------------------//
// Value-grid interpolation: cross sections, range, inverse range.
//
// A grid is uniform in log(E): {log_front, log_delta, size}; values are linear
// in E between nodes; nodes at index >= prime_index hold xs*E
// (/root/reference/src/celeritas/grid/XsGridData.hh,
//  /root/reference/src/celeritas/grid/XsCalculator.hh:107-153,
//  RangeCalculator.hh:80-108, InverseRangeCalculator.hh:86-116).
//---------------------------------------------------------------------------//
#pragma once

#include "views.cuh"

namespace b200
{
struct GridRef
{
    PhysParams const& p;
    u32 id;

    B2_D u32 size() const { return p.grid_size[id]; }
    B2_D real front() const { return p.grid_log_front[id]; }
    B2_D real back() const { return p.grid_log_back[id]; }
    B2_D real delta() const { return p.grid_log_delta[id]; }
    B2_D u32 prime() const { return p.grid_prime[id]; }
    B2_D real value(u32 i) const { return p.reals[p.grid_value_offset[id] + i]; }
    B2_D real log_node(u32 i) const { return front() + delta() * i; }
    //! Energy of node i: exp(log_node(i)), tabulated on the host
    B2_D real energy(u32 i) const { return p.grid_energy[p.grid_energy_offset[id] + i]; }
};

//! Cross section at `energy` (XsCalculator::operator())
B2_D real calc_xs(PhysParams const& p, u32 grid_id, real energy)
{
    GridRef g{p, grid_id};
    real const loge = log(energy);
    u32 const prime = g.prime();
    real const front = g.front();
    if (loge <= front)
    {
        real r = g.value(0);
        if (0 >= prime)
            r /= energy;
        return r;
    }
    u32 const n = g.size();
    if (loge >= g.back())
    {
        real r = g.value(n - 1);
        if (n - 1 >= prime)
            r /= energy;
        return r;
    }
    real const delta = g.delta();
    u32 const lower = static_cast<u32>((loge - front) / delta);
    real const upper_energy = g.energy(lower + 1);
    real upper_xs = g.value(lower + 1);
    if (lower + 1 == prime)
        upper_xs /= upper_energy;
    real result = lerp_points(g.energy(lower), g.value(lower), upper_energy, upper_xs, energy);
    if (lower >= prime)
        result /= energy;
    return result;
}

//! Range at `energy` (RangeCalculator::operator())
B2_D real calc_range(PhysParams const& p, u32 grid_id, real energy)
{
    GridRef g{p, grid_id};
    real const loge = log(energy);
    real const front = g.front();
    if (loge <= front)
    {
        real r = g.value(0);
        r *= exp(real(.5) * (loge - front));
        return r;
    }
    u32 const n = g.size();
    if (loge >= g.back())
        return g.value(n - 1);
    real const delta = g.delta();
    u32 const idx = static_cast<u32>((loge - front) / delta);
    return lerp_points(g.energy(idx), g.value(idx), g.energy(idx + 1), g.value(idx + 1), energy);
}

//! Energy for a given range (InverseRangeCalculator::operator())
B2_D real calc_inverse_range(PhysParams const& p, u32 grid_id, real range)
{
    GridRef g{p, grid_id};
    u32 const n = g.size();
    real const r_front = g.value(0);
    if (range < r_front)
        return g.energy(0) * ipow2(range / r_front);
    real const r_back = g.value(n - 1);
    if (range >= r_back)
        return exp(g.back());
    // lower_bound over the (monotonic) range values, then step back unless
    // exactly on a node (NonuniformGrid::find)
    u32 lo = 0, len = n;
    while (len > 0)
    {
        u32 half = len >> 1;
        u32 mid = lo + half;
        if (g.value(mid) < range)
        {
            lo = mid + 1;
            len -= half + 1;
        }
        else
            len = half;
    }
    u32 idx = lo;
    if (range != g.value(idx))
        --idx;
    return lerp_points(g.value(idx), g.energy(idx), g.value(idx + 1), g.energy(idx + 1), range);
}

//! Tabulated xs at node i (XsCalculator::operator[])
B2_D real calc_xs_at_node(PhysParams const& p, u32 grid_id, u32 i)
{
    GridRef g{p, grid_id};
    real energy = g.energy(i);
    real r = g.value(i);
    if (i >= g.prime())
        r /= energy;
    return r;
}
}  // namespace b200
