//---------------------------------------------------------------------------//
// XORWOW engine over structure-of-arrays state, and sampling distributions.
//
// Bit-exact with the reference's XorwowRngEngine
// (/root/reference/src/celeritas/random/XorwowRngEngine.hh:163-314): same
// xorshift recurrence, Weyl increment 362437, SplitMix64 seeding and
// jump-polynomial skip-ahead; canonical doubles are built from two 32-bit
// draws as (hi << 21 ^ lo) * 2^-53
// (/root/reference/src/celeritas/random/detail/GenerateCanonical32.hh:75-92).
//
// The engine keeps the six state words in registers; it is loaded once when a
// kernel starts working on a slot and stored once at the end.
//---------------------------------------------------------------------------//
#pragma once

#include "views.cuh"

namespace b200
{
struct Rng
{
    u32 x[5];
    u32 d;

    // Layout of StateView::rng (24 B per slot either way):
    //   B2_RNG_PACKED 0  six columns [6][slot]: six 4-byte accesses per slot
    //   B2_RNG_PACKED 1  uint4 {x0..x3}[slot] followed by uint2 {x4, weyl}[slot]: one 16-byte
    //                    and one 8-byte access per slot (2 sectors per slot instead of 6 when
    //                    the slots of a warp are scattered)
    B2_D void load(StateView const& s, u32 slot)
    {
#if B2_RNG_PACKED
        uint4 const a = reinterpret_cast<uint4 const*>(s.rng)[slot];
        uint2 const b = reinterpret_cast<uint2 const*>(s.rng + size_t(4) * s.num_slots)[slot];
        x[0] = a.x;
        x[1] = a.y;
        x[2] = a.z;
        x[3] = a.w;
        x[4] = b.x;
        d = b.y;
#else
#pragma unroll
        for (int k = 0; k < 5; ++k)
            x[k] = s.rng[k * s.num_slots + slot];
        d = s.rng[5 * s.num_slots + slot];
#endif
    }

    B2_D void store(StateView const& s, u32 slot) const
    {
#if B2_RNG_PACKED
        reinterpret_cast<uint4*>(s.rng)[slot] = make_uint4(x[0], x[1], x[2], x[3]);
        reinterpret_cast<uint2*>(s.rng + size_t(4) * s.num_slots)[slot] = make_uint2(x[4], d);
#else
#pragma unroll
        for (int k = 0; k < 5; ++k)
            s.rng[k * s.num_slots + slot] = x[k];
        s.rng[5 * s.num_slots + slot] = d;
#endif
    }

    B2_D void next()
    {
        u32 const t = (x[0] ^ (x[0] >> 2u));
        x[0] = x[1];
        x[1] = x[2];
        x[2] = x[3];
        x[3] = x[4];
        x[4] = (x[4] ^ (x[4] << 4u)) ^ (t ^ (t << 1u));
    }

    B2_D u32 operator()()
    {
        next();
        d += 362437u;
        return d + x[4];
    }

    //! Canonical double on [0, 1)
    B2_D real canonical()
    {
        u32 upper = (*this)();
        u32 lower = (*this)();
        constexpr double nrm = 1.1102230246251565e-16;  // 2^-53
        return nrm
               * static_cast<double>((static_cast<u64>(upper) << 21)
                                     ^ static_cast<u64>(lower));
    }

    //! Apply one jump polynomial (XorwowRngEngine.hh:286-304)
    B2_D void jump_poly(u32 const* poly)
    {
        u32 s[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < 5; ++i)
        {
            u32 w = poly[i];
            for (int j = 0; j < 32; ++j)
            {
                if (w & (1u << j))
                {
#pragma unroll
                    for (int k = 0; k < 5; ++k)
                        s[k] ^= x[k];
                }
                next();
            }
        }
#pragma unroll
        for (int k = 0; k < 5; ++k)
            x[k] = s[k];
    }

    //! Jump ahead `count` units using a table of 4^i-step polynomials
    B2_D void jump(u64 count, u32 const* poly_table)
    {
        u32 idx = 0;
        while (count > 0)
        {
            u32 n = static_cast<u32>(count) & 3u;
            for (u32 i = 0; i < n; ++i)
                jump_poly(poly_table + 5 * idx);
            ++idx;
            count >>= 2;
        }
    }

    //! Seed from (seed, subsequence, offset) (XorwowRngEngine.hh:196-216)
    B2_D void initialize(RngParams const& p, u32 seed, u64 subsequence, u64 offset)
    {
        u64 sm = seed;
        auto splitmix = [&sm]() {
            u64 z = (sm += 0x9e3779b97f4a7c15ull);
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            return z ^ (z >> 31);
        };
        u64 v = splitmix();
        x[0] = static_cast<u32>(v);
        x[1] = static_cast<u32>(v >> 32);
        v = splitmix();
        x[2] = static_cast<u32>(v);
        x[3] = static_cast<u32>(v >> 32);
        v = splitmix();
        x[4] = static_cast<u32>(v);
        d = static_cast<u32>(v >> 32);
        jump(subsequence, p.jump_subsequence);
        jump(offset, p.jump);
        d += static_cast<u32>(offset) * 362437u;
    }
};

//---------------------------------------------------------------------------//
// Distributions (reference: src/celeritas/random/distribution/*.hh)
//---------------------------------------------------------------------------//
//! Exponential with unit rate: -log(xi) (ExponentialDistribution.hh)
B2_D real sample_exponential(Rng& rng, real lambda = 1)
{
    real neg_inv = real(-1) / lambda;
    return log(rng.canonical()) * neg_inv;
}

B2_D bool sample_bernoulli(Rng& rng, real p_true)
{
    return rng.canonical() < p_true;
}

B2_D bool sample_bernoulli(Rng& rng, real scaled_true, real scaled_false)
{
    return rng.canonical() < scaled_true / (scaled_true + scaled_false);
}

//! Uniform on [a, b): fma(b - a, xi, a) (UniformRealDistribution.hh)
B2_D real sample_uniform(Rng& rng, real a, real b)
{
    real delta = b - a;
    return fma(delta, rng.canonical(), a);
}

//! Reciprocal 1/x on [a, b) (ReciprocalDistribution.hh)
struct ReciprocalDist
{
    real a;
    real logratio;
    B2_D ReciprocalDist(real a_, real b_) : a(a_), logratio(log((1 / a_) * b_)) {}
    // one-argument form of the reference: ReciprocalDistribution(a) == (1, a)
    B2_D explicit ReciprocalDist(real b_) : a(1), logratio(log((1 / real(1)) * b_)) {}
    B2_D real operator()(Rng& rng) const { return a * exp(logratio * rng.canonical()); }
};

//! Sample an exiting direction about `dir` with polar cosine costheta
//! (ExitingDirectionSampler, phys/InteractionUtils.hh)
B2_D Real3 sample_exiting_direction(Rng& rng, real costheta, Real3 const& dir)
{
    real phi = sample_uniform(rng, 0, 2 * constants::pi);
    return rotate(from_spherical(costheta, phi), dir);
}

//! Isotropic direction (IsotropicDistribution.hh)
B2_D Real3 sample_isotropic(Rng& rng)
{
    real costheta = sample_uniform(rng, -1, 1);
    real phi = sample_uniform(rng, 0, 2 * constants::pi);
    return from_spherical(costheta, phi);
}
}  // namespace b200
