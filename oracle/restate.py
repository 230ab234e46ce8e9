"""CPU restatement (pure Python / numpy) of the bit-exact pieces of the reference's track loop.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else). It restates, in plain
Python integers and IEEE doubles, the algorithms the CUDA kernels must reproduce bit for bit,
and is itself pinned against the golden vectors in the reference's own tests
(tests/test_cpu_oracle.py):

  * XORWOW engine, SplitMix64 seeding, jump-ahead
      /root/reference/src/celeritas/random/XorwowRngEngine.hh:163-314
      golden: test/celeritas/random/XorwowRngEngine.test.cc:141-186 (initial states)
  * host-side initial state fill: std::mt19937 seeded by std::seed_seq{seed[, stream]}
      /root/reference/src/celeritas/random/XorwowRngData.cc:28-58
  * canonical doubles from 32-bit engines
      /root/reference/src/celeritas/random/detail/GenerateCanonical32.hh:75-92 (XORWOW)
      libstdc++ std::generate_canonical<double, 53> (std::mt19937, used by the interactor
      tests: /root/reference/src/celeritas/random/distribution/GenerateCanonical.hh:61-67)
  * Klein-Nishina sampling
      /root/reference/src/celeritas/em/interactor/KleinNishinaInteractor.hh:104-190
      golden: test/celeritas/em/KleinNishina.test.cc:127-138
  * uniform-log value-grid interpolation
      /root/reference/src/celeritas/grid/XsCalculator.hh:107-153
      golden: test/celeritas/grid/XsCalculator.test.cc
  * RPN logic evaluation over surface senses
      /root/reference/src/orange/univ/detail/LogicEvaluator.hh, LogicStack.hh

The stronger checker, used for whole-loop parity, is oracle/_ref (the reference itself).
"""
import math

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


# --------------------------------------------------------------------------- #
# std::mt19937 / std::seed_seq
# --------------------------------------------------------------------------- #
class Mt19937:
    """std::mt19937 (32-bit Mersenne Twister)."""

    def __init__(self, seed=5489, state=None):
        if state is not None:
            self.mt = list(state)
        else:
            mt = [seed & M32]
            for i in range(1, 624):
                mt.append((1812433253 * (mt[-1] ^ (mt[-1] >> 30)) + i) & M32)
            self.mt = mt
        self.idx = 624

    @classmethod
    def from_seed_seq(cls, seeds):
        return cls(state=seed_seq_generate(seeds, 624))

    def _twist(self):
        mt = self.mt
        for i in range(624):
            y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
            v = mt[(i + 397) % 624] ^ (y >> 1)
            if y & 1:
                v ^= 0x9908B0DF
            mt[i] = v
        self.idx = 0

    def __call__(self):
        if self.idx >= 624:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & M32


def seed_seq_generate(seeds, n):
    """std::seed_seq::generate (C++11 [rand.util.seedseq])."""
    v = [s & M32 for s in seeds]
    s = len(v)
    b = [0x8B8B8B8B] * n
    t = 11 if n >= 623 else 7 if n >= 68 else 5 if n >= 39 else 3 if n >= 7 else (n - 1) // 2
    p = (n - t) // 2
    q = p + t
    m = max(s + 1, n)

    def T(x):
        return x ^ (x >> 27)

    for k in range(m):
        r1 = (1664525 * T(b[k % n] ^ b[(k + p) % n] ^ b[(k - 1) % n])) & M32
        if k == 0:
            r2 = (r1 + s) & M32
        elif k <= s:
            r2 = (r1 + (k % n) + v[k - 1]) & M32
        else:
            r2 = (r1 + (k % n)) & M32
        b[(k + p) % n] = (b[(k + p) % n] + r1) & M32
        b[(k + q) % n] = (b[(k + q) % n] + r2) & M32
        b[k % n] = r2
    for k in range(m, m + n):
        r3 = (1566083941 * T((b[k % n] + b[(k + p) % n] + b[(k - 1) % n]) & M32)) & M32
        r4 = (r3 - (k % n)) & M32
        b[(k + p) % n] ^= r3
        b[(k + q) % n] ^= r4
        b[k % n] = r4
    return b


def canonical_std(rng):
    """libstdc++ std::generate_canonical<double, 53> over a 32-bit engine: two draws."""
    lo = rng()
    hi = rng()
    ret = (lo + hi * 4294967296.0) / 18446744073709551616.0
    if ret >= 1.0:
        ret = math.nextafter(1.0, 0.0)
    return ret


def generate_primaries(options, particle_ids):
    """PrimaryGenerator::operator() over all events (phys/PrimaryGenerator.cc:84-108,
    PrimaryGeneratorOptions.cc:74-140): one mt19937(seed); per primary the position (box:
    x, y, z uniform) and then the direction (isotropic: costheta in [-1, 1), phi in
    [0, 2pi)) are sampled; particle = pdg[i % len(pdg)]. Returns a list of dicts."""
    rng = Mt19937(options.get('seed', 0))

    def uniform(a, b):
        # UniformRealDistribution.hh:71-77: fma(b - a, canonical, a)
        return _fma_exact(b - a, canonical_std(rng), a)

    def spec(value, scalar=False):
        if isinstance(value, dict):
            return value['distribution'], list(value.get('params', []))
        return 'delta', [value] if scalar else list(value)

    (edist, eparams) = spec(options['energy'], scalar=True)
    (pdist, pparams) = spec(options['position'])
    (ddist, dparams) = spec(options['direction'])
    assert edist == 'delta' and pdist in ('delta', 'box') and ddist in ('delta', 'isotropic')
    out = []
    for event in range(options['num_events']):
        for i in range(options['primaries_per_event']):
            if pdist == 'delta':
                pos = list(pparams)
            else:
                pos = [uniform(pparams[k], pparams[3 + k]) for k in range(3)]
            if ddist == 'delta':
                direction = list(dparams)
            else:
                costheta = uniform(-1.0, 1.0)
                phi = uniform(0.0, 2 * math.pi)
                direction = from_spherical(costheta, phi)
            out.append(dict(particle_id=particle_ids[i % len(particle_ids)], event_id=event,
                            energy=eparams[0], pos=pos, dir=direction, time=0.0))
    return out


def initial_xorwow_states(seed, stream, n):
    """initialize_xorwow: n states of 6 words from mt19937(seed_seq{seed[, stream]})."""
    seeds = [seed] if stream == 0 else [seed, stream]
    rng = Mt19937.from_seed_seq(seeds)
    # std::uniform_int_distribution<uint32_t> over the full range returns the raw draw
    return [[rng() for _ in range(6)] for _ in range(n)]


# --------------------------------------------------------------------------- #
# XORWOW
# --------------------------------------------------------------------------- #
class Xorwow:
    def __init__(self, state):
        self.x = list(state[:5])
        self.d = state[5]

    def next(self):
        x = self.x
        t = x[0] ^ (x[0] >> 2)
        x[0], x[1], x[2], x[3] = x[1], x[2], x[3], x[4]
        x[4] = ((x[4] ^ ((x[4] << 4) & M32)) ^ (t ^ ((t << 1) & M32))) & M32

    def __call__(self):
        self.next()
        self.d = (self.d + 362437) & M32
        return (self.d + self.x[4]) & M32

    def canonical(self):
        """GenerateCanonical32<double>: (hi << 21 ^ lo) * 2^-53."""
        upper = self()
        lower = self()
        return 1.1102230246251565e-16 * float((upper << 21) ^ lower)

    def state(self):
        return self.x + [self.d]

    def jump_poly(self, poly):
        s = [0] * 5
        for i in range(5):
            for j in range(32):
                if poly[i] & (1 << j):
                    for k in range(5):
                        s[k] ^= self.x[k]
                self.next()
        self.x = s

    def jump(self, count, table):
        idx = 0
        while count > 0:
            for _ in range(count & 3):
                self.jump_poly(table[idx])
            idx += 1
            count >>= 2

    def discard(self, count, jump_table):
        self.jump(count, jump_table)
        self.d = (self.d + (count & M32) * 362437) & M32

    @classmethod
    def from_seed(cls, seed, subsequence, offset, jump_table, jump_sub_table):
        st = seed & M64

        def splitmix():
            nonlocal st
            st = (st + 0x9E3779B97F4A7C15) & M64
            z = st
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
            return z ^ (z >> 31)

        a, b, c = splitmix(), splitmix(), splitmix()
        r = cls([a & M32, a >> 32, b & M32, b >> 32, c & M32, c >> 32])
        r.jump(subsequence, jump_sub_table)
        r.discard(offset, jump_table)
        return r


# --------------------------------------------------------------------------- #
# Small-vector helpers with the reference's fma usage
# --------------------------------------------------------------------------- #
def fma(a, b, c):
    return math.fma(a, b, c) if hasattr(math, 'fma') else _fma_exact(a, b, c)


def _fma_exact(a, b, c):
    from fractions import Fraction
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def dot(a, b):
    r = 0.0
    for x, y in zip(a, b):
        r = fma(x, y, r)
    return r


def make_unit_vector(v):
    s = 1 / math.sqrt(dot(v, v))
    return [x * s for x in v]


def from_spherical(costheta, phi):
    st = math.sqrt(1 - costheta * costheta)
    return [st * math.cos(phi), st * math.sin(phi), costheta]


def rotate(d, rot):
    sintheta = math.sqrt(1 - rot[2] ** 2)
    if sintheta >= 0.005:
        inv = 1 / sintheta
        cosphi, sinphi = rot[0] * inv, rot[1] * inv
    elif sintheta > 0:
        cosphi = rot[0] / math.sqrt(rot[0] ** 2 + rot[1] ** 2)
        sinphi = math.sqrt(1 - cosphi ** 2)
    else:
        cosphi, sinphi = 1.0, 0.0
    r = [(rot[2] * d[0] + sintheta * d[2]) * cosphi - sinphi * d[1],
         (rot[2] * d[0] + sintheta * d[2]) * sinphi + cosphi * d[1],
         -sintheta * d[0] + rot[2] * d[2]]
    return make_unit_vector(r)


# --------------------------------------------------------------------------- #
# Klein-Nishina
# --------------------------------------------------------------------------- #
def klein_nishina(inc_energy, inc_dir, inv_electron_mass, canonical):
    """Returns (energy, direction, electron_energy or None, electron_direction or None,
    energy_deposition)."""
    k = inc_energy * inv_electron_mass
    eps0 = 1 / (1 + 2 * k)
    f1 = -math.log(eps0)
    f2 = 0.5 * (1 - eps0 ** 2)
    p_f1 = f1 / (f1 + f2)
    logratio = math.log((1 / 1.0) * eps0)
    eps0_sq = eps0 ** 2
    while True:
        if canonical() < p_f1:
            eps = 1.0 * math.exp(logratio * canonical())
            eps_sq = eps * eps
        else:
            eps_sq = fma(1 - eps0_sq, canonical(), eps0_sq)
            eps = math.sqrt(eps_sq)
        omc = (1 - eps) / (eps * k)
        sin_sq = omc * (2 - omc)
        reject = eps * sin_sq / (1 + eps_sq)
        if not (canonical() < reject):
            break
    energy = eps * inc_energy
    phi = fma(2 * math.pi - 0, canonical(), 0.0)
    direction = rotate(from_spherical(1 - omc, phi), inc_dir)
    e_electron = inc_energy - energy
    if e_electron < 1e-4:
        return energy, direction, None, None, e_electron
    r = [inc_dir[i] * inc_energy - direction[i] * energy for i in range(3)]
    return energy, direction, e_electron, make_unit_vector(r), 0.0


# --------------------------------------------------------------------------- #
# Value-grid interpolation
# --------------------------------------------------------------------------- #
def calc_xs(log_front, log_back, values, prime_index, energy):
    """XsCalculator::operator(): uniform grid in log E, linear in E, E-scaled above prime.

    The grid is UniformGridData::from_bounds(front, back, size): delta = (back - front)/(n-1).
    """
    n = len(values)
    loge = math.log(energy)
    back = log_back
    log_delta = (log_back - log_front) / (n - 1)

    def extrap(i):
        r = values[i]
        if i >= prime_index:
            r /= energy
        return r

    if loge <= log_front:
        return extrap(0)
    if loge >= back:
        return extrap(n - 1)
    lo = int((loge - log_front) / log_delta)
    e_hi = math.exp(log_front + log_delta * (lo + 1))
    x_hi = values[lo + 1]
    if lo + 1 == prime_index:
        x_hi /= e_hi
    e_lo = math.exp(log_front + log_delta * lo)
    slope = (x_hi - values[lo]) / (e_hi - e_lo)
    r = fma(slope, energy - e_lo, values[lo])
    if lo >= prime_index:
        r /= energy
    return r


def calc_range(log_front, log_back, values, energy):
    """RangeCalculator::operator() (/root/reference/src/celeritas/grid/RangeCalculator.hh:
    79-108): scaled by sqrt(E/Emin) below the grid, clipped above, linear in E between nodes."""
    n = len(values)
    loge = math.log(energy)
    log_delta = (log_back - log_front) / (n - 1)
    if loge <= log_front:
        return values[0] * math.exp(0.5 * (loge - log_front))
    if loge >= log_back:
        return values[n - 1]
    lo = int((loge - log_front) / log_delta)
    e_lo = math.exp(log_front + log_delta * lo)
    e_hi = math.exp(log_front + log_delta * (lo + 1))
    slope = (values[lo + 1] - values[lo]) / (e_hi - e_lo)
    return fma(slope, energy - e_lo, values[lo])


def calc_inverse_range(log_front, log_back, ranges, rng):
    """InverseRangeCalculator::operator() (grid/InverseRangeCalculator.hh:86-116): energy of a
    particle with the given range; E = Emin (r / r0)^2 below the first node."""
    n = len(ranges)
    log_delta = (log_back - log_front) / (n - 1)
    assert 0 <= rng <= ranges[-1]
    if rng < ranges[0]:
        return math.exp(log_front) * (rng / ranges[0]) ** 2
    if rng >= ranges[-1]:
        return math.exp(log_back)
    # NonuniformGrid::find: the last node not above the value
    lo = max(i for i in range(n) if ranges[i] <= rng)
    e_lo = math.exp(log_front + log_delta * lo)
    e_hi = math.exp(log_front + log_delta * (lo + 1))
    slope = (e_hi - e_lo) / (ranges[lo + 1] - ranges[lo])
    return fma(slope, rng - ranges[lo], e_lo)


# --------------------------------------------------------------------------- #
# ORANGE logic
# --------------------------------------------------------------------------- #
LOGIC_TRUE, LOGIC_OR, LOGIC_AND, LOGIC_NOT = 0xFFFB, 0xFFFC, 0xFFFD, 0xFFFE


def eval_logic(tokens, senses):
    """RPN boolean logic over face senses (True = outside)."""
    stack = []
    for t in tokens:
        if t < 0xFFF9:
            stack.append(bool(senses[t]))
        elif t == LOGIC_TRUE:
            stack.append(True)
        elif t == LOGIC_OR:
            b, a = stack.pop(), stack.pop()
            stack.append(a or b)
        elif t == LOGIC_AND:
            b, a = stack.pop(), stack.pop()
            stack.append(a and b)
        elif t == LOGIC_NOT:
            stack.append(not stack.pop())
    return stack[-1]


def parse_logic(text):
    """'0 1 ~ & 2 &' (the .org.json form) -> token list."""
    m = {'*': LOGIC_TRUE, '|': LOGIC_OR, '&': LOGIC_AND, '~': LOGIC_NOT}
    return [m[t] if t in m else int(t) for t in text.split()]
