// Random-number distributions used by the MCMC host.
//
// The reference draws every proposal / accept decision from one process-wide
// std::mt19937 through Boost.Random distribution classes
// (reference: scicone/src/SingletonRandomGenerator.h:21-47 and the call sites
// listed in SURVEY.md App. B).  Boost is an un-vendored, un-pinned third-party
// dependency of the reference (scicone/CMakeLists.txt:39, any version >= 1.47),
// so this file restates the *published* Boost.Random algorithms:
//
//   uniform_int_distribution   bucketed rejection on one 32-bit engine word;
//                              returns min with NO draw when max == min;
//                              returns the raw engine word when the unsigned
//                              range equals the engine range (e.g. (0,-1)).
//   uniform_real / uniform_01  one 32-bit word times 2^-32 (retry if == max).
//   bernoulli_distribution     one word compared with p * (2^32 - 1).
//   poisson_distribution       table-free inversion for mean < 10.
//   discrete_distribution      Walker alias table (below/above average
//                              stacks), sample = uniform_int(0,n-1) then
//                              uniform_01.
//   normal_distribution        Marsaglia polar method on 2-word canonical
//                              uniforms (Boost uses a ziggurat whose tables
//                              are not restated here; see the class comment).
//
// The oracle build of the unmodified reference (oracle/Makefile) puts thin
// boost/random/*.hpp wrappers around THIS file on its include path, so the
// reference host and the scicone_b200 host consume the engine identically:
// same seed => same proposal stream.  That is the definition of trajectory
// parity used by the tests.
#ifndef SCICONE_B200_RNG_DIST_HPP
#define SCICONE_B200_RNG_DIST_HPP

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <iterator>
#include <numeric>
#include <stdexcept>
#include <utility>
#include <vector>

namespace scicone_rng {

// Engine contract: result_type is 32-bit unsigned, min()==0, max()==2^32-1
// (std::mt19937).
template <class Engine>
inline uint32_t engine_word(Engine& eng) {
    return static_cast<uint32_t>(eng() - (Engine::min)());
}

template <class Real, class Engine>
inline Real uniform_01_draw(Engine& eng) {
    const Real factor = Real(1) / (Real(4294967295.0) + Real(1));
    for (;;) {
        Real result = Real(engine_word(eng)) * factor;
        if (result < Real(1)) return result;
    }
}

template <class IntType = int>
class uniform_int_distribution {
public:
    typedef IntType result_type;
    explicit uniform_int_distribution(IntType mn = 0, IntType mx = 9) : min_(mn), max_(mx) {}
    IntType(min)() const { return min_; }
    IntType(max)() const { return max_; }

    template <class Engine>
    IntType operator()(Engine& eng) const {
        const uint32_t range = static_cast<uint32_t>(max_) - static_cast<uint32_t>(min_);
        const uint32_t brange = 0xFFFFFFFFu;
        if (range == 0) return min_;
        if (range == brange)
            return static_cast<IntType>(engine_word(eng) + static_cast<uint32_t>(min_));
        uint32_t bucket_size = brange / (range + 1u);
        if (brange % (range + 1u) == range) ++bucket_size;
        for (;;) {
            uint32_t result = engine_word(eng) / bucket_size;
            if (result <= range)
                return static_cast<IntType>(result + static_cast<uint32_t>(min_));
        }
    }

private:
    IntType min_, max_;
};

template <class Real = double>
class uniform_real_distribution {
public:
    typedef Real result_type;
    explicit uniform_real_distribution(Real mn = 0, Real mx = 1) : min_(mn), max_(mx) {}
    template <class Engine>
    Real operator()(Engine& eng) const {
        for (;;) {
            Real numerator = static_cast<Real>(engine_word(eng));
            Real divisor = static_cast<Real>(4294967295.0) + 1;
            Real result = numerator / divisor * (max_ - min_) + min_;
            if (result < max_) return result;
        }
    }

private:
    Real min_, max_;
};

template <class Real = double>
class bernoulli_distribution {
public:
    typedef bool result_type;
    explicit bernoulli_distribution(Real p = 0.5) : p_(p) {}
    template <class Engine>
    bool operator()(Engine& eng) const {
        if (p_ == Real(0)) return false;
        return Real(engine_word(eng)) <= p_ * Real(4294967295.0);
    }

private:
    Real p_;
};

template <class IntType = int, class Real = double>
class poisson_distribution {
public:
    typedef IntType result_type;
    explicit poisson_distribution(Real mean = 1) : mean_(mean), exp_mean_(std::exp(-mean)) {
        if (!(mean < 10))
            throw std::domain_error("scicone_rng::poisson_distribution: only mean < 10 (inversion) is implemented");
    }
    template <class Engine>
    IntType operator()(Engine& eng) const {
        Real p = exp_mean_;
        IntType x = 0;
        Real u = uniform_01_draw<Real>(eng);
        while (u > p) {
            u = u - p;
            ++x;
            p = mean_ * p / x;
        }
        return x;
    }

private:
    Real mean_, exp_mean_;
};

template <class IntType = int, class Weight = double>
class discrete_distribution {
public:
    typedef IntType result_type;
    discrete_distribution() { init_empty(); }
    template <class Iter>
    discrete_distribution(Iter first, Iter last) {
        if (first == last) {
            init_empty();
            return;
        }
        const std::size_t n = static_cast<std::size_t>(std::distance(first, last));
        std::vector<std::pair<Weight, IntType> > below, above;
        below.reserve(n);
        above.reserve(n);
        Weight weight_sum = std::accumulate(first, last, static_cast<Weight>(0));
        table_.assign(n, std::pair<Weight, IntType>(Weight(0), IntType(0)));
        const Weight weight_average = weight_sum / static_cast<Weight>(n);
        const Weight normalized_average = Weight(1);
        std::size_t i = 0;
        for (; first != last; ++first, ++i) {
            Weight val = *first / weight_average;
            std::pair<Weight, IntType> elem(val, static_cast<IntType>(i));
            if (val < normalized_average)
                below.push_back(elem);
            else
                above.push_back(elem);
        }
        typename std::vector<std::pair<Weight, IntType> >::iterator b = below.begin(), be = below.end(),
                                                                    a = above.begin(), ae = above.end();
        while (b != be && a != ae) {
            table_[static_cast<std::size_t>(b->second)] = std::make_pair(b->first, a->second);
            a->first -= (Weight(1) - b->first);
            if (a->first < normalized_average)
                *b = *a++;
            else
                ++b;
        }
        for (; b != be; ++b) table_[static_cast<std::size_t>(b->second)].first = Weight(1);
        for (; a != ae; ++a) table_[static_cast<std::size_t>(a->second)].first = Weight(1);
    }

    template <class Engine>
    IntType operator()(Engine& eng) const {
        IntType result =
            uniform_int_distribution<IntType>(0, static_cast<IntType>(table_.size()) - 1)(eng);
        Weight test = uniform_01_draw<Weight>(eng);
        if (test < table_[static_cast<std::size_t>(result)].first) return result;
        return table_[static_cast<std::size_t>(result)].second;
    }

private:
    void init_empty() {
        table_.clear();
        table_.push_back(std::make_pair(Weight(1), IntType(0)));
    }
    std::vector<std::pair<Weight, IntType> > table_;
};

template <class Real = double>
class normal_distribution {
public:
    typedef Real result_type;
    explicit normal_distribution(Real mean = 0, Real sigma = 1) : mean_(mean), sigma_(sigma) {}
    // Marsaglia polar method fed by 53-bit canonical uniforms built from two
    // engine words each (low word first); the second coordinate's variate is
    // returned and the first discarded, i.e. a freshly constructed
    // distribution object per draw, as at the reference call site
    // (scicone/src/Inference.h:925).  With this draw pattern the reference's own
    // test_reproducibility golden values (34.165 / 41.046,
    // scicone/tests/validation.h:174,196) are reproduced by the oracle build.
    template <class Engine>
    Real operator()(Engine& eng) const {
        Real x, y, r2;
        do {
            x = Real(2) * canonical(eng) - Real(1);
            y = Real(2) * canonical(eng) - Real(1);
            r2 = x * x + y * y;
        } while (r2 > Real(1) || r2 == Real(0));
        const Real mult = std::sqrt(Real(-2) * std::log(r2) / r2);
        return y * mult * sigma_ + mean_;
    }

private:
    template <class Engine>
    static Real canonical(Engine& eng) {
        Real sum = Real(0), tmp = Real(1);
        for (int k = 0; k < 2; ++k) {
            sum += Real(engine_word(eng)) * tmp;
            tmp *= Real(4294967296.0);
        }
        Real ret = sum / tmp;
        if (ret >= Real(1)) ret = std::nextafter(Real(1), Real(0));
        return ret;
    }
    Real mean_, sigma_;
};

}  // namespace scicone_rng

#endif  // SCICONE_B200_RNG_DIST_HPP
