// Oracle build shim: maps boost::random::poisson_distribution onto scicone_rng::poisson_distribution so that the
// unmodified reference host consumes the mt19937 stream exactly like the
// scicone_b200 host (see scicone_b200/host/rng_dist.hpp).  Test infrastructure only.
#ifndef ORACLE_SHIM_BOOST_RANDOM_POISSON_DISTRIBUTION_HPP
#define ORACLE_SHIM_BOOST_RANDOM_POISSON_DISTRIBUTION_HPP
#include "rng_dist.hpp"
namespace boost { namespace random { using scicone_rng::poisson_distribution; } }
#endif
