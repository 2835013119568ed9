// TEST INFRASTRUCTURE ONLY -- the reference's unmodified main.cxx with its own predict.hxx
// (host OpenMP), compiled from the sources where they lie, with the same fixed RNG seed as
// oracle/dropin_main.cxx so both binaries remove the same edges.  Output: oracle/_ref/ref_main.
#include <random>
#include "inc/main.hxx"
struct nlp_fixed_seed_device { unsigned operator()() const { return 12345u; } };
#define random_device nlp_fixed_seed_device
#include "main.cxx"
