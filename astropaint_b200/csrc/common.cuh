// Shared host/device helpers for the astropaint_b200 kernels.
//
// Every header under csrc/ that holds arithmetic (healpix.cuh, qagi.cuh, profiles.cuh,
// spline.cuh) compiles both with nvcc for sm_100a (the product) and with g++ for the host
// (csrc/hostcheck.cpp -> libap_hostcheck.so, used ONLY by the CPU test-suite to validate the
// device arithmetic against the oracle / scipy without a GPU; the product never loads it).
#pragma once
#include <cstdint>
#include <cmath>
#include <cfloat>

#if defined(__CUDACC__)
#define AP_HD __host__ __device__ __forceinline__
#define AP_HD_NOINLINE inline __host__ __device__
#else
#define AP_HD inline
#define AP_HD_NOINLINE inline
#endif

// Round-to-nearest multiply/add/sub that the compiler may NOT contract into an FMA.  The
// reference stack (healpix_cxx, QUADPACK, numpy on generic x86-64) never fuses; the disc
// boundaries floor(nr/(2pi)*(phi+-dphi) - shift) must round identically.
#if defined(__CUDA_ARCH__)
#define AP_MUL(a, b) __dmul_rn((a), (b))
#define AP_ADD(a, b) __dadd_rn((a), (b))
#define AP_SUB(a, b) __dsub_rn((a), (b))
#else
// host builds use -ffp-contract=off
#define AP_MUL(a, b) ((a) * (b))
#define AP_ADD(a, b) ((a) + (b))
#define AP_SUB(a, b) ((a) - (b))
#endif

namespace ap {

constexpr double kPi = 3.141592653589793238462643383279502884197;
constexpr double kTwoPi = 6.283185307179586476925286766559005768394;
constexpr double kInvTwoPi = 1.0 / kTwoPi;
constexpr double kHalfPi = 1.570796326794896619231321691639751442099;
constexpr double kTwoThird = 2.0 / 3.0;

}  // namespace ap
