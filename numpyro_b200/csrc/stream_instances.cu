// One translation unit per tile width: compiled five times with -DB2_INST_KS=1|2|4|7|8 (numpyro_b200/build.py builds them
// in parallel) so that the 30 instances of the streaming kernel do not serialise the build in a single nvcc run.
// Only host-side function pointers cross translation units (cudaLaunchCooperativeKernel takes them), no device linking.
#define B2_NO_WIDE 1          // (see for_d_wide in common.cuh)
#include "stream_engine.cuh"

#ifndef B2_INST_KS
#error "compile with -DB2_INST_KS=<1|2|4|7|8>"
#endif
#define B2_CAT2(a, b) a##b
#define B2_CAT(a, b) B2_CAT2(a, b)

using namespace b2;

// lik: LIK_* of families.cuh; mg: more than one chain group
const void* B2_CAT(b2_stream_kernel_ks, B2_INST_KS)(int lik, bool mg) {
    constexpr int KS = B2_INST_KS;
    if (mg)
        return lik == LIK_BERNOULLI ? (const void*)stream_engine_kernel<KS, LIK_BERNOULLI, true>
             : lik == LIK_POISSON   ? (const void*)stream_engine_kernel<KS, LIK_POISSON, true>
                                    : (const void*)stream_engine_kernel<KS, LIK_NORMAL, true>;
    return lik == LIK_BERNOULLI ? (const void*)stream_engine_kernel<KS, LIK_BERNOULLI, false>
         : lik == LIK_POISSON   ? (const void*)stream_engine_kernel<KS, LIK_POISSON, false>
                                : (const void*)stream_engine_kernel<KS, LIK_NORMAL, false>;
}
