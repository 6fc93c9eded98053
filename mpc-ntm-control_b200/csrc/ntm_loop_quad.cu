// ntm_loop_quad.cu -- short-horizon (N <= 24) instantiations of the fused closed loop with four lanes per scenario
// (ntm_quad.cuh): closed_loop_quad_kernel<E, EXT>, E = ceil(N / 4) horizon indices per lane.
#include "ntm_loop.cuh"
#include "ntm_quad.cuh"

namespace ntm {

template <int E, int EXT>
static cudaError_t launch_quad(cudaStream_t st, const DeviceProps &dp, const LoopArgs &a) {
    const size_t region = (size_t)quad_region_doubles(a.N) * sizeof(double);
    // 32 scenarios (4 warps) per CTA when two such CTAs fit an SM, else fewer warps per CTA
    int wpb = 4;
    while (wpb > 1 && 2 * (region * 8 * wpb + 1024) > dp.smem_optin + 1024) wpb >>= 1;
    const size_t smem = region * 8 * wpb;
    if (smem > dp.smem_optin) return cudaErrorInvalidConfiguration;
    static thread_local size_t c_smem = 0;
    static thread_local int c_occ = 0, c_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (c_smem != smem || c_dev != dev) {
        cudaError_t e = raise_smem_attribute(reinterpret_cast<const void *>(closed_loop_quad_kernel<E, EXT>), dev, dp.smem_optin);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, closed_loop_quad_kernel<E, EXT>, 32 * wpb, smem);
        if (e != cudaSuccess) return e;
        c_smem = smem; c_occ = occ; c_dev = dev;
    }
    if (c_occ < 1) return cudaErrorInvalidConfiguration;
    const long long qpc = 8LL * wpb;
    long long grid = (long long)c_occ * dp.sm_count;
    const long long need = ((long long)a.S + qpc - 1) / qpc;
    if (grid > need) grid = need;
    if (grid * qpc > (long long)quad_max_groups(dp)) grid = quad_max_groups(dp) / qpc;      // one global LDL' slab per quad
    closed_loop_quad_kernel<E, EXT><<<(int)grid, 32 * wpb, smem, st>>>(a, (unsigned int)region);
    return cudaGetLastError();
}

cudaError_t launch_closed_loop_quad(cudaStream_t st, const DeviceProps &dp, const LoopArgs &a, long long *launches) {
    const bool rk4 = (a.flags & NTM_PROFILE_PLANT_RK4) != 0;
    const int E = (a.N + 3) / 4;
    cudaError_t e = cudaErrorInvalidValue;
#define NTM_QUAD_CASE(EV) case EV: e = rk4 ? launch_quad<EV, 1>(st, dp, a) : launch_quad<EV, 0>(st, dp, a); break;
    switch (E) {
        NTM_QUAD_CASE(1) NTM_QUAD_CASE(2) NTM_QUAD_CASE(3) NTM_QUAD_CASE(4) NTM_QUAD_CASE(5) NTM_QUAD_CASE(6)
        default: break;
    }
#undef NTM_QUAD_CASE
    if (e == cudaSuccess) ++*launches;
    return e;
}

}  // namespace ntm
