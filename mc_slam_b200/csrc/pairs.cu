// pairs.cu -- builds, on the device, the per key-frame-pair lists that the Schur gather walks
// (for every pair of free key-frame blocks a <= b: the (edge_a, edge_b) of all map points seen by both).
// The host only ships the observation table; a 256-bit key-frame mask per map point turns the list
// construction into ballot compactions in point order, so the lists -- and with them the summation
// order of the Schur complement -- are deterministic.  Requires the observations of a point to be
// ordered by key-frame index (they are: MapPoint::GetObservations() is ordered by KeyFrame id,
// include/MapPoint.h:28, and the key-frame table is sorted by id).
#include "kernels.h"

namespace vilba {

constexpr int kMaskWords = kMaxKF / 64;  // 4

__device__ __forceinline__ bool mask_test(const unsigned long long* m, int k) { return (m[k >> 6] >> (k & 63)) & 1ull; }
__device__ __forceinline__ int mask_rank(const unsigned long long* m, int k) {  // set bits below position k
    int r = 0;
    const int wq = k >> 6;
#pragma unroll
    for (int i = 0; i < kMaskWords; ++i)
        if (i < wq) r += __popcll(m[i]);
    r += __popcll(m[wq] & ((1ull << (k & 63)) - 1ull));
    return r;
}

// one thread per map point: mask of all observing key-frames, mask of the free ones, edge -> point table
__global__ void __launch_bounds__(256) pair_masks_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < w.P; p += gridDim.x * blockDim.x) {
        unsigned long long all[kMaskWords] = {0, 0, 0, 0}, fr[kMaskWords] = {0, 0, 0, 0};
        for (int e = w.pt_obs_begin[p]; e < w.pt_obs_begin[p + 1]; ++e) {
            const int kf = w.obs[e].w & OBS_KF_MASK;
            all[kf >> 6] |= 1ull << (kf & 63);
            if (w.kf_block[kf] >= 0) fr[kf >> 6] |= 1ull << (kf & 63);
            w.edge_pt_rw[e] = p;
        }
#pragma unroll
        for (int i = 0; i < kMaskWords; ++i) {
            w.pt_mask[(size_t)p * 2 * kMaskWords + i] = all[i];
            w.pt_mask[(size_t)p * 2 * kMaskWords + kMaskWords + i] = fr[i];
        }
    }
}

// one warp per block pair; FILL = false counts, FILL = true writes the entries
template <bool FILL>
__global__ void __launch_bounds__(256) pair_lists_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    const int lane = threadIdx.x & 31;
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pair < w.n_pairs; pair += nw) {
        const int kfa = w.blk_kf[w.pair_a[pair]], kfb = w.blk_kf[w.pair_b[pair]];
        int running = FILL ? w.pair_begin_rw[pair] : 0;
        for (int base = 0; base < w.P; base += 32) {
            const int p = base + lane;
            bool has = false;
            const unsigned long long* m = w.pt_mask + (size_t)p * 2 * kMaskWords;
            if (p < w.P) has = mask_test(m + kMaskWords, kfa) && mask_test(m + kMaskWords, kfb);
            const unsigned bal = __ballot_sync(0xffffffffu, has);
            if (FILL && has) {
                const int pos = running + __popc(bal & ((1u << lane) - 1u));
                const int e0 = w.pt_obs_begin[p];
                w.pair_ea_rw[pos] = e0 + mask_rank(m, kfa);
                w.pair_eb_rw[pos] = e0 + mask_rank(m, kfb);
            }
            running += __popc(bal);
        }
        if (!FILL && lane == 0) w.pair_begin_rw[pair + 1] = running;  // counts, shifted by one for the scan
    }
}

// exclusive scan of the counts (single warp, chunks of 32)
__global__ void __launch_bounds__(32) pair_scan_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    const int lane = threadIdx.x;
    int carry = 0;
    if (lane == 0) w.pair_begin_rw[0] = 0;
    for (int base = 0; base < w.n_pairs; base += 32) {
        const int i = base + lane;
        int v = (i < w.n_pairs) ? w.pair_begin_rw[i + 1] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (i < w.n_pairs) w.pair_begin_rw[i + 1] = v + carry;
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
}

cudaError_t launch_build_pair_lists(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    pair_masks_kernel<<<dim3(d.point_grid, d.n_windows), 256, 0, s>>>(wp);
    pair_lists_kernel<false><<<dim3(d.point_grid, d.n_windows), 256, 0, s>>>(wp);
    pair_scan_kernel<<<dim3(1, d.n_windows), 32, 0, s>>>(wp);
    pair_lists_kernel<true><<<dim3(d.point_grid, d.n_windows), 256, 0, s>>>(wp);
    return cudaGetLastError();
}

}  // namespace vilba
