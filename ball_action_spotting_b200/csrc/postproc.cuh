// Post-processing of the raw per-frame probabilities into action spots (src/utils.py:55-64 `post_processing`:
// scipy.ndimage.gaussian_filter(sigma) -> scipy.signal.find_peaks(height, distance)), one CTA per class.
// Index work is exact; the smoothing mirrors scipy's arithmetic (double accumulation, symmetric-kernel summation order
// of NI_Correlate1D, 'reflect' boundary) so that the float32 confidences agree bit for bit.
#pragma once
#include "common.cuh"

namespace mds {

constexpr int kPostMaxRadius = 64;
struct PostParams {
    const float* raw;        // [N][K] probabilities (row = frame, column = class)
    float* smooth;           // [K][N] scratch: filtered series
    unsigned char* state;    // [K][N] scratch: 0 none, 1 candidate peak (undecided), 2 kept, 3 removed
    int* out_index;          // [K][N] positions of the kept peaks, ascending (first *count entries)
    float* out_conf;         // [K][N] smoothed value at those positions
    int* out_count;          // [K]
    double w[2 * kPostMaxRadius + 1];   // gaussian weights, w[radius] is the centre
    int N, K, radius, distance;
    float height;
};

// scipy 'reflect' (half-sample symmetric): d c b a | a b c d | d c b a
__device__ __forceinline__ int reflect_index(int i, int n) {
    if (n == 1) return 0;
    const int period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - 1 - i;
}

__global__ void __launch_bounds__(1024) post_processing_kernel(PostParams p) {
    __shared__ int s_scan[32];
    __shared__ int s_base;
    const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N;
    float* sm = p.smooth + (size_t)k * N;
    unsigned char* st = p.state + (size_t)k * N;

    // 1. gaussian_filter: correlate1d with a symmetric kernel
    for (int i = tid; i < N; i += 1024) {
        double acc = (double)p.raw[(size_t)i * p.K + k] * p.w[p.radius];
        for (int j = -p.radius; j < 0; ++j) {
            const double a = (double)p.raw[(size_t)reflect_index(i + j, N) * p.K + k];
            const double b = (double)p.raw[(size_t)reflect_index(i - j, N) * p.K + k];
            acc += (a + b) * p.w[p.radius + j];
        }
        sm[i] = (float)acc;
        st[i] = 0;
    }
    __syncthreads();

    // 2. local maxima (scipy _local_maxima_1d: plateaus yield their midpoint; the borders are never peaks), height filter
    for (int i = tid; i < N; i += 1024) {
        if (i == 0 || i >= N - 1) continue;
        const float v = sm[i];
        if (!(sm[i - 1] < v)) continue;
        int ahead = i + 1;
        while (ahead < N - 1 && sm[ahead] == v) ++ahead;
        if (sm[ahead] < v) {
            const int mid = (i + ahead - 1) / 2;
            if (sm[mid] >= p.height) st[mid] = 1;
        }
    }
    __syncthreads();

    // 3. minimal distance (scipy _select_by_peak_distance): greedy by decreasing height, evaluated in parallel rounds --
    //    a candidate with no undecided higher candidate closer than `distance` is kept, and removes its neighbours.
    //    Priority ties are broken towards the larger index (a stable argsort walked from its end).
    if (p.distance > 1) {
        while (true) {
            bool pending = false;
            for (int i = tid; i < N; i += 1024) {
                if (st[i] != 1) continue;
                const float v = sm[i];
                bool top = true;
                for (int d = 1; d < p.distance && top; ++d) {
                    const int l = i - d, r = i + d;
                    // a neighbour already marked 4 this round out-ranks this candidate by construction: it still blocks
                    if (l >= 0 && (st[l] == 1 || st[l] == 4) && sm[l] > v) top = false;
                    if (r < N && (st[r] == 1 || st[r] == 4) && sm[r] >= v) top = false;
                }
                if (top) st[i] = 4;         // kept this round (marked apart so that the scan above stays consistent)
                else pending = true;
            }
            __syncthreads();
            for (int i = tid; i < N; i += 1024) {
                if (st[i] != 1) continue;
                bool removed = false;
                for (int d = 1; d < p.distance && !removed; ++d) {
                    const int l = i - d, r = i + d;
                    if ((l >= 0 && st[l] == 4) || (r < N && st[r] == 4)) removed = true;
                }
                if (removed) st[i] = 3;
            }
            __syncthreads();
            for (int i = tid; i < N; i += 1024)
                if (st[i] == 4) st[i] = 2;
            if (!__syncthreads_or(pending)) break;
        }
    } else {
        for (int i = tid; i < N; i += 1024)
            if (st[i] == 1) st[i] = 2;
        __syncthreads();
    }

    // 4. ordered compaction of the kept peaks
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < N; i0 += 1024) {
        const int i = i0 + tid;
        const int flag = (i < N && st[i] == 2) ? 1 : 0;
        int incl = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int v = s_scan[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            s_scan[lane] = v;
        }
        __syncthreads();
        const int base = s_base + (warp > 0 ? s_scan[warp - 1] : 0);
        if (flag) {
            p.out_index[(size_t)k * N + base + incl - 1] = i;
            p.out_conf[(size_t)k * N + base + incl - 1] = sm[i];
        }
        __syncthreads();
        if (tid == 0) s_base += s_scan[31];
        __syncthreads();
    }
    if (tid == 0) p.out_count[k] = s_base;
}

}  // namespace mds
