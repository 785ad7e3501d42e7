// synth_kernels.cuh — device twins of rmdetect_b200/synth.py (no reference counterpart).
// Synthetic genome letters and the stand-in bpp matrices used by benchmarks where ViennaRNA output
// would otherwise be needed; integer hashing only, so host (numpy), oracle (C) and device agree bit
// for bit.  Sequences produced here are pure ACGU.
#pragma once
#include "device_common.cuh"
#include "calib_kernels.cuh"   // mix64_dev

#define SYN_THREADS 128
#define SYN_GOLDEN 0x9E3779B97F4A7C15ULL

// letter at stream position g: top two bits of mix64(seed + GOLDEN * (g + 1)).  rev_total > 0 reads the reverse
// strand of a stream of rev_total letters (lib/sequences.py:138-143: reversed, A<->U, C<->G): position k of the
// strand is 3 - letter(rev_total - 1 - k).
__device__ __forceinline__ uint32_t synth_letter(unsigned long long seed, long long k, long long rev_total) {
    const long long g = rev_total > 0 ? rev_total - 1 - k : k;
    const uint32_t c = (uint32_t)(mix64_dev(seed + SYN_GOLDEN * (unsigned long long)(g + 1)) >> 62);
    return rev_total > 0 ? 3u - c : c;
}

__global__ void synth_codes_kernel(unsigned long long seed, long long first, long long n, long long rev_total, uint32_t *codes2) {
    const long long wi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long base = wi * 16;
    if (base >= n) return;
    uint32_t word = 0;
    for (int k = 0; k < 16 && base + k < n; k++) word |= synth_letter(seed, first + base + k, rev_total) << (2 * k);
    codes2[wi] = word;
}

// letter counts of strand positions [first, first + n): nothing is stored (GC table of a strand that is scanned in pieces)
__global__ void __launch_bounds__(256) synth_counts_kernel(unsigned long long seed, long long first, long long n, long long rev_total,
                                                           unsigned long long *counts /* [4] */) {
    __shared__ unsigned int local[4];
    if (threadIdx.x < 4) local[threadIdx.x] = 0;
    __syncthreads();
    unsigned int c4[4] = {0, 0, 0, 0};
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
        c4[synth_letter(seed, first + k, rev_total)]++;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        unsigned int v = c4[c];
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(FULL_MASK, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&local[c], v);
    }
    __syncthreads();
    if (threadIdx.x < 4 && local[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)local[threadIdx.x]);
}

__device__ __forceinline__ bool can_pair_dev(int a, int b) {
    const int s = a + b;
    return a != b && (s == 3 || (s == 5 && (a < b ? a : b) == 2));
}

__device__ __forceinline__ unsigned long long cell_hash_dev(unsigned long long key, int i, int j, int salt) {
    const unsigned long long v = ((unsigned long long)i << 34) | ((unsigned long long)j << 4) | (unsigned long long)salt;
    return mix64_dev(key + SYN_GOLDEN * (v + 1));
}

// value of cell (i, j) in units of 1e-9 of sqrt(p); 0 = no entry
__device__ __forceinline__ long long synth_cell(const uint8_t *lt, int n, unsigned long long key, int i, int j) {
    long long best = 0;
    for (int t = 0; t < 8; t++) {
        const int si = i - t, sj = j + t;
        if (si < 0 || sj >= n) break;
        if (!can_pair_dev(lt[si], lt[sj])) break;
        const unsigned long long h0 = cell_hash_dev(key, si, sj, 0);
        if ((h0 >> 40) >= (31u << 16)) continue;
        const int len = (0x8765443332221111ULL >> (4 * ((h0 >> 8) & 15))) & 15;   // 1,1,1,1,2,2,2,3,3,3,4,4,5,6,7,8
        if (len <= t) continue;
        const long long bu = (h0 >> 16) & 0xFFFF;
        long long lo, hi;
        if (bu < 49152) { lo = 3162278; hi = 31622777; }
        else if (bu < 60000) { lo = 31622777; hi = 316227767; }
        else { lo = 316227767; hi = 1000000001; }
        const unsigned long long h1 = cell_hash_dev(key, si, sj, 1);
        const long long base = lo + (long long)(h1 % (unsigned long long)(hi - lo));
        const unsigned long long h2 = cell_hash_dev(key, i, j, 2 + t);
        const long long half = base >> 1;
        long long val = half + ((half * (long long)(h2 & 1023)) >> 10);
        if (val < 3162278) val = 3162278;
        if (val > best) best = val;
    }
    return best;
}

// One CTA per window.  FILL == false: count entries (bpp_cnt[w]).  FILL == true: write them, sorted.
template <bool FILL>
__global__ void __launch_bounds__(SYN_THREADS)
synth_bpp_kernel(const DevJob J, unsigned long long seed, uint32_t *bpp_cnt, const long long *bpp_off,
                 uint16_t *oi, uint16_t *oj, double *op) {
    __shared__ uint8_t lt[FAST_W];
    __shared__ unsigned long long s_key;
    __shared__ uint32_t tmp[SYN_THREADS / 32 + 1];
    const long long w = blockIdx.x;
    int seq, n;
    long long start;
    window_geom(J, w, seq, start, n);
    if (n > FAST_W) n = FAST_W;                      // synthetic jobs use windows <= 160 (checked on the host)
    const long long gbase = J.seq_off[seq] + start;
    for (int k = threadIdx.x; k < n; k += SYN_THREADS) lt[k] = (uint8_t)get_code(J, gbase + k);
    __syncthreads();
    if (threadIdx.x == 0) {                           // window key: chained mix over 8-letter words
        unsigned long long key = seed;
        for (int b = 0; b < n; b += 8) {
            unsigned long long wd = 0;
            for (int k = 0; k < 8 && b + k < n; k++) wd |= (unsigned long long)lt[b + k] << (8 * k);
            key = mix64_dev(key ^ wd);
        }
        s_key = mix64_dev(key ^ (unsigned long long)n);
    }
    __syncthreads();
    const unsigned long long key = s_key;
    // rows are dealt to threads in interleaved pairs (i, n-1-i) to even out the triangle
    uint32_t total_all = 0;
    for (int r0 = 0; r0 < n; r0 += SYN_THREADS) {
        const int i = r0 + threadIdx.x;
        uint32_t cnt = 0;
        if (n >= 5 && i < n)
            for (int j = i + 4; j < n; j++) cnt += synth_cell(lt, n, key, i, j) > 0;
        uint32_t total;
        uint32_t off = block_exclusive_scan<SYN_THREADS>(cnt, tmp, &total);
        if (FILL && n >= 5 && i < n && cnt) {
            long long dst = bpp_off[w] + total_all + off;
            for (int j = i + 4; j < n; j++) {
                const long long v = synth_cell(lt, n, key, i, j);
                if (v > 0) {
                    const double root = transport_root((uint32_t)v);
                    oi[dst] = (uint16_t)i; oj[dst] = (uint16_t)j; op[dst] = __dmul_rn(root, root);
                    dst++;
                }
            }
        }
        total_all += total;
    }
    if (!FILL && threadIdx.x == 0) bpp_cnt[w] = total_all;
}
