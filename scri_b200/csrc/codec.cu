// Integer stages of scri's RPXMB waveform codec: XOR of successive time steps, Fletcher-32 checksum and
// the bit-level "multishuffle" - replaces scri/utilities.py:194-407 (numba loops) bit for bit.  All three are HBM-bound
// integer / byte kernels: every element is read once and written once.
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"

namespace scrib200 {

// out[i, c] = in[i, c] ^ in[i-1, c] (row 0 unchanged): what remains of each sample once the bits shared with its predecessor are gone.
__global__ void __launch_bounds__(256)
xor_forward_kernel(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, int64_t n_rows, int64_t n_cols) {
    const int64_t total = n_rows * n_cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = (e >= n_cols) ? (in[e] ^ in[e - n_cols]) : in[e];
}

// The inverse is a prefix XOR along time: chunk totals, an exclusive scan of the totals per column, then the rescan.
constexpr int XOR_CHUNK = 256;
__global__ void __launch_bounds__(128)
xor_chunk_totals_kernel(const unsigned long long* __restrict__ in, int64_t n_rows, int64_t n_cols, unsigned long long* __restrict__ totals) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const int64_t r0 = (int64_t)blockIdx.y * XOR_CHUNK, r1 = min(n_rows, r0 + XOR_CHUNK);
    unsigned long long x = 0;
    for (int64_t r = r0; r < r1; ++r) x ^= in[r * n_cols + c];
    totals[(int64_t)blockIdx.y * n_cols + c] = x;
}
__global__ void __launch_bounds__(128)
xor_scan_totals_kernel(unsigned long long* __restrict__ totals, int64_t n_chunks, int64_t n_cols) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    unsigned long long run = 0;
    for (int64_t k = 0; k < n_chunks; ++k) {
        const unsigned long long t = totals[k * n_cols + c];
        totals[k * n_cols + c] = run;
        run ^= t;
    }
}
__global__ void __launch_bounds__(128)
xor_rescan_kernel(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, int64_t n_rows, int64_t n_cols,
                  const unsigned long long* __restrict__ totals) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const int64_t r0 = (int64_t)blockIdx.y * XOR_CHUNK, r1 = min(n_rows, r0 + XOR_CHUNK);
    unsigned long long run = totals[(int64_t)blockIdx.y * n_cols + c];
    for (int64_t r = r0; r < r1; ++r) {
        run ^= in[r * n_cols + c];
        out[r * n_cols + c] = run;
    }
}

// Fletcher-32 over 16-bit words with modulus 65535: c0 = sum d_j, c1 = sum (size - j) d_j (each word enters every later
// running sum once); the block-wise reductions of the reference only keep 32-bit accumulators from overflowing.
__global__ void __launch_bounds__(256)
fletcher32_kernel(const unsigned short* __restrict__ d, int64_t size, unsigned long long* __restrict__ acc) {
    unsigned long long s0 = 0, s1 = 0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < size; j += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long v = d[j];
        s0 += v;
        s1 += v * (unsigned long long)((size - j) % 65535);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(acc, s0 % 65535ull);
        atomicAdd(acc + 1, s1 % 65535ull);
    }
}

struct ShuffleWidths {
    int n;                      // pieces, lowest significance first (the reference walks its widths in reverse)
    unsigned char width[64];
    unsigned char shift[65];    // bits below piece i inside an element = sum of the widths before it
};

// forward: output word B holds bits [B bw, (B+1) bw) of the stream "piece 0 of every element, piece 1 of every element, ..."
template <typename T>
__global__ void __launch_bounds__(256)
multishuffle_forward_kernel(const T* __restrict__ a, T* __restrict__ b, int64_t n, const ShuffleWidths w) {
    constexpr int BW = 8 * sizeof(T);
    for (int64_t B = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; B < n; B += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long p = (unsigned long long)B * BW;
        int i = 0;
        while (i + 1 < w.n && p >= (unsigned long long)n * w.shift[i + 1]) ++i;
        // position inside piece i: element e, r bits of its piece already taken - one division here, none in the loop
        const unsigned long long q = p - (unsigned long long)n * w.shift[i];
        int wi = w.width[i];
        unsigned long long e = q / wi;
        int r = (int)(q - e * wi);
        unsigned long long word = 0;
        int filled = 0;
        while (filled < BW) {
            const int take = min(wi - r, BW - filled);
            const unsigned long long bits = ((unsigned long long)a[e] >> (w.shift[i] + r)) & ((take == 64) ? ~0ull : ((1ull << take) - 1ull));
            word |= bits << filled;
            filled += take;
            r += take;
            if (r == wi) {
                r = 0;
                if (++e == (unsigned long long)n) {      // piece i of every element is stored: the next piece starts over at element 0
                    e = 0;
                    if (++i < w.n) wi = w.width[i];
                }
            }
        }
        b[B] = (T)word;
    }
}

// forward, 32- and 64-bit elements: a bit transpose through the warp instead of a gather from global memory.  A warp owns 32
// consecutive elements (one coalesced load); piece i of those elements is a run of 32 w_i bits = w_i words, lane l's bits at
// offset l w_i: a ballot for 1-bit pieces, a shuffle butterfly for widths that divide 32 (all words of the run at once), and for
// the rest every word of the run is an OR-reduction (`redux.sync`) over the lanes of their overlap with it.  The 8 warps of a CTA put their segments side by side in shared memory (256 w_i bits per piece, word aligned,
// nobody shares a word), then the CTA writes each piece's run at its place in the stream, n S_i + e0 w_i bits in: funnel-shifted
// by that position's offset inside a word, whole words stored, the first and last (shared with the neighbouring CTAs' runs,
// or with the next piece's region) OR-ed atomically into the zeroed output.
constexpr int MS_WARPS = 8;
constexpr int MS_ELEMS = 32 * MS_WARPS;

template <typename T>
__global__ void __launch_bounds__(32 * MS_WARPS)
multishuffle_forward_warp_kernel(const T* __restrict__ a, unsigned int* __restrict__ out, int64_t n, int64_t n_chunks, const ShuffleWidths w) {
    constexpr int BW = 8 * sizeof(T);
    __shared__ unsigned int run[MS_WARPS * BW + 2];          // the pieces' runs back to back: piece i at 8 S_i, 8 w_i words
    __shared__ unsigned short item_piece[MS_WARPS * BW + 64];
    __shared__ unsigned short item_k[MS_WARPS * BW + 64];
    __shared__ int n_items;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the CTA's output items (piece, word of its shifted run): 8 w_i + 1 per piece
    if (tid < w.n) {
        const int base = MS_WARPS * w.shift[tid] + tid, cnt = MS_WARPS * w.width[tid] + 1;
        for (int k = 0; k < cnt; ++k) {
            item_piece[base + k] = (unsigned short)tid;
            item_k[base + k] = (unsigned short)k;
        }
        if (tid == w.n - 1) n_items = base + cnt;
    }
    __syncthreads();
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int64_t e0 = chunk * MS_ELEMS;
        const int64_t e = e0 + tid;
        const unsigned long long v = e < n ? (unsigned long long)a[e] : 0ull;
        for (int i = 0; i < w.n; ++i) {
            const int wi = w.width[i];
            const unsigned long long val = (v >> w.shift[i]) & (wi == 64 ? ~0ull : ((1ull << wi) - 1ull));
            unsigned int* seg = run + MS_WARPS * w.shift[i] + warp * wi;     // this warp's wi words of piece i
            if (wi == 1) {
                const unsigned int word = __ballot_sync(0xffffffffu, val & 1ull);
                if (lane == 0) seg[0] = word;
            } else if (wi <= 32 && (wi & (wi - 1)) == 0) {
                // 32 / wi pieces per word: a shuffle butterfly builds all wi words at once (lane q 32 / wi ends up with word q)
                unsigned int x = (unsigned int)val;
                for (int step = 1, bits = wi; bits < 32; step <<= 1, bits <<= 1) x |= __shfl_down_sync(0xffffffffu, x, step) << bits;
                const int lg = 6 - __ffs(wi);                                // log2(32 / wi): no division by a run-time width
                if ((lane & ((1 << lg) - 1)) == 0) seg[lane >> lg] = x;
            } else {
                const int o = lane * wi;                                     // my first bit within the segment
                for (int q = 0; q < wi; ++q) {
                    const int d = 32 * q - o;                                // word q starts d bits into my piece
                    unsigned int c = 0;
                    if (d >= 0) {
                        if (d < wi) c = (unsigned int)(val >> d);
                    } else if (d > -32) {
                        c = (unsigned int)(val << (-d));
                    }
                    c = __reduce_or_sync(0xffffffffu, c);
                    if (lane == (q & 31)) seg[q] = c;
                }
            }
        }
        __syncthreads();
        const bool partial = e0 + MS_ELEMS > n;                              // the last chunk: its runs end inside a piece's region
        for (int it = tid; it < n_items; it += 32 * MS_WARPS) {
            const int i = item_piece[it], k = item_k[it], wi = w.width[i], nw = MS_WARPS * wi;
            const unsigned long long pos = (unsigned long long)n * w.shift[i] + (unsigned long long)e0 * wi;   // bit position of the run
            const int sh = (int)(pos & 31ull);
            const unsigned int* r = run + MS_WARPS * w.shift[i];
            const unsigned int lo = k < nw ? r[k] : 0u, prev = k > 0 ? r[k - 1] : 0u;
            const unsigned int word = sh ? ((lo << sh) | (prev >> (32 - sh))) : lo;
            if (k == nw && sh == 0) continue;                                // nothing spills into the word after the run
            const unsigned long long wordidx = (pos >> 5) + (unsigned long long)k;
            if (partial) {
                const unsigned long long end = (unsigned long long)n * (w.shift[i] + wi);      // first bit after piece i's region
                if (wordidx * 32ull >= end) continue;
                if (word) atomicOr(out + wordidx, word);
            } else if ((k == 0 && sh != 0) || k == nw) {
                if (word) atomicOr(out + wordidx, word);
            } else {
                out[wordidx] = word;
            }
        }
        __syncthreads();
    }
}

// reverse: element e collects piece i from bits [n S_i + e w_i, + w_i) of the stream (possibly straddling two words)
template <typename T>
__global__ void __launch_bounds__(256)
multishuffle_reverse_kernel(const T* __restrict__ b, T* __restrict__ a, int64_t n, const ShuffleWidths w) {
    constexpr int BW = 8 * sizeof(T);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long val = 0;
        for (int i = 0; i < w.n; ++i) {
            const int wi = w.width[i];
            const unsigned long long p = (unsigned long long)n * w.shift[i] + (unsigned long long)e * wi;
            const unsigned long long B = p / BW;
            const int off = (int)(p - B * BW);
            unsigned long long bits = (unsigned long long)b[B] >> off;
            if (off + wi > BW) bits |= (unsigned long long)b[B + 1] << (BW - off);
            bits &= (wi == 64) ? ~0ull : ((1ull << wi) - 1ull);
            val |= bits << w.shift[i];
        }
        a[e] = (T)val;
    }
}

// Conjugate-pair form of a mode series (scri/waveform_modes.py:658-703), in place.  Forward: for every (l, m > 0)
//   f[l, m] <- (f[l, m] + conj f[l, -m]) / sqrt(2),  f[l, -m] <- (f[l, m] - conj f[l, -m]) / sqrt(2);
// inverse:  f[l, m] <- (s + d) / sqrt(2),  f[l, -m] <- conj(s - d) / sqrt(2).
// The reference divides a complex array by the real np.sqrt(2): numpy runs its complex division (Smith's algorithm) with a zero
// imaginary divisor, i.e. re = (a_r + a_i * 0) * (1 / sqrt 2), im = (a_i - a_r * 0) * (1 / sqrt 2) - repeated here operation by
// operation (no contraction) so the result is bit-identical, signed zeros included.
__device__ __forceinline__ double2 np_div_by_real(double2 a, double scl) {
    const double rat = 0.0;
    return make_double2(__dmul_rn(__dadd_rn(a.x, __dmul_rn(a.y, rat)), scl), __dmul_rn(__dsub_rn(a.y, __dmul_rn(a.x, rat)), scl));
}

__global__ void __launch_bounds__(256)
conjugate_pairs_kernel(double2* __restrict__ data, int64_t n_times, int n_modes, int ell_min, int ell_max, int n_pairs, int inverse) {
    const double scl = __ddiv_rn(1.0, 1.4142135623730951);
    const int64_t total = n_times * (int64_t)n_pairs;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / n_pairs;
        int p = (int)(e - row * n_pairs);
        // pairs are numbered ell by ell: ell contributes ell pairs; pairs below ell: ell(ell-1)/2 - ell_min(ell_min-1)/2
        int ell = ell_min;
        const int base = ell_min * (ell_min - 1) / 2;
        while ((ell + 1) * ell / 2 - base <= p) ++ell;
        const int m = p - (ell * (ell - 1) / 2 - base) + 1;
        const int centre = ell * (ell + 1) - ell_min * ell_min;
        double2* plus = data + row * n_modes + centre + m;
        double2* minus = data + row * n_modes + centre - m;
        const double2 a = *plus, b = *minus;
        if (!inverse) {
            const double2 bc = make_double2(b.x, -b.y);
            *plus = np_div_by_real(make_double2(__dadd_rn(a.x, bc.x), __dadd_rn(a.y, bc.y)), scl);
            *minus = np_div_by_real(make_double2(__dsub_rn(a.x, bc.x), __dsub_rn(a.y, bc.y)), scl);
        } else {
            *plus = np_div_by_real(make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y)), scl);
            *minus = np_div_by_real(make_double2(__dsub_rn(a.x, b.x), -__dsub_rn(a.y, b.y)), scl);
        }
    }
}

// scri/waveform_modes.py:457-476: every time step is rounded to a multiple of 2^-floor(-log2(|f(t)| tol_per_mode)).  One warp per
// row: |f|^2 by a shuffle reduction, then data <- rint(data * p) / p (np.round rounds half to even; p is a power of two, so the
// scalings are exact).  The row norm is summed in a different order than numpy's pairwise sum: the result can differ from the
// reference only when |f| tol sits within an ulp of a power of two.  A zero row becomes NaN, as in the reference (0 * inf).
__global__ void __launch_bounds__(256)
truncate_kernel(double* __restrict__ data, int64_t n_times, int n_real, double tol_per_mode) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_times; row += warps) {
        double* x = data + row * n_real;
        double s = 0.0;
        for (int c = lane; c < n_real; c += 32) s = fma(x[c], x[c], s);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const double absolute_tolerance = __dmul_rn(sqrt(s), tol_per_mode);
        const double p = exp2(floor(-log2(absolute_tolerance)));
        for (int c = lane; c < n_real; c += 32) x[c] = __ddiv_rn(rint(__dmul_rn(x[c], p)), p);
    }
}

static unsigned grid_for(int64_t n, int threads) {
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = 148 * 16;
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace scrib200

extern "C" int scrib200_xor_timeseries(const void* in, void* out, int64_t n_rows, int64_t n_cols, int reverse, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(in && out, "xor_timeseries: null pointer");
    if (n_rows <= 0 || n_cols <= 0) return SCRIB200_OK;
    auto* src = reinterpret_cast<const unsigned long long*>(in);
    auto* dst = reinterpret_cast<unsigned long long*>(out);
    cudaStream_t st = (cudaStream_t)stream;
    if (!reverse) {
        SCRIB200_REQUIRE(in != out, "xor_timeseries: the forward transform is out of place");
        xor_forward_kernel<<<grid_for(n_rows * n_cols, 256), 256, 0, st>>>(src, dst, n_rows, n_cols);
        SCRIB200_CHECK_LAUNCH("xor_timeseries");
        return SCRIB200_OK;
    }
    const int64_t n_chunks = (n_rows + XOR_CHUNK - 1) / XOR_CHUNK;
    SCRIB200_REQUIRE(workspace && workspace_bytes >= (size_t)(n_chunks * n_cols) * 8, "xor_timeseries: workspace too small (%zu < %zu)",
                     workspace_bytes, (size_t)(n_chunks * n_cols) * 8);
    SCRIB200_REQUIRE(n_chunks <= 65535, "xor_timeseries: series too long (%lld rows)", (long long)n_rows);
    auto* totals = reinterpret_cast<unsigned long long*>(workspace);
    dim3 grid((unsigned)((n_cols + 127) / 128), (unsigned)n_chunks);
    xor_chunk_totals_kernel<<<grid, 128, 0, st>>>(src, n_rows, n_cols, totals);
    SCRIB200_CHECK_LAUNCH("xor_timeseries");
    xor_scan_totals_kernel<<<grid.x, 128, 0, st>>>(totals, n_chunks, n_cols);
    SCRIB200_CHECK_LAUNCH("xor_timeseries");
    xor_rescan_kernel<<<grid, 128, 0, st>>>(src, dst, n_rows, n_cols, totals);
    SCRIB200_CHECK_LAUNCH("xor_timeseries");
    return SCRIB200_OK;
}

extern "C" size_t scrib200_xor_timeseries_workspace_bytes(int64_t n_rows, int64_t n_cols) {
    return (size_t)(((n_rows + scrib200::XOR_CHUNK - 1) / scrib200::XOR_CHUNK) * n_cols) * 8 + 16;
}

extern "C" int scrib200_fletcher32(const void* data, int64_t n_words16, void* acc2, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(data && acc2, "fletcher32: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(acc2, 0, 16, st);
    SCRIB200_REQUIRE(e == cudaSuccess, "fletcher32: %s", cudaGetErrorString(e));
    if (n_words16 <= 0) return SCRIB200_OK;
    fletcher32_kernel<<<grid_for(n_words16, 256), 256, 0, st>>>(reinterpret_cast<const unsigned short*>(data), n_words16,
                                                               reinterpret_cast<unsigned long long*>(acc2));
    SCRIB200_CHECK_LAUNCH("fletcher32");
    return SCRIB200_OK;
}

extern "C" int scrib200_multishuffle(const void* in, void* out, int64_t n, int bit_width, const int* widths, int n_widths, int forward,
                                     void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(in && out && widths, "multishuffle: null pointer");
    SCRIB200_REQUIRE(in != out, "multishuffle: out of place only");
    SCRIB200_REQUIRE(bit_width == 8 || bit_width == 16 || bit_width == 32 || bit_width == 64, "multishuffle: Total bit width must be one of [8, 16, 32, 64], not %d", bit_width);
    SCRIB200_REQUIRE(n_widths >= 1 && n_widths <= 64, "multishuffle: %d shuffle widths", n_widths);
    ShuffleWidths w;
    w.n = n_widths;
    int sum = 0;
    for (int i = 0; i < n_widths; ++i) {           // widths come highest significance first; the stream starts at the lowest
        const int wi = widths[n_widths - 1 - i];
        SCRIB200_REQUIRE(wi >= 1 && wi <= 64, "multishuffle: shuffle width %d", wi);
        w.width[i] = (unsigned char)wi;
        w.shift[i] = (unsigned char)sum;
        sum += wi;
    }
    w.shift[n_widths] = (unsigned char)(sum > 255 ? 255 : sum);
    SCRIB200_REQUIRE(sum == bit_width, "multishuffle: the shuffle widths sum to %d, not to the bit width %d", sum, bit_width);
    if (n <= 0) return SCRIB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned g = grid_for(n, 256);
    if (forward && bit_width >= 32 && getenv("SCRIB200_MULTISHUFFLE_GATHER") == nullptr) {
        // bit transpose through the warp (whole 32-bit words of output: the element types whose streams are word multiples)
        cudaError_t e = cudaMemsetAsync(out, 0, (size_t)n * (bit_width / 8), st);
        SCRIB200_REQUIRE(e == cudaSuccess, "multishuffle: %s", cudaGetErrorString(e));
        const int64_t n_chunks = (n + MS_ELEMS - 1) / MS_ELEMS;
        const unsigned grid = (unsigned)std::min<int64_t>(n_chunks, 148 * 8);
        if (bit_width == 32)
            multishuffle_forward_warp_kernel<unsigned int><<<grid, 32 * MS_WARPS, 0, st>>>(reinterpret_cast<const unsigned int*>(in), reinterpret_cast<unsigned int*>(out), n, n_chunks, w);
        else
            multishuffle_forward_warp_kernel<unsigned long long><<<grid, 32 * MS_WARPS, 0, st>>>(reinterpret_cast<const unsigned long long*>(in), reinterpret_cast<unsigned int*>(out), n, n_chunks, w);
        SCRIB200_CHECK_LAUNCH("multishuffle");
        return SCRIB200_OK;
    }
#define SCRIB200_SHUFFLE(T)                                                                                                   \
    if (forward) multishuffle_forward_kernel<T><<<g, 256, 0, st>>>(reinterpret_cast<const T*>(in), reinterpret_cast<T*>(out), n, w); \
    else multishuffle_reverse_kernel<T><<<g, 256, 0, st>>>(reinterpret_cast<const T*>(in), reinterpret_cast<T*>(out), n, w);
    if (bit_width == 8) { SCRIB200_SHUFFLE(unsigned char) }
    else if (bit_width == 16) { SCRIB200_SHUFFLE(unsigned short) }
    else if (bit_width == 32) { SCRIB200_SHUFFLE(unsigned int) }
    else { SCRIB200_SHUFFLE(unsigned long long) }
#undef SCRIB200_SHUFFLE
    SCRIB200_CHECK_LAUNCH("multishuffle");
    return SCRIB200_OK;
}

extern "C" int scrib200_conjugate_pairs(void* data, int64_t n_times, int ell_min, int ell_max, int inverse, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(data, "conjugate_pairs: null pointer");
    SCRIB200_REQUIRE(ell_min >= 0 && ell_max >= ell_min - 1, "conjugate_pairs: ell range [%d, %d]", ell_min, ell_max);
    const int n_modes = (ell_max + 1) * (ell_max + 1) - ell_min * ell_min;
    const int n_pairs = ell_max * (ell_max + 1) / 2 - ell_min * (ell_min - 1) / 2;
    if (n_times <= 0 || n_pairs <= 0) return SCRIB200_OK;
    conjugate_pairs_kernel<<<grid_for(n_times * n_pairs, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<double2*>(data), n_times, n_modes, ell_min, ell_max, n_pairs, inverse);
    SCRIB200_CHECK_LAUNCH("conjugate_pairs");
    return SCRIB200_OK;
}

extern "C" int scrib200_truncate(void* data, int64_t n_times, int n_complex, double tol_per_mode, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(data, "truncate: null pointer");
    SCRIB200_REQUIRE(n_complex >= 0, "truncate: n_complex=%d", n_complex);
    if (n_times <= 0 || n_complex == 0) return SCRIB200_OK;
    truncate_kernel<<<grid_for(n_times * 32, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<double*>(data), n_times, 2 * n_complex,
                                                                                 tol_per_mode);
    SCRIB200_CHECK_LAUNCH("truncate");
    return SCRIB200_OK;
}
