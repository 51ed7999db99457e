// waf_bitset_nfa_kernel -- bit-parallel NFA simulation for the patterns no DFA unit can hold (nfa_bits.hpp).
//
// One thread per request and unit (blockIdx.y = unit).  The active NFA positions of the request are a bit vector of W
// words in the thread's registers (W <= 8) or local memory (W <= 64); the unit's tables -- byte classes, context map,
// follow rows, accept masks, per-class position masks -- are staged into shared memory when they fit the launch's
// allocation and read from global memory (L1 / L2) otherwise.  A step ORs the follow rows of the active positions
// (walked with ffs) and ANDs the result with the mask of the byte's class; matched patterns fire their atoms into the
// request's bitmap row exactly as the DFA units do.  Launched only for rule sets that contain such a pattern.

struct BitsetParams {
    const uint8_t* col[5];
    const uint32_t* off[5];
    uint32_t n;
    uint32_t* rows;
    uint32_t* info;
    uint32_t atom_words;
    const BitsetUnitDesc* units;   // global memory
    const uint32_t* blob;          // all units' tables
    uint32_t smem_words;           // words of dynamic shared memory the launch carries (a unit whose tables fit is staged)
};

constexpr int kBitsetThreads = 256;

template <int WMAX>
__device__ __forceinline__ void bitset_walk(const BitsetUnitDesc& d, const uint32_t* __restrict__ tab, const uint8_t* __restrict__ bytes,
                                            uint32_t s, uint32_t e, const Sink& sink) {
    const uint8_t* cmap = reinterpret_cast<const uint8_t*>(tab);
    const uint8_t* kind = cmap + 256;
    const uint8_t* ctx = cmap + 512;
    const uint32_t* events = tab + 132;
    const uint32_t W = d.words, P = d.n_pos;
    const uint32_t* follow = tab + d.follow_off;
    const uint32_t* accept = tab + d.accept_off;
    const uint32_t* bmask = tab + d.bmask_off;
    constexpr int kU = WMAX <= 8 ? WMAX : 1;   // register-resident vectors are fully unrolled, the large variant loops over local memory
    uint32_t S[WMAX], R[WMAX];
#pragma unroll(kU)
    for (int w = 0; w < WMAX; ++w) S[w] = 0u;
    uint32_t latch = 0u, pk = 0u, fired = 0u;
    for (uint32_t i = s;; ++i) {
        const bool end = i >= e;
        const uint32_t byte = end ? 0u : (uint32_t)__ldg(bytes + i);
        const uint32_t t = ctx[pk * 4u + (end ? 0u : (uint32_t)kind[byte])];
        const uint32_t* F = follow + (size_t)t * (P + 1u) * W;
        const uint32_t* A = accept + (size_t)t * (P + 1u);
#pragma unroll(kU)
        for (int w = 0; w < WMAX; ++w)
            if ((uint32_t)w < W) R[w] = F[(size_t)P * W + w];
        uint32_t acc = A[P];
#pragma unroll(kU)
        for (int w = 0; w < WMAX; ++w) {
            if ((uint32_t)w >= W) break;
            for (uint32_t x = S[w]; x; x &= x - 1u) {
                const uint32_t p = (uint32_t)w * 32u + (uint32_t)(__ffs((int)x) - 1);
                acc |= A[p];
                const uint32_t* Fp = F + (size_t)p * W;
#pragma unroll(kU)
                for (int v = 0; v < WMAX; ++v)
                    if ((uint32_t)v < W) R[v] |= Fp[v];
            }
        }
        for (uint32_t x = acc; x; x &= x - 1u) {
            const uint32_t j = (uint32_t)(__ffs((int)x) - 1), ev = events[j];
            const uint32_t k = ev >> kEvKindShift, lb = 1u << ((ev >> kEvLatchShift) & 31u);
            if (k == 0u || (k == 1u && (latch & lb))) {
                if (!((fired >> j) & 1u)) fire_atom(sink, ev & kEvAtomMask);   // once is enough: atoms only ever become true
                fired |= 1u << j;
            } else if (k == 2u) latch &= ~lb;
            else if (k == 3u) latch |= lb;
        }
        if (end) break;
        if (d.stop_mask && (fired & d.stop_mask) == d.stop_mask) break;
        const uint32_t* B = bmask + (size_t)cmap[byte] * W;
#pragma unroll(kU)
        for (int w = 0; w < WMAX; ++w)
            if ((uint32_t)w < W) S[w] = R[w] & B[w];
        pk = kind[byte];
    }
}

__global__ void __launch_bounds__(kBitsetThreads) waf_bitset_nfa_kernel(const __grid_constant__ BitsetParams bp) {
    extern __shared__ __align__(16) uint32_t bitset_smem[];
    const BitsetUnitDesc d = bp.units[blockIdx.y];
    const uint32_t* tab = bp.blob + d.blob_off;
    if (d.blob_words <= bp.smem_words) {
        // 16-byte copies: blob_off and blob_words are multiples of 4 words
        const uint4* src = reinterpret_cast<const uint4*>(tab);
        uint4* dst = reinterpret_cast<uint4*>(bitset_smem);
        for (uint32_t i = threadIdx.x; i < d.blob_words / 4u; i += blockDim.x) dst[i] = __ldg(src + i);
        __syncthreads();
        tab = bitset_smem;
    }
    const uint8_t* bytes = bp.col[d.field];
    const uint32_t* off = bp.off[d.field];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < bp.n; r += gridDim.x * blockDim.x) {
        const uint32_t s = __ldg(off + r), e = __ldg(off + r + 1);
        const Sink sink{bp.rows + (size_t)r * bp.atom_words, bp.info + 2u * (size_t)r};
        if (d.words <= 2u) bitset_walk<2>(d, tab, bytes, s, e, sink);
        else if (d.words <= 8u) bitset_walk<8>(d, tab, bytes, s, e, sink);
        else bitset_walk<64>(d, tab, bytes, s, e, sink);
    }
}
