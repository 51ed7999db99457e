// Part of kernels.cu (included inside namespace pgw { namespace { ... } }, one translation unit: device functions are
// not linked across files).  Stand-alone batch kernels: GeoIP lookup and captcha client ids.

// GeoipDB::lookup for a batch of addresses (pingoo/geoip.rs:73-91)
__global__ void geoip_lookup_kernel(const __grid_constant__ KParams p, const uint8_t* __restrict__ ip,
                                    const uint8_t* __restrict__ is_v6, uint32_t n, uint32_t* __restrict__ asn_out,
                                    uint16_t* __restrict__ country_out) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint8_t* ip16 = ip + (size_t)r * 16;
        const bool v6 = is_v6[r] != 0;
        uint32_t asn = 0, country = (uint32_t)'X' | ((uint32_t)'X' << 8);
        bool skip;
        if (!v6) skip = ip16[0] == 127 || (ip16[0] >> 4) == 0xE;
        else {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(ip16);
            skip = ip16[0] == 0xFF || (w[0] == 0 && w[1] == 0 && w[2] == 0 && w[3] == 0x01000000u);
        }
        if (!skip && p.geo_loaded) {
            const LpmLeaf lf = p.leaves[lpm_lookup(p, ip16, v6)];
            asn = lf.asn;
            country = lf.country;
        }
        asn_out[r] = asn;
        country_out[r] = (uint16_t)country;
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// captcha client id: generate_captcha_client_id (pingoo/captcha.rs:409-421) for a batch --
// base64url-no-pad( SHA-256( ip octets (4 or 16) || user_agent || host ) ), 43 characters per request.
// One thread per request; the message is at most 16 + 256 + 256 bytes, i.e. nine 64-byte blocks.
// ---------------------------------------------------------------------------------------------------------------------
__constant__ uint32_t kSha256K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

__global__ void __launch_bounds__(128) captcha_client_id_kernel(const uint8_t* __restrict__ ip, const uint8_t* __restrict__ is_v6,
                                                                const uint8_t* __restrict__ ua_bytes, const uint32_t* __restrict__ ua_off,
                                                                const uint8_t* __restrict__ host_bytes, const uint32_t* __restrict__ host_off, uint32_t n,
                                                                uint8_t* __restrict__ out44) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint8_t* ipp = ip + (size_t)r * 16;
    const uint32_t ipl = is_v6[r] ? 16u : 4u;
    const uint8_t* uap = ua_bytes + ua_off[r];
    const uint32_t ual = ua_off[r + 1] - ua_off[r];
    const uint8_t* hop = host_bytes + host_off[r];
    const uint32_t hol = host_off[r + 1] - host_off[r];
    const uint32_t total = ipl + ual + hol;
    const uint32_t n_blocks = (total + 9u + 63u) / 64u;
    auto msg_byte = [&](uint32_t i) -> uint32_t {
        if (i < ipl) return ipp[i];
        if (i < ipl + ual) return uap[i - ipl];
        if (i < total) return hop[i - ipl - ual];
        return i == total ? 0x80u : 0u;
    };
    uint32_t h[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    for (uint32_t b = 0; b < n_blocks; ++b) {
        uint32_t w[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const uint32_t p = b * 64u + 4u * t;
            w[t] = (msg_byte(p) << 24) | (msg_byte(p + 1) << 16) | (msg_byte(p + 2) << 8) | msg_byte(p + 3);
        }
        if (b == n_blocks - 1) {  // message length in bits, big endian, in the last eight bytes
            w[14] = 0;
            w[15] = total * 8u;
        }
        uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
        for (int t = 0; t < 64; ++t) {
            if (t >= 16) {
                const uint32_t w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
                const uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
                const uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
                w[t & 15] = w[t & 15] + s0 + w[(t + 9) & 15] + s1;
            }
            const uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t t1 = hh + S1 + ch + kSha256K[t] + w[t & 15];
            const uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
            const uint32_t mj = (a & bb) ^ (a & c) ^ (bb & c);
            const uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
        }
        h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    // base64url without padding: 32 bytes -> 43 characters (+ a terminating 0 in the 44th byte)
    auto digest_byte = [&](uint32_t i) -> uint32_t { return i < 32u ? (h[i >> 2] >> (24u - 8u * (i & 3u))) & 0xFFu : 0u; };
    auto b64 = [](uint32_t v) -> uint8_t {
        return (uint8_t)(v < 26u ? 'A' + v : v < 52u ? 'a' + (v - 26u) : v < 62u ? '0' + (v - 52u) : v == 62u ? '-' : '_');
    };
    uint8_t* o = out44 + (size_t)r * 44;
    for (uint32_t i = 0, j = 0; i < 33u; i += 3u, j += 4u) {
        const uint32_t v = (digest_byte(i) << 16) | (digest_byte(i + 1) << 8) | digest_byte(i + 2);
        o[j] = b64(v >> 18);
        o[j + 1] = b64((v >> 12) & 63u);
        if (j + 2 < 43u) o[j + 2] = b64((v >> 6) & 63u);
        if (j + 3 < 43u) o[j + 3] = b64(v & 63u);
    }
    o[43] = 0;
}

