// gen.cuh -- K0 `gen_ndjson`: the synthetic NDJSON stream of SURVEY.md 8(d) S3, generated on the device.
//
// Record g of the stream is line (g mod L) of a template of L records (the reference's parking-citations fixture:
// 1000 records, every one starting with  {"Ticket":"<10 digits>"  ) with the ten Ticket digits replaced by g, zero padded:
// seed-free, reproducible, every record of a 64 GiB stream different from every other, record lengths unchanged.  Records
// are joined by '\n'.  One thread writes 16 output bytes.
#pragma once
#include "common.cuh"

namespace sj {

constexpr uint32_t GEN_TICKET_OFF = 11;  // strlen("{\"Ticket\":\"")
constexpr uint32_t GEN_TICKET_DIGITS = 10;

struct GenParams {
    const uint8_t* tmpl;    // L records, each followed by '\n' (T bytes)
    const uint32_t* off;    // [L + 1] start of each record inside tmpl; off[L] = T
    uint32_t L;
    uint64_t T;
    uint64_t first_record;  // global number of the first record generated (a multiple of L)
    uint64_t out_len;       // bytes to produce (the last record's '\n' is not part of it)
    uint8_t* out;           // 16-byte aligned
};

__global__ void __launch_bounds__(256) gen_ndjson_kernel(const GenParams p) {
    extern __shared__ uint32_t s_off[];  // [L + 1]
    for (uint32_t i = threadIdx.x; i <= p.L; i += blockDim.x) s_off[i] = p.off[i];
    __syncthreads();
    const uint64_t pos0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (pos0 >= p.out_len) return;
    uint64_t cyc = pos0 / p.T;                 // how many whole templates lie in front
    uint32_t r = (uint32_t)(pos0 - cyc * p.T);  // offset inside the template
    uint32_t lo = 0, hi = p.L;                 // line j with off[j] <= r < off[j + 1]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_off[mid] <= r)
            lo = mid;
        else
            hi = mid;
    }
    uint32_t j = lo;
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if (r >= s_off[j + 1]) {  // next record (records are at least 22 bytes long: at most one step per byte)
            j++;
            if (j == p.L) {
                j = 0;
                r = 0;
                cyc++;
            }
        }
        uint32_t c = p.tmpl[r];
        const uint32_t k = r - s_off[j];  // offset inside the record
        if (k - GEN_TICKET_OFF < GEN_TICKET_DIGITS) {
            uint64_t g = (p.first_record + cyc * p.L + j) % 10000000000ull;
            const uint32_t d = GEN_TICKET_DIGITS - 1 - (k - GEN_TICKET_OFF);  // power of ten of this digit
            for (uint32_t q = 0; q < d; q++) g /= 10;
            c = '0' + (uint32_t)(g % 10);
        }
        w[i >> 2] |= c << (8 * (i & 3));
        r++;
    }
    if (pos0 + 16 <= p.out_len) {
        *reinterpret_cast<uint4*>(p.out + pos0) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        for (uint32_t i = 0; pos0 + i < p.out_len; i++) p.out[pos0 + i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
    }
}

}  // namespace sj
