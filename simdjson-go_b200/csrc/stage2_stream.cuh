// stage2_stream.cuh -- the streaming stage 2 on the device: kernels around the portable core (s2s_core.h, s2s_slab.h).
//
//   K2p s2s_count      one warp per 6 KiB slab: per-slab aggregate (tape words, string bytes, brackets, depth, records,
//                      structurals, numbers, bytes behind the last quote)
//   K2q s2s_scan_*     exclusive scan of the aggregates (groups of 1024 + their totals), grand totals -> Stage2Result
//   K2r s2s_emit       the same analysis again, now with every offset known: tape words and Strings.B bytes staged in
//                      shared memory and streamed out with coalesced / 16-byte stores, bracket records for the scope
//                      matching, number list, per-segment grammar masks
//   K2h s2s_numbers    parse_number (parse_number.go:65) over the number list, one number per thread
//   K2d s2_min32 + s2_ansv (stage2.cuh)    scope matching on the brackets
//   K2e s2s_link       per bracket: cross-links of { } [ ] (stage2...go:327-334) and the grammar verdict of the segment
//                      in front of it against the container it lies in
//   K2f s2_roots (stage2.cuh)
//
// The unifiedMachine of the reference (stage2_build_tape_amd64.go:160-446) walks one structural at a time; the
// per-structural kernels of stage2.cuh gave every structural a thread and spent ~460 warp-instructions per 32
// structurals on divergent per-type work.  Here the warp walks the MESSAGE like stage 1 does (lane = 64-byte block),
// and what is left per structural is a short loop.
#pragma once
#include "common.cuh"
#include "number.cuh"
#include "s2s_slab.h"
#include "stage1.cuh"
#include "stage2.cuh"

namespace sj {

static_assert(S2S_SLAB_BYTES == (uint32_t)S1_SLAB_BYTES, "stage 2 takes the in-string state per stage-1 slab");
static_assert(S1_WARPS <= 32, "one bit per slab of a tile");

#ifndef SJ_S2S_WARPS
#define SJ_S2S_WARPS 8
#endif
constexpr int S2S_WARPS = SJ_S2S_WARPS;  // slabs per CTA
constexpr int S2S_THREADS = S2S_WARPS * 32;
constexpr uint32_t S2S_SSTAGE_PAD = (S2S_SSTAGE_BYTES + 15u) & ~15u;
constexpr uint32_t S2S_WARP_SMEM_COUNT = S2S_IMAGE_BYTES + S2S_ESC_SCRATCH;
static_assert(S2S_TSTAGE_WORDS * 8 >= S2S_ESC_SCRATCH, "K2r decodes escapes in the (then idle) tape staging area");
constexpr uint32_t S2S_WARP_SMEM_EMIT = S2S_IMAGE_BYTES + S2S_SSTAGE_PAD + S2S_TSTAGE_WORDS * 8;
constexpr size_t S2S_SMEM_COUNT = (size_t)S2S_WARPS * S2S_WARP_SMEM_COUNT;
constexpr size_t S2S_SMEM_EMIT = (size_t)S2S_WARPS * S2S_WARP_SMEM_EMIT;
#ifndef SJ_S2S_EMIT_MIN_BLOCKS
#define SJ_S2S_EMIT_MIN_BLOCKS 2
#endif
#ifndef SJ_S2S_COUNT_MIN_BLOCKS
#define SJ_S2S_COUNT_MIN_BLOCKS 3
#endif

struct DevWarp {
    __device__ __forceinline__ uint32_t lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ uint32_t ballot(bool p) { return __ballot_sync(FULL, p); }
    __device__ __forceinline__ bool any(bool p) { return __any_sync(FULL, p) != 0; }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, uint32_t src) { return __shfl_sync(FULL, v, (int)src); }
    __device__ __forceinline__ uint32_t shfl_up(uint32_t v, int d) { return __shfl_up_sync(FULL, v, d); }
    __device__ __forceinline__ uint32_t reduce_add(uint32_t v) { return __reduce_add_sync(FULL, v); }
    __device__ __forceinline__ void sync() { __syncwarp(); }
    __device__ __forceinline__ void atomic_and(uint32_t* p, uint32_t v) { atomicAnd(p, v); }
    __device__ __forceinline__ void atomic_or(uint32_t* p, uint32_t v) { atomicOr(p, v); }
    // LDGSTS: 16 bytes global -> shared without a register round trip; .ca keeps the line in L1 for the byte look-ups
    __device__ __forceinline__ void async_copy16(void* dst, const void* src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    }
    __device__ __forceinline__ void async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
    __device__ __forceinline__ void async_wait_prev() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    __device__ __forceinline__ void atomic_or_shared(uint32_t* p, uint32_t v) {
        asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
    }
};

// character / transition / compaction tables of a CTA (shared memory)
struct S2sTables {
    uint8_t ctab[256];
    uint8_t oktab[256];
    uint32_t cmptab[16];
};
__device__ __forceinline__ void s2s_fill_tables(S2sTables& t) {
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        t.ctab[i] = (uint8_t)char_type(i);
        const uint32_t p = i >> 4, c = i & 15;
        t.oktab[i] = (p < 15 && c < 15) ? (uint8_t)transition_mask(p, c) : (uint8_t)0;
    }
    if (threadIdx.x < 16) t.cmptab[threadIdx.x] = compress_sel(threadIdx.x) | ((uint32_t)__popc(threadIdx.x) << 16);
}

__global__ void __launch_bounds__(S2S_THREADS, SJ_S2S_COUNT_MIN_BLOCKS) s2s_count_kernel(const S2sParams p) {
    extern __shared__ __align__(128) uint8_t s2s_smem[];
    __shared__ S2sTables tabs;
    s2s_fill_tables(tabs);
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5;
    S2sWarpMem sm;
    sm.src = s2s_smem + (size_t)warp * S2S_WARP_SMEM_COUNT;
    sm.sstage = nullptr;
    sm.tstage = nullptr;
    sm.esc = sm.src + S2S_IMAGE_BYTES;
    sm.ctab = tabs.ctab;
    sm.oktab = tabs.oktab;
    sm.cmptab = tabs.cmptab;
    DevWarp wp;
    s2s_warp_loop<DevWarp, false>(wp, p, blockIdx.x * S2S_WARPS + warp, gridDim.x * S2S_WARPS, sm);
}

__global__ void __launch_bounds__(S2S_THREADS, SJ_S2S_EMIT_MIN_BLOCKS) s2s_emit_kernel(const S2sParams p) {
    extern __shared__ __align__(128) uint8_t s2s_smem[];
    __shared__ S2sTables tabs;
    s2s_fill_tables(tabs);
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5;
    S2sWarpMem sm;
    uint8_t* base = s2s_smem + (size_t)warp * S2S_WARP_SMEM_EMIT;
    sm.src = base;
    sm.sstage = base + S2S_IMAGE_BYTES;
    sm.tstage = reinterpret_cast<uint64_t*>(base + S2S_IMAGE_BYTES + S2S_SSTAGE_PAD);
    sm.esc = reinterpret_cast<uint8_t*>(sm.tstage);
    sm.ctab = tabs.ctab;
    sm.oktab = tabs.oktab;
    sm.cmptab = tabs.cmptab;
    DevWarp wp;
    s2s_warp_loop<DevWarp, true>(wp, p, blockIdx.x * S2S_WARPS + warp, gridDim.x * S2S_WARPS, sm);
}

// ---------------------------------------------------------------------------------
// K2q: exclusive scan of SlabAgg with agg_combine (not commutative in `trail`)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ SlabAgg agg_shfl_up(const SlabAgg& a, int d) {
    SlabAgg r;
    r.w = __shfl_up_sync(FULL, a.w, d);
    r.str = __shfl_up_sync(FULL, a.str, d);
    r.brk = __shfl_up_sync(FULL, a.brk, d);
    r.rec = __shfl_up_sync(FULL, a.rec, d);
    r.depth = __shfl_up_sync(FULL, a.depth, d);
    r.ns = __shfl_up_sync(FULL, a.ns, d);
    r.num = __shfl_up_sync(FULL, a.num, d);
    r.trail = __shfl_up_sync(FULL, a.trail, d);
    return r;
}
// blockDim.x = 1024; returns the exclusive prefix of the calling thread, `total` = the block's sum
__device__ __forceinline__ SlabAgg block_exclusive_scan_agg(const SlabAgg& v, SlabAgg& total) {
    __shared__ SlabAgg warp_inc[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SlabAgg inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const SlabAgg t = agg_shfl_up(inc, d);
        if (lane >= d) inc = agg_combine(t, inc);
    }
    if (lane == 31) warp_inc[warp + 1] = inc;
    __syncthreads();
    if (warp == 0) {
        SlabAgg wv = warp_inc[lane + 1];  // 32 warps
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const SlabAgg t = agg_shfl_up(wv, d);
            if (lane >= d) wv = agg_combine(t, wv);
        }
        __syncwarp();
        warp_inc[lane + 1] = wv;  // inclusive over warps <= lane
        if (lane == 0) warp_inc[0] = agg_zero();
    }
    __syncthreads();
    total = warp_inc[32];
    SlabAgg ex = agg_shfl_up(inc, 1);
    if (lane == 0) ex = agg_zero();
    const SlabAgg r = agg_combine(warp_inc[warp], ex);
    __syncthreads();  // the shared array is reused by the next call
    return r;
}

__global__ void __launch_bounds__(1024) s2s_scan_groups_kernel(const SlabAgg* in, uint32_t n, SlabAgg* pre, SlabAgg* group_total) {
    const uint32_t i = blockIdx.x * 1024 + threadIdx.x;
    const SlabAgg v = i < n ? in[i] : agg_zero();
    SlabAgg total;
    const SlabAgg e = block_exclusive_scan_agg(v, total);
    if (i < n) pre[i] = e;
    if (threadIdx.x == 0) group_total[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) s2s_scan_top_kernel(const SlabAgg* in, uint32_t n, SlabAgg* pre, Stage2Result* res,
                                                            uint64_t* totals_out, uint64_t msg_bytes) {
    SlabAgg carry = agg_zero();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const SlabAgg v = i < n ? in[i] : agg_zero();
        SlabAgg total;
        const SlabAgg e = block_exclusive_scan_agg(v, total);
        if (i < n) pre[i] = agg_combine(carry, e);
        carry = agg_combine(carry, total);
    }
    if (threadIdx.x == 0) {
        res->tape_len = (uint64_t)carry.w + 2;  // + root open + root close
        res->strings_len = carry.str;
        res->n_brackets = carry.brk;
        res->n_records = carry.rec;
        res->final_depth = carry.depth;
        res->n_numbers = carry.num;
        if (totals_out) {  // sj_shard_totals in device memory, for an exchange that stays on the stream
            totals_out[0] = msg_bytes;
            totals_out[1] = (uint64_t)carry.w + 2;
            totals_out[2] = carry.str;
            totals_out[3] = (uint64_t)carry.rec + 1;
        }
    }
}

// ---------------------------------------------------------------------------------
// K2h: parse_number over the number list
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(S2_THREADS) s2s_numbers_kernel(const uint8_t* msg, uint64_t len, const NumEntry* list, uint32_t count,
                                                                 uint64_t* tape, uint32_t* error) {
    const uint32_t k = blockIdx.x * S2_THREADS + threadIdx.x;
    if (k >= count) return;
    const NumEntry e = list[k];
    uint64_t val = 0;
    uint64_t tag = parse_number_fast(msg + e.pos, len - e.pos, &val);
    if (tag == PN_SLOW) tag = parse_number(msg + e.pos, len - e.pos, &val);  // parse_number.go:65
    if (tag == 0) atomicOr(error, 1u);
    tape[e.slot] = tag;
    tape[(uint64_t)e.slot + 1] = val;
}

// ---------------------------------------------------------------------------------
// K2e: per bracket k (and k = nb for the segment behind the last bracket): is every structural of segment k -- the
// ones behind bracket k-1 up to and including bracket k -- allowed inside the container that is open there?  That
// container is the scope open right after bracket k-1: the bracket itself if it opens, else the parent of the scope
// it closes (par = nearest previous bracket with a smaller depth in front of it, K2d).  Closing brackets cross-link
// the tape words of their pair (stage2...go:327-334).
// ---------------------------------------------------------------------------------
// The same launch writes the root words (K2f, stage2...go:170,207-218,428-441): record r opens at rootpos[r], its close
// sits right in front of the next record's open (or is the last word of the tape).
__global__ void __launch_bounds__(S2_THREADS) s2s_link_kernel(const S2sParams p, const int32_t* par, uint32_t nb, uint64_t n_records,
                                                              uint64_t tape_len) {
    const uint32_t k = blockIdx.x * S2_THREADS + threadIdx.x;
    const uint64_t tape_base = s2s_tape_base(p);
    if (k <= n_records) {
        const uint64_t R = (uint64_t)'r' << 56;
        const uint64_t open = k == 0 ? 0 : p.rootpos[k];
        const uint64_t next_open = k == n_records ? tape_len : p.rootpos[k + 1];
        if (next_open <= tape_len && next_open != 0) {
            p.tape[open] = R | (tape_base + next_open);
            p.tape[next_open - 1] = R | (tape_base + open);
        }
    }
    if (k > nb) return;
    uint32_t ctx = CTX_ROOT;
    if (k > 0) {
        const uint32_t kd = p.brk_kind[k - 1];
        int32_t enc;
        if (kd == T_OBJ_OPEN || kd == T_ARR_OPEN) {
            enc = (int32_t)k - 1;
        } else {
            const int32_t m = par[k - 1];
            enc = m >= 0 ? par[m] : -1;
        }
        ctx = enc >= 0 ? (p.brk_kind[enc] == T_OBJ_OPEN ? CTX_OBJ : CTX_ARR) : CTX_ROOT;
    }
    const uint32_t sg = (p.segmask[k >> 2] >> (8 * (k & 3))) & 0xffu;
    if (!((sg >> ctx) & 1u)) atomicOr(p.error, 1u);
    if (k < nb) {
        const uint32_t kd = p.brk_kind[k];
        if (kd == T_OBJ_CLOSE || kd == T_ARR_CLOSE) {
            const int32_t m = par[k];
            if (m >= 0) {
                const uint32_t otp = p.brk_tp[m], ctp = p.brk_tp[k];
                p.tape[otp] = ((uint64_t)(kd == T_OBJ_CLOSE ? '{' : '[') << 56) | (tape_base + ctp + 1);
                p.tape[ctp] = ((uint64_t)(kd == T_OBJ_CLOSE ? '}' : ']') << 56) | (tape_base + otp);
            }
        }
    }
}

}  // namespace sj
