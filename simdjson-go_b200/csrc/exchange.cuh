// exchange.cuh -- the ONE exchange of the sharded ParseND as a kernel over peer memory.
//
// ParseND returns one ParsedJson (simdjson_amd64.go:82-93): the shards' tapes and string buffers are slices of one
// tape / one Strings.B, so every rank needs the exclusive prefix of (message bytes, tape words, string bytes, records)
// over the ranks in front of it before its emitting kernels can write final root / scope pointers and string offsets.
// That is 32 bytes per rank.  Instead of an NCCL all-gather + prefix kernels enqueued by the host after the counting half
// has synchronised (launch latency of three more kernels sits on the critical path of every call), the counting half ends
// with THIS kernel on the same stream: one warp, lane r <-> rank r,
//     1. PUSH   lane r stores this shard's totals into slot [epoch & 1][my rank] of rank r's exchange buffer (peer
//               memory over NVLink: the buffers are cudaMalloc'ed and opened in every peer with CUDA IPC, or handed in
//               as peer pointers), fences, then stores the epoch number into the slot's sequence word (release, system
//               scope);
//     2. WAIT   lane r polls the sequence word of slot [epoch & 1][r] of the LOCAL buffer (acquire, system scope) until
//               it equals the epoch -- local polling, no traffic on the links while waiting -- and reads rank r's totals;
//     3. PREFIX warp scan over the lanes: bases of this shard + totals of the whole message into `out`, read by the
//               emitting kernels (S2sParams.bases_dev / Stage2Params.bases_dev) and copied back with the counting
//               half's own read-back.
// The kernel never waits long: after `slice_ns` (200 us) without the peers' totals it leaves with status PENDING, and the
// host, which synchronises at the end of the counting half anyway, enqueues wait-only passes until they are there or the
// time limit is over -- a rank that is milliseconds late (or a kernel that cannot be co-scheduled with a spinning one:
// ranks sharing one GPU in the tests) costs a relaunch, not a blocked device.
// Slots are double-buffered by epoch parity: a rank can be at most one epoch ahead of a peer (its push of epoch e + 2
// needs the peer's push of e + 1, which the peer enqueues behind its own wait of e), so a slot is never overwritten
// while a peer still waits on it.  A rank that fails before its totals exist pushes a FAILED marker, so the peers do not
// wait for it; a call whose peers stay silent for the time limit (two seconds unless set otherwise) gives up and reports it.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "stage1.cuh"

namespace sj {

constexpr int XCHG_MAX_WORLD = 32;                        // one lane per rank
constexpr int XCHG_SLOT_WORDS = 8;                        // 64 bytes: { msg, tape, strings, records, seq, pad x3 }
constexpr size_t XCHG_BUFFER_BYTES = 2 * XCHG_MAX_WORLD * XCHG_SLOT_WORDS * sizeof(uint64_t);
constexpr uint64_t XCHG_FAILED = ~0ull;                   // tape_words of a rank whose counting half failed
constexpr uint64_t XCHG_DEFAULT_TIMEOUT_NS = 2000000000ull;
constexpr uint64_t XCHG_SLICE_NS = 200000ull;             // longest wait of one kernel pass

enum : uint64_t { XCHG_OK = 0, XCHG_TIMEOUT = 1, XCHG_PEER_FAILED = 2, XCHG_PENDING = 3 };

// out[0..3] = { msg_base, tape_base, strings_base, records_base }   (the first three are what the emitting kernels read)
// out[4..7] = the same four of the whole message (all ranks)
// out[8]    = status (XCHG_*),  out[9] = epoch
constexpr int XCHG_OUT_WORDS = 10;

struct XchgParams {
    uint64_t* const* peers;  // device array [world]: rank r's exchange buffer as mapped into THIS process
    uint64_t* local;         // this rank's buffer (== peers[rank])
    const uint64_t* totals;  // this shard's { msg_bytes, tape_words, string_bytes, records } (written by the scan's top kernel)
    uint64_t* out;
    uint32_t rank, world;
    uint64_t epoch;          // >= 1, the same on every rank
    uint64_t gap;            // message bytes between this shard's window and the next shard's (the newline they were cut at: 1)
    uint32_t failed;         // push the FAILED marker instead of totals
    uint64_t slice_ns;       // how long this pass waits for the peers' pushes before it leaves with PENDING
    uint32_t wait_only;      // a further pass of the same epoch: the push has been made
    // optional, the streaming counting half (whose host has not seen stage 1's verdict yet when this kernel is enqueued):
    const Stage1Result* s1;     // overflow there = do nothing (the counting half is repeated and publishes then); a failed
                                // stage 1 (stage1_find_marks_amd64.go:115-147) = push the FAILED marker
    const uint32_t* s2_error;   // non-zero = the counting pass met an invalid escape: FAILED
};

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t ld_relaxed_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t xchg_now_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(32) shard_exchange_kernel(const XchgParams x) {
    const uint32_t lane = threadIdx.x;
    bool failed = x.failed != 0;
    if (x.s1 && !x.wait_only) {
        const Stage1Result r = *x.s1;
        if (r.overflow) return;
        const bool closed = r.n_idx != 0 && (r.last_char == '}' || r.last_char == ']');
        failed |= !closed || r.error != 0 || r.ends_in_string != 0;
    }
    if (x.s2_error && *x.s2_error) failed = true;
    const bool live = lane < x.world;
    const size_t slot0 = (size_t)(x.epoch & 1) * XCHG_MAX_WORLD * XCHG_SLOT_WORDS;
    uint64_t mine[4] = {0, XCHG_FAILED, 0, 0};
    if (!failed) {
#pragma unroll
        for (int i = 0; i < 4; i++) mine[i] = x.totals[i];
        if (x.rank + 1 < x.world) mine[0] += x.gap;  // the message bytes this shard accounts for: its window + the cut behind it
    }
    if (live && !x.wait_only) {  // 1. push into rank `lane`'s buffer (lane == rank: the local one)
        uint64_t* dst = x.peers[lane] + slot0 + (size_t)x.rank * XCHG_SLOT_WORDS;
#pragma unroll
        for (int i = 0; i < 4; i++) st_relaxed_sys(dst + i, mine[i]);
        st_release_sys(dst + 4, x.epoch);  // release: the four stores above are visible before the sequence word
    }
    uint64_t v[4] = {0, 0, 0, 0};
    bool late = false;
    if (live) {  // 2. wait for rank `lane`'s push into the local buffer
        const uint64_t* src = x.local + slot0 + (size_t)lane * XCHG_SLOT_WORDS;
        const uint64_t t0 = xchg_now_ns();
        while (ld_acquire_sys(src + 4) != x.epoch) {
            if (xchg_now_ns() - t0 > x.slice_ns) {
                late = true;
                break;
            }
            __nanosleep(100);
        }
        if (!late) {
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] = ld_relaxed_sys(src + i);
        }
    }
    const uint32_t any_late = __ballot_sync(0xffffffffu, late);
    const uint32_t any_failed = __ballot_sync(0xffffffffu, live && !late && v[1] == XCHG_FAILED);
    if (any_failed) v[1] = v[1] == XCHG_FAILED ? 0 : v[1];
    // 3. inclusive warp scan
    uint64_t inc[4];
#pragma unroll
    for (int i = 0; i < 4; i++) inc[i] = v[i];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, inc[i], d);
            if (lane >= (uint32_t)d) inc[i] += t;
        }
    }
    uint64_t whole[4];
#pragma unroll
    for (int i = 0; i < 4; i++) whole[i] = __shfl_sync(0xffffffffu, inc[i], 31);
    if (lane == x.rank) {
#pragma unroll
        for (int i = 0; i < 4; i++) x.out[i] = inc[i] - v[i];
#pragma unroll
        for (int i = 0; i < 4; i++) x.out[4 + i] = whole[i];
        x.out[8] = any_late ? XCHG_PENDING : (any_failed ? XCHG_PEER_FAILED : XCHG_OK);
        x.out[9] = x.epoch;
    }
}

}  // namespace sj
