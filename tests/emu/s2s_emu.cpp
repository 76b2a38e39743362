// s2s_emu.cpp -- host emulation of the streaming stage 2 (test infrastructure, NOT the product).
//
// Compiles the very templates the CUDA kernels instantiate (simdjson-go_b200/csrc/s2s_slab.h) with a "warp" made of
// 32 ucontext fibers that run in lock step from collective to collective (every ballot / shuffle / reduce is one
// round-robin pass over the fibers), so the kernels' logic -- masks, carries, scans, staging, escape patches -- is
// checked bit for bit against the oracle on a machine without a GPU.  Around it: plain host restatements of the small
// kernels (K2q scan, scope matching, links + grammar masks, roots).  Numbers are returned as (position, slot) pairs;
// the test fills them in with the oracle's parse_number.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#include "../../simdjson-go_b200/csrc/s2s_slab.h"

using namespace sj;

namespace {

struct FiberWarp;
FiberWarp* g_warp = nullptr;

struct FiberWarp {
    static constexpr int N = 32;
    ucontext_t ctx[N], main_ctx;
    std::vector<char> stacks[N];
    int cur = 0;
    bool done[N];
    uint32_t slot[2][N];
    int phase[N];
    long collectives = 0;
    std::function<void(FiberWarp&)> body;

    // the order in which the lanes run between two collectives: ascending, or (-DS2S_EMU_REVERSE) descending -- code that
    // is correct on a GPU must not care, so the suite is run both ways
#ifdef S2S_EMU_REVERSE
    static int first_lane() { return N - 1; }
    static int next_lane(int me) { return (me + N - 1) % N; }
    static bool is_last(int me) { return me == 0; }
#else
    static int first_lane() { return 0; }
    static int next_lane(int me) { return (me + 1) % N; }
    static bool is_last(int me) { return me == N - 1; }
#endif
    static void entry() {
        FiberWarp* w = g_warp;
        w->body(*w);
        const int me = w->cur;
        w->done[me] = true;
        if (!is_last(me)) {
            w->cur = next_lane(me);
            swapcontext(&w->ctx[me], &w->ctx[w->cur]);
        } else {
            swapcontext(&w->ctx[me], &w->main_ctx);
        }
    }
    void run(std::function<void(FiberWarp&)> f) {
        body = std::move(f);
        g_warp = this;
        for (int i = 0; i < N; i++) {
            if (stacks[i].empty()) stacks[i].resize(512 << 10);
            getcontext(&ctx[i]);
            ctx[i].uc_stack.ss_sp = stacks[i].data();
            ctx[i].uc_stack.ss_size = stacks[i].size();
            ctx[i].uc_link = nullptr;
            makecontext(&ctx[i], (void (*)())entry, 0);
            done[i] = false;
            phase[i] = 0;
        }
        cur = first_lane();
        swapcontext(&main_ctx, &ctx[cur]);
        for (int i = 0; i < N; i++)
            if (!done[i]) {
                fprintf(stderr, "s2s_emu: lane %d did not finish (non-uniform collectives)\n", i);
                abort();
            }
    }
    // one collective: publish v, let every other lane reach the same point, return the buffer of all 32 values
    const uint32_t* gather(uint32_t v) {
        const int me = cur, ph = phase[me];
        slot[ph][me] = v;
        phase[me] ^= 1;
        const int nx = next_lane(me);
        if (done[nx]) {
            fprintf(stderr, "s2s_emu: lane %d waits in a collective that lane %d never reached\n", me, nx);
            abort();
        }
        collectives++;
        cur = nx;
        swapcontext(&ctx[me], &ctx[nx]);
        return slot[ph];
    }
    uint32_t lane() const { return (uint32_t)cur; }
    uint32_t ballot(bool p) {
        const uint32_t* s = gather(p ? 1u : 0u);
        uint32_t m = 0;
        for (int i = 0; i < N; i++) m |= (s[i] & 1u) << i;
        return m;
    }
    bool any(bool p) { return ballot(p) != 0; }
    uint32_t shfl(uint32_t v, uint32_t src) { return gather(v)[src & 31]; }
    uint32_t shfl_up(uint32_t v, int d) {
        const int me = cur;
        const uint32_t* s = gather(v);
        return me >= d ? s[me - d] : v;
    }
    uint32_t reduce_add(uint32_t v) {
        const uint32_t* s = gather(v);
        uint32_t t = 0;
        for (int i = 0; i < N; i++) t += s[i];
        return t;
    }
    void sync() { gather(0); }
    void atomic_and(uint32_t* p, uint32_t v) { *p &= v; }
    void atomic_or(uint32_t* p, uint32_t v) { *p |= v; }
    void atomic_or_shared(uint32_t* p, uint32_t v) { *p |= v; }
    void async_copy16(void* dst, const void* src) { memcpy(dst, src, 16); }  // (immediately: the emulation has no latency to hide)
    void async_commit() {}
    void async_wait_prev() {}
};

struct Tables {
    uint8_t ctab[256], oktab[256];
    uint32_t cmptab[16];
    Tables() {
        for (int i = 0; i < 256; i++) ctab[i] = (uint8_t)char_type((uint32_t)i);
        memset(oktab, 0, sizeof oktab);
        for (uint32_t p = 0; p < 15; p++)
            for (uint32_t c = 0; c < 15; c++) oktab[p * 16 + c] = (uint8_t)transition_mask(p, c);
        for (uint32_t m = 0; m < 16; m++) cmptab[m] = compress_sel(m) | ((uint32_t)__builtin_popcount(m) << 16);
    }
};

}  // namespace

// msg must be readable up to round_up(len, 16).  idx: stage 1's structural positions (absolute).
// Returns 0 (ok) or 2 (stage-2 failure).  Numbers: (pos, slot) pairs, the tape slots themselves stay 0.
extern "C" int s2s_emu_parse(const uint8_t* msg, size_t len, int ndjson, const uint32_t* idx, size_t n_idx, uint64_t* tape,
                             size_t tape_cap, size_t* tape_len, uint8_t* strings, size_t strings_cap, size_t* strings_len,
                             uint32_t* num_pos, uint32_t* num_slot, size_t num_cap, size_t* n_num, long* collectives) {
    static Tables T;
    const uint32_t SPT = 16;  // slabs per stage-1 tile
    const uint32_t nslabs = (uint32_t)((len + S2S_SLAB_BYTES - 1) / S2S_SLAB_BYTES);
    // in-string state in front of every slab (what stage 1's chain 1 hands over): escapes apply inside and outside strings
    std::vector<uint32_t> slabpar((nslabs + SPT - 1) / SPT + 1, 0);
    {
        bool in = false, esc = false;
        for (size_t i = 0; i < len; i++) {
            if (i % S2S_SLAB_BYTES == 0 && in) slabpar[(i / S2S_SLAB_BYTES) / SPT] |= 1u << ((i / S2S_SLAB_BYTES) % SPT);
            const uint8_t c = msg[i];
            if (esc)
                esc = false;
            else if (c == '\\')
                esc = true;
            else if (c == '"')
                in = !in;
        }
    }
    std::vector<SlabAgg> agg(nslabs), pre(nslabs), grp_pre((nslabs + 1023) / 1024 + 1);
    std::vector<uint8_t> src(S2S_IMAGE_BYTES + 64), sstage(S2S_SSTAGE_BYTES + 64);
    std::vector<uint64_t> tstage(S2S_TSTAGE_WORDS + 8);
    uint32_t error = 0;
    S2sParams p;
    memset(&p, 0, sizeof p);
    p.msg = msg;
    p.len = len;
    p.ndjson = ndjson ? 1 : 0;
    p.idx = idx;
    p.n_idx = (uint32_t)n_idx;
    p.slabpar = slabpar.data();
    p.slabs_per_tile = SPT;
    p.nslabs = nslabs;
    p.agg = agg.data();
    p.pre = pre.data();
    p.grp_pre = grp_pre.data();
    p.error = &error;
    S2sWarpMem sm;
    // 16-byte aligned working memory
    sm.src = (uint8_t*)(((uintptr_t)src.data() + 15) & ~(uintptr_t)15);
    sm.sstage = (uint8_t*)(((uintptr_t)sstage.data() + 15) & ~(uintptr_t)15);
    sm.tstage = tstage.data();
    sm.esc = reinterpret_cast<uint8_t*>(tstage.data());  // (as in K2r; K2p has nothing else in there)
    sm.ctab = T.ctab;
    sm.oktab = T.oktab;
    sm.cmptab = T.cmptab;
    FiberWarp W;
    // ---- K2p ----
    // (three "warps" of a stride-3 grid, so that the hand-over of the image pipeline from slab to slab is exercised)
    for (uint32_t f = 0; f < 3; f++) W.run([&](FiberWarp& w) { s2s_warp_loop<FiberWarp, false>(w, p, f, 3, sm); });
    // ---- K2q: exclusive scan in groups of 1024 + exclusive scan of the group totals ----
    SlabAgg grand = agg_zero();
    for (uint32_t g0 = 0, gi = 0; g0 < nslabs; g0 += 1024, gi++) {
        grp_pre[gi] = grand;
        SlabAgg acc = agg_zero();
        for (uint32_t i = g0; i < nslabs && i < g0 + 1024; i++) {
            pre[i] = acc;
            acc = agg_combine(acc, agg[i]);
        }
        grand = agg_combine(grand, acc);
    }
    const uint64_t tlen = (uint64_t)grand.w + 2;
    *tape_len = tlen;
    *strings_len = grand.str;
    *n_num = grand.num;
    if (collectives) *collectives = W.collectives;
    if (tlen > tape_cap || grand.str > strings_cap || grand.num > num_cap) return 4;
    const uint32_t nb = grand.brk;
    std::vector<uint32_t> brk_tp(nb + 1), segmask((nb + 1 + 3) / 4 + 1, 0xffffffffu), rootpos(grand.rec + 2, 0);
    std::vector<int32_t> brk_depth(nb + 1);
    std::vector<uint8_t> brk_kind(nb + 1);
    std::vector<NumEntry> numlist(grand.num + 1);
    memset(tape, 0, tlen * 8);
    p.tape = tape;
    p.strings = strings;
    p.brk_tp = brk_tp.data();
    p.brk_depth = brk_depth.data();
    p.brk_kind = brk_kind.data();
    p.segmask = segmask.data();
    p.rootpos = rootpos.data();
    p.numlist = numlist.data();
    // ---- K2r ----
    for (uint32_t f = 0; f < 3; f++) W.run([&](FiberWarp& w) { s2s_warp_loop<FiberWarp, true>(w, p, f, 3, sm); });
    if (collectives) *collectives = W.collectives;
    for (uint32_t i = 0; i < grand.num; i++) {
        num_pos[i] = numlist[i].pos;
        num_slot[i] = numlist[i].slot;
    }
    // ---- K2d: nearest previous bracket with a smaller depth-in-front (= the scope stack) ----
    std::vector<int32_t> par(nb + 1, -1), stack;
    for (uint32_t k = 0; k < nb; k++) {
        while (!stack.empty() && brk_depth[stack.back()] >= brk_depth[k]) stack.pop_back();
        par[k] = stack.empty() ? -1 : stack.back();
        stack.push_back((int32_t)k);
    }
    // ---- links + grammar masks (one step per bracket, plus the segment behind the last one) ----
    for (uint32_t k = 0; k <= nb; k++) {
        uint32_t ctx = CTX_ROOT;
        if (k > 0) {
            const uint32_t kd = brk_kind[k - 1];
            int32_t enc;
            if (kd == T_OBJ_OPEN || kd == T_ARR_OPEN)
                enc = (int32_t)k - 1;
            else {
                const int32_t m = par[k - 1];
                enc = m >= 0 ? par[m] : -1;
            }
            ctx = enc >= 0 ? (brk_kind[enc] == T_OBJ_OPEN ? CTX_OBJ : CTX_ARR) : CTX_ROOT;
        }
        const uint32_t sg = (segmask[k >> 2] >> (8 * (k & 3))) & 0xff;
        if (!((sg >> ctx) & 1)) error |= 1;
        if (k < nb && (brk_kind[k] == T_OBJ_CLOSE || brk_kind[k] == T_ARR_CLOSE)) {
            const int32_t m = par[k];
            if (m >= 0) {
                const uint32_t otp = brk_tp[m], ctp = brk_tp[k];
                tape[otp] = ((uint64_t)(brk_kind[k] == T_OBJ_CLOSE ? '{' : '[') << 56) | ((uint64_t)ctp + 1);
                tape[ctp] = ((uint64_t)(brk_kind[k] == T_OBJ_CLOSE ? '}' : ']') << 56) | otp;
            }
        }
    }
    // ---- K2f roots ----
    for (uint64_t r = 0; r <= grand.rec; r++) {
        const uint64_t R = (uint64_t)'r' << 56;
        const uint64_t open = r == 0 ? 0 : rootpos[r];
        const uint64_t next_open = r == grand.rec ? tlen : rootpos[r + 1];
        if (next_open > tlen || next_open == 0) continue;
        tape[open] = R | next_open;
        tape[next_open - 1] = R | open;
    }
    if (error || grand.depth != 0) return 2;
    return 0;
}

// esc_u_fast against its definition: `img` holds one step (2048 message bytes, natural order); every backslash position o
// is decoded both ways.  Returns the number of positions where the fast path applied, or -(o + 1) at the first
// disagreement.  (The fast path reads the swizzled image, the definition reads the bytes in natural order.)
extern "C" long s2s_emu_esc_fast_check(const uint8_t* step, uint32_t avail) {
    alignas(16) static uint8_t img[S2S_STEP_BYTES];
    for (uint32_t i = 0; i < S2S_STEP_BYTES; i++) img[swz(i)] = step[i];
    const GlobalReader g{step, avail};
    long applied = 0;
    for (uint32_t o = 0; o + 1 < avail; o++) {
        if (step[o] != '\\') continue;
        EscInfo f;
        if (!esc_u_fast(img, o, avail, f)) continue;
        const EscInfo d = esc_decode(g, g, o);
        if (d.second || !d.valid || d.c != f.c || d.n != f.n || d.bytes != f.bytes) return -(long)(o + 1);
        applied++;
    }
    return applied;
}

// atom_ok_fast against atom_ok_p at every position of a step that holds 't', 'f' or 'n' (the bytes are whatever the test
// put there): the number of positions where the fast path applied, or -(o + 1) at the first disagreement
extern "C" long s2s_emu_atom_fast_check(const uint8_t* step, uint32_t avail, uint64_t len) {
    alignas(16) static uint8_t img[S2S_STEP_BYTES];
    for (uint32_t i = 0; i < S2S_STEP_BYTES; i++) img[swz(i)] = step[i];
    const GlobalReader g{step, len};
    long applied = 0;
    for (uint32_t o = 0; o < avail; o++) {
        uint32_t ch = 0;
        bool ok = false;
        if (!atom_ok_fast(img, o, avail, &ch, &ok)) continue;
        if (ch != step[o]) return -(long)(o + 1);
        const uint32_t type = char_type(ch);
        if (type != T_TRUE && type != T_FALSE && type != T_NULL) return -(long)(o + 1);
        if (atom_ok_p(g, o, len, type) != ok) return -(long)(o + 1);
        applied++;
    }
    return applied;
}
