// sj_parse.inl -- parseMessage replacement: host trim + upload, K1, K2a..K2f, read-back.
// Included by sj_api.cu.

// ---------------------------------------------------------------------------------
// bytes.TrimSpace (parse_json_amd64.go:55; Go semantics incl. the unicode.IsSpace fall-back
// once a byte >= 0x80 is met -- SURVEY.md appendix C.1).  Host-side pointer work only.
// ---------------------------------------------------------------------------------
namespace {

inline bool ascii_space(uint8_t c) { return c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r' || c == ' '; }

inline bool unicode_space(uint32_t r) {
    if (r < 0x80) return ascii_space((uint8_t)r);
    if (r == 0x85 || r == 0xA0 || r == 0x1680 || r == 0x2028 || r == 0x2029 || r == 0x202F || r == 0x205F || r == 0x3000)
        return true;
    return r >= 0x2000 && r <= 0x200A;
}

// utf8.DecodeRune: width (>= 1 for n > 0); invalid encodings give U+FFFD, width 1
inline int decode_rune(const uint8_t* p, size_t n, uint32_t* r) {
    *r = 0xFFFD;
    if (n == 0) return 0;
    uint8_t c = p[0];
    if (c < 0x80) {
        *r = c;
        return 1;
    }
    int need;
    uint32_t cp, lo;
    if (c >= 0xC2 && c <= 0xDF) {
        need = 1, cp = c & 0x1F, lo = 0x80;
    } else if (c >= 0xE0 && c <= 0xEF) {
        need = 2, cp = c & 0x0F, lo = 0x800;
    } else if (c >= 0xF0 && c <= 0xF4) {
        need = 3, cp = c & 0x07, lo = 0x10000;
    } else {
        return 1;
    }
    if (n < (size_t)need + 1) return 1;
    for (int i = 1; i <= need; i++) {
        if ((p[i] & 0xC0) != 0x80) return 1;
        cp = (cp << 6) | (p[i] & 0x3F);
    }
    if (cp < lo || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return 1;
    *r = cp;
    return need + 1;
}

void trim_runes(const uint8_t* b, size_t* start, size_t* stop) {
    size_t a = *start, e = *stop;
    while (a < e) {
        uint32_t r;
        int w = decode_rune(b + a, e - a, &r);
        if (!unicode_space(r)) break;
        a += (size_t)w;
    }
    while (e > a) {
        uint32_t r = 0xFFFD;
        size_t w = 1;
        if (b[e - 1] < 0x80) {
            r = b[e - 1];
        } else {
            size_t lim = e - a < 4 ? e - a : 4;
            for (size_t back = 1; back <= lim; back++) {
                if ((b[e - back] & 0xC0) != 0x80) {  // first non-continuation byte from the end
                    uint32_t rr;
                    int ww = decode_rune(b + e - back, back, &rr);
                    if ((size_t)ww == back) {
                        r = rr;
                        w = back;
                    }
                    break;
                }
            }
        }
        if (!unicode_space(r)) break;
        e -= w;
    }
    *start = a;
    *stop = e;
}

void trim_space(const uint8_t* b, size_t len, size_t* start_out, size_t* stop_out) {
    size_t start = 0, stop = len;
    bool done = false;
    for (; start < len; start++) {
        if (b[start] >= 0x80) {
            trim_runes(b, &start, &stop);
            done = true;
            break;
        }
        if (!ascii_space(b[start])) break;
    }
    if (!done) {
        for (; stop > start; stop--) {
            if (b[stop - 1] >= 0x80) {
                trim_runes(b, &start, &stop);
                break;
            }
            if (!ascii_space(b[stop - 1])) break;
        }
    }
    *start_out = start;
    *stop_out = stop;
}

}  // namespace

extern "C" void sj_trim_space(const uint8_t* msg, size_t len, size_t* start, size_t* stop) {
    size_t a = 0, b = 0;
    if (len) trim_space(msg, len, &a, &b);
    if (start) *start = a;
    if (stop) *stop = b;
}

namespace {

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Carver {  // sub-allocates one DevBuf
    uint8_t* base;
    size_t off = 0;
    explicit Carver(void* p) : base(reinterpret_cast<uint8_t*>(p)) {}
    template <typename T>
    T* take(size_t count) {
        T* r = reinterpret_cast<T*>(base + off);
        off += align_up(count * sizeof(T) + 16, 256);
        return r;
    }
    static size_t need(std::initializer_list<size_t> bytes) {
        size_t t = 0;
        for (size_t b : bytes) t += align_up(b + 16, 256);
        return t;
    }
};

}  // namespace

// ---------------------------------------------------------------------------------
// stage 2 driver.  d_tape / d_strings may be null: then the context's own output buffers
// are sized from the totals (host-buffer API).
// ---------------------------------------------------------------------------------
static int stage2_verdict(const Stage2Result& r);

// what the counting half of stage 2 leaves for the emitting half (between them the totals are on the host: output
// buffers are sized, and the shards of a multi-GPU ParseND exchange their totals)
struct S2Pending {
    bool valid, stream;
    Stage2Result tot;
    Stage2Params lp;   // per-structural kernels
    S2sParams sp;      // streaming kernels
};
static S2Pending* pending_of(sj_ctx* c) {
    if (!c->pending) {
        c->pending = calloc(1, sizeof(S2Pending));
    }
    return reinterpret_cast<S2Pending*>(c->pending);
}
struct ParseBases {
    uint64_t msg, tape, str;
    const uint64_t* dev;  // or: the three of them in device memory
};
// host destinations of the host-buffer API: copied behind the last kernel, in front of the call's one final sync
struct HostOut {
    uint64_t* tape;
    size_t tape_cap;
    uint8_t* strings;
    size_t strings_cap;
};

// per-structural kernels, counting half: K2a classify / measure, K2b scans, totals to the host
static int legacy_count(sj_ctx* c, const uint8_t* d_msg, size_t len, const uint32_t* d_idx, uint32_t n, uint32_t flags,
                        const uint32_t* d_bsmap, Stage2Result* out, uint64_t* d_totals = nullptr) {
    S2Pending* pd = pending_of(c);
    if (!pd) return SJ_ERR_ARGUMENT;
    pd->valid = false;
    const uint32_t ntiles = (n + S2_TILE - 1) / S2_TILE;
    const uint32_t ngroups = (ntiles + 1023) / 1024;
    c->last_tape = nullptr;  // the device-side results of the previous parse are about to be overwritten
    // ---- phase 1 scratch ----
    size_t need1 = Carver::need({(size_t)n + 16, ((size_t)n + 16) * 4, (size_t)ntiles * sizeof(ScanVal),
                                 (size_t)ntiles * sizeof(ScanVal), (size_t)ntiles * (S2_TILE / 32) * sizeof(ScanVal),
                                 (size_t)ngroups * sizeof(ScanVal), (size_t)ngroups * sizeof(ScanVal)});
    int rc = c->s2a.reserve(need1);
    if (rc) return rc;
    Carver k1(c->s2a.p);
    Stage2Params p;
    memset(&p, 0, sizeof p);
    p.msg = d_msg;
    p.len = len;
    p.idx = d_idx;
    p.bsmap = d_bsmap;
    p.n = n;
    p.ndjson = (flags & SJ_FLAG_NDJSON) ? 1 : 0;
    p.copy_strings = (flags & SJ_FLAG_COPY_STRINGS) ? 1 : 0;
    p.typ = k1.take<uint8_t>(n);
    p.aux = k1.take<uint32_t>(n);
    p.tile_sum = k1.take<ScanVal>(ntiles);
    p.tile_pre = k1.take<ScanVal>(ntiles);
    p.sub_pre = k1.take<ScanVal>((size_t)ntiles * (S2_TILE / 32));
    p.grp_sum = k1.take<ScanVal>(ngroups);
    p.grp_pre = k1.take<ScanVal>(ngroups);
    p.ntiles = ntiles;
    p.ngroups = ngroups;
    Stage2Result* d_res = reinterpret_cast<Stage2Result*>(c->result.as<uint8_t>() + 64);
    p.result = d_res;
    SJ_CUDA_CHECK(cudaMemsetAsync(d_res, 0, sizeof(Stage2Result), c->stream));

    s2_classify_measure_kernel<<<ntiles, S2_THREADS, 0, c->stream>>>(p);
    s2_scan_groups_kernel<<<ngroups, 1024, 0, c->stream>>>(p.tile_sum, ntiles, p.tile_pre, p.grp_sum);
    s2_scan_top_kernel<<<1, 1024, 0, c->stream>>>(p.grp_sum, ngroups, p.grp_pre, d_res, d_totals, (uint64_t)len);
    c->launches += 3;
    SJ_CUDA_CHECK(cudaGetLastError());
    rc = exchange_enqueue(c, d_totals, false, nullptr, nullptr);  // sharded ParseND with the exchange set up: push / wait / prefix, still in front of the read-back
    if (rc) return rc;
    Stage2Result* h_res = reinterpret_cast<Stage2Result*>(reinterpret_cast<uint8_t*>(c->host_result) + 64);
    SJ_CUDA_CHECK(cudaMemcpyAsync(h_res, d_res, sizeof(Stage2Result), cudaMemcpyDeviceToHost, c->stream));
    SJ_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    pd->tot = *h_res;
    pd->lp = p;
    pd->stream = false;
    pd->valid = true;
    *out = pd->tot;
    return SJ_OK;
}

// per-structural kernels, emitting half
static int legacy_emit(sj_ctx* c, uint64_t* d_tape, size_t tape_cap, uint8_t* d_strings, size_t strings_cap, const ParseBases& bases,
                       Stage2Result* out, const HostOut* host = nullptr) {
    S2Pending* pd = pending_of(c);
    if (!pd || !pd->valid || pd->stream) return SJ_ERR_ARGUMENT;
    pd->valid = false;
    Stage2Params p = pd->lp;
    const Stage2Result tot = pd->tot;
    const uint32_t n = p.n, ntiles = p.ntiles;
    const uint8_t* d_msg = p.msg;
    Stage2Result* d_res = p.result;
    Stage2Result* h_res = reinterpret_cast<Stage2Result*>(reinterpret_cast<uint8_t*>(c->host_result) + 64);
    int rc;
    p.tape_base = bases.tape;
    p.str_base = bases.str;
    p.bases_dev = bases.dev;
    p.msg_base = bases.msg;
    p.bases_dev = bases.dev;

    // ---- outputs ----
    if (host && (tot.tape_len > host->tape_cap || tot.strings_len > host->strings_cap)) {
        *out = tot;
        return SJ_ERR_CAPACITY;
    }
    if (!d_tape) {
        rc = c->tape.reserve(tot.tape_len * 8 + 64);
        if (rc) return rc;
        rc = c->strings.reserve(tot.strings_len + 64);
        if (rc) return rc;
        d_tape = c->tape.as<uint64_t>();
        tape_cap = tot.tape_len;
        d_strings = c->strings.as<uint8_t>();
        strings_cap = tot.strings_len;
    } else if (tot.tape_len > tape_cap || tot.strings_len > strings_cap) {
        return SJ_ERR_CAPACITY;
    }
    p.tape = d_tape;
    p.tape_cap = tape_cap;
    p.strings = d_strings;
    p.strings_cap = strings_cap;

    // ---- phase 2 scratch ----
    const size_t nb = (size_t)tot.n_brackets;
    size_t lvl_total = 0;
    {
        size_t sz = nb;
        while (sz > 32) {
            sz = (sz + 31) / 32;
            lvl_total += sz;
        }
    }
    // numbers get their own dense kernels when at least one structural in 16 is a number
    const uint32_t n_num = tot.n_numbers;
    const bool dense_numbers = S2_DENSE_NUMBERS && n_num != 0 && ((uint64_t)n_num << SJ_S2_DENSE_NUMBERS_SHIFT) >= n;
    size_t need2 = Carver::need({nb * 4, nb * 4, nb * 4, nb * 4, nb * 4, nb, (lvl_total + 8) * 4, ((size_t)tot.n_records + 2) * 4,
                                 dense_numbers ? (size_t)n_num * 4 : 0});
    rc = c->s2b.reserve(need2);
    if (rc) return rc;
    Carver k2(c->s2b.p);
    p.brk_i = k2.take<uint32_t>(nb);
    p.brk_tp = k2.take<uint32_t>(nb);
    p.brk_depth = k2.take<int32_t>(nb);
    p.par = k2.take<int32_t>(nb);
    p.enc_after = k2.take<int32_t>(nb);
    p.ctx_after = k2.take<uint8_t>(nb);
    int32_t* lvl_mem = k2.take<int32_t>(lvl_total + 8);
    p.rootpos = k2.take<uint32_t>((size_t)tot.n_records + 2);
    p.numlist = dense_numbers ? k2.take<uint32_t>(n_num) : nullptr;

    s2_emit_kernel<<<(n + S2_THREADS - 1) / S2_THREADS, S2_THREADS, 0, c->stream>>>(p);
    c->launches++;
    if (dense_numbers) {
        s2_numlist_kernel<<<(n + 1023) / 1024, 1024, 0, c->stream>>>(p, n_num);
        s2_numbers_kernel<<<(n_num + S2_THREADS - 1) / S2_THREADS, S2_THREADS, 0, c->stream>>>(p, n_num);
        c->launches += 2;
    }
    if (nb > 0) {
        AnsvLevels L;
        memset(&L, 0, sizeof L);
        L.lv[0] = p.brk_depth;
        L.n[0] = (uint32_t)nb;
        L.nlevels = 1;
        size_t sz = nb;
        int32_t* next = lvl_mem;
        while (sz > 32 && L.nlevels < ANSV_MAX_LEVELS) {
            size_t nsz = (sz + 31) / 32;
            const unsigned threads = 256;
            const unsigned blocks = (unsigned)((nsz * 32 + threads - 1) / threads);
            s2_min32_kernel<<<blocks, threads, 0, c->stream>>>(L.lv[L.nlevels - 1], (uint32_t)sz, next, (uint32_t)nsz);
            c->launches++;
            L.lv[L.nlevels] = next;
            L.n[L.nlevels] = (uint32_t)nsz;
            L.nlevels++;
            next += nsz;
            sz = nsz;
        }
        s2_ansv_kernel<<<(unsigned)((nb + S2_THREADS - 1) / S2_THREADS), S2_THREADS, 0, c->stream>>>(L, p.par);
        s2_scope_kernel<<<(unsigned)((nb + S2_THREADS - 1) / S2_THREADS), S2_THREADS, 0, c->stream>>>(p, (uint32_t)nb);
        c->launches += 2;
    }
    s2_grammar_kernel<<<ntiles, S2_THREADS, 0, c->stream>>>(p);
    {
        const uint64_t nrec = tot.n_records;
        const unsigned blocks = (unsigned)((nrec + 1 + 255) / 256);
        s2_roots_kernel<<<blocks, 256, 0, c->stream>>>(p, nrec, tot.tape_len);
    }
    c->launches += 2;
    SJ_CUDA_CHECK(cudaGetLastError());
    if (host) {  // (on a stage-2 failure the copied words are meaningless; the verdict below says so)
        SJ_CUDA_CHECK(cudaMemcpyAsync(host->tape, d_tape, tot.tape_len * 8, cudaMemcpyDeviceToHost, c->stream));
        if (tot.strings_len) SJ_CUDA_CHECK(cudaMemcpyAsync(host->strings, d_strings, tot.strings_len, cudaMemcpyDeviceToHost, c->stream));
    }
    SJ_CUDA_CHECK(cudaMemcpyAsync(h_res, d_res, sizeof(Stage2Result), cudaMemcpyDeviceToHost, c->stream));
    SJ_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *out = *h_res;
    if (stage2_verdict(*out) == SJ_OK && bases.tape == 0 && !bases.dev) {  // what the tape consumers (sj_consume.inl) may read until the next call
        c->last_rootpos = p.rootpos;
        c->last_records = tot.n_records;
        c->last_tape = d_tape;
        c->last_tape_len = tot.tape_len;
        c->last_strings = d_strings;
        c->last_msg = d_msg;
    }
    return SJ_OK;
}

static int stage2_verdict(const Stage2Result& r) {
    if (r.overflow) return SJ_ERR_CAPACITY;
    if (r.error || r.final_depth != 0) return SJ_ERR_STAGE2;  // stage2...go:428-436: leftover scope
    return SJ_OK;
}

// stage 1 into the context's index buffer (absolute positions), growing it once if needed.  `stream`: stage 2 will
// be the streaming kernels (they take the per-slab in-string bits from K1 instead of the per-block backslash map)
static int stage1_positions(sj_ctx* c, const uint8_t* d_msg, size_t len, bool ndjson, Stage1Result* r, bool stream) {
    size_t dcap = len / 4 + 1024;
    if (c->idx.cap / sizeof(uint32_t) > dcap) dcap = c->idx.cap / sizeof(uint32_t);
    // one bit per 64-byte block, one word per 2 KiB step, rounded up to whole tiles
    const size_t bs_words = ((len + S1_TILE_BYTES - 1) / S1_TILE_BYTES) * (S1_TILE_BYTES / S1_STEP_BYTES) + 64;
    if (!stream) {
        int rcb = c->s2c.reserve(bs_words * sizeof(uint32_t));
        if (rcb) return rcb;
    }
    for (int attempt = 0; attempt < 2; attempt++) {
        int rc = c->idx.reserve(dcap * sizeof(uint32_t));
        if (rc) return rc;
        rc = launch_stage1(c, d_msg, len, ndjson, false, c->idx.as<uint32_t>(), dcap, stream ? nullptr : c->s2c.as<uint32_t>(), stream);
        if (rc) return rc;
        rc = fetch_stage1_result(c, r);
        if (rc) return rc;
        if (!r->overflow) return SJ_OK;
        dcap = (size_t)r->n_idx + 64;
    }
    return SJ_ERR_CAPACITY;
}

static inline bool use_stream_stage2(const sj_ctx* c, uint32_t flags) {
    return c->s2_impl == 0 && (flags & SJ_FLAG_COPY_STRINGS) != 0;
}

// ---------------------------------------------------------------------------------
// stage 2, streaming kernels (stage2_stream.cuh): K2p count -> K2q scan -> [totals to the host] -> K2r emit ->
// K2h numbers, K2d scope matching, K2e links + grammar verdict, K2f roots.  copy_strings only (options.go:13 default).
// ---------------------------------------------------------------------------------
static int stream_count(sj_ctx* c, const uint8_t* d_msg, size_t len, const uint32_t* d_idx, uint32_t n, uint32_t flags, Stage2Result* out,
                        uint64_t* d_totals = nullptr) {
    S2Pending* pd = pending_of(c);
    if (!pd) return SJ_ERR_ARGUMENT;
    pd->valid = false;
    const uint32_t nslabs = (uint32_t)((len + S2S_SLAB_BYTES - 1) / S2S_SLAB_BYTES);
    const uint32_t ngroups = (nslabs + 1023) / 1024;
    c->last_tape = nullptr;
    if (!c->last_slabpar) return SJ_ERR_ARGUMENT;  // stage 1 did not run in streaming mode
    size_t need1 = Carver::need({(size_t)nslabs * sizeof(SlabAgg), (size_t)nslabs * sizeof(SlabAgg), (size_t)ngroups * sizeof(SlabAgg),
                                 (size_t)ngroups * sizeof(SlabAgg)});
    int rc = c->s2a.reserve(need1);
    if (rc) return rc;
    Carver k1(c->s2a.p);
    S2sParams p;
    memset(&p, 0, sizeof p);
    p.msg = d_msg;
    p.len = len;
    p.ndjson = (flags & SJ_FLAG_NDJSON) ? 1 : 0;
    p.idx = d_idx;
    p.n_idx = n;
    p.slabpar = c->last_slabpar;
    p.slabs_per_tile = S1_WARPS;
    p.nslabs = nslabs;
    p.agg = k1.take<SlabAgg>(nslabs);
    SlabAgg* pre = k1.take<SlabAgg>(nslabs);
    SlabAgg* grp_sum = k1.take<SlabAgg>(ngroups);
    SlabAgg* grp_pre = k1.take<SlabAgg>(ngroups);
    p.pre = pre;
    p.grp_pre = grp_pre;
    Stage2Result* d_res = reinterpret_cast<Stage2Result*>(c->result.as<uint8_t>() + 64);
    p.error = &d_res->error;
    SJ_CUDA_CHECK(cudaMemsetAsync(d_res, 0, sizeof(Stage2Result), c->stream));
    // persistent warps: each walks slabs w, w + W, ... with the next step's image always in flight
    unsigned grid = (nslabs + S2S_WARPS - 1) / S2S_WARPS;
    if (grid > (unsigned)(c->sm_count * SJ_S2S_COUNT_MIN_BLOCKS)) grid = (unsigned)(c->sm_count * SJ_S2S_COUNT_MIN_BLOCKS);
    s2s_count_kernel<<<grid, S2S_THREADS, S2S_SMEM_COUNT, c->stream>>>(p);
    s2s_scan_groups_kernel<<<ngroups, 1024, 0, c->stream>>>(p.agg, nslabs, pre, grp_sum);
    s2s_scan_top_kernel<<<1, 1024, 0, c->stream>>>(grp_sum, ngroups, grp_pre, d_res, d_totals, (uint64_t)len);
    c->launches += 3;
    SJ_CUDA_CHECK(cudaGetLastError());
    // sharded ParseND with the exchange set up: push / wait / prefix, still in front of the read-back.  The kernel reads stage
    // 1's verdict on the device: overflow of the index buffer = skip (this counting half is repeated and the repeat publishes)
    rc = exchange_enqueue(c, d_totals, false, c->result.as<Stage1Result>(), &d_res->error);
    if (rc) return rc;
    // one read-back for both stages: the stage-1 result block sits in front of the stage-2 totals, and K1 ran on the same stream
    Stage2Result* h_res = reinterpret_cast<Stage2Result*>(reinterpret_cast<uint8_t*>(c->host_result) + 64);
    SJ_CUDA_CHECK(cudaMemcpyAsync(c->host_result, c->result.p, 64 + sizeof(Stage2Result), cudaMemcpyDeviceToHost, c->stream));
    SJ_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    pd->tot = *h_res;
    pd->sp = p;
    pd->stream = true;
    *out = pd->tot;
    if (pd->tot.error) return SJ_ERR_STAGE2;  // an invalid escape, found by the counting pass: the emitting pass is not run on it
    pd->valid = true;
    return SJ_OK;
}

static int stream_emit(sj_ctx* c, uint64_t* d_tape, size_t tape_cap, uint8_t* d_strings, size_t strings_cap, const ParseBases& bases,
                       Stage2Result* out, const HostOut* host = nullptr) {
    S2Pending* pd = pending_of(c);
    if (!pd || !pd->valid || !pd->stream) return SJ_ERR_ARGUMENT;
    pd->valid = false;
    S2sParams p = pd->sp;
    const Stage2Result tot = pd->tot;
    const uint8_t* d_msg = p.msg;
    const size_t len = p.len;
    unsigned grid = (p.nslabs + S2S_WARPS - 1) / S2S_WARPS;
    if (grid > (unsigned)(c->sm_count * SJ_S2S_EMIT_MIN_BLOCKS)) grid = (unsigned)(c->sm_count * SJ_S2S_EMIT_MIN_BLOCKS);
    Stage2Result* d_res = reinterpret_cast<Stage2Result*>(c->result.as<uint8_t>() + 64);
    Stage2Result* h_res = reinterpret_cast<Stage2Result*>(reinterpret_cast<uint8_t*>(c->host_result) + 64);
    int rc;
    p.tape_base = bases.tape;
    p.str_base = bases.str;
    p.bases_dev = bases.dev;

    if (host && (tot.tape_len > host->tape_cap || tot.strings_len > host->strings_cap)) {
        *out = tot;
        return SJ_ERR_CAPACITY;
    }
    if (!d_tape) {
        rc = c->tape.reserve(tot.tape_len * 8 + 64);
        if (rc) return rc;
        rc = c->strings.reserve(tot.strings_len + 64);
        if (rc) return rc;
        d_tape = c->tape.as<uint64_t>();
        tape_cap = tot.tape_len;
        d_strings = c->strings.as<uint8_t>();
        strings_cap = tot.strings_len;
    } else if (tot.tape_len > tape_cap || tot.strings_len > strings_cap) {
        return SJ_ERR_CAPACITY;
    }
    p.tape = d_tape;
    p.strings = d_strings;

    const size_t nb = (size_t)tot.n_brackets;
    size_t lvl_total = 0;
    {
        size_t sz = nb;
        while (sz > 32) {
            sz = (sz + 31) / 32;
            lvl_total += sz;
        }
    }
    const size_t seg_words = (nb + 1 + 3) / 4;
    const uint32_t n_num = tot.n_numbers;
    size_t need2 = Carver::need({nb * 4, nb * 4, nb, nb * 4, (lvl_total + 8) * 4, seg_words * 4, ((size_t)tot.n_records + 2) * 4,
                                 (size_t)n_num * sizeof(NumEntry)});
    rc = c->s2b.reserve(need2);
    if (rc) return rc;
    Carver k2(c->s2b.p);
    p.brk_tp = k2.take<uint32_t>(nb);
    p.brk_depth = k2.take<int32_t>(nb);
    p.brk_kind = k2.take<uint8_t>(nb);
    int32_t* par = k2.take<int32_t>(nb);
    int32_t* lvl_mem = k2.take<int32_t>(lvl_total + 8);
    p.segmask = k2.take<uint32_t>(seg_words);
    p.rootpos = k2.take<uint32_t>((size_t)tot.n_records + 2);
    p.numlist = k2.take<NumEntry>(n_num);
    SJ_CUDA_CHECK(cudaMemsetAsync(p.segmask, 0xff, seg_words * 4, c->stream));

    s2s_emit_kernel<<<grid, S2S_THREADS, S2S_SMEM_EMIT, c->stream>>>(p);
    c->launches++;
    if (n_num) {
        s2s_numbers_kernel<<<(n_num + S2_THREADS - 1) / S2_THREADS, S2_THREADS, 0, c->stream>>>(d_msg, len, p.numlist, n_num, d_tape, p.error);
        c->launches++;
    }
    if (nb > 0) {
        AnsvLevels L;
        memset(&L, 0, sizeof L);
        L.lv[0] = p.brk_depth;
        L.n[0] = (uint32_t)nb;
        L.nlevels = 1;
        size_t sz = nb;
        int32_t* next = lvl_mem;
        while (sz > 32 && L.nlevels < ANSV_MAX_LEVELS) {
            size_t nsz = (sz + 31) / 32;
            const unsigned threads = 256;
            const unsigned blocks = (unsigned)((nsz * 32 + threads - 1) / threads);
            s2_min32_kernel<<<blocks, threads, 0, c->stream>>>(L.lv[L.nlevels - 1], (uint32_t)sz, next, (uint32_t)nsz);
            c->launches++;
            L.lv[L.nlevels] = next;
            L.n[L.nlevels] = (uint32_t)nsz;
            L.nlevels++;
            next += nsz;
            sz = nsz;
        }
        s2_ansv_kernel<<<(unsigned)((nb + S2_THREADS - 1) / S2_THREADS), S2_THREADS, 0, c->stream>>>(L, par);
        c->launches++;
    }
    {
        const uint64_t nrec = tot.n_records;
        const uint64_t threads = (nb + 1 > nrec + 1 ? nb + 1 : nrec + 1);
        s2s_link_kernel<<<(unsigned)((threads + S2_THREADS - 1) / S2_THREADS), S2_THREADS, 0, c->stream>>>(p, par, (uint32_t)nb, nrec, tot.tape_len);
    }
    c->launches += 1;  // (links and roots share a launch)
    SJ_CUDA_CHECK(cudaGetLastError());
    if (host) {  // (on a stage-2 failure the copied words are meaningless; the verdict below says so)
        SJ_CUDA_CHECK(cudaMemcpyAsync(host->tape, d_tape, tot.tape_len * 8, cudaMemcpyDeviceToHost, c->stream));
        if (tot.strings_len) SJ_CUDA_CHECK(cudaMemcpyAsync(host->strings, d_strings, tot.strings_len, cudaMemcpyDeviceToHost, c->stream));
    }
    SJ_CUDA_CHECK(cudaMemcpyAsync(h_res, d_res, sizeof(Stage2Result), cudaMemcpyDeviceToHost, c->stream));
    SJ_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *out = *h_res;
    if (stage2_verdict(*out) == SJ_OK && bases.tape == 0 && !bases.dev) {
        c->last_rootpos = p.rootpos;
        c->last_records = tot.n_records;
        c->last_tape = d_tape;
        c->last_tape_len = tot.tape_len;
        c->last_strings = d_strings;
        c->last_msg = d_msg;
    }
    return SJ_OK;
}

// stage 2 by whichever implementation the context and the flags select: counting half, emitting half
static int stage2_count_any(sj_ctx* c, const uint8_t* d_msg, size_t len, uint32_t n, uint32_t flags, Stage2Result* out,
                            uint64_t* d_totals = nullptr) {
    if (use_stream_stage2(c, flags)) return stream_count(c, d_msg, len, c->idx.as<uint32_t>(), n, flags, out, d_totals);
    return legacy_count(c, d_msg, len, c->idx.as<uint32_t>(), n, flags, c->s2c.as<uint32_t>(), out, d_totals);
}
static int stage2_emit_any(sj_ctx* c, uint64_t* d_tape, size_t tape_cap, uint8_t* d_strings, size_t strings_cap, const ParseBases& bases,
                           Stage2Result* out, const HostOut* host = nullptr) {
    S2Pending* pd = pending_of(c);
    if (!pd || !pd->valid) return SJ_ERR_ARGUMENT;
    return pd->stream ? stream_emit(c, d_tape, tape_cap, d_strings, strings_cap, bases, out, host)
                      : legacy_emit(c, d_tape, tape_cap, d_strings, strings_cap, bases, out, host);
}

// Stage 1 and the counting half of stage 2.  With the streaming kernels nothing between K1 and the totals needs the
// host, so K1, K2p and K2q are enqueued back to back and ONE read-back delivers the stage-1 verdict and the totals
// (a document of a few hundred KB is launch- and synchronisation-bound: this halves its host round trips).
static int front_half(sj_ctx* c, const uint8_t* d_msg, size_t len, uint32_t flags, Stage1Result* r1, Stage2Result* r2, uint64_t* d_totals = nullptr) {
    const bool ndjson = (flags & SJ_FLAG_NDJSON) != 0;
    if (!use_stream_stage2(c, flags)) {
        int rc = stage1_positions(c, d_msg, len, ndjson, r1, false);
        if (rc) return rc;
        const uint8_t last_char = r1->n_idx && r1->last_pos < len ? (uint8_t)r1->last_char : 0;
        if (!stage1_ok(*r1, last_char)) return SJ_ERR_STAGE1;
        return stage2_count_any(c, d_msg, len, r1->n_idx, flags, r2, d_totals);
    }
    size_t dcap = len / 4 + 1024;
    if (c->idx.cap / sizeof(uint32_t) > dcap) dcap = c->idx.cap / sizeof(uint32_t);
    for (int attempt = 0; attempt < 2; attempt++) {
        int rc = c->idx.reserve(dcap * sizeof(uint32_t));
        if (rc) return rc;
        rc = launch_stage1(c, d_msg, len, ndjson, false, c->idx.as<uint32_t>(), dcap, nullptr, true);
        if (rc) return rc;
        rc = stream_count(c, d_msg, len, c->idx.as<uint32_t>(), 0, flags, r2, d_totals);  // (its read-back brings the stage-1 block too)
        if (rc && rc != SJ_ERR_STAGE2) return rc;
        memcpy(r1, c->host_result, sizeof(Stage1Result));
        if (!r1->overflow) {
            const uint8_t last_char = r1->n_idx && r1->last_pos < len ? (uint8_t)r1->last_char : 0;
            if (!stage1_ok(*r1, last_char)) {  // a stage-1 error wins over a stage-2 one (parse_json_amd64.go:123-126)
                pending_of(c)->valid = false;
                return SJ_ERR_STAGE1;
            }
            return rc;
        }
        dcap = (size_t)r1->n_idx + 64;  // the index buffer was too small (more than one structural in four bytes): once more
        exchange_rearm(c);
    }
    return SJ_ERR_CAPACITY;
}
extern "C" int sj_parse_device(sj_ctx* c, const uint8_t* d_msg, size_t len, uint32_t flags, uint64_t* d_tape,
                               size_t tape_cap, size_t* tape_len, uint8_t* d_strings, size_t strings_cap,
                               size_t* strings_len) {
    if (!c || !tape_len || !strings_len || !d_tape || !d_strings) return SJ_ERR_ARGUMENT;
    *tape_len = 0;
    *strings_len = 0;
    if (len == 0) return SJ_ERR_STAGE1;
    if (len > SJ_MAX_MESSAGE) return SJ_ERR_TOO_LARGE;
    SJ_CUDA_CHECK(cudaSetDevice(c->device));
    Stage1Result r1;
    Stage2Result r2{};
    int rc = front_half(c, d_msg, len, flags, &r1, &r2);
    if (rc) return rc;
    rc = stage2_emit_any(c, d_tape, tape_cap, d_strings, strings_cap, ParseBases{0, 0, 0, nullptr}, &r2);
    if (rc == SJ_OK || rc == SJ_ERR_CAPACITY) {  // the required sizes (simdjson_b200.h): stage 2 got as far as its totals
        *tape_len = r2.tape_len;
        *strings_len = r2.strings_len;
    }
    if (rc) return rc;
    return stage2_verdict(r2);
}

// ---------------------------------------------------------------------------------
// ParseND sharded over several GPUs (simdjson_amd64.go:82-93 returns ONE ParsedJson).  Every rank
// owns a newline-delimited shard of the message.  The counting half of stage 2 ends with the shard's totals on the host
// anyway (that is where output buffers are sized); the ranks exchange them (one all-gather of four integers, NCCL in
// the caller), and the emitting half then writes the shard's tape / Strings.B as the slice of the WHOLE result that
// starts at the exclusive prefix of the totals: root chaining, scope pointers and string offsets are written with the
// bases already added, so there is no separate rebasing pass.
// ---------------------------------------------------------------------------------
extern "C" int sj_parse_nd_sharded_count(sj_ctx* c, const uint8_t* d_msg, size_t len, uint32_t flags, sj_shard_totals* totals,
                                         uint64_t* d_totals) {
    if (!c || !totals) return SJ_ERR_ARGUMENT;
    memset(totals, 0, sizeof *totals);
    SJ_CUDA_CHECK(cudaSetDevice(c->device));
    flags |= SJ_FLAG_NDJSON;
    Stage1Result r1;
    Stage2Result r2{};
    exchange_begin(c, &d_totals);
    int rc = len == 0 ? SJ_ERR_STAGE1 : len > SJ_MAX_MESSAGE ? SJ_ERR_TOO_LARGE : front_half(c, d_msg, len, flags, &r1, &r2, d_totals);
    rc = exchange_end(c, rc);  // (a counting half that failed before its totals existed tells the peers so)
    if (rc) return rc;
    totals->msg_bytes = len;
    totals->tape_words = r2.tape_len;
    totals->string_bytes = r2.strings_len;
    totals->records = r2.n_records + 1;
    return SJ_OK;
}

extern "C" int sj_parse_nd_sharded_emit(sj_ctx* c, uint64_t msg_base, uint64_t tape_base, uint64_t strings_base, const uint64_t* d_bases,
                                        uint64_t* d_tape, size_t tape_cap, uint8_t* d_strings, size_t strings_cap) {
    if (!c || !d_tape || !d_strings) return SJ_ERR_ARGUMENT;
    SJ_CUDA_CHECK(cudaSetDevice(c->device));
    Stage2Result r2{};
    int rc = stage2_emit_any(c, d_tape, tape_cap, d_strings, strings_cap, ParseBases{msg_base, tape_base, strings_base, d_bases}, &r2);
    if (rc) return rc;
    return stage2_verdict(r2);
}

extern "C" int sj_parse(sj_ctx* c, const uint8_t* msg, size_t len, uint32_t flags, uint64_t* tape, size_t tape_cap,
                        size_t* tape_len, uint8_t* strings, size_t strings_cap, size_t* strings_len, size_t* msg_off,
                        size_t* msg_len) {
    if (!c || !tape_len || !strings_len) return SJ_ERR_ARGUMENT;
    *tape_len = 0;
    *strings_len = 0;
    size_t a = 0, b = 0;
    if (len) trim_space(msg, len, &a, &b);
    if (msg_off) *msg_off = a;
    if (msg_len) *msg_len = b - a;
    const size_t n = b - a;
    if (n == 0) return SJ_ERR_STAGE1;  // nothing to index: findStructuralIndices reports failure
    if (n > SJ_MAX_MESSAGE) return SJ_ERR_TOO_LARGE;
    SJ_CUDA_CHECK(cudaSetDevice(c->device));
    int rc = upload_message(c, msg + a, n);
    if (rc) return rc;
    Stage1Result r1;
    Stage2Result r2{};
    rc = front_half(c, c->msg.as<uint8_t>(), n, flags, &r1, &r2);
    if (rc) return rc;
    // the emitting half, the copies of tape and strings into the caller's buffers and the verdict: one synchronisation
    const HostOut host{tape, tape_cap, strings, strings_cap};
    rc = stage2_emit_any(c, nullptr, 0, nullptr, 0, ParseBases{0, 0, 0, nullptr}, &r2, &host);
    if (rc == SJ_OK || rc == SJ_ERR_CAPACITY) {
        *tape_len = r2.tape_len;
        *strings_len = r2.strings_len;
    }
    if (rc) return rc;
    return stage2_verdict(r2);
}

// parseMessage up to the point where tape and strings sit in the context's device buffers
// (c->last_*): trim, upload, stage 1, stage 2, verdict.  Shared by the entry points that do not
// copy into caller-provided host buffers (tape consumers, the NDJSON stream).
static int parse_into_ctx(sj_ctx* c, const uint8_t* msg, size_t len, uint32_t flags, size_t* msg_off, size_t* msg_len,
                          Stage2Result* r2) {
    size_t a = 0, b = 0;
    if (len) trim_space(msg, len, &a, &b);
    if (msg_off) *msg_off = a;
    if (msg_len) *msg_len = b - a;
    const size_t n = b - a;
    if (n == 0) return SJ_ERR_STAGE1;
    if (n > SJ_MAX_MESSAGE) return SJ_ERR_TOO_LARGE;
    SJ_CUDA_CHECK(cudaSetDevice(c->device));
    int rc = upload_message(c, msg + a, n);
    if (rc) return rc;
    Stage1Result r1;
    rc = front_half(c, c->msg.as<uint8_t>(), n, flags, &r1, r2);
    if (rc) return rc;
    rc = stage2_emit_any(c, nullptr, 0, nullptr, 0, ParseBases{0, 0, 0, nullptr}, r2);
    if (rc) return rc;
    return stage2_verdict(*r2);
}

// ---------------------------------------------------------------------------------
// unit-test hooks for the stage-2 leaf routines
// ---------------------------------------------------------------------------------
// one WARP per string: the thread-serial routines (string_measure / string_copy), the warp-cooperative ones
// (warp_string_measure / warp_string_copy) and the all-escapes-at-once one (warp_string_fast, resolved against the
// bound exactly as K2a does) all run and must agree; a disagreement is reported as src_len = ~0
__global__ void test_strings_kernel(const uint8_t* buf, const uint64_t* offs, size_t n, const uint64_t* max_size,
                                    uint8_t* ok, uint64_t* src_len, uint64_t* dst_len, uint8_t* dst, uint8_t* dst2,
                                    uint8_t* dst3) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (i >= n) return;  // warp-uniform
    const uint64_t o = offs[i], e = offs[i + 1];
    StrCursor s{buf + o + 1, e > o ? e - o - 1 : 0};
    uint64_t sl = 0, dl = 0, sl_w = 0, dl_w = 0, sl_f = 0, dl_f = 0;
    const uint64_t mx = max_size[i];
    const bool good = e > o && string_measure(s, mx, &sl, &dl);
    const bool good_w = e > o && warp_string_measure(s, mx, &sl_w, &dl_w);
    bool good_f = false;
    if (e > o && mx != 0) {
        const int r = warp_string_fast<false, 1>(s, mx, nullptr, &sl_f, &dl_f);
        good_f = r == 1;
        if (r == 2 || (r == 1 && sl_f >= mx)) good_f = warp_string_measure(s, mx, &sl_f, &dl_f);
    }
    bool same = good == good_w && good == good_f && (!good || (sl == sl_w && dl == dl_w && sl == sl_f && dl == dl_f));
    if (good && same) {
        if (lane == 0) string_copy(s, dst2 + o);
        warp_string_copy(s, dst3 + o);
        uint64_t sl_c = 0, dl_c = 0;
        const int r = warp_string_fast<true, (int)SJ_S2_FAST_MIN_BACKSLASHES>(s, ~0ull, dst + o, &sl_c, &dl_c);
        __syncwarp();
        bool eq = r == 1 && sl_c == sl && dl_c == dl;
        for (uint64_t k = lane; k < dl; k += 32) eq = eq && dst[o + k] == dst2[o + k] && dst3[o + k] == dst2[o + k];
        same = __all_sync(FULL, eq);
    }
    if (lane == 0) {
        ok[i] = good;
        src_len[i] = same ? sl : ~0ull;
        dst_len[i] = dl;
    }
}

__global__ void test_numbers_kernel(const uint8_t* buf, const uint64_t* offs, size_t n, uint64_t* tag, uint64_t* val, uint64_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t v = 0;
    tag[i] = parse_number(buf + offs[i], offs[i + 1] - offs[i], &v);
    val[i] = v;
    // the one-pass fast path (K2h tries it first) must agree wherever it decides at all
    uint64_t vf = 0;
    // (it may look at up to 24 bytes: give it the rest of the batch -- a number ends at its delimiter inside its own item)
    const uint64_t tf = parse_number_fast(buf + offs[i], total - offs[i], &vf);
    if (tf != PN_SLOW && (tf != tag[i] || vf != v)) tag[i] = 0xBAD0BAD0BAD0BAD0ull;
}

extern "C" int sj_test_parse_strings(sj_ctx* c, const uint8_t* buf, const uint64_t* offs, size_t n,
                                     const uint64_t* max_size, uint8_t* ok, uint64_t* src_len, uint64_t* dst_len,
                                     uint8_t* dst) {
    if (!c || n == 0) return SJ_ERR_ARGUMENT;
    SJ_CUDA_CHECK(cudaSetDevice(c->device));
    const size_t total = offs[n];
    size_t need = Carver::need({total + 64, (n + 1) * 8, n * 8, n, n * 8, n * 8, total + 64, total + 64, total + 64});
    int rc = c->test_in.reserve(need);
    if (rc) return rc;
    Carver k(c->test_in.p);
    uint8_t* d_buf = k.take<uint8_t>(total + 64);
    uint64_t* d_offs = k.take<uint64_t>(n + 1);
    uint64_t* d_max = k.take<uint64_t>(n);
    uint8_t* d_ok = k.take<uint8_t>(n);
    uint64_t* d_sl = k.take<uint64_t>(n);
    uint64_t* d_dl = k.take<uint64_t>(n);
    uint8_t* d_dst = k.take<uint8_t>(total + 64);
    uint8_t* d_dst2 = k.take<uint8_t>(total + 64);
    uint8_t* d_dst3 = k.take<uint8_t>(total + 64);
    SJ_CUDA_CHECK(cudaMemsetAsync(d_buf, 0, total + 64, c->stream));
    SJ_CUDA_CHECK(cudaMemcpyAsync(d_buf, buf, total, cudaMemcpyHostToDevice, c->stream));
    SJ_CUDA_CHECK(cudaMemcpyAsync(d_offs, offs, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    SJ_CUDA_CHECK(cudaMemcpyAsync(d_max, max_size, n * 8, cudaMemcpyHostToDevice, c->stream));
    SJ_CUDA_CHECK(cudaMemsetAsync(d_dst, 0, total + 64, c->stream));
    SJ_CUDA_CHECK(cudaMemsetAsync(d_dst2, 0, total + 64, c->stream));
    SJ_CUDA_CHECK(cudaMemsetAsync(d_dst3, 0, total + 64, c->stream));
    test_strings_kernel<<<(unsigned)((n + 1) / 2), 64, 0, c->stream>>>(d_buf, d_offs, n, d_max, d_ok, d_sl, d_dl, d_dst, d_dst2, d_dst3);
    c->launches++;
    SJ_CUDA_CHECK(cudaGetLastError());
    SJ_CUDA_CHECK(cudaMemcpyAsync(ok, d_ok, n, cudaMemcpyDeviceToHost, c->stream));
    SJ_CUDA_CHECK(cudaMemcpyAsync(src_len, d_sl, n * 8, cudaMemcpyDeviceToHost, c->stream));
    SJ_CUDA_CHECK(cudaMemcpyAsync(dst_len, d_dl, n * 8, cudaMemcpyDeviceToHost, c->stream));
    SJ_CUDA_CHECK(cudaMemcpyAsync(dst, d_dst, total, cudaMemcpyDeviceToHost, c->stream));
    SJ_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return SJ_OK;
}

extern "C" int sj_test_parse_numbers(sj_ctx* c, const uint8_t* buf, const uint64_t* offs, size_t n, uint64_t* tag,
                                     uint64_t* val) {
    if (!c || n == 0) return SJ_ERR_ARGUMENT;
    SJ_CUDA_CHECK(cudaSetDevice(c->device));
    const size_t total = offs[n];
    size_t need = Carver::need({total + 64, (n + 1) * 8, n * 8, n * 8});
    int rc = c->test_in.reserve(need);
    if (rc) return rc;
    Carver k(c->test_in.p);
    uint8_t* d_buf = k.take<uint8_t>(total + 64);
    uint64_t* d_offs = k.take<uint64_t>(n + 1);
    uint64_t* d_tag = k.take<uint64_t>(n);
    uint64_t* d_val = k.take<uint64_t>(n);
    SJ_CUDA_CHECK(cudaMemcpyAsync(d_buf, buf, total, cudaMemcpyHostToDevice, c->stream));
    SJ_CUDA_CHECK(cudaMemcpyAsync(d_offs, offs, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    test_numbers_kernel<<<(unsigned)((n + 63) / 64), 64, 0, c->stream>>>(d_buf, d_offs, n, d_tag, d_val, (uint64_t)total);
    c->launches++;
    SJ_CUDA_CHECK(cudaGetLastError());
    SJ_CUDA_CHECK(cudaMemcpyAsync(tag, d_tag, n * 8, cudaMemcpyDeviceToHost, c->stream));
    SJ_CUDA_CHECK(cudaMemcpyAsync(val, d_val, n * 8, cudaMemcpyDeviceToHost, c->stream));
    SJ_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return SJ_OK;
}
