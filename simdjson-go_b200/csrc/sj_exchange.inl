// sj_exchange.inl -- host side of exchange.cuh: buffers, CUDA IPC handles, the enqueue hook of the counting half.
// Included by sj_api.cu ahead of sj_parse.inl.

struct SjExchange {
    int rank = 0, world = 1;
    uint64_t gap = 1;
    uint64_t timeout_ns = sj::XCHG_DEFAULT_TIMEOUT_NS;
    uint64_t epoch = 0;            // epochs enqueued so far
    uint64_t enqueued_for = 0;     // epoch of the last enqueue (the counting half marks its call with epoch + 1)
    bool armed = false;            // inside sj_parse_nd_sharded_count
    bool connected = false;
    uint64_t* d_local = nullptr;   // XCHG_BUFFER_BYTES, zeroed (sequence 0 = nothing published)
    uint64_t* d_totals = nullptr;  // 4 words: where the scan's top kernel leaves this shard's totals
    uint64_t* d_out = nullptr;     // XCHG_OUT_WORDS
    uint64_t** d_peers = nullptr;  // [world] device array
    uint64_t* h_out = nullptr;     // pinned mirror of d_out
    void* opened[sj::XCHG_MAX_WORLD] = {};  // cudaIpcOpenMemHandle results to close
};

static SjExchange* xchg_of(sj_ctx* c) { return static_cast<SjExchange*>(c->xchg); }

static void exchange_release(sj_ctx* c) {
    SjExchange* x = xchg_of(c);
    if (!x) return;
    for (void* p : x->opened)
        if (p) cudaIpcCloseMemHandle(p);
    if (x->d_local) cudaFree(x->d_local);
    if (x->d_totals) cudaFree(x->d_totals);
    if (x->d_out) cudaFree(x->d_out);
    if (x->d_peers) cudaFree(x->d_peers);
    if (x->h_out) cudaFreeHost(x->h_out);
    delete x;
    c->xchg = nullptr;
}

extern "C" int sj_exchange_create(sj_ctx* c, int rank, int world, uint64_t gap_bytes, void* handle_out) {
    if (!c || world < 1 || world > sj::XCHG_MAX_WORLD || rank < 0 || rank >= world) return SJ_ERR_ARGUMENT;
    SJ_CUDA_CHECK(cudaSetDevice(c->device));
    exchange_release(c);
    SjExchange* x = new (std::nothrow) SjExchange();
    if (!x) return SJ_ERR_ARGUMENT;
    c->xchg = x;
    x->rank = rank;
    x->world = world;
    x->gap = gap_bytes;
    const int rc = [&]() -> int {
        SJ_CUDA_CHECK(cudaMalloc(&x->d_local, sj::XCHG_BUFFER_BYTES));
        SJ_CUDA_CHECK(cudaMemset(x->d_local, 0, sj::XCHG_BUFFER_BYTES));
        SJ_CUDA_CHECK(cudaMalloc(&x->d_totals, 4 * sizeof(uint64_t)));
        SJ_CUDA_CHECK(cudaMemset(x->d_totals, 0, 4 * sizeof(uint64_t)));
        SJ_CUDA_CHECK(cudaMalloc(&x->d_out, sj::XCHG_OUT_WORDS * sizeof(uint64_t)));
        SJ_CUDA_CHECK(cudaMemset(x->d_out, 0, sj::XCHG_OUT_WORDS * sizeof(uint64_t)));
        SJ_CUDA_CHECK(cudaMalloc(&x->d_peers, sj::XCHG_MAX_WORLD * sizeof(uint64_t*)));
        SJ_CUDA_CHECK(cudaHostAlloc(&x->h_out, sj::XCHG_OUT_WORDS * sizeof(uint64_t), cudaHostAllocDefault));
        memset(x->h_out, 0, sj::XCHG_OUT_WORDS * sizeof(uint64_t));
        SJ_CUDA_CHECK(cudaDeviceSynchronize());
        if (handle_out) {
            static_assert(sizeof(cudaIpcMemHandle_t) == SJ_EXCHANGE_HANDLE_BYTES, "handle size of the C ABI");
            cudaIpcMemHandle_t h;
            SJ_CUDA_CHECK(cudaIpcGetMemHandle(&h, x->d_local));
            memcpy(handle_out, &h, sizeof h);
        }
        return SJ_OK;
    }();
    if (rc) exchange_release(c);
    return rc;
}

extern "C" int sj_exchange_set_gap(sj_ctx* c, uint64_t gap_bytes) {
    SjExchange* x = c ? xchg_of(c) : nullptr;
    if (!x) return SJ_ERR_ARGUMENT;
    x->gap = gap_bytes;
    return SJ_OK;
}

extern "C" int sj_exchange_set_timeout_ms(sj_ctx* c, uint32_t ms) {
    SjExchange* x = c ? xchg_of(c) : nullptr;
    if (!x || ms == 0) return SJ_ERR_ARGUMENT;
    x->timeout_ns = (uint64_t)ms * 1000000ull;
    return SJ_OK;
}

extern "C" void* sj_exchange_local(sj_ctx* c) { return c && xchg_of(c) ? xchg_of(c)->d_local : nullptr; }

// `enable_peer`: the buffers were handed in as plain pointers -- those on another device of this process (one process driving
// several GPUs, a goroutine / thread per context) need peer access from this context's device; buffers opened through CUDA
// IPC have it already (cudaIpcMemLazyEnablePeerAccess)
static int exchange_set_table(sj_ctx* c, void* const* peer_buffers, bool enable_peer) {
    SjExchange* x = xchg_of(c);
    uint64_t* tab[sj::XCHG_MAX_WORLD] = {};
    for (int r = 0; r < x->world; r++) {
        tab[r] = r == x->rank ? x->d_local : static_cast<uint64_t*>(peer_buffers[r]);
        if (!tab[r]) return SJ_ERR_ARGUMENT;
        if (enable_peer && r != x->rank) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, tab[r]) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device != c->device) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    cudaGetLastError();
                    return -(1000 + (int)e);
                }
            }
            cudaGetLastError();  // (cudaErrorPeerAccessAlreadyEnabled is sticky-free but reported once)
        }
    }
    SJ_CUDA_CHECK(cudaMemcpy(x->d_peers, tab, sizeof tab, cudaMemcpyHostToDevice));
    x->connected = true;
    return SJ_OK;
}

extern "C" int sj_exchange_connect_ptrs(sj_ctx* c, void* const* peer_buffers) {
    SjExchange* x = c ? xchg_of(c) : nullptr;
    if (!x || !peer_buffers) return SJ_ERR_ARGUMENT;
    SJ_CUDA_CHECK(cudaSetDevice(c->device));
    return exchange_set_table(c, peer_buffers, true);
}

extern "C" int sj_exchange_connect(sj_ctx* c, const void* handles) {
    SjExchange* x = c ? xchg_of(c) : nullptr;
    if (!x || !handles) return SJ_ERR_ARGUMENT;
    SJ_CUDA_CHECK(cudaSetDevice(c->device));
    void* tab[sj::XCHG_MAX_WORLD] = {};
    for (int r = 0; r < x->world; r++) {
        if (r == x->rank) {
            tab[r] = x->d_local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const uint8_t*>(handles) + (size_t)r * sizeof h, sizeof h);
        if (!x->opened[r]) SJ_CUDA_CHECK(cudaIpcOpenMemHandle(&x->opened[r], h, cudaIpcMemLazyEnablePeerAccess));
        tab[r] = x->opened[r];
    }
    return exchange_set_table(c, tab, false);
}

extern "C" const uint64_t* sj_exchange_bases(sj_ctx* c) { return c && xchg_of(c) ? xchg_of(c)->d_out : nullptr; }

// the exchange's result as the counting half's read-back brought it: { bases x4, whole x4, status, epoch }
extern "C" int sj_exchange_result(sj_ctx* c, uint64_t* out10) {
    SjExchange* x = c ? xchg_of(c) : nullptr;
    if (!x || !out10) return SJ_ERR_ARGUMENT;
    memcpy(out10, x->h_out, sj::XCHG_OUT_WORDS * sizeof(uint64_t));
    return SJ_OK;
}

// Enqueued by the counting half right behind the scan's top kernel (which wrote `d_totals`), in front of its read-back;
// `failed`: this rank has no totals (the peers must not wait for them).
static int exchange_enqueue(sj_ctx* c, const uint64_t* d_totals, bool failed, const sj::Stage1Result* d_s1, const uint32_t* d_s2_error,
                            bool wait_only = false) {
    SjExchange* x = xchg_of(c);
    if (!x || !x->armed || !x->connected) return SJ_OK;
    if (!wait_only && x->enqueued_for == x->epoch + 1) return SJ_OK;  // this call has pushed already
    sj::XchgParams p;
    p.peers = x->d_peers;
    p.local = x->d_local;
    p.totals = d_totals ? d_totals : x->d_totals;
    p.out = x->d_out;
    p.rank = (uint32_t)x->rank;
    p.world = (uint32_t)x->world;
    p.epoch = x->epoch + 1;
    p.gap = x->gap;
    p.failed = failed ? 1u : 0u;
    p.slice_ns = sj::XCHG_SLICE_NS;
    p.wait_only = wait_only ? 1u : 0u;
    p.s1 = d_s1;
    p.s2_error = d_s2_error;
    sj::shard_exchange_kernel<<<1, 32, 0, c->stream>>>(p);
    c->launches += 1;
    SJ_CUDA_CHECK(cudaGetLastError());
    SJ_CUDA_CHECK(cudaMemcpyAsync(x->h_out, x->d_out, sj::XCHG_OUT_WORDS * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    x->enqueued_for = x->epoch + 1;
    return SJ_OK;
}

// the counting half is repeated (stage 1 overflowed its index buffer; the kernel of the first attempt skipped itself)
static void exchange_rearm(sj_ctx* c) {
    SjExchange* x = xchg_of(c);
    if (x) x->enqueued_for = 0;
}

static void exchange_begin(sj_ctx* c, uint64_t** d_totals) {
    SjExchange* x = xchg_of(c);
    if (!x || !x->connected) return;
    x->armed = true;
    x->enqueued_for = 0;
    if (!*d_totals) *d_totals = x->d_totals;
}

// rc = what the counting half returned.  Every armed call advances the epoch, and every epoch is published.
static int exchange_end(sj_ctx* c, int rc) {
    SjExchange* x = xchg_of(c);
    if (!x || !x->armed) return rc;
    if (x->enqueued_for != x->epoch + 1) {
        const int e = exchange_enqueue(c, nullptr, true, nullptr, nullptr);
        if (e == SJ_OK) cudaStreamSynchronize(c->stream);
        if (rc == SJ_OK) rc = e ? e : SJ_ERR_EXCHANGE;  // (not reachable: a successful counting half has enqueued)
    }
    // peers that were not there within the kernel's slice: wait-only passes until they are, or the time limit is over
    // (also for a call that failed by itself: its epoch ends when every peer has pushed, like everybody else's)
    const auto t0 = std::chrono::steady_clock::now();
    while (x->h_out[9] == x->epoch + 1 && x->h_out[8] == sj::XCHG_PENDING) {
        const uint64_t waited = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        if (waited > x->timeout_ns) {
            x->h_out[8] = sj::XCHG_TIMEOUT;
            break;
        }
        if (exchange_enqueue(c, nullptr, false, nullptr, nullptr, true) != SJ_OK || cudaStreamSynchronize(c->stream) != cudaSuccess) {
            x->h_out[8] = sj::XCHG_TIMEOUT;
            break;
        }
    }
    x->epoch += 1;
    x->armed = false;
    if (rc == SJ_OK) {
        if (x->h_out[9] != x->epoch || x->h_out[8] == sj::XCHG_TIMEOUT) return SJ_ERR_EXCHANGE;
        if (x->h_out[8] == sj::XCHG_PEER_FAILED) return SJ_ERR_PEER;
    }
    return rc;
}
