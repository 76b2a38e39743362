// sj_stream.inl -- ParseNDStream (simdjson_amd64.go:116-215) inside the library.  Included by sj_api.cu.
//
// The reference reads 10 MiB, extends the chunk to the next '\n' (:157-174), parses at most
// (GOMAXPROCS+1)/2 chunks concurrently (:132) and delivers the results in input order
// (:134-152).  Here:
//   * the caller pushes bytes with sj_stream_write (ANY host memory -- Go slices are pageable);
//     they are copied once, by the calling thread, into the pinned input buffer of the slot that is
//     being filled, and cut at the last '\n' once `chunk_bytes` are there (sized to fill a GPU, not
//     10 MiB); the bytes behind that newline open the next chunk;
//   * every slot owns one sj_ctx (= CUDA stream + device scratch), one worker thread and pinned
//     output buffers: H2D, kernels and D2H of consecutive chunks overlap across slots;
//   * sj_stream_next hands the results out in input order, each one an independent
//     {Message, Tape, Strings} triple exactly like the reference's Stream values; the first failing
//     chunk ends the stream with its error ("parsing input: ...", :196).
// The calls never block on a full pipeline: sj_stream_write reports how much it took, and
// sj_stream_next blocks only while a chunk that was already handed to a worker is still running,
// so a single thread can drive the stream (write until nothing is taken, then take a result).
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace {

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    // grow-only; keeps the first `keep` bytes
    int reserve(size_t bytes, size_t keep = 0) {
        if (bytes <= cap) return 0;
        const size_t want = bytes + (bytes >> 2) + 4096;
        void* q = nullptr;
        cudaError_t e = cudaHostAlloc(&q, want, cudaHostAllocDefault);
        if (e != cudaSuccess) return -(1000 + (int)e);
        if (p) {
            if (keep) memcpy(q, p, keep);
            cudaFreeHost(p);
        }
        p = q;
        cap = want;
        return 0;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

enum SlotState { SLOT_FREE, SLOT_FILLING, SLOT_QUEUED, SLOT_RUNNING, SLOT_DONE, SLOT_HELD };

struct StreamSlot {
    sj_ctx* ctx = nullptr;
    PinnedBuf in, tape, strings;
    size_t in_len = 0, tape_len = 0, strings_len = 0, msg_off = 0, msg_len = 0;
    int rc = 0;
    uint64_t seq = 0;
    SlotState state = SLOT_FREE;
    std::thread worker;
};

}  // namespace

struct sj_stream {
    int device = 0;
    uint32_t flags = 0;
    size_t chunk_bytes = 0;
    std::vector<StreamSlot> slots;
    std::mutex mu;
    std::condition_variable cv;
    int filling = -1;             // slot the writer is filling (owned by the writer, no lock needed for its bytes)
    uint64_t next_seq = 0;        // sequence number of the next chunk handed to a worker
    uint64_t deliver_seq = 0;     // sequence number of the next result to hand out
    std::vector<uint8_t> carry;   // bytes behind the last newline of the previous chunk
    bool closed = false, stop = false;
    int error = 0;                // first failing chunk ends the stream
};

namespace {

// the one copy of the stream's bytes (caller memory -> pinned slot buffer): a single thread moves ~10 GB/s, less than the
// PCIe link takes, so large pieces are split over a few helper threads
void copy_in(uint8_t* dst, const uint8_t* src, size_t n) {
    const size_t PIECE = 8u << 20;
    if (n < 2 * PIECE) {
        memcpy(dst, src, n);
        return;
    }
    unsigned helpers = std::thread::hardware_concurrency();
    helpers = helpers < 2 ? 1 : (helpers > 6 ? 6 : helpers);
    size_t per = (n / helpers + 63) & ~(size_t)63;
    std::vector<std::thread> th;
    size_t done = 0;
    for (unsigned i = 0; i + 1 < helpers && done + per < n; i++) {
        th.emplace_back([=] { memcpy(dst + done, src + done, per); });
        done += per;
    }
    memcpy(dst + done, src + done, n - done);
    for (auto& t : th) t.join();
}

bool all_space(const uint8_t* p, size_t n) {
    for (size_t i = 0; i < n; i++)
        if (!ascii_space(p[i])) return false;
    return true;
}

void stream_worker(sj_stream* s, int k) {
    StreamSlot& sl = s->slots[k];
    cudaSetDevice(s->device);
    std::unique_lock<std::mutex> lk(s->mu);
    for (;;) {
        s->cv.wait(lk, [&] { return s->stop || sl.state == SLOT_QUEUED; });
        if (s->stop) return;
        sl.state = SLOT_RUNNING;
        lk.unlock();
        // parseMessage: trim, upload, stage 1 + the counting half of stage 2 (one read-back: the totals size the slot's
        // pinned output buffers), then the emitting half with tape and strings copied straight into them
        Stage2Result r2;
        memset(&r2, 0, sizeof r2);
        sj_ctx* c = sl.ctx;
        const uint8_t* msg = reinterpret_cast<const uint8_t*>(sl.in.p);
        size_t a = 0, b = 0;
        if (sl.in_len) trim_space(msg, sl.in_len, &a, &b);
        sl.msg_off = a;
        sl.msg_len = b - a;
        int rc = SJ_OK;
        if (b == a)
            rc = SJ_ERR_STAGE1;
        else if (b - a > SJ_MAX_MESSAGE)
            rc = SJ_ERR_TOO_LARGE;
        if (rc == SJ_OK) rc = upload_message(c, msg + a, b - a);
        Stage1Result r1;
        if (rc == SJ_OK) rc = front_half(c, c->msg.as<uint8_t>(), b - a, s->flags, &r1, &r2);
        if (rc == SJ_OK) rc = sl.tape.reserve((size_t)r2.tape_len * 8 + 64);
        if (rc == SJ_OK) rc = sl.strings.reserve((size_t)r2.strings_len + 64);
        if (rc == SJ_OK) {
            const HostOut host{reinterpret_cast<uint64_t*>(sl.tape.p), (size_t)r2.tape_len, reinterpret_cast<uint8_t*>(sl.strings.p),
                               (size_t)r2.strings_len};
            rc = stage2_emit_any(c, nullptr, 0, nullptr, 0, ParseBases{0, 0, 0, nullptr}, &r2, &host);
            if (rc == SJ_OK) rc = stage2_verdict(r2);
            sl.tape_len = (size_t)r2.tape_len;
            sl.strings_len = (size_t)r2.strings_len;
        }
        lk.lock();
        sl.rc = rc;
        sl.state = SLOT_DONE;
        s->cv.notify_all();
    }
}

// hand the slot being filled to its worker (caller holds the lock)
void stream_submit(sj_stream* s) {
    StreamSlot& sl = s->slots[s->filling];
    sl.seq = s->next_seq++;
    sl.state = SLOT_QUEUED;
    s->filling = -1;
    s->cv.notify_all();
}

}  // namespace

extern "C" int sj_stream_create(int device, int inflight, size_t chunk_bytes, uint32_t flags, sj_stream** out) {
    if (!out) return SJ_ERR_ARGUMENT;
    *out = nullptr;
    if (inflight < 1 || inflight > 64 || chunk_bytes == 0 || chunk_bytes > SJ_MAX_MESSAGE / 2) return SJ_ERR_ARGUMENT;
    if (sj_device_count() == 0) return SJ_ERR_NO_DEVICE;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return SJ_ERR_NO_DEVICE;
    sj_stream* s = new (std::nothrow) sj_stream();
    if (!s) return SJ_ERR_ARGUMENT;
    s->device = device;
    s->flags = flags | SJ_FLAG_NDJSON;
    s->chunk_bytes = chunk_bytes;
    s->slots.resize((size_t)inflight);
    int rc = SJ_OK;
    for (int k = 0; k < inflight && rc == SJ_OK; k++) {
        rc = sj_ctx_create(device, &s->slots[k].ctx);
        if (rc == SJ_OK) rc = s->slots[k].in.reserve(chunk_bytes + (64 << 10));
    }
    if (rc != SJ_OK) {
        for (auto& sl : s->slots) {
            sl.in.release();
            sj_ctx_destroy(sl.ctx);
        }
        delete s;
        return rc;
    }
    for (int k = 0; k < inflight; k++) s->slots[k].worker = std::thread(stream_worker, s, k);
    *out = s;
    return SJ_OK;
}

extern "C" void sj_stream_destroy(sj_stream* s) {
    if (!s) return;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->stop = true;
    }
    s->cv.notify_all();
    for (auto& sl : s->slots)
        if (sl.worker.joinable()) sl.worker.join();
    cudaSetDevice(s->device);
    for (auto& sl : s->slots) {
        sl.in.release();
        sl.tape.release();
        sl.strings.release();
        sj_ctx_destroy(sl.ctx);
    }
    delete s;
}

extern "C" int sj_stream_write(sj_stream* s, const uint8_t* data, size_t len, size_t* taken) {
    if (!s || !taken || (len && !data)) return SJ_ERR_ARGUMENT;
    *taken = 0;
    std::unique_lock<std::mutex> lk(s->mu);
    if (s->error) return s->error;
    if (s->closed) return SJ_ERR_ARGUMENT;
    while (len > 0) {
        if (s->filling < 0) {
            int k = -1;
            for (size_t i = 0; i < s->slots.size(); i++)
                if (s->slots[i].state == SLOT_FREE) {
                    k = (int)i;
                    break;
                }
            if (k < 0) break;  // pipeline full: the caller takes a result first
            StreamSlot& sl = s->slots[k];
            sl.state = SLOT_FILLING;
            sl.in_len = 0;
            s->filling = k;
            if (!s->carry.empty()) {  // the bytes behind the previous chunk's last newline come first
                int rc = sl.in.reserve(s->carry.size() + s->chunk_bytes);
                if (rc) return rc;
                memcpy(sl.in.p, s->carry.data(), s->carry.size());
                sl.in_len = s->carry.size();
                s->carry.clear();
            }
        }
        StreamSlot& sl = s->slots[s->filling];
        lk.unlock();  // the slot being filled belongs to the writer alone
        uint8_t* buf = reinterpret_cast<uint8_t*>(sl.in.p);
        size_t room = sl.in_len < s->chunk_bytes ? s->chunk_bytes - sl.in_len : 0;
        bool submit = false;
        int rc = SJ_OK;
        if (room == 0) {
            // a chunk's worth of bytes and still no newline to cut at: one record larger than the chunk, keep growing
            rc = sl.in.reserve(sl.in_len + s->chunk_bytes, sl.in_len);
            buf = reinterpret_cast<uint8_t*>(sl.in.p);
            room = s->chunk_bytes;
            if (sl.in_len + room > SJ_MAX_MESSAGE) rc = SJ_ERR_TOO_LARGE;
        }
        if (rc == SJ_OK) {
            const size_t n = len < room ? len : room;
            copy_in(buf + sl.in_len, data, n);
            sl.in_len += n;
            data += n;
            len -= n;
            *taken += n;
            if (sl.in_len >= s->chunk_bytes) {
                const void* nl = memrchr(buf, '\n', sl.in_len);  // simdjson_amd64.go:165-174: cut at a record boundary
                if (nl) {
                    const size_t cut = (size_t)(reinterpret_cast<const uint8_t*>(nl) - buf) + 1;
                    if (all_space(buf, cut)) {
                        // only blank lines in front of the cut (e.g. "\n{...a record larger than the chunk...}"): a
                        // whitespace-only chunk would fail stage 1, the reference skips blank lines -- drop them and keep filling
                        memmove(buf, buf + cut, sl.in_len - cut);
                        sl.in_len -= cut;
                    } else {
                        s->carry.assign(buf + cut, buf + sl.in_len);
                        sl.in_len = cut;
                        submit = true;
                    }
                }
            }
        }
        lk.lock();
        if (rc) return rc;
        if (submit) stream_submit(s);
    }
    return SJ_OK;
}

extern "C" int sj_stream_close_input(sj_stream* s) {
    if (!s) return SJ_ERR_ARGUMENT;
    std::unique_lock<std::mutex> lk(s->mu);
    if (s->closed) return SJ_OK;
    if (s->filling < 0 && !s->carry.empty() && !all_space(s->carry.data(), s->carry.size())) {
        // the tail sits in `carry` and no slot is being filled: it needs a free slot
        int k = -1;
        for (size_t i = 0; i < s->slots.size(); i++)
            if (s->slots[i].state == SLOT_FREE) {
                k = (int)i;
                break;
            }
        if (k < 0) return SJ_STREAM_BUSY;  // every slot is in use: take a result, then close again
        StreamSlot& sl = s->slots[k];
        int rc = sl.in.reserve(s->carry.size() + 64);
        if (rc) return rc;
        memcpy(sl.in.p, s->carry.data(), s->carry.size());
        sl.in_len = s->carry.size();
        s->carry.clear();
        sl.state = SLOT_FILLING;
        s->filling = k;
    }
    if (s->filling >= 0) {
        StreamSlot& sl = s->slots[s->filling];
        if (all_space(reinterpret_cast<const uint8_t*>(sl.in.p), sl.in_len)) {  // nothing but blanks behind the last record
            sl.state = SLOT_FREE;
            s->filling = -1;
        } else {
            stream_submit(s);
        }
    }
    s->closed = true;
    s->cv.notify_all();
    return SJ_OK;
}

extern "C" int sj_stream_next(sj_stream* s, sj_stream_result* res) {
    if (!s || !res) return SJ_ERR_ARGUMENT;
    memset(res, 0, sizeof *res);
    std::unique_lock<std::mutex> lk(s->mu);
    if (s->error) return s->error;
    if (s->deliver_seq == s->next_seq) return s->closed && s->filling < 0 ? SJ_STREAM_END : SJ_STREAM_EMPTY;
    StreamSlot* sl = nullptr;
    for (auto& x : s->slots)
        if (x.seq == s->deliver_seq && (x.state == SLOT_QUEUED || x.state == SLOT_RUNNING || x.state == SLOT_DONE)) sl = &x;
    if (!sl) return SJ_ERR_ARGUMENT;  // cannot happen: every sequence number below next_seq lives in a slot
    s->cv.wait(lk, [&] { return sl->state == SLOT_DONE; });
    s->deliver_seq++;
    if (sl->rc != SJ_OK) {
        s->error = sl->rc;  // Stream{Error: "parsing input: ..."} and nothing after it (simdjson_amd64.go:196)
        sl->state = SLOT_FREE;
        return sl->rc;
    }
    sl->state = SLOT_HELD;
    res->message = reinterpret_cast<const uint8_t*>(sl->in.p) + sl->msg_off;
    res->message_len = sl->msg_len;
    res->tape = reinterpret_cast<const uint64_t*>(sl->tape.p);
    res->tape_len = sl->tape_len;
    res->strings = reinterpret_cast<const uint8_t*>(sl->strings.p);
    res->strings_len = sl->strings_len;
    res->seq = sl->seq;
    res->slot = sl;
    return SJ_OK;
}

extern "C" int sj_stream_release(sj_stream* s, const sj_stream_result* res) {
    if (!s || !res || !res->slot) return SJ_ERR_ARGUMENT;
    std::lock_guard<std::mutex> lk(s->mu);
    StreamSlot* sl = reinterpret_cast<StreamSlot*>(res->slot);
    if (sl < s->slots.data() || sl >= s->slots.data() + s->slots.size() || sl->state != SLOT_HELD) return SJ_ERR_ARGUMENT;
    sl->state = SLOT_FREE;
    return SJ_OK;
}
