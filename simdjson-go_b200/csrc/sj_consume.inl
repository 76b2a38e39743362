// sj_consume.inl -- C-ABI entry points of the device-side tape consumers (consume.cuh).
// Included by sj_api.cu.
//
// countWhere / countObjects of the reference's NDJSON tests and benchmarks (ndjson_test.go:421-474,
// parse_json_amd64_test.go:115-157) evaluated where the tape is: in HBM.

namespace {

// needles -> device, counters zeroed; layout of tc_small: [counters 2 x u64][key][value]
int tc_stage_needles(sj_ctx* c, const uint8_t* key, size_t klen, const uint8_t* value, size_t vlen, CountParams* p) {
    if (klen > 0x3fffffffu || vlen > 0x3fffffffu || (klen && !key) || (vlen && !value)) return SJ_ERR_ARGUMENT;
    const size_t koff = 64, voff = koff + align_up(klen + 1, 16);
    int rc = c->tc_small.reserve(voff + vlen + 16);
    if (rc) return rc;
    uint8_t* base = c->tc_small.as<uint8_t>();
    SJ_CUDA_CHECK(cudaMemsetAsync(base, 0, 64, c->stream));
    if (klen) SJ_CUDA_CHECK(cudaMemcpyAsync(base + koff, key, klen, cudaMemcpyHostToDevice, c->stream));
    if (vlen) SJ_CUDA_CHECK(cudaMemcpyAsync(base + voff, value, vlen, cudaMemcpyHostToDevice, c->stream));
    p->key = base + koff;
    p->key_len = (uint32_t)klen;
    p->value = base + voff;
    p->value_len = (uint32_t)vlen;
    p->counters = reinterpret_cast<unsigned long long*>(base);
    return SJ_OK;
}

int tc_run_count(sj_ctx* c, const CountParams& p, uint64_t* matches) {
    if (p.n_roots) {
        const unsigned blocks = (unsigned)((p.n_roots + 255) / 256);
        tc_count_where_kernel<<<blocks, 256, 0, c->stream>>>(p);
        c->launches++;
        SJ_CUDA_CHECK(cudaGetLastError());
    }
    unsigned long long* h = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(c->host_result) + 192);
    SJ_CUDA_CHECK(cudaMemcpyAsync(h, p.counters, 16, cudaMemcpyDeviceToHost, c->stream));
    SJ_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *matches = h[1];
    return SJ_OK;
}

}  // namespace

extern "C" int sj_count_where_device(sj_ctx* c, const uint8_t* d_msg, const uint64_t* d_tape, size_t tape_len,
                                     const uint8_t* d_strings, const uint8_t* key, size_t klen, const uint8_t* value,
                                     size_t vlen, uint64_t* roots, uint64_t* matches) {
    if (!c || !d_tape || !roots || !matches) return SJ_ERR_ARGUMENT;
    *roots = 0;
    *matches = 0;
    if (tape_len == 0) return SJ_OK;
    if (tape_len > 0xffffffffull) return SJ_ERR_TOO_LARGE;  // root slots are uint32 (a tape of one parse call always fits)
    SJ_CUDA_CHECK(cudaSetDevice(c->device));
    CountParams p;
    memset(&p, 0, sizeof p);
    int rc = tc_stage_needles(c, key, klen, value, vlen, &p);
    if (rc) return rc;
    p.tape = d_tape;
    p.tape_len = tape_len;
    p.strings = d_strings;
    p.msg = d_msg;
    if (c->last_tape == d_tape && c->last_tape_len == tape_len && c->last_rootpos) {
        // the tape this context has just built: stage 2's own root list is still there
        p.roots = c->last_rootpos;
        p.roots_skip0 = 1;
        p.n_roots = c->last_records + 1;
    } else {
        const uint64_t cap = tape_len / 4 + 1;  // a record is at least  r { } r
        rc = c->tc_roots.reserve(cap * sizeof(uint32_t));
        if (rc) return rc;
        const unsigned blocks = (unsigned)((tape_len + 255) / 256);
        tc_find_roots_kernel<<<blocks, 256, 0, c->stream>>>(d_tape, tape_len, c->tc_roots.as<uint32_t>(), cap, p.counters);
        c->launches++;
        SJ_CUDA_CHECK(cudaGetLastError());
        unsigned long long* h = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(c->host_result) + 192);
        SJ_CUDA_CHECK(cudaMemcpyAsync(h, p.counters, 8, cudaMemcpyDeviceToHost, c->stream));
        SJ_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (h[0] > cap) return SJ_ERR_ARGUMENT;  // not a tape
        p.roots = c->tc_roots.as<uint32_t>();
        p.roots_skip0 = 0;
        p.n_roots = h[0];
    }
    *roots = p.n_roots;
    return tc_run_count(c, p, matches);
}

extern "C" int sj_parse_count_where(sj_ctx* c, const uint8_t* msg, size_t len, uint32_t flags, const uint8_t* key,
                                    size_t klen, const uint8_t* value, size_t vlen, uint64_t* roots, uint64_t* matches) {
    if (!c || !roots || !matches) return SJ_ERR_ARGUMENT;
    *roots = 0;
    *matches = 0;
    Stage2Result r2;
    int rc = parse_into_ctx(c, msg, len, flags, nullptr, nullptr, &r2);
    if (rc) return rc;
    // tape, strings and message stay in HBM; only the two counts travel back
    return sj_count_where_device(c, c->last_msg, c->last_tape, (size_t)c->last_tape_len, c->last_strings, key, klen, value,
                                 vlen, roots, matches);
}
