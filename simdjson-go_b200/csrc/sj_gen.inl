// sj_gen.inl -- host side of K0 (gen.cuh).  Included by sj_api.cu.
extern "C" int sj_gen_ndjson_device(sj_ctx* c, const uint8_t* tmpl, size_t tmpl_len, uint64_t first_record, uint64_t n_records,
                                    uint8_t* d_out, size_t cap, size_t* out_len) {
    if (!c || !tmpl || !out_len || tmpl_len == 0 || n_records == 0) return SJ_ERR_ARGUMENT;
    *out_len = 0;
    // the template: records separated by '\n' (a trailing one is optional), each starting with {"Ticket":"<10 digits>"
    std::vector<uint32_t> off;
    std::vector<uint8_t> t(tmpl, tmpl + tmpl_len);
    if (t.back() != '\n') t.push_back('\n');
    for (size_t i = 0; i < t.size();) {
        const uint8_t* nl = reinterpret_cast<const uint8_t*>(memchr(t.data() + i, '\n', t.size() - i));
        const size_t e = (size_t)(nl - t.data());
        if (e - i < GEN_TICKET_OFF + GEN_TICKET_DIGITS + 2 || memcmp(t.data() + i, "{\"Ticket\":\"", GEN_TICKET_OFF) != 0) return SJ_ERR_ARGUMENT;
        for (uint32_t k = 0; k < GEN_TICKET_DIGITS; k++)
            if (t[i + GEN_TICKET_OFF + k] < '0' || t[i + GEN_TICKET_OFF + k] > '9') return SJ_ERR_ARGUMENT;
        off.push_back((uint32_t)i);
        i = e + 1;
    }
    const uint32_t L = (uint32_t)off.size();
    off.push_back((uint32_t)t.size());
    if (first_record % L != 0 || L > 10000) return SJ_ERR_ARGUMENT;
    const uint64_t T = t.size();
    const uint64_t full = n_records / L, rest = n_records % L;
    const uint64_t total = full * T + off[rest] - 1;  // without the last record's newline
    *out_len = (size_t)total;
    if (total > cap) return SJ_ERR_CAPACITY;
    if ((reinterpret_cast<uintptr_t>(d_out) & 15) != 0) return SJ_ERR_ARGUMENT;
    SJ_CUDA_CHECK(cudaSetDevice(c->device));
    const size_t need = Carver::need({t.size(), off.size() * 4});
    int rc = c->test_in.reserve(need);
    if (rc) return rc;
    Carver k(c->test_in.p);
    uint8_t* d_t = k.take<uint8_t>(t.size());
    uint32_t* d_off = k.take<uint32_t>(off.size());
    SJ_CUDA_CHECK(cudaMemcpyAsync(d_t, t.data(), t.size(), cudaMemcpyHostToDevice, c->stream));
    SJ_CUDA_CHECK(cudaMemcpyAsync(d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice, c->stream));
    GenParams p;
    p.tmpl = d_t;
    p.off = d_off;
    p.L = L;
    p.T = T;
    p.first_record = first_record;
    p.out_len = total;
    p.out = d_out;
    const uint64_t threads = (total + 15) / 16;
    gen_ndjson_kernel<<<(unsigned)((threads + 255) / 256), 256, (L + 1) * 4, c->stream>>>(p);
    c->launches++;
    SJ_CUDA_CHECK(cudaGetLastError());
    SJ_CUDA_CHECK(cudaStreamSynchronize(c->stream));  // (the staging vectors above go out of scope)
    return SJ_OK;
}
