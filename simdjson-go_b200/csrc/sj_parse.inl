// placeholder until stage 2 lands
extern "C" int sj_parse(sj_ctx*, const uint8_t*, size_t, uint32_t, uint64_t*, size_t, size_t*, uint8_t*, size_t, size_t*, size_t*, size_t*) { return SJ_ERR_ARGUMENT; }
extern "C" int sj_parse_device(sj_ctx*, const uint8_t*, size_t, uint32_t, uint64_t*, size_t, size_t*, uint8_t*, size_t, size_t*) { return SJ_ERR_ARGUMENT; }
extern "C" int sj_test_parse_strings(sj_ctx*, const uint8_t*, const uint64_t*, size_t, const uint64_t*, uint8_t*, uint64_t*, uint64_t*, uint8_t*) { return SJ_ERR_ARGUMENT; }
extern "C" int sj_test_parse_numbers(sj_ctx*, const uint8_t*, const uint64_t*, size_t, uint64_t*, uint64_t*) { return SJ_ERR_ARGUMENT; }
