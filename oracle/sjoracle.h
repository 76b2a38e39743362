/*
 * sjoracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C scalar restatement of minio/simdjson-go's parse hot path
 * (stage 1 structural-index discovery + flatten_bits, stage 2 tape build with
 * parse_string / parse_number).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may link or call this.
 * The product path (simdjson-go_b200/csrc) never does.
 *
 * Parity status: PINNED against the reference's own golden vectors
 * (tests/golden/G1..G19, extracted from the reference's Go tests by
 * tests/golden/make_golden.py).  The reference itself (Go + Plan-9 assembly)
 * cannot be built here: there is no Go toolchain in this image.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout, minio/simdjson-go @ 20f0d8f).
 */
#ifndef SJORACLE_H
#define SJORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SJO_FLAG_NDJSON 1u
#define SJO_FLAG_COPY_STRINGS 2u

#define SJO_OK 0
#define SJO_ERR_STAGE1 1 /* "Failed to find all structural indices for stage 1" */
#define SJO_ERR_STAGE2 2 /* "Bad parsing while executing stage 2" */
#define SJO_ERR_CAPACITY -1

#define SJO_INDEX_SIZE 1536               /* parsed_json.go:74 */
#define SJO_INDEX_SIZE_SAFETY (1536 - 128) /* parsed_json.go:75 */

/* ---- stage 1, one 64-byte block each (find_subroutines_amd64.go stubs) ---- */
uint64_t sjo_find_odd_backslash_sequences(const uint8_t *in64, uint64_t *prev_iter_ends_odd_backslash);
uint64_t sjo_find_quote_mask_and_bits(const uint8_t *in64, uint64_t odd_ends, uint64_t *prev_iter_inside_quote,
                                      uint64_t *quote_bits, uint64_t *error_mask);
void sjo_find_whitespace_and_structurals(const uint8_t *in64, uint64_t *whitespace, uint64_t *structurals);
uint64_t sjo_finalize_structurals(uint64_t structurals, uint64_t whitespace, uint64_t quote_mask, uint64_t quote_bits,
                                  uint64_t *prev_iter_ends_pseudo_pred);
uint64_t sjo_find_newline_delimiters(const uint8_t *in64, uint64_t quote_mask);
void sjo_flatten_bits_incremental(uint32_t *base, int *index, uint64_t mask, uint64_t *carried, uint64_t *position);
uint64_t sjo_find_structural_bits(const uint8_t *in64, uint64_t *prev_iter_ends_odd_backslash,
                                  uint64_t *prev_iter_inside_quote, uint64_t *error_mask, uint64_t structurals_in,
                                  uint64_t *prev_iter_ends_pseudo_pred);
uint64_t sjo_find_structural_bits_in_slice(const uint8_t *buf, uint64_t len, uint64_t *prev_iter_ends_odd_backslash,
                                           uint64_t *prev_iter_inside_quote, uint64_t *error_mask,
                                           uint64_t *prev_iter_ends_pseudo_pred, uint32_t *indexes, int *index,
                                           uint64_t *carried, uint64_t *position, uint64_t ndjson);

/* ---- stage 1 driver (stage1_find_marks_amd64.go:41) --------------------- */
/* Emits the concatenation of every index chunk the reference would hand to
 * stage 2 (uint32 deltas; sum(deltas)-1 = absolute offset).  Returns 1 when the
 * reference's findStructuralIndices() returns true.  *n may be written even on
 * failure.  deltas may be NULL to count only. */
int sjo_find_structural_indices(const uint8_t *msg, size_t len, int ndjson, uint32_t *deltas, size_t cap, size_t *n);

/* ---- stage 2 pieces ----------------------------------------------------- */
/* buf points AT the opening quote; avail = readable bytes from buf (bytes past
 * it read as 0, like the reference's zero padded copy). */
int sjo_parse_string_validate_only(const uint8_t *buf, size_t avail, uint64_t max_string_size, uint64_t *src_len,
                                   uint64_t *dst_len);
int sjo_parse_string(const uint8_t *buf, size_t avail, uint8_t *dst, uint64_t *dst_len);
/* returns the tape tag word (tag<<56 | flags) or 0 on failure */
uint64_t sjo_parse_number(const uint8_t *buf, size_t len, uint64_t *val);
int sjo_is_valid_true_atom(const uint8_t *buf, size_t len);
int sjo_is_valid_false_atom(const uint8_t *buf, size_t len);
int sjo_is_valid_null_atom(const uint8_t *buf, size_t len);

/* Go bytes.TrimSpace (simdjson_amd64.go:87, parse_json_amd64.go:55) */
void sjo_trim_space(const uint8_t *buf, size_t len, size_t *start, size_t *stop);

/* ---- whole parse (parse_json_amd64.go:52 parseMessage) ------------------- */
int sjo_parse(const uint8_t *msg, size_t len, uint32_t flags, uint64_t *tape, size_t tape_cap, size_t *tape_len,
              uint8_t *strings, size_t strings_cap, size_t *strings_len, size_t *msg_off, size_t *msg_len);

/* stage-1-only throughput helper for the CPU baseline: runs the reference's
 * chunk loop without keeping the indexes; returns the number of indexes. */
size_t sjo_stage1_count(const uint8_t *msg, size_t len, int ndjson, int *ok);

/* countWhere / countObjects over a finished tape (ndjson_test.go:421-474, Object.FindKey
 * parsed_object.go:97-140): returns the matches, *roots = number of root elements */
uint64_t sjo_count_where(const uint64_t *tape, size_t tape_len, const uint8_t *strings, const uint8_t *msg,
                         const uint8_t *key, size_t klen, const uint8_t *value, size_t vlen, uint64_t *roots);

const char *sjo_isa(void);
#ifdef __cplusplus
}
#endif
#endif
