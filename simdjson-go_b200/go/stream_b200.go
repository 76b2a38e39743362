//go:build cgo && b200

package simdjson

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/.. -lsimdjson_b200 -Wl,-rpath,${SRCDIR}/..
#include "simdjson_b200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"io"
	"unsafe"
)

// ParseNDStream over the library's own pipeline (sj_stream_*, csrc/sj_stream.inl): same signature,
// same channel protocol and same error values as the reference (simdjson_amd64.go:116-215).  The
// chunking at record boundaries (:157-174), the concurrent parses (:132) and the ordered delivery
// (:134-152) happen behind the C ABI; this goroutine only moves bytes in (Go memory is touched
// during sj_stream_write alone: the library stages it into pinned buffers) and results out.
//
// NOTE: like parse_json_b200.go this file cannot be compiled in the build image (no Go toolchain);
// simdjson_b200/stream.py:ParseNDStreamNative is the same loop in Python and is what the GPU
// tests run.
func ParseNDStream(r io.Reader, res chan<- Stream, reuse <-chan *ParsedJson) {
	if !SupportedCPU() {
		go func() {
			res <- Stream{Error: fmt.Errorf("Host CPU does not meet target specs")}
			close(res)
		}()
		return
	}
	const chunkBytes = 256 << 20 // sized to fill a GPU, not the reference's 10 MiB
	const inflight = 3
	go func() {
		defer close(res)
		var s *C.sj_stream
		if rc := C.sj_stream_create(-1, inflight, chunkBytes, C.SJ_FLAG_COPY_STRINGS, &s); rc != C.SJ_OK {
			res <- Stream{Error: errors.New(C.GoString(C.sj_error_string(rc)))}
			return
		}
		defer C.sj_stream_destroy(s)

		// deliver the oldest finished chunk; false once the stream has ended (error sent)
		deliver := func() (more bool, empty bool) {
			var out C.sj_stream_result
			switch rc := C.sj_stream_next(s, &out); rc {
			case C.SJ_OK:
				pj := &ParsedJson{}
				select {
				case v := <-reuse:
					pj = v
				default:
				}
				// copy out of the pinned slot (or wrap with unsafe.Slice and release later)
				pj.Message = append(pj.Message[:0], unsafe.Slice((*byte)(unsafe.Pointer(out.message)), int(out.message_len))...)
				pj.Tape = append(pj.Tape[:0], unsafe.Slice((*uint64)(unsafe.Pointer(out.tape)), int(out.tape_len))...)
				if pj.Strings == nil {
					pj.Strings = &TStrings{}
				}
				pj.Strings.B = append(pj.Strings.B[:0], unsafe.Slice((*byte)(unsafe.Pointer(out.strings)), int(out.strings_len))...)
				C.sj_stream_release(s, &out)
				res <- Stream{Value: pj}
				return true, false
			case C.SJ_STREAM_EMPTY:
				return true, true
			case C.SJ_STREAM_END:
				res <- Stream{Error: io.EOF}
				return false, false
			case C.SJ_ERR_STAGE1:
				res <- Stream{Error: fmt.Errorf("parsing input: %w", errors.New("Failed to find all structural indices for stage 1"))}
				return false, false
			case C.SJ_ERR_STAGE2:
				res <- Stream{Error: fmt.Errorf("parsing input: %w", errors.New("Bad parsing while executing stage 2"))}
				return false, false
			default:
				res <- Stream{Error: errors.New(C.GoString(C.sj_error_string(rc)))}
				return false, false
			}
		}

		buf := make([]byte, 16<<20)
		for {
			n, err := r.Read(buf)
			for off := 0; off < n; {
				var taken C.size_t
				rc := C.sj_stream_write(s, (*C.uint8_t)(unsafe.Pointer(&buf[off])), C.size_t(n-off), &taken)
				if rc != C.SJ_OK { // the stream has already failed: report it and stop
					deliver()
					return
				}
				off += int(taken)
				if taken == 0 { // every slot is busy: hand out the oldest chunk first
					if more, _ := deliver(); !more {
						return
					}
				}
			}
			if err == io.EOF {
				break
			}
			if err != nil {
				res <- Stream{Error: err}
				return
			}
		}
		for C.sj_stream_close_input(s) == C.SJ_STREAM_BUSY {
			if more, _ := deliver(); !more {
				return
			}
		}
		for {
			if more, _ := deliver(); !more {
				return
			}
		}
	}()
}
