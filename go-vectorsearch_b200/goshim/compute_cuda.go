//go:build cuda
// +build cuda

// compute_cuda.go -- the `cuda` build-tag sibling of compute_gonum.go / compute_gorgonia.go.
//
// Drop into github.com/expki/go-vectorsearch/compute next to the existing backends, widen the default
// backend's tag (compute.go:1-2, cosine.go:1-2) to `!gonum && !gorgonia && !cuda`, and build with
//   CGO_ENABLED=1 go build -tags cuda
// quantization.go and types.go are untagged and stay as they are (the Quantize*/Dequantize* Go functions
// keep working; QuantizeMatrixFloat32Cuda below is the bulk device path).
//
// NOT COMPILED in the build container of this repository (no Go toolchain there); the C ABI it binds is
// exercised from C++ (host/selftest.cpp) and Python ctypes (tests/).
package compute

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../lib -lvscuda -Wl,-rpath,${SRCDIR}/../lib
#include <stdlib.h>
#include "vscuda.h"
*/
import "C"

import (
	"fmt"
	"os"
	"runtime"
	"strconv"
	"sync"
	"unsafe"

	"github.com/expki/go-vectorsearch/logger"
)

var initOnce sync.Once

func cudaInit() {
	initOnce.Do(func() {
		dev := 0
		if v, err := strconv.Atoi(os.Getenv("LOCAL_RANK")); err == nil {
			dev = v
		}
		if rc := C.vs_init(C.int(dev)); rc != C.VS_OK {
			// no CPU fallback: the cuda backend refuses to start without a B200
			panic("compute(cuda): " + C.GoString(C.vs_last_error()))
		}
	})
}

// check turns a C status into the reference's own failure mode: panic for the empty-input cases
// (compute.go:13,26,30), logger Fatalf for a dimension mismatch (cosine.go:19-21,77-79).
// check turns a C ABI status into the reference's error behaviour (panic on empty input, compute.go:13,26,30; Fatalf
// on a dimension mismatch, cosine.go:19-21,77-79).  vs_last_error() is thread-local in the library and a goroutine may
// be rescheduled onto another OS thread between two cgo calls, so every caller pins its goroutine to its OS thread
// (pinned(), below) for the span "C call ... check": the message read here belongs to the call that failed.
func check(rc C.int) {
	switch rc {
	case C.VS_OK:
		return
	case C.VS_EDIM:
		logger.Sugar().Fatalf("%s", C.GoString(C.vs_last_error()))
	default:
		panic(C.GoString(C.vs_last_error()))
	}
}

// pinned runs f with the goroutine locked to its OS thread (see check).
func pinned(f func()) {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	f()
}

// ctx is one CUDA stream + scratch arena. The method-form calls (search.go:214, upload.go:245) share a
// default one guarded by a mutex; closures from cosine_cuda.go own theirs.
type ctx struct{ h *C.vs_ctx }

var (
	defaultCtx   *ctx
	defaultCtxMu sync.Mutex
)

func newCtx() *ctx {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	cudaInit()
	c := &ctx{}
	check(C.vs_ctx_create(&c.h))
	return c
}

func (c *ctx) close() {
	if c.h != nil {
		C.vs_ctx_destroy(c.h)
		c.h = nil
	}
}

func withDefaultCtx(f func(c *ctx)) {
	defaultCtxMu.Lock()
	defer defaultCtxMu.Unlock()
	if defaultCtx == nil {
		defaultCtx = newCtx()
	}
	pinned(func() { f(defaultCtx) })
}

// pack copies [][]uint8 into one C buffer: cgo forbids passing Go pointers to Go pointers.
func pack(rows [][]uint8) (buf unsafe.Pointer, n int, rowBytes int) {
	n = len(rows)
	if n == 0 {
		panic("matrix rows are empty") // compute.go:25-27
	}
	rowBytes = len(rows[0])
	if rowBytes-8 <= 0 {
		panic("matrix columns are empty") // compute.go:29-31
	}
	buf = C.malloc(C.size_t(n * rowBytes))
	dst := unsafe.Slice((*uint8)(buf), n*rowBytes)
	for i, r := range rows {
		if len(r) != rowBytes {
			C.free(buf)
			panic(fmt.Sprintf("matrix row %d has %d bytes, row 0 has %d", i, len(r), rowBytes))
		}
		copy(dst[i*rowBytes:], r)
	}
	return buf, n, rowBytes
}

type vectorContainer struct {
	row []uint8 // the quantized query row, immutable
}

type matrixContainer struct {
	h    *C.vs_matrix
	rows int
	cols int
}

func NewVector(vectorQuantized []uint8) Vector {
	cols := len(vectorQuantized) - 8
	if cols <= 0 {
		panic("vector columns are empty") // compute.go:12-14
	}
	cudaInit()
	row := make([]uint8, len(vectorQuantized))
	copy(row, vectorQuantized)
	return &vectorContainer{row: row}
}

func NewMatrix(matrixQuantized [][]uint8) Matrix {
	buf, n, rowBytes := pack(matrixQuantized)
	defer C.free(buf)
	m := &matrixContainer{rows: n, cols: rowBytes - 8}
	withDefaultCtx(func(c *ctx) {
		check(C.vs_matrix_create(c.h, (*C.uint8_t)(buf), C.size_t(n), C.size_t(rowBytes), &m.h))
	})
	runtime.SetFinalizer(m, func(m *matrixContainer) { C.vs_matrix_release(m.h) })
	return m
}

// Clone: device matrices are immutable (nothing is normalized in place), so a clone is a reference bump.
func (v *vectorContainer) Clone() Vector { return &vectorContainer{row: v.row} }

func (m *matrixContainer) Clone() Matrix {
	C.vs_matrix_retain(m.h)
	c := &matrixContainer{h: m.h, rows: m.rows, cols: m.cols}
	runtime.SetFinalizer(c, func(c *matrixContainer) { C.vs_matrix_release(c.h) })
	return c
}

// QuantizeMatrixFloat32Cuda is the device form of QuantizeMatrixFloat32 (quantization.go:142-148) for bulk
// callers; byte-identical output.
func QuantizeMatrixFloat32Cuda(matrix [][]float32) [][]uint8 {
	n := len(matrix)
	if n == 0 {
		return [][]uint8{}
	}
	d := len(matrix[0])
	in := C.malloc(C.size_t(n * d * 4))
	out := C.malloc(C.size_t(n * (8 + d)))
	defer C.free(in)
	defer C.free(out)
	src := unsafe.Slice((*float32)(in), n*d)
	for i, r := range matrix {
		copy(src[i*d:], r)
	}
	withDefaultCtx(func(c *ctx) {
		check(C.vs_quantize_f32(c.h, (*C.float)(in), C.size_t(n), C.size_t(d), (*C.uint8_t)(out)))
	})
	flat := C.GoBytes(out, C.int(n*(8+d)))
	res := make([][]uint8, n)
	for i := range res {
		res[i] = flat[i*(8+d) : (i+1)*(8+d)]
	}
	return res
}
