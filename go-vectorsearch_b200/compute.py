"""Host-side mirror of the reference's `compute` package over libvscuda (cgo stand-in: ctypes).

Same names, argument meaning and error behaviour as the Go package so parity tests read like
tests of the reference would:

  compute/types.go:3-11        Vector, Matrix (Clone, MatrixCosineSimilarity)
  compute/compute.go:10-44     NewVector, NewMatrix  (panic on empty input)
  compute/cosine.go:60-66      VectorMatrixCosineSimilarity() -> (calculate, done)
  compute/cosine.go:129-135    MatrixCosineSimilarity()       -> (calculate, done)
  compute/quantization.go      Quantize*/Dequantize* (Vector and Matrix, Float32 and Float64)

Go `panic(...)` surfaces as ComputePanic, `logger.Sugar().Fatalf(...)` (process exit) as ComputeFatal.
Inputs are never mutated (the reference normalizes in place, which is why its callers Clone();
here Clone() is a reference-count bump on an immutable device matrix).
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import VS_EDIM, VS_EEMPTY, VS_OK


class ComputePanic(Exception):
    """Go panic() in the reference (compute.go:13,26,30)."""


class ComputeFatal(Exception):
    """logger.Sugar().Fatalf in the reference (cosine.go:19-21,77-79): the Go process would exit."""


class ComputeError(RuntimeError):
    """Any other backend failure (CUDA error, out of memory, unsupported size)."""


VS_EFULL = -8


class IndexFull(ComputeError):
    """vs_index_append: a list (or the store) has no room for its new rows; the index is unchanged."""


def _check(rc):
    if rc == VS_OK:
        return
    msg = _lib.last_error()
    if rc == VS_EFULL:
        raise IndexFull(msg)
    if rc == VS_EEMPTY:
        raise ComputePanic(msg)
    if rc == VS_EDIM:
        raise ComputeFatal(msg)
    raise ComputeError(f"libvscuda error {rc}: {msg}")


def _p(a):
    return C.c_void_p(a.ctypes.data)


def debug_set_certify_scale(scale):
    """Test hook (vs_debug_set_certify_scale): inflate the certification half-widths so every score takes the
    literal reference-arithmetic path; results must not change. 1.0 restores production behaviour."""
    _check(_lib.init().vs_debug_set_certify_scale(C.c_float(scale)))


def debug_set_argmax_gemm_min(min_centroids):
    """Test hook: centroid count from which nearest-centroid assignment runs on the tensor cores."""
    _check(_lib.init().vs_debug_set_argmax_gemm_min(int(min_centroids)))


def release_cached_memory():
    """Hand the freed matrix / store memory the library keeps for reuse back to the driver (vs_release_cached_memory)."""
    _check(_lib.init().vs_release_cached_memory())


def debug_set_fused(on):
    """Test hook: False sends single-query searches through the two-launch streaming path instead of the fused kernel."""
    _check(_lib.init().vs_debug_set_fused(1 if on else 0))


def debug_set_list_major(on):
    """Test hook: False sends the list stage of query batches through the query-major scan instead of the list-major one."""
    _check(_lib.init().vs_debug_set_list_major(1 if on else 0))


def debug_set_lm_dense_min(queries_per_list):
    """Test hook: (query, list) pairs per list from which the list-major scan runs on the tensor cores (0 = never; default 4)."""
    _check(_lib.init().vs_debug_set_lm_dense_min(int(queries_per_list)))


class Context:
    """One CUDA stream + scratch arena (vs_ctx). One per closure / goroutine (dnc/dnc.go:349)."""

    def __init__(self, cuda_stream=None):
        """cuda_stream: optional raw cudaStream_t (int) to enqueue on, e.g. the stream NCCL uses."""
        L = _lib.init()
        h = C.c_void_p()
        if cuda_stream is None:
            _check(L.vs_ctx_create(C.byref(h)))
        else:
            _check(L.vs_ctx_create_on_stream(C.c_void_p(int(cuda_stream)), C.byref(h)))
        self._h = h
        self._L = L

    def profile_enable(self, on=True):
        _check(self._L.vs_ctx_profile_enable(self.handle, 1 if on else 0))

    def trace_enable(self, on=True):
        _check(self._L.vs_ctx_trace_enable(self.handle, 1 if on else 0))

    def trace_read(self, stage):
        """[blocks, 16] uint64 %globaltimer stamps (ns) of the latest launch of search stage 1 or 2."""
        out = np.zeros((4096, 16), np.uint64)
        nb = C.c_size_t()
        _check(self._L.vs_ctx_trace_read(self.handle, int(stage), _p(out), 4096, C.byref(nb)))
        return out[:nb.value]

    def profile_read(self):
        """(summed list-scan kernel ms, launches) since the last read."""
        ms, cnt = C.c_double(), C.c_uint64()
        _check(self._L.vs_ctx_profile_read(self.handle, C.byref(ms), C.byref(cnt)))
        return ms.value, int(cnt.value)

    @property
    def handle(self):
        if self._h is None:
            raise ComputeError("context already released (done() was called)")
        return self._h

    def close(self):
        if self._h is not None:
            self._L.vs_ctx_destroy(self._h)
            self._h = None

    def sync(self):
        _check(self._L.vs_ctx_sync(self.handle))

    def launch_count(self):
        return int(self._L.vs_ctx_launch_count(self.handle))

    def slowpath_count(self):
        return int(self._L.vs_ctx_slowpath_count(self.handle))

    def timer_start(self):
        _check(self._L.vs_ctx_timer_start(self.handle))

    def timer_stop(self):
        ms = C.c_float()
        _check(self._L.vs_ctx_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context():
    """Method-form calls (search.go:214, upload.go:245) run on a default stream/context."""
    global _default_ctx
    if _default_ctx is None or _default_ctx._h is None:
        _default_ctx = Context()
    return _default_ctx


def _rows_array(matrixQuantized):
    """[][]uint8 -> contiguous (n, 8+d) uint8 (the cgo shim packs Go rows the same way)."""
    if isinstance(matrixQuantized, np.ndarray):
        if matrixQuantized.ndim != 2:
            raise ComputePanic("matrix rows are empty" if matrixQuantized.size == 0 else "matrix must be 2-D")
        return np.ascontiguousarray(matrixQuantized, dtype=np.uint8)
    rows = [np.frombuffer(bytes(r), dtype=np.uint8) if not isinstance(r, np.ndarray) else r.astype(np.uint8, copy=False)
            for r in matrixQuantized]
    if len(rows) == 0:
        raise ComputePanic("matrix rows are empty")  # compute.go:25-27
    width = rows[0].shape[0]
    for r in rows:
        if r.shape[0] != width:
            # the reference takes cols from row 0 and silently mis-copies ragged rows (compute.go:28,37-39);
            # the device backend refuses instead
            raise ComputePanic("matrix rows have different lengths")
    return np.ascontiguousarray(np.stack(rows), dtype=np.uint8)


class Matrix:
    """compute.Matrix (types.go:8-11) backed by a device-resident vs_matrix."""

    def __init__(self, handle, L):
        self._h = handle
        self._L = L

    @property
    def handle(self):
        return self._h

    @property
    def rows(self):
        return int(self._L.vs_matrix_rows(self._h))

    @property
    def cols(self):
        return int(self._L.vs_matrix_cols(self._h))

    def Clone(self):
        self._L.vs_matrix_retain(self._h)
        return Matrix(self._h, self._L)

    def MatrixCosineSimilarity(self, matrix, ctx=None, want_sims=True):
        """Receiver = centroids, argument = data (cosine.go:70-125) -> (sims[n] float32, argmax[n] int)."""
        ctx = ctx or default_context()
        n = matrix.rows
        idx = np.empty(n, np.int64)
        sims = np.empty(n, np.float32) if want_sims else None
        _check(self._L.vs_argmax_MxN(ctx.handle, self._h, matrix._h, _p(sims) if want_sims else None, _p(idx)))
        return sims, idx

    def FillFloat32Dev(self, first, d_ptr, count, ctx=None):
        """Quantize `count` device float32 rows (raw pointer) into rows [first, first+count)."""
        ctx = ctx or default_context()
        _check(self._L.vs_matrix_fill_f32_dev(ctx.handle, self._h, int(first), C.c_void_p(int(d_ptr)), int(count)))

    def LoadRows(self, first, rows, ctx=None):
        """Overwrite rows from host row776 rows (async on the ctx stream; keep `rows` alive until sync)."""
        ctx = ctx or default_context()
        _check(self._L.vs_matrix_load_rows(ctx.handle, self._h, int(first), _p(rows), rows.shape[0]))

    def ArgmaxDev(self, data, d_idx_ptr, ctx=None):
        """Receiver = centroids; int32 nearest-centroid index per data row written to a device pointer."""
        ctx = ctx or default_context()
        _check(self._L.vs_argmax_MxN_dev(ctx.handle, self._h, data._h, C.c_void_p(int(d_idx_ptr))))

    def SaveSpool(self, path, first=0, count=None, append=False, ctx=None):
        """Write rows as a D&C row spool (createDataset.WriteRow, dnc/dataset.go:53-56, for many rows at once)."""
        ctx = ctx or default_context()
        count = self.rows - first if count is None else count
        _check(self._L.vs_matrix_save_spool(ctx.handle, self._h, int(first), int(count), os.fsencode(path), 1 if append else 0))

    def ReadRows(self, first=0, count=None):
        count = self.rows - first if count is None else count
        out = np.empty((count, 8 + self.cols), np.uint8)
        _check(self._L.vs_matrix_read_rows(default_context().handle, self._h, first, count, _p(out)))
        return out

    def __del__(self):
        try:
            if self._h is not None:
                self._L.vs_matrix_release(self._h)
                self._h = None
        except Exception:
            pass


class Vector:
    """compute.Vector (types.go:3-6): the quantized query row; scored on device."""

    def __init__(self, row):
        self._row = row

    def Clone(self):
        return Vector(self._row)  # immutable: nothing to copy

    def MatrixCosineSimilarity(self, matrix, ctx=None):
        """cosine.go:13-57 -> float32 similarity per matrix row."""
        ctx = ctx or default_context()
        out = np.empty(matrix.rows, np.float32)
        _check(matrix._L.vs_cosine_1xN(ctx.handle, _p(self._row), self._row.shape[0], matrix._h, _p(out)))
        return out

    def IntegerDots(self, matrix, ctx=None):
        """sum_j q[j]*v[j] per row (uint32): the exact integer product the scores are built from."""
        ctx = ctx or default_context()
        out = np.empty(matrix.rows, np.uint32)
        _check(matrix._L.vs_dot_1xN(ctx.handle, _p(self._row), self._row.shape[0], matrix._h, _p(out)))
        return out


def NewVector(vectorQuantized):
    """compute.go:10-21."""
    row = np.ascontiguousarray(np.frombuffer(bytes(vectorQuantized), dtype=np.uint8)
                               if not isinstance(vectorQuantized, np.ndarray) else vectorQuantized, dtype=np.uint8)
    if row.ndim != 1 or row.shape[0] - 8 <= 0:
        raise ComputePanic("vector columns are empty")  # compute.go:12-14
    _lib.init()
    return Vector(row.copy())


def NewMatrix(matrixQuantized, ctx=None):
    """compute.go:23-44 (no dequantize: codes, headers and integer sums go to HBM)."""
    rows = _rows_array(matrixQuantized)
    if rows.shape[0] == 0:
        raise ComputePanic("matrix rows are empty")
    if rows.shape[1] - 8 <= 0:
        raise ComputePanic("matrix columns are empty")  # compute.go:29-31
    L = _lib.init()
    ctx = ctx or default_context()
    h = C.c_void_p()
    _check(L.vs_matrix_create(ctx.handle, _p(rows), rows.shape[0], rows.shape[1], C.byref(h)))
    return Matrix(h, L)


def LoadSpool(path, d, first_row=0, count=0, ctx=None):
    """A device matrix from rows [first_row, first_row+count) (count 0 = to the end) of a D&C row spool: a flat file of
    8+d-byte rows (dnc/dataset.go:19-56,122-146)."""
    L = _lib.init()
    ctx = ctx or default_context()
    h = C.c_void_p()
    _check(L.vs_matrix_load_spool(ctx.handle, os.fsencode(path), 8 + int(d), int(first_row), int(count), C.byref(h)))
    return Matrix(h, L)


def EmptyMatrix(n, d, ctx=None):
    """Device matrix of n x d to be filled by Matrix.FillFloat32Dev / LoadRows (bulk loaders)."""
    L = _lib.init()
    ctx = ctx or default_context()
    h = C.c_void_p()
    _check(L.vs_matrix_create_empty(ctx.handle, int(n), int(d), C.byref(h)))
    return Matrix(h, L)


def VectorMatrixCosineSimilarity():
    """cosine.go:60-66: (calculate, done). The closure owns one stream + scratch arena."""
    ctx = Context()

    def calculate(vector, matrix):
        return vector.MatrixCosineSimilarity(matrix, ctx=ctx)

    def done():
        ctx.close()

    return calculate, done


def MatrixCosineSimilarity():
    """cosine.go:129-135: (calculate, done)."""
    ctx = Context()

    def calculate(matrix1, matrix2):
        return matrix1.MatrixCosineSimilarity(matrix2, ctx=ctx)

    def done():
        ctx.close()

    return calculate, done


# ---- compute/quantization.go ------------------------------------------------------------------
def _quantize(matrix, dtype, fn_name):
    m = np.ascontiguousarray(matrix, dtype=dtype)
    if m.ndim != 2:
        raise ValueError("matrix must be 2-D")
    n, d = m.shape
    out = np.empty((n, 8 + d), np.uint8)
    if n:
        L = _lib.init()
        _check(getattr(L, fn_name)(default_context().handle, _p(m), n, d, _p(out)))
    return out


def QuantizeMatrixFloat32(matrix):
    """quantization.go:142-148."""
    return _quantize(matrix, np.float32, "vs_quantize_f32")


def QuantizeMatrixFloat64(matrix):
    """quantization.go:150-156."""
    return _quantize(matrix, np.float64, "vs_quantize_f64")


def QuantizeVectorFloat32(vector):
    """quantization.go:82-91."""
    v = np.ascontiguousarray(vector, dtype=np.float32).reshape(1, -1)
    return _quantize(v, np.float32, "vs_quantize_f32")[0]


def QuantizeVectorFloat64(vector):
    """quantization.go:93-102."""
    v = np.ascontiguousarray(vector, dtype=np.float64).reshape(1, -1)
    return _quantize(v, np.float64, "vs_quantize_f64")[0]


def _dequantize(rows, dtype, fn_name):
    r = _rows_array(rows)
    n, rb = r.shape
    out = np.empty((n, rb - 8), dtype)
    if n and rb > 8:
        L = _lib.init()
        _check(getattr(L, fn_name)(default_context().handle, _p(r), n, rb, _p(out)))
    return out


def DequantizeMatrixFloat32(matrixQuantized):
    """quantization.go:166-172."""
    return _dequantize(matrixQuantized, np.float32, "vs_dequantize_f32")


def DequantizeMatrixFloat64(matrixQuantized):
    """quantization.go:174-180."""
    return _dequantize(matrixQuantized, np.float64, "vs_dequantize_f64")


def DequantizeVectorFloat32(vectorQuantized):
    """quantization.go:114-122."""
    return _dequantize(np.asarray(vectorQuantized, np.uint8).reshape(1, -1), np.float32, "vs_dequantize_f32")[0]


def DequantizeVectorFloat64(vectorQuantized):
    """quantization.go:124-132."""
    return _dequantize(np.asarray(vectorQuantized, np.uint8).reshape(1, -1), np.float64, "vs_dequantize_f64")[0]
