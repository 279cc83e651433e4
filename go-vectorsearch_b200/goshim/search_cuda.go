//go:build cuda
// +build cuda

// search_cuda.go -- device-resident IVF-Flat index for server/search.go and the Lloyd step for dnc/k_means.go.
// These are extensions beyond the reference's compute API (SURVEY.md 8b): the server loads a category's
// centroids and embeddings into HBM once and then answers Search without touching the database.
package compute

/*
#include <stdlib.h>
#include "vscuda.h"
*/
import "C"

import (
	"runtime"
	"unsafe"
)

// Index is the device-resident store of one category (server/search.go:202-273).
type Index struct{ h *C.vs_index }

// NewIndex streams rows (Embedding.Vector, in primary-key order), their DocumentID and the index of their
// centroid (position of Embedding.CentroidID in `centroids`) into HBM and groups them into posting lists.
func NewIndex(rows [][]uint8, documentIDs []uint64, centroidIndex []uint32, centroids [][]uint8) *Index {
	rbuf, n, rowBytes := pack(rows)
	defer C.free(rbuf)
	cbuf, nc, _ := pack(centroids)
	defer C.free(cbuf)
	ix := &Index{}
	withDefaultCtx(func(c *ctx) {
		check(C.vs_index_build_assigned(c.h, (*C.uint8_t)(rbuf), C.size_t(n), C.size_t(rowBytes),
			(*C.uint64_t)(unsafe.Pointer(&documentIDs[0])), (*C.uint32_t)(unsafe.Pointer(&centroidIndex[0])),
			(*C.uint8_t)(cbuf), C.size_t(nc), &ix.h))
	})
	runtime.SetFinalizer(ix, func(ix *Index) { C.vs_index_release(ix.h) })
	return ix
}

// Search replaces the loop at server/search.go:214-273: nprobe = req.Centroids, k = req.Count+req.Offset.
// Returns document IDs and float32 similarities, similarity desc then ID asc, one entry per document.
func (ix *Index) Search(query []uint8, nprobe int, k int) (documentIDs []uint64, similarities []float32) {
	ids := make([]uint64, k)
	sims := make([]float32, k)
	var count C.int32_t
	withDefaultCtx(func(c *ctx) {
		check(C.vs_search(c.h, ix.h, (*C.uint8_t)(unsafe.Pointer(&query[0])), 1, C.size_t(nprobe), C.size_t(k),
			(*C.uint64_t)(unsafe.Pointer(&ids[0])), (*C.float)(unsafe.Pointer(&sims[0])), &count))
	})
	return ids[:count], sims[:count]
}

// NewIndexLoader reserves the device store of a category from its per-centroid embedding counts (the GROUP BY of
// dnc.go:465-470); Fill then streams the embeddings table into it in primary-key order, FindInBatches chunk by chunk
// (search.go:241-243's own read order), without ever holding a second copy of the rows in HBM.  The index answers
// searches once every counted row has been placed.
func NewIndexLoader(centroids Matrix, rowsPerCentroid []uint64) *Index {
	m := centroids.(*matrixContainer)
	ix := &Index{}
	withDefaultCtx(func(c *ctx) {
		check(C.vs_index_create_empty(c.h, m.h, (*C.uint64_t)(unsafe.Pointer(&rowsPerCentroid[0])), &ix.h))
	})
	runtime.SetFinalizer(ix, func(ix *Index) { C.vs_index_release(ix.h) })
	return ix
}

// Fill places one chunk of embeddings (Embedding.Vector, DocumentID, position of CentroidID in the centroid table).
func (ix *Index) Fill(rows [][]uint8, documentIDs []uint64, centroidIndex []uint32) {
	rbuf, n, rowBytes := pack(rows)
	defer C.free(rbuf)
	withDefaultCtx(func(c *ctx) {
		check(C.vs_index_fill(c.h, ix.h, (*C.uint8_t)(rbuf), C.size_t(n), C.size_t(rowBytes),
			(*C.uint32_t)(unsafe.Pointer(&centroidIndex[0])), (*C.uint64_t)(unsafe.Pointer(&documentIDs[0])), 0))
	})
}

// Upload is the assignment of server/upload.go:239-279 plus the insert into the device store: every new embedding goes to
// its nearest centroid (upload.go:245) and joins that posting list behind the rows already there.  centroidIndex[i] is
// what upload.go:268-271 turns into Embedding.CentroidID (centroids[centroidIndex[i]].ID).  The receiver stays valid for
// searches in flight; the server swaps its category's *Index for the returned one.
func (ix *Index) Upload(rows [][]uint8, documentIDs []uint64) (next *Index, centroidIndex []int64) {
	rbuf, n, rowBytes := pack(rows)
	defer C.free(rbuf)
	centroidIndex = make([]int64, n)
	next = &Index{}
	withDefaultCtx(func(c *ctx) {
		check(C.vs_index_upload(c.h, ix.h, (*C.uint8_t)(rbuf), C.size_t(n), C.size_t(rowBytes),
			(*C.uint64_t)(unsafe.Pointer(&documentIDs[0])), (*C.int64_t)(unsafe.Pointer(&centroidIndex[0])), &next.h))
	})
	runtime.SetFinalizer(next, func(ix *Index) { C.vs_index_release(ix.h) })
	return next, centroidIndex
}

// WithRoom copies the store into one whose every list can grow by max(percent % of its rows, minRows) rows, so that
// Append can write new embeddings in place (vs_index_with_room).
func (ix *Index) WithRoom(percent, minRows int) *Index {
	next := &Index{}
	withDefaultCtx(func(c *ctx) {
		check(C.vs_index_with_room(c.h, ix.h, C.size_t(percent), C.size_t(minRows), &next.h))
	})
	runtime.SetFinalizer(next, func(ix *Index) { C.vs_index_release(ix.h) })
	return next
}

// Append is Upload in place: the new embeddings are written behind the last row of their nearest centroid's list, a cost
// proportional to the new rows only.  ok == false (VS_EFULL) means some list has no room and nothing was changed: the
// caller appends to ix.WithRoom(...) instead and swaps the pointer, like growing a slice.  One appending goroutine per
// index; searches on other contexts see every list either before or after the append.
func (ix *Index) Append(rows [][]uint8, documentIDs []uint64) (centroidIndex []int64, ok bool) {
	rbuf, n, rowBytes := pack(rows)
	defer C.free(rbuf)
	centroidIndex = make([]int64, n)
	ok = true
	withDefaultCtx(func(c *ctx) {
		rc := C.vs_index_append(c.h, ix.h, (*C.uint8_t)(rbuf), C.size_t(n), C.size_t(rowBytes),
			(*C.uint64_t)(unsafe.Pointer(&documentIDs[0])), (*C.int64_t)(unsafe.Pointer(&centroidIndex[0])))
		if rc == C.VS_EFULL {
			ok = false
			return
		}
		check(rc)
	})
	return centroidIndex, ok
}

// KMeansStep is one iteration of dnc/k_means.go:67-117 on a device matrix. means is the [k][d] float32 state the
// reference carries between iterations (flattened); it is updated in place.
func KMeansStep(data Matrix, centroids [][]uint8, means []float32) (counts []int64, newCentroids [][]uint8, converged bool) {
	cbuf, k, rowBytes := pack(centroids)
	defer C.free(cbuf)
	counts = make([]int64, k)
	out := make([]uint8, k*rowBytes)
	var conv C.int
	m := data.(*matrixContainer)
	withDefaultCtx(func(c *ctx) {
		check(C.vs_kmeans_step(c.h, m.h, (*C.uint8_t)(cbuf), C.size_t(k), (*C.float)(unsafe.Pointer(&means[0])), nil,
			(*C.int64_t)(unsafe.Pointer(&counts[0])), (*C.uint8_t)(unsafe.Pointer(&out[0])), &conv))
	})
	newCentroids = make([][]uint8, k)
	for i := range newCentroids {
		newCentroids[i] = out[i*rowBytes : (i+1)*rowBytes]
	}
	return counts, newCentroids, conv != 0
}

// KMeans is the whole kMeans of dnc/k_means.go:19-212 on the device. supersetRows are the distinct random row indices
// the reference draws at k_means.go:35-44 (min(len(data), k*config.SUPERSET_MUL) of them); the two convergence loops,
// the truncation to the first k centroids and the float32 means all stay in HBM. iterLimit = config.KMEANS_ITTERATION_LIMIT.
func KMeans(data Matrix, k int, supersetRows []uint64, iterLimit int) (centroids [][]uint8) {
	m := data.(*matrixContainer)
	rowBytes := m.cols + 8
	out := make([]uint8, k*rowBytes)
	withDefaultCtx(func(c *ctx) {
		check(C.vs_kmeans(c.h, m.h, C.size_t(k), (*C.uint64_t)(unsafe.Pointer(&supersetRows[0])), C.size_t(len(supersetRows)),
			C.size_t(iterLimit), (*C.uint8_t)(unsafe.Pointer(&out[0])), nil))
	})
	centroids = make([][]uint8, k)
	for i := range centroids {
		centroids[i] = out[i*rowBytes : (i+1)*rowBytes]
	}
	return centroids
}

// LoadSpool builds a device matrix straight from a D&C cache file (dnc/dataset.go: a flat file of 8+vectorSize-byte
// rows), replacing the ReadRow loop + NewMatrix re-pack of chunkData (k_means.go:214-221). count 0 = to the end.
func LoadSpool(path string, vectorSize int, firstRow, count uint64) Matrix {
	cpath := C.CString(path)
	defer C.free(unsafe.Pointer(cpath))
	m := &matrixContainer{cols: vectorSize}
	withDefaultCtx(func(c *ctx) {
		check(C.vs_matrix_load_spool(c.h, cpath, C.size_t(8+vectorSize), C.size_t(firstRow), C.size_t(count), &m.h))
	})
	m.rows = int(C.vs_matrix_rows(m.h))
	runtime.SetFinalizer(m, func(m *matrixContainer) { C.vs_matrix_release(m.h) })
	return m
}

// SearchBatch answers many queries at once (queries as packed 8+D-byte rows). With nprobe >= the number of lists and
// 64 or more queries the library scores the batch on the tensor cores (tcgen05 int8 GEMM + fused filter).
func (ix *Index) SearchBatch(queries [][]uint8, nprobe int, k int) (documentIDs [][]uint64, similarities [][]float32) {
	qbuf, nq, _ := pack(queries)
	defer C.free(qbuf)
	ids := make([]uint64, nq*k)
	sims := make([]float32, nq*k)
	counts := make([]int32, nq)
	withDefaultCtx(func(c *ctx) {
		check(C.vs_search(c.h, ix.h, (*C.uint8_t)(qbuf), C.size_t(nq), C.size_t(nprobe), C.size_t(k),
			(*C.uint64_t)(unsafe.Pointer(&ids[0])), (*C.float)(unsafe.Pointer(&sims[0])), (*C.int32_t)(unsafe.Pointer(&counts[0]))))
	})
	documentIDs = make([][]uint64, nq)
	similarities = make([][]float32, nq)
	for i := 0; i < nq; i++ {
		documentIDs[i] = ids[i*k : i*k+int(counts[i])]
		similarities[i] = sims[i*k : i*k+int(counts[i])]
	}
	return documentIDs, similarities
}

// Split is the split loop of dnc.divideNconquer (dnc/dnc.go:363-389) on a device matrix: child j receives the rows of X
// nearest to centroids[j], in X's order (nil for a child without rows).  Together with KMeans and a sample taken with
// Gather this is all divideNconquer needs; see go-vectorsearch_b200/dnc.py:DivideAndConquer for the recursion.
func Split(X Matrix, centroids [][]uint8) []Matrix {
	m := X.(*matrixContainer)
	cbuf, k, _ := pack(centroids)
	defer C.free(cbuf)
	handles := make([]*C.vs_matrix, k)
	counts := make([]C.uint64_t, k)
	withDefaultCtx(func(c *ctx) {
		check(C.vs_matrix_split(c.h, m.h, (*C.uint8_t)(cbuf), C.size_t(k), &handles[0], &counts[0]))
	})
	out := make([]Matrix, k)
	for j, h := range handles {
		if h == nil {
			continue
		}
		child := &matrixContainer{h: h, rows: int(counts[j]), cols: m.cols}
		runtime.SetFinalizer(child, func(m *matrixContainer) { C.vs_matrix_release(m.h) })
		out[j] = child
	}
	return out
}

// Gather is sample() (dnc/sampling.go:12-74) once the sorted row indices are drawn: a new device matrix of those rows.
func Gather(X Matrix, rows []uint64) Matrix {
	m := X.(*matrixContainer)
	g := &matrixContainer{rows: len(rows), cols: m.cols}
	withDefaultCtx(func(c *ctx) {
		check(C.vs_matrix_gather(c.h, m.h, (*C.uint64_t)(unsafe.Pointer(&rows[0])), C.size_t(len(rows)), &g.h))
	})
	runtime.SetFinalizer(g, func(m *matrixContainer) { C.vs_matrix_release(m.h) })
	return g
}

// ReassignRecenter is the tail of KMeansDivideAndConquer (dnc/dnc.go:177-291): the index of the nearest new centroid for
// every row, and every centroid re-centred on its members (recenterDbCentroid, dnc.go:402-456).
func ReassignRecenter(X Matrix, centroids [][]uint8) (assign []int32, recentred [][]uint8, counts []int64) {
	m := X.(*matrixContainer)
	cbuf, k, rowBytes := pack(centroids)
	defer C.free(cbuf)
	assign = make([]int32, m.rows)
	out := make([]uint8, k*rowBytes)
	counts = make([]int64, k)
	withDefaultCtx(func(c *ctx) {
		check(C.vs_reassign_recenter(c.h, m.h, (*C.uint8_t)(cbuf), C.size_t(k), (*C.int32_t)(unsafe.Pointer(&assign[0])), nil,
			(*C.uint8_t)(unsafe.Pointer(&out[0])), (*C.int64_t)(unsafe.Pointer(&counts[0]))))
	})
	recentred = make([][]uint8, k)
	for i := range recentred {
		recentred[i] = out[i*rowBytes : (i+1)*rowBytes]
	}
	return assign, recentred, counts
}

// ShardedIndex is the device-resident store of one category striped over several GPUs of one box, driven from THIS
// process (main.go:31 is one process; server/search.go:115 runs one Search per goroutine): rows go to device
// primaryKey % len(devices), the centroid table is replicated, every device answers on its stripe and device 0 merges
// the shard-local hits it reads from its peers over NVLink (vs_sharded_*, csrc/sharded.cu).  Same hits as one Index.
type ShardedIndex struct{ h *C.vs_sharded }

// NewShardedIndex: rows in primary-key order (row i has key i), their DocumentID and the position of their CentroidID
// in `centroids`.
func NewShardedIndex(devices []int, rows [][]uint8, documentIDs []uint64, centroidIndex []uint32, centroids [][]uint8) *ShardedIndex {
	rbuf, n, rowBytes := pack(rows)
	defer C.free(rbuf)
	cbuf, nc, _ := pack(centroids)
	defer C.free(cbuf)
	devs := make([]C.int, len(devices))
	for i, d := range devices {
		devs[i] = C.int(d)
	}
	sh := &ShardedIndex{}
	pinned(func() {
		check(C.vs_sharded_create(&devs[0], C.size_t(len(devs)), &sh.h))
		check(C.vs_sharded_build_assigned(sh.h, (*C.uint8_t)(rbuf), C.size_t(n), C.size_t(rowBytes),
			(*C.uint64_t)(unsafe.Pointer(&documentIDs[0])), (*C.uint32_t)(unsafe.Pointer(&centroidIndex[0])),
			(*C.uint8_t)(cbuf), C.size_t(nc)))
	})
	runtime.SetFinalizer(sh, func(sh *ShardedIndex) { C.vs_sharded_release(sh.h) })
	return sh
}

// ShardedSearch returns the pair the reference's call sites use (search.go:230): calculate answers a batch of queries
// (nprobe = req.Centroids, k = req.Count+req.Offset), done releases the per-device streams and buffers.  One pair per
// goroutine; pairs run concurrently.
func (sh *ShardedIndex) ShardedSearch() (calculate func(queries [][]uint8, nprobe int, k int) (documentIDs [][]uint64, similarities [][]float32), done func()) {
	var sc *C.vs_sharded_ctx
	pinned(func() { check(C.vs_sharded_ctx_create(sh.h, &sc)) })
	return func(queries [][]uint8, nprobe int, k int) ([][]uint64, [][]float32) {
			qbuf, nq, _ := pack(queries)
			defer C.free(qbuf)
			ids := make([]uint64, nq*k)
			sims := make([]float32, nq*k)
			counts := make([]C.int32_t, nq)
			pinned(func() {
				check(C.vs_sharded_search_ctx(sc, (*C.uint8_t)(qbuf), C.size_t(nq), C.size_t(nprobe), C.size_t(k),
					(*C.uint64_t)(unsafe.Pointer(&ids[0])), (*C.float)(unsafe.Pointer(&sims[0])), &counts[0]))
			})
			outIDs := make([][]uint64, nq)
			outSims := make([][]float32, nq)
			for q := 0; q < nq; q++ {
				outIDs[q] = ids[q*k : q*k+int(counts[q])]
				outSims[q] = sims[q*k : q*k+int(counts[q])]
			}
			return outIDs, outSims
		}, func() {
			C.vs_sharded_ctx_destroy(sc)
		}
}

// Upload is server/upload.go:239-279 on the striped store: the new embeddings take the next primary keys and join the
// list of their nearest centroid on the device that owns their key.  Must not overlap searches on this index.
func (sh *ShardedIndex) Upload(rows [][]uint8, documentIDs []uint64) (centroidIndex []int64) {
	rbuf, n, rowBytes := pack(rows)
	defer C.free(rbuf)
	centroidIndex = make([]int64, n)
	pinned(func() {
		check(C.vs_sharded_upload(sh.h, (*C.uint8_t)(rbuf), C.size_t(n), C.size_t(rowBytes),
			(*C.uint64_t)(unsafe.Pointer(&documentIDs[0])), (*C.int64_t)(unsafe.Pointer(&centroidIndex[0]))))
	})
	return
}
