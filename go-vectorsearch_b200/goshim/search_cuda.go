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
