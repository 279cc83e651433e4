//go:build cuda
// +build cuda

// cosine_cuda.go -- the `cuda` build-tag sibling of cosine_gonum.go / cosine_gorgonia.go.
package compute

/*
#include "vscuda.h"
*/
import "C"

import (
	"unsafe"

	"github.com/expki/go-vectorsearch/logger"
)

func vectorMatrix(c *ctx, vector *vectorContainer, matrix *matrixContainer) []float32 {
	if len(vector.row)-8 != matrix.cols { // cosine.go:19-21
		logger.Sugar().Fatalf("vector/matrix column size does not match: %d != %d", len(vector.row)-8, matrix.cols)
	}
	sims := make([]float32, matrix.rows)
	check(C.vs_cosine_1xN(c.h, (*C.uint8_t)(unsafe.Pointer(&vector.row[0])), C.size_t(len(vector.row)), matrix.h,
		(*C.float)(unsafe.Pointer(&sims[0]))))
	return sims
}

func matrixMatrix(c *ctx, centroids *matrixContainer, data *matrixContainer) ([]float32, []int) {
	if centroids.cols != data.cols { // cosine.go:77-79
		logger.Sugar().Fatalf("matrix/matrix column size does not match: %d != %d", centroids.cols, data.cols)
	}
	sims := make([]float32, data.rows)
	idx64 := make([]int64, data.rows)
	check(C.vs_argmax_MxN(c.h, centroids.h, data.h, (*C.float)(unsafe.Pointer(&sims[0])),
		(*C.int64_t)(unsafe.Pointer(&idx64[0]))))
	idx := make([]int, data.rows)
	for i, v := range idx64 {
		idx[i] = int(v)
	}
	return sims, idx
}

// MatrixCosineSimilarity facilitates the computation of cosine similarity between a vector and a matrix.
func (vector *vectorContainer) MatrixCosineSimilarity(matrix Matrix) (similarity []float32) {
	withDefaultCtx(func(c *ctx) { similarity = vectorMatrix(c, vector, matrix.(*matrixContainer)) })
	return similarity
}

// VectorMatrixCosineSimilarity returns (calculate, done): the closure owns one CUDA stream + scratch arena
// (one per goroutine, server/search.go:230), done() releases it.
func VectorMatrixCosineSimilarity() (calculate func(vector Vector, matrix Matrix) (similarity []float32), done func()) {
	c := newCtx()
	return func(vector Vector, matrix Matrix) (similarity []float32) {
			pinned(func() { similarity = vectorMatrix(c, vector.(*vectorContainer), matrix.(*matrixContainer)) })
			return
		}, func() {
			c.close()
		}
}

// MatrixCosineSimilarity: receiver = centroids, argument = data (cosine.go:70-125).
func (matrix1 *matrixContainer) MatrixCosineSimilarity(matrix2 Matrix) (relativeSimilaritieList []float32, nearestIndexList []int) {
	withDefaultCtx(func(c *ctx) {
		relativeSimilaritieList, nearestIndexList = matrixMatrix(c, matrix1, matrix2.(*matrixContainer))
	})
	return
}

// MatrixCosineSimilarity returns (calculate, done) for the dnc workers (dnc/dnc.go:349, dnc/k_means.go:31).
func MatrixCosineSimilarity() (calculate func(matrix1 Matrix, matrix2 Matrix) (relativeSimilaritieList []float32, nearestIndexList []int), done func()) {
	c := newCtx()
	return func(matrix1 Matrix, matrix2 Matrix) (relativeSimilaritieList []float32, nearestIndexList []int) {
			pinned(func() {
				relativeSimilaritieList, nearestIndexList = matrixMatrix(c, matrix1.(*matrixContainer), matrix2.(*matrixContainer))
			})
			return
		}, func() {
			c.close()
		}
}
