// compute.hpp -- C++ host-side mirror of the reference's Go `compute` package over the C ABI (include/vscuda.h).
//
// The reference is compiled code (Go) whose toolchain is absent from the build image, so the host side
// above the C ABI is written in C++ with the same names, argument meaning and error behaviour:
//   compute/types.go:3-11       Vector, Matrix { Clone, MatrixCosineSimilarity }
//   compute/compute.go:10-44    NewVector, NewMatrix          (Go panic -> compute::Panic)
//   compute/cosine.go:60-66     VectorMatrixCosineSimilarity() -> {calculate, done}
//   compute/cosine.go:129-135   MatrixCosineSimilarity()       -> {calculate, done}
//   compute/quantization.go     Quantize{Vector,Matrix}Float{32,64}, Dequantize...
//   server/search.go:202-273, server/upload.go:239-279 -> Index { Search, Upload, Fill } (device-resident extension)
// logger.Sugar().Fatalf (process exit in Go, cosine.go:19-21,77-79) surfaces as compute::Fatal.
// The Go shim a maintainer adds is goshim/*.go; this header is what the C++ self-test and any C++ caller use.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/vscuda.h"

namespace compute {

struct Panic : std::runtime_error { using std::runtime_error::runtime_error; };   // Go panic()
struct Fatal : std::runtime_error { using std::runtime_error::runtime_error; };   // logger Fatalf
struct Error : std::runtime_error { using std::runtime_error::runtime_error; };   // any other backend failure

inline void check(int rc) {
    if (rc == VS_OK) return;
    std::string msg = vs_last_error();
    if (rc == VS_EEMPTY) throw Panic(msg);
    if (rc == VS_EDIM) throw Fatal(msg);
    throw Error("libvscuda error " + std::to_string(rc) + ": " + msg);
}

inline void Init() {
    static bool done = false;
    if (done) return;
    const char *lr = std::getenv("LOCAL_RANK");
    check(vs_init(lr ? std::atoi(lr) : 0));  // no CPU fallback: throws without a B200
    done = true;
}

// One CUDA stream + scratch arena; what a calculate-closure owns.
class Context {
  public:
    Context() { Init(); check(vs_ctx_create(&h_)); }
    ~Context() { Close(); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    void Close() { if (h_) { vs_ctx_destroy(h_); h_ = nullptr; } }
    vs_ctx *handle() const { if (!h_) throw Error("context released (done() was called)"); return h_; }
  private:
    vs_ctx *h_ = nullptr;
};

inline Context &DefaultContext() { static Context c; return c; }

using Row = std::vector<uint8_t>;
using Rows = std::vector<Row>;

inline std::vector<uint8_t> Pack(const Rows &rows) {
    if (rows.empty()) throw Panic("matrix rows are empty");                      // compute.go:25-27
    const size_t rb = rows[0].size();
    if (rb <= 8) throw Panic("matrix columns are empty");                        // compute.go:29-31
    std::vector<uint8_t> buf(rows.size() * rb);
    for (size_t i = 0; i < rows.size(); i++) {
        if (rows[i].size() != rb) throw Panic("matrix rows have different lengths");
        std::copy(rows[i].begin(), rows[i].end(), buf.begin() + i * rb);
    }
    return buf;
}

class Matrix;
class Vector {
  public:
    explicit Vector(Row row) : row_(std::make_shared<Row>(std::move(row))) {}
    Vector Clone() const { return *this; }                                       // immutable: nothing to copy
    std::vector<float> MatrixCosineSimilarity(const Matrix &m, Context *ctx = nullptr) const;  // cosine.go:13-57
    const Row &row() const { return *row_; }
  private:
    std::shared_ptr<Row> row_;
};

class Matrix {
  public:
    explicit Matrix(vs_matrix *h) : h_(h, [](vs_matrix *p) { vs_matrix_release(p); }) {}
    Matrix Clone() const { return *this; }                                       // reference-count bump
    size_t rows() const { return vs_matrix_rows(h_.get()); }
    size_t cols() const { return vs_matrix_cols(h_.get()); }
    vs_matrix *handle() const { return h_.get(); }
    // receiver = centroids, argument = data (cosine.go:70-125) -> {sims, nearestIndexList}
    std::pair<std::vector<float>, std::vector<int64_t>> MatrixCosineSimilarity(const Matrix &data, Context *ctx = nullptr) const {
        Context &c = ctx ? *ctx : DefaultContext();
        std::vector<float> sims(data.rows());
        std::vector<int64_t> idx(data.rows());
        check(vs_argmax_MxN(c.handle(), h_.get(), data.h_.get(), sims.data(), idx.data()));
        return {std::move(sims), std::move(idx)};
    }
  private:
    std::shared_ptr<vs_matrix> h_;
};

inline std::vector<float> Vector::MatrixCosineSimilarity(const Matrix &m, Context *ctx) const {
    Context &c = ctx ? *ctx : DefaultContext();
    std::vector<float> sims(m.rows());
    check(vs_cosine_1xN(c.handle(), row_->data(), row_->size(), m.handle(), sims.data()));
    return sims;
}

inline Vector NewVector(const Row &vectorQuantized) {                             // compute.go:10-21
    if (vectorQuantized.size() <= 8) throw Panic("vector columns are empty");
    Init();
    return Vector(vectorQuantized);
}

inline Matrix NewMatrix(const Rows &matrixQuantized) {                            // compute.go:23-44
    std::vector<uint8_t> buf = Pack(matrixQuantized);
    vs_matrix *h = nullptr;
    check(vs_matrix_create(DefaultContext().handle(), buf.data(), matrixQuantized.size(), matrixQuantized[0].size(), &h));
    return Matrix(h);
}

struct VectorMatrixClosure {
    std::function<std::vector<float>(const Vector &, const Matrix &)> calculate;
    std::function<void()> done;
};
inline VectorMatrixClosure VectorMatrixCosineSimilarity() {                       // cosine.go:60-66
    auto ctx = std::make_shared<Context>();
    return {[ctx](const Vector &v, const Matrix &m) { return v.MatrixCosineSimilarity(m, ctx.get()); },
            [ctx]() { ctx->Close(); }};
}
struct MatrixMatrixClosure {
    std::function<std::pair<std::vector<float>, std::vector<int64_t>>(const Matrix &, const Matrix &)> calculate;
    std::function<void()> done;
};
inline MatrixMatrixClosure MatrixCosineSimilarity() {                             // cosine.go:129-135
    auto ctx = std::make_shared<Context>();
    return {[ctx](const Matrix &a, const Matrix &b) { return a.MatrixCosineSimilarity(b, ctx.get()); },
            [ctx]() { ctx->Close(); }};
}

// ---- device-resident IVF-Flat store of one category (server/search.go:202-273, server/upload.go:239-279) ----
// The same extension beyond the reference's compute API as goshim/search_cuda.go: the database only loads.
struct Hits {
    std::vector<uint64_t> documentIDs;   // similarity desc (float32), then document id asc; one entry per document
    std::vector<float> similarities;
};

class Index {
  public:
    explicit Index(vs_index *h) : h_(h, [](vs_index *p) { vs_index_release(p); }) {}
    size_t rows() const { return vs_index_rows(h_.get()); }
    size_t lists() const { return vs_index_lists(h_.get()); }
    size_t cols() const { return vs_index_cols(h_.get()); }
    vs_index *handle() const { return h_.get(); }
    // search.go:214-273: nprobe = req.Centroids, k = req.Count + req.Offset
    Hits Search(const Row &query, size_t nprobe, size_t k, Context *ctx = nullptr) const {
        Context &c = ctx ? *ctx : DefaultContext();
        Hits out;
        out.documentIDs.resize(k);
        out.similarities.resize(k);
        int32_t count = 0;
        check(vs_search(c.handle(), h_.get(), query.data(), 1, nprobe, k, out.documentIDs.data(), out.similarities.data(), &count));
        out.documentIDs.resize((size_t)count);
        out.similarities.resize((size_t)count);
        return out;
    }
    // upload.go:239-279: nearest centroid of every new row (upload.go:245), rows appended to their lists; returns the
    // index to swap in and the centroid index of every row (-> Embedding.CentroidID, upload.go:268-271)
    std::pair<Index, std::vector<int64_t>> Upload(const Rows &rows, const std::vector<uint64_t> &documentIDs, Context *ctx = nullptr) const {
        Context &c = ctx ? *ctx : DefaultContext();
        std::vector<uint8_t> buf = Pack(rows);
        if (documentIDs.size() != rows.size()) throw Error("one document id per row");
        std::vector<int64_t> assign(rows.size());
        vs_index *h = nullptr;
        check(vs_index_upload(c.handle(), h_.get(), buf.data(), rows.size(), rows[0].size(), documentIDs.data(), assign.data(), &h));
        return {Index(h), std::move(assign)};
    }
    // Upload in place: a copy whose lists have room (vs_index_with_room), then appends behind the lists (vs_index_append).
    // Append returns false -- and changes nothing -- when a list is full: append to WithRoom(...) of this index instead.
    Index WithRoom(size_t percent, size_t minRows, Context *ctx = nullptr) const {
        Context &c = ctx ? *ctx : DefaultContext();
        vs_index *h = nullptr;
        check(vs_index_with_room(c.handle(), h_.get(), percent, minRows, &h));
        return Index(h);
    }
    bool Append(const Rows &rows, const std::vector<uint64_t> &documentIDs, std::vector<int64_t> *centroidIndex = nullptr,
                Context *ctx = nullptr) {
        Context &c = ctx ? *ctx : DefaultContext();
        std::vector<uint8_t> buf = Pack(rows);
        if (documentIDs.size() != rows.size()) throw Error("one document id per row");
        std::vector<int64_t> assign(rows.size());
        const int rc = vs_index_append(c.handle(), h_.get(), buf.data(), rows.size(), rows[0].size(), documentIDs.data(), assign.data());
        if (rc == VS_EFULL) return false;
        check(rc);
        if (centroidIndex) *centroidIndex = std::move(assign);
        return true;
    }
    // streaming loader, step 2: one chunk of the embeddings table in primary-key order
    void Fill(const Rows &rows, const std::vector<uint64_t> &documentIDs, const std::vector<uint32_t> &centroidIndex, Context *ctx = nullptr) {
        Context &c = ctx ? *ctx : DefaultContext();
        std::vector<uint8_t> buf = Pack(rows);
        if (documentIDs.size() != rows.size() || centroidIndex.size() != rows.size()) throw Error("one id and one list per row");
        check(vs_index_fill(c.handle(), h_.get(), buf.data(), rows.size(), rows[0].size(), centroidIndex.data(), documentIDs.data(), 0));
    }
  private:
    std::shared_ptr<vs_index> h_;
};

// rows = Embedding.Vector in primary-key order, documentIDs = Embedding.DocumentID, centroidIndex = position of
// Embedding.CentroidID in `centroids` (database/model.go:9-18,35-37)
inline Index NewIndex(const Rows &rows, const std::vector<uint64_t> &documentIDs, const std::vector<uint32_t> &centroidIndex,
                      const Rows &centroids) {
    std::vector<uint8_t> rbuf = Pack(rows), cbuf = Pack(centroids);
    if (documentIDs.size() != rows.size() || centroidIndex.size() != rows.size()) throw Error("one id and one list per row");
    vs_index *h = nullptr;
    check(vs_index_build_assigned(DefaultContext().handle(), rbuf.data(), rows.size(), rows[0].size(), documentIDs.data(),
                                  centroidIndex.data(), cbuf.data(), centroids.size(), &h));
    return Index(h);
}

// streaming loader, step 1: the store reserved from the per-centroid embedding counts (the GROUP BY of dnc.go:465-470)
inline Index NewIndexLoader(const Matrix &centroids, const std::vector<uint64_t> &rowsPerCentroid) {
    if (rowsPerCentroid.size() != centroids.rows()) throw Error("one count per centroid");
    vs_index *h = nullptr;
    check(vs_index_create_empty(DefaultContext().handle(), centroids.handle(), rowsPerCentroid.data(), &h));
    return Index(h);
}

// ---- compute/quantization.go ----
inline Rows QuantizeMatrixFloat32(const std::vector<std::vector<float>> &matrix) {  // :142-148
    Rows out(matrix.size());
    if (matrix.empty()) return out;
    const size_t d = matrix[0].size();
    std::vector<float> in(matrix.size() * d);
    for (size_t i = 0; i < matrix.size(); i++) std::copy(matrix[i].begin(), matrix[i].end(), in.begin() + i * d);
    std::vector<uint8_t> buf(matrix.size() * (8 + d));
    check(vs_quantize_f32(DefaultContext().handle(), in.data(), matrix.size(), d, buf.data()));
    for (size_t i = 0; i < matrix.size(); i++) out[i].assign(buf.begin() + i * (8 + d), buf.begin() + (i + 1) * (8 + d));
    return out;
}
inline Row QuantizeVectorFloat32(const std::vector<float> &v) { return QuantizeMatrixFloat32({v})[0]; }  // :82-91
inline Rows QuantizeMatrixFloat64(const std::vector<std::vector<double>> &matrix) {  // :150-156
    Rows out(matrix.size());
    if (matrix.empty()) return out;
    const size_t d = matrix[0].size();
    std::vector<double> in(matrix.size() * d);
    for (size_t i = 0; i < matrix.size(); i++) std::copy(matrix[i].begin(), matrix[i].end(), in.begin() + i * d);
    std::vector<uint8_t> buf(matrix.size() * (8 + d));
    check(vs_quantize_f64(DefaultContext().handle(), in.data(), matrix.size(), d, buf.data()));
    for (size_t i = 0; i < matrix.size(); i++) out[i].assign(buf.begin() + i * (8 + d), buf.begin() + (i + 1) * (8 + d));
    return out;
}
inline Row QuantizeVectorFloat64(const std::vector<double> &v) { return QuantizeMatrixFloat64({v})[0]; }  // :93-102
inline std::vector<float> DequantizeVectorFloat32(const Row &row) {               // :114-122
    std::vector<float> out(row.size() - 8);
    check(vs_dequantize_f32(DefaultContext().handle(), row.data(), 1, row.size(), out.data()));
    return out;
}
inline std::vector<double> DequantizeVectorFloat64(const Row &row) {              // :124-132
    std::vector<double> out(row.size() - 8);
    check(vs_dequantize_f64(DefaultContext().handle(), row.data(), 1, row.size(), out.data()));
    return out;
}

}  // namespace compute
