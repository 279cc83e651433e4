// selftest.cpp -- exercises the C ABI through the C++ mirror of the reference's compute package with the
// hand-derived known answers of tests/golden/kat.json (written out here so the binary is self-contained).
// Exit code 0 = all checks passed. Needs a B200 (no CPU fallback): tests/test_gpu_host_cpp.py runs it.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "compute.hpp"

static int fails = 0;
#define EXPECT(cond, what)                                   \
    do {                                                     \
        if (!(cond)) { std::printf("FAIL: %s\n", what); fails++; } \
    } while (0)

static compute::Row row(float mn, float mx, std::initializer_list<int> codes) {
    compute::Row r(8);
    std::memcpy(r.data(), &mn, 4);
    std::memcpy(r.data() + 4, &mx, 4);
    for (int c : codes) r.push_back((uint8_t)c);
    return r;
}

int main() {
    using namespace compute;
    try {
        // quantization.go:82-91 -- range seeded at 0, truncation
        EXPECT(QuantizeVectorFloat32({0.5f, 1.0f, 0.25f}) == row(0.f, 1.f, {127, 255, 63}), "quantize positive_keeps_min_zero");
        EXPECT(QuantizeVectorFloat32({-1.f, 0.f, 1.f}) == row(-1.f, 1.f, {0, 127, 255}), "quantize symmetric");
        EXPECT(QuantizeVectorFloat32({0.f, 0.f, 0.f}) == row(0.f, 0.f, {0, 0, 0}), "quantize all_zero (NaN -> 0)");
        EXPECT(QuantizeVectorFloat32({0.f, 1.f, 0.999f}) == row(0.f, 1.f, {0, 255, 254}), "quantize truncation");
        EXPECT(QuantizeVectorFloat64({0.1, -0.3}) == row(-0.3f, 0.1f, {255, 0}), "quantize f64 header rounding");
        // quantization.go:124-132
        auto dq = DequantizeVectorFloat64(row(-1.f, 1.f, {0, 255, 51}));
        EXPECT(dq[0] == -1.0 && dq[1] == 1.0 && dq[2] == -1.0 + (51.0 / 255.0) * 2.0, "dequantize f64");
        // cosine.go:13-57
        Vector q = NewVector(row(0.f, 1.f, {255, 0}));
        Matrix m = NewMatrix({row(0.f, 1.f, {0, 255}), row(0.f, 1.f, {255, 255}), row(0.f, 1.f, {255, 0}), row(0.f, 0.f, {7, 9}),
                              row(-1.f, 0.f, {0, 255})});
        auto sims = q.Clone().MatrixCosineSimilarity(m);
        const float s1 = (float)(1.0 / std::sqrt(2.0));
        EXPECT(sims[0] == 0.f && sims[1] == s1 && sims[2] == 1.f && sims[3] == 0.f && sims[4] == -1.f, "cosine 1xN known answers");
        auto cl = VectorMatrixCosineSimilarity();
        EXPECT(cl.calculate(q.Clone(), m.Clone()) == sims, "closure calculate == method");
        cl.done();
        // cosine.go:70-125 -- strict '>' keeps the lowest index
        Matrix cent = NewMatrix({row(0.f, 1.f, {255, 0}), row(0.f, 1.f, {0, 255}), row(0.f, 1.f, {255, 0})});
        Matrix data = NewMatrix({row(0.f, 1.f, {255, 0}), row(0.f, 1.f, {0, 255}), row(0.f, 0.f, {0, 0}), row(0.f, 1.f, {255, 255}),
                                 row(-1.f, 0.f, {0, 255})});
        auto mm = MatrixCosineSimilarity();
        auto res = mm.calculate(cent.Clone(), data.Clone());
        mm.done();
        EXPECT((res.second == std::vector<int64_t>{0, 1, 0, 0, 1}), "argmax known answers");
        // search.go:202-273 + upload.go:239-279 -- the five data rows above behind the first two centroids; by hand: lists
        // {1,0,0,0,1} (row 1 ties -> lowest index; the zero row scores 0 > -1 against centroid 0), query (1,0)
        const Rows cents = {row(0.f, 1.f, {255, 0}), row(0.f, 1.f, {0, 255})};
        const Rows drows = {row(0.f, 1.f, {0, 255}), row(0.f, 1.f, {255, 255}), row(0.f, 1.f, {255, 0}), row(0.f, 0.f, {0, 0}),
                            row(-1.f, 0.f, {0, 255})};
        const Row query = row(0.f, 1.f, {255, 0});
        Index ix = NewIndex(drows, {100, 101, 102, 103, 104}, {1, 0, 0, 0, 1}, cents);
        Hits h1 = ix.Search(query, 1, 10);
        EXPECT((h1.documentIDs == std::vector<uint64_t>{102, 101, 103}), "ivf search nprobe=1 ids");
        EXPECT(h1.similarities.size() == 3 && h1.similarities[0] == 1.f && h1.similarities[1] == s1 && h1.similarities[2] == 0.f,
               "ivf search nprobe=1 similarities");
        Hits h2 = ix.Search(query, 2, 10);
        EXPECT((h2.documentIDs == std::vector<uint64_t>{102, 101, 100, 103, 104}), "ivf search all lists: ties by document id");
        Index first3 = NewIndex({drows[0], drows[1], drows[2]}, {100, 101, 102}, {1, 0, 0}, cents);
        auto up = first3.Upload({drows[3], drows[4]}, {103, 104});
        EXPECT((up.second == std::vector<int64_t>{0, 1}), "upload assignment (upload.go:245)");
        Hits h3 = up.first.Search(query, 2, 10);
        EXPECT(h3.documentIDs == h2.documentIDs && h3.similarities == h2.similarities, "index after upload == index built from all rows");
        EXPECT(first3.rows() == 3 && up.first.rows() == 5, "upload leaves the old index untouched");
        Index ld = NewIndexLoader(NewMatrix(cents), {3, 2});
        EXPECT(ld.Search(query, 2, 10).documentIDs.empty(), "an empty loader index answers with nothing");
        ld.Fill({drows[0], drows[1]}, {100, 101}, {1, 0});
        EXPECT((ld.Search(query, 2, 10).documentIDs == std::vector<uint64_t>{101, 100}), "a half-loaded index answers from the rows placed so far");
        ld.Fill({drows[2], drows[3], drows[4]}, {102, 103, 104}, {0, 0, 1});
        Hits h4 = ld.Search(query, 2, 10);
        EXPECT(h4.documentIDs == h2.documentIDs && h4.similarities == h2.similarities, "streamed index == index built in one piece");
        // upload in place: the same two rows appended behind their lists
        Index roomy = first3.WithRoom(0, 2);
        std::vector<int64_t> where;
        EXPECT((roomy.Append({drows[3], drows[4]}, {103, 104}, &where) && where == std::vector<int64_t>{0, 1}), "append assignment");
        Hits h5 = roomy.Search(query, 2, 10);
        EXPECT(h5.documentIDs == h2.documentIDs && h5.similarities == h2.similarities, "index after append == index built from all rows");
        EXPECT((!roomy.Append({drows[2], drows[2], drows[2]}, {105, 106, 107}) && roomy.rows() == 5), "a full list refuses and changes nothing");
        // error behaviour
        bool panicked = false;
        try { NewVector(compute::Row(8, 0)); } catch (const Panic &) { panicked = true; }
        EXPECT(panicked, "NewVector panics on empty columns (compute.go:12-14)");
        panicked = false;
        try { NewMatrix({}); } catch (const Panic &) { panicked = true; }
        EXPECT(panicked, "NewMatrix panics on empty rows (compute.go:25-27)");
        bool fatal = false;
        try { NewVector(row(0.f, 1.f, {1, 2, 3})).MatrixCosineSimilarity(m); } catch (const Fatal &) { fatal = true; }
        EXPECT(fatal, "dimension mismatch is fatal (cosine.go:19-21)");
    } catch (const std::exception &e) {
        std::printf("FAIL: exception %s\n", e.what());
        fails++;
    }
    std::printf(fails ? "selftest: %d FAILED\n" : "selftest: ok\n", fails);
    return fails ? 1 : 0;
}
